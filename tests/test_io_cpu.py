"""The native XMSI reader / XMSI + XMSO writers (host_io.cpp) against the Python reader of the harness
(xmimsim_b200/xmsi.py) on the four shipped examples, and round trips through the writers."""
import ctypes as C
import os

import numpy as np
import pytest

import xmimsim_b200 as x
from xmimsim_b200 import abi
from inputs import GOLDEN, example


def _read(path):
    L = abi.lib()
    p = C.POINTER(abi.Input)()
    assert L.xmb_input_read_from_xml_file(path.encode(), C.byref(p)) == 1, abi.last_error()
    return p


def _layers(ptr, n):
    return [([ptr[i].Z[j] for j in range(ptr[i].n_elements)], [ptr[i].weight[j] for j in range(ptr[i].n_elements)],
             ptr[i].density, ptr[i].thickness) for i in range(n)]


def _as_tuple(inp):
    """Every field of a ctypes xmb_input as plain python, for comparisons."""
    i = inp.contents
    g, c, ge, ex, ab, de = (i.general.contents, i.composition.contents, i.geometry.contents, i.excitation.contents,
                            i.absorbers.contents, i.detector.contents)
    disc = [tuple(getattr(ex.discrete[k], f) for f, _ in abi.EnergyDiscrete._fields_) for k in range(ex.n_discrete)]
    cont = [tuple(getattr(ex.continuous[k], f) for f, _ in abi.EnergyContinuous._fields_) for k in range(ex.n_continuous)]
    return dict(general=(g.outputfile, g.n_photons_interval, g.n_photons_line, g.n_interactions_trajectory, g.comments),
                layers=_layers(c.layers, c.n_layers), ref=c.reference_layer,
                geometry=(ge.d_sample_source, list(ge.n_sample_orientation), list(ge.p_detector_window), list(ge.n_detector_orientation),
                          ge.area_detector, ge.collimator_height, ge.collimator_diameter, ge.d_source_slit, ge.slit_size_x, ge.slit_size_y),
                disc=disc, cont=cont, exc_layers=_layers(ab.exc_layers, ab.n_exc_layers), det_layers=_layers(ab.det_layers, ab.n_det_layers),
                detector=(de.detector_type, de.live_time, de.pulse_width, de.nchannels, de.gain, de.zero, de.fano, de.noise),
                crystal=_layers(de.crystal_layers, de.n_crystal_layers))


@pytest.mark.parametrize("name", ["srm1155", "srm1412", "srm1132", "In"])
def test_native_reader_matches_python_reader(name):
    path = os.path.join(GOLDEN, name + ".xmsi")
    p = _read(path)
    a = _as_tuple(p)
    ci = x.CInput(example(name))
    b = _as_tuple(C.pointer(ci.input))
    for k in a:
        if k in ("layers", "exc_layers", "det_layers", "crystal"):
            assert len(a[k]) == len(b[k])
            for la, lb in zip(a[k], b[k]):
                assert la[0] == lb[0] and np.allclose(la[1], lb[1], rtol=1e-15, atol=0) and la[2:] == lb[2:]
        else:
            assert a[k] == b[k], k
    abi.lib().xmb_input_free(C.byref(p))


def test_reader_errors_are_reported(tmp_path):
    L = abi.lib()
    p = C.POINTER(abi.Input)()
    assert L.xmb_input_read_from_xml_file(b"/nonexistent.xmsi", C.byref(p)) == 0 and "could not open" in abi.last_error()
    bad = tmp_path / "bad.xmsi"
    bad.write_text("<xmimsim><general><outputfile>a</outputfile></general></xmimsim>")
    assert L.xmb_input_read_from_xml_file(str(bad).encode(), C.byref(p)) == 0 and "missing element" in abi.last_error()
    bad.write_text("<other/>")
    assert L.xmb_input_read_from_xml_file(str(bad).encode(), C.byref(p)) == 0 and "root element" in abi.last_error()
    # well-formed but unusable: rejected as by the reference reader's xmi_input_validate call (src/xmi_xml.c:1337-1343)
    inp = example("srm1155"); inp.n_photons_line = 0; inp.gain = 0.0
    ci = x.CInput(inp)
    assert L.xmb_input_validate(C.byref(ci.input)) == 1 | 32
    assert L.xmb_input_write_to_xml_file(C.byref(ci.input), str(bad).encode()) == 1
    assert L.xmb_input_read_from_xml_file(str(bad).encode(), C.byref(p)) == 0 and "error validating input data" in abi.last_error()


def test_xmsi_round_trip_with_broadened_lines_and_absorbers(tmp_path):
    inp = example("srm1412")                       # has an excitation-path absorber
    inp.discrete[0].distribution_type = 1; inp.discrete[0].scale_parameter = 0.05
    inp.discrete[1].distribution_type = 2; inp.discrete[1].scale_parameter = 0.02
    inp.continuous = [x.ContinuousD(5.0, 1e6, 2e6), x.ContinuousD(9.0, 3e6, 1e6, 0.1, 0.01, 0.2, 0.02)]
    inp.comments = "a < b & c"
    ci = x.CInput(inp)
    out = str(tmp_path / "rt.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), out.encode()) == 1
    p = _read(out)
    a, b = _as_tuple(p), _as_tuple(C.pointer(ci.input))
    assert a["general"] == b["general"] and a["detector"] == b["detector"] and a["ref"] == b["ref"]
    assert np.allclose(np.array(a["disc"], float), np.array(b["disc"], float), rtol=1e-5)    # %g keeps 6 digits
    assert [d[7] for d in a["disc"][:3]] == [1, 2, 0]
    assert np.allclose(np.array(a["cont"], float), np.array(b["cont"], float), rtol=1e-5)
    for k in ("layers", "exc_layers", "det_layers", "crystal"):
        for la, lb in zip(a[k], b[k]):
            assert la[0] == lb[0] and np.allclose(la[1], lb[1], rtol=1e-5)
    # and the Python reader accepts the native writer's file
    again = x.read_xmsi(out)
    assert again.comments == "a < b & c" and len(again.discrete) == len(inp.discrete) and again.exc_layers[0].Z == inp.exc_layers[0].Z
    abi.lib().xmb_input_free(C.byref(p))


def test_xmso_writer_follows_the_reference_layout(tmp_path):
    inp = example("srm1155")
    ci = x.CInput(inp)
    n_int, nch = inp.n_interactions_trajectory, inp.nchannels
    rng = np.random.default_rng(3)
    unconv = np.cumsum(rng.uniform(0, 100, (n_int + 1, nch)), axis=0); unconv[0] = 0
    conv = unconv * 0.9
    vr = np.zeros((100, 385, n_int)); br = np.zeros((100, 385, n_int))
    vr[25, 2, :] = [1e6, 5e4, 700.0, 9.0]          # Fe-KL3
    vr[25, 1, 0] = 5e5                              # Fe-KL2, first order only
    vr[23, 2, 1] = 42.0                             # Cr-KL3, second order only
    vr[25, 383, 0] = 7.0                            # Rayleigh slot: not a line, never written
    rows = (abi.c_double_p * (n_int + 1))(*[C.cast(conv.ctypes.data + i * nch * 8, abi.c_double_p) for i in range(n_int + 1)])
    out = str(tmp_path / "o.xmso")
    ok = abi.lib().xmb_output_write_to_xml_file(C.byref(ci.input), b"in.xmsi", out.encode(), unconv.ctypes.data_as(abi.c_double_p), rows,
                                               br.ctypes.data_as(abi.c_double_p), vr.ctypes.data_as(abi.c_double_p), 0, None)
    assert ok == 1, abi.last_error()
    assert abi.lib().xmb_output_write_to_xml_file(C.byref(ci.input), b"in.xmsi", (out + "2").encode(), unconv.ctypes.data_as(abi.c_double_p),
                                                  rows, None, None, 0, None) == 1            # both histories absent
    assert "<variance_reduction_history/>" in open(out + "2").read()
    txt = open(out).read()
    assert txt.startswith('<?xml version="1.0"?>\n<!DOCTYPE xmimsim-results SYSTEM "http://www.xmi.UGent.be/xml/xmimsim-1.0.dtd">')
    assert "<brute_force_history/>" in txt and '<xmimsim-results version="8.1">' in txt      # VERSION of the reference (src/xmi_xml.c:1465)
    import xml.etree.ElementTree as ET
    root = ET.parse(out).getroot()
    assert [c.tag for c in root] == ["inputfile", "spectrum_conv", "spectrum_unconv", "brute_force_history",
                                     "variance_reduction_history", "xmimsim-input", "svg_graphs"]
    assert root.find("inputfile").text == "in.xmsi"
    chans = root.find("spectrum_conv").findall("channel")
    assert len(chans) == nch and chans[7].find("channelnr").text == "7"
    assert abs(float(chans[7].find("energy").text) - (inp.gain * 7 + inp.zero)) < 1e-6
    got = np.array([[float(c.text) for c in ch.findall("counts")] for ch in chans]).T
    assert got.shape == (n_int, nch) and np.allclose(got, conv[1:], rtol=1e-5)
    assert [c.get("interaction_number") for c in chans[0].findall("counts")] == ["1", "2", "3", "4"]
    got_u = np.array([[float(c.text) for c in ch.findall("counts")] for ch in root.find("spectrum_unconv").findall("channel")]).T
    assert np.allclose(got_u, unconv[1:], rtol=1e-5)
    els = root.find("variance_reduction_history").findall("fluorescence_line_counts")
    assert [(e.get("atomic_number"), e.get("symbol")) for e in els] == [("24", "Cr"), ("26", "Fe")]
    fe = els[1]
    assert abs(float(fe.get("total_counts")) - vr[25, :383].sum()) < 1e-5 * vr[25, :383].sum()
    lines = fe.findall("fluorescence_line")
    assert [l.get("type") for l in lines] == ["KL2", "KL3"]
    assert [c.get("interaction_number") for c in lines[0].findall("counts")] == ["1"]
    assert [float(c.text) for c in lines[1].findall("counts")] == [1e6, 5e4, 700.0, 9.0]
    assert [c.get("interaction_number") for c in els[0].find("fluorescence_line").findall("counts")] == ["2"]
    # the echoed input parses back to the same input with either reader
    p = _read(out)
    assert _as_tuple(p)["detector"] == _as_tuple(C.pointer(ci.input))["detector"]
    abi.lib().xmb_input_free(C.byref(p))


def test_spe_and_csv_files(tmp_path):
    inp = example("srm1155")
    ci = x.CInput(inp)
    nch, n_int = inp.nchannels, inp.n_interactions_trajectory
    spec = np.arange(nch, dtype=np.float64) * 1.5
    spe = str(tmp_path / "s_1.spe")
    assert abi.lib().xmb_write_spe_file(spe.encode(), C.byref(ci.input), spec.ctypes.data_as(abi.c_double_p)) == 1
    lines = open(spe).read().split("\n")
    assert lines[:8] == ["$SPEC_ID:", "", "$MCA_CAL:", "2", "%g %g" % (inp.zero, inp.gain), "", "$DATA:", "0\t%d" % (nch - 1)]
    vals = [float(v) for l in lines[8:] for v in l.split()]
    assert np.allclose(vals, spec) and len(lines[8].split()) == 8
    rows_np = np.stack([spec * k for k in range(n_int + 1)])
    rows = (abi.c_double_p * (n_int + 1))(*[C.cast(rows_np.ctypes.data + i * nch * 8, abi.c_double_p) for i in range(n_int + 1)])
    csv = str(tmp_path / "s.csv")
    assert abi.lib().xmb_write_csv_file(csv.encode(), C.byref(ci.input), rows, 1) == 1
    first = open(csv).readline().strip().split(",")
    assert first[0] == "0" and len(first) == 2 + n_int
    tab = np.loadtxt(csv, delimiter=",")
    assert tab.shape == (nch, 2 + n_int) and np.allclose(tab[:, 2:], rows_np[1:].T, rtol=1e-5)


@pytest.mark.parametrize("broken", ["<xmimsim><general version=1.0></general></xmimsim>", "<xmimsim><a><![CDATA[never closed</a></xmimsim>",
                                    "<xmimsim><general version=\"1.0></general>", "<xmimsim><a>text</a", "<x>" * 200])
def test_malformed_xml_is_an_error_not_a_hang(broken):
    """Unquoted or unterminated attribute values, an unterminated CDATA section or end tag, runaway nesting: the built-in
    parser reports them (the reference leaves this to libxml2)."""
    inp = C.POINTER(abi.Input)()
    assert abi.lib().xmb_input_read_from_xml_string(broken.encode(), C.byref(inp)) == 0
    assert abi.last_error()


def test_xmso_svg_graphs_reproduce_the_shipped_file(tmp_path):
    """The <svg_graphs> block (src/xmi_xml.c:1578-1622, :1880-2022) written from the spectra of the reference's shipped
    examples/srm1155.xmso must reproduce that file's own block: tests/golden/srm1155_svg.npz (tools/make_golden.py) holds its eight
    graphics.  The spectra in the file are printed with six digits, so the coordinates agree to a few 1e-6 of the box, not to the bit."""
    inp = example("srm1155")
    ci = x.CInput(inp)
    n_int, nch = inp.n_interactions_trajectory, inp.nchannels
    g = np.load(os.path.join(GOLDEN, "srm1155_xmso.npz"))
    svg = np.load(os.path.join(GOLDEN, "srm1155_svg.npz"))
    conv = np.zeros((n_int + 1, nch)); conv[1:] = g["conv"]
    unconv = np.zeros((n_int + 1, nch)); unconv[1:] = g["unconv"]
    rows = (abi.c_double_p * (n_int + 1))(*[C.cast(conv.ctypes.data + i * nch * 8, abi.c_double_p) for i in range(n_int + 1)])
    out = str(tmp_path / "svg.xmso")
    assert abi.lib().xmb_output_write_to_xml_file(C.byref(ci.input), b"in.xmsi", out.encode(), unconv.ctypes.data_as(abi.c_double_p), rows,
                                                  None, None, 0, None) == 1, abi.last_error()
    import xml.etree.ElementTree as ET
    graphics = ET.parse(out).getroot().find("svg_graphs").findall("graphic")
    assert len(graphics) == 2 * n_int
    for gi, gr in enumerate(graphics):
        kind, order = svg["g%d_id" % gi]
        assert gr.find("id/name").text == ("convoluted" if kind == 0 else "unconvoluted") and int(gr.find("id/interaction").text) == order
        size = gr.find("rect/size")
        assert gr.find("rect/view") is not None
        box = np.array([float(size.find(k).text) for k in ("width", "height", "min_energy", "max_energy")])
        assert np.allclose(box, svg["g%d_box" % gi], rtol=1e-5)
        xt = np.array([[float(i.find("value").text), float(i.find("name").text)] for i in gr.find("rect/x-axis").findall("index")])
        yt = np.array([[float(i.find("value").text), float(i.find("name").text)] for i in gr.find("rect/y-axis").findall("index")])
        assert xt.shape == svg["g%d_xt" % gi].shape and np.allclose(xt, svg["g%d_xt" % gi], rtol=1e-5, atol=1e-3)
        assert yt.shape == svg["g%d_yt" % gi].shape and np.allclose(yt, svg["g%d_yt" % gi], rtol=1e-5, atol=1e-3)
        assert gr.find("rect/x-axis/name").text == "Energy (keV)" and gr.find("rect/y-axis/name").text == "Intensity (counts)"
        assert gr.find("points/color").text == "blue"
        pts = np.array([[float(p.find("x").text), float(p.find("y").text)] for p in gr.find("points").findall("point")])
        want = svg["g%d_pts" % gi]
        assert pts.shape == want.shape, (gi, pts.shape, want.shape)
        assert np.abs(pts - want).max() < 2e-3, (gi, np.abs(pts - want).max())     # of a 500 x 250 box
