"""End-to-end run of the command-line driver (bin/xmimsim-b200, the replacement for bin/xmimsim.c): XMSI in, XMSO /
SPE / CSV out, compared with the same pipeline driven through the Python mirror."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import xmimsim_b200 as x
from xmimsim_b200 import abi
from inputs import example

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bin", "xmimsim-b200")
SURR = "--surrogate-cross-sections"     # no xraylib in the image: the analytic stand-in has to be asked for


def test_cli_matches_library_pipeline(tmp_path):
    inp = example("srm1155")
    inp.n_photons_line = 2000
    inp.outputfile = str(tmp_path / "out.xmso")
    ci = x.CInput(inp)
    xmsi = str(tmp_path / "in.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode()) == 1
    r = subprocess.run([CLI, SURR, "-v", "--table-quality=0", "--csv-file=" + str(tmp_path / "c.csv"), "--spe-file-unconvoluted=" + str(tmp_path / "u"), xmsi],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for needle in ("Inputfile %s successfully parsed" % xmsi, "Precalculating solid angle grid", "Solid angle calculation finished",
                   "Simulating interactions", "Interactions simulation finished", "Escape peak ratios calculation finished",
                   "Output written to XMSO file"):
        assert needle in out, (needle, out)
    # the same pipeline through the library (the file round trip keeps 6 significant digits of the input)
    inp2 = x.read_xmsi(xmsi)
    sim = x.Simulation(inp2, quality=0)
    sim.solid_angle_calculation(hits_per_single=5000, seed=0)
    ch, br, vr = sim.main_msim(x.main_options())
    er = sim.escape_ratios_calculation()
    raw = ch.copy()
    conv = sim.detector_convolute_all(ch, br, vr, x.main_options(), er.contents)
    res = x.read_xmso(inp.outputfile)
    n_int = inp.n_interactions_trajectory
    assert res["conv"].shape == (n_int, inp.nchannels)
    # <spectrum_unconv> holds the rows AFTER the response corrected them in place (detector absorbers, crystal efficiency,
    # escape peaks), as the reference writes channelsdef after xmi_detector_convolute_all (bin/xmimsim.c:498-642)
    assert np.allclose(res["unconv"], ch[1:], rtol=2e-5, atol=1e-6 * ch.max())
    assert not np.allclose(res["unconv"], raw[1:], rtol=1e-3, atol=1e-6 * raw.max())
    assert "NOT physics-grade" in r.stderr
    assert np.allclose(res["conv"], conv[1:], rtol=2e-5, atol=1e-6 * conv.max())
    fe = res["history"][(26, "KL3")]
    assert abs(fe["counts"][1] - vr[25, 2, 0]) <= 1e-5 * vr[25, 2, 0]
    csv = np.loadtxt(str(tmp_path / "c.csv"), delimiter=",")
    assert csv.shape == (inp.nchannels, 2 + n_int) and np.allclose(csv[:, 2:], conv[1:].T, rtol=2e-5, atol=1e-6 * conv.max())
    assert os.path.exists(str(tmp_path / "u_1.spe")) and os.path.exists(str(tmp_path / "u_4.spe")) and not os.path.exists(str(tmp_path / "u_0.spe"))
    sim.escape_ratios_free(er)
    sim.close()


def test_cli_brute_force_and_errors(tmp_path):
    from inputs import close_detector
    inp = close_detector(n_photons=200000, n_int=2)
    inp.outputfile = str(tmp_path / "b.xmso")
    ci = x.CInput(inp)
    xmsi = str(tmp_path / "b.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode()) == 1
    r = subprocess.run([CLI, SURR, "--disable-variance-reduction", "--disable-escape-peaks", "--table-quality=0", xmsi], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    import xml.etree.ElementTree as ET
    root = ET.parse(inp.outputfile).getroot()
    assert root.find("variance_reduction_history").find("fluorescence_line_counts") is None
    fe = [e for e in root.find("brute_force_history").findall("fluorescence_line_counts") if e.get("atomic_number") == "26"]
    assert fe and float(fe[0].get("total_counts")) > 0
    r = subprocess.run([CLI, SURR, str(tmp_path / "missing.xmsi")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not read" in r.stderr
    r = subprocess.run([CLI, "--no-such-option", xmsi], capture_output=True, text=True)
    assert r.returncode == 1


def test_cli_caches_are_filled_then_reused(tmp_path):
    inp = example("srm1155")
    inp.n_photons_line = 500
    inp.outputfile = str(tmp_path / "c1.xmso")
    ci = x.CInput(inp)
    xmsi = str(tmp_path / "c.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode()) == 1
    sa, er = str(tmp_path / "sa.cache"), str(tmp_path / "er.cache")
    cmd = [CLI, SURR, "-v", "--table-quality=0", "--with-solid-angles-data=" + sa, "--with-escape-ratios-data=" + er, xmsi]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr
    assert "Precalculating solid angle grid" in r1.stdout and "was successfully updated with new solid angle grid" in r1.stdout
    assert "Precalculating escape peak ratios" in r1.stdout and "was successfully updated with new escape peak ratios" in r1.stdout
    first = x.read_xmso(inp.outputfile)
    r2 = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr
    assert "Solid angle grid already present in" in r2.stdout and "Escape peak ratios already present in" in r2.stdout
    assert "Precalculating" not in r2.stdout
    second = x.read_xmso(inp.outputfile)
    assert np.array_equal(first["conv"], second["conv"]) and np.array_equal(first["unconv"], second["unconv"])   # cached == recomputed


def test_cli_custom_detector_response_plugin(tmp_path):
    """tests/test-custom-detector-response.c of the reference: brute force, no escape peaks, response taken from a plugin
    (here the library's own xmi_detector_convolute_all_custom); the result equals the built-in response."""
    from inputs import close_detector
    inp = close_detector(n_photons=300000, n_int=2)
    ci = x.CInput(inp)
    outs = []
    for k, extra in enumerate(([], ["--custom-detector-response=" + abi.LIB_PATH])):
        inp.outputfile = str(tmp_path / ("p%d.xmso" % k))
        ci = x.CInput(inp)
        xmsi = str(tmp_path / ("p%d.xmsi" % k))
        assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode()) == 1
        r = subprocess.run([CLI, SURR, "-v", "--disable-variance-reduction", "--disable-escape-peaks", "--table-quality=0"] + extra + [xmsi],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert ("xmi_detector_convolute_all_custom loaded from" in r.stdout) == bool(extra)
        outs.append(x.read_xmso(inp.outputfile))
    assert outs[0]["conv"].sum() > 0 and np.array_equal(outs[0]["conv"], outs[1]["conv"])
    r = subprocess.run([CLI, SURR, "--custom-detector-response=/nonexistent.so", xmsi], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open" in r.stderr


def test_cli_refuses_to_run_without_xraylib_unless_the_stand_in_is_requested(tmp_path):
    inp = example("srm1155")
    inp.n_photons_line = 10
    inp.outputfile = str(tmp_path / "n.xmso")
    ci = x.CInput(inp)
    xmsi = str(tmp_path / "n.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode()) == 1
    r = subprocess.run([CLI, "--with-xraylib=/nonexistent/libxrl.so", xmsi], capture_output=True, text=True)
    assert r.returncode == 1 and "xraylib is required" in r.stderr and not os.path.exists(inp.outputfile)
