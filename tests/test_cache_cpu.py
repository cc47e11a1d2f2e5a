"""The solid-angle / escape-ratio caches (side-car container with the reference's logical schema and match rules,
src/xmi_solid_angle.c:193-790, src/xmi_detector.c:143-435)."""
import copy
import ctypes as C

import numpy as np

import xmimsim_b200 as x
from xmimsim_b200 import abi
from inputs import example


def _xml(ci):
    s = C.c_void_p()
    assert abi.lib().xmb_input_write_to_xml_string(C.byref(ci.input), C.byref(s)) == 1
    return s


def test_xml_string_round_trip():
    ci = x.CInput(example("srm1412"))
    s = _xml(ci)
    txt = C.string_at(s).decode()
    assert txt.startswith('<?xml version="1.0"?>') and "<xmimsim>" in txt and "<excitation_path>" in txt
    p = C.POINTER(abi.Input)()
    assert abi.lib().xmb_input_read_from_xml_string(txt.encode(), C.byref(p)) == 1
    assert p.contents.composition.contents.n_layers == 2 and p.contents.excitation.contents.n_discrete == 25
    abi.lib().xmb_input_free(C.byref(p))
    C.CDLL(None).free(s)


def test_solid_angle_match_rules():
    L = abi.lib()
    base = example("srm1155")
    A = x.CInput(base)
    assert L.xmb_check_solid_angle_match(C.byref(A.input), C.byref(A.input), None) == 1
    for change, expect in ((dict(area_detector=base.area_detector * 1.01), 0), (dict(collimator_height=base.collimator_height + 0.1), 0),
                           (dict(p_detector_window=[0.0, base.p_detector_window[1] - 0.1, base.p_detector_window[2]]), 0),
                           (dict(n_photons_line=7), 1), (dict(live_time=5.0), 1),
                           # same detector-to-sample offset: moving source distance and window together keeps the grid
                           (dict(d_sample_source=base.d_sample_source + 10, p_detector_window=[base.p_detector_window[0], base.p_detector_window[1], base.p_detector_window[2] + 10]), 1)):
        other = copy.deepcopy(base)
        for k, v in change.items():
            setattr(other, k, v)
        B = x.CInput(other)
        assert L.xmb_check_solid_angle_match(C.byref(A.input), C.byref(B.input), None) == expect, change
    # a harder source probes deeper than the cached grid covers: no match; a softer one is covered
    hard = copy.deepcopy(base); hard.discrete = [x.DiscreteD(60.0, 1e9, 1e9)]
    soft = copy.deepcopy(base); soft.discrete = [x.DiscreteD(15.9, 1e9, 1e9), x.DiscreteD(16.1, 1e9, 1e9)]
    assert L.xmb_check_solid_angle_match(C.byref(A.input), C.byref(x.CInput(hard).input), None) == 0
    assert L.xmb_check_solid_angle_match(C.byref(A.input), C.byref(x.CInput(soft).input), None) == 1


def test_solid_angle_cache_file(tmp_path):
    L = abi.lib()
    path = str(tmp_path / "sa.cache").encode()
    inp = example("srm1155")
    ci = x.CInput(inp)
    rv = C.POINTER(abi.SolidAngle)()
    assert L.xmb_find_solid_angle_match(path, C.byref(ci.input), None, C.byref(rv), None) == 1 and not rv     # no file yet: empty cache
    rng = np.random.default_rng(0)
    grids = []
    for k, name in enumerate(("srm1155", "srm1132")):
        c = x.CInput(example(name))
        g = rng.uniform(0, 1, (6 + k, 5)); r = np.linspace(0.1, 2, 5); t = np.linspace(1e-5, 1.57, 6 + k)
        sa = abi.SolidAngle(g.ctypes.data_as(abi.c_double_p), 5, 6 + k, r.ctypes.data_as(abi.c_double_p), t.ctypes.data_as(abi.c_double_p), _xml(c))
        assert L.xmb_update_solid_angle_cache_file(path, C.byref(sa)) == 1, abi.last_error()
        grids.append((g, r, t))
    for k, name in enumerate(("srm1155", "srm1132")):
        c = x.CInput(example(name))
        rv = C.POINTER(abi.SolidAngle)()
        assert L.xmb_find_solid_angle_match(path, C.byref(c.input), None, C.byref(rv), None) == 1 and rv
        s = rv.contents
        g, r, t = grids[k]
        assert (s.grid_dims_theta_n, s.grid_dims_r_n) == g.shape
        assert np.array_equal(np.ctypeslib.as_array(s.solid_angles, shape=g.shape), g)
        assert np.array_equal(np.ctypeslib.as_array(s.grid_dims_r_vals, shape=r.shape), r)
        assert np.array_equal(np.ctypeslib.as_array(s.grid_dims_theta_vals, shape=t.shape), t)
        assert b"<xmimsim>" in C.string_at(s.xmi_input_string)
        L.xmb_free_solid_angle(rv)
    other = example("In"); other.area_detector *= 3
    rv = C.POINTER(abi.SolidAngle)()
    assert L.xmb_find_solid_angle_match(path, C.byref(x.CInput(other).input), None, C.byref(rv), None) == 1 and not rv
    # a file of the other kind is refused (the reference checks the "kind" attribute)
    er = C.POINTER(abi.EscapeRatios)()
    assert L.xmb_find_escape_ratios_match(path, C.byref(ci.input), C.byref(er), None) == 0 and "kind" in abi.last_error()
    (tmp_path / "junk").write_bytes(b"not a cache")
    assert L.xmb_find_solid_angle_match(str(tmp_path / "junk").encode(), C.byref(ci.input), None, C.byref(rv), None) == 0


def test_escape_ratio_cache_file(tmp_path):
    L = abi.lib()
    path = str(tmp_path / "er.cache").encode()
    inp = example("srm1155")
    ci = x.CInput(inp)
    sim = x.Simulation(inp)
    nE, nO, nZ = 7, 11, 1
    fluo = np.random.default_rng(1).uniform(0, 1e-2, (nE, 109, nZ)); compt = np.random.default_rng(2).uniform(0, 1e-3, (nO, nE))
    e_in = 1.0 + 0.5 * np.arange(nE); e_out = 0.1 + 0.2 * np.arange(nO)
    er = sim.make_escape_ratios([14], fluo, e_in, compt, e_in, e_out)
    er.xmi_input_string = _xml(ci)
    assert L.xmb_update_escape_ratios_cache_file(path, C.byref(er)) == 1, abi.last_error()
    # any input with the same crystal matches, whatever the sample
    other = example("srm1412")
    assert other.crystal_layers[0].Z == inp.crystal_layers[0].Z
    same_crystal = L.xmb_check_escape_ratios_match(C.byref(ci.input), C.byref(x.CInput(other).input))
    rv = C.POINTER(abi.EscapeRatios)()
    assert L.xmb_find_escape_ratios_match(path, C.byref(x.CInput(other).input), C.byref(rv), None) == 1
    assert bool(rv) == bool(same_crystal)
    rv = C.POINTER(abi.EscapeRatios)()
    assert L.xmb_find_escape_ratios_match(path, C.byref(ci.input), C.byref(rv), None) == 1 and rv
    e = rv.contents
    assert (e.n_elements, e.n_fluo_input_energies, e.n_compton_input_energies, e.n_compton_output_energies) == (nZ, nE, nE, nO)
    assert np.array_equal(np.ctypeslib.as_array(e.fluo_escape_ratios, shape=fluo.shape), fluo)
    assert np.array_equal(np.ctypeslib.as_array(e.compton_escape_ratios, shape=compt.shape), compt)
    assert np.array_equal(np.ctypeslib.as_array(e.compton_escape_output_energies, shape=e_out.shape), e_out)
    assert np.array_equal(np.ctypeslib.as_array(e.compton_escape_input_energies, shape=e_in.shape), e_in) and e.Z[0] == 14
    L.xmb_free_escape_ratios(C.byref(rv))
    thick = copy.deepcopy(inp); thick.crystal_layers[0].thickness *= 2
    rv = C.POINTER(abi.EscapeRatios)()
    assert L.xmb_find_escape_ratios_match(path, C.byref(x.CInput(thick).input), C.byref(rv), None) == 1 and not rv
    sim.close()


def test_cache_entries_are_bound_to_their_cross_section_provider(tmp_path):
    """An entry computed with one provider never serves a run on another: the grid axes and the escape ratios depend on
    the cross sections (the reference has a single provider, xraylib; here the analytic stand-in exists beside it)."""
    L = abi.lib()
    path = str(tmp_path / "sa.cache").encode()
    c = x.CInput(example("srm1155"))
    g = np.ones((6, 5)); r = np.linspace(0.1, 2, 5); t = np.linspace(1e-5, 1.57, 6)
    sa = abi.SolidAngle(g.ctypes.data_as(abi.c_double_p), 5, 6, r.ctypes.data_as(abi.c_double_p), t.ctypes.data_as(abi.c_double_p), _xml(c))
    surrogate = L.xmb_xrl_surrogate()
    other = abi.XrlProvider.from_buffer_copy(surrogate.contents)      # same functions under another name
    other.name = b"some other provider"
    try:
        L.xmb_cache_set_provider(surrogate)
        assert L.xmb_update_solid_angle_cache_file(path, C.byref(sa)) == 1, abi.last_error()
        rv = C.POINTER(abi.SolidAngle)()
        assert L.xmb_find_solid_angle_match(path, C.byref(c.input), None, C.byref(rv), None) == 1 and rv
        L.xmb_free_solid_angle(rv)
        L.xmb_cache_set_provider(C.byref(other))
        rv = C.POINTER(abi.SolidAngle)()
        assert L.xmb_find_solid_angle_match(path, C.byref(c.input), None, C.byref(rv), None) == 1 and not rv
        assert L.xmb_update_solid_angle_cache_file(path, C.byref(sa)) == 1
        assert L.xmb_find_solid_angle_match(path, C.byref(c.input), None, C.byref(rv), None) == 1 and rv      # its own entry, behind the first
        L.xmb_free_solid_angle(rv)
    finally:
        L.xmb_cache_set_provider(None)
