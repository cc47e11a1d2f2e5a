"""Host-side checks on the random inputs of tests/random_inputs.py (the GPU parity tests use the same generator): every input
passes the validator, survives an XMSI round trip through the native writer and reader, gives the same derived geometry in the
product's xmb_init_input and the oracle's restatement bit for bit, and the same solid-angle axes."""
import ctypes as C

import numpy as np
import pytest

import orc
import xmimsim_b200 as x
from xmimsim_b200 import abi
from random_inputs import random_input
from test_io_cpu import _as_tuple, _read


@pytest.mark.parametrize("seed", list(range(0, 60, 3)))
def test_random_input_is_valid_and_round_trips(seed, tmp_path):
    inp, _ = random_input(seed)
    ci = x.CInput(inp)
    assert abi.lib().xmb_input_validate(C.byref(ci.input)) == 0
    out = str(tmp_path / "r.xmsi")
    assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), out.encode()) == 1
    p = _read(out)
    a, b = _as_tuple(p), _as_tuple(C.pointer(ci.input))
    assert a["general"] == b["general"] and a["ref"] == b["ref"]
    assert np.allclose(np.array(a["detector"], float), np.array(b["detector"], float), rtol=1e-5)
    assert np.allclose(np.array(a["disc"], float), np.array(b["disc"], float), rtol=1e-5)        # %g keeps 6 digits
    if b["cont"]:
        assert np.allclose(np.array(a["cont"], float), np.array(b["cont"], float), rtol=1e-5)
    g = lambda t: np.concatenate([[t[0]], t[1], t[2], t[3], t[4:]])  # noqa: E731
    assert np.allclose(g(a["geometry"]), g(b["geometry"]), rtol=1e-5, atol=1e-12)
    for k in ("layers", "exc_layers", "det_layers", "crystal"):
        assert len(a[k]) == len(b[k])
        for la, lb in zip(a[k], b[k]):
            assert la[0] == lb[0] and np.allclose(la[1], lb[1], rtol=1e-5) and np.allclose(la[2:], lb[2:], rtol=1e-5)
    abi.lib().xmb_input_free(C.byref(p))


@pytest.mark.parametrize("seed", list(range(1, 40, 4)))
def test_random_input_derived_geometry_and_axes_match_oracle(seed):
    inp, _ = random_input(seed)
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    d = sim.derived
    for f in ("detector_radius", "collimator_present", "collimator_radius", "half_apex", "detector_solid_angle"):
        assert getattr(d, f) == getattr(od, f), f
    assert list(d.ndo_new) == list(od.ndo_new) and list(d.ndo_inv) == list(od.ndo_inv)
    assert list(d.n_sample_orientation_det) == list(od.n_sample_orientation_det)
    for i in range(d.n_layers):
        assert d.Z_coord_begin[i] == od.Z_coord_begin[i] and d.Z_coord_end[i] == od.Z_coord_end[i]
    r, t = sim.solid_angle_inputs()
    r_o, t_o = orc.solid_angle_axes(C.pointer(ci.input), od)
    assert np.array_equal(r, r_o) and np.array_equal(t, t_o)
    assert np.all(np.diff(r) > 0) and np.all(np.diff(t) > 0)
    sim.close()
