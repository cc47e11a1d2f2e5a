"""Pins the forced-detection estimator independently of its restatement: the first-order Ca-KL3 intensity of the
reference's CaSO4 known-answer set-up (tests/libxmimsim-test.c:182-285; the reference asserts its "XAS tool" value
1.725e6 within 1 %, tests/test-xmimsim-main-CaSO4.c:19-20) computed by direct quadrature with the same provider
(tests/fp_closed_form.py: no Monte Carlo, no table bundle, no solid-angle grid) must be reproduced by the ORACLE's
history loop within 1 %.  tests/test_fundamental_parameter_gpu.py asks the same of the engine."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from fp_closed_form import first_order_line_intensity, KL3_LINE
from inputs import caso4, close_detector


def _oracle_line(inp, Z, n_grid=96, hits=3000):
    sim = x.Simulation(inp, quality=0)          # host tables only: no GPU call is made here
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    n_total = orc.lib().orc_total_histories(C.cast(C.pointer(ci.input), C.c_void_p))
    r_full, t_full = sim.solid_angle_inputs()
    r = np.linspace(r_full[0], r_full[-1], n_grid); t = np.linspace(t_full[0], t_full[-1], n_grid)
    g, _ = orc.solid_angle_grid(od, r, np.arange(n_grid), t, np.arange(n_grid), n_grid, hits, 5)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    ch, vr, cnt = orc.main_msim_range(C.pointer(ci.input), od, sim.L.xmb_get_tables(sim.hdf5F), x.main_options(), sa,
                                      0x584D494D53494D, 0, n_total, inp.n_interactions_trajectory, inp.nchannels, 8)
    prov = sim.provider.contents
    out = vr[Z - 1, KL3_LINE - 1, 0] * inp.live_time, prov
    sim.close()
    return out


def test_caso4_ca_kl3_first_order_matches_the_quadrature():
    inp = caso4()
    inp.n_photons_line = 40000
    got, prov = _oracle_line(inp, 20)
    want = first_order_line_intensity(inp, prov, 20)
    physical = first_order_line_intensity(inp, prov, 20, estimator=False)
    assert abs(got / want - 1.0) < 0.01, (got, want)
    # the reference's area-uniform detector point is a good approximation of the solid-angle weighted mean here
    assert abs(want / physical - 1.0) < 0.01, (want, physical)
    # and the quadrature itself is converged
    fine = first_order_line_intensity(inp, prov, 20, n_s=192, n_rad=48, n_phi=96)
    assert abs(fine / want - 1.0) < 1e-4


def test_close_detector_fe_kl3_first_order_matches_the_quadrature():
    """A second geometry (3 cm2 window 2 cm from a one-layer steel slab, no air gap) so that the agreement is not a
    property of one set-up: here the solid angle varies by tens of percent over the interaction depth."""
    inp = close_detector(n_photons=40000, n_int=1)
    got, prov = _oracle_line(inp, 26)
    want = first_order_line_intensity(inp, prov, 26)
    assert abs(got / want - 1.0) < 0.01, (got, want)
