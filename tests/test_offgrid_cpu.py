"""Oracle-only check of the off-grid branch of xmi_get_solid_angle (src/xmi_solid_angle_f.F90:783-789): a grid cut
just behind the sample surface must give the same intensities as the full grid within the Monte Carlo error of the
on-the-spot solid angles -- the branch computes the quantity the grid tabulates."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from inputs import example


def _oracle_run(inp, sim, ci, od, sa, n_total):
    opt = x.main_options()
    ch, vr, cnt = orc.main_msim_range(C.pointer(ci.input), od, sim.L.xmb_get_tables(sim.hdf5F), opt, sa, 0x584D494D53494D, 0,
                                      n_total, inp.n_interactions_trajectory, inp.nchannels, 8)
    return ch, vr, cnt


def test_offgrid_points_are_computed_on_the_spot():
    inp = example("srm1155")
    inp.n_photons_line = 40
    inp.n_interactions_trajectory = 2
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    n_total = orc.lib().orc_total_histories(C.cast(C.pointer(ci.input), C.c_void_p))
    r_full, t_full = sim.solid_angle_inputs()
    n = 48
    t = np.linspace(t_full[0], t_full[-1], n)
    r = np.linspace(r_full[0], r_full[-1], n)
    g_full, _ = orc.solid_angle_grid(od, r, np.arange(n), t, np.arange(n), n, 1500, 5)
    sa_full = sim.make_solid_angle(g_full.copy(), r.copy(), t.copy())
    ch_full, vr_full, cnt_full = _oracle_run(inp, sim, ci, od, sa_full, n_total)
    assert cnt_full[0] < 0.01 * cnt_full[1]          # the reference's own grid leaves ~0.1 % of the points (air-path scatters) outside
    pw = np.array(inp.p_detector_window, float)
    r_cut = np.linalg.norm(pw - np.array([0.0, 0.0, inp.d_sample_source])) * (1.0 + 1e-6)
    r2 = np.linspace(r_full[0], r_cut, n)
    g_cut, _ = orc.solid_angle_grid(od, r2, np.arange(n), t, np.arange(n), n, 1500, 5)
    sa_cut = sim.make_solid_angle(g_cut.copy(), r2.copy(), t.copy())
    orc.lib().orc_set_hits_per_single(1500)
    try:
        ch_cut, vr_cut, cnt_cut = _oracle_run(inp, sim, ci, od, sa_cut, n_total)
    finally:
        orc.lib().orc_set_hits_per_single(5000)
    assert 0.05 * cnt_cut[1] < cnt_cut[0] < 0.98 * cnt_cut[1]
    ratio = ch_cut[-1].sum() / ch_full[-1].sum()
    assert 0.93 < ratio < 1.07, ratio
    sim.close()
