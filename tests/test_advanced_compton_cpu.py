"""CPU tests of the shell-resolved ("advanced") Compton option: table construction
(src/xmi_data_f.F90:1120-1235) and the oracle's energy sampling (src/xmi_main.F90:4785-4983)."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from helpers import Pair
from inputs import close_detector


def test_subshell_tables_are_cdfs_and_inverses():
    inp = close_detector(1000, 1)
    sim = x.Simulation(inp)
    assert sim.tables.n_adv_rows == 0
    assert sim.L.xmb_tables_enable_advanced_compton(sim.hdf5F) == 1
    assert sim.L.xmb_tables_enable_advanced_compton(sim.hdf5F) == 1          # idempotent
    T = sim.tables
    nZ, n_cp, rows = T.nZ, T.n_cp, T.n_adv_rows
    off = [T.adv_off[i] for i in range(nZ + 1)]
    assert off[0] == 0 and off[-1] == rows and all(b > a for a, b in zip(off, off[1:]))
    cdf = np.ctypeslib.as_array(T.adv_cdf, shape=(rows, n_cp))
    qinv = np.ctypeslib.as_array(T.adv_qinv, shape=(rows, n_cp))
    cfg = np.ctypeslib.as_array(T.adv_config, shape=(rows,))
    shells = [T.adv_shell[r] for r in range(rows)]
    for zi in range(nZ):
        assert shells[off[zi]] == 0                                           # K first
        assert abs(cfg[off[zi]:off[zi + 1]].sum() - T.Z[zi]) < 1e-12          # electron configuration sums to Z
    assert np.all(cdf[:, 0] == 0) and np.allclose(cdf[:, -1], 0.5) and np.all(np.diff(cdf, axis=1) >= 0)
    assert np.all(qinv[:, 0] == 0) and np.all(np.diff(qinv, axis=1) >= 0) and np.all(qinv <= 100.0)
    # inverse of the forward table: cdf(qinv(c)) == c on the interior
    dq = 100.0 / (n_cp - 1)
    for r in (0, rows // 2, rows - 1):
        c = 0.5 * np.arange(n_cp) / (n_cp - 1)
        pos = np.minimum((qinv[r] / dq).astype(int), n_cp - 2)
        back = cdf[r, pos] + (cdf[r, pos + 1] - cdf[r, pos]) * (qinv[r] - dq * pos) / dq
        sel = (c > 0.01) & (c < 0.49)
        assert np.abs(back[sel] - c[sel]).max() < 2e-4
    sim.close()


def test_advanced_compton_changes_only_the_compton_part():
    """Same photons, same streams: fluorescence and Rayleigh deposits of the first order are untouched by the option;
    the Compton deposit keeps its total (the subshell weights sum to one) up to the attenuation of the differently
    distributed scattered energies."""
    inp = close_detector(20000, 1)
    P = Pair(inp)
    r_full, t_full = P.sim.solid_angle_inputs()
    n = 16
    r = np.linspace(r_full[0], r_full[-1], n); t = np.linspace(t_full[0], t_full[-1], n)
    sa_g, _ = orc.solid_angle_grid(P.od, r, np.arange(n), t, np.arange(n), n, 2000, 1, n_threads=8)
    sa = P.sim.make_solid_angle(sa_g, r, t)
    ch0, vr0, _ = P.oracle(x.main_options(), sa, 0, n_threads=8)
    assert P.sim.L.xmb_tables_enable_advanced_compton(P.sim.hdf5F) == 1
    ch1, vr1, _ = P.oracle(x.main_options(use_advanced_compton=1), sa, 0, n_threads=8)
    assert np.allclose(vr0[:, :384, 0], vr1[:, :384, 0], rtol=1e-11, atol=0)   # lines + Rayleigh, order 1 (thread-sum order only)
    c0, c1 = vr0[25, 384, 0], vr1[25, 384, 0]
    assert c0 > 0 and abs(c1 / c0 - 1.0) < 0.03          # same cross section; only the escape path sees the other energy distribution
    # the scattered-photon spectrum stays (almost entirely) below the incident energy
    comp = ch1[1] - ch0[1]
    e = inp.zero + inp.gain * np.arange(inp.nchannels)
    assert np.abs(comp[e > 20.05]).sum() < 1e-4 * c1     # Doppler up-shift beyond the line is possible but rare (no E <= E0 retry here)
    P.close()
