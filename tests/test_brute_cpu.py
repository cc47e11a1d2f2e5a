"""CPU tests of the brute-force restatement (use_variance_reduction = 0): the analogue random walk with real
detector hits must agree statistically with the forced-detection estimator of the same quantities -- the
reference's own validation of its variance reduction and a check of both restatements."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from helpers import Pair, DEFAULT_SEED
from inputs import close_detector, example


def _brute(P, opt, n_threads=8, g1=None):
    T = P.sim.L.xmb_get_tables(P.sim.hdf5F)
    ch, br, cnt = orc.main_msim_brute_range(C.pointer(P.ci.input), P.od, T, opt, DEFAULT_SEED, 0, g1 or P.n_total,
                                            P.inp.n_interactions_trajectory, P.inp.nchannels, n_threads)
    return ch * P.inp.live_time, br * P.inp.live_time, cnt


def test_brute_force_agrees_with_forced_detection():
    inp = close_detector(n_photons=3000000, n_int=2)
    P = Pair(inp)
    r_full, t_full = P.sim.solid_angle_inputs()
    n = 24
    r = np.linspace(r_full[0], r_full[-1], n); t = np.linspace(t_full[0], t_full[-1], n)
    sa_g, _ = orc.solid_angle_grid(P.od, r, np.arange(n), t, np.arange(n), n, 100000, 1, n_threads=8)
    sa = P.sim.make_solid_angle(sa_g, r, t)
    kw = dict(use_M_lines=0, use_cascade_auger=0, use_cascade_radiative=0)
    # (a prefix of the photon ids would be biased: polarisation is assigned by photon index, src/xmi_main.F90:682)
    ch_v, vr, _ = P.oracle(x.main_options(**kw), sa, 0, n_threads=8)
    scale = 1.0
    ch_b, br, cnt = _brute(P, x.main_options(use_variance_reduction=0, **kw))
    hits = int(cnt[0])
    assert hits > 3000
    w = inp.discrete[0].horizontal_intensity + inp.discrete[0].vertical_intensity
    w /= inp.n_photons_line                            # weight of one detected photon
    # total detected intensity per order: Poisson error of the brute-force count, 5 sigma
    for k in (1, 2):
        a = (ch_v[k] - ch_v[k - 1]).sum() * scale; b = (ch_b[k] - ch_b[k - 1]).sum()
        assert abs(a - b) < 5 * np.sqrt(b * w) + 0.01 * a, (k, a, b)
    # strongest lines and the two scatter slots of the matrix element, first order
    for slot in (2, 1, 383, 384):                      # Fe KL3, KL2, Rayleigh, Compton
        a = vr[25, slot, 0] * scale; b = br[25, slot, 0]
        assert b > 0 and abs(a - b) < 5 * np.sqrt(b * w) + 0.01 * a, (slot, a, b)
    assert ch_b[0].sum() == 0.0                        # the beam itself never points at the detector
    assert np.all(np.diff(ch_b, axis=0) >= 0)          # rows cumulative
    P.close()


def test_brute_force_cascades_add_offspring_lines():
    """Radiative + Auger cascades (src/xmi_main.F90:2413-4783): a K vacancy in a heavy element hands L vacancies
    to offspring photons; without cascades L lines only come from direct L ionisation."""
    inp = close_detector(n_photons=1500000, n_int=1)
    inp.layers = [x.LayerD([50], [1.0], 7.3, 0.05)]    # Sn: K edge 29.2 keV, L lines 3-4 keV
    inp.discrete = [x.DiscreteD(40.0, 1e9, 1e9)]
    P = Pair(inp)
    base = dict(use_variance_reduction=0, use_M_lines=0)
    ch0, br0, c0 = _brute(P, x.main_options(use_cascade_auger=0, use_cascade_radiative=0, **base))
    ch1, br1, c1 = _brute(P, x.main_options(use_cascade_auger=1, use_cascade_radiative=1, **base))
    assert int(c0[2]) == 0 and int(c1[2]) > 1000       # offspring photons walked
    L0 = br0[49, 29:113, 0].sum(); L1 = br1[49, 29:113, 0].sum(); K0 = br0[49, :29, 0].sum(); K1 = br1[49, :29, 0].sum()
    assert K0 > 0 and L0 > 0
    assert L1 > 1.3 * L0                               # cascade-fed L emission
    assert abs(K1 - K0) < 6 * np.sqrt(K0 * 2e9 / inp.n_photons_line)   # K emission itself is unchanged (statistically)
    # the same photon ids give the same answer for any thread count
    ch2, br2, c2 = _brute(P, x.main_options(use_cascade_auger=1, use_cascade_radiative=1, **base), n_threads=3)
    assert np.array_equal(c1, c2) and np.allclose(br1, br2, rtol=1e-12, atol=0)
    P.close()


def test_detector_segment_test_with_collimator():
    """xmi_check_detector_intersection (src/xmi_aux_f.F90:1622-1833) through the brute-force walk: with the shipped
    conical collimator photons reach the detector only through its aperture, so far fewer hits than without."""
    inp = close_detector(n_photons=600000, n_int=1)
    P0 = Pair(inp)
    opt = x.main_options(use_variance_reduction=0, use_M_lines=0, use_cascade_auger=0, use_cascade_radiative=0)
    _, _, c0 = _brute(P0, opt)
    inp2 = close_detector(n_photons=600000, n_int=1)
    inp2.collimator_height = 0.5; inp2.collimator_diameter = 0.6
    P1 = Pair(inp2)
    _, _, c1 = _brute(P1, opt)
    assert int(c0[0]) > 500 and 0 < int(c1[0]) < 0.5 * int(c0[0])
    P0.close(); P1.close()
