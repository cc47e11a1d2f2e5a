"""GPU parity of the brute-force mode (use_variance_reduction = 0) against the CPU oracle: same tables, same
fixed-address Philox streams, so the same photons reach the detector."""
import numpy as np
import pytest

import xmimsim_b200 as x
from helpers import Pair
from inputs import close_detector, example
from test_brute_cpu import _brute

pytestmark = pytest.mark.gpu


def _check(inp, opt, slack=2):
    """Every detected photon deposits the same weight w: spectra must agree to `slack` photons per bin (a photon
    grazing the detector rim may flip on a last-ulp difference between CUDA and glibc transcendentals)."""
    P = Pair(inp)
    ch, br, vr = P.sim.main_msim(opt, None)
    cnt = P.sim.brute_counters()
    ch_o, br_o, cnt_o = _brute(P, opt, n_threads=16)
    w = max((d.horizontal_intensity + d.vertical_intensity) / inp.n_photons_line for d in inp.discrete) * inp.live_time
    assert np.all(vr == 0)
    assert abs(cnt["hits"] - int(cnt_o[0])) <= slack and abs(cnt["offspring"] - int(cnt_o[2])) <= slack
    assert abs(cnt["interactions"] - int(cnt_o[1])) <= 4 * slack
    assert cnt["no_slot"] == 0
    assert np.abs(ch - ch_o).max() <= slack * w * 1.000001
    assert np.abs(br - br_o).max() <= slack * w * 1.000001
    assert np.abs(ch.sum() - ch_o.sum()) <= 4 * slack * w * (inp.n_interactions_trajectory + 1)
    P.close()
    return ch, br, cnt


def test_brute_matches_oracle_close_detector():
    inp = close_detector(n_photons=2000000, n_int=3)
    ch, br, cnt = _check(inp, x.main_options(use_variance_reduction=0))
    assert cnt["hits"] > 2000 and cnt["offspring"] == 0                          # Cr/Fe/Ni L lines lie below the 1 keV cut
    assert br[25, 2, 0] > 0 and br[25, 383, 0] > 0 and br[25, 384, 0] > 0        # Fe-KL3, Rayleigh, Compton at order 1
    assert np.all(np.diff(ch, axis=0) >= 0)


def test_brute_matches_oracle_cascade_modes_heavy_element():
    inp = close_detector(n_photons=1000000, n_int=2)
    inp.layers = [x.LayerD([50, 82], [0.6, 0.4], 8.0, 0.02)]
    inp.discrete = [x.DiscreteD(40.0, 1e9, 1e9), x.DiscreteD(95.0, 5e8, 5e8)]
    for kw in (dict(use_cascade_auger=0, use_cascade_radiative=0), dict(use_cascade_auger=1, use_cascade_radiative=0),
               dict(use_cascade_auger=0, use_cascade_radiative=1), dict(use_M_lines=0)):
        ch, br, cnt = _check(inp, x.main_options(use_variance_reduction=0, **kw))
        assert (cnt["offspring"] > 100) == bool(kw.get("use_cascade_auger", 1) or kw.get("use_cascade_radiative", 1))


def test_brute_matches_oracle_collimator_and_layers():
    inp = example("srm1155")                     # air + steel, conical collimator
    inp.n_photons_line = 40000
    inp.p_detector_window = [0.0, -1.2, 100.0]   # bring the collimated detector close enough to be hit
    inp.area_detector = 2.0
    inp.collimator_height = 0.4
    inp.collimator_diameter = 1.0
    ch, br, cnt = _check(inp, x.main_options(use_variance_reduction=0))
    assert cnt["hits"] > 50


def test_brute_zero_interaction_row():
    """A detector in the beam: photons that cross the sample without interacting land in channels row 0
    (src/xmi_main.F90:470-485 with n_interactions = 0)."""
    inp = close_detector(n_photons=200000, n_int=2)
    inp.layers = [x.LayerD([6], [1.0], 1.0, 0.01)]
    inp.n_sample_orientation = [0, 0, 1]
    inp.p_detector_window = [0, 0, 102]
    inp.n_detector_orientation = [0, 0, -1]
    ch, br, cnt = _check(inp, x.main_options(use_variance_reduction=0))
    total = (inp.discrete[0].horizontal_intensity + inp.discrete[0].vertical_intensity) * inp.live_time
    assert 0.5 * total < ch[0].sum() <= total
    e0_ch = int((20.0 - inp.zero) / inp.gain)
    assert ch[0, e0_ch] == ch[0].sum()


def test_brute_shards_sum_bit_exactly():
    inp = close_detector(n_photons=300000, n_int=2)
    P = Pair(inp)
    opt = x.main_options(use_variance_reduction=0)
    full, _ = P.sim.main_msim_raw(opt, None)
    parts = [P.sim.main_msim_raw(opt, None, rank=r, n_ranks=3)[0] for r in range(3)]
    assert np.array_equal(full, parts[0] + parts[1] + parts[2])
    a = P.sim.main_msim_finish(full, opt)
    b = P.sim.main_msim(opt, None)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    P.close()
