"""Randomized parity: engine and oracle on random valid inputs (tests/random_inputs.py: 1 - 8 layers, gas gaps, conical /
cylindrical / no collimator, broadened lines, Gaussian sources, continuous blocks, absorbers, 1 - 6 interactions, random
cascade / M-line options).  Same tolerance as the authored cases of tests/test_history_gpu.py.  tools/random_parity_hunt.py
runs the same comparison over many more seeds (record: profiles/r2_random_parity_hunt.txt)."""
import numpy as np
import pytest

import xmimsim_b200 as x
from helpers import assert_spectra_close
from random_inputs import random_input
from test_history_gpu import run_both, RTOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(16)) + [1000, 1002, 1003, 1005, 1007, 1011])   # from 1000: tilted sample / detector
def test_random_input_matches_oracle(seed):
    inp, opts = random_input(seed, n_photons=1500)
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, options=x.main_options(**opts), grid_n=96, hits=300)
    assert ch.shape == (inp.n_interactions_trajectory + 1, inp.nchannels)
    assert np.isfinite(ch).all() and np.isfinite(vr).all()
    assert ch_o[-1].sum() > 0
    assert_spectra_close(ch, ch_o, RTOL, "random %d channels" % seed)
    assert_spectra_close(vr, vr_o, RTOL, "random %d history" % seed)
    assert np.all(np.diff(ch, axis=0) >= 0)


def _hunt():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import random_kernel_hunt
    return random_kernel_hunt


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 8, 1000, 1004, 1009])
def test_random_geometry_solid_angle_hits_match_oracle(seed):
    """Random detector distance / area / collimator: integer hit counts of a sub-grid of the real axes, point for point."""
    ok, note = _hunt().hunt_sa(seed)
    assert ok, note


@pytest.mark.parametrize("seed", [1, 2, 5, 7, 1002, 1006])
def test_random_input_brute_force_matches_oracle(seed):
    """Brute-force mode on random samples (detector brought close): the same photons are detected."""
    ok, note = _hunt().hunt_brute(seed)
    assert ok, note


@pytest.mark.parametrize("seed", [0, 4, 9])
def test_random_input_advanced_compton_matches_oracle(seed):
    ok, note = _hunt().hunt_adv(seed)
    assert ok, note


@pytest.mark.parametrize("seed", [0, 3, 6])
def test_random_crystal_escape_ratios_match_oracle(seed):
    ok, note = _hunt().hunt_esc(seed)
    assert ok, note


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_detector_response_matches_oracle(seed):
    ok, note = _hunt().hunt_det(seed)
    assert ok, note
