"""GPU parity of the shell-resolved Compton option against the oracle, in both scoring modes."""
import numpy as np
import pytest

import xmimsim_b200 as x
from helpers import Pair, assert_spectra_close
from inputs import close_detector, example
from test_brute_cpu import _brute

pytestmark = pytest.mark.gpu
RTOL = 2e-6      # as tests/test_history_gpu.py


@pytest.mark.parametrize("name", ["srm1155", "srm1132"])
def test_advanced_compton_forced_detection_matches_oracle(name):
    inp = example(name)
    inp.n_photons_line = 1200
    P = Pair(inp)
    sa = P.grid(hits_per_single=400)
    opt = x.main_options(use_advanced_compton=1)
    ch, br, vr = P.sim.main_msim(opt, sa)                 # builds the subshell tables on first use
    assert P.sim.tables.n_adv_rows > 0
    ch_o, vr_o, _ = P.oracle(opt, sa, 0)
    assert_spectra_close(ch, ch_o, RTOL, name + " channels")
    assert_spectra_close(vr, vr_o, RTOL, name + " history")
    # and it differs from the default Compton treatment in the scattered part only
    ch_d, _, vr_d = P.sim.main_msim(x.main_options(), sa)
    assert np.allclose(vr[:, :384, 0], vr_d[:, :384, 0], rtol=1e-12, atol=0)
    assert not np.allclose(ch[1], ch_d[1], rtol=1e-6, atol=0)
    P.close()


def test_advanced_compton_brute_force_matches_oracle():
    inp = close_detector(n_photons=600000, n_int=3)
    inp.layers = [x.LayerD([6, 8, 14], [0.5, 0.3, 0.2], 1.5, 0.5)]        # light matrix: Compton dominated
    inp.discrete = [x.DiscreteD(40.0, 1e9, 1e9)]
    P = Pair(inp)
    opt = x.main_options(use_variance_reduction=0, use_advanced_compton=1)
    ch, br, vr = P.sim.main_msim(opt, None)
    cnt = P.sim.brute_counters()
    ch_o, br_o, cnt_o = _brute(P, opt, n_threads=16)
    w = 2e9 / inp.n_photons_line
    assert cnt["hits"] > 3000 and abs(cnt["hits"] - int(cnt_o[0])) <= 2
    assert np.abs(ch - ch_o).max() <= 2 * w * 1.000001 and np.abs(br - br_o).max() <= 2 * w * 1.000001
    assert br[:, 384, :].sum() > 0.5 * br.sum()                            # mostly Compton-scattered photons
    P.close()
