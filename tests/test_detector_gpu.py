"""GPU parity of the detector response against the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
import xmimsim_b200 as x
from inputs import example

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp64 on both sides; the GPU gathers per target channel with a fixed tree where the reference scatters
# sequentially, so sums differ in rounding only.
RTOL = 1e-11


def synthetic_escape_ratios(sim):
    """Plausible Si escape-ratio tables in the reference's layout (src/xmi_detector.c:174-285): 1990 input
    energies from 1 keV in 0.1 keV steps, 1999 Compton output energies from 0.1 keV."""
    n_in, n_out = 1990, 1999
    e_in = 1.0 + 0.1 * np.arange(n_in)
    e_out = 0.1 + 0.1 * np.arange(n_out)
    fl = np.zeros((n_in, 109, 1))
    above = e_in > 1.84
    fl[above, 2, 0] = 0.012 * np.exp(-(e_in[above] - 1.84) / 6.0)     # KL3
    fl[above, 1, 0] = 0.006 * np.exp(-(e_in[above] - 1.84) / 6.0)     # KL2
    fl[above, 5, 0] = 0.0007 * np.exp(-(e_in[above] - 1.84) / 6.0)    # KM3
    co = np.zeros((n_out, n_in))
    for j in range(n_in):
        emax = e_in[j] * (1 - 1 / (1 + 2 * e_in[j] / 511.0))
        sel = e_out < emax
        if sel.any():
            co[sel, j] = 2e-5 * (1 + e_out[sel] / max(emax, 1e-9)) / max(sel.sum(), 1)
    return sim.make_escape_ratios([14], fl, e_in, co, e_in, e_out)


def _golden_rows(name="srm1155"):
    g = np.load(os.path.join(GOLDEN, name + "_xmso.npz"))
    n_int = g["unconv"].shape[0]
    ch = np.zeros((n_int + 1, g["unconv"].shape[1]))
    ch[1:] = g["unconv"]
    return ch


@pytest.mark.parametrize("escape", [False, True])
def test_convolute_all_matches_oracle(escape):
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    ch = _golden_rows()
    er = synthetic_escape_ratios(sim) if escape else None
    o = x.main_options(use_escape_peaks=1 if escape else 0)
    ch_gpu = ch.copy()
    vr = np.zeros((100, 385, inp.n_interactions_trajectory))
    vr[25, 2, :] = [100.0, 10.0, 1.0, 0.1]
    vr_gpu = vr.copy()
    conv = sim.detector_convolute_all(ch_gpu, None, vr_gpu, o, er)
    assert sim.L.xmb_detector_last_launches() >= 3
    for k in range(1, inp.n_interactions_trajectory + 1):
        row, ref = orc.detector_convolute_spectrum(C.pointer(ci.input), ch[k].copy(), o, er, k)
        scale = ref.max()
        assert np.abs(conv[k] - ref).max() <= RTOL * scale, k
        assert np.abs(ch_gpu[k] - row).max() <= RTOL * row.max(), k      # in-place side effect reproduced
    assert np.all(conv[0] == 0) and np.array_equal(ch_gpu[0], ch[0])
    assert np.allclose(vr_gpu, orc.detector_convolute_history(C.pointer(ci.input), vr), rtol=1e-14)
    if escape:
        # escape moves counts 1.74 keV down from the Fe-Ka peak: visible in the efficiency-corrected row
        plain = ch.copy()
        sim.detector_convolute_all(plain, None, None, x.main_options(use_escape_peaks=0), None)
        peak = int(np.argmax(plain[4]))
        esc = peak - int(round(1.7397 / inp.gain))
        assert ch_gpu[4][esc - 2:esc + 3].sum() > 1.5 * plain[4][esc - 2:esc + 3].sum()
        assert ch_gpu[4][peak] < plain[4][peak]
    sim.close()


def test_reference_pile_up_criterion_and_statistics():
    """Reference test tests/test-pile-up.c:56 through the C ABI, plus agreement of the sum-peak content with
    the sequential oracle within counting statistics (independent random streams)."""
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    g = np.load(os.path.join(GOLDEN, "srm1155_xmso.npz"))
    channels = np.ascontiguousarray(g["unconv"][3]).copy()
    o = x.main_options(use_sum_peaks=0, use_escape_peaks=0, use_default_seeds=1)
    without = sim.detector_convolute_spectrum(channels, o, None, 4)
    o.use_sum_peaks = 1
    before = channels.copy()
    with_pu = sim.detector_convolute_spectrum(channels, o, None, 4)
    assert with_pu[1077] / without[1077] > 100.0
    row_o, conv_o = orc.detector_convolute_spectrum(C.pointer(ci.input), before.copy(), o, None, 4)
    # `channels` now holds integer pulse counts (:215); totals agree to sqrt(N), sum-peak window to 5 sigma
    n_gpu, n_orc = channels.sum(), row_o.sum()
    assert abs(n_gpu - n_orc) < 6 * np.sqrt(n_orc)
    w = slice(1060, 1100)
    assert abs(channels[w].sum() - row_o[w].sum()) < 5 * np.sqrt(row_o[w].sum()) + 5
    assert np.all(channels == np.round(channels))
    # deterministic: same call, same result
    again = before.copy()
    with2 = sim.detector_convolute_spectrum(again, o, None, 4)
    assert np.array_equal(with2, with_pu)
    sim.close()


def test_poisson_noise_statistics():
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    spec = np.zeros(inp.nchannels)
    spec[300:1500] = 400.0
    o = x.main_options(use_poisson=1, use_escape_peaks=0)
    clean = sim.detector_convolute_spectrum(spec.copy(), x.main_options(use_escape_peaks=0), None, 1)
    noisy = sim.detector_convolute_spectrum(spec.copy(), o, None, 1)
    sel = clean > 50
    z = (noisy[sel] - clean[sel]) / np.sqrt(clean[sel])
    assert abs(z.mean()) < 0.15 and 0.85 < z.std() < 1.15
    assert np.all(noisy[sel] == np.round(noisy[sel]))
    sim.close()
