"""BASELINE.json configurations at their full per-GPU sizes, checked through size-independent properties
(no oracle can follow at these sizes): bit-identical repetition, shard sums equal the whole bit for bit,
rows cumulative and non-negative, statistical agreement between disjoint shards and with a 100 x smaller run."""
import numpy as np
import pytest

import xmimsim_b200 as x
from inputs import example, synthetic_layers

pytestmark = pytest.mark.gpu


def _norm(limbs):
    """48-bit-split limbs -> python integers per slot (exact)."""
    l = limbs.reshape(-1, 2)
    return [int(a) + (int(b) << 48) for a, b in l[np.nonzero((l[:, 0] | l[:, 1]))[0][:4000]]]


def test_config1_srm1412_1e7_photons_per_line():
    """configs[1]: srm1412, 25 lines x 1e7 photons, 4 interactions, variance reduction on -- 2.5e8 histories."""
    inp = example("srm1412")
    inp.n_photons_line = 10_000_000
    sim = x.Simulation(inp, quality=0)
    g, r, t = sim.solid_angle_calculation(hits_per_single=5000, seed=1)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    opt = x.main_options()
    full, ex = sim.main_msim_raw(opt, sa)
    assert ex.n_histories == 250_000_000 and ex.n_interactions > 6e8
    again, _ = sim.main_msim_raw(opt, sa)
    assert np.array_equal(full, again)                                        # same bits on repetition
    h0, e0 = sim.main_msim_raw(opt, sa, rank=0, n_ranks=2)
    h1, e1 = sim.main_msim_raw(opt, sa, rank=1, n_ranks=2)
    assert e0.n_histories + e1.n_histories == 250_000_000
    # limbs are 48-bit split: add as integers, renormalise, compare with the whole
    def total(a, b):
        s = a.astype(object).reshape(-1, 2) + b.astype(object).reshape(-1, 2)
        return [int(p) + (int(q) << 48) for p, q in s]
    whole = [int(p) + (int(q) << 48) for p, q in full.astype(object).reshape(-1, 2)]
    assert total(h0, h1) == whole                                             # 1 GPU == 2 GPUs, bit for bit
    ch, br, vr = sim.main_msim_finish(full, opt)
    assert np.all(ch >= 0) and np.all(np.diff(ch, axis=0) >= 0) and np.all(br == 0)
    # the two half-jobs are independent estimates of the same spectrum (block-cyclic shards: same mix of lines)
    c0 = sim.main_msim_finish(h0, opt)[0]; c1 = sim.main_msim_finish(h1, opt)[0]
    assert abs(c0[-1].sum() / c1[-1].sum() - 1.0) < 1e-3
    strongest = np.argsort(vr.sum(axis=2).ravel())[-5:]
    v0 = sim.main_msim_finish(h0, opt)[2].sum(axis=2).ravel()[strongest]; v1 = sim.main_msim_finish(h1, opt)[2].sum(axis=2).ravel()[strongest]
    assert np.all(np.abs(v0 / v1 - 1.0) < 2e-3)
    # and a 100 x smaller run gives the same intensities within its own noise
    small = example("srm1412"); small.n_photons_line = 100_000
    sim2 = x.Simulation(small, quality=0)
    ch2, _, vr2 = sim2.main_msim(opt, sa)
    assert abs(ch2[-1].sum() / ch[-1].sum() - 1.0) < 5e-3
    assert np.all(np.abs(vr2.sum(axis=2).ravel()[strongest] / vr.sum(axis=2).ravel()[strongest] - 1.0) < 2e-2)
    sim.close(); sim2.close()


def test_config3_synthetic_ten_layers_one_of_eight_shards():
    """configs[3]: 10 layers, 8 interactions, 1e9 histories over 8 GPUs -- the share of one GPU (1.25e8 histories)."""
    inp = synthetic_layers(n_photons=1_000_000_000, n_int=8)
    sim = x.Simulation(inp, quality=0)
    g, r, t = sim.solid_angle_calculation(hits_per_single=2000, seed=1)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    opt = x.main_options()
    a, ea = sim.main_msim_raw(opt, sa, rank=3, n_ranks=8)
    assert ea.n_histories in (125_000_000 - 512, 125_000_000, 125_000_000 + 512) or abs(ea.n_histories - 125_000_000) <= 1024
    a2, _ = sim.main_msim_raw(opt, sa, rank=3, n_ranks=8)
    assert np.array_equal(a, a2)
    b, eb = sim.main_msim_raw(opt, sa, rank=4, n_ranks=8)
    ca, _, va = sim.main_msim_finish(a, opt); cb, _, vb = sim.main_msim_finish(b, opt)
    assert ca.shape == (9, inp.nchannels) and np.all(np.diff(ca, axis=0) >= 0)
    assert abs(ca[-1].sum() / cb[-1].sum() - 1.0) < 2e-3                       # two shards, same spectrum
    assert ca[8].sum() > ca[4].sum() > ca[1].sum() > 0                         # higher orders keep contributing
    top = np.argsort(va.sum(axis=2).ravel())[-5:]
    assert np.all(np.abs(va.sum(axis=2).ravel()[top] / vb.sum(axis=2).ravel()[top] - 1.0) < 5e-3)
    sim.close()


def test_config4_ebel_tube_with_escape_peaks_and_pile_up():
    """configs[4]: Ebel tube spectrum (Ag anode, 40 kV, 1000 intervals) on srm1155, escape peaks + pile-up in the detector
    response; 1e10 histories over 8 GPUs -- here 1/64 of the job (1.56e8 histories) from the same id space."""
    cont, disc = x.tube_ebel(x.LayerD([47], [1.0], 10.5, 0.0002), 40.0, 1.0, 60.0, 60.0, 0.039, 1e-4)
    inp = example("srm1155")
    inp.continuous = [x.ContinuousD(e, h, v) for e, h, v in cont]
    inp.discrete = [x.DiscreteD(e, h, v) for e, h, v in disc]
    inp.n_photons_interval = 9_950_000                                          # 1000 intervals + lines ~ 1e10 histories
    inp.n_photons_line = 9_950_000
    sim = x.Simulation(inp, quality=0)
    total = sim.L.xmb_msim_total_histories(sim.inputF)
    assert 0.99e10 < total < 1.02e10
    g, r, t = sim.solid_angle_calculation(hits_per_single=2000, seed=1)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    opt = x.main_options(use_sum_peaks=1, use_escape_peaks=1)
    a, ea = sim.main_msim_raw(opt, sa, rank=5, n_ranks=64)
    b, eb = sim.main_msim_raw(opt, sa, rank=6, n_ranks=64)
    assert abs(ea.n_histories - total / 64) <= 1024
    ca, bra, va = sim.main_msim_finish(a, opt); cb, _, vb = sim.main_msim_finish(b, opt)
    assert abs(ca[-1].sum() / cb[-1].sum() - 1.0) < 5e-3
    fe = va[25, 2].sum(), vb[25, 2].sum()
    assert fe[0] > 0 and abs(fe[0] / fe[1] - 1.0) < 1e-2
    # detector response of the shard (scaled to the whole job): escape peaks move counts down, pile-up up; totals obey
    # conv = efficiency-corrected input, within the Poisson noise pile-up introduces
    ch = ca * 64.0
    er = sim.escape_ratios_calculation(sim.escape_ratios_options(n_photons=50000))
    raw = ch.copy()
    conv = sim.detector_convolute_all(ch, bra, va, opt, er.contents)
    assert np.all(conv >= 0) and conv[-1].sum() > 0
    plain = sim.detector_convolute_all(raw.copy(), None, None, x.main_options(use_sum_peaks=0, use_escape_peaks=0), None)
    # pile-up merges pulses (fewer counts in total); the Si escape peak of Fe-K alpha (6.40 - 1.74 keV) gains counts
    assert conv[-1].sum() < plain[-1].sum() * (1 + 1e-9)
    lo, hi = int((4.55 - inp.zero) / inp.gain), int((4.78 - inp.zero) / inp.gain)
    assert conv[-1][lo:hi].sum() > plain[-1][lo:hi].sum()
    sim.escape_ratios_free(er)
    sim.close()
