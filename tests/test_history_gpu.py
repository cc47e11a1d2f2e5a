"""GPU parity of the photon-history engine against the CPU oracle on identical inputs, tables and
Philox streams (tier T0 of SURVEY.md 8c), plus size-independent properties at larger sizes."""
import copy
import os

import numpy as np
import pytest

import xmimsim_b200 as x
from helpers import Pair, assert_spectra_close, diagnose_mismatch
from inputs import example, caso4, synthetic_layers, ebel_like

pytestmark = pytest.mark.gpu

# Tolerance: both sides are fp64 and consume the same tables and random numbers.  Differences come from
# (a) summation order (oracle: per-thread double sums; GPU: exact integers), (b) fused multiply-adds and
# CUDA-vs-glibc transcendentals (a few ulp per history), (c) the rare history whose discrete decision
# (layer / element / channel) flips on such an ulp.  (c) bounds the test: one flipped history changes a
# channel by ~1/N of its content, so spectra must agree to 1e-6 of their maximum at N ~ 5e4 histories.
RTOL = 2e-6


def run_both(inp, options=None, seed=11, grid_n=None, hits=400, provider=None):
    """Engine and oracle once each on the same input, tables, grid and Philox key.  No repetition: a comparison that
    differs is a failure; helpers.diagnose_mismatch records what attributes it (both sides' per-order sums, digests,
    a second engine run, a single-thread oracle run, a fresh Pair) in gpurun_out/ and in the assertion message."""
    options = options or x.main_options()
    P = Pair(inp, provider=provider)
    sa = P.grid(hits_per_single=hits, n=grid_n)
    ch, br, vr = P.sim.main_msim(options, sa)
    ch_o, vr_o, cnt = P.oracle(options, sa, 0)   # seed 0 -> default key on both sides
    if _differs(ch, ch_o) or _differs(vr, vr_o):
        note = diagnose_mismatch(P, options, sa, ch, vr, ch_o, vr_o, grid_n=grid_n, hits=hits)
        P.close()
        raise AssertionError("engine and oracle differ: " + note)
    P.close()
    return ch, br, vr, ch_o, vr_o, cnt


def _differs(a, b):
    scale = np.abs(b).max()
    return bool(scale > 0 and np.abs(a - b).max() > RTOL * scale)


@pytest.mark.parametrize("name,n_line", [("srm1155", 1500), ("srm1412", 1500), ("srm1132", 1500), ("In", 1500)])
def test_examples_match_oracle(name, n_line):
    inp = example(name)
    inp.n_photons_line = n_line
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp)
    assert ch.shape == (inp.n_interactions_trajectory + 1, inp.nchannels)
    assert np.all(ch[0] == 0) and np.all(br == 0)            # no zero-interaction / brute-force content with VR
    assert ch_o[1:].sum() > 0
    assert_spectra_close(ch, ch_o, RTOL, name + " channels")
    assert_spectra_close(vr, vr_o, RTOL, name + " var_red_history")
    # strongest lines agree to 1e-6 relative individually
    idx = np.argsort(vr_o.sum(axis=2).ravel())[-8:]
    a = vr.sum(axis=2).ravel()[idx]; b = vr_o.sum(axis=2).ravel()[idx]
    assert np.all(np.abs(a - b) <= 1e-6 * b)
    # rows are cumulative over interaction order
    assert np.all(np.diff(ch, axis=0) >= 0)


@pytest.mark.parametrize("name", ["srm1155", "srm1412"])
def test_dense_line_set_matches_oracle(name):
    """The stand-in provider with xraylib's line density (~320 active forced-detection lines for srm1155 instead of ~150):
    more and longer line tiles per layer, more groups per tile."""
    from xmimsim_b200 import abi
    inp = example(name)
    inp.n_photons_line = 1200
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128, provider=abi.lib().xmb_xrl_surrogate_dense())
    assert_spectra_close(ch, ch_o, RTOL, name + " dense channels")
    assert_spectra_close(vr, vr_o, RTOL, name + " dense history")
    assert np.count_nonzero(vr.sum(axis=2)) > 250


def test_caso4_single_interaction():
    inp = caso4()
    inp.n_photons_line = 20000
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert_spectra_close(ch, ch_o, RTOL, "caso4 channels")
    assert_spectra_close(vr, vr_o, RTOL, "caso4 history")
    assert vr[19, 2, 0] > 0                                   # Ca-KL3, first order
    assert cnt[1] <= inp.n_photons_line


@pytest.mark.parametrize("opts", [dict(use_M_lines=0), dict(use_cascade_auger=0), dict(use_cascade_radiative=0),
                                   dict(use_cascade_auger=0, use_cascade_radiative=0)])
def test_option_variants(opts):
    inp = example("srm1412")
    inp.n_photons_line = 800
    o = x.main_options(**opts)
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, options=o, grid_n=128)
    assert_spectra_close(ch, ch_o, RTOL, "channels %r" % opts)
    assert_spectra_close(vr, vr_o, RTOL, "history %r" % opts)


def test_synthetic_ten_layers_eight_interactions():
    inp = synthetic_layers(n_photons=30000, n_int=8)
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert ch.shape == (9, inp.nchannels)
    assert_spectra_close(ch, ch_o, RTOL, "synthetic channels")
    assert_spectra_close(vr, vr_o, RTOL, "synthetic history")


@pytest.mark.parametrize("n_layers", [1, 2, 3, 4, 5, 7])
def test_every_layer_count_instantiation_matches_oracle(n_layers):
    """The kernel is compiled for 1, 2, 3 and 4 layers (loops over layers unrolled) and in a generic form: the first
    n layers of the 10-layer sample exercise each instantiation, with three or more layers also the per-layer queues."""
    inp = synthetic_layers(n_photons=12000, n_int=5)
    inp.layers = inp.layers[:n_layers]
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert ch[-1].sum() > 0
    assert_spectra_close(ch, ch_o, RTOL, "%d layers channels" % n_layers)
    assert_spectra_close(vr, vr_o, RTOL, "%d layers history" % n_layers)


@pytest.mark.parametrize("n_photons,n_int", [(1, 1), (1, 4), (31, 3), (1025, 2), (3, 12)])
def test_tiny_and_ragged_sizes(n_photons, n_int):
    """One photon, fewer photons than a warp, one more than a CTA, more interactions than photons: partial batches only."""
    inp = example("srm1155")
    inp.n_photons_line = n_photons
    inp.n_interactions_trajectory = n_int
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=64)
    assert ch.shape == (n_int + 1, inp.nchannels)
    assert_spectra_close(ch, ch_o, RTOL, "tiny channels")
    assert_spectra_close(vr, vr_o, RTOL, "tiny history")


def test_small_detector_and_many_elements_per_layer():
    """Few channels (most deposits fall outside and are dropped, src/xmi_variance_reduction.F90:370-389) and a layer
    with 16 elements (the widest element loop of the pool)."""
    inp = synthetic_layers(n_photons=8000, n_int=3)
    pool = [8, 13, 14, 20, 22, 26, 29, 30, 38, 42, 47, 50, 56, 74, 79, 82]
    inp.layers = [x.LayerD(pool, [1.0 / len(pool)] * len(pool), 4.0, 0.02), inp.layers[1]]
    inp.nchannels = 256
    inp.gain = 0.05
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=96)
    assert_spectra_close(ch, ch_o, RTOL, "few channels")
    assert_spectra_close(vr, vr_o, RTOL, "few channels history")


def test_continuous_and_broadened_sources():
    inp = ebel_like(n_intervals=40, n_photons_interval=600, n_photons_line=1500)
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert_spectra_close(ch, ch_o, RTOL, "continuous channels")
    assert_spectra_close(vr, vr_o, RTOL, "continuous history")


def test_ebel_tube_spectrum_matches_oracle():
    """BASELINE config 5's excitation as the survey specifies it: the product's xmb_tube_ebel spectrum of an Ag anode at 40 kV
    (1000 continuous intervals + 20 discrete lines, xmimsim_b200/workloads.py::ebel_tube), a few photons per segment."""
    from xmimsim_b200 import workloads
    inp = workloads.ebel_tube(n_photons_interval=40)
    assert workloads.n_source_segments(inp) == 1020
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert_spectra_close(ch, ch_o, RTOL, "ebel tube channels")
    assert_spectra_close(vr, vr_o, RTOL, "ebel tube history")


def test_gaussian_source_and_excitation_absorber():
    inp = example("srm1412")       # has an Al excitation-path absorber
    inp.n_photons_line = 600
    for d in inp.discrete:
        d.sigma_x, d.sigma_y, d.sigma_xp, d.sigma_yp = 0.01, 0.02, 1e-4, 2e-4
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=128)
    assert_spectra_close(ch, ch_o, RTOL, "gaussian-source channels")


def test_interaction_points_beyond_the_solid_angle_grid():
    """xmi_get_solid_angle falls back to hits_per_single rays on the spot when a point lies beyond the grid
    (src/xmi_solid_angle_f.F90:783-789): a grid cut just behind the sample surface sends a large share of the
    interactions down that path; the engine (whole CTA shares the rays of a point) and the oracle agree."""
    import orc
    inp = example("srm1155")
    inp.n_photons_line = 300
    P = Pair(inp)
    r_full, t_full = P.sim.solid_angle_inputs()
    pw = np.array(inp.p_detector_window, float)
    r_cut = np.linalg.norm(pw - np.array([0.0, 0.0, inp.d_sample_source])) * (1.0 + 1e-6)
    assert r_full[0] < r_cut < r_full[-1]
    r = np.linspace(r_full[0], r_cut, 128)
    t = np.linspace(t_full[0], t_full[-1], 128)
    g, _ = P.sim.solid_angle_grid(r, t, hits_per_single=400, seed=3)
    sa = P.sim.make_solid_angle(g, r, t)
    opt = x.main_options()
    P.sim.L.xmb_set_hits_per_single(601)
    orc.lib().orc_set_hits_per_single(601)
    try:
        assert P.sim.L.xmb_get_hits_per_single() == 601
        ch, br, vr = P.sim.main_msim(opt, sa)
        ch_o, vr_o, cnt = P.oracle(opt, sa, 0)
    finally:
        P.sim.L.xmb_set_hits_per_single(0)
        orc.lib().orc_set_hits_per_single(5000)
    assert P.sim.L.xmb_get_hits_per_single() == 5000
    n_inter = int(cnt[1])
    assert 0.05 * n_inter < int(cnt[0]) < 0.98 * n_inter, (cnt[0], n_inter)       # both paths are exercised
    assert_spectra_close(ch, ch_o, RTOL, "off-grid channels")
    assert_spectra_close(vr, vr_o, RTOL, "off-grid history")
    # the fallback carries real intensity: with the off-grid points scored as zero the spectrum would be far smaller
    full = P.grid(hits_per_single=400, n=128)
    ch_full, _, _ = P.sim.main_msim(opt, full)
    assert 0.8 < ch[-1].sum() / ch_full[-1].sum() < 1.25
    P.close()


def test_bit_exact_across_rank_counts():
    """Photon-id shards summed in uint64 must reproduce the single-rank accumulators bit for bit, for any
    shard count, and so must a repeat run (atomics order does not matter)."""
    inp = example("srm1155")
    inp.n_photons_line = 1111            # not a multiple of 32 nor of the rank counts
    P = Pair(inp)
    sa = P.grid(n=128)
    o = x.main_options()
    ref, ex = P.sim.main_msim_raw(o, sa)
    assert ex.n_histories == P.n_total and ex.n_launches == 2      # history kernel + limb conversion
    again, _ = P.sim.main_msim_raw(o, sa)
    assert np.array_equal(ref, again)
    for n_ranks in (2, 3, 8):
        tot = np.zeros_like(ref)
        nh = 0
        for r in range(n_ranks):
            limbs, ex = P.sim.main_msim_raw(o, sa, rank=r, n_ranks=n_ranks)
            tot += limbs
            nh += ex.n_histories
        assert nh == P.n_total
        # limbs are 48-bit split words: normalise carries before comparing
        def norm(v):
            v = v.reshape(-1, 2).astype(object)
            return [int(a) + (int(b) << 48) for a, b in v]
        assert norm(tot) == norm(ref), n_ranks
    ch1 = P.sim.main_msim_finish(ref, o)[0]
    ch2 = P.sim.main_msim(o, sa)[0]
    assert np.array_equal(ch1, ch2)
    other, _ = P.sim.main_msim_raw(o, sa, seed=12345)
    assert not np.array_equal(ref, other)
    P.close()


def test_batch_formation_modes_are_bit_identical(monkeypatch):
    """The ways a CTA forms its batches on a many-layer sample -- unsorted, mixed batches sorted by layer or by energy class, one
    queue per (interaction order, layer) -- regroup the same photons: the integer sums must not differ in a single bit
    (XMB_LAYER_SORT is the engine's experiment switch; default = per-layer queues from three layers on)."""
    inp = synthetic_layers(n_photons=150_000, n_int=8)
    P = Pair(inp)
    sa = P.grid(n=128)
    o = x.main_options()
    ref, ex = P.sim.main_msim_raw(o, sa)
    assert ex.n_histories == 150_000
    for mode in ("0", "1", "2", "3"):
        monkeypatch.setenv("XMB_LAYER_SORT", mode)
        limbs, ex2 = P.sim.main_msim_raw(o, sa)
        assert ex2.n_interactions == ex.n_interactions
        assert np.array_equal(limbs, ref), mode
        shard, _ = P.sim.main_msim_raw(o, sa, rank=1, n_ranks=3)
        monkeypatch.delenv("XMB_LAYER_SORT")
        shard_ref, _ = P.sim.main_msim_raw(o, sa, rank=1, n_ranks=3)
        assert np.array_equal(shard, shard_ref), mode
    P.close()


@pytest.mark.parametrize("which", ["synthetic10", "srm1412", "ebel"])
def test_sums_do_not_depend_on_the_launch_shape(which, monkeypatch):
    """Which photons share a batch depends on the number of CTAs, on the CTA size and on the order in which warps reach
    the queues; every deposit is an integer and every draw has a fixed address, so the accumulators must not move by a
    bit.  XMB_HIST_BLOCKS / XMB_HIST_THREADS are the engine's experiment switches (they only reduce the launch)."""
    if which == "synthetic10":
        inp = synthetic_layers(n_photons=60_000, n_int=8)
    elif which == "srm1412":
        inp = example("srm1412"); inp.n_photons_line = 2500
    else:
        inp = ebel_like(n_intervals=60, n_photons_interval=700, n_photons_line=3000)
    P = Pair(inp)
    sa = P.grid(n=128)
    o = x.main_options()
    ref, ex = P.sim.main_msim_raw(o, sa)
    shapes = [(b, t) for b in (1, 5, 37, 148) for t in (1024, 736, 256, 64)]
    for b, t in shapes:
        monkeypatch.setenv("XMB_HIST_BLOCKS", str(b)); monkeypatch.setenv("XMB_HIST_THREADS", str(t))
        for rep in range(2):
            limbs, ex2 = P.sim.main_msim_raw(o, sa)
            assert ex2.n_interactions == ex.n_interactions, (b, t)
            assert np.array_equal(limbs, ref), (b, t, rep)
    monkeypatch.delenv("XMB_HIST_BLOCKS"); monkeypatch.delenv("XMB_HIST_THREADS")
    for rep in range(10):
        limbs, _ = P.sim.main_msim_raw(o, sa)
        assert np.array_equal(limbs, ref), rep
    P.close()


def test_linearity_and_weight_bounds_at_larger_size():
    """Size-independent properties: the spectrum per unit photon converges (two sizes agree within the
    statistical error), order-1 content dominates, all deposits non-negative."""
    inp = example("srm1155")
    res = []
    for n in (20000, 80000):
        d = copy.deepcopy(inp)
        d.n_photons_line = n
        P = Pair(d)
        sa = P.grid(n=256, hits_per_single=1000, seed=5)
        ch, br, vr = P.sim.main_msim(x.main_options(), sa)
        P.close()
        res.append((ch, vr))
    (c1, v1), (c2, v2) = res
    assert np.all(c1 >= 0) and np.all(c2 >= 0)
    t1, t2 = c1[-1].sum(), c2[-1].sum()
    assert abs(t1 / t2 - 1.0) < 5e-3
    fe1, fe2 = v1[25, 2].sum(), v2[25, 2].sum()
    assert abs(fe1 / fe2 - 1.0) < 5e-3
    assert c2[1].sum() > 0.8 * c2[-1].sum()


def test_unsupported_modes_fail_loudly():
    inp = caso4()
    inp.n_photons_line = 100
    P = Pair(inp)
    sa = P.grid(n=64)
    for kw in (dict(escape_ratios_mode=1),):
        with pytest.raises(RuntimeError, match="not implemented"):
            P.sim.main_msim(x.main_options(**kw), sa)
    P.close()


def test_fixed_point_sums_reproduce_the_recorded_digests():
    """Regression pin of the integer accumulators: seven inputs (two- and ten-layer samples, continuous source, all
    cascade modes off, shell-resolved Compton, a 1-of-3 shard) must reproduce, bit for bit, the digests recorded with
    kernel v18 (tests/golden/limbs_digest.json, written by tools/limbs_digest.py; v16 evaluates the line attenuation factors in single precision, v17 multiplies the per-photon factors of the scatter deposits first and v18 rotates directions from their components instead of stored polar angles, each of which moved the sums recorded before by rounding) -- kernel rewrites that regroup
    photons, restage deposits or change launch shapes must not move a single bit."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "tools"))
    import limbs_digest
    want = json.load(open(os.path.join(here, "golden", "limbs_digest.json")))
    got = limbs_digest.collect()
    for name, w in want.items():
        g = got[name]
        assert (g["n"], g["inter"]) == (w["n"], w["inter"]), name
        assert (g["sha"], g["sha_shard"]) == (w["sha"], w["sha_shard"]), name


def test_resident_grid_follows_the_contents_of_the_host_buffer():
    """With keep_on_device the solid-angle grid stays in HBM between calls.  Residency is keyed on the CONTENTS of the caller's
    buffers (a 64-bit hash), not on their address: a caller that refills the same buffer, or a new grid that lands on the address
    of a freed one, must get the new grid (the round-1 library compared host pointer and size only).  Also a grid with the two
    axis lengths swapped on the same handle (per-buffer capacities)."""
    import hashlib
    import torch

    def limbs(sim):
        ptr, n = sim.device_limbs()
        holder = type("H", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}})()
        return hashlib.sha256(torch.as_tensor(holder, device="cuda").cpu().numpy().tobytes()).hexdigest()

    inp = example("srm1155")
    inp.n_photons_line = 300
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r, t = r_full[::16].copy(), t_full[::16].copy()
    rng = np.random.default_rng(3)
    g_a = rng.uniform(1e-4, 2e-4, (t.size, r.size))
    g_b = g_a * 1.5
    opt = x.main_options()
    buf = g_a.copy()
    sa = sim.make_solid_angle(buf, r, t)          # borrows buf's memory
    host_buf = sa._keep[0]
    assert host_buf.ctypes.data == buf.ctypes.data
    sim.main_msim_device(opt, sa); d_a = limbs(sim)
    sim.main_msim_device(opt, sa); assert limbs(sim) == d_a          # unchanged buffer: resident grid, same sums
    host_buf[...] = g_b                                              # the SAME buffer refilled in place
    sim.main_msim_device(opt, sa); d_b = limbs(sim)
    assert d_b != d_a
    fresh = x.Simulation(inp, quality=0)                             # what a run that never saw grid A gives for grid B
    fresh.main_msim_device(opt, fresh.make_solid_angle(g_b.copy(), r, t)); assert limbs(fresh) == d_b
    fresh.close()
    host_buf[...] = g_a
    sim.main_msim_device(opt, sa); assert limbs(sim) == d_a
    # axis lengths swapped (n_r x n_theta, then n_theta x n_r with n_theta > the old n_r capacity): no overflow, right sums
    r2, t2 = r_full[::32].copy(), t_full[::8].copy()
    g2 = rng.uniform(1e-4, 2e-4, (t2.size, r2.size))
    sim.main_msim_device(opt, sim.make_solid_angle(g2.copy(), r2, t2)); d_2 = limbs(sim)
    fresh = x.Simulation(inp, quality=0)
    fresh.main_msim_device(opt, fresh.make_solid_angle(g2.copy(), r2, t2)); assert limbs(fresh) == d_2
    fresh.close()
    sim.close()


def test_very_many_channels_and_the_unstaged_channel_path(monkeypatch):
    """16 384 channels: channels + history slots exceed the shared-memory staging area, the Rayleigh / Compton channel deposits go
    straight to the global accumulators (history slots stay staged).  The path is also forced on an ordinary input
    (XMB_STAGE_CHANNELS=0): integer sums, so the raw accumulators must equal the staged run's bit for bit."""
    inp = example("srm1155")
    inp.n_photons_line = 600
    inp.nchannels = 16384
    inp.gain = 0.0119 / 8
    ch, br, vr, ch_o, vr_o, cnt = run_both(inp, grid_n=96)
    assert ch.shape == (inp.n_interactions_trajectory + 1, 16384) and ch[-1].sum() > 0
    assert_spectra_close(ch, ch_o, RTOL, "16384 channels")
    assert_spectra_close(vr, vr_o, RTOL, "16384 channels history")
    b = example("srm1412"); b.n_photons_line = 2000
    sim = x.Simulation(b, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r, t = r_full[::8].copy(), t_full[::8].copy()
    sa = sim.make_solid_angle(np.random.default_rng(5).uniform(1e-4, 2e-4, (t.size, r.size)), r, t)
    limbs_staged, _ = sim.main_msim_raw(x.main_options(), sa)
    monkeypatch.setenv("XMB_STAGE_CHANNELS", "0")
    limbs_direct, _ = sim.main_msim_raw(x.main_options(), sa)
    sim.close()
    assert np.array_equal(limbs_staged, limbs_direct)
