"""world_size-2 gloo tests (CPU) of the multi-rank path's host logic: photon-id sharding, exact limb summation
with an integer all-reduce, the epilogue; and of the sharding semantics themselves through the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import orc
    import xmimsim_b200 as x
    from helpers import Pair
    from inputs import example
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inp = example("srm1155")
        inp.n_photons_line = 301                      # 26 lines x 301: not divisible by the world size
        P = Pair(inp)
        opt = x.main_options()
        # --- 1. block-cyclic shards partition the id range exactly -----------------------------------------------
        mine_n = P.sim.shard_count(rank, world)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine_n], dtype=torch.int64))
        assert sum(int(c[0]) for c in counts) == P.n_total
        owners = np.array([P.sim.shard_owner(g, world) for g in range(P.n_total)])
        assert int((owners == rank).sum()) == mine_n
        assert max(int(c[0]) for c in counts) - min(int(c[0]) for c in counts) <= 1024
        # --- 2. histories are independent of the sharding: oracle on my shard, summed over ranks == full run ----
        r_full, t_full = P.sim.solid_angle_inputs()
        r = np.linspace(r_full[0], r_full[-1], 48); t = np.linspace(t_full[0], t_full[-1], 48)
        sa_g, _ = orc.solid_angle_grid(P.od, r, np.arange(48), t, np.arange(48), 48, 300, 1, n_threads=2)
        sa = P.sim.make_solid_angle(sa_g, r, t)
        import ctypes as C
        from helpers import DEFAULT_SEED
        ch, vr, _, n_run = orc.main_msim_shard(C.pointer(P.ci.input), P.od, P.sim.L.xmb_get_tables(P.sim.hdf5F), opt, sa,
                                               DEFAULT_SEED, rank, world, inp.n_interactions_trajectory, inp.nchannels, 2)
        assert n_run == mine_n
        ch *= inp.live_time; vr *= inp.live_time
        tch, tvr = torch.from_numpy(ch.copy()), torch.from_numpy(vr.copy())
        dist.all_reduce(tch); dist.all_reduce(tvr)
        if rank == 0:
            ch_all, vr_all, _ = P.oracle(opt, sa, 0, n_threads=4)
            assert np.allclose(tch.numpy(), ch_all, rtol=1e-12, atol=1e-12 * ch_all.max())
            assert np.allclose(tvr.numpy(), vr_all, rtol=1e-12, atol=1e-12 * vr_all.max())
        # --- 3. limb all-reduce + epilogue: integer sums are exact and order independent ------------------------
        hz, hl = P.sim.slot_map(opt)
        n_int, nch = inp.n_interactions_trajectory, inp.nchannels
        row = nch + hz.size
        rng = np.random.default_rng(100 + rank)
        limbs = np.zeros((n_int * row, 2), np.uint64)
        limbs[:, 0] = rng.integers(0, 2 ** 48, n_int * row, dtype=np.uint64)       # low 48-bit words
        limbs[:, 1] = rng.integers(0, 2 ** 20, n_int * row, dtype=np.uint64)       # high words
        tl = torch.from_numpy(limbs.reshape(-1).view(np.int64).copy())
        dist.all_reduce(tl)
        mine = P.sim.main_msim_finish(tl.numpy().view(np.uint64), opt)
        # reference: regenerate every rank's limbs locally, add as python integers
        tot = np.zeros((n_int * row, 2), dtype=object)
        for rr in range(world):
            g = np.random.default_rng(100 + rr)
            tot[:, 0] += g.integers(0, 2 ** 48, n_int * row, dtype=np.uint64).astype(object)
            tot[:, 1] += g.integers(0, 2 ** 20, n_int * row, dtype=np.uint64).astype(object)
        exact = np.array([int(a) + (int(bb) << 48) for a, bb in tot], dtype=object).reshape(n_int, row)
        # channel row k+1 = cumulative sum over orders of (channel slots + line slots binned at their channel)
        Wmax = max(d.horizontal_intensity + d.vertical_intensity for d in inp.discrete) / inp.n_photons_line
        scale = Wmax * inp.live_time / 2.0 ** 56
        vr_out = mine[2]
        for k in range(n_int):
            for s in (0, 1, hz.size - 1):
                assert vr_out[hz[s] - 1, hl[s] - 1, k] == float(exact[k, nch + s]) * scale or \
                    abs(vr_out[hz[s] - 1, hl[s] - 1, k] / (float(exact[k, nch + s]) * scale) - 1) < 1e-15
        assert np.all(np.diff(mine[0], axis=0) >= 0) and np.all(mine[0][0] == 0)
        got = [None] * world
        dist.all_gather_object(got, float(mine[0].sum()))
        assert len(set(got)) == 1                    # every rank finishes to identical bits
        P.close()
        q.put((rank, "ok"))
    except Exception as exc:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_exact_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_helper_edge_cases():
    from xmimsim_b200 import abi
    L = abi.lib()
    for n, w in ((0, 3), (5, 8), (1000, 7), (1024 * 9 + 17, 4), (2 ** 40 + 5, 8)):
        counts = [L.xmb_msim_shard_count(n, r, w) for r in range(w)]
        assert sum(counts) == n
        assert max(counts) - min(counts) <= 1024
        for g in (0, 1023, 1024, n - 1):
            if 0 <= g < n:
                assert L.xmb_msim_shard_owner(g, w) == (g // 1024) % w
    assert L.xmb_msim_shard_count(100, 5, 3) == 0          # rank outside 0..n-1
