"""Shared helpers for the GPU parity tests: run the product and the oracle on the same input."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x


DEFAULT_SEED = 0x584D494D53494D      # "XMIMSIM" 


class Pair:
    def __init__(self, inp, quality=0, provider=None):
        self.inp = inp
        self.sim = x.Simulation(inp, quality=quality, provider=provider)
        self.ci = x.CInput(inp)
        self.od = orc.init_input(C.pointer(self.ci.input))
        self.n_total = orc.lib().orc_total_histories(C.cast(C.pointer(self.ci.input), C.c_void_p))

    def grid(self, hits_per_single=500, seed=3, n=None):
        """Solid-angle grid from the GPU kernel (full 1024^2 axes unless n is given)."""
        if n is None:
            g, r, t = self.sim.solid_angle_calculation(hits_per_single=hits_per_single, seed=seed)
            return self.sim.make_solid_angle(g.copy(), r.copy(), t.copy())
        r_full, t_full = self.sim.solid_angle_inputs()
        r = np.linspace(r_full[0], r_full[-1], n)
        t = np.linspace(t_full[0], t_full[-1], n)
        g, _ = self.sim.solid_angle_grid(r, t, hits_per_single=hits_per_single, seed=seed)
        return self.sim.make_solid_angle(g, r, t)

    def oracle(self, options, sa, seed, g0=0, g1=None, n_threads=16):
        g1 = self.n_total if g1 is None else g1
        seed = seed or DEFAULT_SEED          # the engine maps seed 0 to its default Philox key
        ch, vr, cnt = orc.main_msim_range(C.pointer(self.ci.input), self.od, self.sim.L.xmb_get_tables(self.sim.hdf5F),
                                          options, sa, seed, g0, g1, self.inp.n_interactions_trajectory,
                                          self.inp.nchannels, n_threads)
        return ch * self.inp.live_time, vr * self.inp.live_time, cnt

    def close(self):
        self.sim.close()


def assert_spectra_close(a, b, rtol, what):
    """Element-wise agreement relative to the array's scale: |a-b| <= rtol * max(|b|) per row."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.abs(b).max()
    if scale == 0:
        assert np.abs(a).max() == 0, what
        return 0.0
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max |diff| / max|ref| = %.3e > %.1e" % (what, err, rtol)
    return err


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def diagnose_mismatch(P, options, sa, ch, vr, ch_o, vr_o, grid_n=None, hits=400, tag="history"):
    """Called by a parity test whose first (and only) comparison differs: gathers what tells the two sides apart --
    per-order sums and digests of both, the engine once more on the same handle, the oracle on one thread, both sides
    on a fresh Pair with the same and with a fresh grid -- writes it to gpurun_out/parity_mismatch_<pid>.json (arrays
    next to it as .npz) and returns a one-line summary for the assertion message."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_dir = os.path.join(root, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    inp = P.inp
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    rows = lambda a: [float(v) for v in np.asarray(a).sum(axis=1)]
    n_r, n_t = sa.grid_dims_r_n, sa.grid_dims_theta_n
    grid = np.ctypeslib.as_array(sa.solid_angles, shape=(n_r * n_t,)).copy()
    d = {"tag": tag, "pid": os.getpid(), "n_total": int(P.n_total), "err_ch": rel(ch, ch_o), "err_vr": rel(vr, vr_o),
         "gpu_rows": rows(ch), "orc_rows": rows(ch_o), "gpu_sha": _sha(ch), "orc_sha": _sha(ch_o), "grid_sha": _sha(grid),
         "grid_sum": float(grid.sum()), "env_layer_sort": os.environ.get("XMB_LAYER_SORT")}
    try:
        ch2, br2, vr2 = P.sim.main_msim(options, sa)
        d["engine_repeat_identical"] = bool(np.array_equal(ch, ch2) and np.array_equal(vr, vr2))
        d["engine_repeat_err_vs_oracle"] = rel(ch2, ch_o)
        ch_o1, vr_o1, _ = P.oracle(options, sa, 0, n_threads=1)
        d["oracle_1thread_vs_oracle"] = rel(ch_o1, ch_o)
        d["oracle_1thread_rows"] = rows(ch_o1)
        d["engine_vs_oracle_1thread"] = rel(ch, ch_o1)
        for mode in ("0", "1"):
            os.environ["XMB_LAYER_SORT"] = mode
            try:
                chm, _, _ = P.sim.main_msim(options, sa)
            finally:
                del os.environ["XMB_LAYER_SORT"]
            d["engine_mode%s_identical" % mode] = bool(np.array_equal(chm, ch))
            d["engine_mode%s_err_vs_oracle" % mode] = rel(chm, ch_o)
        P2 = Pair(inp)
        ch3, _, vr3 = P2.sim.main_msim(options, sa)
        ch_o3, vr_o3, _ = P2.oracle(options, sa, 0)
        d["fresh_pair_same_grid"] = {"err": rel(ch3, ch_o3), "gpu_equals_first_gpu": bool(np.array_equal(ch3, ch)),
                                     "orc_vs_first_orc": rel(ch_o3, ch_o), "gpu_rows": rows(ch3), "orc_rows": rows(ch_o3)}
        # cross: first handle's oracle state with the new tables and vice versa tells tables from derived input
        sa4 = P2.grid(hits_per_single=hits, n=grid_n)
        grid4 = np.ctypeslib.as_array(sa4.solid_angles, shape=(sa4.grid_dims_r_n * sa4.grid_dims_theta_n,)).copy()
        ch4, _, _ = P2.sim.main_msim(options, sa4)
        ch_o4, _, _ = P2.oracle(options, sa4, 0)
        d["fresh_pair_fresh_grid"] = {"err": rel(ch4, ch_o4), "grid_equals_first": bool(np.array_equal(grid4, grid)),
                                      "gpu_equals_first_gpu": bool(np.array_equal(ch4, ch)), "orc_vs_first_orc": rel(ch_o4, ch_o)}
        P2.close()
    except Exception as e:                      # the diagnosis must not hide the failure it describes
        d["diagnosis_error"] = repr(e)
    base = os.path.join(out_dir, "parity_mismatch_%d" % os.getpid())
    try:
        np.savez_compressed(base + ".npz", ch=ch, ch_o=ch_o, vr=vr, vr_o=vr_o, grid=grid)
        with open(base + ".json", "a") as f:
            f.write(json.dumps(d) + "\n")
    except OSError:
        pass
    return json.dumps(d)
