"""Repeats the GPU side of tests/test_history_gpu.py::test_synthetic_ten_layers_eight_interactions in one process:
fresh Simulation, kernel-computed 128 x 128 solid-angle grid, main_msim -- every repetition must reproduce the first."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import synthetic_layers, example  # noqa: E402


def one(inp, grid_n):
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r = np.linspace(r_full[0], r_full[-1], grid_n)
    t = np.linspace(t_full[0], t_full[-1], grid_n)
    g, _ = sim.solid_angle_grid(r, t, hits_per_single=400, seed=3)
    sa = sim.make_solid_angle(g, r, t)
    ch, br, vr = sim.main_msim(x.main_options(), sa)
    limbs, ex = sim.main_msim_raw(x.main_options(), sa)
    st = sim.workload_stats()
    sim.close()
    return g.copy(), ch.copy(), vr.copy(), limbs.copy()


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    out = {}
    for name, mk in (("synthetic10", lambda: synthetic_layers(n_photons=30000, n_int=8)),):
        ref = one(mk(), 128)
        bad = {"grid": 0, "channels": 0, "history": 0, "limbs": 0}
        worst = 0.0
        for _ in range(reps):
            cur = one(mk(), 128)
            for key, a, b in zip(("grid", "channels", "history", "limbs"), ref, cur):
                if not np.array_equal(a, b):
                    bad[key] += 1
                    if key == "channels":
                        worst = max(worst, float(np.abs(a - b).max() / np.abs(a).max()))
        out[name] = {"reps": reps, "differ": bad, "worst_channel_diff_rel": worst}
    print(json.dumps({"lib": os.environ.get("XMIMSIM_B200_LIB", "default"), "results": out}))


if __name__ == "__main__":
    main()
