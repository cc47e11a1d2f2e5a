"""Kernel-resident throughput of the history engine on the BASELINE.json configurations that are not the bench
line (configs[0] srm1155, configs[3] synthetic 10 layers / 8 interactions, configs[4] 1000-interval tube spectrum +
detector response), one GPU, inputs resident.  Prints one JSON object; results are kept under profiles/."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from bench import algorithmic_bytes  # noqa: E402
from inputs import example, synthetic_layers, ebel_like  # noqa: E402


def run(name, inp, reps=2, detector=False, **optkw):
    sim = x.Simulation(inp, quality=0)
    g, r, t = sim.solid_angle_calculation(hits_per_single=5000, seed=1)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    opt = x.main_options(**optkw)
    ms = []
    for _ in range(reps + 1):
        ex = sim.main_msim_device(opt, sa)
        ms.append(ex.kernel_ms)
    ms = ms[1:]
    st = sim.workload_stats()
    out = {"workload": name, "histories": int(ex.n_histories), "kernel_ms": [round(m, 2) for m in ms],
           "histories_per_s": ex.n_histories / (min(ms) * 1e-3),
           "interactions_per_history": ex.n_interactions / max(1, ex.n_histories),
           "algorithmic_bytes_per_history": algorithmic_bytes(st, len(st)) / max(1, ex.n_histories)}
    out["algorithmic_GBps"] = out["algorithmic_bytes_per_history"] * out["histories_per_s"] / 1e9
    if detector:
        t0 = time.time()
        ch, br, vr = sim.main_msim(opt, sa)
        t1 = time.time()
        conv = sim.detector_convolute_all(ch, None, vr, opt)
        t2 = time.time()
        out["main_msim_wall_s"] = round(t1 - t0, 3)
        out["detector_response_wall_s"] = round(t2 - t1, 4)
        out["detected_counts_last_order"] = float(np.sum(ch[-1]))
    sim.close()
    return out


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    res = []
    a = example("srm1155")                                                     # configs[0]: the file's own 150 000 photons/line
    res.append(run("configs[0] srm1155.xmsi as shipped (26 lines x 150000 photons, 4 interactions)", a))
    b = synthetic_layers(n_photons=int(125_000_000 * scale), n_int=8)          # configs[3]: one of 8 shards of 1e9
    res.append(run("configs[3] synthetic 10 layers, 8 interactions, 1.25e8 histories (1/8 of 1e9)", b, reps=1))
    c = ebel_like(n_intervals=1000, n_photons_interval=int(100_000 * scale), n_photons_line=int(100_000 * scale))
    res.append(run("configs[4] 1000-interval tube continuum + 5 lines, 1e5 photons each, escape peaks + pile-up", c, reps=1,
                   detector=True, use_sum_peaks=1, use_escape_peaks=0))
    print(json.dumps({"results": res}))


if __name__ == "__main__":
    main()
