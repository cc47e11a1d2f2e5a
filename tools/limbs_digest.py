"""Digest of the raw fixed-point sums (limbs) of the history engine on a set of inputs: two builds of the kernel must
print identical digests (integer accumulation: the regrouping of photons must not change a single bit)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers, ebel_like, caso4  # noqa: E402


def digest(inp, **optkw):
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r, t = r_full[::8], t_full[::8]
    rr = np.random.default_rng(5)
    sa = sim.make_solid_angle(rr.uniform(1e-4, 2e-4, (t.size, r.size)), r.copy(), t.copy())   # synthetic grid: only the kernel is compared
    opt = x.main_options(**optkw)
    limbs, ex = sim.main_msim_raw(opt, sa)
    l2, _ = sim.main_msim_raw(opt, sa, rank=1, n_ranks=3)
    sim.close()
    return {"n": int(ex.n_histories), "inter": int(ex.n_interactions), "ms": round(ex.kernel_ms, 2),
            "sha": hashlib.sha256(limbs.tobytes()).hexdigest()[:16], "sha_shard": hashlib.sha256(l2.tobytes()).hexdigest()[:16]}


def collect():
    out = {}
    a = example("srm1412"); a.n_photons_line = 40000
    out["srm1412"] = digest(a)
    b = example("srm1155"); b.n_photons_line = 20000
    out["srm1155_nocascade"] = digest(b, use_cascade_auger=0, use_cascade_radiative=0)
    out["synthetic10"] = digest(synthetic_layers(n_photons=1_000_000, n_int=8))
    out["synthetic10_100lines"] = digest(synthetic_layers(n_photons=10_000, n_int=8, n_lines=100))
    out["ebel"] = digest(ebel_like(n_intervals=200, n_photons_interval=2000, n_photons_line=2000))
    out["caso4"] = digest(caso4())
    c = example("srm1132"); c.n_photons_line = 20000
    out["srm1132_adv"] = digest(c, use_advanced_compton=1)
    return out


def main():
    print(json.dumps(collect(), indent=1))


if __name__ == "__main__":
    main()
