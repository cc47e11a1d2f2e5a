"""Feature ablation of a random-input parity difference: the input of a seed with one feature removed at a time.
usage: python tools/debug_random_seed.py SEED [n_photons]  (needs a GPU)"""
import copy, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/tests", ROOT + "/oracle"]
import numpy as np, xmimsim_b200 as x
from helpers import Pair
from random_inputs import random_input

seed = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
base, opts = random_input(seed, n_photons=n)


def run(tag, inp, o=None):
    o = o or x.main_options(**opts)
    P = Pair(inp); sa = P.grid(hits_per_single=300, n=96)
    ch, br, vr = P.sim.main_msim(o, sa)
    ch_o, vr_o, cnt = P.oracle(o, sa, 0)
    e = np.abs(vr - vr_o).max() / np.abs(vr_o).max()
    k = np.unravel_index(np.argmax(np.abs(vr - vr_o)), vr.shape)
    rows = [(vr[:, :, i].sum() - vr_o[:, :, i].sum()) / max(vr_o[:, :, i].sum(), 1e-300) for i in range(vr.shape[2])]
    print("%-28s err %.2e at Z %d slot %d order %d | per-order rel %s | cnt %s | offgrid %s" % (tag, e, k[0] + 1, k[1] + 1, k[2] + 1, " ".join("%.1e" % r for r in rows), cnt[:2], P.sim.workload_stats() if hasattr(P.sim, "workload_stats") else ""), flush=True)
    P.close()


run("as generated", base)
d = copy.deepcopy(base)
for e in d.discrete: e.sigma_x = e.sigma_y = e.sigma_xp = e.sigma_yp = 0.0
run("point source", d)
d = copy.deepcopy(base)
for e in d.discrete: e.distribution_type = 0; e.scale_parameter = 0.0
run("monochromatic lines", d)
d = copy.deepcopy(base); d.exc_layers = []
run("no excitation absorber", d)
d = copy.deepcopy(base); d.det_layers = []
run("no detector absorber", d)
for i in range(len(base.discrete)):
    d = copy.deepcopy(base); d.discrete = [d.discrete[i]]
    run("only line %d (%.2f keV)" % (i, d.discrete[0].energy), d)
d = copy.deepcopy(base); d.n_interactions_trajectory = 1
run("one interaction", d)
for i in range(len(base.layers)):
    d = copy.deepcopy(base); d.layers = [d.layers[i]]; d.reference_layer = 1
    run("only layer %d" % i, d)
d = copy.deepcopy(base); d.nchannels = 2048; d.gain *= 2
run("2048 channels", d)
run("all cascades", base, x.main_options())
