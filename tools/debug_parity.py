"""Engine vs oracle on one shipped example: prints the per-order maximum deviation and the worst channels / lines.
usage: python tools/debug_parity.py [example] [photons per line] [interactions]  (needs a GPU)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, xmimsim_b200 as x
from helpers import Pair
from inputs import example
name = sys.argv[1] if len(sys.argv) > 1 else "srm1155"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
inp = example(name); inp.n_photons_line = n
if len(sys.argv) > 3: inp.n_interactions_trajectory = int(sys.argv[3])
P = Pair(inp); sa = P.grid(hits_per_single=400)
o = x.main_options()
ch, br, vr = P.sim.main_msim(o, sa)
ch_o, vr_o, cnt = P.oracle(o, sa, 0)
print("cnt", cnt)
print("order sums gpu", np.diff(ch.sum(axis=1)))
print("order sums orc", np.diff(ch_o.sum(axis=1)))
d = vr - vr_o
for k in range(vr.shape[2]):
    tot_g, tot_o = vr[:, :, k].sum(), vr_o[:, :, k].sum()
    print("order", k + 1, "hist total gpu %.8e orc %.8e rel %.3e" % (tot_g, tot_o, (tot_g - tot_o) / max(tot_o, 1e-300)),
          "rayl rel %.3e" % ((vr[:, 383, k].sum() - vr_o[:, 383, k].sum()) / max(vr_o[:, 383, k].sum(), 1e-300)),
          "compt rel %.3e" % ((vr[:, 384, k].sum() - vr_o[:, 384, k].sum()) / max(vr_o[:, 384, k].sum(), 1e-300)),
          "lines rel %.3e" % ((vr[:, :383, k].sum() - vr_o[:, :383, k].sum()) / max(vr_o[:, :383, k].sum(), 1e-300)))
idx = np.argsort(-np.abs(d).ravel())[:12]
for i in idx:
    z, l, k = np.unravel_index(i, d.shape)
    print("Z", z + 1, "slot", l + 1, "order", k + 1, "gpu %.8e orc %.8e rel %.3e" % (vr[z, l, k], vr_o[z, l, k], d[z, l, k] / max(abs(vr_o[z, l, k]), 1e-300)))
dc = ch - ch_o
for i in np.argsort(-np.abs(dc).ravel())[:8]:
    k, c = np.unravel_index(i, dc.shape)
    print("row", k, "ch", c, "gpu %.8e orc %.8e" % (ch[k, c], ch_o[k, c]))
