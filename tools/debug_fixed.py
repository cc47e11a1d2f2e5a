import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, xmimsim_b200 as x
from helpers import Pair
from inputs import example
for n in (2, 8, 64, 512):
    inp = example("srm1155"); inp.n_photons_line = n; inp.n_interactions_trajectory = 1
    inp.discrete = inp.discrete[:1]
    P = Pair(inp); sa = P.grid(hits_per_single=400, n=128)
    o = x.main_options()
    ch, br, vr = P.sim.main_msim(o, sa)
    ch_o, vr_o, cnt = P.oracle(o, sa, 0, n_threads=1)
    d = inp.discrete[0]
    unit = (d.horizontal_intensity + d.vertical_intensity) / n * inp.live_time * 2.0 ** -56
    print("N", n, "unit %.3e" % unit)
    for (z, l) in [(26, 3), (24, 3), (16, 5), (82, 126), (82, 119), (26, 384), (26, 385)]:
        g, r = vr[z - 1, l - 1, 0], vr_o[z - 1, l - 1, 0]
        print("   Z%d slot %d gpu %.10e orc %.10e diff/unit %.1f" % (z, l, g, r, (g - r) / unit))
    P.close()
