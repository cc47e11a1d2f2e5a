import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import xmimsim_b200 as x
from inputs import example
inp = example("srm1132")
sim = x.Simulation(inp, quality=0)
g, r, t = sim.solid_angle_calculation(hits_per_single=5000, seed=1)
print(g.sum())
