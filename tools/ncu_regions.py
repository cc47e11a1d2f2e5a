#!/usr/bin/env python3
"""Per-source-line and per-region breakdown of an .ncu-rep of the history kernel (source page): warp instructions, stall
samples and the dominant stall reasons of every region of the kernel.  usage: tools/ncu_regions.py x.ncu-rep [n_lines]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, lines = None, None, []
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 2 and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
            lines.append(dict(file=cur, line=int(r[0]), src=r[1].strip(), samples=int(d["# Samples"]), inst=int(d["Instructions Executed"]),
                              tinst=int(d["Thread Instructions Executed"]), stalls=st,
                              local=int(d.get("L2 Theoretical Sectors Local", "0") or 0)))
        except ValueError:
            pass


def region(f, l, srcline):
    if f == "history.cu":
        for hi, name in ((101, "setup"), (146, "transport"), (201, "push"), (228, "sort_batch"), (264, "scheduler"), (351, "pop"), (414, "detector geometry"),
                         (457, "off-grid SA"), (488, "element loop: per-layer setup"), (508, "element loop: Rayleigh"), (543, "element loop: ADV"),
                         (563, "element loop: Compton"), (612, "line loop"), (640, "selection/flush glue")):
            if l <= hi:
                return name
        return "epilogue"
    if f == "history_device.cuh":
        for hi, name in ((57, "node_find"), (62, "row_lerp"), (84, "bilinear"), (100, "to_fixed"), (160, "deposit"), (190, "flush_staged"), (214, "draw_block"),
                         (270, "dirv/elecv"), (325, "compton_energy"), (347, "solid-angle lookup"), (438, "start_photon"), (450, "step_to_plane"),
                         (520, "advanced compton"), (547, "exp_neg"), (700, "select_and_scatter")):
            if l <= hi:
                return name
    if f == "cuda_util.cuh":
        return "philox / u01"
    return f


ts, ti = sum(l["samples"] for l in lines), sum(l["inst"] for l in lines)
R = defaultdict(lambda: dict(samples=0, inst=0, tinst=0, stalls=defaultdict(int), local=0))
for l in lines:
    g = R[region(l["file"], l["line"], l["src"])]
    g["samples"] += l["samples"]; g["inst"] += l["inst"]; g["tinst"] += l["tinst"]; g["local"] += l["local"]
    for k, v in l["stalls"].items():
        g["stalls"][k] += v
print("# %s: %d stall samples, %d warp instructions" % (rep, ts, ti))
print("%-34s %7s %7s %6s  %s" % ("region", "inst%", "smpl%", "lanes", "top stalls (share of region samples)"))
for name, g in sorted(R.items(), key=lambda kv: -kv[1]["inst"]):
    tot = max(1, sum(g["stalls"].values()))
    tops = sorted(g["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print("%-34s %6.1f%% %6.1f%% %6.1f  %s%s" % (name, 100 * g["inst"] / ti, 100 * g["samples"] / ts, g["tinst"] / max(1, g["inst"]),
                                               ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in tops), "  [local sectors %d]" % g["local"] if g["local"] else ""))
print()
print("# hottest lines by instructions")
for l in sorted(lines, key=lambda l: -l["inst"])[:top]:
    print("%-20s %4d  inst %5.2f%%  smpl %5.2f%%  lanes %4.1f  %s" % (l["file"], l["line"], 100 * l["inst"] / ti, 100 * l["samples"] / ts, l["tinst"] / max(1, l["inst"]), l["src"][:110]))
