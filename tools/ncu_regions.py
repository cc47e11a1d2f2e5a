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


import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "xmimsim_b200", "csrc")
# regions of history.cu: (first line matching the marker text, name), found in the source as it is now (run the tool on a
# capture of the same tree)
HIST_MARKERS = [("template <int NL, bool ADV", "setup"), ("auto transport = ", "transport"), ("auto push = ", "push"),
                ("auto energy_class = ", "sort_batch"), ("// ---- scheduler (block-uniform)", "scheduler"), ("Photon p;", "pop / source"),
                ("// ---- forced detection (src/xmi_variance_reduction", "detector geometry"), ("while (__syncthreads_or(sa_pending))", "off-grid SA"),
                ("// warp-uniform loops over layers / elements", "element loop: per-layer setup"), ("// Rayleigh (:342-369)", "element loop: Rayleigh"),
                ("// shell-resolved Compton (xmi_compton_varred,", "element loop: ADV"), ("// Compton (xmi_compton_varred2", "element loop: Compton"),
                ("// ---- fluorescence lines (src/xmi_variance_reduction", "line phase"), ("// ---- atom and interaction selection, scattering", "selection/flush glue"),
                ("// ---- move to the next interaction point and queue there", "move + queue glue"), ("n_inter_local = warp_sum_u64(n_inter_local);", "epilogue")]


def _marks(path, markers):
    out = []
    try:
        lines = open(path).read().splitlines()
    except OSError:
        return out
    for text, name in markers:
        for i, ln in enumerate(lines, 1):
            if text in ln:
                out.append((i, name))
                break
    return sorted(out)


def _functions(path):
    out = []
    try:
        lines = open(path).read().splitlines()
    except OSError:
        return out
    for i, ln in enumerate(lines, 1):
        m = re.match(r"^(?:static )?(?:template <[^>]*>\s*)?__device__ (?:__forceinline__ )?[\w:<> ]+?[ *&](\w+)\(", ln)
        if m:
            out.append((i, m.group(1)))
    return out


_H = _marks(os.path.join(CSRC, "history.cu"), HIST_MARKERS)
_D = _functions(os.path.join(CSRC, "history_device.cuh"))
_GROUP = {"findpos_uniform": "bilinear", "to_fixed": "to_fixed", "fixed_from_scaled": "to_fixed", "warp_sum_u64": "deposit", "smem_u32": "deposit", "stage_red": "deposit",
          "deposit_uniform20": "deposit", "deposit_uniform16": "deposit", "deposit_uniform": "deposit", "deposit_varying": "deposit", "red_global_u64": "flush_staged",
          "normalize3": "dirv/elecv", "update_dirv": "dirv/elecv", "update_elecv": "dirv/elecv", "elec_phi0": "dirv/elecv", "compton_prefetch": "compton_energy",
          "get_solid_angle": "solid-angle lookup", "ran_gaussian": "start_photon", "shard_global_id": "start_photon", "adv_q_from_energy": "advanced compton",
          "adv_energy_from_q": "advanced compton", "adv_shell_cdf": "advanced compton", "adv_sample_q": "advanced compton", "compton_energy_adv": "advanced compton",
          "exp_neg_f32": "exp_neg", "mu_lerp": "row_lerp"}


def region(f, l, srcline):
    if f == "history.cu":
        name = "setup"
        for i, n in _H:
            if i <= l:
                name = n
        return name
    if f == "history_device.cuh":
        name = "history_device.cuh"
        for i, n in _D:
            if i <= l:
                name = _GROUP.get(n, n)
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
