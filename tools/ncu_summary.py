#!/usr/bin/env python3
"""Summarise an .ncu-rep of the history kernel: headline counters + hottest source lines (needs ncu on PATH).
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/prefix"""
import csv, json, subprocess, sys
rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__sass_average_branch_targets_threads_uniform.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__inst_executed_op_global_red.sum',
        'smsp__inst_executed_op_generic_atom_dot_alu.sum', 'lts__t_sectors_srcunit_tex_op_red.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
out = {}
for k in keys:
    for i, h in enumerate(hdr):
        if h == k:
            out[k] = vals[i] + ' ' + units[i]
json.dump(out, open(prefix + "_ncu_full.json", "w"), indent=1)
print(json.dumps(out, indent=1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, lines = None, []
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= 8 and r[0].isdigit():
        try:
            lines.append((cur, int(r[0]), r[1].strip(), int(r[4]), int(r[7])))
        except ValueError:
            pass
ts, ti = sum(l[3] for l in lines), sum(l[4] for l in lines)
lines.sort(key=lambda l: -l[3])
with open(prefix + "_hot_lines.txt", "w") as f:
    f.write("# %s: total stall samples %d, warp instructions %d; file line %%samples %%instructions source\n" % (rep, ts, ti))
    for l in lines[:45]:
        f.write('%-22s %4d  %5.1f%%  %5.1f%%  %s\n' % (l[0], l[1], 100 * l[3] / ts, 100 * l[4] / ti, l[2][:120]))
print(open(prefix + "_hot_lines.txt").read())
