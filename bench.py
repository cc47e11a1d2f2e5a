#!/usr/bin/env python3
"""bench.py -- photon histories/s of the forced-detection history engine (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's B200 engine
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU; the engine's own NCCL all-reduce

A step = one xmi_main_msim pass over the workload (all source lines, every history followed to termination with all
forced-detection deposits).  Headline workload at N GPUs: examples/srm1412.xmsi (BASELINE configs[1]) with
n_photons_line = 1e7 * N (weak scaling: 2.5e8 histories per GPU), 4 interactions, variance reduction on.
`value` times the kernels with every input resident in HBM; `e2e` goes through the public xmi_main_msim-shaped call with
host buffers (solid-angle grid host->device, histograms device->host, epilogue on the host).  For N > 1 both go through
the product's multi-GPU entry points (xmb_main_msim_multi[_raw]: kernel -> limbs -> ncclAllReduce on one stream);
torch.distributed only carries the 128-byte NCCL id, the barriers and the max over ranks of the timings.

Extra keys of the same line: "configs3" (BASELINE configs[3]: synthetic 10 layers, 8 interactions, 1e9 histories in
TOTAL -- strong scaling over N, with a digest of the reduced integer histograms that must not change with N) and
"configs4" (BASELINE configs[4]: 1000-interval continuum, 1.25e9 histories per GPU = 1e10 on 8, plus the detector
response with escape peaks and pile-up, timed)."""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "photon_histories_per_s"
UNIT = "histories/s"
PHOTONS_PER_LINE = 10_000_000
CONFIGS3_TOTAL = 1_000_000_000        # BASELINE configs[3]: 1e9 histories in total, sharded over the GPUs
CONFIGS4_PER_GPU = 1_250_000_000      # BASELINE configs[4]: 1e10 histories on 8 GPUs


def load_workload(name, n_gpus, photons_per_line):
    from xmimsim_b200 import workloads
    inp = workloads.example(name)
    inp.n_photons_line = photons_per_line * n_gpus
    return inp


def workload_config(name, inp, photons_per_line):
    """The `config` object both arms print (same workload string: the driver compares the two lines)."""
    return {"workload": "%s.xmsi (BASELINE configs[1]): %d lines x %.0e photons/line per GPU, %d interactions, variance reduction on, "
                        "M-lines + full cascade" % (name, len(inp.discrete), photons_per_line, inp.n_interactions_trajectory),
            "cross_sections": "analytic surrogate provider (xraylib unavailable offline)",
            "arithmetic": "f64 throughout; the attenuation factor and products of the fluorescence-line deposits in f32 (sums: exact 64-bit integers)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for i, nm in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(s[i].lower().startswith("active") for s in self.samples):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples)}


def algorithmic_bytes(stats, n_layers):
    """bytes one interaction in layer l must read or update once (f64 = 8 B):
    16 nL + 32 + nE(l) (16 nL + 224) + n_active(l) (32 + 8 nL) + 16 nE(l) + 112."""
    total = 0
    for inter, n_el, n_act in stats:
        per = 16 * n_layers + 32 + n_el * (16 * n_layers + 224) + n_act * (32 + 8 * n_layers) + 16 * n_el + 112
        total += inter * per
    return total


def cpu_baseline_run(inp, sa_grid, n_sample_per_line, n_threads):
    """The oracle (CPU restatement of the reference algorithm) on a bounded sample of the same workload:
    same input with n_photons_line reduced (cost is linear in photons, src/xmi_main.F90:618)."""
    import copy
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc  # noqa: F401  (test infrastructure: allowed here as the timed CPU baseline only)
    import xmimsim_b200 as x
    from helpers import Pair
    d = copy.deepcopy(inp)
    d.n_photons_line = n_sample_per_line
    P = Pair(d)
    g, r, t = sa_grid
    sa = P.sim.make_solid_angle(g, r, t)
    t0 = time.perf_counter()
    P.oracle(x.main_options(), sa, 0, n_threads=n_threads)
    dt = time.perf_counter() - t0
    n = P.n_total
    P.close()
    return n, dt


def oracle_solid_angle_grid(inp, n=96, hits=400):
    """A small solid-angle grid from the oracle itself (reference arm must not touch the GPU engine)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    import xmimsim_b200 as x
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    r_full, t_full = orc.solid_angle_axes(C.pointer(ci.input), od)
    r = np.linspace(r_full[0], r_full[-1], n)
    t = np.linspace(t_full[0], t_full[-1], n)
    sa, _ = orc.solid_angle_grid(od, r, np.arange(n), t, np.arange(n), n, hits, 1, n_threads=os.cpu_count() or 1)
    return sa, r, t


_REAL_STDOUT = None


def quiet_stdout():
    """The driver reads ONE JSON line from stdout: libraries that write banners there (NCCL prints its version on
    file descriptor 1 when the first communicator is created) are sent to stderr; emit() writes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Fortran binary cannot be built
    here) with all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    inp = load_workload(args.workload, 1, args.photons_per_line)
    grid = oracle_solid_angle_grid(inp)
    sample_per_line = args.reference_sample_per_line
    times = []
    n = 0
    for i in range(args.warmup + args.steps):
        n, dt = cpu_baseline_run(inp, grid, sample_per_line, cores)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    cfg = workload_config(args.workload, inp, args.photons_per_line)
    cfg["note"] = "CPU oracle port of src/xmi_main.F90 + src/xmi_variance_reduction.F90, OpenMP over photons as the reference"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d histories per step (n_photons_line=%d of %d), cost linear in photons"
                                       % (n, sample_per_line, args.photons_per_line)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def ncu_counters(workload, n, timeout=180):
    """Warp instructions and DRAM bytes of the history kernel, MEASURED in this run: a sub-process runs a bounded sample
    of the workload under ncu (outside every timed region).  Returns a dict, or {"unavailable": why}."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"unavailable": "ncu not found"}
    if any(k.startswith(("CUDA_INJECTION", "NVTX_INJECTION", "NV_NSIGHT_INJECTION")) for k in os.environ):
        return {"unavailable": "this process runs under a profiler already: no nested ncu"}
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = os.path.join(out_dir, "bench_ncu_%s_%d.csv" % (workload, os.getpid()))
    metrics = ("smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
               "smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_sector_hit_rate.pct")
    cmd = [ncu, "--metrics", metrics, "--clock-control", "none", "--print-units", "base", "-k", "regex:xmb_history_kernel", "-s", "1", "-c", "1", "--csv",
           "--log-file", log, sys.executable, os.path.join(ROOT, "tools", "kernel_counters.py"), workload, str(n), "2"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except Exception as exc:
        return {"unavailable": "ncu sub-process: %s" % str(exc)[:100]}
    info = None
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            info = json.loads(ln)
    if info is None or not os.path.exists(log):
        return {"unavailable": "ncu sub-process rc=%d: %s" % (r.returncode, (r.stderr or r.stdout)[-160:])}
    vals = {}
    rows = [ln for ln in open(log) if ln.startswith('"')]
    for row in csv.DictReader(rows):
        try:
            vals[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
            vals[row["Metric Name"] + "|unit"] = row.get("Metric Unit", "")
        except (KeyError, ValueError):
            pass
    if "smsp__inst_executed.sum" not in vals:
        return {"unavailable": "no counters in the ncu log (ERR_NVGPUCTRPERM?)"}

    def in_bytes(name):
        v, u = vals.get(name, 0.0), vals.get(name + "|unit", "byte").lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
    h = info["histories"]
    return {"sample_histories": h, "warp_inst_per_history": vals["smsp__inst_executed.sum"] / h,
            "thread_inst_per_warp_inst": vals.get("smsp__thread_inst_executed.sum", 0.0) / vals["smsp__inst_executed.sum"],
            "dram_bytes_per_history": (in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")) / h,
            "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l2_hit_pct": vals.get("lts__t_sector_hit_rate.pct"),
            "how": "ncu (smsp__inst_executed.sum, dram__bytes_read/write.sum) on one launch of %d histories of the same workload, "
                   "sub-process of this bench run, outside the timed regions" % h}


def run_ours(args):
    import numpy as np
    import torch
    import xmimsim_b200 as x
    from xmimsim_b200 import abi, workloads
    from xmimsim_b200.engine import Comm
    n_gpus = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != n_gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torchrun --nproc-per-node %d)" % (n_gpus, world, n_gpus))
    if not torch.cuda.is_available() or abi.lib().xmb_cuda_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist, comm = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # the engine's own communicator: rank 0 makes the NCCL id, the launcher's channel (here torch.distributed) carries it
        idt = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, device=local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        tt = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def gather(v):
        if dist is None:
            return [v]
        out = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.tensor([v], device="cuda", dtype=torch.float64))
        return [float(t.item()) for t in out]

    opt = x.main_options()

    def device_step(sim, sa, o=opt):
        """Inputs resident in HBM; N > 1: kernel -> limbs -> ncclAllReduce inside the library, reduced limbs stay in HBM."""
        if comm is None:
            return sim.main_msim_device(o, sa, device=local_rank)
        return sim.main_msim_multi_device(comm, o, sa)

    def e2e_step(sim, sa, o=opt):
        """The public call with host buffers: grid host->device, histograms device->host, epilogue on the host."""
        if comm is None:
            return sim.main_msim(o, sa)
        return sim.main_msim_multi(comm, o, sa)

    def timed(fn, steps):
        """`steps` calls bracketed by barrier + synchronize; device time by CUDA events on the stream the library launches on
        (the legacy default stream, which is also torch's current stream), wall clock beside it; max of the two, max over ranks."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        outs = [fn() for _ in range(steps)]
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        return max_over_ranks(max(wall, ev0.elapsed_time(ev1) / 1e3)), outs

    def limbs_digest(sim):
        """SHA-256 of the reduced 128-bit histogram integers.  A slot is two limbs (low 48 bits, the rest); after the sum over
        ranks the low limb carries up to log2(N) extra bits, so the VALUE lo + (hi << 48) is hashed, not the limb pair."""
        ptr, n = sim.device_limbs()
        holder = type("H", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}})()
        a = torch.as_tensor(holder, device="cuda").cpu().numpy().view(np.uint64).reshape(-1, 2)
        lo, hi = a[:, 0], a[:, 1]
        hi = hi + (lo >> np.uint64(48))                       # carry of the low limb (hi < 2^63 here)
        lo = lo & np.uint64(0xFFFFFFFFFFFF)
        return hashlib.sha256(np.stack([lo, hi], axis=1).tobytes()).hexdigest()[:16]

    # ================= headline: BASELINE configs[1] ============================================================
    inp = load_workload(args.workload, n_gpus, args.photons_per_line)
    sim = x.Simulation(inp, quality=args.table_quality)
    # set-up (untimed): solid-angle grid on the GPU, as the reference computes/caches it before xmi_main_msim
    t0 = time.perf_counter()
    grid, r_vals, t_vals = sim.solid_angle_calculation(opt, hits_per_single=5000, seed=1)
    sa_wall = time.perf_counter() - t0
    sa_kernel_ms = sim.L.xmb_solid_angle_last_ms()
    grid_pinned = torch.from_numpy(grid.copy()).pin_memory()
    sa = sim.make_solid_angle(grid_pinned.numpy(), r_vals.copy(), t_vals.copy())
    n_total_job = len(inp.discrete) * inp.n_photons_line

    for _ in range(args.warmup):
        device_step(sim, sa)
    sampler = ClockSampler(local_rank)
    sampler.start()
    elapsed, exs = timed(lambda: device_step(sim, sa), args.steps)
    kernel_ms = [e.kernel_ms for e in exs]
    launches = sum(int(e.n_launches) for e in exs)
    ex = exs[-1]
    per_rank_kernel_ms = gather(sum(kernel_ms) / len(kernel_ms))
    stats = sim.workload_stats()
    n_hist_rank, interactions = int(ex.n_histories), int(ex.n_interactions)
    # ---- end to end through the public call with host buffers ----------------------------------------------------
    e2e_step(sim, sa)
    e2e_elapsed, _ = timed(lambda: e2e_step(sim, sa), args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    _, n_words = sim.device_limbs()
    clocks = sampler.summary()

    line = None
    if rank == 0:
        pk, pk_src = peaks()
        value = n_total_job * args.steps / elapsed
        k_ms = sum(kernel_ms) / len(kernel_ms)
        bytes_launch = algorithmic_bytes(stats, len(inp.layers))
        achieved = bytes_launch / (k_ms * 1e-3) / 1e9
        cfg = workload_config(args.workload, inp, args.photons_per_line)
        cfg.update({"histories_per_step": n_total_job,
                    "l2_policy": "every step re-reads the same ~20 MB of tables (L2-resident by design); accumulators are zeroed "
                                 "(memset) and the per-CTA photon queues (footprint > L2) rewritten each step",
                    "parallelism": ("photon-id shards, 1 ncclAllReduce(uint64) of %d limbs inside the library (xmb_main_msim_multi_raw)" % n_words)
                    if world > 1 else "single GPU"})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": n_total_job * args.steps / e2e_elapsed, "unit": UNIT,
                    "h2d_bytes_per_step": int(grid.nbytes + r_vals.nbytes + t_vals.nbytes),
                    "d2h_bytes_per_step": int(n_words * 8), "steps": args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "hbm_roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                             "traffic": None, "peak_source": pk_src, "kernel": "xmb_history_kernel", "kernel_ms": k_ms,
                             "algorithmic_bytes_per_launch": bytes_launch, "bytes_per_history": bytes_launch / max(1, n_hist_rank),
                             "note": "SURVEY.md 8(d) algorithmic bytes; the tables are L2-resident and warps skip the shells their "
                                     "photons cannot ionise, so DRAM carries a few percent of this (traffic) and frac can exceed 1: "
                                     "not the operative bound"},
            "solid_angle_grid": {"workload": "headline geometry (set-up of the timed runs)", "seconds_wall": sa_wall, "kernel_ms": sa_kernel_ms,
                                 "rays": 1024 * 1024 * 5000},
        }
        line["roofline"] = {"kernel": "xmb_history_kernel", "kernel_ms": k_ms, "per_rank_kernel_ms": per_rank_kernel_ms,
                            "mean_interactions_per_history": interactions / max(1, n_hist_rank)}
    # ================= BASELINE configs[3]: 10 layers, 8 interactions, 1e9 histories in total (strong scaling) ======
    extra = {}
    if not args.headline_only:
        total3 = max(1024 * world, int(args.configs3_total))
        inp3 = workloads.synthetic_layers(n_photons=total3, n_int=8)
        sim3 = x.Simulation(inp3, quality=args.table_quality)
        g3, r3, t3 = sim3.solid_angle_calculation(opt, hits_per_single=5000, seed=1)
        sa3 = sim3.make_solid_angle(g3.copy(), r3.copy(), t3.copy())
        device_step(sim3, sa3)
        el3, ex3 = timed(lambda: device_step(sim3, sa3), 1)
        k3 = gather(ex3[0].kernel_ms)
        dig3 = limbs_digest(sim3)
        extra["configs3"] = {"workload": "BASELINE configs[3]: synthetic 10 layers (Z = 8..82), 8 interactions, %.3g histories in TOTAL sharded over "
                                         "%d GPU(s), product NCCL all-reduce" % (total3, world),
                             "scaling": "strong", "histories": total3, "value": total3 / el3, "unit": UNIT, "ms_per_step": 1e3 * el3,
                             "per_rank_kernel_ms": k3, "steps": 1, "warmup": 1,
                             "limbs_sha256_16": dig3,
                             "digest_note": "SHA-256 of the reduced integer histograms: identical at every GPU count (bit-exactness)"}
        sim3.close()
        # ============= BASELINE configs[4]: 1000-interval continuum, 1.25e9 histories per GPU + detector response ======
        nseg = workloads.n_source_segments(workloads.ebel_tube(1))
        per = max(1, int(args.configs4_per_gpu) * world // nseg)
        inp4 = workloads.ebel_tube(n_photons_interval=per)
        sim4 = x.Simulation(inp4, quality=args.table_quality)
        g4, r4, t4 = sim4.solid_angle_calculation(opt, hits_per_single=5000, seed=1)
        sa4 = sim4.make_solid_angle(g4.copy(), r4.copy(), t4.copy())
        opt4 = x.main_options(use_sum_peaks=1, use_escape_peaks=1)
        device_step(sim4, sa4, opt4)
        el4, ex4 = timed(lambda: device_step(sim4, sa4, opt4), 1)
        total4 = nseg * per
        c4 = {"workload": "BASELINE configs[4]: Ebel tube spectrum (Ag anode, 40 kV, dE 0.039 keV: %d source segments from xmb_tube_ebel), %.4g histories "
                          "(%.3g per GPU), 4 interactions; then escape-ratio Monte Carlo + detector response with escape peaks and pile-up "
                          "(pulse width 1e-6 s)" % (nseg, total4, total4 / world),
              "scaling": "weak", "histories": total4, "value": total4 / el4, "unit": UNIT, "ms_per_step": 1e3 * el4,
              "per_rank_kernel_ms": gather(ex4[0].kernel_ms), "steps": 1, "warmup": 1}
        # detector response of the full result on rank 0 (the reference convolutes on rank 0 after MPI_Reduce, bin/xmimsim.c:413-526)
        ch4, br4, vr4 = e2e_step(sim4, sa4, opt4)
        if rank == 0:
            t0 = time.perf_counter()
            er = sim4.escape_ratios_calculation(options=opt4)
            t1 = time.perf_counter()
            conv4 = sim4.detector_convolute_all(ch4, br4, vr4, opt4, er.contents)
            t2 = time.perf_counter()
            c4["escape_ratios_s"] = t1 - t0
            c4["escape_ratios_kernel_ms"] = sim4.L.xmb_escape_ratios_last_ms()
            c4["detector_response_s"] = t2 - t1
            c4["detector_response_kernel_ms"] = sim4.L.xmb_detector_last_ms()
            c4["detected_counts_last_order"] = float(conv4[-1].sum())
            sim4.escape_ratios_free(er)
        extra["configs4"] = c4
        sim4.close()
        barrier()

    # ============= headline workload with the line density of real xraylib data ======================
    if not args.headline_only:
        simd = x.Simulation(inp, quality=args.table_quality, provider=abi.lib().xmb_xrl_surrogate_dense())
        device_step(simd, sa)
        eld, exd = timed(lambda: device_step(simd, sa), 1)
        zs, ls = simd.slot_map()
        extra["dense_lines"] = {"workload": "the headline workload with the stand-in provider's dense line set (%d active forced-detection lines "
                                            "instead of %d; srm1155 would have 326, xraylib data ~321)" % (int((ls < 384).sum()), int((sim.slot_map()[1] < 384).sum())),
                                "value": n_total_job / eld, "unit": UNIT, "ms_per_step": 1e3 * eld, "per_rank_kernel_ms": gather(exd[0].kernel_ms),
                                "steps": 1, "warmup": 1}
        simd.close()
    if rank == 0:
        line.update(extra)
        # ---- issue roofline, measured in this run (ncu sub-process on a bounded sample; other ranks wait) -------------
        cnt = {"unavailable": "--no-ncu"} if args.no_ncu else ncu_counters(args.workload, args.ncu_sample_per_line)
        sms = 148
        try:
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        except Exception:
            pass
        clock_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        rf = line["roofline"]
        if "unavailable" not in cnt:
            inst_launch = cnt["warp_inst_per_history"] * n_hist_rank
            ach = inst_launch / (rf["kernel_ms"] * 1e-3)
            peak = sms * 4 * clock_hz
            rf.update({"bound": "issue", "achieved": ach, "peak": peak, "unit": "warp-inst/s", "frac": ach / peak,
                       "traffic": cnt["dram_bytes_per_history"] * n_hist_rank,
                       "peak_source": "%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled in the timed region)" % (sms, clock_hz / 1e6),
                       "counters": cnt,
                       "note": "the kernel is bound by instruction issue and dependent-load latency, not by HBM (hbm_roofline.frac is kept for "
                               "continuity): achieved = measured warp instructions per history x histories per launch / kernel time"})
            line["hbm_roofline"]["traffic"] = rf["traffic"]
        else:
            # no counters on this box: fall back to the committed ncu capture of the same kernel, and say so
            fb = os.path.join(ROOT, "profiles", "history_kernel_counters.json")
            rf.update({"bound": "issue", "unit": "warp-inst/s", "counters": cnt})
            if os.path.exists(fb):
                c = json.load(open(fb))
                ach = c["warp_inst_per_history"] * n_hist_rank / (rf["kernel_ms"] * 1e-3)
                peak = sms * 4 * clock_hz
                rf.update({"achieved": ach, "peak": peak, "frac": ach / peak, "traffic": c["dram_bytes_per_history"] * n_hist_rank,
                           "peak_source": "%d SMs x 4 schedulers x %.0f MHz" % (sms, clock_hz / 1e6),
                           "note": "counters from the committed capture profiles/history_kernel_counters.json (ncu unavailable in this run)"})
                line["hbm_roofline"]["traffic"] = rf["traffic"]
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            # the full grid: a sub-sampled one ends below the last r / theta and would send those interaction points
            # through the reference's 5000-ray on-the-spot solid angle (src/xmi_solid_angle_f.F90:783-789)
            n, dt = cpu_baseline_run(inp, (grid, r_vals, t_vals), args.reference_sample_per_line, cores)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d histories (n_photons_line=%d of %d), oracle port of the reference algorithm, "
                                              "cost linear in photons" % (n, args.reference_sample_per_line, args.photons_per_line)}
            # BASELINE's second metric, "solid-angle grid s": the reference's own OpenCL kernel source compiled for the host
            # (oracle/_ref, fp32, all host threads) on every 8th row and column of the same grid, scaled to the full grid
            # BASELINE configs[2]: the grid of the examples/srm1132.xmsi geometry (conical collimator), 1024 x 1024 x 5000 rays
            sim2 = x.Simulation(workloads.example("srm1132"), quality=args.table_quality)
            walls, kms = [], []
            for _ in range(3):                                                         # the call a host makes: axes, kernel, 12 MB device->host
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g2, r2, t2 = sim2.solid_angle_calculation(opt, hits_per_single=5000, seed=1)
                walls.append(time.perf_counter() - t0)
                kms.append(sim2.L.xmb_solid_angle_last_ms())
            sa2_wall, sa2_ms = min(walls), min(kms)
            sg = {"workload": "BASELINE configs[2]: solid-angle grid of the srm1132 geometry, 1024 x 1024 points x 5000 rays",
                  "seconds_wall": sa2_wall, "seconds_wall_all_calls": walls, "kernel_ms": sa2_ms, "rays": 1024 * 1024 * 5000, "rays_per_s": 1024 * 1024 * 5000 / (sa2_ms * 1e-3),
                  "nonzero_points": int((g2 > 0).sum())}
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import ref
                if ref.available():
                    d = sim2.derived
                    t0 = time.perf_counter()
                    ref.solid_angle_grid_cl(r2[::8], t2[::8], d.collimator_present, d.detector_radius, d.collimator_radius,
                                            d.collimator_height, 5000)
                    dt = time.perf_counter() - t0
                    sg["cpu_reference"] = {
                        "seconds_full_grid": dt * 64.0, "cores": cores, "kind": "reference",
                        "sample": "128 x 128 of the 1024 x 1024 points x 5000 rays, src/xmi_kernels.cl compiled for the host, x 64"}
            except Exception as exc:   # the checker is optional for the bench line
                sg["cpu_reference"] = {"unavailable": str(exc)[:120]}
            sim2.close()
            # issue roofline of the grid kernel: 4.15 warp instructions per ray (8.72e9 per 2.10e9 rays, ncu capture
            # profiles/r2_solid_angle_kernel_v2_ncu_full.json: issue slots 75 % busy, fp64 pipe 52 %)
            sg["roofline"] = {"bound": "issue", "unit": "warp-inst/s", "achieved": 4.15 * sg["rays_per_s"], "peak": sms * 4 * clock_hz,
                              "frac": 4.15 * sg["rays_per_s"] / (sms * 4 * clock_hz), "warp_inst_per_ray": 4.15,
                              "source": "instruction count from the committed ncu capture of the same kernel; time measured here"}
            line["solid_angle_configs2"] = sg
        emit(line)
    barrier()
    sim.close()
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="srm1412")
    ap.add_argument("--photons-per-line", type=int, default=PHOTONS_PER_LINE)
    ap.add_argument("--reference-sample-per-line", type=int, default=120000,
                    help="photons per line of the bounded CPU sample (25 lines -> 3e6 histories, ~10-20 s of host work)")
    ap.add_argument("--table-quality", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the configs[3] / configs[4] legs")
    ap.add_argument("--configs3-total", type=float, default=CONFIGS3_TOTAL)
    ap.add_argument("--configs4-per-gpu", type=float, default=CONFIGS4_PER_GPU)
    ap.add_argument("--no-ncu", action="store_true", help="do not start the ncu sub-process that measures instructions / DRAM bytes per history")
    ap.add_argument("--ncu-sample-per-line", type=int, default=400000)
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
