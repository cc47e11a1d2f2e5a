"""bench.py's contract on a box without a GPU: the reference arm (the CPU oracle port, SURVEY.md 8d "CPU side-by-side")
prints one JSON line with the keys the driver reads; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--reference-sample-per-line", "400")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photon_histories_per_s" and d["unit"] == "histories/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_fails_loudly_without_a_gpu():
    from xmimsim_b200 import abi
    if abi.lib().xmb_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines())      # no bench line from a fallback


def test_library_banners_do_not_reach_the_json_line():
    """NCCL prints its version on file descriptor 1 when the first communicator is created: bench.py sends everything written to
    fd 1 during the run to stderr and writes its JSON line to the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.quiet_stdout(); os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); "
            "print('a print from a library'); bench.emit({'metric': 'x', 'value': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-400:]
    assert r.stdout == '{"metric": "x", "value": 1}\n'
    assert "NCCL version" in r.stderr and "a print from a library" in r.stderr
