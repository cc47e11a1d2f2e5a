"""Multi-GPU drivers of the product (multi_gpu.cu): shards + ONE NCCL all-reduce of the integer histograms, the role of
the reference's MPI split and MPI_Reduce (src/xmi_main.F90:314,574; bin/xmimsim.c:396-413).  The outputs must equal the
single-GPU outputs bit for bit at every GPU count the box offers (1 on the default test box; `gpurun --gpus N` for more)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import xmimsim_b200 as x
from xmimsim_b200 import abi
from xmimsim_b200.engine import Comm
from inputs import example, synthetic_layers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_binds_at_run_time():
    assert abi.lib().xmb_nccl_version() >= 22000


def _setup(inp):
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r = np.linspace(r_full[0], r_full[-1], 128); t = np.linspace(t_full[0], t_full[-1], 128)
    g, _ = sim.solid_angle_grid(r, t, hits_per_single=300, seed=3)
    return sim, sim.make_solid_angle(g, r, t)


@pytest.mark.gpu
def test_communicator_of_one_rank_equals_the_plain_call():
    inp = example("srm1155"); inp.n_photons_line = 3000
    sim, sa = _setup(inp)
    ch, br, vr = sim.main_msim(x.main_options(), sa)
    comm = Comm(Comm.unique_id(), 0, 1)
    ch2, br2, vr2 = sim.main_msim_multi(comm, x.main_options(), sa)
    assert np.array_equal(ch, ch2) and np.array_equal(vr, vr2)
    ex = sim.main_msim_multi_device(comm, x.main_options(), sa)
    assert ex.n_histories == sim.shard_count(0, 1) and ex.n_launches == 3          # kernel, limbs, all-reduce
    comm.close()
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["srm1155", "synthetic10", "brute"])
def test_all_devices_is_bit_identical_to_one_device(which):
    n_dev = abi.lib().xmb_cuda_device_count()
    if which == "synthetic10":
        inp = synthetic_layers(n_photons=40000, n_int=8)
    else:
        inp = example("srm1155"); inp.n_photons_line = 4000
    opt = x.main_options(use_variance_reduction=0) if which == "brute" else x.main_options()
    sim, sa = _setup(inp)
    ch, br, vr = sim.main_msim(opt, sa)
    for n in sorted({1, 2, n_dev}):
        if n > n_dev:
            continue
        ch2, br2, vr2, ex = sim.main_msim_all_devices(n, opt, sa)
        assert ex.n_histories == sim.shard_count(0, 1), n
        assert np.array_equal(ch, ch2) and np.array_equal(vr, vr2) and np.array_equal(br, br2), n
    sim.close()


@pytest.mark.gpu
def test_cli_gpus_option_writes_the_same_xmso(tmp_path):
    """bin/xmimsim-b200 --gpus N: the XMSO file is byte-identical at every GPU count of the box."""
    n_dev = abi.lib().xmb_cuda_device_count()
    exe = os.path.join(ROOT, "bin", "xmimsim-b200")
    inp = example("srm1155")
    inp.n_photons_line = 2000
    outs = []
    for n in sorted({1, n_dev}):
        inp.outputfile = str(tmp_path / ("out_%d.xmso" % n))
        xmsi = str(tmp_path / ("in_%d.xmsi" % n))
        ci = x.CInput(inp)
        assert abi.lib().xmb_input_write_to_xml_file(C.byref(ci.input), xmsi.encode())
        r = subprocess.run([exe, "--surrogate-cross-sections", "--table-quality=0", "--gpus=%d" % n, "--disable-escape-peaks", xmsi],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(x.read_xmso(inp.outputfile))
    for o in outs[1:]:
        assert np.array_equal(o["conv"], outs[0]["conv"]), float(np.abs(o["conv"] - outs[0]["conv"]).max())
        assert np.array_equal(o["unconv"], outs[0]["unconv"]), float(np.abs(o["unconv"] - outs[0]["unconv"]).max())
        assert o["history"].keys() == outs[0]["history"].keys()
        for k, v in o["history"].items():
            assert v["counts"] == outs[0]["history"][k]["counts"], k
