"""The reference-named entry points called from C the way the reference's host calls them (tests/c/plugin_harness.c):
xmi_solid_angle_calculation_cl and xmi_detector_convolute_all_custom resolved with dlsym from the plugin file
(src/xmi_solid_angle.c:121-160, bin/xmimsim.c:501-526), xmi_main_msim (include/xmi_main.h:29, bin/xmimsim.c:361) from
libxmimsim-b200-interpose.so, all with an opaque host handle that only the host's own xmi_input_F2C can read.  The
harness compares every output bit for bit with the xmb_* entry points and exits 0 on success."""
import os
import subprocess

import pytest

from xmimsim_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "xmimsim_b200", "lib")
HARNESS = os.path.join(LIBDIR, "plugin_harness")
INTERPOSE = os.path.join(LIBDIR, "libxmimsim-b200-interpose.so")


def test_interpose_file_exports_only_the_reference_name():
    out = subprocess.run(["nm", "-D", "--defined-only", INTERPOSE], capture_output=True, text=True).stdout
    syms = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert syms == ["xmi_main_msim"], syms


@pytest.mark.gpu
@pytest.mark.parametrize("example,photons", [("srm1155", "400"), ("srm1412", "300")])
def test_reference_named_symbols_called_from_c(example, photons):
    xmsi = os.path.join(ROOT, "tests", "golden", example + ".xmsi")
    r = subprocess.run([HARNESS, abi.LIB_PATH, INTERPOSE, xmsi, photons], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "harness OK" in r.stdout and "FAIL" not in r.stdout
    for what in ("xmi_solid_angle_calculation_cl returned a grid", "solid angles bit-identical", "channels bit-identical to xmb_main_msim",
                 "var_red_history bit-identical", "xmi_detector_convolute_all_custom bit-identical to xmb_detector_convolute_all"):
        assert "ok: " + what in r.stdout, what


@pytest.mark.gpu
def test_plugin_refuses_the_stand_in_unless_asked(tmp_path):
    """Without a registered provider and without xraylib the plugin must not silently compute with stand-in physics: the
    solid-angle symbol returns 0 (= the host falls through to its next backend, src/xmi_solid_angle.c:139-160)."""
    code = r'''
import ctypes as C, os, sys
sys.path.insert(0, %r)
import xmimsim_b200 as x
from xmimsim_b200 import abi
L = abi.lib()
inp = x.read_xmsi(%r)
sim = x.Simulation(inp, quality=0)
sa = C.POINTER(abi.SolidAngle)()
opt = x.main_options()
rv = L.xmi_solid_angle_calculation_cl(sim.inputF, C.byref(sa), None, C.byref(opt))
print("rv", rv, bool(sa))
''' % (ROOT, os.path.join(ROOT, "tests", "golden", "srm1155.xmsi"))
    env = dict(os.environ)
    env.pop("XMB_ALLOW_SURROGATE", None)
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert "rv 0 False" in r.stdout, r.stdout + r.stderr
    assert "no cross-section provider" in r.stderr
    env["XMB_ALLOW_SURROGATE"] = "1"
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert "rv 1 True" in r.stdout and "NOT physics-grade" in r.stderr, r.stdout + r.stderr
