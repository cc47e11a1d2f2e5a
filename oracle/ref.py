"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libxmi_ref.so, the reference's own sources
compiled by oracle/build_ref.sh (src/xmi_kernels.cl through an OpenCL-C shim, src/xmi_spline.c).  Only tests/ load it."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libxmi_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libxmi_ref.so missing: run oracle/build_ref.sh where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_solid_angle_calculation_cl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                        C.c_float, C.c_float, C.c_float, C.c_int]
        _lib.ref_solid_angle_calculation_cl.restype = C.c_int
        _lib.ref_cubic_spline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
        _lib.ref_cubic_spline.restype = C.c_double
        if hasattr(_lib, "ref_output_raw2struct_rows"):
            _lib.ref_output_raw2struct_rows.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
            _lib.ref_output_raw2struct_rows.restype = C.c_int
        if hasattr(_lib, "ref_input_validate"):
            _lib.ref_input_validate.argtypes = [C.c_void_p]
            _lib.ref_input_validate.restype = C.c_int
        for name in ("ref_check_solid_angle_match", "ref_check_escape_ratios_match"):
            if hasattr(_lib, name):
                getattr(_lib, name).argtypes = [C.c_void_p, C.c_void_p]
                getattr(_lib, name).restype = C.c_int
    return _lib


def solid_angle_grid_cl(r_vals, theta_vals, collimator_present, detector_radius, collimator_radius, collimator_height,
                        hits_per_single):
    """The reference's OpenCL kernel xmi_solid_angle_calculation (src/xmi_kernels.cl:219-452) on the host, fp32,
    Threefry keyed by the grid indices as the reference does.  Returns solid_angles[theta][r] (float32)."""
    r = np.ascontiguousarray(r_vals, np.float32)
    t = np.ascontiguousarray(theta_vals, np.float32)
    out = np.zeros((t.size, r.size), np.float32)
    lib().ref_solid_angle_calculation_cl(r.ctypes.data, r.size, t.ctypes.data, t.size, out.ctypes.data,
                                         int(collimator_present), float(detector_radius), float(collimator_radius),
                                         float(collimator_height), int(hits_per_single))
    return out


def cubic_spline(x, y, v):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    return lib().ref_cubic_spline(x.ctypes.data, y.ctypes.data, x.size, float(v))


def output_raw2struct_rows(cinput_ptr, brute_history, var_red_history, channels_conv_rows, channels_unconv, use_zero_interactions,
                           which, cap=20000):
    """The reference's xmi_output_raw2struct (src/xmi_data_structs.c:1371-1519) on the raw arrays of xmi_main_msim.
    Returns (rows, unconv_checksum): rows = [(Z, line type, element total, line total, interaction number, counts)] in
    the order of the reference's history list (which = 0 brute force, 1 variance reduction)."""
    Z = np.zeros(cap, np.int32); lt = np.zeros((cap, 10), np.uint8); et = np.zeros(cap); ltot = np.zeros(cap)
    ino = np.zeros(cap, np.int32); cnt = np.zeros(cap); chk = np.zeros(1)
    br = np.ascontiguousarray(brute_history, np.float64)
    vr = None if var_red_history is None else np.ascontiguousarray(var_red_history, np.float64)
    un = np.ascontiguousarray(channels_unconv, np.float64)
    n = lib().ref_output_raw2struct_rows(C.cast(cinput_ptr, C.c_void_p), br.ctypes.data, None if vr is None else vr.ctypes.data,
                                         C.cast(channels_conv_rows, C.c_void_p), un.ctypes.data, int(use_zero_interactions), int(which),
                                         cap, Z.ctypes.data, lt.ctypes.data, et.ctypes.data, ltot.ctypes.data, ino.ctypes.data,
                                         cnt.ctypes.data, chk.ctypes.data)
    if n < 0 or n > cap:
        raise RuntimeError("ref_output_raw2struct_rows: %d rows" % n)
    rows = [(int(Z[i]), bytes(lt[i]).split(b"\0")[0].decode(), float(et[i]), float(ltot[i]), int(ino[i]), float(cnt[i]))
            for i in range(n)]
    return rows, float(chk[0])


def check_solid_angle_match(cached_ptr, fresh_ptr):
    """The reference's xmi_check_solid_angle_match (src/xmi_solid_angle.c:420-673): 1 when the cached grid of input A serves
    input B.  Normalises the orientation vectors of both inputs in place -- pass throw-away copies."""
    return int(lib().ref_check_solid_angle_match(C.cast(cached_ptr, C.c_void_p), C.cast(fresh_ptr, C.c_void_p)))


def check_escape_ratios_match(cached_ptr, fresh_ptr):
    """The reference's xmi_check_escape_ratios_match (src/xmi_detector.c:143-172)."""
    return int(lib().ref_check_escape_ratios_match(C.cast(cached_ptr, C.c_void_p), C.cast(fresh_ptr, C.c_void_p)))


def input_validate(cinput_ptr):
    """The reference's xmi_input_validate (src/xmi_data_structs.c:899-1255): OR of XmiInputFlags, 0 = valid."""
    return int(lib().ref_input_validate(C.cast(cinput_ptr, C.c_void_p)))


def layout_checks():
    """Number of compile-time field checks (offset + size, include/xmimsim_b200.h vs the reference's headers) that
    oracle/ref_shim/ref_layout.c held when oracle/_ref was built; the build fails on a mismatch."""
    f = lib().ref_layout_checks
    f.restype = C.c_int
    return int(f())
