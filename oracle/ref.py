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
