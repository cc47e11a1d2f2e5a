"""ctypes binding of the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")
_lib = None


class OrcDerived(C.Structure):
    _fields_ = [("detector_radius", C.c_double), ("collimator_radius", C.c_double), ("collimator_height", C.c_double),
                ("half_apex", C.c_double), ("vertex", C.c_double * 3), ("collimator_present", C.c_int),
                ("ndo_new", C.c_double * 9), ("ndo_inv", C.c_double * 9), ("detector_solid_angle", C.c_double),
                ("n_sample_orientation_det", C.c_double * 3), ("n_layers", C.c_int),
                ("thickness_along_Z", C.POINTER(C.c_double)), ("Z_coord_begin", C.POINTER(C.c_double)),
                ("Z_coord_end", C.POINTER(C.c_double))]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle library missing: run `make oracle`")
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_poly_solve_quadratic.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)] * 2
        _lib.orc_poly_solve_quadratic.restype = C.c_int
        _lib.orc_init_input.argtypes = [C.c_void_p, C.POINTER(OrcDerived)]
        _lib.orc_init_input.restype = C.c_int
        _lib.orc_single_solid_angle.argtypes = [C.POINTER(OrcDerived), C.c_double, C.c_double, C.c_long, C.c_uint64,
                                                C.c_uint64, C.POINTER(C.c_long)]
        _lib.orc_single_solid_angle.restype = C.c_double
        _lib.orc_solid_angle_grid.argtypes = [C.POINTER(OrcDerived), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_uint64, C.c_void_p,
                                              C.c_void_p, C.c_int]
        _lib.orc_solid_angle_grid.restype = None
        _lib.orc_solid_angle_axes.argtypes = [C.c_void_p, C.POINTER(OrcDerived), C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        _lib.orc_solid_angle_axes.restype = C.c_int
        _lib.xmb_xrl_surrogate.restype = C.c_void_p
        _lib.orc_detector_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        _lib.orc_detector_correction.restype = C.c_double
        _lib.orc_detector_gaussian.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_detector_convolute_spectrum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_int, C.c_uint64]
        _lib.orc_detector_convolute_history.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_set_hits_per_single.argtypes = [C.c_long]
        _lib.orc_set_hits_per_single.restype = None
        _lib.orc_total_histories.argtypes = [C.c_void_p]
        _lib.orc_total_histories.restype = C.c_uint64
        _lib.orc_main_msim_range.argtypes = [C.c_void_p, C.POINTER(OrcDerived), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        _lib.orc_main_msim_range.restype = C.c_uint64
        _lib.orc_main_msim_shard.argtypes = [C.c_void_p, C.POINTER(OrcDerived), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_main_msim_shard.restype = C.c_uint64
        _lib.orc_main_msim_brute_range.argtypes = [C.c_void_p, C.POINTER(OrcDerived), C.c_void_p, C.c_void_p, C.c_uint64,
                                                   C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_main_msim_brute_range.restype = C.c_uint64
        _lib.orc_cubic_spline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
        _lib.orc_cubic_spline.restype = C.c_double
        _lib.orc_tube_ebel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_double] * 6 + [C.c_int, C.c_size_t] + [C.c_void_p] * 7
        _lib.orc_tube_ebel.restype = C.c_int
        _lib.orc_escape_ratios.argtypes = [C.c_void_p, C.POINTER(OrcDerived), C.c_void_p, C.c_uint64, C.c_long, C.c_int,
                                           C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _lib.orc_escape_ratios.restype = C.c_int
    return _lib


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def init_input(cinput_ptr):
    """orc_init_input on a ctypes xmb_input (modified in place: normals normalised)."""
    d = OrcDerived()
    if not lib().orc_init_input(C.cast(cinput_ptr, C.c_void_p), C.byref(d)):
        raise RuntimeError("orc_init_input failed")
    return d


def solid_angle_grid(d, r_vals, r_idx, theta_vals, theta_idx, full_n_r, hits_per_single, seed, n_threads=8):
    r = np.ascontiguousarray(r_vals, np.float64); t = np.ascontiguousarray(theta_vals, np.float64)
    ri = np.ascontiguousarray(r_idx, np.int32); ti = np.ascontiguousarray(theta_idx, np.int32)
    sa = np.zeros((t.size, r.size)); hits = np.zeros((t.size, r.size), np.int32)
    lib().orc_solid_angle_grid(C.byref(d), r.ctypes.data, ri.ctypes.data, r.size, t.ctypes.data, ti.ctypes.data, t.size,
                               full_n_r, hits_per_single, seed, sa.ctypes.data, hits.ctypes.data, n_threads)
    return sa, hits


def solid_angle_axes(cinput_ptr, d, n=1024):
    r = np.zeros(n); t = np.zeros(n)
    ok = lib().orc_solid_angle_axes(C.cast(cinput_ptr, C.c_void_p), C.byref(d), lib().xmb_xrl_surrogate(), r.ctypes.data,
                                    t.ctypes.data, n)
    if not ok:
        raise RuntimeError("orc_solid_angle_axes failed")
    return r, t


def main_msim_range(cinput_ptr, d, tables_ptr, options, sa_struct, seed, g_begin, g_end, n_int, nch, n_threads=8):
    """Oracle histories for global photon ids [g_begin, g_end).  Returns (channels[(n_int+1)][nch] cumulative,
    var_red[100][385][n_int] in the reference's exported C order, counters) -- RAW, without live_time."""
    ch = np.zeros((n_int + 1, nch))
    vr = np.zeros((n_int, 385, 100))
    cnt = np.zeros(2, np.uint64)
    lib().orc_main_msim_range(C.cast(cinput_ptr, C.c_void_p), C.byref(d), C.cast(tables_ptr, C.c_void_p),
                              C.cast(C.pointer(options), C.c_void_p), C.cast(C.pointer(sa_struct), C.c_void_p), seed,
                              g_begin, g_end, n_threads, ch.ctypes.data, vr.ctypes.data, cnt.ctypes.data)
    return ch, np.ascontiguousarray(vr.transpose(2, 1, 0)), cnt


def main_msim_shard(cinput_ptr, d, tables_ptr, options, sa_struct, seed, rank, n_ranks, n_int, nch, n_threads=8):
    """Oracle histories of the block-cyclic shard of `rank`.  Returns (channels, var_red, counters, n_histories)."""
    ch = np.zeros((n_int + 1, nch))
    vr = np.zeros((n_int, 385, 100))
    cnt = np.zeros(2, np.uint64)
    n = lib().orc_main_msim_shard(C.cast(cinput_ptr, C.c_void_p), C.byref(d), C.cast(tables_ptr, C.c_void_p),
                                  C.cast(C.pointer(options), C.c_void_p), C.cast(C.pointer(sa_struct), C.c_void_p), seed,
                                  rank, n_ranks, n_threads, ch.ctypes.data, vr.ctypes.data, cnt.ctypes.data)
    return ch, np.ascontiguousarray(vr.transpose(2, 1, 0)), cnt, int(n)


def main_msim_brute_range(cinput_ptr, d, tables_ptr, options, seed, g_begin, g_end, n_int, nch, n_threads=8):
    """Oracle brute-force histories for photon ids [g_begin, g_end).  Returns (channels[(n_int+1)][nch],
    brute_history[100][385][n_int] in the reference's exported C order, counters[hits, interactions, offspring])."""
    ch = np.zeros((n_int + 1, nch))
    br = np.zeros((n_int, 385, 100))
    cnt = np.zeros(3, np.uint64)
    lib().orc_main_msim_brute_range(C.cast(cinput_ptr, C.c_void_p), C.byref(d), C.cast(tables_ptr, C.c_void_p),
                                    C.cast(C.pointer(options), C.c_void_p), seed, g_begin, g_end, n_threads,
                                    ch.ctypes.data, br.ctypes.data, cnt.ctypes.data)
    return ch, np.ascontiguousarray(br.transpose(2, 1, 0)), cnt


def escape_ratios(cinput_ptr, d, tables_ptr, seed, n_energies, nZ, n_photons, n_out, out_min, out_delta, n_threads=8):
    """Oracle escape-ratio Monte Carlo on an escape-mode input.  Returns (fluo[nE][109][nZ], compton[n_out][nE])."""
    fluo = np.zeros((n_energies, 109, nZ))
    compt = np.zeros((n_out, n_energies))
    lib().orc_escape_ratios(C.cast(cinput_ptr, C.c_void_p), C.byref(d), C.cast(tables_ptr, C.c_void_p), seed, n_photons,
                            n_out, out_min, out_delta, n_threads, fluo.ctypes.data, compt.ctypes.data)
    return fluo, compt


def detector_convolute_spectrum(cinput_ptr, spectrum, options, escape_ratios=None, n_interactions=1, seed=1):
    """Oracle xmi_detector_convolute_spectrum: spectrum modified in place, returns conv."""
    spectrum = np.ascontiguousarray(spectrum, np.float64)
    conv = np.zeros_like(spectrum)
    er = C.cast(C.pointer(escape_ratios), C.c_void_p) if escape_ratios is not None else None
    lib().orc_detector_convolute_spectrum(C.cast(cinput_ptr, C.c_void_p), lib().xmb_xrl_surrogate(), spectrum.ctypes.data,
                                          conv.ctypes.data, C.cast(C.pointer(options), C.c_void_p), er, n_interactions, seed)
    return spectrum, conv


def detector_convolute_history(cinput_ptr, history):
    h = np.ascontiguousarray(history, np.float64).copy()
    lib().orc_detector_convolute_history(C.cast(cinput_ptr, C.c_void_p), lib().xmb_xrl_surrogate(), h.ctypes.data)
    return h
