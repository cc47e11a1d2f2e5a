"""CPU tests of the detector-response oracle: the reference's own pile-up criterion on the golden spectrum
and closed-form properties of the Gaussian response.  (The golden conv/unconv pairs in examples/*.xmso cannot
pin the response numerically: they were written by an unknown version with unknown escape ratios and need
xraylib's Be/Si cross sections -- DESIGN.md 'parity status'.)"""
import ctypes as C
import math
import os

import numpy as np

import orc
import xmimsim_b200 as x
from inputs import example

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ci(name="srm1155"):
    inp = example(name)
    return inp, x.CInput(inp)


def test_gaussian_response_is_normalised_and_has_the_documented_fwhm():
    inp, ci = _ci()
    nch = inp.nchannels
    spec = np.zeros(nch)
    i0 = 800
    spec[i0] = 1e6
    conv = np.zeros(nch)
    orc.lib().orc_detector_gaussian(C.cast(C.pointer(ci.input), C.c_void_p), spec.ctypes.data, conv.ctypes.data)
    assert abs(conv.sum() / 1e6 - 1.0) < 1e-12                     # rows are normalised (src/xmi_detector_f.F90:548-555)
    e0 = inp.zero + inp.gain * i0
    fwhm = math.sqrt(inp.noise ** 2 + 2.3548 ** 2 * 3.85 * inp.fano / 1000.0 * e0)   # :403-404, :505
    half = conv.max() / 2
    above = np.where(conv >= half)[0]
    width = (above[-1] - above[0] + 1) * inp.gain
    assert abs(width - fwhm) <= 2 * inp.gain
    assert conv[: i0 - 200].max() < 1e-3 * conv.max() and conv[i0 + 101:].sum() == 0.0   # tail below, hard cut 100 channels above
    # channels below 1 keV are dropped (:504)
    spec[:] = 0
    spec[10] = 5.0
    orc.lib().orc_detector_gaussian(C.cast(C.pointer(ci.input), C.c_void_p), spec.ctypes.data, conv.ctypes.data)
    assert conv.sum() == 0.0


def test_reference_pile_up_criterion_on_golden_spectrum():
    """tests/test-pile-up.c:17-64 of the reference: srm1155.xmso order-4 unconvoluted spectrum, escape peaks off,
    convoluted without and then with pile-up on the SAME array; channel 1077 (the Fe-Ka sum peak) must grow > 100x."""
    inp, ci = _ci()
    g = np.load(os.path.join(GOLDEN, "srm1155_xmso.npz"))
    channels = np.ascontiguousarray(g["unconv"][3]).copy()
    o = x.main_options(use_sum_peaks=0, use_escape_peaks=0, use_default_seeds=1)
    channels, without = orc.detector_convolute_spectrum(C.pointer(ci.input), channels, o, None, 4)
    o.use_sum_peaks = 1
    channels, with_pu = orc.detector_convolute_spectrum(C.pointer(ci.input), channels, o, None, 4)
    assert with_pu[1077] / without[1077] > 100.0
    assert abs(with_pu.sum() / without.sum() - 1.0) < 0.05           # pulses are conserved up to merged groups


def test_history_correction_uses_line_energies():
    inp, ci = _ci()
    n_int = inp.n_interactions_trajectory
    h = np.zeros((100, 385, n_int))
    h[25, 2, 0] = 1000.0      # Fe-KL3
    h[25, 383, 0] = 7.0       # Rayleigh slot: untouched (loop covers lines 1..383 only, :309)
    out = orc.detector_convolute_history(C.pointer(ci.input), h)
    L = orc.lib()
    xrl = L.xmb_xrl_surrogate()
    L.xmb_xrl_surrogate.restype = C.c_void_p
    corr = out[25, 2, 0] / 1000.0
    assert 0.9 < corr < 1.0 and out[25, 383, 0] == 7.0
