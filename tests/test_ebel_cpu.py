"""The Ebel tube-spectrum generator (xmi_tube_ebel, src/xmi_ebel.F90:114-521): product (C++) against the oracle's
independent restatement, the spline against closed forms, and the model's qualitative laws."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from xmimsim_b200 import abi
from xmimsim_b200.engine import _c_layer


def _oracle(anode, V, current, ae, ax, dE, omega, window=None, filt=None, transmission=0, eff=None):
    keep = []
    la = _c_layer(anode, keep)
    lw = _c_layer(window, keep) if window else None
    lf = _c_layer(filt, keep) if filt else None
    n = int((V - 1.0) / dE) + 3
    cE, cI, dEn, dI = np.zeros(n), np.zeros(n), np.zeros(113), np.zeros(113)
    nd = C.c_int()
    ne, pe, pv = 0, None, None
    if eff is not None:
        ee, ev = [np.ascontiguousarray(a, np.float64) for a in eff]
        ne, pe, pv = ee.size, ee.ctypes.data, ev.ctypes.data
    nc = orc.lib().orc_tube_ebel(orc.lib().xmb_xrl_surrogate(), C.addressof(la), C.addressof(lw) if lw else None,
                                 C.addressof(lf) if lf else None, V, current, ae, ax, dE, omega, transmission, ne, pe, pv,
                                 cE.ctypes.data, cI.ctypes.data, C.addressof(nd), dEn.ctypes.data, dI.ctypes.data)
    return cE[:nc], cI[:nc], dEn[:nd.value], dI[:nd.value]


AG = x.LayerD([47], [1.0], 10.5, 0.0002)
BE = x.LayerD([4], [1.0], 1.85, 0.0125)
AL = x.LayerD([13], [1.0], 2.7, 0.01)


def test_baseline_config5_grid_and_oracle_agreement():
    """Ag anode, 40 kV, 1 mA, 60/60 degrees, no window, 1e-4 sr, dE = 0.039 keV."""
    cont, disc = x.tube_ebel(AG, 40.0, 1.0, 60.0, 60.0, 0.039, 1e-4)
    cE, cI, dE, dI = _oracle(AG, 40.0, 1.0, 60.0, 60.0, 0.039, 1e-4)
    # floor(39/0.039)+1 = 1001 grid points, plus the tube voltage itself because 40/0.039 is not an integer
    # (src/xmi_ebel.F90:203-217): 1000 intervals and a zero-width one the history driver skips
    assert cont.shape[0] == cE.size == 1002
    assert cont[0, 0] == 1.0 and cont[-1, 0] == 40.0 and abs(cont[1, 0] - 1.039) < 1e-12 and abs(cont[-2, 0] - 40.0) < 1e-9
    assert np.array_equal(cont[:, 0], cE) and np.allclose(cont[:, 1], cI, rtol=1e-12, atol=0)
    assert np.array_equal(cont[:, 1], cont[:, 2])                              # unpolarised
    assert disc.shape[0] == dE.size and np.array_equal(disc[:, 0], dE) and np.allclose(disc[:, 1], dI, rtol=1e-12, atol=0)
    # laws of the model: bremsstrahlung vanishes at the Duane-Hunt limit, K lines only above the K edge (25.5 keV)
    assert cont[-1, 1] == 0.0 and cont[-2, 1] == 0.0 and np.all(cont[:-2, 1] > 0) and cont[-6, 1] < 0.02 * cont[:, 1].max()
    assert (disc[:, 0] > 20).sum() >= 2 and (disc[:, 0] < 4.5).sum() >= 3      # Ag K and L series
    _, disc20 = x.tube_ebel(AG, 20.0, 1.0, 60.0, 60.0, 0.039, 1e-4)
    assert (disc20[:, 0] > 20).sum() == 0 and disc20.shape[0] > 0
    # linear in current and solid angle
    cont2, disc2 = x.tube_ebel(AG, 40.0, 3.0, 60.0, 60.0, 0.039, 2e-4)
    assert np.allclose(cont2[:, 1], 6 * cont[:, 1], rtol=1e-12) and np.allclose(disc2[:, 1], 6 * disc[:, 1], rtol=1e-12)


def test_window_filter_transmission_and_efficiency():
    eff = (np.array([1.0, 5.0, 10.0, 20.0, 40.0]), np.array([0.2, 0.6, 0.9, 0.95, 0.8]))
    for tr in (0, 1):
        cont, disc = x.tube_ebel(AG, 35.0, 0.5, 45.0, 30.0, 0.25, 1e-3, window=BE, filt=AL, transmission=tr,
                                 eff_energies=eff[0], efficiencies=eff[1])
        cE, cI, dE, dI = _oracle(AG, 35.0, 0.5, 45.0, 30.0, 0.25, 1e-3, BE, AL, tr, eff)
        assert np.array_equal(cont[:, 0], cE) and np.allclose(cont[:, 1], cI, rtol=1e-11, atol=0)
        assert np.allclose(disc[:, 1], dI, rtol=1e-11, atol=0)
    bare, _ = x.tube_ebel(AG, 35.0, 0.5, 45.0, 30.0, 0.25, 1e-3)
    filt, _ = x.tube_ebel(AG, 35.0, 0.5, 45.0, 30.0, 0.25, 1e-3, window=BE, filt=AL)
    ratio = filt[1:-1, 1] / bare[1:-1, 1]
    assert np.all(ratio < 1) and ratio[2] < ratio[-2]                          # absorbers harden the spectrum


def test_natural_spline():
    xs = np.array([0.0, 1.0, 2.5, 3.0, 4.0, 6.0]); ys = 2.0 - 0.5 * xs       # a straight line is reproduced exactly
    for v in (-1.0, 0.0, 0.3, 2.5, 5.9, 7.0):
        assert abs(orc.lib().orc_cubic_spline(xs.ctypes.data, ys.ctypes.data, xs.size, v) - (2.0 - 0.5 * v)) < 1e-12
    ys = np.sin(xs)
    for i, v in enumerate(xs[:-1]):                                            # interpolates its knots
        assert abs(orc.lib().orc_cubic_spline(xs.ctypes.data, ys.ctypes.data, xs.size, v) - ys[i]) < 1e-12
    # the product's spline (through the efficiency curve): ratio of two runs isolates it
    a, _ = x.tube_ebel(AG, 30.0, 1.0, 60.0, 60.0, 0.5, 1e-4)
    b, _ = x.tube_ebel(AG, 30.0, 1.0, 60.0, 60.0, 0.5, 1e-4, eff_energies=xs * 6 + 1, efficiencies=ys + 2)
    xs6 = xs * 6 + 1; y2 = ys + 2
    for i in (0, 7, 20, 40):
        ref = orc.lib().orc_cubic_spline(xs6.ctypes.data, y2.ctypes.data, xs.size, a[i, 0])
        assert abs(b[i, 1] / a[i, 1] - ref) < 1e-12


def test_ebel_spectrum_drives_the_engine_input():
    """The generated spectrum is a valid excitation for the table builder and the history driver's bookkeeping."""
    cont, disc = x.tube_ebel(AG, 40.0, 1.0, 60.0, 60.0, 0.39, 1e-4)
    from inputs import example
    inp = example("srm1155")
    inp.discrete = [x.DiscreteD(e, h, v) for e, h, v in disc]
    inp.continuous = [x.ContinuousD(e, h, v) for e, h, v in cont]
    inp.n_photons_interval = 10; inp.n_photons_line = 20
    sim = x.Simulation(inp)
    n_valid = sum(1 for i in range(len(cont) - 1) if (cont[i, 1] + cont[i + 1, 1]) > 0)
    assert sim.L.xmb_msim_total_histories(sim.inputF) == n_valid * 10 + len(disc) * 20
    sim.close()


def test_bad_configurations_are_refused():
    """tests/test-ebel.c of the reference: good configurations return 1, a bad one 0."""
    L = abi.lib()
    keep = []
    la = _c_layer(AG, keep)
    out = C.POINTER(abi.Excitation)()
    good = (40.0, 1.0, 60.0, 60.0, 0.1, 1e-4)
    assert L.xmb_tube_ebel(None, C.byref(la), None, None, *good, 0, 0, None, None, C.byref(out)) == 1
    L.xmb_free_excitation(C.byref(out))
    assert L.xmb_tube_ebel(None, C.byref(la), None, None, *good, 1, 0, None, None, C.byref(out)) == 1      # transmission tube
    L.xmb_free_excitation(C.byref(out))
    assert L.xmb_tube_ebel(None, None, None, None, *good, 0, 0, None, None, C.byref(out)) == 0             # no anode
    assert L.xmb_tube_ebel(None, C.byref(la), None, None, 40.0, 1.0, 60.0, 60.0, 0.0, 1e-4, 0, 0, None, None, C.byref(out)) == 0   # zero step
    assert L.xmb_tube_ebel(None, C.byref(la), None, None, 0.5, 1.0, 60.0, 60.0, 0.1, 1e-4, 0, 0, None, None, C.byref(out)) == 0    # below 1 keV
    assert "bad arguments" in abi.last_error()
