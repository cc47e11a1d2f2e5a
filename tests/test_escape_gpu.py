"""GPU parity of the escape-ratio Monte Carlo (xmi_escape_ratios_calculation) against the CPU oracle on the
same escape-mode input, host tables and Philox streams."""
import ctypes as C

import numpy as np
import pytest

import xmimsim_b200 as x
from inputs import example
from xmimsim_b200.xmsi import LayerD
from test_escape_cpu import _setup, _oracle

pytestmark = pytest.mark.gpu

# Both sides are fp64 with the same tables and random numbers; the oracle sums doubles, the GPU exact 2^-40
# fixed-point integers (<= 2^-41 per tally, relative 1e-12 of a ratio of ~1e-3 after >= 10 tallies).  A history whose
# escape decision flips on a last-ulp difference changes a ratio by 1/n_photons; none is expected at these sizes.
RTOL = 1e-9


def test_escape_ratios_match_oracle_si_crystal():
    inp, sim, ero, ein, eh = _setup(n_E=24, n_photons=30000, e_min=1.2, e_delta=1.7)     # 1.2 .. 40.3 keV
    fluo_o, compt_o = _oracle(sim, ero, ein, eh, seed=5, n_threads=16)
    er = sim.escape_ratios_run(ein, eh, ero, seed=5)
    Z, fluo, e_in, compt, e_out = sim.escape_ratios_arrays(er)
    sim.escape_ratios_free(er)
    assert list(Z) == [14] and fluo.shape == fluo_o.shape and compt.shape == compt_o.shape
    assert np.allclose(e_in, 1.2 + 1.7 * np.arange(24)) and np.allclose(e_out, 0.1 + 0.1 * np.arange(300))
    assert fluo_o.sum() > 0 and compt_o.sum() > 0
    assert np.array_equal(fluo > 0, fluo_o > 0) and np.array_equal(compt > 0, compt_o > 0)   # same photons escaped into the same bins
    assert np.abs(fluo - fluo_o).max() <= RTOL * fluo_o.max()
    assert np.abs(compt - compt_o).max() <= RTOL * compt_o.max()
    # deterministic: a second run gives the same bits
    er2 = sim.escape_ratios_run(ein, eh, ero, seed=5)
    fluo2 = sim.escape_ratios_arrays(er2)[1]
    sim.escape_ratios_free(er2)
    assert np.array_equal(fluo, fluo2)
    sim.L.xmb_free_hdf5_F(C.byref(eh)); sim.L.xmb_free_input_F(C.byref(ein)); sim.close()


def test_escape_ratios_two_layer_compound_crystal():
    """Two crystal layers with several elements: layer walk of the analogue step in both directions, element
    indexing of fluo_escape_ratios(element, line, energy)."""
    inp = example("srm1155")
    inp.crystal_layers = [LayerD([31, 33], [0.48, 0.52], 5.32, 0.002), LayerD([14], [1.0], 2.33, 0.05)]
    sim = x.Simulation(inp)
    ero = sim.escape_ratios_options(n_input_energies=10, n_photons=30000, input_energy_min=5.0, input_energy_delta=3.0,
                                    n_compton_output_energies=400)
    ein, eh = sim.escape_ratios_handles(ero)
    fluo_o, compt_o = _oracle(sim, ero, ein, eh, seed=77, n_threads=16)
    er = sim.escape_ratios_run(ein, eh, ero, seed=77)
    Z, fluo, e_in, compt, e_out = sim.escape_ratios_arrays(er)
    sim.escape_ratios_free(er)
    assert list(Z) == [14, 31, 33]
    assert fluo_o[:, :, 1].sum() > 0 and fluo_o[:, :, 2].sum() > 0          # Ga and As K/L escape
    assert np.array_equal(fluo > 0, fluo_o > 0)
    assert np.abs(fluo - fluo_o).max() <= RTOL * fluo_o.max()
    assert np.abs(compt - compt_o).max() <= RTOL * max(compt_o.max(), 1e-300)
    sim.L.xmb_free_hdf5_F(C.byref(eh)); sim.L.xmb_free_input_F(C.byref(ein)); sim.close()


def test_calculated_ratios_drive_the_detector_response():
    """The whole chain of bin/xmimsim.c:470-526: ratios from the Monte Carlo -> escape peaks in the convoluted
    spectrum at E - E(Si-K)."""
    inp = example("srm1155")
    sim = x.Simulation(inp)
    ero = sim.escape_ratios_options(n_input_energies=200, n_photons=20000, input_energy_min=1.0, input_energy_delta=0.1,
                                    n_compton_output_energies=250)
    er = sim.escape_ratios_calculation(ero)
    Z, fluo, e_in, compt, e_out = sim.escape_ratios_arrays(er)
    assert list(Z) == [14] and fluo.shape == (200, 109, 1)
    nch = inp.nchannels
    spec = np.zeros(nch)
    e0 = 8.0
    ch0 = int((e0 - inp.zero) / inp.gain)
    spec[ch0] = 1e7
    opts = x.main_options(use_escape_peaks=1, use_sum_peaks=0, use_poisson=0)
    work = spec.copy()
    conv = sim.detector_convolute_spectrum(work, opts, er.contents)
    opts0 = x.main_options(use_escape_peaks=0, use_sum_peaks=0, use_poisson=0)
    work0 = spec.copy()
    conv0 = sim.detector_convolute_spectrum(work0, opts0, None)
    T = sim.tables
    # strongest Si K line in the ratios at 8 keV
    i8 = int(round((e0 - 1.0) / 0.1))
    line = int(np.argmax(fluo[i8, :, 0])) + 1
    assert fluo[i8, line - 1, 0] > 1e-4
    diff = conv - conv0
    assert diff[ch0 - 5: ch0 + 5].sum() < 0                                   # main peak loses what escapes
    lo = int((e0 - 1.9 - inp.zero) / inp.gain); hi = int((e0 - 1.6 - inp.zero) / inp.gain)
    assert diff[lo:hi].sum() > 0.5 * fluo[i8, :29, 0].sum() * conv0.sum()     # and it shows up ~1.74 keV below
    sim.escape_ratios_free(er)
    sim.close()
