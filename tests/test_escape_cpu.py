"""CPU tests of the escape-ratio Monte Carlo: the escape-mode input the library derives
(src/xmi_detector.c:91-141) and physical properties of the oracle's ratios
(src/xmi_main.F90:5473-5801).  No golden vectors exist for this path in the reference tree (the
ratios live in a user-side HDF5 cache), so the oracle is checked against closed-form limits."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x
from inputs import example


def _setup(n_E=8, n_photons=4000, e_min=1.5, e_delta=0.5):
    inp = example("srm1155")
    sim = x.Simulation(inp)
    ero = sim.escape_ratios_options(n_input_energies=n_E, n_photons=n_photons, input_energy_min=e_min,
                                    input_energy_delta=e_delta, n_compton_output_energies=300)
    ein, eh = sim.escape_ratios_handles(ero)
    return inp, sim, ero, ein, eh


def _oracle(sim, ero, ein, eh, seed=5, n_threads=8):
    L = sim.L
    cin = L.xmb_input_F2C(ein)
    od = orc.init_input(cin)
    T = L.xmb_get_tables(eh)
    return orc.escape_ratios(cin, od, T, seed, ero.n_input_energies, T.contents.nZ, ero.n_photons,
                             ero.n_compton_output_energies, ero.compton_output_energy_min, ero.compton_output_energy_delta,
                             n_threads)


def test_escape_mode_input_follows_the_reference_overrides():
    inp, sim, ero, ein, eh = _setup()
    cin = sim.L.xmb_input_F2C(ein).contents
    comp = cin.composition.contents
    det = sim.c_input().contents.detector.contents
    assert comp.n_layers == det.n_crystal_layers and comp.reference_layer == 1
    for k in range(comp.n_layers):
        a, b = comp.layers[k], det.crystal_layers[k]
        assert a.n_elements == b.n_elements and a.density == b.density and a.thickness == b.thickness
        assert [a.Z[i] for i in range(a.n_elements)] == [b.Z[i] for i in range(b.n_elements)]
    g = cin.geometry.contents
    assert (g.d_sample_source, g.d_source_slit, g.slit_size_x, g.slit_size_y) == (1.0, 1.0, 0.0001, 0.0001)
    assert list(g.n_sample_orientation) == [0.0, 0.0, 1.0]
    assert cin.general.contents.n_interactions_trajectory == 1
    exc = cin.excitation.contents
    assert exc.n_discrete == ero.n_input_energies and exc.n_continuous == 0
    assert [exc.discrete[i].energy for i in range(3)] == [1.5, 2.0, 2.5]
    assert cin.absorbers.contents.n_exc_layers == 0
    od = orc.init_input(sim.L.xmb_input_F2C(ein))
    assert od.Z_coord_begin[0] == 1.0                                  # xmi_init_input_escape_ratios, :1716-1721
    assert abs(od.Z_coord_end[0] - (1.0 + comp.layers[0].thickness)) < 1e-15
    # every input energy is an exact node of the table bundle
    T = sim.L.xmb_get_tables(eh).contents
    nodes = np.ctypeslib.as_array(T.node_E, shape=(T.n_nodes,))
    for i in range(ero.n_input_energies):
        assert (nodes == exc.discrete[i].energy).any()
    sim.L.xmb_free_hdf5_F(C.byref(eh)); sim.L.xmb_free_input_F(C.byref(ein)); sim.close()


def test_oracle_ratios_have_the_physical_limits():
    inp, sim, ero, ein, eh = _setup(n_E=6, n_photons=20000, e_min=1.5, e_delta=1.5)   # 1.5, 3.0 ... 9.0 keV on Si
    fluo, compt = _oracle(sim, ero, ein, eh)
    T = sim.L.xmb_get_tables(eh).contents
    assert T.nZ == 1 and T.Z[0] == 14
    edge_K = T.edge_energy[0]
    assert 1.5 < edge_K < 3.0
    assert np.all(fluo >= 0) and np.all(compt >= 0)
    assert fluo[0].sum() == 0.0                                          # below the K edge: no K escape (L lines < 1 keV are cut)
    k_lines = fluo[1:, :29, 0].sum(axis=1)                               # KL1..KP5
    assert np.all(k_lines > 0)
    yield_K = T.fluor_yield[0]
    assert np.all(k_lines < yield_K)                                     # escape needs a K vacancy AND a radiative decay AND no re-absorption
    # deeper first interactions (higher energy) leave the crystal less often
    assert k_lines[0] > k_lines[-1]
    # Compton escape: a scattered photon is softer than the incident one, never harder
    e_out = ero.compton_output_energy_min + ero.compton_output_energy_delta * np.arange(ero.n_compton_output_energies)
    for i in range(ero.n_input_energies):
        e_in = ero.input_energy_min + i * ero.input_energy_delta
        assert compt[e_out > e_in + 1e-9, i].sum() == 0.0
    assert compt.sum(axis=0).max() < 0.05                                # Si below 10 keV is photo-absorption dominated
    # reproducible: same seed -> same bits, other seed -> statistically compatible
    fluo2, _ = _oracle(sim, ero, ein, eh, n_threads=3)
    assert np.array_equal(fluo, fluo2)
    fluo3, _ = _oracle(sim, ero, ein, eh, seed=6)
    k3 = fluo3[1:, :29, 0].sum(axis=1)
    assert np.all(np.abs(k3 - k_lines) < 6 * np.sqrt(k_lines * yield_K / ero.n_photons) + 1e-4)
    sim.L.xmb_free_hdf5_F(C.byref(eh)); sim.L.xmb_free_input_F(C.byref(ein)); sim.close()
