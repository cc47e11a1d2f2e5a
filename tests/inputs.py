"""Authored inputs for the parity tests (the shipped examples are copied to tests/golden/)."""
import copy
import os

import xmimsim_b200 as x

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def example(name):
    return x.read_xmsi(os.path.join(GOLDEN, name + ".xmsi"))


def no_collimator(inp):
    d = copy.deepcopy(inp)
    d.collimator_height = 0.0
    d.collimator_diameter = 0.0
    return d


def cylindrical_collimator(inp):
    """collimator radius within 1e-6 of the detector radius -> the reference's cylindrical branch
    (src/xmi_solid_angle_f.F90:488-519) while still passing xmi_init_input's conical check."""
    import math
    d = copy.deepcopy(inp)
    d.collimator_height = 0.5
    d.collimator_diameter = 2.0 * (math.sqrt(d.area_detector / math.pi) - 5e-7)
    return d


def caso4():
    """The reference's only physics KAT set-up (tests/libxmimsim-test.c:182-285): CaSO4 1 cm, 1 g/cm3,
    10 keV, default geometry with the detector at y = 10 cm, one interaction."""
    return x.InputD(
        n_photons_line=100000, n_interactions_trajectory=1,
        layers=[x.LayerD([7, 8, 18], [0.7, 0.29, 0.01], 0.001205, 5.0),   # air gap, as the default input has
                x.LayerD([8, 16, 20], [0.470095, 0.235534, 0.294371], 1.0, 1.0)],
        reference_layer=2, d_sample_source=100.0, n_sample_orientation=[0, 1, 1],
        p_detector_window=[0, -1, 100], n_detector_orientation=[0, 1, 0], area_detector=0.3,
        collimator_height=0.0, collimator_diameter=0.0, d_source_slit=100.0, slit_size_x=0.001, slit_size_y=0.001,
        discrete=[x.DiscreteD(10.0, 1e9, 1e9)],
        det_layers=[x.LayerD([4], [1.0], 1.85, 0.002)], detector_type=0, live_time=1.0, pulse_width=1e-5,
        gain=0.02, zero=0.0, fano=0.12, noise=0.1, nchannels=2048,
        crystal_layers=[x.LayerD([14], [1.0], 2.33, 0.5)])


from xmimsim_b200.workloads import synthetic_layers, ebel_like   # noqa: E402,F401  (BASELINE configs[3] / configs[4])


def close_detector(n_photons=200000, n_int=2):
    """A geometry in which analogue (brute-force) photons actually reach the detector: 3 cm2 window 2 cm from a
    45-degree steel-like slab, no collimator, one 20 keV line -- solid angle ~ 5 % of 4 pi."""
    return x.InputD(
        n_photons_line=n_photons, n_interactions_trajectory=n_int,
        layers=[x.LayerD([24, 26, 28], [0.18, 0.72, 0.10], 7.9, 0.05)],
        reference_layer=1, d_sample_source=100.0, n_sample_orientation=[0, 1, 1],
        p_detector_window=[0, -2, 100], n_detector_orientation=[0, 1, 0], area_detector=3.0,
        collimator_height=0.0, collimator_diameter=0.0, d_source_slit=100.0, slit_size_x=0.001, slit_size_y=0.001,
        discrete=[x.DiscreteD(20.0, 1e9, 1e9)],
        det_layers=[x.LayerD([4], [1.0], 1.85, 0.002)], detector_type=0, live_time=1.0, pulse_width=1e-5,
        gain=0.02, zero=0.0, fano=0.12, noise=0.1, nchannels=2048,
        crystal_layers=[x.LayerD([14], [1.0], 2.33, 0.5)])
