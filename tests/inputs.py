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


def synthetic_layers(n_photons=30000, n_int=8, n_lines=1):
    """BASELINE config 4: 10 parallel layers, 3-6 elements each from a Z = 8..82 pool, Dirichlet(1)
    weights, rho ~ U(1,10), thickness ~ logU(1e-4, 1e-1), numpy default_rng(20260101); geometry/detector as
    srm1155; one 28 keV line (or n_lines lines on 20..40 keV)."""
    import numpy as np
    rng = np.random.default_rng(20260101)
    pool = [8, 13, 14, 20, 22, 26, 29, 30, 38, 42, 47, 50, 56, 74, 79, 82]
    base = example("srm1155")
    layers = []
    for _ in range(10):
        k = int(rng.integers(3, 7))
        zs = sorted(int(z) for z in rng.choice(pool, size=k, replace=False))
        w = rng.dirichlet(np.ones(k))
        layers.append(x.LayerD(zs, [float(v) for v in w], float(rng.uniform(1, 10)),
                               float(10 ** rng.uniform(-4, -1))))
    d = copy.deepcopy(base)
    d.layers = layers
    d.reference_layer = 1
    d.n_interactions_trajectory = n_int
    d.n_photons_line = n_photons
    if n_lines == 1:
        d.discrete = [x.DiscreteD(28.0, 1e12, 1e9)]
    else:
        d.discrete = [x.DiscreteD(float(e), 1e10, 1e9) for e in np.linspace(20.0, 40.0, n_lines)]
    d.gain = 0.02
    d.zero = 0.0
    return d


def ebel_like(n_intervals=1000, n_photons_interval=10000, n_photons_line=10000, e_max=40.0):
    """BASELINE config 5 shape: a tube-like continuum of n_intervals trapezoid intervals from 1 keV to
    e_max (Kramers shape, unpolarised) plus Ag K/L characteristic lines, two of them broadened (Gaussian /
    Lorentzian) to exercise those samplers.  Authored synthetically: xmi_tube_ebel needs xraylib."""
    import numpy as np
    d = copy.deepcopy(example("srm1155"))
    es = np.linspace(1.0, e_max, n_intervals + 1)
    cont = []
    for e in es:
        inten = 1e8 * max(e_max / e - 1.0, 0.0) * np.exp(-2.0 / e)
        cont.append(x.ContinuousD(float(e), float(inten / 2), float(inten / 2)))
    d.continuous = cont
    d.discrete = [x.DiscreteD(2.984, 2e8, 2e8), x.DiscreteD(3.151, 1e8, 1e8),
                  x.DiscreteD(21.990, 4e8, 4e8, distribution_type=1, scale_parameter=0.05),
                  x.DiscreteD(22.163, 8e8, 8e8),
                  x.DiscreteD(24.942, 2e8, 2e8, distribution_type=2, scale_parameter=0.02)]
    d.n_photons_interval = n_photons_interval
    d.n_photons_line = n_photons_line
    d.gain = 0.025
    d.zero = 0.0
    return d


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
