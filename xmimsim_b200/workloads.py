"""Named workloads of BASELINE.json that are authored, not shipped as files: configs[3] (synthetic 10-layer stratified
sample) and configs[4] (1000-interval tube-like continuum with characteristic lines).  Used by bench.py, the tools and
the parity tests (tests/inputs.py re-exports them), so that every place measures and checks the same inputs."""
import copy
import os

from .xmsi import ContinuousD, DiscreteD, LayerD, read_xmsi

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def example(name):
    """A shipped example input of the reference (examples/<name>.xmsi, kept under tests/golden/)."""
    return read_xmsi(os.path.join(_GOLDEN, name + ".xmsi"))


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
        layers.append(LayerD(zs, [float(v) for v in w], float(rng.uniform(1, 10)),
                               float(10 ** rng.uniform(-4, -1))))
    d = copy.deepcopy(base)
    d.layers = layers
    d.reference_layer = 1
    d.n_interactions_trajectory = n_int
    d.n_photons_line = n_photons
    if n_lines == 1:
        d.discrete = [DiscreteD(28.0, 1e12, 1e9)]
    else:
        d.discrete = [DiscreteD(float(e), 1e10, 1e9) for e in np.linspace(20.0, 40.0, n_lines)]
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
        cont.append(ContinuousD(float(e), float(inten / 2), float(inten / 2)))
    d.continuous = cont
    d.discrete = [DiscreteD(2.984, 2e8, 2e8), DiscreteD(3.151, 1e8, 1e8),
                  DiscreteD(21.990, 4e8, 4e8, distribution_type=1, scale_parameter=0.05),
                  DiscreteD(22.163, 8e8, 8e8),
                  DiscreteD(24.942, 2e8, 2e8, distribution_type=2, scale_parameter=0.02)]
    d.n_photons_interval = n_photons_interval
    d.n_photons_line = n_photons_line
    d.gain = 0.025
    d.zero = 0.0
    return d
