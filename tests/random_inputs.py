"""Random valid inputs for the randomized parity tests (tests/test_random_inputs_gpu.py, tools/random_parity_hunt.py):
layer stacks, geometry, sources, absorbers, detector and options drawn from a seed.  Every draw stays inside what
xmi_input_validate (src/xmi_data_structs.c:899-1255) accepts."""
import math

import numpy as np

import xmimsim_b200 as x

POOL = [6, 7, 8, 11, 12, 13, 14, 15, 16, 17, 19, 20, 22, 24, 25, 26, 27, 28, 29, 30, 33, 35, 38, 40, 42, 47, 48, 50, 53,
        56, 58, 64, 73, 74, 78, 79, 82, 83]


def _layer(rng, gas=False, thin=False):
    k = int(rng.integers(1, 7))
    zs = sorted(int(z) for z in rng.choice(POOL, size=k, replace=False))
    w = rng.dirichlet(np.ones(k))
    w = [float(v) for v in w / w.sum()]
    if gas:
        return x.LayerD(zs, w, float(10 ** rng.uniform(-3.2, -2.5)), float(rng.uniform(0.5, 4.0)))
    return x.LayerD(zs, w, float(10 ** rng.uniform(-0.3, 1.28)), float(10 ** rng.uniform(-4, -2 if thin else -0.5)))


def random_input(seed, n_photons=1500):
    """One random set-up: 1 - 7 layers (optionally behind a gas gap), the 45 degree / 90 degree geometry of the shipped
    examples with the detector distance, area and collimator varied, 1 - 4 discrete lines (some broadened, some from a
    Gaussian source) and optionally a continuous block, excitation / detector absorbers, 1 - 6 interactions."""
    rng = np.random.default_rng(1000 + seed)
    n_layers = int(rng.integers(1, 8))
    layers = []
    gap = rng.random() < 0.5
    if gap:
        layers.append(_layer(rng, gas=True))
    for _ in range(n_layers):
        layers.append(_layer(rng, thin=n_layers > 3))
    ref = 2 if gap else 1
    if rng.random() < 0.25:
        ref = int(rng.integers(1, len(layers) + 1))
    det_d = float(rng.uniform(0.8, 3.0))
    area = float(rng.uniform(0.1, 1.0))
    coll_h = coll_d = 0.0
    c = rng.random()
    if c < 0.4:                                  # conical collimator in front of the window
        coll_h = float(rng.uniform(0.1, 0.6) * det_d)
        coll_d = float(rng.uniform(0.3, 0.9) * 2.0 * math.sqrt(area / math.pi))
    elif c < 0.55:                               # cylindrical branch (src/xmi_solid_angle_f.F90:488-519)
        coll_h = float(rng.uniform(0.1, 0.5) * det_d)
        coll_d = 2.0 * (math.sqrt(area / math.pi) - 5e-7)
    e_max = float(rng.uniform(8.0, 60.0))
    gauss_source = rng.random() < 0.25
    disc = []
    for _ in range(int(rng.integers(1, 5))):
        e = float(rng.uniform(3.0, e_max))
        h, v = float(10 ** rng.uniform(8, 10)), float(10 ** rng.uniform(8, 10))
        if rng.random() < 0.2:
            v = 0.0                              # fully polarised
        kw = {}
        if gauss_source:
            kw = dict(sigma_x=float(rng.uniform(1e-3, 2e-2)), sigma_xp=float(rng.uniform(1e-4, 1e-3)),
                      sigma_y=float(rng.uniform(1e-3, 2e-2)), sigma_yp=float(rng.uniform(1e-4, 1e-3)))
        t = rng.random()
        if t < 0.15:
            kw.update(distribution_type=1, scale_parameter=float(rng.uniform(0.01, 0.1)))
        elif t < 0.3:
            kw.update(distribution_type=2, scale_parameter=float(rng.uniform(0.005, 0.05)))
        disc.append(x.DiscreteD(e, h, v, **kw))
    disc.sort(key=lambda d: d.energy)
    for a, b in zip(disc[:-1], disc[1:]):        # the validator rejects equal energies
        if b.energy - a.energy < 1e-3:
            b.energy = a.energy + 1e-3
    cont = []
    if rng.random() < 0.4:
        es = np.sort(rng.uniform(2.0, e_max, int(rng.integers(3, 8))))
        es = es[np.concatenate([[True], np.diff(es) > 1e-2])]
        if es.size >= 2:
            for e in es:
                inten = float(10 ** rng.uniform(6, 8))
                cont.append(x.ContinuousD(float(e), inten, inten * float(rng.uniform(0.2, 1.0))))
    exc = [_abs_layer(rng)] if rng.random() < 0.35 else []
    det = [_abs_layer(rng)] if rng.random() < 0.7 else []
    top = max([d.energy * (1.3 if d.distribution_type else 1.0) for d in disc] + [c_.energy for c_ in cont])
    nch = int(rng.choice([512, 1024, 2048, 4096]))
    gain = float(top * rng.uniform(0.8, 1.3) / nch)    # sometimes the top of the source lies beyond the last channel
    inp = x.InputD(
        n_photons_interval=max(1, n_photons // 2), n_photons_line=n_photons, n_interactions_trajectory=int(rng.integers(1, 7)),
        layers=layers, reference_layer=ref, d_sample_source=100.0, n_sample_orientation=[0.0, 0.707107, 0.707107],
        p_detector_window=[0.0, -det_d, 100.0], n_detector_orientation=[0.0, 1.0, 0.0], area_detector=area,
        collimator_height=coll_h, collimator_diameter=coll_d, d_source_slit=100.0,
        slit_size_x=float(10 ** rng.uniform(-3.5, -2)), slit_size_y=float(10 ** rng.uniform(-3.5, -2)),
        discrete=disc, continuous=cont, exc_layers=exc, det_layers=det, detector_type=int(rng.integers(0, 3)),
        live_time=float(rng.uniform(0.5, 100.0)), pulse_width=1e-5, gain=gain, zero=float(rng.uniform(-0.05, 0.05)),
        fano=0.12, noise=0.1, nchannels=nch, crystal_layers=[x.LayerD([14], [1.0], 2.33, 0.5)])
    if seed >= 1000:
        # seeds from 1000: the sample normal and the detector direction leave the 45 / 90 degree set-up of the examples --
        # normal (0, sin a, cos a), a in [25, 65] degrees; the detector looks at the beam spot from a direction up to 15 degrees
        # behind and 30 degrees in front of the perpendicular to the beam, up to 0.2 out of the plane of incidence
        a = math.radians(float(rng.uniform(25.0, 65.0)))
        gam = math.radians(float(rng.uniform(-15.0, 30.0)))
        u = np.array([float(rng.uniform(-0.2, 0.2)), -math.cos(gam), -math.sin(gam)])
        u /= np.linalg.norm(u)
        inp.n_sample_orientation = [0.0, math.sin(a), math.cos(a)]
        inp.p_detector_window = [float(det_d * u[0]), float(det_d * u[1]), float(100.0 + det_d * u[2])]
        inp.n_detector_orientation = [float(-u[0]), float(-u[1]), float(-u[2])]
    opts = dict(use_M_lines=int(rng.random() < 0.7), use_cascade_auger=int(rng.random() < 0.6),
                use_cascade_radiative=int(rng.random() < 0.6))
    return inp, opts


def _abs_layer(rng):
    z = int(rng.choice([4, 6, 13, 14, 29]))
    rho = {4: 1.85, 6: 2.2, 13: 2.7, 14: 2.33, 29: 8.96}[z]
    return x.LayerD([z], [1.0], rho, float(10 ** rng.uniform(-4, -2.3)))
