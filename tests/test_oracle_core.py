"""CPU tests pinning the oracle's building blocks against published / reference-held vectors."""
import ctypes as C
import math

import numpy as np

import orc


def test_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_quadratic_solver_reference_cases():
    """The ten (a, b, c) triples of the reference's tests/test-poly-solve-quadratic.F90:11-13, checked
    against the closed form (the reference checks against fgsl_poly_solve_quadratic to 1e-7)."""
    a = [0.0, 0.0, 0.0, 0.0, 1.0, -5.00, 5.00, 3.00, -2.56, 1e-3]
    b = [0.0, 0.0, 3.0, 3.0, 2.0, 8.50, -8.50, 0.00, 7.25, -2.00]
    c = [0.0, 1.0, 0.0, -5.0, 1.0, -27.0, 27.0, 6.00, -6e-3, -2.00]
    L = orc.lib()
    for ai, bi, ci in zip(a, b, c):
        r1, r2 = C.c_double(), C.c_double()
        n = L.orc_poly_solve_quadratic(ai, bi, ci, C.byref(r1), C.byref(r2))
        if ai == 0.0:
            if bi == 0.0:
                assert n == 0
            else:
                assert n == 1 and abs(r1.value - (-ci / bi)) < 1e-7
            continue
        disc = bi * bi - 4 * ai * ci
        if disc < 0:
            assert n == 0
            continue
        assert n == 2
        roots = sorted([(-bi + math.sqrt(disc)) / (2 * ai), (-bi - math.sqrt(disc)) / (2 * ai)])
        assert abs(r1.value - roots[0]) < 1e-7 and abs(r2.value - roots[1]) < 1e-7
        for r in (r1.value, r2.value):
            assert abs(ai * r * r + bi * r + ci) < 1e-6 * max(1.0, abs(ci))


def test_oracle_solid_angle_on_axis_analytic():
    """No collimator, point on the detector axis (theta = pi/2): Omega = 2 pi (1 - cos(atan(R/r))).  The cone
    the algorithm samples is exactly the detector cone there, so every ray hits."""
    import xmimsim_b200 as x
    from inputs import example, no_collimator
    inp = no_collimator(example("srm1155"))
    ci = x.CInput(inp)
    d = orc.init_input(C.pointer(ci.input))
    R = d.detector_radius
    for r in (0.5, 2.0, 7.5):
        hits = C.c_long()
        sa = orc.lib().orc_single_solid_angle(C.byref(d), r, math.pi / 2, 2000, 1234, 7, C.byref(hits))
        exact = 2 * math.pi * (1 - math.cos(math.atan(R / r)))
        assert hits.value >= 1990          # rays on the rim may round outside
        assert abs(sa - exact) / exact < 6e-3


def test_oracle_solid_angle_off_axis_statistical():
    """Off axis the estimate must agree with a direct numerical quadrature of the disc's solid angle."""
    import xmimsim_b200 as x
    from inputs import example, no_collimator
    inp = no_collimator(example("srm1155"))
    ci = x.CInput(inp)
    d = orc.init_input(C.pointer(ci.input))
    R = d.detector_radius
    r1, th1 = 3.0, 0.6
    p = np.array([0.0, r1 * math.cos(th1), r1 * math.sin(th1)])
    # quadrature over the disc: dOmega = z dA / |x - p|^3
    n = 400
    rr = (np.arange(n) + 0.5) / n * R
    ph = (np.arange(n) + 0.5) / n * 2 * math.pi
    RR, PH = np.meshgrid(rr, ph, indexing="ij")
    X, Y = RR * np.cos(PH), RR * np.sin(PH)
    dist3 = ((X - p[0]) ** 2 + (Y - p[1]) ** 2 + p[2] ** 2) ** 1.5
    exact = float(np.sum(p[2] * RR / dist3) * (R / n) * (2 * math.pi / n))
    hits = C.c_long()
    N = 200000
    sa = orc.lib().orc_single_solid_angle(C.byref(d), r1, th1, N, 99, 3, C.byref(hits))
    pfrac = hits.value / N
    sigma = sa * math.sqrt((1 - pfrac) / (N * pfrac))
    assert abs(sa - exact) < 4 * sigma + 1e-4 * exact
