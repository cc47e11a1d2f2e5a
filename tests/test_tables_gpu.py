"""The GPU table generator (inverse CDFs of the scattering angles and of the Compton profile) against the host
generator: same grids, entries within one integration step of each other."""
import math
import time

import numpy as np
import pytest

import xmimsim_b200 as x
from inputs import example

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("quality", [0, 1])
def test_gpu_generator_matches_host_generator(quality):
    inp = example("srm1155")
    t0 = time.time(); host = x.Simulation(inp, quality=quality); t_host = time.time() - t0
    t0 = time.time(); dev = x.Simulation(inp, quality=quality, gpu_tables=True); t_dev = time.time() - t0
    A, B = host.tables, dev.tables
    nZ, nE, nR, n_cp = A.nZ, A.n_icdf_E, A.n_icdf_R, A.n_cp
    assert (B.nZ, B.n_icdf_E, B.n_icdf_R, B.n_cp) == (nZ, nE, nR, n_cp)
    n_theta = 100000 if quality else 20000
    n_pz = 10000000 if quality else 400000
    for name, step, shape in (("rayl_theta_icdf", math.pi / (n_theta - 1), (nZ, nE, nR)), ("compt_theta_icdf", math.pi / (n_theta - 1), (nZ, nE, nR)),
                              ("cp_icdf", 100.0 / (n_pz - 1), (nZ, n_cp))):
        a = np.ctypeslib.as_array(getattr(A, name), shape=shape); b = np.ctypeslib.as_array(getattr(B, name), shape=shape)
        d = np.abs(a - b)
        # Where the density is tiny (flat CDF: large angles at high momentum transfer, the far tail of the profile) a
        # rounding-level change of the running sum moves the crossing by many steps; such entries are compared in
        # probability instead: the position of the GPU abscissa in the host row.
        Rg = np.arange(shape[-1]) / (shape[-1] - 1.0)
        a2, b2 = a.reshape(-1, shape[-1]), b.reshape(-1, shape[-1])
        worst = 0.0
        for ra, rb in zip(a2, b2):
            far = np.abs(ra - rb) > 1.5 * step
            far[0] = far[-1] = False
            if far.any():
                worst = max(worst, np.abs(np.interp(rb[far], ra, Rg) - Rg[far]).max())
        print(name, "max step diff %.1f, entries off by > half a step: %.4f, worst probability diff of those beyond 1.5 steps: %.2e"
              % (d.max() / step, (d > 0.5 * step).mean(), worst))
        assert worst < 1e-5, (name, worst)
        assert d.max() <= 2.5 * step, (name, d.max() / step)
        assert (d > 0.5 * step).mean() < 0.05, (name, (d > 0.5 * step).mean())
        assert np.all(np.diff(b, axis=-1) >= 0)
    # everything else is the same host code
    for name, n in (("cs_total", nZ * A.n_nodes), ("phi_icdf", A.n_phi_T * nR), ("ff", nZ * A.n_q)):
        assert np.array_equal(np.ctypeslib.as_array(getattr(A, name), shape=(n,)), np.ctypeslib.as_array(getattr(B, name), shape=(n,)))
    print("quality %d: host %.2f s, gpu path %.2f s (kernels %.1f ms)" % (quality, t_host, t_dev, dev.L.xmb_tables_gpu_last_ms()))
    host.close(); dev.close()


def test_histories_with_gpu_tables_agree_statistically():
    """A simulation on GPU-generated tables gives the same spectrum as one on host-generated tables up to the few
    histories whose sampled angle sits on a shifted table entry."""
    inp = example("srm1155")
    inp.n_photons_line = 4000
    res = []
    for gpu_tables in (False, True):
        sim = x.Simulation(inp, quality=0, gpu_tables=gpu_tables)
        sim.solid_angle_calculation(hits_per_single=300, seed=2)
        ch, br, vr = sim.main_msim(x.main_options())
        res.append((ch, vr))
        sim.close()
    (c0, v0), (c1, v1) = res
    assert abs(c1[-1].sum() / c0[-1].sum() - 1.0) < 2e-3
    assert abs(v1[25, 2].sum() / v0[25, 2].sum() - 1.0) < 2e-3
