"""The engine's first-order line intensities against direct quadrature (tests/fp_closed_form.py), the anchor that does not
pass through the oracle: CaSO4 known-answer set-up of the reference (tests/test-xmimsim-main-CaSO4.c:19-20 asserts 1 %)
and a close-detector steel slab.  Tables, solid-angle grid and histories all come from the product here."""
import numpy as np
import pytest

import xmimsim_b200 as x
from fp_closed_form import first_order_line_intensity, KL3_LINE
from inputs import caso4, close_detector

pytestmark = pytest.mark.gpu


def _engine_line(inp, Z, gpu_tables=False):
    sim = x.Simulation(inp, quality=0, gpu_tables=gpu_tables)
    g, r, t = sim.solid_angle_calculation(hits_per_single=2000, seed=9)         # full 1024 x 1024 grid
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    ch, br, vr = sim.main_msim(x.main_options(), sa)
    prov = sim.provider.contents
    out = vr[Z - 1, KL3_LINE - 1, 0], ch, prov
    sim.close()
    return out


@pytest.mark.parametrize("gpu_tables", [False, True])
def test_caso4_ca_kl3_first_order_matches_the_quadrature(gpu_tables):
    inp = caso4()
    inp.n_photons_line = 400000
    got, ch, prov = _engine_line(inp, 20, gpu_tables)
    want = first_order_line_intensity(inp, prov, 20)
    assert abs(got / want - 1.0) < 0.01, (got, want)
    # the line's channel carries it (channel spectrum is rebuilt from the history slots)
    e_line = prov.LineEnergy(20, -KL3_LINE)
    c = int((e_line - inp.zero) / inp.gain)
    assert ch[1, c] >= got * (1.0 - 1e-12)


def test_close_detector_fe_kl3_first_order_matches_the_quadrature():
    inp = close_detector(n_photons=400000, n_int=1)
    got, ch, prov = _engine_line(inp, 26)
    want = first_order_line_intensity(inp, prov, 26)
    assert abs(got / want - 1.0) < 0.01, (got, want)
