"""CPU tests of the host side: library loads and exports the declared ABI, input handling, init_input
against the oracle's independent restatement, table sanity."""
import ctypes as C
import math

import numpy as np
import pytest

import orc
import xmimsim_b200 as x
from xmimsim_b200 import abi
from inputs import example, caso4, no_collimator


def test_library_exports_every_declared_symbol():
    L = abi.lib()
    names = abi.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match_reference_sizes():
    # LP64 sizes of the reference structs (include/xmi_data_structs.h) -- a layout drift breaks the drop-in
    assert C.sizeof(abi.General) == 48
    assert C.sizeof(abi.Layer) == 40
    assert C.sizeof(abi.Geometry) == 128
    assert C.sizeof(abi.EnergyDiscrete) == 72
    assert C.sizeof(abi.EnergyContinuous) == 56
    assert C.sizeof(abi.Detector) == 72
    assert C.sizeof(abi.MainOptions) == 64
    assert C.sizeof(abi.SolidAngle) == 48
    assert C.sizeof(abi.EscapeRatios) == 72


def test_main_options_defaults():
    o = x.main_options()
    assert (o.use_M_lines, o.use_cascade_auger, o.use_cascade_radiative, o.use_variance_reduction) == (1, 1, 1, 1)
    assert (o.use_sum_peaks, o.use_escape_peaks, o.use_poisson, o.use_advanced_compton) == (0, 1, 0, 0)


@pytest.mark.parametrize("name", ["srm1155", "srm1412", "srm1132", "In"])
def test_init_input_matches_oracle(name):
    inp = example(name)
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    d = sim.derived
    assert d.detector_radius == od.detector_radius
    assert d.collimator_present == od.collimator_present
    assert d.collimator_radius == od.collimator_radius
    assert d.half_apex == od.half_apex
    assert list(d.ndo_new) == list(od.ndo_new)
    assert list(d.ndo_inv) == list(od.ndo_inv)
    assert d.detector_solid_angle == od.detector_solid_angle
    assert list(d.n_sample_orientation_det) == list(od.n_sample_orientation_det)
    for i in range(d.n_layers):
        assert d.Z_coord_begin[i] == od.Z_coord_begin[i]
        assert d.Z_coord_end[i] == od.Z_coord_end[i]
    # frame sanity: inverse * new = identity, x' = detector normal
    A = np.array(list(d.ndo_new)).reshape(3, 3)
    B = np.array(list(d.ndo_inv)).reshape(3, 3)
    assert np.allclose(B @ A, np.eye(3), atol=1e-14)
    assert np.allclose(A[:, 0], list(d.n_detector_orientation))
    sim.close()


def test_xmsi_reader_conventions(srm1155):
    # elements sorted by Z, weights normalised (src/xmi_xml.c:1209-1263), lines sorted (:966)
    for l in srm1155.layers:
        assert l.Z == sorted(l.Z)
        assert abs(sum(l.weight) - 1.0) < 1e-12
    es = [d.energy for d in srm1155.discrete]
    assert es == sorted(es) and len(es) == 26
    assert srm1155.n_photons_line == 150000 and srm1155.n_interactions_trajectory == 4
    assert srm1155.detector_type == 2 and srm1155.nchannels == 2048


def test_input_validation():
    inp = caso4()
    inp.layers[0].Z = [120]
    with pytest.raises(RuntimeError):
        x.Simulation(inp)
    inp = caso4()
    inp.collimator_height = 1.0
    inp.collimator_diameter = 5.0      # wider than the detector: the reference exits, we return 0
    with pytest.raises(RuntimeError, match="Non conical"):
        x.Simulation(inp)


def test_solid_angle_axes_match_oracle():
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    r, t = sim.solid_angle_inputs()
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    ro, to = orc.solid_angle_axes(C.pointer(ci.input), od)
    assert np.array_equal(r, ro) and np.array_equal(t, to)
    assert r.size == 1024 and t.size == 1024 and t[0] == 1e-5 and abs(t[-1] - math.pi / 2) < 1e-15
    assert abs(r[0] - r[-1] / 1024) < 1e-15
    sim.close()


def test_tables_sanity():
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    T = sim.tables
    P = sim.provider.contents
    nZ, nN = T.nZ, T.n_nodes
    Z = [T.Z[i] for i in range(nZ)]
    assert Z == sorted(set(z for l in inp.layers for z in l.Z))
    E = np.ctypeslib.as_array(T.node_E, shape=(nN,))
    assert np.all(np.diff(E) > 0) and E[0] == 0.1 and E[-1] >= max(d.energy for d in inp.discrete)
    cs = np.ctypeslib.as_array(T.cs_total, shape=(nZ, nN))
    pr = np.ctypeslib.as_array(T.p_rayl, shape=(nZ, nN))
    prc = np.ctypeslib.as_array(T.p_rayl_compt, shape=(nZ, nN))
    assert np.all(cs > 0) and np.all(pr > 0) and np.all(prc > pr) and np.all(prc < 1)
    # tabulation error of linear interpolation between nodes vs the provider, away from edges
    rng = np.random.default_rng(1)
    iz = Z.index(26)
    for e in rng.uniform(1.0, 16.0, 200):
        k = np.searchsorted(E, e) - 1
        f = (e - E[k]) / (E[k + 1] - E[k])
        lin = cs[iz, k] * (1 - f) + cs[iz, k + 1] * f
        assert abs(lin - P.CS_Total_Kissel(26, e)) / lin < 5e-4
    # inverse CDFs: monotone in R, end points pinned (src/xmi_data_f.F90:1036-1037)
    ric = np.ctypeslib.as_array(T.rayl_theta_icdf, shape=(nZ, T.n_icdf_E, T.n_icdf_R))
    cic = np.ctypeslib.as_array(T.compt_theta_icdf, shape=(nZ, T.n_icdf_E, T.n_icdf_R))
    for a in (ric, cic):
        assert np.all(a[:, :, 0] == 0.0) and np.all(np.abs(a[:, :, -1] - math.pi) < 1e-12)
        assert np.all(np.diff(a, axis=2) >= 0)
    phi = np.ctypeslib.as_array(T.phi_icdf, shape=(T.n_phi_T, T.n_icdf_R))
    assert np.all(np.diff(phi, axis=1) >= 0) and np.allclose(phi[0], np.linspace(0, 2 * math.pi, T.n_icdf_R), atol=2e-3)
    # corrected yields (src/xmi_data_f.F90:1244-1251): L1 = w1 + f12 w2 + (f13 + f12 f23) w3
    fy = np.ctypeslib.as_array(T.fluor_yield, shape=(nZ, 9))
    fyc = np.ctypeslib.as_array(T.fluor_yield_corr, shape=(nZ, 9))
    ck = np.ctypeslib.as_array(T.cos_kron, shape=(nZ, 13))
    i82 = Z.index(82)
    exp_l1 = fy[i82, 1] + ck[i82, 0] * fy[i82, 2] + (ck[i82, 1] + ck[i82, 0] * ck[i82, 2]) * fy[i82, 3]
    assert abs(fyc[i82, 1] - exp_l1) < 1e-15
    assert fyc[i82, 0] == fy[i82, 0] and fyc[i82, 3] == fy[i82, 3]
    # radiative rates of a shell's line range sum to one (assumed at src/xmi_main.F90:5422-5426)
    rr = np.ctypeslib.as_array(T.rad_rate, shape=(nZ, 384))
    assert abs(rr[i82, 1:30].sum() - 1.0) < 1e-12 and abs(rr[i82, 86:114].sum() - 1.0) < 1e-12
    sim.close()


@pytest.mark.parametrize("name", ["srm1412", "srm1155"])
def test_node_tables_equal_the_provider_at_every_sampled_node(name):
    """Every energy-dependent table holds the provider's value AT its nodes (interpolation happens between them): cross
    sections, branching ratios, partial and vacancy cross sections for the four cascade modes, mu of the layers, the exciter
    absorbers' optical depth -- checked entry by entry on a random sample of nodes, against calls of the provider's own
    functions.  Engine and oracle read the same bundle, so only this kind of test sees an error of the generator."""
    inp = example(name)
    sim = x.Simulation(inp, quality=0)
    T, P = sim.tables, sim.provider.contents
    nZ, nN, nL = T.nZ, T.n_nodes, T.n_layers
    Z = [T.Z[i] for i in range(nZ)]
    E = np.ctypeslib.as_array(T.node_E, shape=(nN,))
    arr = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
    cs, ph = arr(T.cs_total, (nZ, nN)), arr(T.cs_photo_total, (nZ, nN))
    pr, prc = arr(T.p_rayl, (nZ, nN)), arr(T.p_rayl_compt, (nZ, nN))
    part, vac = arr(T.cs_photo_partial, (nZ, 9, nN)), arr(T.cs_vacancy, (4, nZ, 9, nN))
    mu, exc = arr(T.mu_layer, (nL, nN)), arr(T.exc_murhod, (nN,))
    rng = np.random.default_rng(7)
    nodes = np.unique(np.concatenate([rng.integers(0, nN, 60), [0, nN - 1]]))
    buf = (C.c_double * 9)()
    for n in nodes:
        e = float(E[n])
        for iz, z in enumerate(Z):
            tot = P.CS_Total_Kissel(z, e)
            assert cs[iz, n] == tot and ph[iz, n] == P.CS_Photo_Total(z, e)
            assert pr[iz, n] == P.CS_Rayl(z, e) / tot
            assert prc[iz, n] == P.CS_Compt(z, e) / tot + P.CS_Rayl(z, e) / tot
            for sh in range(9):
                assert part[iz, sh, n] == P.CS_Photo_Partial(z, sh, e)
            for mode in (1, 2, 3, 4):
                for sh in range(9):
                    buf[sh] = P.VacancyCS(z, sh, e, mode, buf)
                    assert vac[mode - 1, iz, sh, n] == buf[sh], (z, mode, sh, e)
        for k, lay in enumerate(inp.layers):
            w = np.array(lay.weight) / np.sum(lay.weight)
            want = sum(wi * P.CS_Total_Kissel(int(zz), e) for zz, wi in zip(lay.Z, w))
            assert abs(mu[k, n] - want) <= 4e-16 * want, (k, e)
        want = sum(l.density * l.thickness * sum(wi * P.CS_Total_Kissel(int(zz), e) for zz, wi in zip(l.Z, np.array(l.weight) / np.sum(l.weight)))
                   for l in inp.exc_layers)
        assert abs(exc[n] - want) <= 1e-15 * max(want, 1e-300)
    # every fluorescence line of every element and every source line is a node (exact lookups: precalc_mu_cs, precalc_xrf_cs)
    le = arr(T.line_energy, (nZ, 384))
    for iz in range(nZ):
        for l in range(1, 220):
            if 0.1 <= le[iz, l] < E[-1]:
                assert le[iz, l] in E
    for d in inp.discrete:
        assert d.energy in E
    sim.close()


def test_plugin_file_exports_reference_symbol_names():
    """The drop-in plugin file (name the reference's GModule loader opens, src/xmi_solid_angle.c:121-136) exports the
    reference's symbol names (src/xmi_solid_angle_cl.c:118-120; bin/xmimsim.c:513)."""
    import os
    path = os.path.join(os.path.dirname(abi.LIB_PATH), "xmimsim-cl.so")
    assert os.path.exists(path)
    P = C.CDLL(path)
    assert hasattr(P, "xmi_solid_angle_calculation_cl") and hasattr(P, "xmi_detector_convolute_all_custom")


def test_inverse_cdf_tables_reproduce_the_analytic_distributions():
    """Tier T1: the table generator against quadratures that need no reference data.
    theta tables: moments of the sampled angle vs direct integration of the same differential cross sections
    (src/xmi_data_f.F90:1002-1060); phi table: CDF(phi; a) = (phi - a sin 2 phi) / 2 pi inverted exactly (:1508-1545);
    Compton-profile table: P(|pz| < q) (:1162-1186)."""
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=1)
    T = sim.tables
    P = sim.provider.contents
    nZ, nE, nR = T.nZ, T.n_icdf_E, T.n_icdf_R
    rayl = np.ctypeslib.as_array(T.rayl_theta_icdf, shape=(nZ, nE, nR))
    compt = np.ctypeslib.as_array(T.compt_theta_icdf, shape=(nZ, nE, nR))
    E = np.ctypeslib.as_array(T.icdf_E, shape=(nE,))
    Z = [T.Z[i] for i in range(nZ)]
    th = np.linspace(0.0, math.pi, 20001)
    for z in (26, 8):
        iz = Z.index(z)
        for je in (5, nE - 3):
            q = E[je] / 12.39841930 * np.sin(th / 2)
            F = np.array([P.FF_Rayl(z, v) for v in q]); S = np.array([P.SF_Compt(z, v) for v in q])
            k = 1.0 / (1.0 + E[je] / 510.998928 * (1 - np.cos(th)))
            f_r = (1 + np.cos(th) ** 2) * F ** 2 * np.sin(th)
            f_c = k ** 2 * (k + 1 / k - np.sin(th) ** 2) * S * np.sin(th)
            for tab, f in ((rayl, f_r), (compt, f_c)):
                w = np.trapezoid(f, th)
                mean_cos = np.trapezoid(f * np.cos(th), th) / w
                mean_cos2 = np.trapezoid(f * np.cos(th) ** 2, th) / w
                icdf = tab[iz, je]
                assert icdf[0] == 0.0 and abs(icdf[-1] - math.pi) < 1e-12 and np.all(np.diff(icdf) >= 0)
                # sampling theta = icdf(R), R uniform: trapezoid over the R grid
                s1 = np.trapezoid(np.cos(icdf), dx=1.0 / (nR - 1)); s2 = np.trapezoid(np.cos(icdf) ** 2, dx=1.0 / (nR - 1))
                assert abs(s1 - mean_cos) < 2e-3 and abs(s2 - mean_cos2) < 2e-3, (z, je, s1, mean_cos)
    nT = T.n_phi_T
    phi = np.ctypeslib.as_array(T.phi_icdf, shape=(nT, nR)); a = np.ctypeslib.as_array(T.phi_T, shape=(nT,))
    R = np.arange(nR) / (nR - 1.0)
    for it in (0, nT // 2, nT - 1):
        back = (phi[it] - a[it] * np.sin(2 * phi[it])) / (2 * math.pi)
        assert np.abs(back - R).max() < 2e-4 and phi[it, 0] == 0.0 and abs(phi[it, -1] - 2 * math.pi) < 1e-12
    n_cp = T.n_cp
    cp = np.ctypeslib.as_array(T.cp_icdf, shape=(nZ, n_cp))
    iz = Z.index(26)
    pz = np.linspace(0, 100, 400001)
    J = np.array([P.ComptonProfile(26, v) for v in pz[::40]])
    cdf = np.concatenate([[0], np.cumsum((J[1:] + J[:-1]) / 2 * np.diff(pz[::40]))]); cdf /= cdf[-1]
    for r in (0.1, 0.5, 0.9, 0.99):
        q_tab = cp[iz, int(round(r * (n_cp - 1)))]
        assert abs(np.interp(q_tab, pz[::40], cdf) - r) < 2e-3
    sim.close()
