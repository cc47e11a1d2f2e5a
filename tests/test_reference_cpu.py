"""The oracle restatement against the reference's OWN code run here (oracle/_ref/libxmi_ref.so, built by
oracle/build_ref.sh from /root/reference/src/xmi_kernels.cl and src/xmi_spline.c; see oracle/ref_shim/README.md), and
against the known answers of the reference's tests/test-cubic-spline.c.

Solid angle (north_star: "the deterministic solid-angle grid must agree to a stated relative float
tolerance"): the reference's OpenCL kernel is an fp32 Monte Carlo estimate with Threefry streams keyed by the grid
indices, the oracle (and the CUDA kernel, which matches the oracle hit for hit) an fp64 one with Philox streams.  Two
independent N-ray estimates of the same cone fraction p differ by a binomial error: the stated tolerance is therefore
|a - b| <= 5 sigma + 2e-4 relative per grid point (sigma^2 = both binomial variances; 2e-4 covers the fp32 arithmetic
of the reference kernel), a reduced chi-square of the whole grid within [0.8, 1.25], and the same 5 sigma + 2e-4 rule for the sum over the grid."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402
import ref  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, no_collimator, cylindrical_collimator  # noqa: E402

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")

# tests/test-cubic-spline.c:8-13 (x = 0..4, y = 100, 50, -100, 50, 200; evaluated at i/10, tolerance 1.0 there)
SPLINE_KAT = [100, 99.7732, 99.2571, 98.1625, 96.2, 93.0804, 88.5143, 82.2125, 73.8857, 63.2446,
              50, 34.0518, 16.0571, -3.1375, -22.6857, -41.7411, -59.4571, -74.9875, -87.4857, -96.1054,
              -100, -98.5804, -92.2857, -81.8125, -67.8571, -51.1161, -32.2857, -12.0625, 8.85714, 29.7768,
              50, 68.9696, 86.6857, 103.287, 118.914, 133.705, 147.8, 161.338, 174.457, 187.298, 200.0]


def test_spline_known_answers_reference_and_oracle():
    xs = np.arange(5.0)
    ys = np.array([100.0, 50.0, -100.0, 50.0, 200.0])
    for i, want in enumerate(SPLINE_KAT):
        v = i / 10.0
        r = ref.cubic_spline(xs, ys, v)
        o = orc.lib().orc_cubic_spline(xs.ctypes.data, ys.ctypes.data, xs.size, v)
        assert abs(r - want) < 1.0                      # the reference test's own criterion
        assert abs(r - want) < 6e-4 * max(1.0, abs(want))   # the table is printed to 6 significant digits
        assert abs(o - r) <= 1e-12 * max(1.0, abs(r)), (v, o, r)


def test_spline_oracle_equals_reference_on_random_knots():
    rng = np.random.default_rng(3)
    for n in (3, 4, 7, 30, 200):
        xs = np.cumsum(rng.uniform(0.05, 2.0, n))
        ys = rng.normal(0.0, 10.0, n)
        for v in np.concatenate([xs[:3], rng.uniform(xs[0], xs[-1], 40), [xs[-1]]]):
            r = ref.cubic_spline(xs, ys, v)
            o = orc.lib().orc_cubic_spline(xs.ctypes.data, ys.ctypes.data, xs.size, float(v))
            assert abs(o - r) <= 1e-10 * max(1.0, abs(r)), (n, v, o, r)


def _geometry(variant):
    inp = example("srm1132")                      # BASELINE configs[2]: the srm1132 geometry
    if variant == "none":
        inp = no_collimator(inp)
    elif variant == "cylindrical":
        inp = cylindrical_collimator(inp)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    r_full, t_full = orc.solid_angle_axes(C.pointer(ci.input), od)
    return ci, od, r_full, t_full


def compare_with_reference_kernel(sa, hits, r, t, od, n_rays):
    """sa / hits: an fp64 estimate [theta][r] with its integer hit counts; returns the statistics asserted below."""
    sa_ref = ref.solid_angle_grid_cl(r, t, od.collimator_present, od.detector_radius, od.collimator_radius,
                                     od.collimator_height, n_rays).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        cone = np.where(hits > 0, sa * n_rays / np.maximum(hits, 1), np.nan)     # full-cone solid angle where known
    p_a = hits / n_rays
    p_b = np.clip(np.where(np.isfinite(cone), sa_ref / cone, 0.0), 0.0, 1.0)
    var = cone ** 2 * (p_a * (1 - p_a) + p_b * (1 - p_b)) / n_rays
    known = np.isfinite(cone)
    # where the oracle saw no hit the cone is unknown: the reference must be (nearly) empty there as well
    if np.any(~known):
        assert np.count_nonzero(sa_ref[~known] > 0) <= 0.02 * np.count_nonzero(~known) + 2
    d = (sa - sa_ref)[known]
    sig = np.sqrt(var[known])
    tol = 5.0 * sig + 2e-4 * np.maximum(sa[known], sa_ref[known]) + 1e-12
    stat = sig > 0
    chi2 = float(np.mean((d[stat] / sig[stat]) ** 2)) if np.any(stat) else 1.0
    sum_tol = 5.0 * float(np.sqrt(np.sum(sig ** 2))) + 2e-4 * float(np.sum(sa_ref[known]))   # the grid sum, same rule
    return d, tol, chi2, sa_ref, known, sum_tol


@pytest.mark.parametrize("variant", ["conical", "none", "cylindrical"])
def test_oracle_solid_angle_matches_reference_opencl_kernel(variant):
    ci, od, r_full, t_full = _geometry(variant)
    ri = np.unique(np.concatenate([np.arange(0, 1024, 24), [1023]]))
    ti = np.unique(np.concatenate([np.arange(0, 1024, 24), [1023]]))
    r, t = r_full[ri], t_full[ti]
    n_rays = 5000
    sa, hits = orc.solid_angle_grid(od, r, np.arange(r.size), t, np.arange(t.size), r.size, n_rays, 20260101,
                                    n_threads=os.cpu_count() or 1)
    d, tol, chi2, sa_ref, known, sum_tol = compare_with_reference_kernel(sa, hits, r, t, od, n_rays)
    assert sa_ref.sum() > 0 and np.count_nonzero(known) > 400      # a conical collimator shadows most of the grid
    assert np.all(np.abs(d) <= tol), (float(np.max(np.abs(d) / tol)), int(np.argmax(np.abs(d) / tol)))
    assert 0.8 < chi2 < 1.25, chi2
    assert abs(float(np.sum(d))) <= sum_tol, (float(np.sum(d)), sum_tol)
    # deterministic part: points where every ray hits (or none does) carry no Monte Carlo error at all
    full = known & (hits == n_rays)
    if np.any(full):
        assert np.allclose(sa[full], sa_ref[full], rtol=2e-4)


def test_xmso_history_mapping_matches_the_reference_function(tmp_path):
    """The XMSO writer's history lists (xmb_output_write_to_xml_file, host_io.cpp) against the reference's own
    xmi_output_raw2struct (src/xmi_data_structs.c:1371-1519, compiled from /root/reference into oracle/_ref): same
    elements, lines and interaction entries in the same order, same counts and totals (the writer prints 6 digits), for
    both histories; rows the reference drops (zero entries, the Rayleigh / Compton slots 384 / 385, elements that are
    not in the sample) are dropped.  Line energies are xraylib's and are not compared."""
    import ctypes as C
    import xml.etree.ElementTree as ET
    import xmimsim_b200 as x
    from xmimsim_b200 import abi
    from inputs import example
    if not hasattr(ref.lib(), "ref_output_raw2struct_rows"):
        pytest.skip("oracle/_ref built without the raw2struct shim")
    inp = example("srm1155")
    ci = x.CInput(inp)
    n_int, nch = inp.n_interactions_trajectory, inp.nchannels
    rng = np.random.default_rng(11)
    unconv = np.cumsum(rng.uniform(0, 100, (n_int + 1, nch)), axis=0)
    conv = unconv * 0.9
    sample_Z = sorted({z for l in inp.layers for z in l.Z})
    hist = []
    for seed in (1, 2):
        r = np.random.default_rng(seed)
        h = np.zeros((100, 385, n_int))
        for z in sample_Z[::2] + [79, 3]:                     # 79 / 3: not in the sample -> never listed
            for line in r.choice(np.arange(1, 384), size=12, replace=False):
                k = r.integers(1, n_int + 1)
                h[z - 1, line - 1, :k] = r.uniform(1e-3, 1e6, k) * (r.uniform(size=k) > 0.3)   # some orders empty
            h[z - 1, 383, 0] = 5.0; h[z - 1, 384, 1] = 6.0   # Rayleigh / Compton slots: never listed
        h[sample_Z[1] - 1, :383, 0] = 1.0 + np.arange(383)    # one element with every line: all 383 names of src/xmi_lines.c
        hist.append(h)
    br, vr = hist
    rows_p = (abi.c_double_p * (n_int + 1))(*[C.cast(conv.ctypes.data + i * nch * 8, abi.c_double_p) for i in range(n_int + 1)])
    out = str(tmp_path / "o.xmso")
    assert abi.lib().xmb_output_write_to_xml_file(C.byref(ci.input), b"in.xmsi", out.encode(), unconv.ctypes.data_as(abi.c_double_p),
                                                  rows_p, br.ctypes.data_as(abi.c_double_p), vr.ctypes.data_as(abi.c_double_p), 0,
                                                  None) == 1, abi.last_error()
    root = ET.parse(out).getroot()
    for which, tag, arr in ((0, "brute_force_history", br), (1, "variance_reduction_history", vr)):
        want, _ = ref.output_raw2struct_rows(C.pointer(ci.input), br, vr, rows_p, unconv, 0, which)
        got = []
        for el in root.find(tag).findall("fluorescence_line_counts"):
            for ln in el.findall("fluorescence_line"):
                for c in ln.findall("counts"):
                    got.append((int(el.get("atomic_number")), ln.get("type"), float(el.get("total_counts")), float(ln.get("total_counts")),
                                int(c.get("interaction_number")), float(c.text)))
        assert len(want) > 420 and len(got) == len(want), (tag, len(got), len(want))
        assert len({w[1] for w in want}) == 383
        assert [(g[0], g[1], g[4]) for g in got] == [(w[0], w[1], w[4]) for w in want], tag
        assert {w[0] for w in want} <= set(sample_Z)
        for g, w in zip(got, want):
            assert abs(g[2] - w[2]) <= 2e-5 * w[2] and abs(g[3] - w[3]) <= 2e-5 * w[3] and abs(g[5] - w[5]) <= 2e-5 * w[5], (tag, g, w)


def _variants(base):
    """(label, input) pairs around `base`: each changes what one clause of the reference's match rules looks at, by an
    amount on either side of that clause's threshold."""
    import copy
    import xmimsim_b200 as x
    out = [("identical", copy.deepcopy(base))]

    def var(label, **kw):
        d = copy.deepcopy(base)
        for k, v in kw.items():
            setattr(d, k, v)
        out.append((label, d))
    for f in (1e-13, 1e-9, 1e-4, 1e-2):
        var("area_detector x(1+%g)" % f, area_detector=base.area_detector * (1 + f))
        var("collimator_height +%g" % f, collimator_height=base.collimator_height + f)
        var("collimator_diameter +%g" % f, collimator_diameter=base.collimator_diameter + f)
        var("window y +%g" % f, p_detector_window=[base.p_detector_window[0], base.p_detector_window[1] + f, base.p_detector_window[2]])
        var("window x +%g" % f, p_detector_window=[base.p_detector_window[0] + f, base.p_detector_window[1], base.p_detector_window[2]])
        var("window z +%g" % f, p_detector_window=[base.p_detector_window[0], base.p_detector_window[1], base.p_detector_window[2] + f])
        var("source distance +%g with window" % f, d_sample_source=base.d_sample_source + f,
            p_detector_window=[base.p_detector_window[0], base.p_detector_window[1], base.p_detector_window[2] + f])
        var("detector normal tilt %g" % f, n_detector_orientation=[base.n_detector_orientation[0] + f, base.n_detector_orientation[1], base.n_detector_orientation[2]])
        var("sample normal tilt %g" % f, n_sample_orientation=[base.n_sample_orientation[0], base.n_sample_orientation[1] + f, base.n_sample_orientation[2]])
    var("detector normal scaled", n_detector_orientation=[2.0 * v for v in base.n_detector_orientation])
    var("sample normal scaled", n_sample_orientation=[3.0 * v for v in base.n_sample_orientation])
    var("no collimator", collimator_height=0.0, collimator_diameter=0.0)
    var("photons / live time", n_photons_line=7, live_time=5.0)
    for e in (5.0, 12.0, 15.9, 16.3, 20.0, 40.0, 60.0):
        var("one line at %g keV" % e, discrete=[x.DiscreteD(e, 1e9, 1e9)])
    var("continuum 5..30 keV", continuous=[x.ContinuousD(5.0, 1e6, 1e6), x.ContinuousD(30.0, 1e6, 1e6)])
    var("continuum 15.9..16.1 keV", continuous=[x.ContinuousD(15.9, 1e6, 1e6), x.ContinuousD(16.1, 1e6, 1e6)])
    for f in (0.5, 0.999, 1.001, 2.0):
        d = copy.deepcopy(base); d.layers[-1].thickness *= f; out.append(("last layer thickness x%g" % f, d))
        d = copy.deepcopy(base); d.layers[-1].density *= f; out.append(("last layer density x%g" % f, d))
        d = copy.deepcopy(base); d.layers[0].thickness *= f; out.append(("first layer thickness x%g" % f, d))
    d = copy.deepcopy(base); d.layers = d.layers[:1]; d.reference_layer = 1; out.append(("one layer", d))
    d = copy.deepcopy(base); d.layers = [d.layers[0], d.layers[1], copy.deepcopy(d.layers[1])]; out.append(("three layers", d))
    d = copy.deepcopy(base); d.reference_layer = 1; out.append(("reference layer 1", d))
    return out


@pytest.mark.parametrize("name", ["srm1155", "srm1412"])
def test_solid_angle_cache_match_rule_equals_the_reference_function(name):
    """xmb_check_solid_angle_match (host_cache.cpp) against the reference's xmi_check_solid_angle_match
    (src/xmi_solid_angle.c:420-673, compiled from /root/reference into oracle/_ref) on ~80 input pairs, both directions
    (cached vs fresh is not symmetric: the cached grid must cover the fresh input's depth range)."""
    import ctypes as C
    import xmimsim_b200 as x
    from xmimsim_b200 import abi
    from inputs import example
    if not hasattr(ref.lib(), "ref_check_solid_angle_match"):
        pytest.skip("oracle/_ref built without the match-rule shim")
    base = example(name)
    L = abi.lib()
    seen = {0: 0, 1: 0}
    for label, other in _variants(base):
        for cached, fresh in ((base, other), (other, base)):
            c1, f1 = x.CInput(cached), x.CInput(fresh)
            ours = L.xmb_check_solid_angle_match(C.byref(c1.input), C.byref(f1.input), None)
            a, b = x.CInput(cached), x.CInput(fresh)          # throw-away copies: the reference normalises them in place
            want = ref.check_solid_angle_match(C.pointer(a.input), C.pointer(b.input))
            assert ours == want, (label, "cached=base" if cached is base else "cached=variant", ours, want)
            seen[want] += 1
    assert seen[0] > 20 and seen[1] > 20, seen


def test_escape_ratio_cache_match_rule_equals_the_reference_function():
    """xmb_check_escape_ratios_match against xmi_check_escape_ratios_match (src/xmi_detector.c:143-172)."""
    import copy
    import ctypes as C
    import xmimsim_b200 as x
    from xmimsim_b200 import abi
    from inputs import example
    if not hasattr(ref.lib(), "ref_check_escape_ratios_match"):
        pytest.skip("oracle/_ref built without the match-rule shim")
    base = example("srm1155")
    variants = [copy.deepcopy(base)]
    for f in (1 + 1e-13, 1 + 1e-9, 1.01):
        d = copy.deepcopy(base); d.crystal_layers[0].thickness *= f; variants.append(d)
        d = copy.deepcopy(base); d.crystal_layers[0].density *= f; variants.append(d)
    d = copy.deepcopy(base); d.crystal_layers = [x.LayerD([31, 33], [0.48, 0.52], 5.3, 0.05)]; variants.append(d)
    d = copy.deepcopy(base); d.crystal_layers = d.crystal_layers + [x.LayerD([14], [1.0], 2.33, 0.01)]; variants.append(d)
    d = copy.deepcopy(base); d.area_detector *= 2; d.live_time = 9.0; variants.append(d)          # not looked at
    seen = {0: 0, 1: 0}
    for other in variants:
        for cached, fresh in ((base, other), (other, base)):
            c1, f1, c2, f2 = x.CInput(cached), x.CInput(fresh), x.CInput(cached), x.CInput(fresh)
            ours = abi.lib().xmb_check_escape_ratios_match(C.byref(c1.input), C.byref(f1.input))
            want = ref.check_escape_ratios_match(C.pointer(c2.input), C.pointer(f2.input))
            assert ours == want, (ours, want)
            seen[want] += 1
    assert seen[0] >= 6 and seen[1] >= 6, seen


def _invalid_variants(base):
    """(label, mutate) pairs: each breaks (or leaves intact) one clause of the reference's xmi_input_validate."""
    import xmimsim_b200 as x
    out = [("valid", lambda d: None)]

    def a(label, fn):
        out.append((label, fn))
    a("n_photons_line 0", lambda d: setattr(d, "n_photons_line", 0))
    a("n_photons_interval -1", lambda d: setattr(d, "n_photons_interval", -1))
    a("n_interactions 0", lambda d: setattr(d, "n_interactions_trajectory", 0))
    a("empty outputfile", lambda d: setattr(d, "outputfile", ""))
    a("reference layer 0", lambda d: setattr(d, "reference_layer", 0))
    a("reference layer beyond", lambda d: setattr(d, "reference_layer", len(d.layers) + 1))
    a("Z 95", lambda d: d.layers[0].Z.__setitem__(0, 95))
    a("Z 0", lambda d: d.layers[-1].Z.__setitem__(0, 0))
    a("weight > 1", lambda d: d.layers[0].weight.__setitem__(0, 1.5))
    a("weight < 0", lambda d: d.layers[0].weight.__setitem__(0, -0.1))
    a("density 0", lambda d: setattr(d.layers[0], "density", 0.0))
    a("thickness -1", lambda d: setattr(d.layers[-1], "thickness", -1.0))
    a("sample normal z 0", lambda d: d.n_sample_orientation.__setitem__(2, 0.0))
    a("sample normal z < 0", lambda d: d.n_sample_orientation.__setitem__(2, -0.5))
    for f in ("d_sample_source", "area_detector", "d_source_slit", "slit_size_x", "slit_size_y"):
        a(f + " 0", lambda d, f=f: setattr(d, f, 0.0))
    a("collimator_height < 0", lambda d: setattr(d, "collimator_height", -1.0))
    a("collimator_height 0", lambda d: setattr(d, "collimator_height", 0.0))
    a("collimator_diameter < 0", lambda d: setattr(d, "collimator_diameter", -1.0))
    a("no source", lambda d: setattr(d, "discrete", []))
    a("one continuous point", lambda d: setattr(d, "continuous", [x.ContinuousD(5.0, 1.0, 1.0)]))
    a("continuum only", lambda d: (setattr(d, "discrete", []), setattr(d, "continuous", [x.ContinuousD(5.0, 1.0, 1.0), x.ContinuousD(9.0, 1.0, 1.0)])))
    a("two dark points", lambda d: setattr(d, "continuous", [x.ContinuousD(5.0, 0.0, 0.0), x.ContinuousD(9.0, 0.0, 0.0)]))
    a("dark start", lambda d: setattr(d, "continuous", [x.ContinuousD(e, i, i) for e, i in ((5, 0), (6, 0), (7, 1), (8, 1))]))
    a("dark end", lambda d: setattr(d, "continuous", [x.ContinuousD(e, i, i) for e, i in ((5, 1), (6, 1), (7, 0), (8, 0))]))
    a("dark middle", lambda d: setattr(d, "continuous", [x.ContinuousD(e, i, i) for e, i in ((5, 1), (6, 0), (7, 0), (8, 0), (9, 1))]))
    a("single dark points", lambda d: setattr(d, "continuous", [x.ContinuousD(e, i, i) for e, i in ((5, 1), (6, 0), (7, 1), (8, 0), (9, 1))]))
    a("three points dark pair", lambda d: setattr(d, "continuous", [x.ContinuousD(e, i, i) for e, i in ((5, 0), (6, 0), (7, 1))]))
    a("continuous energy < 0", lambda d: setattr(d, "continuous", [x.ContinuousD(-1.0, 1.0, 1.0), x.ContinuousD(9.0, 1.0, 1.0)]))
    a("continuous intensity < 0", lambda d: setattr(d, "continuous", [x.ContinuousD(5.0, -1.0, 2.0), x.ContinuousD(9.0, 1.0, 1.0)]))
    a("line energy 0", lambda d: setattr(d.discrete[0], "energy", 0.0))
    a("line intensity < 0", lambda d: setattr(d.discrete[0], "horizontal_intensity", -1.0))
    a("line dark", lambda d: (setattr(d.discrete[0], "horizontal_intensity", 0.0), setattr(d.discrete[0], "vertical_intensity", 0.0)))
    a("line sigma_x < 0", lambda d: setattr(d.discrete[0], "sigma_x", -1.0))
    a("line sigma_y < 0", lambda d: setattr(d.discrete[0], "sigma_y", -1.0))
    a("line distribution 3", lambda d: setattr(d.discrete[0], "distribution_type", 3))
    a("gaussian line without width", lambda d: (setattr(d.discrete[0], "distribution_type", 1), setattr(d.discrete[0], "scale_parameter", 0.0)))
    a("gaussian line", lambda d: (setattr(d.discrete[0], "distribution_type", 1), setattr(d.discrete[0], "scale_parameter", 0.1)))
    a("detector absorber density 0", lambda d: setattr(d.det_layers[0], "density", 0.0))
    a("excitation absorber Z 0", lambda d: setattr(d, "exc_layers", [x.LayerD([0], [1.0], 2.7, 0.01)]))
    a("excitation absorber weight 2", lambda d: setattr(d, "exc_layers", [x.LayerD([13], [2.0], 2.7, 0.01)]))
    for f in ("live_time", "pulse_width", "gain", "fano", "noise"):
        a(f + " 0", lambda d, f=f: setattr(d, f, 0.0))
    a("nchannels 9", lambda d: setattr(d, "nchannels", 9))
    a("nchannels 10", lambda d: setattr(d, "nchannels", 10))
    a("no crystal", lambda d: setattr(d, "crystal_layers", []))
    a("crystal thickness 0", lambda d: setattr(d.crystal_layers[0], "thickness", 0.0))
    a("two sections", lambda d: (setattr(d, "gain", 0.0), setattr(d, "n_photons_line", 0), d.layers[0].Z.__setitem__(0, 99)))
    return out


@pytest.mark.parametrize("name", ["srm1155", "srm1412"])
def test_input_validation_equals_the_reference_function(name):
    """xmb_input_validate (host_io.cpp) against the reference's xmi_input_validate (src/xmi_data_structs.c:899-1255,
    compiled from /root/reference into oracle/_ref): the same XmiInputFlags for ~60 inputs, each broken in one clause."""
    import copy
    import ctypes as C
    import xmimsim_b200 as x
    from xmimsim_b200 import abi
    from inputs import example
    if not hasattr(ref.lib(), "ref_input_validate"):
        pytest.skip("oracle/_ref built without the validate shim")
    base = example(name)
    flags_seen = set()
    for label, mutate in _invalid_variants(base):
        d = copy.deepcopy(base)
        mutate(d)
        ci = x.CInput(d)
        ours = abi.lib().xmb_input_validate(C.byref(ci.input))
        want = ref.input_validate(C.pointer(ci.input))
        assert ours == want, (label, ours, want)
        flags_seen.add(want)
    assert {0, 1, 2, 4, 8, 16, 32} <= flags_seen and any(f not in (0, 1, 2, 4, 8, 16, 32) for f in flags_seen), flags_seen


def test_abi_struct_layouts_equal_the_reference_headers():
    """include/xmimsim_b200.h against the reference's include/xmi_data_structs.h, xmi_solid_angle.h and xmi_detector.h:
    oracle/ref_shim/ref_layout.c holds a _Static_assert per struct (size) and per field (offset and size) of the 14 structs that
    cross the C ABI, plus the enumeration values; oracle/build_ref.sh fails on a mismatch, so a loadable oracle/_ref with
    the expected number of checks is the pass."""
    if not hasattr(ref.lib(), "ref_layout_checks"):
        pytest.skip("oracle/_ref built without the layout checks")
    assert ref.layout_checks() == 103
    # and the ctypes mirror the tests use has the sizes of the reference's structs
    import ctypes as C
    from xmimsim_b200 import abi
    sizes = (C.c_int * 13)()
    assert ref.lib().ref_struct_sizes(sizes) == 13
    mirror = (abi.General, abi.Layer, abi.Composition, abi.Geometry, abi.EnergyDiscrete, abi.EnergyContinuous, abi.Excitation,
              abi.Absorbers, abi.Detector, abi.Input, abi.MainOptions, abi.SolidAngle, abi.EscapeRatios)
    assert [C.sizeof(m) for m in mirror] == list(sizes)


def test_default_options_equal_the_reference_defaults():
    """xmb_main_options_defaults and xmb_get_default_escape_ratios_options against the reference's own initialisers
    (src/xmi_data_structs.c:2531-2547 -- omp_num_threads is set to the host's thread count afterwards, :2561 --,
    src/xmi_detector.c:643-646), compiled into oracle/_ref."""
    import ctypes as C
    import xmimsim_b200 as x
    from xmimsim_b200 import abi
    if not hasattr(ref.lib(), "ref_default_main_options"):
        pytest.skip("oracle/_ref built without the defaults")
    want = abi.MainOptions()
    ref.lib().ref_default_main_options(C.byref(want))
    ours = x.main_options()
    for f, _ in abi.MainOptions._fields_:
        if f == "omp_num_threads":
            assert getattr(ours, f) >= 1
        else:
            assert getattr(ours, f) == getattr(want, f), f
    want_e = abi.EscapeRatiosOptions()
    ref.lib().ref_default_escape_ratios_options(C.byref(want_e))
    ours_e = abi.lib().xmb_get_default_escape_ratios_options()
    assert [getattr(ours_e, f) for f, _ in abi.EscapeRatiosOptions._fields_] == [getattr(want_e, f) for f, _ in abi.EscapeRatiosOptions._fields_]
    assert want_e.n_input_energies == 1990 and want_e.n_photons == 500000
