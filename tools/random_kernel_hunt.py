#!/usr/bin/env python3
"""The other Monte Carlo kernels against the oracle on random inputs (tests/random_inputs.py), one line per seed:
  sa    solid-angle sub-grid (32 x 24 points of the real axes, 500 rays): integer hit counts point for point
  brute brute-force mode (detector brought close): hit / interaction / offspring counters, spectra within 2 photon weights
  adv   shell-resolved Compton with forced detection: spectra within the history tolerance (2e-6 of the maximum)
  esc   escape-ratio Monte Carlo on random detector crystals: same non-zero bins, ratios within 1e-9 of their maximum
  det   detector response (efficiency, escape peaks, Gaussian + tail convolution) on random detectors and spectra: 1e-11
usage: tools/random_kernel_hunt.py WHAT FIRST LAST   (needs a GPU)"""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import orc  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from helpers import Pair, DEFAULT_SEED  # noqa: E402
from random_inputs import random_input  # noqa: E402


def hunt_sa(seed):
    inp, _ = random_input(seed)
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    r_full, t_full = sim.solid_angle_inputs()
    rng = np.random.default_rng(seed)
    ri = np.unique(np.concatenate([np.arange(0, 8), rng.integers(8, 1024, 24)]))
    ti = np.unique(np.concatenate([np.arange(0, 6), rng.integers(6, 1024, 18)]))
    r, t = r_full[ri], t_full[ti]
    hps = 500 + seed % 2
    sa_g, hits_g = sim.solid_angle_grid(r, t, hits_per_single=hps, seed=seed + 1)
    sa_o, hits_o = orc.solid_angle_grid(od, r, np.arange(r.size), t, np.arange(t.size), r.size, hps, seed + 1)
    d = hits_g.astype(np.int64) - hits_o
    same = d == 0
    rel = float(np.abs(sa_g[same] / np.where(sa_o[same] == 0, 1, sa_o[same]) - 1)[sa_o[same] > 0].max()) if (sa_o[same] > 0).any() else 0.0
    sim.close()
    ok = np.count_nonzero(d) <= 2 and np.abs(d).max() <= 1 and rel < 1e-12
    return ok, "coll h %.2f d %.2f  points %d  rays hit %d  points differing %d (max %d)  rel %.1e" % (
        inp.collimator_height, inp.collimator_diameter, d.size, int(hits_o.sum()), np.count_nonzero(d), np.abs(d).max(), rel)


def hunt_brute(seed):
    inp, opts = random_input(seed, n_photons=150000)
    inp.p_detector_window = [0.0, -min(-inp.p_detector_window[1], 1.5), 100.0]
    inp.area_detector = max(inp.area_detector, 0.8) * 2.5
    if inp.collimator_height > 0:
        inp.collimator_height = min(inp.collimator_height, 0.3)
        inp.collimator_diameter = 0.8 * 2.0 * (inp.area_detector / np.pi) ** 0.5
    inp.continuous = []
    inp.n_interactions_trajectory = min(inp.n_interactions_trajectory, 3)
    o = x.main_options(use_variance_reduction=0, **opts)
    P = Pair(inp)
    ch, br, vr = P.sim.main_msim(o, None)
    cnt = P.sim.brute_counters()
    T = P.sim.L.xmb_get_tables(P.sim.hdf5F)
    ch_o, br_o, cnt_o = orc.main_msim_brute_range(C.pointer(P.ci.input), P.od, T, o, DEFAULT_SEED, 0, P.n_total,
                                                  inp.n_interactions_trajectory, inp.nchannels, 16)
    ch_o = ch_o * inp.live_time; br_o = br_o * inp.live_time
    w = max((d.horizontal_intensity + d.vertical_intensity) / inp.n_photons_line for d in inp.discrete) * inp.live_time
    dh, di, do = cnt["hits"] - int(cnt_o[0]), cnt["interactions"] - int(cnt_o[1]), cnt["offspring"] - int(cnt_o[2])
    dch = float(np.abs(ch - ch_o).max() / w); dbr = float(np.abs(br - br_o).max() / w)
    P.close()
    ok = abs(dh) <= 2 and abs(do) <= 2 and abs(di) <= 8 and dch <= 2.000002 and dbr <= 2.000002 and cnt["no_slot"] == 0
    return ok, "hits %d (diff %d)  interactions diff %d  offspring %d (diff %d)  max |channel diff| %.2f w  |history diff| %.2f w" % (
        cnt["hits"], dh, di, cnt["offspring"], do, dch, dbr)


def hunt_adv(seed):
    inp, opts = random_input(seed, n_photons=800)
    o = x.main_options(use_advanced_compton=1, **opts)
    P = Pair(inp)
    sa = P.grid(hits_per_single=300, n=64)
    ch, br, vr = P.sim.main_msim(o, sa)
    ch_o, vr_o, cnt = P.oracle(o, sa, 0)
    P.close()
    e_ch = float(np.abs(ch - ch_o).max() / max(np.abs(ch_o).max(), 1e-300))
    e_vr = float(np.abs(vr - vr_o).max() / max(np.abs(vr_o).max(), 1e-300))
    return bool(np.isfinite(ch).all() and e_ch <= 2e-6 and e_vr <= 2e-6), "err channels %.2e  history %.2e" % (e_ch, e_vr)


CRYSTALS = [([14], [1.0], 2.33), ([32], [1.0], 5.32), ([31, 33], [0.48, 0.52], 5.32), ([48, 52], [0.47, 0.53], 5.85),
            ([48, 30, 52], [0.43, 0.03, 0.54], 5.78)]


def hunt_esc(seed):
    from xmimsim_b200.xmsi import LayerD
    from inputs import example
    rng = np.random.default_rng(5000 + seed)
    inp = example("srm1155")
    layers = []
    for _ in range(int(rng.integers(1, 3))):
        z, w, rho = CRYSTALS[int(rng.integers(0, len(CRYSTALS)))]
        layers.append(LayerD(list(z), list(w), rho, float(10 ** rng.uniform(-3, -0.3))))
    inp.crystal_layers = layers
    sim = x.Simulation(inp)
    ero = sim.escape_ratios_options(n_input_energies=int(rng.integers(6, 14)), n_photons=20000,
                                    input_energy_min=float(rng.uniform(1.2, 6.0)), input_energy_delta=float(rng.uniform(1.5, 4.0)),
                                    n_compton_output_energies=int(rng.integers(200, 500)))
    ein, eh = sim.escape_ratios_handles(ero)
    L = sim.L
    cin = L.xmb_input_F2C(ein)
    od = orc.init_input(cin)
    T = L.xmb_get_tables(eh)
    fluo_o, compt_o = orc.escape_ratios(cin, od, T, seed + 3, ero.n_input_energies, T.contents.nZ, ero.n_photons,
                                        ero.n_compton_output_energies, ero.compton_output_energy_min,
                                        ero.compton_output_energy_delta, 16)
    er = sim.escape_ratios_run(ein, eh, ero, seed=seed + 3)
    Z, fluo, e_in, compt, e_out = sim.escape_ratios_arrays(er)
    fluo = fluo.copy(); compt = compt.copy()
    sim.escape_ratios_free(er)
    same_bins = bool(np.array_equal(fluo > 0, fluo_o > 0) and np.array_equal(compt > 0, compt_o > 0))
    ef = float(np.abs(fluo - fluo_o).max() / max(fluo_o.max(), 1e-300))
    ec = float(np.abs(compt - compt_o).max() / max(compt_o.max(), 1e-300))
    L.xmb_free_hdf5_F(C.byref(eh)); L.xmb_free_input_F(C.byref(ein)); sim.close()
    ok = same_bins and ef <= 1e-9 and ec <= 1e-9 and fluo_o.sum() > 0
    return ok, "crystal %s  energies %d  same bins %s  err fluo %.1e  compton %.1e" % (
        "+".join("".join(str(z) + "." for z in l.Z) for l in layers), ero.n_input_energies, same_bins, ef, ec)


def hunt_det(seed):
    from inputs import example
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_detector_gpu import synthetic_escape_ratios
    rng = np.random.default_rng(7000 + seed)
    inp = example("srm1155")
    inp.detector_type = int(rng.integers(0, 3))
    inp.nchannels = int(rng.choice([512, 1024, 2048, 4096]))
    inp.gain = float(40.0 * rng.uniform(0.5, 1.2) / inp.nchannels)
    inp.zero = float(rng.uniform(-0.05, 0.05))
    inp.fano = float(rng.uniform(0.08, 0.15)); inp.noise = float(rng.uniform(0.03, 0.2))
    inp.n_interactions_trajectory = int(rng.integers(1, 5))
    if rng.random() < 0.5:
        inp.det_layers = []
    escape = rng.random() < 0.6
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    n_int = inp.n_interactions_trajectory
    ch = np.zeros((n_int + 1, inp.nchannels))
    e = inp.zero + inp.gain * np.arange(inp.nchannels)
    for k in range(1, n_int + 1):
        row = 1e3 * np.exp(-e / 10.0) * rng.uniform(0.5, 1.5, inp.nchannels)
        for _ in range(6):
            c = int(rng.integers(20, inp.nchannels - 20))
            row[c] += 10 ** rng.uniform(4, 7)
        ch[k] = ch[k - 1] + row
    er = synthetic_escape_ratios(sim) if escape else None
    o = x.main_options(use_escape_peaks=1 if escape else 0)
    ch_gpu = ch.copy()
    conv = sim.detector_convolute_all(ch_gpu, None, None, o, er)
    worst = 0.0
    for k in range(1, n_int + 1):
        row, ref = orc.detector_convolute_spectrum(C.pointer(ci.input), ch[k].copy(), o, er, k)
        worst = max(worst, float(np.abs(conv[k] - ref).max() / ref.max()), float(np.abs(ch_gpu[k] - row).max() / row.max()))
    sim.close()
    return worst <= 1e-11, "type %d  channels %d  gain %.4f  escape %d  rows %d  worst %.1e" % (
        inp.detector_type, inp.nchannels, inp.gain, escape, n_int, worst)


def main():
    what, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    fn = {"sa": hunt_sa, "brute": hunt_brute, "adv": hunt_adv, "esc": hunt_esc, "det": hunt_det}[what]
    bad = []
    for seed in range(first, last):
        t0 = time.time()
        try:
            ok, note = fn(seed)
        except Exception as exc:   # an input the product rejects is reported, not hidden
            ok, note = False, "EXCEPTION %r" % (exc,)
        if not ok:
            bad.append(seed)
        print("%s seed %4d  %s  %s  %.1fs" % (what, seed, note, "ok" if ok else "MISMATCH", time.time() - t0), flush=True)
    print("%s seeds %d..%d: mismatches: %s" % (what, first, last - 1, bad or "none"))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
