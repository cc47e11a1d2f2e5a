"""Shared helpers for the GPU parity tests: run the product and the oracle on the same input."""
import ctypes as C

import numpy as np

import orc
import xmimsim_b200 as x


DEFAULT_SEED = 0x584D494D53494D      # "XMIMSIM" 


class Pair:
    def __init__(self, inp, quality=0):
        self.inp = inp
        self.sim = x.Simulation(inp, quality=quality)
        self.ci = x.CInput(inp)
        self.od = orc.init_input(C.pointer(self.ci.input))
        self.n_total = orc.lib().orc_total_histories(C.cast(C.pointer(self.ci.input), C.c_void_p))

    def grid(self, hits_per_single=500, seed=3, n=None):
        """Solid-angle grid from the GPU kernel (full 1024^2 axes unless n is given)."""
        if n is None:
            g, r, t = self.sim.solid_angle_calculation(hits_per_single=hits_per_single, seed=seed)
            return self.sim.make_solid_angle(g.copy(), r.copy(), t.copy())
        r_full, t_full = self.sim.solid_angle_inputs()
        r = np.linspace(r_full[0], r_full[-1], n)
        t = np.linspace(t_full[0], t_full[-1], n)
        g, _ = self.sim.solid_angle_grid(r, t, hits_per_single=hits_per_single, seed=seed)
        return self.sim.make_solid_angle(g, r, t)

    def oracle(self, options, sa, seed, g0=0, g1=None, n_threads=16):
        g1 = self.n_total if g1 is None else g1
        seed = seed or DEFAULT_SEED          # the engine maps seed 0 to its default Philox key
        ch, vr, cnt = orc.main_msim_range(C.pointer(self.ci.input), self.od, self.sim.L.xmb_get_tables(self.sim.hdf5F),
                                          options, sa, seed, g0, g1, self.inp.n_interactions_trajectory,
                                          self.inp.nchannels, n_threads)
        return ch * self.inp.live_time, vr * self.inp.live_time, cnt

    def close(self):
        self.sim.close()


def assert_spectra_close(a, b, rtol, what):
    """Element-wise agreement relative to the array's scale: |a-b| <= rtol * max(|b|) per row."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.abs(b).max()
    if scale == 0:
        assert np.abs(a).max() == 0, what
        return 0.0
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max |diff| / max|ref| = %.3e > %.1e" % (what, err, rtol)
    return err
