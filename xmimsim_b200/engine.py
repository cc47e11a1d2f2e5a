"""Host-side mirror of the reference's call sequence around the hot path (bin/xmimsim.c:262-526):

    xmi_input_C2F -> xmi_init_input -> xmi_init_from_hdf5/xmi_update_input_from_hdf5
    -> xmi_solid_angle_calculation -> xmi_main_msim -> xmi_detector_convolute_all

Every method is a thin ctypes call into libxmimsim_b200.so with the reference's argument meaning.
"""
import ctypes as C

import numpy as np

from . import abi
from .xmsi import CInput, InputD


def main_options(**kw):
    """xmi_main_options_new defaults (src/xmi_data_structs.c:2531-2565) with overrides."""
    o = abi.MainOptions()
    abi.lib().xmb_main_options_defaults(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _np_from_ptr(ptr, shape, dtype=np.float64):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = np.ctypeslib.as_array(ptr, shape=(n,))
    return buf.reshape(shape)


def _c_layer(l, keep):
    z = (C.c_int * len(l.Z))(*l.Z)
    w = (C.c_double * len(l.weight))(*l.weight)
    keep += [z, w]
    return abi.Layer(len(l.Z), C.cast(z, abi.c_int_p), C.cast(w, abi.c_double_p), l.density, l.thickness)


def tube_ebel(anode, voltage, current=1.0, angle_electron=60.0, angle_xray=60.0, delta_energy=0.1, solid_angle=1e-4,
              window=None, filt=None, transmission=False, eff_energies=None, efficiencies=None, provider=None):
    """xmi_tube_ebel: returns (continuous[(E, I_h, I_v)], discrete[(E, I_h, I_v)]) as two numpy arrays; anode / window /
    filt are LayerD.  Feed the result into InputD.continuous / .discrete (ContinuousD / DiscreteD)."""
    L = abi.lib()
    keep = []
    la = _c_layer(anode, keep)
    lw = _c_layer(window, keep) if window is not None else None
    lf = _c_layer(filt, keep) if filt is not None else None
    n_eff, pe, pv = 0, None, None
    if eff_energies is not None:
        ee = np.ascontiguousarray(eff_energies, np.float64); ev = np.ascontiguousarray(efficiencies, np.float64)
        keep += [ee, ev]
        n_eff, pe, pv = ee.size, ee.ctypes.data_as(abi.c_double_p), ev.ctypes.data_as(abi.c_double_p)
    out = C.POINTER(abi.Excitation)()
    ok = L.xmb_tube_ebel(provider, C.byref(la), C.byref(lw) if lw is not None else None, C.byref(lf) if lf is not None else None,
                         voltage, current, angle_electron, angle_xray, delta_energy, solid_angle, int(bool(transmission)),
                         n_eff, pe, pv, C.byref(out))
    if not ok:
        raise RuntimeError("xmb_tube_ebel: " + abi.last_error())
    e = out.contents
    cont = np.array([(e.continuous[i].energy, e.continuous[i].horizontal_intensity, e.continuous[i].vertical_intensity)
                     for i in range(e.n_continuous)]).reshape(-1, 3)
    disc = np.array([(e.discrete[i].energy, e.discrete[i].horizontal_intensity, e.discrete[i].vertical_intensity)
                     for i in range(e.n_discrete)]).reshape(-1, 3)
    L.xmb_free_excitation(C.byref(out))
    return cont, disc


class Comm:
    """NCCL communicator of the engine, one rank per GPU.  Rank 0 creates the id (Comm.unique_id()), the launcher's own
    channel carries the 128 bytes to the other ranks (torch.distributed / MPI), every rank builds Comm(id, rank, n)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(abi.COMM_ID_BYTES)
        if not abi.lib().xmb_comm_unique_id(buf):
            raise RuntimeError("xmb_comm_unique_id: " + abi.last_error())
        return buf.raw

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int = -1):
        assert len(unique_id) == abi.COMM_ID_BYTES
        self.handle = C.c_void_p()
        if not abi.lib().xmb_comm_init_rank(unique_id, rank, n_ranks, device, C.byref(self.handle)):
            raise RuntimeError("xmb_comm_init_rank: " + abi.last_error())
        self.rank, self.n_ranks = rank, n_ranks

    def close(self):
        if self.handle:
            abi.lib().xmb_comm_free(C.byref(self.handle))


class Simulation:
    """Owns the two opaque handles (input + tables) of one simulation."""

    def __init__(self, inp: InputD, quality: int = 0, provider=None, gpu_tables: bool = False):
        L = abi.lib()
        self.L = L
        self.inp = inp
        self.cinput = CInput(inp)
        self.inputF = C.c_void_p()
        if not L.xmb_input_C2F(C.byref(self.cinput.input), C.byref(self.inputF)):
            raise RuntimeError("xmb_input_C2F: " + abi.last_error())
        if not L.xmb_init_input(C.byref(self.inputF)):
            raise RuntimeError("xmb_init_input: " + abi.last_error())
        self.provider = provider if provider is not None else L.xmb_xrl_surrogate()
        self.hdf5F = C.c_void_p()
        init = L.xmb_init_from_provider_gpu if gpu_tables else L.xmb_init_from_provider
        if not init(self.provider, self.inputF, quality, C.byref(self.hdf5F)):
            raise RuntimeError("xmb_init_from_provider: " + abi.last_error())
        self._sa = None

    # -- views ---------------------------------------------------------------------------------
    @property
    def derived(self):
        return self.L.xmb_get_derived(self.inputF).contents

    @property
    def tables(self):
        return self.L.xmb_get_tables(self.hdf5F).contents

    def c_input(self):
        """The handle's own (normalised) C tree, as xmi_input_F2C would return it."""
        return self.L.xmb_input_F2C(self.inputF)

    # -- solid angle -----------------------------------------------------------------------------
    def solid_angle_inputs(self):
        sa = C.POINTER(abi.SolidAngle)()
        if not self.L.xmb_solid_angle_inputs(self.inputF, self.hdf5F, C.byref(sa)):
            raise RuntimeError("xmb_solid_angle_inputs: " + abi.last_error())
        r = _np_from_ptr(sa.contents.grid_dims_r_vals, (sa.contents.grid_dims_r_n,)).copy()
        t = _np_from_ptr(sa.contents.grid_dims_theta_vals, (sa.contents.grid_dims_theta_n,)).copy()
        self.L.xmb_free_solid_angle(sa)
        return r, t

    def solid_angle_calculation(self, options=None, hits_per_single=5000, seed=0):
        """xmi_solid_angle_calculation (GPU backend).  Returns (grid[theta][r], r_vals, theta_vals)."""
        options = options or main_options()
        sa = C.POINTER(abi.SolidAngle)()
        if not self.L.xmb_solid_angle_calculation(self.inputF, self.hdf5F, C.byref(sa), None, C.byref(options),
                                                  hits_per_single, seed):
            raise RuntimeError("xmb_solid_angle_calculation: " + abi.last_error())
        if self._sa is not None:
            self.L.xmb_free_solid_angle(self._sa)
        self._sa = sa
        s = sa.contents
        grid = _np_from_ptr(s.solid_angles, (s.grid_dims_theta_n, s.grid_dims_r_n))
        return grid, _np_from_ptr(s.grid_dims_r_vals, (s.grid_dims_r_n,)), _np_from_ptr(s.grid_dims_theta_vals, (s.grid_dims_theta_n,))

    def solid_angle_grid(self, r_vals, theta_vals, hits_per_single=5000, seed=0, verbose=0):
        """Grid over caller-given axes.  Returns (solid_angles[theta][r], hits[theta][r])."""
        r = np.ascontiguousarray(r_vals, dtype=np.float64)
        t = np.ascontiguousarray(theta_vals, dtype=np.float64)
        out = np.zeros((t.size, r.size), dtype=np.float64)
        hits = np.zeros((t.size, r.size), dtype=np.int32)
        ok = self.L.xmb_solid_angle_grid(self.inputF, r.ctypes.data_as(abi.c_double_p), r.size,
                                         t.ctypes.data_as(abi.c_double_p), t.size, hits_per_single, seed, verbose,
                                         out.ctypes.data_as(abi.c_double_p), hits.ctypes.data_as(C.POINTER(C.c_int32)))
        if not ok:
            raise RuntimeError("xmb_solid_angle_grid: " + abi.last_error())
        return out, hits

    def solid_angle_struct(self):
        return self._sa

    def make_solid_angle(self, grid, r_vals, theta_vals):
        """Wrap numpy arrays (grid[theta][r]) as an xmi_solid_angle struct (borrowed memory)."""
        g = np.ascontiguousarray(grid, np.float64)
        r = np.ascontiguousarray(r_vals, np.float64)
        t = np.ascontiguousarray(theta_vals, np.float64)
        sa = abi.SolidAngle(g.ctypes.data_as(abi.c_double_p), r.size, t.size, r.ctypes.data_as(abi.c_double_p),
                            t.ctypes.data_as(abi.c_double_p), None)
        sa._keep = (g, r, t)
        return sa

    # -- photon histories ----------------------------------------------------------------------------
    def _sa_arg(self, sa, options=None):
        if sa is None:
            if options is not None and not options.use_variance_reduction:
                return None                      # brute-force mode needs no solid-angle grid
            if self._sa is None:
                raise RuntimeError("no solid-angle grid: call solid_angle_calculation first")
            return self._sa
        return C.pointer(sa) if isinstance(sa, abi.SolidAngle) else sa

    def main_msim(self, options=None, sa=None, n_mpi_hosts=1):
        """xmi_main_msim.  Returns (channels[(n_int+1)][nch], brute_history[100][385][n_int],
        var_red_history[100][385][n_int]) -- the reference's three output arrays, x live_time."""
        options = options or main_options()
        ch, br, vr = abi.c_double_p(), abi.c_double_p(), abi.c_double_p()
        if not self.L.xmb_main_msim(self.inputF, self.hdf5F, n_mpi_hosts, C.byref(ch), C.byref(options), C.byref(br),
                                    C.byref(vr), self._sa_arg(sa, options)):
            raise RuntimeError("xmb_main_msim: " + abi.last_error())
        return self._take(ch, br, vr)

    def _take(self, ch, br, vr):
        n_int, nch = self.inp.n_interactions_trajectory, self.inp.nchannels
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        out = []
        for ptr, shape in ((ch, (n_int + 1, nch)), (br, (100, 385, n_int)), (vr, (100, 385, n_int))):
            if not ptr:                      # var_red_history is NULL without variance reduction, as in the reference
                out.append(np.zeros(shape))
                continue
            out.append(_np_from_ptr(ptr, shape).copy())
            libc.free(ptr)
        return tuple(out)

    def main_msim_device(self, options=None, sa=None, rank=0, n_ranks=1, seed=0, device=-1):
        """History kernels with every input and output resident in HBM (the solid-angle grid is uploaded
        only when its host buffer changes).  Returns MsimEx; read the limbs with device_limbs()."""
        options = options or main_options()
        ex = abi.MsimEx(rank, n_ranks, seed, device, 1, 0, 0.0, 0, 0)
        acc = C.POINTER(C.c_uint64)()
        n = C.c_size_t()
        if not self.L.xmb_main_msim_raw(self.inputF, self.hdf5F, C.byref(options), self._sa_arg(sa, options), C.byref(ex),
                                        C.byref(acc), C.byref(n)):
            raise RuntimeError("xmb_main_msim_raw: " + abi.last_error())
        return ex

    # -- multi-GPU (multi_gpu.cu): shards + one NCCL all-reduce of the integer histograms -----------------------
    def main_msim_multi(self, comm, options=None, sa=None):
        """xmi_main_msim over a communicator (Comm): the three output arrays, summed over all ranks, on every rank."""
        options = options or main_options()
        ch, br, vr = abi.c_double_p(), abi.c_double_p(), abi.c_double_p()
        if not self.L.xmb_main_msim_multi(self.inputF, self.hdf5F, comm.handle, C.byref(ch), C.byref(options), C.byref(br),
                                          C.byref(vr), self._sa_arg(sa, options)):
            raise RuntimeError("xmb_main_msim_multi: " + abi.last_error())
        return self._take(ch, br, vr)

    def main_msim_multi_device(self, comm, options=None, sa=None, seed=0):
        """History kernel -> limbs -> ncclAllReduce on one stream, the reduced limbs left in HBM (device_limbs());
        the solid-angle grid is uploaded only when its content changed.  Returns MsimEx (this rank's counters)."""
        options = options or main_options()
        ex = abi.MsimEx(0, 1, seed, -1, 1, 0, 0.0, 0, 0)
        if not self.L.xmb_main_msim_multi_raw(self.inputF, self.hdf5F, C.byref(options), self._sa_arg(sa, options),
                                              comm.handle, C.byref(ex)):
            raise RuntimeError("xmb_main_msim_multi_raw: " + abi.last_error())
        return ex

    def main_msim_all_devices(self, n_devices=0, options=None, sa=None, seed=0, devices=None):
        """One process, n_devices GPUs (0: all visible).  Returns (channels, brute_history, var_red_history, MsimEx)."""
        options = options or main_options()
        ch, br, vr = abi.c_double_p(), abi.c_double_p(), abi.c_double_p()
        ex = abi.MsimEx(0, 1, seed, -1, 0, 0, 0.0, 0, 0)
        dv = (C.c_int * len(devices))(*devices) if devices else None
        if not self.L.xmb_main_msim_all_devices(self.inputF, self.hdf5F, len(devices) if devices else n_devices, dv, C.byref(ch),
                                                C.byref(options), C.byref(br), C.byref(vr), self._sa_arg(sa, options), C.byref(ex)):
            raise RuntimeError("xmb_main_msim_all_devices: " + abi.last_error())
        return self._take(ch, br, vr) + (ex,)

    def device_limbs(self):
        """(device pointer, n_words) of the last run's uint64 limbs."""
        ptr, n = C.c_void_p(), C.c_size_t()
        if not self.L.xmb_msim_device_limbs(self.hdf5F, C.byref(ptr), C.byref(n)):
            raise RuntimeError(abi.last_error())
        return ptr.value, n.value

    def workload_stats(self):
        """Per layer: (interactions of the last run, n_elements, active line records)."""
        buf = (C.c_uint64 * 128)()
        if not self.L.xmb_msim_workload_stats(self.inputF, self.hdf5F, buf, 128):
            raise RuntimeError(abi.last_error())
        nl = int(buf[0])
        return [(int(buf[1 + 3 * k]), int(buf[2 + 3 * k]), int(buf[3 + 3 * k])) for k in range(nl)]

    def brute_counters(self):
        """dict of the last brute-force run's counters."""
        buf = (C.c_uint64 * 8)()
        if not self.L.xmb_msim_brute_counters(self.hdf5F, buf, 8):
            raise RuntimeError(abi.last_error())
        return {"interactions": int(buf[1]), "hits": int(buf[3]), "offspring": int(buf[4]), "no_slot": int(buf[5])}

    def shard_count(self, rank, n_ranks):
        """Number of histories `rank` simulates (block-cyclic shard of the global photon ids)."""
        return int(self.L.xmb_msim_shard_count(self.L.xmb_msim_total_histories(self.inputF), rank, n_ranks))

    def shard_owner(self, photon_id, n_ranks):
        return int(self.L.xmb_msim_shard_owner(photon_id, n_ranks))

    def slot_map(self, options=None):
        """(Z[n_hist_slots], line[n_hist_slots]) of the history slots that follow the nch channel slots in a row."""
        options = options or main_options()
        n = self.L.xmb_msim_slot_map(self.inputF, self.hdf5F, C.byref(options), None, None, 0)
        if n <= 0:
            raise RuntimeError("xmb_msim_slot_map: " + abi.last_error())
        z = np.zeros(n, np.int32); ln = np.zeros(n, np.int32)
        self.L.xmb_msim_slot_map(self.inputF, self.hdf5F, C.byref(options), z.ctypes.data_as(C.POINTER(C.c_int32)),
                                 ln.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return z, ln

    def main_msim_raw(self, options=None, sa=None, rank=0, n_ranks=1, seed=0, device=-1):
        """History kernels only: returns (limbs uint64[2*n_slots], MsimEx) -- exact fixed-point partial sums of
        this rank's photon-id shard, safe to add across ranks in uint64."""
        options = options or main_options()
        ex = abi.MsimEx(rank, n_ranks, seed, device, 0, 0, 0.0, 0, 0)
        acc = C.POINTER(C.c_uint64)()
        n = C.c_size_t()
        if not self.L.xmb_main_msim_raw(self.inputF, self.hdf5F, C.byref(options), self._sa_arg(sa, options), C.byref(ex),
                                        C.byref(acc), C.byref(n)):
            raise RuntimeError("xmb_main_msim_raw: " + abi.last_error())
        limbs = np.ctypeslib.as_array(acc, shape=(2 * n.value,)).copy()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(acc)
        return limbs, ex

    def main_msim_finish(self, limbs, options=None):
        options = options or main_options()
        limbs = np.ascontiguousarray(limbs, np.uint64)
        ch, br, vr = abi.c_double_p(), abi.c_double_p(), abi.c_double_p()
        if not self.L.xmb_main_msim_finish(self.inputF, self.hdf5F, C.byref(options),
                                           limbs.ctypes.data_as(C.POINTER(C.c_uint64)), limbs.size // 2, C.byref(ch),
                                           C.byref(br), C.byref(vr)):
            raise RuntimeError("xmb_main_msim_finish: " + abi.last_error())
        return self._take(ch, br, vr)

    # -- detector response ----------------------------------------------------------------------------
    def make_escape_ratios(self, Z, fluo_ratios, fluo_E, compt_ratios, compt_Ein, compt_Eout):
        """xmi_escape_ratios struct from numpy arrays: fluo_ratios[n_E][109][n_Z] (element fastest, the
        reference's Fortran (n_elements, 109, n_E) order), compt_ratios[n_out][n_in] (input fastest)."""
        zz = np.ascontiguousarray(Z, np.int32)
        a = [np.ascontiguousarray(v, np.float64) for v in (fluo_ratios, fluo_E, compt_ratios, compt_Ein, compt_Eout)]
        er = abi.EscapeRatios(zz.size, a[1].size, a[3].size, a[4].size, zz.ctypes.data_as(abi.c_int_p),
                              a[0].ctypes.data_as(abi.c_double_p), a[1].ctypes.data_as(abi.c_double_p),
                              a[2].ctypes.data_as(abi.c_double_p), a[3].ctypes.data_as(abi.c_double_p),
                              a[4].ctypes.data_as(abi.c_double_p), None)
        er._keep = (zz, a)
        return er

    def escape_ratios_options(self, **kw):
        """xmi_get_default_escape_ratios_options with overrides."""
        ero = self.L.xmb_get_default_escape_ratios_options()
        for k, v in kw.items():
            if not hasattr(ero, k):
                raise AttributeError(k)
            setattr(ero, k, v)
        return ero

    def escape_ratios_handles(self, ero, quality=0):
        """(inputF, hdf5F) of the escape-mode run (composition = crystal): for the parity tests, which hand the same
        host tables to the oracle.  Caller frees with xmb_free_hdf5_F / xmb_free_input_F."""
        ein, eh = C.c_void_p(), C.c_void_p()
        if not self.L.xmb_escape_ratios_input(C.byref(self.cinput.input), C.byref(ero), C.byref(ein)):
            raise RuntimeError("xmb_escape_ratios_input: " + abi.last_error())
        if not self.L.xmb_init_from_provider(self.provider, ein, quality, C.byref(eh)):
            raise RuntimeError("xmb_init_from_provider: " + abi.last_error())
        return ein, eh

    def escape_ratios_run(self, ein, eh, ero, seed=0):
        er = C.POINTER(abi.EscapeRatios)()
        if not self.L.xmb_escape_ratios_run(ein, eh, C.byref(ero), seed, C.byref(er), None):
            raise RuntimeError("xmb_escape_ratios_run: " + abi.last_error())
        return er

    def escape_ratios_calculation(self, ero=None, options=None, seed=0):
        """xmi_escape_ratios_calculation: returns a pointer to a malloc'ed xmi_escape_ratios
        (free with escape_ratios_free); `.contents` is what detector_convolute_all takes."""
        ero = ero or self.escape_ratios_options()
        options = options or main_options()
        er = C.POINTER(abi.EscapeRatios)()
        if not self.L.xmb_escape_ratios_calculation(C.byref(self.cinput.input), C.byref(er), None, self.provider,
                                                    C.byref(options), ero, seed):
            raise RuntimeError("xmb_escape_ratios_calculation: " + abi.last_error())
        return er

    @staticmethod
    def escape_ratios_arrays(er):
        """numpy copies of an xmi_escape_ratios: (Z, fluo[nE][109][nZ], E_in, compton[n_out][n_in], E_out)."""
        e = er.contents
        nZ, nE, nO = e.n_elements, e.n_fluo_input_energies, e.n_compton_output_energies
        Z = np.array([e.Z[i] for i in range(nZ)], np.int32)
        return (Z, _np_from_ptr(e.fluo_escape_ratios, (nE, 109, nZ)).copy(), _np_from_ptr(e.fluo_escape_input_energies, (nE,)).copy(),
                _np_from_ptr(e.compton_escape_ratios, (nO, nE)).copy(), _np_from_ptr(e.compton_escape_output_energies, (nO,)).copy())

    def escape_ratios_free(self, er):
        self.L.xmb_free_escape_ratios(C.byref(er))

    def detector_convolute_all(self, channels, brute_history=None, var_red_history=None, options=None, escape_ratios=None,
                               zero_interaction=0):
        """xmi_detector_convolute_all.  `channels` [(n_int+1)][nch] is modified in place (as the reference does);
        returns channels_conv [(n_int+1)][nch] (row 0 zero unless zero_interaction)."""
        options = options or main_options()
        n_int, nch = self.inp.n_interactions_trajectory, self.inp.nchannels
        assert channels.shape == (n_int + 1, nch) and channels.dtype == np.float64 and channels.flags.c_contiguous
        rows = (abi.c_double_p * (n_int + 1))()
        for i in range(n_int + 1):
            rows[i] = C.cast(channels.ctypes.data + i * nch * 8, abi.c_double_p)
        conv = (abi.c_double_p * (n_int + 1))()
        bh = brute_history.ctypes.data_as(abi.c_double_p) if brute_history is not None else None
        vh = var_red_history.ctypes.data_as(abi.c_double_p) if var_red_history is not None else None
        er = C.byref(escape_ratios) if escape_ratios is not None else None
        self.L.xmb_detector_convolute_all(self.inputF, self.hdf5F, rows, conv, bh, vh, C.byref(options), er, n_int,
                                          zero_interaction)
        out = np.zeros((n_int + 1, nch))
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for i in range(0 if zero_interaction else 1, n_int + 1):
            if not conv[i]:
                raise RuntimeError("xmb_detector_convolute_all: " + abi.last_error())
            out[i] = _np_from_ptr(conv[i], (nch,))
            libc.free(conv[i])
        return out

    def detector_convolute_spectrum(self, spectrum, options=None, escape_ratios=None, n_interactions=1):
        """xmi_detector_convolute_spectrum: spectrum[nch] modified in place, returns the convoluted spectrum."""
        options = options or main_options()
        assert spectrum.dtype == np.float64 and spectrum.flags.c_contiguous
        conv = abi.c_double_p()
        er = C.byref(escape_ratios) if escape_ratios is not None else None
        self.L.xmb_detector_convolute_spectrum(self.inputF, self.hdf5F, spectrum.ctypes.data_as(abi.c_double_p), C.byref(conv),
                                               C.byref(options), er, n_interactions)
        if not conv:
            raise RuntimeError("xmb_detector_convolute_spectrum: " + abi.last_error())
        out = _np_from_ptr(conv, (spectrum.size,)).copy()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(conv)
        return out

    def close(self):
        if self._sa is not None:
            self.L.xmb_free_solid_angle(self._sa)
            self._sa = None
        if self.hdf5F:
            self.L.xmb_free_hdf5_F(C.byref(self.hdf5F))
        if self.inputF:
            self.L.xmb_free_input_F(C.byref(self.inputF))
