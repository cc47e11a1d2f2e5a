"""ctypes mirror of include/xmimsim_b200.h and loader of the product library.

The structures have the reference's C layouts (include/xmi_data_structs.h:40-336, :479-495;
include/xmi_solid_angle.h:28-35; include/xmi_detector.h:28-40).  The product library is mandatory:
there is no Python or CPU fallback for any compute entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XMIMSIM_B200_LIB") or os.path.join(_HERE, "lib", "libxmimsim_b200.so")   # override: experiment builds

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class General(C.Structure):
    _fields_ = [("version", C.c_float), ("outputfile", C.c_char_p), ("n_photons_interval", C.c_long),
                ("n_photons_line", C.c_long), ("n_interactions_trajectory", C.c_int), ("comments", C.c_char_p)]


class Layer(C.Structure):
    _fields_ = [("n_elements", C.c_int), ("Z", c_int_p), ("weight", c_double_p), ("density", C.c_double),
                ("thickness", C.c_double)]


class Composition(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("layers", C.POINTER(Layer)), ("reference_layer", C.c_int)]


class Geometry(C.Structure):
    _fields_ = [("d_sample_source", C.c_double), ("n_sample_orientation", C.c_double * 3),
                ("p_detector_window", C.c_double * 3), ("n_detector_orientation", C.c_double * 3),
                ("area_detector", C.c_double), ("collimator_height", C.c_double), ("collimator_diameter", C.c_double),
                ("d_source_slit", C.c_double), ("slit_size_x", C.c_double), ("slit_size_y", C.c_double)]


class EnergyDiscrete(C.Structure):
    _fields_ = [("energy", C.c_double), ("horizontal_intensity", C.c_double), ("vertical_intensity", C.c_double),
                ("sigma_x", C.c_double), ("sigma_xp", C.c_double), ("sigma_y", C.c_double), ("sigma_yp", C.c_double),
                ("distribution_type", C.c_int), ("scale_parameter", C.c_double)]


class EnergyContinuous(C.Structure):
    _fields_ = [("energy", C.c_double), ("horizontal_intensity", C.c_double), ("vertical_intensity", C.c_double),
                ("sigma_x", C.c_double), ("sigma_xp", C.c_double), ("sigma_y", C.c_double), ("sigma_yp", C.c_double)]


class Excitation(C.Structure):
    _fields_ = [("n_discrete", C.c_int), ("discrete", C.POINTER(EnergyDiscrete)), ("n_continuous", C.c_int),
                ("continuous", C.POINTER(EnergyContinuous))]


class Absorbers(C.Structure):
    _fields_ = [("n_exc_layers", C.c_int), ("exc_layers", C.POINTER(Layer)), ("n_det_layers", C.c_int),
                ("det_layers", C.POINTER(Layer))]


class Detector(C.Structure):
    _fields_ = [("detector_type", C.c_int), ("live_time", C.c_double), ("pulse_width", C.c_double),
                ("gain", C.c_double), ("zero", C.c_double), ("fano", C.c_double), ("noise", C.c_double),
                ("nchannels", C.c_int), ("n_crystal_layers", C.c_int), ("crystal_layers", C.POINTER(Layer))]


class Input(C.Structure):
    _fields_ = [("general", C.POINTER(General)), ("composition", C.POINTER(Composition)),
                ("geometry", C.POINTER(Geometry)), ("excitation", C.POINTER(Excitation)),
                ("absorbers", C.POINTER(Absorbers)), ("detector", C.POINTER(Detector))]


class MainOptions(C.Structure):
    _fields_ = [("use_M_lines", C.c_int), ("use_cascade_auger", C.c_int), ("use_cascade_radiative", C.c_int),
                ("use_variance_reduction", C.c_int), ("use_sum_peaks", C.c_int), ("use_escape_peaks", C.c_int),
                ("escape_ratios_mode", C.c_int), ("verbose", C.c_int), ("use_poisson", C.c_int), ("use_gpu", C.c_int),
                ("omp_num_threads", C.c_int), ("extra_verbose", C.c_int), ("custom_detector_response", C.c_char_p),
                ("use_advanced_compton", C.c_int), ("use_default_seeds", C.c_int)]


class SolidAngle(C.Structure):
    _fields_ = [("solid_angles", c_double_p), ("grid_dims_r_n", C.c_long), ("grid_dims_theta_n", C.c_long),
                ("grid_dims_r_vals", c_double_p), ("grid_dims_theta_vals", c_double_p),
                ("xmi_input_string", C.c_void_p)]


class EscapeRatios(C.Structure):
    _fields_ = [("n_elements", C.c_int), ("n_fluo_input_energies", C.c_int), ("n_compton_input_energies", C.c_int),
                ("n_compton_output_energies", C.c_int), ("Z", c_int_p), ("fluo_escape_ratios", c_double_p),
                ("fluo_escape_input_energies", c_double_p), ("compton_escape_ratios", c_double_p),
                ("compton_escape_input_energies", c_double_p), ("compton_escape_output_energies", c_double_p),
                ("xmi_input_string", C.c_void_p)]


class EscapeRatiosOptions(C.Structure):
    _fields_ = [("n_input_energies", C.c_long), ("n_compton_output_energies", C.c_long), ("n_photons", C.c_long),
                ("input_energy_min", C.c_double), ("input_energy_delta", C.c_double),
                ("compton_output_energy_min", C.c_double), ("compton_output_energy_delta", C.c_double)]


class Derived(C.Structure):
    _fields_ = [("n_sample_orientation", C.c_double * 3), ("n_detector_orientation", C.c_double * 3),
                ("detector_radius", C.c_double), ("collimator_present", C.c_int), ("collimator_radius", C.c_double),
                ("collimator_height", C.c_double), ("half_apex", C.c_double), ("vertex", C.c_double * 3),
                ("ndo_new", C.c_double * 9), ("ndo_inv", C.c_double * 9), ("detector_solid_angle", C.c_double),
                ("n_sample_orientation_det", C.c_double * 3), ("n_layers", C.c_int),
                ("thickness_along_Z", c_double_p), ("Z_coord_begin", c_double_p), ("Z_coord_end", c_double_p)]


class TablesHost(C.Structure):
    _fields_ = [("nZ", C.c_int), ("Z", c_int_p), ("uniqZ", c_int_p), ("atomic_weight", c_double_p),
                ("n_nodes", C.c_int), ("node_E", c_double_p), ("bucket_E0", C.c_double), ("bucket_inv_dE", C.c_double),
                ("n_buckets", C.c_int), ("bucket_start", c_int_p),
                ("cs_total", c_double_p), ("cs_photo_total", c_double_p), ("p_rayl", c_double_p),
                ("p_rayl_compt", c_double_p), ("cs_photo_partial", c_double_p), ("cs_vacancy", c_double_p),
                ("n_icdf_E", C.c_int), ("n_icdf_R", C.c_int), ("icdf_E", c_double_p), ("icdf_R", c_double_p),
                ("rayl_theta_icdf", c_double_p), ("compt_theta_icdf", c_double_p),
                ("n_phi_T", C.c_int), ("phi_T", c_double_p), ("phi_icdf", c_double_p),
                ("n_cp", C.c_int), ("cp_R", c_double_p), ("cp_icdf", c_double_p),
                ("n_q", C.c_int), ("q_max", C.c_double), ("ff", c_double_p), ("sf", c_double_p),
                ("fluor_yield", c_double_p), ("fluor_yield_corr", c_double_p), ("cos_kron", c_double_p),
                ("rad_rate", c_double_p), ("line_energy", c_double_p), ("edge_energy", c_double_p),
                ("n_layers", C.c_int), ("mu_layer", c_double_p), ("exc_murhod", c_double_p),
                ("auger_rate", c_double_p),
                ("n_adv_rows", C.c_int), ("adv_off", c_int_p), ("adv_shell", c_int_p), ("adv_config", c_double_p),
                ("adv_edge", c_double_p), ("adv_cdf", c_double_p), ("adv_qinv", c_double_p)]


class XrlProvider(C.Structure):
    _D, _I = C.c_double, C.c_int
    _fields_ = [("name", C.c_char_p),
                ("AtomicWeight", C.CFUNCTYPE(_D, _I)), ("EdgeEnergy", C.CFUNCTYPE(_D, _I, _I)),
                ("LineEnergy", C.CFUNCTYPE(_D, _I, _I)), ("FluorYield", C.CFUNCTYPE(_D, _I, _I)),
                ("RadRate", C.CFUNCTYPE(_D, _I, _I)), ("CosKronTransProb", C.CFUNCTYPE(_D, _I, _I)),
                ("JumpFactor", C.CFUNCTYPE(_D, _I, _I)), ("CS_Total_Kissel", C.CFUNCTYPE(_D, _I, _D)),
                ("CS_Photo_Total", C.CFUNCTYPE(_D, _I, _D)), ("CS_Photo_Partial", C.CFUNCTYPE(_D, _I, _I, _D)),
                ("CS_Rayl", C.CFUNCTYPE(_D, _I, _D)), ("CS_Compt", C.CFUNCTYPE(_D, _I, _D)),
                ("FF_Rayl", C.CFUNCTYPE(_D, _I, _D)), ("SF_Compt", C.CFUNCTYPE(_D, _I, _D)),
                ("ComptonProfile", C.CFUNCTYPE(_D, _I, _D)),
                ("VacancyCS", C.CFUNCTYPE(_D, _I, _I, _D, _I, c_double_p)),
                ("AugerRate", C.CFUNCTYPE(_D, _I, _I, _I, _I)),
                ("ElectronConfig_Biggs", C.CFUNCTYPE(_D, _I, _I)), ("ComptonProfile_Partial", C.CFUNCTYPE(_D, _I, _I, _D))]


class MsimEx(C.Structure):
    _fields_ = [("rank", C.c_int), ("n_ranks", C.c_int), ("seed", C.c_uint64), ("device", C.c_int),
                ("keep_on_device", C.c_int), ("n_histories", C.c_uint64), ("kernel_ms", C.c_double),
                ("n_launches", C.c_uint64), ("n_interactions", C.c_uint64)]


COMM_ID_BYTES = 128     # XMB_COMM_ID_BYTES = sizeof(ncclUniqueId)

_lib = None


def declared_symbols():
    """Names declared in include/xmimsim_b200.h (parsed from the header itself)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "xmimsim_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(xm[bi]_[a-z0-9_A-Z]+)\s*\(", txt)) - {"xmb_discrete_distribution"})


def lib():
    """Load libxmimsim_b200.so (built by `make lib` / __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("xmimsim_b200: %s is missing -- run `make lib` (or __graft_entry__.build()); "
                           "there is no CPU fallback" % LIB_PATH)
    # NCCL is bound at run time by the library (multi_gpu.cu).  A Python host usually also hosts PyTorch, whose own
    # libnccl.so.2 must be THE libnccl of the process (the loader keys on the soname: whichever is loaded first wins, and
    # an older system NCCL loaded first breaks `import torch`): point the library at the pip-installed one when there is one.
    if "XMB_NCCL_LIBRARY" not in os.environ:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for d in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["XMB_NCCL_LIBRARY"] = cand
                    break
        except Exception:
            pass
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    L.xmb_xrl_surrogate.restype = C.POINTER(XrlProvider)
    L.xmb_xrl_surrogate_dense.restype = C.POINTER(XrlProvider)
    L.xmb_xrl_from_library.argtypes = [C.c_char_p]; L.xmb_xrl_from_library.restype = C.POINTER(XrlProvider)
    L.xmb_input_C2F.argtypes = [C.POINTER(Input), vpp]; L.xmb_input_C2F.restype = C.c_int
    L.xmb_input_F2C.argtypes = [vp]; L.xmb_input_F2C.restype = C.POINTER(Input)
    L.xmb_free_input_F.argtypes = [vpp]; L.xmb_free_input_F.restype = None
    L.xmb_init_input.argtypes = [vpp]; L.xmb_init_input.restype = C.c_int
    L.xmb_get_derived.argtypes = [vp]; L.xmb_get_derived.restype = C.POINTER(Derived)
    L.xmb_init_from_provider.argtypes = [C.POINTER(XrlProvider), vp, C.c_int, vpp]; L.xmb_init_from_provider.restype = C.c_int
    L.xmb_init_from_provider_gpu.argtypes = [C.POINTER(XrlProvider), vp, C.c_int, vpp]; L.xmb_init_from_provider_gpu.restype = C.c_int
    L.xmb_tables_gpu_last_ms.restype = C.c_double
    L.xmb_get_tables.argtypes = [vp]; L.xmb_get_tables.restype = C.POINTER(TablesHost)
    L.xmb_free_hdf5_F.argtypes = [vpp]; L.xmb_free_hdf5_F.restype = None
    L.xmb_tables_enable_advanced_compton.argtypes = [vp]; L.xmb_tables_enable_advanced_compton.restype = C.c_int
    L.xmb_solid_angle_inputs.argtypes = [vp, vp, C.POINTER(C.POINTER(SolidAngle))]; L.xmb_solid_angle_inputs.restype = C.c_int
    L.xmb_solid_angle_calculation.argtypes = [vp, vp, C.POINTER(C.POINTER(SolidAngle)), C.c_void_p,
                                              C.POINTER(MainOptions), C.c_long, C.c_uint64]
    L.xmb_solid_angle_calculation.restype = C.c_int
    L.xmb_solid_angle_grid.argtypes = [vp, c_double_p, C.c_long, c_double_p, C.c_long, C.c_long, C.c_uint64, C.c_int,
                                       c_double_p, C.POINTER(C.c_int32)]
    L.xmb_solid_angle_grid.restype = C.c_int
    L.xmb_solid_angle_last_hits.argtypes = [C.POINTER(C.c_int32), C.c_long]; L.xmb_solid_angle_last_hits.restype = C.c_long
    L.xmb_solid_angle_last_ms.restype = C.c_double
    L.xmb_set_hits_per_single.argtypes = [C.c_long]; L.xmb_set_hits_per_single.restype = None
    L.xmb_get_hits_per_single.restype = C.c_long
    L.xmb_free_solid_angle.argtypes = [C.POINTER(SolidAngle)]; L.xmb_free_solid_angle.restype = None
    L.xmb_main_options_defaults.argtypes = [C.POINTER(MainOptions)]; L.xmb_main_options_defaults.restype = None
    L.xmb_main_msim.argtypes = [vp, vp, C.c_int, C.POINTER(c_double_p), C.POINTER(MainOptions), C.POINTER(c_double_p),
                                C.POINTER(c_double_p), C.POINTER(SolidAngle)]
    L.xmb_main_msim.restype = C.c_int
    L.xmb_main_msim_raw.argtypes = [vp, vp, C.POINTER(MainOptions), C.POINTER(SolidAngle), C.POINTER(MsimEx),
                                    C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_size_t)]
    L.xmb_main_msim_raw.restype = C.c_int
    L.xmb_main_msim_finish.argtypes = [vp, vp, C.POINTER(MainOptions), C.POINTER(C.c_uint64), C.c_size_t,
                                       C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p)]
    L.xmb_main_msim_finish.restype = C.c_int
    L.xmb_msim_device_limbs.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.xmb_msim_device_limbs.restype = C.c_int
    L.xmb_msim_workload_stats.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.c_int]
    L.xmb_msim_workload_stats.restype = C.c_int
    L.xmb_msim_brute_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
    L.xmb_msim_brute_counters.restype = C.c_int
    pp = C.POINTER(c_double_p)
    L.xmb_detector_convolute_all.argtypes = [vp, vp, pp, pp, c_double_p, c_double_p, C.POINTER(MainOptions),
                                             C.POINTER(EscapeRatios), C.c_int, C.c_int]
    L.xmb_detector_convolute_all.restype = None
    L.xmb_detector_convolute_spectrum.argtypes = [vp, vp, c_double_p, pp, C.POINTER(MainOptions),
                                                  C.POINTER(EscapeRatios), C.c_int]
    L.xmb_detector_convolute_spectrum.restype = None
    L.xmb_detector_convolute_history.argtypes = [vp, vp, c_double_p, C.POINTER(MainOptions)]
    L.xmb_detector_convolute_history.restype = None
    L.xmb_detector_last_ms.restype = C.c_double
    L.xmb_input_read_from_xml_file.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Input))]; L.xmb_input_read_from_xml_file.restype = C.c_int
    L.xmb_input_validate.argtypes = [C.POINTER(Input)]; L.xmb_input_validate.restype = C.c_int
    L.xmb_input_free.argtypes = [C.POINTER(C.POINTER(Input))]; L.xmb_input_free.restype = None
    L.xmb_input_write_to_xml_file.argtypes = [C.POINTER(Input), C.c_char_p]; L.xmb_input_write_to_xml_file.restype = C.c_int
    L.xmb_output_write_to_xml_file.argtypes = [C.POINTER(Input), C.c_char_p, C.c_char_p, c_double_p, pp, c_double_p, c_double_p,
                                               C.c_int, C.POINTER(XrlProvider)]
    L.xmb_output_write_to_xml_file.restype = C.c_int
    L.xmb_write_spe_file.argtypes = [C.c_char_p, C.POINTER(Input), c_double_p]; L.xmb_write_spe_file.restype = C.c_int
    L.xmb_write_csv_file.argtypes = [C.c_char_p, C.POINTER(Input), pp, C.c_int]; L.xmb_write_csv_file.restype = C.c_int
    L.xmb_input_read_from_xml_string.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Input))]; L.xmb_input_read_from_xml_string.restype = C.c_int
    L.xmb_input_write_to_xml_string.argtypes = [C.POINTER(Input), C.POINTER(C.c_void_p)]; L.xmb_input_write_to_xml_string.restype = C.c_int
    L.xmb_check_solid_angle_match.argtypes = [C.POINTER(Input), C.POINTER(Input), C.POINTER(XrlProvider)]; L.xmb_check_solid_angle_match.restype = C.c_int
    L.xmb_check_escape_ratios_match.argtypes = [C.POINTER(Input), C.POINTER(Input)]; L.xmb_check_escape_ratios_match.restype = C.c_int
    L.xmb_cache_set_provider.argtypes = [C.POINTER(XrlProvider)]; L.xmb_cache_set_provider.restype = None
    L.xmb_find_solid_angle_match.argtypes = [C.c_char_p, C.POINTER(Input), C.POINTER(XrlProvider), C.POINTER(C.POINTER(SolidAngle)), C.POINTER(MainOptions)]
    L.xmb_find_solid_angle_match.restype = C.c_int
    L.xmb_update_solid_angle_cache_file.argtypes = [C.c_char_p, C.POINTER(SolidAngle)]; L.xmb_update_solid_angle_cache_file.restype = C.c_int
    L.xmb_find_escape_ratios_match.argtypes = [C.c_char_p, C.POINTER(Input), C.POINTER(C.POINTER(EscapeRatios)), C.POINTER(MainOptions)]
    L.xmb_find_escape_ratios_match.restype = C.c_int
    L.xmb_update_escape_ratios_cache_file.argtypes = [C.c_char_p, C.POINTER(EscapeRatios)]; L.xmb_update_escape_ratios_cache_file.restype = C.c_int
    L.xmb_tube_ebel.argtypes = [C.POINTER(XrlProvider), C.POINTER(Layer), C.POINTER(Layer), C.POINTER(Layer), C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_size_t,
                                c_double_p, c_double_p, C.POINTER(C.POINTER(Excitation))]
    L.xmb_tube_ebel.restype = C.c_int
    L.xmb_free_excitation.argtypes = [C.POINTER(C.POINTER(Excitation))]; L.xmb_free_excitation.restype = None
    L.xmb_get_default_escape_ratios_options.restype = EscapeRatiosOptions
    L.xmb_escape_ratios_input.argtypes = [C.POINTER(Input), C.POINTER(EscapeRatiosOptions), vpp]
    L.xmb_escape_ratios_input.restype = C.c_int
    L.xmb_escape_ratios_run.argtypes = [vp, vp, C.POINTER(EscapeRatiosOptions), C.c_uint64,
                                        C.POINTER(C.POINTER(EscapeRatios)), C.c_void_p]
    L.xmb_escape_ratios_run.restype = C.c_int
    L.xmb_escape_ratios_calculation.argtypes = [C.POINTER(Input), C.POINTER(C.POINTER(EscapeRatios)), C.c_void_p,
                                                C.POINTER(XrlProvider), C.POINTER(MainOptions), EscapeRatiosOptions,
                                                C.c_uint64]
    L.xmb_escape_ratios_calculation.restype = C.c_int
    L.xmb_free_escape_ratios.argtypes = [C.POINTER(C.POINTER(EscapeRatios))]; L.xmb_free_escape_ratios.restype = None
    L.xmb_escape_ratios_last_ms.restype = C.c_double
    L.xmb_detector_last_launches.restype = C.c_uint64
    L.xmi_solid_angle_calculation_cl.argtypes = [vp, C.POINTER(C.POINTER(SolidAngle)), C.c_void_p, C.POINTER(MainOptions)]
    L.xmi_solid_angle_calculation_cl.restype = C.c_int
    L.xmi_detector_convolute_all_custom.argtypes = [vp, pp, pp, c_double_p, c_double_p, C.POINTER(MainOptions),
                                                    C.POINTER(EscapeRatios), C.c_int, C.c_int]
    L.xmi_detector_convolute_all_custom.restype = None
    L.xmb_msim_shard_owner.argtypes = [C.c_uint64, C.c_int]; L.xmb_msim_shard_owner.restype = C.c_int
    L.xmb_msim_shard_count.argtypes = [C.c_uint64, C.c_int, C.c_int]; L.xmb_msim_shard_count.restype = C.c_uint64
    L.xmb_msim_total_histories.argtypes = [vp]; L.xmb_msim_total_histories.restype = C.c_uint64
    L.xmb_msim_slot_map.argtypes = [vp, vp, C.POINTER(MainOptions), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
    L.xmb_msim_slot_map.restype = C.c_int
    L.xmb_comm_unique_id.argtypes = [C.c_char_p]; L.xmb_comm_unique_id.restype = C.c_int
    L.xmb_comm_init_rank.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, vpp]; L.xmb_comm_init_rank.restype = C.c_int
    L.xmb_comm_free.argtypes = [vpp]; L.xmb_comm_free.restype = None
    L.xmb_comm_rank.argtypes = [vp]; L.xmb_comm_rank.restype = C.c_int
    L.xmb_comm_size.argtypes = [vp]; L.xmb_comm_size.restype = C.c_int
    L.xmb_nccl_version.restype = C.c_int
    L.xmb_main_msim_multi.argtypes = [vp, vp, vp, C.POINTER(c_double_p), C.POINTER(MainOptions), C.POINTER(c_double_p),
                                      C.POINTER(c_double_p), C.POINTER(SolidAngle)]
    L.xmb_main_msim_multi.restype = C.c_int
    L.xmb_main_msim_multi_raw.argtypes = [vp, vp, C.POINTER(MainOptions), C.POINTER(SolidAngle), vp, C.POINTER(MsimEx)]
    L.xmb_main_msim_multi_raw.restype = C.c_int
    L.xmb_main_msim_all_devices.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(c_double_p), C.POINTER(MainOptions),
                                            C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(SolidAngle), C.POINTER(MsimEx)]
    L.xmb_main_msim_all_devices.restype = C.c_int
    L.xmb_plugin_set_provider.argtypes = [C.POINTER(XrlProvider)]; L.xmb_plugin_set_provider.restype = None
    L.xmb_plugin_provider.restype = C.POINTER(XrlProvider)
    L.xmb_plugin_resolve_input.argtypes = [vp, vpp, C.POINTER(C.c_int)]; L.xmb_plugin_resolve_input.restype = C.c_int
    L.xmb_version.restype = C.c_char_p
    L.xmb_last_error.restype = C.c_char_p
    L.xmb_cuda_device_count.restype = C.c_int
    _lib = L
    return L


def last_error():
    return lib().xmb_last_error().decode()
