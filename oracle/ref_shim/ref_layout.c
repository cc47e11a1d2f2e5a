/* oracle/ref_shim/ref_layout.c -- TEST INFRASTRUCTURE: include/xmimsim_b200.h against the reference's own headers
 * (include/xmi_data_structs.h, xmi_solid_angle.h, xmi_detector.h, compiled from /root/reference through the GLib
 * stand-in): every struct that crosses the C ABI has the reference's size and every field the reference's offset and
 * size.  The checks are _Static_asserts: a mismatch fails oracle/build_ref.sh.  ref_layout_checks() returns how many
 * field checks this file holds (tests/test_reference_cpu.py asserts it is the number it expects). */
#include <stddef.h>
#include "xmi_data_structs.h"
#include "xmi_solid_angle.h"
#include "xmi_detector.h"
#include "xmimsim_b200.h"

#define FIELD_SIZE(t, f) sizeof(((t *)0)->f)
#define SAME_STRUCT(a, b) _Static_assert(sizeof(a) == sizeof(b), "sizeof " #a " != sizeof " #b)
#define SAME_FIELD(a, b, f)                                                                         \
	_Static_assert(offsetof(a, f) == offsetof(b, f), "offset of " #f " differs: " #a " / " #b);   \
	_Static_assert(FIELD_SIZE(a, f) == FIELD_SIZE(b, f), "size of " #f " differs: " #a " / " #b); \
	enum { check_##a##_##f = __COUNTER__ }

SAME_STRUCT(xmb_general, xmi_general);
SAME_FIELD(xmb_general, xmi_general, version); SAME_FIELD(xmb_general, xmi_general, outputfile);
SAME_FIELD(xmb_general, xmi_general, n_photons_interval); SAME_FIELD(xmb_general, xmi_general, n_photons_line);
SAME_FIELD(xmb_general, xmi_general, n_interactions_trajectory); SAME_FIELD(xmb_general, xmi_general, comments);
SAME_STRUCT(xmb_layer, xmi_layer);
SAME_FIELD(xmb_layer, xmi_layer, n_elements); SAME_FIELD(xmb_layer, xmi_layer, Z); SAME_FIELD(xmb_layer, xmi_layer, weight);
SAME_FIELD(xmb_layer, xmi_layer, density); SAME_FIELD(xmb_layer, xmi_layer, thickness);
SAME_STRUCT(xmb_composition, xmi_composition);
SAME_FIELD(xmb_composition, xmi_composition, n_layers); SAME_FIELD(xmb_composition, xmi_composition, layers);
SAME_FIELD(xmb_composition, xmi_composition, reference_layer);
SAME_STRUCT(xmb_geometry, xmi_geometry);
SAME_FIELD(xmb_geometry, xmi_geometry, d_sample_source); SAME_FIELD(xmb_geometry, xmi_geometry, n_sample_orientation);
SAME_FIELD(xmb_geometry, xmi_geometry, p_detector_window); SAME_FIELD(xmb_geometry, xmi_geometry, n_detector_orientation);
SAME_FIELD(xmb_geometry, xmi_geometry, area_detector); SAME_FIELD(xmb_geometry, xmi_geometry, collimator_height);
SAME_FIELD(xmb_geometry, xmi_geometry, collimator_diameter); SAME_FIELD(xmb_geometry, xmi_geometry, d_source_slit);
SAME_FIELD(xmb_geometry, xmi_geometry, slit_size_x); SAME_FIELD(xmb_geometry, xmi_geometry, slit_size_y);
SAME_STRUCT(xmb_energy_discrete, xmi_energy_discrete);
SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, energy); SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, horizontal_intensity);
SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, vertical_intensity); SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, sigma_x);
SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, sigma_xp); SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, sigma_y);
SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, sigma_yp); SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, distribution_type);
SAME_FIELD(xmb_energy_discrete, xmi_energy_discrete, scale_parameter);
SAME_STRUCT(xmb_energy_continuous, xmi_energy_continuous);
SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, energy); SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, horizontal_intensity);
SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, vertical_intensity); SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, sigma_x);
SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, sigma_xp); SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, sigma_y);
SAME_FIELD(xmb_energy_continuous, xmi_energy_continuous, sigma_yp);
SAME_STRUCT(xmb_excitation, xmi_excitation);
SAME_FIELD(xmb_excitation, xmi_excitation, n_discrete); SAME_FIELD(xmb_excitation, xmi_excitation, discrete);
SAME_FIELD(xmb_excitation, xmi_excitation, n_continuous); SAME_FIELD(xmb_excitation, xmi_excitation, continuous);
SAME_STRUCT(xmb_absorbers, xmi_absorbers);
SAME_FIELD(xmb_absorbers, xmi_absorbers, n_exc_layers); SAME_FIELD(xmb_absorbers, xmi_absorbers, exc_layers);
SAME_FIELD(xmb_absorbers, xmi_absorbers, n_det_layers); SAME_FIELD(xmb_absorbers, xmi_absorbers, det_layers);
SAME_STRUCT(xmb_detector, xmi_detector);
SAME_FIELD(xmb_detector, xmi_detector, detector_type); SAME_FIELD(xmb_detector, xmi_detector, live_time);
SAME_FIELD(xmb_detector, xmi_detector, pulse_width); SAME_FIELD(xmb_detector, xmi_detector, gain); SAME_FIELD(xmb_detector, xmi_detector, zero);
SAME_FIELD(xmb_detector, xmi_detector, fano); SAME_FIELD(xmb_detector, xmi_detector, noise); SAME_FIELD(xmb_detector, xmi_detector, nchannels);
SAME_FIELD(xmb_detector, xmi_detector, n_crystal_layers); SAME_FIELD(xmb_detector, xmi_detector, crystal_layers);
SAME_STRUCT(xmb_input, xmi_input);
SAME_FIELD(xmb_input, xmi_input, general); SAME_FIELD(xmb_input, xmi_input, composition); SAME_FIELD(xmb_input, xmi_input, geometry);
SAME_FIELD(xmb_input, xmi_input, excitation); SAME_FIELD(xmb_input, xmi_input, absorbers); SAME_FIELD(xmb_input, xmi_input, detector);
SAME_STRUCT(xmb_main_options, xmi_main_options);
SAME_FIELD(xmb_main_options, xmi_main_options, use_M_lines); SAME_FIELD(xmb_main_options, xmi_main_options, use_cascade_auger);
SAME_FIELD(xmb_main_options, xmi_main_options, use_cascade_radiative); SAME_FIELD(xmb_main_options, xmi_main_options, use_variance_reduction);
SAME_FIELD(xmb_main_options, xmi_main_options, use_sum_peaks); SAME_FIELD(xmb_main_options, xmi_main_options, use_escape_peaks);
SAME_FIELD(xmb_main_options, xmi_main_options, escape_ratios_mode); SAME_FIELD(xmb_main_options, xmi_main_options, verbose);
SAME_FIELD(xmb_main_options, xmi_main_options, use_poisson); SAME_FIELD(xmb_main_options, xmi_main_options, use_gpu);
SAME_FIELD(xmb_main_options, xmi_main_options, omp_num_threads); SAME_FIELD(xmb_main_options, xmi_main_options, extra_verbose);
SAME_FIELD(xmb_main_options, xmi_main_options, custom_detector_response); SAME_FIELD(xmb_main_options, xmi_main_options, use_advanced_compton);
SAME_FIELD(xmb_main_options, xmi_main_options, use_default_seeds);
SAME_STRUCT(xmb_solid_angle, xmi_solid_angle);
SAME_FIELD(xmb_solid_angle, xmi_solid_angle, solid_angles); SAME_FIELD(xmb_solid_angle, xmi_solid_angle, grid_dims_r_n);
SAME_FIELD(xmb_solid_angle, xmi_solid_angle, grid_dims_theta_n); SAME_FIELD(xmb_solid_angle, xmi_solid_angle, grid_dims_r_vals);
SAME_FIELD(xmb_solid_angle, xmi_solid_angle, grid_dims_theta_vals); SAME_FIELD(xmb_solid_angle, xmi_solid_angle, xmi_input_string);
SAME_STRUCT(xmb_escape_ratios, xmi_escape_ratios);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, n_elements); SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, n_fluo_input_energies);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, n_compton_input_energies); SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, n_compton_output_energies);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, Z); SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, fluo_escape_ratios);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, fluo_escape_input_energies); SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, compton_escape_ratios);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, compton_escape_input_energies);
SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, compton_escape_output_energies); SAME_FIELD(xmb_escape_ratios, xmi_escape_ratios, xmi_input_string);
SAME_STRUCT(xmb_escape_ratios_options, xmi_escape_ratios_options);
SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, n_input_energies); SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, n_compton_output_energies);
SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, n_photons); SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, input_energy_min);
SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, input_energy_delta); SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, compton_output_energy_min);
SAME_FIELD(xmb_escape_ratios_options, xmi_escape_ratios_options, compton_output_energy_delta);
/* the flag values of xmb_input_validate and the enumerations stored in int fields */
_Static_assert(XMB_INPUT_GENERAL == XMI_INPUT_GENERAL && XMB_INPUT_COMPOSITION == XMI_INPUT_COMPOSITION && XMB_INPUT_GEOMETRY == XMI_INPUT_GEOMETRY &&
               XMB_INPUT_EXCITATION == XMI_INPUT_EXCITATION && XMB_INPUT_ABSORBERS == XMI_INPUT_ABSORBERS && XMB_INPUT_DETECTOR == XMI_INPUT_DETECTOR, "XmiInputFlags");
_Static_assert((int)XMB_DISCRETE_MONOCHROMATIC == (int)XMI_ENERGY_DISCRETE_DISTRIBUTION_MONOCHROMATIC && (int)XMB_DISCRETE_GAUSSIAN == (int)XMI_ENERGY_DISCRETE_DISTRIBUTION_GAUSSIAN &&
               (int)XMB_DISCRETE_LORENTZIAN == (int)XMI_ENERGY_DISCRETE_DISTRIBUTION_LORENTZIAN, "XmiEnergyDiscreteDistribution");
_Static_assert((int)XMB_DETECTOR_SILI == (int)XMI_DETECTOR_CONVOLUTION_PROFILE_SILI && (int)XMB_DETECTOR_GE == (int)XMI_DETECTOR_CONVOLUTION_PROFILE_GE &&
               (int)XMB_DETECTOR_SI_SDD == (int)XMI_DETECTOR_CONVOLUTION_PROFILE_SI_SDD, "XmiDetectorConvolutionProfile");

int ref_layout_checks(void) { return __COUNTER__; }
/* sizeof of the reference's structs, in the order general, layer, composition, geometry, energy_discrete,
 * energy_continuous, excitation, absorbers, detector, input, main_options, solid_angle, escape_ratios */
int ref_struct_sizes(int *out) {
	const int s[13] = {sizeof(xmi_general), sizeof(xmi_layer), sizeof(xmi_composition), sizeof(xmi_geometry), sizeof(xmi_energy_discrete),
	                   sizeof(xmi_energy_continuous), sizeof(xmi_excitation), sizeof(xmi_absorbers), sizeof(xmi_detector), sizeof(xmi_input),
	                   sizeof(xmi_main_options), sizeof(xmi_solid_angle), sizeof(xmi_escape_ratios)};
	for (int i = 0; i < 13; i++) out[i] = s[i];
	return 13;
}
