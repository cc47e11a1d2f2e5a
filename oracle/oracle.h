/* oracle.h -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * import, link or execute anything under oracle/.  The product (xmimsim_b200/) never does.
 *
 * Each function follows the reference file:line cited at its definition.  fp64 throughout, as the
 * Fortran reference.  Random numbers: the reference uses one MT19937 per OpenMP thread; the oracle
 * uses the same per-history / per-grid-point Philox4x32-10 counter streams as the GPU engine so
 * that results are comparable history-by-history (the reference itself is only statistically
 * reproducible across thread counts).
 *
 * PARITY STATUS: the Fortran reference cannot be built here (no Fortran compiler, no xraylib, no
 * HDF5/GLib), so the oracle is pinned by: the Random123 Philox known-answer vectors,
 * the reference's quadratic-solver test cases (tests/test-poly-solve-quadratic.F90), analytic
 * solid angles, and the golden convoluted spectra in examples/.xmso for the detector response.
 * Spectra parity against examples/.xmso is "parity unpinned": it needs xraylib data.
 */
#ifndef XMB_ORACLE_H
#define XMB_ORACLE_H
#include <stdint.h>
#include "xmimsim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Random123; Salmon et al. SC'11).  out[4] = philox(ctr[4], key[2]). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* xmi_poly_solve_quadratic  (src/xmi_aux_f.F90:1872-1905).  Returns number of roots (0,1,2). */
int orc_poly_solve_quadratic(double a, double b, double c, double *rv1, double *rv2);

/* derived geometry, restating xmi_init_input (src/xmi_main.F90:1741-1918).  input is modified:
 * normals normalised, sample normal flipped to +z.  Layer arrays are malloc'ed into *d. */
typedef struct orc_derived {
	double detector_radius, collimator_radius, collimator_height, half_apex, vertex[3];
	int collimator_present;
	double ndo_new[9], ndo_inv[9];   /* row-major */
	double detector_solid_angle, n_sample_orientation_det[3];
	int n_layers;
	double *thickness_along_Z, *Z_coord_begin, *Z_coord_end;
} orc_derived;
int orc_init_input(xmb_input *input, orc_derived *d);
void orc_free_derived(orc_derived *d);

/* xmi_single_solid_angle_calculation (src/xmi_solid_angle_f.F90:432-710) for grid point index
 * `point_id` (= theta_index * n_r + r_index), Philox stream (seed, point_id).  Returns the solid
 * angle; *hits receives the number of rays that reached the detector. */
double orc_single_solid_angle(const orc_derived *d, double r1, double theta1, long hits_per_single,
                              uint64_t seed, uint64_t point_id, long *hits);
/* ... for an interaction point of photon g beyond the grid (src/xmi_solid_angle_f.F90:783-789) */
double orc_single_solid_angle_photon(const orc_derived *d, double r1, double theta1, long hits_per_single,
                                     uint64_t seed, uint64_t g, int order, long *hits);
/* the reference's global hits_per_single (src/xmi_solid_angle_f.F90:43-44) as the history loop sees it */
void orc_set_hits_per_single(long n);
/* xmi_solid_angle_calculation_f grid loop (src/xmi_solid_angle_f.F90:303-429) over a caller-given
 * sub-grid: solid_angles[it*n_r + ir], hits likewise.  point ids use the FULL grid width
 * full_n_r so that a sub-grid reproduces the same streams: id = theta_idx[it]*full_n_r + r_idx[ir]. */
void orc_solid_angle_grid(const orc_derived *d, const double *r_vals, const int *r_idx, int n_r,
                          const double *theta_vals, const int *theta_idx, int n_theta, long full_n_r,
                          long hits_per_single, uint64_t seed, double *solid_angles, int32_t *hits,
                          int n_threads);
/* xmi_solid_angle_inputs_f (src/xmi_solid_angle_f.F90:62-301): grid axes from penetration depths;
 * mu of the layers supplied by the provider.  r_vals/theta_vals: caller arrays of 1024. */
int orc_solid_angle_axes(const xmb_input *input, const orc_derived *d, const xmb_xrl_provider *xrl,
                         double *r_vals, double *theta_vals, long n);

/* Photon histories with forced detection: xmi_main_msim (src/xmi_main.F90:66-954) restricted to the
 * global photon ids [g_begin, g_end) (ids enumerate valid continuous intervals, then discrete lines,
 * photon-minor).  RAW sums, no live_time:
 *   channels[(n_int+1)][nch]  cumulative over interaction order (reference's channels(0:n_int,:))
 *   var_red[n_int][385][100]  = Fortran var_red_history(Z, slot, k)
 * counters[0] = solid-angle lookups that fell off the grid, counters[1] = interactions simulated. */
uint64_t orc_total_histories(const xmb_input *in);
uint64_t orc_main_msim_range(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T,
                             const xmb_main_options *opt, const xmb_solid_angle *sa, uint64_t seed,
                             uint64_t g_begin, uint64_t g_end, int n_threads, double *channels,
                             double *var_red, uint64_t *counters);

/* Same for the block-cyclic shard of `rank` (blocks of 1024 ids, block b to rank b % n_ranks). */
uint64_t orc_main_msim_shard(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T,
                             const xmb_main_options *opt, const xmb_solid_angle *sa, uint64_t seed, int rank,
                             int n_ranks, int n_threads, double *channels, double *var_red, uint64_t *counters);

/* Brute-force mode (use_variance_reduction = 0; src/xmi_main.F90:1229-1416, :1920-1984, :2231-4783;
 * src/xmi_aux_f.F90:1622-1833): analogue walk, detector/collimator hit tests, Auger and radiative cascade offspring.
 * channels[(n_int+1)][nch] cumulative from the photon's interaction count; brute[n_int][385][100] =
 * Fortran brute_history(Z, slot, k).  RAW sums.  counters[0] detector hits, [1] interactions, [2] offspring. */
uint64_t orc_main_msim_brute_range(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T,
                                   const xmb_main_options *opt, uint64_t seed, uint64_t g_begin, uint64_t g_end,
                                   int n_threads, double *channels, double *brute, uint64_t *counters);

/* Escape-peak ratios of the crystal (src/xmi_main.F90:5473-5801) on the escape-mode input (its discrete lines
 * are the input energies).  fluo[(i*109 + line-1)*nZ + zi] and compt[c*nE + i] are the Fortran layouts
 * fluo_escape_ratios(element, line, i) / compton_escape_ratios(i, c), already divided by photons_interacted. */
int orc_escape_ratios(const xmb_input *in, const orc_derived *d, const xmb_tables_host *T, uint64_t seed,
                      long n_photons, int n_out, double out_min, double out_delta, int n_threads,
                      double *fluo, double *compt);

/* X-ray tube spectrum (xmi_tube_ebel, src/xmi_ebel.F90:114-521) and its spline (src/xmi_spline.c). */
double orc_cubic_spline(const double *x, const double *y, size_t n, double v);
int orc_tube_ebel(const xmb_xrl_provider *xrl, const xmb_layer *anode, const xmb_layer *window, const xmb_layer *filter,
                  double V, double current, double angle_e, double angle_x, double dE, double solid_angle,
                  int transmission, size_t n_eff, const double *eff_E, const double *eff, double *cont_E,
                  double *cont_I, int *ndisc_out, double *disc_E, double *disc_I);

/* Detector response (src/xmi_detector_f.F90).  noconv[nch] is modified IN PLACE by the absorption
 * correction, escape peaks and pile-up, as the reference does (:412-413); conv[nch] is the result. */
double orc_detector_correction(const xmb_input *in, const xmb_xrl_provider *xrl, double E);
void orc_detector_gaussian(const xmb_input *in, const double *temp, double *conv);
void orc_detector_convolute_spectrum(const xmb_input *in, const xmb_xrl_provider *xrl, double *noconv, double *conv,
                                     const xmb_main_options *opt, const xmb_escape_ratios *er, int n_interactions,
                                     uint64_t seed);
void orc_detector_convolute_history(const xmb_input *in, const xmb_xrl_provider *xrl, double *history);

#ifdef __cplusplus
}
#endif
#endif
