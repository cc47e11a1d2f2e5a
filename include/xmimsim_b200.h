/* xmimsim_b200.h -- C ABI of the B200-native XMI-MSIM photon-transport engine.
 *
 * Plain C, plain pointers and sizes: this is the drop-in boundary.  Every entry
 * point cites the reference interface (file:line under the reference tree) it replaces.  The
 * struct layouts of the input tree, the options, the solid-angle grid and the escape ratios are
 * identical to the reference's C structs so that a reference-side caller can hand its own
 * objects over without conversion.
 *
 * Nothing in here is torch; device memory, streams and (optional) NCCL are private to the .so.
 */
#ifndef XMIMSIM_B200_H
#define XMIMSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Input tree -- same field order/types as include/xmi_data_structs.h:40-329 (reference).
 * ------------------------------------------------------------------------------------------ */
typedef struct xmb_general {          /* reference: struct _xmi_general, xmi_data_structs.h:40-47 */
	float version;
	char *outputfile;
	long n_photons_interval;
	long n_photons_line;
	int n_interactions_trajectory;
	char *comments;
} xmb_general;

typedef struct xmb_layer {            /* struct _xmi_layer, xmi_data_structs.h:66-72 */
	int n_elements;
	int *Z;
	double *weight;
	double density;
	double thickness;
} xmb_layer;

typedef struct xmb_composition {      /* struct _xmi_composition, :91-95 */
	int n_layers;
	xmb_layer *layers;
	int reference_layer;              /* 1-based */
} xmb_composition;

typedef struct xmb_geometry {         /* struct _xmi_geometry, :119-130 */
	double d_sample_source;
	double n_sample_orientation[3];
	double p_detector_window[3];
	double n_detector_orientation[3];
	double area_detector;
	double collimator_height;
	double collimator_diameter;
	double d_source_slit;
	double slit_size_x;
	double slit_size_y;
} xmb_geometry;

typedef enum {                        /* XmiEnergyDiscreteDistribution, :145-149 */
	XMB_DISCRETE_MONOCHROMATIC = 0,
	XMB_DISCRETE_GAUSSIAN = 1,
	XMB_DISCRETE_LORENTZIAN = 2
} xmb_discrete_distribution;

typedef struct xmb_energy_discrete {  /* struct _xmi_energy_discrete, :167-177 */
	double energy;
	double horizontal_intensity;
	double vertical_intensity;
	double sigma_x;
	double sigma_xp;
	double sigma_y;
	double sigma_yp;
	int distribution_type;            /* xmb_discrete_distribution */
	double scale_parameter;
} xmb_energy_discrete;

typedef struct xmb_energy_continuous {/* struct _xmi_energy_continuous, :197-205 */
	double energy;
	double horizontal_intensity;
	double vertical_intensity;
	double sigma_x;
	double sigma_xp;
	double sigma_y;
	double sigma_yp;
} xmb_energy_continuous;

typedef struct xmb_excitation {       /* struct _xmi_excitation, :223-228 */
	int n_discrete;
	xmb_energy_discrete *discrete;
	int n_continuous;
	xmb_energy_continuous *continuous;
} xmb_excitation;

typedef struct xmb_absorbers {        /* struct _xmi_absorbers, :249-254 */
	int n_exc_layers;
	xmb_layer *exc_layers;
	int n_det_layers;
	xmb_layer *det_layers;
} xmb_absorbers;

typedef enum {                        /* XmiDetectorConvolutionProfile, :273-277 */
	XMB_DETECTOR_SILI = 0,
	XMB_DETECTOR_GE = 1,
	XMB_DETECTOR_SI_SDD = 2
} xmb_detector_profile;

typedef struct xmb_detector {         /* struct _xmi_detector, :296-307 */
	int detector_type;                /* xmb_detector_profile */
	double live_time;
	double pulse_width;
	double gain;
	double zero;
	double fano;
	double noise;
	int nchannels;
	int n_crystal_layers;
	xmb_layer *crystal_layers;
} xmb_detector;

typedef struct xmb_input {            /* struct _xmi_input, :329-336 */
	xmb_general *general;
	xmb_composition *composition;
	xmb_geometry *geometry;
	xmb_excitation *excitation;
	xmb_absorbers *absorbers;
	xmb_detector *detector;
} xmb_input;

typedef struct xmb_main_options {     /* struct _xmi_main_options, xmi_data_structs.h:479-495 */
	int use_M_lines;                  /* default 1 */
	int use_cascade_auger;            /* default 1 */
	int use_cascade_radiative;        /* default 1 */
	int use_variance_reduction;       /* default 1 */
	int use_sum_peaks;                /* default 0 */
	int use_escape_peaks;             /* default 1 */
	int escape_ratios_mode;           /* default 0 */
	int verbose;                      /* default 0 */
	int use_poisson;                  /* default 0 */
	int use_gpu;                      /* default 1 */
	int omp_num_threads;              /* default: all */
	int extra_verbose;                /* default 0 */
	char *custom_detector_response;   /* default NULL */
	int use_advanced_compton;         /* default 0 */
	int use_default_seeds;            /* default 0 */
} xmb_main_options;

typedef struct xmb_solid_angle {      /* struct _xmi_solid_angle, include/xmi_solid_angle.h:28-35 */
	double *solid_angles;             /* [grid_dims_theta_n][grid_dims_r_n], r fastest */
	long int grid_dims_r_n;
	long int grid_dims_theta_n;
	double *grid_dims_r_vals;
	double *grid_dims_theta_vals;
	char *xmi_input_string;
} xmb_solid_angle;

typedef struct xmb_escape_ratios {    /* struct _xmi_escape_ratios, include/xmi_detector.h:28-40 */
	int n_elements;
	int n_fluo_input_energies;
	int n_compton_input_energies;
	int n_compton_output_energies;
	int *Z;
	double *fluo_escape_ratios;       /* Fortran (n_elements, 109, n_fluo_input_energies): element fastest */
	double *fluo_escape_input_energies;
	double *compton_escape_ratios;    /* Fortran (n_compton_input_energies, n_compton_output_energies) */
	double *compton_escape_input_energies;
	double *compton_escape_output_energies;
	char *xmi_input_string;
} xmb_escape_ratios;

/* ------------------------------------------------------------------------------------------
 * Cross-section provider: the xraylib calls the reference makes (SURVEY.md 8c lists the call
 * sites).  Fill it with libxrl's functions to run on real data; xmb_xrl_surrogate() returns the
 * analytic stand-in shipped with this repo (xraylib is not available offline).
 * Shell codes K=0..M5=8 (..Q3=30); lines are xraylib's negative macros (KL1=-1 ... P3P5=-383);
 * Coster-Kronig transitions use XMB_F* below.
 * ------------------------------------------------------------------------------------------ */
enum { XMB_FL12 = 0, XMB_FL13, XMB_FL23, XMB_FM12, XMB_FM13, XMB_FM14, XMB_FM15, XMB_FM23,
       XMB_FM24, XMB_FM25, XMB_FM34, XMB_FM35, XMB_FM45, XMB_N_CK };

typedef struct xmb_xrl_provider {
	const char *name;
	double (*AtomicWeight)(int Z);
	double (*EdgeEnergy)(int Z, int shell);
	double (*LineEnergy)(int Z, int line);
	double (*FluorYield)(int Z, int shell);
	double (*RadRate)(int Z, int line);
	double (*CosKronTransProb)(int Z, int trans);
	double (*JumpFactor)(int Z, int shell);
	double (*CS_Total_Kissel)(int Z, double E);
	double (*CS_Photo_Total)(int Z, double E);
	double (*CS_Photo_Partial)(int Z, int shell, double E);
	double (*CS_Rayl)(int Z, double E);
	double (*CS_Compt)(int Z, double E);
	double (*FF_Rayl)(int Z, double q);
	double (*SF_Compt)(int Z, double q);
	double (*ComptonProfile)(int Z, double pz);
	/* vacancy-production cross section of `shell` under cascade mode (1 none, 2 non-radiative,
	 * 3 radiative, 4 full: src/xmi_aux_f.F90:662-665), given the already-evaluated P[] of the
	 * deeper shells (P[0]=PK ...).  Restates xraylib's P{L1..M5}_{pure,auger,rad,full}_kissel. */
	double (*VacancyCS)(int Z, int shell, double E, int cascade, const double *P);
	/* xraylib AugerRate(Z, <shell>_<new1><new2>_AUGER): fraction of the non-radiative decays of a
	 * vacancy in `shell` (K, L1..L3) that leave vacancies in new1 and new2 (ordered pair, as
	 * xraylib's macros are).  Shell numbers are xraylib's (K 0, L1 1 ... M5 8, N1 9 ... Q3 30).
	 * Only used by the brute-force mode (src/xmi_main.F90:2413-2481).  May be NULL: no Auger
	 * cascade offspring are then simulated. */
	double (*AugerRate)(int Z, int shell, int shell_new1, int shell_new2);
	/* xraylib ElectronConfig_Biggs(Z, shell) and ComptonProfile_Partial(Z, shell, pz), shell 0..30
	 * (K .. Q3).  Only used with options->use_advanced_compton (src/xmi_main.F90:4785-4983,
	 * src/xmi_variance_reduction.F90:752-947; tables src/xmi_data_f.F90:1120-1235).  May be NULL. */
	double (*ElectronConfig_Biggs)(int Z, int shell);
	double (*ComptonProfile_Partial)(int Z, int shell, double pz);
} xmb_xrl_provider;

const xmb_xrl_provider *xmb_xrl_surrogate(void);
/* The same stand-in with the line density of real data: every transition of xraylib's enumeration whose subshells exist
 * carries a rate (forbidden satellites ~1e-3), ~320 instead of ~150 active forced-detection lines for srm1155. */
const xmb_xrl_provider *xmb_xrl_surrogate_dense(void);
/* Provider backed by a libxrl loaded at run time (dlopen of `path`, or of libxrl.so.11 / .7 / libxrl.so when NULL):
 * the functions the reference links from xraylib >= 3.99 (configure.ac:115-116).  AugerRate is left NULL.
 * Returns NULL (xmb_last_error says why) when the library or one of its symbols is missing. */
const xmb_xrl_provider *xmb_xrl_from_library(const char *path);

/* ------------------------------------------------------------------------------------------
 * Opaque handles replacing xmi_inputFPtr / xmi_hdf5FPtr (pointers to Fortran derived types in
 * the reference, include/xmi_data_structs.h:511, include/xmi_data.h).
 * ------------------------------------------------------------------------------------------ */
typedef void *xmb_inputFPtr;
typedef void *xmb_hdf5FPtr;

/* Geometry derived by xmb_init_input (fields the reference stores inside its Fortran xmi_input:
 * src/xmi_main.F90:1741-1918). */
typedef struct xmb_derived {
	double n_sample_orientation[3];       /* normalised, z >= 0 */
	double n_detector_orientation[3];     /* normalised */
	double detector_radius;
	int collimator_present;
	double collimator_radius;
	double collimator_height;
	double half_apex;
	double vertex[3];
	double ndo_new[9];                    /* row-major 3x3, columns = new x,y,z axes in lab coords */
	double ndo_inv[9];                    /* row-major inverse */
	double detector_solid_angle;
	double n_sample_orientation_det[3];
	int n_layers;
	const double *thickness_along_Z;      /* [n_layers] */
	const double *Z_coord_begin;
	const double *Z_coord_end;
} xmb_derived;

/* Reference-shaped physics tables for the unique elements of one input (the contents of
 * xmimsimdata.h5 + what the reference fetches from xraylib inside the loop, pre-tabulated on the
 * common energy-node grid).  All arrays row-major, last index fastest.  Read-only view owned by
 * the hdf5 handle; the CPU oracle consumes exactly this view. */
typedef struct xmb_tables_host {
	int nZ;
	const int *Z;                  /* [nZ] ascending */
	const int *uniqZ;              /* [95]: Z -> index or -1 */
	const double *atomic_weight;   /* [nZ] */
	/* energy nodes (uniform grid + edge doublets, sorted) with O(1) bucket index */
	int n_nodes;
	const double *node_E;          /* [n_nodes] */
	double bucket_E0, bucket_inv_dE;
	int n_buckets;
	const int *bucket_start;       /* [n_buckets+1]: first node with E >= bucket lower bound */
	const double *cs_total;        /* [nZ][n_nodes]  CS_Total_Kissel */
	const double *cs_photo_total;  /* [nZ][n_nodes] */
	const double *p_rayl;          /* [nZ][n_nodes]  CS_Rayl/CS_Total */
	const double *p_rayl_compt;    /* [nZ][n_nodes]  (CS_Rayl+CS_Compt)/CS_Total */
	const double *cs_photo_partial;/* [nZ][9][n_nodes] */
	const double *cs_vacancy;      /* [4][nZ][9][n_nodes] cascade mode 1..4 -> index mode-1
	                                  (xraylib's P{K..M5}_{pure,auger,rad,full}_kissel) */
	/* scattering-angle inverse CDFs (xmimsimdata.h5 shapes: src/xmi_data.c:150-152) */
	int n_icdf_E, n_icdf_R;
	const double *icdf_E;          /* [n_icdf_E] uniform */
	const double *icdf_R;          /* [n_icdf_R] uniform 0..1 */
	const double *rayl_theta_icdf; /* [nZ][n_icdf_E][n_icdf_R] */
	const double *compt_theta_icdf;/* [nZ][n_icdf_E][n_icdf_R] */
	int n_phi_T;                   /* phi table: parameter axis 0..0.5 */
	const double *phi_T;           /* [n_phi_T] */
	const double *phi_icdf;        /* [n_phi_T][n_icdf_R] */
	int n_cp;                      /* Compton profile ICDF points (10001) */
	const double *cp_R;            /* [n_cp] uniform 0..1 */
	const double *cp_icdf;         /* [nZ][n_cp] */
	/* form factor / scattering function on a uniform q grid (for DCSP_Rayl / DCSP_Compt) */
	int n_q;
	double q_max;
	const double *ff;              /* [nZ][n_q] F(Z,q) */
	const double *sf;              /* [nZ][n_q] S(Z,q) */
	/* atomic constants */
	const double *fluor_yield;     /* [nZ][9] */
	const double *fluor_yield_corr;/* [nZ][9] */
	const double *cos_kron;        /* [nZ][XMB_N_CK] */
	const double *rad_rate;        /* [nZ][384] by |line| */
	const double *line_energy;     /* [nZ][384] */
	const double *edge_energy;     /* [nZ][9] */
	/* Per-layer tables on the nodes.  Every fluorescence-line energy (|line| <= 219, E >= 0.1 keV) of
	 * every element present and every monochromatic source-line energy IS a node, so a lookup at such
	 * an energy returns the provider's value exactly: this is how the reference's precalc_mu_cs
	 * (src/xmi_main.F90:227-237), precalc_xrf_cs (src/xmi_data_f.F90:1307-1476) and initial_mus
	 * (:597) are represented. */
	int n_layers;
	const double *mu_layer;        /* [n_layers][n_nodes]  sum_i w_i CS_Total_Kissel(Z_i, E) */
	const double *exc_murhod;      /* [n_nodes] sum over excitation-path absorbers of mu*rho*t */
	/* Auger transition rates in the reference's enumeration (src/xmi_main.F90:2482-4418):
	 * index 0..239 = K_<X><Y>, X in L1..M5 (8), Y in L1..Q3 (30): x*30 + y;
	 * 240 + 135*(s-1) + x*27 + y = L<s>_<X><Y>, X in M1..M5 (5), Y in M1..Q3 (27). */
	const double *auger_rate;      /* [nZ][XMB_N_AUGER] */
	/* Shell-resolved Compton profiles (use_advanced_compton), built on demand by
	 * xmb_tables_enable_advanced_compton; n_adv_rows = 0 until then.  Element zi owns rows
	 * adv_off[zi] .. adv_off[zi+1]-1, one per occupied subshell (src/xmi_data_f.F90:1126-1145). */
	int n_adv_rows;
	const int *adv_off;            /* [nZ+1] */
	const int *adv_shell;          /* [n_adv_rows] xraylib shell number */
	const double *adv_config;      /* [n_adv_rows] ElectronConfig_Biggs */
	const double *adv_edge;        /* [n_adv_rows] EdgeEnergy of the subshell (0 when unknown) */
	const double *adv_cdf;         /* [n_adv_rows][n_cp] profile_partial_cdf on Q = 100 j/(n_cp-1), normalised to 0.5 at Q = 100 */
	const double *adv_qinv;        /* [n_adv_rows][n_cp] Qs_inv on cdf = 0.5 j/(n_cp-1) */
} xmb_tables_host;
#define XMB_N_AUGER 645

/* ------------------------------------------------------------------------------------------
 * Entry points
 * ------------------------------------------------------------------------------------------ */

/* Replaces xmi_input_C2F (src/xmi_aux_f.F90:782-1025): deep copy, layer weights renormalised to
 * sum 1 (:857).  Returns 1 on success, 0 on failure. */
int xmb_input_C2F(const xmb_input *input, xmb_inputFPtr *out);
/* Replaces xmi_input_F2C (src/xmi_aux_f.F90:766-776): borrowed pointer to the handle's C tree. */
const xmb_input *xmb_input_F2C(xmb_inputFPtr inputF);
/* Replaces xmi_free_input_F. */
void xmb_free_input_F(xmb_inputFPtr *inputF);
/* Replaces xmi_init_input (src/xmi_main.F90:1741-1918).  Returns 1 / 0 (non-conical collimator
 * is reported as 0 instead of the reference's exit(1), :1785-1789). */
int xmb_init_input(xmb_inputFPtr *inputF);
/* Read-only view of what xmb_init_input derived.  NULL before xmb_init_input. */
const xmb_derived *xmb_get_derived(xmb_inputFPtr inputF);

/* Replaces xmi_init_from_hdf5 + xmi_update_input_from_hdf5 (include/xmi_data.h:45-50;
 * src/xmi_data_f.F90:98-848) and the table generator xmi_db (src/xmi_data.c:149-487,
 * src/xmi_data_f.F90:880-1551): builds the table bundle for the elements of `inputF` from a
 * cross-section provider.  `quality` scales the integration resolution of the inverse CDFs
 * (1 = reference's 1e5 theta steps / 1e7 pz steps; 0 = fast setting for tests).
 * Returns 1 / 0. */
int xmb_init_from_provider(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality,
                           xmb_hdf5FPtr *out);
/* Same bundle, with the inverse CDFs of the scattering angles and of the Compton profile integrated on the GPU
 * (the loops of xmi_db_Z_specific, src/xmi_data_f.F90:1002-1060, 1162-1186): the provider is sampled once per element
 * on fine grids instead of at every integration step.  Entries differ from xmb_init_from_provider by at most one
 * integration step (summation order).  Needs a CUDA device.  Returns 1 / 0. */
int xmb_init_from_provider_gpu(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality,
                               xmb_hdf5FPtr *out);
double xmb_tables_gpu_last_ms(void);
const xmb_tables_host *xmb_get_tables(xmb_hdf5FPtr hdf5F);
/* Builds the shell-resolved Compton tables (idempotent; host, OpenMP).  xmb_main_msim calls it itself when
 * options->use_advanced_compton is set.  Returns 1 / 0 (provider lacks the two partial-profile calls). */
int xmb_tables_enable_advanced_compton(xmb_hdf5FPtr hdf5F);
void xmb_free_hdf5_F(xmb_hdf5FPtr *hdf5F);

/* Replaces xmi_solid_angle_inputs_f (src/xmi_solid_angle_f.F90:62-301): allocates the grid
 * struct (malloc; free with xmb_free_solid_angle) and fills the r/theta axes.  Needs the tables
 * for mu(E) of the layers.  Returns 1 / 0. */
int xmb_solid_angle_inputs(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, xmb_solid_angle **out);
/* Replaces xmi_solid_angle_calculation_cl (src/xmi_solid_angle_cl.c:118-439; typedef
 * XmiSolidAngleCalculation src/xmi_solid_angle.c:51): computes the whole grid on the GPU.
 * input_string is stored (not copied) in (*solid_angle)->xmi_input_string, as the reference does.
 * Returns 1 on success, 0 = "fall through to next backend".  hits_per_single is the reference's
 * global (src/xmi_solid_angle_f.F90:43); seed selects the Philox key. */
int xmb_solid_angle_calculation(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F,
                                xmb_solid_angle **solid_angle, char *input_string,
                                const xmb_main_options *options, long hits_per_single,
                                uint64_t seed);
/* The grid kernel on caller-given axes (any n_r x n_theta): solid_angles[theta][r] and, if non-NULL,
 * the raw hit counts.  xmb_solid_angle_calculation = xmb_solid_angle_inputs + this.  Returns 1 / 0. */
int xmb_solid_angle_grid(xmb_inputFPtr inputF, const double *r_vals, long n_r, const double *theta_vals,
                         long n_theta, long hits_per_single, uint64_t seed, int verbose,
                         double *solid_angles, int32_t *hits);
/* Device time (ms, CUDA events) of the last grid computed. */
double xmb_solid_angle_last_ms(void);
/* Raw hit counts of the last grid computed by xmb_solid_angle_calculation on this thread's
 * device (int32 [theta][r]); for parity tests.  Returns number of points copied. */
long xmb_solid_angle_last_hits(int32_t *hits, long capacity);
/* The reference's mutable global `hits_per_single` (src/xmi_solid_angle_f.F90:43-44, default 5000): rays per point of
 * the plugin-shaped grid call and of the on-the-spot solid angle of an interaction point beyond the grid
 * (xmi_get_solid_angle, :783-789).  get: the value set here, else the host process's exported `hits_per_single`
 * symbol when there is one, else 5000. */
void xmb_set_hits_per_single(long n);
long xmb_get_hits_per_single(void);
void xmb_free_solid_angle(xmb_solid_angle *sa);   /* xmi_free_solid_angle, src/xmi_solid_angle.c:792-800 */

/* Replaces xmi_main_msim (include/xmi_main.h:29; src/xmi_main.F90:66-954).
 * channels:        malloc'ed double[(n_int+1)][nchannels], rows cumulative over interaction order, x live_time
 * brute_history:   malloc'ed double[100][385][n_int]  (all zero with variance reduction on)
 * var_red_history: malloc'ed double[100][385][n_int]  (NULL with variance reduction off, src/xmi_main.F90:942)
 * options->use_variance_reduction = 0 selects the brute-force mode (analogue walk, detector/collimator hit
 * tests src/xmi_aux_f.F90:1622-1833, Auger/radiative cascade offspring src/xmi_main.F90:2413-4783);
 * solid_angles may then be NULL and channels row 0 holds the photons detected without interaction.
 * n_mpi_hosts > 1 (the reference's MPI build): this rank -- OMPI_COMM_WORLD_RANK / PMI_RANK / RANK -- simulates its
 * block-cyclic shard of the photon ids of every source line; the outputs are PARTIAL sums which the host adds over
 * ranks, as the reference's host does with MPI_Reduce (bin/xmimsim.c:396-413).  xmb_main_msim_multi does the sum
 * itself (NCCL) and is what a multi-GPU caller should use.
 * Returns 1 / 0. */
int xmb_main_msim(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, int n_mpi_hosts, double **channels,
                  const xmb_main_options *options, double **brute_history,
                  double **var_red_history, const xmb_solid_angle *solid_angles);

/* Extended driver used by the multi-GPU harness: photon-id sharding and raw fixed-point output
 * so that the cross-rank sum is order-independent and bit-exact at any GPU count. */
typedef struct xmb_msim_ex {
	int rank, n_ranks;             /* shard [rank*N/n_ranks, (rank+1)*N/n_ranks) of every line/interval */
	uint64_t seed;                 /* Philox key; 0 -> default 0x584D494D53494D */
	int device;                    /* CUDA device ordinal, -1 = current */
	int keep_on_device;            /* 1: leave accumulators resident (read with xmb_msim_accum_*) */
	/* outputs */
	uint64_t n_histories;          /* histories simulated by this rank */
	double kernel_ms;              /* device time of the history kernel(s), CUDA events */
	uint64_t n_launches;           /* kernels launched */
	uint64_t n_interactions;       /* total interactions simulated (for bytes/history) */
} xmb_msim_ex;

/* Runs the history kernels and leaves the exact fixed-point accumulators in a host buffer of
 * uint64 limbs: accum[2*i], accum[2*i+1] = low / high 48-bit-split words of slot i.
 * Slot layout: channels [n_int][nch] (row k = deposits made at interaction k+1, NOT cumulative),
 * then history [n_int][n_hist_slots].  *n_slots receives the slot count.  The buffer is malloc'ed. */
int xmb_main_msim_raw(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                      const xmb_solid_angle *solid_angles, xmb_msim_ex *ex,
                      uint64_t **accum, size_t *n_slots);
/* Device pointer to the limbs of the last xmb_main_msim_raw run on this handle (uint64[n_words],
 * valid until the next run): lets the multi-GPU harness all-reduce in HBM (NCCL, int64 sum). */
int xmb_msim_device_limbs(xmb_hdf5FPtr hdf5F, uint64_t **dev_ptr, size_t *n_words);
/* Workload description of the last run: out[0] = n_layers, then per layer (interactions simulated,
 * n_elements, active forced-detection line records of its elements).  Feeds the algorithmic-bytes
 * figure of SURVEY.md 8(d) / DESIGN.md. */
int xmb_msim_workload_stats(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, uint64_t *out, int capacity);
/* Counters of the last brute-force run (options->use_variance_reduction = 0): out[1] interactions,
 * out[3] photons that reached the detector, out[4] cascade offspring photons walked, out[5] detected
 * photons whose line has no history slot. */
int xmb_msim_brute_counters(xmb_hdf5FPtr hdf5F, uint64_t *out, int capacity);
/* Host-side helpers of the sharded driver (no GPU needed): the block-cyclic photon-id shard of a rank (ids are
 * dealt in blocks of 1024, block b to rank b % n_ranks, so every rank simulates the same share of every source
 * line, as the reference's MPI split does: src/xmi_main.F90:314,574), the total
 * number of histories of an input, and the accumulator slot map: a row (one per interaction order) is
 * nchannels channel slots followed by n_hist_slots history slots; history slot s holds (out_Z[s], out_line[s]),
 * line 384 / 385 = Rayleigh / Compton.  xmb_msim_slot_map returns n_hist_slots (pass NULL arrays to query). */
int xmb_msim_shard_owner(uint64_t photon_id, int n_ranks);
uint64_t xmb_msim_shard_count(uint64_t n_total, int rank, int n_ranks);
uint64_t xmb_msim_total_histories(xmb_inputFPtr inputF);
int xmb_msim_slot_map(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                      int32_t *out_Z, int32_t *out_line, int capacity);
/* ---- multi-GPU: histories sharded over the GPUs of one box, ONE NCCL all-reduce of the histograms ----------------
 * Replaces the reference's MPI split (n_photons / n_mpi_hosts of every line per host, src/xmi_main.F90:314,574) and
 * its three MPI_Reduce(MPI_SUM) to rank 0 (bin/xmimsim.c:396-413).  A rank = one GPU.  The sum runs on the exact
 * integer limbs in HBM, on the stream of the history kernel (kernel -> limb conversion -> ncclAllReduce(uint64, sum),
 * no host synchronisation in between); every rank then holds the full result, bit-identical at any rank count.
 * NCCL (libnccl.so.2) is bound at run time; the calls fail with xmb_last_error when it cannot be loaded.
 *
 * One process per GPU (mpirun / torchrun): rank 0 calls xmb_comm_unique_id, the launcher broadcasts the
 * XMB_COMM_ID_BYTES bytes (MPI_Bcast / torch.distributed), every rank calls xmb_comm_init_rank (device < 0: current). */
#define XMB_COMM_ID_BYTES 128
typedef struct xmb_comm xmb_comm;
int xmb_comm_unique_id(char *id /* [XMB_COMM_ID_BYTES] */);
int xmb_comm_init_rank(const char *id, int rank, int n_ranks, int device, xmb_comm **out);
void xmb_comm_free(xmb_comm **comm);
int xmb_comm_rank(const xmb_comm *comm);
int xmb_comm_size(const xmb_comm *comm);
int xmb_nccl_version(void);              /* e.g. 22703; 0 when NCCL cannot be loaded */
/* xmi_main_msim over a communicator: same outputs as xmb_main_msim, on EVERY rank the sum over all ranks. */
int xmb_main_msim_multi(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, xmb_comm *comm, double **channels,
                        const xmb_main_options *options, double **brute_history,
                        double **var_red_history, const xmb_solid_angle *solid_angles);
/* Same, leaving the reduced limbs in HBM (xmb_msim_device_limbs; xmb_main_msim_finish converts them).  ex->seed and
 * ex->keep_on_device (solid-angle grid residency) are inputs; rank / n_ranks / device are taken from comm. */
int xmb_main_msim_multi_raw(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                            const xmb_solid_angle *solid_angles, xmb_comm *comm, xmb_msim_ex *ex);
/* One process driving n_devices GPUs (ordinals devices[0..n), or 0..n_devices-1 when NULL; n_devices <= 0: all
 * visible): tables and grid replicated, every device launched before any is waited for, ncclCommInitAll + one grouped
 * all-reduce.  ex (optional): seed in; total histories / interactions / launches and the slowest kernel time out. */
int xmb_main_msim_all_devices(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, int n_devices, const int *devices,
                              double **channels, const xmb_main_options *options, double **brute_history,
                              double **var_red_history, const xmb_solid_angle *solid_angles, xmb_msim_ex *ex);
/* Converts (summed) raw accumulators to the reference's three output arrays. */
int xmb_main_msim_finish(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                         const uint64_t *accum, size_t n_slots, double **channels,
                         double **brute_history, double **var_red_history);

/* Replaces xmi_detector_convolute_all / the plugin symbol xmi_detector_convolute_all_custom
 * (include/xmi_main.h:35-37; src/xmi_detector_f.F90:219-289).  channels_noconv[i] point at the rows of the
 * raw array (bin/xmimsim.c:496-498) and are MODIFIED IN PLACE by the efficiency correction, escape peaks
 * and pile-up, exactly as the reference's pointer remap does (src/xmi_detector_f.F90:412-413);
 * channels_conv[i] (i from zero_interaction?0:1 to n_interactions_all) are malloc'ed double[nchannels];
 * the two histories are corrected in place (var_red_history may be NULL); escape_ratios may be NULL
 * when escape peaks are off.  hdf5F supplies the cross-section provider (NULL: surrogate). */
void xmb_detector_convolute_all(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, double **channels_noconv,
                                double **channels_conv, double *brute_history,
                                double *var_red_history, const xmb_main_options *options,
                                const xmb_escape_ratios *escape_ratios, int n_interactions_all,
                                int zero_interaction);
/* Replaces xmi_detector_convolute_spectrum (include/xmi_main.h:31; src/xmi_detector_f.F90:339-580);
 * channels_noconv is modified in place, *channels_conv is malloc'ed (NULL on failure). */
void xmb_detector_convolute_spectrum(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F,
                                     double *channels_noconv, double **channels_conv,
                                     const xmb_main_options *options,
                                     const xmb_escape_ratios *escape_ratios, int n_interactions);
/* Replaces xmi_detector_convolute_history (include/xmi_main.h:33; src/xmi_detector_f.F90:291-337). */
void xmb_detector_convolute_history(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, double *history,
                                    const xmb_main_options *options);
/* ------------------------------------------------------------------------------------------
 * On-disk formats (host_io.cpp; no libxml2: a small built-in XML reader).
 * ------------------------------------------------------------------------------------------ */
/* Replaces xmi_input_read_from_xml_file (include/xmi_xml.h; src/xmi_xml.c:966-1340) with the reference reader's
 * conventions: weight fractions in any positive scale, |w| < 1e-20 dropped, elements sorted by Z and normalised,
 * lines and continuous points sorted by energy, nchannels optional (2048).  Also accepts an .xmso (reads its
 * <xmimsim-input>).  *input is malloc'ed: xmb_input_free.  Returns 1 / 0. */
int xmb_input_read_from_xml_file(const char *xmsifile, xmb_input **input);
void xmb_input_free(xmb_input **input);
/* Replaces xmi_input_validate (include/xmi_data_structs.h:508-515; src/xmi_data_structs.c:899-1255): 0 for a usable
 * input, else the OR of the flags of the sections that are not (the reference's XmiInputFlags values).  The readers
 * above and below reject an input with a non-zero result, as xmi_input_read_from_xml_file does (src/xmi_xml.c:1337). */
enum { XMB_INPUT_GENERAL = 1, XMB_INPUT_COMPOSITION = 2, XMB_INPUT_GEOMETRY = 4, XMB_INPUT_EXCITATION = 8,
       XMB_INPUT_ABSORBERS = 16, XMB_INPUT_DETECTOR = 32 };
int xmb_input_validate(const xmb_input *input);
/* Replaces xmi_input_write_to_xml_file (src/xmi_xml.c:1405-1450). */
int xmb_input_write_to_xml_file(const xmb_input *input, const char *xmsifile);
/* Replaces xmi_output_new + xmi_output_write_to_xml_file (src/xmi_data_structs.c:1369-1519;
 * src/xmi_xml.c:1453-1700): channels_unconv is the raw [(n_int+1)][nchannels] array of xmb_main_msim,
 * channels_conv the rows returned by the detector response, the histories [100][385][n_int] (NULL = empty).
 * The optional <svg_graphs> block is not written.  xrl NULL = surrogate (line energies in the history). */
int xmb_output_write_to_xml_file(const xmb_input *input, const char *inputfile, const char *xmsofile,
                                 const double *channels_unconv, double *const *channels_conv,
                                 const double *brute_history, const double *var_red_history,
                                 int use_zero_interactions, const xmb_xrl_provider *xrl);
/* String forms (xmi_input_read_from_xml_string / xmi_input_write_to_xml_string, src/xmi_xml.c:1342-1403):
 * *xmlstring is malloc'ed.  The caches below store this string as the key of an entry. */
int xmb_input_read_from_xml_string(const char *xmsistring, xmb_input **input);
int xmb_input_write_to_xml_string(const xmb_input *input, char **xmlstring);
/* The solid-angle and escape-ratio caches (xmi_find_solid_angle_match / xmi_update_solid_angle_hdf5_file,
 * src/xmi_solid_angle.c:193-790; xmi_find_escape_ratios_match / xmi_update_escape_ratios_hdf5_file,
 * src/xmi_detector.c:174-435) with the reference's match rules (xmi_check_solid_angle_match :420-670,
 * xmi_check_escape_ratios_match src/xmi_detector.c:143-172).  libhdf5 is not available to this build, so the
 * files are a documented side-car container with the same logical schema (host_cache.cpp).  find: returns 1 with
 * *rv = NULL when nothing matches (or the file does not exist yet), 0 on error; results are malloc'ed
 * (xmb_free_solid_angle / xmb_free_escape_ratios) and own their xmi_input_string.  update: appends an entry
 * (creating the file); the struct must carry its xmi_input_string. */
int xmb_check_solid_angle_match(const xmb_input *cached, const xmb_input *fresh, const xmb_xrl_provider *xrl);
int xmb_check_escape_ratios_match(const xmb_input *cached, const xmb_input *fresh);
/* The cross-section provider whose name is stamped into new cache entries and required of matching ones (the axes of a
 * solid-angle grid and the escape ratios depend on the cross sections).  Call once before using the cache functions. */
void xmb_cache_set_provider(const xmb_xrl_provider *xrl);
int xmb_find_solid_angle_match(const char *cache_file, const xmb_input *input, const xmb_xrl_provider *xrl,
                               xmb_solid_angle **rv, const xmb_main_options *options);
int xmb_update_solid_angle_cache_file(const char *cache_file, const xmb_solid_angle *solid_angle);
int xmb_find_escape_ratios_match(const char *cache_file, const xmb_input *input, xmb_escape_ratios **rv,
                                 const xmb_main_options *options);
int xmb_update_escape_ratios_cache_file(const char *cache_file, const xmb_escape_ratios *escape_ratios);
/* The SPE / CSV spectrum files of bin/xmimsim.c:546-640 (--spe-file*, --csv-file*). */
int xmb_write_spe_file(const char *filename, const xmb_input *input, const double *spectrum);
int xmb_write_csv_file(const char *filename, const xmb_input *input, double *const *rows, int first_row);

/* ------------------------------------------------------------------------------------------
 * X-ray tube source generator.
 * ------------------------------------------------------------------------------------------ */
/* Replaces xmi_tube_ebel (include/xmi_ebel.h; src/xmi_ebel.F90:114-521): Ebel's bremsstrahlung continuum on
 * 1 keV .. voltage in steps of delta_energy plus the anode's K and L lines, attenuated by the window and filter
 * (first element of each layer, as the reference), scaled by solid angle and current, optionally by a
 * transmission-efficiency curve (natural cubic spline, src/xmi_spline.c).  xrl NULL = surrogate.  *spectrum and
 * its arrays are malloc'ed (xmb_free_excitation).  The reference's CS_Total is taken from CS_Total_Kissel.
 * Returns 1 / 0. */
int xmb_tube_ebel(const xmb_xrl_provider *xrl, const xmb_layer *tube_anode, const xmb_layer *tube_window,
                  const xmb_layer *tube_filter, double tube_voltage, double tube_current,
                  double tube_angle_electron, double tube_angle_xray, double tube_delta_energy,
                  double tube_solid_angle, int tube_transmission, size_t tube_nefficiencies,
                  const double *tube_energies, const double *tube_efficiencies, xmb_excitation **spectrum);
void xmb_free_excitation(xmb_excitation **spectrum);

/* ------------------------------------------------------------------------------------------
 * Escape-peak ratios of the detector crystal (Monte Carlo, one interaction per photon).
 * ------------------------------------------------------------------------------------------ */
typedef struct xmb_escape_ratios_options {   /* struct _xmi_escape_ratios_options, include/xmi_detector.h:42-51 */
	long n_input_energies;            /* default 1990 */
	long n_compton_output_energies;   /* default 1999 */
	long n_photons;                   /* default 500000 */
	double input_energy_min;          /* default 1.0 keV */
	double input_energy_delta;        /* default 0.1 keV */
	double compton_output_energy_min; /* default 0.1 keV */
	double compton_output_energy_delta;/* default 0.1 keV */
} xmb_escape_ratios_options;
/* xmi_get_default_escape_ratios_options (src/xmi_detector.c:643-655). */
xmb_escape_ratios_options xmb_get_default_escape_ratios_options(void);
/* The input the reference simulates for the ratios (src/xmi_detector.c:91-141 +
 * xmi_init_input_escape_ratios, src/xmi_main.F90:1687-1738): composition = the crystal layers,
 * pencil beam at normal incidence one cm upstream, one interaction per trajectory; the input
 * energies of `ero` become its discrete lines (so that they are nodes of the table bundle).
 * Host only.  Returns an initialised handle (xmb_init_input already applied) or 0. */
int xmb_escape_ratios_input(const xmb_input *input, const xmb_escape_ratios_options *ero, xmb_inputFPtr *out);
/* Replaces xmi_escape_ratios_calculation (include/xmi_detector.h:57; src/xmi_detector.c:91-141 and
 * xmi_escape_ratios_calculation_fortran, src/xmi_main.F90:5473-5801).  The HDF5 file argument of the
 * reference becomes the cross-section provider (NULL: surrogate).  *escape_ratios and its arrays are
 * malloc'ed (xmb_free_escape_ratios); the struct keeps its own copy of input_string (the reference's driver
 * hands over a g_strdup, src/xmi_detector.c:139, and xmi_free_escape_ratios frees it, :566).
 * `seed` 0 = library default.  Returns 1 / 0 (the reference is void and exits on failure). */
int xmb_escape_ratios_calculation(const xmb_input *input, xmb_escape_ratios **escape_ratios, char *input_string,
                                  const xmb_xrl_provider *xrl, const xmb_main_options *options,
                                  xmb_escape_ratios_options ero, uint64_t seed);
/* The Monte Carlo proper on an escape-mode handle pair (the two calls above + xmb_init_from_provider). */
int xmb_escape_ratios_run(xmb_inputFPtr esc_inputF, xmb_hdf5FPtr esc_hdf5F, const xmb_escape_ratios_options *ero,
                          uint64_t seed, xmb_escape_ratios **escape_ratios, char *input_string);
void xmb_free_escape_ratios(xmb_escape_ratios **escape_ratios);
double xmb_escape_ratios_last_ms(void);

/* Device time (ms) and kernel launches of the last detector-response call. */
double xmb_detector_last_ms(void);
uint64_t xmb_detector_last_launches(void);

/* ------------------------------------------------------------------------------------------
 * Plugin symbols under the reference's own names (plugin_shim.cpp).  They take the reference's
 * opaque xmi_inputFPtr OR one of this library's handles.
 * ------------------------------------------------------------------------------------------ */
/* XmiSolidAngleCalculation (src/xmi_solid_angle.c:51), symbol name from src/xmi_solid_angle_cl.c:118-120. */
int xmi_solid_angle_calculation_cl(void *inputFPtr, xmb_solid_angle **solid_angle, char *input_string,
                                   xmb_main_options *options);
/* XmiDetectorConvoluteAll (include/xmi_main.h:37), symbol looked up at bin/xmimsim.c:513. */
void xmi_detector_convolute_all_custom(void *inputFPtr, double **channels_noconv, double **channels_conv,
                                       double *brute_history, double *var_red_history,
                                       xmb_main_options *options, xmb_escape_ratios *escape_ratios,
                                       int n_interactions_all, int zero_interaction);
/* Cross-section provider used by the two plugin symbols (NULL: analytic surrogate). */
void xmb_plugin_set_provider(const xmb_xrl_provider *provider);
/* The provider the plugin symbols use: the registered one, else xraylib bound at run time, else -- only with
 * XMB_ALLOW_SURROGATE=1 -- the analytic stand-in; NULL + xmb_last_error otherwise. */
const xmb_xrl_provider *xmb_plugin_provider(void);
/* Turns the handle a reference host passes (its opaque Fortran xmi_inputFPtr, converted with the host's own
 * xmi_input_F2C, src/xmi_aux_f.F90:766-776) or one of this library's handles into an xmb_inputFPtr; *owned = 1 when
 * a temporary handle was created (free it with xmb_free_input_F).  Used by libxmimsim-b200-interpose.so. */
int xmb_plugin_resolve_input(void *inputFPtr, xmb_inputFPtr *out, int *owned);

/* xmi_main_options_new defaults (src/xmi_data_structs.c:2531-2565). */
void xmb_main_options_defaults(xmb_main_options *options);

/* Library / device info. */
const char *xmb_version(void);
int xmb_cuda_device_count(void);
const char *xmb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
