// history.cuh -- device-side data layout of the photon-history engine (see DESIGN.md "HBM layout").
#pragma once
#include <cstdint>
#include "solid_angle_device.cuh"

#define XMB_MAX_LAYERS 32
#define XMB_N_COUNTERS (8 + XMB_MAX_LAYERS + 16)   // [0..7] see below, [8 + L] interactions in layer L, [40 + i] phase clocks (experiment builds)
#define XMB_SHARD_BLOCK 1024
#define XMB_SHARD_SHIFT 10
#define XMB_ELEM_STRIDE 22         // per-element doubles in a node row
#define XMB_EO_CS_TOTAL 0
#define XMB_EO_P_RAYL 1
#define XMB_EO_P_RAYL_COMPT 2
#define XMB_EO_PHOTO_TOTAL 3
#define XMB_EO_PHOTO_PARTIAL 4     // 9 values
#define XMB_EO_VACANCY 13          // 9 values (selected cascade mode)
#define XMB_FIXED_SHIFT 56         // deposits are accumulated as round(w_rel * 2^56) in 128-bit integers

struct XmbSegDev {                 // one source segment: a valid continuous interval or a discrete line
	int is_cont, distribution_type;
	double energy, scale_parameter;          // discrete
	double weight_rel;                       // discrete: (I_h+I_v) exc_corr / n_photons_line / W_max
	double hor_ver_ratio;                    // discrete: I_h n_photons_line / (I_h+I_v)  (index rule)
	double x1, x2, y1, y2, h1, h2;           // continuous: interval ends, total and horizontal intensities
	double total_rel;                        // continuous: 0.5 (y1+y2)(x2-x1) / n_photons_interval / W_max
	double sigma_x, sigma_y, sigma_xp, sigma_yp;
};

struct XmbLayerDev {
	int n_elements, elem_begin;              // into elem_zi / elem_w
	double density, Z_begin, Z_end;
};

// Forced-detection line deposits ("line tiles").  The active line records of a layer -- (element, shell, line) with
// yield * rate > 0, E_line >= 1 keV -- are sorted by the edge energy of their shell and cut into tiles of 32 records:
// in the line phase a LANE owns a RECORD of the tile and the warp walks its 32 photons, so a record's deposits of the
// warp are summed in a register and leave as one staged add per tile (history.cu).  One blob per layer, 16-byte aligned,
// staged in shared memory by a bulk copy (cp.async.bulk + mbarrier):
//   int4 {n_tiles, n_groups, lanes_off, tile_bytes}
//   XmbLineTile  tiles[n_tiles]
//   XmbLineGroup groups[n_groups]            the (element, shell) groups in edge order; a tile spans <= XMB_TILE_GROUPS of them
//   per tile: float wy[32]                   weight fraction * fluorescence yield * radiative rate * 2^56 (0: padding lane)
//             float mu[nL][32]               -mu log2(e) of every layer at the line energy
//             int slot[32], gs[32]           history slot of the line; group of the record, relative to the tile's first group
#ifndef XMB_TILE_GROUPS
#define XMB_TILE_GROUPS 8
#endif
#define XMB_WPRE_STRIDE 33                   // doubles per group row of the per-warp scratch (33: lanes of different groups read different banks)
// doubles of per-warp scratch in the line phase: the photons' factors per shell group of a tile, float [tile_groups][XMB_WPRE_STRIDE]
__host__ __device__ inline int xmb_warp_scratch_doubles(int tile_groups) { return (4 * tile_groups * XMB_WPRE_STRIDE + 15) / 16 * 2; }
struct __align__(16) XmbLineTile {
	int g_begin, n_groups;                   // groups [g_begin, g_begin + n_groups) of the layer
	double min_edge;                         // lowest shell edge of the tile minus the edge-doublet half width: photons below it deposit exactly 0
};
struct __align__(16) XmbLineGroup {
	int row_off;                             // vacancy cross section of the shell in a node row
	int zi;
	double edgeK;                            // K shell: K edge (the reference skips the K lines below it), other shells: 0
};

struct XmbHistParams {
	// run
	uint64_t seed;
	// block-cyclic shard of the global photon ids [0, n_total): blocks of XMB_SHARD_BLOCK ids, block b belongs to
	// rank b % shard_n; n_local_span = owned blocks * XMB_SHARD_BLOCK (local index range, the last block may be partial)
	uint64_t n_total, n_local_span;
	int shard_rank, shard_n;
	uint64_t n_cont_seg, n_per_interval, n_per_line;
	int n_seg, n_int, nch, nL, nZ;
	int use_M_lines;
	int layer_sort;                          // batch formation: 0 as queued, 1 counting sort of a batch by layer, 2 one queue per (order, layer), 3 counting sort of a batch by energy class
	int n_ecls;                              // energy classes (mode 3): class = number of ecls_thr[] at or below the photon energy
	double ekey_scale;                       // energy buckets of the batch sort: key = (E - 1 keV) * ekey_scale
	double ecls_thr[30];                     // ascending shell-edge energies that split the line records into equal shares
	double zero, gain;
	const XmbSegDev *segs;
	// geometry
	double n_sample[3], p_window[3], n_detector[3];
	double ndo_new[9], ndo_inv[9];
	double detector_radius, slit_x1_max, slit_y1_max, d_source_slit;
	const XmbLayerDev *layers;
	const int *elem_zi;                      // unique-element index per (layer, element)
	const double *elem_w;                    // weight fraction
	const double *elem_wa;                   // weight fraction * N_A / A of the element (forced-detection scatter deposits)
	// node grid rows
	int n_nodes, n_buckets, row_stride, off_exc, off_elem;
	double bucket_E0, bucket_inv_dE;
	const double *node_E;
	const int *bucket_start;
	const double *rows;                      // [n_nodes][row_stride]
	const double *mu_tab;                    // [n_nodes][nL]: mu of the layers, compact (also the first nL entries of a row)
	// inverse CDFs
	int n_icdf_E, n_icdf_R, n_phi_T, n_cp, n_q;
	double cp_dR, cp_inv_dR, q_max;
	const double *icdf_E, *icdf_R, *phi_T, *cp_R;   // axis arrays
	const double *rayl_icdf, *compt_icdf;    // [nZ][n_icdf_E][n_icdf_R]
	const double *phi_icdf;                  // [n_phi_T][n_icdf_R]
	const double *cp_icdf;                   // [nZ][n_cp]
	// shell-resolved Compton profiles (use_advanced_compton); rows adv_off[zi] .. adv_off[zi+1]-1
	const int *adv_off;
	const double *adv_config, *adv_edge, *adv_cdf, *adv_qinv;
	const double *ff, *sf;                   // [nZ][n_q]
	// per unique element
	const double *atomic_weight;             // [nZ]
	const double *avog_over_A;               // [nZ] N_A / A (xraylib AVOGNUM units)
	const double *edge_K;                    // [nZ]
	const double *fluor_yield_corr;          // [nZ][9]
	const double *cos_kron;                  // [nZ][13]
	const double *rad_rate;                  // [nZ][384]
	const double *line_energy;               // [nZ][384]
	// forced-detection line tiles (see XmbLineTile): one blob per layer
	const char *lblob;                       // all blobs, layer L at lblob + lblob_off[L]
	int lblob_off[XMB_MAX_LAYERS + 1];
	int stage_nch;                           // channel slots of the shared-memory staging area: nch, or 0 when nch + history slots do not fit -- channel deposits then go straight to the global accumulators
	int lblob_stage_bytes;                   // shared memory reserved for one staged blob (largest layer); 0: tiles are read in place
	int tile_groups;                         // shell groups a tile may span (XMB_TILE_GROUPS, or half of it when shared memory is short)
	int lblob_main_layer;                    // the layer with the most records (staged once when batches mix layers)
	double rec_wy_max;                       // largest wy of the records (range check of the line deposits)
	const int *hist_base;                    // [nZ] first history slot of the element (+0 Rayleigh, +1 Compton)
	int n_hist_slots;
	// solid-angle grid
	const double *sa_grid;                   // [n_theta][n_r]
	int sa_nr, sa_nt;
	const double *sa_r_vals, *sa_t_vals;
	SaDetector sa_det;                       // for points beyond the grid: Monte Carlo on the spot
	int sa_hits_per_single;
	// global accumulators: [n_int][nch + n_hist_slots] 128-bit integers as (lo, hi) uint64 pairs
	double *queue;                           // per-CTA compaction queues [CTA][order][field][2T]
	unsigned long long *acc;
	unsigned long long *counters;            // [0] off-grid solid angle lookups, [1] interactions, [2] fixed-point range errors
};
