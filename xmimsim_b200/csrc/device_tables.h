// device_tables.h -- host-side handle of everything a simulation keeps in HBM (tables, accumulators, queues) and the
// entry points the three kernel files share.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "engine.h"
#include "history.cuh"

struct XmbDeviceTables {
	int cascade = 0, use_M_lines = -1, device = -1;
	std::vector<void *> allocs;
	XmbHistParams P{};
	// host metadata for the epilogue
	std::vector<int> rec_slot, rec_channel, rec_line, rec_zi, hist_base;
	int n_rec = 0, n_hist_slots = 0, max_nE = 1;
	size_t n_line_tiles_bytes = 0;     // all line-tile blobs (history.cuh)
	double W_max = 0.0;
	uint64_t n_total = 0;
	// solid-angle grid + accumulators (re-used across calls)
	double *sa_grid = nullptr, *sa_r = nullptr, *sa_t = nullptr;
	size_t sa_cap = 0, sa_r_cap = 0, sa_t_cap = 0;      // capacity of each of the three buffers, in doubles
	size_t sa_nr = 0, sa_nt = 0;                        // dimensions of the grid now in HBM
	uint64_t sa_hash = 0;                               // content hash of that grid (values + axes), see sa_content_hash()
	bool sa_valid = false;                              // the grid in HBM was uploaded whole with keep_on_device
	unsigned long long *acc = nullptr, *limbs = nullptr, *counters = nullptr;
	size_t acc_slots = 0;
	double *queue = nullptr;
	size_t queue_doubles = 0;
	int *line_slot = nullptr;          // [nZ][384] compact history slot of a line (brute-force scoring)
	double *auger_rate = nullptr;      // [nZ][XMB_N_AUGER]
	// the run in flight / last run (xmb_msim_launch / xmb_msim_collect)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	bool run_brute = false;
	size_t run_slots = 0;
	uint64_t run_launches = 0;
	unsigned long long brute_counters[8] = {0};
	unsigned long long layer_interactions[XMB_MAX_LAYERS] = {0};
	~XmbDeviceTables() {
		for (void *p : allocs) cudaFree(p);
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		cudaFree(sa_grid); cudaFree(sa_r); cudaFree(sa_t); cudaFree(acc); cudaFree(limbs); cudaFree(counters); cudaFree(queue);
	}
};

// cascade type 1 none, 2 non-radiative, 3 radiative, 4 full (src/xmi_main.F90:141-153)
int xmb_cascade_mode(const xmb_main_options *o);
// the handle's device tables for these options on the current device (built on first use, rebuilt when the cascade
// mode, the M-line switch or the device changed); NULL + xmb_last_error on failure
XmbDeviceTables *xmb_device_tables_get(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *opt);

// the two halves of xmb_main_msim_raw (history.cu): enqueue on the current (or ex->device's) default stream / wait and read counters
int xmb_msim_launch(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *options, const xmb_solid_angle *sa, xmb_msim_ex *ex,
                    XmbDeviceTables **D_out);
int xmb_msim_collect(XmbDeviceTables *D, const xmb_main_options *options, xmb_msim_ex *ex);

// brute-force kernel (brute.cu)
struct XmbBruteParams {
	int use_auger, use_rad;
	double collimator_height, collimator_radius, half_apex, vertex_x, vertex_y, vertex_z;
	int collimator_present;
	const int *line_slot;        // [nZ][384]: compact history slot of a line, -1 = not an active line
	const double *auger_rate;    // [nZ][XMB_N_AUGER] running sums within each block (K: 240, L1..L3: 135 each)
};
// launches xmb_brute_kernel on the current stream (one persistent CTA per SM); returns cudaGetLastError()
cudaError_t xmb_brute_launch(const XmbHistParams &P, const XmbBruteParams &B, bool advanced_compton, int sms, uint64_t n_histories);
