// escape.cu -- escape-peak ratios of the detector crystal on the GPU (xmi_escape_ratios_calculation).
#include <cstdio>
#include <vector>
#include <algorithm>
#include "history_device.cuh"
#include "device_tables.h"

// =====================================================================================================
// Escape-peak ratios of the detector crystal (src/xmi_main.F90:5473-5801): per input energy, n_photons
// pencil-beam photons are forced to interact once in the crystal (weight = interaction probability);
// a photon whose secondary (Compton-scattered or K/L fluorescence) leaves the crystal without a second
// interaction is tallied.  Streams: photon id g = energy index * n_photons + j; order 0 = source
// (slit x, slit y, polarisation angle), order 1 = the interaction (same addresses as the history kernel),
// order 2 stage 1 block 0 word 0 = free path of the secondary.
// Tallies are exact: weights <= 1 in 2^-40 fixed point, 64-bit integer sums.
// =====================================================================================================
#define XMB_ESC_SHIFT 40
struct XmbEscParams {
	uint64_t n_photons;
	int n_out;
	double out_min, out_delta;
	unsigned long long *fluo;        // [nE][109][nZ]
	unsigned long long *compt;       // [n_out][nE]
	unsigned long long *interacted;  // [nE]
};

__device__ __forceinline__ unsigned long long esc_fixed(double w) { return (unsigned long long)(w * (double)(1ULL << XMB_ESC_SHIFT) + 0.5); }

template <int NL>
#ifndef XMB_ESC_MINB
#define XMB_ESC_MINB 4
#endif
__global__ void __launch_bounds__(256, XMB_ESC_MINB) xmb_escape_kernel(const __grid_constant__ XmbHistParams P, const XmbEscParams R) {
	const int nL = NL > 0 ? NL : P.nL;
	constexpr int NLA = NL > 0 ? NL : XMB_MAX_LAYERS;
	const int iE = blockIdx.y;
	const int nE = gridDim.y;
	const double E0 = P.segs[iE].energy;
	double mus0[NLA];
	{
		const NodePos np = node_find(P, E0);
		for (int i = 0; i < nL; i++) mus0[i] = mu_lerp(P, np, i);
	}
	unsigned long long interacted = 0;
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < R.n_photons; j += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t g = (uint64_t)iE * R.n_photons + j;
		Photon p;
		double mus[NLA], rd[NLA];
		for (int i = 0; i < nL; i++) mus[i] = mus0[i];
		// ---- source (:5641-5668): point source through the slit, random polarisation ---------------------
		XmbRng rng;
		rng.init(P.seed, g, XMB_TAG_HISTORY);
		p.energy = E0; p.weight = 1.0; p.alive = true; p.n_interactions = 0;
		const double x1 = P.slit_x1_max * (-1.0 + 2.0 * rng.uniform());
		const double y1 = P.slit_y1_max * (-1.0 + 2.0 * rng.uniform());
		p.cx = p.cy = p.cz = 0.0;
		p.dx = tan(x1); p.dy = tan(y1); p.dz = 1.0;
		normalize3(p.dx, p.dy, p.dz);
		{
			double se, ce;
			sincos(rng.uniform() * M_PI * 2.0, &se, &ce);
			p.ex = ce; p.ey = se; p.ez = 0.0;
			const double cosalfa = p.ex * p.dx + p.ey * p.dy + p.ez * p.dz;
			const double c_ae = 1.0 / sin(acos(cosalfa)), c_be = -c_ae * cosalfa;
			p.ex = c_ae * p.ex + c_be * p.dx; p.ey = c_ae * p.ey + c_be * p.dy; p.ez = c_ae * p.ez + c_be * p.dz;
		}
		// xmi_photon_shift_first_layer (:1140-1186); the source sits upstream of the crystal
		{
			double d;
			if (!step_to_plane(P, p.cx, p.cy, p.cz, p.dx, p.dy, p.dz, P.layers[0].Z_begin, d)) continue;
			p.layer = 0;
		}
		// ---- first iteration: forced interaction (:1417-1518), weight_escape = weight (:1462-1464) --------
		const uint4 b0 = draw_block(P.seed, g, 1, 1, 0, 0);
		{
			const double interactionR = xmb_u01(b0.x);
			double lx = p.cx, ly = p.cy, lz = p.cz, Pabs = 0.0;
			bool ok = true;
			for (int i = 0; i < nL; i++) {     // moving towards higher layers (dirv . n > 0)
				double dist;
				if (!step_to_plane(P, lx, ly, lz, p.dx, p.dy, p.dz, P.layers[i].Z_end, dist)) { ok = false; break; }
				rd[i] = dist;
				Pabs += mus[i] * P.layers[i].density * dist;
			}
			if (!ok) continue;
			const double Pabs2 = -1.0 * expm1(-1.0 * Pabs);
			p.weight *= Pabs2;
			const double l1p = log1p(-1.0 * interactionR * Pabs2);
			const double negln = -1.0 * l1p;
			int my_index = 0;
			double my_sum = 0.0;
			for (int i = 0; i < nL; i++) {
				my_sum += mus[i] * P.layers[i].density * rd[i];
				if (my_sum > negln) { my_index = i; break; }
			}
			const double murho_idx = mus[my_index] * P.layers[my_index].density;
			double temp_sum = 0.0;
			for (int i = 0; i <= my_index; i++) temp_sum += (1.0 - (mus[i] * P.layers[i].density / murho_idx)) * rd[i];
			temp_sum = temp_sum - 1.0 * l1p / murho_idx;
			p.cx += temp_sum * p.dx; p.cy += temp_sum * p.dy; p.cz += temp_sum * p.dz;
			p.layer = my_index;
			p.n_interactions = 1;
		}
		double weight_escape = p.weight;
		interacted += esc_fixed(p.weight);   // photons_interacted (:5685-5688): every photon interacts, forced
		int type = 0, zi = 0, line = 0, shell_unused;
		// the lanes that arrive here together meet again in front of the common scatter tail (see select_and_scatter)
		select_and_scatter<NL, 1>(P, p, g, 1, mus, 1, b0.w, weight_escape, type, zi, line, shell_unused, __activemask());
		// ---- second iteration: analogue free path (:1229-1413); escaped = no interaction before the surface ----
		if (p.energy < ENERGY_THRESHOLD) continue;   // EXIT main with inside still true (:1229-1231)
		bool escaped = true;
		{
			// trip count and signed step: one copy of the loop body for both directions of flight
			const bool up = p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0;
			const int step_dir = up ? 1 : -1;
			const int n_steps = up ? nL - p.layer : p.layer + 1;
			const double interactionR = xmb_u01(draw_block(P.seed, g, 2, 1, 0, 0).x);
			double blbs = 1.0, max_random_layer = 0.0;
			double lx = p.cx, ly = p.cy, lz = p.cz;
			for (int k = 0, i = p.layer; k < n_steps; k++, i += step_dir) {
				double dist;
				if (!step_to_plane(P, lx, ly, lz, p.dx, p.dy, p.dz, up ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { escaped = false; break; }
				const double temp_prod = -1.0 * dist * P.layers[i].density * mus[i];
				const double tempexp = exp(temp_prod);
				max_random_layer = max_random_layer - blbs * expm1(temp_prod);
				if (interactionR <= max_random_layer) { escaped = false; break; }
				blbs = blbs * tempexp;
			}
		}
		if (!escaped) continue;
		if (type == 2) {
			const int ci = (int)((p.energy - R.out_min) / R.out_delta);   // 0-based (:5705-5713)
			if (ci >= 0 && ci < R.n_out) atomicAdd(&R.compt[(size_t)ci * nE + iE], esc_fixed(p.weight));
		} else if (type == 3 && line >= 1 && line <= 109) {
			atomicAdd(&R.fluo[((size_t)iE * 109 + (line - 1)) * P.nZ + zi], esc_fixed(weight_escape));
		}
	}
	interacted = warp_sum_u64(interacted);
	if ((threadIdx.x & 31) == 0 && interacted) atomicAdd(&R.interacted[iE], interacted);
}

static double g_escape_ms = 0.0;
extern "C" double xmb_escape_ratios_last_ms(void) { return g_escape_ms; }

extern "C" void xmb_free_escape_ratios(xmb_escape_ratios **p) {
	if (!p || !*p) return;
	xmb_escape_ratios *e = *p;
	free(e->Z); free(e->fluo_escape_ratios); free(e->fluo_escape_input_energies); free(e->compton_escape_ratios);
	free(e->compton_escape_output_energies);   // compton_escape_input_energies aliases fluo_escape_input_energies (:5525)
	free(e->xmi_input_string);                 // owned by the struct, as in xmi_free_escape_ratios (src/xmi_detector.c:566)
	free(e);
	*p = nullptr;
}

extern "C" int xmb_escape_ratios_run(xmb_inputFPtr esc_inputF, xmb_hdf5FPtr esc_hdf5F, const xmb_escape_ratios_options *ero,
                                     uint64_t seed, xmb_escape_ratios **out, char *input_string) {
	XmbInputF *in = xmb_as_input(esc_inputF);
	XmbHdf5F *h = xmb_as_hdf5(esc_hdf5F);
	if (!in || !h || !in->inited || !ero || !out) { xmb_set_error("xmb_escape_ratios_run: bad arguments"); return 0; }
	if (in->in.excitation->n_discrete != ero->n_input_energies || in->in.excitation->n_continuous != 0 ||
	    in->in.general->n_photons_line != ero->n_photons) {
		xmb_set_error("xmb_escape_ratios_run: handle was not made by xmb_escape_ratios_input with these options");
		return 0;
	}
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: xmb_escape_ratios_calculation has no CPU fallback"); return 0; }
	// options of the reference's escape run (:5561-5569): no M lines, no cascade
	xmb_main_options opt;
	xmb_main_options_defaults(&opt);
	opt.use_M_lines = 0; opt.use_cascade_auger = 0; opt.use_cascade_radiative = 0; opt.use_variance_reduction = 0;
	opt.escape_ratios_mode = 1;
	XmbDeviceTables *D = xmb_device_tables_get(in, h, &opt);
	if (!D) return 0;
	XmbHistParams P = D->P;
	P.seed = seed ? seed : XMB_DEFAULT_SEED;
	const int nE = (int)ero->n_input_energies, nO = (int)ero->n_compton_output_energies, nZ = P.nZ;
	const size_t n_fluo = (size_t)nE * 109 * nZ, n_compt = (size_t)nE * nO;
	unsigned long long *d_all = nullptr;
	XMB_CUDA_OK(cudaMalloc(&d_all, sizeof(unsigned long long) * (n_fluo + n_compt + nE)));
	XMB_CUDA_OK(cudaMemsetAsync(d_all, 0, sizeof(unsigned long long) * (n_fluo + n_compt + nE)));
	XmbEscParams R;
	R.n_photons = (uint64_t)ero->n_photons; R.n_out = nO; R.out_min = ero->compton_output_energy_min; R.out_delta = ero->compton_output_energy_delta;
	R.fluo = d_all; R.compt = d_all + n_fluo; R.interacted = d_all + n_fluo + n_compt;
	const int threads = 256;
	// each thread walks >= 64 photons when there are that many; the grid is nE rows of gx CTAs
	unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(64, (R.n_photons + (uint64_t)threads * 64 - 1) / ((uint64_t)threads * 64)));
	dim3 grid(gx, (unsigned)nE);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	switch (P.nL) {
	case 1: xmb_escape_kernel<1><<<grid, threads>>>(P, R); break;
	case 2: xmb_escape_kernel<2><<<grid, threads>>>(P, R); break;
	default: xmb_escape_kernel<0><<<grid, threads>>>(P, R); break;
	}
	cudaEventRecord(e1);
	XMB_CUDA_OK(cudaGetLastError());
	XMB_CUDA_OK(cudaEventSynchronize(e1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	g_escape_ms = ms;
	std::vector<unsigned long long> hst(n_fluo + n_compt + nE);
	XMB_CUDA_OK(cudaMemcpy(hst.data(), d_all, sizeof(unsigned long long) * hst.size(), cudaMemcpyDeviceToHost));
	cudaFree(d_all);
	// ---- reference-shaped result (:5521-5557, :5762-5783) ---------------------------------------------------------
	xmb_escape_ratios *er = (xmb_escape_ratios *)calloc(1, sizeof(xmb_escape_ratios));
	er->n_elements = nZ;
	er->n_fluo_input_energies = nE; er->n_compton_input_energies = nE; er->n_compton_output_energies = nO;
	er->Z = (int *)malloc(sizeof(int) * nZ);
	for (int z = 0; z < nZ; z++) er->Z[z] = h->view.Z[z];
	er->fluo_escape_input_energies = (double *)malloc(sizeof(double) * nE);
	er->compton_escape_input_energies = er->fluo_escape_input_energies;
	for (int i = 0; i < nE; i++) er->fluo_escape_input_energies[i] = ero->input_energy_min + i * ero->input_energy_delta;
	er->compton_escape_output_energies = (double *)malloc(sizeof(double) * nO);
	for (int i = 0; i < nO; i++) er->compton_escape_output_energies[i] = ero->compton_output_energy_min + i * ero->compton_output_energy_delta;
	er->fluo_escape_ratios = (double *)malloc(sizeof(double) * n_fluo);
	er->compton_escape_ratios = (double *)malloc(sizeof(double) * n_compt);
	const unsigned long long *h_fluo = hst.data(), *h_compt = hst.data() + n_fluo, *h_int = hst.data() + n_fluo + n_compt;
	// ratio of two exact integer sums; the common 2^-40 scale cancels
	for (int i = 0; i < nE; i++) {
		const double den = (double)h_int[i];
		for (size_t k = 0; k < (size_t)109 * nZ; k++) er->fluo_escape_ratios[(size_t)i * 109 * nZ + k] = (double)h_fluo[(size_t)i * 109 * nZ + k] / den;
		for (int c = 0; c < nO; c++) er->compton_escape_ratios[(size_t)c * nE + i] = (double)h_compt[(size_t)c * nE + i] / den;
	}
	er->xmi_input_string = input_string;
	*out = er;
	return 1;
}

extern "C" int xmb_escape_ratios_calculation(const xmb_input *input, xmb_escape_ratios **escape_ratios, char *input_string,
                                             const xmb_xrl_provider *xrl, const xmb_main_options *options,
                                             xmb_escape_ratios_options ero, uint64_t seed) {
	xmb_inputFPtr ein = nullptr;
	xmb_hdf5FPtr eh = nullptr;
	if (!xmb_escape_ratios_input(input, &ero, &ein)) return 0;
	if (!xmb_init_from_provider(xrl, ein, 1, &eh)) { xmb_free_input_F(&ein); return 0; }
	// the struct owns a copy of the string (the reference's driver hands a g_strdup, src/xmi_detector.c:139)
	const int rv = xmb_escape_ratios_run(ein, eh, &ero, seed, escape_ratios, input_string ? strdup(input_string) : nullptr);
	if (rv && options && options->verbose) { printf("Escape peak ratios calculation finished\n"); fflush(stdout); }
	xmb_free_hdf5_F(&eh);
	xmb_free_input_F(&ein);
	return rv;
}

