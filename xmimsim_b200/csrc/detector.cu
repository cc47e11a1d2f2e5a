// detector.cu -- detector response on the GPU (north_star item 4).
//
// Replaces xmi_detector_convolute_all / xmi_detector_convolute_spectrum / xmi_detector_escape /
// xmi_detector_sum_peaks / xmi_detector_poisson / xmi_detector_convolute_history
// (src/xmi_detector_f.F90:56-905), i.e. the built-in twin of the plugin hook
// xmi_detector_convolute_all_custom (include/xmi_main.h:37).
//
// Stages, all fp64 and deterministic (fixed-order tree reductions, integer atomics only):
//   1. per-channel efficiency: absorbers exp(-mu rho t), crystal 1-exp(-mu rho t)         (:437-454)
//   2. escape peaks as a gather: out[t] = in[t](1-S_t) + sum_{i>t} r(i->t) in[i]            (:582-809)
//      -- the reference's ascending in-place loop only ever moves counts downwards, so every source
//         channel is read before anything is added to it; the gather form is the same map
//   3. pile-up: the sequential pulse train (:147-202) is a renewal process; 2^14 independent Philox
//      streams each simulate an equal share of the Nt pulses, closing on a group boundary like the
//      reference does
//   4. Gaussian + tail/shelf response, row-normalised, gathered per target channel        (:502-558)
//   5. optional Poisson noise, one Philox stream per (order, channel)                      (:862-905)
// All interaction orders are processed by the same launches (the reference loops over them with OpenMP,
// :255-261).  The work is O(nch^2) per order -- milliseconds; it is here for completeness of the path,
// not because it is a bottleneck.
#include <cstdio>
#include <vector>
#include <cmath>
#include <algorithm>
#include "cuda_util.cuh"
#include "xmb_lines.h"

#define DET_THREADS 128
#define PILEUP_STREAMS 16384

struct DetParams {
	int nch, n_rows, detector_type;
	double gain, zero, noise, fano, live_time, pulse_width;
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
	// fixed-order tree: deterministic
	const int tid = threadIdx.x;
	sh[tid] = v;
	__syncthreads();
	for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
		if (tid < s) sh[tid] += sh[tid + s];
		__syncthreads();
	}
	const double r = sh[0];
	__syncthreads();
	return r;
}

__global__ void det_corr_kernel(DetParams P, const double *__restrict__ corr, double *__restrict__ spec) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < P.nch * P.n_rows) spec[i] *= corr[i % P.nch];
}

// ---- escape peaks -----------------------------------------------------------------------------------
struct EscParams {
	int n_el, n_fe, n_ci, n_co;
	const double *fluo_ratios;    // (element, 109, energy), element fastest
	const double *fluo_E;
	const double *compt_ratios;   // (input, output), input fastest
	const double *compt_Ein, *compt_Eout;
	const double *edge;           // [n_el][4]  K, L1, L2, L3
	const double *line_E;         // [n_el][110]
};

__device__ __forceinline__ int findpos_lin(const double *a, int n, double x) {
	if (fabs(x - a[0]) < 1e-10) return 0;
	// arrays are uniform in every file the reference writes, but follow findpos' definition (first a[i] >= x)
	int lo = 1, hi = n - 1, ans = -1;
	while (lo <= hi) { const int mid = (lo + hi) >> 1; if (x <= a[mid]) { ans = mid - 1; hi = mid - 1; } else lo = mid + 1; }
	return ans;
}

// one block per source channel i: fills column i of M (M[t*nch + i] = r(i->t)) and S[i]
__global__ void __launch_bounds__(DET_THREADS) escape_matrix_kernel(DetParams P, EscParams E, double *__restrict__ M, double *__restrict__ S) {
	__shared__ double sh[DET_THREADS];
	const int i = blockIdx.x, tid = threadIdx.x, nch = P.nch;
	const double channel_e = ((double)(float)i + 0.5) * P.gain + P.zero;
	const double channel_1e = 0.5 * P.gain + P.zero;
	double sum_ratio = 0.0;
	// fluorescence escape: a handful of lines; thread 0 walks them in the reference's order
	if (tid == 0) {
		const int first[4] = {1, XMB_L1M1, XMB_L2M1, 86}, last[4] = {29, 58, 85, 109};
		for (int j = 0; j < E.n_el; j++)
			for (int sh_i = 0; sh_i < 4; sh_i++) {
				if (!(channel_e > E.edge[j * 4 + sh_i])) continue;
				for (int k = first[sh_i]; k <= last[sh_i]; k++) {
					const double line_e = E.line_E[j * 110 + k];
					if ((channel_e - line_e) >= channel_1e && channel_e >= E.fluo_E[0] && channel_e < E.fluo_E[E.n_fe - 1]) {
						const int pos = findpos_lin(E.fluo_E, E.n_fe, channel_e);
						if (pos < 0) continue;
						const double a1 = E.fluo_E[pos], b1 = E.fluo_E[pos + 1];
						const double a2 = E.fluo_ratios[((size_t)pos * 109 + (k - 1)) * E.n_el + j];
						const double b2 = E.fluo_ratios[((size_t)(pos + 1) * 109 + (k - 1)) * E.n_el + j];
						const double ratio = a2 + ((b2 - a2) * (channel_e - a1) / (b1 - a1));
						sum_ratio += ratio;
						const int esc = (int)((channel_e - line_e - P.zero) / P.gain);
						const bool ok = sh_i == 0 ? (esc >= 0 && esc < nch - 1) : (esc >= 0 && esc <= nch - 1);
						if (ok) M[(size_t)esc * nch + i] += ratio;
					}
				}
			}
	}
	__syncthreads();
	// Compton escape to every lower channel
	double part = 0.0;
	const double out_diff = E.compt_Eout[1] - E.compt_Eout[0];
	for (int j = tid; j < i; j += blockDim.x) {
		const double channel_c = ((double)(float)j + 0.5) * P.gain + P.zero;
		const double d = channel_e - channel_c;
		if (channel_e >= E.compt_Ein[0] && channel_e < E.compt_Ein[E.n_ci - 1] && d >= E.compt_Eout[0] && d < E.compt_Eout[E.n_co - 1]) {
			const int p1 = findpos_lin(E.compt_Ein, E.n_ci, channel_e), p2 = findpos_lin(E.compt_Eout, E.n_co, d);
			if (p1 < 0 || p2 < 0) continue;
			const double *x1 = E.compt_Ein, *x2 = E.compt_Eout;
			const double denom = (x1[p1 + 1] - x1[p1]) * (x2[p2 + 1] - x2[p2]);
			const double c1 = (x1[p1 + 1] - channel_e) * (x2[p2 + 1] - d) / denom, c2 = (channel_e - x1[p1]) * (x2[p2 + 1] - d) / denom;
			const double c3 = (x1[p1 + 1] - channel_e) * (d - x2[p2]) / denom, c4 = (channel_e - x1[p1]) * (d - x2[p2]) / denom;
			const double *A = E.compt_ratios;
			const double v = c1 * A[(size_t)p2 * E.n_ci + p1] + c2 * A[(size_t)p2 * E.n_ci + p1 + 1] + c3 * A[(size_t)(p2 + 1) * E.n_ci + p1] +
			                 c4 * A[(size_t)(p2 + 1) * E.n_ci + p1 + 1];
			const double ratio = v * P.gain / out_diff;
			M[(size_t)j * nch + i] += ratio;
			part += ratio;
		}
	}
	const double tot = block_sum(part, sh);
	if (tid == 0) S[i] = sum_ratio + tot;
}

// one block per (target channel t, row k)
__global__ void __launch_bounds__(DET_THREADS) escape_apply_kernel(DetParams P, const double *__restrict__ M, const double *__restrict__ S,
                                                                  const double *__restrict__ in, double *__restrict__ out) {
	__shared__ double sh[DET_THREADS];
	const int t = blockIdx.x, k = blockIdx.y, nch = P.nch;
	const double *row = in + (size_t)k * nch;
	double part = 0.0;
	for (int i = t + 1 + threadIdx.x; i < nch; i += blockDim.x) part += M[(size_t)t * nch + i] * row[i];
	const double tot = block_sum(part, sh);
	if (threadIdx.x == 0) out[(size_t)k * nch + t] = row[t] * (1.0 - S[t]) + M[(size_t)t * nch + t] * row[t] + tot;
}

// ---- pile-up -------------------------------------------------------------------------------------------
// cdf[k][nch] inclusive prefix sums of max(counts,0); one thread = one independent pulse-train share.  A block counts into a
// histogram in shared memory and adds it to the global row once at the end: with per-pulse global atomics the strongest
// channels (a few percent of 8e8 pulses each) serialised in L2 and set the kernel time (50.8 ms for BASELINE configs[4],
// profiles/r2_detector_kernels_ncu.json; the sums are integers, so the result is the same).
__global__ void pileup_kernel(DetParams P, const double *__restrict__ cdf, unsigned long long *__restrict__ counts, uint64_t seed, int row_first) {
	extern __shared__ unsigned int pu_hist[];   // [nch]
	const int k = blockIdx.y, nch = P.nch;
	for (int i = threadIdx.x; i < nch; i += blockDim.x) pu_hist[i] = 0u;
	__syncthreads();
	const int sid = blockIdx.x * blockDim.x + threadIdx.x;
	const double *c = cdf + (size_t)k * nch;
	const double total = c[nch - 1];
	const long long Nt_long = (long long)total;
	const long long quota = Nt_long <= 0 || sid >= PILEUP_STREAMS ? 0 : Nt_long / PILEUP_STREAMS + (sid < Nt_long % PILEUP_STREAMS ? 1 : 0);
	if (quota > 0) {
		const double mu = 1.0 / (total / P.live_time);
		XmbRng rng;
		rng.init(seed, ((uint64_t)(row_first + k) << 32) | (uint64_t)sid, XMB_TAG_DETECTOR);
		long long done = 0;
		int npulses = 0;
		long long psum = 0;   // sum of 1-based pulse channels of the open group
		int first = 0;
		for (;;) {
			npulses++; done++;
			if (npulses > 100) break;   // reference: "pulsetrain maximum reached" (aborts the run)
			const double u = rng.uniform() * total;
			int lo = 0, hi = nch - 1;
			while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid] > u) hi = mid; else lo = mid + 1; }
			if (npulses == 1) first = lo + 1;
			psum += lo + 1;
			const double deltaT = -mu * log(1.0 - rng.uniform());
			if (deltaT > P.pulse_width) {
				if (npulses == 1) atomicAdd(&pu_hist[first - 1], 1u);
				else {
					// energies_sum = sum(p*gain + zero); pulses_sum = (energies_sum - zero)/gain   (:187-191)
					const double energies_sum = (double)psum * P.gain + (double)npulses * P.zero;
					const long long pulses_sum = (long long)((energies_sum - P.zero) / P.gain);
					if (pulses_sum > 0 && pulses_sum <= nch) atomicAdd(&pu_hist[pulses_sum - 1], 1u);
				}
				if (done >= quota) break;
				npulses = 0; psum = 0;
			}
		}
	}
	__syncthreads();
	unsigned long long *out = counts + (size_t)k * nch;
	for (int i = threadIdx.x; i < nch; i += blockDim.x) if (pu_hist[i]) atomicAdd(&out[i], (unsigned long long)pu_hist[i]);
}

__global__ void prefix_kernel(DetParams P, const double *__restrict__ spec, double *__restrict__ cdf) {
	// one thread per row: sequential inclusive scan (2048 adds) -- deterministic
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= P.n_rows) return;
	double run = 0.0;
	for (int i = 0; i < P.nch; i++) { const double v = spec[(size_t)k * P.nch + i]; run += v > 0.0 ? v : 0.0; cdf[(size_t)k * P.nch + i] = run; }
}

__global__ void counts_to_double_kernel(int n, const unsigned long long *__restrict__ c, double *__restrict__ spec) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) spec[i] = (double)c[i];
}

// ---- Gaussian + tail response ----------------------------------------------------------------------------
struct RespConst { double B0, A0, A3, A4, ALFA, E0; int valid; };

__device__ __forceinline__ RespConst resp_const(const DetParams &P, int I0) {
	RespConst c;
	const double a = P.noise * P.noise, b = (2.3548) * (2.3548) * 3.85 * P.fano / 1000.0;
	const double cc = 1.4142135623730951 / (2.0 * sqrt(2.0 * log(2.0)));
	c.E0 = P.zero + P.gain * I0;
	c.valid = !(c.E0 < 1.0);
	const double FWHM = sqrt(a + b * c.E0);
	c.B0 = cc * FWHM;
	c.A0 = 1.0 / (c.B0 * 1.77245385090551602729816748334);
	c.A3 = 2.73E-3 * exp(-0.21 * c.E0) + 1.E-4;
	c.A4 = 0.000188 * exp(-0.00296 * pow(c.E0, 0.763)) + 1.355E-5 * exp(0.968 * pow(c.E0, 0.498));
	c.ALFA = 1.179 * exp(8.6E-4 * pow(c.E0, 1.877)) - 7.793 * exp(-3.81 * pow(c.E0, -0.0716));
	return c;
}
__device__ __forceinline__ double resp_R(const DetParams &P, const RespConst &c, int I) {
	const double E = P.zero + P.gain * I, X = (E - c.E0) / c.B0, G = exp(-X * X);
	if (c.E0 > 50.0) return c.A0 * G;
	const double F = erfc(X);
	const double k3 = P.detector_type == XMB_DETECTOR_SI_SDD ? 0.63 : 2.7;
	return c.A0 * G + 1.0 * (k3 * c.A3 + 15.0 * c.A4 * exp(c.ALFA * (E - c.E0))) * F;
}

// one block per source channel I0: inv_sum[I0] = 1 / sum_I R(I0 -> I)
__global__ void __launch_bounds__(DET_THREADS) response_norm_kernel(DetParams P, double *__restrict__ inv_sum) {
	__shared__ double sh[DET_THREADS];
	const int I0 = blockIdx.x;
	const RespConst c = resp_const(P, I0);
	double part = 0.0;
	if (c.valid) {
		const int last = min(I0 + 100, P.nch - 1);
		for (int I = threadIdx.x; I <= last; I += blockDim.x) part += resp_R(P, c, I);
	}
	const double tot = block_sum(part, sh);
	if (threadIdx.x == 0) inv_sum[I0] = c.valid ? 1.0 / tot : 0.0;
}

// one block per target channel I, all rows: conv[k][I] = sum_{I0 >= I-100} R(I0->I) temp[k][I0] / sum(I0)
__global__ void __launch_bounds__(DET_THREADS) convolve_kernel(DetParams P, const double *__restrict__ inv_sum, const double *__restrict__ temp,
                                                              double *__restrict__ conv) {
	__shared__ double sh[DET_THREADS];
	const int I = blockIdx.x, nch = P.nch;
	double part[8];
	for (int k = 0; k < 8; k++) part[k] = 0.0;
	for (int I0 = max(0, I - 100) + threadIdx.x; I0 < nch; I0 += blockDim.x) {
		const double is = inv_sum[I0];
		if (is == 0.0) continue;
		const RespConst c = resp_const(P, I0);
		const double r = resp_R(P, c, I) * is;
		for (int k = 0; k < P.n_rows && k < 8; k++) part[k] += r * temp[(size_t)k * nch + I0];
	}
	for (int k = 0; k < P.n_rows && k < 8; k++) {
		const double tot = block_sum(part[k], sh);
		if (threadIdx.x == 0) conv[(size_t)k * nch + I] = tot;
	}
}

// ---- Poisson ------------------------------------------------------------------------------------------------
__device__ double ran_poisson(XmbRng &r, double lam) {
	if (lam < 10.0) {
		const double L = exp(-lam);
		double p = 1.0;
		long long k = 0;
		do { k++; p *= r.uniform(); } while (p > L);
		return (double)(k - 1);
	}
	// PTRS, W. Hoermann, Insurance: Mathematics and Economics 12 (1993) 39-45
	const double slam = sqrt(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b, inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
	for (;;) {
		const double U = r.uniform() - 0.5, V = r.uniform(), us = 0.5 - fabs(U);
		const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
		if (us >= 0.07 && V <= vr) return k;
		if (k < 0 || (us < 0.013 && V > us)) continue;
		if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * log(lam) - lgamma(k + 1.0)) return k;
	}
}
__global__ void poisson_kernel(DetParams P, double *__restrict__ conv, uint64_t seed, int row_first) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.nch * P.n_rows) return;
	const double v = conv[i];
	if (v > 4294967295.0 || !(v > 1.0)) return;
	XmbRng rng;
	rng.init(seed, ((uint64_t)(1000 + row_first + i / P.nch) << 32) | (uint64_t)(i % P.nch), XMB_TAG_DETECTOR);
	conv[i] = ran_poisson(rng, v);
}

// =============================================================================================================
// Host side
// =============================================================================================================
static double host_det_corr(const xmb_input &in, const xmb_xrl_provider *xrl, double E) {
	double c = 1.0;
	for (int j = 0; j < in.absorbers->n_det_layers; j++) {
		const xmb_layer &l = in.absorbers->det_layers[j];
		c = c * std::exp(-1.0 * l.density * l.thickness * xmb_host_mu_layer(xrl, &l, E));
	}
	for (int j = 0; j < in.detector->n_crystal_layers; j++) {
		const xmb_layer &l = in.detector->crystal_layers[j];
		c = -1.0 * c * std::expm1(-1.0 * l.density * l.thickness * xmb_host_mu_layer(xrl, &l, E));
	}
	return c;
}

struct DevBuf {
	std::vector<void *> p;
	~DevBuf() { for (void *q : p) cudaFree(q); }
	template <typename T> T *alloc(size_t n) { T *d = nullptr; if (cudaMalloc(&d, sizeof(T) * (n ? n : 1)) != cudaSuccess) return nullptr; p.push_back(d); return d; }
	template <typename T> T *put(const T *h, size_t n) { T *d = alloc<T>(n); if (d && n) cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice); return d; }
};

static double g_last_det_ms = 0.0;
static uint64_t g_last_det_launches = 0;
extern "C" double xmb_detector_last_ms(void) { return g_last_det_ms; }
extern "C" uint64_t xmb_detector_last_launches(void) { return g_last_det_launches; }

// rows: n_rows spectra of nch channels, modified in place (efficiency, escape, pile-up) -> conv (n_rows x nch)
static int convolute_rows(XmbInputF *in, const xmb_xrl_provider *xrl, double *rows_h, double *conv_h, int n_rows, int row_first,
                          const xmb_main_options *opt, const xmb_escape_ratios *er, uint64_t seed) {
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: the detector response has no CPU fallback"); return 0; }
	const xmb_detector &det = *in->in.detector;
	const int nch = det.nchannels;
	if (n_rows > 8) { xmb_set_error("more than 8 spectra per call"); return 0; }
	DetParams P{nch, n_rows, det.detector_type, det.gain, det.zero, det.noise, det.fano, det.live_time, det.pulse_width};
	DevBuf B;
	std::vector<double> corr(nch);
	for (int i = 0; i < nch; i++) corr[i] = host_det_corr(in->in, xrl, i * det.gain + det.zero);
	double *d_corr = B.put(corr.data(), nch);
	double *d_spec = B.put(rows_h, (size_t)n_rows * nch);
	double *d_tmp = B.alloc<double>((size_t)n_rows * nch);
	double *d_conv = B.alloc<double>((size_t)n_rows * nch);
	double *d_inv = B.alloc<double>(nch);
	if (!d_corr || !d_spec || !d_tmp || !d_conv || !d_inv) { xmb_set_error("detector: device allocation failed"); return 0; }
	uint64_t launches = 0;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	const int tot = n_rows * nch;
	det_corr_kernel<<<(tot + 255) / 256, 256>>>(P, d_corr, d_spec); launches++;
	if (opt->use_escape_peaks == 1 && er) {
		EscParams E{};
		E.n_el = er->n_elements; E.n_fe = er->n_fluo_input_energies; E.n_ci = er->n_compton_input_energies; E.n_co = er->n_compton_output_energies;
		E.fluo_ratios = B.put(er->fluo_escape_ratios, (size_t)E.n_el * 109 * E.n_fe);
		E.fluo_E = B.put(er->fluo_escape_input_energies, E.n_fe);
		E.compt_ratios = B.put(er->compton_escape_ratios, (size_t)E.n_ci * E.n_co);
		E.compt_Ein = B.put(er->compton_escape_input_energies, E.n_ci);
		E.compt_Eout = B.put(er->compton_escape_output_energies, E.n_co);
		std::vector<double> edge((size_t)E.n_el * 4), lineE((size_t)E.n_el * 110, 0.0);
		for (int j = 0; j < E.n_el; j++) {
			for (int s = 0; s < 4; s++) edge[j * 4 + s] = xrl->EdgeEnergy(er->Z[j], s);
			for (int k = 1; k <= 109; k++) lineE[j * 110 + k] = xrl->LineEnergy(er->Z[j], -k);
		}
		E.edge = B.put(edge.data(), edge.size());
		E.line_E = B.put(lineE.data(), lineE.size());
		double *d_M = B.alloc<double>((size_t)nch * nch), *d_S = B.alloc<double>(nch);
		if (!d_M || !d_S) { xmb_set_error("detector: device allocation failed"); return 0; }
		cudaMemsetAsync(d_M, 0, sizeof(double) * nch * nch);
		escape_matrix_kernel<<<nch, DET_THREADS>>>(P, E, d_M, d_S); launches++;
		escape_apply_kernel<<<dim3(nch, n_rows), DET_THREADS>>>(P, d_M, d_S, d_spec, d_tmp); launches++;
		cudaMemcpyAsync(d_spec, d_tmp, sizeof(double) * tot, cudaMemcpyDeviceToDevice);
	}
	if (opt->use_sum_peaks == 1) {
		unsigned long long *d_cnt = B.alloc<unsigned long long>(tot);
		if (!d_cnt) { xmb_set_error("detector: device allocation failed"); return 0; }
		cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * tot);
		prefix_kernel<<<1, 32>>>(P, d_spec, d_tmp); launches++;
		pileup_kernel<<<dim3(PILEUP_STREAMS / 128, n_rows), 128, sizeof(unsigned int) * nch>>>(P, d_tmp, d_cnt, seed ? seed : XMB_DEFAULT_SEED, row_first); launches++;
		counts_to_double_kernel<<<(tot + 255) / 256, 256>>>(tot, d_cnt, d_spec); launches++;
	}
	response_norm_kernel<<<nch, DET_THREADS>>>(P, d_inv); launches++;
	convolve_kernel<<<nch, DET_THREADS>>>(P, d_inv, d_spec, d_conv); launches++;
	if (opt->use_poisson == 1) { poisson_kernel<<<(tot + 255) / 256, 256>>>(P, d_conv, seed ? seed : XMB_DEFAULT_SEED, row_first); launches++; }
	cudaEventRecord(e1);
	XMB_CUDA_OK(cudaGetLastError());
	XMB_CUDA_OK(cudaEventSynchronize(e1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	g_last_det_ms = ms;
	g_last_det_launches = launches;
	XMB_CUDA_OK(cudaMemcpy(rows_h, d_spec, sizeof(double) * tot, cudaMemcpyDeviceToHost));
	XMB_CUDA_OK(cudaMemcpy(conv_h, d_conv, sizeof(double) * tot, cudaMemcpyDeviceToHost));
	return 1;
}

// xmi_detector_convolute_history (src/xmi_detector_f.F90:291-337): host-side, a few hundred non-zero entries
static void convolute_history(XmbInputF *in, const xmb_xrl_provider *xrl, double *history) {
	const int n_int = in->in.general->n_interactions_trajectory;
	for (int k = 0; k < 100; k++)
		for (int j = 0; j < 383; j++) {
			double corr = -1.0;
			for (int i = 0; i < n_int; i++) {
				double &c = history[((size_t)k * 385 + j) * n_int + i];
				if (c > 0.0) {
					if (corr < 0.0) {
						const double le = xrl->LineEnergy(k + 1, -(j + 1));
						corr = le > 0.0 ? host_det_corr(in->in, xrl, le) : 1.0;
					}
					c *= corr;
				}
			}
		}
}

static const xmb_xrl_provider *provider_of(xmb_hdf5FPtr hdf5F) {
	XmbHdf5F *h = hdf5F ? xmb_as_hdf5(hdf5F) : nullptr;
	return h ? h->xrl : xmb_xrl_surrogate();
}

extern "C" void xmb_detector_convolute_spectrum(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, double *channels_noconv, double **channels_conv,
                                                const xmb_main_options *options, const xmb_escape_ratios *escape_ratios, int n_interactions) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in || !channels_noconv || !channels_conv || !options) { xmb_set_error("xmb_detector_convolute_spectrum: bad arguments"); return; }
	const int nch = in->in.detector->nchannels;
	double *conv = (double *)malloc(sizeof(double) * nch);
	if (!convolute_rows(in, provider_of(hdf5F), channels_noconv, conv, 1, n_interactions, options, escape_ratios, 0)) {
		free(conv);
		*channels_conv = nullptr;
		fprintf(stderr, "xmb_detector_convolute_spectrum: %s\n", xmb_last_error());
		return;
	}
	*channels_conv = conv;
}

extern "C" void xmb_detector_convolute_all(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, double **channels_noconv, double **channels_conv,
                                           double *brute_history, double *var_red_history, const xmb_main_options *options,
                                           const xmb_escape_ratios *escape_ratios, int n_interactions_all, int zero_interaction) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in || !channels_noconv || !channels_conv || !options) { xmb_set_error("xmb_detector_convolute_all: bad arguments"); return; }
	const xmb_xrl_provider *xrl = provider_of(hdf5F);
	const int nch = in->in.detector->nchannels;
	const int start = zero_interaction == 1 ? 0 : 1;      // :244-248
	const int n_rows = n_interactions_all + 1 - start;
	for (int i = 0; i <= n_interactions_all; i++) channels_conv[i] = nullptr;
	// batches of up to 8 orders per set of launches (the reference loops over orders with OpenMP, :255-261)
	for (int b0 = 0; b0 < n_rows; b0 += 8) {
		const int nb = std::min(8, n_rows - b0);
		std::vector<double> rows((size_t)nb * nch), conv((size_t)nb * nch);
		for (int r = 0; r < nb; r++) memcpy(&rows[(size_t)r * nch], channels_noconv[start + b0 + r], sizeof(double) * nch);
		if (!convolute_rows(in, xrl, rows.data(), conv.data(), nb, start + b0, options, escape_ratios, 0)) {
			fprintf(stderr, "xmb_detector_convolute_all: %s\n", xmb_last_error());
			return;
		}
		for (int r = 0; r < nb; r++) {
			memcpy(channels_noconv[start + b0 + r], &rows[(size_t)r * nch], sizeof(double) * nch);   // in place, as the reference (:412-413)
			double *c = (double *)malloc(sizeof(double) * nch);
			memcpy(c, &conv[(size_t)r * nch], sizeof(double) * nch);
			channels_conv[start + b0 + r] = c;
		}
	}
	if (options->use_variance_reduction == 1 && var_red_history) {
		if (options->verbose == 1) printf("Calculating variance reduction history detector absorption correction\n");
		convolute_history(in, xrl, var_red_history);
	}
	if (brute_history) {
		if (options->verbose == 1) printf("Calculating brute force history detector absorption correction\n");
		convolute_history(in, xrl, brute_history);
	}
}

extern "C" void xmb_detector_convolute_history(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, double *history, const xmb_main_options *options) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in || !history) return;
	(void)options;
	convolute_history(in, provider_of(hdf5F), history);
}
