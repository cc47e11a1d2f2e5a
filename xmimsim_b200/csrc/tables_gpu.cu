// tables_gpu.cu -- the inverse-CDF part of the physics-table generator on the GPU.
//
// Replaces the integration loops of xmi_db_Z_specific (src/xmi_data_f.F90:1002-1060: scattering angle theta of Rayleigh
// and Compton scattering per element and energy, 1e5 trapezoid steps each; :1162-1186: Compton profile, 1e7 steps per
// element).  The reference evaluates xraylib at every step (94 elements x 400 energies x 1e5 angles x 2 = 7.5e9 calls);
// here the provider is sampled once per element on fine uniform grids (form factor, scattering function: 2^18 points
// in q; Compton profile: 2^20 points in pz) and the kernels interpolate.  One CTA integrates one (element, energy,
// process) row: per-thread segment sums -> block scan -> thresholds crossed per step -> the reference's "at most one
// output per step" rule (src/xmi_data_f.F90:1022-1037) as a prefix maximum.  Rows differ from the host generator by
// summation order only (a threshold may be crossed one step earlier or later).
#include <cstdio>
#include <vector>
#include "cuda_util.cuh"
#include "engine.h"

#define TG_THREADS 1024
static const double KEV2ANGST = 12.39841930;
static const double MEC2 = 510.998928;

struct TgParams {
	int nZ, nE, nR, n_fine;
	long n_theta;
	double q_max;
	const double *E;        // [nE]
	const double *ff, *sf;  // [nZ][n_fine] on q = q_max i / (n_fine - 1)
	double *rayl, *compt;   // [nZ][nE][nR]
};

__device__ __forceinline__ double tg_lerp(const double *t, int n, double x_over_max) {
	const double u = x_over_max * (n - 1);
	int i = (int)u;
	if (i > n - 2) i = n - 2;
	if (i < 0) i = 0;
	return t[i] + (t[i + 1] - t[i]) * (u - i);
}

// integrand of row (z, e, process) at step k (theta_k = pi k / (n_theta - 1)); constants that cancel are dropped
__device__ __forceinline__ double tg_f(const TgParams &P, int z, double E, bool compton, long k) {
	const double th = M_PI * (double)k / (double)(P.n_theta - 1);
	double s, c;
	sincos(th, &s, &c);
	const double q = E / KEV2ANGST * sin(th * 0.5);
	if (!compton) {
		const double F = tg_lerp(P.ff + (size_t)z * P.n_fine, P.n_fine, q / P.q_max);
		return (1.0 + c * c) * F * F * s;
	}
	const double kk = 1.0 / (1.0 + E / MEC2 * (1.0 - c));
	const double S = tg_lerp(P.sf + (size_t)z * P.n_fine, P.n_fine, q / P.q_max);
	return kk * kk * (kk + 1.0 / kk - s * s) * S * s;
}

// Generic row inversion shared by both kernels.  F(k) = integrand at step k, k = 0 .. n-1; masses m_l = (F(l) + F(l+1)) h / 2,
// l = 0 .. n-2.  out[m], m = 0 .. nR-1: abscissa l h0 of the step at which the running sum first reaches rs_m = m / (nR - 1)
// of the total, never two outputs on one step; out[0] = first, out[nR-1] = last.
template <typename Func>
__device__ void tg_invert_row(Func F, long n, double step, int nR, double *out, double last, int *s_first, double *s_scan) {
	const int tid = threadIdx.x, T = blockDim.x;
	const long n_mass = n - 1;
	const long per = (n_mass + T - 1) / T;
	const long l0 = (long)tid * per, l1 = min(n_mass, l0 + per);
	// pass 1: segment sums
	double seg = 0.0;
	if (l0 < l1) {
		double prev = F(l0);
		for (long l = l0; l < l1; l++) { const double next = F(l + 1); seg += (prev + next) * step * 0.5; prev = next; }
	}
	// exclusive block scan of the segment sums (T <= 1024)
	s_scan[tid] = seg;
	__syncthreads();
	for (int o = 1; o < T; o <<= 1) {
		const double v = tid >= o ? s_scan[tid - o] : 0.0;
		__syncthreads();
		s_scan[tid] += v;
		__syncthreads();
	}
	const double total = s_scan[T - 1];
	double run = s_scan[tid] - seg;
	for (int m = tid; m < nR; m += T) s_first[m] = (int)(n_mass - 1);
	__syncthreads();
	// pass 2: which thresholds does each step cross?  rs_m * total <= running sum after step l
	if (l0 < l1 && total > 0.0) {
		const double scale = (double)(nR - 1) / total;
		double prev = F(l0);
		// thresholds already reached by earlier segments: rs_m total <= running sum before this segment (none for the first)
		int m_next = 0;
		if (l0 > 0) {
			m_next = (int)floor(run * scale) + 1;
			while (m_next > 0 && (double)(m_next - 1) / scale > run) m_next--;
			while (m_next < nR && (double)m_next / scale <= run) m_next++;
		}
		for (long l = l0; l < l1; l++) {
			const double next = F(l + 1);
			run += (prev + next) * step * 0.5;
			prev = next;
			while (m_next < nR && (double)m_next / scale <= run) { s_first[m_next] = (int)l; m_next++; }
		}
	}
	__syncthreads();
	// "one output per step": l_m = m + max_{j <= m} (first_j - j)  (prefix maximum; nR <= 2 T handled in two strides)
	for (int m = tid; m < nR; m += T) s_first[m] -= m;
	__syncthreads();
	if (tid == 0) { int best = s_first[0]; for (int m = 0; m < nR; m++) { best = max(best, s_first[m]); s_first[m] = best; } }
	__syncthreads();
	for (int m = tid; m < nR; m += T) {
		long l = (long)s_first[m] + m;
		if (l > n_mass - 1) l = n_mass - 1;
		out[m] = (double)l * step;
	}
	__syncthreads();
	if (tid == 0) { out[0] = 0.0; out[nR - 1] = last; }
}

__global__ void __launch_bounds__(TG_THREADS) tg_theta_kernel(const TgParams P) {
	extern __shared__ unsigned char tg_smem[];
	double *s_scan = reinterpret_cast<double *>(tg_smem);
	int *s_first = reinterpret_cast<int *>(s_scan + TG_THREADS);
	const int row = blockIdx.x;                   // ((z * nE) + e) * 2 + process
	const bool compton = row & 1;
	const int e = (row >> 1) % P.nE, z = (row >> 1) / P.nE;
	const double E = P.E[e];
	double *out = (compton ? P.compt : P.rayl) + ((size_t)z * P.nE + e) * P.nR;
	const double step = M_PI / (double)(P.n_theta - 1);
	tg_invert_row([&](long k) { return tg_f(P, z, E, compton, k); }, P.n_theta, step, P.nR, out, M_PI, s_first, s_scan);
}

struct TgProfileParams {
	int n_cp, n_fine;
	long n_pz;
	double max_pz;
	const double *J;        // [nZ][n_fine] on pz = max_pz i / (n_fine - 1)
	double *icdf;           // [nZ][n_cp]
};

__global__ void __launch_bounds__(TG_THREADS) tg_profile_kernel(const TgProfileParams P) {
	extern __shared__ unsigned char tg_smem[];
	double *s_scan = reinterpret_cast<double *>(tg_smem);
	int *s_first = reinterpret_cast<int *>(s_scan + TG_THREADS);
	const int z = blockIdx.x;
	const double step = P.max_pz / (double)(P.n_pz - 1);
	const double *J = P.J + (size_t)z * P.n_fine;
	tg_invert_row([&](long k) { return tg_lerp(J, P.n_fine, (double)k / (double)(P.n_pz - 1)); }, P.n_pz, step, P.n_cp,
	              P.icdf + (size_t)z * P.n_cp, P.max_pz, s_first, s_scan);
}

static double g_tables_gpu_ms = 0.0;
extern "C" double xmb_tables_gpu_last_ms(void) { return g_tables_gpu_ms; }

// Same result as xmb_init_from_provider; the theta and Compton-profile inverse CDFs are integrated on the GPU.
extern "C" int xmb_init_from_provider_gpu(const xmb_xrl_provider *xrl, xmb_inputFPtr inputF, int quality, xmb_hdf5FPtr *out) {
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: xmb_init_from_provider_gpu has no CPU fallback (use xmb_init_from_provider)"); return 0; }
	if (!xmb_build_tables(xrl, inputF, quality, out, true)) return 0;
	XmbHdf5F *h = xmb_as_hdf5(*out);
	const xmb_tables_host &v = h->view;
	const int nZ = v.nZ, nE = v.n_icdf_E, nR = v.n_icdf_R, n_cp = v.n_cp;
	const long n_theta = quality >= 1 ? 100000 : 20000, n_pz = quality >= 1 ? 10000000 : 400000;
	const int n_fine_q = 1 << 18, n_fine_pz = 1 << 20;
	const double q_max = v.icdf_E[nE - 1] / KEV2ANGST * 1.0000001;   // the energy grid's last point lies above the source maximum
	// ---- sample the provider once per element ------------------------------------------------------------------
	std::vector<double> ff((size_t)nZ * n_fine_q), sf((size_t)nZ * n_fine_q), J((size_t)nZ * n_fine_pz);
	for (int z = 0; z < nZ; z++) { (void)xrl->FF_Rayl(v.Z[z], 0.1); (void)xrl->ComptonProfile(v.Z[z], 0.1); }
#pragma omp parallel for schedule(static) collapse(2)
	for (int z = 0; z < nZ; z++)
		for (int i = 0; i < n_fine_q; i++) {
			const double q = q_max * i / (n_fine_q - 1.0);
			ff[(size_t)z * n_fine_q + i] = xrl->FF_Rayl(v.Z[z], q);
			sf[(size_t)z * n_fine_q + i] = xrl->SF_Compt(v.Z[z], q);
		}
#pragma omp parallel for schedule(static) collapse(2)
	for (int z = 0; z < nZ; z++)
		for (int i = 0; i < n_fine_pz; i++) J[(size_t)z * n_fine_pz + i] = xrl->ComptonProfile(v.Z[z], 100.0 * i / (n_fine_pz - 1.0));
	// ---- device buffers -------------------------------------------------------------------------------------------
	double *d_ff = nullptr, *d_sf = nullptr, *d_J = nullptr, *d_E = nullptr, *d_r = nullptr, *d_c = nullptr, *d_cp = nullptr;
	const size_t n_rows = (size_t)nZ * nE * nR;
	auto fail = [&](const char *what) {
		xmb_set_error("xmb_init_from_provider_gpu: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
		cudaFree(d_ff); cudaFree(d_sf); cudaFree(d_J); cudaFree(d_E); cudaFree(d_r); cudaFree(d_c); cudaFree(d_cp);
		xmb_free_hdf5_F(out);
		return 0;
	};
	if (cudaMalloc(&d_ff, sizeof(double) * ff.size()) != cudaSuccess || cudaMalloc(&d_sf, sizeof(double) * sf.size()) != cudaSuccess ||
	    cudaMalloc(&d_J, sizeof(double) * J.size()) != cudaSuccess || cudaMalloc(&d_E, sizeof(double) * nE) != cudaSuccess ||
	    cudaMalloc(&d_r, sizeof(double) * n_rows) != cudaSuccess || cudaMalloc(&d_c, sizeof(double) * n_rows) != cudaSuccess ||
	    cudaMalloc(&d_cp, sizeof(double) * (size_t)nZ * n_cp) != cudaSuccess)
		return fail("cudaMalloc");
	cudaMemcpy(d_ff, ff.data(), sizeof(double) * ff.size(), cudaMemcpyHostToDevice);
	cudaMemcpy(d_sf, sf.data(), sizeof(double) * sf.size(), cudaMemcpyHostToDevice);
	cudaMemcpy(d_J, J.data(), sizeof(double) * J.size(), cudaMemcpyHostToDevice);
	cudaMemcpy(d_E, v.icdf_E, sizeof(double) * nE, cudaMemcpyHostToDevice);
	TgParams P{nZ, nE, nR, n_fine_q, n_theta, q_max, d_E, d_ff, d_sf, d_r, d_c};
	TgProfileParams Q{n_cp, n_fine_pz, n_pz, 100.0, d_J, d_cp};
	const size_t smem_theta = sizeof(double) * TG_THREADS + sizeof(int) * nR, smem_cp = sizeof(double) * TG_THREADS + sizeof(int) * n_cp;
	cudaFuncSetAttribute(tg_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cp);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	tg_theta_kernel<<<(unsigned)(nZ * nE * 2), TG_THREADS, smem_theta>>>(P);
	tg_profile_kernel<<<(unsigned)nZ, TG_THREADS, smem_cp>>>(Q);
	cudaEventRecord(e1);
	if (cudaGetLastError() != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) return fail("kernel");
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	g_tables_gpu_ms = ms;
	cudaMemcpy(h->rayl_theta_icdf.data(), d_r, sizeof(double) * n_rows, cudaMemcpyDeviceToHost);
	cudaMemcpy(h->compt_theta_icdf.data(), d_c, sizeof(double) * n_rows, cudaMemcpyDeviceToHost);
	cudaMemcpy(h->cp_icdf.data(), d_cp, sizeof(double) * (size_t)nZ * n_cp, cudaMemcpyDeviceToHost);
	cudaFree(d_ff); cudaFree(d_sf); cudaFree(d_J); cudaFree(d_E); cudaFree(d_r); cudaFree(d_c); cudaFree(d_cp);
	return 1;
}
