// solid_angle.cu -- solid-angle grid on the GPU.
//
// Replaces the OpenCL plugin xmi_solid_angle_calculation_cl (src/xmi_solid_angle_cl.c:118-439,
// kernel src/xmi_kernels.cl:219-452) and its Fortran fallback (src/xmi_solid_angle_f.F90:303-710).
//
// Mapping: one warp per (r, theta) grid point; the point's hits_per_single rays are dealt to the 32
// lanes in Philox blocks (one Philox4x32-10 call = 4 words = 2 rays, as the OpenCL kernel's
// Threefry use, src/xmi_kernels.cl:395-406); hits are summed with one warp reduction.  The whole
// 1024 x 1024 grid is ONE launch (the OpenCL host issued 64 blocking launches, :379-401).
// Arithmetic is fp64 like the Fortran path: the fp32 OpenCL kernel quantises narrow cones
// (1 - rn*(1-cos(apex)) has only (1-cos(apex))/6e-8 distinct values in fp32).  No input traffic
// beyond two 8 KB axes: the kernel is issue-bound.
#include <cstdio>
#include <vector>
#include "cuda_util.cuh"
#include "solid_angle_device.cuh"

struct SaParams {
	SaDetector det;
	long n_r, n_theta, hits_per_single;
	uint64_t seed;
};

__global__ void __launch_bounds__(256) xmb_solid_angle_kernel(SaParams P, const double *__restrict__ r_vals,
                                                             const double *__restrict__ theta_vals, long theta_begin,
                                                             long theta_end, double *__restrict__ solid_angles,
                                                             int *__restrict__ hits_out) {
	const int lane = threadIdx.x & 31;
	const long warps_per_grid = (long)gridDim.x * (blockDim.x >> 5);
	const long n_points = (theta_end - theta_begin) * P.n_r;
	const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
	const double det_r2 = P.det.detector_radius * P.det.detector_radius, col_r2 = P.det.collimator_radius * P.det.collimator_radius;
	for (long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_points; w += warps_per_grid) {
		const long it = theta_begin + w / P.n_r, ir = w % P.n_r;
		const uint64_t id = (uint64_t)it * (uint64_t)P.n_r + (uint64_t)ir;
		const SaCone cone = sa_cone_setup(P.det, r_vals[ir], theta_vals[it]);
		if (cone.dead) {   // warp-uniform
			if (lane == 0) { solid_angles[id] = 0.0; if (hits_out) hits_out[id] = 0; }
			continue;
		}
		int hits = 0;
		const long n_pairs = (P.hits_per_single + 1) >> 1;
		for (long p = lane; p < n_pairs; p += 32) {
			const uint4 rnd = xmb_philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)p, XMB_TAG_SOLID_ANGLE), key);
			hits += sa_pair_hits(cone, det_r2, col_r2, rnd, 2 * p + 1 < P.hits_per_single);
		}
		hits = __reduce_add_sync(0xffffffffu, hits);
		if (lane == 0) {
			solid_angles[id] = cone.cone_sa * (double)hits / (double)P.hits_per_single;
			if (hits_out) hits_out[id] = hits;
		}
	}
}

// last grid's raw hit counts, kept on the host for parity tests
static std::vector<int32_t> g_last_hits;
static double g_last_sa_ms = 0.0;

extern "C" long xmb_solid_angle_last_hits(int32_t *hits, long capacity) {
	long n = (long)g_last_hits.size();
	if (hits) for (long i = 0; i < n && i < capacity; i++) hits[i] = g_last_hits[i];
	return n;
}
extern "C" double xmb_solid_angle_last_ms(void) { return g_last_sa_ms; }

extern "C" int xmb_cuda_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// Grid over caller-provided axes (any n_r x n_theta); the plugin entry below wraps it.
extern "C" int xmb_solid_angle_grid(xmb_inputFPtr inputF, const double *r_vals, long n_r, const double *theta_vals,
                                    long n_theta, long hits_per_single, uint64_t seed, int verbose,
                                    double *solid_angles, int32_t *hits) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in || !in->inited) { xmb_set_error("xmb_solid_angle_grid: input not initialised"); return 0; }
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: the solid-angle grid has no CPU fallback"); return 0; }
	SaParams P;
	P.det.collimator_present = in->der.collimator_present;
	P.det.detector_radius = in->der.detector_radius;
	P.det.collimator_radius = in->der.collimator_radius;
	P.det.collimator_height = in->der.collimator_height;
	P.n_r = n_r; P.n_theta = n_theta; P.hits_per_single = hits_per_single;
	P.seed = seed ? seed : XMB_DEFAULT_SEED;
	double *d_r = nullptr, *d_t = nullptr, *d_sa = nullptr;
	int *d_hits = nullptr;
	const size_t n = (size_t)n_r * n_theta;
	XMB_CUDA_OK(cudaMalloc(&d_r, sizeof(double) * n_r));
	XMB_CUDA_OK(cudaMalloc(&d_t, sizeof(double) * n_theta));
	XMB_CUDA_OK(cudaMalloc(&d_sa, sizeof(double) * n));
	XMB_CUDA_OK(cudaMalloc(&d_hits, sizeof(int) * n));
	XMB_CUDA_OK(cudaMemcpy(d_r, r_vals, sizeof(double) * n_r, cudaMemcpyHostToDevice));
	XMB_CUDA_OK(cudaMemcpy(d_t, theta_vals, sizeof(double) * n_theta, cudaMemcpyHostToDevice));
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int blocks = sms * 8;   // 8 resident CTAs of 256 threads per SM
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	if (!verbose) {
		xmb_solid_angle_kernel<<<blocks, 256>>>(P, d_r, d_t, 0, n_theta, d_sa, d_hits);
	} else {
		// progress lines the reference prints (src/xmi_solid_angle_cl.c:398-399): ten row chunks
		for (int c = 0; c < 10; c++) {
			long t0 = n_theta * c / 10, t1 = n_theta * (c + 1) / 10;
			if (t1 > t0) xmb_solid_angle_kernel<<<blocks, 256>>>(P, d_r, d_t, t0, t1, d_sa, d_hits);
			XMB_CUDA_OK(cudaStreamSynchronize(0));
			printf("Solid angle calculation at %3i %%\n", (c + 1) * 10);
			fflush(stdout);
		}
	}
	cudaEventRecord(e1);
	XMB_CUDA_OK(cudaGetLastError());
	XMB_CUDA_OK(cudaEventSynchronize(e1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	g_last_sa_ms = ms;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	XMB_CUDA_OK(cudaMemcpy(solid_angles, d_sa, sizeof(double) * n, cudaMemcpyDeviceToHost));
	g_last_hits.resize(n);
	XMB_CUDA_OK(cudaMemcpy(g_last_hits.data(), d_hits, sizeof(int) * n, cudaMemcpyDeviceToHost));
	if (hits) memcpy(hits, g_last_hits.data(), sizeof(int32_t) * n);
	cudaFree(d_r); cudaFree(d_t); cudaFree(d_sa); cudaFree(d_hits);
	if (verbose) { printf("Solid angle calculation finished\n"); fflush(stdout); }
	return 1;
}

extern "C" int xmb_solid_angle_calculation(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, xmb_solid_angle **solid_angle,
                                           char *input_string, const xmb_main_options *options, long hits_per_single,
                                           uint64_t seed) {
	if (!solid_angle) return 0;
	xmb_solid_angle *sa = nullptr;
	if (!xmb_solid_angle_inputs(inputF, hdf5F, &sa)) return 0;
	if (!xmb_solid_angle_grid(inputF, sa->grid_dims_r_vals, sa->grid_dims_r_n, sa->grid_dims_theta_vals,
	                          sa->grid_dims_theta_n, hits_per_single > 0 ? hits_per_single : 5000, seed,
	                          options ? options->verbose : 0, sa->solid_angles, nullptr)) {
		sa->xmi_input_string = nullptr;
		xmb_free_solid_angle(sa);
		return 0;   // reference semantics: 0 = fall through to the next backend
	}
	sa->xmi_input_string = input_string;
	*solid_angle = sa;
	return 1;
}
