// brute.cu -- the brute-force mode of the history engine (options->use_variance_reduction = 0).
#include <algorithm>
#include "history_device.cuh"
#include "device_tables.h"

// =====================================================================================================
// Brute-force mode (options->use_variance_reduction = 0): analogue random walk; a photon is scored only
// when it reaches the detector (src/xmi_main.F90:1229-1416, :1525-1533, :1920-1984; detector / collimator
// segment tests src/xmi_aux_f.F90:1622-1833); Auger and radiative cascades spawn one offspring photon
// (src/xmi_main.F90:2413-4783), walked by the same thread after its parent.
// Random-number addresses: counter word 2 = (gen<<31)|(order<<20)|(stage<<16)|(elem<<8)|block, gen = 1 for
// the offspring's own walk; stage 1 block 0 {free path, -, -, atom}, block 1 {type, s0, s1, s2}; stage 3
// Doppler trials / photo (phi, yield check, Coster-Kronig hops); stage 4 elem 0 Auger transition + the
// parent's re-emission, elem 1 the offspring's; stage 5 radiative cascade.
// Deposits are rare (detector hits): exact 128-bit integer adds straight into the global accumulators,
// rows 0..n_int (row = interactions before detection), [nch channels | history slots].
// =====================================================================================================
#define XMB_GEN_BIT 0x800
enum { XMB_DET_NONE = 0, XMB_DET_HIT = 1, XMB_DET_COLLIMATOR = 2, XMB_DET_BAD = 3 };


__device__ __forceinline__ void add128(unsigned long long *acc, size_t slot, unsigned long long v) {
	const unsigned long long old = atomicAdd(&acc[2 * slot], v);
	if (old + v < old) atomicAdd(&acc[2 * slot + 1], 1ULL);   // carry out of the low word: exact, order independent
}

// detector frame: x along the detector normal (n_detector_orientation_inverse * (r - p_detector_window))
__device__ __forceinline__ void to_detector_frame(const XmbHistParams &P, double x, double y, double z, bool point, double *o) {
	const double *B = P.ndo_inv;
	if (point) { x -= P.p_window[0]; y -= P.p_window[1]; z -= P.p_window[2]; }
	o[0] = B[0] * x + B[1] * y + B[2] * z;
	o[1] = B[3] * x + B[4] * y + B[5] * z;
	o[2] = B[6] * x + B[7] * y + B[8] * z;
}

// xmi_check_detector_intersection (src/xmi_aux_f.F90:1622-1833) for the segment b -> e (lab coordinates)
__device__ int check_detector_intersection(const XmbHistParams &P, const XmbBruteParams &B, double bx, double by, double bz,
                                           double ex, double ey, double ez) {
	double b[3], e[3];
	to_detector_frame(P, bx, by, bz, true, b);
	to_detector_frame(P, ex, ey, ez, true, e);
	const double d0 = e[0] - b[0], d1 = e[1] - b[1], d2 = e[2] - b[2];
	if (!B.collimator_present) {
		if (b[0] * e[0] > 0) return XMB_DET_NONE;
		// (the reference assigns the scalar norm to the direction here, :1662; the segment direction is used instead)
		if (d0 == 0.0) return XMB_DET_NONE;
		const double t = (0.0 - e[0]) / d0;
		const double iy = t * d1 + e[1], iz = t * d2 + e[2];
		if (sqrt(iy * iy + iz * iz) <= P.detector_radius) return d0 >= 0.0 ? XMB_DET_BAD : XMB_DET_HIT;
		return XMB_DET_NONE;
	}
	if (d0 == 0.0) return XMB_DET_NONE;
	const double t_begin = (b[0] - e[0]) / d0, t_end = 0.0;
	const double l0 = e[0] - B.vertex_x, l1 = e[1] - B.vertex_y, l2 = e[2] - B.vertex_z;
	const double ch = cos(B.half_apex);
	const double cos2theta = ch * ch;
	const double M0 = 1.0 - cos2theta, M1 = -cos2theta;
	const double c2 = (d0 * M0) * d0 + (d1 * M1) * d1 + (d2 * M1) * d2;
	const double c1 = (d0 * M0) * l0 + (d1 * M1) * l1 + (d2 * M1) * l2;
	const double c0 = (l0 * M0) * l0 + (l1 * M1) * l1 + (l2 * M1) * l2;
	const double disc = c1 * c1 - c0 * c2;
	if (disc < 0.0) return XMB_DET_NONE;
	const double sq = sqrt(disc);
	const double t1 = (-c1 + sq) / c2, t2 = (-c1 - sq) / c2;
	const double X1x = e[0] + t1 * d0, X2x = e[0] + t2 * d0;
	const bool v1 = -(X1x - B.vertex_x) >= 0.0, v2 = -(X2x - B.vertex_x) >= 0.0;
	const double tmax = fmax(t_begin, t_end), tmin = fmin(t_begin, t_end);
	const bool in1 = t1 <= tmax && t1 >= tmin && X1x <= B.collimator_height;
	const bool in2 = t2 <= tmax && t2 >= tmin && X2x <= B.collimator_height;
	if (!v1 && !v2) return XMB_DET_NONE;
	if (v1 && v2) return (in1 || in2) ? XMB_DET_COLLIMATOR : XMB_DET_NONE;
	if (v1 ? in1 : in2) return XMB_DET_COLLIMATOR;
	const double t = (0.0 - e[0]) / d0;
	const double iy = t * d1 + e[1], iz = t * d2 + e[2];
	const double db = sqrt(b[0] * b[0] + (b[1] - iy) * (b[1] - iy) + (b[2] - iz) * (b[2] - iz));
	const double de = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
	if (sqrt(iy * iy + iz * iz) <= P.detector_radius && db <= de) return d0 >= 0.0 ? XMB_DET_BAD : XMB_DET_HIT;
	return XMB_DET_NONE;
}

// xmi_check_photon_detector_hit (src/xmi_main.F90:1920-1984): a photon that left the sample
__device__ bool check_photon_detector_hit(const XmbHistParams &P, const XmbBruteParams &B, const Photon &p) {
	if (p.dx * P.n_detector[0] + p.dy * P.n_detector[1] + p.dz * P.n_detector[2] >= 0.0) return false;
	double dd[3], cd[3];
	to_detector_frame(P, p.dx, p.dy, p.dz, false, dd);
	to_detector_frame(P, p.cx, p.cy, p.cz, true, cd);
	if (dd[0] == 0.0) return false;
	double t = (0.0 - cd[0]) / dd[0];
	double ix = t * dd[0] + cd[0], iy = t * dd[1] + cd[1], iz = t * dd[2] + cd[2];
	if (sqrt(ix * ix + iy * iy + iz * iz) > P.detector_radius) return false;
	if (!B.collimator_present) return true;
	t = (B.collimator_height - cd[0]) / dd[0];
	iy = t * dd[1] + cd[1]; iz = t * dd[2] + cd[2];
	return !(sqrt(iy * iy + iz * iz) > B.collimator_radius);
}

__device__ __forceinline__ int ck_walk(const XmbHistParams &P, int zi, int shell, SubStream &xs) {   // xmi_coster_kronig_check (:5184-5323)
	const double *ck = P.cos_kron + zi * XMB_N_CK;
	while (shell == 1 || shell == 2 || (shell >= 4 && shell <= 7)) {
		const int first = shell == 1 ? XMB_FL12 : shell == 2 ? XMB_FL23 : shell == 4 ? XMB_FM12 : shell == 5 ? XMB_FM23 : shell == 6 ? XMB_FM34 : XMB_FM45;
		const int ntr = shell == 1 ? 2 : shell == 2 ? 1 : shell == 4 ? 4 : shell == 5 ? 3 : shell == 6 ? 2 : 1;
		const double rr = xs.uniform();
		double sz = 0.0;
		int found = -1;
		for (int t = 0; t < ntr; t++) { sz += ck[first + t]; if (rr < sz) { found = t; break; } }
		if (found < 0) break;
		shell = shell + 1 + found;
	}
	return shell;
}

// one vacancy of a cascade: yield check (:5325-5350), Coster-Kronig, line (:5352-5437); returns the line or 0
__device__ int cascade_vacancy(const XmbHistParams &P, int zi, int shell, SubStream &xs) {
	if (shell > 8) return 0;
	if (shell >= 4 && !P.use_M_lines) return 0;
	if (xs.uniform() > P.fluor_yield_corr[zi * 9 + shell]) return 0;
	shell = ck_walk(P, zi, shell, xs);
	const double rl = xs.uniform();
	double sl = 0.0;
	int line = 0;
	const int lf = d_shell_line_first[shell], ll = d_shell_line_last[shell];
	for (int l = lf; l <= ll; l++) { sl += P.rad_rate[(size_t)zi * 384 + l]; if (rl < sl) { line = l; break; } }
	if (!line) return 0;
	if (P.line_energy[(size_t)zi * 384 + line] <= ENERGY_THRESHOLD) return 0;
	return line;
}

// isotropic re-emission of a cascade photon (:4455-4481, :4733-4767)
template <int NL>
__device__ void cascade_emit(const XmbHistParams &P, Photon &q, double *mus, int zi, int line, SubStream &xs) {
	const int nL = NL > 0 ? NL : P.nL;
	q.energy = P.line_energy[(size_t)zi * 384 + line];
	const NodePos lp = node_find(P, q.energy);
	for (int i = 0; i < nL; i++) mus[i] = mu_lerp(P, lp, i);
	const double theta = acos(2.0 * xs.uniform() - 1.0);
	const double phi = 2.0 * M_PI * xs.uniform();
	q.dx = sin(theta) * cos(phi); q.dy = sin(theta) * sin(phi); q.dz = cos(theta);
	const double r = 2.0 * M_PI * xs.uniform();
	q.ex = cos(r); q.ey = sin(r); q.ez = 0.0;
	const double cosalfa = q.ex * q.dx + q.ey * q.dy + q.ez * q.dz;
	const double c_ae = 1.0 / sin(acos(cosalfa)), c_be = -c_ae * cosalfa;
	q.ex = c_ae * q.ex + c_be * q.dx; q.ey = c_ae * q.ey + c_be * q.dy; q.ez = c_ae * q.ez + c_be * q.dz;
}

// Persistent lanes, phase-synchronous CTA.  The first version (one thread = one history, start to end) ran at 7.7 of
// 32 threads per instruction and 14 % issue utilisation with 8.8 warps stalled on instruction fetch
// (profiles/r1_brute_kernel_v1_*): histories differ in length and every warp sat somewhere else in ~200 KB of code.
// Here every iteration of the CTA is: refill (a lane without a photon takes its pending cascade offspring, else the
// next unsimulated photon id) | __syncthreads | analogue step + detector tests + scoring | __syncthreads | interaction
// + cascades | __syncthreads -- all lanes busy in every phase, all warps in the same code.  Photon ids are handed out
// by a warp-aggregated atomic counter; results do not depend on the assignment (fixed-address random numbers,
// integer deposits).
#ifndef XMB_BRUTE_THREADS
#define XMB_BRUTE_THREADS 1024
#endif
template <int NL, bool ADV = false>
__global__ void __launch_bounds__(XMB_BRUTE_THREADS, 1) xmb_brute_kernel(const __grid_constant__ XmbHistParams P, const XmbBruteParams B) {
	const int nL = NL > 0 ? NL : P.nL;
	constexpr int NLA = NL > 0 ? NL : XMB_MAX_LAYERS;
	const size_t acc_row = (size_t)P.nch + P.n_hist_slots;
	const int lane = threadIdx.x & 31;
	unsigned long long n_inter = 0, n_hits = 0, n_off = 0, n_noslot = 0;
	Photon p, off;
	double mus[NLA], off_mus[NLA];
	uint64_t g = 0;
	bool have = false, exhausted = false, pending_off = false, co_auger = false, co_rad = false;
	int gen_bit = 0, last_type = 0, last_zi = 0, last_line = 0, off_zi = 0, off_line = 0;
	p.alive = false; p.energy = 0.0; p.n_interactions = 0; p.layer = 0;
	for (;;) {
		// ---- phase 0: refill ---------------------------------------------------------------------------------
		if (!have && pending_off) {
			// walk the offspring next (its cascades are switched off, src/xmi_main.F90:4509-4511, :4729-4731)
			p = off;
			for (int i = 0; i < nL; i++) mus[i] = off_mus[i];
			last_type = 3; last_zi = off_zi; last_line = off_line;
			co_auger = co_rad = false;
			gen_bit = XMB_GEN_BIT;
			pending_off = false;
			have = true;
			n_off++;
		}
		{
			const bool want = !have && !exhausted;
			const unsigned m = __ballot_sync(0xffffffffu, want);
			if (m) {
				unsigned long long base = 0;
				if (lane == __ffs(m) - 1) base = atomicAdd(&P.counters[6], (unsigned long long)__popc(m));
				base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
				if (want) {
					const uint64_t lid = base + __popc(m & ((1u << lane) - 1u));
					g = shard_global_id(P, lid);
					if (lid >= P.n_local_span) exhausted = true;
					else if (g < P.n_total) {
						XmbRng rng;
						rng.init(P.seed, g, XMB_TAG_HISTORY);
						start_photon<NL>(P, p, rng, g, mus, 1);
						have = p.alive;
						gen_bit = 0;
						co_auger = B.use_auger != 0; co_rad = B.use_rad != 0;
						last_type = 0; last_zi = 0; last_line = 0;
					}
				}
			}
		}
		if (!__syncthreads_or(have ? 1 : 0)) break;
		// ---- phase 1: analogue step through the layer stack, detector / collimator tests (:1229-1416, :1525-1533) ----------
		bool interact = false, hit = false;
		uint4 b0 = make_uint4(0u, 0u, 0u, 0u);
		int order = 0;
		if (have) {
			if (p.energy < ENERGY_THRESHOLD) have = false;
			else {
				// trip count and signed step instead of a direction-dependent loop condition: one copy of the loop body for
				// both directions of flight (see the transport of the history kernel)
				const bool up = p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0;
				const int step_dir = up ? 1 : -1;
				const int n_steps = up ? nL - p.layer : p.layer + 1;
				order = (p.n_interactions + 1) | gen_bit;
				b0 = draw_block(P.seed, g, order, 1, 0, 0);
				const double interactionR = xmb_u01(b0.x);
				double blbs = 1.0, max_random_layer = 0.0;
				bool stop = false;
				for (int k = 0, i = p.layer; k < n_steps; k++, i += step_dir) {
					double nx = p.cx, ny = p.cy, nz = p.cz, dist;
					if (!step_to_plane(P, nx, ny, nz, p.dx, p.dy, p.dz, up ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { stop = true; break; }
					const double temp_prod = -1.0 * dist * P.layers[i].density * mus[i];
					const double tempexp = exp(temp_prod);
					const double min_random_layer = max_random_layer;
					max_random_layer = max_random_layer - blbs * expm1(temp_prod);
					if (interactionR <= max_random_layer) {
						dist = -1.0 * log1p(-1.0 * (interactionR - min_random_layer) / blbs) / mus[i] / P.layers[i].density;
						const double ox = p.cx, oy = p.cy, oz = p.cz;
						p.cx += dist * p.dx; p.cy += dist * p.dy; p.cz += dist * p.dz;
						const int rv = check_detector_intersection(P, B, ox, oy, oz, p.cx, p.cy, p.cz);
						if (rv == XMB_DET_COLLIMATOR || rv == XMB_DET_BAD) { stop = true; break; }
						if (rv == XMB_DET_HIT) { hit = true; stop = true; break; }
						p.layer = i;
						interact = true;
						break;
					}
					const int rv = check_detector_intersection(P, B, p.cx, p.cy, p.cz, nx, ny, nz);
					if (rv == XMB_DET_COLLIMATOR || rv == XMB_DET_BAD) { stop = true; break; }
					if (rv == XMB_DET_HIT) { hit = true; stop = true; break; }
					p.cx = nx; p.cy = ny; p.cz = nz;
					blbs = blbs * tempexp;
				}
				if (!stop && !interact) hit = check_photon_detector_hit(P, B, p);   // left the sample (:1525-1533)
				if (interact && p.n_interactions == P.n_int) interact = false;       // :1536-1539
				if (!interact) have = false;
			}
			// ---- scoring (src/xmi_main.F90:443-523) ---------------------------------------------------------------
			if (hit) {
				n_hits++;
				const unsigned long long fx = to_fixed(p.weight, P.counters);
				const int k = p.n_interactions;
				if (p.energy >= ENERGY_THRESHOLD) {
					const int ch = (int)((p.energy - P.zero) / P.gain);
					if (ch >= 0 && ch < P.nch) add128(P.acc, (size_t)k * acc_row + ch, fx);
				}
				if (k > 0) {
					int slot = -1;
					if (last_type == 1) slot = P.hist_base[last_zi];
					else if (last_type == 2) slot = P.hist_base[last_zi] + 1;
					else if (last_type == 3 && last_line) slot = B.line_slot[(size_t)last_zi * 384 + last_line];
					if (slot >= 0) add128(P.acc, (size_t)k * acc_row + P.nch + slot, fx);
					else n_noslot++;
				}
			}
		}
		__syncthreads();
		// ---- phase 2: interaction, cascades ----------------------------------------------------------------------------
		const unsigned sel_mask = __ballot_sync(0xffffffffu, have);   // the lanes that meet in front of the scatter tail
		if (have) {
			p.n_interactions++;
			n_inter++;
			double we_unused = 0.0;
			int shell = -1;
			select_and_scatter<NL, 2, ADV>(P, p, g, order, mus, 1, b0.w, we_unused, last_type, last_zi, last_line, shell, sel_mask);
			if (last_type == 4) {
				// xmi_simulate_photon_cascade_auger (:2413-4594): the primary vacancy decays without radiation
				last_type = 3;
				if (co_auger && shell >= 0 && shell <= 3) {
					// running sums of the block's rates, accumulated on the host in the reference's order (:2471-2477), so that the
					// first k with r < sum_k is found by bisection instead of a 240-step walk by the few lanes that need it
					const double *a = B.auger_rate + (size_t)last_zi * XMB_N_AUGER;
					const int first = shell == 0 ? 0 : 240 + 135 * (shell - 1), n = shell == 0 ? 240 : 135;
					SubStream xs;
					xs.init(P.seed, g, order, 4, 0);
					const double r = xs.uniform();
					int lo = 0, hi = n;                       // smallest k in [0, n) with r < a[first + k]; n if none
					while (lo < hi) { const int mid = (lo + hi) >> 1; if (r < a[first + mid]) hi = mid; else lo = mid + 1; }
					const int found = lo < n ? lo : -1;
					if (found >= 0) {
						const int new1 = shell == 0 ? 1 + found / 30 : 4 + found / 27, new2 = shell == 0 ? 1 + found % 30 : 4 + found % 27;
						off = p;   // the offspring starts as a copy of the parent (:4421-4440)
						const int l1 = cascade_vacancy(P, last_zi, new1, xs);
						if (l1) { co_auger = co_rad = false; last_line = l1; cascade_emit<NL>(P, p, mus, last_zi, l1, xs); }
						SubStream ys;
						ys.init(P.seed, g, order, 4, 1);
						const int l2 = cascade_vacancy(P, last_zi, new2, ys);
						if (l2) { cascade_emit<NL>(P, off, off_mus, last_zi, l2, ys); pending_off = true; off_zi = last_zi; off_line = l2; }
					}
				}
			} else if (last_type == 3 && last_line && co_rad) {
				// xmi_simulate_photon_cascade_radiative (:4596-4783): the vacancy the emitted line left behind
				int shell_new = -1;
				if (shell == 0) { if (last_line >= 1 && last_line <= XMB_KM5) shell_new = last_line; }
				else if (shell >= 1 && shell <= 3 && P.use_M_lines) {
					const int base = shell == 1 ? XMB_L1M1 : shell == 2 ? XMB_L2M1 : XMB_L3M1;
					if (last_line >= base && last_line <= base + 4) shell_new = 4 + (last_line - base);
				}
				if (shell_new >= 0 && !(shell_new >= 4 && !P.use_M_lines)) {
					SubStream xs;
					xs.init(P.seed, g, order, 5, 0);
					const int l = cascade_vacancy(P, last_zi, shell_new, xs);
					if (l) {
						off = p;
						co_auger = co_rad = false;
						cascade_emit<NL>(P, off, off_mus, last_zi, l, xs);
						pending_off = true; off_zi = last_zi; off_line = l;
					}
				}
			}
			if (p.energy < ENERGY_THRESHOLD) have = false;   // absorbed: the lane refills in the next phase 0 instead of idling a round
		}
		__syncthreads();
	}
	n_inter = warp_sum_u64(n_inter); n_hits = warp_sum_u64(n_hits); n_off = warp_sum_u64(n_off); n_noslot = warp_sum_u64(n_noslot);
	if (lane == 0) {
		if (n_inter) atomicAdd(&P.counters[1], n_inter);
		if (n_hits) atomicAdd(&P.counters[3], n_hits);
		if (n_off) atomicAdd(&P.counters[4], n_off);
		if (n_noslot) atomicAdd(&P.counters[5], n_noslot);
	}
}

cudaError_t xmb_brute_launch(const XmbHistParams &P, const XmbBruteParams &B, bool advanced_compton, int sms, uint64_t n_histories) {
	const int bt = XMB_BRUTE_THREADS;
	const uint64_t want = (n_histories + bt - 1) / bt;
	const unsigned bg = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms));   // persistent: one CTA per SM
	if (n_histories == 0) return cudaSuccess;
	if (advanced_compton) xmb_brute_kernel<0, true><<<bg, bt>>>(P, B);
	else switch (P.nL) {
	case 1: xmb_brute_kernel<1><<<bg, bt>>>(P, B); break;
	case 2: xmb_brute_kernel<2><<<bg, bt>>>(P, B); break;
	case 3: xmb_brute_kernel<3><<<bg, bt>>>(P, B); break;
	default: xmb_brute_kernel<0><<<bg, bt>>>(P, B); break;
	}
	return cudaGetLastError();
}
