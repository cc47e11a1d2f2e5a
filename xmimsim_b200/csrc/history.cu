// history.cu -- photon histories with forced detection on the GPU (north_star items 1 and 2).
//
// Replaces the OpenMP photon loops of xmi_main_msim (src/xmi_main.F90:280-867) and everything they
// call per photon: source sampling (:957-1186), xmi_simulate_photon's variance-reduction branch
// (:1188-1685), Rayleigh / Compton / photo-electric interactions with Coster-Kronig and line selection
// (:1986-2411, :4985-5437) and the forced-detection scoring xmi_variance_reduction
// (src/xmi_variance_reduction.F90:29-1101) with the solid-angle lookup (src/xmi_solid_angle_f.F90:712-801).
//
// Mapping: one thread = one history; a warp owns 32 consecutive global photon ids and walks
// the interaction orders in lock step (with forced interactions every live photon interacts exactly once
// per iteration, so the interaction order is warp-uniform).  The forced-detection loops over layers,
// elements, shells and line records are warp-uniform; lanes that are dead or sit in another layer
// contribute zero.  Every deposit is converted to 2^-56 fixed point per lane, summed exactly across the
// warp with integer shuffles and added with ONE 128-bit (lo/hi + carry) atomic: the totals are
// independent of scheduling, launch shape and GPU count, bit for bit.
// XRF deposits only touch the per-line history slot; the channel spectrum is rebuilt from those slots in
// the epilogue (a line's channel is a constant), which halves the atomics of the inner loop.
// Random numbers: Philox4x32-10 keyed by the run seed; every draw of photon g has a fixed counter address
// (g, interaction order, stage, element, block), see draw_block().
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cmath>
#include "cuda_util.cuh"
#include "history.cuh"
#include "xmb_lines.h"

#define ENERGY_THRESHOLD 1.0
#define ENERGY_MAX 200.0
#define XMI_MEC2 (9.10938188e-31 * 2.99792458e8 * 2.99792458e8 / 1.602176487e-19 / 1000.0)
#define KEV2ANGST 12.39841930
#define AVOGNUM 0.602252
#define RE2 0.07940775
#ifndef HIST_THREADS
#define HIST_THREADS 1024        // one CTA per SM: every warp of the SM is in the same phase (I-cache locality)
#endif
#ifndef XMB_REC_UNROLL
#define XMB_REC_UNROLL 2
#endif
#define XMB_MAX_ORDERS 64
#define XMB_STATE_FIELDS 15      // 13 doubles of photon state + photon id + layer (mus[nL] follow)
#define XMB_PRAGMA(x) _Pragma(#x)
#define XMB_UNROLL_NL _Pragma("unroll")
#define XMB_UNROLL(n) XMB_PRAGMA(unroll n)
#ifndef HIST_MIN_BLOCKS
#define HIST_MIN_BLOCKS 1
#endif

__constant__ short d_shell_line_first[9] = {1, 30, 59, 86, 118, 140, 161, 182, 201};   // = xmb_shell_line_first
__constant__ short d_shell_line_last[9] = {29, 58, 85, 113, 136, 158, 180, 200, 219};    // = xmb_shell_line_last

struct NodePos { int pos; double f; };

__device__ __forceinline__ NodePos node_find(const XmbHistParams &P, double E) {
	int b = (int)floor((E - P.bucket_E0) * P.bucket_inv_dE);
	b = max(0, min(b, P.n_buckets - 1));
	// bit 31 of bucket_start marks a bucket whose only node is the uniform-grid node at its lower bound and whose
	// upper bound is the next node: the bracket is known without scanning (one dependent load less on the chain)
	const int bs = P.bucket_start[b];
	int i = bs & 0x7FFFFFFF;
	if (bs >= 0 || E < P.node_E[i] || E >= P.node_E[i + 1]) {
		while (i > 0 && P.node_E[i] > E) i--;
		while (i + 1 < P.n_nodes - 1 && P.node_E[i + 1] <= E) i++;
		i = min(i, P.n_nodes - 2);
	}
	NodePos p;
	p.pos = i;
	const double e0 = P.node_E[i], e1 = P.node_E[i + 1];
	p.f = (E - e0) / (e1 - e0);
	return p;
}
__device__ __forceinline__ double row_lerp(const XmbHistParams &P, NodePos np, int off) {
	const double *r0 = P.rows + (size_t)np.pos * P.row_stride + off;
	const double a = r0[0], b = r0[P.row_stride];
	return a + (b - a) * np.f;
}

// findpos on a uniform axis with the reference's interval convention axis(i) < x <= axis(i+1)
// (src/xmi_aux_f.F90:1305-1335)
__device__ __forceinline__ int findpos_uniform(const double *ax, int n, double x) {
	const double x0 = ax[0], dx = ax[1] - ax[0];
	if (fabs(x - x0) < 1e-10) return 0;
	int i = (int)ceil((x - x0) / dx) - 1;
	i = max(0, min(i, n - 2));
	while (i > 0 && x <= x0 + dx * i) i--;
	while (i < n - 2 && x > x0 + dx * (i + 1)) i++;
	return i;
}
// bilinear_interpolation (src/xmi_aux_f.F90:1337-1428); a[i1][i2], i2 fastest
__device__ __forceinline__ double bilinear(const double *a, int n2, const double *ax1, int n1, const double *ax2, double x1, double x2) {
	const int p1 = findpos_uniform(ax1, n1, x1), p2 = findpos_uniform(ax2, n2, x2);
	const double a1l = ax1[p1], a1h = ax1[p1 + 1], a2l = ax2[p2], a2h = ax2[p2 + 1];
	const double denom = (a1h - a1l) * (a2h - a2l);
	const double c1 = (a1h - x1) * (a2h - x2) / denom, c2 = (x1 - a1l) * (a2h - x2) / denom;
	const double c3 = (a1h - x1) * (x2 - a2l) / denom, c4 = (x1 - a1l) * (x2 - a2l) / denom;
	const double *q = a + (size_t)p1 * n2 + p2;
	return c1 * q[0] + c2 * q[n2] + c3 * q[1] + c4 * q[n2 + 1];
}

// ---- exact accumulation ---------------------------------------------------------------------------
// Every deposit is a non-negative 2^-56 fixed-point integer.  Deposits of the batch a CTA is working on (one
// interaction order) are staged in shared memory: a slot is two 64-bit words, A accumulates the low 32 bits of
// each addend and B the high 32 bits, so a deposit is two carry-free shared-memory atomics (total = A + (B<<32),
// exact for < 2^32 addends of < 2^63).  After the batch the CTA folds every non-zero slot into the global 128-bit
// (lo, hi) accumulator.  Ablation on B200 (profiles/r1_history_ablation.txt): with per-lane global REDs the Compton
// peak's ~50 hot channel words serialised in L2 and cost 47 % of the kernel.
__device__ __forceinline__ unsigned long long to_fixed(double w, unsigned long long *counters) {
	const double s = w * 72057594037927936.0;   // 2^56
	if (!(s < 2.8e17)) { if (s == s) atomicAdd(&counters[2], 1ULL); return 0ULL; }   // w >= ~4: counted, never wrapped
	return __double2ull_rn(s);
}
// hot-loop variant: no branch; out-of-range / NaN inputs are flagged in `bad` (reported once per thread at the end)
__device__ __forceinline__ unsigned long long to_fixed_fast(double w, bool &bad) {
	const double s = w * 72057594037927936.0;
	bad |= !(s < 2.8e17);
	return __double2ull_rn(fmin(s, 2.8e17));
}
// staged deposit: four native 32-bit shared-memory atomics on the 16-bit pieces of v (64-bit shared atomics are CAS
// spin loops on sm_100a -- ATOMS.CAST.SPIN.64 -- and collapse when the lanes of a warp hit the same channel)
__device__ __forceinline__ void red128(unsigned int *stage, size_t slot, unsigned long long v) {
	if (v == 0ULL) return;
	unsigned int *w = stage + 4 * slot;
	atomicAdd(&w[0], (unsigned int)(v & 0xFFFFULL));
	atomicAdd(&w[1], (unsigned int)((v >> 16) & 0xFFFFULL));
	atomicAdd(&w[2], (unsigned int)((v >> 32) & 0xFFFFULL));
	const unsigned int top = (unsigned int)(v >> 48);
	if (top) atomicAdd(&w[3], top);
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
// all 32 lanes call; slot is warp-uniform.  The four 16-bit pieces are summed across the warp with REDUX (sums < 2^21)
// and lane 0 adds them to the slot's four staging words.
#ifndef XMB_REDUX_PIECES
#define XMB_REDUX_PIECES 1
#endif
__device__ __forceinline__ void deposit_uniform(unsigned int *acc, size_t slot, unsigned long long v, int lane) {
#if XMB_REDUX_PIECES
	const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
	const unsigned int s0 = __reduce_add_sync(0xffffffffu, lo & 0xFFFFu), s1 = __reduce_add_sync(0xffffffffu, lo >> 16);
	const unsigned int s2 = __reduce_add_sync(0xffffffffu, hi & 0xFFFFu), s3 = __reduce_add_sync(0xffffffffu, hi >> 16);
	if (lane == 0) {
		unsigned int *w = acc + 4 * slot;
		if (s0) atomicAdd(&w[0], s0);
		if (s1) atomicAdd(&w[1], s1);
		if (s2) atomicAdd(&w[2], s2);
		if (s3) atomicAdd(&w[3], s3);
	}
#else
	v = warp_sum_u64(v);
	if (lane == 0) red128(acc, slot, v);
#endif
}
// all 32 lanes call; slot may differ per lane (slot < 0: nothing to add)
__device__ __forceinline__ void deposit_varying(unsigned int *acc, long slot, unsigned long long v, int lane) {
	const long s0 = __shfl_sync(0xffffffffu, slot, 0);
	if (__all_sync(0xffffffffu, slot == s0)) {
		if (s0 >= 0) deposit_uniform(acc, (size_t)s0, v, lane);
	} else if (slot >= 0) red128(acc, (size_t)slot, v);
}
// fold the CTA's staged slots into the global (lo, hi) accumulators of interaction order `order` and clear them
__device__ __forceinline__ void flush_staged(unsigned int *stage, unsigned long long *global_row, int n_slots, int tid, int T) {
	for (int i = tid; i < n_slots; i += T) {
		const uint4 w = *reinterpret_cast<uint4 *>(stage + 4 * i);
		if ((w.x | w.y | w.z | w.w) == 0u) continue;
		*reinterpret_cast<uint4 *>(stage + 4 * i) = make_uint4(0u, 0u, 0u, 0u);
		// total = w0 + w1 2^16 + w2 2^32 + w3 2^48 as a 128-bit integer
		const unsigned long long t01 = (unsigned long long)w.x + ((unsigned long long)w.y << 16);     // < 2^49
		const unsigned long long t2 = (unsigned long long)w.z << 32, t3 = (unsigned long long)w.w << 48;
		unsigned long long lo = t01 + t2;
		unsigned long long hi = (lo < t2 ? 1ULL : 0ULL) + ((unsigned long long)w.w >> 16);
		const unsigned long long lo2 = lo + t3;
		if (lo2 < lo) hi++;
		lo = lo2;
		const unsigned long long old = atomicAdd(&global_row[2 * i], lo);
		if (old + lo < old) hi++;
		if (hi) atomicAdd(&global_row[2 * i + 1], hi);
	}
}

// Random-number layout: counter = (photon id lo, hi, (order << 20) | (stage << 16) | (element << 8) | block, tag);
// see DESIGN.md "Random numbers".  Every draw has a fixed address: all lanes of a warp generate their blocks at
// the same program point (no divergent refills) and no generator state lives across the interaction loop.
__device__ __forceinline__ uint4 draw_block(uint64_t seed, uint64_t g, int order, int stage, int elem, int block) {
	return xmb_philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32),
	                                    ((uint32_t)order << 20) | ((uint32_t)stage << 16) | ((uint32_t)elem << 8) | (uint32_t)block,
	                                    XMB_TAG_HISTORY),
	                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
struct SubStream {   // sequential cursor over the blocks of one (order, stage, element) sub-stream
	uint64_t seed, g;
	int order, stage, elem, block, have;
	uint4 buf;
	__device__ __forceinline__ void init(uint64_t seed_, uint64_t g_, int order_, int stage_, int elem_) {
		seed = seed_; g = g_; order = order_; stage = stage_; elem = elem_; block = 0; have = 0;
	}
	__device__ __forceinline__ double uniform() {
		if (have == 0) { buf = draw_block(seed, g, order, stage, elem, block & 0xFF); block++; have = 4; }
		const uint32_t w = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
		have--;
		return xmb_u01(w);
	}
};

struct Photon {
	double cx, cy, cz, dx, dy, dz, ex, ey, ez;
	double energy, weight, theta, phi;
	int layer;
	int n_interactions;
	bool alive;
};

__device__ __forceinline__ void normalize3(double &x, double &y, double &z) {
	const double n = sqrt(x * x + y * y + z * z);
	x /= n; y /= n; z /= n;
}

// xmi_update_photon_dirv (src/xmi_main.F90:5071-5148)
__device__ __forceinline__ void update_dirv(Photon &p, double theta_i, double phi_i) {
	double phi_new = phi_i;
	if (phi_i > 2.0 * M_PI) phi_new = phi_i - 2.0 * M_PI;
	else if (phi_i < 0.0) phi_new = phi_i + 2.0 * M_PI;
	double sph, cph, sth, cth, sti, cti, spn, cpn;
	sincos(p.phi, &sph, &cph);
	sincos(p.theta, &sth, &cth);
	sincos(theta_i, &sti, &cti);
	sincos(phi_new, &spn, &cpn);
	const double v0 = sti * cpn, v1 = sti * spn, v2 = cti;
	p.dx = cth * cph * v0 + (-sph) * v1 + sth * cph * v2;
	p.dy = cth * sph * v0 + cph * v1 + sth * sph * v2;
	p.dz = (-sth) * v0 + 0.0 * v1 + cth * v2;
	normalize3(p.dx, p.dy, p.dz);
	p.theta = acos(p.dz);
	p.phi = atan2(p.dy, p.dx);
	if (p.phi > 2.0 * M_PI) p.phi -= 2.0 * M_PI;
	else if (p.phi < 0.0) p.phi += 2.0 * M_PI;
}
// xmi_update_photon_elecv (:5150-5182)
__device__ __forceinline__ void update_elecv(Photon &p) {
	const double cosalfa = p.dx * p.ex + p.dy * p.ey + p.dz * p.ez;
	const double sinalfa = sin(acos(cosalfa));
	const double c_ae = 1.0 / sinalfa, c_be = -c_ae * cosalfa;
	p.ex = c_ae * p.ex + c_be * p.dx;
	p.ey = c_ae * p.ey + c_be * p.dy;
	p.ez = c_ae * p.ez + c_be * p.dz;
	normalize3(p.ex, p.ey, p.ez);
}
// phi0 of the electric vector in the photon frame (:2055-2066)
__device__ __forceinline__ double elec_phi0(const Photon &p) {
	double sph, cph, sth, cth;
	sincos(p.phi, &sph, &cph);
	sincos(p.theta, &sth, &cth);
	double cosphi0 = p.ex * (cph * cth) + p.ey * (cth * sph) + p.ez * (-sth);
	const double sinphi0 = p.ex * sph + p.ey * (-cph) + p.ez * 0.0;
	if (fabs(cosphi0) > 1.0) cosphi0 = cosphi0 > 0 ? 1.0 : -1.0;
	double phi0 = acos(cosphi0);
	if (sinphi0 > 0.0) phi0 = -phi0;
	return phi0;
}

// Doppler-broadened Compton energy (src/xmi_main.F90:4985-5067; forced-detection variant
// src/xmi_variance_reduction.F90:1010-1101)
// sth2 = sin(theta/2) and c_lamb0 = 1.2399e-6 / (1000 E0) are hoisted by the callers (same for every element).
// Two trials per Philox block: (pz, sign), (pz, sign).
// The first half-trial's random block and its two inverse-CDF entries may be handed in (software prefetch by the
// element loop: the gather of element e+1 overlaps the dependent chain of element e).
struct ComptonPrefetch { uint4 w; double i0, i1; int zi; double F0, F1, S0, S1; };

__device__ __forceinline__ void compton_prefetch(const XmbHistParams &P, int zi, uint64_t g, int order, int elem, int qi, ComptonPrefetch &pf) {
	pf.zi = zi;
	pf.w = draw_block(P.seed, g, order, 2, elem, 0);
	const int pos = min((int)(xmb_u01(pf.w.x) * P.cp_inv_dR), P.n_cp - 2);
	const double *icdf = P.cp_icdf + (size_t)zi * P.n_cp + pos;
	pf.i0 = icdf[0]; pf.i1 = icdf[1];
	const double *f = P.ff + (size_t)zi * P.n_q + qi, *sfp = P.sf + (size_t)zi * P.n_q + qi;
	pf.F0 = f[0]; pf.F1 = f[1]; pf.S0 = sfp[0]; pf.S1 = sfp[1];
}

__device__ __forceinline__ double compton_energy(const XmbHistParams &P, int zi, double E0, double c_lamb0, double sth2, uint64_t g, int order,
                                                 int stage, int elem, bool varred, const ComptonPrefetch *pf = nullptr) {
	const double cc = 1.2399E-6, c0 = 4.85E-12, c1 = 1.456E-2;
	const double *icdf = P.cp_icdf + (size_t)zi * P.n_cp;
	const double shift = c0 * sth2 * sth2, slope = c1 * c_lamb0 * sth2;
	double energy = 0.0;
	int tries = 0;
	for (int blk = 0;; blk++) {
		const uint4 w = (pf && blk == 0) ? pf->w : draw_block(P.seed, g, order, stage, elem, blk & 0xFF);
		bool done = false;
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const double r = xmb_u01(h ? w.z : w.x), rs = xmb_u01(h ? w.w : w.y);
			const double rs_ = r * P.cp_inv_dR;      // uniform axis: position in units of the step
			int pos = (int)rs_;
			if (varred && pos == P.n_cp - 2) continue;
			pos = min(pos, P.n_cp - 2);
			const bool use_pf = pf && blk == 0 && h == 0;
			const double ia = use_pf ? pf->i0 : icdf[pos], ib = use_pf ? pf->i1 : icdf[pos + 1];
			double pz = ia + (ib - ia) * (rs_ - pos);
			if (rs < 0.5) pz = -pz;
			const double c_lamb = c_lamb0 + (shift - slope * pz);
			energy = (cc / 1000.0) / c_lamb;
			if (energy <= E0 || tries == (varred ? 100 : 500)) { done = true; break; }
			tries++;
		}
		if (done) break;
	}
	return energy;
}

// xmi_get_solid_angle (src/xmi_solid_angle_f.F90:712-801); off-grid points are counted and score zero
__device__ __forceinline__ double get_solid_angle(const XmbHistParams &P, const Photon &p) {
	double vx = p.cx - P.p_window[0], vy = p.cy - P.p_window[1], vz = p.cz - P.p_window[2];
	const double r = sqrt(vx * vx + vy * vy + vz * vz);
	normalize3(vx, vy, vz);
	double temp_theta = acos(vx * P.n_detector[0] + vy * P.n_detector[1] + vz * P.n_detector[2]);
	if (temp_theta > M_PI / 2.0) temp_theta = M_PI - temp_theta;
	const double theta = (M_PI / 2.0) - temp_theta;
	const double *R = P.sa_r_vals, *Th = P.sa_t_vals;
	if (theta < Th[0]) return 0.0;
	if (r > R[P.sa_nr - 1] || r < R[0] - 1e-10 || theta > Th[P.sa_nt - 1]) { atomicAdd(&P.counters[0], 1ULL); return 0.0; }
	const int p1 = findpos_uniform(R, P.sa_nr, r), p2 = findpos_uniform(Th, P.sa_nt, theta);
	const double rl = R[p1], rh = R[p1 + 1], tl = Th[p2], th = Th[p2 + 1];
	const double denom = (rh - rl) * (th - tl);
	const double c1 = (rh - r) * (th - theta) / denom, c2 = (r - rl) * (th - theta) / denom;
	const double c3 = (rh - r) * (theta - tl) / denom, c4 = (r - rl) * (theta - tl) / denom;
	const double *A = P.sa_grid + (size_t)p2 * P.sa_nr + p1;
	return c1 * A[0] + c2 * A[1] + c3 * A[P.sa_nr] + c4 * A[P.sa_nr + 1];
}

__device__ __forceinline__ double ran_gaussian(XmbRng &rng, double sigma) {
	const double u1 = rng.uniform(), u2 = rng.uniform();
	return sigma * sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2);
}

// local index of this rank -> global photon id (block-cyclic: block b of XMB_SHARD_BLOCK ids belongs to rank b % n)
__device__ __forceinline__ uint64_t shard_global_id(const XmbHistParams &P, uint64_t lid) {
	return (((lid >> XMB_SHARD_SHIFT) * (uint64_t)P.shard_n + (uint64_t)P.shard_rank) << XMB_SHARD_SHIFT) | (lid & (XMB_SHARD_BLOCK - 1));
}

// ---- source sampling (src/xmi_main.F90:319-438, :579-724, :957-1186) -----------------------------------
template <int NL>
__device__ void start_photon(const XmbHistParams &P, Photon &p, XmbRng &rng, uint64_t g, double *mus /* [nL] stride T */, int T) {
	const int nL = NL > 0 ? NL : P.nL;
	int s;
	uint64_t j;
	const uint64_t n_cont = P.n_cont_seg * P.n_per_interval;
	if (g < n_cont) { s = (int)(g / P.n_per_interval); j = g - (uint64_t)s * P.n_per_interval; }
	else { const uint64_t k = (g - n_cont) / P.n_per_line; s = (int)(P.n_cont_seg + k); j = g - n_cont - k * P.n_per_line; }
	const XmbSegDev &S = P.segs[s];
	p.alive = true;
	p.n_interactions = 0;
	double hor_ver_ratio;
	if (S.is_cont) {
		// xmi_ran_trap (src/xmi_aux_f.F90:1841-1941)
		const double m = (S.y2 - S.y1) / (S.x2 - S.x1);
		const double denom = (S.x2 - S.x1) * (S.y1 - S.x1 * m) + m * (S.x2 * S.x2 - S.x1 * S.x1) / 2.0;
		const double a = m / 2.0, b = S.y1 - S.x1 * m, c = -S.x1 * S.y1 + m * S.x1 * S.x1 / 2.0 - denom * rng.uniform();
		double rv1, rv2;
		if (a == 0.0) { rv1 = -1.0 * c / b; rv2 = rv1; }
		else {
			const double delta = b * b - 4.0 * a * c;
			if (delta <= 0.0) { rv1 = -b / 2.0 / a; rv2 = rv1; }
			else { const double sq = sqrt(delta), t1 = (-b + sq) / 2.0 / a, t2 = (-b - sq) / 2.0 / a; rv1 = fmin(t1, t2); rv2 = fmax(t1, t2); }
		}
		p.energy = (S.x1 <= rv1 && rv1 <= S.x2) ? rv1 : rv2;
		const double hi = S.h1 + (S.h2 - S.h1) * (p.energy - S.x1) / (S.x2 - S.x1);
		const double ti = S.y1 + (S.y2 - S.y1) * (p.energy - S.x1) / (S.x2 - S.x1);
		hor_ver_ratio = hi / ti;
		const NodePos np = node_find(P, p.energy);
		p.weight = S.total_rel * exp(-row_lerp(P, np, P.off_exc));
		XMB_UNROLL_NL
for (int i = 0; i < nL; i++) mus[i * T] = row_lerp(P, np, i);
	} else {
		hor_ver_ratio = S.hor_ver_ratio;
		p.weight = S.weight_rel;
		if (S.distribution_type == XMB_DISCRETE_GAUSSIAN) p.energy = ran_gaussian(rng, S.scale_parameter) + S.energy;
		else if (S.distribution_type == XMB_DISCRETE_LORENTZIAN) p.energy = S.scale_parameter * tan(M_PI * rng.uniform()) + S.energy;
		else p.energy = S.energy;
		if (p.energy <= ENERGY_THRESHOLD || p.energy > ENERGY_MAX) { p.alive = false; return; }
		const NodePos np = node_find(P, p.energy);
		XMB_UNROLL_NL
for (int i = 0; i < nL; i++) mus[i * T] = row_lerp(P, np, i);
	}
	double x1, y1;
	if (fabs(S.sigma_x * S.sigma_y) < 1.0E-20) {
		x1 = P.slit_x1_max * (-1.0 + 2.0 * rng.uniform());
		y1 = P.slit_y1_max * (-1.0 + 2.0 * rng.uniform());
		p.cx = p.cy = p.cz = 0.0;
	} else {
		x1 = ran_gaussian(rng, S.sigma_xp);
		y1 = ran_gaussian(rng, S.sigma_yp);
		p.cx = ran_gaussian(rng, S.sigma_x) - P.d_source_slit * sin(x1);
		p.cy = ran_gaussian(rng, S.sigma_y) - P.d_source_slit * sin(y1);
		p.cz = 0.0;
	}
	p.dx = tan(x1); p.dy = tan(y1); p.dz = 1.0;
	normalize3(p.dx, p.dy, p.dz);
	p.theta = acos(p.dz);
	p.phi = atan2(p.dy, p.dx);
	bool horizontal;
	if (S.is_cont) horizontal = rng.uniform() <= hor_ver_ratio;
	else horizontal = (double)(j + 1) <= hor_ver_ratio;
	if (horizontal) { p.ex = 0.0; p.ey = 1.0; p.ez = 0.0; } else { p.ex = 1.0; p.ey = 0.0; p.ez = 0.0; }
	const double cosalfa = p.ex * p.dx + p.ey * p.dy + p.ez * p.dz;
	const double c_ae = 1.0 / sin(acos(cosalfa)), c_be = -c_ae * cosalfa;
	p.ex = c_ae * p.ex + c_be * p.dx; p.ey = c_ae * p.ey + c_be * p.dy; p.ez = c_ae * p.ez + c_be * p.dz;
	// xmi_photon_shift_first_layer (:1140-1186)
	p.layer = -1;
	if (p.cz >= P.layers[0].Z_begin) {
		for (int i = 0; i < nL; i++) if (p.cz < P.layers[i].Z_end) { p.layer = i; break; }
		if (p.layer < 0) { p.alive = false; return; }
	} else {
		const double ItimesN = p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2];
		if (ItimesN == 0.0) { p.alive = false; return; }
		const double d = ((0.0 - p.cx) * P.n_sample[0] + (0.0 - p.cy) * P.n_sample[1] + (P.layers[0].Z_begin - p.cz) * P.n_sample[2]) / ItimesN;
		p.cx = d * p.dx + p.cx; p.cy = d * p.dy + p.cy; p.cz = d * p.dz + p.cz;
		p.layer = 0;
	}
}

// distance along (dx,dy,dz) from (x,y,z) to the plane through (0,0,zp) with the sample normal; also moves the point
__device__ __forceinline__ bool step_to_plane(const XmbHistParams &P, double &x, double &y, double &z, double dx, double dy, double dz,
                                              double zp, double &dist) {
	const double ItimesN = dx * P.n_sample[0] + dy * P.n_sample[1] + dz * P.n_sample[2];
	if (ItimesN == 0.0) return false;
	const double d = ((0.0 - x) * P.n_sample[0] + (0.0 - y) * P.n_sample[1] + (zp - z) * P.n_sample[2]) / ItimesN;
	const double nx = d * dx + x, ny = d * dy + y, nz = d * dz + z;
	dist = sqrt((x - nx) * (x - nx) + (y - ny) * (y - ny) + (z - nz) * (z - nz));
	x = nx; y = ny; z = nz;
	return true;
}

// ---- shell-resolved ("advanced") Compton: src/xmi_aux_f.F90:1951-2073, src/xmi_main.F90:4785-4983,
//      src/xmi_variance_reduction.F90:752-947 ----------------------------------------------------------------
__device__ __forceinline__ double adv_q_from_energy(double e0, double e1, double ct) {
	const double Q = 137.0 * (e1 - e0 + (1.0 - ct) * e0 * e1 / XMI_MEC2);
	return Q / sqrt(e1 * e1 + e0 * e0 - 2.0 * e0 * e1 * ct);
}
__device__ double adv_energy_from_q(double e0, double Q, double theta) {
	const double a = e0, b = XMI_MEC2, c = cos(theta);
	if (fabs(c - 1.0) < 1E-8) return 0.0;
	if (fabs(Q) < 1E-4) return e0 / (1.0 + e0 * (1.0 - c) / XMI_MEC2);
	const double d = 1.0 + a / b - a * c / b;
	const double aq = 137.0 * 137.0 * d * d - Q * Q;
	const double bq = -2.0 * 137.0 * 137.0 * a * d + 2.0 * a * c * Q * Q;
	const double cq = 137.0 * 137.0 * a * a - a * a * Q * Q;
	double E1, E2;
	if (aq == 0.0) {                                   // xmi_poly_solve_quadratic (src/xmi_aux_f.F90:1872-1905)
		if (bq == 0.0) return 0.0;
		E1 = E2 = -1.0 * cq / bq;
	} else {
		const double delta = bq * bq - 4.0 * aq * cq;
		if (delta < 0.0) return 0.0;
		if (delta == 0.0) E1 = E2 = -bq / 2.0 / aq;
		else { const double sq = sqrt(delta), t1 = (-bq + sq) / 2.0 / aq, t2 = (-bq - sq) / 2.0 / aq; E1 = fmin(t1, t2); E2 = fmax(t1, t2); }
	}
	const double Q1 = adv_q_from_energy(e0, E1, c), Q2 = adv_q_from_energy(e0, E2, c);
	if (Q * Q1 > 0.0) return E1;
	if (Q * Q2 > 0.0) return E2;
	if (fabs(E1 - E2) < 1E-10 || fabs(Q1 - Q2) < 1E-10) return E1;
	return 0.0;
}
__device__ double adv_shell_cdf(const XmbHistParams &P, int r, double energy, double theta) {
	const double Ii = P.adv_edge[r];
	double Qimax = 0.0;
	if (!(Ii != 0.0 && energy < Ii)) {
		const double EminIi = energy - Ii, costheta = cos(theta);
		Qimax = 137.0 * (EminIi * energy * (1.0 - costheta) / XMI_MEC2 - Ii);
		Qimax = Qimax / sqrt(EminIi * EminIi + energy * energy - 2.0 * EminIi * energy * costheta);
	}
	if (Qimax < -100.0) return 0.0;
	if (Qimax > 100.0) return 1.0;
	const double *cdf = P.adv_cdf + (size_t)r * P.n_cp;
	const double dq = 100.0 / (P.n_cp - 1.0), qa = fabs(Qimax);
	const int pos = min((int)(qa / dq), P.n_cp - 2);
	const double v = cdf[pos] + (cdf[pos + 1] - cdf[pos]) * (qa - dq * pos) / dq;
	return Qimax < 0.0 ? 1.0 - (0.5 + v) : 0.5 + v;
}
__device__ double adv_sample_q(const XmbHistParams &P, int r, double cdf) {
	const double *qinv = P.adv_qinv + (size_t)r * P.n_cp;
	const double dc = 0.5 / (P.n_cp - 1.0), cp = cdf < 0.5 ? 0.5 - cdf : cdf - 0.5;
	const int pos = min((int)(cp / dc), P.n_cp - 2);
	const double q = qinv[pos] + (qinv[pos + 1] - qinv[pos]) * (cp - dc * pos) / dc;
	return cdf < 0.5 ? -q : q;
}
// xmi_update_photon_energy_compton (:4785-4983): two draws {subshell, Q}
__device__ double compton_energy_adv(const XmbHistParams &P, int zi, double E0, double theta_i, double u_shell, double u_q) {
	const int r0 = P.adv_off[zi], r1 = P.adv_off[zi + 1];
	double cdf_sum = 0.0;
	for (int r = r0; r < r1; r++) cdf_sum += P.adv_config[r] * adv_shell_cdf(P, r, E0, theta_i);
	if (cdf_sum == 0.0) return 0.0;
	double temp_sum = 0.0, cdf_i = 0.0;
	int i = r1 - 1;
	for (int r = r0; r < r1; r++) {
		cdf_i = adv_shell_cdf(P, r, E0, theta_i);
		temp_sum += P.adv_config[r] * cdf_i / cdf_sum;
		if (u_shell <= temp_sum) { i = r; break; }
	}
	return adv_energy_from_q(E0, adv_sample_q(P, i, u_q * cdf_i), theta_i);
}

// exp(-t) for t >= 0 in ~15 instructions (the library exp is ~30 and the line loop evaluates one per active line and
// interaction): 2^(-y) with y = t log2(e) = (j + r) / 64, |r| <= 1/2; 2^(-j/64) = 2^(-(j >> 6)) tab[j & 63] with a 64-entry
// table in shared memory, and exp(-r ln2 / 64) by its Taylor polynomial of degree 5 (|x| < 0.0055: remainder 4e-17).
// Relative error <= 2^-53 t + 3e-16, i.e. 1e-13 at the largest exponents that still matter.
__device__ __forceinline__ double exp_neg(double t, const double *tab) {
	const double y = t * (64.0 * 1.4426950408889634074);
	if (!(y < 64.0 * 1000.0)) return 0.0;                       // exp(-693) = 1e-301: below anything a deposit can represent
	const double jf = rint(y);
	const int j = (int)jf;
	const double x = (jf - y) * (0.69314718055994530942 / 64.0);
	double pl = fma(x, 1.0 / 120.0, 1.0 / 24.0);
	pl = fma(pl, x, 1.0 / 6.0);
	pl = fma(pl, x, 0.5);
	pl = fma(pl, x, 1.0);
	pl = fma(pl, x, 1.0);
	const double scale = __hiloint2double((1023 - (j >> 6)) << 20, 0);   // 2^(-(j >> 6)), j >> 6 <= 1000
	return pl * tab[j & 63] * scale;
}

// ---- atom and interaction selection, scattering (src/xmi_main.F90:1558-1652) ------------------------------
// MODE 0: forced detection (the fluorescence yield multiplies the weight); 1: escape-ratio mode (it goes to
// weight_escape, src/xmi_variance_reduction.F90:697-750); 2: brute force (analogue yield check, :2297-2319 / :5325-5350:
// on failure the energy is zeroed and out_type = 4 tells the caller to run the Auger cascade on out_shell).
// out_type: 1 Rayleigh, 2 Compton, 3 photo-electric (4: Auger); out_zi: element slot; out_line: |line macro| or 0;
// out_shell: the ionised shell, after Coster-Kronig when a line was emitted.
template <int NL, int MODE, bool ADV = false>
__device__ __forceinline__ void select_and_scatter(const XmbHistParams &P, Photon &p, uint64_t g, int order, double *mus, int T,
                                                   uint32_t atom_word, double &weight_escape, int &out_type, int &out_zi, int &out_line,
                                                   int &out_shell) {
	const int nL = NL > 0 ? NL : P.nL;
	out_line = 0;
	out_shell = -1;
	const XmbLayerDev lay = P.layers[p.layer];
	const NodePos ep = node_find(P, p.energy);
	const uint4 b1 = draw_block(P.seed, g, order, 1, 0, 1);   // {interaction type, s0, s1, s2}
	double R2 = xmb_u01(atom_word);
	double thr = 0.0;
	int zi = 0;
	const double mu_cur = mus[p.layer * T];
	for (int i = 0; i < lay.n_elements; i++) {
		zi = P.elem_zi[lay.elem_begin + i];
		thr += P.elem_w[lay.elem_begin + i] * row_lerp(P, ep, P.off_elem + zi * XMB_ELEM_STRIDE + XMB_EO_CS_TOTAL) / mu_cur;
		if (R2 < thr) break;
	}
	const int eoff = P.off_elem + zi * XMB_ELEM_STRIDE;
	R2 = xmb_u01(b1.x);
	const double s0 = xmb_u01(b1.y), s1 = xmb_u01(b1.z), s2 = xmb_u01(b1.w);
	const double pr = row_lerp(P, ep, eoff + XMB_EO_P_RAYL), prc = row_lerp(P, ep, eoff + XMB_EO_P_RAYL_COMPT);
	out_zi = zi;
	// The three interaction branches only decide (theta_i, phi_rot, new energy); the lookups they share and the
	// rotation of the direction / polarisation vectors run once, after the branches, with the warp converged
	// (profiles/r1_history_kernel_v8_*: the rotation code ran at 10 of 32 lanes when it was inlined per branch).
	const bool is_rayl = R2 < pr, is_compt = !is_rayl && R2 < prc;
	double theta_i = 0.0, phi_i = 0.0, phi_rot = 0.0;
	bool rotate = false, new_energy = false;
	if (is_rayl || is_compt) {
		out_type = is_rayl ? 1 : 2;
		// Rayleigh (:1986-2101) / Compton (:2103-2229): theta from the element's inverse CDF, phi from the polarisation table
		const double *icdf = (is_rayl ? P.rayl_icdf : P.compt_icdf) + (size_t)zi * P.n_icdf_E * P.n_icdf_R;
		theta_i = bilinear(icdf, P.n_icdf_R, P.icdf_E, P.n_icdf_E, P.icdf_R, p.energy, s0);
		double sti, cti;
		sincos(theta_i, &sti, &cti);
		double tt = sti * sti;
		if (is_rayl) tt = tt / (4.0 - 2.0 * tt);
		else {
			const double K0K = 1.0 + p.energy * (1.0 - cti) / XMI_MEC2;
			tt = tt / (K0K + (1.0 / K0K) - tt) / 2.0;
		}
		phi_i = bilinear(P.phi_icdf, P.n_icdf_R, P.phi_T, P.n_phi_T, P.icdf_R, tt, s1);
		phi_rot = phi_i + elec_phi0(p);
		rotate = true;
		if (is_compt) {
			if (ADV) {
				const uint4 w = draw_block(P.seed, g, order, 3, 0, 0);
				p.energy = compton_energy_adv(P, zi, p.energy, theta_i, xmb_u01(w.x), xmb_u01(w.y));
			} else
				p.energy = compton_energy(P, zi, p.energy, 1.2399E-6 / (p.energy * 1000.0), sin(theta_i / 2.0), g, order, 3, 0, false);
			new_energy = true;
			rotate = p.energy != 0.0;
		}
	} else {
		out_type = 3;
		// photo-electric effect with fluorescence (:2231-2411)
		const double photo_total = row_lerp(P, ep, eoff + XMB_EO_PHOTO_TOTAL);
		double sumz = 0.0;
		const double r = s0;
		const int max_shell = P.use_M_lines ? 8 : 3;
		int shell = -1;
		for (int s = 0; s <= max_shell; s++) {
			sumz += row_lerp(P, ep, eoff + XMB_EO_PHOTO_PARTIAL + s) / photo_total;
			if (r < sumz) { shell = s; break; }
		}
		if (shell < 0) { p.energy = 0.0; }
		else {
			// (the reference draws one unused number here, xmi_variance_reduction.F90:737; not reproduced)
			if (MODE == 1) weight_escape *= P.fluor_yield_corr[zi * 9 + shell];
			else if (MODE == 0) p.weight *= P.fluor_yield_corr[zi * 9 + shell];
			SubStream xs;
			xs.init(P.seed, g, order, 3, 0);
			const double u_phi = xs.uniform();
			out_shell = shell;
			if (MODE == 2 && xs.uniform() > P.fluor_yield_corr[zi * 9 + shell]) { p.energy = 0.0; out_type = 4; return; }
			// Coster-Kronig (:5184-5323)
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
			// line (:5352-5437)
			const double rl = s1;
			double sl = 0.0;
			int line = 0;
			const int lf = d_shell_line_first[shell], ll = d_shell_line_last[shell];
			for (int l = lf; l <= ll; l++) { sl += P.rad_rate[(size_t)zi * 384 + l]; if (rl < sl) { line = l; break; } }
			if (!line) p.energy = 0.0;
			else {
				out_line = line;
				out_shell = shell;
				p.energy = P.line_energy[(size_t)zi * 384 + line];
				new_energy = true;
				theta_i = acos(-2.0 * s2 + 1.0);
				phi_rot = 2.0 * M_PI * u_phi;
				rotate = true;
			}
		}
	}
	// ---- common tail: attenuation coefficients at the new energy, rotation of direction and polarisation ----------
	if (new_energy) {
		const NodePos cp = node_find(P, p.energy);
		XMB_UNROLL_NL
for (int i = 0; i < nL; i++) mus[i * T] = row_lerp(P, cp, i);
	}
	if (rotate) {
		update_dirv(p, theta_i, phi_rot);
		update_elecv(p);
		if (is_compt) {
			// depolarisation of the Compton-scattered photon (:2201-2211)
			double spi, cpi;
			sincos(phi_i, &spi, &cpi);
			const double cti = cos(theta_i);
			double pp = 2.0 * ((cti * cpi) * (cti * cpi) + spi * spi);
			const double rat = 1.0 / (1.0 + (1 - cti) * p.energy / 510.998910);
			const double rk = rat - 2.0 + 1.0 / rat;
			pp = pp / (rk + pp);
			const double w_h = (1.0 + pp) / 2.0;
			if (s2 > w_h) {
				const double tx = p.dy * p.ez - p.dz * p.ey, ty = p.dz * p.ex - p.dx * p.ez, tz = p.dx * p.ey - p.dy * p.ex;
				p.ex = tx; p.ey = ty; p.ez = tz;
			}
		}
	}
}

// NL > 0: number of layers known at compile time (loops over layers fully unrolled); NL = 0: generic.
template <int NL, bool ADV = false>
__global__ void __launch_bounds__(HIST_THREADS, HIST_MIN_BLOCKS) xmb_history_kernel(const __grid_constant__ XmbHistParams P) {
	const int nL = NL > 0 ? NL : P.nL;
	extern __shared__ double smem[];
	const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
	double *mus = smem + tid;                 // mus[j*T]   : mu of layer j at the photon energy
	double *rd = smem + (size_t)nL * T + tid;   // rd[j*T]    : distances, then rho_j * d_j towards the detector
	unsigned int *stage = reinterpret_cast<unsigned int *>(smem + (size_t)2 * nL * T);   // [nch + n_hist_slots][4] 16-bit pieces
	const uint64_t n_total = P.n_local_span;
	const uint64_t n_chunks = (n_total + T - 1) / T;
	const size_t acc_row = (size_t)P.nch + P.n_hist_slots;
	unsigned long long n_inter_local = 0;
	bool bad_fixed = false;
	__shared__ unsigned int s_layer_cnt[XMB_MAX_LAYERS];
	__shared__ double s_exp_tab[64];              // 2^(-k/64), see exp_neg()
	if (tid < XMB_MAX_LAYERS) s_layer_cnt[tid] = 0;
	if (tid < 64) s_exp_tab[tid] = exp2(-(double)tid / 64.0);
	__syncthreads();

	// The kernel is ~260 KB of SASS (fp64 transcendentals inlined at every site) against a 32 KB L1.5 I-cache: a
	// CTA therefore walks the phases of an interaction in lock step (__syncthreads between phases), so all warps
	// of the SM fetch the same few KB at any time (profiles/: stall_no_instruction 7.3 -> see r1 v3).
	//
	// CTA-local wavefront with compaction: survivors of order k are appended, densely, to the CTA's queue for order
	// k+1 (structure-of-arrays in global memory, coalesced); the CTA always runs the deepest order that has a full
	// batch of T photons, otherwise samples a fresh chunk of source photons, and drains the queues at the end.
	// Every batch therefore has all lanes alive and one interaction order; exact integer deposits and fixed-address
	// random numbers make the result independent of this regrouping.
	__shared__ int s_qcount[XMB_MAX_ORDERS];      // photons waiting to run order k+1
	__shared__ int s_wsum[32];
	if (tid < XMB_MAX_ORDERS) s_qcount[tid] = 0;
	for (int i = tid; i < 4 * (P.nch + P.n_hist_slots); i += T) stage[i] = 0u;
	__syncthreads();
	const int NF = XMB_STATE_FIELDS + nL;
	const size_t qcap = 2 * (size_t)T;
	double *qbase = P.queue + (size_t)blockIdx.x * P.n_int * NF * qcap;
	uint64_t next_chunk = blockIdx.x;
	for (;;) {
		// ---- scheduler (block-uniform) ---------------------------------------------------------------
		int k = -1;
		for (int kk = P.n_int - 1; kk >= 1; kk--) if (s_qcount[kk] >= T) { k = kk; break; }
		bool from_source = false;
		if (k < 0) {
			if (next_chunk < n_chunks) from_source = true;
			else {
				for (int kk = P.n_int - 1; kk >= 1; kk--) if (s_qcount[kk] > 0) { k = kk; break; }
				if (k < 0) break;
			}
		}
		uint64_t g = 0;
		Photon p;
		p.alive = false;
		p.layer = 0; p.n_interactions = 0; p.energy = 0.0; p.weight = 0.0;
		int order = 1;
		if (from_source) {
			const uint64_t lid = next_chunk * T + tid;
			g = shard_global_id(P, lid);
			next_chunk += gridDim.x;
			p.alive = lid < P.n_local_span && g < P.n_total;
			if (p.alive) {
				XmbRng rng;   // order 0, stage 0: sequential words, counter word 2 = block
				rng.init(P.seed, g, XMB_TAG_HISTORY);
				start_photon<NL>(P, p, rng, g, mus, T);
			}
		} else {
			const int have = s_qcount[k], n = min(T, have), base = have - n;
			order = k + 1;
			if (tid < n) {
				const double *q = qbase + (size_t)k * NF * qcap + base + tid;
				p.cx = q[0 * qcap]; p.cy = q[1 * qcap]; p.cz = q[2 * qcap];
				p.dx = q[3 * qcap]; p.dy = q[4 * qcap]; p.dz = q[5 * qcap];
				p.ex = q[6 * qcap]; p.ey = q[7 * qcap]; p.ez = q[8 * qcap];
				p.energy = q[9 * qcap]; p.weight = q[10 * qcap]; p.theta = q[11 * qcap]; p.phi = q[12 * qcap];
				g = (uint64_t)__double_as_longlong(q[13 * qcap]);
				p.layer = (int)__double_as_longlong(q[14 * qcap]);
				XMB_UNROLL_NL
for (int j = 0; j < nL; j++) mus[j * T] = q[(XMB_STATE_FIELDS + j) * qcap];
				p.n_interactions = order - 1;
				p.alive = true;
			}
			__syncthreads();
			if (tid == 0) s_qcount[k] = base;
		}
		{
			// ---- forced interaction (src/xmi_main.F90:1229-1518) ------------------------------------
			if (p.alive && p.energy < ENERGY_THRESHOLD) p.alive = false;
			const uint4 b0 = draw_block(P.seed, g, order, 1, 0, 0);   // {path length, detector r, detector phi, atom}
			double interactionR = 0.0;
			int step_max = 0, step_dir = 1;
			if (p.alive) {
				if (p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0) { step_max = nL - 1; step_dir = 1; }
				else { step_max = 0; step_dir = -1; }
				interactionR = xmb_u01(b0.x);
				double lx = p.cx, ly = p.cy, lz = p.cz;
				double Pabs = 0.0;
				for (int i = p.layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {
					double dist;
					if (!step_to_plane(P, lx, ly, lz, p.dx, p.dy, p.dz, step_dir == 1 ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { p.alive = false; break; }
					rd[i * T] = dist;
					Pabs += mus[i * T] * P.layers[i].density * dist;
				}
				if (p.alive) {
					const double Pabs2 = -1.0 * expm1(-1.0 * Pabs);
					p.weight *= Pabs2;
					const double l1p = log1p(-1.0 * interactionR * Pabs2);
					const double negln = -1.0 * l1p;
					int my_index = p.layer;
					double my_sum = 0.0;
					for (int i = p.layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {
						my_sum += mus[i * T] * P.layers[i].density * rd[i * T];
						if (my_sum > negln) { my_index = i; break; }
					}
					const double murho_idx = mus[my_index * T] * P.layers[my_index].density;
					double temp_sum = 0.0;
					for (int i = p.layer; step_dir > 0 ? i <= my_index : i >= my_index; i += step_dir)
						temp_sum += (1.0 - (mus[i * T] * P.layers[i].density / murho_idx)) * rd[i * T];
					temp_sum = temp_sum - 1.0 * l1p / murho_idx;
					p.cx += temp_sum * p.dx; p.cy += temp_sum * p.dy; p.cz += temp_sum * p.dz;
					p.layer = my_index;
					p.n_interactions++;
					n_inter_local++;
#ifndef XMB_NO_LAYER_CNT
					atomicAdd(&s_layer_cnt[my_index], 1u);
#endif
				}
			}
			__syncthreads();   // phase barrier after transport
			const int n_ia = order;   // == p.n_interactions for every live lane
			unsigned int *acc_k = stage;   // deposits of this batch are staged in shared memory, flushed below

			// ---- forced detection (src/xmi_variance_reduction.F90:29-726) -----------------------------
			bool vr = p.alive && p.energy > ENERGY_THRESHOLD;
			double theta = 0.0, phi = 0.0, Pesc_rayl = 0.0, omega = 0.0;
			NodePos np;
			np.pos = 0; np.f = 0.0;
			if (vr) {
				const double radius = sqrt(xmb_u01(b0.y)) * P.detector_radius;
				const double th = 2.0 * M_PI * xmb_u01(b0.z);
				double sdp, cdp;
				sincos(th, &sdp, &cdp);
				const double dp0 = 0.0, dp1 = cdp * radius, dp2 = sdp * radius;
				const double rx = p.cx - P.p_window[0], ry = p.cy - P.p_window[1], rz = p.cz - P.p_window[2];
				const double *B = P.ndo_inv, *A = P.ndo_new;
				const double lp0 = B[0] * rx + B[1] * ry + B[2] * rz, lp1 = B[3] * rx + B[4] * ry + B[5] * rz, lp2 = B[6] * rx + B[7] * ry + B[8] * rz;
				double d0 = B[0] * p.dx + B[1] * p.dy + B[2] * p.dz, d1 = B[3] * p.dx + B[4] * p.dy + B[5] * p.dz, d2 = B[6] * p.dx + B[7] * p.dy + B[8] * p.dz;
				double l0 = dp0 - lp0, l1 = dp1 - lp1, l2 = dp2 - lp2;
				if (l0 >= 0.0) vr = false;
				else {
					double total_distance = sqrt((dp0 - lp0) * (dp0 - lp0) + (dp1 - lp1) * (dp1 - lp1) + (dp2 - lp2) * (dp2 - lp2));
					normalize3(d0, d1, d2);
					normalize3(l0, l1, l2);
					const double n0 = A[0] * l0 + A[1] * l1 + A[2] * l2, n1 = A[3] * l0 + A[4] * l1 + A[5] * l2, n2 = A[6] * l0 + A[7] * l1 + A[8] * l2;
					double dotprod = d0 * l0 + d1 * l1 + d2 * l2;
					dotprod = fmin(1.0, fmax(-1.0, dotprod));
					theta = acos(dotprod);
					const double dpp = n0 * p.dx + n1 * p.dy + n2 * p.dz;
					double q0 = n0 - dpp * p.dx, q1 = n1 - dpp * p.dy, q2 = n2 - dpp * p.dz;
					normalize3(q0, q1, q2);
					const double en = sqrt(p.ex * p.ex + p.ey * p.ey + p.ez * p.ez);
					dotprod = q0 * (p.ex / en) + q1 * (p.ey / en) + q2 * (p.ez / en);
					dotprod = fmin(1.0, fmax(-1.0, dotprod));
					phi = acos(dotprod);
					int vmax, vdir;
					if (n0 * P.n_sample[0] + n1 * P.n_sample[1] + n2 * P.n_sample[2] > 0.0) { vmax = nL - 1; vdir = 1; } else { vmax = 0; vdir = -1; }
					XMB_UNROLL_NL
for (int i = 0; i < nL; i++) rd[i * T] = 0.0;
					double tx = p.cx, ty = p.cy, tz = p.cz;
					double temp_murhod = 0.0;
					for (int i = p.layer; vdir > 0 ? i <= vmax : i >= vmax; i += vdir) {
						double dist;
						if (!step_to_plane(P, tx, ty, tz, n0, n1, n2, vdir == 1 ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { vr = false; break; }
						bool last = false;
						if (dist > total_distance) { dist = total_distance; last = true; }
						rd[i * T] = P.layers[i].density * dist;
						temp_murhod += mus[i * T] * P.layers[i].density * dist;
						if (last) break;
						total_distance -= dist;
					}
					Pesc_rayl = exp(-temp_murhod);
					omega = get_solid_angle(P, p);
					np = node_find(P, p.energy);
				}
			}
			__syncthreads();   // phase: scatter deposits of every element
			// warp-uniform loops over layers / elements / shells / line records
			for (int L = 0; L < nL; L++) {
				const bool mine = vr && p.layer == L;
				if (!__any_sync(0xffffffffu, mine)) continue;
				const XmbLayerDev lay = P.layers[L];
				const double inv_mu = mine ? 1.0 / mus[L * T] : 0.0;
				double qf = 0.0, sin2cos2 = 0.0, k0k = 1.0, c_lamb0 = 0.0, sth2 = 0.0, dcsp_kn = 0.0;
				int qi = 0;
				long ch_rayl = -1;
				if (mine) {
					sth2 = sin(theta / 2.0);
					c_lamb0 = 1.2399E-6 / (p.energy * 1000.0);
					const double q = p.energy / KEV2ANGST * sth2;
					const double qx = q / P.q_max * (P.n_q - 1);
					qi = min((int)qx, P.n_q - 2);
					qf = qx - qi;
					double st, ct, cp = cos(phi);
					sincos(theta, &st, &ct);
					sin2cos2 = st * st * cp * cp;
					k0k = 1.0 / (1.0 + (1.0 - ct) * p.energy / 510.998928);
					dcsp_kn = RE2 / 2.0 * k0k * k0k * (k0k + 1.0 / k0k - 2.0 * sin2cos2);
					const int ch = (int)((p.energy - P.zero) / P.gain);
					if (p.energy >= ENERGY_THRESHOLD && ch >= 0 && ch <= P.nch - 1) ch_rayl = ch;
				}

				for (int e = 0; e < lay.n_elements; e++) {
					const int zi = P.elem_zi[lay.elem_begin + e];
					const double wfrac = P.elem_w[lay.elem_begin + e];
					const size_t hbase = (size_t)P.nch + P.hist_base[zi];
					// the element's random block, first inverse-CDF bracket and form factors are requested together, ahead of
					// the dependent chain Compton energy -> energy bracket -> mu rows -> exp
					ComptonPrefetch pf;
					if (mine && !ADV) compton_prefetch(P, zi, g, order, e, qi, pf);
					if (mine && ADV) { const double *f = P.ff + (size_t)zi * P.n_q + qi, *sfp = P.sf + (size_t)zi * P.n_q + qi; pf.F0 = f[0]; pf.F1 = f[1]; pf.S0 = sfp[0]; pf.S1 = sfp[1]; }
					// Rayleigh (:342-369)
					unsigned long long fx = 0ULL;
					double Pconv = 0.0;
					if (mine) {
						Pconv = wfrac * inv_mu;
						const double F = pf.F0 * (1.0 - qf) + pf.F1 * qf;
						const double dcsp = P.avog_over_A[zi] * F * F * RE2 * (1.0 - sin2cos2);
						fx = to_fixed(Pconv * (omega * dcsp) * Pesc_rayl * p.weight, P.counters);
					}
					deposit_uniform(acc_k, hbase + 0, fx, lane);
					deposit_varying(acc_k, ch_rayl, fx, lane);
					if (ADV) {
						// shell-resolved Compton (xmi_compton_varred, :752-947): one deposit per occupied subshell
						const int r0 = P.adv_off[zi], r1 = P.adv_off[zi + 1];
						double cdf_sum = 0.0, Pdir = 0.0;
						if (mine) {
							for (int r = r0; r < r1; r++) cdf_sum += P.adv_config[r] * adv_shell_cdf(P, r, p.energy, theta);
							const double S = pf.S0 * (1.0 - qf) + pf.S1 * qf;
							Pdir = omega * P.avog_over_A[zi] * S * dcsp_kn;
						}
						for (int r = r0; r < r1; r++) {
							fx = 0ULL;
							long ch_c = -1;
							if (mine && cdf_sum != 0.0) {
								const double cdf_r = adv_shell_cdf(P, r, p.energy, theta);
								const double shell_weight = P.adv_config[r] * cdf_r / cdf_sum;
								if (shell_weight != 0.0) {
									const uint4 w = draw_block(P.seed, g, order, 2, e, (r - r0) >> 2);
									const int k = (r - r0) & 3;
									const double u = xmb_u01(k == 0 ? w.x : k == 1 ? w.y : k == 2 ? w.z : w.w);
									const double e_c = adv_energy_from_q(p.energy, adv_sample_q(P, r, u * cdf_r), theta);
									if (e_c != 0.0) {
										const NodePos cp = node_find(P, e_c);
										double tm = 0.0;
										for (int j = 0; j < nL; j++) tm += row_lerp(P, cp, j) * rd[j * T];
										fx = to_fixed(Pconv * Pdir * exp(-tm) * p.weight * shell_weight, P.counters);
										const int ch = (int)((e_c - P.zero) / P.gain);
										if (e_c >= ENERGY_THRESHOLD && ch >= 0 && ch <= P.nch - 1) ch_c = ch;
									}
								}
							}
							deposit_uniform(acc_k, hbase + 1, fx, lane);
							deposit_varying(acc_k, ch_c, fx, lane);
						}
						continue;
					}
					// Compton (xmi_compton_varred2, :949-1008)
					fx = 0ULL;
					long ch_c = -1;
					if (mine) {
						const double e_c = compton_energy(P, zi, p.energy, c_lamb0, sth2, g, order, 2, e, true, &pf);
						const NodePos cp = node_find(P, e_c);
						double tm = 0.0;
						XMB_UNROLL_NL
for (int j = 0; j < nL; j++) tm += row_lerp(P, cp, j) * rd[j * T];
						const double S = pf.S0 * (1.0 - qf) + pf.S1 * qf;
						const double Pdir = omega * P.avog_over_A[zi] * S * dcsp_kn;
						fx = to_fixed(Pconv * Pdir * exp_neg(tm, s_exp_tab) * p.weight, P.counters);
						const int ch = (int)((e_c - P.zero) / P.gain);
						if (e_c >= ENERGY_THRESHOLD && ch >= 0 && ch <= P.nch - 1) ch_c = ch;
					}
					deposit_uniform(acc_k, hbase + 1, fx, lane);
					deposit_varying(acc_k, ch_c, fx, lane);
				}
			}
			__syncthreads();   // phase: fluorescence-line deposits (small loop body: exp + exact warp sum + RED)
			for (int L = 0; L < nL; L++) {
				const bool mine = vr && p.layer == L;
				if (!__any_sync(0xffffffffu, mine)) continue;
				const XmbLayerDev lay = P.layers[L];
				const double inv_mu = mine ? 1.0 / mus[L * T] : 0.0;
				for (int e = 0; e < lay.n_elements; e++) {
					const int zi = P.elem_zi[lay.elem_begin + e];
					const double wfrac = P.elem_w[lay.elem_begin + e];
					// fluorescence lines (:391-709): per shell, vacancy cross section at the photon energy (after a
					// fluorescence interaction that energy is a node of the grid: the reference's precalc_xrf_cs)
					const double common = mine ? wfrac * inv_mu * (omega / 4.0 / M_PI) * p.weight : 0.0;
					const bool aboveK = mine && p.energy >= P.edge_K[zi];
					const int eoff = P.off_elem + zi * XMB_ELEM_STRIDE;
					const int n_sh = P.use_M_lines ? 9 : 4;
					for (int s = 0; s < n_sh; s++) {
						const int r0 = P.rec_begin[zi * 10 + s], r1 = P.rec_begin[zi * 10 + s + 1];
						if (r0 == r1) continue;
						double Ps = 0.0;
						if (mine && (s > 0 || aboveK)) Ps = row_lerp(P, np, eoff + XMB_EO_VACANCY + s);
						if (!__any_sync(0xffffffffu, Ps != 0.0)) continue;
						const double pre = common * Ps;
XMB_UNROLL(XMB_REC_UNROLL)
						for (int r = r0; r < r1; r++) {
							const double *mu = P.rec_mu + (size_t)r * nL;
							double tm = 0.0;
							XMB_UNROLL_NL
for (int j = 0; j < nL; j++) tm += mu[j] * rd[j * T];
							const double tw = pre * P.rec_yr[r] * exp_neg(tm, s_exp_tab);
							deposit_uniform(acc_k, (size_t)P.nch + P.rec_slot[r], mine ? to_fixed_fast(tw, bad_fixed) : 0ULL, lane);
						}
					}
				}
			}

			// ---- atom and interaction selection, scattering (src/xmi_main.F90:1558-1652) ----------------
			__syncthreads();   // phase: selection + scattering (and: every deposit of the batch is staged)
			flush_staged(stage, P.acc + 2 * (size_t)(n_ia - 1) * acc_row, (int)acc_row, tid, T);
			if (p.alive) {
				double we_unused = 0.0;
				int t_unused, z_unused, l_unused, s_unused;
				select_and_scatter<NL, 0, ADV>(P, p, g, order, mus, T, b0.w, we_unused, t_unused, z_unused, l_unused, s_unused);
			}
		}
		// ---- compaction: survivors go, densely packed, to the queue of the next order -------------------
		if (order < P.n_int) {
			const bool surv = p.alive && p.energy >= ENERGY_THRESHOLD;
			const unsigned bal = __ballot_sync(0xffffffffu, surv);
			if (lane == 0) s_wsum[tid >> 5] = __popc(bal);
			__syncthreads();
			int off = 0, tot = 0;
			for (int w = 0; w < (T >> 5); w++) { const int c = s_wsum[w]; if (w < (tid >> 5)) off += c; tot += c; }
			const int have = s_qcount[order];
			if (surv) {
				double *q = qbase + (size_t)order * NF * qcap + have + off + __popc(bal & ((1u << lane) - 1u));
				q[0 * qcap] = p.cx; q[1 * qcap] = p.cy; q[2 * qcap] = p.cz;
				q[3 * qcap] = p.dx; q[4 * qcap] = p.dy; q[5 * qcap] = p.dz;
				q[6 * qcap] = p.ex; q[7 * qcap] = p.ey; q[8 * qcap] = p.ez;
				q[9 * qcap] = p.energy; q[10 * qcap] = p.weight; q[11 * qcap] = p.theta; q[12 * qcap] = p.phi;
				q[13 * qcap] = __longlong_as_double((long long)g);
				q[14 * qcap] = __longlong_as_double((long long)p.layer);
				XMB_UNROLL_NL
for (int j = 0; j < nL; j++) q[(XMB_STATE_FIELDS + j) * qcap] = mus[j * T];
			}
			__syncthreads();
			if (tid == 0) s_qcount[order] = have + tot;
		}
		__syncthreads();
	}
	n_inter_local = warp_sum_u64(n_inter_local);
	if (lane == 0 && n_inter_local) atomicAdd(&P.counters[1], n_inter_local);
	if (bad_fixed) atomicAdd(&P.counters[2], 1ULL);
	__syncthreads();
	if (tid < nL && s_layer_cnt[tid]) atomicAdd(&P.counters[8 + tid], (unsigned long long)s_layer_cnt[tid]);
}

// raw (lo, hi) accumulators -> two 48-bit-split words per slot (safe to sum over ranks in 64-bit integers)
__global__ void xmb_limbs_kernel(const unsigned long long *__restrict__ acc, unsigned long long *__restrict__ limbs, size_t n_slots) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long lo = acc[2 * i], hi = acc[2 * i + 1];   // 128-bit total
		limbs[2 * i] = lo & 0xFFFFFFFFFFFFULL;
		limbs[2 * i + 1] = (lo >> 48) | (hi << 16);
	}
}

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

struct XmbBruteParams {
	int use_auger, use_rad;
	double collimator_height, collimator_radius, half_apex, vertex_x, vertex_y, vertex_z;
	int collimator_present;
	const int *line_slot;        // [nZ][384]: compact history slot of a line, -1 = not an active line
	const double *auger_rate;    // [nZ][XMB_N_AUGER] running sums within each block (K: 240, L1..L3: 135 each)
};

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
	for (int i = 0; i < nL; i++) mus[i] = row_lerp(P, lp, i);
	q.theta = acos(2.0 * xs.uniform() - 1.0);
	q.phi = 2.0 * M_PI * xs.uniform();
	q.dx = sin(q.theta) * cos(q.phi); q.dy = sin(q.theta) * sin(q.phi); q.dz = cos(q.theta);
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
				int step_max, step_dir;
				if (p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0) { step_max = nL - 1; step_dir = 1; }
				else { step_max = 0; step_dir = -1; }
				order = (p.n_interactions + 1) | gen_bit;
				b0 = draw_block(P.seed, g, order, 1, 0, 0);
				const double interactionR = xmb_u01(b0.x);
				double blbs = 1.0, max_random_layer = 0.0;
				bool stop = false;
				for (int i = p.layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {
					double nx = p.cx, ny = p.cy, nz = p.cz, dist;
					if (!step_to_plane(P, nx, ny, nz, p.dx, p.dy, p.dz, step_dir == 1 ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { stop = true; break; }
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
		if (have) {
			p.n_interactions++;
			n_inter++;
			double we_unused = 0.0;
			int shell = -1;
			select_and_scatter<NL, 2, ADV>(P, p, g, order, mus, 1, b0.w, we_unused, last_type, last_zi, last_line, shell);
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

// =====================================================================================================
// Host side: device layouts, launch, exact reduction epilogue.
// =====================================================================================================
struct XmbDeviceTables {
	int cascade = 0, use_M_lines = -1, device = -1;
	std::vector<void *> allocs;
	XmbHistParams P{};
	// host metadata for the epilogue
	std::vector<int> rec_slot, rec_channel, rec_line, rec_zi, hist_base;
	int n_rec = 0, n_hist_slots = 0, max_nE = 1;
	double W_max = 0.0;
	uint64_t n_total = 0;
	// solid-angle grid + accumulators (re-used across calls)
	double *sa_grid = nullptr, *sa_r = nullptr, *sa_t = nullptr;
	size_t sa_cap = 0, sa_n = 0;
	const double *sa_host = nullptr;
	unsigned long long *acc = nullptr, *limbs = nullptr, *counters = nullptr;
	size_t acc_slots = 0;
	double *queue = nullptr;
	size_t queue_doubles = 0;
	int *line_slot = nullptr;          // [nZ][384] compact history slot of a line (brute-force scoring)
	double *auger_rate = nullptr;      // [nZ][XMB_N_AUGER]
	unsigned long long brute_counters[8] = {0};
	unsigned long long layer_interactions[XMB_MAX_LAYERS] = {0};
	~XmbDeviceTables() {
		for (void *p : allocs) cudaFree(p);
		cudaFree(sa_grid); cudaFree(sa_r); cudaFree(sa_t); cudaFree(acc); cudaFree(limbs); cudaFree(counters); cudaFree(queue);
	}
};

void xmb_free_device_tables(XmbDeviceTables *dev) { delete dev; }

static bool g_layout_only = false;   // build_device_tables: host-side layout without touching CUDA

template <typename T>
static T *upload(XmbDeviceTables *D, const T *src, size_t n, bool &ok) {
	T *d = nullptr;
	if (g_layout_only) return nullptr;
	if (n == 0) n = 1;
	if (cudaMalloc(&d, sizeof(T) * n) != cudaSuccess) { ok = false; return nullptr; }
	D->allocs.push_back(d);
	if (src && cudaMemcpy(d, src, sizeof(T) * n, cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
	return d;
}

// VR line -> shell classification (src/xmi_variance_reduction.F90:587-654)
static int vr_shell_of_line(int l) {
	if (l >= 1 && l <= 29) return 0;
	if (l >= XMB_L1M1 && l <= 58) return 1;
	if (l >= XMB_L2M1 && l <= 85) return 2;
	if (l >= 86 && l <= 113) return 3;
	if (l >= 118 && l <= 136) return 4;
	if (l >= 140 && l <= 158) return 5;
	if (l >= 161 && l <= 180) return 6;
	if (l >= 182 && l <= 200) return 7;
	if (l >= 201 && l <= 219) return 8;
	return -1;
}

static int cascade_mode(const xmb_main_options *o) {   // src/xmi_main.F90:141-153
	return 1 + (o->use_cascade_auger ? 1 : 0) + (o->use_cascade_radiative ? 2 : 0);
}

static XmbDeviceTables *build_device_tables(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *opt) {
	const xmb_tables_host &T = h->view;
	const xmb_input &I = in->in;
	const int nL = I.composition->n_layers, nZ = T.nZ, nN = T.n_nodes;
	if (nL > XMB_MAX_LAYERS) { xmb_set_error("more than %d layers", XMB_MAX_LAYERS); return nullptr; }
	XmbDeviceTables *D = new XmbDeviceTables();
	D->cascade = cascade_mode(opt);
	D->use_M_lines = opt->use_M_lines ? 1 : 0;
	if (g_layout_only) D->device = -2; else cudaGetDevice(&D->device);
	XmbHistParams &P = D->P;
	bool ok = true;
	// ---- node rows -----------------------------------------------------------------------------------
	P.nL = nL; P.nZ = nZ;
	P.off_exc = nL;
	P.off_elem = nL + 1;
	P.row_stride = (nL + 1 + nZ * XMB_ELEM_STRIDE + 1) & ~1;
	std::vector<double> rows((size_t)nN * P.row_stride, 0.0);
	for (int n = 0; n < nN; n++) {
		double *r = &rows[(size_t)n * P.row_stride];
		for (int k = 0; k < nL; k++) r[k] = T.mu_layer[(size_t)k * nN + n];
		r[P.off_exc] = T.exc_murhod[n];
		for (int z = 0; z < nZ; z++) {
			double *e = r + P.off_elem + z * XMB_ELEM_STRIDE;
			e[XMB_EO_CS_TOTAL] = T.cs_total[(size_t)z * nN + n];
			e[XMB_EO_P_RAYL] = T.p_rayl[(size_t)z * nN + n];
			e[XMB_EO_P_RAYL_COMPT] = T.p_rayl_compt[(size_t)z * nN + n];
			e[XMB_EO_PHOTO_TOTAL] = T.cs_photo_total[(size_t)z * nN + n];
			for (int s = 0; s < 9; s++) {
				e[XMB_EO_PHOTO_PARTIAL + s] = T.cs_photo_partial[((size_t)z * 9 + s) * nN + n];
				e[XMB_EO_VACANCY + s] = T.cs_vacancy[(((size_t)(D->cascade - 1) * nZ + z) * 9 + s) * nN + n];
			}
		}
	}
	P.rows = upload(D, rows.data(), rows.size(), ok);
	P.n_nodes = nN; P.n_buckets = T.n_buckets; P.bucket_E0 = T.bucket_E0; P.bucket_inv_dE = T.bucket_inv_dE;
	P.node_E = upload(D, T.node_E, nN, ok);
	{
		std::vector<int> bs(T.bucket_start, T.bucket_start + T.n_buckets + 1);
		for (int b = 0; b + 1 <= T.n_buckets; b++) {
			const int i = T.bucket_start[b];
			const double lo = T.bucket_E0 + b / T.bucket_inv_dE, hi = T.bucket_E0 + (b + 1) / T.bucket_inv_dE;
			// simple: node i sits at the bucket's lower bound (within rounding) and node i+1 is at/after its upper bound
			if (i + 1 < nN && std::fabs(T.node_E[i] - lo) < 1e-9 && T.node_E[i + 1] >= hi - 1e-9) bs[b] = i | (int)0x80000000;
		}
		P.bucket_start = upload(D, bs.data(), bs.size(), ok);
	}
	// ---- inverse CDFs, form factors ---------------------------------------------------------------------
	P.n_icdf_E = T.n_icdf_E; P.n_icdf_R = T.n_icdf_R; P.n_phi_T = T.n_phi_T; P.n_cp = T.n_cp; P.n_q = T.n_q;
	P.q_max = T.q_max; P.cp_dR = T.cp_R[1] - T.cp_R[0]; P.cp_inv_dR = (double)(T.n_cp - 1);
	P.icdf_E = upload(D, T.icdf_E, T.n_icdf_E, ok);
	P.icdf_R = upload(D, T.icdf_R, T.n_icdf_R, ok);
	P.phi_T = upload(D, T.phi_T, T.n_phi_T, ok);
	P.cp_R = upload(D, T.cp_R, T.n_cp, ok);
	P.rayl_icdf = upload(D, T.rayl_theta_icdf, (size_t)nZ * T.n_icdf_E * T.n_icdf_R, ok);
	P.compt_icdf = upload(D, T.compt_theta_icdf, (size_t)nZ * T.n_icdf_E * T.n_icdf_R, ok);
	P.phi_icdf = upload(D, T.phi_icdf, (size_t)T.n_phi_T * T.n_icdf_R, ok);
	P.cp_icdf = upload(D, T.cp_icdf, (size_t)nZ * T.n_cp, ok);
	if (T.n_adv_rows > 0) {
		P.adv_off = upload(D, T.adv_off, nZ + 1, ok);
		P.adv_config = upload(D, T.adv_config, T.n_adv_rows, ok);
		P.adv_edge = upload(D, T.adv_edge, T.n_adv_rows, ok);
		P.adv_cdf = upload(D, T.adv_cdf, (size_t)T.n_adv_rows * T.n_cp, ok);
		P.adv_qinv = upload(D, T.adv_qinv, (size_t)T.n_adv_rows * T.n_cp, ok);
	}
	P.ff = upload(D, T.ff, (size_t)nZ * T.n_q, ok);
	P.sf = upload(D, T.sf, (size_t)nZ * T.n_q, ok);
	// ---- per-element constants ----------------------------------------------------------------------------
	P.atomic_weight = upload(D, T.atomic_weight, nZ, ok);
	{
		std::vector<double> aoa(nZ);
		for (int z = 0; z < nZ; z++) aoa[z] = AVOGNUM / T.atomic_weight[z];
		P.avog_over_A = upload(D, aoa.data(), nZ, ok);
	}
	std::vector<double> edgeK(nZ);
	for (int z = 0; z < nZ; z++) edgeK[z] = T.edge_energy[z * 9 + 0];
	P.edge_K = upload(D, edgeK.data(), nZ, ok);
	P.fluor_yield_corr = upload(D, T.fluor_yield_corr, (size_t)nZ * 9, ok);
	P.cos_kron = upload(D, T.cos_kron, (size_t)nZ * XMB_N_CK, ok);
	P.rad_rate = upload(D, T.rad_rate, (size_t)nZ * 384, ok);
	P.line_energy = upload(D, T.line_energy, (size_t)nZ * 384, ok);
	// ---- layers ---------------------------------------------------------------------------------------------
	std::vector<XmbLayerDev> layers(nL);
	std::vector<int> elem_zi;
	std::vector<double> elem_w;
	for (int k = 0; k < nL; k++) {
		const xmb_layer &l = I.composition->layers[k];
		layers[k].n_elements = l.n_elements;
		layers[k].elem_begin = (int)elem_zi.size();
		layers[k].density = l.density;
		layers[k].Z_begin = in->Z_coord_begin[k];
		layers[k].Z_end = in->Z_coord_end[k];
		for (int e = 0; e < l.n_elements; e++) { elem_zi.push_back(T.uniqZ[l.Z[e]]); elem_w.push_back(l.weight[e]); }
	}
	for (int k = 0; k < nL; k++) D->max_nE = std::max(D->max_nE, layers[k].n_elements);
	P.layers = upload(D, layers.data(), nL, ok);
	P.elem_zi = upload(D, elem_zi.data(), elem_zi.size(), ok);
	P.elem_w = upload(D, elem_w.data(), elem_w.size(), ok);
	// ---- forced-detection line records (active lines only) and history slots --------------------------------
	const int line_last = opt->use_M_lines ? XMB_M5P5 : XMB_L3Q1;
	const xmb_detector &det = *I.detector;
	std::vector<int> rec_begin((size_t)nZ * 10, 0);
	std::vector<double> rec_yr, rec_mu;
	D->hist_base.assign(nZ, 0);
	int slot = 0;
	for (int z = 0; z < nZ; z++) {
		D->hist_base[z] = slot;
		slot += 2;   // +0 Rayleigh (history slot 384), +1 Compton (385)
		for (int s = 0; s < 9; s++) {
			rec_begin[z * 10 + s] = (int)rec_yr.size();
			for (int l = 1; l <= line_last; l++) {
				if (vr_shell_of_line(l) != s) continue;
				const double E = T.line_energy[(size_t)z * 384 + l];
				if (E < ENERGY_THRESHOLD) continue;                      // :579
				const double yr = T.fluor_yield[z * 9 + s] * T.rad_rate[(size_t)z * 384 + l];
				if (yr <= 0.0) continue;
				// a line above the table window lies above the highest source energy: its shell can never be
				// ionised (E_line < E_edge), the reference's P_shell == 0 skip (:589-645)
				if (E >= T.node_E[nN - 1]) continue;
				rec_yr.push_back(yr);
				// mu of every layer at the line energy: exact node lookup (precalc_mu_cs, src/xmi_main.F90:227-237)
				const double *ne = std::lower_bound(T.node_E, T.node_E + nN, E);
				const int node = (int)(ne - T.node_E);
				if (node >= nN || T.node_E[node] != E) { xmb_set_error("line energy is not a table node"); delete D; return nullptr; }
				for (int k = 0; k < nL; k++) rec_mu.push_back(T.mu_layer[(size_t)k * nN + node]);
				D->rec_slot.push_back(slot++);
				int ch = -1;
				if (E >= ENERGY_THRESHOLD) { ch = (int)((E - det.zero) / det.gain); if (ch < 0 || ch > det.nchannels - 1) ch = -1; }
				D->rec_channel.push_back(ch);
				D->rec_line.push_back(l);
				D->rec_zi.push_back(z);
			}
		}
		rec_begin[z * 10 + 9] = (int)rec_yr.size();
	}
	D->n_rec = (int)rec_yr.size();
	D->n_hist_slots = slot;
	P.n_hist_slots = slot;
	P.rec_begin = upload(D, rec_begin.data(), rec_begin.size(), ok);
	P.rec_yr = upload(D, rec_yr.data(), rec_yr.size(), ok);
	P.rec_mu = upload(D, rec_mu.data(), rec_mu.size(), ok);
	P.rec_slot = upload(D, D->rec_slot.data(), D->rec_slot.size(), ok);
	P.hist_base = upload(D, D->hist_base.data(), nZ, ok);
	{
		std::vector<int> ls((size_t)nZ * 384, -1);
		for (int r = 0; r < D->n_rec; r++) ls[(size_t)D->rec_zi[r] * 384 + D->rec_line[r]] = D->rec_slot[r];
		D->line_slot = upload(D, ls.data(), ls.size(), ok);
		std::vector<double> cdf((size_t)nZ * XMB_N_AUGER);
		for (int z = 0; z < nZ; z++) {
			const double *a = T.auger_rate + (size_t)z * XMB_N_AUGER;
			double *c = &cdf[(size_t)z * XMB_N_AUGER];
			const int first[5] = {0, 240, 375, 510, 645};
			for (int b = 0; b < 4; b++) { double sum = 0.0; for (int k = first[b]; k < first[b + 1]; k++) { sum += a[k]; c[k] = sum; } }
		}
		D->auger_rate = upload(D, cdf.data(), cdf.size(), ok);
	}
	// ---- source segments (src/xmi_main.F90:319-338, :579-601) --------------------------------------------------
	const xmb_excitation &exc = *I.excitation;
	const xmb_general &gen = *I.general;
	std::vector<XmbSegDev> segs;
	double Wmax = 0.0;
	for (int i = 0; i + 1 < exc.n_continuous; i++) {
		const xmb_energy_continuous &a = exc.continuous[i], &b = exc.continuous[i + 1];
		const double y1 = a.vertical_intensity + a.horizontal_intensity, y2 = b.vertical_intensity + b.horizontal_intensity;
		const double total = (y1 + y2) * (b.energy - a.energy) / 2.0;
		if (total == 0.0) continue;
		XmbSegDev s{};
		s.is_cont = 1; s.x1 = a.energy; s.x2 = b.energy; s.y1 = y1; s.y2 = y2; s.h1 = a.horizontal_intensity; s.h2 = b.horizontal_intensity;
		s.total_rel = total / (double)gen.n_photons_interval;
		s.sigma_x = a.sigma_x; s.sigma_y = a.sigma_y; s.sigma_xp = a.sigma_xp; s.sigma_yp = a.sigma_yp;
		Wmax = std::max(Wmax, s.total_rel);
		segs.push_back(s);
	}
	const uint64_t n_cont_seg = segs.size();
	auto exc_corr = [&](double E) {
		double s = 0.0;
		for (int k = 0; k < I.absorbers->n_exc_layers; k++)
			s += I.absorbers->exc_layers[k].density * I.absorbers->exc_layers[k].thickness * xmb_host_mu_layer(h->xrl, &I.absorbers->exc_layers[k], E);
		return std::exp(-s);
	};
	for (int i = 0; i < exc.n_discrete; i++) {
		const xmb_energy_discrete &e = exc.discrete[i];
		const double total = e.vertical_intensity + e.horizontal_intensity;
		XmbSegDev s{};
		s.is_cont = 0; s.distribution_type = e.distribution_type; s.energy = e.energy; s.scale_parameter = e.scale_parameter;
		s.weight_rel = total * exc_corr(e.energy) / (double)gen.n_photons_line;
		s.hor_ver_ratio = e.horizontal_intensity * (double)gen.n_photons_line / total;
		s.sigma_x = e.sigma_x; s.sigma_y = e.sigma_y; s.sigma_xp = e.sigma_xp; s.sigma_yp = e.sigma_yp;
		Wmax = std::max(Wmax, s.weight_rel);
		segs.push_back(s);
	}
	if (segs.empty() || Wmax <= 0.0) { xmb_set_error("no excitation"); delete D; return nullptr; }
	for (auto &s : segs) { s.weight_rel /= Wmax; s.total_rel /= Wmax; }
	D->W_max = Wmax;
	P.segs = upload(D, segs.data(), segs.size(), ok);
	P.n_seg = (int)segs.size();
	P.n_cont_seg = n_cont_seg;
	P.n_per_interval = (uint64_t)gen.n_photons_interval;
	P.n_per_line = (uint64_t)gen.n_photons_line;
	D->n_total = n_cont_seg * P.n_per_interval + (uint64_t)exc.n_discrete * P.n_per_line;
	// ---- geometry, detector ---------------------------------------------------------------------------------------
	const xmb_geometry &g = *I.geometry;
	for (int i = 0; i < 3; i++) { P.n_sample[i] = g.n_sample_orientation[i]; P.p_window[i] = g.p_detector_window[i]; P.n_detector[i] = g.n_detector_orientation[i]; }
	for (int i = 0; i < 9; i++) { P.ndo_new[i] = in->der.ndo_new[i]; P.ndo_inv[i] = in->der.ndo_inv[i]; }
	P.detector_radius = in->der.detector_radius;
	P.slit_x1_max = std::atan(g.slit_size_x / g.d_source_slit / 2.0);
	P.slit_y1_max = std::atan(g.slit_size_y / g.d_source_slit / 2.0);
	P.d_source_slit = g.d_source_slit;
	P.n_int = gen.n_interactions_trajectory;
	P.nch = det.nchannels; P.zero = det.zero; P.gain = det.gain;
	P.use_M_lines = D->use_M_lines;
	if (!ok) { xmb_set_error("device table upload failed: %s", cudaGetErrorString(cudaGetLastError())); delete D; return nullptr; }
	return D;
}

extern "C" int xmb_main_msim_raw(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                                 const xmb_solid_angle *sa, xmb_msim_ex *ex, uint64_t **accum, size_t *n_slots) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !in->inited || !options || !ex || !accum || !n_slots) { xmb_set_error("xmb_main_msim_raw: bad arguments"); return 0; }
	const bool brute = !options->use_variance_reduction;
	if (options->use_advanced_compton && !xmb_tables_enable_advanced_compton(hdf5F)) return 0;   // builds the subshell tables once
	if (options->escape_ratios_mode) { xmb_set_error("escape_ratios_mode is not implemented on the GPU path"); return 0; }
	if (!brute && (!sa || !sa->solid_angles)) { xmb_set_error("variance reduction needs a solid-angle grid"); return 0; }
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: xmb_main_msim has no CPU fallback"); return 0; }
	if (ex->device >= 0) XMB_CUDA_OK(cudaSetDevice(ex->device));
	int dev = 0;
	cudaGetDevice(&dev);
	XmbDeviceTables *D = h->dev;
	if (!D || D->cascade != cascade_mode(options) || D->use_M_lines != (options->use_M_lines ? 1 : 0) || D->device != dev) {
		if (D) delete D;
		h->dev = D = build_device_tables(in, h, options);
		if (!D) return 0;
	}
	XmbHistParams P = D->P;
	// solid-angle grid: an argument of the call -> copied host->device every call
	const size_t nsa = brute ? 0 : (size_t)sa->grid_dims_r_n * sa->grid_dims_theta_n;
	if (!brute && (D->sa_cap < nsa || !D->sa_grid)) {
		cudaFree(D->sa_grid); cudaFree(D->sa_r); cudaFree(D->sa_t);
		XMB_CUDA_OK(cudaMalloc(&D->sa_grid, sizeof(double) * nsa));
		XMB_CUDA_OK(cudaMalloc(&D->sa_r, sizeof(double) * sa->grid_dims_r_n));
		XMB_CUDA_OK(cudaMalloc(&D->sa_t, sizeof(double) * sa->grid_dims_theta_n));
		D->sa_cap = nsa;
	}
	const bool resident = brute || (ex->keep_on_device && D->sa_host == sa->solid_angles && D->sa_n == nsa);
	if (!resident) {
		XMB_CUDA_OK(cudaMemcpy(D->sa_grid, sa->solid_angles, sizeof(double) * nsa, cudaMemcpyHostToDevice));
		XMB_CUDA_OK(cudaMemcpy(D->sa_r, sa->grid_dims_r_vals, sizeof(double) * sa->grid_dims_r_n, cudaMemcpyHostToDevice));
		XMB_CUDA_OK(cudaMemcpy(D->sa_t, sa->grid_dims_theta_vals, sizeof(double) * sa->grid_dims_theta_n, cudaMemcpyHostToDevice));
		D->sa_host = sa->solid_angles; D->sa_n = nsa;
	}
	P.sa_grid = D->sa_grid; P.sa_r_vals = D->sa_r; P.sa_t_vals = D->sa_t;
	if (!brute) { P.sa_nr = (int)sa->grid_dims_r_n; P.sa_nt = (int)sa->grid_dims_theta_n; }
	// accumulators: one row per interaction order (brute force: rows 0..n_int, row = interactions before detection)
	const size_t slots = (size_t)(P.n_int + (brute ? 1 : 0)) * ((size_t)P.nch + P.n_hist_slots);
	if (D->acc_slots != slots) {
		cudaFree(D->acc); cudaFree(D->limbs); cudaFree(D->counters);
		XMB_CUDA_OK(cudaMalloc(&D->acc, sizeof(unsigned long long) * 2 * slots));
		XMB_CUDA_OK(cudaMalloc(&D->limbs, sizeof(unsigned long long) * 2 * slots));
		XMB_CUDA_OK(cudaMalloc(&D->counters, sizeof(unsigned long long) * (8 + XMB_MAX_LAYERS)));
		D->acc_slots = slots;
	}
	XMB_CUDA_OK(cudaMemsetAsync(D->acc, 0, sizeof(unsigned long long) * 2 * slots));
	XMB_CUDA_OK(cudaMemsetAsync(D->counters, 0, sizeof(unsigned long long) * (8 + XMB_MAX_LAYERS)));
	P.acc = D->acc; P.counters = D->counters;
	// shard of global photon ids
	const int nr = ex->n_ranks > 0 ? ex->n_ranks : 1, rk = ex->rank;
	if (rk < 0 || rk >= nr) { xmb_set_error("rank %d outside 0..%d", rk, nr - 1); return 0; }
	P.seed = ex->seed ? ex->seed : XMB_DEFAULT_SEED;
	P.n_total = D->n_total; P.shard_rank = rk; P.shard_n = nr;
	{
		const uint64_t blocks = (D->n_total + XMB_SHARD_BLOCK - 1) / XMB_SHARD_BLOCK;
		P.n_local_span = (blocks / nr + ((uint64_t)rk < blocks % nr ? 1 : 0)) * XMB_SHARD_BLOCK;
	}
	ex->n_histories = xmb_msim_shard_count(D->n_total, rk, nr);
	// launch
	int sms = 148, occ = 1;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (brute) {
		XmbBruteParams B{};
		B.use_auger = options->use_cascade_auger ? 1 : 0; B.use_rad = options->use_cascade_radiative ? 1 : 0;
		B.collimator_present = in->der.collimator_present; B.collimator_height = in->der.collimator_height;
		B.collimator_radius = in->der.collimator_radius; B.half_apex = in->der.half_apex;
		B.vertex_x = in->der.vertex[0]; B.vertex_y = in->der.vertex[1]; B.vertex_z = in->der.vertex[2];
		B.line_slot = D->line_slot; B.auger_rate = D->auger_rate;
		const int bt = XMB_BRUTE_THREADS;
		const uint64_t want = (ex->n_histories + bt - 1) / bt;
		const unsigned bg = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms));   // persistent: one CTA per SM
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0);
		if (ex->n_histories > 0) {
			if (options->use_advanced_compton) xmb_brute_kernel<0, true><<<bg, bt>>>(P, B);
			else switch (P.nL) {
			case 1: xmb_brute_kernel<1><<<bg, bt>>>(P, B); break;
			case 2: xmb_brute_kernel<2><<<bg, bt>>>(P, B); break;
			case 3: xmb_brute_kernel<3><<<bg, bt>>>(P, B); break;
			default: xmb_brute_kernel<0><<<bg, bt>>>(P, B); break;
			}
		}
		cudaEventRecord(e1);
		xmb_limbs_kernel<<<sms, 256>>>(D->acc, D->limbs, slots);
		XMB_CUDA_OK(cudaGetLastError());
		XMB_CUDA_OK(cudaEventSynchronize(e1));
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		cudaEventDestroy(e0); cudaEventDestroy(e1);
		ex->kernel_ms = ms;
		ex->n_launches = (ex->n_histories > 0 ? 1 : 0) + 1;
		unsigned long long cnt[8];
		XMB_CUDA_OK(cudaMemcpy(cnt, D->counters, sizeof(cnt), cudaMemcpyDeviceToHost));
		for (int i = 0; i < 8; i++) D->brute_counters[i] = cnt[i];
		ex->n_interactions = cnt[1];
		if (cnt[2]) { xmb_set_error("%llu deposits fell outside the fixed-point range", cnt[2]); return 0; }
		if (cnt[5] && options->verbose) fprintf(stderr, "detected photons of lines without a history slot: %llu\n", cnt[5]);
		*n_slots = slots;
		if (ex->keep_on_device) { *accum = nullptr; return 1; }
		uint64_t *out = (uint64_t *)malloc(sizeof(uint64_t) * 2 * slots);
		XMB_CUDA_OK(cudaMemcpy(out, D->limbs, sizeof(uint64_t) * 2 * slots, cudaMemcpyDeviceToHost));
		*accum = out;
		return 1;
	}
	// threads per CTA: as many as the per-thread shared arrays (2 nL doubles) allow within 200 KB
	int threads = HIST_THREADS;
	const size_t stage_bytes = sizeof(unsigned long long) * 2 * ((size_t)P.nch + P.n_hist_slots);
	if (stage_bytes > 160 * 1024) { xmb_set_error("nchannels + history slots do not fit the shared-memory staging area"); return 0; }
	while (threads > 64 && stage_bytes + sizeof(double) * 2 * P.nL * threads > 200 * 1024) threads -= 32;
	// a staged 16-bit piece holds < 2^16 per addend and the word 2^32: at most 2^16 addends per slot and batch
	const size_t per_photon = (size_t)std::max(1, D->max_nE) * (options->use_advanced_compton ? 32 : 1);   // + one addend per subshell
	while (threads > 64 && (size_t)threads * per_photon > 60000) threads -= 32;
	const size_t smem = stage_bytes + sizeof(double) * 2 * P.nL * threads;
	void (*kernel)(const XmbHistParams) = P.nL == 1 ? xmb_history_kernel<1> : P.nL == 2 ? xmb_history_kernel<2> : P.nL == 3 ? xmb_history_kernel<3>
	                                     : P.nL == 4 ? xmb_history_kernel<4> : xmb_history_kernel<0>;
	if (options->use_advanced_compton) kernel = xmb_history_kernel<0, true>;   // opt-in physics: one generic-nL instantiation
	XMB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	XMB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
	if (occ < 1) occ = 1;
	const uint64_t n_chunks = (ex->n_histories + threads - 1) / threads;
	uint64_t blocks = (uint64_t)sms * occ;
	blocks = std::max<uint64_t>(1, std::min<uint64_t>(blocks, n_chunks));
	if (P.n_int > XMB_MAX_ORDERS) { xmb_set_error("more than %d interactions per trajectory", XMB_MAX_ORDERS); return 0; }
	// per-CTA compaction queues: n_int orders x 2T photons x (15 + nL) doubles (structure of arrays)
	const size_t qd = (size_t)blocks * P.n_int * (XMB_STATE_FIELDS + P.nL) * 2 * threads;
	if (D->queue_doubles < qd) {
		cudaFree(D->queue);
		D->queue = nullptr;
		XMB_CUDA_OK(cudaMalloc(&D->queue, sizeof(double) * qd));
		D->queue_doubles = qd;
	}
	P.queue = D->queue;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	if (ex->n_histories > 0) kernel<<<(unsigned)blocks, threads, smem>>>(P);
	cudaEventRecord(e1);
	xmb_limbs_kernel<<<sms, 256>>>(D->acc, D->limbs, slots);
	XMB_CUDA_OK(cudaGetLastError());
	XMB_CUDA_OK(cudaEventSynchronize(e1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	ex->kernel_ms = ms;
	ex->n_launches = (ex->n_histories > 0 ? 1 : 0) + 1;
	unsigned long long cnt[8 + XMB_MAX_LAYERS];
	XMB_CUDA_OK(cudaMemcpy(cnt, D->counters, sizeof(cnt), cudaMemcpyDeviceToHost));
	for (int i = 0; i < XMB_MAX_LAYERS; i++) D->layer_interactions[i] = cnt[8 + i];
	ex->n_interactions = cnt[1];
	if (cnt[2]) { xmb_set_error("%llu deposits fell outside the fixed-point range", cnt[2]); return 0; }
	if (cnt[0] && options->verbose) fprintf(stderr, "detector_solid_angle_not_found: %llu\n", cnt[0]);
	*n_slots = slots;
	if (ex->keep_on_device) { *accum = nullptr; return 1; }   // limbs stay in HBM: xmb_msim_device_limbs()
	uint64_t *out = (uint64_t *)malloc(sizeof(uint64_t) * 2 * slots);
	XMB_CUDA_OK(cudaMemcpy(out, D->limbs, sizeof(uint64_t) * 2 * slots, cudaMemcpyDeviceToHost));
	*accum = out;
	return 1;
}

// Block-cyclic shard of the global photon ids [0, n_total): ids are dealt in blocks of XMB_SHARD_BLOCK, block b to
// rank b % n_ranks -- every rank gets the same mix of source lines (the reference gives each MPI host
// n_photons/n_hosts photons of EVERY line, src/xmi_main.F90:314,574), so the ranks finish together.
extern "C" int xmb_msim_shard_owner(uint64_t g, int n_ranks) {
	if (n_ranks < 1) n_ranks = 1;
	return (int)((g >> XMB_SHARD_SHIFT) % (uint64_t)n_ranks);
}
extern "C" uint64_t xmb_msim_shard_count(uint64_t n_total, int rank, int n_ranks) {
	if (n_ranks < 1) n_ranks = 1;
	if (rank < 0 || rank >= n_ranks) return 0;
	const uint64_t blocks = (n_total + XMB_SHARD_BLOCK - 1) / XMB_SHARD_BLOCK;
	if (blocks == 0) return 0;
	uint64_t owned = blocks / n_ranks + ((uint64_t)rank < blocks % n_ranks ? 1 : 0);
	uint64_t n = owned * XMB_SHARD_BLOCK;
	// the last global block may be partial
	if ((blocks - 1) % n_ranks == (uint64_t)rank) n -= blocks * XMB_SHARD_BLOCK - n_total;
	return n;
}

static XmbDeviceTables *ensure_layout(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *opt) {
	if (h->dev && h->dev->cascade == cascade_mode(opt) && h->dev->use_M_lines == (opt->use_M_lines ? 1 : 0)) return h->dev;
	if (h->dev) { delete h->dev; h->dev = nullptr; }
	g_layout_only = true;
	h->dev = build_device_tables(in, h, opt);
	g_layout_only = false;
	return h->dev;
}

extern "C" uint64_t xmb_msim_total_histories(xmb_inputFPtr inputF) {
	XmbInputF *in = xmb_as_input(inputF);
	if (!in) return 0;
	const xmb_excitation &exc = *in->in.excitation;
	uint64_t n = 0;
	for (int i = 0; i + 1 < exc.n_continuous; i++) {
		const double y1 = exc.continuous[i].vertical_intensity + exc.continuous[i].horizontal_intensity;
		const double y2 = exc.continuous[i + 1].vertical_intensity + exc.continuous[i + 1].horizontal_intensity;
		if ((y1 + y2) * (exc.continuous[i + 1].energy - exc.continuous[i].energy) / 2.0 != 0.0) n += (uint64_t)in->in.general->n_photons_interval;
	}
	return n + (uint64_t)exc.n_discrete * (uint64_t)in->in.general->n_photons_line;
}

// slot map of the accumulator rows: slots [0, nch) are channels; slot nch + s is history slot s with
// (Z, line) = (out_Z[s], out_line[s]); line 384 / 385 = Rayleigh / Compton.  Returns n_hist_slots (0 on error).
extern "C" int xmb_msim_slot_map(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options, int32_t *out_Z,
                                 int32_t *out_line, int capacity) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !options) return 0;
	XmbDeviceTables *D = ensure_layout(in, h, options);
	if (!D) return 0;
	if (out_Z && out_line) {
		if (capacity < D->n_hist_slots) { xmb_set_error("slot map capacity"); return 0; }
		for (int z = 0; z < h->view.nZ; z++) {
			out_Z[D->hist_base[z]] = h->view.Z[z]; out_line[D->hist_base[z]] = 384;
			out_Z[D->hist_base[z] + 1] = h->view.Z[z]; out_line[D->hist_base[z] + 1] = 385;
		}
		for (int r = 0; r < D->n_rec; r++) { out_Z[D->rec_slot[r]] = h->view.Z[D->rec_zi[r]]; out_line[D->rec_slot[r]] = D->rec_line[r]; }
	}
	return D->n_hist_slots;
}

extern "C" int xmb_main_msim_finish(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                                    const uint64_t *accum, size_t n_slots, double **channels, double **brute_history,
                                    double **var_red_history) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !accum || !options) { xmb_set_error("xmb_main_msim_finish: bad arguments"); return 0; }
	XmbDeviceTables *D = ensure_layout(in, h, options);
	if (!D) return 0;
	const int n_int = D->P.n_int, nch = D->P.nch;
	const size_t row = (size_t)nch + D->n_hist_slots;
	const bool brute = !options->use_variance_reduction;
	if (n_slots != (size_t)(n_int + (brute ? 1 : 0)) * row) { xmb_set_error("xmb_main_msim_finish: slot count mismatch"); return 0; }
	const double live_time = in->in.detector->live_time;
	const double scale = D->W_max * live_time;
	auto slot128 = [&](size_t i) { return (unsigned __int128)accum[2 * i] + ((unsigned __int128)accum[2 * i + 1] << 48); };
	auto to_double = [&](unsigned __int128 v) {
		const double hi = (double)(uint64_t)(v >> 64), lo = (double)(uint64_t)v;
		return (hi * 18446744073709551616.0 + lo) * (1.0 / 72057594037927936.0) * scale;
	};
	double *ch = (double *)calloc((size_t)(n_int + 1) * nch, sizeof(double));
	double *vr = (double *)calloc((size_t)100 * 385 * n_int, sizeof(double));
	double *br = (double *)calloc((size_t)100 * 385 * n_int, sizeof(double));
	std::vector<unsigned __int128> cum(nch, 0), cur(nch);
	const xmb_tables_host &T = h->view;
	if (brute) {
		// rows 0..n_int hold what photons with that many interactions left in the detector: channels(k:, ch) += w
		// (src/xmi_main.F90:470-485); history slots go to brute_history(Z, slot, k) (:497-523)
		for (int k = 0; k <= n_int; k++) {
			for (int c = 0; c < nch; c++) { cum[c] += slot128((size_t)k * row + c); ch[(size_t)k * nch + c] = to_double(cum[c]); }
			if (k == 0) continue;
			for (int r = 0; r < D->n_rec; r++)
				br[((size_t)(T.Z[D->rec_zi[r]] - 1) * 385 + (D->rec_line[r] - 1)) * n_int + (k - 1)] =
				    to_double(slot128((size_t)k * row + nch + D->rec_slot[r]));
			for (int z = 0; z < T.nZ; z++) {
				br[((size_t)(T.Z[z] - 1) * 385 + 383) * n_int + (k - 1)] = to_double(slot128((size_t)k * row + nch + D->hist_base[z] + 0));
				br[((size_t)(T.Z[z] - 1) * 385 + 384) * n_int + (k - 1)] = to_double(slot128((size_t)k * row + nch + D->hist_base[z] + 1));
			}
		}
		if (channels) *channels = ch; else free(ch);
		free(vr);
		if (var_red_history) *var_red_history = nullptr;   // as the reference: no array without variance reduction (src/xmi_main.F90:942)
		if (brute_history) *brute_history = br; else free(br);
		return 1;
	}
	for (int k = 0; k < n_int; k++) {
		for (int c = 0; c < nch; c++) cur[c] = slot128((size_t)k * row + c);
		// XRF deposits: channel content rebuilt from the per-line slots (exact integer sums)
		for (int r = 0; r < D->n_rec; r++) {
			const unsigned __int128 v = slot128((size_t)k * row + nch + D->rec_slot[r]);
			if (D->rec_channel[r] >= 0) cur[D->rec_channel[r]] += v;
			// var_red_history[Z-1][|line|-1][k]   (C order of the export, src/xmi_main.F90:934-940)
			vr[((size_t)(T.Z[D->rec_zi[r]] - 1) * 385 + (D->rec_line[r] - 1)) * n_int + k] = to_double(v);
		}
		for (int z = 0; z < T.nZ; z++) {
			vr[((size_t)(T.Z[z] - 1) * 385 + 383) * n_int + k] = to_double(slot128((size_t)k * row + nch + D->hist_base[z] + 0));
			vr[((size_t)(T.Z[z] - 1) * 385 + 384) * n_int + k] = to_double(slot128((size_t)k * row + nch + D->hist_base[z] + 1));
		}
		// rows are cumulative over interaction order: channels(n_ia:, ch) += w
		for (int c = 0; c < nch; c++) { cum[c] += cur[c]; ch[(size_t)(k + 1) * nch + c] = to_double(cum[c]); }
	}
	if (channels) *channels = ch; else free(ch);
	if (var_red_history) *var_red_history = vr; else free(vr);
	if (brute_history) *brute_history = br; else free(br);
	(void)options;
	return 1;
}

static int env_rank(int n) {
	const char *names[] = {"OMPI_COMM_WORLD_RANK", "PMI_RANK", "RANK"};
	for (const char *nm : names) { const char *v = getenv(nm); if (v) { int r = atoi(v); if (r >= 0 && r < n) return r; } }
	return -1;
}

extern "C" int xmb_main_msim(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, int n_mpi_hosts, double **channels,
                             const xmb_main_options *options, double **brute_history, double **var_red_history,
                             const xmb_solid_angle *solid_angles) {
	xmb_msim_ex ex{};
	ex.n_ranks = n_mpi_hosts > 0 ? n_mpi_hosts : 1;
	ex.rank = 0;
	ex.device = -1;
	if (ex.n_ranks > 1) {
		ex.rank = env_rank(ex.n_ranks);
		if (ex.rank < 0) { xmb_set_error("n_mpi_hosts > 1 but no rank in OMPI_COMM_WORLD_RANK / PMI_RANK / RANK"); return 0; }
	}
	if (options && options->verbose) { printf("Simulating interactions\n"); fflush(stdout); }
	uint64_t *acc = nullptr;
	size_t n = 0;
	if (!xmb_main_msim_raw(inputF, hdf5F, options, solid_angles, &ex, &acc, &n)) return 0;
	const int rv = xmb_main_msim_finish(inputF, hdf5F, options, acc, n, channels, brute_history, var_red_history);
	free(acc);
	if (rv && options && options->verbose) { printf("Simulating interactions at 100 %%\nInteractions simulation finished\n"); fflush(stdout); }
	return rv;
}

// Device-resident view of the last run's limbs (for an NCCL all-reduce without host staging).
extern "C" int xmb_msim_device_limbs(xmb_hdf5FPtr hdf5F, uint64_t **dev_ptr, size_t *n_words) {
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!h || !h->dev || !h->dev->limbs) { xmb_set_error("no device accumulators"); return 0; }
	*dev_ptr = (uint64_t *)h->dev->limbs;
	*n_words = 2 * h->dev->acc_slots;
	return 1;
}

// Counters of the last brute-force run: [1] interactions, [3] detector hits, [4] offspring photons walked,
// [5] detected photons whose line has no history slot.
extern "C" int xmb_msim_brute_counters(xmb_hdf5FPtr hdf5F, uint64_t *out, int capacity) {
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!h || !h->dev || !out) { xmb_set_error("no run to describe"); return 0; }
	for (int i = 0; i < capacity && i < 8; i++) out[i] = h->dev->brute_counters[i];
	return 1;
}

// Workload description of the last run, for the algorithmic-bytes figure of DESIGN.md / SURVEY.md 8d:
// out[0] = n_layers, then per layer: interactions, n_elements, active line records of its elements.
extern "C" int xmb_msim_workload_stats(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, uint64_t *out, int capacity) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !h->dev) { xmb_set_error("no run to describe"); return 0; }
	XmbDeviceTables *D = h->dev;
	const xmb_composition &c = *in->in.composition;
	if (capacity < 1 + 3 * c.n_layers) return 0;
	std::vector<int> per_z(h->view.nZ, 0);
	for (int r = 0; r < D->n_rec; r++) per_z[D->rec_zi[r]]++;
	out[0] = c.n_layers;
	for (int k = 0; k < c.n_layers; k++) {
		uint64_t act = 0;
		for (int e = 0; e < c.layers[k].n_elements; e++) act += per_z[h->view.uniqZ[c.layers[k].Z[e]]];
		out[1 + 3 * k] = D->layer_interactions[k];
		out[2 + 3 * k] = c.layers[k].n_elements;
		out[3 + 3 * k] = act;
	}
	return 1;
}

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
		for (int i = 0; i < nL; i++) mus0[i] = row_lerp(P, np, i);
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
		p.theta = acos(p.dz);
		p.phi = atan2(p.dy, p.dx);
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
		select_and_scatter<NL, 1>(P, p, g, 1, mus, 1, b0.w, weight_escape, type, zi, line, shell_unused);
		// ---- second iteration: analogue free path (:1229-1413); escaped = no interaction before the surface ----
		if (p.energy < ENERGY_THRESHOLD) continue;   // EXIT main with inside still true (:1229-1231)
		bool escaped = true;
		{
			int step_max, step_dir;
			if (p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0) { step_max = nL - 1; step_dir = 1; }
			else { step_max = 0; step_dir = -1; }
			const double interactionR = xmb_u01(draw_block(P.seed, g, 2, 1, 0, 0).x);
			double blbs = 1.0, max_random_layer = 0.0;
			double lx = p.cx, ly = p.cy, lz = p.cz;
			for (int i = p.layer; step_dir > 0 ? i <= step_max : i >= step_max; i += step_dir) {
				double dist;
				if (!step_to_plane(P, lx, ly, lz, p.dx, p.dy, p.dz, step_dir == 1 ? P.layers[i].Z_end : P.layers[i].Z_begin, dist)) { escaped = false; break; }
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
	int dev = 0;
	cudaGetDevice(&dev);
	XmbDeviceTables *D = h->dev;
	if (!D || D->cascade != cascade_mode(&opt) || D->use_M_lines != 0 || D->device != dev) {
		if (D) delete D;
		h->dev = D = build_device_tables(in, h, &opt);
		if (!D) return 0;
	}
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

