// history_device.cuh -- device code shared by the three Monte Carlo kernels (history.cu: forced detection,
// brute.cu: analogue walk with detector hits, escape.cu: escape-peak ratios): table lookups, exact deposits, the
// fixed-address Philox streams, source sampling (src/xmi_main.F90:319-438, :579-724, :957-1186), the interaction
// (atom and type selection, Rayleigh / Compton / photo-electric with Coster-Kronig and line selection,
// src/xmi_main.F90:1558-1652, :1986-2411, :4785-5437).
#pragma once
#include <cstdint>
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
#define XMB_REC_UNROLL 4
#endif
#define XMB_MAX_ORDERS 64
#ifndef XMB_MAX_QL
#define XMB_MAX_QL 1024          // (order, layer) queues of a CTA when batches are formed per layer
#endif
#define XMB_STATE_FIELDS 13      // 11 doubles of photon state + photon id + layer (mus[nL] follow)
#define XMB_PRAGMA(x) _Pragma(#x)
#define XMB_UNROLL_NL _Pragma("unroll")
#define XMB_UNROLL(n) XMB_PRAGMA(unroll n)
#ifndef HIST_MIN_BLOCKS
#define HIST_MIN_BLOCKS 1
#endif

static __constant__ short d_shell_line_first[9] = {1, 30, 59, 86, 118, 140, 161, 182, 201};   // = xmb_shell_line_first
static __constant__ short d_shell_line_last[9] = {29, 58, 85, 113, 136, 158, 180, 200, 219};    // = xmb_shell_line_last

struct NodePos { int pos; double f; };

__device__ __forceinline__ NodePos node_find(const XmbHistParams &P, double E) {
	int b = (int)floor((E - P.bucket_E0) * P.bucket_inv_dE);
	b = max(0, min(b, P.n_buckets - 1));
	// bucket_start = the last node at or below the bucket's lower bound.  Bit 31: no node lies inside the bucket or within
	// 1e-9 keV of its bounds, so every energy that maps to the bucket has this bracket and the node energies need not be
	// compared first -- the gathers that depend on the position start one memory round trip earlier (build_device_tables:
	// buckets 16 x finer than the uniform node spacing; the first bucket of a cell and those around edge doublets and
	// line energies take the checked path).
	const int bs = P.bucket_start[b];
	int i = bs & 0x7FFFFFFF;
	if (bs >= 0) {
		if (E < P.node_E[i] || E >= P.node_E[i + 1]) {
			while (i > 0 && P.node_E[i] > E) i--;
			while (i + 1 < P.n_nodes - 1 && P.node_E[i + 1] <= E) i++;
			i = min(i, P.n_nodes - 2);
		}
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
// mu of layer j: the same interpolation on the compact table mu_tab[node][nL] (the two nodes of a bracket are adjacent
// in memory and the nodes a Compton line scatters into fill a few KB, instead of one 32-byte sector per 3 KB node row)
__device__ __forceinline__ double mu_lerp(const XmbHistParams &P, NodePos np, int j) {
	const double *r0 = P.mu_tab + (size_t)np.pos * P.nL + j;
	const double a = r0[0], b = r0[P.nL];
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
// interaction order) are staged in shared memory: a slot is four 32-bit words holding 16- or 20-bit pieces of the
// addends (native 32-bit ATOMS.ADD: 64-bit shared atomics are CAS spin loops on sm_100a -- ATOMS.CAST.SPIN.64 -- and
// collapse when the lanes of a warp hit the same channel; layouts below).  After the batch the CTA folds every non-zero
// slot into the global 128-bit (lo, hi) accumulator.  Ablation on B200 (profiles/r1_history_ablation.txt): with
// per-lane global REDs the Compton peak's ~50 hot channel words serialised in L2 and cost 47 % of the kernel.
__device__ __forceinline__ unsigned long long to_fixed(double w, unsigned long long *counters) {
	const double s = w * 72057594037927936.0;   // 2^56
	if (!(s < 2.8e17)) { if (s == s) atomicAdd(&counters[2], 1ULL); return 0ULL; }   // w >= ~4: counted, never wrapped
	return __double2ull_rn(s);
}
// hot-loop variant: the caller has scaled by 2^56 and bounded the value already (see the line loop of the history
// kernel: one range check per shell instead of one per line)
__device__ __forceinline__ unsigned long long fixed_from_scaled(double s) { return __double2ull_rn(s); }
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
// ---- staged deposits, hot path ------------------------------------------------------------------------
// The staging area is addressed through its 32-bit shared-window address (one register, computed once per thread):
// with generic pointers every deposit rebuilt that address (S2R SR_CgaCtaId + LEA + IMAD chain, ~10 instructions, in
// the dependent chain of the atomic) because the kernel sits at its 64-register cap (profiles/r1_history_kernel_v10_*).
// No "memory" clobber: the staging words are only read back by flush_staged() behind a __syncthreads().
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_red(unsigned addr, unsigned v) {   // v == 0: no atomic issued
	asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q red.shared.add.u32 [%0], %1; }" ::"r"(addr), "r"(v));
}
// Two piece layouts of a staged slot (four 32-bit words):
//  * 16-bit pieces in words 0..3 -- channel slots: a channel can take max nE addends per photon and batch, any lane mix;
//  * 20-bit pieces in words 0..2 (P20) -- history slots (a line, or Rayleigh / Compton of an element): at most ONE
//    addend per photon and batch, always through the warp sum below: a word receives <= 32 warp sums < 2^25 per batch.
//    One REDUX, one atomic and the mask/shift of a piece less than the 16-bit layout on the most repeated code of the
//    kernel (one deposit per active line and interaction).
// all 32 lanes call; slot is warp-uniform; v < 2^60
__device__ __forceinline__ void deposit_uniform20(unsigned stage_s32, unsigned slot, unsigned long long v, int lane) {
	const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
	const unsigned int s0 = __reduce_add_sync(0xffffffffu, lo & 0xFFFFFu);
	const unsigned int s1 = __reduce_add_sync(0xffffffffu, __funnelshift_r(lo, hi, 20) & 0xFFFFFu);
	const unsigned int s2 = __reduce_add_sync(0xffffffffu, hi >> 8);
	unsigned int piece;   // lane 0: s0, lane 1: s1, lane 2: s2, others 0 -- selects, not branches
	asm("{ .reg .pred a, b, c;\n\t"
	    "setp.eq.u32 a, %4, 0; setp.eq.u32 b, %4, 1; setp.gt.u32 c, %4, 2;\n\t"
	    "selp.u32 %0, %1, %3, a; selp.u32 %0, %2, %0, b; selp.u32 %0, 0, %0, c; }"
	    : "=&r"(piece) : "r"(s0), "r"(s1), "r"(s2), "r"(lane));
	stage_red(stage_s32 + slot * 16u + (unsigned)lane * 4u, piece);
}
// same with 16-bit pieces (channel slots; every slot when the kernel runs without the 20-bit layout)
__device__ __forceinline__ void deposit_uniform16(unsigned stage_s32, unsigned slot, unsigned long long v, int lane) {
	const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
	const unsigned int s0 = __reduce_add_sync(0xffffffffu, lo & 0xFFFFu), s1 = __reduce_add_sync(0xffffffffu, lo >> 16);
	const unsigned int s2 = __reduce_add_sync(0xffffffffu, hi & 0xFFFFu), s3 = __reduce_add_sync(0xffffffffu, hi >> 16);
	unsigned int piece;
	asm("{ .reg .pred a, b, c, d;\n\t"
	    "setp.eq.u32 a, %5, 0; setp.eq.u32 b, %5, 1; setp.eq.u32 c, %5, 2; setp.gt.u32 d, %5, 3;\n\t"
	    "selp.u32 %0, %1, %4, a; selp.u32 %0, %2, %0, b; selp.u32 %0, %3, %0, c; selp.u32 %0, 0, %0, d; }"
	    : "=&r"(piece) : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(lane));
	stage_red(stage_s32 + slot * 16u + (unsigned)lane * 4u, piece);
}
template <bool P20>
__device__ __forceinline__ void deposit_uniform(unsigned stage_s32, unsigned slot, unsigned long long v, int lane) {
	if (P20) deposit_uniform20(stage_s32, slot, v, lane); else deposit_uniform16(stage_s32, slot, v, lane);
}
// channel deposit: all 32 lanes call; the channel may differ per lane (< 0: nothing to add); 16-bit pieces
__device__ __forceinline__ void deposit_varying(unsigned stage_s32, int slot, unsigned long long v, int lane) {
	const int s0 = __shfl_sync(0xffffffffu, slot, 0);
	if (__all_sync(0xffffffffu, slot == s0)) {
		if (s0 >= 0) deposit_uniform16(stage_s32, (unsigned)s0, v, lane);
	} else if (slot >= 0) {
		const unsigned a = stage_s32 + (unsigned)slot * 16u;
		const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
		stage_red(a, lo & 0xFFFFu); stage_red(a + 4u, lo >> 16); stage_red(a + 8u, hi & 0xFFFFu); stage_red(a + 12u, hi >> 16);
	}
}
// Fold the CTA's staged slots into the global accumulators of one interaction order and clear them.  A global slot is
// FOUR 64-bit words, one per 16-bit piece position (value = w0 + w1 2^16 + w2 2^32 + w3 2^48): each staged word is added to
// its own global word with a fire-and-forget RED -- no carry between words, so no atomic has to return a value (v15 added
// (lo, hi) pairs with a returning atomic for the carry: the flush waited a round trip to L2 per slot, 8 % of the kernel
// time on srm1412 and 13 % on the 10-layer sample, profiles/r2_history_phase_clocks.txt).  A staged word is < 2^32, so a
// global word takes 2^32 flushes; xmb_limbs4_kernel normalises the four words into limbs.
__device__ __forceinline__ void red_global_u64(unsigned long long *addr, unsigned long long v) {
	asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
template <bool P20>
__device__ __forceinline__ void flush_staged(unsigned int *stage, unsigned long long *global_row, int n_slots, int n_ch, int tid, int T) {
	static_assert(!P20, "the history kernel stages 16-bit pieces in every slot");
	for (int i = tid; i < n_slots; i += T) {
		const uint4 w = *reinterpret_cast<uint4 *>(stage + 4 * i);
		if ((w.x | w.y | w.z | w.w) == 0u) continue;
		*reinterpret_cast<uint4 *>(stage + 4 * i) = make_uint4(0u, 0u, 0u, 0u);
		unsigned long long *gs = global_row + 4 * (size_t)i;
		if (w.x) red_global_u64(gs + 0, w.x);
		if (w.y) red_global_u64(gs + 1, w.y);
		if (w.z) red_global_u64(gs + 2, w.z);
		if (w.w) red_global_u64(gs + 3, w.w);
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
	double energy, weight;   // (the polar angles of the direction, which the reference keeps beside it, are functions of dx, dy, dz: frame_trig())
	int layer;
	int n_interactions;
	bool alive;
};

__device__ __forceinline__ void normalize3(double &x, double &y, double &z) {
	const double n = sqrt(x * x + y * y + z * z);
	x /= n; y /= n; z /= n;
}

// Sines and cosines of the polar angles theta = acos(dz), phi = atan2(dy, dx) of the (unit) direction, which the reference stores
// with the photon and feeds to sin / cos at every scatter (src/xmi_main.F90:5071-5148, :2055-2066): taken from the components
// themselves -- no acos / atan2 when a direction is set, no two sincos when it is used.  atan2(0, 0) = 0: cos phi = 1, sin phi = 0.
__device__ __forceinline__ void frame_trig(const Photon &p, double &sth, double &cth, double &sph, double &cph) {
	cth = p.dz;
	sth = sqrt(p.dx * p.dx + p.dy * p.dy);
	if (sth > 0.0) { const double inv = 1.0 / sth; cph = p.dx * inv; sph = p.dy * inv; }
	else { cph = 1.0; sph = 0.0; }
}
// xmi_update_photon_dirv (src/xmi_main.F90:5071-5148): the new direction has polar angle theta_i (sti, cti) and azimuth phi_i
// (spn, cpn) in the frame of the old one
__device__ __forceinline__ void update_dirv(Photon &p, double sti, double cti, double spn, double cpn) {
	double sth, cth, sph, cph;
	frame_trig(p, sth, cth, sph, cph);
	const double v0 = sti * cpn, v1 = sti * spn, v2 = cti;
	p.dx = cth * cph * v0 + (-sph) * v1 + sth * cph * v2;
	p.dy = cth * sph * v0 + cph * v1 + sth * sph * v2;
	p.dz = (-sth) * v0 + 0.0 * v1 + cth * v2;
	normalize3(p.dx, p.dy, p.dz);
}
// xmi_update_photon_elecv (:5150-5182); sin(acos(c)) = sqrt(1 - c^2)
__device__ __forceinline__ void update_elecv(Photon &p) {
	const double cosalfa = p.dx * p.ex + p.dy * p.ey + p.dz * p.ez;
	const double sinalfa = sqrt(fmax(0.0, 1.0 - cosalfa * cosalfa));
	const double c_ae = 1.0 / sinalfa, c_be = -c_ae * cosalfa;
	p.ex = c_ae * p.ex + c_be * p.dx;
	p.ey = c_ae * p.ey + c_be * p.dy;
	p.ez = c_ae * p.ez + c_be * p.dz;
	normalize3(p.ex, p.ey, p.ez);
}
// phi0 of the electric vector in the photon frame (:2055-2066): returned as (sin phi0, cos phi0); the reference takes
// phi0 = acos(cos phi0) and flips its sign when the sine of the projection is positive
__device__ __forceinline__ void elec_phi0(const Photon &p, double &s0, double &c0) {
	double sth, cth, sph, cph;
	frame_trig(p, sth, cth, sph, cph);
	double cosphi0 = p.ex * (cph * cth) + p.ey * (cth * sph) + p.ez * (-sth);
	const double sinphi0 = p.ex * sph + p.ey * (-cph) + p.ez * 0.0;
	if (fabs(cosphi0) > 1.0) cosphi0 = cosphi0 > 0 ? 1.0 : -1.0;
	const double mag = sqrt(fmax(0.0, 1.0 - cosphi0 * cosphi0));   // sin(acos(c)) >= 0
	c0 = cosphi0;
	s0 = sinphi0 > 0.0 ? -mag : mag;
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
	const double *f = P.ff + (size_t)zi * P.n_q + qi;
	pf.F0 = f[0]; pf.F1 = f[1];   // (the scattering function is gathered next to the energy bracket: four registers less across the Doppler loop)
}

__device__ __forceinline__ double compton_energy(const XmbHistParams &P, int zi, double E0, double c_lamb0, double sth2, uint64_t g, int order,
                                                 int stage, int elem, bool varred, const ComptonPrefetch *pf = nullptr) {
	const double cc = 1.2399E-6, c0 = 4.85E-12, c1 = 1.456E-2;
	const double *icdf = P.cp_icdf + (size_t)zi * P.n_cp;
	const double shift = c0 * sth2 * sth2, slope = c1 * c_lamb0 * sth2;
	// The acceptance test of a trial is "energy = (cc / 1000) / c_lamb <= E0".  c_lamb0 equals (cc / 1000) / E0 to a few
	// ulp, so outside a 1e-13 relative band around c_lamb0 the comparison of the wavelengths decides it and the division
	// (in the dependent chain of every retry) is only needed for the accepted trial; inside the band the quotient itself
	// is compared -- the decisions, and the returned energy, are those of the quotient test bit for bit.
	const double c_hi = c_lamb0 * (1.0 + 1e-13), c_lo = c_lamb0 * (1.0 - 1e-13);
	double c_lamb = c_lamb0;
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
			c_lamb = c_lamb0 + (shift - slope * pz);
			bool accept = c_lamb >= c_hi;
			if (!accept && !(c_lamb > 0.0 && c_lamb <= c_lo)) accept = (cc / 1000.0) / c_lamb <= E0;
			if (accept || tries == (varred ? 100 : 500)) { done = true; break; }
			tries++;
		}
		if (done) break;
	}
	return (cc / 1000.0) / c_lamb;
}

// xmi_get_solid_angle (src/xmi_solid_angle_f.F90:712-801).  A point beyond the last r or theta of the grid (findpos
// = -1, src/xmi_aux_f.F90:1305-1335; a value below the first r extrapolates from the first cell, as findpos returns 1)
// is reported through `offgrid` with its (r, theta): the caller runs the reference's on-the-fly Monte Carlo (:783-789).
__device__ __forceinline__ double get_solid_angle(const XmbHistParams &P, const Photon &p, bool &offgrid, double &r_og, double &theta_og) {
	double vx = p.cx - P.p_window[0], vy = p.cy - P.p_window[1], vz = p.cz - P.p_window[2];
	const double r = sqrt(vx * vx + vy * vy + vz * vz);
	normalize3(vx, vy, vz);
	double temp_theta = acos(vx * P.n_detector[0] + vy * P.n_detector[1] + vz * P.n_detector[2]);
	if (temp_theta > M_PI / 2.0) temp_theta = M_PI - temp_theta;
	const double theta = (M_PI / 2.0) - temp_theta;
	const double *R = P.sa_r_vals, *Th = P.sa_t_vals;
	if (theta < Th[0]) return 0.0;
	if (r > R[P.sa_nr - 1] || theta > Th[P.sa_nt - 1]) { offgrid = true; r_og = r; theta_og = theta; return 0.0; }
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
for (int i = 0; i < nL; i++) mus[i * T] = mu_lerp(P, np, i);
	} else {
		hor_ver_ratio = S.hor_ver_ratio;
		p.weight = S.weight_rel;
		if (S.distribution_type == XMB_DISCRETE_GAUSSIAN) p.energy = ran_gaussian(rng, S.scale_parameter) + S.energy;
		else if (S.distribution_type == XMB_DISCRETE_LORENTZIAN) p.energy = S.scale_parameter * tan(M_PI * rng.uniform()) + S.energy;
		else p.energy = S.energy;
		if (p.energy <= ENERGY_THRESHOLD || p.energy > ENERGY_MAX) { p.alive = false; return; }
		const NodePos np = node_find(P, p.energy);
		XMB_UNROLL_NL
for (int i = 0; i < nL; i++) mus[i * T] = mu_lerp(P, np, i);
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
	bool horizontal;
	if (S.is_cont) horizontal = rng.uniform() <= hor_ver_ratio;
	else horizontal = (double)(j + 1) <= hor_ver_ratio;
	if (horizontal) { p.ex = 0.0; p.ey = 1.0; p.ez = 0.0; } else { p.ex = 1.0; p.ey = 0.0; p.ez = 0.0; }
	const double cosalfa = p.ex * p.dx + p.ey * p.dy + p.ez * p.dz;
	const double c_ae = 1.0 / sqrt(fmax(0.0, 1.0 - cosalfa * cosalfa)), c_be = -c_ae * cosalfa;   // sin(acos(c))
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
static __device__ double adv_energy_from_q(double e0, double Q, double theta) {
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
static __device__ double adv_shell_cdf(const XmbHistParams &P, int r, double energy, double theta) {
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
static __device__ double adv_sample_q(const XmbHistParams &P, int r, double cdf) {
	const double *qinv = P.adv_qinv + (size_t)r * P.n_cp;
	const double dc = 0.5 / (P.n_cp - 1.0), cp = cdf < 0.5 ? 0.5 - cdf : cdf - 0.5;
	const int pos = min((int)(cp / dc), P.n_cp - 2);
	const double q = qinv[pos] + (qinv[pos + 1] - qinv[pos]) * (cp - dc * pos) / dc;
	return cdf < 0.5 ? -q : q;
}
// xmi_update_photon_energy_compton (:4785-4983): two draws {subshell, Q}
static __device__ double compton_energy_adv(const XmbHistParams &P, int zi, double E0, double theta_i, double u_shell, double u_q) {
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
// Branch-free: j = rint(y) by the 1.5 * 2^52 shift (two DADD instead of FRND + F2I on the quarter-rate pipe; same
// round-to-nearest-even integer for 0 <= y < 2^31), the out-of-range case (exp(-693) = 1e-301: below anything a deposit
// can represent) is a final select, the polynomial coefficients are constant-bank operands of the DFMAs (as immediates
// each costs two UMOV per evaluation at the kernel's register cap), the table is read through its shared-window address.
static __constant__ double d_exp_poly[3] = {1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};
__device__ __forceinline__ double exp_neg(double t, unsigned tab_s32) {
	const double y = t * (64.0 * 1.4426950408889634074);
	const double shifter = 6755399441055744.0;
	const double tj = y + shifter;
	const int j = __double2loint(tj);
	const double jf = tj - shifter;
	const double x = (jf - y) * (0.69314718055994530942 / 64.0);
	double pl = fma(x, d_exp_poly[0], d_exp_poly[1]);
	pl = fma(pl, x, d_exp_poly[2]);
	pl = fma(pl, x, 0.5);
	pl = fma(pl, x, 1.0);
	pl = fma(pl, x, 1.0);
	double tabv;
	asm("ld.shared.f64 %0, [%1];" : "=d"(tabv) : "r"(tab_s32 + (((unsigned)j & 63u) << 3)));
	const double scale = __hiloint2double((1023 - (j >> 6)) << 20, 0);   // 2^(-(j >> 6)), j >> 6 <= 1000
	const double r = pl * tabv * scale;
	return y < 64.0 * 1000.0 ? r : 0.0;
}

// exp(-t) through the SFU in single precision (relative error ~2e-7): for attenuation factors of forced-detection deposits
__device__ __forceinline__ double exp_neg_f32(double t) {
	float ex;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((float)(t * -1.4426950408889634074)));
	return (double)ex;
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
                                                   int &out_shell, unsigned conv_mask = 0u, const NodePos *ep_known = nullptr) {
	const int nL = NL > 0 ? NL : P.nL;
	out_line = 0;
	out_shell = -1;
	const XmbLayerDev lay = P.layers[p.layer];
	const NodePos ep = ep_known ? *ep_known : node_find(P, p.energy);   // the history kernel has looked the photon energy up already
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
	// the scatter: polar angle (sti, cti) and azimuth (srot, crot) of the new direction in the frame of the old one
	double sti = 0.0, cti = 1.0, srot = 0.0, crot = 1.0, spi = 0.0, cpi = 1.0;
	bool rotate = false, new_energy = false;
	if (is_rayl || is_compt) {
		out_type = is_rayl ? 1 : 2;
		// Rayleigh (:1986-2101) / Compton (:2103-2229): theta from the element's inverse CDF, phi from the polarisation table
		const double *icdf = (is_rayl ? P.rayl_icdf : P.compt_icdf) + (size_t)zi * P.n_icdf_E * P.n_icdf_R;
		const double theta_i = bilinear(icdf, P.n_icdf_R, P.icdf_E, P.n_icdf_E, P.icdf_R, p.energy, s0);
		sincos(theta_i, &sti, &cti);
		double tt = sti * sti;
		if (is_rayl) tt = tt / (4.0 - 2.0 * tt);
		else {
			const double K0K = 1.0 + p.energy * (1.0 - cti) / XMI_MEC2;
			tt = tt / (K0K + (1.0 / K0K) - tt) / 2.0;
		}
		const double phi_i = bilinear(P.phi_icdf, P.n_icdf_R, P.phi_T, P.n_phi_T, P.icdf_R, tt, s1);
		sincos(phi_i, &spi, &cpi);
		{   // azimuth phi_i + phi0 by the addition theorems
			double s0e, c0e;
			elec_phi0(p, s0e, c0e);
			srot = spi * c0e + cpi * s0e; crot = cpi * c0e - spi * s0e;
		}
		rotate = true;
		if (is_compt) {
			if (ADV) {
				const uint4 w = draw_block(P.seed, g, order, 3, 0, 0);
				p.energy = compton_energy_adv(P, zi, p.energy, theta_i, xmb_u01(w.x), xmb_u01(w.y));
			} else
				p.energy = compton_energy(P, zi, p.energy, 1.2399E-6 / (p.energy * 1000.0), sqrt(0.5 * (1.0 - cti)), g, order, 3, 0, false);   // sin(theta_i / 2)
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
			// (no early return: every lane reaches the meeting point in front of the common tail)
			const bool auger = MODE == 2 && xs.uniform() > P.fluor_yield_corr[zi * 9 + shell];
			if (auger) { p.energy = 0.0; out_type = 4; }
			// Coster-Kronig (:5184-5323)
			const double *ck = P.cos_kron + zi * XMB_N_CK;
			while (!auger && (shell == 1 || shell == 2 || (shell >= 4 && shell <= 7))) {
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
			for (int l = lf; l <= ll && !auger; l++) { sl += P.rad_rate[(size_t)zi * 384 + l]; if (rl < sl) { line = l; break; } }
			if (auger) { }
			else if (!line) p.energy = 0.0;
			else {
				out_line = line;
				out_shell = shell;
				p.energy = P.line_energy[(size_t)zi * 384 + line];
				new_energy = true;
				cti = -2.0 * s2 + 1.0;                         // theta_i = acos(1 - 2 s2), isotropic
				sti = 2.0 * sqrt(fmax(0.0, s2 * (1.0 - s2)));
				sincospi(2.0 * u_phi, &srot, &crot);
				rotate = true;
			}
		}
	}
	// ---- common tail: attenuation coefficients at the new energy, rotation of direction and polarisation ----------
	// conv_mask = the lanes of the warp that made this call (history kernel): they meet here.  Without it the compiler
	// joins the interaction branches only at the end of the function and every group of lanes runs the tail -- four
	// sincos, acos, atan2 -- on its own: 2.7 passes per warp at 11.5 of 32 lanes (profiles/r1_history_kernel_v12_*).
	if (conv_mask) __syncwarp(conv_mask);
	if (new_energy) {
		const NodePos cp = node_find(P, p.energy);
		XMB_UNROLL_NL
for (int i = 0; i < nL; i++) mus[i * T] = mu_lerp(P, cp, i);
	}
	if (rotate) {
		update_dirv(p, sti, cti, srot, crot);
		update_elecv(p);
		if (is_compt) {
			// depolarisation of the Compton-scattered photon (:2201-2211)
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

