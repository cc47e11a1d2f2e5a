// history.cu -- photon histories with forced detection on the GPU (north_star items 1 and 2): the kernel, the device
// layouts and the host driver behind xmb_main_msim.
//
// Replaces the OpenMP photon loops of xmi_main_msim (src/xmi_main.F90:280-867) and everything they call per photon:
// source sampling (:957-1186), xmi_simulate_photon's variance-reduction branch (:1188-1685), the interactions
// (:1986-2411, :4985-5437; shared device code in history_device.cuh) and the forced-detection scoring
// xmi_variance_reduction (src/xmi_variance_reduction.F90:29-1101) with the solid-angle lookup
// (src/xmi_solid_angle_f.F90:712-801).
//
// Mapping: one thread = one history; a 1024-thread CTA per SM walks one interaction of a batch of
// photons in lock step through four phases (transport + detector geometry | scatter deposits per element | line
// deposits | selection + scattering); survivors are compacted into per-order queues so that every batch has all lanes
// alive and one interaction order.  Deposits are 2^-56 fixed-point integers staged in shared memory and flushed into
// 128-bit global accumulators: totals do not depend on scheduling, launch shape or GPU count, bit for bit.
// XRF deposits only touch the per-line history slot; the channel spectrum is rebuilt from those slots in the epilogue.
// Random numbers: Philox4x32-10 keyed by the run seed; every draw of photon g has a fixed counter address
// (g, interaction order, stage, element, block), see draw_block().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cmath>
#include "history_device.cuh"
#include "solid_angle_device.cuh"
#include "device_tables.h"

#ifndef XMB_SA_ROUND
#define XMB_SA_ROUND 32
#endif
#ifndef XMB_BUCKET_FINE
#define XMB_BUCKET_FINE 16
#endif
#ifndef XMB_PUSH_ATOMIC
#define XMB_PUSH_ATOMIC 1
#endif
#ifndef XMB_SKIP_LAST
#define XMB_SKIP_LAST 1
#endif
#ifndef XMB_ECLS_PER_LAYER
#define XMB_ECLS_PER_LAYER 1
#endif
#ifndef XMB_CONV_TAIL
#define XMB_CONV_TAIL 1
#endif
// -DXMB_PHASE_CLOCKS=1 (experiment builds): lane 0 of every warp accumulates the SM clock per phase of the batch loop,
// the sums land in counters[40 + phase] (XMB_PHASES=1 prints them after a run)
// queue traffic with the streaming (evict-first) cache operator, or plain
#ifndef XMB_QUEUE_STREAMING
#define XMB_QUEUE_STREAMING 0
#endif
#if XMB_QUEUE_STREAMING
#define XMB_QST(p, v) __stcs((p), (v))
#define XMB_QLD(p) __ldcs(p)
#else
#define XMB_QST(p, v) (*(p) = (v))
#define XMB_QLD(p) (*(p))
#endif
#ifndef XMB_FINE_ENERGY_KEY
#define XMB_FINE_ENERGY_KEY 1   // batches sorted by 255 uniform energy buckets (0: by the <= 31 shell-edge classes of v14)
#endif
#define XMB_FINE_KEYS 255
#ifndef XMB_PREFETCH_EARLY
#define XMB_PREFETCH_EARLY 0
#endif
// the layers' mu travel through the compaction queues with the photon when there are one or two layers; with more they are
// looked up again when the photon is taken from the queue (10-layer sample 65.7 -> 64.2 ms; two layers: 145.6 -> 146.1, kept queued)
#ifndef XMB_QUEUE_MUS_FOR
#define XMB_QUEUE_MUS_FOR(nl) ((nl) <= 2)
#endif
#ifndef XMB_COMPTON_EXP_F32
#define XMB_COMPTON_EXP_F32 0
#endif
#ifndef XMB_PHASE_CLOCKS
#define XMB_PHASE_CLOCKS 0
#endif
#if XMB_PHASE_CLOCKS
// a barrier whose result is consumed: the clock read behind it cannot be issued before the barrier completes (BAR.SYNC.DEFER_BLOCKING lets a warp run on)
#define XMB_SYNC_TIMED() do { if (__syncthreads_or(0)) ph_last_ = 0; } while (0)
#define XMB_PHW(i) do { __syncwarp(); XMB_PH(i); } while (0)
#define XMB_PH(i) do { if (lane == 0) { const long long c_ = clock64(); ph_[i] += c_ - ph_last_; ph_last_ = c_; } } while (0)
#else
#define XMB_SYNC_TIMED() __syncthreads()
#define XMB_PH(i) do { } while (0)
#define XMB_PHW(i) do { } while (0)
#endif
#ifndef XMB_BARRIERS
#define XMB_BARRIERS 5   // phase barriers kept: 1 before the scatter deposits, 2 before the line deposits, 4 before selection + scattering
#endif

// NL > 0: number of layers known at compile time (loops over layers fully unrolled); NL = 0: generic.
// MAXT: the largest CTA the instantiation is launched with.  One CTA per SM owns the register file, so a launch that cannot use
// 1024 threads (shared memory: 2 nL doubles per thread; 16-bit pieces: T x max nE addends) takes the instantiation compiled for
// its size and gets the registers of the absent warps: 72 per thread at 896 threads (the 10-layer sample: 67.6 -> 65.7 ms), 80 at 768.
// UNSTAGED: the channel slots are not in the shared-memory staging area (very many channels, P.stage_nch == 0): generic nL only.
template <int NL, bool ADV = false, int MAXT = HIST_THREADS, bool UNSTAGED = false>
__global__ void __launch_bounds__(MAXT, HIST_MIN_BLOCKS) xmb_history_kernel(const __grid_constant__ XmbHistParams P) {
	const int nL = NL > 0 ? NL : P.nL;
	extern __shared__ __align__(16) double smem[];
	const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
	double *mus = smem + tid;                 // mus[j*T]   : mu of layer j at the photon energy
	double *rd = smem + (size_t)nL * T + tid;   // rd[j*T]    : distances, then rho_j * d_j towards the detector
	unsigned int *stage = reinterpret_cast<unsigned int *>(smem + (size_t)2 * nL * T);   // [nch + n_hist_slots][4] pieces
	// line phase (lane = record): per-warp scratch of the photons' factors per shell group, and the staged line tiles
	const int wscr = xmb_warp_scratch_doubles(P.tile_groups);
	float *wpre_w = reinterpret_cast<float *>(smem + (size_t)2 * nL * T + 2 * ((size_t)P.stage_nch + P.n_hist_slots) + (size_t)(tid >> 5) * wscr);
	const char *sblob = reinterpret_cast<const char *>(smem + (size_t)2 * nL * T + 2 * ((size_t)P.stage_nch + P.n_hist_slots) + (size_t)(T >> 5) * wscr);
	constexpr bool P20 = false;   // every staged slot holds four 16-bit pieces (a line slot receives one per-lane sum per tile and warp)
	const uint64_t n_total = P.n_local_span;
	const uint64_t n_chunks = (n_total + T - 1) / T;
	const size_t acc_row = (size_t)P.nch + P.n_hist_slots;
	unsigned long long n_inter_local = 0;
	bool bad_fixed = false;
	__shared__ unsigned int s_layer_cnt[XMB_MAX_LAYERS];
	__shared__ double s_exp_tab[64];              // 2^(-k/64), see exp_neg()
	const unsigned stage_s32 = smem_u32(stage);   // shared-window addresses of the staging area and of the exp table:
	const unsigned tab_s32 = smem_u32(s_exp_tab); // see stage_red() / exp_neg()
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
#if !XMB_PUSH_ATOMIC
	__shared__ int s_wsum[32];
#endif
	// solid angles of interaction points beyond the grid, computed by the whole CTA, XMB_SA_ROUND points at a time
	__shared__ int s_sa_n, s_sa_hits[XMB_SA_ROUND];
	__shared__ double s_sa_pt[2 * XMB_SA_ROUND];
	__shared__ uint64_t s_sa_g[XMB_SA_ROUND];
	__shared__ SaCone s_sa_cone[XMB_SA_ROUND];
	__shared__ int s_lcnt[XMB_FINE_KEYS + 1];     // batch members per key (counting sort of a batch by layer or by energy)
	__shared__ int s_qcl[XMB_MAX_QL];             // layer_sort 2: photons waiting in the queue of (order k, layer L), [k * nL + L]
	__shared__ int s_sched[3];                    // layer_sort 2: the scheduler's choice {k, L or -1 (mixed batch), source?}
	__shared__ unsigned short s_perm[HIST_THREADS];
	if (tid < XMB_MAX_ORDERS) s_qcount[tid] = 0;
	for (int i = tid; i < XMB_MAX_QL; i += T) s_qcl[i] = 0;
	for (int i = tid; i < 4 * (P.stage_nch + P.n_hist_slots); i += T) stage[i] = 0u;
	// Line tiles of ONE layer are staged in shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, UBLKCP): the
	// layer of the batch when batches are formed per layer (the copy is issued when the batch is chosen and lands during the
	// geometry and scatter-deposit phases), else the layer with the most records, once.  Other layers are read in place.
	__shared__ __align__(8) unsigned long long s_mbar;
	const unsigned mbar_s32 = smem_u32(&s_mbar), sblob_s32 = smem_u32(sblob);
	int staged_layer = -1;
	unsigned stage_parity = 0;
	bool stage_pending = false;
	if (tid == 0) { xmb_mbar_init(mbar_s32, 1); xmb_fence_proxy_async(); }
	__syncthreads();
	auto stage_layer = [&](int L) {   // block-uniform; callers guarantee that no thread still reads the staged blob
		if (P.lblob_stage_bytes == 0) return;   // no room in shared memory (many layers): the tiles are read in place
		const unsigned bytes = (unsigned)(P.lblob_off[L + 1] - P.lblob_off[L]);
		if (tid == 0) { xmb_fence_proxy_async(); xmb_bulk_g2s(sblob_s32, P.lblob + P.lblob_off[L], bytes, mbar_s32); }
		staged_layer = L; stage_pending = true;
	};
	const int NF = XMB_STATE_FIELDS + (XMB_QUEUE_MUS_FOR(nL) ? nL : 0);
	const size_t qcap = 2 * (size_t)T;
	const bool per_layer = NL != 1 && P.layer_sort == 2;   // one queue per (order, layer): see the scheduler below
	double *qbase = P.queue + (size_t)blockIdx.x * P.n_int * (per_layer ? nL : 1) * NF * qcap;
	uint64_t next_chunk = blockIdx.x;
	// Forced interaction (src/xmi_main.F90:1229-1518): moves the photon to the point of its interaction number
	// order + 1.  Done before the photon is queued, so that the layer it will interact in is known when a batch is formed.
	auto transport = [&](Photon &p, uint64_t g, int order) {
		if (p.alive && p.energy < ENERGY_THRESHOLD) p.alive = false;
		if (p.alive) {
			const uint4 b1 = draw_block(P.seed, g, order + 1, 1, 0, 0);   // .x = path length
			// loops over the layers in the direction of flight, written with a trip count and a signed step (one loop body
			// for both directions: the lanes of a warp fly both ways after a fluorescence interaction)
			const bool up = p.dx * P.n_sample[0] + p.dy * P.n_sample[1] + p.dz * P.n_sample[2] > 0.0;
			const int step_dir = up ? 1 : -1;
			const int n_steps = up ? nL - p.layer : p.layer + 1;
			const double interactionR = xmb_u01(b1.x);
			double lx = p.cx, ly = p.cy, lz = p.cz;
			double Pabs = 0.0;
			for (int k = 0, i = p.layer; k < n_steps; k++, i += step_dir) {
				double dist;
				const XmbLayerDev &lay = P.layers[i];
				if (!step_to_plane(P, lx, ly, lz, p.dx, p.dy, p.dz, up ? lay.Z_end : lay.Z_begin, dist)) { p.alive = false; break; }
				rd[i * T] = dist;
				Pabs += mus[i * T] * lay.density * dist;
			}
			if (p.alive) {
				const double Pabs2 = -1.0 * expm1(-1.0 * Pabs);
				p.weight *= Pabs2;
				const double l1p = log1p(-1.0 * interactionR * Pabs2);
				const double negln = -1.0 * l1p;
				int my_index = p.layer, my_steps = 1;
				double my_sum = 0.0;
				for (int k = 0, i = p.layer; k < n_steps; k++, i += step_dir) {
					my_sum += mus[i * T] * P.layers[i].density * rd[i * T];
					if (my_sum > negln) { my_index = i; my_steps = k + 1; break; }
				}
				const double murho_idx = mus[my_index * T] * P.layers[my_index].density;
				double temp_sum = 0.0;
				for (int k = 0, i = p.layer; k < my_steps; k++, i += step_dir)
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
	};
	// compaction: survivors go, densely packed, to the queue of the next order
	auto push = [&](const Photon &p, uint64_t g, int order) {
		const bool surv = p.alive;
		if (per_layer) {
			// queue of (order, layer of the next interaction point): the lanes of a warp that go to the same queue take
			// consecutive places behind one shared-memory atomic of their leader
			const int myL = surv ? p.layer : -1;
			const unsigned peers = __match_any_sync(0xffffffffu, myL);
			const int leader = __ffs(peers) - 1;
			int wbase = 0;
			if (surv && lane == leader) wbase = atomicAdd(&s_qcl[order * nL + myL], __popc(peers));
			wbase = __shfl_sync(0xffffffffu, wbase, leader);
			if (surv) {
				double *q = qbase + (size_t)(order * nL + myL) * NF * qcap + wbase + __popc(peers & ((1u << lane) - 1u));
				// (streaming stores / loads: 28 GB of queue traffic per 5e7 histories must not evict the tables and the grid from L2)
				XMB_QST(&q[0 * qcap], p.cx); XMB_QST(&q[1 * qcap], p.cy); XMB_QST(&q[2 * qcap], p.cz);
				XMB_QST(&q[3 * qcap], p.dx); XMB_QST(&q[4 * qcap], p.dy); XMB_QST(&q[5 * qcap], p.dz);
				XMB_QST(&q[6 * qcap], p.ex); XMB_QST(&q[7 * qcap], p.ey); XMB_QST(&q[8 * qcap], p.ez);
				XMB_QST(&q[9 * qcap], p.energy); XMB_QST(&q[10 * qcap], p.weight);
				XMB_QST(&q[11 * qcap], __longlong_as_double((long long)g));
				XMB_QST(&q[12 * qcap], __longlong_as_double((long long)p.layer));
				if (XMB_QUEUE_MUS_FOR(nL)) {
					for (int j = 0; j < nL; j++) XMB_QST(&q[(XMB_STATE_FIELDS + j) * qcap], mus[j * T]);
				}
			}
			return;   // the caller's __syncthreads() publishes the counts
		}
		const unsigned bal = __ballot_sync(0xffffffffu, surv);
#if XMB_PUSH_ATOMIC
		// the survivors of a warp take consecutive places behind one shared-memory atomic of lane 0: no CTA-wide prefix
		// sum, no barriers (the place in the queue is irrelevant: exact sums, fixed-address random numbers)
		int have = 0;
		if (lane == 0 && bal) have = atomicAdd(&s_qcount[order], __popc(bal));
		have = __shfl_sync(0xffffffffu, have, 0);
		const int off = 0;
#else
		if (lane == 0) s_wsum[tid >> 5] = __popc(bal);
		__syncthreads();
		int off = 0, tot = 0;
		for (int w = 0; w < (T >> 5); w++) { const int c = s_wsum[w]; if (w < (tid >> 5)) off += c; tot += c; }
		const int have = s_qcount[order];
#endif
		if (surv) {
			double *q = qbase + (size_t)order * NF * qcap + have + off + __popc(bal & ((1u << lane) - 1u));
			XMB_QST(&q[0 * qcap], p.cx); XMB_QST(&q[1 * qcap], p.cy); XMB_QST(&q[2 * qcap], p.cz);
			XMB_QST(&q[3 * qcap], p.dx); XMB_QST(&q[4 * qcap], p.dy); XMB_QST(&q[5 * qcap], p.dz);
			XMB_QST(&q[6 * qcap], p.ex); XMB_QST(&q[7 * qcap], p.ey); XMB_QST(&q[8 * qcap], p.ez);
			XMB_QST(&q[9 * qcap], p.energy); XMB_QST(&q[10 * qcap], p.weight);
			XMB_QST(&q[11 * qcap], __longlong_as_double((long long)g));
			XMB_QST(&q[12 * qcap], __longlong_as_double((long long)p.layer));
			if (XMB_QUEUE_MUS_FOR(nL)) {
				XMB_UNROLL_NL
for (int j = 0; j < nL; j++) XMB_QST(&q[(XMB_STATE_FIELDS + j) * qcap], mus[j * T]);
			}
		}
#if !XMB_PUSH_ATOMIC
		__syncthreads();
		if (tid == 0) s_qcount[order] = have + tot;
#endif
	};
	// Energy class of a photon: the number of class thresholds (shell edges, build_device_tables) at or below its energy.
	// A shell group of the line phase is skipped when no lane of the warp can ionise the shell; after the first
	// interaction the photons are fluorescence lines of every energy, and a warp of one class skips the shells above it.
	auto energy_class = [&](double e) {
#if XMB_FINE_ENERGY_KEY
		// uniform energy buckets between the lowest tracked energy and the top of the table window: photons of one
		// fluorescence line (the bulk of a batch behind the first interaction) share a bucket, so a warp reads the same
		// node rows and takes the same branches; monotone in the energy, as the line phase's per-tile photon range needs
		const int c = (int)((e - ENERGY_THRESHOLD) * P.ekey_scale);
		return max(0, min(c, XMB_FINE_KEYS - 1));
#else
		int c = 0;
		for (int i = 0; i < P.n_ecls; i++) c += e >= P.ecls_thr[i] ? 1 : 0;
		return c;
#endif
	};
	// Counting sort of a batch by a small key (0 .. nK-1; -1: idle lane, sorted last): returns the batch position whose
	// photon this thread takes.  Four CTA barriers; the first warp turns the counts into start positions.
	auto sort_batch = [&](int key, int nK) {
		if (key < 0) key = nK;
		for (int i = tid; i <= nK; i += T) s_lcnt[i] = 0;
		__syncthreads();
		const unsigned peers = __match_any_sync(0xffffffffu, key);
		const int leader = __ffs(peers) - 1;
		int wbase = 0;
		if (lane == leader) wbase = atomicAdd(&s_lcnt[key], __popc(peers));
		wbase = __shfl_sync(0xffffffffu, wbase, leader);
		const int rank = wbase + __popc(peers & ((1u << lane) - 1u));
		__syncthreads();
		if (tid < 32) {   // exclusive scan of the nK + 1 counts, in place
			const int per = (nK + 1 + 31) >> 5, i0 = lane * per, i1 = min(i0 + per, nK + 1);
			int sum = 0;
			for (int i = i0; i < i1; i++) sum += s_lcnt[i];
			int incl = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
			int run = incl - sum;
			for (int i = i0; i < i1; i++) { const int c = s_lcnt[i]; s_lcnt[i] = run; run += c; }
		}
		__syncthreads();
		s_perm[s_lcnt[key] + rank] = (unsigned short)tid;
		__syncthreads();
		return (int)s_perm[tid];
	};
	if (!per_layer) stage_layer(P.lblob_main_layer);
#if XMB_PHASE_CLOCKS
	long long ph_[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	long long ph_last_ = clock64();
#endif
	for (;;) {
		// ---- scheduler (block-uniform) ---------------------------------------------------------------
		int k = -1, Lsel = -1;
		bool from_source = false;
		if (per_layer) {
			// Batches of ONE layer: every warp of the CTA then runs the element and line loops of the same layer, and the
			// phases end together.  With mixed (sorted) batches the warps of the layer with the most lines set the
			// duration of every batch: on the 10-layer configuration the warps spent 15.8 issue slots at the phase
			// barriers per instruction issued (profiles/r1_history_kernel_v10_synthetic10_*).  The deepest order with a
			// full layer queue runs first; else fresh source photons; at the end the remainders drain as mixed batches
			// (the queues of one order concatenated: sorted by layer by construction).
			if (tid == 0) {
				int kq = -1, Lq = -1, src = 0;
				for (int kk = P.n_int - 1; kk >= 0 && kq < 0; kk--)
					for (int l = 0; l < nL; l++) if (s_qcl[kk * nL + l] >= T) { kq = kk; Lq = l; break; }
				if (kq < 0) {
					if (next_chunk < n_chunks) src = 1;
					else
						for (int kk = P.n_int - 1; kk >= 0 && kq < 0; kk--)
							for (int l = 0; l < nL; l++) if (s_qcl[kk * nL + l] > 0) { kq = kk; break; }
				}
				s_sched[0] = kq; s_sched[1] = Lq; s_sched[2] = src;
			}
			__syncthreads();
			k = s_sched[0]; Lsel = s_sched[1]; from_source = s_sched[2] != 0;
			if (k < 0 && !from_source) break;
			// a batch of one layer: its line tiles are copied to shared memory while the batch runs its first phases (every
			// thread is behind the barrier that ended the previous batch: nobody reads the blob staged before)
			if (!from_source && Lsel >= 0 && Lsel != staged_layer) stage_layer(Lsel);
		} else {
			for (int kk = P.n_int - 1; kk >= 0; kk--) if (s_qcount[kk] >= T) { k = kk; break; }
			if (k < 0) {
				if (next_chunk < n_chunks) from_source = true;
				else {
					for (int kk = P.n_int - 1; kk >= 0; kk--) if (s_qcount[kk] > 0) { k = kk; break; }
					if (k < 0) break;
				}
			}
		}
		uint64_t g = 0;
		Photon p;
		p.alive = false;
		p.layer = 0; p.n_interactions = 0; p.energy = 0.0; p.weight = 0.0;
		int order = 0;   // interactions this batch has behind it once the step is done (0: fresh source photons)
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
			transport(p, g, 0);
			if (P.layer_sort == 1 || P.layer_sort == 2) {   // first interactions are batched by layer as well: through queue 0
				push(p, g, 0);
				__syncthreads();
				continue;
			}
			order = 1;
		} else if (per_layer) {
			order = k + 1;
			// one full layer queue, or (draining) the tails of the order's queues one after the other
			int myL = -1, at = 0, taken = 0;
			if (Lsel >= 0) {
				myL = Lsel; at = s_qcl[k * nL + Lsel] - T + tid;
				if (XMB_ECLS_PER_LAYER && P.n_ecls > 0)   // a batch of one layer: its lanes sorted by energy class
					at = s_qcl[k * nL + Lsel] - T + sort_batch(energy_class(qbase[(size_t)(k * nL + Lsel) * NF * qcap + 9 * qcap + at]), XMB_FINE_ENERGY_KEY ? XMB_FINE_KEYS : P.n_ecls + 1);
			} else {
				for (int l = 0; l < nL; l++) {
					const int c = s_qcl[k * nL + l], n_l = min(c, T - taken);
					if (myL < 0 && tid < taken + n_l) { myL = l; at = c - n_l + (tid - taken); }
					taken += n_l;
				}
			}
			if (myL >= 0) {
				const double *q = qbase + (size_t)(k * nL + myL) * NF * qcap + at;
				p.cx = XMB_QLD(&q[0 * qcap]); p.cy = XMB_QLD(&q[1 * qcap]); p.cz = XMB_QLD(&q[2 * qcap]);
				p.dx = XMB_QLD(&q[3 * qcap]); p.dy = XMB_QLD(&q[4 * qcap]); p.dz = XMB_QLD(&q[5 * qcap]);
				p.ex = XMB_QLD(&q[6 * qcap]); p.ey = XMB_QLD(&q[7 * qcap]); p.ez = XMB_QLD(&q[8 * qcap]);
				p.energy = XMB_QLD(&q[9 * qcap]); p.weight = XMB_QLD(&q[10 * qcap]);
				g = (uint64_t)__double_as_longlong(XMB_QLD(&q[11 * qcap]));
				p.layer = myL;
				if (XMB_QUEUE_MUS_FOR(nL)) {
					for (int j = 0; j < nL; j++) mus[j * T] = XMB_QLD(&q[(XMB_STATE_FIELDS + j) * qcap]);
				} else {
					{ const NodePos nq = node_find(P, p.energy); for (int j = 0; j < nL; j++) mus[j * T] = mu_lerp(P, nq, j); }
				}
				p.n_interactions = order;
				p.alive = true;
			}
			__syncthreads();
			if (tid == 0) {
				if (Lsel >= 0) s_qcl[k * nL + Lsel] -= T;
				else {
					int left = T;
					for (int l = 0; l < nL; l++) { const int n_l = min(s_qcl[k * nL + l], left); s_qcl[k * nL + l] -= n_l; left -= n_l; }
				}
			}
		} else {
			const int have = s_qcount[k], n = min(T, have), base = have - n;
			order = k + 1;
			const double *qk = qbase + (size_t)k * NF * qcap + base;
			int src = tid;
			if ((NL != 1 && P.layer_sort == 1) || P.layer_sort == 3) {
				// Mode 1: the key is the layer of the interaction point (known: the photon was moved there before it was
				// queued).  Every loop of the deposit phases is per (layer, element); with warps of one layer a warp runs
				// the loops of its own layer only instead of those of every layer its lanes are in.  Mode 3: energy class.
				const bool by_energy = P.layer_sort == 3;
				int key = -1;
				if (tid < n) key = by_energy ? energy_class(qk[9 * qcap + tid]) : (int)__double_as_longlong(qk[12 * qcap + tid]);
				src = sort_batch(key, by_energy ? (XMB_FINE_ENERGY_KEY ? XMB_FINE_KEYS : P.n_ecls + 1) : nL);
			}
			if (tid < n) {
				const double *q = qk + src;
				p.cx = XMB_QLD(&q[0 * qcap]); p.cy = XMB_QLD(&q[1 * qcap]); p.cz = XMB_QLD(&q[2 * qcap]);
				p.dx = XMB_QLD(&q[3 * qcap]); p.dy = XMB_QLD(&q[4 * qcap]); p.dz = XMB_QLD(&q[5 * qcap]);
				p.ex = XMB_QLD(&q[6 * qcap]); p.ey = XMB_QLD(&q[7 * qcap]); p.ez = XMB_QLD(&q[8 * qcap]);
				p.energy = XMB_QLD(&q[9 * qcap]); p.weight = XMB_QLD(&q[10 * qcap]);
				g = (uint64_t)__double_as_longlong(XMB_QLD(&q[11 * qcap]));
				p.layer = (int)__double_as_longlong(XMB_QLD(&q[12 * qcap]));
				if (XMB_QUEUE_MUS_FOR(nL)) {
					XMB_UNROLL_NL
for (int j = 0; j < nL; j++) mus[j * T] = XMB_QLD(&q[(XMB_STATE_FIELDS + j) * qcap]);
				} else {
					{
						// mu of the layers is a function of the photon energy alone (every writer of mus[] is mu_lerp at node_find(energy)):
						// looked up again here instead of travelling through the queue -- nL doubles less per photon written and read
						// (10 of 23 on the 10-layer sample); photons of one fluorescence line share the two table rows
						const NodePos nq = node_find(P, p.energy);
						XMB_UNROLL_NL
for (int j = 0; j < nL; j++) mus[j * T] = mu_lerp(P, nq, j);
					}
				}
				p.n_interactions = order;
				p.alive = true;
			}
			__syncthreads();
			if (tid == 0) s_qcount[k] = base;
		}
		XMB_PH(0);   // scheduler + batch formation
		{
			const uint4 b0 = draw_block(P.seed, g, order, 1, 0, 0);   // {path length (used when the photon was moved), detector r, detector phi, atom}
			const int n_ia = order;   // == p.n_interactions for every live lane
			const unsigned acc_k = stage_s32;   // deposits of this batch are staged in shared memory, flushed below
			const unsigned hist0 = (unsigned)P.stage_nch;
			// channel deposits: staged like every other deposit, or -- very many channels: nch + history slots beyond the shared
			// memory of an SM -- added straight to the global accumulators of the order (low and high half of the 2^-56
			// fixed-point value into piece words 0 and 2: each takes 2^32 deposits)
			unsigned long long *const grow = P.acc + 4 * (size_t)(n_ia - 1) * acc_row;
			auto deposit_channel = [&](int ch, unsigned long long v) {
				if (!UNSTAGED) deposit_varying(acc_k, ch, v, lane);
				else if (ch >= 0 && v) { red_global_u64(grow + 4 * (size_t)ch, v & 0xFFFFFFFFULL); red_global_u64(grow + 4 * (size_t)ch + 2, v >> 32); }
			};

			// ---- forced detection (src/xmi_variance_reduction.F90:29-726) -----------------------------
			bool vr = p.alive && p.energy > ENERGY_THRESHOLD;
			double theta = 0.0, phi = 0.0, Pesc_rayl = 0.0, omega = 0.0;
			NodePos np;
			np.pos = 0; np.f = 0.0;
			int pj_lo = nL, pj_hi = -1;   // layers the path to the detector crosses (rd[] is zero outside)
			bool sa_pending = false;      // interaction point beyond the solid-angle grid
			double sa_r = 0.0, sa_theta = 0.0;
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
					pj_lo = min(p.layer, vmax); pj_hi = max(p.layer, vmax);
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
					omega = get_solid_angle(P, p, sa_pending, sa_r, sa_theta);
				}
			}
			// Points beyond the grid: the reference computes their solid angle on the spot with hits_per_single rays
			// (src/xmi_solid_angle_f.F90:783-789).  About one interaction in a thousand (air-path scatters far from the
			// window), and 5000 rays by one lane would stall the whole CTA: the CTA collects the points of the batch (32
			// per round) and all its threads share their rays (Philox address: photon id, (order << 20) | ray pair,
			// XMB_TAG_SA_FALLBACK -- fixed per photon, so the result does not depend on the batch either).
			XMB_PH(1);   // detector geometry, solid-angle lookup
			while (__syncthreads_or(sa_pending)) {
				if (tid == 0) s_sa_n = 0;
				if (tid < XMB_SA_ROUND) s_sa_hits[tid] = 0;
				__syncthreads();
				int slot = -1;
				if (sa_pending) {
					slot = atomicAdd(&s_sa_n, 1);
					if (slot < XMB_SA_ROUND) { s_sa_pt[2 * slot] = sa_r; s_sa_pt[2 * slot + 1] = sa_theta; s_sa_g[slot] = g; }
				}
				__syncthreads();
				const int npt = min(s_sa_n, XMB_SA_ROUND);
				if (tid < npt) s_sa_cone[tid] = sa_cone_setup(P.sa_det, s_sa_pt[2 * tid], s_sa_pt[2 * tid + 1]);
				__syncthreads();
				{
					const double det_r2 = P.sa_det.detector_radius * P.sa_det.detector_radius, col_r2 = P.sa_det.collimator_radius * P.sa_det.collimator_radius;
					const int n_pairs = (P.sa_hits_per_single + 1) >> 1, total = npt * n_pairs;
					for (int i = tid; i < total; i += T) {
						const int pt = i / n_pairs, pr = i - pt * n_pairs;
						const SaCone cone = s_sa_cone[pt];
						if (cone.dead) continue;
						const uint64_t gp = s_sa_g[pt];
						const uint4 rnd = xmb_philox4x32_10(make_uint4((uint32_t)gp, (uint32_t)(gp >> 32), ((uint32_t)order << 20) | (uint32_t)pr, XMB_TAG_SA_FALLBACK),
						                                    make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
						const int h = sa_pair_hits(cone, det_r2, col_r2, rnd, 2 * pr + 1 < P.sa_hits_per_single);
						if (h) atomicAdd(&s_sa_hits[pt], h);
					}
				}
				__syncthreads();
				if (slot >= 0 && slot < XMB_SA_ROUND) {
					omega = s_sa_cone[slot].dead ? 0.0 : s_sa_cone[slot].cone_sa * (double)s_sa_hits[slot] / (double)P.sa_hits_per_single;
					sa_pending = false;
					atomicAdd(&P.counters[0], 1ULL);
				}
			}
			// generic layer count: the optical-depth sums below skip the layers no lane of the warp crosses (terms that
			// are exactly zero); with a compile-time layer count the loops are unrolled over all layers
			XMB_PH(2);   // off-grid solid angles (and the barrier that closes the geometry phase)
			const int jlo = NL > 0 ? 0 : __reduce_min_sync(0xffffffffu, pj_lo);
			const int jhi = NL > 0 ? nL - 1 : __reduce_max_sync(0xffffffffu, pj_hi);
#if XMB_BARRIERS & 1
			XMB_SYNC_TIMED();   // phase: scatter deposits of every element
#endif
			XMB_PH(3);
			// warp-uniform loops over layers / elements / shells / line records
			for (int L = 0; L < nL; L++) {
				const bool mine = vr && p.layer == L;
				if (!__any_sync(0xffffffffu, mine)) continue;
				const XmbLayerDev lay = P.layers[L];
				const double inv_mu = mine ? 1.0 / mus[L * T] : 0.0;
				double qf = 0.0, sin2cos2 = 0.0, k0k = 1.0, c_lamb0 = 0.0, sth2 = 0.0, dcsp_kn = 0.0;
				int qi = 0;
				int ch_rayl = -1;
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
				// The factors of a deposit that belong to the photon are multiplied once, in front of the element loop -- the loop
				// then keeps two products alive instead of six factors (it runs at the kernel's register cap):
				//   Rayleigh deposit = [w_e N_A/A_e] F^2 * Rp,   Rp = 1/mu * Omega * r_e^2 (1 - sin^2 theta cos^2 phi) * P_esc * weight
				//   Compton  deposit = [w_e N_A/A_e] S * exp(-tau(E')) * Cp,   Cp = 1/mu * Omega * dsigma_KN * weight
				const double Rp = mine ? inv_mu * omega * (RE2 * (1.0 - sin2cos2)) * Pesc_rayl * p.weight : 0.0;
				const double Cp = mine ? inv_mu * omega * dcsp_kn * p.weight : 0.0;

				// Rayleigh deposits of all elements go to the channel of the photon's energy: summed here (integers), one
				// channel deposit after the element loop
				unsigned long long fx_rayl = 0ULL;
				for (int e = 0; e < lay.n_elements; e++) {
					const int zi = P.elem_zi[lay.elem_begin + e];
					const double wfrac = P.elem_w[lay.elem_begin + e];
					const double wa = P.elem_wa[lay.elem_begin + e];   // weight fraction * N_A / A
					const unsigned hbase = hist0 + (unsigned)P.hist_base[zi];
					// the element's random block, first inverse-CDF bracket and form factors are requested together, ahead of
					// the dependent chain Compton energy -> energy bracket -> mu rows -> exp
					ComptonPrefetch pf;
#if XMB_PREFETCH_EARLY
					if (mine && !ADV) compton_prefetch(P, zi, g, order, e, qi, pf);
#else
					if (mine && !ADV) { const double *f = P.ff + (size_t)zi * P.n_q + qi; pf.F0 = f[0]; pf.F1 = f[1]; }
#endif
					if (mine && ADV) { const double *f = P.ff + (size_t)zi * P.n_q + qi, *sfp = P.sf + (size_t)zi * P.n_q + qi; pf.F0 = f[0]; pf.F1 = f[1]; pf.S0 = sfp[0]; pf.S1 = sfp[1]; }
					// Rayleigh (:342-369)
					unsigned long long fx = 0ULL;
					double Pconv = 0.0;
					if (mine) {
						Pconv = wfrac * inv_mu;   // (shell-resolved Compton below)
						const double F = pf.F0 * (1.0 - qf) + pf.F1 * qf;
						fx = to_fixed(wa * (F * F) * Rp, P.counters);
					}
					deposit_uniform<P20>(acc_k, hbase + 0, fx, lane);
					fx_rayl += fx;
					XMB_PHW(11);   // (element phase, per element) prefetch + Rayleigh deposit
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
							int ch_c = -1;
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
										for (int j = 0; j < nL; j++) tm += mu_lerp(P, cp, j) * rd[j * T];
										fx = to_fixed(Pconv * Pdir * exp(-tm) * p.weight * shell_weight, P.counters);
										const int ch = (int)((e_c - P.zero) / P.gain);
										if (e_c >= ENERGY_THRESHOLD && ch >= 0 && ch <= P.nch - 1) ch_c = ch;
									}
								}
							}
							deposit_uniform<P20>(acc_k, hbase + 1, fx, lane);
							deposit_channel(ch_c, fx);
						}
						continue;
					}
					// Compton (xmi_compton_varred2, :949-1008)
					fx = 0ULL;
					int ch_c = -1;
#if XMB_PHASE_CLOCKS
					double e_c_keep = 0.0;
#endif
					if (mine) {
#if !XMB_PREFETCH_EARLY
						// the element's first random block and inverse-CDF bracket are requested here, behind the Rayleigh deposit: ahead
						// of it they had to survive the Rayleigh arithmetic in spilled registers (seven local stores per element)
						compton_prefetch(P, zi, g, order, e, qi, pf);
#endif
						const double e_c = compton_energy(P, zi, p.energy, c_lamb0, sth2, g, order, 2, e, true, &pf);
#if XMB_PHASE_CLOCKS
						e_c_keep = e_c;
					}
					XMB_PHW(12);   // Doppler trials
					if (mine) {
						const double e_c = e_c_keep;
#endif
						const double *sfp = P.sf + (size_t)zi * P.n_q + qi;
						const double S0 = sfp[0], S1 = sfp[1];
						const NodePos cp = node_find(P, e_c);
						double tm = 0.0;
						XMB_UNROLL_NL
for (int j = jlo; j <= jhi; j++) tm += mu_lerp(P, cp, j) * rd[j * T];
						const double S = S0 * (1.0 - qf) + S1 * qf;
						fx = to_fixed(wa * S * (XMB_COMPTON_EXP_F32 ? exp_neg_f32(tm) : exp_neg(tm, tab_s32)) * Cp, P.counters);
						const int ch = (int)((e_c - P.zero) / P.gain);
						if (e_c >= ENERGY_THRESHOLD && ch >= 0 && ch <= P.nch - 1) ch_c = ch;
					}
					XMB_PHW(13);   // bracket, mu, exp, fixed point
					deposit_uniform<P20>(acc_k, hbase + 1, fx, lane);
					deposit_channel(ch_c, fx);
					XMB_PHW(14);   // staged adds
				}
				deposit_channel(ch_rayl, fx_rayl);
			}
			XMB_PH(4);   // Rayleigh / Compton deposits per element
#if XMB_BARRIERS & 2
			__syncthreads();   // phase: fluorescence-line deposits
#endif
			// ---- fluorescence lines (src/xmi_variance_reduction.F90:391-709) -----------------------------------------
			// deposit(photon, record) = [w_layer / mu * Omega / 4 pi * weight](photon) * P_shell(photon energy)
			//                           * [w_element * yield * rate](record) * exp(-sum_j mu_j(E_line) rho_j d_j(photon))
			// In a tile a LANE owns a RECORD: the photons' factors per shell group go through a per-warp scratch, the warp walks
			// its photons, and the deposits of the warp on a record are one 64-bit integer sum in a register, staged once per
			// tile (v15: one exact warp sum -- 3 REDUX -- and one staged add per record and interaction).
			if (stage_pending) { xmb_mbar_wait(mbar_s32, stage_parity); stage_parity ^= 1u; stage_pending = false; }
			if (p.alive) np = node_find(P, p.energy);   // bracket of the photon energy for the line phase and the selection (looked up here: not alive across the element loop)
			// The attenuation factor exp(-sum_j mu_j rho_j d_j) and the products of a deposit are evaluated in SINGLE precision
			// ("geometry/angles can be fp32 after validation against the fp64 oracle"): a deposit carries a relative
			// rounding error of ~1e-7, independent from photon to photon -- five orders below the Monte Carlo error of a line --
			// and the phase leaves the half-rate fp64 pipe, which bounded it (19 fp64 operations per photon and record;
			// profiles/r2_history_kernel_v16_*).  Sums stay exact 64-bit integers, so GPU-count independence is untouched.
			{
				float *wpref = wpre_w;                                                  // [XMB_TILE_GROUPS][XMB_WPRE_STRIDE]
				// rho d of the path to the detector is dead in double once the scatter deposits are made: every photon lane
				// rewrites its slots in place as floats, which the record lanes of its warp read (rdf[2 * (j * T + q)])
				if (NL > 0) {
					XMB_UNROLL_NL
for (int j = 0; j < (NL > 0 ? NL : 1); j++) { const float f = vr ? (float)rd[j * T] : 0.f; *reinterpret_cast<float *>(&rd[j * T]) = f; }
				} else {
					for (int j = jlo; j <= jhi; j++) { const float f = vr ? (float)rd[j * T] : 0.f; *reinterpret_cast<float *>(&rd[j * T]) = f; }
				}
				__syncwarp();
				const float *rdf = reinterpret_cast<const float *>(smem + (size_t)nL * T + (tid & ~31));
				for (int L = 0; L < nL; L++) {
					const bool mine = vr && p.layer == L;
					if (!__any_sync(0xffffffffu, mine)) continue;
					const char *blob = L == staged_layer ? sblob : P.lblob + P.lblob_off[L];
					const int4 bh = *reinterpret_cast<const int4 *>(blob);   // {n_tiles, n_groups, lanes_off, tile_bytes}
					// (after a fluorescence interaction the photon energy is a node of the grid: the reference's precalc_xrf_cs)
					const double commonp = mine ? (1.0 / mus[L * T]) * (omega / 4.0 / M_PI) * p.weight : 0.0;
					const double e_mine = mine ? p.energy : -1.0;
					const XmbLineGroup *G = reinterpret_cast<const XmbLineGroup *>(blob + 16 + 16 * bh.x);
					for (int t = 0; t < bh.x; t++) {
						const XmbLineTile th = *reinterpret_cast<const XmbLineTile *>(blob + 16 + 16 * t);
						const unsigned pmask = __ballot_sync(0xffffffffu, e_mine >= th.min_edge);
						if (!pmask) continue;   // no photon of the warp can ionise a shell of this tile
						for (int gi = 0; gi < th.n_groups; gi++) {
							const XmbLineGroup g = G[th.g_begin + gi];
							double pre = 0.0;
							if (e_mine >= th.min_edge && e_mine >= g.edgeK) pre = commonp * row_lerp(P, np, g.row_off);
							bad_fixed |= !(pre * P.rec_wy_max < 1.4e17);   // a deposit is < 2^57: 32 of them fit the lane's 64-bit sum
							wpref[gi * XMB_WPRE_STRIDE + lane] = (float)pre;
						}
						__syncwarp();
						const char *lt = blob + bh.z + (size_t)t * bh.w;
						const float wyf = reinterpret_cast<const float *>(lt)[lane];
						const float *mulf = reinterpret_cast<const float *>(lt) + 32 + lane;
						const int2 sg = make_int2(reinterpret_cast<const int *>(lt + 128 * (1 + nL))[lane], reinterpret_cast<const int *>(lt + 128 * (2 + nL))[lane]);   // mulf[j * 32] = -mu_j log2(e)
						const float *prow = wpref + sg.y * XMB_WPRE_STRIDE;
						float muv[NL > 0 ? NL : 1];
						if (NL > 0) {
							XMB_UNROLL_NL
for (int j = 0; j < (NL > 0 ? NL : 1); j++) muv[j] = mulf[j * 32];
						}
						unsigned long long acc = 0ULL;
						const int p_lo = __ffs(pmask) - 1, p_hi = 31 - __clz(pmask);
XMB_UNROLL(4)
						for (int q = p_lo; q <= p_hi; q++) {
							float t2 = 0.f;
							if (NL > 0) {
								XMB_UNROLL_NL
for (int j = 0; j < (NL > 0 ? NL : 1); j++) t2 = __fmaf_rn(muv[j], rdf[2 * (j * T + q)], t2);
							} else {
								const float *mj = mulf + jlo * 32, *rj = rdf + 2 * (jlo * T + q);   // pointer walk: the strides fold into the loads
XMB_UNROLL(4)
								for (int n = jhi - jlo + 1; n > 0; n--, mj += 32, rj += 2 * T) t2 = __fmaf_rn(*mj, *rj, t2);
							}
							float ex;
							asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(t2));
							acc += __float2ull_rn(prow[q] * wyf * ex);
						}
						__syncwarp();   // the scratch is rewritten by the next tile
						if (acc) {
							const unsigned a = acc_k + (hist0 + (unsigned)sg.x) * 16u;
							const unsigned int lo = (unsigned int)acc, hi = (unsigned int)(acc >> 32);
							stage_red(a, lo & 0xFFFFu); stage_red(a + 4u, lo >> 16); stage_red(a + 8u, hi & 0xFFFFu); stage_red(a + 12u, hi >> 16);
						}
					}
				}
			}

			XMB_PH(5);   // line deposits
			// ---- atom and interaction selection, scattering (src/xmi_main.F90:1558-1652) ----------------
#if XMB_BARRIERS & 4
			XMB_SYNC_TIMED();   // phase: selection + scattering (and: every deposit of the batch is staged)
			XMB_PH(6);   // waiting for the slowest warp's deposits
			flush_staged<P20>(stage, P.acc + 4 * ((size_t)(n_ia - 1) * acc_row + (size_t)(P.nch - P.stage_nch)), P.stage_nch + P.n_hist_slots, P.nch, tid, T);
			XMB_PH(7);
#endif
			// (the interaction of the last order is scored above; what it does to the photon is never used)
			const bool do_sel = p.alive && (!XMB_SKIP_LAST || order < P.n_int);
			const unsigned sel_mask = __ballot_sync(0xffffffffu, do_sel);
			if (do_sel) {
				double we_unused = 0.0;
				int t_unused, z_unused, l_unused, s_unused;
				select_and_scatter<NL, 0, ADV>(P, p, g, order, mus, T, b0.w, we_unused, t_unused, z_unused, l_unused, s_unused, XMB_CONV_TAIL ? sel_mask : 0u, &np);
			}
			__syncwarp();
			XMB_PH(8);   // selection + scattering
		}
		// ---- move to the next interaction point and queue there --------------------------------------------
		if (order < P.n_int) {
			transport(p, g, order);
			push(p, g, order);
		}
		XMB_PH(9);   // forced move + queue
		XMB_SYNC_TIMED();
		XMB_PH(10);
#if !(XMB_BARRIERS & 4)
		// every deposit of the batch is staged (barrier above); the next batch's first deposit comes behind the barriers of
		// its formation and of the off-grid solid-angle round
		flush_staged<P20>(stage, P.acc + 4 * ((size_t)(order - 1) * acc_row + (size_t)(P.nch - P.stage_nch)), P.stage_nch + P.n_hist_slots, P.nch, tid, T);
#endif
	}
#if XMB_PHASE_CLOCKS
	if (lane == 0) for (int i = 0; i < 16; i++) atomicAdd(&P.counters[40 + i], (unsigned long long)ph_[i]);
#endif
	n_inter_local = warp_sum_u64(n_inter_local);
	if (lane == 0 && n_inter_local) atomicAdd(&P.counters[1], n_inter_local);
	if (bad_fixed) atomicAdd(&P.counters[2], 1ULL);
	__syncthreads();
	if (tid < nL && s_layer_cnt[tid]) atomicAdd(&P.counters[8 + tid], (unsigned long long)s_layer_cnt[tid]);
}

// raw (lo, hi) accumulators (brute-force kernel) -> two 48-bit-split words per slot (safe to sum over ranks in 64-bit integers)
__global__ void xmb_limbs_kernel(const unsigned long long *__restrict__ acc, unsigned long long *__restrict__ limbs, size_t n_slots) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long lo = acc[2 * i], hi = acc[2 * i + 1];   // 128-bit total
		limbs[2 * i] = lo & 0xFFFFFFFFFFFFULL;
		limbs[2 * i + 1] = (lo >> 48) | (hi << 16);
	}
}
// four-word accumulators of the history kernel (w0 + w1 2^16 + w2 2^32 + w3 2^48, see flush_staged) -> the same limbs
__global__ void xmb_limbs4_kernel(const unsigned long long *__restrict__ acc, unsigned long long *__restrict__ limbs, size_t n_slots) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned __int128 v = (unsigned __int128)acc[4 * i] + ((unsigned __int128)acc[4 * i + 1] << 16) + ((unsigned __int128)acc[4 * i + 2] << 32) +
		                            ((unsigned __int128)acc[4 * i + 3] << 48);
		limbs[2 * i] = (unsigned long long)v & 0xFFFFFFFFFFFFULL;
		limbs[2 * i + 1] = (unsigned long long)(v >> 48);
	}
}

// =====================================================================================================
// Host side: device layouts, launch, exact reduction epilogue.
// =====================================================================================================

void xmb_free_device_tables(XmbDeviceTables *dev) {
	if (!dev) return;
	int cur = 0;
	const bool on_device = dev->device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != dev->device;
	if (on_device) cudaSetDevice(dev->device);   // cudaFree wants the owning device's context current
	delete dev;
	if (on_device) cudaSetDevice(cur);
}
void xmb_free_all_device_tables(XmbHdf5F *h) {
	for (XmbDeviceTables *d : h->devs) xmb_free_device_tables(d);
	h->devs.clear();
	h->dev = nullptr;
}

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

int xmb_cascade_mode(const xmb_main_options *o) {   // src/xmi_main.F90:141-153
	return 1 + (o->use_cascade_auger ? 1 : 0) + (o->use_cascade_radiative ? 2 : 0);
}

static XmbDeviceTables *build_device_tables(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *opt) {
	const xmb_tables_host &T = h->view;
	const xmb_input &I = in->in;
	const int nL = I.composition->n_layers, nZ = T.nZ, nN = T.n_nodes;
	if (nL > XMB_MAX_LAYERS) { xmb_set_error("more than %d layers", XMB_MAX_LAYERS); return nullptr; }
	XmbDeviceTables *D = new XmbDeviceTables();
	D->cascade = xmb_cascade_mode(opt);
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
	{
		std::vector<double> mt((size_t)nN * nL);
		for (int n = 0; n < nN; n++) for (int k = 0; k < nL; k++) mt[(size_t)n * nL + k] = T.mu_layer[(size_t)k * nN + n];
		P.mu_tab = upload(D, mt.data(), mt.size(), ok);
	}
	P.n_nodes = nN;
	P.node_E = upload(D, T.node_E, nN, ok);
	P.ekey_scale = (double)XMB_FINE_KEYS / std::max(T.node_E[nN - 1] - ENERGY_THRESHOLD, 1e-3);
	{
		// Device bucket index, XMB_BUCKET_FINE x finer than the host's (whose buckets are the cells of the uniform part of
		// the node grid): bucket b covers [E0 + b dE, E0 + (b + 1) dE) and stores the last node at or below its lower
		// bound; bit 31 marks a bucket without a node inside -- its bracket is known without scanning.  Edge doublets,
		// line energies and source lines are extra nodes inside the uniform cells: with the host's cell-wide buckets every
		// lookup in such a cell scanned (2.7 % of the kernel's instructions, profiles/r1_history_kernel_v11_hot_lines.txt).
		const int fine = XMB_BUCKET_FINE;
		const int nBf = T.n_buckets * fine;
		const double inv_dE = T.bucket_inv_dE * fine;
		std::vector<int> bs((size_t)nBf + 1, 0);
		int i = 0;
		for (int b = 0; b < nBf; b++) {
			const double lo = T.bucket_E0 + b / inv_dE, hi = T.bucket_E0 + (b + 1) / inv_dE;
			while (i + 1 < nN && T.node_E[i + 1] <= lo + 1e-9) i++;
			const int start = std::min(i, std::max(nN - 2, 0));
			bs[b] = start;
			// strictly inside the bracket, with a margin far above the rounding of the device's bucket arithmetic
			if (start + 1 < nN && T.node_E[start] <= lo - 1e-9 && T.node_E[start + 1] >= hi + 1e-9) bs[b] = start | (int)0x80000000;
		}
		bs[nBf] = std::max(nN - 2, 0);
		P.n_buckets = nBf; P.bucket_E0 = T.bucket_E0; P.bucket_inv_dE = inv_dE;
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
	{
		std::vector<double> wa(elem_w.size());
		for (size_t i = 0; i < wa.size(); i++) wa[i] = elem_w[i] * (AVOGNUM / T.atomic_weight[elem_zi[i]]);
		P.elem_wa = upload(D, wa.data(), wa.size(), ok);
	}
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
	{
		// Line tiles (history.cuh): per layer the non-empty (element, shell) groups in the order of their shell edges, their
		// records cut into tiles of 32 (one record per lane), a tile spanning at most XMB_TILE_GROUPS groups.  A photon
		// below a tile's lowest edge deposits exactly zero on all of its records (no vacancy cross section below the shell's
		// own edge; the lower node of the edge doublet lies 1e-5 keV below it), so the tile is skipped for it.
		struct Grp { int zi, s, r0, r1, row_off; double wfrac, edge, edgeK; };
		const int n_sh = opt->use_M_lines ? 9 : 4;
		std::vector<char> blob;
		std::vector<std::pair<double, int>> edges;   // (shell edge, records) of every group of every layer: energy classes below
		P.rec_wy_max = 0.0;
		// (With many layers the per-thread arrays mus[nL], rd[nL] leave little shared memory and the CTA runs fewer threads: 864
		// on the 10-layer sample.  Halving the per-warp scratch -- tiles of 4 groups -- and reading the tiles in place keeps 1024
		// threads but measured 2 % slower, profiles/r2_history_ablation.txt.)
		const bool short_smem = false;
		P.tile_groups = XMB_TILE_GROUPS;
		P.lblob_stage_bytes = 16;
		int most = -1;
		P.lblob_main_layer = 0;
		for (int k = 0; k < nL; k++) {
			std::vector<Grp> grp;
			for (int e = 0; e < layers[k].n_elements; e++) {
				const int z = elem_zi[layers[k].elem_begin + e];
				for (int sh = 0; sh < n_sh; sh++) {
					Grp g;
					g.r0 = rec_begin[z * 10 + sh]; g.r1 = rec_begin[z * 10 + sh + 1];
					if (g.r0 == g.r1) continue;
					g.zi = z; g.s = sh;
					g.row_off = P.off_elem + z * XMB_ELEM_STRIDE + XMB_EO_VACANCY + sh;
					g.wfrac = elem_w[layers[k].elem_begin + e];
					g.edge = T.edge_energy[z * 9 + sh];
					g.edgeK = sh == 0 ? edgeK[z] : 0.0;
					grp.push_back(g);
				}
			}
			std::stable_sort(grp.begin(), grp.end(), [](const Grp &a, const Grp &b) { return a.edge < b.edge; });
			int n_rec_layer = 0;
			for (const Grp &g : grp) { edges.push_back(std::make_pair(g.edge, g.r1 - g.r0)); n_rec_layer += g.r1 - g.r0; }
			if (n_rec_layer > most) { most = n_rec_layer; P.lblob_main_layer = k; }
			// tiles: (group, record) pairs in order; a tile closes at 32 records or XMB_TILE_GROUPS groups
			struct Lane { int g, r; };
			std::vector<std::vector<Lane>> tiles;
			std::vector<int> tile_g0;
			for (int g = 0; g < (int)grp.size(); g++)
				for (int r = grp[g].r0; r < grp[g].r1; r++) {
					if (tiles.empty() || tiles.back().size() == 32 || g - tile_g0.back() >= P.tile_groups) { tiles.push_back({}); tile_g0.push_back(g); }
					tiles.back().push_back(Lane{g, r});
				}
			const int n_tiles = (int)tiles.size(), n_grp = (int)grp.size();
			const int tile_bytes = 128 * (3 + nL);
			const int lanes_off = 16 + 16 * n_tiles + 16 * n_grp;
			const size_t base = blob.size();
			blob.resize(base + lanes_off + (size_t)n_tiles * tile_bytes, 0);
			char *bp = blob.data() + base;
			const int hdr[4] = {n_tiles, n_grp, lanes_off, tile_bytes};
			memcpy(bp, hdr, 16);
			for (int t = 0; t < n_tiles; t++) {
				XmbLineTile th;
				th.g_begin = tile_g0[t];
				th.n_groups = tiles[t].back().g - tile_g0[t] + 1;
				th.min_edge = grp[tile_g0[t]].edge - 1.5e-5;
				memcpy(bp + 16 + 16 * t, &th, 16);
				char *lt = bp + lanes_off + (size_t)t * tile_bytes;
				float *wyf = (float *)lt, *muf = wyf + 32;
				int *slot = (int *)(lt + 128 * (1 + nL)), *gs = slot + 32;
				for (int l = 0; l < 32; l++) {
					wyf[l] = 0.f; slot[l] = 0; gs[l] = 0;                 // padding lane: deposits nothing
					for (int j = 0; j < nL; j++) muf[j * 32 + l] = 0.f;
					if (l >= (int)tiles[t].size()) continue;
					const Lane &ln = tiles[t][l];
					const double wy = grp[ln.g].wfrac * rec_yr[ln.r] * 72057594037927936.0;
					wyf[l] = (float)wy;
					P.rec_wy_max = std::max(P.rec_wy_max, wy * (1.0 + 1e-6));
					for (int j = 0; j < nL; j++) muf[j * 32 + l] = (float)(-rec_mu[(size_t)ln.r * nL + j] * 1.4426950408889634074);
					slot[l] = D->rec_slot[ln.r];
					gs[l] = ln.g - tile_g0[t];
				}
			}
			for (int g = 0; g < n_grp; g++) {
				XmbLineGroup lg;
				lg.row_off = grp[g].row_off; lg.zi = grp[g].zi; lg.edgeK = grp[g].edgeK;
				memcpy(bp + 16 + 16 * n_tiles + 16 * g, &lg, 16);
			}
			P.lblob_off[k] = (int)base;
			P.lblob_stage_bytes = std::max(P.lblob_stage_bytes, (int)(blob.size() - base));
		}
		if (short_smem || P.lblob_stage_bytes > 48 * 1024) P.lblob_stage_bytes = 0;   // very large line sets: tiles are read in place
		P.lblob_off[nL] = (int)blob.size();
		P.lblob = upload(D, blob.data(), blob.size(), ok);
		D->n_line_tiles_bytes = blob.size();
		// energy classes for batches sorted by photon energy (layer_sort 3): shell edges that split the records of all
		// groups into (nearly) equal shares
		{
			int n_all = 0;
			for (const auto &e : edges) n_all += e.second;
			std::sort(edges.begin(), edges.end());
			int max_thr = (int)(sizeof(P.ecls_thr) / sizeof(P.ecls_thr[0]));
			if (const char *e = getenv("XMB_ECLS_MAX")) max_thr = std::max(0, std::min(max_thr, atoi(e)));   // experiments
			P.n_ecls = 0;
			int cum = 0;
			for (size_t i = 0; i < edges.size() && P.n_ecls < max_thr; i++) {
				// a threshold in front of the group that crosses the next share boundary
				if (cum * (max_thr + 1) / std::max(n_all, 1) >= P.n_ecls + 1 && edges[i].first > (P.n_ecls ? P.ecls_thr[P.n_ecls - 1] : 0.0))
					P.ecls_thr[P.n_ecls++] = edges[i].first;
				cum += edges[i].second;
			}
			for (int i = P.n_ecls; i < (int)(sizeof(P.ecls_thr) / sizeof(P.ecls_thr[0])); i++) P.ecls_thr[i] = 1e300;
		}
	}
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

XmbDeviceTables *xmb_device_tables_get(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *opt) {
	int dev = 0;
	cudaGetDevice(&dev);
	for (size_t i = 0; i < h->devs.size(); i++) {
		XmbDeviceTables *D = h->devs[i];
		if (D->device != dev) continue;
		if (D->cascade == xmb_cascade_mode(opt) && D->use_M_lines == (opt->use_M_lines ? 1 : 0)) return h->dev = D;
		xmb_free_device_tables(D);          // same device, other options: rebuilt below
		h->devs.erase(h->devs.begin() + i);
		break;
	}
	h->dev = nullptr;
	XmbDeviceTables *D = build_device_tables(in, h, opt);
	if (D) { h->devs.push_back(D); h->dev = D; }
	return D;
}

// 64-bit content hash of a solid-angle grid (values + both axes), four independent multiply-xorshift lanes over the
// 8-byte words: ~2 ms for the 1024 x 1024 grid.  Decides whether the grid already in HBM may be reused.
static uint64_t sa_content_hash(const xmb_solid_angle *sa) {
	auto mix = [](uint64_t h, uint64_t v) { h ^= v; h *= 0x9E3779B97F4A7C15ULL; return h ^ (h >> 29); };
	auto hash_words = [&](const double *p, size_t n, uint64_t h0) {
		uint64_t h[4] = {h0, h0 ^ 0xA5A5A5A5A5A5A5A5ULL, h0 + 0x1234567ULL, ~h0};
		size_t i = 0;
		uint64_t v[4];
		for (; i + 4 <= n; i += 4) { memcpy(v, p + i, 32); for (int k = 0; k < 4; k++) h[k] = mix(h[k], v[k]); }
		for (; i < n; i++) { memcpy(v, p + i, 8); h[0] = mix(h[0], v[0]); }
		return mix(mix(mix(h[0], h[1]), h[2]), h[3]);
	};
	const size_t nr = (size_t)sa->grid_dims_r_n, nt = (size_t)sa->grid_dims_theta_n;
	uint64_t h = hash_words(sa->solid_angles, nr * nt, 0x584D42ULL ^ (nr << 20) ^ nt);
	h = hash_words(sa->grid_dims_r_vals, nr, h);
	return hash_words(sa->grid_dims_theta_vals, nt, h);
}

// One run of the history (or brute-force) kernel in two halves, so that a driver can have every GPU of the box working
// before it waits for any of them (multi_gpu.cu):
//   xmb_msim_launch  : device tables for the current device, solid-angle grid host->device (content-hash residency with
//                      keep_on_device), accumulators zeroed, kernel + limb conversion ENQUEUED on the device's default
//                      stream; returns without waiting.  The limbs (D->limbs, 2 x slots uint64) are complete in stream order.
//   xmb_msim_collect : waits for the kernel, reads the counters, fills ex's outputs, reports range errors.
int xmb_msim_launch(XmbInputF *in, XmbHdf5F *h, const xmb_main_options *options, const xmb_solid_angle *sa, xmb_msim_ex *ex,
                    XmbDeviceTables **D_out) {
	if (!in || !h || !in->inited || !options || !ex) { xmb_set_error("xmb_main_msim_raw: bad arguments"); return 0; }
	const bool brute = !options->use_variance_reduction;
	if (options->use_advanced_compton && !xmb_tables_enable_advanced_compton((xmb_hdf5FPtr)h)) return 0;   // builds the subshell tables once
	if (options->escape_ratios_mode) { xmb_set_error("escape_ratios_mode is not implemented on the GPU path"); return 0; }
	if (!brute && (!sa || !sa->solid_angles)) { xmb_set_error("variance reduction needs a solid-angle grid"); return 0; }
	if (xmb_cuda_device_count() < 1) { xmb_set_error("no CUDA device: xmb_main_msim has no CPU fallback"); return 0; }
	if (ex->device >= 0) XMB_CUDA_OK(cudaSetDevice(ex->device));
	int dev = 0;
	cudaGetDevice(&dev);
	XmbDeviceTables *D = xmb_device_tables_get(in, h, options);
	if (!D) return 0;
	if (D_out) *D_out = D;
	XmbHistParams P = D->P;
	// solid-angle grid: an argument of the call -> copied host->device every call; with keep_on_device the copy is
	// skipped when the grid in HBM has the same CONTENT (dimensions + a 64-bit hash of values and axes: a caller may
	// refill the same host buffer, and a new buffer may land on the address of a freed one)
	const size_t nsa = brute ? 0 : (size_t)sa->grid_dims_r_n * sa->grid_dims_theta_n;
	if (!brute) {
		const size_t nr_ = (size_t)sa->grid_dims_r_n, nt_ = (size_t)sa->grid_dims_theta_n;
		if (nr_ < 2 || nt_ < 2 || !sa->grid_dims_r_vals || !sa->grid_dims_theta_vals) { xmb_set_error("solid-angle grid needs two axes of at least two values"); return 0; }
		bool fresh = false;
		if (D->sa_cap < nsa || !D->sa_grid) { cudaFree(D->sa_grid); D->sa_grid = nullptr; XMB_CUDA_OK(cudaMalloc(&D->sa_grid, sizeof(double) * nsa)); D->sa_cap = nsa; fresh = true; }
		if (D->sa_r_cap < nr_ || !D->sa_r) { cudaFree(D->sa_r); D->sa_r = nullptr; XMB_CUDA_OK(cudaMalloc(&D->sa_r, sizeof(double) * nr_)); D->sa_r_cap = nr_; fresh = true; }
		if (D->sa_t_cap < nt_ || !D->sa_t) { cudaFree(D->sa_t); D->sa_t = nullptr; XMB_CUDA_OK(cudaMalloc(&D->sa_t, sizeof(double) * nt_)); D->sa_t_cap = nt_; fresh = true; }
		uint64_t hash = 0;
		bool resident = false;
		if (ex->keep_on_device) {
			hash = sa_content_hash(sa);
			resident = !fresh && D->sa_valid && D->sa_hash == hash && D->sa_nr == nr_ && D->sa_nt == nt_;
		}
		if (!resident) {
			D->sa_valid = false;
			XMB_CUDA_OK(cudaMemcpy(D->sa_grid, sa->solid_angles, sizeof(double) * nsa, cudaMemcpyHostToDevice));
			XMB_CUDA_OK(cudaMemcpy(D->sa_r, sa->grid_dims_r_vals, sizeof(double) * nr_, cudaMemcpyHostToDevice));
			XMB_CUDA_OK(cudaMemcpy(D->sa_t, sa->grid_dims_theta_vals, sizeof(double) * nt_, cudaMemcpyHostToDevice));
			D->sa_nr = nr_; D->sa_nt = nt_;
			if (ex->keep_on_device) { D->sa_hash = hash; D->sa_valid = true; }
		}
	}
	P.sa_grid = D->sa_grid; P.sa_r_vals = D->sa_r; P.sa_t_vals = D->sa_t;
	if (!brute) { P.sa_nr = (int)sa->grid_dims_r_n; P.sa_nt = (int)sa->grid_dims_theta_n; }
	// batches sorted by layer pay when photons interact in several layers with different element lists; with one or two
	// layers (a sample behind an air gap) nearly every interaction is in the same layer and the sort is skipped
	// 2: one queue per (order, layer); 1: one queue per order, batches sorted by layer; 3: one queue per order, batches
	// sorted by energy class; 0: batches as queued
	P.layer_sort = P.nL >= 3 ? 2 : (P.n_ecls > 0 ? 3 : 0);
	if (const char *e = getenv("XMB_LAYER_SORT")) P.layer_sort = std::max(0, std::min(3, atoi(e)));   // experiments / tests: force a mode
	if (P.layer_sort == 3 && P.n_ecls == 0) P.layer_sort = 0;
	if (P.layer_sort == 2 && (P.nL < 2 || P.n_int * P.nL > XMB_MAX_QL)) P.layer_sort = 1;
	if (P.layer_sort == 1 && P.nL < 2) P.layer_sort = 0;
	P.sa_det.collimator_present = in->der.collimator_present; P.sa_det.detector_radius = in->der.detector_radius;
	P.sa_det.collimator_radius = in->der.collimator_radius; P.sa_det.collimator_height = in->der.collimator_height;
	P.sa_hits_per_single = (int)std::min<long>(xmb_get_hits_per_single(), 1L << 20);
	// accumulators: one row per interaction order (brute force: rows 0..n_int, row = interactions before detection)
	const size_t slots = (size_t)(P.n_int + (brute ? 1 : 0)) * ((size_t)P.nch + P.n_hist_slots);
	if (D->acc_slots != slots) {
		cudaFree(D->acc); cudaFree(D->limbs); cudaFree(D->counters);
		D->acc = D->limbs = D->counters = nullptr; D->acc_slots = 0;
		XMB_CUDA_OK(cudaMalloc(&D->acc, sizeof(unsigned long long) * 4 * slots));   // history kernel: 4 words per slot; brute force: (lo, hi)
		XMB_CUDA_OK(cudaMalloc(&D->limbs, sizeof(unsigned long long) * 2 * slots));
		XMB_CUDA_OK(cudaMalloc(&D->counters, sizeof(unsigned long long) * XMB_N_COUNTERS));
		D->acc_slots = slots;
	}
	XMB_CUDA_OK(cudaMemsetAsync(D->acc, 0, sizeof(unsigned long long) * (brute ? 2 : 4) * slots));
	XMB_CUDA_OK(cudaMemsetAsync(D->counters, 0, sizeof(unsigned long long) * XMB_N_COUNTERS));
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
	if (!D->ev0) { XMB_CUDA_OK(cudaEventCreate(&D->ev0)); XMB_CUDA_OK(cudaEventCreate(&D->ev1)); }
	D->run_brute = brute; D->run_slots = slots; D->run_launches = 1;
	int sms = 148, occ = 1;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (brute) {
		XmbBruteParams B{};
		B.use_auger = options->use_cascade_auger ? 1 : 0; B.use_rad = options->use_cascade_radiative ? 1 : 0;
		B.collimator_present = in->der.collimator_present; B.collimator_height = in->der.collimator_height;
		B.collimator_radius = in->der.collimator_radius; B.half_apex = in->der.half_apex;
		B.vertex_x = in->der.vertex[0]; B.vertex_y = in->der.vertex[1]; B.vertex_z = in->der.vertex[2];
		B.line_slot = D->line_slot; B.auger_rate = D->auger_rate;
		cudaEventRecord(D->ev0);
		XMB_CUDA_OK(xmb_brute_launch(P, B, options->use_advanced_compton != 0, sms, ex->n_histories));
		cudaEventRecord(D->ev1);
		if (ex->n_histories > 0) D->run_launches++;
		xmb_limbs_kernel<<<sms, 256>>>(D->acc, D->limbs, slots);
		XMB_CUDA_OK(cudaGetLastError());
		return 1;
	}
	// threads per CTA: as many as the per-thread shared arrays (2 nL doubles) allow within 200 KB
	int threads = HIST_THREADS;
	// staging area: channels + history slots; beyond 160 KB (more than ~10 000 channels) the history slots alone, and the
	// Rayleigh / Compton channel deposits go straight to the global accumulators (slower: the peak channels serialise in L2)
	P.stage_nch = P.nch;
	if (const char *e = getenv("XMB_STAGE_CHANNELS")) { if (atoi(e) == 0) P.stage_nch = 0; }   // tests: force the unstaged path
	if (sizeof(unsigned long long) * 2 * ((size_t)P.nch + P.n_hist_slots) > 160 * 1024) P.stage_nch = 0;
	const size_t stage_bytes = sizeof(unsigned long long) * 2 * ((size_t)P.stage_nch + P.n_hist_slots);
	if (stage_bytes > 160 * 1024) { xmb_set_error("the history slots of the sample's lines do not fit the shared-memory staging area"); return 0; }
	// + the line phase: XMB_TILE_GROUPS x XMB_WPRE_STRIDE doubles of scratch per warp and one staged line-tile blob
	const size_t fixed_bytes = stage_bytes + (size_t)P.lblob_stage_bytes;
	auto per_cta = [&](int t) { return fixed_bytes + sizeof(double) * 2 * P.nL * t + sizeof(double) * xmb_warp_scratch_doubles(P.tile_groups) * (t / 32); };
	if (fixed_bytes > 180 * 1024) { xmb_set_error("channels, history slots and line tiles do not fit the shared memory of an SM"); return 0; }
	while (threads > 64 && per_cta(threads) > 216 * 1024) threads -= 32;
	// a staged 16-bit piece holds < 2^16 per addend and the word 2^32: at most 2^16 addends per slot and batch; a photon
	// adds to a channel slot once per element (Compton) plus once for the summed Rayleigh deposits
	const size_t per_photon = ((size_t)std::max(1, D->max_nE) + 1) * (options->use_advanced_compton ? 32 : 1);   // + one addend per subshell
	while (threads > 64 && (size_t)threads * per_photon > 60000) threads -= 32;
	// experiments / tests: force a smaller CTA (the sums do not depend on the launch shape: test_history_gpu.py)
	if (const char *e = getenv("XMB_HIST_THREADS")) { const int t = atoi(e) & ~31; if (t >= 32 && t < threads) threads = t; }
	const size_t smem = per_cta(threads);
	void (*kernel)(const XmbHistParams) = P.nL == 1 ? xmb_history_kernel<1> : P.nL == 2 ? xmb_history_kernel<2> : P.nL == 3 ? xmb_history_kernel<3>
	                                     : P.nL == 4 ? xmb_history_kernel<4> : xmb_history_kernel<0>;
	if (P.nL > 4 && threads <= 896) kernel = threads <= 640 ? xmb_history_kernel<0, false, 640> : threads <= 768 ? xmb_history_kernel<0, false, 768> : xmb_history_kernel<0, false, 896>;
	if (options->use_advanced_compton) kernel = xmb_history_kernel<0, true>;   // opt-in physics: one generic-nL instantiation (compiled for 256 threads and 225 registers it is slower: 79 -> 113 ms on srm1132)
	if (P.stage_nch == 0) kernel = options->use_advanced_compton ? xmb_history_kernel<0, true, HIST_THREADS, true> : xmb_history_kernel<0, false, HIST_THREADS, true>;   // channel deposits unstaged
	XMB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	if (const char *e = getenv("XMB_SMEM_CARVEOUT")) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));   // experiments: percent of the SM's L1 + shared array
	XMB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
	if (occ < 1) occ = 1;
	const uint64_t n_chunks = (ex->n_histories + threads - 1) / threads;
	uint64_t blocks = (uint64_t)sms * occ;
	blocks = std::max<uint64_t>(1, std::min<uint64_t>(blocks, n_chunks));
	if (const char *e = getenv("XMB_HIST_BLOCKS")) { const long b = atol(e); if (b >= 1 && (uint64_t)b < blocks) blocks = (uint64_t)b; }
	if (P.n_int > XMB_MAX_ORDERS) { xmb_set_error("more than %d interactions per trajectory", XMB_MAX_ORDERS); return 0; }
	// per-CTA compaction queues: n_int orders (x nL layers) x 2T photons x 13 doubles, + nL with one or two layers (structure of arrays)
	size_t qd = (size_t)blocks * P.n_int * (XMB_STATE_FIELDS + (XMB_QUEUE_MUS_FOR(P.nL) ? P.nL : 0)) * 2 * threads;
	if (P.layer_sort == 2) {
		if (qd * P.nL * sizeof(double) > ((size_t)24 << 30)) P.layer_sort = 1;   // keep the queues within 24 GB of HBM
		else qd *= P.nL;
	}
	if (D->queue_doubles < qd) {
		cudaFree(D->queue);
		D->queue = nullptr; D->queue_doubles = 0;
		XMB_CUDA_OK(cudaMalloc(&D->queue, sizeof(double) * qd));
		D->queue_doubles = qd;
	}
	P.queue = D->queue;
	cudaEventRecord(D->ev0);
	if (ex->n_histories > 0) { kernel<<<(unsigned)blocks, threads, smem>>>(P); D->run_launches++; }
	cudaEventRecord(D->ev1);
	xmb_limbs4_kernel<<<sms, 256>>>(D->acc, D->limbs, slots);
	XMB_CUDA_OK(cudaGetLastError());
	return 1;
}

int xmb_msim_collect(XmbDeviceTables *D, const xmb_main_options *options, xmb_msim_ex *ex) {
	int cur = 0;
	cudaGetDevice(&cur);
	if (cur != D->device) XMB_CUDA_OK(cudaSetDevice(D->device));
	XMB_CUDA_OK(cudaEventSynchronize(D->ev1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, D->ev0, D->ev1);
	ex->kernel_ms = ms;
	ex->n_launches = D->run_launches;
	unsigned long long cnt[XMB_N_COUNTERS];
	XMB_CUDA_OK(cudaMemcpy(cnt, D->counters, sizeof(cnt), cudaMemcpyDeviceToHost));
	ex->n_interactions = cnt[1];
	if (D->run_brute) {
		for (int i = 0; i < 8; i++) D->brute_counters[i] = cnt[i];
		if (cnt[5] && options->verbose) fprintf(stderr, "detected photons of lines without a history slot: %llu\n", cnt[5]);
	} else {
		for (int i = 0; i < XMB_MAX_LAYERS; i++) D->layer_interactions[i] = cnt[8 + i];
		if (cnt[0] && options->verbose) fprintf(stderr, "detector_solid_angle_not_found: %llu\n", cnt[0]);
	}
	if (getenv("XMB_PHASES")) {   // experiment builds (-DXMB_PHASE_CLOCKS=1): share of the warps' time per phase of the batch loop
		static const char *nm[16] = {"scheduler+batch", "geometry", "off-grid+barrier", "barrier 1", "element: rest", "line deposits", "barrier 4", "flush",
		                             "select+scatter", "move+queue", "end barrier", "element: prefetch+Rayleigh", "element: Doppler trials",
		                             "element: bracket..fixed", "element: staged adds", "-"};
		double tot = 0.0;
		for (int i = 0; i < 16; i++) tot += (double)cnt[40 + i];
		if (tot > 0.0) for (int i = 0; i < 16; i++) fprintf(stderr, "phase %-18s %5.1f %%\n", nm[i], 100.0 * (double)cnt[40 + i] / tot);
	}
	if (cnt[2]) { xmb_set_error("%llu deposits fell outside the fixed-point range", cnt[2]); return 0; }
	return 1;
}

extern "C" int xmb_main_msim_raw(xmb_inputFPtr inputF, xmb_hdf5FPtr hdf5F, const xmb_main_options *options,
                                 const xmb_solid_angle *sa, xmb_msim_ex *ex, uint64_t **accum, size_t *n_slots) {
	XmbInputF *in = xmb_as_input(inputF);
	XmbHdf5F *h = xmb_as_hdf5(hdf5F);
	if (!in || !h || !options || !ex || !accum || !n_slots) { xmb_set_error("xmb_main_msim_raw: bad arguments"); return 0; }
	XmbDeviceTables *D = nullptr;
	if (!xmb_msim_launch(in, h, options, sa, ex, &D)) return 0;
	if (!xmb_msim_collect(D, options, ex)) return 0;
	*n_slots = D->run_slots;
	if (ex->keep_on_device) { *accum = nullptr; return 1; }   // limbs stay in HBM: xmb_msim_device_limbs()
	uint64_t *out = (uint64_t *)malloc(sizeof(uint64_t) * 2 * D->run_slots);
	XMB_CUDA_OK(cudaMemcpy(out, D->limbs, sizeof(uint64_t) * 2 * D->run_slots, cudaMemcpyDeviceToHost));
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
	auto fits = [&](XmbDeviceTables *D) { return D && D->cascade == xmb_cascade_mode(opt) && D->use_M_lines == (opt->use_M_lines ? 1 : 0); };
	if (fits(h->dev)) return h->dev;
	for (XmbDeviceTables *D : h->devs) if (fits(D)) return D;
	// host-side layout only (slot map, epilogue): a handle without device memory, device = -2
	for (size_t i = 0; i < h->devs.size(); i++) if (h->devs[i]->device == -2) { if (h->dev == h->devs[i]) h->dev = nullptr; delete h->devs[i]; h->devs.erase(h->devs.begin() + i); break; }
	g_layout_only = true;
	XmbDeviceTables *D = build_device_tables(in, h, opt);
	g_layout_only = false;
	if (D) { h->devs.push_back(D); if (!h->dev) h->dev = D; }
	return D;
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

