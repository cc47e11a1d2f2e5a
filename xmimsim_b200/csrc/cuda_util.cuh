// cuda_util.cuh -- error handling + Philox4x32-10 for the device code.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "engine.h"

#define XMB_CUDA_OK(call)                                                                         \
	do {                                                                                          \
		cudaError_t e__ = (call);                                                                 \
		if (e__ != cudaSuccess) {                                                                 \
			xmb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
			return 0;                                                                             \
		}                                                                                         \
	} while (0)

#define XMB_TAG_SOLID_ANGLE 0x5Au
#define XMB_TAG_SA_FALLBACK 0x5Bu
#define XMB_TAG_HISTORY 0x48u
#define XMB_TAG_DETECTOR 0x44u

// ---- bulk asynchronous copy global -> shared with an mbarrier (the TMA engine's 1-D path: UBLKCP in SASS) -----------
// One thread arms the barrier with the byte count and issues the copy; every consumer waits on the barrier's phase
// parity.  Addresses are shared-window addresses (__cvta_generic_to_shared); src / dst 16-byte aligned, bytes % 16 == 0.
#ifdef __CUDACC__
__device__ __forceinline__ void xmb_mbar_init(unsigned mbar_s32, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s32), "r"(count) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void xmb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void xmb_bulk_g2s(unsigned dst_s32, const void *src, unsigned bytes, unsigned mbar_s32) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s32), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(dst_s32), "l"(src), "r"(bytes), "r"(mbar_s32) : "memory");
}
__device__ __forceinline__ void xmb_mbar_wait(unsigned mbar_s32, unsigned parity) {
	unsigned done;
	do {
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
		             : "=r"(done) : "r"(mbar_s32), "r"(parity) : "memory");
	} while (!done);
}
#endif

// Philox4x32-10 (Random123): one call = 4 x 32 random bits from (counter[4], key[2]).
__host__ __device__ __forceinline__ uint4 xmb_philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
	for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
		const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
#else
		const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
		const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
		c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
		k.x += 0x9E3779B9u;
		k.y += 0xBB67AE85u;
	}
	return c;
}

// uniform in [0,1), 32-bit resolution (what gsl/easyRNG return for MT19937)
__host__ __device__ __forceinline__ double xmb_u01(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// per-stream sequential cursor: ctr = (stream_lo, stream_hi, block, tag)
struct XmbRng {
	uint2 key;
	uint32_t s_lo, s_hi, block, tag;
	uint4 buf;
	int have;
	__device__ __forceinline__ void init(uint64_t seed, uint64_t stream, uint32_t tag_) {
		key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
		s_lo = (uint32_t)stream; s_hi = (uint32_t)(stream >> 32); block = 0; tag = tag_; have = 0;
	}
	__device__ __forceinline__ uint32_t u32() {
		if (have == 0) { buf = xmb_philox4x32_10(make_uint4(s_lo, s_hi, block, tag), key); block++; have = 4; }
		uint32_t w = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
		have--;
		return w;
	}
	__device__ __forceinline__ double uniform() { return xmb_u01(u32()); }
};
