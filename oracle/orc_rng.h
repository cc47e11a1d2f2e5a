/* orc_rng.h -- per-stream Philox cursor used by every oracle routine (test infrastructure). */
#ifndef ORC_RNG_H
#define ORC_RNG_H
#include <stdint.h>
#include "oracle.h"
typedef struct orc_rng {
	uint32_t key[2];
	uint32_t ctr[4];   /* ctr[0..1] = stream id (lo, hi), ctr[2] = block counter, ctr[3] = stream tag */
	uint32_t buf[4];
	int have;          /* unread words left in buf */
} orc_rng;
static inline void orc_rng_init(orc_rng *r, uint64_t seed, uint64_t stream, uint32_t tag) {
	r->key[0] = (uint32_t)seed; r->key[1] = (uint32_t)(seed >> 32);
	r->ctr[0] = (uint32_t)stream; r->ctr[1] = (uint32_t)(stream >> 32);
	r->ctr[2] = 0; r->ctr[3] = tag; r->have = 0;
}
static inline uint32_t orc_rng_u32(orc_rng *r) {
	if (r->have == 0) { orc_philox4x32_10(r->ctr, r->key, r->buf); r->ctr[2]++; r->have = 4; }
	return r->buf[4 - r->have--];
}
/* uniform in [0,1) with 32-bit resolution, as gsl/easyRNG rng_uniform on MT19937 */
static inline double orc_rng_uniform(orc_rng *r) { return orc_rng_u32(r) * (1.0 / 4294967296.0); }
#define ORC_TAG_SOLID_ANGLE 0x5Au
#define ORC_TAG_SA_FALLBACK 0x5Bu
#define ORC_TAG_HISTORY 0x48u
#define ORC_TAG_DETECTOR 0x44u
#endif
