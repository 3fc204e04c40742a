/*
 * gg_hostutil.h — host-side (product) helpers: bit-exact segment routing.
 *   hash_uint32 / hashint8      src/backend/access/hash/hashfunc.c:527,52
 *   cdbhash / cdbhashreduce     src/backend/cdb/cdbhash.c:191-287
 *   jump_consistent_hash        src/backend/cdb/cdbhash.c:549-560
 * Used by the synthetic loader (DISTRIBUTED BY placement) and by the Motion host logic.
 */
#ifndef GG_HOSTUTIL_H
#define GG_HOSTUTIL_H
#include <stdint.h>

static inline uint32_t ggh_rot(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

static inline uint32_t ggh_hash_uint32(uint32_t k)
{
	uint32_t a, b, c;
	a = b = c = 0x9e3779b9u + 4u + 3923095u;
	a += k;
	c ^= b; c -= ggh_rot(b, 14); a ^= c; a -= ggh_rot(c, 11); b ^= a; b -= ggh_rot(a, 25);
	c ^= b; c -= ggh_rot(b, 16); a ^= c; a -= ggh_rot(c, 4);  b ^= a; b -= ggh_rot(a, 14);
	c ^= b; c -= ggh_rot(b, 24);
	return c;
}

static inline uint32_t ggh_hashint8(int64_t val)
{
	uint32_t lo = (uint32_t) val, hi = (uint32_t) ((uint64_t) val >> 32);
	lo ^= (val >= 0) ? hi : ~hi;
	return ggh_hash_uint32(lo);
}

static inline int32_t ggh_jump_consistent_hash(uint64_t key, int32_t nsegs)
{
	int64_t b = -1, j = 0;
	while (j < nsegs)
	{
		b = j;
		key = key * 2862933555777941757ULL + 1;
		j = (int64_t) ((double) (b + 1) * ((double) (1LL << 31) / (double) ((key >> 33) + 1)));
	}
	return (int32_t) b;
}

/* segment of a row distributed by one int8 key: cdbhashinit (h=0), cdbhash (rotl1 ^ hashint8), reduce */
static inline int32_t ggh_seg_of_int8(int64_t key, int32_t nsegs)
{
	uint32_t h = 0;
	h = (h << 1) | (h >> 31);
	h ^= ggh_hashint8(key);
	return ggh_jump_consistent_hash((uint64_t) h, nsegs);
}
#endif
