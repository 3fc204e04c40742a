/*
 * gg_motion_host.c — host side of Motion: where does a row go?
 *
 * Replaces evalHashKey (src/backend/executor/nodeMotion.c:1481-1530) for rows that are already on the
 * host as Datums (the handful of partial-aggregate rows a slice emits): cdbhashinit / cdbhash /
 * cdbhashreduce (src/backend/cdb/cdbhash.c:173-287) with the default hash opclass functions
 * hashint4 / hashint8 / hashfloat8 / hashbpchar / hashtext (hashfunc.c:46-125, varchar.c:906),
 * bit-exact so that a GPU segment and a CPU segment agree on placement.  Bulk redistribution of
 * scanned rows happens on the device (csrc/gg_motion.cu).
 */
#include <stdint.h>
#include <string.h>
#include "../../include/gg_plan.h"
#include "gg_hostutil.h"

#define MIX(a, b, c) \
	{ a -= c; a ^= ggh_rot(c, 4);  c += b; b -= a; b ^= ggh_rot(a, 6);  a += c; \
	  c -= b; c ^= ggh_rot(b, 8);  b += a; a -= c; a ^= ggh_rot(c, 16); c += b; \
	  b -= a; b ^= ggh_rot(a, 19); a += c; c -= b; c ^= ggh_rot(b, 4);  b += a; }
#define FINAL(a, b, c) \
	{ c ^= b; c -= ggh_rot(b, 14); a ^= c; a -= ggh_rot(c, 11); b ^= a; b -= ggh_rot(a, 25); \
	  c ^= b; c -= ggh_rot(b, 16); a ^= c; a -= ggh_rot(c, 4);  b ^= a; b -= ggh_rot(a, 14); \
	  c ^= b; c -= ggh_rot(b, 24); }

/* hash_any, hashfunc.c:302 (little-endian byte path) */
uint32_t gg_hash_any(const unsigned char *k, int keylen)
{
	uint32_t a, b, c, len = (uint32_t) keylen;
	a = b = c = 0x9e3779b9u + len + 3923095u;
	while (len >= 12)
	{
		a += k[0] + ((uint32_t) k[1] << 8) + ((uint32_t) k[2] << 16) + ((uint32_t) k[3] << 24);
		b += k[4] + ((uint32_t) k[5] << 8) + ((uint32_t) k[6] << 16) + ((uint32_t) k[7] << 24);
		c += k[8] + ((uint32_t) k[9] << 8) + ((uint32_t) k[10] << 16) + ((uint32_t) k[11] << 24);
		MIX(a, b, c);
		k += 12;
		len -= 12;
	}
	switch (len)
	{
		case 11: c += (uint32_t) k[10] << 24;	/* fall through */
		case 10: c += (uint32_t) k[9] << 16;	/* fall through */
		case 9:  c += (uint32_t) k[8] << 8;		/* fall through */
		case 8:  b += (uint32_t) k[7] << 24;	/* fall through */
		case 7:  b += (uint32_t) k[6] << 16;	/* fall through */
		case 6:  b += (uint32_t) k[5] << 8;		/* fall through */
		case 5:  b += k[4];						/* fall through */
		case 4:  a += (uint32_t) k[3] << 24;	/* fall through */
		case 3:  a += (uint32_t) k[2] << 16;	/* fall through */
		case 2:  a += (uint32_t) k[1] << 8;		/* fall through */
		case 1:  a += k[0];
	}
	FINAL(a, b, c);
	return c;
}

static uint32_t hash_datum(int32_t typid, int64_t v, int32_t len)
{
	switch (typid)
	{
		case GG_INT4OID: case GG_DATEOID: return ggh_hash_uint32((uint32_t) (int32_t) v);
		case GG_INT8OID: case GG_TIMESTAMPOID: return ggh_hashint8(v);
		case GG_FLOAT8OID:
		{
			double d;
			memcpy(&d, &v, 8);
			if (d == 0.0) return 0;
			return gg_hash_any((const unsigned char *) &v, 8);
		}
		case GG_BPCHAROID:
		{
			/* packed strings are already blank-stripped (bcTruelen); a packed Datum holds at most 8 bytes */
			return gg_hash_any((const unsigned char *) &v, len < 0 ? 0 : (len > 8 ? 8 : len));
		}
		case GG_VARCHAROID: case GG_TEXTOID:
			return gg_hash_any((const unsigned char *) &v, len < 0 ? 0 : (len > 8 ? 8 : len));
		case GG_BOOLOID:
			return ggh_hash_uint32((uint32_t) (int32_t) (int8_t) v);
	}
	return 0;
}

/* destination segment of a row whose distribution keys are given as Datums */
int32_t gg_cdbhash_route(const int32_t *typids, const int64_t *vals, const int32_t *lens, const int32_t *isnull,
                         int nkeys, int nsegs)
{
	uint32_t h = 0;
	int i;
	for (i = 0; i < nkeys; i++)
	{
		h = (h << 1) | (h >> 31);
		if (!isnull[i]) h ^= hash_datum(typids[i], vals[i], lens[i]);
	}
	return ggh_jump_consistent_hash((uint64_t) h, nsegs);
}

/* the same routing for an array of aggregate rows keyed by their group columns: dest[i] = receiving segment of rows[i]
 * (what evalHashKey does per tuple in the sending Motion's loop, nodeMotion.c:1481) */
void gg_cdbhash_route_aggrows(const gg_aggrow *rows, int n, const int32_t *typids, int nkeys, int nsegs, int32_t *dest)
{
	int r;
	for (r = 0; r < n; r++)
		dest[r] = gg_cdbhash_route(typids, rows[r].key, rows[r].keylen, rows[r].keyisnull, nkeys, nsegs);
}

