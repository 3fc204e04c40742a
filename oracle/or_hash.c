/*
 * or_hash.c — ORACLE (test infrastructure): restatement of the reference's
 * hash functions and segment routing.
 *
 *   hash_any / hash_uint32     src/backend/access/hash/hashfunc.c:241-552 (Jenkins lookup3, PG variant)
 *   hashint4/int8/float8       src/backend/access/hash/hashfunc.c:46-125
 *   hashbpchar / bcTruelen     src/backend/utils/adt/varchar.c:653-671,906-922
 *   bpchareq / bpcharcmp       src/backend/utils/adt/varchar.c:702-726,840-860 (C locale => memcmp)
 *   cdbhashinit/cdbhash/reduce src/backend/cdb/cdbhash.c:173-287
 *   jump_consistent_hash       src/backend/cdb/cdbhash.c:549-560
 *
 * Pinned by tests/golden/hash_kat.json, produced from the reference's own
 * hashfunc.o / cdbhash.o (oracle/ref_build) by tests/golden/make_golden.py.
 */
#include <string.h>
#include "gg_oracle.h"

#define ROT(x, k) (((x) << (k)) | ((x) >> (32 - (k))))

#define MIX(a, b, c) \
	do { \
		a -= c; a ^= ROT(c, 4);  c += b; \
		b -= a; b ^= ROT(a, 6);  a += c; \
		c -= b; c ^= ROT(b, 8);  b += a; \
		a -= c; a ^= ROT(c, 16); c += b; \
		b -= a; b ^= ROT(a, 19); a += c; \
		c -= b; c ^= ROT(b, 4);  b += a; \
	} while (0)

#define FINAL(a, b, c) \
	do { \
		c ^= b; c -= ROT(b, 14); \
		a ^= c; a -= ROT(c, 11); \
		b ^= a; b -= ROT(a, 25); \
		c ^= b; c -= ROT(b, 16); \
		a ^= c; a -= ROT(c, 4);  \
		b ^= a; b -= ROT(a, 14); \
		c ^= b; c -= ROT(b, 24); \
	} while (0)

/* hashfunc.c:302.  Little-endian: the aligned and unaligned paths of the
 * reference give the same value, so only the byte-wise path is restated. */
uint32_t
or_hash_any(const unsigned char *k, int keylen)
{
	uint32_t a, b, c, len;

	len = (uint32_t) keylen;
	a = b = c = 0x9e3779b9 + len + 3923095;

	while (len >= 12)
	{
		a += (k[0] + ((uint32_t) k[1] << 8) + ((uint32_t) k[2] << 16) + ((uint32_t) k[3] << 24));
		b += (k[4] + ((uint32_t) k[5] << 8) + ((uint32_t) k[6] << 16) + ((uint32_t) k[7] << 24));
		c += (k[8] + ((uint32_t) k[9] << 8) + ((uint32_t) k[10] << 16) + ((uint32_t) k[11] << 24));
		MIX(a, b, c);
		k += 12;
		len -= 12;
	}
	switch (len)
	{
		case 11: c += ((uint32_t) k[10] << 24);	/* fall through */
		case 10: c += ((uint32_t) k[9] << 16);	/* fall through */
		case 9:  c += ((uint32_t) k[8] << 8);	/* lowest byte of c is reserved for the length */
			/* fall through */
		case 8:  b += ((uint32_t) k[7] << 24);	/* fall through */
		case 7:  b += ((uint32_t) k[6] << 16);	/* fall through */
		case 6:  b += ((uint32_t) k[5] << 8);	/* fall through */
		case 5:  b += k[4];						/* fall through */
		case 4:  a += ((uint32_t) k[3] << 24);	/* fall through */
		case 3:  a += ((uint32_t) k[2] << 16);	/* fall through */
		case 2:  a += ((uint32_t) k[1] << 8);	/* fall through */
		case 1:  a += k[0];
	}
	FINAL(a, b, c);
	return c;
}

/* hashfunc.c:527 */
uint32_t
or_hash_uint32(uint32_t k)
{
	uint32_t a, b, c;

	a = b = c = 0x9e3779b9 + (uint32_t) sizeof(uint32_t) + 3923095;
	a += k;
	FINAL(a, b, c);
	return c;
}

uint32_t
or_hashint4(int32_t v)
{
	return or_hash_uint32((uint32_t) v);
}

/* hashfunc.c:52 — xor the high half (or its complement when negative) so that
 * int8 values in int4 range hash like hashint4 */
uint32_t
or_hashint8(int64_t val)
{
	uint32_t lohalf = (uint32_t) val;
	uint32_t hihalf = (uint32_t) (val >> 32);

	lohalf ^= (val >= 0) ? hihalf : ~hihalf;
	return or_hash_uint32(lohalf);
}

/* hashfunc.c:110 */
uint32_t
or_hashfloat8(double key)
{
	if (key == (double) 0)
		return 0;
	return or_hash_any((const unsigned char *) &key, sizeof(key));
}

/* varchar.c:653 */
int
or_bctruelen(const char *s, int len)
{
	int i;

	for (i = len - 1; i >= 0; i--)
		if (s[i] != ' ')
			break;
	return i + 1;
}

/* varchar.c:906 */
uint32_t
or_hashbpchar(const char *s, int len)
{
	return or_hash_any((const unsigned char *) s, or_bctruelen(s, len));
}

/* varchar.c:702 */
int
or_bpchareq(const char *a, int la, const char *b, int lb)
{
	int len1 = or_bctruelen(a, la);
	int len2 = or_bctruelen(b, lb);

	if (len1 != len2)
		return 0;
	return memcmp(a, b, len1) == 0;
}

/* varchar.c:840 -> varstr_cmp (varlena.c) with C collation: memcmp of the
 * common prefix, then the shorter string sorts first */
int
or_bpcharcmp(const char *a, int la, const char *b, int lb)
{
	int len1 = or_bctruelen(a, la);
	int len2 = or_bctruelen(b, lb);
	int r = memcmp(a, b, len1 < len2 ? len1 : len2);

	if (r == 0 && len1 != len2)
		r = (len1 < len2) ? -1 : 1;
	return r;
}

/* Hash proc of the default hash opclass for the supported types
 * (cdb_hashproc_in_opfamily, cdbhash.c:135-165; execTuplesHashPrepare, execGrouping.c:220).
 * Strings arrive packed (<= 8 blank-stripped bytes, LSB first) with their length. */
uint32_t
or_hash_datum(int32_t typid, int64_t datum, int32_t len)
{
	switch (typid)
	{
		case GG_INT4OID:
		case GG_DATEOID:		/* date uses hashint4 (pg_amproc: date_ops -> hashint4) */
			return or_hashint4((int32_t) datum);
		case GG_INT8OID:
		case GG_TIMESTAMPOID:	/* timestamp_hash = hashint8 with integer datetimes */
			return or_hashint8(datum);
		case GG_FLOAT8OID:
		{
			double d;

			memcpy(&d, &datum, 8);
			return or_hashfloat8(d);
		}
		case GG_BPCHAROID:
			return or_hashbpchar((const char *) &datum, len);
		case GG_VARCHAROID:
		case GG_TEXTOID:		/* hashtext: hash_any over the payload, no blank stripping */
			return or_hash_any((const unsigned char *) &datum, len);
		case GG_BOOLOID:		/* hashchar */
			return or_hash_uint32((uint32_t) (int32_t) (int8_t) datum);
	}
	return 0;
}

/* ---- cdbhash.c ---- */

uint32_t
or_cdbhash_init(void)
{
	return 0;					/* cdbhash.c:176, non-legacy */
}

/* cdbhash.c:191-219: rotate left one bit, xor the column's hash unless NULL */
uint32_t
or_cdbhash_add(uint32_t hashkey, uint32_t hkey, int isnull)
{
	hashkey = (hashkey << 1) | ((hashkey & 0x80000000) ? 1 : 0);
	if (!isnull)
		hashkey ^= hkey;
	return hashkey;
}

/* cdbhash.c:549-560.  The (double) division is the reference's; keep it a
 * plain IEEE double divide followed by a double multiply and a truncating
 * conversion (no FMA contraction: compile with -ffp-contract=off). */
int32_t
or_jump_consistent_hash(uint64_t key, int32_t num_segments)
{
	int64_t b = -1;
	int64_t j = 0;

	while (j < num_segments)
	{
		b = j;
		key = key * 2862933555777941757ULL + 1;
		j = (int64_t) ((double) (b + 1) * ((double) (1LL << 31) / (double) ((key >> 33) + 1)));
	}
	return (int32_t) b;
}

/* cdbhash.c:255-287, REDUCE_JUMP_HASH (the non-legacy default, cdbhash.c:118) */
int32_t
or_cdbhash_reduce(uint32_t h, int32_t nsegs)
{
	return or_jump_consistent_hash((uint64_t) h, nsegs);
}

int32_t
or_route_datums(const int32_t *typids, const int64_t *vals, const int32_t *lens,
				const int32_t *isnull, int nkeys, int nsegs)
{
	uint32_t h = or_cdbhash_init();
	int i;

	for (i = 0; i < nkeys; i++)
	{
		uint32_t hk = isnull[i] ? 0 : or_hash_datum(typids[i], vals[i], lens[i]);

		h = or_cdbhash_add(h, hk, isnull[i]);
	}
	return or_cdbhash_reduce(h, nsegs);
}
