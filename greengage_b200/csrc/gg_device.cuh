/*
 * gg_device.cuh — device-side building blocks shared by the kernels:
 *   - mbarrier / TMA bulk-copy PTX wrappers (sm_100a)
 *   - heap page / tuple decoding (bufpage.h:153, itemid.h:24, htup_details.h:139,
 *     tupmacs.h:23-175, postgres.h:158-300 big-endian varlena headers)
 *   - the accumulator-machine interpreter for compiled expressions (gg_program.h)
 *   - bit-exact Jenkins hash / cdbhash / jump-consistent-hash (hashfunc.c:241-552,
 *     cdbhash.c:191-287,549-560)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gg_program.h"

#define GG_FULL_MASK 0xffffffffu

namespace ggd {

/* ---------------- PTX wrappers ---------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	uint32_t addr = smem_u32(bar);
	do
	{
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(addr), "r"(parity) : "memory");
	} while (!done);
}
/* TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ---------------- hashing: bit-exact with hashfunc.c ---------------- */
__device__ __forceinline__ uint32_t rot32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

#define GGD_FINAL(a, b, c) \
	{ c ^= b; c -= rot32(b, 14); a ^= c; a -= rot32(c, 11); b ^= a; b -= rot32(a, 25); \
	  c ^= b; c -= rot32(b, 16); a ^= c; a -= rot32(c, 4);  b ^= a; b -= rot32(a, 14); \
	  c ^= b; c -= rot32(b, 24); }
#define GGD_MIX(a, b, c) \
	{ a -= c; a ^= rot32(c, 4);  c += b; b -= a; b ^= rot32(a, 6);  a += c; \
	  c -= b; c ^= rot32(b, 8);  b += a; a -= c; a ^= rot32(c, 16); c += b; \
	  b -= a; b ^= rot32(a, 19); a += c; c -= b; c ^= rot32(b, 4);  b += a; }

/* hash_uint32, hashfunc.c:527 */
__device__ __forceinline__ uint32_t hash_uint32(uint32_t k)
{
	uint32_t a, b, c;
	a = b = c = 0x9e3779b9u + 4u + 3923095u;
	a += k;
	GGD_FINAL(a, b, c);
	return c;
}
/* hash_any over <= 8 bytes held LSB-first in a register (hashfunc.c:302; tail switch cases 1..8) */
__device__ __forceinline__ uint32_t hash_any_le8(uint64_t v, int len)
{
	uint32_t a, b, c;
	a = b = c = 0x9e3779b9u + (uint32_t) len + 3923095u;
	uint64_t m = len >= 8 ? ~0ull : ((1ull << (8 * len)) - 1ull);
	v &= m;
	a += (uint32_t) v;
	b += (uint32_t) (v >> 32);
	GGD_FINAL(a, b, c);
	return c;
}
/* hashint8, hashfunc.c:52 */
__device__ __forceinline__ uint32_t hashint8(int64_t val)
{
	uint32_t lo = (uint32_t) val, hi = (uint32_t) ((uint64_t) val >> 32);
	lo ^= (val >= 0) ? hi : ~hi;
	return hash_uint32(lo);
}
/* hashfloat8, hashfunc.c:110 */
__device__ __forceinline__ uint32_t hashfloat8(uint64_t bits)
{
	double d = __longlong_as_double((long long) bits);
	if (d == 0.0) return 0;
	return hash_any_le8(bits, 8);
}
/* cdbhash.c:197-219: rotate left 1, xor the column hash unless NULL */
__device__ __forceinline__ uint32_t cdbhash_add(uint32_t h, uint32_t hk, bool isnull)
{
	h = (h << 1) | (h >> 31);
	return isnull ? h : (h ^ hk);
}
/* jump_consistent_hash, cdbhash.c:549-560.  Same IEEE operations as the C code:
 * int->double conversions, one double divide (round-to-nearest), one double multiply,
 * truncating conversion; __dmul_rn/__ddiv_rn keep the compiler from contracting them. */
__device__ __forceinline__ int32_t jump_consistent_hash(uint64_t key, int32_t nsegs)
{
	int64_t b = -1, j = 0;
	while (j < nsegs)
	{
		b = j;
		key = key * 2862933555777941757ULL + 1;
		double q = __ddiv_rn((double) (1LL << 31), (double) ((key >> 33) + 1));
		j = (int64_t) __dmul_rn((double) (b + 1), q);
	}
	return (int32_t) b;
}

/* ---------------- tuple decoding ---------------- */
__device__ __forceinline__ uint32_t align_nominal(uint32_t off, int attalign)
{
	/* tupmacs.h:121-130 */
	uint32_t m = attalign == 'd' ? 7u : attalign == 'i' ? 3u : attalign == 's' ? 1u : 0u;
	return (off + m) & ~m;
}
/* VARSIZE_ANY for inline datums, postgres.h:276.  0x80 (external TOAST pointer) and
 * compressed 4-byte headers are reported through *bad. */
__device__ __forceinline__ uint32_t varsize_any(const uint8_t *p, bool *bad)
{
	uint32_t h = p[0];
	if (h & 0x80)
	{
		if (h == 0x80) { *bad = true; return 4; }
		return h & 0x7F;
	}
	if (h & 0x40) *bad = true;   /* compressed in line */
	uint32_t l = ((h & 0x3F) << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3];
	if (l < 4) { *bad = true; return 4; }
	return l;
}

/* Per-lane view of one tuple after the attribute walk */
struct TupleView {
	const uint8_t *tp;      /* start of user data (tuple + t_hoff) */
	uint32_t colnull;       /* bit s: column slot s is NULL */
};

/* The attribute walk: slot_deform_tuple (heaptuple.c:1119-1213) restricted to the attributes the
 * program references.  Offsets of referenced columns go to offs[slot*32 + lane] (shared memory).
 * Returns false (and sets err bits) if the tuple is malformed. */
__device__ __forceinline__ void walk_tuple(const ggp_side &S, const uint8_t *tup, uint32_t tuplen,
                                           uint16_t *offs, int lane, TupleView &tv, uint32_t &err)
{
	uint32_t infomask = *(const uint16_t *) (tup + 20);
	uint32_t tnatts = *(const uint16_t *) (tup + 18) & GG_HEAP_NATTS_MASK;
	uint32_t hoff = tup[22];
	bool hasnulls = (infomask & GG_HEAP_HASNULL) != 0;
	const uint8_t *bp = tup + GG_HEAP_HDR_SIZE;
	const uint8_t *tp = tup + hoff;
	uint32_t datalen = tuplen > hoff ? tuplen - hoff : 0;
	uint32_t colnull = 0;
	bool bad = false;
	int a0 = 0;
	uint32_t off = 0;

	tv.tp = tp;
	if (!hasnulls)
	{
		a0 = S.first_walk > 0 ? S.first_walk - 1 : 0;
		if (a0 >= S.natts_walk) a0 = S.natts_walk;     /* everything referenced has a constant offset */
		/* constant offsets (attcacheoff) for the fixed-width prefix */
		for (int s = 0; s < S.ncols; s++)
		{
			int a = S.colatt[s];
			if (a < a0 || a0 == S.natts_walk)
			{
				if ((uint32_t) a < tnatts) offs[s * 32 + lane] = (uint16_t) S.att[a].cacheoff;
				else colnull |= 1u << s;               /* attribute added after the tuple was written */
			}
		}
		if (a0 < S.natts_walk) off = a0 > 0 ? (uint32_t) S.att[a0].cacheoff : 0;
	}
	for (int a = a0; a < S.natts_walk; a++)
	{
		const ggp_attr at = S.att[a];
		if ((uint32_t) a >= tnatts || (hasnulls && !(bp[a >> 3] & (1 << (a & 7)))))
		{
			if (at.slot >= 0) colnull |= 1u << at.slot;
			continue;
		}
		if (at.attlen == -1)
		{
			/* att_align_pointer: a zero byte is padding (or an aligned 4-byte header) */
			if (off < datalen && tp[off] == 0) off = align_nominal(off, at.attalign);
		}
		else
			off = align_nominal(off, at.attalign);
		if (at.slot >= 0) offs[at.slot * 32 + lane] = (uint16_t) off;
		if (off >= datalen) { bad = true; break; }
		off += at.attlen > 0 ? (uint32_t) at.attlen : varsize_any(tp + off, &bad);
		if (off > datalen) { bad = true; break; }
	}
	if (bad) err |= GGP_EF_BADPAGE;
	tv.colnull = colnull;
}

/* load a referenced column into the 64-bit accumulator */
__device__ __forceinline__ uint64_t load_col(const ggp_side &S, int slot, const TupleView &tv,
                                             const uint16_t *offs, int lane, uint32_t &err)
{
	const uint8_t *p = tv.tp + offs[slot * 32 + lane];
	switch (S.coltype[slot])
	{
		case GGP_LD_I4: return (uint64_t) (int64_t) * (const int32_t *) p;
		case GGP_LD_I8: return *(const uint64_t *) p;
		case GGP_LD_BOOL: return (uint64_t) (p[0] != 0);
		default:
		{
			/* short string: VARDATA_ANY / VARSIZE_ANY_EXHDR, then bcTruelen for bpchar (varchar.c:653) */
			uint32_t h = p[0], len;
			const uint8_t *d;
			if (h & 0x80)
			{
				if (h == 0x80) { err |= GGP_EF_STRING_TOO_LONG; return 0; }
				len = (h & 0x7F) - 1; d = p + 1;
			}
			else
			{
				if (h & 0x40) { err |= GGP_EF_STRING_TOO_LONG; return 0; }
				len = ((((h & 0x3F) << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3])) - 4;
				d = p + 4;
			}
			if (S.coltype[slot] == GGP_LD_BPCHAR)
				while (len > 0 && d[len - 1] == ' ') len--;
			if (len > 8) { err |= GGP_EF_STRING_TOO_LONG; return 0; }
			uint64_t v = 0;
			for (uint32_t i = 0; i < len; i++) v |= (uint64_t) d[i] << (8 * i);
			return v;
		}
	}
}

__device__ __forceinline__ int f8_cmp(double a, double b)
{
	/* float8_cmp_internal, float.c:964: NaN = NaN, NaN > everything */
	bool na = a != a, nb = b != b;
	if (na) return nb ? 0 : 1;
	if (nb) return -1;
	return (a > b) - (a < b);
}
__device__ __forceinline__ bool test_cc(int c, int cc)
{
	switch (cc)
	{
		case GGP_LT: return c < 0;
		case GGP_LE: return c <= 0;
		case GGP_EQ: return c == 0;
		case GGP_NE: return c != 0;
		case GGP_GT: return c > 0;
		default: return c >= 0;
	}
}
__device__ __forceinline__ bool f8_isinf(double x) { return fabs(x) == __longlong_as_double(0x7ff0000000000000LL); }

/* Run one compiled expression.  Result in acc/accnull.  NULLABLE=false compiles the null tracking out. */
template <bool NULLABLE, bool HAS_INNER>
__device__ __forceinline__ void run_span(const ggp_program &P, ggp_span sp,
                                         const TupleView &tv, const uint16_t *offs,
                                         const ggp_side *IS, const TupleView *itv, const uint16_t *ioffs,
                                         int lane, uint64_t &acc, bool &accnull, uint32_t &err)
{
	uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
	uint32_t tnull = 0;
	acc = 0;
	accnull = false;
	const int end = sp.start + sp.len;
	for (int pc = sp.start; pc < end; pc++)
	{
		const ggp_op o = P.code[pc];
		uint64_t sv = 0;
		bool sn = false;
		switch (o.src)
		{
			case GGP_SRC_COL:
				if (NULLABLE) sn = (tv.colnull >> o.idx) & 1;
				if (!sn) sv = load_col(P.outer, o.idx, tv, offs, lane, err);
				break;
			case GGP_SRC_ICOL:
				if (HAS_INNER)
				{
					if (NULLABLE) sn = (itv->colnull >> o.idx) & 1;
					if (!sn) sv = load_col(*IS, o.idx, *itv, ioffs, lane, err);
				}
				break;
			case GGP_SRC_CONST:
				sv = (uint64_t) P.consts[o.idx];
				if (NULLABLE) sn = (P.constnull >> o.idx) & 1;
				break;
			case GGP_SRC_TEMP:
				sv = o.idx == 0 ? t0 : o.idx == 1 ? t1 : o.idx == 2 ? t2 : t3;
				if (NULLABLE) sn = (tnull >> o.idx) & 1;
				break;
			default: break;
		}
		switch (o.op)
		{
			case GGP_LOAD: acc = sv; accnull = sn; break;
			case GGP_STORE:
				if (o.idx == 0) t0 = acc; else if (o.idx == 1) t1 = acc; else if (o.idx == 2) t2 = acc; else t3 = acc;
				if (NULLABLE) tnull = (tnull & ~(1u << o.idx)) | ((uint32_t) accnull << o.idx);
				break;
			case GGP_F8ADD: case GGP_F8SUB: case GGP_F8RSUB: case GGP_F8MUL: case GGP_F8DIV: case GGP_F8RDIV:
			{
				double x = __longlong_as_double((long long) acc), y = __longlong_as_double((long long) sv), r;
				if (o.op == GGP_F8RSUB || o.op == GGP_F8RDIV) { double t = x; x = y; y = t; }
				bool isnull = NULLABLE && (accnull || sn);
				if (o.op == GGP_F8ADD) r = __dadd_rn(x, y);
				else if (o.op == GGP_F8SUB || o.op == GGP_F8RSUB) r = __dsub_rn(x, y);
				else if (o.op == GGP_F8MUL) r = __dmul_rn(x, y);
				else
				{
					if (y == 0.0 && !isnull) err |= GGP_EF_DIV_ZERO;
					r = __ddiv_rn(x, y);
				}
				if (!isnull)
				{
					/* CHECKFLOATVAL, float_utils.h:28 */
					if (f8_isinf(r) && !(f8_isinf(x) || f8_isinf(y))) err |= GGP_EF_FLOAT_OVERFLOW;
					if (o.op == GGP_F8MUL && r == 0.0 && !(x == 0.0 || y == 0.0)) err |= GGP_EF_FLOAT_UNDERFLOW;
					if ((o.op == GGP_F8DIV || o.op == GGP_F8RDIV) && r == 0.0 && x != 0.0 && y != 0.0) err |= GGP_EF_FLOAT_UNDERFLOW;
				}
				acc = (uint64_t) __double_as_longlong(r);
				accnull = isnull;
				break;
			}
			case GGP_CMPF8:
				acc = test_cc(f8_cmp(__longlong_as_double((long long) acc), __longlong_as_double((long long) sv)), o.aux);
				if (NULLABLE) accnull = accnull || sn;
				break;
			case GGP_CMPI:
			{
				int64_t x = (int64_t) acc, y = (int64_t) sv;
				acc = test_cc((x > y) - (x < y), o.aux);
				if (NULLABLE) accnull = accnull || sn;
				break;
			}
			case GGP_CMPSTR:
				acc = (o.aux == GGP_EQ) ? (acc == sv) : (acc != sv);
				if (NULLABLE) accnull = accnull || sn;
				break;
			case GGP_DATE2TS:
			{
				/* date2timestamp, date.c:457 */
				int32_t d = (int32_t) acc;
				int64_t r;
				if (d == INT32_MIN) r = INT64_MIN;
				else if (d == INT32_MAX) r = INT64_MAX;
				else
				{
					r = (int64_t) d * 86400000000LL;      /* wraps like the reference's int64 multiply */
					if (r / 86400000000LL != d && !(NULLABLE && accnull)) err |= GGP_EF_DATE_RANGE;
				}
				acc = (uint64_t) r;
				break;
			}
			case GGP_I2F8:
				acc = (uint64_t) __double_as_longlong((double) (int64_t) acc);
				break;
			case GGP_AND:
			case GGP_OR:
			{
				bool a = acc != 0, b = sv != 0, an = NULLABLE && accnull, bn = NULLABLE && sn;
				if (o.op == GGP_AND)
				{
					if ((!an && !a) || (!bn && !b)) { acc = 0; accnull = false; }
					else if (an || bn) { acc = 0; accnull = true; }
					else { acc = 1; accnull = false; }
				}
				else
				{
					if ((!an && a) || (!bn && b)) { acc = 1; accnull = false; }
					else if (an || bn) { acc = 0; accnull = true; }
					else { acc = 0; accnull = false; }
				}
				break;
			}
			case GGP_NOT: acc = (acc == 0); break;
			case GGP_ISNULL: acc = accnull; accnull = false; break;
			case GGP_ISNOTNULL: acc = !accnull; accnull = false; break;
			default: break;
		}
	}
}

/* normalise a grouping key so that bitwise equality == SQL equality for grouping */
__device__ __forceinline__ uint64_t normalize_key(uint64_t v, int keytype)
{
	if (keytype == 2)
	{
		double d = __longlong_as_double((long long) v);
		if (d == 0.0) return 0;                               /* -0 = +0 (float8eq) */
		if (d != d) return 0x7ff8000000000000ull;             /* all NaNs are equal (float.c:964) */
	}
	return v;
}

}  // namespace ggd
