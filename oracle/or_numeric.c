/*
 * or_numeric.c — ORACLE (test infrastructure): restatement of the numeric arithmetic the hot path needs, tuple at a time,
 * with the display-scale rules of src/backend/utils/adt/numeric.c:
 *   on-disk form -> value               numeric.c:95-190 (NumericShort / NumericLong, base-10000 digits)
 *   numeric_add / numeric_sub           :1659,1698 -> add_var / sub_var: res_dscale = Max(dscale1, dscale2)
 *   numeric_mul                         :1735 -> mul_var with rscale = dscale1 + dscale2
 *   numeric_cmp                         :1512 -> cmp_var: values, not representations
 *   numeric_avg_accum / numeric_sum / numeric_avg   :3057,3205,3173: exact running sum and N; avg = numeric_div(sum, N)
 *   numeric_div                         :1773 -> select_div_scale + div_var(round = true)
 * A value is held as a 128-bit integer scaled by 10^dscale (the product computes on 64-bit integers with scales fixed at
 * plan time: a different route to the same exact answers).  What does not fit 128 bits is OR_ERR_UNSUPPORTED.
 * Pinned against the reference's own numeric.o: tests/golden/numeric_kat.json, tests/test_oracle_numeric.py.
 */
#include <string.h>
#include "gg_oracle.h"
#include "or_internal.h"

typedef __int128 i128;
typedef unsigned __int128 u128;

static const i128 I128_MAX = (i128) (((u128) 1 << 127) - 1);

static i128 get128(const or_datum *d) { return (i128) (((u128) (uint64_t) d->hi << 64) | (u128) (uint64_t) d->v); }
static void put128(or_datum *d, i128 x, int dscale)
{
	d->v = (int64_t) (uint64_t) (u128) x;
	d->hi = (int64_t) (uint64_t) ((u128) x >> 64);
	d->dscale = dscale;
	d->isnull = 0; d->len = 0; d->ptr = NULL;
}

static int mul_chk(i128 a, i128 b, i128 *r) { return __builtin_mul_overflow(a, b, r); }
static int add_chk(i128 a, i128 b, i128 *r) { return __builtin_add_overflow(a, b, r); }

static int rescale(i128 *x, int from, int to)
{
	for (; from < to; from++)
		if (mul_chk(*x, 10, x)) return OR_ERR_UNSUPPORTED;
	return 0;
}

/* numeric payload (behind the varlena header) -> value at its own display scale */
int
or_numeric_decode(const uint8_t *p, int len, or_datum *res)
{
	unsigned hdr;
	int neg, dscale, weight, nd, i;
	const uint8_t *dp;
	u128 v = 0;

	if (len < 2) return OR_ERR_UNSUPPORTED;
	hdr = p[0] | (p[1] << 8);
	if ((hdr & 0xC000) == 0xC000) return OR_ERR_UNSUPPORTED;			/* NaN */
	if (hdr & 0x8000)
	{
		neg = (hdr & 0x2000) != 0;
		dscale = (hdr & 0x1F80) >> 7;
		weight = (hdr & 0x3F) | ((hdr & 0x40) ? ~0x3F : 0);
		dp = p + 2;
	}
	else
	{
		if (len < 4) return OR_ERR_UNSUPPORTED;
		neg = (hdr & 0xC000) == 0x4000;
		dscale = hdr & 0x3FFF;
		weight = (int16_t) (p[2] | (p[3] << 8));
		dp = p + 4;
	}
	nd = (int) ((p + len - dp) / 2);
	/* value = sum digit[i] * 10000^(weight - i), wanted scaled by 10^dscale: digit i contributes at decimal exponent
	 * 4 * (weight - i) + dscale, which is never negative for a value whose dscale covers its stored digits */
	for (i = 0; i < nd; i++)
	{
		int e = 4 * (weight - i) + dscale, k;
		u128 t = (u128) (dp[2 * i] | (dp[2 * i + 1] << 8));

		if (e < 0)
		{
			/* trailing part of the last digit beyond dscale must be zeros (numeric.c keeps dscale >= stored fraction) */
			for (k = e; k < 0; k++) { if (t % 10) return OR_ERR_UNSUPPORTED; t /= 10; }
			e = 0;
		}
		for (k = 0; k < e; k++)
		{
			if (t > ((u128) I128_MAX) / 10) return OR_ERR_UNSUPPORTED;
			t *= 10;
		}
		if (v + t < v || v + t > (u128) I128_MAX) return OR_ERR_UNSUPPORTED;
		v += t;
	}
	put128(res, neg ? -(i128) v : (i128) v, dscale);
	return 0;
}

/* a constant of the plan: unscaled integer + display scale (gg_plan.h) */
void
or_numeric_const(int64_t unscaled, int dscale, or_datum *res)
{
	put128(res, (i128) unscaled, dscale);
}

int
or_numeric_func(int funcid, const or_datum *a, const or_datum *b, or_datum *res)
{
	i128 x = get128(a), y = get128(b), r;
	int sc, rc;

	switch (funcid)
	{
		case GG_F_NUMERIC_ADD: case GG_F_NUMERIC_SUB:
			sc = a->dscale > b->dscale ? a->dscale : b->dscale;
			if ((rc = rescale(&x, a->dscale, sc)) != 0 || (rc = rescale(&y, b->dscale, sc)) != 0) return rc;
			if (funcid == GG_F_NUMERIC_SUB) { if (y == -I128_MAX - 1) return OR_ERR_UNSUPPORTED; y = -y; }
			if (add_chk(x, y, &r)) return OR_ERR_UNSUPPORTED;
			put128(res, r, sc);
			return 0;
		case GG_F_NUMERIC_MUL:
			if (mul_chk(x, y, &r)) return OR_ERR_UNSUPPORTED;
			put128(res, r, a->dscale + b->dscale);
			return 0;
		case GG_F_NUMERIC_EQ: case GG_F_NUMERIC_NE: case GG_F_NUMERIC_LT:
		case GG_F_NUMERIC_LE: case GG_F_NUMERIC_GT: case GG_F_NUMERIC_GE:
		{
			int c;

			sc = a->dscale > b->dscale ? a->dscale : b->dscale;
			if ((rc = rescale(&x, a->dscale, sc)) != 0 || (rc = rescale(&y, b->dscale, sc)) != 0) return rc;
			c = (x > y) - (x < y);
			res->isnull = 0; res->len = 0; res->ptr = NULL; res->hi = 0; res->dscale = 0;
			res->v = funcid == GG_F_NUMERIC_EQ ? c == 0 : funcid == GG_F_NUMERIC_NE ? c != 0 : funcid == GG_F_NUMERIC_LT ? c < 0 :
				funcid == GG_F_NUMERIC_LE ? c <= 0 : funcid == GG_F_NUMERIC_GT ? c > 0 : c >= 0;
			return 0;
		}
	}
	return OR_ERR_UNSUPPORTED;
}

/* numeric_avg_accum: sum += x (the sum's display scale follows the largest input's), N += 1 */
int
or_numeric_accum(int64_t *sum_lo, int64_t *sum_hi, int *sum_dscale, int64_t *n, const or_datum *x)
{
	i128 s = (i128) (((u128) (uint64_t) *sum_hi << 64) | (u128) (uint64_t) *sum_lo), v = get128(x), r;
	int sc = *sum_dscale > x->dscale ? *sum_dscale : x->dscale, rc;

	if ((rc = rescale(&s, *sum_dscale, sc)) != 0 || (rc = rescale(&v, x->dscale, sc)) != 0) return rc;
	if (add_chk(s, v, &r)) return OR_ERR_UNSUPPORTED;
	*sum_lo = (int64_t) (uint64_t) (u128) r;
	*sum_hi = (int64_t) (uint64_t) ((u128) r >> 64);
	*sum_dscale = sc;
	(*n)++;
	return 0;
}

/* weight and first digit of a non-negative magnitude / 10^dscale in base 10000 (a NumericVar's after strip_var) */
static void
weight_first(u128 mag, int dscale, int *weight, int *first)
{
	char dec[64];
	int nd = 0, i, pos;

	*weight = 0; *first = 0;
	if (mag == 0) return;
	while (mag) { dec[nd++] = (char) (mag % 10); mag /= 10; }			/* least significant first */
	while (nd <= dscale) dec[nd++] = 0;									/* at least one integer digit position */
	/* dec[i] sits at decimal exponent i - dscale; the base-10000 digit of exponent e has weight floor(e / 4) */
	for (i = nd - 1; i >= 0; i--)
		if (dec[i]) break;
	pos = i - dscale;													/* decimal exponent of the first non-zero digit */
	*weight = pos >= 0 ? pos / 4 : -((-pos + 3) / 4);
	{
		/* the base-10000 digit at that weight: decimal exponents 4 * weight + 3 .. 4 * weight */
		int e, val = 0;

		for (e = 4 * *weight + 3; e >= 4 * *weight; e--)
		{
			int idx = e + dscale;										/* index in dec[] */

			val = val * 10 + ((idx >= 0 && idx < nd) ? dec[idx] : 0);
		}
		*first = val;
	}
}

/* numeric_div(sum, N) as numeric_avg computes it: *out at display scale *rscale */
int
or_numeric_avg(int64_t sum_lo, int64_t sum_hi, int sum_dscale, int64_t n, int64_t *out_lo, int64_t *out_hi, int *rscale)
{
	i128 s = (i128) (((u128) (uint64_t) sum_hi << 64) | (u128) (uint64_t) sum_lo);
	int neg = s < 0, w1, f1, w2, f2, qweight, rs, i;
	u128 mag = neg ? (u128) (-s) : (u128) s, q, r;

	weight_first(mag, sum_dscale, &w1, &f1);
	weight_first((u128) n, 0, &w2, &f2);
	qweight = w1 - w2;
	if (f1 <= f2) qweight--;
	rs = 16 - qweight * 4;												/* NUMERIC_MIN_SIG_DIGITS, DEC_DIGITS */
	if (rs < sum_dscale) rs = sum_dscale;
	if (rs < 0) rs = 0;
	if (rs > 1000) rs = 1000;
	for (i = sum_dscale; i < rs; i++)
	{
		if (mag > (~(u128) 0) / 10) return OR_ERR_UNSUPPORTED;
		mag *= 10;
	}
	q = mag / (u128) n;
	r = mag % (u128) n;
	if (r * 2 >= (u128) n) q++;											/* round_var: half away from zero */
	if (q > (u128) I128_MAX) return OR_ERR_UNSUPPORTED;
	s = neg ? -(i128) q : (i128) q;
	*out_lo = (int64_t) (uint64_t) (u128) s;
	*out_hi = (int64_t) (uint64_t) ((u128) s >> 64);
	*rscale = rs;
	return 0;
}
