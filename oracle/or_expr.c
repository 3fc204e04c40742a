/*
 * or_expr.c — ORACLE (test infrastructure): restatement of the interpreted
 * expression evaluator for the operators that occur on the hot path.
 *
 *   ExecQual / ExecEvalAnd / Or / Not   src/backend/executor/execQual.c:6260,3373-3520
 *   ExecEvalScalarVar -> slot_getattr   src/backend/executor/execQual.c:690, executor/tuptable.h:324
 *   ExecEvalOper (strict fn => NULL)    src/backend/executor/execQual.c:2169-2260,2652
 *   float8pl/mi/mul/div + CHECKFLOATVAL src/backend/utils/adt/float.c:782-850, utils/float_utils.h:28
 *   float8_cmp_internal                 src/backend/utils/adt/float.c:964-988
 *   date_*_timestamp / date2timestamp   src/backend/utils/adt/date.c:457-482,560-640
 *   int4/int8/date comparisons          src/backend/utils/adt/int.c, int8.c, date.c
 *   bpchareq / bpcharne                 src/backend/utils/adt/varchar.c:702-750
 */
#include <math.h>
#include <string.h>
#include "gg_oracle.h"
#include "or_internal.h"

#define USECS_PER_DAY 86400000000LL

static inline double
as_f8(int64_t v)
{
	double d;

	memcpy(&d, &v, 8);
	return d;
}

static inline int64_t
f8_bits(double d)
{
	int64_t v;

	memcpy(&v, &d, 8);
	return v;
}

/* slot_getattr (tuptable.h:324): deform lazily up to the attribute asked for */
static void
row_getattr(or_row *r, int attno, or_datum *res)
{
	const gg_attr *att;

	if (r == NULL)						/* null-extended side of an outer join */
	{
		res->isnull = 1; res->v = 0; res->len = 0; res->ptr = NULL;
		return;
	}
	att = &r->desc->attrs[attno - 1];
	if (r->nvalid < attno)
	{
		or_heap_deform(r->desc, r->tuple, attno, r->values, r->isnull);
		r->nvalid = attno;
	}
	res->isnull = r->isnull[attno - 1];
	res->v = r->values[attno - 1];
	res->len = 0;
	res->ptr = NULL;
	if (!res->isnull && att->attlen == -1)
	{
		int len;

		res->ptr = or_varlena_payload(r->tuple + r->values[attno - 1], &len);
		res->len = len;
	}
}

static int
is_numeric_func(int funcid)
{
	return funcid >= GG_F_NUMERIC_EQ && funcid <= GG_F_NUMERIC_MUL;
}

/* float.c:964 */
static int
float8_cmp_internal(double a, double b)
{
	if (isnan(a))
		return isnan(b) ? 0 : 1;
	if (isnan(b))
		return -1;
	if (a > b)
		return 1;
	if (a < b)
		return -1;
	return 0;
}

/* date.c:457 date2timestamp; returns 0 or an error */
static int
date2timestamp(int32_t dateVal, int64_t *ts)
{
	if (dateVal == INT32_MIN)
		*ts = INT64_MIN;					/* TIMESTAMP_NOBEGIN */
	else if (dateVal == INT32_MAX)
		*ts = INT64_MAX;					/* TIMESTAMP_NOEND */
	else
	{
		int64_t result = (int64_t) ((uint64_t) (int64_t) dateVal * (uint64_t) USECS_PER_DAY);

		if (result / USECS_PER_DAY != dateVal)
			return OR_ERR_UNSUPPORTED;		/* "date out of range for timestamp" */
		*ts = result;
	}
	return 0;
}

static int
cmp_i64(int64_t a, int64_t b)
{
	return (a > b) - (a < b);
}

static void
str_of(const or_datum *d, const char **s, int *len)
{
	if (d->ptr)
	{
		*s = (const char *) d->ptr;
		*len = d->len;
	}
	else
	{
		*s = (const char *) &d->v;			/* packed constant */
		*len = d->len;
	}
}

static int
eval_func(int funcid, const or_datum *a, const or_datum *b, or_datum *res)
{
	double x, y, r;
	int c;

	res->isnull = 0;
	res->len = 0;
	res->ptr = NULL;
	switch (funcid)
	{
		/* float8 arithmetic, float.c:782-850 */
		case GG_F_FLOAT8PL:
			x = as_f8(a->v); y = as_f8(b->v); r = x + y;
			if (isinf(r) && !(isinf(x) || isinf(y)))
				return OR_ERR_FLOAT_OVERFLOW;
			res->v = f8_bits(r);
			return 0;
		case GG_F_FLOAT8MI:
			x = as_f8(a->v); y = as_f8(b->v); r = x - y;
			if (isinf(r) && !(isinf(x) || isinf(y)))
				return OR_ERR_FLOAT_OVERFLOW;
			res->v = f8_bits(r);
			return 0;
		case GG_F_FLOAT8MUL:
			x = as_f8(a->v); y = as_f8(b->v); r = x * y;
			if (isinf(r) && !(isinf(x) || isinf(y)))
				return OR_ERR_FLOAT_OVERFLOW;
			if (r == 0.0 && !(x == 0 || y == 0))
				return OR_ERR_FLOAT_UNDERFLOW;
			res->v = f8_bits(r);
			return 0;
		case GG_F_FLOAT8DIV:
			x = as_f8(a->v); y = as_f8(b->v);
			if (y == 0.0)
				return OR_ERR_DIV_ZERO;
			r = x / y;
			if (isinf(r) && !(isinf(x) || isinf(y)))
				return OR_ERR_FLOAT_OVERFLOW;
			if (r == 0.0 && !(x == 0))
				return OR_ERR_FLOAT_UNDERFLOW;
			res->v = f8_bits(r);
			return 0;
		case GG_F_FLOAT8EQ: case GG_F_FLOAT8NE: case GG_F_FLOAT8LT:
		case GG_F_FLOAT8LE: case GG_F_FLOAT8GT: case GG_F_FLOAT8GE:
			c = float8_cmp_internal(as_f8(a->v), as_f8(b->v));
			res->v = funcid == GG_F_FLOAT8EQ ? c == 0 : funcid == GG_F_FLOAT8NE ? c != 0 :
				funcid == GG_F_FLOAT8LT ? c < 0 : funcid == GG_F_FLOAT8LE ? c <= 0 :
				funcid == GG_F_FLOAT8GT ? c > 0 : c >= 0;
			return 0;
		/* int4 / date comparisons (int.c int4eq.., date.c date_eq..: plain integer compares) */
		case GG_F_INT4EQ: case GG_F_DATE_EQ: res->v = (int32_t) a->v == (int32_t) b->v; return 0;
		case GG_F_INT4NE: case GG_F_DATE_NE: res->v = (int32_t) a->v != (int32_t) b->v; return 0;
		case GG_F_INT4LT: case GG_F_DATE_LT: res->v = (int32_t) a->v < (int32_t) b->v; return 0;
		case GG_F_INT4LE: case GG_F_DATE_LE: res->v = (int32_t) a->v <= (int32_t) b->v; return 0;
		case GG_F_INT4GT: case GG_F_DATE_GT: res->v = (int32_t) a->v > (int32_t) b->v; return 0;
		case GG_F_INT4GE: case GG_F_DATE_GE: res->v = (int32_t) a->v >= (int32_t) b->v; return 0;
		case GG_F_INT8EQ: res->v = a->v == b->v; return 0;
		case GG_F_INT8NE: res->v = a->v != b->v; return 0;
		case GG_F_INT8LT: res->v = a->v < b->v; return 0;
		case GG_F_INT8LE: res->v = a->v <= b->v; return 0;
		case GG_F_INT8GT: res->v = a->v > b->v; return 0;
		case GG_F_INT8GE: res->v = a->v >= b->v; return 0;
		/* casts */
		case GG_F_INT48: res->v = (int64_t) (int32_t) a->v; return 0;			/* int8.c int48 */
		case GG_F_I4TOD: res->v = f8_bits((double) (int32_t) a->v); return 0;	/* float.c i4tod */
		case GG_F_I8TOD: res->v = f8_bits((double) a->v); return 0;				/* int8.c i8tod */
		/* date vs timestamp, date.c:560-640 */
		case GG_F_DATE_LT_TIMESTAMP: case GG_F_DATE_LE_TIMESTAMP: case GG_F_DATE_EQ_TIMESTAMP:
		case GG_F_DATE_GT_TIMESTAMP: case GG_F_DATE_GE_TIMESTAMP: case GG_F_DATE_NE_TIMESTAMP:
		{
			int64_t dt1;
			int rc = date2timestamp((int32_t) a->v, &dt1);

			if (rc)
				return rc;
			c = cmp_i64(dt1, b->v);
			res->v = funcid == GG_F_DATE_LT_TIMESTAMP ? c < 0 : funcid == GG_F_DATE_LE_TIMESTAMP ? c <= 0 :
				funcid == GG_F_DATE_EQ_TIMESTAMP ? c == 0 : funcid == GG_F_DATE_GT_TIMESTAMP ? c > 0 :
				funcid == GG_F_DATE_GE_TIMESTAMP ? c >= 0 : c != 0;
			return 0;
		}
		case GG_F_BPCHAREQ: case GG_F_BPCHARNE:
		{
			const char *s1, *s2;
			int l1, l2;

			str_of(a, &s1, &l1);
			str_of(b, &s2, &l2);
			c = or_bpchareq(s1, l1, s2, l2);
			res->v = funcid == GG_F_BPCHAREQ ? c : !c;
			return 0;
		}
	}
	return OR_ERR_UNSUPPORTED;
}

int
or_eval(const gg_exprpool *pool, int root, or_row *outer, or_row *inner, or_datum *res)
{
	const gg_expr *e = &pool->nodes[root];
	or_datum a, b;
	int rc;

	memset(&a, 0, sizeof a);
	memset(&b, 0, sizeof b);
	switch (e->kind)
	{
		case GG_E_VAR:
			row_getattr(e->varno == 1 ? inner : outer, e->varattno, res);
			if (e->rettype == GG_NUMERICOID && !res->isnull)
			{
				/* the on-disk digits become a value here (numeric.c:95-190) */
				const uint8_t *p = res->ptr;
				int len = res->len;

				return or_numeric_decode(p, len, res);
			}
			return 0;
		case GG_E_CONST:
			res->isnull = e->constisnull;
			res->v = e->constvalue;
			res->len = e->constlen;
			res->ptr = NULL;
			if (e->rettype == GG_NUMERICOID && !e->constisnull)
				or_numeric_const(e->constvalue, e->constlen, res);
			return 0;
		case GG_E_FUNC:
			/* ExecMakeFunctionResultNoSets: a strict function with any NULL argument yields NULL
			 * without being called (execQual.c:2215-2232) */
			if ((rc = or_eval(pool, e->args[0], outer, inner, &a)) != 0)
				return rc;
			if (e->nargs > 1 && (rc = or_eval(pool, e->args[1], outer, inner, &b)) != 0)
				return rc;
			if (a.isnull || (e->nargs > 1 && b.isnull))
			{
				res->isnull = 1;
				res->v = 0;
				res->len = 0;
				res->ptr = NULL;
				return 0;
			}
			if (is_numeric_func(e->funcid))
				return or_numeric_func(e->funcid, &a, &b, res);
			return eval_func(e->funcid, &a, &b, res);
		case GG_E_AND:
			/* ExecEvalAnd (execQual.c:3455): FALSE wins, else NULL if any NULL, else TRUE */
		{
			int anynull = 0;

			if ((rc = or_eval(pool, e->args[0], outer, inner, &a)) != 0)
				return rc;
			if (a.isnull)
				anynull = 1;
			else if (!a.v)
			{
				res->isnull = 0; res->v = 0; return 0;
			}
			if ((rc = or_eval(pool, e->args[1], outer, inner, &b)) != 0)
				return rc;
			if (b.isnull)
				anynull = 1;
			else if (!b.v)
			{
				res->isnull = 0; res->v = 0; return 0;
			}
			res->isnull = anynull;
			res->v = !anynull;
			return 0;
		}
		case GG_E_OR:
			/* ExecEvalOr (execQual.c:3404): TRUE wins, else NULL if any NULL, else FALSE */
		{
			int anynull = 0;

			if ((rc = or_eval(pool, e->args[0], outer, inner, &a)) != 0)
				return rc;
			if (a.isnull)
				anynull = 1;
			else if (a.v)
			{
				res->isnull = 0; res->v = 1; return 0;
			}
			if ((rc = or_eval(pool, e->args[1], outer, inner, &b)) != 0)
				return rc;
			if (b.isnull)
				anynull = 1;
			else if (b.v)
			{
				res->isnull = 0; res->v = 1; return 0;
			}
			res->isnull = anynull;
			res->v = 0;
			return 0;
		}
		case GG_E_NOT:
			if ((rc = or_eval(pool, e->args[0], outer, inner, &a)) != 0)
				return rc;
			res->isnull = a.isnull;
			res->v = a.isnull ? 0 : !a.v;
			return 0;
		case GG_E_ISNULL:
		case GG_E_ISNOTNULL:
			if ((rc = or_eval(pool, e->args[0], outer, inner, &a)) != 0)
				return rc;
			res->isnull = 0;
			res->v = (e->kind == GG_E_ISNULL) ? a.isnull : !a.isnull;
			return 0;
	}
	return OR_ERR_UNSUPPORTED;
}

const char *
or_strerror(int code)
{
	switch (code)
	{
		case 0: return "ok";
		case OR_ERR_FLOAT_OVERFLOW: return "value out of range: overflow";
		case OR_ERR_FLOAT_UNDERFLOW: return "value out of range: underflow";
		case OR_ERR_DIV_ZERO: return "division by zero";
		case OR_ERR_INT_OVERFLOW: return "bigint out of range";
		case OR_ERR_UNSUPPORTED: return "unsupported expression or type";
		case OR_ERR_VISIBILITY: return "tuple visibility needs clog/snapshot";
		case OR_ERR_NOMEM: return "out of memory / output capacity";
	}
	return "unknown";
}
