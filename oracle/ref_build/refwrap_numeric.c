/*
 * refwrap_numeric.c — ORACLE infrastructure: the reference's own numeric.o (utils/adt/numeric.c, compiled in place) behind
 * a plain C surface for tests/golden/make_golden.py: text <-> on-disk form (numeric_in / numeric_out, numeric.c:468,559),
 * numeric_add / numeric_sub / numeric_mul / numeric_div (:1659,1698,1735,1773), the comparison numeric_cmp (:1512).
 * sum(numeric) is a fold of numeric_add and avg(numeric) is numeric_div(sum, N::numeric) — what numeric_sum / numeric_avg
 * compute from their state (numeric.c:3173,3205).  Test infrastructure only.
 */
#include "postgres.h"
#include "fmgr.h"
#include "utils/builtins.h"
#include "utils/numeric.h"
#include <setjmp.h>

extern sigjmp_buf *ref_err_jmp;			/* shim.c: where ereport(ERROR) lands */

/* DirectFunctionCall*: numeric.c calls its own functions through them */
Datum
DirectFunctionCall1Coll(PGFunction func, Oid collation, Datum arg1)
{
	FunctionCallInfoData fcinfo;

	memset(&fcinfo, 0, sizeof fcinfo);
	fcinfo.nargs = 1;
	fcinfo.arg[0] = arg1;
	return (*func) (&fcinfo);
}

Datum
DirectFunctionCall2Coll(PGFunction func, Oid collation, Datum arg1, Datum arg2)
{
	FunctionCallInfoData fcinfo;

	memset(&fcinfo, 0, sizeof fcinfo);
	fcinfo.nargs = 2;
	fcinfo.arg[0] = arg1;
	fcinfo.arg[1] = arg2;
	return (*func) (&fcinfo);
}

Datum
DirectFunctionCall3Coll(PGFunction func, Oid collation, Datum arg1, Datum arg2, Datum arg3)
{
	FunctionCallInfoData fcinfo;

	memset(&fcinfo, 0, sizeof fcinfo);
	fcinfo.nargs = 3;
	fcinfo.arg[0] = arg1;
	fcinfo.arg[1] = arg2;
	fcinfo.arg[2] = arg3;
	return (*func) (&fcinfo);
}

static Datum
num_in(const char *s, int32 typmod)
{
	return DirectFunctionCall3Coll(numeric_in, 0, CStringGetDatum(s), ObjectIdGetDatum(0), Int32GetDatum(typmod));
}

/* text -> the varlena (4-byte header form) numeric_in builds, typmod applied (-1: none).  Returns its total length, -1 on ERROR */
int
ref_numeric_in(const char *s, int32 typmod, uint8 *out, int cap)
{
	sigjmp_buf jb, *saved = ref_err_jmp;
	int len = -1;

	ref_err_jmp = &jb;
	if (sigsetjmp(jb, 0) == 0)
	{
		struct varlena *v = (struct varlena *) DatumGetPointer(num_in(s, typmod));

		len = VARSIZE(v);
		if (len <= cap) memcpy(out, v, len);
	}
	ref_err_jmp = saved;
	return len;
}

/* on-disk numeric (4-byte header form) -> text */
int
ref_numeric_out(const uint8 *v, char *out, int cap)
{
	char *s = DatumGetCString(DirectFunctionCall1Coll(numeric_out, 0, PointerGetDatum(v)));
	int n = (int) strlen(s);

	if (n < cap) memcpy(out, s, n + 1);
	return n;
}

/* op: '+', '-', '*', '/' on two numerics given as text; result as text.  -1 on ERROR */
int
ref_numeric_binop(int op, const char *a, const char *b, char *out, int cap)
{
	sigjmp_buf jb, *saved = ref_err_jmp;
	int n = -1;

	ref_err_jmp = &jb;
	if (sigsetjmp(jb, 0) == 0)
	{
		Datum x = num_in(a, -1), y = num_in(b, -1), r;
		PGFunction f = op == '+' ? numeric_add : op == '-' ? numeric_sub : op == '*' ? numeric_mul : numeric_div;
		char *s;

		r = DirectFunctionCall2Coll(f, 0, x, y);
		s = DatumGetCString(DirectFunctionCall1Coll(numeric_out, 0, r));
		n = (int) strlen(s);
		if (n < cap) memcpy(out, s, n + 1);
	}
	ref_err_jmp = saved;
	return n;
}

int
ref_numeric_cmp(const char *a, const char *b)
{
	return DatumGetInt32(DirectFunctionCall2Coll(numeric_cmp, 0, num_in(a, -1), num_in(b, -1)));
}
