/*
 * refwrap.c — ORACLE build infrastructure.  Thin exports over the REFERENCE'S OWN
 * leaf objects (compiled in place from /root/reference by ./Makefile) so that
 * tests/golden/make_golden.py and the oracle tests can call them through ctypes.
 * No algorithm lives here: every function fills an fmgr call frame or a
 * TupleDesc by hand and calls the reference symbol named in its comment.
 */
#include "postgres.h"
#include <setjmp.h>
#include "fmgr.h"
#include "access/htup_details.h"
#include "access/tupdesc.h"
#include "access/hash.h"
#include "catalog/pg_attribute.h"
#include "catalog/pg_type.h"
#include "cdb/cdbhash.h"
#include "utils/array.h"
#include "utils/builtins.h"
#include "utils/date.h"
#include "utils/int8.h"
#include "utils/timestamp.h"

extern sigjmp_buf *ref_err_jmp;		/* shim.c: where ereport(ERROR) lands */
extern char ref_err_msg[256];

#define REF_TRY(errvar) \
	sigjmp_buf _jb; sigjmp_buf *_save = ref_err_jmp; *(errvar) = 0; \
	ref_err_jmp = &_jb; \
	if (sigsetjmp(_jb, 0) != 0) { ref_err_jmp = _save; *(errvar) = 1; } else
#define REF_END() ref_err_jmp = _save

static Datum
call1(PGFunction fn, Datum a, bool *isnull)
{
	FunctionCallInfoData fcinfo;

	InitFunctionCallInfoData(fcinfo, NULL, 1, InvalidOid, NULL, NULL);
	fcinfo.arg[0] = a; fcinfo.argnull[0] = false;
	{
		Datum r = fn(&fcinfo);

		if (isnull) *isnull = fcinfo.isnull;
		return r;
	}
}

static Datum
call2(PGFunction fn, Datum a, Datum b, bool *isnull)
{
	FunctionCallInfoData fcinfo;

	InitFunctionCallInfoData(fcinfo, NULL, 2, InvalidOid, NULL, NULL);
	fcinfo.arg[0] = a; fcinfo.argnull[0] = false;
	fcinfo.arg[1] = b; fcinfo.argnull[1] = false;
	{
		Datum r = fn(&fcinfo);

		if (isnull) *isnull = fcinfo.isnull;
		return r;
	}
}

/* ---- hashfunc.c ---- */
uint32 ref_hash_any(const unsigned char *k, int len) { return DatumGetUInt32(hash_any(k, len)); }
uint32 ref_hash_uint32(uint32 k) { return DatumGetUInt32(hash_uint32(k)); }
uint32 ref_hashint4(int32 v) { return DatumGetUInt32(call1(hashint4, Int32GetDatum(v), NULL)); }
uint32 ref_hashint8(int64 v) { return DatumGetUInt32(call1(hashint8, Int64GetDatum(v), NULL)); }
uint32 ref_hashfloat8(double v) { return DatumGetUInt32(call1(hashfloat8, Float8GetDatum(v), NULL)); }

static struct varlena *
mk_varlena(const char *payload, int len)
{
	struct varlena *v = (struct varlena *) malloc(len + VARHDRSZ);

	SET_VARSIZE(v, len + VARHDRSZ);
	memcpy(VARDATA(v), payload, len);
	return v;
}

/* varchar.c hashbpchar / bpchareq */
uint32
ref_hashbpchar(const char *payload, int len)
{
	struct varlena *v = mk_varlena(payload, len);
	uint32 r = DatumGetUInt32(call1(hashbpchar, PointerGetDatum(v), NULL));

	free(v);
	return r;
}

int
ref_bpchareq(const char *a, int la, const char *b, int lb)
{
	struct varlena *x = mk_varlena(a, la), *y = mk_varlena(b, lb);
	int r = DatumGetBool(call2(bpchareq, PointerGetDatum(x), PointerGetDatum(y), NULL));

	free(x); free(y);
	return r;
}

/* ---- cdbhash.c: cdbhashinit / cdbhash / cdbhashreduce (jump_consistent_hash is static inside) ---- */
int
ref_cdbhash_route(const int32 *typids, const int64 *vals, const int32 *lens, const int32 *isnull,
				  int nkeys, int nsegs)
{
	CdbHash h;
	FmgrInfo fi[8];
	struct varlena *tofree[8];
	int i, seg;

	memset(&h, 0, sizeof h);
	memset(fi, 0, sizeof fi);
	h.numsegs = nsegs;
	h.reducealg = REDUCE_JUMP_HASH;
	h.is_legacy_hash = false;
	h.natts = nkeys;
	h.hashfuncs = fi;
	cdbhashinit(&h);
	for (i = 0; i < nkeys; i++)
	{
		Datum d = (Datum) vals[i];

		tofree[i] = NULL;
		fi[i].fn_nargs = 1;
		fi[i].fn_strict = true;
		switch (typids[i])
		{
			case INT4OID: case DATEOID: fi[i].fn_addr = hashint4; d = Int32GetDatum((int32) vals[i]); break;
			case INT8OID: fi[i].fn_addr = hashint8; break;
			case FLOAT8OID: fi[i].fn_addr = hashfloat8; break;
			case BPCHAROID:
				fi[i].fn_addr = hashbpchar;
				tofree[i] = mk_varlena((const char *) &vals[i], lens[i]);
				d = PointerGetDatum(tofree[i]);
				break;
			default: return -1;
		}
		cdbhash(&h, i + 1, d, isnull[i] != 0);
	}
	seg = (int) cdbhashreduce(&h);
	for (i = 0; i < nkeys; i++)
		if (tofree[i]) free(tofree[i]);
	return seg;
}

/* ---- heaptuple.c: heap_form_tuple / heap_deform_tuple ---- */
typedef struct ref_attr { int32 atttypid; int32 atttypmod; int16 attlen; int8 attalign; int8 attbyval; int8 attnotnull; int8 pad[3]; } ref_attr;

static TupleDesc
mk_desc(int natts, const ref_attr *a)
{
	TupleDesc d = (TupleDesc) calloc(1, sizeof(struct tupleDesc));
	int i;

	d->natts = natts;
	d->attrs = (Form_pg_attribute *) calloc(natts, sizeof(Form_pg_attribute));
	d->tdtypeid = RECORDOID;
	d->tdtypmod = -1;
	d->tdrefcount = -1;
	for (i = 0; i < natts; i++)
	{
		Form_pg_attribute att = (Form_pg_attribute) calloc(1, ATTRIBUTE_FIXED_PART_SIZE);

		att->atttypid = a[i].atttypid;
		att->attlen = a[i].attlen;
		att->attnum = i + 1;
		att->attcacheoff = -1;
		att->atttypmod = a[i].atttypmod;
		att->attbyval = a[i].attbyval;
		att->attalign = a[i].attalign;
		att->attstorage = (a[i].attlen == -1) ? 'x' : 'p';	/* packable varlenas, as bpchar/varchar are */
		att->attnotnull = a[i].attnotnull;
		d->attrs[i] = att;
	}
	return d;
}

static void
free_desc(TupleDesc d)
{
	int i;

	for (i = 0; i < d->natts; i++) free(d->attrs[i]);
	free(d->attrs);
	free(d);
}

/* vals[]: Datum bits, or pointer to payload for varlenas (+lens). Returns tuple length; bytes in out. */
int
ref_heap_form_tuple(int natts, const ref_attr *a, const int64 *vals, const int32 *lens,
					const uint8 *isnull, uint8 *out, int outcap)
{
	TupleDesc d = mk_desc(natts, a);
	Datum values[64];
	bool nulls[64];
	struct varlena *tofree[64];
	HeapTuple tup;
	int i, len;

	for (i = 0; i < natts; i++)
	{
		tofree[i] = NULL;
		nulls[i] = isnull && isnull[i];
		values[i] = (Datum) vals[i];
		if (!nulls[i] && a[i].attlen == -1)
		{
			tofree[i] = mk_varlena((const char *) (uintptr_t) vals[i], lens[i]);
			values[i] = PointerGetDatum(tofree[i]);
		}
		else if (a[i].attlen == 4)
			values[i] = Int32GetDatum((int32) vals[i]);
	}
	tup = heap_form_tuple(d, values, nulls);
	len = (int) tup->t_len;
	if (len <= outcap)
		memcpy(out, tup->t_data, len);
	free(tup);
	for (i = 0; i < natts; i++)
		if (tofree[i]) free(tofree[i]);
	free_desc(d);
	return len;
}

/* heap_deform_tuple: values[] = Datum bits or, for varlenas, byte offset from tuple start */
int
ref_heap_deform_tuple(int natts, const ref_attr *a, uint8 *tuple, int tuplen, int64 *vals, uint8 *isnull)
{
	TupleDesc d = mk_desc(natts, a);
	HeapTupleData htup;
	Datum values[64];
	bool nulls[64];
	int i;

	memset(&htup, 0, sizeof htup);
	htup.t_len = tuplen;
	htup.t_data = (HeapTupleHeader) tuple;
	heap_deform_tuple(&htup, d, values, nulls);
	for (i = 0; i < natts; i++)
	{
		isnull[i] = nulls[i];
		if (nulls[i])
			vals[i] = 0;
		else if (a[i].attlen == -1)
			vals[i] = (int64) ((uint8 *) DatumGetPointer(values[i]) - tuple);
		else if (a[i].attlen == 4)
			vals[i] = (int64) DatumGetInt32(values[i]);
		else
			vals[i] = (int64) values[i];
	}
	free_desc(d);
	return natts;
}

/* ---- float.c ---- */
static double
f8op(PGFunction fn, double a, double b, int *err)
{
	volatile double r = 0;
	REF_TRY(err) { r = DatumGetFloat8(call2(fn, Float8GetDatum(a), Float8GetDatum(b), NULL)); REF_END(); }
	return r;
}
double ref_float8pl(double a, double b, int *err) { return f8op(float8pl, a, b, err); }
double ref_float8mi(double a, double b, int *err) { return f8op(float8mi, a, b, err); }
double ref_float8mul(double a, double b, int *err) { return f8op(float8mul, a, b, err); }
double ref_float8div(double a, double b, int *err) { return f8op(float8div, a, b, err); }

static int
f8cmp(PGFunction fn, double a, double b)
{
	return DatumGetBool(call2(fn, Float8GetDatum(a), Float8GetDatum(b), NULL));
}
int ref_float8eq(double a, double b) { return f8cmp(float8eq, a, b); }
int ref_float8lt(double a, double b) { return f8cmp(float8lt, a, b); }
int ref_float8le(double a, double b) { return f8cmp(float8le, a, b); }
int ref_btfloat8cmp(double a, double b) { return DatumGetInt32(call2(btfloat8cmp, Float8GetDatum(a), Float8GetDatum(b), NULL)); }

static ArrayType *
mk_f8array(const double *v, int n)
{
	int sz = ARR_OVERHEAD_NONULLS(1) + n * sizeof(float8);
	ArrayType *arr = (ArrayType *) calloc(1, sz);

	SET_VARSIZE(arr, sz);
	arr->ndim = 1;
	arr->dataoffset = 0;
	arr->elemtype = FLOAT8OID;
	ARR_DIMS(arr)[0] = n;
	ARR_LBOUND(arr)[0] = 1;
	memcpy(ARR_DATA_PTR(arr), v, n * sizeof(float8));
	return arr;
}

/* float8_accum (float.c:1878): state[3] updated in place */
void
ref_float8_accum(double *state, double x, int *err)
{
	ArrayType *arr = mk_f8array(state, 3);
	REF_TRY(err)
	{
		ArrayType *r = (ArrayType *) DatumGetPointer(call2(float8_accum, PointerGetDatum(arr), Float8GetDatum(x), NULL));

		memcpy(state, ARR_DATA_PTR(r), 3 * sizeof(float8));
		if (r != arr) free(r);
		REF_END();
	}
	free(arr);
}

/* float8_combine (float.c:1842) */
void
ref_float8_combine(double *s1, const double *s2, int *err)
{
	ArrayType *a = mk_f8array(s1, 3), *b = mk_f8array(s2, 3);
	REF_TRY(err)
	{
		ArrayType *r = (ArrayType *) DatumGetPointer(call2(float8_combine, PointerGetDatum(a), PointerGetDatum(b), NULL));

		memcpy(s1, ARR_DATA_PTR(r), 3 * sizeof(float8));
		REF_END();
	}
	free(a); free(b);
}

/* float8_avg (float.c:1982) */
double
ref_float8_avg(const double *state, int *isnull)
{
	ArrayType *a = mk_f8array(state, 3);
	bool n = false;
	double r = DatumGetFloat8(call1(float8_avg, PointerGetDatum(a), &n));

	*isnull = n;
	free(a);
	return n ? 0.0 : r;
}

/* ---- int8.c ---- */
int64
ref_int8inc(int64 v, int *err)
{
	volatile int64 r = 0;
	REF_TRY(err) { r = DatumGetInt64(call1(int8inc, Int64GetDatum(v), NULL)); REF_END(); }
	return r;
}

int64
ref_int8pl(int64 a, int64 b, int *err)
{
	volatile int64 r = 0;
	REF_TRY(err) { r = DatumGetInt64(call2(int8pl, Int64GetDatum(a), Int64GetDatum(b), NULL)); REF_END(); }
	return r;
}

/* ---- date.c ---- */
int
ref_date_cmp_timestamp(int op, int32 d, int64 ts, int *err)
{
	static const PGFunction fns[6] = { date_lt_timestamp, date_le_timestamp, date_eq_timestamp,
		date_gt_timestamp, date_ge_timestamp, date_ne_timestamp };
	volatile int r = 0;
	REF_TRY(err) { r = DatumGetBool(call2(fns[op], DateADTGetDatum(d), TimestampGetDatum(ts), NULL)); REF_END(); }
	return r;
}

const char *ref_last_error(void) { return ref_err_msg; }
