/*
 * gg_compile.cpp — host-side plan compiler: gg_plan.h PODs -> ggp_program.
 *
 * Stands where ExecInitExpr / ExecInitAgg prepare ExprState trees and per-agg
 * state (execQual.c:5378, nodeAgg.c:1875-2300): the plan is checked for
 * eligibility (anything outside the accelerated subset returns
 * GG_ERR_UNSUPPORTED so the caller keeps the CPU node), referenced columns get
 * slots, attcacheoff is precomputed the way slot_deform_tuple memoises it
 * (heaptuple.c:1160-1190), and every expression is flattened to accumulator
 * code (gg_program.h).
 */
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include "gg_program.h"
#include "../../include/ggb200.h"

namespace {

struct Ctx {
	const gg_exprpool *pool;
	ggp_program *prog;
	ggp_side *outer;
	ggp_side *inner;          /* may be null */
	const gg_tupdesc *odesc;
	const gg_tupdesc *idesc;
	char *err;
	int errlen;
	int temps_used;
	bool failed;
};

void fail(Ctx &c, const char *fmt, ...)
{
	if (c.failed) return;
	c.failed = true;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(c.err, (size_t) c.errlen, fmt, ap);
	va_end(ap);
}

int loadtype_of(int32_t typid)
{
	switch (typid)
	{
		case GG_INT4OID: case GG_DATEOID: return GGP_LD_I4;
		case GG_INT8OID: case GG_FLOAT8OID: case GG_TIMESTAMPOID: return GGP_LD_I8;
		case GG_BPCHAROID: return GGP_LD_BPCHAR;
		case GG_VARCHAROID: case GG_TEXTOID: return GGP_LD_VARCHAR;
		case GG_BOOLOID: return GGP_LD_BOOL;
	}
	return 0;
}

/* slot of (varno, attno), allocating on first use */
int col_slot(Ctx &c, int varno, int attno)
{
	ggp_side *s = varno == 1 ? c.inner : c.outer;
	const gg_tupdesc *d = varno == 1 ? c.idesc : c.odesc;

	if (!s || !d) { fail(c, "Var references a side that does not exist (varno %d)", varno); return 0; }
	if (attno < 1 || attno > d->natts) { fail(c, "Var attno %d out of range", attno); return 0; }
	int a = attno - 1;
	if (s->att[a].slot >= 0) return s->att[a].slot;
	int lt = loadtype_of(d->attrs[a].atttypid);
	if (!lt) { fail(c, "column %d: type %d not supported on the GPU path", attno, d->attrs[a].atttypid); return 0; }
	if (s->ncols >= GGP_MAX_COLS) { fail(c, "too many referenced columns"); return 0; }
	int slot = s->ncols++;
	s->att[a].slot = (int8_t) slot;
	s->coltype[slot] = (uint8_t) lt;
	s->colatt[slot] = (uint8_t) a;
	if (a + 1 > s->natts_walk) s->natts_walk = a + 1;
	if (!d->attrs[a].attnotnull) c.prog->nullable = 1;
	return slot;
}

int add_const(Ctx &c, int64_t v, bool isnull)
{
	ggp_program *p = c.prog;
	for (int i = 0; i < p->nconst; i++)
		if (p->consts[i] == v && (((p->constnull >> i) & 1) != 0) == isnull) return i;
	if (p->nconst >= GGP_MAX_CONSTS) { fail(c, "too many constants"); return 0; }
	p->consts[p->nconst] = v;
	if (isnull) { p->constnull |= 1 << p->nconst; p->nullable = 1; }
	return p->nconst++;
}

void emit(Ctx &c, int op, int src, int idx, int aux = 0)
{
	ggp_program *p = c.prog;
	if (p->ncode >= GGP_MAX_CODE) { fail(c, "expression program too long"); return; }
	ggp_op o;
	o.op = (uint8_t) op; o.src = (uint8_t) src; o.idx = (uint8_t) idx; o.aux = (uint8_t) aux;
	p->code[p->ncode++] = o;
}

bool is_leaf(const gg_expr &e) { return e.kind == GG_E_VAR || e.kind == GG_E_CONST; }

void leaf_operand(Ctx &c, const gg_expr &e, int *src, int *idx)
{
	if (e.kind == GG_E_VAR)
	{
		*src = e.varno == 1 ? GGP_SRC_ICOL : GGP_SRC_COL;
		*idx = col_slot(c, e.varno, e.varattno);
	}
	else
	{
		*src = GGP_SRC_CONST;
		*idx = add_const(c, e.constvalue, e.constisnull != 0);
	}
}

int swap_cc(int cc)
{
	switch (cc) { case GGP_LT: return GGP_GT; case GGP_LE: return GGP_GE; case GGP_GT: return GGP_LT; case GGP_GE: return GGP_LE; }
	return cc;
}

struct BinInfo { int op, rop, cc; bool isbin; bool commut; };

/* map a pg_proc OID to machine ops */
bool func_info(int funcid, BinInfo *b, int *unary_op)
{
	*unary_op = 0;
	b->isbin = true; b->commut = false; b->cc = 0; b->rop = 0;
	switch (funcid)
	{
		case GG_F_FLOAT8PL:  b->op = GGP_F8ADD; b->rop = GGP_F8ADD; b->commut = true; return true;
		case GG_F_FLOAT8MUL: b->op = GGP_F8MUL; b->rop = GGP_F8MUL; b->commut = true; return true;
		case GG_F_FLOAT8MI:  b->op = GGP_F8SUB; b->rop = GGP_F8RSUB; return true;
		case GG_F_FLOAT8DIV: b->op = GGP_F8DIV; b->rop = GGP_F8RDIV; return true;
#define CMPCASE(F, OP, CC) case F: b->op = OP; b->rop = OP; b->cc = CC; return true;
		CMPCASE(GG_F_FLOAT8LT, GGP_CMPF8, GGP_LT) CMPCASE(GG_F_FLOAT8LE, GGP_CMPF8, GGP_LE)
		CMPCASE(GG_F_FLOAT8EQ, GGP_CMPF8, GGP_EQ) CMPCASE(GG_F_FLOAT8NE, GGP_CMPF8, GGP_NE)
		CMPCASE(GG_F_FLOAT8GT, GGP_CMPF8, GGP_GT) CMPCASE(GG_F_FLOAT8GE, GGP_CMPF8, GGP_GE)
		CMPCASE(GG_F_INT4LT, GGP_CMPI, GGP_LT) CMPCASE(GG_F_INT4LE, GGP_CMPI, GGP_LE)
		CMPCASE(GG_F_INT4EQ, GGP_CMPI, GGP_EQ) CMPCASE(GG_F_INT4NE, GGP_CMPI, GGP_NE)
		CMPCASE(GG_F_INT4GT, GGP_CMPI, GGP_GT) CMPCASE(GG_F_INT4GE, GGP_CMPI, GGP_GE)
		CMPCASE(GG_F_INT8LT, GGP_CMPI, GGP_LT) CMPCASE(GG_F_INT8LE, GGP_CMPI, GGP_LE)
		CMPCASE(GG_F_INT8EQ, GGP_CMPI, GGP_EQ) CMPCASE(GG_F_INT8NE, GGP_CMPI, GGP_NE)
		CMPCASE(GG_F_INT8GT, GGP_CMPI, GGP_GT) CMPCASE(GG_F_INT8GE, GGP_CMPI, GGP_GE)
		CMPCASE(GG_F_DATE_LT, GGP_CMPI, GGP_LT) CMPCASE(GG_F_DATE_LE, GGP_CMPI, GGP_LE)
		CMPCASE(GG_F_DATE_EQ, GGP_CMPI, GGP_EQ) CMPCASE(GG_F_DATE_NE, GGP_CMPI, GGP_NE)
		CMPCASE(GG_F_DATE_GT, GGP_CMPI, GGP_GT) CMPCASE(GG_F_DATE_GE, GGP_CMPI, GGP_GE)
		CMPCASE(GG_F_BPCHAREQ, GGP_CMPSTR, GGP_EQ) CMPCASE(GG_F_BPCHARNE, GGP_CMPSTR, GGP_NE)
#undef CMPCASE
		case GG_F_INT48: b->isbin = false; *unary_op = -1; return true;     /* loads already sign-extend */
		case GG_F_I4TOD: case GG_F_I8TOD: b->isbin = false; *unary_op = GGP_I2F8; return true;
	}
	return false;
}

int date_ts_cc(int funcid)
{
	switch (funcid)
	{
		case GG_F_DATE_LT_TIMESTAMP: return GGP_LT; case GG_F_DATE_LE_TIMESTAMP: return GGP_LE;
		case GG_F_DATE_EQ_TIMESTAMP: return GGP_EQ; case GG_F_DATE_GT_TIMESTAMP: return GGP_GT;
		case GG_F_DATE_GE_TIMESTAMP: return GGP_GE; case GG_F_DATE_NE_TIMESTAMP: return GGP_NE;
	}
	return -1;
}

void gen(Ctx &c, int root);

/* acc = acc OP right-subtree */
void gen_binary(Ctx &c, const gg_expr &e, int op, int rop, int cc, bool commut)
{
	const gg_expr &l = c.pool->nodes[e.args[0]];
	const gg_expr &r = c.pool->nodes[e.args[1]];
	int src, idx;

	if (is_leaf(r))
	{
		gen(c, e.args[0]);
		leaf_operand(c, r, &src, &idx);
		emit(c, op, src, idx, cc);
	}
	else if (is_leaf(l))
	{
		/* evaluate the complex right side, then apply with the leaf on the LEFT */
		gen(c, e.args[1]);
		leaf_operand(c, l, &src, &idx);
		if (commut) emit(c, op, src, idx, cc);
		else if (op == GGP_CMPF8 || op == GGP_CMPI || op == GGP_CMPSTR) emit(c, op, src, idx, swap_cc(cc));
		else emit(c, rop, src, idx, cc);
	}
	else
	{
		gen(c, e.args[0]);
		if (c.temps_used >= 4) { fail(c, "expression too deep"); return; }
		int t = c.temps_used++;
		emit(c, GGP_STORE, GGP_SRC_TEMP, t);
		gen(c, e.args[1]);
		/* acc = right, temp = left: need left OP right */
		if (commut) emit(c, op, GGP_SRC_TEMP, t, cc);
		else if (op == GGP_CMPF8 || op == GGP_CMPI || op == GGP_CMPSTR) emit(c, op, GGP_SRC_TEMP, t, swap_cc(cc));
		else emit(c, rop, GGP_SRC_TEMP, t, cc);
		c.temps_used--;
	}
}

void gen(Ctx &c, int root)
{
	if (c.failed) return;
	if (root < 0 || root >= c.pool->nnodes) { fail(c, "bad expression index %d", root); return; }
	const gg_expr &e = c.pool->nodes[root];
	int src, idx;

	switch (e.kind)
	{
		case GG_E_VAR:
		case GG_E_CONST:
			leaf_operand(c, e, &src, &idx);
			emit(c, GGP_LOAD, src, idx);
			return;
		case GG_E_FUNC:
		{
			int cc = date_ts_cc(e.funcid);
			if (cc >= 0)
			{
				/* date_xx_timestamp(date, ts): promote the date, then integer compare (date.c:560-640) */
				gen(c, e.args[0]);
				emit(c, GGP_DATE2TS, 0, 0);
				const gg_expr &r = c.pool->nodes[e.args[1]];
				if (is_leaf(r)) { leaf_operand(c, r, &src, &idx); emit(c, GGP_CMPI, src, idx, cc); }
				else
				{
					if (c.temps_used >= 4) { fail(c, "expression too deep"); return; }
					int t = c.temps_used++;
					emit(c, GGP_STORE, GGP_SRC_TEMP, t);
					gen(c, e.args[1]);
					emit(c, GGP_CMPI, GGP_SRC_TEMP, t, swap_cc(cc));
					c.temps_used--;
				}
				return;
			}
			BinInfo b; int un;
			if (!func_info(e.funcid, &b, &un)) { fail(c, "function %d not supported on the GPU path", e.funcid); return; }
			if (!b.isbin)
			{
				gen(c, e.args[0]);
				if (un > 0) emit(c, un, 0, 0);
				return;
			}
			gen_binary(c, e, b.op, b.rop, b.cc, b.commut);
			return;
		}
		case GG_E_AND:
		case GG_E_OR:
			gen_binary(c, e, e.kind == GG_E_AND ? GGP_AND : GGP_OR, 0, 0, true);
			return;
		case GG_E_NOT: gen(c, e.args[0]); emit(c, GGP_NOT, 0, 0); return;
		case GG_E_ISNULL: gen(c, e.args[0]); emit(c, GGP_ISNULL, 0, 0); c.prog->nullable = 1; return;
		case GG_E_ISNOTNULL: gen(c, e.args[0]); emit(c, GGP_ISNOTNULL, 0, 0); c.prog->nullable = 1; return;
	}
	fail(c, "expression kind %d not supported", e.kind);
}

ggp_span gen_span(Ctx &c, int root)
{
	ggp_span s;
	s.start = (int16_t) c.prog->ncode;
	c.temps_used = 0;
	gen(c, root);
	s.len = (int16_t) (c.prog->ncode - s.start);
	return s;
}

bool expr_equal(const gg_exprpool *pool, int a, int b)
{
	if (a == b) return true;
	if (a < 0 || b < 0) return false;
	const gg_expr &x = pool->nodes[a], &y = pool->nodes[b];
	if (x.kind != y.kind) return false;
	switch (x.kind)
	{
		case GG_E_VAR: return x.varno == y.varno && x.varattno == y.varattno;
		case GG_E_CONST: return x.constvalue == y.constvalue && x.constisnull == y.constisnull && x.constlen == y.constlen;
		case GG_E_FUNC:
			if (x.funcid != y.funcid || x.nargs != y.nargs) return false;
			for (int i = 0; i < x.nargs; i++) if (!expr_equal(pool, x.args[i], y.args[i])) return false;
			return true;
		case GG_E_NOT: case GG_E_ISNULL: case GG_E_ISNOTNULL: return expr_equal(pool, x.args[0], y.args[0]);
		default: return expr_equal(pool, x.args[0], y.args[0]) && expr_equal(pool, x.args[1], y.args[1]);
	}
}

void init_side(ggp_side *s, const gg_tupdesc *d)
{
	memset(s, 0, sizeof *s);
	s->natts = d->natts;
	for (int i = 0; i < d->natts && i < GG_MAX_ATTS; i++)
	{
		s->att[i].attlen = d->attrs[i].attlen;
		s->att[i].attalign = d->attrs[i].attalign;
		s->att[i].slot = -1;
		s->att[i].cacheoff = -1;
		s->att[i].notnull = d->attrs[i].attnotnull;
	}
}

long align_nominal(long off, char a)
{
	switch (a) { case 'i': return (off + 3) & ~3L; case 'c': return off; case 'd': return (off + 7) & ~7L; default: return (off + 1) & ~1L; }
}

/* attcacheoff as slot_deform_tuple memoises it (heaptuple.c:1160-1200): valid for the fixed-width
 * prefix, and for the first varlena iff its offset is already aligned */
void finish_side(ggp_side *s)
{
	long off = 0;
	int i;
	s->first_walk = s->natts;
	for (i = 0; i < s->natts; i++)
	{
		if (s->att[i].attlen == -1)
		{
			if (off == align_nominal(off, s->att[i].attalign)) s->att[i].cacheoff = (int16_t) off;
			else { s->first_walk = i; break; }
			s->first_walk = i + 1;      /* offset of the varlena itself is constant; what follows is not */
			break;
		}
		off = align_nominal(off, s->att[i].attalign);
		s->att[i].cacheoff = (int16_t) off;
		off += s->att[i].attlen;
	}
	if (s->first_walk > s->natts) s->first_walk = s->natts;
}

}  // namespace

int ggp_compile_scanagg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool,
                        ggp_program *prog, ggp_aggmap *aggmap, char *err, int errlen)
{
	Ctx c;
	memset(prog, 0, sizeof *prog);
	c.pool = pool; c.prog = prog; c.outer = &prog->outer; c.inner = nullptr;
	c.odesc = &scan->desc; c.idesc = nullptr; c.err = err; c.errlen = errlen;
	c.temps_used = 0; c.failed = false;
	if (err && errlen) err[0] = 0;

	if (scan->desc.natts < 0 || scan->desc.natts > GG_MAX_ATTS) { fail(c, "too many attributes"); return GG_ERR_UNSUPPORTED; }
	for (int i = 0; i < scan->desc.natts; i++)
	{
		const gg_attr &a = scan->desc.attrs[i];
		if (!(a.attlen == -1 || a.attlen == 1 || a.attlen == 2 || a.attlen == 4 || a.attlen == 8))
		{ fail(c, "attribute %d: attlen %d not supported", i + 1, a.attlen); return GG_ERR_UNSUPPORTED; }
	}
	init_side(&prog->outer, &scan->desc);

	if (scan->qual >= 0) prog->qual = gen_span(c, scan->qual);
	if (agg->numCols < 0 || agg->numCols > GG_MAX_KEYS) fail(c, "too many grouping columns");
	prog->nkeys = agg->numCols;
	for (int i = 0; i < agg->numCols && !c.failed; i++)
	{
		prog->key[i] = gen_span(c, agg->grpCol[i]);
		int32_t t = pool->nodes[agg->grpCol[i]].rettype;
		switch (t)
		{
			case GG_INT4OID: case GG_INT8OID: case GG_DATEOID: case GG_TIMESTAMPOID: case GG_BOOLOID: prog->keytype[i] = 1; break;
			case GG_FLOAT8OID: prog->keytype[i] = 2; break;
			case GG_BPCHAROID: case GG_VARCHAROID: case GG_TEXTOID: prog->keytype[i] = 3; break;
			default: fail(c, "grouping column type %d not supported", t);
		}
	}
	if (agg->numAggs < 0 || agg->numAggs > GG_MAX_AGGS) fail(c, "too many aggregates");
	if (agg->aggstage == GG_AGGSTAGE_FINAL) fail(c, "FINAL stage runs through gg_agg_final");

	/* aggregate arguments -> deduplicated accumulator columns */
	int accroot[GGP_MAX_ACCS];
	for (int i = 0; i < agg->numAggs && !c.failed; i++)
	{
		const gg_aggref &ar = agg->aggs[i];
		int kind = 0, sq = 0;
		switch (ar.aggfnoid)
		{
			case GG_AGG_COUNT_STAR: aggmap[i].col = -1; continue;
			case GG_AGG_COUNT_ANY: kind = GGP_ACC_COUNT; break;
			case GG_AGG_SUM_FLOAT8: kind = GGP_ACC_F8SUM; break;
			case GG_AGG_AVG_FLOAT8: kind = GGP_ACC_F8SUM; sq = 1; break;
			case GG_AGG_MIN_FLOAT8: kind = GGP_ACC_F8MIN; break;
			case GG_AGG_MAX_FLOAT8: kind = GGP_ACC_F8MAX; break;
			case GG_AGG_SUM_INT4: kind = GGP_ACC_I8SUM; break;
			case GG_AGG_MIN_INT4: case GG_AGG_MIN_INT8: case GG_AGG_MIN_DATE: kind = GGP_ACC_I8MIN; break;
			case GG_AGG_MAX_INT4: case GG_AGG_MAX_INT8: case GG_AGG_MAX_DATE: kind = GGP_ACC_I8MAX; break;
			default: fail(c, "aggregate %d not supported on the GPU path", ar.aggfnoid); continue;
		}
		if (ar.arg < 0) { fail(c, "aggregate %d needs an argument", ar.aggfnoid); continue; }
		int found = -1;
		for (int j = 0; j < prog->nacc; j++)
		{
			bool compat = prog->acckind[j] == kind ||
				(kind == GGP_ACC_COUNT && prog->acckind[j] != GGP_ACC_COUNT) ;
			if (compat && expr_equal(pool, accroot[j], ar.arg)) { found = j; break; }
		}
		if (found < 0)
		{
			if (prog->nacc >= GGP_MAX_ACCS) { fail(c, "too many distinct aggregate arguments"); continue; }
			found = prog->nacc++;
			accroot[found] = ar.arg;
			prog->acckind[found] = (uint8_t) kind;
			prog->acc[found] = gen_span(c, ar.arg);
		}
		if (sq) prog->accsq[found] = 1;
		aggmap[i].col = found;
	}
	finish_side(&prog->outer);
	if (c.failed) return GG_ERR_UNSUPPORTED;
	return GG_OK;
}
