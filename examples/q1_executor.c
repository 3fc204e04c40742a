/*
 * q1_executor.c — TPC-H Q1 through the executor-node surface, in plain C (what the Postgres-side module of
 * INTEGRATION.md does after translating the plan): load a synthetic lineitem segment, build
 *     Sort <- Agg <- SeqScan
 * run it with GgExecInitNode / GgExecProcNode / GgExecEndNode and print the rows.
 *
 *   gcc -std=c11 -Iinclude examples/q1_executor.c -Lgreengage_b200 -lggexec -lggb200 -lgghost \
 *       -Wl,-rpath,$PWD/greengage_b200 -o q1_executor && ./q1_executor [rows]
 *
 * Without a CUDA device gg_engine_create fails and the program says so (there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gg_executor.h"
#include "gg_synth.h"

static int32_t add_node(gg_exprpool *p, gg_expr e) { p->nodes[p->nnodes] = e; return p->nnodes++; }

static int32_t var(gg_exprpool *p, int attno, int32_t typid)
{
	gg_expr e; memset(&e, 0, sizeof e);
	e.kind = GG_E_VAR; e.varattno = (int16_t) attno; e.rettype = typid;
	return add_node(p, e);
}

static int32_t constant(gg_exprpool *p, int32_t typid, int64_t bits)
{
	gg_expr e; memset(&e, 0, sizeof e);
	e.kind = GG_E_CONST; e.rettype = typid; e.constvalue = bits;
	return add_node(p, e);
}

static int32_t func2(gg_exprpool *p, int32_t funcid, int32_t rettype, int32_t a, int32_t b)
{
	gg_expr e; memset(&e, 0, sizeof e);
	e.kind = GG_E_FUNC; e.funcid = funcid; e.rettype = rettype; e.nargs = 2; e.args[0] = a; e.args[1] = b;
	return add_node(p, e);
}

int main(int argc, char **argv)
{
	const uint64_t rows = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000000;
	gg_engine *eng = NULL;
	int rc = gg_engine_create(0, &eng);
	if (rc != GG_OK) { fprintf(stderr, "gg_engine_create: %s\n", gg_last_error()); return 2; }

	/* the segment's relation: synthetic lineitem (narrow layout), generated on the host, loaded into HBM */
	gg_synth_spec spec; memset(&spec, 0, sizeof spec);
	spec.table = GG_TAB_LINEITEM_NARROW; spec.seed = 42; spec.ncand = rows; spec.norders = rows / 4 ? rows / 4 : 1; spec.nsegs = 1;
	uint64_t nblocks = 0, nrows = 0;
	if (gg_synth_measure(&spec, 8, &nblocks, &nrows) != 0) return 3;
	void *pages = malloc((size_t) nblocks * GG_BLCKSZ);
	uint64_t nb2 = 0, nr2 = 0;
	if (!pages || gg_synth_generate(&spec, 8, pages, nblocks, &nb2, &nr2) != 0) return 3;
	gg_relation *rel = NULL;
	if ((rc = gg_relation_create(eng, nblocks, &rel)) != GG_OK || (rc = gg_relation_load(rel, 0, pages, nblocks)) != GG_OK)
	{ fprintf(stderr, "relation: %s\n", gg_last_error()); return 2; }

	/* Q1: WHERE l_shipdate <= timestamp; GROUP BY l_returnflag, l_linestatus; sums, avgs, count; ORDER BY the keys.
	 * Narrow layout: 1 orderkey, 2 quantity, 3 extendedprice, 4 discount, 5 tax, 6 returnflag, 7 linestatus, 8 shipdate */
	static gg_exprpool pool;
	double one = 1.0; int64_t onebits; memcpy(&onebits, &one, 8);
	int32_t qty = var(&pool, 2, GG_FLOAT8OID), price = var(&pool, 3, GG_FLOAT8OID), disc = var(&pool, 4, GG_FLOAT8OID), tax = var(&pool, 5, GG_FLOAT8OID);
	int32_t flag = var(&pool, 6, GG_BPCHAROID), status = var(&pool, 7, GG_BPCHAROID), shipdate = var(&pool, 8, GG_DATEOID);
	int32_t cutoff = constant(&pool, GG_TIMESTAMPOID, (int64_t) (-396 - 90) * 86400000000LL);      /* 1998-12-01 - 90 days */
	int32_t qual = func2(&pool, GG_F_DATE_LE_TIMESTAMP, GG_BOOLOID, shipdate, cutoff);
	int32_t k1 = constant(&pool, GG_FLOAT8OID, onebits);
	int32_t disc_price = func2(&pool, GG_F_FLOAT8MUL, GG_FLOAT8OID, price, func2(&pool, GG_F_FLOAT8MI, GG_FLOAT8OID, k1, disc));
	int32_t charge = func2(&pool, GG_F_FLOAT8MUL, GG_FLOAT8OID,
	                       func2(&pool, GG_F_FLOAT8MUL, GG_FLOAT8OID, price, func2(&pool, GG_F_FLOAT8MI, GG_FLOAT8OID, k1, disc)),
	                       func2(&pool, GG_F_FLOAT8PL, GG_FLOAT8OID, k1, tax));

	static GgSeqScan scan; static GgAgg agg; static GgSort sort;
	scan.plan.type = T_GgSeqScan; scan.plan.qual = qual; scan.scanrelid = 0;
	if (gg_synth_tupdesc(GG_TAB_LINEITEM_NARROW, &scan.desc) != 0) return 3;
	agg.plan.type = T_GgAgg; agg.plan.qual = -1; agg.plan.lefttree = &scan.plan;
	agg.agg.aggstage = GG_AGGSTAGE_NORMAL; agg.agg.numCols = 2; agg.agg.grpCol[0] = flag; agg.agg.grpCol[1] = status;
	const gg_aggref aggs[8] = { { GG_AGG_SUM_FLOAT8, qty }, { GG_AGG_SUM_FLOAT8, price }, { GG_AGG_SUM_FLOAT8, disc_price }, { GG_AGG_SUM_FLOAT8, charge },
	                            { GG_AGG_AVG_FLOAT8, qty }, { GG_AGG_AVG_FLOAT8, price }, { GG_AGG_AVG_FLOAT8, disc }, { GG_AGG_COUNT_STAR, -1 } };
	agg.agg.numAggs = 8; memcpy(agg.agg.aggs, aggs, sizeof aggs);
	sort.plan.type = T_GgSort; sort.plan.qual = -1; sort.plan.lefttree = &agg.plan; sort.numCols = 2;
	sort.keys[0].col = 0; sort.keys[0].typid = GG_BPCHAROID; sort.keys[1].col = 1; sort.keys[1].typid = GG_BPCHAROID;

	GgEState es; memset(&es, 0, sizeof es);
	es.engine = eng; es.pool = &pool; es.relations[0] = rel; es.nsegs = 1;
	GgPlanState *ps = GgExecInitNode(&sort.plan, &es, 0);
	if (!ps) { fprintf(stderr, "ExecInitNode: %s\n", GgExecLastError()); return 2; }
	GgTupleTableSlot *slot;
	while ((slot = GgExecProcNode(ps)) != NULL && !GgTupIsNull(slot))
	{
		double f[7]; int i;
		for (i = 0; i < 7; i++) memcpy(&f[i], &slot->tts_values[2 + i], 8);
		printf("%c %c  sum_qty %.2f  sum_base %.2f  sum_disc %.4f  sum_charge %.6f  avg_qty %.6f  avg_price %.6f  avg_disc %.8f  count %lld\n",
		       (char) slot->tts_values[0], (char) slot->tts_values[1], f[0], f[1], f[2], f[3], f[4], f[5], f[6], (long long) slot->tts_values[9]);
	}
	if (GgExecLastErrorCode() != GG_OK) { fprintf(stderr, "ExecProcNode: %s\n", GgExecLastError()); return 2; }
	GgExecEndNode(ps);
	gg_relation_free(rel);
	gg_engine_free(eng);
	free(pages);
	return 0;
}
