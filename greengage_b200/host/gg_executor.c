/*
 * gg_executor.c — ExecInitNode / ExecProcNode / ExecEndNode for the B200 segment engine (include/gg_executor.h).
 *
 * Host C above the C-ABI of libggb200.so; no CUDA here.  What the reference does tuple-at-a-time through
 * ExecProcNode dispatch (execProcnode.c:925-1100), this layer does pipeline-at-a-time: ExecInitNode fuses the
 * slice into device pipelines, the first ExecProcNode call on a pipeline's top node runs it on the device, and
 * every call hands out one row of the result as a virtual tuple — the contract the node above sees is the
 * reference's (one TupleTableSlot per call, NULL at end of stream, ExecReScan restarts, ExecSquelchNode stops
 * early; nodeAgg.c:1123, nodeHashjoin.c:78, nodeSort.c:48, nodeMotion.c:180).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/gg_executor.h"

int32_t gg_cdbhash_route(const int32_t *typids, const int64_t *vals, const int32_t *lens, const int32_t *isnull,
                         int nkeys, int nsegs);      /* gg_motion_host.c */

enum { K_SCANAGG = 1, K_JOINAGG, K_AGGFINAL, K_SORT, K_MOTION };

struct GgPlanState {
	int kind;
	GgPlan *plan;
	GgEState *estate;
	struct GgPlanState *child;          /* Sort / Motion / final Agg: the pipeline below */
	/* device pipelines */
	gg_scanagg *sa;
	gg_joinagg *ja;
	gg_relation *rel, *inner_rel;
	gg_agg agg;
	/* result set: filled by the first ExecProcNode, then handed out row by row */
	int done;                           /* pipeline has run */
	int squelched;
	int32_t ncols;
	int64_t nrows, next;
	int64_t *values;
	uint8_t *isnull;
	int32_t typid[GG_MAX_OUTCOLS];
	int32_t *lens;                      /* [nrows][ncols], strings only */
	GgTupleTableSlot slot;
};

static _Thread_local char g_err[512];
static _Thread_local int g_errcode;

static void *exec_fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	g_errcode = code;
	return NULL;
}

const char *GgExecLastError(void) { return g_err; }
int GgExecLastErrorCode(void) { return g_errcode; }

const char *GgExecNodeKind(GgPlanState *s)
{
	if (!s) return "";
	switch (s->kind)
	{
		case K_SCANAGG: return "scanagg";
		case K_JOINAGG: return "joinagg";
		case K_AGGFINAL: return "aggfinal";
		case K_SORT: return "sort";
		case K_MOTION: return "motion";
	}
	return "";
}

/* ---- output layout of an Agg node: group keys, then aggregates (a PARTIAL avg is its float8[3] state) ---- */
static int agg_ncols_of(const gg_agg *agg, int i)
{
	return (agg->aggs[i].aggfnoid == GG_AGG_AVG_FLOAT8 && agg->aggstage == GG_AGGSTAGE_PARTIAL) ? 3 : 1;
}

static int32_t agg_result_type(int32_t fn)
{
	switch (fn)
	{
		case GG_AGG_COUNT_ANY: case GG_AGG_COUNT_STAR: case GG_AGG_SUM_INT4: case GG_AGG_MAX_INT8: case GG_AGG_MIN_INT8:
			return GG_INT8OID;
		case GG_AGG_MAX_INT4: case GG_AGG_MIN_INT4: return GG_INT4OID;
		case GG_AGG_MAX_DATE: case GG_AGG_MIN_DATE: return GG_DATEOID;
		default: return GG_FLOAT8OID;
	}
}

static int64_t f8bits(double d) { int64_t v; memcpy(&v, &d, 8); return v; }
static double bitsf8(int64_t v) { double d; memcpy(&d, &v, 8); return d; }

static int alloc_result(GgPlanState *s, int64_t nrows, int32_t ncols)
{
	free(s->values); free(s->isnull); free(s->lens);
	s->nrows = nrows; s->ncols = ncols; s->next = 0;
	s->values = calloc((size_t) (nrows > 0 ? nrows : 1) * (size_t) ncols, 8);
	s->isnull = calloc((size_t) (nrows > 0 ? nrows : 1) * (size_t) ncols, 1);
	s->lens = calloc((size_t) (nrows > 0 ? nrows : 1) * (size_t) ncols, 4);
	return (s->values && s->isnull && s->lens) ? 0 : -1;
}

/* gg_aggrow[] -> result columns */
static int rows_from_aggrows(GgPlanState *s, const gg_agg *agg, const int32_t *keytypes, const gg_aggrow *rows, int n)
{
	int ncols = agg->numCols, i, c, r;
	for (i = 0; i < agg->numAggs; i++) ncols += agg_ncols_of(agg, i);
	if (ncols > GG_MAX_OUTCOLS) { exec_fail(GG_ERR_UNSUPPORTED, "too many output columns"); return -1; }
	if (alloc_result(s, n, ncols)) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
	for (c = 0; c < agg->numCols; c++) s->typid[c] = keytypes[c];
	for (i = 0, c = agg->numCols; i < agg->numAggs; i++)
	{
		int w = agg_ncols_of(agg, i), k;
		for (k = 0; k < w; k++) s->typid[c + k] = (w == 3) ? GG_FLOAT8OID : agg_result_type(agg->aggs[i].aggfnoid);
		c += w;
	}
	for (r = 0; r < n; r++)
	{
		int64_t *v = s->values + (size_t) r * ncols;
		uint8_t *nl = s->isnull + (size_t) r * ncols;
		int32_t *ln = s->lens + (size_t) r * ncols;
		for (c = 0; c < agg->numCols; c++)
		{
			v[c] = rows[r].key[c]; nl[c] = (uint8_t) rows[r].keyisnull[c]; ln[c] = rows[r].keylen[c];
		}
		for (i = 0, c = agg->numCols; i < agg->numAggs; i++)
		{
			const gg_aggval *a = &rows[r].agg[i];
			int w = agg_ncols_of(agg, i);
			if (w == 3) { v[c] = f8bits(a->f[0]); v[c + 1] = f8bits(a->f[1]); v[c + 2] = f8bits(a->f[2]); }
			else
			{
				nl[c] = (uint8_t) a->isnull;
				v[c] = s->typid[c] == GG_FLOAT8OID ? f8bits(a->f[0]) : a->i;
			}
			c += w;
		}
	}
	return 0;
}

/* result columns of a PARTIAL Agg (as they come out of a Motion) -> gg_aggrow[] for the FINAL stage */
static gg_aggrow *aggrows_from_rows(const GgPlanState *child, const gg_agg *agg)
{
	int64_t r;
	gg_aggrow *out = calloc((size_t) (child->nrows > 0 ? child->nrows : 1), sizeof *out);
	if (!out) return NULL;
	for (r = 0; r < child->nrows; r++)
	{
		const int64_t *v = child->values + (size_t) r * child->ncols;
		const uint8_t *nl = child->isnull + (size_t) r * child->ncols;
		const int32_t *ln = child->lens + (size_t) r * child->ncols;
		int c, i;
		for (c = 0; c < agg->numCols; c++)
		{
			out[r].key[c] = v[c]; out[r].keyisnull[c] = nl[c]; out[r].keylen[c] = ln[c];
		}
		for (i = 0, c = agg->numCols; i < agg->numAggs; i++)
		{
			gg_aggval *a = &out[r].agg[i];
			int32_t fn = agg->aggs[i].aggfnoid;
			if (fn == GG_AGG_AVG_FLOAT8)
			{
				a->f[0] = bitsf8(v[c]); a->f[1] = bitsf8(v[c + 1]); a->f[2] = bitsf8(v[c + 2]);
				c += 3;
			}
			else
			{
				a->isnull = nl[c];
				if (agg_result_type(fn) == GG_FLOAT8OID) a->f[0] = bitsf8(v[c]); else a->i = v[c];
				c += 1;
			}
		}
	}
	return out;
}

static void free_state(GgPlanState *s)
{
	if (!s) return;
	if (s->sa) gg_scanagg_free(s->sa);
	if (s->ja) gg_joinagg_free(s->ja);
	free(s->values); free(s->isnull); free(s->lens);
	free(s);
}

static gg_relation *relation_of(GgEState *es, const GgSeqScan *scan)
{
	if (scan->scanrelid < 0 || scan->scanrelid >= GG_MAX_RELATIONS || !es->relations[scan->scanrelid])
		return exec_fail(GG_ERR_ARG, "SeqScan: relation %d is not resident on the device", scan->scanrelid);
	return es->relations[scan->scanrelid];
}

static int32_t expr_type(const gg_exprpool *pool, int32_t root) { return pool->nodes[root].rettype; }

#define GG_MAX_PLAN_DEPTH 32        /* the deepest accelerated slice is Motion <- Sort <- Agg <- Motion <- Agg <- HashJoin <- Hash <- SeqScan */

static GgPlanState *init_node(GgPlan *node, GgEState *estate, int eflags, int depth);

GgPlanState *GgExecInitNode(GgPlan *node, GgEState *estate, int eflags)
{
	g_err[0] = 0; g_errcode = GG_OK;
	return init_node(node, estate, eflags, 0);
}

static GgPlanState *init_node(GgPlan *node, GgEState *estate, int eflags, int depth)
{
	GgPlanState *s;
	(void) eflags;
	if (!node) return NULL;                                   /* ExecInitNode(NULL) is NULL, execProcnode.c:268 */
	if (!estate || !estate->engine || !estate->pool) return exec_fail(GG_ERR_ARG, "EState without engine or expression pool");
	if (depth > GG_MAX_PLAN_DEPTH) return exec_fail(GG_ERR_ARG, "plan tree deeper than %d nodes (a cycle?)", GG_MAX_PLAN_DEPTH);
	s = calloc(1, sizeof *s);
	if (!s) return exec_fail(GG_ERR_NOMEM, "out of memory");
	s->plan = node; s->estate = estate;
	switch (node->type)
	{
		case T_GgAgg:
		{
			GgAgg *an = (GgAgg *) node;
			GgPlan *below = node->lefttree;
			int rc;
			s->agg = an->agg;
			if (an->agg.numCols < 0 || an->agg.numCols > GG_MAX_KEYS || an->agg.numAggs < 0 || an->agg.numAggs > GG_MAX_AGGS)
			{
				exec_fail(GG_ERR_ARG, "Agg with %d grouping columns and %d aggregates", an->agg.numCols, an->agg.numAggs);
				free_state(s);
				return NULL;
			}
			if (an->agg.aggstage == GG_AGGSTAGE_FINAL)
			{
				/* the receiving half of a two-stage aggregate: combine what the Motion below delivers */
				s->kind = K_AGGFINAL;
				s->child = init_node(below, estate, eflags, depth + 1);
				if (!s->child) { free_state(s); return NULL; }
				return s;
			}
			if (below && below->type == T_GgSeqScan)
			{
				GgSeqScan *sc = (GgSeqScan *) below;
				gg_scan scan;
				memset(&scan, 0, sizeof scan);
				scan.desc = sc->desc; scan.qual = below->qual;
				if (!(s->rel = relation_of(estate, sc))) { free_state(s); return NULL; }
				s->kind = K_SCANAGG;
				rc = gg_scanagg_create(estate->engine, &scan, &an->agg, estate->pool, &s->sa);
				if (rc != GG_OK) { exec_fail(rc, "Agg <- SeqScan: %s", gg_last_error()); free_state(s); return NULL; }
				return s;
			}
			if (below && below->type == T_GgHashJoin)
			{
				GgHashJoin *hj = (GgHashJoin *) below;
				GgPlan *outer = below->lefttree, *hash = below->righttree, *inner = hash ? hash->lefttree : NULL;
				gg_scan oscan, iscan;
				if (!outer || outer->type != T_GgSeqScan || !hash || hash->type != T_GgHash || !inner || inner->type != T_GgSeqScan)
				{
					exec_fail(GG_ERR_UNSUPPORTED, "HashJoin: only SeqScan ⋈ Hash(SeqScan) is fused on the device");
					free_state(s);
					return NULL;
				}
				memset(&oscan, 0, sizeof oscan); memset(&iscan, 0, sizeof iscan);
				oscan.desc = ((GgSeqScan *) outer)->desc; oscan.qual = outer->qual;
				iscan.desc = ((GgSeqScan *) inner)->desc; iscan.qual = inner->qual;
				if (!(s->rel = relation_of(estate, (GgSeqScan *) outer)) || !(s->inner_rel = relation_of(estate, (GgSeqScan *) inner)))
				{ free_state(s); return NULL; }
				s->kind = K_JOINAGG;
				rc = gg_joinagg_create(estate->engine, &oscan, &iscan, &hj->hj, &an->agg, estate->pool, &s->ja);
				if (rc != GG_OK) { exec_fail(rc, "Agg <- HashJoin: %s", gg_last_error()); free_state(s); return NULL; }
				return s;
			}
			exec_fail(GG_ERR_UNSUPPORTED, "Agg: child node type %d is not on the accelerated path", below ? (int) below->type : 0);
			free_state(s);
			return NULL;
		}
		case T_GgSort:
		{
			GgSort *so = (GgSort *) node;
			if (so->numCols < 1 || so->numCols > GG_MAX_SORTKEYS) { exec_fail(GG_ERR_UNSUPPORTED, "Sort with %d keys", so->numCols); free_state(s); return NULL; }
			s->kind = K_SORT;
			s->child = init_node(node->lefttree, estate, eflags, depth + 1);
			if (!s->child) { free_state(s); return NULL; }
			return s;
		}
		case T_GgMotion:
		{
			GgMotion *mo = (GgMotion *) node;
			if (!estate->transport && estate->nsegs > 1) { exec_fail(GG_ERR_ARG, "Motion: %d segments but no transport", estate->nsegs); free_state(s); return NULL; }
			if (mo->motionType == GG_MOTIONTYPE_HASH && (mo->numHashCols < 1 || mo->numHashCols > GG_MAX_KEYS))
			{ exec_fail(GG_ERR_UNSUPPORTED, "Redistribute Motion with %d hash columns", mo->numHashCols); free_state(s); return NULL; }
			s->kind = K_MOTION;
			s->child = init_node(node->lefttree, estate, eflags, depth + 1);
			if (!s->child) { free_state(s); return NULL; }
			return s;
		}
		case T_GgSeqScan: case T_GgHashJoin: case T_GgHash:
			/* bare scans / joins return whole tuples to a CPU parent: nothing to accelerate without an Agg on top */
			exec_fail(GG_ERR_UNSUPPORTED, "node type %d is only accelerated underneath an Agg", (int) node->type);
			free_state(s);
			return NULL;
	}
	exec_fail(GG_ERR_UNSUPPORTED, "unknown node type %d", (int) node->type);
	free_state(s);
	return NULL;
}

/* drain a child pipeline (what tuplesort_puttupleslot / the Motion sender loop do, nodeSort.c:139, nodeMotion.c:340) */
static int run_node(GgPlanState *s);

static int run_child(GgPlanState *s)
{
	if (!s->child->done && run_node(s->child)) return -1;
	return 0;
}

static int run_node(GgPlanState *s)
{
	GgEState *es = s->estate;
	int rc;
	switch (s->kind)
	{
		case K_SCANAGG:
		case K_JOINAGG:
		{
			int cap = 4096, n = 0, c;
			int32_t keytypes[GG_MAX_KEYS] = { 0 };
			gg_aggrow *rows = NULL;
			if (s->kind == K_SCANAGG)
				rc = gg_scanagg_run(s->sa, s->rel, 0, gg_relation_nblocks(s->rel));
			else
			{
				rc = gg_joinagg_build(s->ja, s->inner_rel, 0, gg_relation_nblocks(s->inner_rel));
				if (rc == GG_OK) rc = gg_joinagg_probe(s->ja, s->rel, 0, gg_relation_nblocks(s->rel));
			}
			/* the result stays on the device until fetched: grow the row buffer until every group fits */
			while (rc == GG_OK)
			{
				free(rows);
				rows = malloc(sizeof(gg_aggrow) * (size_t) cap);
				if (!rows) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
				rc = s->kind == K_SCANAGG ? gg_scanagg_fetch(s->sa, rows, cap, &n, NULL, NULL) : gg_joinagg_fetch(s->ja, rows, cap, &n, NULL);
				if (rc != GG_ERR_NOMEM || cap >= (1 << 24)) break;
				cap *= 16;
				rc = GG_OK;
			}
			if (rc != GG_OK) { exec_fail(rc, "%s", gg_last_error()); free(rows); return -1; }
			for (c = 0; c < s->agg.numCols; c++) keytypes[c] = expr_type(es->pool, s->agg.grpCol[c]);
			rc = rows_from_aggrows(s, &s->agg, keytypes, rows, n);
			free(rows);
			if (rc) return -1;
			break;
		}
		case K_AGGFINAL:
		{
			gg_aggrow *in, *out;
			int n = 0, cap;
			gg_agg part = s->agg;
			if (run_child(s)) return -1;
			part.aggstage = GG_AGGSTAGE_PARTIAL;      /* layout of the incoming rows */
			{
				int want = part.numCols, i;
				for (i = 0; i < part.numAggs; i++) want += agg_ncols_of(&part, i);
				if (want != s->child->ncols && s->child->nrows > 0)
				{ exec_fail(GG_ERR_ARG, "FINAL Agg expects %d columns of partial state, the node below delivers %d", want, s->child->ncols); return -1; }
			}
			in = aggrows_from_rows(s->child, &part);
			cap = s->child->nrows > 0 ? (int) s->child->nrows : 1;
			out = malloc(sizeof(gg_aggrow) * (size_t) cap);
			if (!in || !out) { free(in); free(out); exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
			rc = gg_agg_final(es->engine, &s->agg, in, (int) s->child->nrows, out, cap, &n);
			free(in);
			if (rc != GG_OK) { exec_fail(rc, "%s", gg_last_error()); free(out); return -1; }
			rc = rows_from_aggrows(s, &s->agg, s->agg.grpCol /* key type OIDs at the FINAL stage */, out, n);
			free(out);
			if (rc) return -1;
			break;
		}
		case K_SORT:
		{
			GgSort *so = (GgSort *) s->plan;
			GgPlanState *ch = s->child;
			uint64_t *perm;
			gg_sortkey keys[GG_MAX_SORTKEYS];
			int64_t r;
			int k;
			if (run_child(s)) return -1;
			for (k = 0; k < so->numCols; k++)
			{
				keys[k] = so->keys[k];
				if (keys[k].col < 0 || keys[k].col >= ch->ncols) { exec_fail(GG_ERR_ARG, "Sort key column %d out of range", keys[k].col); return -1; }
				if (!keys[k].typid) keys[k].typid = ch->typid[keys[k].col];
			}
			if (alloc_result(s, ch->nrows, ch->ncols)) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
			memcpy(s->typid, ch->typid, sizeof s->typid);
			perm = malloc(8 * (size_t) (ch->nrows > 0 ? ch->nrows : 1));
			if (!perm) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
			rc = gg_sort_rows(es->engine, keys, so->numCols, ch->ncols, ch->values, ch->isnull, (uint64_t) ch->nrows, perm);
			if (rc != GG_OK) { exec_fail(rc, "%s", gg_last_error()); free(perm); return -1; }
			for (r = 0; r < ch->nrows; r++)
			{
				memcpy(s->values + (size_t) r * ch->ncols, ch->values + (size_t) perm[r] * ch->ncols, 8 * (size_t) ch->ncols);
				memcpy(s->isnull + (size_t) r * ch->ncols, ch->isnull + (size_t) perm[r] * ch->ncols, (size_t) ch->ncols);
				memcpy(s->lens + (size_t) r * ch->ncols, ch->lens + (size_t) perm[r] * ch->ncols, 4 * (size_t) ch->ncols);
			}
			free(perm);
			break;
		}
		case K_MOTION:
		{
			GgMotion *mo = (GgMotion *) s->plan;
			GgPlanState *ch = s->child;
			int64_t r;
			int c;
			if (run_child(s)) return -1;
			if (!es->transport)
			{
				/* one segment: sender and receiver are the same process */
				if (alloc_result(s, ch->nrows, ch->ncols)) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
				memcpy(s->values, ch->values, 8 * (size_t) ch->nrows * ch->ncols);
				memcpy(s->isnull, ch->isnull, (size_t) ch->nrows * ch->ncols);
				memcpy(s->lens, ch->lens, 4 * (size_t) ch->nrows * ch->ncols);
			}
			else
			{
				GgRowBatch send, recv;
				int32_t *dest;
				if (mo->motionType == GG_MOTIONTYPE_HASH)
					for (c = 0; c < mo->numHashCols; c++)
						if (mo->hashCol[c] < 0 || mo->hashCol[c] >= ch->ncols)
						{ exec_fail(GG_ERR_ARG, "Motion hash column %d out of range (the node below has %d columns)", mo->hashCol[c], ch->ncols); return -1; }
				dest = malloc(4 * (size_t) (ch->nrows > 0 ? ch->nrows : 1));
				if (!dest) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
				for (r = 0; r < ch->nrows; r++)
				{
					if (mo->motionType == GG_MOTIONTYPE_HASH)
					{
						/* evalHashKey (nodeMotion.c:1481): cdbhash over the hash columns, reduced to a segment */
						int32_t t[GG_MAX_KEYS], ln[GG_MAX_KEYS], nn[GG_MAX_KEYS];
						int64_t v[GG_MAX_KEYS];
						for (c = 0; c < mo->numHashCols; c++)
						{
							int col = mo->hashCol[c];
							t[c] = ch->typid[col];
							v[c] = ch->values[(size_t) r * ch->ncols + col];
							ln[c] = ch->lens[(size_t) r * ch->ncols + col];
							nn[c] = ch->isnull[(size_t) r * ch->ncols + col];
						}
						dest[r] = gg_cdbhash_route(t, v, ln, nn, mo->numHashCols, es->nsegs);
					}
					else
						dest[r] = mo->motionType == GG_MOTIONTYPE_BROADCAST ? -1 : 0;
				}
				send.ncols = ch->ncols; send.nrows = ch->nrows; send.values = ch->values; send.isnull = ch->isnull;
				memset(&recv, 0, sizeof recv);
				rc = es->transport->exchange(es->transport->ctx, mo->motionID, mo->motionType, &send, dest, &recv);
				free(dest);
				if (rc) { exec_fail(GG_ERR_CUDA, "Motion %d: transport failed (%d)", mo->motionID, rc); return -1; }
				if (alloc_result(s, recv.nrows, ch->ncols)) { exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
				if (recv.nrows)
				{
					memcpy(s->values, recv.values, 8 * (size_t) recv.nrows * ch->ncols);
					memcpy(s->isnull, recv.isnull, (size_t) recv.nrows * ch->ncols);
				}
				/* string lengths do not travel: recompute from the packed bytes */
				for (r = 0; r < recv.nrows; r++)
					for (c = 0; c < ch->ncols; c++)
						if (ch->typid[c] == GG_BPCHAROID || ch->typid[c] == GG_VARCHAROID || ch->typid[c] == GG_TEXTOID)
						{
							uint64_t u = (uint64_t) s->values[(size_t) r * ch->ncols + c];
							int l = 0;
							while (l < 8 && ((u >> (8 * l)) & 0xff)) l++;
							s->lens[(size_t) r * ch->ncols + c] = l;
						}
				free(recv.values); free(recv.isnull);
			}
			memcpy(s->typid, ch->typid, sizeof s->typid);
			if (mo->numSortCols > 0 && s->nrows > 1)
			{
				/* sorted receive: the merged order of sorted streams is the sorted order of their union; the comparator is
				 * the Sort node's (tuplesort_mk.c:2816), ties in unspecified order as in the reference's merge */
				gg_sortkey keys[GG_MAX_SORTKEYS];
				uint64_t *perm = malloc(8 * (size_t) s->nrows);
				int64_t *v2 = malloc(8 * (size_t) s->nrows * s->ncols);
				uint8_t *n2 = malloc((size_t) s->nrows * s->ncols);
				int32_t *l2 = malloc(4 * (size_t) s->nrows * s->ncols);
				int k;
				if (mo->numSortCols > GG_MAX_SORTKEYS) { exec_fail(GG_ERR_UNSUPPORTED, "Motion with %d merge keys", mo->numSortCols); free(perm); free(v2); free(n2); free(l2); return -1; }
				if (!perm || !v2 || !n2 || !l2) { free(perm); free(v2); free(n2); free(l2); exec_fail(GG_ERR_NOMEM, "out of memory"); return -1; }
				for (k = 0; k < mo->numSortCols; k++)
				{
					keys[k] = mo->sortKeys[k];
					if (keys[k].col < 0 || keys[k].col >= s->ncols) { free(perm); free(v2); free(n2); free(l2); exec_fail(GG_ERR_ARG, "Motion merge key column %d out of range", keys[k].col); return -1; }
					if (!keys[k].typid) keys[k].typid = s->typid[keys[k].col];
				}
				rc = gg_sort_rows(es->engine, keys, mo->numSortCols, s->ncols, s->values, s->isnull, (uint64_t) s->nrows, perm);
				if (rc != GG_OK) { exec_fail(rc, "%s", gg_last_error()); free(perm); free(v2); free(n2); free(l2); return -1; }
				for (r = 0; r < s->nrows; r++)
				{
					memcpy(v2 + (size_t) r * s->ncols, s->values + (size_t) perm[r] * s->ncols, 8 * (size_t) s->ncols);
					memcpy(n2 + (size_t) r * s->ncols, s->isnull + (size_t) perm[r] * s->ncols, (size_t) s->ncols);
					memcpy(l2 + (size_t) r * s->ncols, s->lens + (size_t) perm[r] * s->ncols, 4 * (size_t) s->ncols);
				}
				free(s->values); free(s->isnull); free(s->lens); free(perm);
				s->values = v2; s->isnull = n2; s->lens = l2;
			}
			break;
		}
		default:
			exec_fail(GG_ERR_ARG, "bad plan state");
			return -1;
	}
	s->done = 1;
	s->next = 0;
	return 0;
}

GgTupleTableSlot *GgExecProcNode(GgPlanState *s)
{
	int c;
	if (!s || s->squelched) return NULL;
	if (!s->done)
	{
		g_err[0] = 0; g_errcode = GG_OK;
		if (run_node(s)) return NULL;                 /* the C wrapper on the Postgres side turns this into ereport(ERROR) */
	}
	if (s->next >= s->nrows)
	{
		s->slot.tts_isempty = 1;                      /* ExecClearTuple: end of stream */
		return NULL;
	}
	s->slot.tts_isempty = 0;
	s->slot.tts_nvalid = s->ncols;
	for (c = 0; c < s->ncols; c++)
	{
		s->slot.tts_values[c] = s->values[(size_t) s->next * s->ncols + c];
		s->slot.tts_isnull[c] = s->isnull[(size_t) s->next * s->ncols + c];
		s->slot.tts_typid[c] = s->typid[c];
		s->slot.tts_len[c] = s->lens[(size_t) s->next * s->ncols + c];
	}
	s->next++;
	s->estate->es_processed++;
	return &s->slot;
}

int GgExecReScan(GgPlanState *s)
{
	int rc = GG_OK;
	if (!s) return GG_ERR_ARG;
	if (s->child && (rc = GgExecReScan(s->child)) != GG_OK) return rc;
	if (s->sa) rc = gg_scanagg_reset(s->sa);
	if (s->ja && rc == GG_OK) rc = gg_joinagg_reset(s->ja);
	s->done = 0; s->squelched = 0; s->next = 0; s->nrows = 0;
	return rc;
}

void GgExecSquelchNode(GgPlanState *s)
{
	/* the node above needs no more rows (LIMIT satisfied, nodeLimit.c): stop handing them out */
	for (; s; s = s->child) s->squelched = 1;
}

void GgExecEndNode(GgPlanState *s)
{
	if (!s) return;
	GgExecEndNode(s->child);
	free_state(s);
}

/* ---- per-node entry points (executor/node*.h names) ---- */
static GgPlanState *init_tagged(GgPlan *plan, GgNodeTag tag, GgEState *estate, int eflags)
{
	if (!plan || plan->type != tag) return exec_fail(GG_ERR_ARG, "node tag %d where %d was expected", plan ? (int) plan->type : 0, (int) tag);
	return GgExecInitNode(plan, estate, eflags);
}

GgPlanState *GgExecInitAgg(GgAgg *node, GgEState *estate, int eflags) { return init_tagged(&node->plan, T_GgAgg, estate, eflags); }
GgTupleTableSlot *GgExecAgg(GgPlanState *node) { return GgExecProcNode(node); }
void GgExecEndAgg(GgPlanState *node) { GgExecEndNode(node); }
int GgExecReScanAgg(GgPlanState *node) { return GgExecReScan(node); }
void GgExecSquelchAgg(GgPlanState *node) { GgExecSquelchNode(node); }

GgPlanState *GgExecInitSort(GgSort *node, GgEState *estate, int eflags) { return init_tagged(&node->plan, T_GgSort, estate, eflags); }
GgTupleTableSlot *GgExecSort(GgPlanState *node) { return GgExecProcNode(node); }
void GgExecEndSort(GgPlanState *node) { GgExecEndNode(node); }
int GgExecReScanSort(GgPlanState *node) { return GgExecReScan(node); }
void GgExecSquelchSort(GgPlanState *node) { GgExecSquelchNode(node); }

GgPlanState *GgExecInitMotion(GgMotion *node, GgEState *estate, int eflags) { return init_tagged(&node->plan, T_GgMotion, estate, eflags); }
GgTupleTableSlot *GgExecMotion(GgPlanState *node) { return GgExecProcNode(node); }
void GgExecEndMotion(GgPlanState *node) { GgExecEndNode(node); }
int GgExecReScanMotion(GgPlanState *node) { return GgExecReScan(node); }
void GgExecSquelchMotion(GgPlanState *node) { GgExecSquelchNode(node); }

GgPlanState *GgExecInitHashJoin(GgHashJoin *node, GgEState *estate, int eflags) { return init_tagged(&node->plan, T_GgHashJoin, estate, eflags); }
GgPlanState *GgExecInitSeqScan(GgSeqScan *node, GgEState *estate, int eflags) { return init_tagged(&node->plan, T_GgSeqScan, estate, eflags); }
