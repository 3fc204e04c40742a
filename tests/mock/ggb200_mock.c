/*
 * ggb200_mock.c — TEST INFRASTRUCTURE: a stand-in for libggb200.so that answers the C-ABI calls of the executor-node
 * surface (greengage_b200/host/gg_executor.c) with the ORACLE, so the host logic above the device engine — pipeline
 * fusion, slot formation, ReScan / Squelch, and above all the N > 1 path through Motion (routing, transport, FINAL stage on
 * the receiving segments, merging Gather) — can run on a CPU-only box with several gloo ranks.
 * Built by tests/test_executor_multiseg.py into a temporary directory together with the product's gg_executor.c and
 * gg_motion_host.c; never shipped, never loaded by the product.
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/ggb200.h"
#include "../../oracle/gg_oracle.h"

struct gg_engine { int unused; };
struct gg_relation { const uint8_t *pages; uint64_t nblocks; };
struct gg_scanagg { gg_scan scan; gg_agg agg; gg_exprpool pool; const uint8_t *pages; uint64_t nblocks; int fed; };
struct gg_joinagg { gg_scan outer, inner; gg_hashjoin hj; gg_agg agg; gg_exprpool pool;
                    const uint8_t *opages, *ipages; uint64_t onb, inb; };

static char last_error[256] = "";
const char *gg_last_error(void) { return last_error; }

/* The product's own plan compiler (greengage_b200/csrc/gg_compile.cpp, host C++, linked in) decides what *_create accepts,
 * exactly as in libggb200.so: a plan it refuses never reaches the oracle. */
#include <stdarg.h>
#include <stdio.h>
void gg_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(last_error, sizeof last_error, fmt, ap); va_end(ap); }
int mock_compile_scanagg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool, char *err, int errlen);
int mock_compile_join(const gg_scan *outer, const gg_scan *inner, const gg_hashjoin *hj, const gg_agg *agg, const gg_exprpool *pool,
                      char *err, int errlen);

static int fail(int rc, const char *what)
{
	if (rc) { strncpy(last_error, what, sizeof last_error - 1); }
	return rc == 0 ? GG_OK : (rc == OR_ERR_NOMEM ? GG_ERR_NOMEM : rc);      /* OR_ERR_* use the GG_ERR_* numbers */
}

/* test-only constructors */
gg_engine *mock_engine(void) { static gg_engine e; return &e; }
gg_relation *mock_relation(const uint8_t *pages, uint64_t nblocks)
{
	gg_relation *r = calloc(1, sizeof *r);
	r->pages = pages; r->nblocks = nblocks;
	return r;
}

uint64_t gg_relation_nblocks(gg_relation *r) { return r->nblocks; }

int gg_scanagg_create(gg_engine *e, const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool, gg_scanagg **out)
{
	gg_scanagg *p;
	int rc = mock_compile_scanagg(scan, agg, pool, last_error, sizeof last_error);
	(void) e;
	if (rc != GG_OK) return rc;
	p = calloc(1, sizeof *p);
	p->scan = *scan; p->agg = *agg; p->pool = *pool;
	*out = p;
	return GG_OK;
}

int gg_scanagg_run(gg_scanagg *p, gg_relation *r, uint64_t first_block, uint64_t nblocks)
{
	if (p->fed) return fail(GG_ERR_ARG, "mock: one run per accumulation");
	p->pages = r->pages + first_block * (uint64_t) GG_BLCKSZ;
	p->nblocks = nblocks;
	p->fed = 1;
	return GG_OK;
}

int gg_scanagg_fetch(gg_scanagg *p, gg_aggrow *out, int outcap, int *nout, uint64_t *rows_scanned, uint64_t *rows_passed)
{
	return fail(or_seqscan_agg(&p->scan, &p->agg, &p->pool, p->pages, p->nblocks, out, outcap, nout, rows_scanned, rows_passed),
	            "mock: or_seqscan_agg failed");
}

int gg_scanagg_reset(gg_scanagg *p) { p->fed = 0; p->pages = NULL; p->nblocks = 0; return GG_OK; }
void gg_scanagg_free(gg_scanagg *p) { free(p); }

int gg_agg_final(gg_engine *e, const gg_agg *agg, const gg_aggrow *in, int nin, gg_aggrow *out, int outcap, int *nout)
{
	(void) e;
	return fail(or_agg_final(agg, in, nin, out, outcap, nout), "mock: or_agg_final failed");
}

int gg_joinagg_create(gg_engine *e, const gg_scan *outer, const gg_scan *inner, const gg_hashjoin *hj,
                      const gg_agg *agg, const gg_exprpool *pool, gg_joinagg **out)
{
	gg_joinagg *p;
	int rc = mock_compile_join(outer, inner, hj, agg, pool, last_error, sizeof last_error);
	(void) e;
	if (rc != GG_OK) return rc;
	p = calloc(1, sizeof *p);
	p->outer = *outer; p->inner = *inner; p->hj = *hj; p->agg = *agg; p->pool = *pool;
	*out = p;
	return GG_OK;
}

int gg_joinagg_build(gg_joinagg *p, gg_relation *inner, uint64_t first_block, uint64_t nblocks)
{
	p->ipages = inner->pages + first_block * (uint64_t) GG_BLCKSZ; p->inb = nblocks;
	return GG_OK;
}

int gg_joinagg_probe(gg_joinagg *p, gg_relation *outer, uint64_t first_block, uint64_t nblocks)
{
	p->opages = outer->pages + first_block * (uint64_t) GG_BLCKSZ; p->onb = nblocks;
	return GG_OK;
}

int gg_joinagg_set_work_mem(gg_joinagg *p, uint64_t bytes) { (void) p; (void) bytes; return GG_OK; }
int gg_joinagg_nbatch(gg_joinagg *p) { (void) p; return 1; }
int gg_joinagg_run(gg_joinagg *p, gg_relation *inner, gg_relation *outer)
{
	p->ipages = inner->pages; p->inb = inner->nblocks;
	p->opages = outer->pages; p->onb = outer->nblocks;
	return GG_OK;
}

int gg_joinagg_fetch(gg_joinagg *p, gg_aggrow *out, int outcap, int *nout, uint64_t *rows_joined)
{
	return fail(or_hashjoin_agg(&p->outer, &p->inner, &p->hj, &p->agg, &p->pool, p->opages, p->onb, p->ipages, p->inb,
	                            out, outcap, nout, rows_joined), "mock: or_hashjoin_agg failed");
}

int gg_joinagg_reset(gg_joinagg *p) { p->opages = NULL; p->onb = 0; return GG_OK; }
void gg_joinagg_free(gg_joinagg *p) { free(p); }

int gg_sort_rows(gg_engine *e, const gg_sortkey *keys, int nkeys, int ncols, const int64_t *host_rows, const uint8_t *host_nulls,
                 uint64_t n, uint64_t *host_perm)
{
	(void) e;
	return fail(or_sort_perm(keys, nkeys, ncols, host_rows, host_nulls, n, host_perm), "mock: or_sort_perm failed");
}

/* device-resident results and the NCCL interconnect have no stand-in: the executor falls back to host rows and to the
 * transport callback, which is what these tests exercise */
static int unsupported(void) { return fail(GG_ERR_UNSUPPORTED, "mock: no device-resident results"); }
int gg_scanagg_groups(gg_scanagg *p, gg_groups **out) { (void) p; *out = NULL; return unsupported(); }
int gg_joinagg_groups(gg_joinagg *p, gg_groups **out) { (void) p; *out = NULL; return unsupported(); }
int gg_groups_final(gg_engine *e, gg_groups *in, gg_groups **out) { (void) e; (void) in; *out = NULL; return unsupported(); }
int gg_groups_fetch(gg_groups *g, gg_aggrow *out, int outcap, int *nout, uint64_t *a, uint64_t *b)
{ (void) g; (void) out; (void) outcap; (void) nout; (void) a; (void) b; return unsupported(); }
void gg_groups_set_nonreceiver(gg_groups *g) { (void) g; }
void gg_groups_free(gg_groups *g) { (void) g; }
int gg_ic_create(gg_engine *e, const void *id, int nsegs, int seg, gg_interconnect **out) { (void) e; (void) id; (void) nsegs; (void) seg; *out = NULL; return unsupported(); }
void gg_ic_teardown(gg_interconnect *ic, int has_errors) { (void) ic; (void) has_errors; }
int gg_ic_motion_groups(gg_interconnect *ic, int t, int root, int nhash, const int32_t *hc, const int32_t *ht, gg_groups *in, int lerr, gg_groups **out)
{ (void) ic; (void) t; (void) root; (void) nhash; (void) hc; (void) ht; (void) in; (void) lerr; *out = NULL; return unsupported(); }
int gg_ic_exchange_rows(gg_interconnect *ic, const void *s, const uint64_t *c, uint64_t rc, int w, void *r, uint64_t cap, uint64_t *n)
{ (void) ic; (void) s; (void) c; (void) rc; (void) w; (void) r; (void) cap; (void) n; return unsupported(); }
int gg_ic_exchange_host(gg_interconnect *ic, int ncols, int64_t nrows, const int64_t *v, const uint8_t *nl, const int32_t *d, int err,
                        int64_t *on, int64_t **ov, uint8_t **onl)
{ (void) ic; (void) ncols; (void) nrows; (void) v; (void) nl; (void) d; (void) err; (void) on; (void) ov; (void) onl; return unsupported(); }
int gg_scanagg_run_host(gg_scanagg *p, const void *host_pages, uint64_t nblocks)
{
	if (p->fed) return fail(GG_ERR_ARG, "mock: one run per accumulation");
	p->pages = host_pages; p->nblocks = nblocks; p->fed = 1;
	return GG_OK;
}
int gg_relation_create(gg_engine *e, uint64_t nblocks, gg_relation **out) { (void) e; (void) nblocks; *out = NULL; return unsupported(); }
int gg_relation_attach_rows(gg_engine *e, void *rows, uint64_t nrows, int ncols, gg_relation **out) { (void) e; (void) rows; (void) nrows; (void) ncols; *out = NULL; return unsupported(); }
int gg_relation_read(gg_relation *r, uint64_t first, void *host, uint64_t n) { (void) r; (void) first; (void) host; (void) n; return unsupported(); }
void *gg_relation_device_ptr(gg_relation *r) { (void) r; return NULL; }
void gg_relation_free(gg_relation *r) { free(r); }
int gg_motion_partition(gg_engine *e, const gg_scan *scan, const gg_exprpool *pool, const int32_t *hk, int nk, const int32_t *pl, int np,
                        int nsegs, gg_relation *r, uint64_t fb, uint64_t nb, void *out, uint64_t cap, uint64_t *hc, uint64_t *ho)
{ (void) e; (void) scan; (void) pool; (void) hk; (void) nk; (void) pl; (void) np; (void) nsegs; (void) r; (void) fb; (void) nb; (void) out; (void) cap; (void) hc; (void) ho; return unsupported(); }
int gg_sort_datumrows(gg_engine *e, const gg_sortkey *keys, int nkeys, int ncols, const void *rows, uint64_t n, void *out, uint64_t *nlive, int *passes)
{ (void) e; (void) keys; (void) nkeys; (void) ncols; (void) rows; (void) n; (void) out; (void) nlive; (void) passes; return unsupported(); }
int gg_engine_set_snapshot(gg_engine *e, const gg_snapshot *s) { (void) e; or_set_snapshot(s); return GG_OK; }   /* the oracle's scans stand in for the device's */
int gg_relation_count_rows(gg_relation *r, uint64_t *n) { (void) r; (void) n; return unsupported(); }
int gg_scanagg_scan_kernel_ms(gg_scanagg *p, float *ms, int *launches) { (void) p; if (ms) *ms = 0; if (launches) *launches = 0; return GG_OK; }
int gg_scanagg_variant(gg_scanagg *p) { (void) p; return -1; }
int gg_joinagg_variant(gg_joinagg *p) { (void) p; return -1; }
int gg_joinagg_stats(gg_joinagg *p, uint64_t *a, uint64_t *b, float *c, float *d) { (void) p; (void) a; (void) b; (void) c; (void) d; return unsupported(); }
