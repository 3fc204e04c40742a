/* or_internal.h — ORACLE (test infrastructure) internals shared between or_*.c */
#ifndef OR_INTERNAL_H
#define OR_INTERNAL_H
#include "gg_oracle.h"

/* HeapScanDescData subset (src/include/access/relscan.h): rs_vistuples etc. */
typedef struct or_heapscan {
	const gg_tupdesc *desc;
	const uint8_t *pages;
	uint64_t nblocks;
	uint64_t cblock;
	int inited;
	int ntuples;                 /* rs_ntuples */
	int cindex;                  /* rs_cindex */
	int error;
	uint16_t vistuples[GG_BLCKSZ / 28 + 1];   /* MaxHeapTuplesPerPage-ish bound */
} or_heapscan;

void or_scan_begin(or_heapscan *s, const gg_tupdesc *desc, const uint8_t *pages, uint64_t nblocks);
const uint8_t *or_scan_next(or_heapscan *s, uint64_t *tid);

static inline void or_row_store(or_row *r, const gg_tupdesc *desc, const uint8_t *tuple)
{
	r->desc = desc; r->tuple = tuple; r->nvalid = 0;
}

/* group-key / agg machinery shared by or_agg.c and or_join.c */
typedef struct or_aggtable or_aggtable;
or_aggtable *or_aggtable_create(const gg_agg *agg, const gg_exprpool *pool, long max_entries);
int  or_aggtable_advance(or_aggtable *t, or_row *outer, or_row *inner);
int  or_aggtable_emit(or_aggtable *t, gg_aggrow *out, int outcap, int *nout);
void or_aggtable_free(or_aggtable *t);

double or_now(void);
#endif
