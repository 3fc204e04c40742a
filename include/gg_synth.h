/*
 * gg_synth.h — deterministic TPC-H-shaped heap relations for tests and benchmarks
 * (SURVEY §8d).  The loader writes byte-exact Greengage heap pages (32 KB,
 * bufpage.h:153; frozen all-visible tuples, htup_details.h:139) into host
 * memory, the way COPY + VACUUM FREEZE would leave them on disk; the executor
 * then reads them like any other relation.  Column order and types follow the
 * reference's TPC-H schema (src/test/regress/sql/tpch500GB.sql:66-83,102-112)
 * with numeric(15,2) replaced by float8 as BASELINE.json's north_star asks.
 *
 * Every value is a pure function of (seed, table, candidate id, column), so any
 * segment, extent or page can be generated independently and in parallel.
 */
#ifndef GG_SYNTH_H
#define GG_SYNTH_H

#include <stdint.h>
#include "gg_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

enum gg_synth_table {
	GG_TAB_LINEITEM_WIDE = 1,    /* 16 columns, ~172 B/row incl. line pointer */
	GG_TAB_LINEITEM_NARROW = 2,  /* 8 columns Q1 needs + l_orderkey, 76 B/row */
	GG_TAB_ORDERS = 3            /* 9 columns, ~148 B/row */
};

enum gg_synth_policy {
	GG_DIST_RANDOM = 0,          /* DISTRIBUTED RANDOMLY: candidate id mod nsegs */
	GG_DIST_HASH = 1             /* DISTRIBUTED BY (orderkey): cdbhash + jump consistent hash (cdbhash.c:191-287) */
};

#define GG_SYNTH_EXTENT 65536    /* candidate ids per extent; every extent starts on a fresh page */

typedef struct gg_synth_spec {
	int32_t  table;
	int32_t  policy;
	uint64_t seed;
	uint64_t ncand;              /* candidate ids [0, ncand): rows of the whole table over all segments */
	uint64_t norders;            /* size of the orders key space l_orderkey draws from (orders: == ncand) */
	int32_t  nsegs;
	int32_t  seg;
} gg_synth_spec;

int      gg_synth_tupdesc(int table, gg_tupdesc *out);
/* blocks and rows this segment's relation will have */
int      gg_synth_measure(const gg_synth_spec *spec, int nthreads, uint64_t *nblocks, uint64_t *nrows);
/* write the pages; cap_blocks must be >= the measured block count */
int      gg_synth_generate(const gg_synth_spec *spec, int nthreads, uint8_t *pages, uint64_t cap_blocks,
                           uint64_t *nblocks, uint64_t *nrows);
/* values of candidate row `cand` (Datum bits; strings: pointer into strbuf + len), and whether it
 * belongs to this segment — for verification */
int      gg_synth_row(const gg_synth_spec *spec, uint64_t cand, int64_t *vals, int32_t *lens,
                      char *strbuf, int strcap, int *mine);
/* The same relation (same rows, same order) stored append-only column-oriented: one column file per requested
 * attribute (0-based attribute numbers in cols[]), written by include/gg_aocs.h's writer the way an INSERT into a table
 * `WITH (appendonly=true, orientation=column)` would.  out[i] / outcap[i]: buffer of column cols[i]
 * (gg_aocs_file_bound() sizes it); outbytes[i] receives the file length.  One thread per column, at most nthreads. */
int      gg_synth_aocs_generate(const gg_synth_spec *spec, int nthreads, const int32_t *cols, int ncols,
                                uint8_t *const *out, const int64_t *outcap, int blocksize, int checksum,
                                int64_t *outbytes, uint64_t *nrows);
/* the dbgen-style sparse order key of order index o: 8 keys per 32 */
int64_t  gg_synth_orderkey(uint64_t o);

#ifdef __cplusplus
}
#endif
#endif
