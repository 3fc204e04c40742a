/*
 * or_sort.c — ORACLE (test infrastructure): ordering semantics of Sort.
 *
 *   ExecSort                       src/backend/executor/nodeSort.c:48-256
 *   inlineApplySortFunction        src/backend/utils/sort/tuplesort_mk.c:2816-2850
 *   float8_cmp_internal            src/backend/utils/adt/float.c:964-988 (NaN = NaN, NaN > all)
 *   btint4cmp/btint8cmp/date_cmp   src/backend/access/nbtree/nbtcompare.c, utils/adt/date.c
 *   bpcharcmp                      src/backend/utils/adt/varchar.c:840 (C locale: memcmp on stripped)
 *
 * mk_qsort (tuplesort_mkqsort.c) is an unstable multi-key quicksort: the order
 * of rows with equal keys is unspecified, so only the comparator is the
 * contract.  The permutation is produced with the C library's qsort_r; tests
 * compare tie groups as multisets.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gg_oracle.h"

static int
cmp_datum(int32_t typid, int64_t a, int64_t b)
{
	switch (typid)
	{
		case GG_INT4OID:
		case GG_DATEOID:
		{
			int32_t x = (int32_t) a, y = (int32_t) b;

			return (x > y) - (x < y);
		}
		case GG_FLOAT8OID:
		{
			double x, y;

			memcpy(&x, &a, 8);
			memcpy(&y, &b, 8);
			if (isnan(x))
				return isnan(y) ? 0 : 1;
			if (isnan(y))
				return -1;
			return (x > y) - (x < y);
		}
		case GG_BPCHAROID:
		case GG_VARCHAROID:
		case GG_TEXTOID:
		{
			/* packed strings: <= 8 bytes, LSB first, zero padded; C-locale memcmp
			 * then shorter-first (varstr_cmp); a NUL can never be a payload byte */
			const unsigned char *p = (const unsigned char *) &a, *q = (const unsigned char *) &b;
			int i;

			for (i = 0; i < 8; i++)
				if (p[i] != q[i])
					return p[i] < q[i] ? -1 : 1;
			return 0;
		}
		default:
			return (a > b) - (a < b);
	}
}

int
or_sort_compare(const gg_sortkey *keys, int nkeys, int ncols,
				const int64_t *a, const uint8_t *an, const int64_t *b, const uint8_t *bn)
{
	int k;

	(void) ncols;
	for (k = 0; k < nkeys; k++)
	{
		int c = keys[k].col, compare;
		int n1 = an ? an[c] : 0, n2 = bn ? bn[c] : 0;

		if (n1)
			compare = n2 ? 0 : (keys[k].nulls_first ? -1 : 1);
		else if (n2)
			compare = keys[k].nulls_first ? 1 : -1;
		else
		{
			compare = cmp_datum(keys[k].typid, a[c], b[c]);
			if (keys[k].desc)
				compare = -compare;
		}
		if (compare)
			return compare;
	}
	return 0;
}

typedef struct sort_ctx {
	const gg_sortkey *keys;
	int nkeys, ncols;
	const int64_t *rows;
	const uint8_t *nulls;
} sort_ctx;

static int
perm_cmp(const void *pa, const void *pb, void *vc)
{
	const sort_ctx *c = vc;
	uint64_t i = *(const uint64_t *) pa, j = *(const uint64_t *) pb;

	return or_sort_compare(c->keys, c->nkeys, c->ncols,
						   c->rows + i * c->ncols, c->nulls ? c->nulls + i * c->ncols : NULL,
						   c->rows + j * c->ncols, c->nulls ? c->nulls + j * c->ncols : NULL);
}

int
or_sort_perm(const gg_sortkey *keys, int nkeys, int ncols, const int64_t *rows,
			 const uint8_t *nulls, uint64_t n, uint64_t *perm_out)
{
	sort_ctx c;
	uint64_t i;

	c.keys = keys; c.nkeys = nkeys; c.ncols = ncols; c.rows = rows; c.nulls = nulls;
	for (i = 0; i < n; i++)
		perm_out[i] = i;
	qsort_r(perm_out, n, sizeof(uint64_t), perm_cmp, &c);
	return 0;
}
