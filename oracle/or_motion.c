/*
 * or_motion.c — ORACLE (test infrastructure): destination routing of Motion.
 *
 *   execMotionSender / doSendTuple (MOTIONTYPE_HASH)   src/backend/executor/nodeMotion.c:270-374,1574-1687
 *   evalHashKey                                        src/backend/executor/nodeMotion.c:1481-1530
 *   cdbhashinit / cdbhash / cdbhashreduce              src/backend/cdb/cdbhash.c:173-287
 *
 * The interconnect itself (cdbmotion.c, ic_udpifc.c) is replaced wholesale by
 * NCCL in the product; what must match bit for bit is WHERE each row goes.
 */
#include <stdlib.h>
#include <string.h>
#include "gg_oracle.h"
#include "or_internal.h"

int
or_motion_route(const gg_scan *scan, const gg_exprpool *pool, const int32_t *hashkeys, int nkeys,
				int nsegs, const uint8_t *pages, uint64_t nblocks,
				int32_t *dest_out, uint64_t cap, uint64_t *nrows)
{
	or_heapscan *hs = malloc(sizeof *hs);
	const uint8_t *tup;
	uint64_t n = 0;
	or_row row;
	int rc = 0, i;

	or_scan_begin(hs, &scan->desc, pages, nblocks);
	while ((tup = or_scan_next(hs, NULL)) != NULL)
	{
		uint32_t h;

		or_row_store(&row, &scan->desc, tup);
		if (scan->qual >= 0)
		{
			or_datum q;

			if ((rc = or_eval(pool, scan->qual, &row, NULL, &q)) != 0)
				break;
			if (q.isnull || !q.v)
				continue;
		}
		/* evalHashKey: cdbhashinit; per key cdbhash(datum, isnull); cdbhashreduce */
		h = or_cdbhash_init();
		for (i = 0; i < nkeys; i++)
		{
			or_datum d;
			uint32_t hk = 0;

			if ((rc = or_eval(pool, hashkeys[i], &row, NULL, &d)) != 0)
				goto out;
			if (!d.isnull)
			{
				int32_t typ = pool->nodes[hashkeys[i]].rettype;

				if (d.ptr)
				{
					if (typ == GG_BPCHAROID)
						hk = or_hashbpchar((const char *) d.ptr, d.len);
					else
						hk = or_hash_any(d.ptr, d.len);
				}
				else
					hk = or_hash_datum(typ, d.v, d.len);
			}
			h = or_cdbhash_add(h, hk, d.isnull);
		}
		if (n >= cap)
		{
			rc = OR_ERR_NOMEM;
			break;
		}
		dest_out[n++] = or_cdbhash_reduce(h, nsegs);
	}
out:
	if (rc == 0 && hs->error)
		rc = hs->error;
	free(hs);
	*nrows = n;
	return rc;
}
