/*
 * or_heap.c — ORACLE (test infrastructure): restatement of the heap tuple and
 * heap page formats and of the forward page-mode scan.
 *
 *   heap_compute_data_size   src/backend/access/common/heaptuple.c:68-129
 *   heap_fill_tuple          src/backend/access/common/heaptuple.c:149-272
 *   heap_form_tuple          src/backend/access/common/heaptuple.c:664-760
 *   slot_deform_tuple        src/backend/access/common/heaptuple.c:1119-1213
 *   att_* macros             src/include/access/tupmacs.h:23-175
 *   varlena header           src/include/postgres.h:158-300 (GPDB: big-endian always)
 *   PageInit / PageAddItem   src/backend/storage/page/bufpage.c:41-60,176-330
 *   ItemIdData               src/include/storage/itemid.h:24-40
 *   HeapTupleSatisfiesMVCC   src/backend/utils/time/tqual.c:997-1140 (frozen fast path only)
 *   heapgetpage              src/backend/access/heap/heapam.c:312-463
 *
 * Pinned by tests/golden/heap_kat.json: tuples formed by the reference's own
 * heap_form_tuple and deformed by its heap_deform_tuple (oracle/ref_build).
 */
#include <string.h>
#include "gg_oracle.h"
#include "or_internal.h"

/* tupmacs.h:121-130 */
static inline long
att_align_nominal(long off, char attalign)
{
	switch (attalign)
	{
		case 'i': return (off + 3) & ~3L;
		case 'c': return off;
		case 'd': return (off + 7) & ~7L;
		default:  return (off + 1) & ~1L;	/* 's' */
	}
}

/* VARATT_CAN_MAKE_SHORT + VARATT_CONVERTED_SHORT_SIZE (postgres.h:228-233): a
 * payload of n bytes becomes a 1-byte-header datum when n + 1 <= 0x7F. */
static inline int
varlena_stored_size(int payload, int *is_short)
{
	if (payload + 1 <= 0x7F)
	{
		*is_short = 1;
		return payload + 1;
	}
	*is_short = 0;
	return payload + 4;
}

/* heaptuple.c:68-129 (att_align_datum for packable varlenas, tupmacs.h:73-78) */
int
or_heap_compute_data_size(const gg_tupdesc *desc, const int64_t *val, const int32_t *len,
						  const uint8_t *isnull)
{
	long data_length = 0;
	int i;

	(void) val;
	for (i = 0; i < desc->natts; i++)
	{
		const gg_attr *att = &desc->attrs[i];

		if (isnull && isnull[i])
			continue;
		if (att->attlen == -1)
		{
			int is_short;
			int sz = varlena_stored_size(len[i], &is_short);

			if (!is_short)
				data_length = att_align_nominal(data_length, att->attalign);
			data_length += sz;
		}
		else
		{
			data_length = att_align_nominal(data_length, att->attalign);
			data_length += att->attlen;
		}
	}
	return (int) data_length;
}

/* heap_form_tuple (heaptuple.c:664-760) + heap_fill_tuple (:149-272).
 * The header is stamped the way a frozen, never-updated tuple looks on disk:
 * t_xmin = FrozenTransactionId, t_xmax = 0, HEAP_XMIN_FROZEN | HEAP_XMAX_INVALID.
 * t_ctid is left (0,0) and is set by the page builder. */
int
or_heap_form_tuple(const gg_tupdesc *desc, const int64_t *val, const int32_t *len,
				   const uint8_t *isnull, uint8_t *out, int outcap)
{
	int natts = desc->natts;
	int hasnull = 0;
	int hoff, data_len, total, i;
	uint16_t infomask = GG_HEAP_XMIN_FROZEN | GG_HEAP_XMAX_INVALID;
	uint8_t *data;
	uint8_t *bitP = NULL;
	int bitmask = 0;

	for (i = 0; i < natts; i++)
		if (isnull && isnull[i])
			hasnull = 1;

	hoff = GG_HEAP_HDR_SIZE;
	if (hasnull)
		hoff += (natts + 7) / 8;				/* BITMAPLEN */
	hoff = (int) GG_MAXALIGN(hoff);
	data_len = or_heap_compute_data_size(desc, val, len, isnull);
	total = hoff + data_len;
	if (total > outcap)
		return -1;
	memset(out, 0, (size_t) total);

	if (hasnull)
	{
		bitP = out + GG_HEAP_HDR_SIZE - 1;		/* &bit[-1] */
		bitmask = 0x80;							/* HIGHBIT */
		infomask |= GG_HEAP_HASNULL;
	}
	data = out + hoff;
	{
		uint8_t *start = data;

		for (i = 0; i < natts; i++)
		{
			const gg_attr *att = &desc->attrs[i];
			long off = data - start;

			if (hasnull)
			{
				if (bitmask != 0x80)
					bitmask <<= 1;
				else
				{
					bitP += 1;
					*bitP = 0;
					bitmask = 1;
				}
				if (isnull[i])
					continue;
				*bitP |= (uint8_t) bitmask;
			}
			if (att->attlen == -1)
			{
				int is_short;
				int sz = varlena_stored_size(len[i], &is_short);
				const uint8_t *payload = (const uint8_t *) (uintptr_t) val[i];

				infomask |= GG_HEAP_HASVARWIDTH;
				if (is_short)
				{
					/* SET_VARSIZE_1B: len | 0x80 (postgres.h:218) */
					data[0] = (uint8_t) (sz | 0x80);
					memcpy(data + 1, payload, (size_t) len[i]);
				}
				else
				{
					uint32_t l = (uint32_t) sz & 0x3FFFFFFF;

					off = att_align_nominal(off, att->attalign);
					data = start + off;
					/* SET_VARSIZE_4B: htonl(len) — network byte order on every platform (postgres.h:214) */
					data[0] = (uint8_t) (l >> 24);
					data[1] = (uint8_t) (l >> 16);
					data[2] = (uint8_t) (l >> 8);
					data[3] = (uint8_t) l;
					memcpy(data + 4, payload, (size_t) len[i]);
				}
				data += sz;
			}
			else
			{
				off = att_align_nominal(off, att->attalign);
				data = start + off;
				/* store_att_byval (tupmacs.h:177-200) */
				switch (att->attlen)
				{
					case 1: { int8_t v = (int8_t) val[i]; memcpy(data, &v, 1); break; }
					case 2: { int16_t v = (int16_t) val[i]; memcpy(data, &v, 2); break; }
					case 4: { int32_t v = (int32_t) val[i]; memcpy(data, &v, 4); break; }
					default: memcpy(data, &val[i], 8); break;
				}
				data += att->attlen;
			}
		}
	}

	/* header: t_xmin(4) t_xmax(4) t_cid(4) t_ctid(6) t_infomask2(2) t_infomask(2) t_hoff(1) */
	{
		uint32_t xmin = GG_FROZEN_XID;
		uint16_t infomask2 = (uint16_t) (natts & GG_HEAP_NATTS_MASK);

		memcpy(out + 0, &xmin, 4);
		memcpy(out + 18, &infomask2, 2);
		memcpy(out + 20, &infomask, 2);
		out[22] = (uint8_t) hoff;
	}
	return total;
}

/* bufpage.c:41 — PageInit(page, BLCKSZ, 0) */
void
or_page_init(uint8_t *page)
{
	uint16_t v;

	memset(page, 0, GG_BLCKSZ);
	v = GG_PAGE_HEADER_SIZE; memcpy(page + 12, &v, 2);		/* pd_lower */
	v = (uint16_t) GG_BLCKSZ; memcpy(page + 14, &v, 2);		/* pd_upper (32768 fits uint16) */
	memcpy(page + 16, &v, 2);								/* pd_special */
	v = (uint16_t) (GG_BLCKSZ | GG_PAGE_VERSION); memcpy(page + 18, &v, 2);	/* bufpage.h:207 */
}

void
or_page_set_all_visible(uint8_t *page)
{
	uint16_t f;

	memcpy(&f, page + 10, 2);
	f |= GG_PD_ALL_VISIBLE;
	memcpy(page + 10, &f, 2);
}

int
or_page_nitems(const uint8_t *page)
{
	uint16_t lower;

	memcpy(&lower, page + 12, 2);
	return lower <= GG_PAGE_HEADER_SIZE ? 0 : (lower - GG_PAGE_HEADER_SIZE) / GG_ITEMID_SIZE;
}

/* bufpage.c:176 — PageAddItem(page, item, size, InvalidOffsetNumber, false, true)
 * on a page without free line pointers: append at `limit`. Returns the offset
 * number (1-based) or 0 when the tuple does not fit.  pd_upper/pd_special of an
 * empty 32 KB page are 32768 = 0x8000, which still fits LocationIndex (uint16,
 * bufpage.h:100); lp_off has 15 bits but no tuple ever starts at 32768. */
int
or_page_add_item(uint8_t *page, const uint8_t *item, int size)
{
	uint16_t lower16, upper16;
	int lower, upper, offnum;
	uint32_t lp;

	memcpy(&lower16, page + 12, 2);
	memcpy(&upper16, page + 14, 2);
	upper = upper16;
	offnum = (lower16 - GG_PAGE_HEADER_SIZE) / GG_ITEMID_SIZE + 1;
	lower = lower16 + GG_ITEMID_SIZE;
	upper -= (int) GG_MAXALIGN(size);
	if (lower > upper)
		return 0;
	/* ItemIdSetNormal: lp_off:15 | lp_flags:2 | lp_len:15 (itemid.h:24-29) */
	lp = ((uint32_t) upper & 0x7FFF) | ((uint32_t) GG_LP_NORMAL << 15) | ((uint32_t) size << 17);
	memcpy(page + GG_PAGE_HEADER_SIZE + (offnum - 1) * GG_ITEMID_SIZE, &lp, 4);
	memcpy(page + upper, item, (size_t) size);
	lower16 = (uint16_t) lower;
	upper16 = (uint16_t) upper;
	memcpy(page + 12, &lower16, 2);
	memcpy(page + 14, &upper16, 2);
	return offnum;
}

/* VARSIZE_ANY / VARDATA_ANY (postgres.h:276-300) for inline, uncompressed datums */
const uint8_t *
or_varlena_payload(const uint8_t *p, int *len)
{
	if (p[0] & 0x80)
	{
		*len = (p[0] & 0x7F) - 1;				/* VARSIZE_1B - VARHDRSZ_SHORT */
		return p + 1;
	}
	*len = (int) ((((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3])
				  & 0x3FFFFFFF) - 4;
	return p + 4;
}

static inline int
varsize_any(const uint8_t *p)
{
	if (p[0] & 0x80)
		return p[0] & 0x7F;
	return (int) ((((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3])
				  & 0x3FFFFFFF);
}

/* slot_deform_tuple (heaptuple.c:1119-1213), without the attcacheoff memo:
 * the memo only short-circuits the same arithmetic.  att_align_pointer
 * (tupmacs.h:99-104) peeks at the byte to tell a pad byte from a 1-byte header. */
int
or_heap_deform(const gg_tupdesc *desc, const uint8_t *tup, int natts_wanted,
			   int64_t *values, uint8_t *isnull)
{
	uint16_t infomask, infomask2;
	int hasnulls, natts, attnum;
	const uint8_t *bp = tup + GG_HEAP_HDR_SIZE;
	const uint8_t *tp;
	long off = 0;

	memcpy(&infomask2, tup + 18, 2);
	memcpy(&infomask, tup + 20, 2);
	hasnulls = (infomask & GG_HEAP_HASNULL) != 0;
	natts = infomask2 & GG_HEAP_NATTS_MASK;
	if (natts > natts_wanted)
		natts = natts_wanted;
	tp = tup + tup[22];							/* t_hoff */

	for (attnum = 0; attnum < natts; attnum++)
	{
		const gg_attr *att = &desc->attrs[attnum];

		if (hasnulls && !(bp[attnum >> 3] & (1 << (attnum & 7))))	/* att_isnull */
		{
			values[attnum] = 0;
			isnull[attnum] = 1;
			continue;
		}
		isnull[attnum] = 0;
		if (att->attlen == -1)
		{
			if (tp[off] == 0)					/* pad byte (or aligned 4B header): align */
				off = att_align_nominal(off, att->attalign);
			values[attnum] = (int64_t) ((tp + off) - tup);
			off += varsize_any(tp + off);
		}
		else
		{
			off = att_align_nominal(off, att->attalign);
			switch (att->attlen)				/* fetch_att, tupmacs.h:44-70 */
			{
				case 1: { int8_t v; memcpy(&v, tp + off, 1); values[attnum] = v; break; }
				case 2: { int16_t v; memcpy(&v, tp + off, 2); values[attnum] = v; break; }
				case 4: { int32_t v; memcpy(&v, tp + off, 4); values[attnum] = v; break; }
				default: memcpy(&values[attnum], tp + off, 8); break;
			}
			off += att->attlen;
		}
	}
	/* attributes beyond the tuple's natts read as NULL (heaptuple.c:1252-1258) */
	for (; attnum < natts_wanted; attnum++)
	{
		values[attnum] = 0;
		isnull[attnum] = 1;
	}
	return natts_wanted;
}

/* tqual.c:997 HeapTupleSatisfiesMVCC, restricted to what can be decided
 * without clog or a snapshot: a frozen xmin is visible to everyone
 * (tqual.c:1009 HeapTupleHeaderXminFrozen) and an invalid xmax means never
 * deleted (tqual.c:1119).  Everything else needs the transaction machinery. */
int
or_tuple_visible(const uint8_t *tup)
{
	uint16_t infomask;

	memcpy(&infomask, tup + 20, 2);
	if ((infomask & GG_HEAP_XMIN_FROZEN) == GG_HEAP_XMIN_FROZEN)
	{
		if (infomask & GG_HEAP_XMAX_INVALID)
			return 1;
		return -1;
	}
	if ((infomask & GG_HEAP_XMIN_INVALID) && !(infomask & GG_HEAP_XMIN_COMMITTED))
		return 0;								/* tqual.c:1003: xmin aborted */
	return -1;
}

/* ---- HeapTupleSatisfiesMVCC (tqual.c:997-1238) against a snapshot ----
 * The snapshot of the scans that follow, like the one ExecutorStart leaves in es_snapshot for heap_beginscan
 * (heapam.c:1573); NULL: hint bits only (or_tuple_visible above).  Pinned against the reference's own tqual.o + transam.o
 * by tests/golden/mvcc_kat.json (oracle/ref_build/refwrap_tqual.c). */
static const gg_snapshot *or_snapshot;

void
or_set_snapshot(const gg_snapshot *snap)
{
	or_snapshot = snap;
}

/* TransactionIdPrecedes, transam.c:300: modulo-2^32 for normal xids, plain for the three special ones */
static int
or_xid_precedes(uint32_t a, uint32_t b)
{
	if (a < 3 || b < 3)
		return a < b;
	return (int32_t) (a - b) < 0;
}

/* TransactionIdDidCommit, transam.c:125 via TransactionLogFetch :52-100: bootstrap and frozen xids count as
 * committed, the invalid xid as aborted, everything else is what pg_clog says.  -1: not answerable from the bits given
 * (outside the range, or sub-committed: transam.c:146 would ask pg_subtrans for the parent) */
static int
or_xid_did_commit(uint32_t xid, const gg_snapshot *snap)
{
	uint32_t d;
	int st;

	if (xid == 1 || xid == 2)
		return 1;
	if (xid == 0)
		return 0;
	d = xid - snap->clog_base;
	if (d >= snap->clog_n)
		return -1;
	st = (snap->clog[d >> 2] >> ((d & 3) * 2)) & 3;		/* clog.c:73-78 */
	if (st == 3)
		return -1;
	return st == 1;
}

/* XidInMVCCSnapshot_Local, tqual.c:1600-1650 (no subxip, not overflowed, not taken during recovery) */
static int
or_xid_in_snapshot(uint32_t xid, const gg_snapshot *snap)
{
	uint32_t i;

	if (or_xid_precedes(xid, snap->xmin))
		return 0;
	if (!or_xid_precedes(xid, snap->xmax))				/* TransactionIdFollowsOrEquals */
		return 1;
	for (i = 0; i < snap->xcnt; i++)
		if (snap->xip[i] == xid)
			return 1;
	return 0;
}

/* 1 visible, 0 not, -1 the rule needs the server (multixact, combo cid, moved tuples, unknown status).
 * TransactionIdIsInProgress (procarray) and "did not commit" lead to the same answer on both the xmin and the xmax
 * side (tqual.c:1100-1110, 1198-1209), so the status bits alone decide. */
int
or_tuple_satisfies_mvcc(const uint8_t *tup, const gg_snapshot *snap)
{
	uint32_t xmin, xmax, cid;
	uint16_t infomask;
	int locked_only, c;

	memcpy(&xmin, tup, 4);
	memcpy(&xmax, tup + 4, 4);
	memcpy(&cid, tup + 8, 4);
	memcpy(&infomask, tup + 20, 2);
	/* HEAP_XMAX_IS_LOCKED_ONLY, htup_details.h:211 */
	locked_only = (infomask & GG_HEAP_XMAX_LOCK_ONLY) ||
		(infomask & (GG_HEAP_XMAX_IS_MULTI | GG_HEAP_XMAX_EXCL_LOCK | GG_HEAP_XMAX_KEYSHR_LOCK)) == GG_HEAP_XMAX_EXCL_LOCK;

	if (!(infomask & GG_HEAP_XMIN_COMMITTED))				/* tqual.c:1009 */
	{
		if (infomask & GG_HEAP_XMIN_INVALID)				/* :1011 */
			return 0;
		if (infomask & GG_HEAP_MOVED)						/* :1015-1052, pre-9.0 VACUUM FULL */
			return -1;
		if (snap->own_xid && xmin == snap->own_xid)			/* :1053 TransactionIdIsCurrentTransactionId */
		{
			if (infomask & GG_HEAP_COMBOCID)
				return -1;
			if (cid >= snap->curcid)						/* :1055 inserted after scan started */
				return 0;
			if (infomask & GG_HEAP_XMAX_INVALID)			/* :1058 */
				return 1;
			if (locked_only)								/* :1061 */
				return 1;
			if (infomask & GG_HEAP_XMAX_IS_MULTI)			/* :1064 */
				return -1;
			if (xmax != snap->own_xid)						/* :1083 deleting subtransaction aborted */
				return 1;
			return cid >= snap->curcid;						/* :1097-1100 */
		}
		c = or_xid_did_commit(xmin, snap);					/* :1102-1113 */
		if (c < 0)
			return -1;
		if (!c)
			return 0;
	}
	/* :1120-1135 the inserting transaction committed: before the snapshot? */
	if ((infomask & GG_HEAP_XMIN_FROZEN) != GG_HEAP_XMIN_FROZEN && or_xid_in_snapshot(xmin, snap))
		return 0;
	if (infomask & GG_HEAP_XMAX_INVALID)					/* :1137 */
		return 1;
	if (locked_only)										/* :1140 */
		return 1;
	if (infomask & GG_HEAP_XMAX_IS_MULTI)					/* :1143 */
		return -1;
	if (!(infomask & GG_HEAP_XMAX_COMMITTED))				/* :1186 */
	{
		if (snap->own_xid && xmax == snap->own_xid)			/* :1188 */
		{
			if (infomask & GG_HEAP_COMBOCID)
				return -1;
			return cid >= snap->curcid;
		}
		c = or_xid_did_commit(xmax, snap);					/* :1196-1209 */
		if (c < 0)
			return -1;
		if (!c)
			return 1;
	}
	return or_xid_in_snapshot(xmax, snap);					/* :1219-1233 */
}

/* ---- forward page-mode scan: heapgetpage (heapam.c:312-463) + heapgettup_pagemode (:767-1006) ---- */

void
or_scan_begin(or_heapscan *s, const gg_tupdesc *desc, const uint8_t *pages, uint64_t nblocks)
{
	s->desc = desc;
	s->pages = pages;
	s->nblocks = nblocks;
	s->cblock = 0;
	s->inited = 0;
	s->ntuples = 0;
	s->cindex = 0;
	s->error = 0;
}

/* heapgetpage: collect the visible LP_NORMAL items of one page into rs_vistuples[] */
static void
or_scan_getpage(or_heapscan *s, uint64_t blk)
{
	const uint8_t *dp = s->pages + blk * (uint64_t) GG_BLCKSZ;
	uint16_t flags;
	int lines = or_page_nitems(dp);
	int all_visible, lineoff, n = 0;

	memcpy(&flags, dp + 10, 2);
	all_visible = (flags & GG_PD_ALL_VISIBLE) != 0;		/* heapam.c:391 */
	for (lineoff = 1; lineoff <= lines; lineoff++)
	{
		uint32_t lp;

		memcpy(&lp, dp + GG_PAGE_HEADER_SIZE + (lineoff - 1) * GG_ITEMID_SIZE, 4);
		if (((lp >> 15) & 3) != GG_LP_NORMAL)				/* ItemIdIsNormal */
			continue;
		if (!all_visible)
		{
			int v = or_tuple_visible(dp + (lp & 0x7FFF));

			if (v < 0 && or_snapshot)
				v = or_tuple_satisfies_mvcc(dp + (lp & 0x7FFF), or_snapshot);
			if (v < 0)
				s->error = OR_ERR_VISIBILITY;
			if (v <= 0)
				continue;
		}
		s->vistuples[n++] = (uint16_t) lineoff;
	}
	s->ntuples = n;
	s->cblock = blk;
	s->cindex = 0;
}

/* heap_getnext: returns the next visible tuple or NULL at end of relation */
const uint8_t *
or_scan_next(or_heapscan *s, uint64_t *tid)
{
	for (;;)
	{
		if (!s->inited)
		{
			if (s->nblocks == 0)
				return NULL;
			or_scan_getpage(s, 0);
			s->inited = 1;
		}
		if (s->cindex < s->ntuples)
		{
			const uint8_t *dp = s->pages + s->cblock * (uint64_t) GG_BLCKSZ;
			int lineoff = s->vistuples[s->cindex++];
			uint32_t lp;

			memcpy(&lp, dp + GG_PAGE_HEADER_SIZE + (lineoff - 1) * GG_ITEMID_SIZE, 4);
			if (tid)
				*tid = (s->cblock << 16) | (uint64_t) lineoff;
			return dp + (lp & 0x7FFF);
		}
		if (s->cblock + 1 >= s->nblocks)
			return NULL;
		or_scan_getpage(s, s->cblock + 1);
	}
}
