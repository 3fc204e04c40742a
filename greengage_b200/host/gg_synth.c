/*
 * gg_synth.c — synthetic TPC-H-shaped relation loader (see include/gg_synth.h).
 *
 * Writes the on-disk format directly:
 *   page   PageInit + PageAddItem            src/backend/storage/page/bufpage.c:41,176
 *   tuple  heap_form_tuple / heap_fill_tuple src/backend/access/common/heaptuple.c:149,664
 *   short varlena headers (big-endian in GPDB)  src/include/postgres.h:158-219
 * tests/test_synth.py checks every generated tuple against the oracle's (and, when built, the
 * reference's own) heap_form_tuple / heap_deform_tuple.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/gg_synth.h"
#include "../../include/gg_aocs.h"
#include "gg_hostutil.h"

#define D_1992_01_02 (-2921)
#define D_1998_12_01 (-396)
#define D_1995_06_17 (-1659)
#define D_1992_01_01 (-2922)
#define D_1998_08_02 (-517)

static inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
/* counter-based random stream: value k of candidate c */
static inline uint64_t rnd(const gg_synth_spec *s, uint64_t c, uint32_t k)
{
	return mix64(s->seed + mix64(c * 0x9E3779B97F4A7C15ULL + ((uint64_t) s->table << 56) + k));
}
static inline int64_t uni(const gg_synth_spec *s, uint64_t c, uint32_t k, int64_t lo, int64_t hi)
{
	return lo + (int64_t) (rnd(s, c, k) % (uint64_t) (hi - lo + 1));
}

int64_t gg_synth_orderkey(uint64_t o)
{
	return (int64_t) ((o >> 3) * 32 + (o & 7) + 1);
}

static void set_attr(gg_attr *a, int32_t typ, int16_t len, char align, int byval, int32_t typmod)
{
	memset(a, 0, sizeof *a);
	a->atttypid = typ; a->attlen = len; a->attalign = align; a->attbyval = (int8_t) byval;
	a->attnotnull = 1; a->atttypmod = typmod;
}

int gg_synth_tupdesc(int table, gg_tupdesc *d)
{
	int i = 0;
	memset(d, 0, sizeof *d);
#define I8()   set_attr(&d->attrs[i++], GG_INT8OID, 8, 'd', 1, -1)
#define I4()   set_attr(&d->attrs[i++], GG_INT4OID, 4, 'i', 1, -1)
#define F8()   set_attr(&d->attrs[i++], GG_FLOAT8OID, 8, 'd', 1, -1)
#define DT()   set_attr(&d->attrs[i++], GG_DATEOID, 4, 'i', 1, -1)
#define BP(n)  set_attr(&d->attrs[i++], GG_BPCHAROID, -1, 'i', 0, (n) + 4)
#define VC(n)  set_attr(&d->attrs[i++], GG_VARCHAROID, -1, 'i', 0, (n) + 4)
	switch (table)
	{
		case GG_TAB_LINEITEM_WIDE:      /* tpch500GB.sql:66-83 */
			I8(); I4(); I4(); I4(); F8(); F8(); F8(); F8(); BP(1); BP(1); DT(); DT(); DT(); BP(25); BP(10); VC(44);
			break;
		case GG_TAB_LINEITEM_NARROW:
			I8(); F8(); F8(); F8(); F8(); BP(1); BP(1); DT();
			break;
		case GG_TAB_ORDERS:             /* tpch500GB.sql:102-112 */
			I8(); I4(); BP(1); F8(); DT(); BP(15); BP(15); I4(); VC(79);
			break;
		default:
			return -1;
	}
	d->natts = i;
	return 0;
}

static const char *const SHIPINSTRUCT[4] = { "DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN" };
static const char *const SHIPMODE[7] = { "REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB" };
static const char *const PRIORITY[5] = { "1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW" };
static const char ALPHA[33] = "abcdefghijklmnopqrstuvwxyz ,.-;: ";

static inline int64_t f8bits(double d) { int64_t v; memcpy(&v, &d, 8); return v; }

static int padded(char *dst, const char *src, int n)
{
	int l = (int) strlen(src);
	memcpy(dst, src, (size_t) l);
	memset(dst + l, ' ', (size_t) (n - l));
	return n;
}

static int comment(const gg_synth_spec *s, uint64_t c, uint32_t k, int lo, int hi, char *dst)
{
	int len = (int) uni(s, c, k, lo, hi), i;
	uint64_t r = 0;
	for (i = 0; i < len; i++)
	{
		if ((i & 7) == 0) r = rnd(s, c, k + 1 + (uint32_t) (i >> 3));
		dst[i] = ALPHA[(r >> ((i & 7) * 8)) & 31];
	}
	if (dst[len - 1] == ' ') dst[len - 1] = 'x';     /* varchar keeps trailing blanks; avoid ambiguity anyway */
	return len;
}

static int owner_seg(const gg_synth_spec *s, uint64_t c, int64_t distkey)
{
	if (s->nsegs <= 1) return 0;
	if (s->policy == GG_DIST_HASH) return ggh_seg_of_int8(distkey, s->nsegs);
	return (int) (c % (uint64_t) s->nsegs);
}

/* Row values of candidate c.  strbuf receives the string payloads back to back. */
int gg_synth_row(const gg_synth_spec *s, uint64_t c, int64_t *v, int32_t *len, char *sb, int cap, int *mine)
{
	int i, sp = 0;
	if (cap < 200) return -1;
	for (i = 0; i < GG_MAX_ATTS; i++) len[i] = 0;
	if (s->table == GG_TAB_ORDERS)
	{
		int64_t okey = gg_synth_orderkey(c);
		if (mine) *mine = owner_seg(s, c, okey) == s->seg;
		v[0] = okey;
		v[1] = uni(s, c, 1, 1, (int64_t) (s->ncand / 10 > 0 ? s->ncand / 10 : 1));
		sb[sp] = "FOP"[uni(s, c, 2, 0, 2)]; v[2] = (int64_t) (intptr_t) (sb + sp); len[2] = 1; sp += 1;
		v[3] = f8bits((double) uni(s, c, 3, 100000, 50000000) / 100.0);
		v[4] = uni(s, c, 4, D_1992_01_01, D_1998_08_02);
		v[5] = (int64_t) (intptr_t) (sb + sp); len[5] = padded(sb + sp, PRIORITY[uni(s, c, 5, 0, 4)], 15); sp += 15;
		{
			char clerk[16];
			int64_t id = uni(s, c, 6, 1, 1000), p;
			memcpy(clerk, "Clerk#000000000", 15);
			for (p = 14; id > 0 && p >= 6; p--, id /= 10) clerk[p] = (char) ('0' + id % 10);
			memcpy(sb + sp, clerk, 15);
			v[6] = (int64_t) (intptr_t) (sb + sp); len[6] = 15; sp += 15;
		}
		v[7] = 0;
		v[8] = (int64_t) (intptr_t) (sb + sp); len[8] = comment(s, c, 16, 19, 78, sb + sp); sp += len[8];
		return 9;
	}
	else
	{
		uint64_t o = rnd(s, c, 0) % (s->norders ? s->norders : 1);
		int64_t okey = gg_synth_orderkey(o);
		int64_t qty = uni(s, c, 4, 1, 50);
		int64_t cents = uni(s, c, 5, 90000, 209999);
		int64_t shipdate = uni(s, c, 10, D_1992_01_02, D_1998_12_01);
		int64_t commitdate = shipdate + uni(s, c, 11, -30, 30);
		int64_t receiptdate = shipdate + uni(s, c, 12, 1, 30);
		char flag = receiptdate > D_1995_06_17 ? 'N' : (uni(s, c, 8, 0, 1) ? 'R' : 'A');
		char status = shipdate > D_1995_06_17 ? 'O' : 'F';
		double price = (double) (qty * cents) / 100.0;
		double disc = (double) uni(s, c, 6, 0, 10) / 100.0;
		double tax = (double) uni(s, c, 7, 0, 8) / 100.0;

		if (mine) *mine = owner_seg(s, c, okey) == s->seg;
		if (s->table == GG_TAB_LINEITEM_NARROW)
		{
			v[0] = okey; v[1] = f8bits((double) qty); v[2] = f8bits(price); v[3] = f8bits(disc); v[4] = f8bits(tax);
			sb[sp] = flag; v[5] = (int64_t) (intptr_t) (sb + sp); len[5] = 1; sp++;
			sb[sp] = status; v[6] = (int64_t) (intptr_t) (sb + sp); len[6] = 1; sp++;
			v[7] = shipdate;
			return 8;
		}
		v[0] = okey;
		v[1] = uni(s, c, 1, 1, (int64_t) (s->ncand / 30 > 1000 ? s->ncand / 30 : 1000));
		v[2] = uni(s, c, 2, 1, (int64_t) (s->ncand / 600 > 100 ? s->ncand / 600 : 100));
		v[3] = uni(s, c, 3, 1, 7);
		v[4] = f8bits((double) qty); v[5] = f8bits(price); v[6] = f8bits(disc); v[7] = f8bits(tax);
		sb[sp] = flag; v[8] = (int64_t) (intptr_t) (sb + sp); len[8] = 1; sp++;
		sb[sp] = status; v[9] = (int64_t) (intptr_t) (sb + sp); len[9] = 1; sp++;
		v[10] = shipdate; v[11] = commitdate; v[12] = receiptdate;
		v[13] = (int64_t) (intptr_t) (sb + sp); len[13] = padded(sb + sp, SHIPINSTRUCT[uni(s, c, 13, 0, 3)], 25); sp += 25;
		v[14] = (int64_t) (intptr_t) (sb + sp); len[14] = padded(sb + sp, SHIPMODE[uni(s, c, 14, 0, 6)], 10); sp += 10;
		v[15] = (int64_t) (intptr_t) (sb + sp); len[15] = comment(s, c, 16, 10, 43, sb + sp); sp += len[15];
		return 16;
	}
}

/* heap_form_tuple for a row without NULLs: 24-byte header (t_hoff = MAXALIGN(23)), then the
 * attributes with att_align_nominal for fixed-width ones and 1-byte-header varlenas unaligned. */
static int form_tuple(const gg_tupdesc *d, const int64_t *v, const int32_t *len, uint8_t *out)
{
	uint32_t off = 24;
	uint16_t infomask = GG_HEAP_XMIN_FROZEN | GG_HEAP_XMAX_INVALID;
	uint16_t infomask2 = (uint16_t) d->natts;
	uint32_t xmin = GG_FROZEN_XID;
	int i;

	memset(out, 0, 24);
	for (i = 0; i < d->natts; i++)
	{
		const gg_attr *a = &d->attrs[i];
		if (a->attlen == -1)
		{
			uint32_t n = (uint32_t) len[i];
			infomask |= GG_HEAP_HASVARWIDTH;
			if (n + 1 <= 0x7F)
			{
				out[off] = (uint8_t) ((n + 1) | 0x80);
				memcpy(out + off + 1, (const void *) (intptr_t) v[i], n);
				off += n + 1;
			}
			else
			{
				uint32_t o2 = (off + 3) & ~3u, l = n + 4;
				memset(out + off, 0, o2 - off);
				off = o2;
				out[off] = (uint8_t) (l >> 24); out[off + 1] = (uint8_t) (l >> 16);
				out[off + 2] = (uint8_t) (l >> 8); out[off + 3] = (uint8_t) l;
				memcpy(out + off + 4, (const void *) (intptr_t) v[i], n);
				off += l;
			}
		}
		else
		{
			uint32_t m = a->attalign == 'd' ? 7u : a->attalign == 'i' ? 3u : a->attalign == 's' ? 1u : 0u;
			uint32_t o2 = (off - 24 + m) & ~m;
			o2 += 24;
			memset(out + off, 0, o2 - off);
			off = o2;
			if (a->attlen == 8) memcpy(out + off, &v[i], 8);
			else if (a->attlen == 4) { int32_t x = (int32_t) v[i]; memcpy(out + off, &x, 4); }
			else if (a->attlen == 2) { int16_t x = (int16_t) v[i]; memcpy(out + off, &x, 2); }
			else out[off] = (uint8_t) v[i];
			off += (uint32_t) a->attlen;
		}
	}
	memcpy(out, &xmin, 4);
	memcpy(out + 18, &infomask2, 2);
	memcpy(out + 20, &infomask, 2);
	out[22] = 24;
	return (int) off;
}

typedef struct page_writer {
	uint8_t *pages;              /* NULL: measure only */
	uint64_t blk;                /* current block index (relative to the extent's first block) */
	uint64_t base;               /* first block of the extent */
	uint32_t lower, upper;
	int open;
	uint64_t nrows;
} page_writer;

static void page_open(page_writer *w)
{
	w->lower = GG_PAGE_HEADER_SIZE;
	w->upper = GG_BLCKSZ;
	w->open = 1;
	if (w->pages)
	{
		uint8_t *p = w->pages + (w->base + w->blk) * (uint64_t) GG_BLCKSZ;
		uint16_t x;
		memset(p, 0, GG_BLCKSZ);
		x = GG_PD_ALL_VISIBLE; memcpy(p + 10, &x, 2);
		x = (uint16_t) GG_BLCKSZ; memcpy(p + 16, &x, 2);
		x = (uint16_t) (GG_BLCKSZ | GG_PAGE_VERSION); memcpy(p + 18, &x, 2);
	}
}

static void page_close(page_writer *w)
{
	if (!w->open) return;
	if (w->pages)
	{
		uint8_t *p = w->pages + (w->base + w->blk) * (uint64_t) GG_BLCKSZ;
		uint16_t x;
		x = (uint16_t) w->lower; memcpy(p + 12, &x, 2);
		x = (uint16_t) w->upper; memcpy(p + 14, &x, 2);
	}
	w->open = 0;
	w->blk++;
}

static void page_add(page_writer *w, const uint8_t *tup, int len)
{
	uint32_t aligned = (uint32_t) GG_MAXALIGN(len);
	if (w->open && w->lower + GG_ITEMID_SIZE > w->upper - aligned) page_close(w);
	if (!w->open) page_open(w);
	w->upper -= aligned;
	if (w->pages)
	{
		uint8_t *p = w->pages + (w->base + w->blk) * (uint64_t) GG_BLCKSZ;
		uint32_t offnum = (w->lower - GG_PAGE_HEADER_SIZE) / GG_ITEMID_SIZE + 1;
		uint32_t lp = (w->upper & 0x7FFF) | ((uint32_t) GG_LP_NORMAL << 15) | ((uint32_t) len << 17);
		uint8_t *t = p + w->upper;
		uint64_t b = w->base + w->blk;
		uint16_t bh = (uint16_t) (b >> 16), bl = (uint16_t) b, on = (uint16_t) offnum;
		memcpy(p + w->lower, &lp, 4);
		memcpy(t, tup, (size_t) len);
		/* t_ctid = (block, offset) of the tuple itself (ItemPointerData: bi_hi, bi_lo, ip_posid) */
		memcpy(t + 12, &bh, 2); memcpy(t + 14, &bl, 2); memcpy(t + 16, &on, 2);
	}
	w->lower += GG_ITEMID_SIZE;
	w->nrows++;
}

typedef struct extent_job {
	const gg_synth_spec *spec;
	gg_tupdesc desc;
	uint64_t *ext_blocks;        /* [nextents] out (measure) / in (generate: prefix offsets) */
	uint64_t *ext_rows;
	uint64_t nextents;
	uint8_t *pages;
	volatile uint64_t *next;
} extent_job;

static void run_extent(const extent_job *j, uint64_t e, uint64_t base, uint64_t *nblk, uint64_t *nrow)
{
	const gg_synth_spec *s = j->spec;
	uint64_t c0 = e * GG_SYNTH_EXTENT, c1 = c0 + GG_SYNTH_EXTENT, c;
	page_writer w;
	int64_t v[GG_MAX_ATTS];
	int32_t len[GG_MAX_ATTS];
	char sb[256];
	uint8_t tup[512];

	if (c1 > s->ncand) c1 = s->ncand;
	memset(&w, 0, sizeof w);
	w.pages = j->pages;
	w.base = base;
	for (c = c0; c < c1; c++)
	{
		int mine, tl;
		if (s->nsegs > 1 && s->policy == GG_DIST_RANDOM && (int) (c % (uint64_t) s->nsegs) != s->seg) continue;
		gg_synth_row(s, c, v, len, sb, sizeof sb, &mine);
		if (!mine) continue;
		tl = form_tuple(&j->desc, v, len, tup);
		page_add(&w, tup, tl);
	}
	page_close(&w);
	*nblk = w.blk;
	*nrow = w.nrows;
}

static void *worker(void *p)
{
	extent_job *j = p;
	for (;;)
	{
		uint64_t e = __sync_fetch_and_add(j->next, 1), nb, nr;
		if (e >= j->nextents) break;
		if (j->pages)
			run_extent(j, e, j->ext_blocks[e], &nb, &nr);
		else
		{
			run_extent(j, e, 0, &nb, &nr);
			j->ext_blocks[e] = nb;
			j->ext_rows[e] = nr;
		}
	}
	return NULL;
}

static int run_all(const gg_synth_spec *spec, int nthreads, uint8_t *pages, uint64_t cap,
                   uint64_t *nblocks, uint64_t *nrows)
{
	extent_job j;
	uint64_t next = (spec->ncand + GG_SYNTH_EXTENT - 1) / GG_SYNTH_EXTENT, e, tot = 0, rows = 0;
	volatile uint64_t counter = 0;
	pthread_t th[256];
	int t;

	if (gg_synth_tupdesc(spec->table, &j.desc)) return -1;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	j.spec = spec;
	j.nextents = next;
	j.ext_blocks = calloc(next + 1, sizeof(uint64_t));
	j.ext_rows = calloc(next + 1, sizeof(uint64_t));
	if (j.ext_blocks == NULL || j.ext_rows == NULL) { free(j.ext_blocks); free(j.ext_rows); return -3; }
	j.pages = NULL;
	j.next = &counter;
	for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
	for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	for (e = 0; e < next; e++)
	{
		uint64_t nb = j.ext_blocks[e];
		j.ext_blocks[e] = tot;              /* exclusive prefix: first block of the extent */
		tot += nb;
		rows += j.ext_rows[e];
	}
	if (nblocks) *nblocks = tot;
	if (nrows) *nrows = rows;
	if (pages)
	{
		if (tot > cap) { free(j.ext_blocks); free(j.ext_rows); return -2; }
		j.pages = pages;
		counter = 0;
		for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
		for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	}
	free(j.ext_blocks);
	free(j.ext_rows);
	return 0;
}

int gg_synth_measure(const gg_synth_spec *spec, int nthreads, uint64_t *nblocks, uint64_t *nrows)
{
	return run_all(spec, nthreads, NULL, 0, nblocks, nrows);
}

int gg_synth_generate(const gg_synth_spec *spec, int nthreads, uint8_t *pages, uint64_t cap_blocks,
                      uint64_t *nblocks, uint64_t *nrows)
{
	return run_all(spec, nthreads, pages, cap_blocks, nblocks, nrows);
}

/* ---- the same relation as AOCS column files ---- */

typedef struct aocs_job {
	const gg_synth_spec *spec;
	gg_tupdesc desc;
	const int32_t *cols;
	int ncols;
	uint8_t *const *out;
	const int64_t *outcap;
	int blocksize, checksum;
	int64_t *outbytes;
	uint64_t *colrows;
	int *rc;
	volatile uint64_t *next;
} aocs_job;

static void *aocs_worker(void *p)
{
	aocs_job *j = p;
	const gg_synth_spec *s = j->spec;

	for (;;)
	{
		uint64_t k = __sync_fetch_and_add(j->next, 1), c, rows = 0;
		gg_aocs_writer *w;
		int64_t v[GG_MAX_ATTS];
		int32_t len[GG_MAX_ATTS];
		char sb[256];
		int col, rc;

		if (k >= (uint64_t) j->ncols) break;
		col = j->cols[k];
		rc = gg_aocs_writer_create(&j->desc.attrs[col], j->blocksize, j->checksum, 1, j->out[k], j->outcap[k], &w);
		for (c = 0; rc == GG_OK && c < s->ncand; c++)
		{
			int mine;
			if (s->nsegs > 1 && s->policy == GG_DIST_RANDOM && (int) (c % (uint64_t) s->nsegs) != s->seg) continue;
			gg_synth_row(s, c, v, len, sb, sizeof sb, &mine);
			if (!mine) continue;
			rc = gg_aocs_writer_put(w, v[col], len[col], 0);
			rows++;
		}
		if (rc == GG_OK)
			rc = gg_aocs_writer_finish(w, &j->outbytes[k]);
		else if (w != NULL)
			(void) gg_aocs_writer_finish(w, NULL);
		j->colrows[k] = rows;
		j->rc[k] = rc;
	}
	return NULL;
}

int gg_synth_aocs_generate(const gg_synth_spec *spec, int nthreads, const int32_t *cols, int ncols,
                           uint8_t *const *out, const int64_t *outcap, int blocksize, int checksum,
                           int64_t *outbytes, uint64_t *nrows)
{
	aocs_job j;
	pthread_t th[GG_MAX_ATTS];
	uint64_t colrows[GG_MAX_ATTS];
	int rcs[GG_MAX_ATTS], t, i, rc = 0;
	volatile uint64_t counter = 0;

	if (spec == NULL || cols == NULL || out == NULL || outcap == NULL || outbytes == NULL || ncols < 1 || ncols > GG_MAX_ATTS)
		return -1;
	if (gg_synth_tupdesc(spec->table, &j.desc)) return -1;
	for (i = 0; i < ncols; i++)
		if (cols[i] < 0 || cols[i] >= j.desc.natts) return -1;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > ncols) nthreads = ncols;
	j.spec = spec; j.cols = cols; j.ncols = ncols; j.out = out; j.outcap = outcap;
	j.blocksize = blocksize; j.checksum = checksum; j.outbytes = outbytes; j.colrows = colrows; j.rc = rcs;
	j.next = &counter;
	for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, aocs_worker, &j);
	for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	for (i = 0; i < ncols; i++)
		if (rcs[i]) rc = rcs[i];
	if (nrows) *nrows = colrows[0];
	return rc;
}
