/*
 * or_aocs.c — ORACLE (test infrastructure only): CPU restatement of the column-oriented append-only (AOCS)
 * on-disk format, compresstype=none, and of a SeqScan over it feeding the aggregate of or_agg.c.
 * the first format the accelerated path widens to after heap pages.
 *
 * An AOCS relation keeps one file per column per segment file number.  A column file is a sequence of
 * Append-Only STORAGE BLOCKS (src/include/cdb/cdbappendonlystorage_int.h:18-150):
 *     bytes 0..7    AOSmallContentHeader, two native-endian uint32:
 *                     w0: reserved0:1 | headerKind:3 (=1 SmallContent) | hasFirstRowNum:1 | executorBlockKind:3 |
 *                         rowCount:14 | dataLength[20..11]:10
 *                     w1: dataLength[10..0]:11 | compressedLength:21 (0 = stored uncompressed)
 *     bytes 8..11   block checksum   CRC-32C over everything after the header checksum      } only when the
 *     bytes 12..15  header checksum  CRC-32C over bytes 0..11                               } table has checksum=true
 *                   (neither is bit-inverted at the end: cdbappendonlystorageformat.c:38-47,67-76)
 *     next 8        firstRowNum (int64) when hasFirstRowNum — always, for column files (datumstream.c:895-899)
 *     then          dataLength bytes of content, zero padded to a multiple of 8 (AOStorage_RoundUp, format version >= 2)
 * and the content of an executorBlockKind 1 (AOCSBK_BLOCK) block is a DATUM STREAM BLOCK, "Original" version
 * (src/include/utils/datumstreamblock.h:68-81; writer datumstreamblock.c:1486-1748,3644-3710; reader :150-330 and
 * datumstreamblock.h:1216-1570):
 *     int16 version (0) | int16 flags (1 = has NULL bitmap) | int16 ndatum (rows incl. NULLs) | int16 unused
 *     int32 nullsz (MAXALIGNed bitmap bytes) | int32 sz (bytes of datum data)
 *     NULL bitmap, one bit per row, LSB first, 1 = NULL, zero padded to nullsz      (only when flags & 1)
 *     datum data from MAXALIGN(16 + nullsz): the non-NULL values only, in row order:
 *         fixed length   attlen bytes each, native endian, no padding between items
 *         varlena        the heap's own inline form — 1-byte header when the value fits 126 payload bytes, else the
 *                        4-byte header aligned to typalign with zero padding in front (att_align_zero)
 *
 * Every function cites what it restates.  Pinned by tests/golden/aocs_kat.json: column files WRITTEN BY THE REFERENCE'S
 * OWN datumstreamblock.o / cdbappendonlystorageformat.o (oracle/ref_build/refwrap_aocs.c) must be read back identically,
 * and or_aocs_write_column must reproduce them byte for byte.  Large objects (AOCSBK_BLOB), bulk compression and the
 * Dense (RLE_TYPE / delta) block versions are out of scope of this restatement and are refused.
 */
#include <stdlib.h>
#include <string.h>
#include "or_internal.h"

#define AO_HDR 8
#define AO_KIND_SMALLCONTENT 1
#define AOCSBK_BLOCK 1
#define AO_MAXROWS 0x3FFF					/* AOSmallContentHeader_MaxRowCount = MAXDATUM_PER_AOCS_ORIG_BLOCK */
#define DSB_HDR 16
#define DSB_HAS_NULLBITMAP 1
#define MAXALIGN8(x) (((x) + 7) & ~(int64_t) 7)

/* ---- CRC-32C as the storage layer uses it (port/pg_crc32c_sb8.c; INIT 0xFFFFFFFF, no final inversion) ---- */
static uint32_t crc_table[256];

static void
crc_init(void)
{
	uint32_t i, j, c;

	for (i = 0; i < 256; i++)
	{
		c = i;
		for (j = 0; j < 8; j++)
			c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;	/* reflected Castagnoli polynomial */
		crc_table[i] = c;
	}
}

uint32_t
or_aocs_crc32c(const uint8_t *p, int64_t n)
{
	uint32_t c = 0xFFFFFFFFu;
	int64_t i;

	if (crc_table[1] == 0)
		crc_init();
	for (i = 0; i < n; i++)
		c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
	return c;
}

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

static inline int64_t
align_off(int64_t off, int attalign)		/* att_align_nominal (tupmacs.h:121-130) on an offset */
{
	switch (attalign)
	{
		case 'i': return (off + 3) & ~(int64_t) 3;
		case 'd': return (off + 7) & ~(int64_t) 7;
		case 's': return (off + 1) & ~(int64_t) 1;
		default:  return off;
	}
}

static inline int
ao_header_len(int checksum)					/* AoHeader_Size(false, checksum, true), cdbappendonlystorage.h:94-113 */
{
	return AO_HDR + (checksum ? 8 : 0) + 8;
}

/* ------------------------------------------------ writer ------------------------------------------------ */

typedef struct blockwriter {				/* DatumStreamBlockWrite, the Original-version members */
	const gg_attr *att;
	int maxrows, maxdata;					/* maxDatumPerBlock, maxDataBlockSize */
	int nth, nullcount_bits;				/* nth; always_null_bitmap_count */
	int has_null;
	uint8_t *nullmap;						/* one BYTE per row here; packed when the block is formatted */
	uint8_t *data;
	int64_t datalen;						/* datump - datum_buffer */
} blockwriter;

/* DatumStreamBlockWrite_OrigHasSpace (datumstreamblock.c:1486-1543) */
static int
orig_has_space(const blockwriter *w, int null, int sz)
{
	int64_t nullsize;

	if (w->nth + 1 >= w->maxrows)
		return 0;
	nullsize = (null || w->has_null) ? MAXALIGN8((w->nullcount_bits + 1 + 7) >> 3) : 0;
	return DSB_HDR + nullsize + w->datalen + sz < w->maxdata;
}

/* DatumStreamBlockWrite_PutOrig (datumstreamblock.c:1546-1748): 0/size on success, < 0 when the block is full */
static int
orig_put(blockwriter *w, int64_t value, int32_t len, int null)
{
	const gg_attr *att = w->att;

	if (null)
	{
		if (!orig_has_space(w, 1, 0))
			return -1;
		w->has_null = 1;					/* MakeNullBitMapSpace: earlier rows of the block become 0 bits */
		w->nullmap[w->nth] = 1;
		w->nullcount_bits++;
		w->nth++;
		return 0;
	}
	if (att->attlen == -1)
	{
		const uint8_t *payload = (const uint8_t *) (uintptr_t) value;
		int sz;

		if (len + 1 <= 0x7F)				/* value_type_could_short: VARATT_CONVERTED_SHORT_SIZE, 1-byte header */
		{
			sz = len + 1;
			if (!orig_has_space(w, 0, sz))
				return -sz;
			w->data[w->datalen] = (uint8_t) (sz | 0x80);		/* VARSIZE_TO_SHORT_D */
			memcpy(w->data + w->datalen + 1, payload, (size_t) len);
		}
		else
		{
			int64_t aligned = align_off(w->datalen, att->attalign);
			uint32_t l;

			sz = len + 4;
			/* att_align_zero happens BEFORE the space check (:1666-1671): the padding stays even if the item does not */
			memset(w->data + w->datalen, 0, (size_t) (aligned - w->datalen));
			w->datalen = aligned;
			if (!orig_has_space(w, 0, sz))
				return -sz;
			l = (uint32_t) sz & 0x3FFFFFFF;
			w->data[w->datalen] = (uint8_t) (l >> 24);			/* SET_VARSIZE_4B: network byte order (postgres.h:214) */
			w->data[w->datalen + 1] = (uint8_t) (l >> 16);
			w->data[w->datalen + 2] = (uint8_t) (l >> 8);
			w->data[w->datalen + 3] = (uint8_t) l;
			memcpy(w->data + w->datalen + 4, payload, (size_t) len);
		}
		w->datalen += sz;
		w->nullmap[w->nth] = 0;
		w->nullcount_bits++;
		w->nth++;
		return sz;
	}
	if (!orig_has_space(w, 0, att->attlen))
		return -att->attlen;
	/* DatumStreamBlockWrite_PutFixedLength (:1271-1312): by-value Datums are stored in attlen native-endian bytes */
	if (att->attbyval)
		memcpy(w->data + w->datalen, &value, (size_t) att->attlen);		/* little-endian host */
	else
		memcpy(w->data + w->datalen, (const void *) (uintptr_t) value, (size_t) att->attlen);
	w->datalen += att->attlen;
	w->nullmap[w->nth] = 0;
	w->nullcount_bits++;
	w->nth++;
	return att->attlen;
}

/* DatumStreamBlockWrite_BlockOrig (:3644-3710) + AppendOnlyStorageWrite_FinishBuffer (cdbappendonlystoragewrite.c:1327-1361)
 * + AppendOnlyStorageFormat_MakeSmallContentHeader / AddFirstRowNum / AddBlockHeaderChecksums
 * (cdbappendonlystorageformat.c:89-160,241-321).  Returns the storage block's length. */
static int64_t
finish_block(blockwriter *w, uint8_t *out, int checksum, int64_t firstrow)
{
	int hdrlen = ao_header_len(checksum);
	uint8_t *c = out + hdrlen, *p;
	int32_t nullsz = 0, unaligned = 0;
	int64_t contentlen, rounded, overall;
	uint32_t w0, w1;
	int i;

	if (w->has_null)
	{
		unaligned = (w->nth + 7) >> 3;
		nullsz = (int32_t) MAXALIGN8(unaligned);
	}
	c[0] = 0; c[1] = 0;											/* version = DatumStreamVersion_Original */
	c[2] = w->has_null ? DSB_HAS_NULLBITMAP : 0; c[3] = 0;			/* flags */
	c[4] = (uint8_t) (w->nth & 0xFF); c[5] = (uint8_t) (w->nth >> 8);	/* ndatum */
	c[6] = 0; c[7] = 0;												/* unused */
	wr32(c + 8, (uint32_t) nullsz);
	wr32(c + 12, (uint32_t) w->datalen);
	p = c + DSB_HDR;
	if (w->has_null)
	{
		memset(p, 0, (size_t) nullsz);
		for (i = 0; i < w->nth; i++)
			if (w->nullmap[i])
				p[i >> 3] |= (uint8_t) (1 << (i & 7));			/* DatumStreamBitMapWrite_AddBit: LSB first */
		p += nullsz;
	}
	memcpy(p, w->data, (size_t) w->datalen);
	p += w->datalen;
	contentlen = p - c;
	rounded = MAXALIGN8(contentlen);							/* AOStorage_RoundUp8 + AOStorage_ZeroPad */
	memset(c + contentlen, 0, (size_t) (rounded - contentlen));
	overall = hdrlen + rounded;

	w0 = ((uint32_t) AO_KIND_SMALLCONTENT << 28) | (1u << 27) | ((uint32_t) AOCSBK_BLOCK << 24) |
		(((uint32_t) w->nth << 10) & 0x00FFFC00u) | (((uint32_t) contentlen >> 11) & 0x3FFu);
	w1 = ((uint32_t) contentlen << 21) & 0xFFE00000u;			/* compressedLength = 0 */
	wr32(out, w0);
	wr32(out + 4, w1);
	memcpy(out + AO_HDR + (checksum ? 8 : 0), &firstrow, 8);
	if (checksum)
	{
		wr32(out + 8, or_aocs_crc32c(out + 16, overall - 16));	/* block checksum first: the header checksum covers it */
		wr32(out + 12, or_aocs_crc32c(out, 12));
	}
	/* DatumStreamBlockWrite_GetReady */
	w->nth = 0; w->nullcount_bits = 0; w->has_null = 0; w->datalen = 0;
	return overall;
}

/* aocs_insert_values (aocsam.c:964-1016) for one column: put; when the block is full write it and put again */
int64_t
or_aocs_write_column(const gg_attr *att, const int64_t *values, const int32_t *lens, const uint8_t *nulls, int64_t nrows,
					 int blocksize, int checksum, int64_t first_rownum, uint8_t *out, int64_t outcap)
{
	blockwriter w;
	int64_t pos = 0, r, blockfirst = first_rownum;
	int rc = 0;

	if (att->attlen == 0 || att->attlen < -1 || (att->attbyval && att->attlen != 1 && att->attlen != 2 && att->attlen != 4 && att->attlen != 8))
		return OR_ERR_UNSUPPORTED;
	memset(&w, 0, sizeof w);
	w.att = att;
	w.maxrows = AO_MAXROWS;
	w.maxdata = blocksize - ao_header_len(checksum);
	w.nullmap = calloc(AO_MAXROWS + 1, 1);
	w.data = malloc((size_t) blocksize + 16);
	for (r = 0; r < nrows && rc == 0; r++)
	{
		int null = nulls != NULL && nulls[r] != 0;
		int32_t len = lens ? lens[r] : 0;

		if (orig_put(&w, values[r], len, null) < 0)
		{
			if (w.nth > 0)
			{
				if (pos + blocksize > outcap) { rc = OR_ERR_NOMEM; break; }
				pos += finish_block(&w, out + pos, checksum, blockfirst);
				blockfirst = first_rownum + r;
			}
			if (orig_put(&w, values[r], len, null) < 0)
				rc = OR_ERR_UNSUPPORTED;			/* datumstreamwrite_lob: one value per AOCSBK_BLOB block, out of scope */
		}
	}
	if (rc == 0 && w.nth > 0)
	{
		if (pos + blocksize > outcap)
			rc = OR_ERR_NOMEM;
		else
			pos += finish_block(&w, out + pos, checksum, blockfirst);
	}
	free(w.nullmap);
	free(w.data);
	return rc ? rc : pos;
}

/* ------------------------------------------------ reader ------------------------------------------------ */

/* One storage block: AppendOnlyStorageFormat_GetHeaderInfo / GetSmallContentHeaderInfo / Verify*Checksum
 * (cdbappendonlystorageformat.c:1202-1417,1661-1721).  Returns 0 or OR_ERR_*. */
static int
parse_block(const uint8_t *b, int64_t avail, int checksum, int *rowcount, int64_t *firstrow,
			int32_t *offset, int32_t *datalen, int64_t *overall)
{
	uint32_t w0, w1;
	int hasfirst;

	if (avail < AO_HDR + (checksum ? 8 : 0))
		return OR_ERR_UNSUPPORTED;
	w0 = rd32(b);
	w1 = rd32(b + 4);
	if (w0 == 0 || (w0 >> 31) != 0)							/* AOHeaderCheckFirst32BitsAllZeroes / ReservedBit0Not0 */
		return OR_ERR_UNSUPPORTED;
	if (((w0 >> 28) & 7) != AO_KIND_SMALLCONTENT)			/* large / dense content: out of scope */
		return OR_ERR_UNSUPPORTED;
	if (checksum && rd32(b + 12) != or_aocs_crc32c(b, 12))
		return OR_ERR_UNSUPPORTED;
	hasfirst = (w0 >> 27) & 1;
	if (((w0 >> 24) & 7) != AOCSBK_BLOCK)
		return OR_ERR_UNSUPPORTED;
	*rowcount = (int) ((w0 >> 10) & 0x3FFF);
	*datalen = (int32_t) (((w0 & 0x3FF) << 11) | (w1 >> 21));
	if ((w1 & 0x001FFFFF) != 0)								/* compressedLength: bulk compression is out of scope */
		return OR_ERR_UNSUPPORTED;
	*offset = AO_HDR + (checksum ? 8 : 0);
	*firstrow = -1;
	if (hasfirst)
	{
		memcpy(firstrow, b + *offset, 8);
		*offset += 8;
	}
	*overall = *offset + MAXALIGN8(*datalen);
	if (*overall > avail)									/* AOHeaderCheckInvalidOverallBlockLen */
		return OR_ERR_UNSUPPORTED;
	if (checksum && rd32(b + 8) != or_aocs_crc32c(b + 16, *overall - 16))
		return OR_ERR_UNSUPPORTED;
	return 0;
}

/* values[]: by-value Datums (zero-extended from attlen bytes, as DatumStreamBlockRead_GetOrig does); for attlen -1 and
 * fixed-length by-reference types the byte OFFSET in `file` of the stored datum.  Returns the row count or OR_ERR_*. */
int64_t
or_aocs_read_column(const gg_attr *att, const uint8_t *file, int64_t nbytes, int checksum,
					int64_t *values, uint8_t *nulls, int64_t cap,
					int64_t *firstrows, int32_t *rowcounts, int blockcap, int *nblocks)
{
	int64_t pos = 0, n = 0;
	int nb = 0;

	while (pos < nbytes)
	{
		int rowcount, rc, has_null, ndatum, i;
		int32_t offset, datalen, nullsz, sz;
		int64_t firstrow, overall, d, dend;
		const uint8_t *c, *bitmap;

		if ((rc = parse_block(file + pos, nbytes - pos, checksum, &rowcount, &firstrow, &offset, &datalen, &overall)) != 0)
			return rc;
		if (nb < blockcap)
		{
			if (firstrows) firstrows[nb] = firstrow;
			if (rowcounts) rowcounts[nb] = rowcount;
		}
		nb++;
		/* DatumStreamBlockRead_GetReadyOrig (datumstreamblock.c:150-330) */
		c = file + pos + offset;
		if (datalen < DSB_HDR || c[0] != 0 || c[1] != 0)		/* version must be Original */
			return OR_ERR_UNSUPPORTED;
		has_null = (c[2] & DSB_HAS_NULLBITMAP) != 0;
		if ((c[2] & ~DSB_HAS_NULLBITMAP) != 0 || c[3] != 0)
			return OR_ERR_UNSUPPORTED;
		ndatum = c[4] | (c[5] << 8);
		nullsz = (int32_t) rd32(c + 8);
		sz = (int32_t) rd32(c + 12);
		if (ndatum != rowcount)								/* "logical_row_count == rowCount" (:268) */
			return OR_ERR_UNSUPPORTED;
		bitmap = c + DSB_HDR;
		d = MAXALIGN8(DSB_HDR + (has_null ? nullsz : 0));
		dend = d + sz;
		if (dend > datalen || (has_null && nullsz < ((ndatum + 7) >> 3)))
			return OR_ERR_UNSUPPORTED;
		/* DatumStreamBlockRead_AdvanceOrig / GetOrig (datumstreamblock.h:1216-1570) */
		for (i = 0; i < ndatum; i++)
		{
			if (n >= cap)
				return OR_ERR_NOMEM;
			if (has_null && (bitmap[i >> 3] >> (i & 7)) & 1)
			{
				nulls[n] = 1;
				values[n++] = 0;
				continue;
			}
			nulls[n] = 0;
			if (att->attlen == -1)
			{
				int len;

				/* zero bytes in front of an item are alignment padding of a 4-byte header (:1529-1536) */
				if (d < dend && c[d] == 0)
					d = align_off(d, att->attalign);
				if (d >= dend)
					return OR_ERR_UNSUPPORTED;
				values[n] = (int64_t) (c + d - file);
				(void) or_varlena_payload(c + d, &len);
				d += len + ((c[d] & 0x80) ? 1 : 4);
			}
			else
			{
				if (d + att->attlen > dend)
					return OR_ERR_UNSUPPORTED;
				if (att->attbyval)
				{
					uint64_t v = 0;

					memcpy(&v, c + d, (size_t) att->attlen);	/* *(uint8|uint16|uint32|Datum *) datump */
					values[n] = (int64_t) v;
				}
				else
					values[n] = (int64_t) (c + d - file);
				d += att->attlen;
			}
			n++;
		}
		pos += overall;
	}
	if (nblocks)
		*nblocks = nb;
	return n;
}

/* ------------------------------ SeqScan over the column files -> qual -> Agg ------------------------------ */

/* aocs_getnext (aocsam.c:700-800) reads the same row of every projected column and fills the slot's Datum arrays;
 * ExecScan applies the qual; the aggregate is or_agg.c's.  colfiles[i] may be NULL for a column the plan never
 * references (AOCS opens only the projected columns, aocsam.c:160-200): it then reads as NULL. */
int
or_aocs_seqscan_agg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool,
					const uint8_t *const *colfiles, const int64_t *colbytes, int checksum, int64_t nrows_hint,
					gg_aggrow *out, int outcap, int *nout, uint64_t *rows_scanned, uint64_t *rows_passed)
{
	int natts = scan->desc.natts, a, rc = 0;
	int64_t **vals = calloc((size_t) natts, sizeof *vals);
	uint8_t **nul = calloc((size_t) natts, sizeof *nul);
	int64_t nrows = -1, r;
	uint64_t npass = 0;
	or_aggtable *t = NULL;
	or_row row;
	int32_t signext[GG_MAX_ATTS];

	for (a = 0; a < natts && rc == 0; a++)
	{
		int64_t n;

		signext[a] = 0;
		if (colfiles[a] == NULL)
			continue;
		vals[a] = malloc(sizeof(int64_t) * (size_t) (nrows_hint + 1));
		nul[a] = malloc((size_t) nrows_hint + 1);
		n = or_aocs_read_column(&scan->desc.attrs[a], colfiles[a], colbytes[a], checksum, vals[a], nul[a], nrows_hint, NULL, NULL, 0, NULL);
		if (n < 0)
			rc = (int) n;
		else if (nrows >= 0 && n != nrows)
			rc = OR_ERR_UNSUPPORTED;							/* the columns of one segment file hold the same rows */
		else
			nrows = n;
		/* by-value Datums of signed types arrive zero-extended from the block; the expression evaluator takes
		 * int4 / date Datums sign-extended like DatumGetInt32 (postgres.h) does */
		if (scan->desc.attrs[a].attbyval && scan->desc.attrs[a].attlen == 4)
			signext[a] = 1;
	}
	if (rc == 0)
	{
		t = or_aggtable_create(agg, pool, 0);
		memset(&row, 0, sizeof row);
		row.desc = &scan->desc;
		row.nvalid = natts;
		for (r = 0; r < (nrows < 0 ? 0 : nrows) && rc == 0; r++)
		{
			const uint8_t *base = NULL;

			for (a = 0; a < natts; a++)
			{
				if (vals[a] == NULL) { row.isnull[a] = 1; row.values[a] = 0; continue; }
				row.isnull[a] = nul[a][r];
				row.values[a] = signext[a] ? (int64_t) (int32_t) vals[a][r] : vals[a][r];
				if (scan->desc.attrs[a].attlen == -1 && !nul[a][r])
				{
					/* row_getattr adds the datum's offset to row.tuple: make offsets relative to one base */
					if (base == NULL)
						base = colfiles[a];
					row.values[a] = (int64_t) ((colfiles[a] + vals[a][r]) - base);
				}
			}
			row.tuple = base;
			if (scan->qual >= 0)
			{
				or_datum q;

				if ((rc = or_eval(pool, scan->qual, &row, NULL, &q)) != 0)
					break;
				if (q.isnull || !q.v)
					continue;
			}
			npass++;
			rc = or_aggtable_advance(t, &row, NULL);
		}
		if (rc == 0)
			rc = or_aggtable_emit(t, out, outcap, nout);
		or_aggtable_free(t);
	}
	for (a = 0; a < natts; a++) { free(vals[a]); free(nul[a]); }
	free(vals);
	free(nul);
	if (rows_scanned) *rows_scanned = nrows < 0 ? 0 : (uint64_t) nrows;
	if (rows_passed) *rows_passed = npass;
	return rc;
}
