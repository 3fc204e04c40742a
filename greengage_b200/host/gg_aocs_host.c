/*
 * gg_aocs_host.c — loader-side handling of append-only column-oriented (AOCS) column files (include/gg_aocs.h).
 * Host C, part of libgghost.so.  Format citations are in the header; this file is the product's own reader of block
 * headers and its own writer, not shared with oracle/ (which is test infrastructure and checks this code).
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/gg_aocs.h"

#if defined(__x86_64__)
#include <nmmintrin.h>
#define GG_HAVE_SSE42_PATH 1
#endif

/* ------------------------------------------------ CRC-32C ------------------------------------------------ */

static uint32_t crc_tab[8][256];
static int crc_ready;

static void crc_build(void)
{
	uint32_t i, k;
	int t;

	for (i = 0; i < 256; i++)
	{
		uint32_t c = i;
		for (k = 0; k < 8; k++)
			c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
		crc_tab[0][i] = c;
	}
	for (t = 1; t < 8; t++)
		for (i = 0; i < 256; i++)
			crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
	__atomic_store_n(&crc_ready, 1, __ATOMIC_RELEASE);
}

static uint32_t crc_soft(uint32_t c, const uint8_t *p, int64_t n)
{
	if (!__atomic_load_n(&crc_ready, __ATOMIC_ACQUIRE))
		crc_build();
	while (n >= 8)							/* slicing by 8 */
	{
		uint32_t lo, hi;
		memcpy(&lo, p, 4);
		memcpy(&hi, p + 4, 4);
		lo ^= c;
		c = crc_tab[7][lo & 0xFF] ^ crc_tab[6][(lo >> 8) & 0xFF] ^ crc_tab[5][(lo >> 16) & 0xFF] ^ crc_tab[4][lo >> 24] ^
			crc_tab[3][hi & 0xFF] ^ crc_tab[2][(hi >> 8) & 0xFF] ^ crc_tab[1][(hi >> 16) & 0xFF] ^ crc_tab[0][hi >> 24];
		p += 8;
		n -= 8;
	}
	while (n-- > 0)
		c = crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
	return c;
}

#ifdef GG_HAVE_SSE42_PATH
__attribute__((target("sse4.2")))
static uint32_t crc_hw(uint32_t c, const uint8_t *p, int64_t n)
{
	uint64_t c64 = c;

	while (n >= 8)
	{
		uint64_t v;
		memcpy(&v, p, 8);
		c64 = _mm_crc32_u64(c64, v);
		p += 8;
		n -= 8;
	}
	c = (uint32_t) c64;
	while (n-- > 0)
		c = _mm_crc32_u8(c, *p++);
	return c;
}
#endif

uint32_t gg_aocs_crc32c(const uint8_t *p, int64_t n)
{
#ifdef GG_HAVE_SSE42_PATH
	if (__builtin_cpu_supports("sse4.2"))
		return crc_hw(0xFFFFFFFFu, p, n);
#endif
	return crc_soft(0xFFFFFFFFu, p, n);
}

/* ------------------------------------------------ layout helpers ------------------------------------------------ */

enum { STORAGE_HDR = 8, CHECKSUMS = 8, FIRSTROW = 8, STREAM_HDR = 16 };
enum { KIND_SMALL = 1, EXEC_BLOCK = 1, FLAG_NULLMAP = 1, ROWS_LIMIT = 0x3FFF };

static inline int64_t up8(int64_t x) { return (x + 7) & ~(int64_t) 7; }
static inline uint32_t get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

static inline int64_t type_align(int64_t off, int align)
{
	int a = align == 'd' ? 8 : align == 'i' ? 4 : align == 's' ? 2 : 1;
	return (off + a - 1) & ~(int64_t) (a - 1);
}

static int attr_ok(const gg_attr *att)
{
	if (att->attlen == -1)
		return 1;
	if (att->attlen <= 0)
		return 0;
	if (att->attbyval)
		return att->attlen == 1 || att->attlen == 2 || att->attlen == 4 || att->attlen == 8;
	return 1;
}

/* ------------------------------------------------ index (loader) ------------------------------------------------ */

static int32_t count_null_bits(const uint8_t *bitmap, int32_t nbits);

/* Stored size shared by every value of a varlena block, or 0.  Only 1-byte-header values qualify (first byte 0x80 | size,
 * postgres.h VARATT_IS_1B in this tree): a zero byte is alignment padding in front of a 4-byte header. */
static int32_t uniform_varlena_size(const uint8_t *d, int64_t len)
{
	int64_t at = 0;
	int32_t size = 0;

	if (len == 0)
		return 1;							/* a block of NULLs only: nothing is ever addressed */
	while (at < len)
	{
		int32_t sz;

		if (!(d[at] & 0x80))
			return 0;
		sz = d[at] & 0x7F;
		if (sz < 1 || (size != 0 && sz != size))
			return 0;
		size = sz;
		at += sz;
	}
	return at == len ? size : 0;
}

int gg_aocs_index_column(const gg_attr *att, const uint8_t *file, int64_t nbytes, int checksum,
                         gg_aocs_block *dir, int64_t cap, int64_t *nblocks, int64_t *nrows)
{
	int64_t at = 0, nb = 0, rows = 0, expect_row = -1;
	const int fixed = STORAGE_HDR + (checksum ? CHECKSUMS : 0);

	if (att == NULL || file == NULL || nbytes < 0 || !attr_ok(att))
		return GG_ERR_ARG;
	while (at < nbytes)
	{
		const uint8_t *b = file + at, *s;
		uint32_t w0, w1, kind, contentlen, flags, nullbytes, datalen;
		int64_t hdr, blocklen, firstrow = -1;
		int nrow, ndatum, nvalues;

		if (nbytes - at < fixed)
			return GG_ERR_BADPAGE;
		w0 = get32(b);
		w1 = get32(b + 4);
		if (w0 == 0 || (w0 & 0x80000000u))			/* all-zero word / reserved bit: not a header */
			return GG_ERR_BADPAGE;
		kind = (w0 >> 28) & 7;
		if (kind == 0 || kind > 4)
			return GG_ERR_BADPAGE;
		if (kind == 4)
			return GG_ERR_UNSUPPORTED;					/* bulk dense content: 16-byte header, bulk-compressed RLE blocks */
		if (checksum && get32(b + 12) != gg_aocs_crc32c(b, 12))
			return GG_ERR_BADPAGE;						/* header checksum: only now are the fields trustworthy */
		if (kind != KIND_SMALL)
			return GG_ERR_UNSUPPORTED;					/* large content (a value spanning blocks), non-bulk dense content */
		if (((w0 >> 24) & 7) != EXEC_BLOCK)
			return GG_ERR_UNSUPPORTED;					/* AOCSBK_BLOB */
		if (w1 & 0x001FFFFFu)
			return GG_ERR_UNSUPPORTED;					/* compressedLength != 0: bulk-compressed block */
		nrow = (int) ((w0 >> 10) & 0x3FFF);
		contentlen = ((w0 & 0x3FF) << 11) | (w1 >> 21);
		hdr = fixed;
		if (w0 & 0x08000000u)
		{
			if (nbytes - at < hdr + FIRSTROW)
				return GG_ERR_BADPAGE;
			memcpy(&firstrow, b + hdr, 8);
			hdr += FIRSTROW;
		}
		blocklen = hdr + up8(contentlen);
		if (blocklen > nbytes - at)
			return GG_ERR_BADPAGE;
		if (checksum && get32(b + 8) != gg_aocs_crc32c(b + 16, blocklen - 16))
			return GG_ERR_BADPAGE;

		/* the datum-stream block inside */
		s = b + hdr;
		if (contentlen < STREAM_HDR)
			return GG_ERR_BADPAGE;
		if (s[0] != 0 || s[1] != 0)
			return GG_ERR_UNSUPPORTED;					/* version 1 / 2: Dense (RLE_TYPE, delta) */
		flags = (uint32_t) s[2] | ((uint32_t) s[3] << 8);
		if (flags & ~(uint32_t) FLAG_NULLMAP)
			return GG_ERR_UNSUPPORTED;
		ndatum = s[4] | (s[5] << 8);
		nullbytes = get32(s + 8);
		datalen = get32(s + 12);
		if (ndatum != nrow)
			return GG_ERR_BADPAGE;
		if (flags & FLAG_NULLMAP)
		{
			if ((int64_t) nullbytes < ((int64_t) nrow + 7) / 8 || (nullbytes & 7))
				return GG_ERR_BADPAGE;
		}
		else
			nullbytes = 0;
		if (up8(STREAM_HDR + (int64_t) nullbytes) + (int64_t) datalen > (int64_t) contentlen)
			return GG_ERR_BADPAGE;
		if (att->attlen > 0 && datalen % (uint32_t) att->attlen != 0)
			return GG_ERR_BADPAGE;
		/* the value area holds exactly the non-NULL rows: the device addresses value i at data_off + i * stride without
		 * looking at data_len again, so a block that claims more values than it stores must not get into the directory */
		nvalues = nrow - ((flags & FLAG_NULLMAP) ? count_null_bits(s + STREAM_HDR, nrow) : 0);
		if (att->attlen > 0 && (int64_t) datalen != (int64_t) nvalues * att->attlen)
			return GG_ERR_BADPAGE;
		/* row numbers of consecutive blocks of one segment file are contiguous unless rows were appended by separate
		 * inserts with gaps in the fast sequence (allowed): they must at least never go backwards */
		if (firstrow >= 0 && expect_row >= 0 && firstrow < expect_row)
			return GG_ERR_BADPAGE;
		if (firstrow >= 0)
			expect_row = firstrow + nrow;

		if (dir != NULL)
		{
			if (nb >= cap)
				return GG_ERR_NOMEM;
			dir[nb].first_row = firstrow;
			dir[nb].null_off = (flags & FLAG_NULLMAP) ? at + hdr + STREAM_HDR : -1;
			dir[nb].data_off = at + hdr + up8(STREAM_HDR + (int64_t) nullbytes);
			dir[nb].nrows = nrow;
			dir[nb].data_len = (int32_t) datalen;
			dir[nb].stride = att->attlen > 0 ? att->attlen
				: uniform_varlena_size(b + hdr + up8(STREAM_HDR + (int64_t) nullbytes), (int64_t) datalen);
			if (att->attlen < 0 && dir[nb].stride > 0 && nvalues > 0 && (int64_t) dir[nb].stride * nvalues != (int64_t) datalen)
				dir[nb].stride = 0;						/* fewer values than rows claim: irregular, the device refuses the block */
			dir[nb].pad = 0;
		}
		nb++;
		rows += nrow;
		at += blocklen;
	}
	if (nblocks) *nblocks = nb;
	if (nrows) *nrows = rows;
	return GG_OK;
}

/* ------------------------------------------------ tile plan ------------------------------------------------ */

static int32_t count_null_bits(const uint8_t *bitmap, int32_t nbits)
{
	int32_t n = 0, i = 0;

	for (; i + 64 <= nbits; i += 64)
	{
		uint64_t w;
		memcpy(&w, bitmap + (i >> 3), 8);
		n += __builtin_popcountll(w);
	}
	for (; i < nbits; i++)
		n += (bitmap[i >> 3] >> (i & 7)) & 1;
	return n;
}

int gg_aocs_plan_tiles(const gg_aocs_block *dir, int64_t nblocks, const uint8_t *file, int32_t tile_rows,
                       gg_aocs_tile *tiles, int64_t ntiles)
{
	int64_t b = 0, pos = 0, t;		/* pos: file position (row ordinal) of block b's first row */
	int64_t total = 0, i;

	if (dir == NULL || tiles == NULL || tile_rows <= 0 || nblocks < 0 || ntiles < 0)
		return GG_ERR_ARG;
	for (i = 0; i < nblocks; i++)
		total += dir[i].nrows;
	if (ntiles != (total + tile_rows - 1) / tile_rows)
		return GG_ERR_ARG;
	for (t = 0; t < ntiles; t++)
	{
		int64_t r = t * (int64_t) tile_rows;

		while (b < nblocks && r >= pos + dir[b].nrows)
			pos += dir[b++].nrows;
		if (b >= nblocks)
			return GG_ERR_BADPAGE;
		tiles[t].block = (int32_t) b;
		tiles[t].row_in_block = (int32_t) (r - pos);
		tiles[t].nulls_before = 0;
		tiles[t].pad = 0;
		if (dir[b].null_off >= 0 && r > pos)
		{
			if (file == NULL)
				return GG_ERR_ARG;
			tiles[t].nulls_before = count_null_bits(file + dir[b].null_off, (int32_t) (r - pos));
		}
	}
	return GG_OK;
}

/* ------------------------------------------------ writer ------------------------------------------------ */

struct gg_aocs_writer {
	gg_attr att;
	int checksum;
	int hdrlen;					/* storage header + checksums + first row number */
	int room;					/* bytes a block's content must stay below: blocksize - hdrlen */
	int blocksize;
	uint8_t *out;
	int64_t outcap, outlen;
	int64_t next_row;			/* row number the next put gets */
	int64_t block_first;		/* row number of the open block's first row */
	int rows;					/* rows in the open block */
	int any_null;				/* the open block has a NULL bitmap */
	uint8_t *bits;				/* packed NULL bitmap of the open block, always maintained (cheap) */
	uint8_t *vals;				/* stored values of the open block */
	int64_t vlen;
};

int64_t gg_aocs_file_bound(const gg_attr *att, int64_t nrows, int32_t maxlen, int blocksize, int checksum)
{
	int64_t per = att->attlen > 0 ? att->attlen : (int64_t) maxlen + 4 + 8;		/* header + worst alignment padding */
	int64_t payload = nrows * per + nrows / 8 + 8;
	int64_t overhead = STORAGE_HDR + (checksum ? CHECKSUMS : 0) + FIRSTROW + STREAM_HDR + 16;
	int64_t blocks = payload / (blocksize / 2) + nrows / GG_AOCS_MAX_BLOCK_ROWS + 2;

	return payload + blocks * overhead + blocksize;
}

int gg_aocs_writer_create(const gg_attr *att, int blocksize, int checksum, int64_t first_rownum,
                          uint8_t *out, int64_t outcap, gg_aocs_writer **wp)
{
	gg_aocs_writer *w;

	if (att == NULL || out == NULL || wp == NULL || !attr_ok(att) || blocksize < 8192 || blocksize > 2 * 1024 * 1024 || (blocksize & 7))
		return GG_ERR_ARG;
	w = calloc(1, sizeof *w);
	if (w == NULL)
		return GG_ERR_NOMEM;
	w->att = *att;
	w->checksum = checksum != 0;
	w->hdrlen = STORAGE_HDR + (w->checksum ? CHECKSUMS : 0) + FIRSTROW;
	w->room = blocksize - w->hdrlen;
	w->blocksize = blocksize;
	w->out = out;
	w->outcap = outcap;
	w->next_row = w->block_first = first_rownum;
	w->bits = calloc(ROWS_LIMIT / 8 + 16, 1);
	w->vals = malloc((size_t) blocksize + 16);
	if (w->bits == NULL || w->vals == NULL)
	{
		free(w->bits); free(w->vals); free(w);
		return GG_ERR_NOMEM;
	}
	*wp = w;
	return GG_OK;
}

/* would the open block still be a legal block with one more row of `add` value bytes? */
static int fits(const gg_aocs_writer *w, int isnull, int64_t add)
{
	int64_t bitmap = (isnull || w->any_null) ? up8(((int64_t) w->rows + 1 + 7) / 8) : 0;

	if (w->rows + 1 >= ROWS_LIMIT)
		return 0;
	return STREAM_HDR + bitmap + w->vlen + add < w->room;
}

static int flush_block(gg_aocs_writer *w)
{
	uint8_t *b = w->out + w->outlen, *s;
	int64_t bitmap = w->any_null ? up8(((int64_t) w->rows + 7) / 8) : 0;
	int64_t content = STREAM_HDR + bitmap + w->vlen, padded = up8(content), total = w->hdrlen + padded;
	int sums = w->checksum ? CHECKSUMS : 0;

	if (w->rows == 0)
		return GG_OK;
	if (w->outlen + total > w->outcap)
		return GG_ERR_NOMEM;
	put32(b, ((uint32_t) KIND_SMALL << 28) | 0x08000000u | ((uint32_t) EXEC_BLOCK << 24) |
			 ((uint32_t) w->rows << 10) | (uint32_t) (content >> 11));
	put32(b + 4, (uint32_t) (content & 0x7FF) << 21);
	memcpy(b + STORAGE_HDR + sums, &w->block_first, 8);
	s = b + w->hdrlen;
	memset(s, 0, STREAM_HDR);
	s[2] = w->any_null ? FLAG_NULLMAP : 0;
	s[4] = (uint8_t) (w->rows & 0xFF);
	s[5] = (uint8_t) (w->rows >> 8);
	put32(s + 8, (uint32_t) bitmap);
	put32(s + 12, (uint32_t) w->vlen);
	if (w->any_null)
	{
		int64_t used = ((int64_t) w->rows + 7) / 8;

		memcpy(s + STREAM_HDR, w->bits, (size_t) used);
		memset(s + STREAM_HDR + used, 0, (size_t) (bitmap - used));
	}
	memcpy(s + STREAM_HDR + bitmap, w->vals, (size_t) w->vlen);
	memset(s + content, 0, (size_t) (padded - content));
	if (w->checksum)
	{
		put32(b + 8, gg_aocs_crc32c(b + 16, total - 16));
		put32(b + 12, gg_aocs_crc32c(b, 12));
	}
	w->outlen += total;
	memset(w->bits, 0, (size_t) (((int64_t) w->rows + 7) / 8));
	w->rows = 0;
	w->any_null = 0;
	w->vlen = 0;
	w->block_first = w->next_row;
	return GG_OK;
}

/* append one value to the open block; 0 = stored, 1 = the block is full */
static int try_put(gg_aocs_writer *w, int64_t value, int32_t len, int isnull)
{
	if (isnull)
	{
		if (!fits(w, 1, 0))
			return 1;
		w->any_null = 1;
		w->bits[w->rows >> 3] |= (uint8_t) (1u << (w->rows & 7));
		w->rows++;
		return 0;
	}
	if (w->att.attlen > 0)
	{
		if (!fits(w, 0, w->att.attlen))
			return 1;
		if (w->att.attbyval)
			memcpy(w->vals + w->vlen, &value, (size_t) w->att.attlen);
		else
			memcpy(w->vals + w->vlen, (const void *) (uintptr_t) value, (size_t) w->att.attlen);
		w->vlen += w->att.attlen;
	}
	else if (len <= 126)
	{
		/* short varlena: one header byte 0x80 | total length (postgres.h SET_VARSIZE_1B in this tree) */
		if (!fits(w, 0, (int64_t) len + 1))
			return 1;
		w->vals[w->vlen] = (uint8_t) (0x80 | (len + 1));
		memcpy(w->vals + w->vlen + 1, (const void *) (uintptr_t) value, (size_t) len);
		w->vlen += len + 1;
	}
	else
	{
		/* 4-byte header (big-endian total length), aligned to typalign with zero bytes in front.  The reference pads
		 * BEFORE it checks for room (datumstreamblock.c:1666-1674), so the padding stays in a block that then turns
		 * out to be full — kept, or the files would differ. */
		int64_t start = type_align(w->vlen, w->att.attalign);
		uint32_t total = (uint32_t) len + 4;

		memset(w->vals + w->vlen, 0, (size_t) (start - w->vlen));
		w->vlen = start;
		if (!fits(w, 0, (int64_t) total))
			return 1;
		w->vals[start] = (uint8_t) (total >> 24);
		w->vals[start + 1] = (uint8_t) (total >> 16);
		w->vals[start + 2] = (uint8_t) (total >> 8);
		w->vals[start + 3] = (uint8_t) total;
		memcpy(w->vals + start + 4, (const void *) (uintptr_t) value, (size_t) len);
		w->vlen += total;
	}
	w->rows++;
	return 0;
}

int gg_aocs_writer_put(gg_aocs_writer *w, int64_t value, int32_t len, int isnull)
{
	int rc;

	if (w == NULL || (!isnull && w->att.attlen == -1 && (len < 0 || len > 0x3FFFFFFF - 4)))
		return GG_ERR_ARG;
	if (try_put(w, value, len, isnull) != 0)
	{
		if ((rc = flush_block(w)) != GG_OK)
			return rc;
		if (try_put(w, value, len, isnull) != 0)
			return GG_ERR_UNSUPPORTED;		/* a value larger than a block: the reference stores it as large content */
	}
	w->next_row++;
	return GG_OK;
}

int gg_aocs_writer_finish(gg_aocs_writer *w, int64_t *nbytes)
{
	int rc;

	if (w == NULL)
		return GG_ERR_ARG;
	rc = flush_block(w);
	if (nbytes)
		*nbytes = w->outlen;
	free(w->bits);
	free(w->vals);
	free(w);
	return rc;
}
