/*
 * gg_tupser.c — MemTuple and tuple-chunk wire format (include/gg_tupser.h), host C.
 *
 * What the reference does in memtuple.c (binding, form, deform) and tupser.c (SerializeTuple, CvtChunksToTup), written for
 * rows held as Datum arrays.  The layout rules followed, each with its source:
 *   - physical attribute order: 8-byte aligned fixed-width first, then 4-byte aligned (and, in the large layout, varlena
 *     offset words of 4 bytes), then 2-byte aligned (small layout: varlena offset words of 2 bytes), then 1-byte aligned
 *     (memtuple.c:238-381); the NULL bitmap is indexed by physical position
 *   - fixed area starts at 8 when any attribute is 8-byte aligned, else at 4; a NULL bitmap of more than 4 bytes (or any, when
 *     the columns are 4-byte aligned) pushes everything by null_bitmap_extra (memtuple.c:60-72,692-697)
 *   - a NULL attribute takes no space: everything physically behind it moves up by its len_aligned — its length padded to the
 *     alignment of the attribute that physically follows (memtuple.c:124-141,546-549)
 *   - varlena bodies follow the fixed area: a value that fits a 1-byte header (payload <= 126 bytes) is stored with one,
 *     unaligned; longer ones with the 4-byte big-endian header, aligned to the attribute's alignment (memtuple.c:762-797;
 *     postgres.h:158-230 — GPDB's varlena headers are big-endian)
 *   - total length padded to 8, stored in the first word with bit 31 set; bit 0 = has NULLs, bit 1 = large layout
 *     (memtup.h:66-80)
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/gg_tupser.h"
#include "../../include/ggb200.h"

#define MT_LEAD_BIT 0x80000000u
#define MT_LEN_MASK 0x3FFFFFF8u
#define MT_HASNULL 1u
#define MT_LARGE 2u
#define MT_HASEXT 4u
#define MT_FITSHORT 0xFFF0u
#define SHORT_MAX_PAYLOAD 126          /* VARATT_SHORT_MAX 0x7F includes the 1-byte header */

static uint32_t align_to(uint32_t off, int a)
{
	uint32_t m = a == 'd' ? 7u : a == 'i' ? 3u : a == 's' ? 1u : 0u;
	return (off + m) & ~m;
}

static int is_varlena(const gg_attr *a) { return a->attlen == -1; }

/* one layout (small: 2-byte varlena offsets; large: 4-byte), memtuple.c:175-417 */
static void make_layout(const gg_attr *attrs, int natts, int column_align, int large, gg_mt_layout *L)
{
	uint32_t cur = column_align == 8 ? 8 : 4;
	int phys = 0, pass, i, prev = -1;
	static const char pass_align[4] = { 'd', 'i', 's', 'c' };
	memset(L, 0, sizeof *L);
	for (pass = 0; pass < 4; pass++)
		for (i = 0; i < natts; i++)
		{
			const gg_attr *a = &attrs[i];
			int take = 0, len = 0;
			if (pass == 0) take = a->attlen > 0 && a->attalign == 'd';
			else if (pass == 1) take = (a->attlen > 0 && a->attalign == 'i') || (large && is_varlena(a));
			else if (pass == 2) take = (a->attlen > 0 && a->attalign == 's') || (!large && is_varlena(a));
			else take = a->attlen > 0 && a->attalign == 'c';
			if (!take) continue;
			len = a->attlen > 0 ? a->attlen : (pass == 1 ? 4 : 2);
			L->att[i].offset = (int32_t) align_to(cur, pass_align[pass]);
			L->att[i].len = (int16_t) len;
			L->att[i].flag = (uint8_t) (is_varlena(a) ? 3 : (a->attbyval ? 1 : 2));
			L->att[i].null_byte = (uint8_t) (phys >> 3);
			L->att[i].null_mask = (uint8_t) (1u << (phys & 7));
			L->att[i].phys = (uint8_t) phys;
			/* the previous attribute's length as far as the NULL accounting goes: padded for this one's alignment */
			if (prev >= 0) L->att[prev].len_aligned = (int16_t) align_to((uint32_t) L->att[prev].len, pass_align[pass]);
			prev = i;
			phys++;
			cur = (uint32_t) L->att[i].offset + (uint32_t) len;
		}
	if (prev >= 0) L->att[prev].len_aligned = L->att[prev].len;
	L->var_start = natts ? (int32_t) cur : 8;
}

int gg_memtuple_bind(const gg_attr *attrs, int natts, gg_memtuple_binding *out)
{
	int i, nbytes, avail;
	if (!attrs || !out || natts < 0) return GG_ERR_ARG;
	if (natts > GG_MT_MAX_ATTS) return GG_ERR_UNSUPPORTED;
	memset(out, 0, sizeof *out);
	out->natts = natts;
	out->column_align = 4;
	for (i = 0; i < natts; i++)
	{
		if (attrs[i].attlen == -2 || attrs[i].attlen == 0 || attrs[i].attlen < -2) return GG_ERR_UNSUPPORTED;
		if (attrs[i].attlen > 0 && attrs[i].attalign == 'd') out->column_align = 8;
		out->attrs[i] = attrs[i];
	}
	/* compute_null_bitmap_extra_size, memtuple.c:60 */
	nbytes = (natts + 7) >> 3;
	avail = out->column_align == 4 ? 0 : 4;
	out->null_bitmap_extra = nbytes <= avail ? 0 : (int32_t) ((nbytes - avail + out->column_align - 1) / out->column_align * out->column_align);
	make_layout(out->attrs, natts, out->column_align, 0, &out->small);
	make_layout(out->attrs, natts, out->column_align, 1, &out->large);
	return GG_OK;
}

static uint32_t varlena_len(const int32_t *lens, int i) { return lens ? (uint32_t) lens[i] : 0; }

/* compute_memtuple_size_using_bind, memtuple.c:441 */
static uint32_t size_with(const gg_memtuple_binding *b, const gg_mt_layout *L, const uint8_t *isnull, const int32_t *lens, int hasnull,
                          uint32_t *nullsaves)
{
	uint32_t n = (uint32_t) L->var_start;
	int i;
	*nullsaves = 0;
	if (hasnull)
	{
		n += (uint32_t) b->null_bitmap_extra;
		for (i = 0; i < b->natts; i++)
			if (isnull[i]) { *nullsaves += (uint32_t) L->att[i].len_aligned; n -= (uint32_t) L->att[i].len_aligned; }
	}
	for (i = 0; i < b->natts; i++)
	{
		uint32_t pl;
		if ((isnull && isnull[i]) || L->att[i].flag != 3) continue;
		pl = varlena_len(lens, i);
		if (pl <= SHORT_MAX_PAYLOAD) n += pl + 1;
		else { n = align_to(n, b->attrs[i].attalign); n += pl + 4; }
	}
	return (n + 7) & ~7u;
}

/* bytes saved by the NULL attributes that physically precede physical position `phys` */
static uint32_t null_save_before(const gg_memtuple_binding *b, const gg_mt_layout *L, const uint8_t *bitmap, int phys)
{
	uint32_t s = 0;
	int i;
	if (!bitmap) return 0;
	for (i = 0; i < b->natts; i++)
		if (L->att[i].phys < phys && (bitmap[L->att[i].null_byte] & L->att[i].null_mask)) s += (uint32_t) L->att[i].len_aligned;
	return s;
}

static void store_le(uint8_t *p, uint64_t v, int n) { int k; for (k = 0; k < n; k++) p[k] = (uint8_t) (v >> (8 * k)); }
static uint64_t load_le(const uint8_t *p, int n) { uint64_t v = 0; int k; for (k = 0; k < n; k++) v |= (uint64_t) p[k] << (8 * k); return v; }

int gg_memtuple_form(const gg_memtuple_binding *b, const int64_t *values, const uint8_t *isnull, const int32_t *lens,
                     const void *const *ptrs, uint8_t *out, uint32_t cap, uint32_t *len)
{
	const gg_mt_layout *L;
	uint32_t n, nullsaves = 0, word, start = 0, vs;
	uint8_t *bitmap = NULL;
	int hasnull = 0, i;
	if (!b || !values || !len) return GG_ERR_ARG;
	for (i = 0; i < b->natts; i++) if (isnull && isnull[i]) hasnull = 1;
	n = size_with(b, &b->small, isnull, lens, hasnull, &nullsaves);
	L = &b->small;
	if (n > MT_FITSHORT) { n = size_with(b, &b->large, isnull, lens, hasnull, &nullsaves); L = &b->large; }
	*len = n;
	if (!out || cap < n) return GG_ERR_NOMEM;
	if (n > MT_LEN_MASK) return GG_ERR_UNSUPPORTED;             /* longer than the length field holds */
	memset(out, 0, n);
	word = n | MT_LEAD_BIT;
	if (L == &b->large) word |= MT_LARGE;
	if (hasnull) word |= MT_HASNULL;
	store_le(out, word, 4);
	vs = (uint32_t) L->var_start - nullsaves;
	if (hasnull)
	{
		bitmap = out + 4;
		start = (uint32_t) b->null_bitmap_extra;
		vs += (uint32_t) b->null_bitmap_extra;
		for (i = 0; i < b->natts; i++) if (isnull[i]) bitmap[L->att[i].null_byte] |= L->att[i].null_mask;
	}
	for (i = 0; i < b->natts; i++)
	{
		const gg_mt_attbind *ab = &L->att[i];
		uint8_t *p;
		if (isnull && isnull[i]) continue;
		p = out + start + (uint32_t) ab->offset - null_save_before(b, L, bitmap, ab->phys);
		if (ab->flag == 1) store_le(p, (uint64_t) values[i], ab->len);
		else if (ab->flag == 2)
		{
			if (!ptrs || !ptrs[i]) return GG_ERR_ARG;
			memcpy(p, ptrs[i], (size_t) ab->len);
		}
		else
		{
			const uint32_t pl = varlena_len(lens, i);
			uint8_t packed[8];
			const uint8_t *src = ptrs && ptrs[i] ? (const uint8_t *) ptrs[i] : packed;
			if (!(ptrs && ptrs[i]))
			{
				if (pl > 8) return GG_ERR_ARG;
				store_le(packed, (uint64_t) values[i], 8);
			}
			if (pl <= SHORT_MAX_PAYLOAD)
			{
				out[vs] = (uint8_t) (0x80u | (pl + 1));            /* 1-byte header: total length incl. itself (postgres.h:206) */
				memcpy(out + vs + 1, src, pl);
				store_le(p, vs - start, ab->len);
				vs += pl + 1;
			}
			else
			{
				const uint32_t tot = pl + 4;
				vs = align_to(vs, b->attrs[i].attalign);
				out[vs] = (uint8_t) ((tot >> 24) & 0x3F); out[vs + 1] = (uint8_t) (tot >> 16); out[vs + 2] = (uint8_t) (tot >> 8); out[vs + 3] = (uint8_t) tot;
				memcpy(out + vs + 4, src, pl);
				store_le(p, vs - start, ab->len);
				vs += tot;
			}
		}
	}
	return GG_OK;
}

uint32_t gg_memtuple_size(const uint8_t *mt) { return mt ? (uint32_t) load_le(mt, 4) & MT_LEN_MASK : 0; }

int gg_memtuple_deform(const gg_memtuple_binding *b, const uint8_t *mt, uint32_t len, int64_t *values, uint8_t *isnull, int32_t *lens)
{
	uint32_t word, n, start = 0;
	const gg_mt_layout *L;
	const uint8_t *bitmap = NULL;
	int i;
	if (!b || !mt || !values || !isnull || len < 8) return GG_ERR_ARG;
	word = (uint32_t) load_le(mt, 4);
	n = word & MT_LEN_MASK;
	if (!(word & MT_LEAD_BIT) || n > len || n < 8) return GG_ERR_BADPAGE;
	if (word & MT_HASEXT) return GG_ERR_UNSUPPORTED;            /* toasted attributes never travel (tupser.c:424) */
	L = (word & MT_LARGE) ? &b->large : &b->small;
	if (word & MT_HASNULL)
	{
		bitmap = mt + 4;
		start = (uint32_t) b->null_bitmap_extra;
		if (4u + (uint32_t) ((b->natts + 7) >> 3) > n) return GG_ERR_BADPAGE;
	}
	for (i = 0; i < b->natts; i++)
	{
		const gg_mt_attbind *ab = &L->att[i];
		uint32_t at;
		if (lens) lens[i] = 0;
		isnull[i] = bitmap && (bitmap[ab->null_byte] & ab->null_mask) ? 1 : 0;
		values[i] = 0;
		if (isnull[i]) continue;
		at = start + (uint32_t) ab->offset - null_save_before(b, L, bitmap, ab->phys);
		if (at + (uint32_t) ab->len > n) return GG_ERR_BADPAGE;
		if (ab->flag == 1)
		{
			uint64_t v = load_le(mt + at, ab->len);
			/* fetch_att sign-extends int2 / int4 Datums (tupmacs.h:44-68) */
			if (ab->len == 4) v = (uint64_t) (int64_t) (int32_t) v;
			else if (ab->len == 2) v = (uint64_t) (int64_t) (int16_t) v;
			else if (ab->len == 1) v = (uint64_t) (int64_t) (int8_t) v;
			values[i] = (int64_t) v;
		}
		else if (ab->flag == 2) values[i] = (int64_t) at;
		else
		{
			const uint32_t off = start + (uint32_t) load_le(mt + at, ab->len);
			uint32_t pl, body;
			if (off >= n) return GG_ERR_BADPAGE;
			if (mt[off] & 0x80)
			{
				if (mt[off] == 0x80) return GG_ERR_UNSUPPORTED;       /* external TOAST pointer */
				pl = (uint32_t) (mt[off] & 0x7F) - 1; body = off + 1;
			}
			else
			{
				uint32_t tot;
				if (off + 4 > n || (mt[off] & 0x40)) return (mt[off] & 0x40) ? GG_ERR_UNSUPPORTED : GG_ERR_BADPAGE;
				tot = ((uint32_t) (mt[off] & 0x3F) << 24) | ((uint32_t) mt[off + 1] << 16) | ((uint32_t) mt[off + 2] << 8) | mt[off + 3];
				if (tot < 4) return GG_ERR_BADPAGE;
				pl = tot - 4; body = off + 4;
			}
			if (body + pl > n) return GG_ERR_BADPAGE;
			values[i] = (int64_t) body;
			if (lens) lens[i] = (int32_t) pl;
		}
	}
	return GG_OK;
}

/* ---- tuple chunks (tupchunk.h:21-49; tupchunklist.c; tupser.c:400-603) ---- */
static void chunk_header(uint8_t *p, uint32_t size, uint32_t type) { store_le(p, size, 2); store_le(p + 2, type, 2); }

static int64_t chunk_bytes(const uint8_t *data, uint32_t n, int max_chunk, uint8_t *out, uint64_t cap, int32_t *nchunks)
{
	/* addByteStringToChunkList (tupser.c:320-370): the byte string fills chunks of at most max_chunk bytes, header included;
	 * one chunk: TC_WHOLE; several: PARTIAL_START, PARTIAL_MID ..., PARTIAL_END.  Each chunk's data is padded to
	 * TUPLE_CHUNK_ALIGN in the packet (ic_common.c:200): not at all on x86 */
	const uint32_t room = (uint32_t) max_chunk - GG_TUPLE_CHUNK_HEADER_SIZE;
	uint32_t done = 0;
	uint64_t pos = 0;
	int k = 0, total;
	if (max_chunk <= GG_TUPLE_CHUNK_HEADER_SIZE) return GG_ERR_ARG;
	total = n == 0 ? 1 : (int) ((n + room - 1) / room);
	do
	{
		const uint32_t take = n - done < room ? n - done : room;
		const uint32_t padded = (take + GG_TUPLE_CHUNK_ALIGN - 1) & ~(uint32_t) (GG_TUPLE_CHUNK_ALIGN - 1);
		uint32_t type = GG_TC_WHOLE;
		if (total > 1) type = k == 0 ? GG_TC_PARTIAL_START : (k == total - 1 ? GG_TC_PARTIAL_END : GG_TC_PARTIAL_MID);
		if (pos + GG_TUPLE_CHUNK_HEADER_SIZE + padded > cap) return GG_ERR_NOMEM;
		chunk_header(out + pos, take, type);
		memcpy(out + pos + GG_TUPLE_CHUNK_HEADER_SIZE, data + done, take);
		memset(out + pos + GG_TUPLE_CHUNK_HEADER_SIZE + take, 0, padded - take);
		pos += GG_TUPLE_CHUNK_HEADER_SIZE + padded;
		done += take;
		k++;
	} while (done < n);
	if (nchunks) *nchunks = k;
	return (int64_t) pos;
}

int64_t gg_tupser_serialize(const gg_memtuple_binding *b, const int64_t *values, const uint8_t *isnull, const int32_t *lens,
                            const void *const *ptrs, int max_chunk, uint8_t *out, uint64_t cap, int32_t *nchunks)
{
	uint32_t len = 0;
	uint8_t stackbuf[1024], *mt = stackbuf;
	int64_t rc;
	if (!b || !out) return GG_ERR_ARG;
	if (b->natts == 0)
	{
		/* a row without attributes is one TC_EMPTY chunk (tupser.c:417-424) */
		if (cap < GG_TUPLE_CHUNK_HEADER_SIZE) return GG_ERR_NOMEM;
		chunk_header(out, 0, GG_TC_EMPTY);
		if (nchunks) *nchunks = 1;
		return GG_TUPLE_CHUNK_HEADER_SIZE;
	}
	rc = gg_memtuple_form(b, values, isnull, lens, ptrs, mt, sizeof stackbuf, &len);
	if (rc == GG_ERR_NOMEM)
	{
		mt = malloc(len);
		if (!mt) return GG_ERR_NOMEM;
		rc = gg_memtuple_form(b, values, isnull, lens, ptrs, mt, len, &len);
	}
	if (rc == GG_OK) rc = chunk_bytes(mt, len, max_chunk, out, cap, nchunks);      /* the MemTuple's own 8-byte padding is the chunk padding */
	if (mt != stackbuf) free(mt);
	return rc;
}

int gg_tupser_eos(uint8_t *out, uint64_t cap)
{
	if (!out || cap < GG_TUPLE_CHUNK_HEADER_SIZE) return GG_ERR_NOMEM;
	chunk_header(out, 0, GG_TC_END_OF_STREAM);
	return GG_TUPLE_CHUNK_HEADER_SIZE;
}

/* the data area of a heap tuple (slot_deform_tuple's walk, heaptuple.c:1119-1213) as it arrives behind a TupSerHeader */
static int deform_heap_data(const gg_memtuple_binding *b, const uint8_t *bits, int natts_tuple, const uint8_t *data, uint32_t datalen,
                            int64_t *values, uint8_t *isnull, int32_t *lens, const uint8_t *base)
{
	uint32_t off = 0;
	int i;
	for (i = 0; i < b->natts; i++)
	{
		const gg_attr *a = &b->attrs[i];
		values[i] = 0; lens[i] = 0;
		if (i >= natts_tuple || (bits && !(bits[i >> 3] & (1u << (i & 7))))) { isnull[i] = 1; continue; }
		isnull[i] = 0;
		if (a->attlen == -1)
		{
			uint32_t pl, body;
			if (off < datalen && data[off] == 0) off = align_to(off, a->attalign);      /* att_align_pointer, tupmacs.h:99 */
			if (off >= datalen) return GG_ERR_BADPAGE;
			if (data[off] & 0x80)
			{
				if (data[off] == 0x80) return GG_ERR_UNSUPPORTED;
				pl = (uint32_t) (data[off] & 0x7F) - 1; body = off + 1;
			}
			else
			{
				uint32_t tot;
				if (off + 4 > datalen) return GG_ERR_BADPAGE;
				if (data[off] & 0x40) return GG_ERR_UNSUPPORTED;
				tot = ((uint32_t) (data[off] & 0x3F) << 24) | ((uint32_t) data[off + 1] << 16) | ((uint32_t) data[off + 2] << 8) | data[off + 3];
				if (tot < 4) return GG_ERR_BADPAGE;
				pl = tot - 4; body = off + 4;
			}
			if (body + pl > datalen) return GG_ERR_BADPAGE;
			values[i] = (int64_t) (data + body - base);
			lens[i] = (int32_t) pl;
			off = body + pl;
		}
		else
		{
			uint64_t v;
			off = align_to(off, a->attalign);
			if (off + (uint32_t) a->attlen > datalen) return GG_ERR_BADPAGE;
			if (!a->attbyval) { values[i] = (int64_t) (data + off - base); off += (uint32_t) a->attlen; continue; }
			v = load_le(data + off, a->attlen);
			if (a->attlen == 4) v = (uint64_t) (int64_t) (int32_t) v;
			else if (a->attlen == 2) v = (uint64_t) (int64_t) (int16_t) v;
			else if (a->attlen == 1) v = (uint64_t) (int64_t) (int8_t) v;
			values[i] = (int64_t) v;
			off += (uint32_t) a->attlen;
		}
	}
	return GG_OK;
}

int gg_tupser_deserialize(const gg_memtuple_binding *b, const uint8_t *chunks, uint64_t nbytes, uint64_t *consumed,
                          int64_t *values, uint8_t *isnull, int32_t *lens, uint8_t *strbuf, uint32_t strcap)
{
	uint64_t pos = 0, total = 0;
	uint8_t *buf = NULL;
	const uint8_t *ser;
	uint32_t type, size, sp = 0;
	int rc = GG_OK, i, first = 1;
	if (!b || !chunks || !values || !isnull || !lens || nbytes < GG_TUPLE_CHUNK_HEADER_SIZE) return GG_ERR_ARG;
	size = (uint32_t) load_le(chunks, 2); type = (uint32_t) load_le(chunks + 2, 2);
	if (type == GG_TC_END_OF_STREAM) { if (consumed) *consumed = GG_TUPLE_CHUNK_HEADER_SIZE; return 1; }
	if (type == GG_TC_EMPTY)
	{
		for (i = 0; i < b->natts; i++) { values[i] = 0; isnull[i] = 1; lens[i] = 0; }
		if (consumed) *consumed = GG_TUPLE_CHUNK_HEADER_SIZE;
		return GG_OK;
	}
	if (type == GG_TC_WHOLE)
	{
		if (GG_TUPLE_CHUNK_HEADER_SIZE + (uint64_t) size > nbytes) return GG_ERR_BADPAGE;
		ser = chunks + GG_TUPLE_CHUNK_HEADER_SIZE;
		total = size;
		pos = GG_TUPLE_CHUNK_HEADER_SIZE + (((uint64_t) size + GG_TUPLE_CHUNK_ALIGN - 1) & ~(uint64_t) (GG_TUPLE_CHUNK_ALIGN - 1));
	}
	else if (type == GG_TC_PARTIAL_START)
	{
		/* reassemble: START, MID ..., END (tupser.c:662-720) */
		uint64_t p = 0;
		for (;;)
		{
			uint32_t t, s;
			if (p + GG_TUPLE_CHUNK_HEADER_SIZE > nbytes) return GG_ERR_BADPAGE;
			s = (uint32_t) load_le(chunks + p, 2); t = (uint32_t) load_le(chunks + p + 2, 2);
			if (p + GG_TUPLE_CHUNK_HEADER_SIZE + s > nbytes) return GG_ERR_BADPAGE;
			if (first ? t != GG_TC_PARTIAL_START : (t != GG_TC_PARTIAL_MID && t != GG_TC_PARTIAL_END)) return GG_ERR_BADPAGE;
			first = 0;
			total += s;
			p += GG_TUPLE_CHUNK_HEADER_SIZE + (((uint64_t) s + GG_TUPLE_CHUNK_ALIGN - 1) & ~(uint64_t) (GG_TUPLE_CHUNK_ALIGN - 1));
			if (t == GG_TC_PARTIAL_END) break;
		}
		buf = malloc(total ? total : 1);
		if (!buf) return GG_ERR_NOMEM;
		{
			uint64_t q = 0, w = 0;
			while (q < p)
			{
				const uint32_t s = (uint32_t) load_le(chunks + q, 2);
				memcpy(buf + w, chunks + q + GG_TUPLE_CHUNK_HEADER_SIZE, s);
				w += s;
				q += GG_TUPLE_CHUNK_HEADER_SIZE + (((uint64_t) s + GG_TUPLE_CHUNK_ALIGN - 1) & ~(uint64_t) (GG_TUPLE_CHUNK_ALIGN - 1));
			}
		}
		ser = buf;
		pos = p;
	}
	else
		return GG_ERR_BADPAGE;
	if (consumed) *consumed = pos;
	if (total < 8) { free(buf); return GG_ERR_BADPAGE; }
	if (load_le(ser, 4) & MT_LEAD_BIT)
		rc = gg_memtuple_deform(b, ser, (uint32_t) total, values, isnull, lens);
	else
	{
		/* TupSerHeader { uint32 tuplen; uint16 natts; uint16 infomask } ‖ null bitmap (pad 4) ‖ data (tupser.c:282-287,497-548) */
		const uint32_t tuplen = (uint32_t) load_le(ser, 4);
		const int tnatts = (int) load_le(ser + 4, 2);
		const uint32_t infomask = (uint32_t) load_le(ser + 6, 2);
		const uint32_t nullslen = (infomask & 0x0001) ? (uint32_t) ((tnatts + 7) / 8) : 0;        /* HEAP_HASNULL, BITMAPLEN */
		const uint32_t hdr = 8 + ((nullslen + GG_TUPLE_CHUNK_ALIGN - 1) & ~(uint32_t) (GG_TUPLE_CHUNK_ALIGN - 1));
		if (tuplen > total || hdr > tuplen) { free(buf); return GG_ERR_BADPAGE; }
		rc = deform_heap_data(b, nullslen ? ser + 8 : NULL, tnatts, ser + hdr, tuplen - hdr, values, isnull, lens, ser);
	}
	/* varlena payloads out of the (possibly temporary) buffer */
	for (i = 0; rc == GG_OK && i < b->natts; i++)
	{
		if (isnull[i] || (b->attrs[i].attlen != -1 && b->attrs[i].attbyval)) continue;
		{
			const uint32_t l = b->attrs[i].attlen == -1 ? (uint32_t) lens[i] : (uint32_t) b->attrs[i].attlen;
			if (!strbuf || sp + l > strcap) { rc = GG_ERR_NOMEM; break; }
			memcpy(strbuf + sp, ser + values[i], l);
			values[i] = (int64_t) sp;
			if (b->attrs[i].attlen != -1) lens[i] = (int32_t) l;
			sp += l;
		}
	}
	free(buf);
	return rc;
}

/* ArrayType for float8[3]: ndim, dataoffset, elemtype, dim, lbound, then the elements (utils/array.h:75-81,170-183) */
int gg_float8_array3(double n, double sumx, double sumx2, uint8_t *out)
{
	const double v[3] = { n, sumx, sumx2 };
	store_le(out, 1, 4); store_le(out + 4, 0, 4); store_le(out + 8, 701, 4);
	store_le(out + 12, 3, 4); store_le(out + 16, 1, 4);
	memcpy(out + 20, v, 24);
	return 44;
}

int gg_float8_array3_read(const uint8_t *payload, int len, double *out3)
{
	if (!payload || len != 44 || load_le(payload, 4) != 1 || load_le(payload + 8, 4) != 701 || load_le(payload + 12, 4) != 3) return GG_ERR_ARG;
	memcpy(out3, payload + 20, 24);
	return GG_OK;
}
