/*
 * gg_tupser.h — the row formats a Motion shares with CPU segments (libgghost.so, host C): MemTuple and the tuple-chunk
 * wire format.
 *
 * A GPU segment that sits on the same Motion as a CPU segment — a CPU FINAL-stage Agg above a GPU PARTIAL stage, a mixed
 * cluster — has to put its rows on the interconnect exactly as the reference's senders do and read what they send:
 *     MemTuple            src/include/access/memtup.h:52-80, src/backend/access/common/memtuple.c:24-56 (layout),
 *                         :175-417 (create_col_bind), :551 (memtuple_form_to), :917 (memtuple_getattr)
 *     tuple chunks        src/include/cdb/tupchunk.h:21-49 (4-byte header {u16 size, u16 type}), TUPLE_CHUNK_ALIGN (cdbvars.h:30-34:
 *                         1, i.e. no padding, on everything but sparc), src/backend/cdb/motion/tupser.c:400-603 (SerializeTuple:
 *                         a virtual / MemTuple slot travels as its MemTuple; a heap tuple without toasted attributes as
 *                         TupSerHeader ‖ null bitmap ‖ data, each padded to TUPLE_CHUNK_ALIGN, tupser.c:282-287,497-548),
 *                         :609 (CvtChunksToTup)
 * Byte-for-byte against the reference's own memtuple.o / tupser.o: tests/golden/memtuple_kat.json, tests/test_tupser.py.
 *
 * Rows cross this API the way the executor's slots hold them: one 64-bit Datum per column + null flags; a varlena column
 * (bpchar / varchar / text, or any other varlena such as avg's float8[3] transition array) is its payload bytes (no header):
 * ptrs[i] + lens[i], or for strings of at most 8 bytes the packed value in values[i] with lens[i] (ptrs[i] == NULL).
 */
#ifndef GG_TUPSER_H
#define GG_TUPSER_H

#include <stdint.h>
#include "gg_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GG_MT_MAX_ATTS 64

typedef struct gg_mt_attbind {
	int32_t offset;             /* of the attribute (fixed-width value, or the 2/4-byte offset word of a varlena) */
	int16_t len;                /* bytes it occupies in the fixed area */
	int16_t len_aligned;        /* len padded for the physically following attribute: what a NULL saves */
	uint8_t flag;               /* 1 by value, 2 fixed-length by reference, 3 varlena (MemTupleBindFlag, memtup.h:20-26) */
	uint8_t null_byte, null_mask;
	uint8_t phys;               /* physical position (8-byte aligned attributes first, then 4, 2, 1: memtuple.c:238-251) */
} gg_mt_attbind;

typedef struct gg_mt_layout {
	gg_mt_attbind att[GG_MT_MAX_ATTS];
	int32_t var_start;          /* where the varlena bodies begin when nothing is NULL */
	int32_t pad;
} gg_mt_layout;

typedef struct gg_memtuple_binding {     /* MemTupleBinding, memtup.h:48-56 */
	int32_t natts;
	int32_t column_align;       /* 8 if any attribute is 8-byte aligned, else 4 */
	int32_t null_bitmap_extra;  /* bytes the NULL bitmap needs beyond the 4 that are free after the length word */
	int32_t pad;
	gg_attr attrs[GG_MT_MAX_ATTS];
	gg_mt_layout small, large;  /* varlena offsets as 2 bytes (tuples up to 0xFFF0 bytes) / 4 bytes */
} gg_memtuple_binding;

/* create_memtuple_binding (memtuple.c:420).  GG_ERR_UNSUPPORTED: cstring (attlen -2) attributes, more than GG_MT_MAX_ATTS */
int gg_memtuple_bind(const gg_attr *attrs, int natts, gg_memtuple_binding *out);
/* memtuple_form_to (memtuple.c:551): *len receives the tuple's length; GG_ERR_NOMEM if cap is too small (len still set) */
int gg_memtuple_form(const gg_memtuple_binding *b, const int64_t *values, const uint8_t *isnull, const int32_t *lens,
                     const void *const *ptrs, uint8_t *out, uint32_t cap, uint32_t *len);
/* memtuple_deform (memtuple.c:917): values[i] = Datum bits; varlena: offset of the payload from the tuple's start, lens[i]
 * its length.  GG_ERR_BADPAGE if the tuple is not a well-formed MemTuple of this binding within `len` bytes. */
int gg_memtuple_deform(const gg_memtuple_binding *b, const uint8_t *mt, uint32_t len, int64_t *values, uint8_t *isnull, int32_t *lens);
uint32_t gg_memtuple_size(const uint8_t *mt);          /* memtuple_get_size */

/* ---- tuple chunks ---- */
enum { GG_TC_WHOLE = 0, GG_TC_PARTIAL_START = 1, GG_TC_PARTIAL_MID = 2, GG_TC_PARTIAL_END = 3, GG_TC_END_OF_STREAM = 4, GG_TC_EMPTY = 5 };
#define GG_TUPLE_CHUNK_HEADER_SIZE 4
#define GG_TUPLE_CHUNK_ALIGN 1          /* cdbvars.h:30-34: 4 on sparc only; everywhere else chunks and their parts are not padded */

/* SerializeTuple for a row that travels as a MemTuple (tupser.c:436-494): the chunks, each {u16 size, u16 type} + data,
 * back to back in out.  max_chunk = Gp_max_tuple_chunk_size (header included), e.g. 8192 - the packet header - 4.
 * Returns the number of bytes written (< 0: GG_ERR_*); *nchunks the chunk count. */
int64_t gg_tupser_serialize(const gg_memtuple_binding *b, const int64_t *values, const uint8_t *isnull, const int32_t *lens,
                            const void *const *ptrs, int max_chunk, uint8_t *out, uint64_t cap, int32_t *nchunks);
/* End-of-stream chunk (SendEndOfStream, cdbmotion.c:532) */
int gg_tupser_eos(uint8_t *out, uint64_t cap);
/* CvtChunksToTup (tupser.c:609) for one tuple's chunks at `chunks`: reassembles PARTIAL_* chunks and deforms either form —
 * a MemTuple, or the TupSerHeader form of a heap tuple.  Varlena payloads are copied to strbuf: values[i] = offset in strbuf,
 * lens[i] = length.  *consumed = bytes of chunks used.  Returns GG_OK, 1 for an end-of-stream chunk, or GG_ERR_*. */
int gg_tupser_deserialize(const gg_memtuple_binding *b, const uint8_t *chunks, uint64_t nbytes, uint64_t *consumed,
                          int64_t *values, uint8_t *isnull, int32_t *lens, uint8_t *strbuf, uint32_t strcap);
/* avg(float8)'s transition value as the reference ships it between Agg stages: a float8[3] array {N, sumX, sumX2}
 * (utils/array.h:75-81, SURVEY App. A "two-stage interchange"): 44 bytes of varlena payload (the varlena header is the
 * row format's, not the payload's).  Returns 44. */
int gg_float8_array3(double n, double sumx, double sumx2, uint8_t *out44);
int gg_float8_array3_read(const uint8_t *payload, int len, double *out3);

#ifdef __cplusplus
}
#endif
#endif /* GG_TUPSER_H */
