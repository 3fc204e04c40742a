/*
 * refwrap_motion.c — ORACLE infrastructure: drives the reference's own memtuple.o, tupser.o and tupchunklist.o
 * (compiled in place from /root/reference by this directory's Makefile) so that tests/golden/make_golden.py can record
 * what the reference itself writes for a row that travels through a Motion:
 *     create_memtuple_binding / memtuple_form_to / memtuple_deform    access/common/memtuple.c:420,551,917
 *     SerializeTuple / CvtChunksToTup                                  cdb/motion/tupser.c:400,609
 * Test infrastructure only; nothing here is product code, and nothing is copied from the reference: the calls are
 * sequenced the way ExecFetchSlotMemTuple + SendTuple (cdbmotion.c:434) sequence them.
 */
#include "postgres.h"
#include "access/htup_details.h"
#include "access/memtup.h"
#include "access/tupmacs.h"
#include "catalog/pg_type.h"
#include "cdb/cdbmotion.h"
#include "cdb/tupser.h"
#include "cdb/tupchunklist.h"
#include "executor/tuptable.h"

/* what the reference's objects expect from the backend around them */
MemoryContext TopMemoryContext = NULL;
int Gp_max_tuple_chunk_size = 8192 - 64 - 4;
MemoryContext AllocSetContextCreate(MemoryContext parent, const char *name, Size a, Size b, Size c) { (void) parent; (void) name; (void) a; (void) b; (void) c; return (MemoryContext) 1; }
void MemoryContextReset(MemoryContext c) { (void) c; }

typedef struct ref_attr { int32 atttypid; int32 atttypmod; int16 attlen; int8 attalign; int8 attbyval; int8 attnotnull; int8 pad[3]; } ref_attr;

static TupleDesc
mk_desc2(int natts, const ref_attr *a)
{
	TupleDesc d = (TupleDesc) calloc(1, sizeof(struct tupleDesc));
	int i;

	d->natts = natts;
	d->attrs = (Form_pg_attribute *) calloc(natts, sizeof(Form_pg_attribute));
	d->tdtypeid = RECORDOID;
	d->tdtypmod = -1;
	d->tdrefcount = -1;
	for (i = 0; i < natts; i++)
	{
		Form_pg_attribute att = (Form_pg_attribute) calloc(1, ATTRIBUTE_FIXED_PART_SIZE);

		att->atttypid = a[i].atttypid;
		att->attlen = a[i].attlen;
		att->attnum = i + 1;
		att->attcacheoff = -1;
		att->atttypmod = a[i].atttypmod;
		att->attbyval = a[i].attbyval;
		att->attalign = a[i].attalign;
		att->attstorage = (a[i].attlen == -1) ? 'x' : 'p';
		att->attnotnull = a[i].attnotnull;
		d->attrs[i] = att;
	}
	return d;
}

static struct varlena *
mk_varlena4(const char *payload, int len)
{
	struct varlena *v = (struct varlena *) malloc(len + VARHDRSZ + 8);

	SET_VARSIZE(v, len + VARHDRSZ);
	memcpy(VARDATA(v), payload, len);
	return v;
}

static void
to_datums(int natts, const ref_attr *a, const int64 *vals, const int32 *lens, const uint8 *isnull, Datum *values, bool *nulls, void **tofree)
{
	int i;

	for (i = 0; i < natts; i++)
	{
		tofree[i] = NULL;
		nulls[i] = isnull && isnull[i];
		values[i] = (Datum) vals[i];
		if (!nulls[i] && a[i].attlen == -1)
		{
			tofree[i] = mk_varlena4((const char *) (uintptr_t) vals[i], lens[i]);
			values[i] = PointerGetDatum(tofree[i]);
		}
		else if (a[i].attlen == 4)
			values[i] = Int32GetDatum((int32) vals[i]);
		else if (a[i].attlen == 2)
			values[i] = Int16GetDatum((int16) vals[i]);
		else if (a[i].attlen == 1)
			values[i] = CharGetDatum((char) vals[i]);
	}
}

/* vals[]: Datum bits, or pointer to the payload bytes of a varlena (+ lens[]).  Returns the MemTuple's length. */
int
ref_memtuple_form(int natts, const ref_attr *a, const int64 *vals, const int32 *lens, const uint8 *isnull, uint8 *out, int outcap)
{
	TupleDesc d = mk_desc2(natts, a);
	MemTupleBinding *b = create_memtuple_binding(d);
	Datum values[128];
	bool nulls[128];
	void *tofree[128];
	MemTuple mt;
	int i, len;

	to_datums(natts, a, vals, lens, isnull, values, nulls, tofree);
	mt = memtuple_form_to(b, values, nulls, NULL, NULL, false);
	len = (int) memtuple_get_size(mt);
	if (len <= outcap) memcpy(out, mt, len);
	free(mt);
	for (i = 0; i < natts; i++) if (tofree[i]) free(tofree[i]);
	return len;
}

/* memtuple_deform: vals[] = Datum bits or, for varlenas, the byte offset of the datum from the tuple's start */
int
ref_memtuple_deform(int natts, const ref_attr *a, uint8 *mt, int64 *vals, uint8 *isnull)
{
	TupleDesc d = mk_desc2(natts, a);
	MemTupleBinding *b = create_memtuple_binding(d);
	Datum values[128];
	bool nulls[128];
	int i;

	memtuple_deform((MemTuple) mt, b, values, nulls);
	for (i = 0; i < natts; i++)
	{
		isnull[i] = nulls[i];
		if (nulls[i]) vals[i] = 0;
		else if (a[i].attlen == -1) vals[i] = (int64) ((uint8 *) DatumGetPointer(values[i]) - mt);
		else if (a[i].attlen == 4) vals[i] = (int64) DatumGetInt32(values[i]);
		else if (a[i].attlen == 2) vals[i] = (int64) DatumGetInt16(values[i]);
		else if (a[i].attlen == 1) vals[i] = (int64) DatumGetChar(values[i]);
		else vals[i] = (int64) values[i];
	}
	return natts;
}

/* the binding itself: per attribute { offset, len, len_aligned, flag, null_byte, null_mask } of the short or the large
 * layout; returns var_start.  info[1] = null_bitmap_extra_size, info[0] = column_align */
int
ref_memtuple_binding(int natts, const ref_attr *a, int large, int32 *out6, int32 *info)
{
	TupleDesc d = mk_desc2(natts, a);
	MemTupleBinding *b = create_memtuple_binding(d);
	MemTupleBindingCols *c = large ? &b->large_bind : &b->bind;
	int i;

	for (i = 0; i < natts; i++)
	{
		out6[i * 6 + 0] = c->bindings[i].offset;
		out6[i * 6 + 1] = c->bindings[i].len;
		out6[i * 6 + 2] = c->bindings[i].len_aligned;
		out6[i * 6 + 3] = c->bindings[i].flag;
		out6[i * 6 + 4] = c->bindings[i].null_byte;
		out6[i * 6 + 5] = c->bindings[i].null_mask;
	}
	info[0] = b->column_align;
	info[1] = b->null_bitmap_extra_size;
	return (int) c->var_start;
}

/* SerializeTuple of one row held in a slot, out of line (the chunk list path), with the given maximum chunk size: the
 * chunks exactly as they would go into packets, concatenated (each padded to TUPLE_CHUNK_ALIGN the way the packet
 * assembly does, ic_common.c:200).  as_heap != 0: the slot holds a heap tuple (TupSerHeader form), else a MemTuple. */
int
ref_serialize_tuple(int natts, const ref_attr *a, const int64 *vals, const int32 *lens, const uint8 *isnull, int as_heap,
					int max_chunk, uint8 *out, int outcap, int32 *nchunks)
{
	TupleDesc d = mk_desc2(natts, a);
	SerTupInfo info;
	TupleTableSlot slot;
	TupleChunkListData tcl;
	TupleChunkListItem it;
	struct directTransportBuffer b;
	Datum values[128];
	bool nulls[128];
	void *tofree[128];
	int i, total = 0, n = 0, saved = Gp_max_tuple_chunk_size;

	Gp_max_tuple_chunk_size = max_chunk;
	memset(&info, 0, sizeof info);
	info.tupdesc = d;
	info.chunkCache.len = 0;
	info.chunkCache.items = NULL;
	memset(&slot, 0, sizeof slot);
	slot.tts_tupleDescriptor = d;
	slot.tts_mt_bind = create_memtuple_binding(d);
	to_datums(natts, a, vals, lens, isnull, values, nulls, tofree);
	if (as_heap)
		slot.PRIVATE_tts_heaptuple = heap_form_tuple(d, values, nulls);
	else
		slot.PRIVATE_tts_memtuple = memtuple_form_to(slot.tts_mt_bind, values, nulls, NULL, NULL, false);
	memset(&b, 0, sizeof b);
	memset(&tcl, 0, sizeof tcl);
	(void) SerializeTuple(&slot, &info, &b, &tcl, 0);
	for (it = tcl.p_first; it; it = it->p_next)
	{
		int len = it->chunk_length, padded = TYPEALIGN(TUPLE_CHUNK_ALIGN, len);
		uint16 sz = (uint16) (len - TUPLE_CHUNK_HEADER_SIZE);

		if (total + padded <= outcap)
		{
			memcpy(out + total, it->chunk_data, len);
			memcpy(out + total, &sz, 2);          /* SendTuple sets the size when it hands the chunk on (cdbmotion.c:487) */
			memset(out + total + len, 0, padded - len);
		}
		total += padded;
		n++;
	}
	*nchunks = n;
	Gp_max_tuple_chunk_size = saved;
	for (i = 0; i < natts; i++) if (tofree[i]) free(tofree[i]);
	return total;
}

/* CvtChunksToTup over chunks laid out as above; the resulting tuple is deformed into vals/isnull (varlenas: payload copied
 * to strout, vals = offset into strout, lens = payload length).  Returns 1 if the tuple came back as a MemTuple, 0 heap. */
int
ref_deserialize_tuple(int natts, const ref_attr *a, const uint8 *chunks, int nbytes, int64 *vals, int32 *lens, uint8 *isnull,
					  uint8 *strout, int strcap)
{
	TupleDesc d = mk_desc2(natts, a);
	SerTupInfo info;
	TupleChunkListData tcl;
	GenericTuple tup;
	Datum values[128];
	bool nulls[128];
	int pos = 0, i, sp = 0, ismem;

	memset(&info, 0, sizeof info);
	info.tupdesc = d;
	info.values = values;
	info.nulls = nulls;
	memset(&tcl, 0, sizeof tcl);
	while (pos < nbytes)
	{
		uint16 sz;
		TupleChunkListItem it;

		memcpy(&sz, chunks + pos, 2);
		it = (TupleChunkListItem) calloc(1, sizeof(TupleChunkListItemData) + sz + TUPLE_CHUNK_HEADER_SIZE + 8);
		it->chunk_length = sz + TUPLE_CHUNK_HEADER_SIZE;
		memcpy(it->chunk_data, chunks + pos, it->chunk_length);
		appendChunkToTCList(&tcl, it);
		pos += TYPEALIGN(TUPLE_CHUNK_ALIGN, it->chunk_length);
	}
	tup = CvtChunksToTup(&tcl, &info, NULL);
	ismem = is_memtuple(tup);
	if (ismem)
		memtuple_deform((MemTuple) tup, create_memtuple_binding(d), values, nulls);
	else
		heap_deform_tuple((HeapTuple) tup, d, values, nulls);
	for (i = 0; i < natts; i++)
	{
		isnull[i] = nulls[i];
		lens[i] = 0;
		if (nulls[i]) { vals[i] = 0; continue; }
		if (a[i].attlen == -1)
		{
			struct varlena *v = (struct varlena *) DatumGetPointer(values[i]);
			int l = VARSIZE_ANY_EXHDR(v);

			if (sp + l <= strcap) memcpy(strout + sp, VARDATA_ANY(v), l);
			vals[i] = sp; lens[i] = l; sp += l;
		}
		else if (a[i].attlen == 4) vals[i] = (int64) DatumGetInt32(values[i]);
		else if (a[i].attlen == 2) vals[i] = (int64) DatumGetInt16(values[i]);
		else if (a[i].attlen == 1) vals[i] = (int64) DatumGetChar(values[i]);
		else vals[i] = (int64) values[i];
	}
	return ismem;
}
