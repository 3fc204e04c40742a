/*
 * refwrap_aocs.c — ORACLE build infrastructure for the column-oriented append-only format.
 * Thin exports over the REFERENCE'S OWN objects (compiled in place from /root/reference by ./Makefile):
 *     utils/datumstream/datumstreamblock.c   DatumStreamBlockWrite_* / DatumStreamBlockRead_*  (the block content)
 *     cdb/cdbappendonlystorageformat.c       AppendOnlyStorageFormat_* (the storage block header + checksums)
 *     port/pg_crc32c_sb8.c                   CRC-32C
 * so tests/golden/make_golden.py can have the reference write column files and read them back.
 *
 * The reference's file layer (datumstream.c, cdbappendonlystoragewrite.c, aocsam.c) drags in smgr, WAL, the catalog
 * and the buffered-append machinery and is not linked.  The two loops below stand in for it and only sequence calls
 * into the reference's functions:
 *     writer  aocs_insert_values (aocsam.c:964-1016: put; on "no room" flush the block and put again)
 *             + datumstreamwrite_block_orig (datumstream.c:889-930) + AppendOnlyStorageWrite_FinishBuffer
 *             (cdbappendonlystoragewrite.c:1327-1361: zero pad to the rounded length, make the small content header)
 *     reader  AppendOnlyStorageRead_* header walk (GetHeaderInfo / GetSmallContentHeaderInfo) + datumstreamread_block_content
 */
#include "postgres.h"
#include <setjmp.h>
#include "catalog/pg_appendonly.h"
#include "catalog/pg_type.h"
#include "cdb/cdbappendonlystorage.h"
#include "cdb/cdbappendonlystorageformat.h"
#include "port/pg_crc32c.h"
#include "utils/datumstreamblock.h"

extern sigjmp_buf *ref_err_jmp;		/* shim.c: where ereport(ERROR) lands */

#define REF_TRY(errvar) \
	sigjmp_buf _jb; sigjmp_buf *_save = ref_err_jmp; *(errvar) = 0; \
	ref_err_jmp = &_jb; \
	if (sigsetjmp(_jb, 0) != 0) { ref_err_jmp = _save; *(errvar) = 1; } else
#define REF_END() ref_err_jmp = _save

/* GUCs the two objects consult (guc_gp.c); all tracing off, integrity checks on */
bool Debug_appendonly_print_insert = false;
bool Debug_appendonly_print_insert_tuple = false;
bool Debug_appendonly_print_scan = false;
bool Debug_appendonly_print_scan_tuple = false;
bool Debug_appendonly_print_storage_headers = false;
bool Debug_appendonly_print_verify_write_block = false;
bool Debug_datumstream_block_read_check_integrity = true;
bool Debug_datumstream_block_write_check_integrity = true;
bool Debug_datumstream_read_check_large_varlena_integrity = false;
bool Debug_datumstream_read_print_varlena_info = false;
bool Debug_datumstream_write_print_small_varlena_info = false;
bool Debug_datumstream_write_use_small_initial_buffers = false;
/* CurrentMemoryContext: shim.c */

/* tuptoaster.c varattrib_untoast_ptr_len for a datum that is neither external nor compressed: nothing to free */
void
varattrib_untoast_ptr_len(Datum d, char **datastart, int *len, void **tofree)
{
	struct varlena *v = (struct varlena *) DatumGetPointer(d);

	*tofree = NULL;
	*datastart = VARDATA_ANY(v);
	*len = VARSIZE_ANY_EXHDR(v);
}

#define MAXDATUM_PER_AOCS_ORIG_BLOCK AOSmallContentHeader_MaxRowCount	/* utils/datumstream.h:33 (that header needs the catalog) */

static int no_detail(void *arg) { (void) arg; return 0; }

static void
typeinfo(DatumStreamTypeInfo *ti, int typid, int attlen, int attbyval, char attalign)
{
	ti->datumlen = attlen;				/* init_datumstream_typeinfo (datumstream.c:316-324) */
	ti->typid = typid;
	ti->align = attalign;
	ti->byval = attbyval != 0;
}

static int64
finish_block(DatumStreamBlockWrite *dsw, uint8 *out, int64 pos, int64 outcap, int blocksize, int checksum, int64 firstrow)
{
	int32 hdrlen = AoHeader_Size( /* isLong */ false, checksum, /* hasFirstRowNum */ true);
	int rowcount = DatumStreamBlockWrite_Nth(dsw);
	int64 writesz;
	int32 rounded;

	if (pos + blocksize > outcap)
		elog(ERROR, "output buffer too small");
	writesz = DatumStreamBlockWrite_Block(dsw, out + pos + hdrlen);
	rounded = AOStorage_RoundUp((int32) writesz, AORelationVersion_GetLatest());
	AOStorage_ZeroPad((out + pos + hdrlen), writesz, rounded);
	AppendOnlyStorageFormat_MakeSmallContentHeader(out + pos, checksum, /* hasFirstRowNum */ true,
												   AORelationVersion_GetLatest(), firstrow,
												   /* executorKind AOCSBK_BLOCK */ 1, rowcount, (int32) writesz, 0);
	DatumStreamBlockWrite_GetReady(dsw);
	return pos + hdrlen + rounded;
}

/*
 * One column of an AOCS segment file, compresstype=none (DatumStreamVersion_Original).  values[]: by-value Datums, or
 * for attlen -1 a pointer to the payload bytes with lens[] their length (stored the way heap_form_tuple would pass them
 * on: a 4-byte-header varlena, which the block writer turns into a short one when it fits).  Returns the file length.
 */
int64
ref_aocs_write_column(int typid, int attlen, int attbyval, char attalign,
					  const int64 *values, const int32 *lens, const uint8 *nulls, int64 nrows,
					  int blocksize, int checksum, int64 first_rownum, uint8 *out, int64 outcap, int *err)
{
	volatile int64 pos = 0;

	REF_TRY(err)
	{
		DatumStreamBlockWrite dsw;
		DatumStreamTypeInfo ti;
		int64 r, blockfirst = first_rownum;

		typeinfo(&ti, typid, attlen, attbyval, attalign);
		memset(&dsw, 0, sizeof dsw);
		DatumStreamBlockWrite_Init(&dsw, &ti, DatumStreamVersion_Original, false, false,
								   MAXDATUM_PER_AOCS_ORIG_BLOCK, MAXDATUM_PER_AOCS_ORIG_BLOCK,
								   blocksize - AoHeader_Size(false, checksum, true),
								   no_detail, NULL, no_detail, NULL);
		DatumStreamBlockWrite_GetReady(&dsw);
		for (r = 0; r < nrows; r++)
		{
			bool isnull = nulls != NULL && nulls[r] != 0;
			struct varlena *v = NULL;
			Datum d = (Datum) values[r];
			void *tofree;

			if (!isnull && attlen == -1)
			{
				v = (struct varlena *) malloc(lens[r] + VARHDRSZ);
				SET_VARSIZE(v, lens[r] + VARHDRSZ);
				memcpy(VARDATA(v), (const void *) (uintptr_t) values[r], lens[r]);
				d = PointerGetDatum(v);
			}
			if (DatumStreamBlockWrite_Put(&dsw, d, isnull, &tofree) < 0)
			{
				if (DatumStreamBlockWrite_Nth(&dsw) > 0)
				{
					pos = finish_block(&dsw, out, pos, outcap, blocksize, checksum, blockfirst);
					blockfirst = first_rownum + r;
				}
				if (DatumStreamBlockWrite_Put(&dsw, d, isnull, &tofree) < 0)
					elog(ERROR, "datum does not fit a block (large objects are out of scope)");
			}
			free(v);
		}
		if (DatumStreamBlockWrite_Nth(&dsw) > 0)
			pos = finish_block(&dsw, out, pos, outcap, blocksize, checksum, blockfirst);
		DatumStreamBlockWrite_Finish(&dsw);
		REF_END();
	}
	return pos;
}

/*
 * Read a column file back with the reference's header parser, checksum verifier and block reader.  values[]: by-value
 * Datums; for attlen -1 the byte OFFSET in `file` of the stored varlena (its header byte).  firstrows[]/rowcounts[] get
 * one entry per storage block (up to blockcap).  Returns the number of rows, -1 on a format error.
 */
int64
ref_aocs_read_column(int typid, int attlen, int attbyval, char attalign,
					 uint8 *file, int64 nbytes, int checksum,
					 int64 *values, uint8 *nulls, int64 cap,
					 int64 *firstrows, int32 *rowcounts, int blockcap, int *nblocks, int *err)
{
	volatile int64 n = 0;

	*nblocks = 0;
	REF_TRY(err)
	{
		DatumStreamBlockRead dsr;
		DatumStreamTypeInfo ti;
		int64 pos = 0;

		typeinfo(&ti, typid, attlen, attbyval, attalign);
		memset(&dsr, 0, sizeof dsr);
		DatumStreamBlockRead_Init(&dsr, &ti, DatumStreamVersion_Original, false, no_detail, NULL, no_detail, NULL);
		while (pos < nbytes)
		{
			AoHeaderKind kind;
			int32 hdrlen, overall, offset, uncompressed, compressed;
			int execkind, rowcount;
			bool hasfirst, adjusted;
			int64 firstrow;
			int32 adjcount;

			if (AppendOnlyStorageFormat_GetHeaderInfo(file + pos, checksum, &kind, &hdrlen) != AOHeaderCheckOk)
				elog(ERROR, "%s", AppendOnlyStorageFormat_GetHeaderCheckErrorStr());
			if (kind != AoHeaderKind_SmallContent)
				elog(ERROR, "unexpected header kind %d", (int) kind);
			if (checksum)
			{
				pg_crc32 stored, computed;

				if (!AppendOnlyStorageFormat_VerifyHeaderChecksum(file + pos, &stored, &computed))
					elog(ERROR, "header checksum does not match");
			}
			if (AppendOnlyStorageFormat_GetSmallContentHeaderInfo(file + pos, hdrlen, checksum, (int32) Min(nbytes - pos, 0x7fffffff),
																  &overall, &offset, &uncompressed, &execkind, &hasfirst,
																  AORelationVersion_GetLatest(), &firstrow, &rowcount,
																  &adjusted /* isCompressed */, &compressed) != AOHeaderCheckOk)
				elog(ERROR, "%s", AppendOnlyStorageFormat_GetHeaderCheckErrorStr());
			if (checksum)
			{
				pg_crc32 stored, computed;

				if (!AppendOnlyStorageFormat_VerifyBlockChecksum(file + pos, overall, &stored, &computed))
					elog(ERROR, "block checksum does not match");
			}
			if (*nblocks < blockcap)
			{
				firstrows[*nblocks] = hasfirst ? firstrow : -1;
				rowcounts[*nblocks] = rowcount;
			}
			(*nblocks)++;
			DatumStreamBlockRead_Reset(&dsr);		/* datumstreamread_block_content (datumstream.c:1153-1170) */
			DatumStreamBlockRead_GetReady(&dsr, file + pos + offset, uncompressed, firstrow, rowcount, &adjusted, &adjcount);
			while (DatumStreamBlockRead_Advance(&dsr) != 0)
			{
				Datum d = 0;
				bool isnull = false;

				if (n >= cap)
					elog(ERROR, "row capacity too small");
				DatumStreamBlockRead_Get(&dsr, &d, &isnull);
				nulls[n] = isnull;
				if (isnull)
					values[n] = 0;
				else if (attlen == -1)
					values[n] = (int64) ((uint8 *) DatumGetPointer(d) - file);
				else
					values[n] = (int64) d;
				n++;
			}
			pos += overall;
		}
		DatumStreamBlockRead_Finish(&dsr);
		REF_END();
	}
	return *err ? -1 : n;
}

/* port/pg_crc32c_sb8.c with the storage layer's conventions (INIT_CRC32C, no final inversion: cdbappendonlystorageformat.c:38-47) */
uint32
ref_aocs_crc32c(const uint8 *data, int64 len)
{
	pg_crc32c crc;

	INIT_CRC32C(crc);
	COMP_CRC32C(crc, data, len);
	return crc;
}
