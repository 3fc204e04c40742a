/*
 * gg_aocs.h — host side of the append-only column-oriented (AOCS) scan: what the loader does with a segment's column
 * files before they go to the device (libgghost.so, host C, no GPU needed).  SURVEY §8f rank 1; DESIGN.md §8.1.
 * The device kernels that consume the directory: gg_aocs_decode_rows (two-pass) and the fused scan gg_scanagg_run_aocs (ggb200.h).
 *
 * On-disk format handled (compresstype=none):
 *   storage blocks   AOSmallContentHeader + optional CRC-32C checksums + first row number
 *                    src/include/cdb/cdbappendonlystorage_int.h:18-150, src/backend/cdb/cdbappendonlystorageformat.c:26-321,1202-1417
 *   block content    DatumStreamBlock_Orig: 16-byte header, NULL bitmap, packed non-NULL values
 *                    src/include/utils/datumstreamblock.h:68-81, src/backend/utils/datumstream/datumstreamblock.c:150-330,1486-1748,3644-3710
 * Anything else (bulk compression, RLE_TYPE / delta "Dense" blocks, large-object blocks) is GG_ERR_UNSUPPORTED: the
 * caller keeps the CPU scan for that relation.
 */
#ifndef GG_AOCS_H
#define GG_AOCS_H

#include <stdint.h>
#include "gg_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef GG_OK
#define GG_OK               0
#define GG_ERR_UNSUPPORTED (-6)
#define GG_ERR_NOMEM       (-8)
#define GG_ERR_BADPAGE     (-9)
#define GG_ERR_ARG         (-10)
#endif

#define GG_AOCS_DEFAULT_BLOCKSIZE 32768     /* DEFAULT_APPENDONLY_BLOCK_SIZE (cdb/cdbappendonlyam.h) */
#define GG_AOCS_MAX_BLOCK_ROWS    16382     /* a small-content block takes rows while nth + 1 < 0x3FFF (datumstreamblock.c:1497) */

/* One storage block of a column file, as the device scan addresses it: datapath(row) = data_off + index of the row among
 * the block's non-NULL rows * attlen (fixed-width types), NULL-ness = bit (row - first_row) of the bitmap at null_off. */
typedef struct gg_aocs_block {
	int64_t first_row;      /* firstRowNum: row number of the block's first row within the segment file (1-based) */
	int64_t data_off;       /* byte offset in the column file of the first stored value */
	int64_t null_off;       /* byte offset of the NULL bitmap (one bit per row, LSB first, 1 = NULL); -1: no NULLs */
	int32_t nrows;          /* logical rows of the block, NULLs included */
	int32_t data_len;       /* bytes of stored values */
	int32_t stride;         /* bytes from one stored value to the next: attlen for fixed-width types; for a varlena column the
	                         * common stored size when every value of the block is a 1-byte-header varlena of the same length
	                         * (bpchar(n), n <= 126 — the loader walks the block to make sure), else 0 = irregular */
	int32_t pad;
} gg_aocs_block;            /* 40 bytes */

/* CRC-32C the way the append-only storage layer computes it (port/pg_crc32c_sb8.c; initial value 0xFFFFFFFF and, "by
 * historical accident", no final inversion — cdbappendonlystorageformat.c:38-47).  Uses the SSE4.2 instruction when the
 * CPU has it. */
#ifndef __CUDACC_RTC__      /* the run-time kernel compiler only needs the types of this header */
uint32_t gg_aocs_crc32c(const uint8_t *p, int64_t n);

/* Walk a column file: validate every storage-block header (and both checksums when the relation has checksum=true),
 * validate the datum-stream block header inside, and fill the block directory.  AppendOnlyStorageRead's header walk
 * (cdbappendonlystorageread.c) + DatumStreamBlockRead_GetReadyOrig (datumstreamblock.c:150-330).
 * dir may be NULL (count only).  GG_ERR_BADPAGE: checksum or structure error; GG_ERR_UNSUPPORTED: a block kind outside the
 * format above; GG_ERR_NOMEM: more than cap blocks. */
int gg_aocs_index_column(const gg_attr *att, const uint8_t *file, int64_t nbytes, int checksum,
                         gg_aocs_block *dir, int64_t cap, int64_t *nblocks, int64_t *nrows);
#endif

/* The scan's unit of work is a TILE of tile_rows consecutive rows (the same rows of every projected column; the columns'
 * storage blocks end at different rows because their widths differ).  For one column, per tile: the storage block that
 * holds the tile's first row and how many of that block's rows before it are NULL — so that a thread handling row r of
 * the tile finds its value without searching:
 *     j = row_in_block + (r - tile's first row);  while j >= nrows of the block: j -= nrows, next block;
 *     value index in the block = j - (NULLs among the block's rows [0, j)), of which `nulls_before` are already counted
 *     for the tile's first block (the rest is a popcount over at most tile_rows bitmap bits).
 * Rows are addressed by POSITION in the file (0-based, counting every row of every block in order), which is what
 * aocs_getnext returns them in; first_row numbers may have gaps between blocks and are not used for addressing. */
typedef struct gg_aocs_tile {
	int32_t block;          /* index in the directory of the block holding the tile's first row */
	int32_t row_in_block;   /* position of the tile's first row inside that block */
	int32_t nulls_before;   /* NULL rows among the block's rows [0, row_in_block) */
	int32_t pad;
} gg_aocs_tile;             /* 16 bytes */

/* ntiles must be ceil(nrows / tile_rows) for the directory's row total; file is needed to count NULL bits. */
#ifndef __CUDACC_RTC__
int gg_aocs_plan_tiles(const gg_aocs_block *dir, int64_t nblocks, const uint8_t *file, int32_t tile_rows,
                       gg_aocs_tile *tiles, int64_t ntiles);
#endif

/* ---- what the device side takes (libggb200.so: gg_aocs_decode_rows in ggb200.h; kernel in csrc/gg_aocs.cu) ---- */

/* how a stored value becomes the 64-bit Datum word of a GG_FMT_DATUMROWS row (gg_plan.h) */
enum gg_aocs_kind {
	GG_AOCS_K_W8 = 0,       /* 8-byte by-value types (int8, float8, timestamp): the bits */
	GG_AOCS_K_I4 = 1,       /* int4, date: sign-extended like DatumGetInt32 */
	GG_AOCS_K_I2 = 2,       /* int2: sign-extended */
	GG_AOCS_K_B1 = 3,       /* bool / "char": the byte */
	GG_AOCS_K_BPCHAR = 4,   /* 1-byte-header varlena of <= 8 payload bytes, trailing blanks stripped (bcTruelen), packed LSB first */
	GG_AOCS_K_TEXT = 5      /* the same without stripping (varchar, text) */
};

#define GG_AOCS_E_RANGE     1u      /* the walk ran off the directory: plan and directory disagree */
#define GG_AOCS_E_IRREGULAR 2u      /* a block whose values have no common stride (stride 0) or a string too long to pack */

/* one projected column as the kernel sees it (device pointers).
 * The fused scan (gg_scanagg_run_aocs) bulk-copies runs of a column's values into shared memory in 16-byte units: `file` should
 * start on a 16-byte boundary and be readable up to the next multiple of 16 past its last byte (any cudaMalloc'ed buffer is; a
 * loader that packs several files into one allocation pads each to 16).  A file that does not start on such a boundary is
 * read value by value instead — same answer, slower. */
typedef struct gg_aocs_devcol {
	const uint8_t *file;
	const gg_aocs_block *dir;
	const gg_aocs_tile *tiles;
	int64_t nblocks;
	int32_t kind;           /* enum gg_aocs_kind */
	int32_t pad;
} gg_aocs_devcol;           /* 40 bytes */

/* Streaming writer of one column file (what an INSERT / COPY into the relation appends: aocs_insert_values,
 * aocsam.c:964-1016 -> datumstreamwrite_put / datumstreamwrite_block_orig -> AppendOnlyStorageWrite_FinishBuffer).
 * Used by the synthetic loader and the tests; byte-identical to what the reference writes for the same values. */
#ifndef __CUDACC_RTC__
typedef struct gg_aocs_writer gg_aocs_writer;
int gg_aocs_writer_create(const gg_attr *att, int blocksize, int checksum, int64_t first_rownum,
                          uint8_t *out, int64_t outcap, gg_aocs_writer **w);
/* value: the Datum of a by-value type, or a pointer to attlen bytes (fixed by-reference) / to len payload bytes (varlena) */
int gg_aocs_writer_put(gg_aocs_writer *w, int64_t value, int32_t len, int isnull);
/* flush the open block, report the file length, free the writer */
int gg_aocs_writer_finish(gg_aocs_writer *w, int64_t *nbytes);
/* a safe output-buffer size for nrows values of at most maxlen payload bytes each */
int64_t gg_aocs_file_bound(const gg_attr *att, int64_t nrows, int32_t maxlen, int blocksize, int checksum);
#endif

#ifdef __cplusplus
}
#endif
#endif /* GG_AOCS_H */
