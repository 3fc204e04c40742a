/*
 * gg_aocs_decode.h — how ONE thread finds and loads ONE value of an append-only column-oriented (AOCS) column file
 * resident in device memory, through the loader's block directory and tile plan (include/gg_aocs.h).
 * The same source is the device function of gg_aocs.cu and, compiled by gcc in tests/test_aocs_decode.py, the function the
 * CPU tests hold against the oracle on the reference-written files — plain C, no CUDA intrinsics except the popcount.
 *
 * Reading side of DatumStreamBlockRead_AdvanceOrig / GetOrig (src/include/utils/datumstreamblock.h:1216-1570): rows of a
 * block in order, NULL rows take no space in the value area, value i of the block sits at data_off + i * stride.
 */
#ifndef GG_AOCS_DECODE_H
#define GG_AOCS_DECODE_H

#include <stdint.h>
#include "../../include/gg_aocs.h"

#if defined(__CUDACC__)
#define GG_AOCS_FN __host__ __device__ __forceinline__
#else
#define GG_AOCS_FN static inline
#endif

GG_AOCS_FN int gg_aocs_popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}

/* NULL bits among rows [from, to) of a block's bitmap (8-byte aligned in the file: storage blocks start at multiples of 8,
 * the bitmap 24 or 32 + 16 bytes in) */
GG_AOCS_FN int32_t gg_aocs_nulls_between(const uint8_t *bitmap, int32_t from, int32_t to)
{
	const uint64_t *w = (const uint64_t *) bitmap;
	int32_t n = 0, i;

	if (to <= from)
		return 0;
	for (i = from >> 6; i <= (to - 1) >> 6; i++)
	{
		uint64_t m = w[i];
		const int32_t lo = i << 6;

		if (from > lo)
			m &= ~(uint64_t) 0 << (from - lo);
		if (to < lo + 64)
			m &= ~(~(uint64_t) 0 << (to - lo));
		n += gg_aocs_popc64(m);
	}
	return n;
}

/* a stored value at p -> the 64-bit Datum word of a datum row (enum gg_aocs_kind); 0, or GG_AOCS_E_IRREGULAR for a string that
 * does not pack into 8 bytes */
GG_AOCS_FN uint32_t gg_aocs_value(int kind, const uint8_t *p, uint64_t *word)
{
	uint64_t v = 0;

	switch (kind)
	{
		case GG_AOCS_K_W8:
			v = *(const uint64_t *) p;						/* data_off is 8-aligned, stride 8 */
			break;
		case GG_AOCS_K_I4:
			v = (uint64_t) (int64_t) *(const int32_t *) p;
			break;
		case GG_AOCS_K_I2:
			v = (uint64_t) (int64_t) *(const int16_t *) p;
			break;
		case GG_AOCS_K_B1:
			v = *p;
			break;
		default:
		{
			int n = (p[0] & 0x7F) - 1, i;					/* payload bytes after the 1-byte header */

			if (kind == GG_AOCS_K_BPCHAR)
				while (n > 0 && p[n] == ' ')				/* bcTruelen (varchar.c:653) */
					n--;
			if (n > 8)
				return GG_AOCS_E_IRREGULAR;
			for (i = 0; i < n; i++)
				v |= (uint64_t) p[1 + i] << (8 * i);
			break;
		}
	}
	*word = v;
	return 0;
}

/* Row `lane_row` of tile `tile`: 0 and *word / *isnull set, or a GG_AOCS_E_* bit. */
GG_AOCS_FN uint32_t gg_aocs_fetch(const gg_aocs_devcol *c, int64_t tile, int32_t lane_row, uint64_t *word, int *isnull)
{
	const gg_aocs_tile t = c->tiles[tile];
	int64_t b = t.block;
	int64_t j = (int64_t) t.row_in_block + lane_row;
	int32_t counted_to = t.row_in_block, nulls = t.nulls_before;
	const uint8_t *p;
	gg_aocs_block blk;

	while (b < c->nblocks && j >= c->dir[b].nrows)		/* the column's next storage block; bounded by the directory */
	{
		j -= c->dir[b].nrows;
		b++;
		counted_to = 0;
		nulls = 0;
	}
	if (b >= c->nblocks)
		return GG_AOCS_E_RANGE;
	blk = c->dir[b];
	*isnull = 0;
	*word = 0;
	if (blk.null_off >= 0)
	{
		const uint8_t *bitmap = c->file + blk.null_off;

		if ((bitmap[j >> 3] >> (j & 7)) & 1)
		{
			*isnull = 1;
			return 0;
		}
		nulls += gg_aocs_nulls_between(bitmap, counted_to, (int32_t) j);
	}
	if (blk.stride <= 0)
		return GG_AOCS_E_IRREGULAR;
	p = c->file + blk.data_off + (j - nulls) * (int64_t) blk.stride;
	return gg_aocs_value(c->kind, p, word);
}

#endif /* GG_AOCS_DECODE_H */
