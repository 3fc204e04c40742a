/* Host build of greengage_b200/csrc/gg_aocs_decode.h for tests/test_aocs_decode.py: the loop of gg_aocs_rows_kernel
 * (csrc/gg_aocs.cu) with the thread index turned into a for loop, calling the SAME gg_aocs_fetch the device kernel calls.
 * Test infrastructure: nothing in the product links this. */
#include <stdint.h>
#include "../greengage_b200/csrc/gg_aocs_decode.h"

uint32_t harness_decode_rows(const gg_aocs_devcol *cols, int ncols, uint64_t nrows, int32_t tile_rows, uint64_t *out)
{
	const uint32_t W = 1u + (uint32_t) ncols;
	uint32_t err = 0;
	uint64_t r;
	int c;

	for (r = 0; r < nrows; r++)
	{
		const int64_t tile = (int64_t) (r / (uint64_t) tile_rows);
		const int32_t lane_row = (int32_t) (r - (uint64_t) tile * (uint64_t) tile_rows);
		uint64_t mask = 0;

		for (c = 0; c < ncols; c++)
		{
			uint64_t w = 0;
			int isnull = 0;
			const uint32_t rc = gg_aocs_fetch(&cols[c], tile, lane_row, &w, &isnull);

			err |= rc;
			if (rc || isnull) { w = 0; mask |= 1ull << c; }
			out[r * W + 1 + c] = w;
		}
		out[r * W] = mask;
	}
	return err;
}
