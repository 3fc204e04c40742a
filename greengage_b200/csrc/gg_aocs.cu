/*
 * gg_aocs.cu — append-only column-oriented (AOCS) column files resident in device memory -> GG_FMT_DATUMROWS rows.
 *
 * First device step of the AOCS scan: the projected columns of a segment file are
 * decoded into the row format every operator of this engine already scans (gg_relation_attach_rows: SeqScan, HashJoin
 * build / probe, Motion send, Sort), the way aocs_getnext (aocsam.c:700-800) fills a slot's Datum arrays from the
 * per-column datum streams.  Only the projected columns are ever read — or moved over PCIe when the files come from
 * host memory, which is what bounds the end-to-end number.
 *
 * One thread per row: the loader's tile plan (gg_aocs_plan_tiles) gives the storage block and NULL count at the tile's
 * first row of every column, gg_aocs_fetch (gg_aocs_decode.h — the same source the CPU tests run against the oracle)
 * walks forward from there.  Reads are coalesced (consecutive rows = consecutive values of a block); the row-major
 * writes are strided and merge in L2.  A fused kernel that feeds the accumulator program straight from the column tiles
 * is the follow-up; this kernel is HBM-bound at (projected bytes + 8 * (1 + ncols)) per row.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "gg_engine.h"
#include "gg_aocs_decode.h"

#define GG_AOCS_MAX_COLS GG_MAX_ATTS

struct AocsCols {
	gg_aocs_devcol c[GG_AOCS_MAX_COLS];
};

__global__ void __launch_bounds__(256)
gg_aocs_rows_kernel(const AocsCols cols, int ncols, uint64_t nrows, int32_t tile_rows, uint64_t *out, uint32_t *errflags)
{
	const uint32_t W = 1u + (uint32_t) ncols;
	uint32_t err = 0;

	for (uint64_t r = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; r < nrows; r += (uint64_t) gridDim.x * blockDim.x)
	{
		const int64_t tile = (int64_t) (r / (uint64_t) tile_rows);
		const int32_t lane_row = (int32_t) (r - (uint64_t) tile * (uint64_t) tile_rows);
		uint64_t mask = 0;

		for (int c = 0; c < ncols; c++)
		{
			uint64_t w = 0;
			int isnull = 0;
			const uint32_t rc = gg_aocs_fetch(&cols.c[c], tile, lane_row, &w, &isnull);

			err |= rc;
			if (rc || isnull) { w = 0; mask |= 1ull << c; }
			out[r * W + 1 + c] = w;
		}
		out[r * W] = mask;
	}
	if (err)
		atomicOr(errflags, err);
}

extern "C" int gg_aocs_decode_rows(gg_engine *e, const struct gg_aocs_devcol *cols, int ncols, uint64_t nrows, int32_t tile_rows,
                                   void *device_rows)
{
	if (!e || !cols || ncols < 1 || ncols > GG_AOCS_MAX_COLS || tile_rows < 1 || (nrows && !device_rows)) return GG_ERR_ARG;
	if (((uintptr_t) device_rows) & 15) { gg_set_error("row buffer must be 16-byte aligned"); return GG_ERR_ARG; }
	GG_CUDA(cudaSetDevice(e->device));
	if (nrows == 0) return GG_OK;
	AocsCols k;
	for (int c = 0; c < ncols; c++)
	{
		if (!cols[c].file || !cols[c].dir || !cols[c].tiles || cols[c].nblocks < 1 || cols[c].kind < GG_AOCS_K_W8 || cols[c].kind > GG_AOCS_K_TEXT)
		{
			gg_set_error("AOCS column %d: incomplete descriptor", c);
			return GG_ERR_ARG;
		}
		k.c[c] = cols[c];
	}
	uint32_t *d_err = nullptr;
	GG_CUDA(cudaMalloc((void **) &d_err, 4));
	cudaError_t ce = cudaMemsetAsync(d_err, 0, 4, e->stream);
	uint32_t h_err = 0;
	if (ce == cudaSuccess)
	{
		const uint64_t want = (nrows + 255) / 256;
		const int grid = (int) (want < (uint64_t) e->sm_count * 16 ? want : (uint64_t) e->sm_count * 16);
		ce = cudaEventRecord(e->ev_start, e->stream);
		gg_aocs_rows_kernel<<<grid, 256, 0, e->stream>>>(k, ncols, nrows, tile_rows, (uint64_t *) device_rows, d_err);
		e->launches++;
		if (ce == cudaSuccess) ce = cudaGetLastError();
		if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_stop, e->stream);
		if (ce == cudaSuccess) ce = cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, e->stream);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
	}
	cudaFree(d_err);
	if (ce != cudaSuccess) return gg_cuda_fail(ce, "gg_aocs_decode_rows");
	e->timed = true;
	if (h_err & GG_AOCS_E_RANGE) { gg_set_error("AOCS decode: tile plan and block directory disagree"); return GG_ERR_BADPAGE; }
	if (h_err & GG_AOCS_E_IRREGULAR)
	{
		gg_set_error("AOCS decode: a projected column has blocks without a common value stride (long or mixed-length strings)");
		return GG_ERR_UNSUPPORTED;
	}
	return GG_OK;
}
