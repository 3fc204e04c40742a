/*
 * gg_scanagg.cu — fused heap SeqScan -> qual -> projection -> partial HashAggregate.
 *
 * Replaces, for one segment, the per-tuple loop
 *   ExecAgg -> agg_hash_initial_pass -> ExecSeqScan -> heap_getnext -> heapgetpage
 *   (nodeAgg.c:1123, execHHashagg.c:905, nodeSeqscan.c:128, heapam.c:312-463,767-1006;
 *    SURVEY §3.3 hot loops B and C)
 * with one persistent, warp-specialised kernel:
 *
 *   producer warp    TMA bulk-copies whole 32 KB heap pages HBM -> shared-memory ring
 *                    (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP)
 *   consumer warps   lane = one line pointer: ItemId decode, visibility, attribute walk
 *                    (slot_deform_tuple semantics), then ONE compiled program per row
 *                    (gg_program.h): scan qual -> grouping keys -> group lookup in a per-block
 *                    key table -> aggregate arguments
 *   accumulate       MODE_PRIV: every consumer thread owns a private (group x value-slot) float8
 *                               accumulator array in shared memory, laid out [group][slot][thread]
 *                               so a warp's 32 read-modify-writes hit 32 different banks:
 *                               3 instructions per value, no atomics, no shuffles
 *                    MODE_TR  : "lane owns (group, slot)": the warp's 32 values are transposed
 *                               through shared memory and folded by the owning lane into
 *                               registers (<= 32 groups, <= 128 pairs; min/max/int sums, NULLs)
 *                    Reduction trees are fixed => sums are deterministic run to run.
 *   epilogue         threads -> one record per group and block -> global; a single-block merge
 *                    kernel folds records with equal keys, again in a fixed order.
 *
 * HBM-bound by construction: algorithmic bytes = nblocks * 32768, each read exactly once.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "gg_device.cuh"
#include "gg_engine.h"

using namespace ggd;

#define GG_NROUNDS (GGP_MAX_PAIRS / 32)

enum { MODE_PRIV = 0, MODE_TR = 1, MODE_TRN = 2 };

struct ScanAggParams {
	const uint8_t *pages;
	uint64_t nblocks;
	ggp_grec *block_recs;                 /* [gridDim.x][GGP_FAST_GROUPS] */
	uint32_t *errflags;
	unsigned long long *counters;         /* [0] rows scanned (visible), [1] rows passed */
	int nstage;
	int gcap;                             /* groups this variant holds per block */
	int scratch_per_warp;                 /* bytes: column offsets (+ TR: transposed values, group ids, null masks) */
	uint32_t scratch_off;                 /* byte offsets from the start of dynamic shared memory */
	uint32_t cnt_off, acc_off;            /* MODE_PRIV: per-thread row counts [gcap][NT] u32, sums [gcap][nslots][NT] f64 */
};

struct BlockTable {                       /* per-block group table in shared memory */
	uint64_t key[GGP_FAST_GROUPS][GG_MAX_KEYS];
	uint32_t keynull[GGP_FAST_GROUPS];
	volatile int n;
	int lock;
};

__device__ __forceinline__ bool key_eq(const BlockTable *T, int i, const uint64_t *k, uint32_t knull, int nkeys)
{
	if (T->keynull[i] != knull) return false;
	for (int c = 0; c < nkeys; c++)
		if (T->key[i][c] != k[c]) return false;
	return true;
}

/* find the group of each lane's key, inserting new groups under a block-level lock
 * (lookup_agg_hash_entry, execHHashagg.c:456: NULL keys compare equal to each other) */
__device__ __forceinline__ int find_or_insert(BlockTable *T, const uint64_t *k, uint32_t knull, int nkeys,
                                              int gcap, bool want, int lane, uint32_t &err)
{
	int gid = -1;
	bool need = want;
	if (need)
	{
		int n = T->n;
		for (int i = 0; i < n; i++)
			if (key_eq(T, i, k, knull, nkeys)) { gid = i; break; }
		need = gid < 0;
	}
	unsigned m = __ballot_sync(GG_FULL_MASK, need);
	while (m)
	{
		int leader = __ffs(m) - 1;
		if (lane == leader)
		{
			while (atomicCAS(&T->lock, 0, 1) != 0) { }
			__threadfence_block();
			int n = T->n, found = -1;
			for (int i = 0; i < n; i++)
				if (key_eq(T, i, k, knull, nkeys)) { found = i; break; }
			if (found < 0)
			{
				if (n < gcap)
				{
					for (int c = 0; c < GG_MAX_KEYS; c++) T->key[n][c] = c < nkeys ? k[c] : 0;
					T->keynull[n] = knull;
					__threadfence_block();
					T->n = n + 1;
					found = n;
				}
				else
					err |= GGP_EF_GROUP_OVERFLOW;
			}
			__threadfence_block();
			atomicExch(&T->lock, 0);
			gid = found;
			need = false;
		}
		__syncwarp();
		if (need)
		{
			int n = T->n;
			for (int i = 0; i < n; i++)
				if (key_eq(T, i, k, knull, nkeys)) { gid = i; break; }
			need = gid < 0;
		}
		m = __ballot_sync(GG_FULL_MASK, need);
	}
	return gid;
}

/* post-action sink shared by the kernel variants */
template <int MODE>
struct RowSink {
	const ggp_program *P;
	BlockTable *T;
	uint64_t k0, k1, k2, k3;
	uint32_t knull;
	int nkeys, gcap, lane, gid;
	uint32_t *err;
	unsigned long long npassed;
	bool nonfinite;
	/* MODE_PRIV */
	uint32_t acc_thread;         /* shared address of this thread's slot-0/group-0 accumulator */
	uint32_t cnt_thread;
	uint32_t gstride, sstride, cstride;
	/* MODE_TR */
	uint32_t sv;                 /* shared address of the warp's transposed values [slot][33] f64 */
	uint32_t vnull;

	__device__ __forceinline__ void begin_row()
	{
		k0 = k1 = k2 = k3 = 0; knull = 0; gid = -1; vnull = 0;
	}
	__device__ __forceinline__ bool filter(bool pass) { return pass; }
	__device__ __forceinline__ void key(int kc, uint64_t v, bool isnull)
	{
		if (isnull) { knull |= 1u << kc; return; }
		v = normalize_key(v, P->keytype[kc]);
		if (kc == 0) k0 = v; else if (kc == 1) k1 = v; else if (kc == 2) k2 = v; else k3 = v;
	}
	__device__ __forceinline__ bool group(bool live)
	{
		if (nkeys == 0) gid = live ? 0 : -1;
		else
		{
			uint64_t k[GG_MAX_KEYS] = { k0, k1, k2, k3 };
			gid = find_or_insert(T, k, knull, nkeys, gcap, live, lane, *err);
		}
		if (live) npassed++;
		if (MODE == MODE_PRIV && gid >= 0)
		{
			uint32_t a = cnt_thread + (uint32_t) gid * cstride;
			sts32(a, lds32(a) + 1);
		}
		return gid >= 0;
	}
	__device__ __forceinline__ void out(int slot, double v, bool isnull)
	{
		if (MODE == MODE_PRIV)
		{
			if (gid >= 0)
			{
				uint32_t a = acc_thread + (uint32_t) gid * gstride + (uint32_t) slot * sstride;
				stsf64(a, __dadd_rn(ldsf64(a), v));
				nonfinite |= !f8_finite(v);
			}
		}
		else
		{
			stsf64(sv + (uint32_t) (slot * 33 + lane) * 8, v);
			if (isnull) vnull |= 1u << slot;
			else if (gid >= 0) nonfinite |= !f8_finite(v);
		}
	}
};

template <int MODE>
__global__ void __launch_bounds__(MODE == MODE_PRIV ? 480 : 256, MODE == MODE_PRIV ? 1 : 2)
gg_scanagg_kernel(const __grid_constant__ ggp_program P, const ScanAggParams prm)
{
	constexpr bool NULLABLE = (MODE == MODE_TRN);
	extern __shared__ __align__(128) uint8_t smem[];
	const int nstage = prm.nstage;
	const int ncons = (blockDim.x >> 5) - 1;          /* consumer warps; the last warp is the producer */
	const int NT = ncons * 32;                         /* consumer threads */
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	const uint32_t smem_base = smem_u32(smem);
	const uint32_t ring = smem_base;
	const uint32_t full_bar = ring + (uint32_t) nstage * GG_BLCKSZ;
	const uint32_t empty_bar = full_bar + (uint32_t) nstage * 8;
	BlockTable *T = (BlockTable *) (smem + (size_t) nstage * GG_BLCKSZ + (size_t) nstage * 16);

	const int nkeys = P.nkeys;
	const int nslots = P.nslots;
	const int V = nslots > 0 ? nslots : 1;
	const int gcap = prm.gcap;

	if (threadIdx.x == 0)
	{
		for (int s = 0; s < nstage; s++)
		{
			mbar_init(full_bar + s * 8, 1);
			mbar_init(empty_bar + s * 8, ncons);
		}
		T->n = (nkeys == 0) ? 1 : 0;               /* plain aggregation: the single group always exists */
		T->lock = 0;
		if (nkeys == 0) { T->keynull[0] = 0; for (int c = 0; c < GG_MAX_KEYS; c++) T->key[0][c] = 0; }
		mbar_fence_init();
	}
	if (MODE == MODE_PRIV && warp < ncons)
	{
		/* zero this thread's private accumulators */
		for (int g = 0; g < gcap; g++)
		{
			sts32(smem_base + prm.cnt_off + (uint32_t) (g * NT + (int) threadIdx.x) * 4, 0);
			for (int s = 0; s < nslots; s++)
				sts64(smem_base + prm.acc_off + (uint32_t) ((g * nslots + s) * NT + (int) threadIdx.x) * 8, 0);
		}
	}
	__syncthreads();

	/* pages of this block: blockIdx.x, +gridDim.x, ... */
	const uint64_t first = blockIdx.x, stride = gridDim.x;
	const uint64_t npages = first < prm.nblocks ? (prm.nblocks - first + stride - 1) / stride : 0;

	/* MODE_TR accumulators: round r of this lane owns pair p = r*32+lane -> (g = p / V, slot = p % V) */
	double acc_sum[MODE == MODE_PRIV ? 1 : GG_NROUNDS];
	uint32_t acc_cnt[MODE == MODE_PRIV ? 1 : GG_NROUNDS], acc_n[MODE == MODE_PRIV ? 1 : GG_NROUNDS];
	int pair_g[MODE == MODE_PRIV ? 1 : GG_NROUNDS], pair_j[MODE == MODE_PRIV ? 1 : GG_NROUNDS], pair_kind[MODE == MODE_PRIV ? 1 : GG_NROUNDS];
	if (MODE != MODE_PRIV)
	{
#pragma unroll
		for (int r = 0; r < GG_NROUNDS; r++)
		{
			int p = r * 32 + lane;
			pair_g[r] = p / V;
			pair_j[r] = p % V;
			pair_kind[r] = P.nacc == 0 ? GGP_ACC_COUNT : (pair_j[r] < P.nacc ? P.acckind[pair_j[r]] : GGP_ACC_F8SUM);
			acc_sum[r] = 0.0; acc_cnt[r] = 0; acc_n[r] = 0;
		}
	}
	uint32_t err = 0;
	unsigned long long n_scanned = 0, n_passed = 0;

	if (warp == ncons)
	{
		/* ===== producer: one elected lane streams pages through the ring ===== */
		if (lane == 0)
		{
			for (uint64_t it = 0; it < npages; it++)
			{
				int s = (int) (it % nstage);
				uint32_t ph = (uint32_t) ((it / nstage) & 1);
				mbar_wait(empty_bar + s * 8, ph ^ 1, 256);
				mbar_arrive_expect_tx(full_bar + s * 8, GG_BLCKSZ);
				tma_load_1d(ring + (uint32_t) s * GG_BLCKSZ,
				            prm.pages + (first + it * stride) * (uint64_t) GG_BLCKSZ, GG_BLCKSZ, full_bar + s * 8);
			}
		}
	}
	else
	{
		/* ===== consumers ===== */
		const uint32_t myscr = smem_base + prm.scratch_off + (uint32_t) warp * prm.scratch_per_warp;
		const uint32_t offs = myscr;                                               /* [ncols][32] u16 */
		const uint32_t sv = myscr + ((P.outer.ncols * 64 + 15) & ~15);             /* TR: [V][33] f64 */
		const uint32_t sg = sv + (uint32_t) V * 33 * 8;                            /* TR: [32] i32 group of each tuple, -1 = none */
		const uint32_t snull = sg + 128;                                           /* TR: [32] u32 bit s: value slot s is NULL */

		EvalCtx X;
		X.P = &P; X.offs = offs; X.IS = nullptr; X.ioffs = 0; X.ifast = false; X.lane = lane;
		X.itv.tp = 0; X.itv.colnull = 0;
		RowSink<MODE> sink;
		sink.P = &P; sink.T = T; sink.nkeys = nkeys; sink.gcap = gcap; sink.lane = lane; sink.err = &err;
		sink.npassed = 0; sink.nonfinite = false;
		sink.acc_thread = smem_base + prm.acc_off + threadIdx.x * 8;
		sink.cnt_thread = smem_base + prm.cnt_off + threadIdx.x * 4;
		sink.sstride = (uint32_t) NT * 8;
		sink.gstride = (uint32_t) nslots * NT * 8;
		sink.cstride = (uint32_t) NT * 4;
		sink.sv = sv;

		for (uint64_t it = 0; it < npages; it++)
		{
			int s = (int) (it % nstage);
			uint32_t ph = (uint32_t) ((it / nstage) & 1);
			if (lane == 0) mbar_wait(full_bar + s * 8, ph, 32);
			__syncwarp();
			const uint32_t pg = ring + (uint32_t) s * GG_BLCKSZ;

			/* page header, bufpage.h:153-166; sanity rules of PageAddItem (bufpage.c:196-204) */
			const uint32_t w2 = lds32(pg + 8), w3 = lds32(pg + 12), w4 = lds32(pg + 16);
			const uint32_t pd_flags = w2 >> 16, pd_lower = w3 & 0xFFFF, pd_upper = w3 >> 16, pd_special = w4 & 0xFFFF;
			int nitems = 0;
			if (pd_lower < GG_PAGE_HEADER_SIZE || pd_lower > pd_upper || pd_upper > pd_special || pd_special > GG_BLCKSZ)
			{
				/* an all-zero page is a valid empty page (PageIsNew) */
				if (pd_upper != 0 || pd_lower != 0) err |= GGP_EF_BADPAGE;
			}
			else
				nitems = (int) ((pd_lower - GG_PAGE_HEADER_SIZE) >> 2);
			const bool all_visible = (pd_flags & GG_PD_ALL_VISIBLE) != 0;     /* heapam.c:391 */
			const int nchunks = (nitems + 31) >> 5;

			for (int c = (int) ((warp + it) % ncons); c < nchunks; c += ncons)
			{
				const int idx = c * 32 + lane;
				bool live = false;
				uint32_t tup = pg, tuplen = 64;
				if (idx < nitems)
				{
					/* ItemIdData: lp_off:15 | lp_flags:2 | lp_len:15 (itemid.h:24-29) */
					const uint32_t lp = lds32(pg + GG_PAGE_HEADER_SIZE + idx * 4);
					const uint32_t lp_off = lp & 0x7FFF, lp_flags = (lp >> 15) & 3, lp_len = lp >> 17;
					if (lp_flags == GG_LP_NORMAL)
					{
						if (lp_off < pd_upper || lp_off + lp_len > pd_special || lp_len < GG_HEAP_HDR_SIZE + 1 || (lp_off & 7))
							err |= GGP_EF_BADPAGE;
						else
						{
							tup = pg + lp_off;
							tuplen = lp_len;
							live = true;
						}
					}
				}
				/* t_infomask2 | t_infomask | t_hoff live in bytes 18..22 of the header (htup_details.h:139-162) */
				const uint32_t hw = lds32(tup + 20);                 /* infomask (lo 16) | t_hoff (byte 2) */
				const uint32_t infomask = hw & 0xFFFF, hoff = (hw >> 16) & 0xFF;
				if (live && !all_visible)
				{
					/* HeapTupleSatisfiesMVCC fast path (tqual.c:1009,1119): frozen xmin + invalid xmax */
					if ((infomask & GG_HEAP_XMIN_FROZEN) == GG_HEAP_XMIN_FROZEN && (infomask & GG_HEAP_XMAX_INVALID)) { }
					else if ((infomask & GG_HEAP_XMIN_INVALID) && !(infomask & GG_HEAP_XMIN_COMMITTED)) live = false;
					else { err |= GGP_EF_VISIBILITY; live = false; }
				}
				if (live && (hoff > tuplen || (hoff & 7) || hoff < 24)) { err |= GGP_EF_BADPAGE; live = false; }

				/* dead lanes get a harmless view (the page header) so that the warp-uniform program can run */
				const bool hasnulls = live && (infomask & GG_HEAP_HASNULL);
				const bool fast = !__any_sync(GG_FULL_MASK, hasnulls);
				X.fast = fast;
				X.tv.tp = pg;
				X.tv.colnull = 0;
				if (live)
				{
					n_scanned++;
					const uint32_t e0 = err;
					walk_tuple(P.outer, tup, tuplen, fast, offs, lane, X.tv, err);
					if (err != e0 && (err & GGP_EF_BADPAGE)) live = false;
					if (!NULLABLE && X.tv.colnull) { err |= GGP_EF_NOTNULL_VIOLATED; live = false; }
				}
				if (!live)
				{
					/* offsets of walked columns must be readable: point them at offset 0 */
					X.tv.tp = pg;
					for (int sl = 0; sl < P.outer.ncols; sl++) sts16(offs + (uint32_t) (sl * 32 + lane) * 2, 0);
					X.fast = false;
				}
				/* X.fast must be warp-uniform only in the sense that every lane reads valid memory:
				 * a dead lane with fast=false reads offset 0 of the page, which is always mapped */

				sink.begin_row();
				run_prog<NULLABLE, false>(X, live, err, sink);

				if (MODE != MODE_PRIV)
				{
					sts32(sg + lane * 4, (uint32_t) sink.gid);
					if (NULLABLE) sts32(snull + lane * 4, sink.vnull);
					__syncwarp();

					/* ---- lane-owns-(group, slot) accumulate ---- */
					const int G = T->n;
#pragma unroll
					for (int r = 0; r < GG_NROUNDS; r++)
					{
						if (r * 32 < G * V)              /* warp-uniform */
						{
							const int g = pair_g[r], j = pair_j[r], kind = pair_kind[r];
							double s0 = acc_sum[r];
							uint32_t cnt = acc_cnt[r], nn = acc_n[r];
#pragma unroll 8
							for (int i = 0; i < 32; i++)
							{
								if ((int) lds32(sg + i * 4) == g)
								{
									cnt++;
									bool isn = NULLABLE && ((lds32(snull + i * 4) >> j) & 1);
									if (!isn && P.nacc > 0)
									{
										double v = ldsf64(sv + (uint32_t) (j * 33 + i) * 8);
										if (kind == GGP_ACC_F8SUM) s0 = __dadd_rn(s0, v);
										else if (kind == GGP_ACC_I8SUM)
											s0 = __longlong_as_double(__double_as_longlong(s0) + __double_as_longlong(v));
										else if (kind == GGP_ACC_F8MIN) { if (nn == 0 || f8_cmp(v, s0) < 0) s0 = v; }
										else if (kind == GGP_ACC_F8MAX) { if (nn == 0 || f8_cmp(v, s0) > 0) s0 = v; }
										else if (kind == GGP_ACC_I8MIN) { if (nn == 0 || __double_as_longlong(v) < __double_as_longlong(s0)) s0 = v; }
										else if (kind == GGP_ACC_I8MAX) { if (nn == 0 || __double_as_longlong(v) > __double_as_longlong(s0)) s0 = v; }
										nn++;
									}
								}
							}
							acc_sum[r] = s0; acc_cnt[r] = cnt; acc_n[r] = nn;
						}
					}
					__syncwarp();
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(empty_bar + s * 8);
		}
		n_passed = sink.npassed;
		if (sink.nonfinite) err |= GGP_EF_SAW_INF;     /* an infinite/NaN input legitimises an infinite sum */
	}

	/* ===== epilogue: threads -> block records, all in fixed order ===== */
	__syncthreads();
	if (warp < ncons)
	{
		/* per-warp counters and error bits */
		for (int o = 16; o > 0; o >>= 1)
		{
			n_scanned += __shfl_xor_sync(GG_FULL_MASK, n_scanned, o);
			n_passed += __shfl_xor_sync(GG_FULL_MASK, n_passed, o);
			err |= __shfl_xor_sync(GG_FULL_MASK, err, o);
		}
		if (lane == 0)
		{
			if (n_scanned) atomicAdd(&prm.counters[0], n_scanned);
			if (n_passed) atomicAdd(&prm.counters[1], n_passed);
			if (err) atomicOr(prm.errflags, err);
		}
	}
	const int G = T->n;
	ggp_grec *out = prm.block_recs + (size_t) blockIdx.x * GGP_FAST_GROUPS;
	for (int g = G + (int) threadIdx.x; g < GGP_FAST_GROUPS; g += blockDim.x) out[g].valid = 0;

	if (MODE == MODE_PRIV)
	{
		/* one warp per (group, slot): lane l folds threads l, l+32, ... in order, then a fixed butterfly */
		for (int e = warp; e < G * V; e += (int) (blockDim.x >> 5))
		{
			const int g = e / V, sl = e % V;
			double s0 = 0.0;
			unsigned long long cnt = 0;
			for (int t = lane; t < NT; t += 32)
			{
				if (nslots > 0) s0 = __dadd_rn(s0, ldsf64(smem_base + prm.acc_off + (uint32_t) ((g * nslots + sl) * NT + t) * 8));
				if (sl == 0) cnt += lds32(smem_base + prm.cnt_off + (uint32_t) (g * NT + t) * 4);
			}
			for (int o = 16; o > 0; o >>= 1)
			{
				s0 = __dadd_rn(s0, __shfl_xor_sync(GG_FULL_MASK, s0, o));
				cnt += __shfl_xor_sync(GG_FULL_MASK, cnt, o);
			}
			if (lane == 0)
			{
				if (nslots > 0)
				{
					if (sl < P.nacc) { out[g].sum[sl] = s0; if (P.accsq[sl] < 0) out[g].sumsq[sl] = 0.0; }
					else
						for (int j = 0; j < P.nacc; j++)
							if (P.accsq[j] == sl) out[g].sumsq[j] = s0;
				}
				if (sl == 0)
				{
					out[g].count = cnt;
					for (int j = 0; j < P.nacc; j++) out[g].n[j] = cnt;     /* NOT NULL inputs: every row counts */
					out[g].keynull = T->keynull[g];
					for (int c = 0; c < GG_MAX_KEYS; c++) out[g].key[c] = T->key[g][c];
					out[g].valid = 1;
				}
			}
		}
		return;
	}

	struct Red { double sum; unsigned long long cnt, n; };
	Red *red = (Red *) smem;                       /* [ncons][GGP_MAX_PAIRS]: 21 KB for 7 warps, inside the ring */
	if (warp < ncons)
	{
#pragma unroll
		for (int r = 0; r < GG_NROUNDS; r++)
		{
			Red x;
			x.sum = acc_sum[r]; x.cnt = acc_cnt[r]; x.n = acc_n[r];
			red[warp * GGP_MAX_PAIRS + r * 32 + lane] = x;
		}
	}
	__syncthreads();
	for (int p = threadIdx.x; p < GGP_MAX_PAIRS; p += blockDim.x)
	{
		int g = p / V, sl = p % V;
		if (g >= G || g >= GGP_FAST_GROUPS) continue;
		int kind = P.nacc == 0 ? GGP_ACC_COUNT : (sl < P.nacc ? P.acckind[sl] : GGP_ACC_F8SUM);
		double s0 = 0.0;
		unsigned long long cnt = 0, nn = 0;
		for (int w = 0; w < ncons; w++)
		{
			Red x = red[w * GGP_MAX_PAIRS + p];
			cnt += x.cnt;
			if (x.n)
			{
				if (kind == GGP_ACC_F8SUM) s0 = __dadd_rn(s0, x.sum);
				else if (kind == GGP_ACC_I8SUM) s0 = __longlong_as_double(__double_as_longlong(s0) + __double_as_longlong(x.sum));
				else if (kind == GGP_ACC_F8MIN) { if (nn == 0 || f8_cmp(x.sum, s0) < 0) s0 = x.sum; }
				else if (kind == GGP_ACC_F8MAX) { if (nn == 0 || f8_cmp(x.sum, s0) > 0) s0 = x.sum; }
				else if (kind == GGP_ACC_I8MIN) { if (nn == 0 || __double_as_longlong(x.sum) < __double_as_longlong(s0)) s0 = x.sum; }
				else if (kind == GGP_ACC_I8MAX) { if (nn == 0 || __double_as_longlong(x.sum) > __double_as_longlong(s0)) s0 = x.sum; }
				nn += x.n;
			}
		}
		if (P.nacc > 0)
		{
			if (sl < P.nacc) { out[g].sum[sl] = s0; out[g].n[sl] = nn; if (P.accsq[sl] < 0) out[g].sumsq[sl] = 0.0; }
			else
				for (int j = 0; j < P.nacc; j++)
					if (P.accsq[j] == sl) out[g].sumsq[j] = s0;
		}
		if (sl == 0)
		{
			out[g].count = cnt;
			out[g].keynull = T->keynull[g];
			for (int c = 0; c < GG_MAX_KEYS; c++) out[g].key[c] = T->key[g][c];
			out[g].valid = 1;
		}
	}
}

/* ---- merge kernel: fold group records with equal keys, in record order (deterministic) ----
 * One block.  A: ordered compaction of the valid records.  B: every valid record finds (or, under a
 * lock, creates) its merged group.  C: thread (group, column) folds that group's records in record
 * order.  This is also the combine step of a FINAL-stage Agg (float8pl / float8_combine / int8pl,
 * nodeAgg.c:2123-2148) when the records come from other segments. */
__global__ void __launch_bounds__(256, 1)
gg_merge_recs_kernel(const ggp_grec *recs, int nrecs, int nkeys, int nacc, ggp_acckinds kinds,
                     ggp_grec *out, int outcap, int *nout, int *vidx /* [nrecs] */, int *vmap /* [nrecs] */,
                     uint32_t *errflags)
{
	__shared__ int s_nvalid, s_nout, s_lock;
	__shared__ int s_warpsum[8];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) { s_nvalid = 0; s_nout = 0; s_lock = 0; }
	__syncthreads();

	/* A: ordered compaction (chunks of 256 records, in order) */
	for (int base = 0; base < nrecs; base += 256)
	{
		int i = base + tid;
		bool v = i < nrecs && recs[i].valid != 0;
		unsigned b = __ballot_sync(GG_FULL_MASK, v);
		if (lane == 0) s_warpsum[warp] = __popc(b);
		__syncthreads();
		int pre = 0;
		for (int w = 0; w < warp; w++) pre += s_warpsum[w];
		int tot = 0;
		for (int w = 0; w < 8; w++) tot += s_warpsum[w];
		int pos = s_nvalid + pre + __popc(b & ((1u << lane) - 1));
		if (v) vidx[pos] = i;
		__syncthreads();
		if (tid == 0) s_nvalid += tot;
		__syncthreads();
	}
	const int nvalid = s_nvalid;

	/* B: group assignment */
	for (int base = 0; base < nvalid; base += 256)
	{
		int k = base + tid;
		bool need = k < nvalid;
		const ggp_grec *x = need ? &recs[vidx[k]] : nullptr;
		int f = -1;
		if (need)
		{
			int n = *(volatile int *) &s_nout;
			for (int m = 0; m < n && f < 0; m++)
			{
				bool eq = out[m].keynull == x->keynull;
				for (int c = 0; eq && c < nkeys; c++) eq = out[m].key[c] == x->key[c];
				if (eq) f = m;
			}
			need = f < 0;
		}
		unsigned mm = __ballot_sync(GG_FULL_MASK, need);
		while (mm)
		{
			int leader = __ffs(mm) - 1;
			if (lane == leader)
			{
				while (atomicCAS(&s_lock, 0, 1) != 0) { }
				__threadfence();
				int n = *(volatile int *) &s_nout;
				for (int m = 0; m < n && f < 0; m++)
				{
					bool eq = ((volatile ggp_grec *) out)[m].keynull == x->keynull;
					for (int c = 0; eq && c < nkeys; c++) eq = ((volatile ggp_grec *) out)[m].key[c] == x->key[c];
					if (eq) f = m;
				}
				if (f < 0)
				{
					if (n < outcap)
					{
						for (int c = 0; c < GG_MAX_KEYS; c++) out[n].key[c] = x->key[c];
						out[n].keynull = x->keynull;
						out[n].valid = 1;
						__threadfence();
						*(volatile int *) &s_nout = n + 1;
						f = n;
					}
					else
						atomicOr(errflags, GGP_EF_GROUP_OVERFLOW);
				}
				__threadfence();
				atomicExch(&s_lock, 0);
				need = false;
			}
			__syncwarp();
			if (need)
			{
				int n = *(volatile int *) &s_nout;
				for (int m = 0; m < n && f < 0; m++)
				{
					bool eq = ((volatile ggp_grec *) out)[m].keynull == x->keynull;
					for (int c = 0; eq && c < nkeys; c++) eq = ((volatile ggp_grec *) out)[m].key[c] == x->key[c];
					if (eq) f = m;
				}
				need = f < 0;
			}
			mm = __ballot_sync(GG_FULL_MASK, need);
		}
		if (k < nvalid) vmap[k] = f;
	}
	__threadfence();
	__syncthreads();

	/* C: fold.  One warp per (merged group, column): lane l folds valid records l, l+32, ... in order,
	 * then the 32 partials are combined by a fixed butterfly => deterministic result. */
	const int n = s_nout;
	if (tid == 0) *nout = n;
	const int V = nacc > 0 ? nacc : 1;
	const bool saw_inf = (*errflags & GGP_EF_SAW_INF) != 0;
	for (int t = warp; t < n * V; t += (int) (blockDim.x >> 5))
	{
		int mg = t / V, j = t % V;
		int kind = nacc > 0 ? kinds.k[j] : GGP_ACC_COUNT;
		double s0 = 0.0, s1 = 0.0;
		unsigned long long cnt = 0, nn = 0;
		for (int k = lane; k < nvalid; k += 32)
		{
			if (vmap[k] != mg) continue;
			const ggp_grec &x = recs[vidx[k]];
			cnt += x.count;
			if (nacc > 0 && x.n[j])
			{
				if (kind == GGP_ACC_F8SUM) { s0 = __dadd_rn(s0, x.sum[j]); s1 = __dadd_rn(s1, x.sumsq[j]); }
				else if (kind == GGP_ACC_I8SUM)
				{
					long long a = __double_as_longlong(s0), b = __double_as_longlong(x.sum[j]), r = (long long) ((unsigned long long) a + (unsigned long long) b);
					if (((a ^ r) & (b ^ r)) < 0) atomicOr(errflags, GGP_EF_INT_OVERFLOW);   /* int8pl, int8.c:526 */
					s0 = __longlong_as_double(r);
				}
				else if (kind == GGP_ACC_F8MIN) { if (nn == 0 || f8_cmp(x.sum[j], s0) < 0) s0 = x.sum[j]; }
				else if (kind == GGP_ACC_F8MAX) { if (nn == 0 || f8_cmp(x.sum[j], s0) > 0) s0 = x.sum[j]; }
				else if (kind == GGP_ACC_I8MIN) { if (nn == 0 || __double_as_longlong(x.sum[j]) < __double_as_longlong(s0)) s0 = x.sum[j]; }
				else if (kind == GGP_ACC_I8MAX) { if (nn == 0 || __double_as_longlong(x.sum[j]) > __double_as_longlong(s0)) s0 = x.sum[j]; }
				nn += x.n[j];
			}
		}
		for (int o = 16; o > 0; o >>= 1)
		{
			double os0 = __shfl_xor_sync(GG_FULL_MASK, s0, o), os1 = __shfl_xor_sync(GG_FULL_MASK, s1, o);
			unsigned long long ocnt = __shfl_xor_sync(GG_FULL_MASK, cnt, o), onn = __shfl_xor_sync(GG_FULL_MASK, nn, o);
			/* both partners must compute the identical combined value: order operands by lane */
			bool lo = (lane & o) == 0;
			double a0 = lo ? s0 : os0, b0 = lo ? os0 : s0;
			unsigned long long an = lo ? nn : onn, bn = lo ? onn : nn;
			if (kind == GGP_ACC_F8SUM) { s0 = __dadd_rn(a0, b0); s1 = __dadd_rn(lo ? s1 : os1, lo ? os1 : s1); }
			else if (kind == GGP_ACC_I8SUM)
			{
				long long a = __double_as_longlong(a0), b = __double_as_longlong(b0), r = (long long) ((unsigned long long) a + (unsigned long long) b);
				if (((a ^ r) & (b ^ r)) < 0) atomicOr(errflags, GGP_EF_INT_OVERFLOW);
				s0 = __longlong_as_double(r);
			}
			else if (kind == GGP_ACC_F8MIN) s0 = an == 0 ? b0 : bn == 0 ? a0 : (f8_cmp(b0, a0) < 0 ? b0 : a0);
			else if (kind == GGP_ACC_F8MAX) s0 = an == 0 ? b0 : bn == 0 ? a0 : (f8_cmp(b0, a0) > 0 ? b0 : a0);
			else if (kind == GGP_ACC_I8MIN) s0 = an == 0 ? b0 : bn == 0 ? a0 : (__double_as_longlong(b0) < __double_as_longlong(a0) ? b0 : a0);
			else if (kind == GGP_ACC_I8MAX) s0 = an == 0 ? b0 : bn == 0 ? a0 : (__double_as_longlong(b0) > __double_as_longlong(a0) ? b0 : a0);
			cnt += ocnt;
			nn += onn;
		}
		if (lane == 0)
		{
			if (nacc > 0)
			{
				/* CHECKFLOATVAL of float8pl / float8_accum / float8_combine (float.c:782,1842,1878): a sum that
				 * became infinite although no input was infinite is an overflow ERROR */
				if (kind == GGP_ACC_F8SUM && !saw_inf && (!f8_finite(s0) || !f8_finite(s1)))
					atomicOr(errflags, GGP_EF_FLOAT_OVERFLOW);
				out[mg].sum[j] = s0; out[mg].sumsq[j] = s1; out[mg].n[j] = nn;
			}
			if (j == 0) out[mg].count = cnt;
		}
	}
}

/* =====================================================================================
 * host side: the pipeline object behind gg_scanagg_* (include/ggb200.h)
 * ===================================================================================== */
#include <cstring>
#include <vector>

#define GG_MERGE_CAP 1024          /* merged groups the fast path holds per segment */
#define GG_STREAM_CHUNK_BLOCKS 8192 /* 256 MB staging chunks for gg_scanagg_run_host */

struct gg_scanagg {
	gg_engine *eng = nullptr;
	gg_scan scan;
	gg_agg agg;
	gg_exprpool pool;
	ggp_program prog;
	ggp_aggmap aggmap[GG_MAX_AGGS];
	int grid = 0, threads = 0, nstage = 0, scratch_per_warp = 0;
	int mode = MODE_PRIV;           /* kernel variant; escalates PRIV -> TR when a run overflows its group capacity */
	int ctas_per_sm = 2, gcap = 0;
	uint32_t scratch_off = 0, cnt_off = 0, acc_off = 0;
	size_t smem = 0;
	/* device state */
	ggp_grec *recs = nullptr;       /* [GG_MERGE_CAP (previous merged)] ++ [grid * GGP_FAST_GROUPS (block records)] */
	ggp_grec *merged = nullptr;     /* [GG_MERGE_CAP] output of the merge kernel */
	int *vidx = nullptr, *vmap = nullptr, *d_nout = nullptr;
	uint32_t *d_err = nullptr;
	unsigned long long *d_counters = nullptr;
	int nrecs_total = 0, nrecs_cap = 0;
	/* inputs of the current accumulation, kept so that a group-capacity overflow can be replayed on a wider variant */
	struct Fed { const uint8_t *dev; const void *host; uint64_t nblocks; };
	std::vector<Fed> fed;
	bool has_state = false;
	/* host staging for the streamed path */
	uint8_t *stage[2] = { nullptr, nullptr };
	cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_consumed[2] = { nullptr, nullptr };
};

/* launch configuration and shared-memory layout for the current kernel variant:
 *   ring[nstage][32 KB] | full/empty mbarriers | BlockTable | per-warp scratch | (PRIV) counts | (PRIV) sums */
static int scanagg_configure(gg_scanagg *p)
{
	gg_engine *e = p->eng;
	const int nslots = p->prog.nslots;
	const int V = nslots > 0 ? nslots : 1;
	int scr = (p->prog.outer.ncols * 64 + 15) & ~15;             /* column offsets [ncols][32] u16 */
	if (p->mode != MODE_PRIV) scr += V * 33 * 8 + 128 + 128;      /* + transposed values, group ids, null masks */
	p->scratch_per_warp = (scr + 15) & ~15;
	p->nstage = 3;
	if (p->mode == MODE_PRIV)
	{
		/* 1 CTA/SM: 14 consumer warps + producer.  What the ring and the scratch leave of the 227 KB goes to
		 * the per-thread private accumulators; that fixes how many groups this variant holds. */
		p->ctas_per_sm = 1;
		p->threads = 15 * 32;
		const int ncons = 14, NT = ncons * 32;
		size_t fixed = (size_t) p->nstage * GG_BLCKSZ + (size_t) p->nstage * 16 + sizeof(BlockTable) + 16 +
		               (size_t) ncons * p->scratch_per_warp + 16;
		if (fixed + (size_t) NT * (8 * nslots + 4) > e->smem_optin) { gg_set_error("plan needs too much shared memory"); return GG_ERR_UNSUPPORTED; }
		int gcap = (int) ((e->smem_optin - fixed) / ((size_t) NT * (8 * nslots + 4)));
		if (gcap > GGP_FAST_GROUPS) gcap = GGP_FAST_GROUPS;
		p->gcap = gcap;
		p->scratch_off = (uint32_t) (((size_t) p->nstage * GG_BLCKSZ + (size_t) p->nstage * 16 + sizeof(BlockTable) + 15) & ~(size_t) 15);
		p->cnt_off = p->scratch_off + (uint32_t) ncons * p->scratch_per_warp;
		p->acc_off = (p->cnt_off + (uint32_t) gcap * NT * 4 + 15) & ~15u;
		p->smem = p->acc_off + (size_t) gcap * nslots * NT * 8;
	}
	else
	{
		/* 2 CTAs/SM x (7 consumer warps + producer = 8 warps) */
		p->ctas_per_sm = 2;
		p->threads = 8 * 32;
		const int ncons = 7;
		p->gcap = GGP_MAX_PAIRS / V < GGP_FAST_GROUPS ? GGP_MAX_PAIRS / V : GGP_FAST_GROUPS;
		p->scratch_off = (uint32_t) (((size_t) p->nstage * GG_BLCKSZ + (size_t) p->nstage * 16 + sizeof(BlockTable) + 15) & ~(size_t) 15);
		p->cnt_off = p->acc_off = 0;
		const size_t per_cta_2 = (e->smem_optin + 1024) / 2 - 1024;   /* 1 KB reserved per CTA */
		p->smem = p->scratch_off + (size_t) ncons * p->scratch_per_warp;
		if (p->smem > per_cta_2)
		{
			p->nstage = 2;
			p->scratch_off = (uint32_t) (((size_t) p->nstage * GG_BLCKSZ + (size_t) p->nstage * 16 + sizeof(BlockTable) + 15) & ~(size_t) 15);
			p->smem = p->scratch_off + (size_t) ncons * p->scratch_per_warp;
		}
		if (p->smem > per_cta_2) { gg_set_error("plan needs too much shared memory"); return GG_ERR_UNSUPPORTED; }
		if (p->smem < 32 * 1024) p->smem = 32 * 1024;             /* the epilogue reuses the ring as reduction scratch */
	}
	p->grid = e->sm_count * p->ctas_per_sm;
	if (p->mode == MODE_PRIV)
		GG_CUDA(cudaFuncSetAttribute(gg_scanagg_kernel<MODE_PRIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p->smem));
	else if (p->mode == MODE_TR)
		GG_CUDA(cudaFuncSetAttribute(gg_scanagg_kernel<MODE_TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p->smem));
	else
		GG_CUDA(cudaFuncSetAttribute(gg_scanagg_kernel<MODE_TRN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p->smem));
	return GG_OK;
}

static int scanagg_launch(gg_scanagg *p, const uint8_t *dev_pages, uint64_t nblocks, cudaStream_t st)
{
	gg_engine *e = p->eng;
	ScanAggParams prm;
	prm.pages = dev_pages;
	prm.nblocks = nblocks;
	prm.block_recs = p->recs + GG_MERGE_CAP;
	prm.errflags = p->d_err;
	prm.counters = p->d_counters;
	prm.nstage = p->nstage;
	prm.gcap = p->gcap;
	prm.scratch_per_warp = p->scratch_per_warp;
	prm.scratch_off = p->scratch_off;
	prm.cnt_off = p->cnt_off;
	prm.acc_off = p->acc_off;
	if (p->mode == MODE_PRIV)
		gg_scanagg_kernel<MODE_PRIV><<<p->grid, p->threads, p->smem, st>>>(p->prog, prm);
	else if (p->mode == MODE_TR)
		gg_scanagg_kernel<MODE_TR><<<p->grid, p->threads, p->smem, st>>>(p->prog, prm);
	else
		gg_scanagg_kernel<MODE_TRN><<<p->grid, p->threads, p->smem, st>>>(p->prog, prm);
	GG_CUDA(cudaGetLastError());
	e->launches++;
	/* fold the block records (and the previously merged groups) */
	ggp_acckinds kinds;
	memcpy(kinds.k, p->prog.acckind, sizeof kinds.k);
	gg_merge_recs_kernel<<<1, 256, 0, st>>>(p->recs, p->nrecs_total, p->prog.nkeys, p->prog.nacc, kinds,
	                                        p->merged, GG_MERGE_CAP, p->d_nout, p->vidx, p->vmap, p->d_err);
	GG_CUDA(cudaGetLastError());
	e->launches++;
	/* merged -> first GG_MERGE_CAP slots of recs (input of the next fold); slots beyond nout are invalidated
	 * by copying the whole (zero-initialised) merged array */
	GG_CUDA(cudaMemcpyAsync(p->recs, p->merged, sizeof(ggp_grec) * GG_MERGE_CAP, cudaMemcpyDeviceToDevice, st));
	GG_CUDA(cudaMemsetAsync(p->merged, 0, sizeof(ggp_grec) * GG_MERGE_CAP, st));
	p->has_state = true;
	return GG_OK;
}

extern "C" {

int gg_scanagg_create(gg_engine *e, const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool,
                      gg_scanagg **out)
{
	if (!e || !scan || !agg || !pool || !out) return GG_ERR_ARG;
	*out = nullptr;
	GG_CUDA(cudaSetDevice(e->device));
	gg_scanagg *p = new gg_scanagg();
	p->eng = e;
	p->scan = *scan;
	p->agg = *agg;
	p->pool = *pool;
	char msg[256];
	int rc = ggp_compile_scanagg(scan, agg, pool, &p->prog, p->aggmap, msg, sizeof msg);
	if (rc != GG_OK) { gg_set_error("%s", msg); delete p; return rc; }

	/* kernel variant: private accumulators when the plan allows it (NOT NULL float8 sums); the planner's
	 * group estimate decides whether they can hold the groups */
	p->mode = p->prog.nullable ? MODE_TRN : MODE_TR;
	if (p->prog.priv_ok) p->mode = MODE_PRIV;
	{
		const char *force = getenv("GGB200_SCAN_MODE");       /* experiments: 0 PRIV, 1 TR, 2 TRN */
		if (force) { int m = atoi(force); if (m == MODE_PRIV && !p->prog.priv_ok) m = MODE_TR; if (p->prog.nullable) m = MODE_TRN; p->mode = m; }
	}
	rc = scanagg_configure(p);
	if (rc == GG_OK && p->mode == MODE_PRIV && agg->numGroups > p->gcap)
	{
		p->mode = MODE_TR;
		rc = scanagg_configure(p);
	}
	if (rc) { delete p; return rc; }
	/* block records are sized for the largest grid either configuration uses */
	p->nrecs_cap = GG_MERGE_CAP + e->sm_count * 4 * GGP_FAST_GROUPS;
	p->nrecs_total = GG_MERGE_CAP + p->grid * GGP_FAST_GROUPS;
	GG_CUDA(cudaMalloc((void **) &p->recs, sizeof(ggp_grec) * p->nrecs_cap));
	GG_CUDA(cudaMalloc((void **) &p->merged, sizeof(ggp_grec) * GG_MERGE_CAP));
	GG_CUDA(cudaMalloc((void **) &p->vidx, sizeof(int) * p->nrecs_cap));
	GG_CUDA(cudaMalloc((void **) &p->vmap, sizeof(int) * p->nrecs_cap));
	GG_CUDA(cudaMalloc((void **) &p->d_nout, sizeof(int)));
	GG_CUDA(cudaMalloc((void **) &p->d_err, sizeof(uint32_t)));
	GG_CUDA(cudaMalloc((void **) &p->d_counters, 2 * sizeof(unsigned long long)));
	*out = p;
	return gg_scanagg_reset(p);
}

int gg_scanagg_reset(gg_scanagg *p)
{
	if (!p) return GG_ERR_ARG;
	cudaStream_t st = p->eng->stream;
	GG_CUDA(cudaSetDevice(p->eng->device));
	GG_CUDA(cudaMemsetAsync(p->recs, 0, sizeof(ggp_grec) * p->nrecs_cap, st));
	p->fed.clear();
	GG_CUDA(cudaMemsetAsync(p->merged, 0, sizeof(ggp_grec) * GG_MERGE_CAP, st));
	GG_CUDA(cudaMemsetAsync(p->d_nout, 0, sizeof(int), st));
	GG_CUDA(cudaMemsetAsync(p->d_err, 0, sizeof(uint32_t), st));
	GG_CUDA(cudaMemsetAsync(p->d_counters, 0, 2 * sizeof(unsigned long long), st));
	p->has_state = false;
	return GG_OK;
}

int gg_scanagg_run(gg_scanagg *p, gg_relation *r, uint64_t first_block, uint64_t nblocks)
{
	if (!p || !r || first_block + nblocks > r->nblocks) return GG_ERR_ARG;
	gg_engine *e = p->eng;
	GG_CUDA(cudaSetDevice(e->device));
	GG_CUDA(cudaEventRecord(e->ev_start, e->stream));
	int rc = scanagg_launch(p, r->pages + first_block * GG_BLCKSZ, nblocks, e->stream);
	if (rc) return rc;
	p->fed.push_back({ r->pages + first_block * GG_BLCKSZ, nullptr, nblocks });
	GG_CUDA(cudaEventRecord(e->ev_stop, e->stream));
	e->timed = true;
	return GG_OK;
}

/* Streamed end-to-end path: pages live in HOST memory (the segment's shared buffers / file cache).
 * Double-buffered 256 MB chunks: H2D on the copy stream overlaps the scan kernel of the previous
 * chunk on the compute stream. */
static int scanagg_stream_host(gg_scanagg *p, const void *host_pages, uint64_t nblocks);

int gg_scanagg_run_host(gg_scanagg *p, const void *host_pages, uint64_t nblocks)
{
	if (!p || (!host_pages && nblocks)) return GG_ERR_ARG;
	int rc = scanagg_stream_host(p, host_pages, nblocks);
	if (rc == GG_OK) p->fed.push_back({ nullptr, host_pages, nblocks });
	return rc;
}

static int scanagg_stream_host(gg_scanagg *p, const void *host_pages, uint64_t nblocks)
{
	gg_engine *e = p->eng;
	GG_CUDA(cudaSetDevice(e->device));
	const uint64_t chunk = GG_STREAM_CHUNK_BLOCKS;
	for (int b = 0; b < 2; b++)
	{
		if (!p->stage[b])
		{
			GG_CUDA(cudaMalloc((void **) &p->stage[b], (size_t) chunk * GG_BLCKSZ));
			GG_CUDA(cudaEventCreateWithFlags(&p->ev_copied[b], cudaEventDisableTiming));
			GG_CUDA(cudaEventCreateWithFlags(&p->ev_consumed[b], cudaEventDisableTiming));
		}
	}
	GG_CUDA(cudaEventRecord(e->ev_start, e->stream));
	/* the copy stream must not overtake earlier work on the compute stream that used the staging buffers */
	GG_CUDA(cudaEventRecord(p->ev_consumed[0], e->stream));
	GG_CUDA(cudaEventRecord(p->ev_consumed[1], e->stream));
	uint64_t done = 0;
	int i = 0;
	while (done < nblocks)
	{
		uint64_t n = nblocks - done < chunk ? nblocks - done : chunk;
		int b = i & 1;
		GG_CUDA(cudaStreamWaitEvent(e->copy_stream, p->ev_consumed[b], 0));
		GG_CUDA(cudaMemcpyAsync(p->stage[b], (const uint8_t *) host_pages + done * GG_BLCKSZ, (size_t) n * GG_BLCKSZ,
		                        cudaMemcpyHostToDevice, e->copy_stream));
		GG_CUDA(cudaEventRecord(p->ev_copied[b], e->copy_stream));
		GG_CUDA(cudaStreamWaitEvent(e->stream, p->ev_copied[b], 0));
		int rc = scanagg_launch(p, p->stage[b], n, e->stream);
		if (rc) return rc;
		GG_CUDA(cudaEventRecord(p->ev_consumed[b], e->stream));
		done += n;
		i++;
	}
	GG_CUDA(cudaEventRecord(e->ev_stop, e->stream));
	e->timed = true;
	return GG_OK;
}

/* finalize_aggregate (nodeAgg.c:871-999) over the merged group records: O(groups) scalar work */
static void finalize_rows(const gg_agg *agg, const ggp_aggmap *aggmap, const ggp_program *prog, int final_stage,
                          const ggp_grec *recs, int n, gg_aggrow *out)
{
	for (int g = 0; g < n; g++)
	{
		const ggp_grec &x = recs[g];
		gg_aggrow &row = out[g];
		memset(&row, 0, sizeof row);
		for (int c = 0; c < agg->numCols; c++)
		{
			row.keyisnull[c] = (x.keynull >> c) & 1;
			row.key[c] = (int64_t) x.key[c];
			if (prog->keytype[c] == 3 && !row.keyisnull[c])
			{
				int len = 0;
				while (len < 8 && ((x.key[c] >> (8 * len)) & 0xff)) len++;
				row.keylen[c] = len;
			}
		}
		for (int i = 0; i < agg->numAggs; i++)
		{
			gg_aggval &v = row.agg[i];
			int col = aggmap[i].col;
			int fn = agg->aggs[i].aggfnoid;
			bool partial = agg->aggstage == GG_AGGSTAGE_PARTIAL;
			if (col < 0)                      /* count(*) */
			{
				v.i = (int64_t) x.count;
				continue;
			}
			uint64_t nn = x.n[col];
			int64_t ibits;
			memcpy(&ibits, &x.sum[col], 8);
			switch (fn)
			{
				case GG_AGG_COUNT_ANY:
					v.i = final_stage ? ibits : (int64_t) nn;
					break;
				case GG_AGG_COUNT_STAR:       /* FINAL stage only: int8pl over partial counts */
					v.i = ibits;
					break;
				case GG_AGG_SUM_FLOAT8:
				case GG_AGG_MIN_FLOAT8:
				case GG_AGG_MAX_FLOAT8:
					v.isnull = nn == 0;
					v.f[0] = nn ? x.sum[col] : 0.0;
					break;
				case GG_AGG_AVG_FLOAT8:
					if (partial)
					{
						v.f[0] = (double) nn; v.f[1] = x.sum[col]; v.f[2] = x.sumsq[col];
					}
					else if (nn == 0)
						v.isnull = 1;            /* float8_avg: N == 0 => NULL (float.c:1995) */
					else
						v.f[0] = x.sum[col] / (double) nn;
					break;
				default:                      /* int sum/min/max */
					v.isnull = nn == 0;
					v.i = nn ? ibits : 0;
					break;
			}
		}
	}
}

int gg_scanagg_fetch(gg_scanagg *p, gg_aggrow *out, int outcap, int *nout,
                     uint64_t *rows_scanned, uint64_t *rows_passed)
{
	if (!p || !nout) return GG_ERR_ARG;
	gg_engine *e = p->eng;
	GG_CUDA(cudaSetDevice(e->device));
	GG_CUDA(cudaStreamSynchronize(e->copy_stream));
	GG_CUDA(cudaStreamSynchronize(e->stream));
	uint32_t flags = 0;
	unsigned long long counters[2] = { 0, 0 };
	int n = 0;
	GG_CUDA(cudaMemcpy(&flags, p->d_err, sizeof flags, cudaMemcpyDeviceToHost));
	GG_CUDA(cudaMemcpy(counters, p->d_counters, sizeof counters, cudaMemcpyDeviceToHost));
	GG_CUDA(cudaMemcpy(&n, p->d_nout, sizeof n, cudaMemcpyDeviceToHost));
	if ((flags & GGP_EF_GROUP_OVERFLOW) && p->mode == MODE_PRIV)
	{
		/* more groups than the private-accumulator variant holds (the planner's numGroups was low or
		 * absent): replay the fed inputs on the transposed variant, like the reference's hybrid hash
		 * aggregate re-reading spilled input (execHHashagg.c:1093) */
		std::vector<gg_scanagg::Fed> replay = p->fed;
		p->mode = p->prog.nullable ? MODE_TRN : MODE_TR;
		int rc2 = scanagg_configure(p);
		if (rc2) return rc2;
		p->nrecs_total = GG_MERGE_CAP + p->grid * GGP_FAST_GROUPS;
		rc2 = gg_scanagg_reset(p);
		if (rc2) return rc2;
		for (const auto &f : replay)
		{
			rc2 = f.dev ? scanagg_launch(p, f.dev, f.nblocks, e->stream) : scanagg_stream_host(p, f.host, f.nblocks);
			if (rc2) return rc2;
		}
		p->fed = replay;
		return gg_scanagg_fetch(p, out, outcap, nout, rows_scanned, rows_passed);
	}
	if (rows_scanned) *rows_scanned = counters[0];
	if (rows_passed) *rows_passed = counters[1];
	int rc = gg_errflags_to_code(flags);
	if (rc) return rc;
	if (!p->has_state) n = 0;
	/* plain aggregation over zero rows still yields one row (nodeAgg.c:1247-1400) */
	std::vector<ggp_grec> recs((size_t) (n > 0 ? n : 1));
	if (n > 0) GG_CUDA(cudaMemcpy(recs.data(), p->recs, sizeof(ggp_grec) * n, cudaMemcpyDeviceToHost));
	if (n == 0 && p->agg.numCols == 0) { memset(&recs[0], 0, sizeof(ggp_grec)); n = 1; }
	if (n > outcap) { gg_set_error("output capacity %d < %d groups", outcap, n); return GG_ERR_NOMEM; }
	finalize_rows(&p->agg, p->aggmap, &p->prog, 0, recs.data(), n, out);
	*nout = n;
	return GG_OK;
}

void gg_scanagg_free(gg_scanagg *p)
{
	if (!p) return;
	cudaSetDevice(p->eng->device);
	cudaStreamSynchronize(p->eng->stream);
	cudaFree(p->recs); cudaFree(p->merged); cudaFree(p->vidx); cudaFree(p->vmap);
	cudaFree(p->d_nout); cudaFree(p->d_err); cudaFree(p->d_counters);
	for (int b = 0; b < 2; b++)
	{
		if (p->stage[b]) cudaFree(p->stage[b]);
		if (p->ev_copied[b]) cudaEventDestroy(p->ev_copied[b]);
		if (p->ev_consumed[b]) cudaEventDestroy(p->ev_consumed[b]);
	}
	delete p;
}

/* FINAL-stage Agg: the partial rows of all segments become group records (one accumulator column per
 * aggregate) and go through the same deterministic merge kernel; combine functions float8pl /
 * float8_combine / int8pl (nodeAgg.c:2123-2148, float.c:1842, int8.c:513). */
int gg_agg_final(gg_engine *e, const gg_agg *agg, const gg_aggrow *in, int nin,
                 gg_aggrow *out, int outcap, int *nout)
{
	if (!e || !agg || !nout || (nin && !in)) return GG_ERR_ARG;
	if (agg->numAggs > GGP_MAX_ACCS) { gg_set_error("too many aggregates"); return GG_ERR_UNSUPPORTED; }
	GG_CUDA(cudaSetDevice(e->device));
	ggp_program prog;
	ggp_aggmap aggmap[GG_MAX_AGGS];
	ggp_acckinds kinds;
	memset(&prog, 0, sizeof prog);
	memset(&kinds, 0, sizeof kinds);
	prog.nkeys = agg->numCols;
	prog.nacc = agg->numAggs;
	for (int c = 0; c < agg->numCols; c++)
	{
		int32_t t = agg->grpCol[c];
		prog.keytype[c] = (t == GG_FLOAT8OID) ? 2 : (t == GG_BPCHAROID || t == GG_VARCHAROID || t == GG_TEXTOID) ? 3 : 1;
	}
	for (int i = 0; i < agg->numAggs; i++)
	{
		aggmap[i].col = i;
		switch (agg->aggs[i].aggfnoid)
		{
			case GG_AGG_COUNT_STAR: case GG_AGG_COUNT_ANY: case GG_AGG_SUM_INT4: kinds.k[i] = GGP_ACC_I8SUM; break;
			case GG_AGG_SUM_FLOAT8: case GG_AGG_AVG_FLOAT8: kinds.k[i] = GGP_ACC_F8SUM; break;
			case GG_AGG_MIN_FLOAT8: kinds.k[i] = GGP_ACC_F8MIN; break;
			case GG_AGG_MAX_FLOAT8: kinds.k[i] = GGP_ACC_F8MAX; break;
			case GG_AGG_MIN_INT4: case GG_AGG_MIN_INT8: case GG_AGG_MIN_DATE: kinds.k[i] = GGP_ACC_I8MIN; break;
			case GG_AGG_MAX_INT4: case GG_AGG_MAX_INT8: case GG_AGG_MAX_DATE: kinds.k[i] = GGP_ACC_I8MAX; break;
			default: gg_set_error("aggregate %d not supported", agg->aggs[i].aggfnoid); return GG_ERR_UNSUPPORTED;
		}
		prog.acckind[i] = kinds.k[i];
	}
	std::vector<ggp_grec> recs((size_t) (nin > 0 ? nin : 1));
	uint32_t hostflags = 0;
	for (int r = 0; r < nin; r++)
	{
		ggp_grec &x = recs[r];
		memset(&x, 0, sizeof x);
		x.valid = 1;
		for (int c = 0; c < agg->numCols; c++)
		{
			if (in[r].keyisnull[c]) x.keynull |= 1u << c;
			else
			{
				uint64_t k = (uint64_t) in[r].key[c];
				if (prog.keytype[c] == 2)
				{
					double d; memcpy(&d, &k, 8);
					if (d == 0.0) k = 0; else if (d != d) k = 0x7ff8000000000000ull;
				}
				else if (agg->grpCol[c] == GG_INT4OID || agg->grpCol[c] == GG_DATEOID)
					k = (uint64_t) (int64_t) (int32_t) k;
				x.key[c] = k;
			}
		}
		for (int i = 0; i < agg->numAggs; i++)
		{
			const gg_aggval &v = in[r].agg[i];
			switch (agg->aggs[i].aggfnoid)
			{
				case GG_AGG_AVG_FLOAT8:
					x.n[i] = (uint64_t) v.f[0]; x.sum[i] = v.f[1]; x.sumsq[i] = v.f[2];
					if (!(fabs(v.f[1]) < INFINITY) || !(fabs(v.f[2]) < INFINITY)) hostflags |= GGP_EF_SAW_INF;
					/* float8_combine adds N even when it is 0; a zero-N state contributes nothing */
					break;
				case GG_AGG_SUM_FLOAT8: case GG_AGG_MIN_FLOAT8: case GG_AGG_MAX_FLOAT8:
					x.n[i] = v.isnull ? 0 : 1; x.sum[i] = v.f[0];
					if (!v.isnull && !(fabs(v.f[0]) < INFINITY)) hostflags |= GGP_EF_SAW_INF;
					break;
				default:
					x.n[i] = v.isnull ? 0 : 1; memcpy(&x.sum[i], &v.i, 8);
					break;
			}
		}
	}
	ggp_grec *d_recs = nullptr, *d_out = nullptr;
	int *d_vidx = nullptr, *d_vmap = nullptr, *d_n = nullptr;
	uint32_t *d_err = nullptr;
	int n = 0;
	uint32_t flags = 0;
	const int cap = nin > 0 ? nin : 1;
	GG_CUDA(cudaMalloc((void **) &d_recs, sizeof(ggp_grec) * cap));
	GG_CUDA(cudaMalloc((void **) &d_out, sizeof(ggp_grec) * cap));
	GG_CUDA(cudaMalloc((void **) &d_vidx, sizeof(int) * cap));
	GG_CUDA(cudaMalloc((void **) &d_vmap, sizeof(int) * cap));
	GG_CUDA(cudaMalloc((void **) &d_n, sizeof(int)));
	GG_CUDA(cudaMalloc((void **) &d_err, sizeof(uint32_t)));
	GG_CUDA(cudaMemcpyAsync(d_recs, recs.data(), sizeof(ggp_grec) * (size_t) nin, cudaMemcpyHostToDevice, e->stream));
	GG_CUDA(cudaMemcpyAsync(d_err, &hostflags, sizeof hostflags, cudaMemcpyHostToDevice, e->stream));
	GG_CUDA(cudaMemsetAsync(d_out, 0, sizeof(ggp_grec) * cap, e->stream));
	gg_merge_recs_kernel<<<1, 256, 0, e->stream>>>(d_recs, nin, agg->numCols, agg->numAggs, kinds,
	                                               d_out, cap, d_n, d_vidx, d_vmap, d_err);
	cudaError_t le = cudaGetLastError();
	e->launches++;
	if (le == cudaSuccess) le = cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, e->stream);
	if (le == cudaSuccess) le = cudaMemcpyAsync(&flags, d_err, sizeof flags, cudaMemcpyDeviceToHost, e->stream);
	if (le == cudaSuccess) le = cudaStreamSynchronize(e->stream);
	std::vector<ggp_grec> merged((size_t) (n > 0 ? n : 1));
	if (le == cudaSuccess && n > 0) le = cudaMemcpy(merged.data(), d_out, sizeof(ggp_grec) * n, cudaMemcpyDeviceToHost);
	cudaFree(d_recs); cudaFree(d_out); cudaFree(d_vidx); cudaFree(d_vmap); cudaFree(d_n); cudaFree(d_err);
	if (le != cudaSuccess) return gg_cuda_fail(le, "gg_agg_final");
	int rc = gg_errflags_to_code(flags);
	if (rc) return rc;
	if (n == 0 && agg->numCols == 0) { memset(&merged[0], 0, sizeof(ggp_grec)); n = 1; }
	if (n > outcap) { gg_set_error("output capacity %d < %d groups", outcap, n); return GG_ERR_NOMEM; }
	gg_agg fin = *agg;
	fin.aggstage = GG_AGGSTAGE_FINAL;
	finalize_rows(&fin, aggmap, &prog, 1, merged.data(), n, out);
	*nout = n;
	return GG_OK;
}

}  /* extern "C" */
