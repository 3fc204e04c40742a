/*
 * gg_pipeline.h — the host object behind gg_scanagg (and, as its probe side, gg_joinagg): shared by gg_scanagg.cu and
 * gg_join.cu.
 */
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include "gg_scanagg_kernel.cuh"
#include "gg_engine.h"
#include "gg_jit.h"

/* Launch configuration of the kernels that run two small blocks per SM by default (hash build, Motion send, general
 * HashAggregate, the transposed / nullable scan and probe variants): consumer warps per block, ring stages, team size
 * (ScanAggParams.team) and blocks per SM.  GGB200_NP_CONFIG="ncons,stages,team,ctas" overrides it for experiments; block
 * sizes other than the default need the run-time specialised kernel (the interpreter kernels are built for 256 threads). */
struct gg_npconfig { int ncons, nstage, team, ctas; bool forced; };
static inline gg_npconfig gg_np_config(int ncons, int nstage)
{
	gg_npconfig c = { ncons, nstage, 0, 2, false };
	const char *env = getenv("GGB200_NP_CONFIG");
	int a, b, t = 0, k = 2;
	if (env && sscanf(env, "%d,%d,%d,%d", &a, &b, &t, &k) >= 2 && a >= 1 && a <= 30 && b >= 2 && b <= 6 && t >= 0 && t <= a && k >= 1 && k <= 4)
	{ c.ncons = a; c.nstage = b; c.team = t; c.ctas = k; c.forced = true; }
	/* at most one team per ring slot (a team's pages arrive on its own barrier set, BlockTable::teamfull) */
	if (c.team > 0 && c.ncons / c.team > c.nstage) c.ncons = c.team * c.nstage;
	return c;
}

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
	int mode = ggd::MODE_PRIV;           /* kernel variant; escalates PRIV -> TR when a run overflows its group capacity */
	int ctas_per_sm = 2, gcap = 0;
	uint32_t scratch_off = 0, cnt_off = 0, acc_off = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kev;   /* events around every scan kernel launch since reset */
	size_t kev_used = 0;
	gg_jit_kernel *jit = nullptr;   /* plan-specialised kernel for the current variant, or nullptr: interpreter */
	gg_jit_kernel *jit_snap = nullptr;  /* the same with the snapshot rule built in (gg_jit.h mvcc), compiled when a launch first finds
	                                     * the engine holding a snapshot; follows `jit` through every reconfiguration */
	const gg_jit_kernel *jit_snap_of = nullptr;     /* the `jit` that jit_snap was compiled next to */
	int regslots = -1;              /* private-accumulator variant: trailing value slots kept in registers (-1: not decided) */
	int chunks_per_page = 0;        /* 32-row chunks per page of the relation being scanned (0: not sampled yet) */
	int items_per_page = 0;         /* line pointers of the sampled page */
	int team = 0;                   /* consumer warps per team (ScanAggParams.team); 0: chunks dealt across all warps */
	bool np_forced = false;         /* launch configuration came from GGB200_NP_CONFIG */
	bool is_join = false;           /* probe side of a gg_joinagg: prog = the probe program, jt = the built table */
	ggd::HashAggTable ha = {};           /* MODE_HASH: the group table in HBM */
	void *ha_mem = nullptr;
	uint64_t ha_cap = 0;
	unsigned long long *d_nout64 = nullptr;
	ggd::JoinTable jt = {};
	int join_probe_pc = -1;
	size_t smem = 0;
	/* device state */
	ggp_grec *recs = nullptr;       /* [GG_MERGE_CAP (previous merged)] ++ [grid * GGP_FAST_GROUPS (block records)] */
	ggp_grec *merged = nullptr;     /* [GG_MERGE_CAP] output of the merge kernel */
	int *vidx = nullptr, *vmap = nullptr;
	/* status words of the pipeline: one device block, mirrored into pinned host memory by ONE copy per fetch
	 * (together with the first merged group records) */
	struct Status { uint32_t err; int nout; unsigned long long counters[2]; };
	Status *d_status = nullptr;
	int *d_nout = nullptr;                  /* = &d_status->nout */
	uint32_t *d_err = nullptr;              /* = &d_status->err */
	unsigned long long *d_counters = nullptr;   /* = d_status->counters */
	struct HostMirror { Status st; ggp_grec recs[GGP_FAST_GROUPS]; };
	HostMirror *h_mirror = nullptr;         /* pinned */
	int nrecs_total = 0, nrecs_cap = 0;
	/* inputs of the current accumulation, kept so that a group-capacity overflow can be replayed on a wider variant */
	struct Fed { const uint8_t *dev; const void *host; uint64_t nblocks; uint64_t nrows; bool fill; int32_t tile_rows; };   /* tile_rows > 0: dev = the column descriptors of an AOCS feed */
	gg_aocs_devcol *d_aocs = nullptr;       /* device copy of the column descriptors of gg_scanagg_run_aocs */
	int32_t aocs_unit_rows = 0;             /* rows of every projected column staged per ring slot (set by gg_scanagg_run_aocs) */
	std::vector<Fed> fed;
	bool has_state = false;
	/* set by a batched join: how to feed the inputs again (its batches, each with its own hash table) when fetch has to
	 * replay them on a wider kernel variant; empty: replay `fed` */
	std::function<int()> replay_hook;
	/* host staging for the streamed path */
	uint8_t *stage[2] = { nullptr, nullptr };
	cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_consumed[2] = { nullptr, nullptr };
};

/* gg_scanagg.cu */
int scanagg_finish_create(gg_scanagg *p, gg_scanagg **out);      /* after p->prog / p->aggmap are compiled */
/* one launch of the pipeline's kernel over device pages (fill_inner: the HJ_FILL_INNER_TUPLES pass of a right/full join) */
int scanagg_launch(gg_scanagg *p, const uint8_t *dev_pages, uint64_t nblocks, cudaStream_t st, uint64_t nrows = 0, bool fill_inner = false, int32_t aocs_tile_rows = 0);
/* gg_motion.cu: the partitioning kernel behind gg_motion_partition, with the routing rule as a parameter (route 0: segments by
 * cdbhash + jump consistent hash; route 1: hash-join batches by the batch bits above `shift`) */
int gg_partition_rows(gg_engine *e, const gg_scan *scan, const gg_exprpool *pool,
                      const int32_t *hashkeys, int nkeys, const int32_t *payload, int npayload,
                      int nsegs, int route, int shift, gg_relation *r, uint64_t first_block, uint64_t nblocks,
                      void *device_out_rows, uint64_t out_cap_rows,
                      uint64_t *host_counts, uint64_t *host_offsets);
/* gg_join.cu: the probe-side kernels of a join pipeline (interpreter path) */
int gg_probe_kernel_prepare(gg_scanagg *p);
int gg_probe_kernel_launch(gg_scanagg *p, const ggd::ScanAggParams &prm, cudaStream_t st);
