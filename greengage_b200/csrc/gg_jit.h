/*
 * gg_jit.h — plan-specialised kernels.
 *
 * The interpreter kernels (gg_scanagg.cu) run any supported plan.  For a plan that will scan a
 * large relation, the same kernel body (gg_scanagg_kernel.cuh) is compiled once more with the
 * plan's program expanded into straight-line exec_op()/walk_step() calls carrying literal
 * operands, so the compiler folds dispatch, operand decoding, constant offsets and attribute
 * properties away — the role PostgreSQL's own expression JIT plays on the CPU.  Compilation uses
 * NVRTC (dlopen'ed; optional) for sm_100a and the result is cached per (program, variant).
 * Without NVRTC, or with GGB200_JIT=0, the interpreter kernels are used: same results.
 */
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "gg_program.h"

struct gg_jit_kernel {
	cudaLibrary_t lib = nullptr;
	cudaKernel_t kernel = nullptr;
	double compile_ms = 0;
	bool precompiled = false;       /* from the build-time plan cache: `kernel` is a __global__ function address */
};

/* C++ source of the specialised translation unit (also used to pre-generate kernels offline) */
/* mode: the kernel role (MODE_* of gg_scanagg_kernel.cuh); join_probe_pc >= 0: the probe side of a join pipeline */
/* ctas: resident blocks per SM the launch bounds ask for (0: one for the private-accumulator variant, else two) */
/* mvcc: the kernel carries HeapTupleSatisfiesMVCC for tuples whose hint bits do not decide (needed when the scan has a snapshot);
 * without it such a tuple raises GGP_EF_VISIBILITY as it does when no snapshot was given — the rule's code stays out of the
 * kernels that never reach it (it cost the headline scan 7 % through sheer code size, profiles/r2g_ab_scan.jsonl) */
std::string gg_jit_scanagg_source(const ggp_program *prog, int mode, int threads, const char *suffix, int join_probe_pc = -1, int regslots = 0, int ctas = 0,
                                  int mvcc = 0);
int gg_priv_regslots(const ggp_program *prog, int mode, long long num_groups, int join_probe_pc);
uint64_t gg_plan_hash(const ggp_program *prog, int mode);
/* build-time plan cache (csrc/plans/gg_plan_cache.cu): address of the kernel specialised for this hash, or nullptr */
const void *gg_plan_cache_lookup(uint64_t hash, int threads, const ggp_program *prog);
/* compile (or fetch from the cache) the specialised scan+agg kernel; returns nullptr and fills err when JIT is
 * unavailable or fails */
gg_jit_kernel *gg_jit_scanagg(const ggp_program *prog, int mode, int threads, int device, char *err, int errlen, int join_probe_pc = -1,
                              int regslots = 0, int ctas = 0, int mvcc = 0);
