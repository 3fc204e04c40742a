"""Host logic of the executor-node surface (libggexec.so, include/gg_executor.h) that needs no GPU:
every declared symbol is exported, shapes outside the accelerated subset are refused the way the reference's
ExecInitNode refuses an unknown node (execProcnode.c:785: elog(ERROR, "unrecognized node type")), and the
Motion transport moves rows between two segments over gloo exactly as routed."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from greengage_b200 import capi, executor as ex, tpch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exec_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gg_executor.h")).read()
    decl = set(re.findall(r"\b(GgExec\w+)\s*\(", hdr))
    assert len(decl) >= 8
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "greengage_b200", "libggexec.so")]).decode()
    exp = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert [s for s in decl if s not in exp] == []


class FakeEngine:
    h = C.c_void_p(1)          # never dereferenced on the paths exercised here


def test_struct_sizes_match_the_header():
    src = r'''
    #include <stdio.h>
    #include "gg_executor.h"
    int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(GgPlan), sizeof(GgSeqScan), sizeof(GgAgg), sizeof(GgHashJoin),
                            sizeof(GgSort), sizeof(GgMotion), sizeof(GgTupleTableSlot), sizeof(GgEState), sizeof(GgInstrumentation)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(t) for t in (ex.GgPlan, ex.GgSeqScan, ex.GgAgg, ex.GgHashJoin, ex.GgSort, ex.GgMotion, ex.GgTupleTableSlot, ex.GgEState,
                                  ex.GgInstrumentation)]
    assert got == want


def test_shapes_outside_the_accelerated_subset_are_refused():
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_NARROW)
    b = ex.PlanBuilder()
    bare = b.seqscan(0, scan.desc, scan.qual)
    with pytest.raises(ex.ExecError) as e:
        ex.Executor(FakeEngine, pool, [None], bare)
    assert e.value.code == -6 and "underneath an Agg" in str(e.value)
    # Agg over a relation that is not resident
    with pytest.raises(ex.ExecError) as e:
        ex.Executor(FakeEngine, pool, [None], b.agg(b.seqscan(0, scan.desc, scan.qual), agg))
    assert e.value.code == -10
    # Motion with several segments needs a transport
    with pytest.raises(ex.ExecError) as e:
        ex.Executor(FakeEngine, pool, [None], b.motion(b.agg(b.seqscan(0, scan.desc, scan.qual), agg), ex.MOTION_GATHER), nsegs=4)
    assert e.value.code == -10 and "transport" in str(e.value)
    # a HashJoin whose inner side is not Hash(SeqScan)
    outer, inner, hj, jagg, jpool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, "count")
    bad = b.hashjoin(b.seqscan(0, outer.desc), b.seqscan(1, inner.desc), hj)
    with pytest.raises(ex.ExecError) as e:
        ex.Executor(FakeEngine, jpool, [None, None], b.agg(bad, jagg))
    assert e.value.code == -6
    with pytest.raises(ex.ExecError) as e:
        ex.Executor(FakeEngine, pool, [None], b.sort(b.seqscan(0, scan.desc), []))
    assert e.value.code == -6


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, ROOT)
        from greengage_b200 import executor as ex2
        tr = ex2.TorchTransport()
        rng = np.random.default_rng(100 + rank)
        n = 50 + rank * 7
        vals = rng.integers(-1000, 1000, (n, 3)).astype(np.int64)
        vals[:, 0] = rank                                             # sender tag
        nulls = (rng.random((n, 3)) < 0.2).astype(np.uint8)
        dest = rng.integers(0, world, n).astype(np.int32)
        rv, rn = tr.exchange_arrays(vals, nulls, dest, ex2.MOTION_HASH)
        bv, bn = tr.exchange_arrays(vals[:5], nulls[:5], np.full(5, -1, dtype=np.int32), ex2.MOTION_BROADCAST)
        gv, gn = tr.exchange_arrays(vals, nulls, np.zeros(n, dtype=np.int32), ex2.MOTION_GATHER)
        q.put(("ok", rank, vals, nulls, dest, rv, rn, bv, gv))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put(("err", traceback.format_exc()))


def test_transport_moves_rows_as_routed_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[0] == "ok", r[1]
    by = {r[1]: r for r in res}
    for me in (0, 1):
        want = np.concatenate([by[s][2][by[s][4] == me] for s in (0, 1)])      # sender order, like the UDP receiver's per-sender queues
        wantn = np.concatenate([by[s][3][by[s][4] == me] for s in (0, 1)])
        assert np.array_equal(by[me][5], want) and np.array_equal(by[me][6], wantn)
        assert np.array_equal(by[me][7], np.concatenate([by[0][2][:5], by[1][2][:5]]))     # broadcast: everyone gets everything
    assert np.array_equal(by[0][8], np.concatenate([by[0][2], by[1][2]])) and by[1][8].shape[0] == 0   # gather on segment 0


def test_c_example_builds_as_plain_c_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/q1_executor.c drives Q1 through GgExecInitNode / GgExecProcNode in C11 (-pedantic): the headers are a C
    ABI, not C++.  On a machine without a CUDA device the program must stop at gg_engine_create with the no-fallback
    message (on the GPU box it prints the Q1 rows)."""
    exe = tmp_path / "q1_executor"
    lib = os.path.join(ROOT, "greengage_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "q1_executor.c"), "-L", lib, "-lggexec", "-lggb200", "-lgghost",
                           "-Wl,-rpath," + lib, "-o", str(exe)])
    r = subprocess.run([str(exe), "2000"], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and "count" in r.stdout
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
