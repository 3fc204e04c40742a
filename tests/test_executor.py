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


def test_the_external_sorts_merge_order_is_the_references_comparators():
    """sort_row_cmp (gg_executor.c) orders rows by the keys the device sorts by; held here, pair by pair, to the order the oracle's
    restatement of the reference's comparators (btint8cmp, btfloat8cmp with NaN last and -0 = +0, bpcharcmp on packed strings,
    date_cmp, NULLS FIRST / LAST, DESC) gives the same rows."""
    import random
    from oracle import pyoracle as po
    L = ex.exec_lib()
    L.GgExecDebugSortCompare.argtypes = [C.POINTER(capi.gg_sortkey), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
    rnd = random.Random(3)
    f8 = lambda d: int(np.float64(d).view(np.int64))
    specials = [0.0, -0.0, float("nan"), float("inf"), float("-inf"), 1.5, -1.5, 1e300, -1e-300]
    n = 400
    rows = np.zeros((n, 4), dtype=np.int64)
    nulls = (np.array([[rnd.random() < 0.15 for _ in range(4)] for _ in range(n)])).astype(np.uint8)
    for i in range(n):
        rows[i, 0] = rnd.choice([-5, 0, 7, 2**40, -2**40, rnd.randint(-100, 100)])                          # int8
        rows[i, 1] = f8(rnd.choice(specials) if rnd.random() < 0.5 else rnd.uniform(-10, 10))               # float8
        s = bytes(rnd.choice(b"AB ab") for _ in range(rnd.randint(0, 8))).rstrip(b" ")
        rows[i, 2] = int.from_bytes(s.ljust(8, b"\0"), "little", signed=True)                               # packed bpchar
        rows[i, 3] = rnd.randint(-3000, 3000)                                                                # date
    for trial in range(20):
        cols = rnd.sample(range(4), rnd.randint(1, 4))
        typ = {0: capi.INT8OID, 1: capi.FLOAT8OID, 2: capi.BPCHAROID, 3: capi.DATEOID}
        keys = [capi.make_sortkey(c, typ[c], desc=rnd.random() < 0.5, nulls_first=rnd.choice([None, True, False])) for c in cols]
        perm = po.sort_perm(keys, 4, rows, nulls)
        ka = (capi.gg_sortkey * len(keys))(*keys)
        cmp = lambda a, b: L.GgExecDebugSortCompare(ka, len(keys), 4, rows.ctypes.data, nulls.ctypes.data, int(a), int(b))
        # the oracle's order never goes down under the product's comparator, and equal neighbours are equal both ways
        for a, b in zip(perm[:-1], perm[1:]):
            c = cmp(a, b)
            assert c <= 0, (trial, [(k.col, k.desc, k.nulls_first) for k in keys], rows[a], rows[b])
            assert cmp(b, a) == -c
