"""Fuzz the executor-node surface (greengage_b200/host/gg_executor.c) under AddressSanitizer + UBSan with malformed plan
trees: node tags, child pointers, counts, column numbers, relation ids, Motion fields and EState fields set to boundary and
garbage values.  GgExecInitNode must refuse (NULL + error) or accept; an accepted plan is run to the end.  The device
library is the oracle-backed stand-in of tests/mock.  Run through scripts/fuzz/run_executor_fuzz.sh."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from greengage_b200 import capi, executor as ex, tpch
from test_gpu_executor import q1_sorted_plan

L = ex.bind(C.CDLL(os.environ["FZ_LIB"]))
L.mock_engine.restype = C.c_void_p
L.mock_relation.restype = C.c_void_p
L.mock_relation.argtypes = [C.c_void_p, C.c_uint64]
ex._lib = L
eng = L.mock_engine()


class Rel:
    def __init__(self, pages):
        self.pages, self.h = pages, L.mock_relation(pages.ctypes.data, pages.size // 32768)


li, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 3000, seed=6, norders=600))
od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 500, seed=6))
rels = [Rel(li), Rel(od)]
rng = np.random.default_rng(5)
vals = [-2 ** 31, -1000, -2, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 15, 16, 17, 31, 32, 33, 100, 65535, 2 ** 31 - 1]


def build(kind):
    b = ex.PlanBuilder()
    if kind == 0:
        scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_NARROW)
        return b, q1_sorted_plan(b, scan, agg, bool(rng.integers(0, 2))), pool
    outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, "q3ish", capi.JOIN_INNER)
    plan = b.sort(b.agg(b.hashjoin(b.seqscan(0, outer.desc, outer.qual), b.hash(b.seqscan(1, inner.desc, inner.qual)), hj), agg),
                  [capi.make_sortkey(0, capi.BPCHAROID, desc=True)])
    if rng.random() < 0.5:
        plan = b.motion(plan, ex.MOTION_GATHER, [], 1)
    return b, plan, pool


def mutate(b):
    n = b.nodes[int(rng.integers(0, len(b.nodes)))]
    v = int(rng.choice(vals))
    c = int(rng.integers(0, 6))
    if c == 0: n.plan.type = v
    elif c == 1: n.plan.qual = v
    elif c == 2:
        if rng.random() < 0.5: n.plan.lefttree = None
        else: n.plan.righttree = ex._as_plan(b.nodes[int(rng.integers(0, len(b.nodes)))])       # may even make a cycle
    elif isinstance(n, ex.GgSeqScan): n.scanrelid = v if v not in (0, 1) else n.scanrelid    # a wrong but resident relation is the caller's garbage in, and the oracle behind the stand-in does not survive mismatched pages
    elif isinstance(n, ex.GgSort):
        if rng.random() < 0.5: n.numCols = v
        else: n.keys[int(rng.integers(0, 4))].col = v
    elif isinstance(n, ex.GgMotion):
        f = int(rng.integers(0, 5))
        if f == 0: n.motionType = v
        elif f == 1: n.numHashCols = v
        elif f == 2: n.hashCol[int(rng.integers(0, 4))] = v
        elif f == 3: n.numSortCols = v
        else: n.sortKeys[int(rng.integers(0, 4))].col = v
    elif isinstance(n, ex.GgAgg):
        if rng.random() < 0.5: n.agg.numCols = v
        else: n.agg.numAggs = v
    elif isinstance(n, ex.GgHashJoin):
        n.hj.nkeys = v


stats = {"accepted": 0, "refused": 0, "ran": 0, "run_error": 0}
for it in range(int(os.environ.get("FZ_ITERS", "6000"))):
    b, plan, pool = build(int(rng.integers(0, 2)))
    for _ in range(int(rng.integers(1, 4))):
        mutate(b)
    top = plan if rng.random() < 0.9 else b.nodes[int(rng.integers(0, len(b.nodes)))]
    nsegs, seg = (1, 0) if rng.random() < 0.8 else (int(rng.choice(vals)), int(rng.choice(vals)))
    try:
        x = ex.Executor(eng, pool, rels if rng.random() < 0.9 else [rels[0], None], top, nsegs=nsegs, segindex=seg)
    except ex.ExecError:
        stats["refused"] += 1
        continue
    stats["accepted"] += 1
    try:
        x.rows()
        if rng.random() < 0.3:
            x.rescan()
            x.rows(limit=1)
        stats["ran"] += 1
    except ex.ExecError:
        stats["run_error"] += 1
    x.end()
print(stats)
