"""The executor-node surface across SEVERAL segments on CPU: the product's host C (greengage_b200/host/gg_executor.c +
gg_motion_host.c) linked against a stand-in device library that answers the C-ABI with the oracle
(tests/mock/ggb200_mock.c), one process per segment over gloo.  What runs for real here is everything above the device
engine: the fusion of the dispatched plan into pipelines, the Redistribute Motion (cdbhash routing of partial rows, the
TorchTransport exchange), the FINAL stage on the receiving segments, the sorted Gather to segment 0, slots, ReScan, and
the join pipeline under a Gather.  The answers are the single-segment oracle's / the reference's golden Q1."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build_mock(outdir):
    so = os.path.join(outdir, "libggexec_mock.so")
    host = os.path.join(ROOT, "greengage_b200", "host")
    objs = []
    for src, cc, std in ((os.path.join(host, "gg_executor.c"), "gcc", "-std=gnu11"), (os.path.join(host, "gg_motion_host.c"), "gcc", "-std=gnu11"),
                         (os.path.join(host, "gg_tupser.c"), "gcc", "-std=gnu11"),
                         (os.path.join(HERE, "mock", "ggb200_mock.c"), "gcc", "-std=gnu11"),
                         (os.path.join(HERE, "mock", "compile_glue.cpp"), "g++", "-std=c++17"),
                         (os.path.join(ROOT, "greengage_b200", "csrc", "gg_compile.cpp"), "g++", "-std=c++17")):   # the product's plan compiler
        obj = os.path.join(outdir, os.path.basename(src) + ".o")
        subprocess.check_call([cc, "-O1", "-g", "-fPIC", "-Wall", "-Wextra", std, "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + ["-L", os.path.join(ROOT, "oracle"), "-lggoracle",
                                                                 "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-z,defs", "-lm"])
    return so


class MockRel:
    def __init__(self, L, pages):
        self.pages = pages
        self.h = L.mock_relation(pages.ctypes.data, pages.size // 32768)


def _worker(rank, world, port, mock, case, q):
    try:
        dist = None
        if world > 1:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
            os.environ["MASTER_PORT"] = str(port)
            import torch.distributed as dist
            dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        from greengage_b200 import capi, executor as ex, tpch
        L = ex.bind(C.CDLL(mock))
        L.mock_engine.restype = C.c_void_p
        L.mock_relation.restype = C.c_void_p
        L.mock_relation.argtypes = [C.c_void_p, C.c_uint64]
        ex._lib = L                                                   # the Executor class now drives the mock-linked host code
        eng = L.mock_engine()
        tr = ex.TorchTransport() if world > 1 else None
        b = ex.PlanBuilder()
        if case == "single":
            q.put(("ok", rank, "single", single_segment_checks(L, eng, ex), None, 0))
            return
        if case in ("fail", "squelch", "plainfinal"):
            # one segment's slice fails / one segment stops before its first row / a plain FINAL aggregate above a Gather
            spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 20000, nsegs=world, seg=rank)
            pages, nb, nr = tpch.synth_generate(spec, nthreads=1)
            if case == "fail" and rank == 1:
                pages = pages.copy()
                pages[12:16] = 0xFF                                   # pd_lower / pd_upper of the first page: PageAddItem's sanity rules fail
            if case == "plainfinal":
                scan, agg, pool = tpch.count_star_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL)
                fin = capi.gg_agg.from_buffer_copy(bytes(agg))
                fin.aggstage = capi.AGGSTAGE_FINAL
                plan = b.agg(b.motion(b.agg(b.seqscan(0, scan.desc, scan.qual), agg), ex.MOTION_GATHER, [], 1), fin)
            else:
                from test_gpu_executor import q1_sorted_plan
                scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE)
                plan = q1_sorted_plan(b, scan, agg, True)
            x = ex.Executor(eng, pool, [MockRel(L, pages)], plan, nsegs=world, segindex=rank, transport=tr)
            try:
                rows = x.rows(limit=0) if (case == "squelch" and rank == 0) else x.rows()
                outcome = ("rows", len(rows), [r[0] for r in rows])
            except ex.ExecError as e:
                outcome = ("error", e.code, str(e))
            x.end()
            q.put(("ok", rank, "ctl", outcome, None, nr))
            if dist is not None:
                dist.destroy_process_group()
            return
        if case == "q1":
            from test_gpu_executor import q1_sorted_plan
            spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 60000, nsegs=world, seg=rank)
            pages, nb, nr = tpch.synth_generate(spec, nthreads=1)
            scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE)
            plan = q1_sorted_plan(b, scan, agg, True)
            rels = [MockRel(L, pages)]
        else:
            # co-located join (both sides DISTRIBUTED BY the order key) under a partial Agg, Gather, FINAL on segment 0 is
            # not expressible without a Result node; use: Gather Motion <- Agg(NORMAL, group key = order priority-like column)
            # per segment and combine on the host side of the test instead
            li, _, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 40000, seed=6, norders=8000, nsegs=world, seg=rank,
                                                            policy=capi.DIST_HASH), nthreads=1)
            od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 8000, seed=6, nsegs=world, seg=rank, policy=capi.DIST_HASH), nthreads=1)
            outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, "q3ish", capi.JOIN_INNER)
            plan = b.motion(b.agg(b.hashjoin(b.seqscan(0, outer.desc, outer.qual), b.hash(b.seqscan(1, inner.desc, inner.qual)), hj), agg),
                            ex.MOTION_GATHER, [], 1)
            rels = [MockRel(L, li), MockRel(L, od)]
        x = ex.Executor(eng, pool, rels, plan, nsegs=world, segindex=rank, transport=tr)
        kind = x.kind()
        rows = x.rows()
        again = None
        if case == "q1":
            x.rescan()                                                # every segment takes part in the rescan's exchanges
            again = x.rows()
        x.end()
        q.put(("ok", rank, kind, rows, again, nr))
        if dist is not None:
            dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put(("err", rank, traceback.format_exc(), None, None, 0))


def single_segment_checks(L, eng, ex):
    """what tests/test_gpu_executor.py checks on the device, with the oracle behind the C-ABI: node-surface control flow"""
    from _util import golden, lineitem_fixture_pages
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    from test_gpu_executor import b2f, q1_sorted_plan
    done = []
    desc, pages, n = lineitem_fixture_pages()
    exp = golden("q1_expected.json")
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_NORMAL, interval_days=exp["interval_days"], desc=desc)
    for two_stage in (False, True):
        x = ex.Executor(eng, pool, [MockRel(L, pages)], q1_sorted_plan(ex.PlanBuilder(), scan, agg, two_stage))
        assert x.kind() == ("motion" if two_stage else "sort")
        rows = x.rows()
        assert len(rows) == len(exp["rows"])
        for (v, nl, ty, ln), w in zip(rows, exp["rows"]):              # ORDER BY l_returnflag, l_linestatus
            assert capi.unpack_str(v[0], ln[0]) == w["returnflag"] and capi.unpack_str(v[1], ln[1]) == w["linestatus"]
            assert v[9] == w["count_order"]
            for col, name in ((2, "sum_qty"), (3, "sum_base_price"), (4, "sum_disc_price"), (5, "sum_charge"),
                              (6, "avg_qty"), (7, "avg_price"), (8, "avg_disc")):
                assert abs(b2f(v[col]) - float(w[name])) <= 1e-6 * abs(float(w[name]))
        assert x.rows() == []                                           # end of stream stays end of stream
        x.rescan()
        assert len(x.rows()) == len(rows)                               # ExecReScan runs the slice again
        x.rescan()
        assert len(x.rows(limit=2)) == 2 and x.rows() == []              # squelched after LIMIT
        x.end()
        done.append("q1-two-stage" if two_stage else "q1-one-stage")
    li, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 30_000, seed=6, norders=6_000))
    od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 5_000, seed=6))
    outer, inner, hj, jagg, jpool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, "q3ish", capi.JOIN_INNER)
    want, _ = po.hashjoin_agg(outer, inner, hj, jagg, jpool, li, od)
    b = ex.PlanBuilder()
    plan = b.sort(b.agg(b.hashjoin(b.seqscan(0, outer.desc, outer.qual), b.hash(b.seqscan(1, inner.desc, inner.qual)), hj), jagg),
                  [capi.make_sortkey(0, capi.BPCHAROID, desc=True)])
    x = ex.Executor(eng, jpool, [MockRel(L, li), MockRel(L, od)], plan)
    rows = x.rows()
    keys = [capi.unpack_str(v[0], ln[0]) for v, nl, ty, ln in rows]
    assert keys == sorted(keys, reverse=True) and len(rows) == len(want)
    byk = {r.key[0]: r for r in want}
    for v, nl, ty, ln in rows:
        assert v[1] == byk[v[0]].agg[0].i and v[3] == byk[v[0]].agg[2].i
    x.end()
    done.append("join-sort-desc")
    # a Sort whose rows exceed the operator's memory goes external: sorted runs, merged on the host (tuplesort_mk.c:2019) —
    # same rows, same order as the in-memory sort, for a total order and for one with many ties and NULLS FIRST / DESC keys
    p = capi.ExprPool()
    c = tpch.LI_NARROW_COLS
    okey, price = p.var(c["orderkey"], capi.INT8OID), p.var(c["extendedprice"], capi.FLOAT8OID)
    gagg = capi.make_agg(capi.AGGSTAGE_NORMAL, [okey], [(capi.AGG_COUNT_STAR, -1), (capi.AGG_SUM_FLOAT8, price)], num_groups=6000)
    lscan = capi.make_scan(capi.synth_tupdesc(capi.TAB_LINEITEM_NARROW), -1)
    L.GgExecSortRuns.argtypes = [C.c_void_p]
    for keys in ([capi.make_sortkey(1, capi.INT8OID, desc=True), capi.make_sortkey(0, capi.INT8OID)],           # count desc, orderkey: total
                 [capi.make_sortkey(1, capi.INT8OID)],                                                            # count only: ties keep input order
                 [capi.make_sortkey(2, capi.FLOAT8OID, desc=True, nulls_first=True), capi.make_sortkey(0, capi.INT8OID, desc=True)]):
        got = []
        for mem in (0, 16 * 1024):
            b = ex.PlanBuilder()
            x = ex.Executor(eng, p.pool, [MockRel(L, li)], b.sort(b.agg(b.seqscan(0, lscan.desc, lscan.qual), gagg), keys), operator_mem=mem)
            got.append([tuple(v) for v, nl, ty, ln in x.rows()])
            runs = L.GgExecSortRuns(x.state)
            assert (runs == 1) if mem == 0 else (runs >= 10), (mem, runs)
            # Instrumentation the way EXPLAIN ANALYZE reads it: the Sort handed up every row once, in one execution; a ReScan
            # and a second drain double the rows and the loops
            ins = dict((k, i) for k, i in x.instrumentation())
            assert ins["sort"].ntuples == len(got[-1]) and ins["sort"].nloops == 1 and ins["sort"].sort_runs == runs
            x.rescan()
            assert len(x.rows()) == len(got[-1])
            ins = dict((k, i) for k, i in x.instrumentation())
            assert ins["sort"].ntuples == 2 * len(got[-1]) and ins["sort"].nloops == 2
            x.end()
        assert len(got[0]) > 4000 and got[0] == got[1]
        k0 = keys[0]
        col = [r[k0.col] if k0.typid != capi.FLOAT8OID else b2f(r[k0.col]) for r in got[1]]
        assert col == sorted(col, reverse=bool(k0.desc))
    done.append("external-sort")
    # a Motion over several segments without a transport is refused at init, not at run time
    try:
        ex.Executor(eng, pool, [MockRel(L, pages)], q1_sorted_plan(ex.PlanBuilder(), scan, agg, True), nsegs=2, segindex=0)
        raise AssertionError("accepted")
    except ex.ExecError as e:
        assert e.code == -10
    done.append("motion-needs-transport")
    # malformed plan trees are error codes, never crashes (scripts/fuzz/run_executor_fuzz.sh found these under ASan)
    b = ex.PlanBuilder()
    plan = q1_sorted_plan(b, scan, agg, True)
    sort = next(n for n in b.nodes if isinstance(n, ex.GgSort))
    sort.plan.lefttree = ex._as_plan(plan)                               # a cycle: Sort's child is the top Motion again
    try:
        ex.Executor(eng, pool, [MockRel(L, pages)], plan)
        raise AssertionError("accepted a cyclic plan")
    except ex.ExecError as e:
        assert e.code == -10 and "deeper" in str(e)
    b = ex.PlanBuilder()
    plan = q1_sorted_plan(b, scan, agg, True)
    fin = next(n for n in b.nodes if isinstance(n, ex.GgAgg) and n.agg.aggstage == capi.AGGSTAGE_FINAL)
    fin.agg.numAggs = 12                                                 # a FINAL stage that expects more state columns than arrive
    x = ex.Executor(eng, pool, [MockRel(L, pages)], plan)
    try:
        x.rows()
        raise AssertionError("ran a FINAL Agg over too few columns")
    except ex.ExecError as e:
        assert e.code == -10 and "FINAL Agg expects" in str(e)
    x.end()
    fin.agg.numAggs = 10 ** 6
    try:
        ex.Executor(eng, pool, [MockRel(L, pages)], plan)
        raise AssertionError("accepted")
    except ex.ExecError as e:
        assert e.code == -10
    done.append("malformed-trees-refused")
    return done


def run(world, case, tmp_path):
    import torch.multiprocessing as mp
    mock = build_mock(str(tmp_path))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() * 3 + world * 11 + len(case)) % 600
    procs = [ctx.Process(target=_worker, args=(r, world, port, mock, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[0] == "ok", r[2]
    return {r[1]: r for r in res}


def b2f(v):
    return np.int64(v).view(np.float64).item()


@pytest.mark.parametrize("world", [2, 3])
def test_dispatched_two_stage_q1_over_segments(world, tmp_path):
    """Gather Motion(merge) <- Sort <- Agg(FINAL) <- Redistribute Motion <- Agg(PARTIAL) <- SeqScan, one slice set per segment
    (expected/tpch500GB.out:1771-1782): segment 0 returns the ordered answer of the whole table, the others nothing."""
    sys.path.insert(0, ROOT)
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    by = run(world, "q1", tmp_path)
    assert sum(by[r][5] for r in by) == 60000
    pages, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 60000))
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE)
    want, _, _ = po.seqscan_agg(scan, agg, pool, pages)
    want = sorted(want, key=lambda r: (r.key[0] & 0xFF, r.key[1] & 0xFF))
    for r in range(1, world):
        assert by[r][2] == "motion" and by[r][3] == [] and by[r][4] == []
    rows = by[0][3]
    assert len(rows) == len(want) == 4
    for (v, nl, ty, ln), w in zip(rows, want):                        # merged order = ORDER BY l_returnflag, l_linestatus
        assert (v[0], v[1]) == (w.key[0], w.key[1]) and v[9] == w.agg[7].i
        for col in range(7):                                          # 4 sums, 3 avgs (float8_avg of the combined states)
            assert abs(b2f(v[2 + col]) - w.agg[col].f[0]) <= 1e-9 * abs(w.agg[col].f[0])
    assert [r[0][:2] + [r[0][9]] for r in by[0][4]] == [r[0][:2] + [r[0][9]] for r in rows]      # the rescan returned the same groups


def test_colocated_join_under_a_gather(tmp_path):
    """HashJoin(SeqScan, Hash(SeqScan)) + Agg on every segment over co-located relations, Gather Motion to segment 0: the
    per-segment partial answers add up to the single-segment join"""
    sys.path.insert(0, ROOT)
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    by = run(2, "join", tmp_path)
    li, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 40000, seed=6, norders=8000))
    od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 8000, seed=6))
    outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, "q3ish", capi.JOIN_INNER)
    want, nj = po.hashjoin_agg(outer, inner, hj, agg, pool, li, od)
    assert by[1][3] == [] and by[0][2] == "motion"
    got = {}
    for v, nl, ty, ln in by[0][3]:                                    # a group arrives once per segment that has it
        g = got.setdefault(v[0], [0, 0.0, None])
        g[0] += v[1]                                                  # count(*): int8pl
        g[1] += b2f(v[2])                                             # sum(revenue)
        g[2] = v[3] if g[2] is None else min(g[2], v[3])              # min(o_orderdate)
    assert len(got) == len(want) and sum(g[0] for g in got.values()) == nj
    for w in want:
        g = got[w.key[0]]
        assert g[0] == w.agg[0].i and g[2] == w.agg[2].i
        assert abs(g[1] - w.agg[1].f[0]) <= 1e-9 * abs(w.agg[1].f[0])


def test_node_surface_control_flow_on_one_segment(tmp_path):
    """ReScan, end of stream, Squelch after a LIMIT, Sort DESC above a join pipeline, the loopback Motions of the two-stage
    plan — tests/test_gpu_executor.py's checks with the oracle behind the C-ABI, in a child process"""
    by = run(1, "single", tmp_path)
    assert by[0][3] == ["q1-one-stage", "q1-two-stage", "join-sort-desc", "external-sort", "motion-needs-transport", "malformed-trees-refused"]


def test_a_failing_segment_does_not_leave_its_peers_in_the_exchange(tmp_path):
    """One segment's scan hits a corrupted page.  It still enters the Motion (with no rows and its status), so the other
    segment is not left waiting in a collective, and BOTH come back with an error: the failing one with its own, the peer
    with GG_ERR_PEER (the reference: error propagation + SendStopMessage, cdbmotion.c:342, nodeMotion.c:1730)."""
    by = run(2, "fail", tmp_path)
    assert by[1][3][0] == "error" and by[1][3][1] not in (0, -12)          # its own error, not the peer notice
    assert by[0][3][0] == "error" and by[0][3][1] == -12


def test_a_segment_squelched_before_its_first_row_still_takes_part(tmp_path):
    """LIMIT 0 on segment 0: ExecSquelchNode reaches a Motion that has not run — it runs (the peers are in the exchange) and
    then hands out nothing; segment 1 finishes normally"""
    by = run(2, "squelch", tmp_path)
    assert by[0][3] == ("rows", 0, []) and by[1][3][0] == "rows"


def test_plain_final_aggregate_above_a_gather_yields_its_row_on_the_receiver_only(tmp_path):
    """Agg(FINAL, no keys) <- Gather <- Agg(PARTIAL): the slice above a Gather exists on the receiving segment only
    (nodeMotion.c:1036-1053); a non-receiving segment must not synthesise the empty-input row count = 0"""
    by = run(2, "plainfinal", tmp_path)
    assert by[0][3][0] == "rows" and by[0][3][1] == 1 and by[0][3][2][0][0] == by[0][5] + by[1][5]
    assert by[1][3] == ("rows", 0, [])
