"""Byte-compare what two builds of the plan compiler produce (ggp_program + aggregate map, ggp_joinprog) for every plan the
tests know: the random generators' plans, the TPC-H plans, the reference-golden plans.  Used to show that a change to
gg_compile.cpp (validation, an experiment switch that is off) leaves the programs round 1 validated on the GPU untouched.
    OLD=/path/old.so NEW=/path/new.so python scripts/fuzz/compiler_compare.py        (both built with compiler_wrap.cpp)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from greengage_b200 import capi, tpch
from greengage_b200.capi import ExprPool
import test_gpu_random_plans as rp, test_gpu_random_joins as rj
from _util import (make_desc, tpch_q4_plan, tpch_q12_plan, tpch_q6_plan, tpch_join_fixture, lineitem_fixture_pages, j1j2_fixture, j1j2_join,
                   j1j2_agg, J1J2_QUERIES, onek_fixture, onek_plans, gp_hashagg_case, golden)

OLD, NEW = C.CDLL(os.environ["OLD"]), C.CDLL(os.environ["NEW"])
b1, b2 = C.create_string_buffer(1 << 17), C.create_string_buffer(1 << 17)
n = [0, 0]


def sa(scan, agg, pool):
    r1 = OLD.fz_scanagg_raw(C.byref(scan), C.byref(agg), C.byref(pool), b1, 1 << 17)
    r2 = NEW.fz_scanagg_raw(C.byref(scan), C.byref(agg), C.byref(pool), b2, 1 << 17)
    assert r1 == r2, (r1, r2)
    if r1 > 0:
        assert b1.raw[:r1] == b2.raw[:r1]
        n[0] += 1
    else:
        n[1] += 1


def jn(o, i, hj, agg, pool):
    r1 = OLD.fz_join_raw(C.byref(o), C.byref(i), C.byref(hj), C.byref(agg), C.byref(pool), b1, 1 << 17)
    r2 = NEW.fz_join_raw(C.byref(o), C.byref(i), C.byref(hj), C.byref(agg), C.byref(pool), b2, 1 << 17)
    assert r1 == r2, (r1, r2)
    if r1 > 0:
        assert b1.raw[:r1] == b2.raw[:r1]
        n[0] += 1
    else:
        n[1] += 1


desc = make_desc([(capi.INT4OID, 4, "i", 1, 1), (capi.INT4OID, 4, "i", 1, 1), (capi.INT4OID, 4, "i", 1, 0), (capi.FLOAT8OID, 8, "d", 1, 1),
                  (capi.FLOAT8OID, 8, "d", 1, 0), (capi.BPCHAROID, -1, "i", 0, 0), (capi.DATEOID, 4, "i", 1, 1), (capi.INT8OID, 8, "d", 1, 1)])
for seed in range(400):
    rng = np.random.default_rng(1000 + seed)
    p = ExprPool(); g = rp.Gen(rng, p)
    qual = g.boolean(2) if rng.random() < 0.8 else -1
    aggs = [(capi.AGG_COUNT_STAR, -1)]
    for _ in range(int(rng.integers(1, 5))):
        fn = int(rng.choice([capi.AGG_SUM_FLOAT8, capi.AGG_AVG_FLOAT8, capi.AGG_MIN_FLOAT8, capi.AGG_MAX_FLOAT8, capi.AGG_COUNT_ANY]))
        aggs.append((fn, g.f8(2)))
    if rng.random() < 0.5:
        aggs.append((int(rng.choice([capi.AGG_SUM_INT4, capi.AGG_MIN_INT4, capi.AGG_MAX_INT4])), p.var(int(rng.choice([2, 3])), capi.INT4OID)))
    keys = [[], [p.var(1, capi.INT4OID)], [p.var(1, capi.INT4OID), p.var(6, capi.BPCHAROID)]][int(rng.integers(0, 3))]
    stage = capi.AGGSTAGE_PARTIAL if rng.random() < 0.3 else capi.AGGSTAGE_NORMAL
    sa(capi.make_scan(desc, qual), capi.make_agg(stage, keys, aggs, num_groups=int(rng.choice([0, 20, 500]))), p.pool)
for seed in range(100):
    outer, inner, hj, agg, p, opages, ipages, what = rj.random_join_case(seed)
    jn(outer, inner, hj, agg, p.pool)
for t in (capi.TAB_LINEITEM_WIDE, capi.TAB_LINEITEM_NARROW):
    for st in (capi.AGGSTAGE_NORMAL, capi.AGGSTAGE_PARTIAL):
        sa(*tpch.q1_plan(t, st))
    sa(*tpch.count_star_plan(t))
    for kind in ("count", "q3ish"):
        for jt in range(7):
            jn(*tpch.join_plan(t, kind, jt))
li_desc, li_pages, ord_desc, ord_pages, shipmode_code, priority_code = tpch_join_fixture()
jn(*tpch_q4_plan(li_desc, ord_desc, golden("tpch_join_expected.json")["q4"]))
for high in (True, False):
    jn(*tpch_q12_plan(li_desc, ord_desc, golden("tpch_join_expected.json")["q12"], shipmode_code, priority_code, high))
d, _, _ = lineitem_fixture_pages(); sa(*tpch_q6_plan(d))
d1, p1, d2, p2, g = j1j2_fixture()
for name in J1J2_QUERIES:
    p, o, i, hj = j1j2_join(d1, d2, name); jn(o, i, hj, j1j2_agg(p), p.pool)
d, pg, exp = onek_fixture()
for pl in onek_plans(d, exp): sa(*pl)
d, pg, scan, agg, pool, want = gp_hashagg_case(); sa(scan, agg, pool)
print("byte-identical programs for %d plans (%d refused by both)" % (n[0], n[1]))
