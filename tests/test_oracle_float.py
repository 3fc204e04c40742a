"""float8 / int8 / date semantics of the oracle against the reference's own float.o, int8.o, date.o
(golden float_kat.json): CHECKFLOATVAL overflow/underflow ERRORs, division by zero, NaN ordering,
float8_accum / float8_combine / float8_avg, int8pl overflow, date vs timestamp promotion."""

import pytest

from _util import b2f, f2b, golden, make_desc
from greengage_b200 import capi
from greengage_b200.capi import ExprPool
from oracle import pyoracle as po

K = golden("float_kat.json")
FN = {"pl": capi.F_FLOAT8PL, "mi": capi.F_FLOAT8MI, "mul": capi.F_FLOAT8MUL, "div": capi.F_FLOAT8DIV}


def test_arith_and_errors():
    for fn, a, b, err, r, _msg in K["arith"]:
        p = ExprPool()
        root = p.func(FN[fn], capi.FLOAT8OID, p.const(capi.FLOAT8OID, b2f(a)), p.const(capi.FLOAT8OID, b2f(b)))
        # NaN constants lose nothing: const() goes through float(), bits preserved for the values used
        p.pool.nodes[0].constvalue = int(a)
        p.pool.nodes[1].constvalue = int(b)
        rc, v, isnull = po.eval_expr(p.pool, root)
        if err:
            assert rc != 0, (fn, b2f(a), b2f(b))
        else:
            assert rc == 0 and not isnull
            x, y = b2f(v), b2f(r)
            assert v == int(r) or (x != x and y != y), (fn, b2f(a), b2f(b), x, y)


def test_comparisons_nan_order():
    for a, b, eq, lt, le, cmp3 in K["cmp"]:
        for fid, want in ((capi.F_FLOAT8EQ, eq), (capi.F_FLOAT8LT, lt), (capi.F_FLOAT8LE, le)):
            p = ExprPool()
            root = p.func(fid, capi.BOOLOID, p.const(capi.FLOAT8OID, 0.0), p.const(capi.FLOAT8OID, 0.0))
            p.pool.nodes[0].constvalue = int(a)
            p.pool.nodes[1].constvalue = int(b)
            rc, v, isnull = po.eval_expr(p.pool, root)
            assert rc == 0 and int(v) == want


def _one_col_relation(values):
    desc = make_desc([(capi.FLOAT8OID, 8, 'd', 1, 1)])
    return desc, po.build_pages(desc, [[x] for x in values])


def test_float8_accum_sequences():
    for seq, errs, state in K["accum"]:
        vals = [b2f(x) for x in seq[:len(errs)]]
        desc, pages = _one_col_relation(vals)
        p = ExprPool()
        x = p.var(1, capi.FLOAT8OID)
        scan = capi.make_scan(desc, -1)
        agg = capi.make_agg(capi.AGGSTAGE_PARTIAL, [], [(capi.AGG_AVG_FLOAT8, x)])
        if errs[-1]:
            with pytest.raises(po.OracleError):
                po.seqscan_agg(scan, agg, p.pool, pages)
        else:
            rows, _, _ = po.seqscan_agg(scan, agg, p.pool, pages)
            got = [f2b(rows[0].agg[0].f[i]) for i in range(3)]
            want = [int(s) for s in state]
            for g, w in zip(got, want):
                assert g == w or (b2f(g) != b2f(g) and b2f(w) != b2f(w)), (vals, got, want)


def test_float8_combine_and_avg():
    agg = capi.make_agg(capi.AGGSTAGE_FINAL, [], [(capi.AGG_AVG_FLOAT8, -1)])
    for a, b, err, res in K["combine"]:
        rows = []
        for st in (a, b):
            r = capi.gg_aggrow()
            for i in range(3):
                r.agg[0].f[i] = b2f(st[i])
            rows.append(r)
        part = capi.make_agg(capi.AGGSTAGE_PARTIAL, [], [(capi.AGG_AVG_FLOAT8, -1)])
        # combine only (no final function): run the FINAL machinery and recompute avg = sumX / N
        if err:
            with pytest.raises(po.OracleError):
                po.agg_final(agg, rows)
            continue
        out = po.agg_final(agg, rows)
        n, sx = b2f(res[0]), b2f(res[1])
        if n == 0.0:
            assert out[0].agg[0].isnull
        elif sx == sx:
            want = sx / n
            assert out[0].agg[0].f[0] == want or (want != want and out[0].agg[0].f[0] != out[0].agg[0].f[0])
    for st, isnull, r in K["avg"]:
        row = capi.gg_aggrow()
        for i in range(3):
            row.agg[0].f[i] = b2f(st[i])
        out = po.agg_final(agg, [row])
        assert out[0].agg[0].isnull == isnull
        if not isnull:
            x, y = out[0].agg[0].f[0], b2f(r)
            # the FINAL stage combines into the initial state {0,0,0} first (0.0 + -0.0 = +0.0), so compare
            # numerically rather than bit-wise for signed zeros
            assert x == y or (x != x and y != y)


def test_int8pl_overflow_in_count_combine():
    agg = capi.make_agg(capi.AGGSTAGE_FINAL, [], [(capi.AGG_COUNT_STAR, -1)])
    for a, b, err, r in K["int8pl"]:
        rows = []
        for v in (a, b):
            row = capi.gg_aggrow()
            row.agg[0].i = int(v)
            rows.append(row)
        # count's FINAL state starts at 0, so the combine chain is 0 + a + b; skip cases where 0 + a + b
        # differs in overflow behaviour from a + b (it cannot: 0 + a never overflows)
        if err:
            with pytest.raises(po.OracleError):
                po.agg_final(agg, rows)
        else:
            assert po.agg_final(agg, rows)[0].agg[0].i == int(r)


def test_date_vs_timestamp():
    fids = [capi.F_DATE_LT_TIMESTAMP, capi.F_DATE_LE_TIMESTAMP, capi.F_DATE_EQ_TIMESTAMP,
            capi.F_DATE_GT_TIMESTAMP, capi.F_DATE_GE_TIMESTAMP, capi.F_DATE_NE_TIMESTAMP]
    for op, d, ts, err, r in K["date_ts"]:
        p = ExprPool()
        root = p.func(fids[op], capi.BOOLOID, p.const(capi.DATEOID, d), p.const(capi.TIMESTAMPOID, int(ts)))
        rc, v, isnull = po.eval_expr(p.pool, root)
        if err:
            assert rc != 0
        else:
            assert rc == 0 and int(v) == r, (op, d, ts)


def short_circuit_case():
    """`b <> 0 AND a / b > 1` and `b = 0 OR a / b > 1` over rows where b = 0: ExecEvalAnd / ExecEvalOr (execQual.c:3455,3385)
    never evaluate the division for those rows.  -> (desc, pages, [(scan, agg, pool, expected count)])"""
    from _util import make_desc
    desc = make_desc([(capi.FLOAT8OID, 8, 'd', 1, 1), (capi.FLOAT8OID, 8, 'd', 1, 1)])
    pages = po.build_pages(desc, [[4.0, 2.0], [1.0, 0.0], [9.0, 3.0], [5.0, 0.0], [1.0, 4.0]])
    plans = []
    for kind, first, want in ((capi.E_AND, capi.F_FLOAT8NE, 2), (capi.E_OR, capi.F_FLOAT8EQ, 4)):
        p = capi.ExprPool()
        a, b = p.var(1, capi.FLOAT8OID), p.var(2, capi.FLOAT8OID)
        guard = p.func(first, capi.BOOLOID, b, p.const(capi.FLOAT8OID, 0.0))
        ratio = p.func(capi.F_FLOAT8GT, capi.BOOLOID, p.func(capi.F_FLOAT8DIV, capi.FLOAT8OID, a, b), p.const(capi.FLOAT8OID, 1.0))
        plans.append((capi.make_scan(desc, p.boolop(kind, guard, ratio)), capi.make_agg(0, [], [(capi.AGG_COUNT_STAR, -1)]), p.pool, want))
    return desc, pages, plans


def test_and_or_skip_the_arm_that_would_raise():
    desc, pages, plans = short_circuit_case()
    for scan, agg, pool, want in plans:
        rows, sc, ps = po.seqscan_agg(scan, agg, pool, pages)
        assert (sc, ps, rows[0].agg[0].i) == (5, want, want)
    # with the arms swapped the division runs first and the ERROR is the reference's too
    p = capi.ExprPool()
    a, b = p.var(1, capi.FLOAT8OID), p.var(2, capi.FLOAT8OID)
    ratio = p.func(capi.F_FLOAT8GT, capi.BOOLOID, p.func(capi.F_FLOAT8DIV, capi.FLOAT8OID, a, b), p.const(capi.FLOAT8OID, 1.0))
    guard = p.func(capi.F_FLOAT8NE, capi.BOOLOID, b, p.const(capi.FLOAT8OID, 0.0))
    with pytest.raises(po.OracleError):
        po.seqscan_agg(capi.make_scan(desc, p.boolop(capi.E_AND, ratio, guard)), capi.make_agg(0, [], [(capi.AGG_COUNT_STAR, -1)]), p.pool, pages)
