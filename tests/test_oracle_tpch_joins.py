"""The oracle's Hash / HashJoin restatement against the REFERENCE'S OWN golden answers: the two TPC-H queries of its
regression suite that are a lineitem-orders hash join under an aggregate — Q4 (semi join) and Q12 (inner join) —
over its own regression data (src/test/regress/output/rpt_tpch.source 'mpph4' / 'mpph12'; tests/golden/make_golden.py)."""
from _util import golden, tpch_join_fixture, tpch_q4_plan, tpch_q12_plan
from greengage_b200 import capi
from oracle import pyoracle as po


def test_q4_semi_join_counts_are_the_references():
    li_desc, li_pages, ord_desc, ord_pages, shipmode_code, priority_code = tpch_join_fixture()
    exp = golden("tpch_join_expected.json")["q4"]
    outer, inner, hj, agg, pool = tpch_q4_plan(li_desc, ord_desc, exp)
    rows, nj = po.hashjoin_agg(outer, inner, hj, agg, pool, ord_pages, li_pages)
    got = {capi.unpack_str(r.key[0], r.keylen[0]): r.agg[0].i for r in rows}
    assert got == {priority_code[e["orderpriority"]]: e["order_count"] for e in exp["rows"]}
    assert nj == sum(e["order_count"] for e in exp["rows"])


def test_q12_inner_join_counts_are_the_references():
    li_desc, li_pages, ord_desc, ord_pages, shipmode_code, priority_code = tpch_join_fixture()
    exp = golden("tpch_join_expected.json")["q12"]
    for high, col in ((True, "high_line_count"), (False, "low_line_count")):
        outer, inner, hj, agg, pool = tpch_q12_plan(li_desc, ord_desc, exp, shipmode_code, priority_code, high)
        rows, nj = po.hashjoin_agg(outer, inner, hj, agg, pool, li_pages, ord_pages)
        got = {capi.unpack_str(r.key[0], r.keylen[0]): r.agg[0].i for r in rows}
        assert got == {shipmode_code[e["shipmode"]]: e[col] for e in exp["rows"]}, (col, got)
