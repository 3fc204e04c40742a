"""The node-level restatement (SeqScan -> qual -> hybrid HashAgg, two-stage combine) pinned by the
reference's own golden Q1 answer (src/test/regress/output/rpt_tpch.source:288-315) over the reference's
own regression data, re-expressed as float8 heap pages: count exact, sums/avgs within 1e-6 relative of the
exact numeric answers."""
from _util import assert_aggrows_match, golden, lineitem_fixture_pages
from greengage_b200 import capi, tpch
from oracle import pyoracle as po

EXP = golden("q1_expected.json")


def _check_against_golden(rows):
    got = {(capi.unpack_str(r.key[0], r.keylen[0]), capi.unpack_str(r.key[1], r.keylen[1])): r for r in rows}
    assert len(got) == 4
    for e in EXP["rows"]:
        r = got[(e["returnflag"], e["linestatus"])]
        assert r.agg[7].i == e["count_order"]
        for i, name in enumerate(["sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"]):
            want = float(e[name])
            assert abs(r.agg[i].f[0] - want) <= 1e-6 * abs(want), (name, r.agg[i].f[0], want)


def test_q1_single_stage_matches_reference_golden():
    desc, pages, n = lineitem_fixture_pages()
    assert n == EXP["nrows_loaded"] == 60175
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_NORMAL, interval_days=EXP["interval_days"], desc=desc)
    rows, scanned, passed = po.seqscan_agg(scan, agg, pool, pages)
    assert scanned == n and passed == sum(e["count_order"] for e in EXP["rows"])
    _check_against_golden(rows)


def test_q1_two_stage_three_segments_matches_reference_golden():
    """The reference's plan: partial HashAggregate per segment -> Redistribute -> final HashAggregate
    (expected/tpch500GB.out:1771-1782), here over 3 segments like the gpdemo cluster."""
    desc, pages, n = lineitem_fixture_pages()
    scan, part, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL, interval_days=EXP["interval_days"], desc=desc)
    nb = pages.size // capi.GG_BLCKSZ
    cuts = [0, nb // 3, 2 * nb // 3, nb]
    partial = []
    for s in range(3):
        rows, _, _ = po.seqscan_agg(scan, part, pool, pages[cuts[s] * capi.GG_BLCKSZ:cuts[s + 1] * capi.GG_BLCKSZ])
        partial.extend(rows)
    final = po.agg_final(tpch.q1_final_agg(part), partial)
    _check_against_golden(final)
    # and the single-stage answer agrees with the two-stage one to the last few ulps
    scan1, agg1, pool1 = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_NORMAL, interval_days=EXP["interval_days"], desc=desc)
    single, _, _ = po.seqscan_agg(scan1, agg1, pool1, pages)
    assert_aggrows_match(final, single, agg1, rel=1e-12)


def test_streaming_bottom_stage_duplicates_are_tolerated():
    """A streaming partial stage may emit the same group twice (execHHashagg.c:996-1002); FINAL combines them."""
    desc, pages, n = lineitem_fixture_pages()
    scan, part, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL, interval_days=108, desc=desc)
    half = (pages.size // capi.GG_BLCKSZ // 2) * capi.GG_BLCKSZ
    a, _, _ = po.seqscan_agg(scan, part, pool, pages[:half])
    b, _, _ = po.seqscan_agg(scan, part, pool, pages[half:])
    _check_against_golden(po.agg_final(tpch.q1_final_agg(part), a + b))


def test_count_star_two_segment_plumbing():
    """BASELINE config 0: SELECT count(*) over a 2-segment table: SeqScan -> partial Agg -> Gather -> final Agg."""
    import ctypes as C
    import numpy as np
    segs = []
    total = 0
    for s in range(2):
        spec = tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 100000, nsegs=2, seg=s)
        pages, nb, nr = tpch.synth_generate(spec)
        segs.append((pages, nb))
        total += nr
    assert total == 100000
    ptrs = (C.c_void_p * 2)(*[p.ctypes.data for p, _ in segs])
    nbs = (C.c_uint64 * 2)(*[nb for _, nb in segs])
    assert po.lib().or_count_star_2stage(ptrs, nbs, 2) == 100000


def test_q6_revenue_is_the_references():
    """A second golden answer of the same suite for the scan -> qual -> aggregate path: Q6 ('mpph6', a plain aggregate
    behind five range predicates)."""
    from _util import Q6_GOLDEN_REVENUE, lineitem_fixture_pages, tpch_q6_plan
    from oracle import pyoracle as po
    desc, pages, n = lineitem_fixture_pages()
    scan, agg, pool = tpch_q6_plan(desc)
    rows, sc, ps = po.seqscan_agg(scan, agg, pool, pages)
    assert sc == n and len(rows) == 1 and 0 < ps < n
    assert abs(rows[0].agg[0].f[0] - Q6_GOLDEN_REVENUE) <= 1e-6 * Q6_GOLDEN_REVENUE, rows[0].agg[0].f[0]


def test_onek_aggregates_are_the_references():
    """Integer aggregates and a hashed GROUP BY on an int4 key, against expected/aggregates.out over the suite's onek table."""
    from _util import onek_check, onek_fixture, onek_plans
    from oracle import pyoracle as po
    from test_compile import disasm
    desc, pages, exp = onek_fixture()
    plain, grouped = onek_plans(desc, exp)
    onek_check(exp, po.seqscan_agg(*plain, pages)[0], po.seqscan_agg(*grouped, pages)[0])
    assert disasm(*plain) and disasm(*grouped)               # and the device compiler takes both plans


def test_gp_hashagg_text_key_answer_is_the_references():
    from _util import gp_hashagg_case
    from oracle import pyoracle as po
    from test_compile import disasm
    desc, pages, scan, agg, pool, want = gp_hashagg_case()
    rows, sc, ps = po.seqscan_agg(scan, agg, pool, pages)
    assert (sc, ps) == (6, 6)
    assert {capi.unpack_str(r.key[0], r.keylen[0]): r.agg[0].i for r in rows} == want
    assert disasm(scan, agg, pool)
