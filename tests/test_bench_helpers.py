"""bench.py's parity helpers on CPU: per-segment one-stage oracle answers combine into the whole table's answer, and the
comparison flags what it should."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from greengage_b200 import capi, tpch  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def _oracle(pages, threads=2):
    scan, part, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL)
    rows, secs, scanned = po.seqscan_agg_mt(scan, part, tpch.q1_final_agg(part), pool, pages, threads)
    return bench.q1_rows_from_oracle(rows), scanned


def test_segment_answers_combine_to_the_whole_tables_answer():
    whole, _, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 60_000), nthreads=2)
    want, scanned = _oracle(whole)
    assert scanned == nr
    parts = []
    for seg in range(3):
        pages, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 60_000, nsegs=3, seg=seg), nthreads=2)
        parts.append(_oracle(pages)[0])
    got = bench.q1_combine(parts)
    par = bench.q1_compare(got, want)
    assert par["ok"] and par["counts_equal"] and par["rows"] == 4 and par["max_rel_err"] < 1e-9


def test_compare_flags_a_wrong_count_and_a_wrong_sum():
    pages, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 20_000), nthreads=2)
    want, _ = _oracle(pages)
    bad = copy.deepcopy(want)
    k = next(iter(bad))
    bad[k] = (bad[k][0], bad[k][1] + 1)
    assert not bench.q1_compare(bad, want)["ok"]
    bad = copy.deepcopy(want)
    bad[k][0][2] *= 1.00001
    assert not bench.q1_compare(bad, want)["ok"]
    bad = copy.deepcopy(want)
    del bad[k]
    assert not bench.q1_compare(bad, want)["ok"]


def test_slot_rows_decode_like_oracle_rows():
    import numpy as np
    f = lambda d: int(np.float64(d).view(np.int64))
    rows = [([65, 70, f(1.5), f(2.5), f(3.5), f(4.5), f(5.5), f(6.5), f(7.5), 42], [0] * 10, [0] * 10, [0] * 10)]
    got = bench.q1_rows_from_slots(rows)
    assert got == {(65, 70): ([1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5], 42)}
