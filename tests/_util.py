"""Shared helpers for the test-suite: golden vectors, fixture relations, result comparison."""
import ctypes as C
import json
import os
import struct

import numpy as np

from greengage_b200 import capi, tpch
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def golden(name):
    return json.load(open(os.path.join(GOLD, name)))


def b2f(bits):
    return struct.unpack("<d", struct.pack("<q", int(bits)))[0]


def f2b(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def make_desc(spec):
    """spec: list of (typid, attlen, attalign, byval[, notnull])"""
    d = capi.gg_tupdesc()
    d.natts = len(spec)
    for i, sp in enumerate(spec):
        a = d.attrs[i]
        a.atttypid, a.attlen, a.attalign, a.attbyval = sp[0], sp[1], ord(sp[2]), sp[3]
        a.attnotnull = sp[4] if len(sp) > 4 else 0
        a.atttypmod = -1
    return d


_fixture_cache = {}


def lineitem_fixture_pages():
    """The reference's own regression lineitem data (tests/golden/lineitem_q1.npz) as LI-wide heap pages
    (float8 in place of numeric), built with the oracle's heap_form_tuple / PageAddItem restatement."""
    if "li" in _fixture_cache:
        return _fixture_cache["li"]
    zf = np.load(os.path.join(GOLD, "lineitem_q1.npz"))
    z = {k: zf[k] for k in zf.files}          # NpzFile decompresses on every access: materialise once
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    instr = [str(x) for x in z["shipinstruct_names"]]
    modes = [str(x) for x in z["shipmode_names"]]
    n = len(z["orderkey"])
    rows = []
    for i in range(n):
        rows.append([int(z["orderkey"][i]), int(z["partkey"][i]), int(z["suppkey"][i]), int(z["linenumber"][i]),
                     float(z["quantity"][i]), float(z["extendedprice"][i]), float(z["discount"][i]), float(z["tax"][i]),
                     bytes([z["returnflag"][i]]), bytes([z["linestatus"][i]]),
                     int(z["shipdate"][i]), int(z["commitdate"][i]), int(z["receiptdate"][i]),
                     instr[z["shipinstruct"][i]].ljust(25).encode(), modes[z["shipmode"][i]].ljust(10).encode(),
                     b"c" * int(z["comment_len"][i])])
    pages = po.build_pages(desc, rows)
    _fixture_cache["li"] = (desc, pages, n)
    return _fixture_cache["li"]


def rows_by_key(rows, nkeys=2):
    out = {}
    for r in rows:
        k = tuple((r.key[i], r.keylen[i], r.keyisnull[i]) for i in range(nkeys))
        assert k not in out, "duplicate group in output"
        out[k] = r
    return out


def assert_aggrows_match(got, want, agg, rel=1e-6, float_exact=False):
    """Integer results and keys bit-exact; float8 sums/avgs within `rel` (the tolerance BASELINE.json states)."""
    nkeys = agg.numCols
    g, w = rows_by_key(got, nkeys), rows_by_key(want, nkeys)
    assert set(g) == set(w), (sorted(g), sorted(w))
    for k in w:
        for i in range(agg.numAggs):
            a, b = g[k].agg[i], w[k].agg[i]
            assert a.isnull == b.isnull, (k, i, a.isnull, b.isnull)
            if a.isnull:
                continue
            fn = agg.aggs[i].aggfnoid
            if fn in (capi.AGG_COUNT_STAR, capi.AGG_COUNT_ANY, capi.AGG_SUM_INT4, capi.AGG_MAX_INT4, capi.AGG_MIN_INT4,
                      capi.AGG_MAX_INT8, capi.AGG_MIN_INT8, capi.AGG_MAX_DATE, capi.AGG_MIN_DATE):
                assert a.i == b.i, (k, i, a.i, b.i)
            else:
                nf = 3 if (fn == capi.AGG_AVG_FLOAT8 and agg.aggstage == capi.AGGSTAGE_PARTIAL) else 1
                for j in range(nf):
                    x, y = a.f[j], b.f[j]
                    if fn in (capi.AGG_MAX_FLOAT8, capi.AGG_MIN_FLOAT8) or float_exact or (nf == 3 and j == 0):
                        assert f2b(x) == f2b(y) or (x != x and y != y), (k, i, j, x, y)
                    elif y == 0 or y != y or abs(y) == float("inf"):
                        assert x == y or (x != x and y != y), (k, i, j, x, y)
                    else:
                        assert abs(x - y) <= rel * abs(y), (k, i, j, x, y)
