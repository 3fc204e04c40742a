"""The oracle's Sort comparator (oracle/or_sort.c, restating inlineApplySortFunction + the btree comparison functions)
pinned against what the reference's own objects computed: btfloat8cmp for every pair of the float golden vectors
(tests/golden/float_kat.json, column 6: NaN = NaN, NaN above everything, -0 = +0), plus the NULLS FIRST/LAST and DESC rules of
tuplesort_mk.c:2816-2850 and C-locale bpchar ordering on packed strings."""
import ctypes as C
import itertools

import numpy as np

from _util import golden
from greengage_b200 import capi
from oracle import pyoracle as po


def compare(keys, a, b, an=None, bn=None):
    L = po.lib()
    L.or_sort_compare.argtypes = [C.POINTER(capi.gg_sortkey), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ka = (capi.gg_sortkey * len(keys))(*keys)
    n = len(a)
    av, bv = (C.c_int64 * n)(*a), (C.c_int64 * n)(*b)
    anv = (C.c_uint8 * n)(*(an or [0] * n))
    bnv = (C.c_uint8 * n)(*(bn or [0] * n))
    return L.or_sort_compare(ka, len(keys), n, av, anv, bv, bnv)


def sign(x):
    return (x > 0) - (x < 0)


def test_float8_order_is_the_reference_btfloat8cmp():
    kat = golden("float_kat.json")["cmp"]
    assert len(kat) >= 400
    asc, desc = [capi.make_sortkey(0, capi.FLOAT8OID)], [capi.make_sortkey(0, capi.FLOAT8OID, desc=True)]
    for a, b, eq, lt, le, cmp3 in kat:
        assert sign(compare(asc, [int(a)], [int(b)])) == cmp3
        assert sign(compare(desc, [int(a)], [int(b)])) == -cmp3


def test_nulls_first_last_and_desc():
    for desc, nf in itertools.product((False, True), (False, True)):
        k = [capi.make_sortkey(0, capi.INT8OID, desc, nf)]
        assert sign(compare(k, [5], [9])) == (1 if desc else -1)
        assert sign(compare(k, [0], [9], [1], [0])) == (-1 if nf else 1)        # NULL vs value: only nulls_first decides
        assert sign(compare(k, [9], [0], [0], [1])) == (1 if nf else -1)
        assert compare(k, [1], [2], [1], [1]) == 0                              # NULL = NULL for ordering
    # PostgreSQL defaults: ASC -> NULLS LAST, DESC -> NULLS FIRST
    assert capi.make_sortkey(0, capi.INT8OID).nulls_first == 0 and capi.make_sortkey(0, capi.INT8OID, desc=True).nulls_first == 1


def test_multi_key_and_types():
    keys = [capi.make_sortkey(0, capi.BPCHAROID), capi.make_sortkey(1, capi.INT4OID, desc=True), capi.make_sortkey(2, capi.DATEOID)]
    s = lambda t: capi.pack_str(t)[0]
    assert sign(compare(keys, [s("A"), 1, 0], [s("B"), 9, 0])) == -1           # first key decides
    assert sign(compare(keys, [s("AB"), 1, 0], [s("A"), 1, 0])) == 1           # shorter string first (memcmp then length)
    assert sign(compare(keys, [s("A "), 1, 0], [s("A"), 1, 0])) == 0           # bpchar: trailing blanks do not count
    assert sign(compare(keys, [s("A"), 1, 0], [s("A"), 9, 0])) == 1            # second key DESC
    assert sign(compare(keys, [s("A"), 1, -5], [s("A"), 1, 7])) == -1          # third key: signed dates
    # int4 columns compare as 32-bit signed values
    assert sign(compare([capi.make_sortkey(0, capi.INT4OID)], [-1], [1])) == -1


def test_sort_perm_is_sorted_under_the_comparator():
    rng = np.random.default_rng(4)
    n = 3000
    f = np.where(rng.random(n) < 0.2, rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0], n), rng.normal(size=n))
    rows = np.stack([rng.integers(0, 5, n), f.view(np.int64)], axis=1).astype(np.int64)
    nulls = (rng.random((n, 2)) < 0.1).astype(np.uint8)
    keys = [capi.make_sortkey(0, capi.INT8OID, True), capi.make_sortkey(1, capi.FLOAT8OID)]
    perm = po.sort_perm(keys, 2, rows, nulls).astype(np.int64)
    assert sorted(perm.tolist()) == list(range(n))
    for i in range(n - 1):
        a, b = perm[i], perm[i + 1]
        assert compare(keys, rows[a].tolist(), rows[b].tolist(), nulls[a].tolist(), nulls[b].tolist()) <= 0


def test_order_by_answers_are_the_references():
    """Golden ORDER BY outputs of the reference's sort regression test (expected/sort.out: int8, char, date, float8, int4
    columns ASC and DESC; text COLLATE "C" with NULLS LAST and NULLS FIRST)."""
    from _util import sort_golden_cases
    cases = sort_golden_cases()
    assert len(cases) == 12
    for name, keys, rows, nulls, want, wantnulls in cases:
        perm = po.sort_perm(keys, 1, rows, nulls).astype(np.int64)
        assert np.array_equal(rows[perm], want) and np.array_equal(nulls[perm], wantnulls), name
