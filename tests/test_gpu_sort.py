"""GPU parity tests for Sort through the C-ABI (gg_sort_rows) against the oracle's comparator
(oracle/or_sort.c: inlineApplySortFunction + float8_cmp_internal + bpcharcmp).  mk_qsort is unstable, so the
contract is: the output is a permutation, adjacent rows compare <= 0 under the reference comparator, and the
sequence of sort keys equals the oracle's."""
import numpy as np
import pytest

from _util import f2b
from greengage_b200 import capi
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from greengage_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def check_sorted(keys, rows, nulls, perm):
    n, ncols = rows.shape
    assert sorted(perm.tolist()) == list(range(n))
    want = po.sort_perm(keys, ncols, rows, nulls)
    kc = [k.col for k in keys]
    nl = nulls if nulls is not None else np.zeros_like(rows, dtype=np.uint8)
    def norm(p):
        v = rows[p][:, kc].copy()
        m = nl[p][:, kc]
        for j, k in enumerate(keys):
            if k.typid == capi.FLOAT8OID:
                f = v[:, j].view(np.float64)
                v[:, j] = np.where(np.isnan(f), f2b(float("nan")), np.where(f == 0.0, 0, v[:, j]))
        v[m != 0] = 0
        return v, m
    gv, gm = norm(perm.astype(np.int64))
    wv, wm = norm(want.astype(np.int64))
    assert np.array_equal(gm, wm)
    assert np.array_equal(gv, wv)


def test_int64_keys_large(eng):
    from greengage_b200.engine import sort_rows
    rng = np.random.default_rng(1)
    n = 1_000_003                                    # not a multiple of the tile
    rows = np.stack([rng.integers(-2**62, 2**62, n), np.arange(n)], axis=1).astype(np.int64)
    keys = [capi.make_sortkey(0, capi.INT8OID)]
    perm = sort_rows(eng, keys, rows)
    assert sorted(perm.tolist()) == list(range(n))
    assert np.all(np.diff(rows[perm.astype(np.int64), 0]) >= 0)


@pytest.mark.parametrize("desc", [False, True])
@pytest.mark.parametrize("nulls_first", [False, True])
def test_multi_key_with_nulls_and_specials(eng, desc, nulls_first):
    from greengage_b200.engine import sort_rows
    rng = np.random.default_rng(7)
    n = 20_000
    specials = np.array([0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1.5, -1.5, 1e-300, -1e300])
    f = np.where(rng.random(n) < 0.3, rng.choice(specials, n), rng.normal(size=n) * 100).astype(np.float64)
    s = np.array([capi.pack_str("".join(rng.choice(list("ABC "), rng.integers(0, 4))), True)[0] for _ in range(n)], dtype=np.int64)
    i4 = rng.integers(-5, 5, n).astype(np.int64)
    d = rng.integers(-40000, 40000, n).astype(np.int64)
    rows = np.stack([i4, f.view(np.int64), s, d], axis=1)
    nulls = (rng.random((n, 4)) < 0.1).astype(np.uint8)
    keys = [capi.make_sortkey(0, capi.INT4OID, desc, nulls_first), capi.make_sortkey(2, capi.BPCHAROID, not desc, nulls_first),
            capi.make_sortkey(1, capi.FLOAT8OID, desc, not nulls_first), capi.make_sortkey(3, capi.DATEOID, False, nulls_first)]
    perm = sort_rows(eng, keys, rows, nulls)
    check_sorted(keys, rows, nulls, perm)


def test_q1_order_by(eng):
    """ORDER BY l_returnflag, l_linestatus over the four Q1 groups (rpt_tpch.source:307)."""
    from greengage_b200.engine import sort_rows
    rows = np.array([[capi.pack_str(a)[0], capi.pack_str(b)[0]] for a, b in (("R", "F"), ("N", "O"), ("A", "F"), ("N", "F"))], dtype=np.int64)
    keys = [capi.make_sortkey(0, capi.BPCHAROID), capi.make_sortkey(1, capi.BPCHAROID)]
    perm = sort_rows(eng, keys, rows)
    assert perm.tolist() == [2, 3, 1, 0]


def test_degenerate_inputs(eng):
    from greengage_b200.engine import sort_rows
    keys = [capi.make_sortkey(0, capi.INT8OID)]
    assert sort_rows(eng, keys, np.zeros((0, 1), dtype=np.int64)).size == 0
    same = np.full((5000, 1), 7, dtype=np.int64)
    perm = sort_rows(eng, keys, same)
    assert sorted(perm.tolist()) == list(range(5000))
    with pytest.raises(capi.GGError) as e:
        sort_rows(eng, [capi.make_sortkey(0, 1700)], same)          # numeric: not on the GPU path
    assert e.value.code == -6
