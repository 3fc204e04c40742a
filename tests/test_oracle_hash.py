"""The oracle's hashing and segment routing against golden vectors computed by the reference's own
hashfunc.o / varchar.o / cdbhash.o (tests/golden/make_golden.py), and against those objects directly
when oracle/_ref/libggref.so is present.  The product's host-side routing (libgghost) is held to the
same vectors."""
import ctypes as C
import random

from _util import golden
from greengage_b200 import capi
from oracle import pyoracle as po

K = golden("hash_kat.json")
L = po.lib()


def test_hash_any_golden():
    for hexs, want in K["hash_any"]:
        b = bytes.fromhex(hexs)
        assert L.or_hash_any(b, len(b)) == want
        assert capi.host_lib().gg_hash_any(b, len(b)) == want


def test_scalar_hashes_golden():
    for v, want in K["hash_uint32"]:
        assert L.or_hash_uint32(v) == want
    for v, want in K["hashint4"]:
        assert L.or_hashint4(v) == want
    for v, want in K["hashint8"]:
        assert L.or_hashint8(int(v)) == want
    for bits, want in K["hashfloat8"]:
        assert L.or_hashfloat8(C.c_double.from_buffer_copy(C.c_int64(int(bits))).value) == want


def test_known_answers_from_survey():
    # values obtained from the reference's hashfunc.o
    assert L.or_hash_uint32(0) == 4022255791 and L.or_hash_uint32(1) == 2389907270 and L.or_hash_uint32(42) == 1509752520
    for s, want in ((b"A", 1656725486), (b"N", 1706742859), (b"R", 4055972430), (b"F", 1874189369), (b"O", 2962905310)):
        assert L.or_hash_any(s, 1) == want
    assert L.or_hashfloat8(1.0) == 376496956


def test_bpchar_golden():
    for hexs, want in K["hashbpchar"]:
        b = bytes.fromhex(hexs)
        assert L.or_hashbpchar(b, len(b)) == want
    for a, b, want in K["bpchareq"]:
        a, b = bytes.fromhex(a), bytes.fromhex(b)
        assert L.or_bpchareq(a, len(a), b, len(b)) == want


def test_routing_golden_oracle_and_product():
    H = capi.host_lib()
    for r in K["route"]:
        n = len(r["typ"])
        t = (C.c_int32 * n)(*r["typ"])
        v = (C.c_int64 * n)(*[int(x) for x in r["val"]])
        ln = (C.c_int32 * n)(*r["len"])
        nu = (C.c_int32 * n)(*r["null"])
        assert L.or_route_datums(t, v, ln, nu, n, r["nsegs"]) == r["seg"]
        assert H.gg_cdbhash_route(t, v, ln, nu, n, r["nsegs"]) == r["seg"]


def test_against_reference_objects_when_built():
    R = po.ref_lib()
    if R is None:
        import pytest
        pytest.skip("oracle/_ref not built (no /root/reference on this box); golden vectors cover it")
    rng = random.Random(7)
    for _ in range(20000):
        n = rng.randint(0, 48)
        b = bytes(rng.getrandbits(8) for _ in range(n))
        assert R.ref_hash_any(b, n) == L.or_hash_any(b, n)
        v = rng.getrandbits(64) - (1 << 63)
        assert R.ref_hashint8(v) == L.or_hashint8(v)
        ns = rng.choice([1, 2, 3, 5, 8, 13, 64, 999])
        t, vv, ln, nu = (C.c_int32 * 1)(20), (C.c_int64 * 1)(v), (C.c_int32 * 1)(0), (C.c_int32 * 1)(0)
        assert R.ref_cdbhash_route(t, vv, ln, nu, 1, ns) == L.or_route_datums(t, vv, ln, nu, 1, ns)


def test_bulk_routing_of_aggregate_rows_matches_the_oracle():
    """greengage_b200.motion.route_rows_raw (one C call over a gg_aggrow array, what bench.py's Redistribute uses) against
    the oracle's cdbhash restatement, row by row: int, float8 (incl. -0), packed-string and NULL keys."""
    import ctypes as C
    import numpy as np
    from greengage_b200 import capi, motion
    from oracle import pyoracle as po
    rng = np.random.default_rng(31)
    n = 500
    rows = (capi.gg_aggrow * n)()
    typids = [capi.INT8OID, capi.BPCHAROID, capi.FLOAT8OID, capi.INT4OID]
    for r in rows:
        r.key[0] = int(rng.integers(-2**40, 2**40))
        s, ln = capi.pack_str("".join(rng.choice(list("ABCxyz"), int(rng.integers(0, 6)))))
        r.key[1], r.keylen[1] = s, ln
        f = float(rng.choice([0.0, -0.0, 1.5, -2.25, 1e300, float(rng.normal())]))
        r.key[2] = np.float64(f).view(np.int64).item()
        r.key[3] = int(rng.integers(-1000, 1000))
        for c in range(4):
            r.keyisnull[c] = int(rng.random() < 0.15)
    buf = np.frombuffer(rows, dtype=np.uint8)
    for nsegs in (1, 2, 3, 8, 64):
        got = motion.route_rows_raw(buf, n, typids, nsegs)
        t = (C.c_int32 * 4)(*typids)
        for i, r in enumerate(rows):
            v = (C.c_int64 * 4)(*[r.key[c] for c in range(4)])
            ln = (C.c_int32 * 4)(*[r.keylen[c] for c in range(4)])
            nn = (C.c_int32 * 4)(*[r.keyisnull[c] for c in range(4)])
            assert got[i] == po.lib().or_route_datums(t, v, ln, nn, 4, nsegs), (i, nsegs)
