"""HeapTupleSatisfiesMVCC against a snapshot (tqual.c:997-1238): the oracle's restatement and the device's
rule (gg_device.cuh heap_tuple_satisfies_mvcc, compiled for the host by tests/emu) against tests/golden/mvcc_kat.json — answers
of the reference's own tqual.o + transam.o (oracle/ref_build/refwrap_tqual.c)."""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from greengage_b200 import capi
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
KAT = json.load(open(os.path.join(HERE, "golden", "mvcc_kat.json")))


def header(infomask, xmin, xmax, cid):
    """the 24 header bytes of a heap tuple (htup_details.h:139-162): xmin, xmax, t_cid, ctid, infomask2, infomask, hoff"""
    return struct.pack("<IIIHHHHHB", xmin, xmax, cid, 0, 0, 1, 8, infomask, 24) + b"\0"


def snapshot(s):
    return capi.make_snapshot(s["xmin"], s["xmax"], s["xip"], s["curcid"], s["own_xid"], s["clog_base"], bytes.fromhex(s["clog"]), s["clog_n"])


def device_words(s):
    """the snapshot as gg_engine_set_snapshot lays it out in device memory"""
    clog = bytes.fromhex(s["clog"])
    w = [s["xmin"], s["xmax"], len(s["xip"]), s["curcid"], s["own_xid"], s["clog_base"], s["clog_n"], 0] + list(s["xip"])
    raw = struct.pack("<%dI" % len(w), *w) + clog + b"\0" * (-len(clog) % 4)
    return (C.c_uint32 * (len(raw) // 4)).from_buffer_copy(raw)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_mvcc") / "libemu.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-I", os.path.join(HERE, "emu"), "-shared", "-o", so,
                           os.path.join(HERE, "emu", "device_emu.cpp"), os.path.join(ROOT, "greengage_b200", "csrc", "gg_compile.cpp")])
    L = C.CDLL(so)
    L.emu_tuple_satisfies_mvcc.argtypes = [C.c_char_p, C.c_void_p]
    return L


def test_oracle_rule_equals_the_reference():
    snaps = [snapshot(s) for s in KAT["snapshots"]]
    seen = set()
    for si, infomask, xmin, xmax, cid, want in KAT["cases"]:
        got = po.tuple_satisfies_mvcc(header(infomask, xmin, xmax, cid), snaps[si])
        assert got == want, (KAT["snapshots"][si], hex(infomask), xmin, xmax, cid)
        seen.add((want, bool(infomask & 0x100), bool(infomask & 0x800)))
    assert len(seen) >= 7              # visible and invisible answers on every side of the hint bits
    assert len(KAT["cases"]) > 3000


def test_device_rule_equals_the_reference(emu):
    words = [device_words(s) for s in KAT["snapshots"]]
    for si, infomask, xmin, xmax, cid, want in KAT["cases"]:
        assert emu.emu_tuple_satisfies_mvcc(header(infomask, xmin, xmax, cid), words[si]) == want, (si, hex(infomask), xmin, xmax, cid)


def test_what_only_the_server_can_decide_is_refused(emu):
    """multixact xmax, combo command ids, HEAP_MOVED_*, a status outside the given range or sub-committed: -1 (the oracle) /
    GGP_EF_VISIBILITY (the device) whenever the rule reaches the question — never a guess"""
    snaps = [snapshot(s) for s in KAT["snapshots"]]
    words = [device_words(s) for s in KAT["snapshots"]]
    refused = 0
    for si, infomask, xmin, xmax, cid, kind in KAT["unsupported"]:
        h = header(infomask, xmin, xmax, cid)
        o = po.tuple_satisfies_mvcc(h, snaps[si])
        assert emu.emu_tuple_satisfies_mvcc(h, words[si]) == o, (kind, hex(infomask))
        refused += o == -1
    assert refused > len(KAT["unsupported"]) // 3
    # TRANSACTION_STATUS_SUB_COMMITTED (clog.h:28): the parent's status decides (transam.c:146) — not known here
    s = dict(KAT["snapshots"][0], clog="ff" * 16)
    h = header(0x0802, s["clog_base"] + 5, 0, 0)
    assert po.tuple_satisfies_mvcc(h, snapshot(s)) == -1 and emu.emu_tuple_satisfies_mvcc(h, device_words(s)) == -1


def test_scan_with_a_snapshot_sees_what_the_rule_says():
    """or_scan_* with or_set_snapshot: pages that are not all-visible, tuples of committed / aborted / in-progress inserters
    and deleters (tests/_util.py MVCC_PATTERNS); an all-visible page skips the rule (heapam.c:391)"""
    from _util import make_desc, mvcc_snapshot, stamp_visibility
    desc = make_desc([(capi.INT4OID, 4, "i", 1, 1), (capi.FLOAT8OID, 8, "d", 1, 1)])
    rows = [[i, float(i)] for i in range(9000)]
    pg, vis = stamp_visibility(po.build_pages(desc, rows, [[False, False]] * len(rows)), all_visible_every=3)
    snap = mvcc_snapshot()
    p = capi.ExprPool()
    agg = capi.make_agg(0, [], [(capi.AGG_COUNT_STAR, -1)])
    po.set_snapshot(snap)
    try:
        got, sc, ps = po.seqscan_agg(capi.make_scan(desc, -1), agg, p.pool, pg)
    finally:
        po.set_snapshot(None)
    assert got[0].agg[0].i == sum(vis) and 0 < sum(vis) < len(rows) and len(vis) == len(rows)
    with pytest.raises(Exception):
        po.seqscan_agg(capi.make_scan(desc, -1), agg, p.pool, pg)            # without a snapshot the same pages are refused
