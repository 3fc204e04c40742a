"""The product's synthetic relation loader (libgghost) writes byte-exact Greengage heap pages: every tuple
re-forms identically under the oracle's heap_form_tuple restatement (itself pinned to the reference's
heaptuple.o), pages follow PageAddItem's layout, and generation is deterministic and shardable."""
import ctypes as C

import numpy as np
import pytest

from greengage_b200 import capi, tpch
from oracle import pyoracle as po

H = capi.host_lib()


def _row(spec, c):
    vals, lens = (C.c_int64 * 32)(), (C.c_int32 * 32)()
    sb = C.create_string_buffer(256)
    mine = C.c_int(0)
    n = H.gg_synth_row(C.byref(spec), c, vals, lens, sb, 256, C.byref(mine))
    return n, vals, lens, sb, mine.value


@pytest.mark.parametrize("table", [capi.TAB_LINEITEM_WIDE, capi.TAB_LINEITEM_NARROW, capi.TAB_ORDERS])
def test_pages_match_oracle_tuple_format(table):
    spec = tpch.synth_spec(table, 3000)
    pages, nb, nr = tpch.synth_generate(spec, nthreads=2)
    assert nr == 3000
    desc = capi.synth_tupdesc(table)
    c = 0
    for blk in range(nb):
        pg = pages[blk * capi.GG_BLCKSZ:(blk + 1) * capi.GG_BLCKSZ]
        hdr = pg[:24].view(np.uint16)
        n = po.lib().or_page_nitems(pg.ctypes.data_as(C.c_void_p))
        assert hdr[5] & 4 and hdr[6] == 24 + 4 * n and hdr[8] == 32768 and hdr[9] == (32768 | 14)
        upper = 32768
        for i in range(n):
            lp = int(pg[24 + 4 * i:28 + 4 * i].view(np.uint32)[0])
            off, flags, ln = lp & 0x7FFF, (lp >> 15) & 3, lp >> 17
            assert flags == 1 and off % 8 == 0 and off == upper - ((ln + 7) & ~7)
            upper = off
            nn, vals, lens, sb, mine = _row(spec, c)
            assert mine
            row = []
            for a in range(nn):
                if desc.attrs[a].attlen == -1:
                    row.append(C.string_at(vals[a], lens[a]))
                elif desc.attrs[a].atttypid == capi.FLOAT8OID:
                    row.append(C.c_double.from_buffer_copy(C.c_int64(vals[a])).value)
                else:
                    row.append(int(vals[a]))
            want = bytearray(po.form_tuple(desc, row))
            got = bytes(pg[off:off + ln])
            # t_ctid = (block, offset) is stamped by the loader
            want[12:18] = got[12:18]
            assert got == bytes(want), (table, blk, i)
            assert got[12:18] == bytes([0, 0]) + int(blk).to_bytes(2, "little") + int(i + 1).to_bytes(2, "little")
            c += 1
        assert hdr[7] == upper
    assert c == nr


def test_rows_per_page_match_survey():
    for table, per_page in ((capi.TAB_LINEITEM_NARROW, 430), (capi.TAB_LINEITEM_WIDE, (186, 196)), (capi.TAB_ORDERS, (215, 230))):
        spec = tpch.synth_spec(table, 20000)
        pages, nb, nr = tpch.synth_generate(spec)
        n0 = po.lib().or_page_nitems(pages[:capi.GG_BLCKSZ].ctypes.data_as(C.c_void_p))
        if isinstance(per_page, tuple):
            assert per_page[0] <= n0 <= per_page[1], (table, n0)
        else:
            assert n0 == per_page


def test_deterministic_and_thread_independent():
    spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 150000)
    a, nba, _ = tpch.synth_generate(spec, nthreads=1)
    b, nbb, _ = tpch.synth_generate(spec, nthreads=5)
    assert nba == nbb and np.array_equal(a, b)


def test_segments_partition_the_table():
    total = 50000
    # DISTRIBUTED RANDOMLY
    rows = sum(tpch.synth_measure(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, total, nsegs=4, seg=s))[1] for s in range(4))
    assert rows == total
    # DISTRIBUTED BY (o_orderkey): every row sits on the segment cdbhash + jump hash sends its key to
    seen = 0
    for s in range(3):
        spec = tpch.synth_spec(capi.TAB_ORDERS, total, nsegs=3, seg=s, policy=capi.DIST_HASH)
        pages, nb, nr = tpch.synth_generate(spec)
        seen += nr
        desc = capi.synth_tupdesc(capi.TAB_ORDERS)
        for blk in (0, nb - 1):
            for row in po.deform_page(desc, pages, blk):
                t, v, ln, nu = (C.c_int32 * 1)(20), (C.c_int64 * 1)(row[0]), (C.c_int32 * 1)(0), (C.c_int32 * 1)(0)
                assert po.lib().or_route_datums(t, v, ln, nu, 1, 3) == s
    assert seen == total


def test_lineitem_orderkeys_reference_orders():
    spec = tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 5000, norders=1000)
    pages, nb, nr = tpch.synth_generate(spec)
    keys = {H.gg_synth_orderkey(o) for o in range(1000)}
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_NARROW)
    for row in po.deform_page(desc, pages, 0):
        assert row[0] in keys
