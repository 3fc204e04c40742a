"""Heap tuple format: the oracle's heap_form_tuple / slot_deform_tuple restatement against tuples formed
and deformed by the reference's own heaptuple.o (golden heap_kat.json): NULL bitmaps, 1-byte and 4-byte
(big-endian) varlena headers, alignment padding."""
import ctypes as C

import numpy as np

from _util import golden, make_desc
from greengage_b200 import capi
from oracle import pyoracle as po

K = golden("heap_kat.json")
L = po.lib()


def _desc(name):
    return make_desc([(t, l, al, bv) for t, l, al, bv in K["descs"][name]])


def _case_values(desc, case):
    vals, nulls = [], []
    for i, v in enumerate(case["values"]):
        a = desc.attrs[i]
        if v is None:
            vals.append(None); nulls.append(1); continue
        nulls.append(0)
        if a.attlen == -1:
            vals.append(bytes.fromhex(v))
        elif a.atttypid == capi.FLOAT8OID:
            vals.append(C.c_double.from_buffer_copy(C.c_int64(int(v))).value)
        else:
            vals.append(int(v))
    return vals, nulls


def test_form_matches_reference_bytes():
    for case in K["cases"]:
        desc = _desc(case["desc"])
        vals, nulls = _case_values(desc, case)
        mine = po.form_tuple(desc, vals, nulls)
        ref = bytes.fromhex(case["tuple"])
        assert len(mine) == len(ref)
        # bytes 0..17 are transaction fields (xmin/xmax/cid/ctid: the reference leaves a DatumTupleFields
        # header there, a stored tuple carries xmin etc.); from t_infomask2 on everything must agree except
        # the visibility hint bits the oracle stamps (HEAP_XMIN_FROZEN | HEAP_XMAX_INVALID)
        assert mine[18:20] == ref[18:20]
        im_m, im_r = int.from_bytes(mine[20:22], "little"), int.from_bytes(ref[20:22], "little")
        assert im_m & 0x000F == im_r & 0x000F          # HASNULL / HASVARWIDTH / HASEXTERNAL / HASOID
        assert im_m & 0x0B00 == 0x0B00
        assert mine[22:] == ref[22:], case["desc"]


def test_deform_matches_reference():
    for case in K["cases"]:
        desc = _desc(case["desc"])
        tup = np.frombuffer(bytes.fromhex(case["tuple"]), dtype=np.uint8).copy()
        n = desc.natts
        v, nu = (C.c_int64 * n)(), (C.c_uint8 * n)()
        L.or_heap_deform(C.byref(desc), tup.ctypes.data_as(C.c_void_p), n, v, nu)
        assert [int(x) for x in nu] == case["deform_null"]
        for i in range(n):
            if not nu[i]:
                assert int(v[i]) == int(case["deform"][i]), (case["desc"], i)


def test_page_add_item_layout():
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_NARROW)
    rows = [[i, 1.0 * i, 2.0, 0.05, 0.02, b"A", b"F", 100 + i] for i in range(1000)]
    pages = po.build_pages(desc, rows)
    nb = pages.size // capi.GG_BLCKSZ
    assert nb == 3                                     # 430 narrow rows per 32 KB page
    assert L.or_page_nitems(pages[:capi.GG_BLCKSZ].ctypes.data_as(C.c_void_p)) == 430
    hdr = pages[:24].view(np.uint16)
    assert hdr[5] == 0x0004 and hdr[6] == 24 + 4 * 430 and hdr[8] == 32768 and hdr[9] == (32768 | 14)
    lp0 = int(pages[24:28].view(np.uint32)[0])
    assert lp0 & 0x7FFF == 32768 - 72 and (lp0 >> 15) & 3 == 1 and lp0 >> 17 == 72
    back = po.deform_page(desc, pages, 2)
    assert len(back) == 1000 - 860 and back[0][0] == 860 and back[-1][7] == 1099
