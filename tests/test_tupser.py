"""MemTuple and the tuple-chunk wire format (include/gg_tupser.h, libgghost.so) held to the reference itself:
tests/golden/memtuple_kat.json was written by the reference's own memtuple.o / tupser.o / tupchunklist.o
(oracle/ref_build/refwrap_motion.c; generator tests/golden/make_golden.py memtuple_kat).  Byte for byte:
bindings, formed tuples, what deform reads back, the chunks SerializeTuple emits at several chunk sizes, and reading the
reference's chunks — both the MemTuple form and the heap-tuple (TupSerHeader) form."""
import ctypes as C
import hashlib
import json
import os
import struct

import pytest

from greengage_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "memtuple_kat.json")))
MAXA = 64


class AttBind(C.Structure):
    _fields_ = [("offset", C.c_int32), ("len", C.c_int16), ("len_aligned", C.c_int16), ("flag", C.c_uint8), ("null_byte", C.c_uint8),
                ("null_mask", C.c_uint8), ("phys", C.c_uint8)]


class Layout(C.Structure):
    _fields_ = [("att", AttBind * MAXA), ("var_start", C.c_int32), ("pad", C.c_int32)]


class Binding(C.Structure):
    _fields_ = [("natts", C.c_int32), ("column_align", C.c_int32), ("null_bitmap_extra", C.c_int32), ("pad", C.c_int32),
                ("attrs", capi.gg_attr * MAXA), ("small", Layout), ("large", Layout)]


def lib():
    L = capi.host_lib()
    L.gg_tupser_serialize.restype = C.c_int64
    L.gg_memtuple_size.restype = C.c_uint32
    return L


def binding(name):
    cols = KAT["descs"][name]["cols"]
    attrs = (capi.gg_attr * len(cols))()
    for i, (t, l, al, bv) in enumerate(cols):
        attrs[i].atttypid, attrs[i].attlen, attrs[i].attalign, attrs[i].attbyval, attrs[i].atttypmod = t, l, ord(al), bv, -1
    b = Binding()
    assert lib().gg_memtuple_bind(attrs, len(cols), C.byref(b)) == 0
    return b, cols


def row_arrays(cols, pyvals):
    n = len(cols)
    vals, lens, nulls, ptrs, keep = (C.c_int64 * n)(), (C.c_int32 * n)(), (C.c_uint8 * n)(), (C.c_void_p * n)(), []
    for i, ((t, l, al, bv), v) in enumerate(zip(cols, pyvals)):
        if v is None:
            nulls[i] = 1
        elif l == -1:
            if v.startswith("x"):
                ln, byte = v[1:].split(":")
                pay = bytes([int(byte)]) * int(ln)
            else:
                pay = bytes.fromhex(v)
            if len(pay) <= 8 and t != 1022 and i % 2 == 0:
                vals[i] = int.from_bytes(pay.ljust(8, b"\0"), "little", signed=True)      # the executor's packed short strings
            else:
                buf = C.create_string_buffer(pay, max(len(pay), 1))
                keep.append(buf)
                ptrs[i] = C.addressof(buf)
            lens[i] = len(pay)
        else:
            vals[i] = int(v)
    return vals, lens, nulls, ptrs, keep


@pytest.mark.parametrize("name", sorted(KAT["descs"]))
def test_binding_equals_the_references(name):
    b, cols = binding(name)
    meta = KAT["descs"][name]
    assert (b.column_align, b.null_bitmap_extra) == (meta["column_align"], meta["null_bitmap_extra"])
    for which, lay in (("small", b.small), ("large", b.large)):
        want = meta["bind"][which]
        assert lay.var_start == want["var_start"]
        for i, w in enumerate(want["att"]):
            a = lay.att[i]
            assert [a.offset, a.len, a.len_aligned, a.flag, a.null_byte, a.null_mask] == w, (which, i)


def test_formed_tuples_are_the_references_byte_for_byte():
    L = lib()
    nbig = 0
    for case in KAT["cases"]:
        b, cols = binding(case["desc"])
        vals, lens, nulls, ptrs, keep = row_arrays(cols, case["values"])
        ln = C.c_uint32(0)
        out = (C.c_uint8 * 200000)()
        assert L.gg_memtuple_form(C.byref(b), vals, nulls, lens, ptrs, out, len(out), C.byref(ln)) == 0
        got = bytes(out[:ln.value])
        if "memtuple" in case:
            assert got.hex() == case["memtuple"], case["desc"]
        else:
            assert ln.value == case["memtuple_len"] and hashlib.sha1(got).hexdigest() == case["memtuple_sha1"]
            large = bool(struct.unpack("<I", got[:4])[0] & 2)               # the large (4-byte offset) layout ...
            assert large == (ln.value > 0xFFF0)                             # ... exactly beyond MEMTUPLE_LEN_FITSHORT
            nbig += large
        assert L.gg_memtuple_size(out) == ln.value
        # a buffer that is too small reports the length it needs
        need = C.c_uint32(0)
        assert L.gg_memtuple_form(C.byref(b), vals, nulls, lens, ptrs, out, 4, C.byref(need)) == -8       # GG_ERR_NOMEM
        assert need.value == ln.value
    assert nbig >= 1


def test_deform_reads_back_what_the_reference_reads():
    L = lib()
    for case in KAT["cases"]:
        if "memtuple" not in case:
            continue
        b, cols = binding(case["desc"])
        mt = bytes.fromhex(case["memtuple"])
        n = len(cols)
        v, nl, ln = (C.c_int64 * n)(), (C.c_uint8 * n)(), (C.c_int32 * n)()
        assert L.gg_memtuple_deform(C.byref(b), mt, len(mt), v, nl, ln) == 0
        assert [int(x) for x in nl] == case["deform_null"]
        for i, (t, l, al, bv) in enumerate(cols):
            if nl[i]:
                continue
            if l == -1:
                # the reference returns a pointer to the datum (header included); ours points at the payload
                hdr = 1 if mt[int(case["deform"][i])] & 0x80 else 4
                assert v[i] == int(case["deform"][i]) + hdr
                want = case["values"][i]
                assert mt[v[i]:v[i] + ln[i]].hex() == want
            else:
                assert v[i] == int(case["deform"][i]), (case["desc"], i)
        # truncated tuples are refused, not read past
        assert L.gg_memtuple_deform(C.byref(b), mt[:len(mt) - 8], len(mt) - 8, v, nl, ln) != 0


def test_chunks_are_the_references_and_come_back_as_the_row():
    L = lib()
    seen_partial = 0
    for case in KAT["cases"]:
        if "chunks" not in case:
            continue
        b, cols = binding(case["desc"])
        vals, lens, nulls, ptrs, keep = row_arrays(cols, case["values"])
        n = len(cols)
        for mc, (want_hex, want_n) in case["chunks"].items():
            out = (C.c_uint8 * 40000)()
            nch = C.c_int32(0)
            got = L.gg_tupser_serialize(C.byref(b), vals, nulls, lens, ptrs, int(mc), out, len(out), C.byref(nch))
            assert got > 0 and bytes(out[:got]).hex() == want_hex and nch.value == want_n, (case["desc"], mc)
            seen_partial += want_n > 1
            # and back: reassembly + deform
            v, nl, ln, sb = (C.c_int64 * n)(), (C.c_uint8 * n)(), (C.c_int32 * n)(), (C.c_uint8 * 8192)()
            used = C.c_uint64(0)
            assert L.gg_tupser_deserialize(C.byref(b), out, got, C.byref(used), v, nl, ln, sb, len(sb)) == 0
            assert used.value == got
            check_row(cols, case["values"], v, nl, ln, bytes(sb))
        # the heap-tuple form the reference sends for a tuple straight off a heap page
        hc = bytes.fromhex(case["heap_chunks"])
        v, nl, ln, sb = (C.c_int64 * n)(), (C.c_uint8 * n)(), (C.c_int32 * n)(), (C.c_uint8 * 8192)()
        used = C.c_uint64(0)
        assert L.gg_tupser_deserialize(C.byref(b), hc, len(hc), C.byref(used), v, nl, ln, sb, len(sb)) == 0
        assert used.value == len(hc)
        check_row(cols, case["values"], v, nl, ln, bytes(sb))
    assert seen_partial > 50


def check_row(cols, pyvals, v, nl, ln, strbuf):
    for i, ((t, l, al, bv), want) in enumerate(zip(cols, pyvals)):
        if want is None:
            assert nl[i] == 1
            continue
        assert nl[i] == 0
        if l == -1:
            assert strbuf[v[i]:v[i] + ln[i]].hex() == want
        elif l == 1:
            assert (v[i] & 1) == (int(want) & 1)
        else:
            assert v[i] == int(want)


def test_end_of_stream_empty_rows_and_malformed_chunks():
    L = lib()
    b, cols = binding("ints4")
    out = (C.c_uint8 * 64)()
    assert L.gg_tupser_eos(out, 64) == 4 and bytes(out[:4]) == b"\x00\x00\x04\x00"
    n = len(cols)
    v, nl, ln, sb = (C.c_int64 * n)(), (C.c_uint8 * n)(), (C.c_int32 * n)(), (C.c_uint8 * 64)()
    used = C.c_uint64(0)
    assert L.gg_tupser_deserialize(C.byref(b), out, 4, C.byref(used), v, nl, ln, sb, 64) == 1 and used.value == 4
    bad = bytes([200, 0, 0, 0]) + bytes(16)                 # a WHOLE chunk claiming 200 bytes in a 20-byte buffer
    assert L.gg_tupser_deserialize(C.byref(b), bad, len(bad), C.byref(used), v, nl, ln, sb, 64) < 0
    bad = bytes([8, 0, 2, 0]) + bytes(8)                     # a PARTIAL_MID chunk cannot start a tuple
    assert L.gg_tupser_deserialize(C.byref(b), bad, len(bad), C.byref(used), v, nl, ln, sb, 64) < 0


def test_float8_array_state_is_the_references_array_layout():
    L = lib()
    out = (C.c_uint8 * 44)()
    assert L.gg_float8_array3(C.c_double(3.0), C.c_double(1.5), C.c_double(9.25), out) == 44
    assert bytes(out) == struct.pack("<iiIii3d", 1, 0, 701, 3, 1, 3.0, 1.5, 9.25)
    back = (C.c_double * 3)()
    assert L.gg_float8_array3_read(out, 44, back) == 0 and list(back) == [3.0, 1.5, 9.25]


# ---- the slot's MemTuple form at the node surface (GgExecFetchSlotMemTuple / GgExecStoreMemTuple, libggexec.so) ----

def _exec_lib():
    from greengage_b200 import executor as ex
    L = ex.exec_lib() if hasattr(ex, "exec_lib") else C.CDLL(os.path.join(os.path.dirname(capi.__file__), "libggexec.so"))
    L.GgExecFetchSlotMemTuple.restype = C.c_int64
    L.GgExecFetchSlotMemTuple.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    L.GgExecStoreMemTuple.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_char_p, C.c_uint32]
    return L


def _slot_type():
    from greengage_b200 import executor as ex
    for name in ("GgTupleTableSlot", "gg_slot", "Slot"):
        if hasattr(ex, name):
            return getattr(ex, name)
    raise AssertionError("executor.py has no slot structure")


def test_a_slots_memtuple_is_the_references_byte_for_byte_and_comes_back():
    """ExecFetchSlotMemTuple / ExecStoreMemTuple for the node surface's virtual slots: over the reference-written cases whose
    columns are slot types (int4, date, short bpchar) the formed MemTuple is the reference's, and storing it back gives the slot."""
    L = _exec_lib()
    Slot = _slot_type()
    used = 0
    for case in KAT["cases"]:
        cols = KAT["descs"][case["desc"]]["cols"]
        if case["desc"] != "ints4" or "memtuple" not in case:
            continue
        pay = [None if v is None else (bytes.fromhex(v) if not str(v).startswith("x") else bytes([int(v[1:].split(":")[1])]) * int(v[1:].split(":")[0]))
               if l == -1 else int(v) for (t, l, al, bv), v in zip(cols, case["values"])]
        if any(isinstance(p, bytes) and len(p) > 8 for p in pay):
            continue
        s = Slot()
        s.tts_nvalid, s.tts_isempty = len(cols), 0
        for i, ((t, l, al, bv), p) in enumerate(zip(cols, pay)):
            s.tts_typid[i] = t
            if p is None:
                s.tts_isnull[i] = 1
            elif isinstance(p, bytes):
                s.tts_values[i] = int.from_bytes(p.ljust(8, b"\0"), "little", signed=True)
                s.tts_len[i] = len(p)
            else:
                s.tts_values[i] = p
        out = (C.c_uint8 * 4096)()
        need = C.c_uint32(0)
        n = L.GgExecFetchSlotMemTuple(C.byref(s), out, len(out), C.byref(need))
        assert n > 0 and need.value == n and bytes(out[:n]).hex() == case["memtuple"], case
        assert L.GgExecFetchSlotMemTuple(C.byref(s), out, 4, C.byref(need)) == -8 and need.value == n      # too small: says how much
        back = Slot()
        typids = (C.c_int32 * len(cols))(*[c[0] for c in cols])
        assert L.GgExecStoreMemTuple(C.byref(back), typids, len(cols), bytes(out[:n]), n) == 0
        assert back.tts_nvalid == len(cols) and not back.tts_isempty
        for i in range(len(cols)):
            assert back.tts_isnull[i] == s.tts_isnull[i]
            if not s.tts_isnull[i]:
                assert (back.tts_values[i], back.tts_len[i], back.tts_typid[i]) == (s.tts_values[i], s.tts_len[i], s.tts_typid[i])
        used += 1
    assert used >= 10


def test_slot_memtuples_of_every_slot_type_round_trip_and_malformed_ones_are_refused():
    import random
    L = _exec_lib()
    Slot = _slot_type()
    rnd = random.Random(5)
    types = [capi.BOOLOID, capi.INT4OID, capi.DATEOID, capi.INT8OID, capi.FLOAT8OID, capi.TIMESTAMPOID, capi.BPCHAROID, capi.VARCHAROID, capi.TEXTOID]
    for _ in range(300):
        n = rnd.randint(1, 16)
        s = Slot()
        s.tts_nvalid, s.tts_isempty = n, 0
        for i in range(n):
            t = rnd.choice(types)
            s.tts_typid[i] = t
            if rnd.random() < 0.2:
                s.tts_isnull[i] = 1
            elif t in (capi.BPCHAROID, capi.VARCHAROID, capi.TEXTOID):
                ln = rnd.randint(0, 8)
                b = bytes(rnd.randint(1, 255) for _ in range(ln))
                s.tts_values[i], s.tts_len[i] = int.from_bytes(b.ljust(8, b"\0"), "little", signed=True), ln
            elif t == capi.BOOLOID:
                s.tts_values[i] = rnd.randint(0, 1)
            elif t in (capi.INT4OID, capi.DATEOID):
                s.tts_values[i] = rnd.randint(-2**31, 2**31 - 1)
            else:
                s.tts_values[i] = rnd.randint(-2**63, 2**63 - 1)
        out = (C.c_uint8 * 4096)()
        need = C.c_uint32(0)
        ln = L.GgExecFetchSlotMemTuple(C.byref(s), out, len(out), C.byref(need))
        assert ln > 0
        back = Slot()
        typids = (C.c_int32 * n)(*[s.tts_typid[i] for i in range(n)])
        assert L.GgExecStoreMemTuple(C.byref(back), typids, n, bytes(out[:ln]), ln) == 0
        for i in range(n):
            assert back.tts_isnull[i] == s.tts_isnull[i]
            if not s.tts_isnull[i]:
                assert (back.tts_values[i], back.tts_len[i]) == (s.tts_values[i], s.tts_len[i]), (i, s.tts_typid[i])
        # a truncated tuple is refused, not read past its end
        assert L.GgExecStoreMemTuple(C.byref(back), typids, n, bytes(out[:ln]), max(ln - 5, 1)) != 0
    empty = Slot()
    empty.tts_isempty = 1
    assert L.GgExecFetchSlotMemTuple(C.byref(empty), None, 0, None) == -10
