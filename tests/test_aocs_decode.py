"""The device function of the AOCS decode kernel (greengage_b200/csrc/gg_aocs_decode.h: gg_aocs_fetch) compiled for the
host by gcc and run row by row exactly as gg_aocs_rows_kernel runs it (tests/aocs_decode_harness.c), against the column
files the REFERENCE wrote (tests/golden/aocs_kat.npz) and against the oracle — so the addressing, the NULL prefix counts,
the sign extension and the string packing of the kernel are checked here without a GPU; the GPU tests then only have to
show that the same code gives the same rows on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _util import GOLD
from greengage_b200 import aocs, capi, tpch
from oracle import pyoracle as po
from test_oracle_aocs import attr, TYPES

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("aocs") / "harness.so")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "aocs_decode_harness.c")])
    L = C.CDLL(so)
    L.harness_decode_rows.restype = C.c_uint32
    L.harness_decode_rows.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int32, C.c_void_p]
    return L


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLD, "aocs_kat.npz"))


def host_decode(L, atts, files, tile_rows, checksum=True):
    """what aocs.DeviceColumns + gg_aocs_decode_rows do, with host pointers"""
    keep, cols, nrows = [], (aocs.gg_aocs_devcol * len(atts))(), None
    for i, (att, f) in enumerate(zip(atts, files)):
        f = np.ascontiguousarray(np.concatenate([f, np.zeros(16, dtype=np.uint8)]))      # the loader's 16 bytes of slack
        d, n = aocs.index_column(att, f[:-16], checksum)
        t = aocs.plan_tiles(d, f, tile_rows)
        assert nrows in (None, n)
        nrows = n
        keep += [f, d, t]
        cols[i].file, cols[i].dir, cols[i].tiles = f.ctypes.data, d.ctypes.data, t.ctypes.data
        cols[i].nblocks, cols[i].kind = len(d), aocs.KIND_OF_TYPE[att.atttypid]
    out = np.zeros((nrows, 1 + len(atts)), dtype=np.uint64)
    err = L.harness_decode_rows(cols, len(atts), nrows, tile_rows, out.ctypes.data)
    return out, err


@pytest.mark.parametrize("tile_rows", [64, 1024, 5000])
def test_rows_of_reference_written_files(harness, kat, tile_rows):
    """every fixed-width and char(1) file the reference wrote decodes to the values its own reader returned"""
    for key in kat["cases"]:
        name, cs = str(key).split("_")[0], str(key).endswith("c1")
        if name == "text":
            continue
        a = attr(name)
        rows, err = host_decode(harness, [a], [kat[key + "_file"]], tile_rows, cs)
        assert err == 0
        nulls, vals = kat[key + "_nulls"], kat[key + "_vals"]
        assert np.array_equal(rows[:, 0], nulls.astype(np.uint64)), key
        live = nulls == 0
        if name == "bpchar1":
            f = kat[key + "_file"]
            want = np.array([int(f[v + 1]) for v in vals[live]], dtype=np.uint64)          # the character after the header byte
        elif TYPES[name][1] == 4:
            want = vals[live].astype(np.uint32).view(np.int32).astype(np.int64).view(np.uint64)   # DatumGetInt32: sign-extended
        else:
            want = vals[live].view(np.uint64)
        assert np.array_equal(rows[live, 1], want), key
        assert not rows[~live, 1].any()


def test_irregular_and_long_strings_are_refused_not_misread(harness, kat):
    rows, err = host_decode(harness, [attr("text")], [kat["text_n0_c1_file"]], 1024)
    assert err & 2
    # char(12): uniform stride, but the value does not fit a packed Datum
    a = attr("bpchar1")
    f = po.aocs_write_column(a, [b"abcdefghijkl"] * 100)
    rows, err = host_decode(harness, [a], [f], 64)
    assert err & 2
    # ... unless the blanks bring it to <= 8 bytes: bcTruelen strips them (varchar.c:653)
    f = po.aocs_write_column(a, [b"abc         ", b"R           ", b"            "] * 50)
    rows, err = host_decode(harness, [a], [f], 64)
    assert err == 0
    assert [capi.unpack_str(int(w), 8).rstrip("\0") for w in rows[:3, 1]] == ["abc", "R", ""]


def test_q1_columns_of_the_synthetic_relation(harness):
    """the seven columns Q1 projects, decoded from column files, are the datum rows the heap relation's tuples hold —
    and carry a row plan to the heap answer (oracle on both sides)"""
    spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 50000, nsegs=3, seg=2)
    pages, nb, nr = tpch.synth_generate(spec)
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    cols = [4, 5, 6, 7, 8, 9, 10]
    files, nrows = aocs.synth_columns(spec, cols, nr)
    rows, err = host_decode(harness, [desc.attrs[c] for c in cols], [files[c] for c in cols], 1024)
    assert err == 0 and nrows == nr == len(rows) and not rows[:, 0].any()
    heap = [r for b in range(nb) for r in po.deform_page(desc, pages, b)]
    for i in list(range(0, nr, 997)) + [nr - 1]:
        h = heap[i]
        want = [np.float64(h[4]).view(np.uint64), np.float64(h[5]).view(np.uint64), np.float64(h[6]).view(np.uint64),
                np.float64(h[7]).view(np.uint64), h[8][0], h[9][0], np.int64(h[10]).view(np.uint64)]
        assert [int(x) for x in rows[i, 1:]] == [int(x) for x in want], i


def test_null_mask_of_a_multi_column_row(harness, kat):
    """columns with independent NULLs and different block boundaries: word 0 of every row carries one bit per projected column"""
    keys = ["int8_n2_c1", "date_n2_c1", "float8_n0_c1", "int4_n2_c0"]
    names = [k.split("_")[0] for k in keys]
    # the golden files of one type share their values but every file drew its own NULLs; all hold 2500 rows
    out = []
    for cs in (True, False):
        sel = [i for i, k in enumerate(keys) if k.endswith("c1") == cs]
        rows, err = host_decode(harness, [attr(names[i]) for i in sel], [kat[keys[i] + "_file"] for i in sel], 777, cs)
        assert err == 0 and len(rows) == 2500
        want_mask = np.zeros(2500, dtype=np.uint64)
        for bit, i in enumerate(sel):
            want_mask |= kat[keys[i] + "_nulls"].astype(np.uint64) << np.uint64(bit)
        assert np.array_equal(rows[:, 0], want_mask)
        out.append(rows)
    assert out[0][:, 0].max() >= 3 and (out[0][:, 0] == 0).any()      # rows with both NULL bits, rows with none
