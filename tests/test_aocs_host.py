"""Product host code of the AOCS scan (include/gg_aocs.h, greengage_b200/host/gg_aocs_host.c in libgghost.so; no GPU):
the loader's block directory, checksum verification, the column-file writer and the synthetic relations as column
files — held to the column files the REFERENCE'S OWN objects wrote (tests/golden/aocs_kat.npz) and to the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from _util import GOLD
from greengage_b200 import aocs, capi, tpch
from oracle import pyoracle as po
from test_oracle_aocs import attr, inputs, TYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLD, "aocs_kat.npz"))


def test_host_library_exports_every_declared_symbol():
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gg_aocs.h")).read(), flags=re.S)
    decl = sorted(set(re.findall(r"\b(gg_aocs_[a-z0-9_]+)\s*\(", txt)))
    assert len(decl) == 7, decl
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "greengage_b200", "libgghost.so")]).decode()
    exp = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert [s for s in decl if s not in exp] == []
    assert aocs.BLOCK_DTYPE.itemsize == 40 and aocs.TILE_DTYPE.itemsize == 16


def test_crc32c_known_answers(kat):
    data = kat["crc_inputs"]
    for ln, want in zip(kat["crc_lens"], kat["crc_values"]):
        assert aocs.crc32c(data[:int(ln)]) == int(want)
    # unaligned start and the byte tail of the hardware path
    for a, b in ((1, 4096), (3, 77), (7, 8), (5, 5)):
        assert aocs.crc32c(data[a:b]) == po.lib().or_aocs_crc32c(data[a:b].ctypes.data, b - a)


def test_directory_of_reference_written_files(kat):
    """every block the reference wrote is found with the reference's own first row numbers and row counts, and a fixed-width
    column decoded through the directory alone equals what the reference's reader returned"""
    for key in kat["cases"]:
        name, cs = str(key).split("_")[0], str(key).endswith("c1")
        f = kat[key + "_file"]
        d, nrows = aocs.index_column(attr(name), f, checksum=cs)
        assert nrows == len(kat[key + "_vals"])
        assert np.array_equal(d["first_row"], kat[key + "_firstrows"]) and np.array_equal(d["nrows"], kat[key + "_rowcounts"])
        assert np.all(d["data_off"] % 8 == 0)
        # stride: attlen for fixed-width columns, 2 for char(1) (header byte + character), 0 for the mixed-length text column
        assert np.all(d["stride"] == {"bpchar1": 2, "text": 0}.get(name, TYPES[name][1]))
        has_nulls = kat[key + "_nulls"].any()
        assert np.all((d["null_off"] >= 0) == has_nulls) or has_nulls     # a block of a NULL-bearing column may have none
        if TYPES[name][1] > 0:
            v, nl = aocs.fixed_column_values(attr(name), f, d)
            assert np.array_equal(nl, kat[key + "_nulls"]) and np.array_equal(v, kat[key + "_vals"]), key
        else:
            # varlena: the first stored value of every block sits at data_off
            vals, nulls = kat[key + "_vals"], kat[key + "_nulls"]
            row = 0
            for b in d:
                live = np.nonzero(nulls[row:row + b["nrows"]] == 0)[0]
                if len(live):
                    assert vals[row + live[0]] == b["data_off"]
                row += b["nrows"]


@pytest.mark.parametrize("tile_rows", [64, 1000, 4096])
def test_tile_plan_addresses_every_row(kat, tile_rows):
    """the per-tile starting points (gg_aocs_plan_tiles) + the forward walk a thread does from them reach exactly the values
    the reference's reader returned, for every row of the reference-written fixed-width files (NULL-bearing ones too)"""
    for key in kat["cases"]:
        name, cs = str(key).split("_")[0], str(key).endswith("c1")
        if TYPES[name][1] < 0 or not cs:
            continue
        a, f = attr(name), kat[key + "_file"]
        d, nrows = aocs.index_column(a, f)
        tiles = aocs.plan_tiles(d, f, tile_rows)
        assert len(tiles) == (nrows + tile_rows - 1) // tile_rows
        got = [aocs.tile_values(a, f, d, tiles, tile_rows, t) for t in range(len(tiles))]
        v, nl = np.concatenate([g[0] for g in got]), np.concatenate([g[1] for g in got])
        assert np.array_equal(nl, kat[key + "_nulls"]), key
        assert np.array_equal(v, kat[key + "_vals"]), key
        # tiles that start inside a block carry the NULL count of the rows before them
        pos = np.concatenate([[0], np.cumsum(d["nrows"])])
        for t in range(len(tiles)):
            b = int(tiles[t]["block"])
            assert pos[b] <= t * tile_rows < pos[b + 1] and tiles[t]["row_in_block"] == t * tile_rows - pos[b]
            assert tiles[t]["nulls_before"] == int(kat[key + "_nulls"][pos[b]:t * tile_rows].sum())


def test_writer_reproduces_the_references_files_byte_for_byte(kat):
    for key in kat["cases"]:
        name, cs = str(key).split("_")[0], str(key).endswith("c1")
        nulls = kat[key + "_innulls"] if key + "_innulls" in kat else None
        f = aocs.write_column(attr(name), inputs(kat, name), nulls, blocksize=8192, checksum=cs)
        assert f.size == kat[key + "_file"].size and np.array_equal(f, kat[key + "_file"]), key


def test_writer_equals_the_oracle_on_random_columns():
    rng = np.random.default_rng(3)
    for name in TYPES:
        a = attr(name)
        n = 3000 if name == "text" else 40000
        if name == "text":
            vals = [bytes(rng.integers(97, 123, int(ln)).astype(np.uint8)) for ln in rng.choice([0, 2, 126, 127, 129, 2000, 9000], n)]
        elif name == "bpchar1":
            vals = [bytes([65 + int(x)]) for x in rng.integers(0, 3, n)]
        elif name == "float8":
            vals = [float(x) for x in rng.normal(size=n)]
        else:
            vals = [int(x) for x in rng.integers(-2**31, 2**31, n)]
        for nullfrac in (0.0, 0.3, 1.0):
            nulls = (rng.random(n) < nullfrac).astype(np.uint8) if nullfrac else None
            mine = aocs.write_column(a, vals, nulls)
            want = po.aocs_write_column(a, vals, nulls)
            assert mine.size == want.size and np.array_equal(mine, want), (name, nullfrac)
            d, nrows = aocs.index_column(a, mine)
            assert nrows == n and d["nrows"].max() <= aocs.MAX_BLOCK_ROWS


def test_corruption_and_foreign_blocks_are_refused(kat):
    a = attr("int8")
    good = np.array(kat["int8_n2_c1_file"])
    for pos in (0, 5, 9, 13, 17, 40, good.size - 1):
        f = good.copy()
        f[pos] ^= 0x10
        with pytest.raises(capi.GGError) as e:
            aocs.index_column(a, f)
        assert e.value.code in (-9, -6), pos
    with pytest.raises(capi.GGError):
        aocs.index_column(a, good[:-8])                       # truncated file
    # a datum-stream block of the Dense (RLE_TYPE) version, checksums off: unsupported, not misread
    f = np.array(kat["int8_n0_c0_file"])
    f[16] = 1
    with pytest.raises(capi.GGError) as e:
        aocs.index_column(a, f, checksum=False)
    assert e.value.code == -6
    # a bulk-compressed block (compressedLength != 0)
    f = np.array(kat["int8_n0_c0_file"])
    f[4] |= 1
    with pytest.raises(capi.GGError) as e:
        aocs.index_column(a, f, checksum=False)
    assert e.value.code == -6
    # an int4 reader on an int8 file: the value area is not a whole number of rows
    with pytest.raises(capi.GGError):
        aocs.index_column(attr("int4"), np.array(kat["int8_n0_c1_file"]))


def test_synthetic_relation_as_column_files_is_the_heap_relation():
    """gg_synth_aocs_generate stores the rows gg_synth_generate stores, in the same order: every column decodes (oracle
    reader) to the heap pages' column (oracle deform), and the oracle's Q1 over the column files is bit-identical to its
    Q1 over the heap pages"""
    spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 30000, nsegs=2, seg=1)
    pages, nb, nr = tpch.synth_generate(spec)
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    cols = [0, 4, 5, 6, 7, 8, 9, 10, 15]
    files, nrows = aocs.synth_columns(spec, cols, nr)
    assert nrows == nr
    heap_rows = [r for b in range(nb) for r in po.deform_page(desc, pages, b)]
    for c in cols:
        a = desc.attrs[c]
        v, nl, fr, rc = po.aocs_read_column(a, files[c], nr)
        assert len(v) == nr and not nl.any()
        for i in (0, 1, 2, nr // 2, nr - 1):
            want = heap_rows[i][c]
            if a.attlen == -1:
                hdr = int(files[c][v[i]])
                assert hdr & 0x80 and bytes(files[c][v[i] + 1:v[i] + (hdr & 0x7F)]) == want
            elif a.atttypid == capi.FLOAT8OID:
                assert v[i] == np.float64(want).view(np.int64)
            else:
                assert int(v[i]) == want & ((1 << (8 * a.attlen)) - 1)
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_NORMAL)
    colfiles = [files.get(i) for i in range(desc.natts)]
    arows, asc, aps = po.aocs_seqscan_agg(scan, agg, pool, colfiles, nr)
    hrows, hsc, hps = po.seqscan_agg(scan, agg, pool, pages)
    assert (asc, aps) == (hsc, hps) and len(arows) == len(hrows)
    hgot = {(r.key[0], r.key[1]): r for r in hrows}
    for r in arows:
        h = hgot[(r.key[0], r.key[1])]
        assert all(r.agg[j].f[0] == h.agg[j].f[0] and r.agg[j].i == h.agg[j].i for j in range(8))
