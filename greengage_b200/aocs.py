"""Host side of the append-only column-oriented (AOCS) scan: ctypes mirror of include/gg_aocs.h (libgghost.so, no GPU
needed) — the loader's block directory, the column-file writer and the synthetic relations as column files.
SURVEY §8f rank 1; DESIGN.md §8.1.  The device kernel that consumes the directory is round-2 work."""
import ctypes as C
import os

import numpy as np

from . import capi

DEFAULT_BLOCKSIZE = 32768
MAX_BLOCK_ROWS = 16382

# gg_aocs_block (40 bytes)
BLOCK_DTYPE = np.dtype([("first_row", np.int64), ("data_off", np.int64), ("null_off", np.int64), ("nrows", np.int32),
                        ("data_len", np.int32), ("stride", np.int32), ("pad", np.int32)])


# gg_aocs_tile (16 bytes)
TILE_DTYPE = np.dtype([("block", np.int32), ("row_in_block", np.int32), ("nulls_before", np.int32), ("pad", np.int32)])


def _check(rc):
    if rc != 0:
        raise capi.GGError(rc, "gg_aocs: error %d" % rc)


def crc32c(buf):
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    return capi.host_lib().gg_aocs_crc32c(a.ctypes.data, a.size)


def index_column(att, file, checksum=True):
    """Validate a column file and return its block directory (structured array of BLOCK_DTYPE) and the row count."""
    L = capi.host_lib()
    f = np.ascontiguousarray(file, dtype=np.uint8)
    nb, nr = C.c_int64(0), C.c_int64(0)
    _check(L.gg_aocs_index_column(C.byref(att), f.ctypes.data, f.size, int(checksum), None, 0, C.byref(nb), C.byref(nr)))
    d = np.zeros(nb.value, dtype=BLOCK_DTYPE)
    _check(L.gg_aocs_index_column(C.byref(att), f.ctypes.data, f.size, int(checksum), d.ctypes.data, d.size, C.byref(nb), C.byref(nr)))
    return d, nr.value


def plan_tiles(directory, file, tile_rows):
    """Per-tile starting points of one column (structured array of TILE_DTYPE): see gg_aocs_plan_tiles."""
    f = np.ascontiguousarray(file, dtype=np.uint8)
    d = np.ascontiguousarray(directory)
    total = int(d["nrows"].sum())
    tiles = np.zeros((total + tile_rows - 1) // tile_rows, dtype=TILE_DTYPE)
    _check(capi.host_lib().gg_aocs_plan_tiles(d.ctypes.data, d.size, f.ctypes.data, tile_rows, tiles.ctypes.data, tiles.size))
    return tiles


def tile_values(att, file, directory, tiles, tile_rows, t):
    """What ONE tile's threads compute, written the way the device kernel will address the column: start at the tile's
    block, walk forward, NULL prefix by popcount.  Returns (values int64[rows of the tile], nulls uint8) for a
    fixed-width column — the emulation the CPU tests hold against the oracle."""
    assert att.attlen > 0
    f = np.ascontiguousarray(file, dtype=np.uint8)
    total = int(directory["nrows"].sum())
    n = min(tile_rows, total - t * tile_rows)
    b, j, skipped = int(tiles[t]["block"]), int(tiles[t]["row_in_block"]), int(tiles[t]["nulls_before"])
    vals, nulls = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.uint8)
    for lane in range(n):
        while j >= int(directory[b]["nrows"]):                   # next storage block of this column
            b, j, skipped = b + 1, 0, 0
        blk = directory[b]
        isnull = 0
        if blk["null_off"] >= 0:
            isnull = (int(f[blk["null_off"] + (j >> 3)]) >> (j & 7)) & 1
        if isnull:
            nulls[lane] = 1
            skipped += 1
        else:
            at = int(blk["data_off"]) + (j - skipped) * att.attlen
            x = int.from_bytes(f[at:at + att.attlen].tobytes(), "little")        # zero-extended, like the block reader
            vals[lane] = x - (1 << 64) if x >= 1 << 63 else x
        j += 1
    return vals, nulls


def write_column(att, values, nulls=None, blocksize=DEFAULT_BLOCKSIZE, checksum=True, first_rownum=1):
    """One column file from python values (bytes for varlena attributes); returns a uint8 array."""
    L = capi.host_lib()
    n = len(values)
    maxlen = max([len(v) for i, v in enumerate(values) if att.attlen == -1 and not (nulls is not None and nulls[i])] or [0])
    out = np.zeros(L.gg_aocs_file_bound(C.byref(att), n, maxlen, blocksize, int(checksum)), dtype=np.uint8)
    w = C.c_void_p()
    _check(L.gg_aocs_writer_create(C.byref(att), blocksize, int(checksum), first_rownum, out.ctypes.data, out.size, C.byref(w)))
    rc = 0
    for i, v in enumerate(values):
        if nulls is not None and nulls[i]:
            rc = L.gg_aocs_writer_put(w, 0, 0, 1)
        elif att.attlen == -1:
            b = v.encode() if isinstance(v, str) else bytes(v)
            buf = C.create_string_buffer(b, len(b))
            rc = L.gg_aocs_writer_put(w, C.addressof(buf), len(b), 0)
        else:
            word = C.c_int64.from_buffer_copy(C.c_double(float(v))).value if att.atttypid == capi.FLOAT8OID else int(v)
            if att.attbyval:
                rc = L.gg_aocs_writer_put(w, word, 0, 0)
            else:
                # fixed-width by-reference attribute: the writer takes a pointer to attlen bytes
                buf = C.create_string_buffer((word & (2**64 - 1)).to_bytes(8, "little"), 8)
                rc = L.gg_aocs_writer_put(w, C.addressof(buf), 0, 0)
        if rc:
            break
    nbytes = C.c_int64(0)
    rc2 = L.gg_aocs_writer_finish(w, C.byref(nbytes))
    _check(rc or rc2)
    return out[:nbytes.value].copy()


def synth_columns(spec, cols, nrows_bound, blocksize=DEFAULT_BLOCKSIZE, checksum=True, nthreads=None):
    """The synthetic relation of `spec` (greengage_b200.tpch.synth_spec) as AOCS column files for the 0-based attribute
    numbers in `cols`; returns ({col: uint8 array}, nrows).  nrows_bound: an upper bound of this segment's rows
    (tpch.synth_measure gives the exact count)."""
    L = capi.host_lib()
    desc = capi.synth_tupdesc(spec.table)
    k = len(cols)
    bufs, caps = [], (C.c_int64 * k)()
    for i, c in enumerate(cols):
        cap = L.gg_aocs_file_bound(C.byref(desc.attrs[c]), nrows_bound, 80, blocksize, int(checksum))
        bufs.append(np.zeros(cap, dtype=np.uint8))
        caps[i] = cap
    ptrs = (C.c_void_p * k)(*[b.ctypes.data for b in bufs])
    outb = (C.c_int64 * k)()
    nrows = C.c_uint64(0)
    rc = L.gg_synth_aocs_generate(C.byref(spec), nthreads or min(k, os.cpu_count() or 1), (C.c_int32 * k)(*cols), k, ptrs, caps,
                                  blocksize, int(checksum), outb, C.byref(nrows))
    _check(rc)
    return {c: bufs[i][:outb[i]] for i, c in enumerate(cols)}, nrows.value


def fixed_column_values(att, file, directory):
    """Decode a fixed-width column through the directory alone — what the device kernel's addressing amounts to:
    value of the j-th non-NULL row of block b = file[data_off + j * attlen ...].  Returns (values int64, nulls uint8)."""
    assert att.attlen > 0
    f = np.ascontiguousarray(file, dtype=np.uint8)
    vals, nulls = [], []
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[att.attlen]
    for b in directory:
        n = int(b["nrows"])
        if b["null_off"] >= 0:
            bits = np.unpackbits(f[b["null_off"]:b["null_off"] + (n + 7) // 8], bitorder="little")[:n]
        else:
            bits = np.zeros(n, dtype=np.uint8)
        stored = f[b["data_off"]:b["data_off"] + b["data_len"]].view(dt).astype(np.uint64).view(np.int64)
        assert len(stored) == n - int(bits.sum())
        v = np.zeros(n, dtype=np.int64)
        v[bits == 0] = stored
        vals.append(v)
        nulls.append(bits)
    if not vals:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.uint8)
    return np.concatenate(vals), np.concatenate(nulls)


# ---- device side: column files in HBM -> datum rows (gg_aocs_decode_rows, csrc/gg_aocs.cu) ----

K_W8, K_I4, K_I2, K_B1, K_BPCHAR, K_TEXT = range(6)
KIND_OF_TYPE = {capi.INT8OID: K_W8, capi.FLOAT8OID: K_W8, capi.TIMESTAMPOID: K_W8, capi.INT4OID: K_I4, capi.DATEOID: K_I4,
                capi.BOOLOID: K_B1, capi.BPCHAROID: K_BPCHAR, capi.VARCHAROID: K_TEXT, capi.TEXTOID: K_TEXT}


class gg_aocs_devcol(C.Structure):
    _fields_ = [("file", C.c_void_p), ("dir", C.c_void_p), ("tiles", C.c_void_p), ("nblocks", C.c_int64),
                ("kind", C.c_int32), ("pad", C.c_int32)]


def _pad16(n):
    return (n + 15) & ~15


class DeviceColumns:
    """The projected column files of one segment file resident in HBM with their block directories and tile plans, and
    their decoding into GG_FMT_DATUMROWS rows every operator scans.

    desc: the table's descriptor; cols: 0-based attribute numbers to project (the row's columns, in this order);
    files: {attribute number: uint8 array}.  The loader work (checksums, directory, tile plan) runs on the host
    (libgghost.so); the decoding on the device (gg_aocs_decode_rows).  Fails loudly without a GPU like everything else."""

    def __init__(self, eng, desc, cols, files, checksum=True, tile_rows=1024, pinned=False, file_shift=0):
        """pinned: stage the arena in pinned host memory and keep it, so that upload() can repeat the host -> device copy
        (the end-to-end measurement of bench.py).  Column files start on 16-byte boundaries of the arena, which is what the
        fused scan's bulk copies want; file_shift = 8 puts them 8 bytes off (tests: the scan then reads those files value by
        value and must give the same answer)."""
        from concurrent.futures import ThreadPoolExecutor
        from .engine import Relation
        self.eng, self.cols, self.tile_rows = eng, list(cols), tile_rows
        self.typids = [desc.attrs[c].atttypid for c in self.cols]
        parts, layout, off, self.nrows = [], [], 0, None
        for c in self.cols:
            if desc.attrs[c].atttypid not in KIND_OF_TYPE:
                raise capi.GGError(-6, "AOCS column %d: type %d is not decodable on the device" % (c, desc.attrs[c].atttypid))

        def load_one(c):
            # the loader's work for one column file: checksums, block directory, tile plan (libgghost.so releases the GIL)
            f = np.ascontiguousarray(files[c], dtype=np.uint8)
            d, nrows = index_column(desc.attrs[c], f, checksum)
            return f, d, nrows, plan_tiles(d, f, tile_rows)

        with ThreadPoolExecutor(max_workers=max(1, min(len(self.cols), os.cpu_count() or 1))) as tp:
            loaded = list(tp.map(load_one, self.cols))
        for c, (f, d, nrows, t) in zip(self.cols, loaded):
            att = desc.attrs[c]
            if self.nrows is not None and nrows != self.nrows:
                raise capi.GGError(-9, "AOCS column files of one segment file disagree on the row count")
            self.nrows = nrows
            entry = {"kind": KIND_OF_TYPE[att.atttypid], "nblocks": len(d)}
            for name, arr in (("file", f), ("dir", d.view(np.uint8).reshape(-1)), ("tiles", t.view(np.uint8).reshape(-1))):
                at = off + (file_shift if name == "file" else 0)
                entry[name] = at
                parts.append((at, arr))
                off += _pad16(arr.size + 16 + file_shift)     # 16 bytes of slack: 8-byte loads at the tail of a bitmap / value area,
                                                              # bulk copies rounded out to 16 bytes
            layout.append(entry)
        self.bytes_in = sum(np.ascontiguousarray(files[c]).size for c in self.cols)
        nb = (off + capi.GG_BLCKSZ - 1) // capi.GG_BLCKSZ
        self.arena_bytes = nb * capi.GG_BLCKSZ
        self._pinned_addr = None
        if pinned:
            from .engine import host_alloc
            self._pinned_addr, arena = host_alloc(self.arena_bytes)
        else:
            arena = np.empty(self.arena_bytes, dtype=np.uint8)
        for o, arr in parts:
            arena[o:o + arr.size] = arr
        self.arena_host = arena if pinned else None
        self.arena = Relation(eng, host_pages=arena)
        base = self.arena.device_ptr()
        self.devcols = (gg_aocs_devcol * len(self.cols))()
        for i, e in enumerate(layout):
            self.devcols[i].file, self.devcols[i].dir, self.devcols[i].tiles = base + e["file"], base + e["dir"], base + e["tiles"]
            self.devcols[i].nblocks, self.devcols[i].kind = e["nblocks"], e["kind"]
        self.rows = None

    def rows_tupdesc(self, notnull=None):
        return capi.rows_tupdesc(self.typids, notnull)

    def upload(self):
        """host -> device copy of the arena again (column files + directories + tile plans), asynchronous on the engine's
        stream: the next kernel is ordered behind it.  Needs pinned=True."""
        assert self.arena_host is not None
        self.arena.load(0, self.arena_host)

    def decode(self):
        """-> engine.RowRelation over the decoded rows (kept until free())"""
        from .engine import Relation, RowRelation, dev_lib
        W = 1 + len(self.cols)
        nb = (self.nrows * W * 8 + 64 + capi.GG_BLCKSZ - 1) // capi.GG_BLCKSZ
        self.rows = Relation(self.eng, nblocks=max(nb, 1))
        L = dev_lib()
        L.gg_aocs_decode_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int32, C.c_void_p]
        capi.check(L.gg_aocs_decode_rows(self.eng.h, self.devcols, len(self.cols), self.nrows, self.tile_rows,
                                         C.c_void_p(self.rows.device_ptr())))
        return RowRelation(self.eng, self.rows.device_ptr(), self.nrows, len(self.cols))

    def read_rows(self):
        """the decoded rows back on the host as an int64 array [nrows, 1 + ncols] (tests)"""
        W = 1 + len(self.cols)
        return self.rows.read().view(np.int64)[:self.nrows * W].reshape(self.nrows, W)

    def free(self):
        for r in (self.rows, self.arena):
            if r is not None:
                r.free()
        self.rows = self.arena = None
        if self._pinned_addr is not None:
            from .engine import host_free
            self.arena_host = None
            host_free(self._pinned_addr)
            self._pinned_addr = None
