"""Fuzz the AOCS loader (greengage_b200/host/gg_aocs_host.c: gg_aocs_index_column, gg_aocs_plan_tiles) under
AddressSanitizer + UBSan with corrupted and truncated column files: every call must return GG_OK or an error code and never
read outside the buffer it was given (the buffers are exact-size heap allocations, so ASan sees any over-read).
Run through scripts/fuzz/run_aocs_fuzz.sh."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from greengage_b200 import aocs, capi
from test_oracle_aocs import attr

L = C.CDLL(os.environ["FZ_LIB"])
L.gg_aocs_index_column.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
L.gg_aocs_plan_tiles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
L.harness_decode_rows.restype = C.c_uint32
L.harness_decode_rows.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int32, C.c_void_p]
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p
libc.malloc.argtypes = [C.c_size_t]
libc.free.argtypes = [C.c_void_p]
kat = np.load(os.path.join(ROOT, "tests", "golden", "aocs_kat.npz"))
rng = np.random.default_rng(11)
stats = {"ok": 0, "refused": 0, "planned": 0, "decoded": 0, "decode_refused": 0}
cases = [str(k) for k in kat["cases"]]
for it in range(40000):
    key = cases[int(rng.integers(0, len(cases)))]
    name, cs = key.split("_")[0], key.endswith("c1")
    f = np.array(kat[key + "_file"])
    if cs and rng.random() < 0.7:
        cs = False                                   # read a checksummed file as if it had none: structure is garbage from block 2 on
    for _ in range(int(rng.integers(0, 4))):
        pos = int(rng.integers(0, min(f.size, 200))) if rng.random() < 0.7 else int(rng.integers(0, f.size))
        f[pos] = int(rng.integers(0, 256)) if rng.random() < 0.5 else f[pos] ^ (1 << int(rng.integers(0, 8)))
    if rng.random() < 0.3:
        f = f[:int(rng.integers(0, f.size + 1))]
    n = int(f.size)
    p = libc.malloc(n + 16)                           # the device arena's 16 bytes of slack, nothing more: an over-read is a report
    C.memmove(p, f.ctypes.data, n)
    C.memset(p + n, 0, 16)
    a = attr(name if rng.random() < 0.8 else str(rng.choice(["int8", "int4", "bpchar1", "text", "float8", "date"])))
    nb, nr = C.c_int64(0), C.c_int64(0)
    cap = 4096
    d = np.zeros(cap, dtype=aocs.BLOCK_DTYPE)
    rc = L.gg_aocs_index_column(C.byref(a), p, n, int(cs), d.ctypes.data, cap, C.byref(nb), C.byref(nr))
    if rc == 0:
        stats["ok"] += 1
        tr = int(rng.choice([1, 7, 64, 1000, 5000]))
        nt = (nr.value + tr - 1) // tr
        t = np.zeros(max(nt, 1), dtype=aocs.TILE_DTYPE)
        if L.gg_aocs_plan_tiles(d.ctypes.data, nb.value, p, tr, t.ctypes.data, nt) == 0:
            stats["planned"] += 1
            if a.atttypid in aocs.KIND_OF_TYPE and nr.value:
                # the device function, exactly as the kernel runs it (tests/aocs_decode_harness.c), over the same exact-size buffer
                col = aocs.gg_aocs_devcol()
                col.file, col.dir, col.tiles, col.nblocks, col.kind = p, d.ctypes.data, t.ctypes.data, nb.value, aocs.KIND_OF_TYPE[a.atttypid]
                out = np.zeros((nr.value, 2), dtype=np.uint64)
                err = L.harness_decode_rows(C.byref(col), 1, nr.value, tr, out.ctypes.data)
                stats["decoded" if err == 0 else "decode_refused"] += 1
    else:
        assert rc in (-6, -8, -9, -10), rc
        stats["refused"] += 1
    libc.free(p)
print(stats)
