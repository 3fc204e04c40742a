"""Fuzz the plan compiler (greengage_b200/csrc/gg_compile.cpp) under AddressSanitizer + UBSan: every plan the random
test generators build must compile, and tens of thousands of single-field mutations of them (indices, counts, type and
attribute fields set to boundary and garbage values) must come back as a result or an error code — never a crash.
Run through scripts/fuzz/run_compiler_fuzz.sh, which builds the compiler standalone with the sanitizers."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from greengage_b200 import capi, tpch
from greengage_b200.capi import ExprPool
import test_gpu_random_plans as rp, test_gpu_random_joins as rj
from _util import make_desc
L = C.CDLL(os.environ['FZ_LIB'])
L.fz_last_error.restype = C.c_char_p
buf = C.create_string_buffer(1 << 17)
def scanagg(scan, agg, pool): return L.fz_scanagg(C.byref(scan), C.byref(agg), C.byref(pool), buf, 1 << 17)
def join(o, i, hj, agg, pool): return L.fz_join(C.byref(o), C.byref(i), C.byref(hj), C.byref(agg), C.byref(pool), buf, 1 << 17)
def clone(x): return type(x).from_buffer_copy(bytes(x))

desc = make_desc([(capi.INT4OID, 4, "i", 1, 1), (capi.INT4OID, 4, "i", 1, 1), (capi.INT4OID, 4, "i", 1, 0), (capi.FLOAT8OID, 8, "d", 1, 1),
                  (capi.FLOAT8OID, 8, "d", 1, 0), (capi.BPCHAROID, -1, "i", 0, 0), (capi.DATEOID, 4, "i", 1, 1), (capi.INT8OID, 8, "d", 1, 1)])
stats = {"valid_ok": 0, "valid_refused": 0, "mut_ok": 0, "mut_refused": 0}
plans = []
reasons = {}
for seed in range(300):
    rng = np.random.default_rng(1000 + seed)
    p = ExprPool(); g = rp.Gen(rng, p)
    qual = g.boolean(3) if rng.random() < 0.8 else -1
    aggs = [(capi.AGG_COUNT_STAR, -1)]
    for _ in range(int(rng.integers(1, 8))):
        fn = int(rng.choice([capi.AGG_SUM_FLOAT8, capi.AGG_AVG_FLOAT8, capi.AGG_MIN_FLOAT8, capi.AGG_MAX_FLOAT8, capi.AGG_COUNT_ANY]))
        aggs.append((fn, g.f8(3)))
    keys = [[], [p.var(1, capi.INT4OID)], [p.var(1, capi.INT4OID), p.var(6, capi.BPCHAROID)]][int(rng.integers(0, 3))]
    agg = capi.make_agg(int(rng.integers(0, 3)), keys, aggs, num_groups=int(rng.choice([0, 20, 500])))
    scan = capi.make_scan(desc, qual)
    rc = scanagg(scan, agg, p.pool)
    stats["valid_ok" if rc >= 0 else "valid_refused"] += 1
    if rc < 0: reasons[L.fz_last_error().decode()[:40]] = reasons.get(L.fz_last_error().decode()[:40], 0) + 1
    plans.append((scan, agg, p.pool))
jplans = []
for seed in range(60):
    outer, inner, hj, agg, p, opages, ipages, what = rj.random_join_case(seed)
    rc = join(outer, inner, hj, agg, p.pool)
    stats["valid_ok" if rc >= 0 else "valid_refused"] += 1
    jplans.append((outer, inner, hj, agg, p.pool))
print(stats, reasons, flush=True)

# mutations: corrupt single fields of valid plans with boundary / garbage values; the compiler must answer (>= 0 or an error code), never crash
rng = np.random.default_rng(7)
vals = [-2**31, -1000, -2, -1, 0, 1, 2, 3, 31, 32, 33, 63, 64, 255, 256, 1000, 65535, 2**31 - 1]
def mutate_pool(pool):
    n = max(0, min(pool.nnodes, 96))
    k = int(rng.integers(0, 6))
    if k == 0: pool.nnodes = int(rng.choice(vals))
    elif n > 0:
        e = pool.nodes[int(rng.integers(0, max(n, 1)))]
        f = ["kind", "funcid", "rettype", "varno", "varattno", "nargs"][int(rng.integers(0, 6))]
        v = int(rng.choice(vals))
        if f in ("varno", "varattno"): v = max(-32768, min(32767, v))
        if k < 4: setattr(e, f, v)
        else: e.args[int(rng.integers(0, 2))] = int(rng.choice(vals + list(range(n + 2))))
        if rng.random() < 0.1: pool.nnodes = int(rng.integers(0, 97))
def mutate_agg(agg):
    f = int(rng.integers(0, 6))
    v = int(rng.choice(vals))
    if f == 0: agg.numCols = v
    elif f == 1: agg.numAggs = v
    elif f == 2: agg.aggstage = v
    elif f == 3: agg.grpCol[int(rng.integers(0, 4))] = v
    elif f == 4: agg.aggs[int(rng.integers(0, 16))].aggfnoid = v
    else: agg.aggs[int(rng.integers(0, 16))].arg = v
def mutate_scan(scan):
    f = int(rng.integers(0, 5))
    v = int(rng.choice(vals))
    if f == 0: scan.qual = v
    elif f == 1: scan.desc.natts = v
    elif f == 2: scan.desc.format = v
    else:
        a = scan.desc.attrs[int(rng.integers(0, 32))]
        g = ["atttypid", "attlen", "attalign", "attbyval", "attnotnull"][int(rng.integers(0, 5))]
        if g == "attlen": v = max(-32768, min(32767, v))
        if g in ("attalign", "attbyval", "attnotnull"): v = max(-128, min(127, v))
        setattr(a, g, v)
for it in range(30000):
    scan, agg, pool = plans[int(rng.integers(0, len(plans)))]
    scan, agg, pool = clone(scan), clone(agg), clone(pool)
    for _ in range(int(rng.integers(1, 4))):
        [lambda: mutate_pool(pool), lambda: mutate_agg(agg), lambda: mutate_scan(scan)][int(rng.integers(0, 3))]()
    rc = scanagg(scan, agg, pool)
    stats["mut_ok" if rc >= 0 else "mut_refused"] += 1
    if it % 5 == 0:
        o, i, hj, ja, jp = jplans[int(rng.integers(0, len(jplans)))]
        o, i, hj, ja, jp = clone(o), clone(i), clone(hj), clone(ja), clone(jp)
        for _ in range(int(rng.integers(1, 4))):
            c = int(rng.integers(0, 5))
            if c == 0: mutate_pool(jp)
            elif c == 1: mutate_agg(ja)
            elif c == 2: mutate_scan(o if rng.random() < 0.5 else i)
            else:
                f = int(rng.integers(0, 5)); v = int(rng.choice(vals))
                if f == 0: hj.jointype = v
                elif f == 1: hj.nkeys = v
                elif f == 2: hj.joinqual = v
                elif f == 3: hj.outerkey[int(rng.integers(0, 2))] = v
                else: hj.innerkey[int(rng.integers(0, 2))] = v
        rc = join(o, i, hj, ja, jp)
        stats["mut_ok" if rc >= 0 else "mut_refused"] += 1
print(stats)
