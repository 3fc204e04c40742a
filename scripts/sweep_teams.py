#!/usr/bin/env python
"""Sweep of the scan kernel's launch configuration (consumer warps, ring stages, team size, register slots) on Q1 over
lineitem-wide / lineitem-narrow.  Kernels are specialised at run time (GGB200_PLAN_CACHE=0) so any block size works.
Every configuration's result is held to the first one's (counts exact, sums 1e-9).  One JSON line per configuration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from greengage_b200 import capi, tpch  # noqa: E402
from greengage_b200.engine import Engine, Relation, ScanAgg  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "wide,narrow"
os.environ["GGB200_PLAN_CACHE"] = "0"
eng = Engine(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

# (GGB200_PRIV_CONFIG "consumer warps,stages,team", GGB200_REG_SLOTS, GGB200_KEYCACHE); None = the engine's own choice
CONFIGS = {
    "wide": [(None, None, None), ("20,5,0", None, None), ("21,5,7", None, None), ("21,5,7", None, "0"), ("20,5,0", None, "0"), ("21,4,7", None, None), ("14,5,7", None, None),
             ("14,4,7", None, None), ("21,3,7", None, None), ("24,5,8", None, None), ("24,4,8", None, None), ("16,5,8", None, None), ("28,4,7", None, None),
             ("21,5,7", "2", None), ("21,4,7", "0", None), ("14,5,7", "0", None)],
    "narrow": [(None, None, None), ("20,3,0", None, None), ("20,3,0", "0", None), ("21,3,7", None, None), ("21,3,7", "0", None), ("24,3,8", None, None), ("24,3,8", "0", None),
               ("16,3,8", "0", None), ("16,4,8", None, None), ("24,4,8", None, None), ("14,4,7", "0", None), ("15,3,15", "0", None), ("20,4,5", None, None), ("21,4,7", None, None)],
}
for tname in which.split(","):
    table = capi.TAB_LINEITEM_WIDE if tname == "wide" else capi.TAB_LINEITEM_NARROW
    pages, nb, nr = tpch.synth_generate(tpch.synth_spec(table, rows))
    rel = Relation(eng, host_pages=pages)
    del pages
    scan, agg, pool = tpch.q1_plan(table, capi.AGGSTAGE_NORMAL)
    ref = None
    for cfg, regs, kc in CONFIGS[tname]:
        for k, v in (("GGB200_PRIV_CONFIG", cfg), ("GGB200_REG_SLOTS", regs), ("GGB200_KEYCACHE", kc)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        try:
            sa = ScanAgg(eng, scan, agg, pool)
        except Exception as exc:
            print(json.dumps({"table": tname, "config": cfg, "regslots": regs, "keycache": kc, "error": str(exc)[:200]}), flush=True)
            continue
        best = None
        try:
            for it in range(5):
                sa.reset()
                sa.run(rel)
                got, sc, ps = sa.fetch()
                ms = sa.scan_kernel_ms()[0]
                best = ms if best is None or ms < best else best
            res = sorted((r.key[0], r.key[1], r.agg[7].i, [r.agg[i].f[0] for i in range(7)]) for r in got)
            ok = True
            if ref is None:
                ref = res
            else:
                ok = len(res) == len(ref) and all(a[:3] == b[:3] and all(abs(x - y) <= 1e-9 * abs(y) for x, y in zip(a[3], b[3])) for a, b in zip(res, ref))
            print(json.dumps({"table": tname, "config": cfg, "regslots": regs, "keycache": kc, "variant": sa.variant(), "ms": best, "GBps": nb * 32768 / best / 1e6,
                              "frac": nb * 32768 / best / 1e6 / PEAK, "Grows_s": nr / best / 1e6, "scanned": sc, "equal_to_first": ok}), flush=True)
        except Exception as exc:
            print(json.dumps({"table": tname, "config": cfg, "regslots": regs, "error": str(exc)[:200]}), flush=True)
        sa.free()
    rel.free()
