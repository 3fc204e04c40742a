#!/usr/bin/env python
"""A/B measurement of the headline scan kernel: Q1 over 10^8-row lineitem-wide, the engine's own launch configuration, the
build-time specialised kernel of whichever libggb200.so GGB200_DEVLIB names (scripts/ab_build.sh).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from greengage_b200 import capi, tpch  # noqa: E402
from greengage_b200.engine import Engine, Relation, ScanAgg  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
eng = Engine(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, rows))
rel = Relation(eng, host_pages=pages)
del pages
scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_NORMAL)
sa = ScanAgg(eng, scan, agg, pool)
ms = []
for it in range(13):
    sa.reset()
    sa.run(rel)
    got, sc, ps = sa.fetch()
    if it >= 3:
        ms.append(sa.scan_kernel_ms()[0])
mean = sum(ms) / len(ms)
print(json.dumps({"lib": os.environ.get("GGB200_DEVLIB", "default"), "priv_config": os.environ.get("GGB200_PRIV_CONFIG"), "variant": sa.variant(),
                  "ms_mean": mean, "ms_min": min(ms), "frac_mean": nb * 32768 / mean / 1e6 / PEAK, "scanned": sc, "passed": ps,
                  "count_sum": sum(r.agg[7].i for r in got)}), flush=True)
