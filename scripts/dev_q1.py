"""Development driver: Q1 on synthetic lineitem through the C-ABI, timed with CUDA events."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from greengage_b200 import capi, tpch
from greengage_b200.engine import Engine, Relation, ScanAgg

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
table = capi.TAB_LINEITEM_NARROW if (len(sys.argv) > 2 and sys.argv[2] == "narrow") else capi.TAB_LINEITEM_WIDE
check = len(sys.argv) > 3 and sys.argv[3] == "check"
stage = capi.AGGSTAGE_PARTIAL if "partial" in sys.argv[3:] else capi.AGGSTAGE_NORMAL
spec = tpch.synth_spec(table, n)
t = time.time(); pages, nb, nr = tpch.synth_generate(spec); print("gen %.2fs blocks=%d rows=%d" % (time.time() - t, nb, nr), flush=True)
scan, agg, pool = tpch.q1_plan(table, stage)
eng = Engine(0)
rel = Relation(eng, host_pages=pages)
sa = ScanAgg(eng, scan, agg, pool)
for it in range(6):
    sa.reset()
    sa.run(rel)
    rows, sc, ps = sa.fetch()
    ms = eng.last_kernel_ms()
    print("iter %d: %.3f ms  %.1f GB/s  %.2f Grows/s  scanned=%d passed=%d groups=%d" % (
        it, ms, nb * 32768 / ms / 1e6, sc / ms / 1e6, sc, ps, len(rows)), flush=True)
for r in sorted(rows, key=lambda r: (r.key[0], r.key[1])):
    print(capi.unpack_str(r.key[0], r.keylen[0]), capi.unpack_str(r.key[1], r.keylen[1]),
          [r.agg[i].f[0] for i in range(7)], r.agg[7].i)
if check:
    from oracle import pyoracle as po
    t = time.time(); want, wsc, wps = po.seqscan_agg(scan, agg, pool, pages); print("oracle %.2fs" % (time.time() - t))
    got = {(r.key[0], r.key[1]): r for r in rows}
    assert sc == wsc and ps == wps and len(got) == len(want)
    worst = 0
    for w in want:
        g = got[(w.key[0], w.key[1])]
        assert g.agg[7].i == w.agg[7].i
        for i in range(7):
            worst = max(worst, abs(g.agg[i].f[0] - w.agg[i].f[0]) / abs(w.agg[i].f[0]))
    print("parity ok, worst rel err %.3e" % worst)
