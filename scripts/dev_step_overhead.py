"""Where the time of one N > 1 bench step goes (single process, world_size 1 NCCL group): kernel vs host-side pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
import torch, torch.distributed as dist
from greengage_b200 import capi, tpch, motion
from greengage_b200.engine import Engine, Relation, ScanAgg, agg_final_raw
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
eng = Engine(0)
pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 30_000_000))
rel = Relation(eng, host_pages=pages)
scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL)
fin = tpch.q1_final_agg(agg)
sa = ScanAgg(eng, scan, agg, pool)
keyt = [capi.BPCHAROID, capi.BPCHAROID]
dev = torch.device("cuda", 0)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)
for it in range(60):
    if it == 10: T.clear()
    t = time.perf_counter(); sa.reset(); tick("reset", t)
    t = time.perf_counter(); sa.run(rel); tick("run(launch)", t)
    t = time.perf_counter(); buf, n, sc, ps = sa.fetch_raw(256); tick("fetch (waits for the kernel)", t)
    t = time.perf_counter(); mine, nm = motion.redistribute_small_raw(buf, n, keyt, device=dev); tick("redistribute", t)
    t = time.perf_counter(); fb, nf = agg_final_raw(eng, fin, mine, nm, cap=256); tick("final agg", t)
    t = time.perf_counter(); gb, ng = motion.gather_small_raw(fb, nf, 0, device=dev); tick("gather", t)
ms, k = sa.scan_kernel_ms()
print("scan kernel ms/step (last):", ms / max(k, 1))
for k2, v in T.items():
    print("%-32s %8.1f us/step" % (k2, v / 50 * 1e6))
dist.destroy_process_group()
