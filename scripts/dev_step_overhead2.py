"""dev_step_overhead with TWO processes (gloo, both on cuda:0): is the N > 1 step slow because of the Motion path?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from greengage_b200 import capi, tpch, motion
from greengage_b200.engine import Engine, Relation, ScanAgg, agg_final_raw
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
backend = os.environ.get("BACKEND", "gloo")
if backend == "nccl":
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    dev, gpu = torch.device("cuda", rank), rank
else:
    dist.init_process_group("gloo")
    dev, gpu = None, 0
eng = Engine(gpu)
pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 30_000_000 * world, nsegs=world, seg=rank))
rel = Relation(eng, host_pages=pages)
scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE, capi.AGGSTAGE_PARTIAL)
fin = tpch.q1_final_agg(agg)
sa = ScanAgg(eng, scan, agg, pool)
keyt = [capi.BPCHAROID, capi.BPCHAROID]
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)
for it in range(40):
    if it == 10: T.clear()
    t = time.perf_counter(); sa.reset(); tick("reset", t)
    t = time.perf_counter(); sa.run(rel); tick("run(launch)", t)
    t = time.perf_counter(); buf, n, sc, ps = sa.fetch_raw(256); tick("fetch (waits for the kernel)", t)
    t = time.perf_counter(); mine, nm = motion.redistribute_small_raw(buf, n, keyt, device=dev); tick("redistribute", t)
    t = time.perf_counter(); fb, nf = agg_final_raw(eng, fin, mine, nm, cap=256) if nm else (mine, 0); tick("final agg", t)
    t = time.perf_counter(); gb, ng = motion.gather_small_raw(fb, nf, 0, device=dev); tick("gather", t)
ms, k = sa.scan_kernel_ms()
if rank == 0:
    print("rank0 scan kernel ms/step (last):", ms / max(k, 1), "mine:", nm, "final:", nf)
    for k2, v in T.items():
        print("%-32s %8.1f us/step" % (k2, v / 30 * 1e6))
dist.destroy_process_group()
