#!/usr/bin/env python
"""bench.py — headline benchmark: rows/sec of TPC-H lineitem SeqScan + HashAggregate on B200 segments.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Q1's scan + filter + GROUP BY (l_returnflag, l_linestatus) with its
8 aggregates (no ORDER BY) over a 10^8-row synthetic lineitem heap relation (16 columns, 32 KB pages,
17.25 GB) per GPU segment.  One step = one pass of the hot path over that relation:
  value  pages already resident in HBM (the segment's buffer pool), result rows fetched to the host
  e2e    the same call with the pages in pinned HOST memory: H2D copies inside the timed region
N > 1 is weak scaling: every rank scans its own 10^8-row segment (DISTRIBUTED RANDOMLY), the partial
aggregate rows go through a Redistribute Motion on the group keys (NCCL all-to-all), a FINAL-stage
aggregate and a Gather Motion to rank 0 — the reference's two-stage plan (tpch500GB.out:1771-1782).

The reference arm times the CPU executor restatement (oracle/, one thread per segment over all host
cores) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from greengage_b200 import capi, tpch  # noqa: E402

METRIC = "rows_per_sec_lineitem_scan_hashagg"
UNIT = "rows/s"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_quota():
    """CPUs the container may actually use (cgroup quota), or None: the affinity mask can be far wider than that."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return float(q) / float(per)
    except Exception:
        pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return q / per
    except Exception:
        pass
    return None


def cores_note(cores):
    q = cpu_quota()
    return "%d threads (affinity mask)%s" % (cores, "" if q is None else ", cgroup CPU quota %.1f" % q)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def has_sample(self):
        try:
            return os.path.getsize(self.path) > 0
        except OSError:
            return False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_baseline_run(table, rows, threads, steps=1, warmup=0):
    """The oracle (CPU restatement of the reference executor) on `rows` synthetic rows, `threads` segments."""
    from oracle import pyoracle as po
    spec = tpch.synth_spec(table, rows)
    pages, nb, nr = tpch.synth_generate(spec)
    scan, part, pool = tpch.q1_plan(table, capi.AGGSTAGE_PARTIAL)
    fin = tpch.q1_final_agg(part)
    best = None
    for i in range(warmup + steps):
        out, secs, scanned = po.seqscan_agg_mt(scan, part, fin, pool, pages, threads)
        assert scanned == nr and len(out) >= 1
        if i >= warmup:
            best = secs if best is None else min(best, secs)
    return nr, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    table = capi.TAB_LINEITEM_NARROW if args.table == "narrow" else capi.TAB_LINEITEM_WIDE
    # bounded sample: ~3 s of CPU work per step at ~3 M rows/s/core
    rows = int(min(args.rows * min(max(world, 1), 2), 8_000_000 * cores, 200_000_000))      # <= 34 GB of pages, a few seconds per step
    t0 = time.time()
    total_secs = 0.0
    nr = 0
    from oracle import pyoracle as po
    spec = tpch.synth_spec(table, rows)
    pages, nb, nr = tpch.synth_generate(spec)
    scan, part, pool = tpch.q1_plan(table, capi.AGGSTAGE_PARTIAL)
    fin = tpch.q1_final_agg(part)
    for i in range(args.warmup + args.steps):
        out, secs, scanned = po.seqscan_agg_mt(scan, part, fin, pool, pages, cores)
        if i >= args.warmup:
            total_secs += secs
    value = nr * args.steps / total_secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total_secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Q1 scan+filter+hashagg (no ORDER BY), lineitem-%s, %d rows/GPU" % (args.table, args.rows),
                   "sample_rows": nr, "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of the workload's rows, one oracle thread (= CPU segment) per host core, pages in RAM; %s" % (nr, cores_note(cores))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": round(time.time() - t0 - total_secs, 1),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows", type=float, default=1e8, help="rows per GPU segment")
    ap.add_argument("--table", default="wide", choices=["wide", "narrow"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.rows = int(args.rows)
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from greengage_b200.engine import Engine, Relation, ScanAgg, agg_final_raw, host_alloc, host_free

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    eng = Engine(device)
    table = capi.TAB_LINEITEM_NARROW if args.table == "narrow" else capi.TAB_LINEITEM_WIDE

    # ---- the segment's relation: generated on the host (pinned), loaded into HBM ----
    t_setup = time.time()
    spec = tpch.synth_spec(table, args.rows * world, nsegs=world, seg=rank)
    nb, nr = tpch.synth_measure(spec)
    nbytes = nb * capi.GG_BLCKSZ
    pinned = True
    try:
        haddr, hview = host_alloc(nbytes)
    except Exception:
        pinned = False
        hview = np.empty(nbytes, dtype=np.uint8)
        haddr = hview.ctypes.data
    tpch.synth_generate(spec, out=haddr)
    rel = Relation(eng, nblocks=nb)
    rel.load(0, hview)
    eng.sync()
    setup_s = time.time() - t_setup

    if world == 1:
        scan, agg, pool = tpch.q1_plan(table, capi.AGGSTAGE_NORMAL)
        fin = None
    else:
        scan, agg, pool = tpch.q1_plan(table, capi.AGGSTAGE_PARTIAL)
        fin = tpch.q1_final_agg(agg)
    sa = ScanAgg(eng, scan, agg, pool)
    key_typids = [capi.BPCHAROID, capi.BPCHAROID]
    from greengage_b200 import motion

    def finish(buf, n):
        """everything above the partial aggregate: Redistribute -> FINAL Agg -> Gather (N > 1); raw row buffers"""
        if world == 1:
            return n
        dev = torch.device("cuda", local_rank)
        mine, nm = motion.redistribute_small_raw(buf, n, key_typids, device=dev)
        fbuf, nf = agg_final_raw(eng, fin, mine, nm, cap=256) if nm else (mine, 0)
        gbuf, ng = motion.gather_small_raw(fbuf, nf, 0, device=dev)
        return ng

    scan_ms_tot, scan_launches = 0.0, 0
    fetched = [0]                 # result rows the last step copied device -> host on this rank

    def step_resident():
        nonlocal scan_ms_tot, scan_launches
        sa.reset()
        sa.run(rel)
        buf, n, scanned, passed = sa.fetch_raw(256)
        fetched[0] = n
        ms, k = sa.scan_kernel_ms()
        scan_ms_tot += ms
        scan_launches += k
        return finish(buf, n), scanned

    def step_e2e():
        sa.reset()
        sa.run_host(haddr, nb)
        buf, n, scanned, passed = sa.fetch_raw(256)
        return finish(buf, n), scanned

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        eng.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident: W warm-up steps, then exactly K timed steps ----
    # The clock sampler (nvidia-smi) is started first and the GPU is kept under load, untimed, until its first line is
    # out: NVML initialisation can stall CUDA calls for hundreds of milliseconds, which must not land in the K steps.
    # Every sample it takes from then on is under load (pre-steps, warm-up, timed steps).
    sampler = ClockSampler(device) if rank == 0 else None
    t_pre = time.time()
    while True:
        more = 1 if (rank == 0 and sampler.proc is not None and not sampler.has_sample() and time.time() - t_pre < 8.0) else 0
        if dist is not None:
            flag = torch.tensor([more], dtype=torch.int32, device=torch.device("cuda", local_rank))
            dist.broadcast(flag, 0)
            more = int(flag.item())
        if not more:
            break
        step_resident()
    for _ in range(args.warmup):
        result, scanned = step_resident()
    assert scanned == nr, (scanned, nr)
    scan_ms_tot, scan_launches = 0.0, 0
    barrier()
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(args.steps):
        result, scanned = step_resident()
    ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)
    total_rows = sum_over_ranks(float(nr))
    value = total_rows * args.steps / (ms / 1000.0)
    scan_ms = scan_ms_tot / max(scan_launches, 1)
    variant = sa.variant()

    # ---- end to end: pages start in host memory every step ----
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        eng.timer_start()
        for _ in range(args.e2e_steps):
            result_e, scanned = step_e2e()
        ems = max_over_ranks(eng.timer_stop())
        barrier()
        e2e = {"value": total_rows * args.e2e_steps / (ems / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": int(nbytes * world), "d2h_bytes_per_step": int((fetched[0] * C.sizeof(capi.gg_aggrow) + 28) * max(world, 1)),
               "steps": args.e2e_steps, "ms_per_step": ems / args.e2e_steps,
               "host_memory": "pinned" if pinned else "pageable"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        sample = int(min(args.rows, 4_000_000 * cores))
        srows, secs = cpu_baseline_run(table, sample, cores)
        cpu = {"value": srows / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the workload's %d rows, one oracle thread (= one CPU segment) per host core, pages in RAM; %s" % (srows, args.rows, cores_note(cores))}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = nbytes / (scan_ms / 1000.0) / 1e9 if scan_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_scanagg_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("table") == args.table:
                    traffic = tj["dram_bytes_per_block"] * nb
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Q1 scan+filter+hashagg (no ORDER BY), lineitem-%s, %d rows/GPU" % (args.table, args.rows),
                       "rows_per_gpu": nr, "blocks_per_gpu": nb, "bytes_per_gpu": nbytes,
                       "l2": "input %.1f GB per GPU >> 126 MB L2, streamed once per step" % (nbytes / 1e9),
                       "plan": "Agg(NORMAL)<-SeqScan" if world == 1 else "Gather<-Agg(FINAL)<-Redistribute<-Agg(PARTIAL)<-SeqScan",
                       "kernel_variant": variant},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "gg scan+agg (TMA page ring)", "kernel_ms": scan_ms,
                         "algorithmic_bytes": nbytes, "peak_source": peak_src},
            "gpu_launches": int(launches), "clocks": clocks, "setup_s": round(setup_s, 1),
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    sa.free()
    rel.free()
    if pinned:
        host_free(haddr)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
