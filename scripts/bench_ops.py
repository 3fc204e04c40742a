#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other operators of the hot path on one segment, and
the Redistribute-HashJoin across segments.  Prints one JSON line per measurement; CUDA-event timing, inputs larger
than L2, >= 3 warm-up passes.

  python scripts/bench_ops.py join  [--rows 1e8] [--orders 2.5e7] [--kind count|q3ish]      BASELINE config 2
  python scripts/bench_ops.py sort  [--rows 1e8]
  python scripts/bench_ops.py motion [--rows 1e8] [--nsegs 8]                                sending side only
  python scripts/bench_ops.py groupby [--rows 1e8]                                           general HashAggregate
  python scripts/bench_ops.py aocs  [--rows 1e8]                                             Q1 over AOCS column files (decode + scan)
  torchrun ... scripts/bench_ops.py rjoin [--rows 1e8] [--orders 2.5e7]                       BASELINE config 3 (N GPUs)
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from greengage_b200 import capi, tpch  # noqa: E402
from greengage_b200.engine import Engine, JoinAgg, Relation, RowRelation, motion_partition  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def gen(eng, table, ncand, **kw):
    spec = tpch.synth_spec(table, int(ncand), **kw)
    pages, nb, nr = tpch.synth_generate(spec)
    rel = Relation(eng, host_pages=pages)
    return rel, nb, nr


def bench_join(args):
    eng = Engine(0)
    li, lnb, lnr = gen(eng, capi.TAB_LINEITEM_NARROW, args.rows, norders=int(args.orders))
    od, onb, onr = gen(eng, capi.TAB_ORDERS, args.orders)
    outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, args.kind, capi.JOIN_INNER)
    ja = JoinAgg(eng, outer, inner, hj, agg, pool)
    build_ms, probe_ms, tot_ms = [], [], []
    for it in range(args.warmup + args.steps):
        eng.timer_start()
        ja.build(od)
        ja.reset()
        ja.probe(li)
        rows, nj = ja.fetch()
        ms = eng.timer_stop()
        st = ja.stats()
        if it >= args.warmup:
            build_ms.append(st["build_ms"]); probe_ms.append(st["probe_ms"]); tot_ms.append(ms)
    b, p, t = np.mean(build_ms), np.mean(probe_ms), np.mean(tot_ms)
    sector = 32 * lnr                                   # one random 32-byte sector per probe
    line = {"op": "hashjoin+agg", "workload": "lineitem-narrow ⋈ orders on l_orderkey (int64), %s" % args.kind,
            "outer_rows": lnr, "inner_rows": onr, "rows_joined": nj, "groups": len(rows),
            "build_ms": b, "probe_ms": p, "total_ms": t, "table_bytes": st["table_bytes"],
            "rows_per_s": (lnr + onr) / (t / 1e3), "probe_rows_per_s": lnr / (p / 1e3),
            "roofline": {"bound": "hbm", "achieved": (lnb * 32768 + sector) / (p / 1e3) / 1e9, "peak": peak(), "unit": "GB/s",
                         "frac": (lnb * 32768 + sector) / (p / 1e3) / 1e9 / peak(),
                         "algorithmic_bytes": "outer pages %d B + one 32 B table sector per probe" % (lnb * 32768)},
            "build_roofline_frac": (onb * 32768 + st["table_bytes"]) / (b / 1e3) / 1e9 / peak()}
    print(json.dumps(line), flush=True)


def bench_sort(args):
    eng = Engine(0)
    L = capi.dev_lib()
    n = int(args.rows)
    rng = np.random.default_rng(1)
    rows = rng.integers(0, 6 * 10**9, n, dtype=np.int64).reshape(n, 1)      # l_orderkey-like: 33 significant bits
    buf = Relation(eng, nblocks=(n * 8 + 64 + 32767) // 32768)
    perm = Relation(eng, nblocks=(n * 4 + 64 + 32767) // 32768)
    capi.check(L.gg_relation_load(buf.h, 0, rows.ctypes.data_as(C.c_void_p), (n * 8) // 32768))
    rem = (n * 8) % 32768
    eng.sync()
    keys = (capi.gg_sortkey * 1)(capi.make_sortkey(0, capi.INT8OID))
    passes = C.c_int(0)
    ms = []
    n_full = ((n * 8) // 32768) * 32768 // 8                                  # rows covered by whole pages
    for it in range(args.warmup + args.steps):
        capi.check(L.gg_sort_device(eng.h, keys, 1, 1, C.c_void_p(buf.device_ptr()), None, n_full, C.c_void_p(perm.device_ptr()), C.byref(passes)))
        if it >= args.warmup:
            ms.append(eng.last_kernel_ms())
    t = np.mean(ms)
    out = perm.read().view(np.uint32)[:n_full]
    srt = rows[:n_full, 0][out.astype(np.int64)]
    assert np.all(np.diff(srt) >= 0)
    algo = n_full * (8 + 8 + passes.value * 32)
    print(json.dumps({"op": "sort", "workload": "%d int64 keys (33 significant bits)" % n_full, "ms": t, "passes": passes.value,
                      "rows_per_s": n_full / (t / 1e3),
                      "roofline": {"bound": "hbm", "achieved": algo / (t / 1e3) / 1e9, "peak": peak(), "unit": "GB/s",
                                   "frac": algo / (t / 1e3) / 1e9 / peak(),
                                   "algorithmic_bytes": "16 B/row key build + 32 B/row per executed radix pass"}}), flush=True)


def bench_groupby(args):
    """The general HashAggregate: GROUP BY l_orderkey (2.5 * 10^7 groups at the default size), count(*) + sum(float8)."""
    from greengage_b200.engine import ScanAgg
    eng = Engine(0)
    li, nb, nr = gen(eng, capi.TAB_LINEITEM_NARROW, args.rows)
    c = tpch.LI_NARROW_COLS
    p = capi.ExprPool()
    agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [p.var(c["orderkey"], capi.INT8OID)],
                        [(capi.AGG_COUNT_STAR, -1), (capi.AGG_SUM_FLOAT8, p.var(c["extendedprice"], capi.FLOAT8OID))], num_groups=int(args.rows // 4))
    sa = ScanAgg(eng, capi.make_scan(capi.synth_tupdesc(capi.TAB_LINEITEM_NARROW), -1), agg, p.pool)
    ms = []
    for it in range(args.warmup + args.steps):
        sa.reset()
        sa.run(li)
        eng.sync()
        if it >= args.warmup:
            ms.append(sa.scan_kernel_ms()[0])
    t = np.mean(ms)
    print(json.dumps({"op": "hashagg-general", "workload": "GROUP BY l_orderkey over %d rows (~%d groups), count(*) + sum(float8)" % (nr, args.rows // 4),
                      "ms": t, "rows_per_s": nr / (t / 1e3), "variant": sa.variant(),
                      "roofline": {"bound": "hbm", "achieved": nb * 32768 / (t / 1e3) / 1e9, "peak": peak(), "unit": "GB/s",
                                   "frac": nb * 32768 / (t / 1e3) / 1e9 / peak(), "algorithmic_bytes": "pages read once (+ 4 random table sectors per row)"}}), flush=True)


def bench_aocs(args):
    """Q1 over the LI-wide segment stored append-only column-oriented: the seven projected column files are loaded
    (host index + tile plan, one H2D of the files), decoded to datum rows on the device (gg_aocs_decode_rows) and scanned.
    Reports the decode kernel, the scan over the decoded rows, and the load (host loader + H2D) beside them."""
    import time
    from greengage_b200 import aocs
    from greengage_b200.engine import ScanAgg
    eng = Engine(0)
    spec = tpch.synth_spec(capi.TAB_LINEITEM_WIDE, args.rows)
    nb, nr = tpch.synth_measure(spec)
    cols = [4, 5, 6, 7, 8, 9, 10]
    t0 = time.time()
    files, nrows = aocs.synth_columns(spec, cols, nr)
    gen_s = time.time() - t0
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    t0 = time.time()
    dc = aocs.DeviceColumns(eng, desc, cols, files)
    load_s = time.time() - t0
    names = dict(quantity=1, extendedprice=2, discount=3, tax=4, returnflag=5, linestatus=6, shipdate=7)
    scan, agg, pool = tpch.q1_plan(stage=capi.AGGSTAGE_NORMAL, desc=dc.rows_tupdesc([1] * len(cols)), cols=names)
    sa = ScanAgg(eng, scan, agg, pool)
    dec_ms, scan_ms = [], []
    rel = None
    for it in range(args.warmup + args.steps):
        if rel is not None:
            rel.free()
            dc.rows.free()
        rel = dc.decode()
        d = eng.last_kernel_ms()
        sa.reset()
        sa.run(rel)
        rows, sc, ps = sa.fetch()
        if it >= args.warmup:
            dec_ms.append(d); scan_ms.append(sa.scan_kernel_ms()[0])
    d, t = float(np.mean(dec_ms)), float(np.mean(scan_ms))
    rowbytes = 8 * (1 + len(cols))
    dec_bytes = dc.bytes_in + nr * rowbytes
    print(json.dumps({"op": "aocs-q1", "workload": "Q1 over %d rows of LI-wide stored as AOCS column files, 7 projected columns" % nr,
                      "column_bytes": dc.bytes_in, "bytes_per_row": dc.bytes_in / nr, "heap_bytes_per_row": 172,
                      "decode_ms": d, "scan_ms": t, "rows_per_s": nr / ((d + t) / 1e3), "groups": len(rows),
                      "host_generate_s": gen_s, "host_index_plan_and_h2d_s": load_s,
                      "roofline": {"bound": "hbm", "kernel": "gg_aocs_rows_kernel", "achieved": dec_bytes / (d / 1e3) / 1e9, "peak": peak(),
                                   "unit": "GB/s", "frac": dec_bytes / (d / 1e3) / 1e9 / peak(),
                                   "algorithmic_bytes": "column files read once + %d B/row of datum rows written" % rowbytes}}), flush=True)


def li_payload():
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_NARROW)
    c = tpch.LI_NARROW_COLS
    p = capi.ExprPool()
    key = p.var(c["orderkey"], capi.INT8OID)
    payload = [key, p.var(c["extendedprice"], capi.FLOAT8OID), p.var(c["discount"], capi.FLOAT8OID), p.var(c["shipdate"], capi.DATEOID)]
    types = [capi.INT8OID, capi.FLOAT8OID, capi.FLOAT8OID, capi.DATEOID]
    return desc, p, key, payload, types


def od_payload():
    desc = capi.synth_tupdesc(capi.TAB_ORDERS)
    p = capi.ExprPool()
    key = p.var(tpch.ORDERS_COLS["orderkey"], capi.INT8OID)
    payload = [key, p.var(tpch.ORDERS_COLS["orderdate"], capi.DATEOID), p.var(tpch.ORDERS_COLS["orderstatus"], capi.BPCHAROID)]
    types = [capi.INT8OID, capi.DATEOID, capi.BPCHAROID]
    return desc, p, key, payload, types


def bench_motion(args):
    eng = Engine(0)
    li, nb, nr = gen(eng, capi.TAB_LINEITEM_NARROW, args.rows)
    desc, p, key, payload, types = li_payload()
    W = 1 + len(payload)
    cap = (int(nr / args.nsegs * 1.1) + 4096) * args.nsegs
    out = Relation(eng, nblocks=(cap * W * 8 + 64 + 32767) // 32768)
    scan = capi.make_scan(desc, -1)
    ms = []
    for it in range(args.warmup + args.steps):
        counts, offs = motion_partition(eng, scan, p.pool, [key], payload, args.nsegs, li, out.device_ptr(), cap)
        if it >= args.warmup:
            ms.append(eng.last_kernel_ms())
    t = np.mean(ms)
    algo = nb * 32768 + nr * W * 8
    print(json.dumps({"op": "motion-send", "workload": "Redistribute lineitem-narrow on l_orderkey to %d segments, %d columns travel" % (args.nsegs, len(payload)),
                      "rows": nr, "ms": t, "rows_per_s": nr / (t / 1e3), "counts_min_max": [min(counts), max(counts)],
                      "roofline": {"bound": "hbm", "achieved": algo / (t / 1e3) / 1e9, "peak": peak(), "unit": "GB/s",
                                   "frac": algo / (t / 1e3) / 1e9 / peak(), "algorithmic_bytes": "pages read once + %d B/row written" % (W * 8)}}), flush=True)


def bench_rjoin(args):
    """Redistribute both sides on the join key (NCCL all-to-all of the device regions), then HashJoin -> Agg locally,
    partial results gathered on rank 0.  Weak scaling: every rank holds args.rows lineitem rows and args.orders orders."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(local)
    li, lnb, lnr = gen(eng, capi.TAB_LINEITEM_NARROW, args.rows * world, norders=int(args.orders * world), nsegs=world, seg=rank)
    od, onb, onr = gen(eng, capi.TAB_ORDERS, args.orders * world, nsegs=world, seg=rank)
    ldesc, lp, lkey, lpay, ltypes = li_payload()
    odesc, op, okey, opay, otypes = od_payload()
    Wl, Wo = 1 + len(lpay), 1 + len(opay)

    def cap_for(n):
        return (int(n / world * 1.15) + 8192) * world

    lcap, ocap = cap_for(lnr), cap_for(onr)
    lsend = torch.empty(lcap * Wl + 8, dtype=torch.int64, device=dev)
    osend = torch.empty(ocap * Wo + 8, dtype=torch.int64, device=dev)
    lrd = capi.rows_tupdesc(ltypes, notnull=[1] * len(ltypes))
    ord_ = capi.rows_tupdesc(otypes, notnull=[1] * len(otypes))
    outer, inner, hj, agg, pool = tpch.join_plan(kind="q3ish", jointype=capi.JOIN_INNER, li_desc=lrd, ord_desc=ord_,
                                                 li_cols=dict(orderkey=1, extendedprice=2, discount=3, shipdate=4),
                                                 ord_cols=dict(orderkey=1, orderdate=2, orderstatus=3))
    ja = JoinAgg(eng, outer, inner, hj, agg, pool)

    def exchange(send, counts, offs, W):
        """all-to-all-v of the per-destination regions; returns (tensor of received rows, nrows)"""
        if world == 1:
            return send[offs[0] * W: (offs[0] + counts[0]) * W + 2], counts[0]
        sc = torch.tensor(counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        rcounts = [int(x) for x in rc.tolist()]
        packed = torch.cat([send[offs[d] * W:(offs[d] + counts[d]) * W] for d in range(world)])
        recv = torch.empty(sum(rcounts) * W + 8, dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv[:sum(rcounts) * W], packed, output_split_sizes=[c * W for c in rcounts],
                               input_split_sizes=[c * W for c in counts])
        return recv, sum(rcounts)

    def step():
        lc, lo = motion_partition(eng, capi.make_scan(ldesc, -1), lp.pool, [lkey], lpay, world, li, lsend.data_ptr(), lcap)
        oc, oo = motion_partition(eng, capi.make_scan(odesc, -1), op.pool, [okey], opay, world, od, osend.data_ptr(), ocap)
        eng.sync()
        lrecv, ln = exchange(lsend, lc, lo, Wl)
        orecv, on = exchange(osend, oc, oo, Wo)
        torch.cuda.synchronize()
        lrel = RowRelation(eng, lrecv.data_ptr(), ln, len(ltypes))
        orel = RowRelation(eng, orecv.data_ptr(), on, len(otypes))
        ja.build(orel)
        ja.reset()
        ja.probe(lrel)
        rows, nj = ja.fetch()
        lrel.free(); orel.free()
        return rows, nj

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rows, nj = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, float(nj), float(lnr + onr)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, nj_tot, rows_tot = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        nj_tot, rows_tot = float(nj), float(lnr + onr)
    if rank == 0:
        print(json.dumps({"op": "redistribute-hashjoin", "n_gpus": world, "scaling": "weak",
                          "workload": "lineitem-narrow ⋈ orders, both redistributed on the join key, q3ish aggregate",
                          "rows_in_per_gpu": lnr + onr, "rows_joined": nj_tot, "ms_per_step": ms / args.steps,
                          "rows_per_s": rows_tot * args.steps / (ms / 1e3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["join", "sort", "motion", "rjoin", "groupby", "aocs"])
    ap.add_argument("--rows", type=float, default=1e8)
    ap.add_argument("--orders", type=float, default=2.5e7)
    ap.add_argument("--kind", default="count")
    ap.add_argument("--nsegs", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    a.rows, a.orders = int(a.rows), int(a.orders)
    {"join": bench_join, "sort": bench_sort, "motion": bench_motion, "rjoin": bench_rjoin, "groupby": bench_groupby, "aocs": bench_aocs}[a.op](a)
