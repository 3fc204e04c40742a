"""Motion host plumbing over 2- and 4-rank gloo groups on CPU (the N > 1 path of bench.py without GPUs):
Redistribute on the group keys routes every partial row to the segment the reference's cdbhash picks,
the FINAL stage combines there, Gather brings the result to rank 0 — and the answer equals the
single-segment answer.  The per-segment partial aggregates come from the oracle (this is a test)."""
import os
import sys

import pytest


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from greengage_b200 import capi, motion, tpch
        from oracle import pyoracle as po
        spec = tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 40000, nsegs=world, seg=rank)
        pages, nb, nr = tpch.synth_generate(spec, nthreads=1)
        scan, part, pool = tpch.q1_plan(capi.TAB_LINEITEM_NARROW, capi.AGGSTAGE_PARTIAL)
        rows, _, _ = po.seqscan_agg(scan, part, pool, pages)
        keyt = [capi.BPCHAROID, capi.BPCHAROID]
        mine = motion.redistribute_aggrows(rows, keyt)
        for r in mine:
            assert motion.route_aggrow(r, keyt, world) == rank
        # the one-collective path for small row counts delivers the same rows in the same order
        mine2 = motion.redistribute_small(rows, keyt)
        assert [bytes(r) for r in mine2] == [bytes(r) for r in mine]
        old = motion.SMALL_MOTION_ROWS
        motion.SMALL_MOTION_ROWS = 1                  # force the collective fallback decision
        assert [bytes(r) for r in motion.redistribute_small(rows, keyt)] == [bytes(r) for r in mine]
        motion.SMALL_MOTION_ROWS = old
        final = po.agg_final(tpch.q1_final_agg(part), mine) if mine else []
        gathered = motion.gather_aggrows(final, 0)
        # the raw-buffer variants bench.py uses: same rows, same order
        raw = motion._rows_to_bytes(rows)
        mbuf, nm = motion.redistribute_small_raw(raw, len(rows), keyt)
        assert nm == len(mine) and mbuf.tobytes() == b"".join(bytes(r) for r in mine)
        fraw = motion._rows_to_bytes(final)
        gbuf, ng = motion.gather_small_raw(fraw, len(final), 0)
        assert ng == len(gathered) and gbuf.tobytes() == b"".join(bytes(r) for r in gathered)
        assert [bytes(r) for r in motion.gather_small(final, 0)] == [bytes(r) for r in gathered]
        if rank == 0:
            out = [(r.key[0], r.key[1], r.agg[7].i, r.agg[0].f[0], r.agg[4].f[0]) for r in gathered]
            q.put(("ok", sorted(out), nr))
        else:
            assert gathered == []
            q.put(("ok", None, nr))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("err", traceback.format_exc(), 0))


@pytest.mark.parametrize("world", [2, 4])
def test_multi_segment_q1_through_motion(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for st, payload, _ in res:
        assert st == "ok", payload
    assert sum(n for _, _, n in res) == 40000
    got = [p for _, p, _ in res if p is not None][0]
    # single-segment answer from the oracle
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 40000))
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_NARROW)
    want, _, _ = po.seqscan_agg(scan, agg, pool, pages)
    want = sorted((r.key[0], r.key[1], r.agg[7].i, r.agg[0].f[0], r.agg[4].f[0]) for r in want)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[:3] == w[:3]
        assert abs(g[3] - w[3]) <= 1e-9 * abs(w[3]) and abs(g[4] - w[4]) <= 1e-9 * abs(w[4])
