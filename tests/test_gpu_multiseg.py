"""Several GPU segments on one box: one process per GPU, the C interconnect over NCCL (gg_ic_*), plans through the
executor-node surface.  Needs >= 2 visible GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, case, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)          # plumbing only: hands the NCCL id around
        from greengage_b200 import capi, executor as ex, tpch
        from greengage_b200.engine import Engine, Interconnect, Relation
        eng = Engine(rank)
        uid = [Interconnect.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        ic = Interconnect(eng, world, rank, uid[0])
        b = ex.PlanBuilder()
        if case == "q1":
            pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 200_000, nsegs=world, seg=rank), nthreads=2)
            plan, pool = tpch.q1_exec_plan(b, two_stage=True)
            rels = [Relation(eng, host_pages=pages)]
        elif case == "fail":
            pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 50_000, nsegs=world, seg=rank), nthreads=2)
            if rank == 1:
                pages = pages.copy()
                pages[12:16] = 0xFF
            plan, pool = tpch.q1_exec_plan(b, two_stage=True)
            rels = [Relation(eng, host_pages=pages)]
        else:
            li, _, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 200_000, seed=9, norders=50_000, nsegs=world, seg=rank), nthreads=2)
            od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 50_000, seed=9, nsegs=world, seg=rank), nthreads=2)
            plan, pool, _, _ = tpch.rjoin_exec_plan(b, case)
            rels = [Relation(eng, host_pages=li), Relation(eng, host_pages=od)]
        x = ex.Executor(eng, pool, rels, plan, nsegs=world, segindex=rank, interconnect=ic)
        try:
            rows = x.rows()
            loc = x.locations()
            x.rescan()
            again = x.rows()
            out = ("rows", rows, loc, again)
        except ex.ExecError as e:
            out = ("error", e.code, str(e), None)
        x.end()
        for r in rels:
            r.free()
        ic.close(has_errors=out[0] == "error")
        eng.close()
        q.put(("ok", rank, out, nr))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put(("err", rank, traceback.format_exc(), 0))


def run(world, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() * 7 + world * 13 + len(case)) % 90
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=240) for _ in procs]          # the cases take seconds; a stuck exchange must not cost the box
    except Exception:
        for p in procs:
            if p.is_alive():
                p.kill()
        raise AssertionError("a segment did not answer within 240 s (GGB200_IC_TRACE=1 GGB200_EXEC_TRACE=1 show where each one is)")
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[0] == "ok", r[2]
    return {r[1]: r for r in res}


def b2f(v):
    return np.int64(v).view(np.float64).item()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_stage_q1_over_nccl_segments():
    sys.path.insert(0, ROOT)
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    world = min(_ngpus(), 4)
    by = run(world, "q1")
    pages, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, 200_000))
    scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_WIDE)
    want, _, _ = po.seqscan_agg(scan, agg, pool, pages)
    assert sum(by[r][3] for r in by) == 200_000
    for r in range(1, world):
        assert by[r][2][0] == "rows" and by[r][2][1] == [] and by[r][2][3] == []
    kind, rows, loc, again = by[0][2]
    assert [l for _, l in loc[1:3]] == ["device-groups", "device-groups"]
    got = {(v[0], v[1]): v for v, nl, ty, ln in rows}
    assert len(got) == len(want) == 4
    for w in want:
        v = got[(w.key[0], w.key[1])]
        assert v[9] == w.agg[7].i
        for col in range(7):
            assert abs(b2f(v[2 + col]) - w.agg[col].f[0]) <= 1e-9 * abs(w.agg[col].f[0])
    assert sorted(r[0][:2] + [r[0][9]] for r in again) == sorted(r[0][:2] + [r[0][9]] for r in rows)


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("kind", ["q3ish", "survey"])
def test_redistribute_hashjoin_over_nccl_segments(kind):
    sys.path.insert(0, ROOT)
    from greengage_b200 import capi, tpch
    from oracle import pyoracle as po
    world = 2
    by = run(world, kind)
    li, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 200_000, seed=9, norders=50_000))
    od, _, _ = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, 50_000, seed=9))
    outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_NARROW, kind, capi.JOIN_INNER)
    want, nj = po.hashjoin_agg(outer, inner, hj, agg, pool, li, od)
    assert by[1][2][0] == "rows" and by[1][2][1] == []
    rows = by[0][2][1]
    assert len(rows) == len(want)
    if kind == "survey":
        v, w = rows[0][0], want[0]
        assert v[0] == w.agg[0].i == nj and v[1] == w.agg[1].i
        assert abs(b2f(v[2]) - w.agg[2].f[0]) <= 1e-6 * abs(w.agg[2].f[0])
    else:
        bw = {r.key[0]: r for r in want}
        for v, nl, ty, ln in rows:
            w = bw[v[0]]
            assert v[1] == w.agg[0].i and v[3] == w.agg[2].i
            assert abs(b2f(v[2]) - w.agg[1].f[0]) <= 1e-6 * abs(w.agg[1].f[0])


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_an_error_on_one_gpu_segment_reaches_all_of_them():
    by = run(2, "fail")
    assert by[0][2][0] == "error" and by[1][2][0] == "error"
