#!/usr/bin/env python
"""Launch-configuration sweep of the other kernels of the path on one GPU: hash build, hash probe, Motion send, general
HashAggregate.  GGB200_NP_CONFIG="ncons,stages,team,ctas" (build / send / hash aggregate / transposed probes) and
GGB200_PRIV_CONFIG="ncons,stages,team" (private-accumulator probe) are read when a pipeline is created, so one process
measures them all over the same resident relations.  One JSON line per measurement."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from greengage_b200 import capi, tpch  # noqa: E402
from greengage_b200.engine import Engine, JoinAgg, Relation, ScanAgg, motion_partition  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
which = (sys.argv[2] if len(sys.argv) > 2 else "build,probe,motion,groupby").split(",")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
eng = Engine(0)
norders = rows // 4


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def emit(**kw):
    print(json.dumps(kw), flush=True)


pages, lnb, lnr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_WIDE, rows, norders=norders))
li = Relation(eng, host_pages=pages)
del pages
pages, onb, onr = tpch.synth_generate(tpch.synth_spec(capi.TAB_ORDERS, norders))
od = Relation(eng, host_pages=pages)
del pages

NP = [None, "7,2,7,2", "7,3,7,2", "10,2,0,2", "14,4,7,1", "21,5,7,1", "28,6,7,1", "21,4,7,1", "28,5,7,1", "14,3,7,1", "14,3,0,1", "21,4,0,1", "7,2,0,3", "9,2,0,3"]

if "build" in which or "probe" in which:
    outer, inner, hj, agg, pool = tpch.join_plan(capi.TAB_LINEITEM_WIDE, "survey", capi.JOIN_INNER)
    ref = None
    if "build" in which:
        for cfg in NP:
            setenv(GGB200_NP_CONFIG=cfg, GGB200_PRIV_CONFIG=None)
            try:
                ja = JoinAgg(eng, outer, inner, hj, agg, pool)
                best = None
                for it in range(4):
                    ja.build(od)
                    st = ja.stats()
                    best = st["build_ms"] if best is None or st["build_ms"] < best else best
                ja.reset()
                ja.probe(li)
                got, nj = ja.fetch()
                res = (nj, got[0].agg[0].i, got[0].agg[1].i)
                ref = ref or res
                emit(op="build", config=cfg, ms=best, rows=onr, frac=(onb * 32768 + 16 * onr) / best / 1e6 / PEAK, table_bytes=st["table_bytes"], equal=res == ref)
                ja.free()
            except Exception as exc:
                emit(op="build", config=cfg, error=str(exc)[:200])
    if "probe" in which:
        setenv(GGB200_NP_CONFIG=None)
        # SWEEP_PROBE="cfg;cfg;..." overrides the list; a cfg is "ncons,stages,team[,ctas]", "-" = the engine's own choice, and a
        # trailing "/u0" runs it with GGB200_JOIN_UNIQUE=0 (the probe scans on to the empty slot although the keys are distinct)
        plist = os.environ.get("SWEEP_PROBE")
        plist = plist.split(";") if plist else ["-", "20,5,0", "18,5,6", "18,4,6", "24,4,6", "12,5,6", "24,3,6", "21,5,7", "30,3,6", "30,4,6", "27,4,9"]
        for ent in plist:
            cfg, _, flag = ent.partition("/")
            cfg = None if cfg == "-" else cfg
            setenv(GGB200_PRIV_CONFIG=cfg, GGB200_JOIN_UNIQUE="0" if flag == "u0" else None)
            try:
                ja = JoinAgg(eng, outer, inner, hj, agg, pool)
                ja.build(od)
                best = None
                for it in range(4):
                    ja.reset()
                    ja.probe(li)
                    got, nj = ja.fetch()
                    ms = ja.stats()["probe_ms"]
                    best = ms if best is None or ms < best else best
                res = (nj, got[0].agg[0].i, got[0].agg[1].i)
                ref = ref or res
                emit(op="probe", config=ent, variant=capi.dev_lib().gg_joinagg_variant(ja.h), ms=best, rows=lnr, frac=(lnb * 32768 + 32 * lnr) / best / 1e6 / PEAK, equal=res == ref)
                ja.free()
            except Exception as exc:
                emit(op="probe", config=ent, error=str(exc)[:200])
        setenv(GGB200_PRIV_CONFIG=None, GGB200_JOIN_UNIQUE=None)

if "motion" in which:
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    c = tpch.LI_WIDE_COLS
    p = capi.ExprPool()
    key = p.var(c["orderkey"], capi.INT8OID)
    payload = [key, p.var(c["extendedprice"], capi.FLOAT8OID)]
    W = 1 + len(payload)
    scan = capi.make_scan(desc, -1)
    for nsegs in (8, 1):
        cap = (int(lnr / nsegs * 1.25) + 8192) * nsegs
        out = Relation(eng, nblocks=(cap * W * 8 + 64 + 32767) // 32768)
        for cfg, win in [(None, "0"), (None, None), ("7,2,7,2", None), ("14,4,7,1", None), ("21,5,7,1", None), ("28,6,7,1", None), ("12,4,6,1", None), ("18,5,6,1", None),
                         ("24,6,6,1", None), ("7,2,0,3", None), (None, "128"), (None, "1024")]:
            setenv(GGB200_NP_CONFIG=cfg, GGB200_MOTION_WINDOW=win)
            try:
                best = None
                for it in range(4):
                    counts, offs = motion_partition(eng, scan, p.pool, [key], payload, nsegs, li, out.device_ptr(), cap)
                    ms = eng.last_kernel_ms()
                    best = ms if best is None or ms < best else best
                emit(op="motion", nsegs=nsegs, config=cfg, window=win, ms=best, rows=lnr, slots=sum(counts), frac=(lnb * 32768 + lnr * W * 8) / best / 1e6 / PEAK)
            except Exception as exc:
                emit(op="motion", nsegs=nsegs, config=cfg, window=win, error=str(exc)[:200])
        out.free()
    setenv(GGB200_NP_CONFIG=None, GGB200_MOTION_WINDOW=None)

if "groupby" in which:
    c = tpch.LI_WIDE_COLS
    p = capi.ExprPool()
    agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [p.var(c["orderkey"], capi.INT8OID)],
                        [(capi.AGG_COUNT_STAR, -1), (capi.AGG_SUM_FLOAT8, p.var(c["extendedprice"], capi.FLOAT8OID))], num_groups=norders)
    scan = capi.make_scan(capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE), -1)
    for cfg in NP:
        setenv(GGB200_NP_CONFIG=cfg)
        try:
            sa = ScanAgg(eng, scan, agg, p.pool)
            best = None
            for it in range(3):
                sa.reset()
                sa.run(li)
                eng.sync()
                ms = sa.scan_kernel_ms()[0]
                best = ms if best is None or ms < best else best
            emit(op="groupby", config=cfg, variant=sa.variant(), ms=best, rows=lnr, frac=lnb * 32768 / best / 1e6 / PEAK)
            sa.free()
        except Exception as exc:
            emit(op="groupby", config=cfg, error=str(exc)[:200])
    setenv(GGB200_NP_CONFIG=None)
