"""TPC-H-shaped plans and synthetic relations used by tests and bench.py.

q1_plan() builds exactly the reference's Q1 (src/test/regress/output/rpt_tpch.source:288-307):
the scan qual is planned as date_le_timestamp against a timestamp constant
(src/test/regress/expected/tpch500GB.out:1782).
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .capi import (AGG_AVG_FLOAT8, AGG_COUNT_STAR, AGG_SUM_FLOAT8, BOOLOID, BPCHAROID, DATEOID, FLOAT8OID,
                   TIMESTAMPOID, ExprPool)

USECS_PER_DAY = 86400000000
D_1998_12_01 = -396          # days since 2000-01-01

# attribute numbers (1-based) of the columns Q1 touches, per table layout
LI_WIDE_COLS = dict(orderkey=1, quantity=5, extendedprice=6, discount=7, tax=8, returnflag=9, linestatus=10, shipdate=11)
LI_NARROW_COLS = dict(orderkey=1, quantity=2, extendedprice=3, discount=4, tax=5, returnflag=6, linestatus=7, shipdate=8)


def q1_plan(table=capi.TAB_LINEITEM_WIDE, stage=capi.AGGSTAGE_NORMAL, interval_days=90, desc=None, cols=None, flags=0):
    """Q1: scan + filter + group by (l_returnflag, l_linestatus) + 4 sums, 3 avgs, count(*).
    `desc` / `cols` override the relation layout (e.g. datum rows a Motion delivered)."""
    cols = cols or (LI_WIDE_COLS if table == capi.TAB_LINEITEM_WIDE else LI_NARROW_COLS)
    desc = desc or capi.synth_tupdesc(table)
    p = ExprPool()
    qty = p.var(cols["quantity"], FLOAT8OID)
    price = p.var(cols["extendedprice"], FLOAT8OID)
    disc = p.var(cols["discount"], FLOAT8OID)
    tax = p.var(cols["tax"], FLOAT8OID)
    flag = p.var(cols["returnflag"], BPCHAROID)
    status = p.var(cols["linestatus"], BPCHAROID)
    shipdate = p.var(cols["shipdate"], DATEOID)
    cutoff = p.const(TIMESTAMPOID, (D_1998_12_01 - interval_days) * USECS_PER_DAY)
    qual = p.func(capi.F_DATE_LE_TIMESTAMP, BOOLOID, shipdate, cutoff)
    one = p.const(FLOAT8OID, 1.0)
    disc_price = p.func(capi.F_FLOAT8MUL, FLOAT8OID, price, p.func(capi.F_FLOAT8MI, FLOAT8OID, one, disc))
    disc_price2 = p.func(capi.F_FLOAT8MUL, FLOAT8OID, price, p.func(capi.F_FLOAT8MI, FLOAT8OID, one, disc))
    charge = p.func(capi.F_FLOAT8MUL, FLOAT8OID, disc_price2, p.func(capi.F_FLOAT8PL, FLOAT8OID, one, tax))
    scan = capi.make_scan(desc, qual)
    agg = capi.make_agg(stage, [flag, status], [
        (AGG_SUM_FLOAT8, qty), (AGG_SUM_FLOAT8, price), (AGG_SUM_FLOAT8, disc_price), (AGG_SUM_FLOAT8, charge),
        (AGG_AVG_FLOAT8, qty), (AGG_AVG_FLOAT8, price), (AGG_AVG_FLOAT8, disc), (AGG_COUNT_STAR, -1)], flags=flags)
    return scan, agg, p.pool


def numeric_lineitem_desc():
    """lineitem as the reference's regression suite declares it (input/rpt_tpch.source:16-35): the four measures are
    numeric(15,2); the rest as in the float8 table"""
    desc = capi.synth_tupdesc(capi.TAB_LINEITEM_WIDE)
    for a in (4, 5, 6, 7):
        at = desc.attrs[a]
        at.atttypid, at.attlen, at.attalign, at.attbyval, at.atttypmod = capi.NUMERICOID, -1, ord("i"), 0, ((15 << 16) | 2) + 4
    return desc


def q1_plan_numeric(desc=None, interval_days=90):
    """The reference's Q1 over numeric columns (output/rpt_tpch.source:288-307): numeric_mul / numeric_sub / numeric_add,
    sum / avg over numeric — exact arithmetic, results to the last digit of the golden answer."""
    desc = desc or numeric_lineitem_desc()
    c = LI_WIDE_COLS
    N = capi.NUMERICOID
    p = ExprPool()
    qty, price, disc, tax = (p.var(c[k], N) for k in ("quantity", "extendedprice", "discount", "tax"))
    flag, status, shipdate = p.var(c["returnflag"], BPCHAROID), p.var(c["linestatus"], BPCHAROID), p.var(c["shipdate"], DATEOID)
    qual = p.func(capi.F_DATE_LE_TIMESTAMP, BOOLOID, shipdate, p.const(TIMESTAMPOID, (D_1998_12_01 - interval_days) * USECS_PER_DAY))
    one = p.const(N, "1")
    disc_price = p.func(capi.F_NUMERIC_MUL, N, price, p.func(capi.F_NUMERIC_SUB, N, one, disc))
    disc_price2 = p.func(capi.F_NUMERIC_MUL, N, price, p.func(capi.F_NUMERIC_SUB, N, one, disc))
    charge = p.func(capi.F_NUMERIC_MUL, N, disc_price2, p.func(capi.F_NUMERIC_ADD, N, one, tax))
    S, A = capi.AGG_SUM_NUMERIC, capi.AGG_AVG_NUMERIC
    agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [flag, status], [(S, qty), (S, price), (S, disc_price), (S, charge), (A, qty), (A, price), (A, disc),
                                                                (AGG_COUNT_STAR, -1)])
    return capi.make_scan(desc, qual), agg, p.pool


def q1_final_agg(agg):
    """The FINAL-stage Agg above the Redistribute Motion: same aggregates, grpCol carries key type OIDs."""
    fin = capi.gg_agg()
    C.memmove(C.byref(fin), C.byref(agg), C.sizeof(capi.gg_agg))
    fin.aggstage = capi.AGGSTAGE_FINAL
    fin.grpCol[0] = BPCHAROID
    fin.grpCol[1] = BPCHAROID
    return fin


def count_star_plan(table=capi.TAB_LINEITEM_WIDE, stage=capi.AGGSTAGE_NORMAL):
    """BASELINE config 0: SELECT count(*) FROM t"""
    desc = capi.synth_tupdesc(table)
    p = ExprPool()
    scan = capi.make_scan(desc, -1)
    agg = capi.make_agg(stage, [], [(AGG_COUNT_STAR, -1)])
    return scan, agg, p.pool


ORDERS_COLS = dict(orderkey=1, custkey=2, orderstatus=3, totalprice=4, orderdate=5, orderpriority=6, clerk=7,
                   shippriority=8, comment=9)
D_1995_03_15 = -1753


def join_plan(table=capi.TAB_LINEITEM_NARROW, kind="count", jointype=capi.JOIN_INNER, li_desc=None, ord_desc=None,
              li_cols=None, ord_cols=None):
    """lineitem ⋈ orders on l_orderkey = o_orderkey (BASELINE config 2), Agg on top.
    kind "count":  SELECT count(*)
    kind "q3ish":  SELECT o_orderstatus, count(*), sum(l_extendedprice * (1 - l_discount)), min(o_orderdate)
                   WHERE o_orderdate < date '1995-03-15' AND l_shipdate > o_orderdate  (join qual)  GROUP BY o_orderstatus
    Var numbering: varno 0 = outer (lineitem, the probe side), varno 1 = inner (orders, hashed)."""
    cols = li_cols or (LI_WIDE_COLS if table == capi.TAB_LINEITEM_WIDE else LI_NARROW_COLS)
    ocols = ord_cols or ORDERS_COLS
    li_desc = li_desc or capi.synth_tupdesc(table)
    ord_desc = ord_desc or capi.synth_tupdesc(capi.TAB_ORDERS)
    p = ExprPool()
    lkey = p.var(cols["orderkey"], capi.INT8OID, varno=0)
    okey = p.var(ocols["orderkey"], capi.INT8OID, varno=1)
    if kind == "count":
        outer = capi.make_scan(li_desc, -1)
        inner = capi.make_scan(ord_desc, -1)
        hj = capi.make_hashjoin(jointype, [lkey], [okey])
        agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [], [(AGG_COUNT_STAR, -1)])
        return outer, inner, hj, agg, p.pool
    if kind == "survey":
        # SELECT count(*), sum(o_custkey), sum(l_extendedprice) — count and the integer sum are the bit-exact
        # checks, the float8 sum the 1e-6 one
        custkey = p.var(ocols["custkey"], capi.INT4OID, varno=1)
        price = p.var(cols["extendedprice"], FLOAT8OID, varno=0)
        outer = capi.make_scan(li_desc, -1)
        inner = capi.make_scan(ord_desc, -1)
        hj = capi.make_hashjoin(jointype, [lkey], [okey])
        agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [], [(AGG_COUNT_STAR, -1), (capi.AGG_SUM_INT4, custkey), (AGG_SUM_FLOAT8, price)])
        return outer, inner, hj, agg, p.pool
    odate = p.var(ocols["orderdate"], DATEOID, varno=1)
    ostatus = p.var(ocols["orderstatus"], BPCHAROID, varno=1)
    price = p.var(cols["extendedprice"], FLOAT8OID, varno=0)
    disc = p.var(cols["discount"], FLOAT8OID, varno=0)
    shipdate = p.var(cols["shipdate"], DATEOID, varno=0)
    iqual = p.func(capi.F_DATE_LT, BOOLOID, odate, p.const(DATEOID, D_1995_03_15))
    jqual = p.func(capi.F_DATE_GT, BOOLOID, shipdate, odate)
    one = p.const(FLOAT8OID, 1.0)
    rev = p.func(capi.F_FLOAT8MUL, FLOAT8OID, price, p.func(capi.F_FLOAT8MI, FLOAT8OID, one, disc))
    outer = capi.make_scan(li_desc, -1)
    inner = capi.make_scan(ord_desc, iqual)
    hj = capi.make_hashjoin(jointype, [lkey], [okey], jqual)
    agg = capi.make_agg(capi.AGGSTAGE_NORMAL, [ostatus], [(AGG_COUNT_STAR, -1), (AGG_SUM_FLOAT8, rev), (capi.AGG_MIN_DATE, odate)])
    return outer, inner, hj, agg, p.pool


def synth_spec(table, ncand, seed=42, nsegs=1, seg=0, policy=capi.DIST_RANDOM, norders=None):
    s = capi.gg_synth_spec()
    s.table, s.policy, s.seed, s.ncand = table, policy, seed, ncand
    if norders is None:
        norders = ncand if table == capi.TAB_ORDERS else max(ncand // 4, 1)
    s.norders, s.nsegs, s.seg = norders, nsegs, seg
    return s


def synth_measure(spec, nthreads=None):
    nthreads = nthreads or min(os.cpu_count() or 1, 64)
    nb, nr = C.c_uint64(0), C.c_uint64(0)
    assert capi.host_lib().gg_synth_measure(C.byref(spec), nthreads, C.byref(nb), C.byref(nr)) == 0
    return nb.value, nr.value


def synth_generate(spec, out=None, nthreads=None, measured=None):
    """Generate this segment's pages into a numpy uint8 array (or a caller-provided buffer address).
    measured: (nblocks, nrows) from an earlier synth_measure of the same spec (saves one pass over the candidates)."""
    nthreads = nthreads or min(os.cpu_count() or 1, 64)
    nb, nr = measured if measured is not None else synth_measure(spec, nthreads)
    if out is None:
        out = np.empty(nb * capi.GG_BLCKSZ, dtype=np.uint8)
        ptr = out.ctypes.data_as(C.c_void_p)
    else:
        ptr = C.c_void_p(out)
    nb2, nr2 = C.c_uint64(0), C.c_uint64(0)
    rc = capi.host_lib().gg_synth_generate(C.byref(spec), nthreads, ptr, nb, C.byref(nb2), C.byref(nr2))
    assert rc == 0 and nb2.value == nb
    return out, nb, nr


def final_agg_of(agg, key_typids):
    """The FINAL-stage Agg above a Motion: same aggregates, grpCol carries the key type OIDs."""
    fin = capi.gg_agg()
    C.memmove(C.byref(fin), C.byref(agg), C.sizeof(capi.gg_agg))
    fin.aggstage = capi.AGGSTAGE_FINAL
    for i, t in enumerate(key_typids):
        fin.grpCol[i] = t
    return fin


def q1_exec_plan(b, table=capi.TAB_LINEITEM_WIDE, two_stage=False, interval_days=90, desc=None):
    """The executor plan tree of the headline benchmark (scan + filter + hash aggregate, no ORDER BY):
       one segment     Agg(NORMAL) <- SeqScan
       several         Gather Motion <- Agg(FINAL) <- Redistribute Motion(l_returnflag, l_linestatus) <- Agg(PARTIAL) <- SeqScan
    (expected/tpch500GB.out:1771-1782 without the Sort).  Returns (plan, pool)."""
    from . import executor as ex
    if not two_stage:
        scan, agg, pool = q1_plan(table, capi.AGGSTAGE_NORMAL, interval_days, desc)
        return b.agg(b.seqscan(0, scan.desc, scan.qual), agg), pool
    # this engine's own FINAL stage combines the partial rows: float8_avg never reads sumX2 (float.c:1982-1996)
    scan, part, pool = q1_plan(table, capi.AGGSTAGE_PARTIAL, interval_days, desc, flags=capi.AGGF_DEVICE_FINAL)
    fin = final_agg_of(part, [BPCHAROID, BPCHAROID])
    return b.motion(b.agg(b.motion(b.agg(b.seqscan(0, scan.desc, scan.qual), part), ex.MOTION_HASH, [0, 1], 1), fin), ex.MOTION_GATHER, [], 2), pool


def rjoin_exec_plan(b, kind="q3ish", table=capi.TAB_LINEITEM_NARROW, redistribute=True):
    """BASELINE config 3: lineitem JOIN orders with both sides redistributed on the join key, aggregated in two stages:
         Agg(FINAL) <- Gather Motion <- Agg(PARTIAL) <- HashJoin( Redistribute Motion <- SeqScan(lineitem),
                                                                  Hash <- Redistribute Motion <- SeqScan(orders) )
    (nodeMotion.c:1481-1687, nodeHashjoin.c:78-509).  The scans project the columns the join needs (targets), so only those
    travel.  relations: [lineitem, orders].  redistribute=False: the same join over the base relations (co-located data).
    Returns (plan, pool, lineitem_targets, orders_targets)."""
    from . import executor as ex
    ldesc, odesc = capi.synth_tupdesc(table), capi.synth_tupdesc(capi.TAB_ORDERS)
    lc = LI_WIDE_COLS if table == capi.TAB_LINEITEM_WIDE else LI_NARROW_COLS
    if kind == "survey":
        lnames, ltypes = ["orderkey", "extendedprice"], [capi.INT8OID, FLOAT8OID]
        onames, otypes = ["orderkey", "custkey"], [capi.INT8OID, capi.INT4OID]
    else:
        lnames, ltypes = ["orderkey", "extendedprice", "discount", "shipdate"], [capi.INT8OID, FLOAT8OID, FLOAT8OID, DATEOID]
        onames, otypes = ["orderkey", "orderdate", "orderstatus"], [capi.INT8OID, DATEOID, BPCHAROID]
    if not redistribute:
        outer, inner, hj, agg, pool = join_plan(table, kind)
        part = capi.gg_agg.from_buffer_copy(bytes(agg))
        part.aggstage, part.flags = capi.AGGSTAGE_PARTIAL, capi.AGGF_DEVICE_FINAL
        fin = final_agg_of(part, [BPCHAROID] if part.numCols else [])
        join = b.hashjoin(b.seqscan(0, outer.desc, outer.qual), b.hash(b.seqscan(1, inner.desc, inner.qual)), hj)
        return b.agg(b.motion(b.agg(join, part), ex.MOTION_GATHER, [], 3), fin), pool, [], []
    lrd = capi.rows_tupdesc(ltypes, notnull=[1] * len(ltypes))
    ord_ = capi.rows_tupdesc(otypes, notnull=[1] * len(otypes))
    outer, inner, hj, agg, pool = join_plan(table, kind, li_desc=lrd, ord_desc=ord_,
                                            li_cols={n: i + 1 for i, n in enumerate(lnames)}, ord_cols={n: i + 1 for i, n in enumerate(onames)})
    ep = ExprPool()
    ep.pool = pool                                    # the plan's one expression pool: the scans' target lists go into it too
    lt = [ep.var(lc[n], t) for n, t in zip(lnames, ltypes)]
    ot = [ep.var(ORDERS_COLS[n], t) for n, t in zip(onames, otypes)]
    oqual = -1
    if kind == "q3ish":                               # the inner scan qual sits on the base SeqScan, below the Motion
        oqual = ep.func(capi.F_DATE_LT, BOOLOID, ep.var(ORDERS_COLS["orderdate"], DATEOID), ep.const(DATEOID, D_1995_03_15))
    part = capi.gg_agg.from_buffer_copy(bytes(agg))
    part.aggstage, part.flags = capi.AGGSTAGE_PARTIAL, capi.AGGF_DEVICE_FINAL
    fin = final_agg_of(part, [BPCHAROID] if part.numCols else [])
    lside = b.motion(b.seqscan(0, ldesc, -1, targets=lt), ex.MOTION_HASH, [0], 1)
    oside = b.motion(b.seqscan(1, odesc, oqual, targets=ot), ex.MOTION_HASH, [0], 2)
    join = b.hashjoin(lside, b.hash(oside), hj)
    return b.agg(b.motion(b.agg(join, part), ex.MOTION_GATHER, [], 3), fin), pool, lt, ot
