import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from greengage_b200 import capi, tpch
from greengage_b200.engine import Engine, Relation, ScanAgg
from _util import f2b
eng = Engine(0)
pages, nb, nr = tpch.synth_generate(tpch.synth_spec(capi.TAB_LINEITEM_NARROW, 300000, seed=3))
scan, agg, pool = tpch.q1_plan(capi.TAB_LINEITEM_NARROW)
def run(host=False, newrel=False):
    sa = ScanAgg(eng, scan, agg, pool)
    rel = Relation(eng, host_pages=pages)
    if host: sa.run_host(pages.ctypes.data, nb)
    else: sa.run(rel)
    rows, sc, ps = sa.fetch()
    out = {(r.key[0], r.key[1]): [f2b(r.agg[i].f[0]) for i in range(7)] for r in rows}
    sa.free(); rel.free()
    return out
a = run(); b = run(); c = run(host=True); d = run(host=True)
print("resident==resident", a == b, " host==host", c == d, " resident==host", a == c)
for k in a:
    if a[k] != c[k]: print(k, [x - y for x, y in zip(a[k], c[k])])
