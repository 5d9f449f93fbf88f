import sys, time, json; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import bench
from xmap_b200 import engine as E, _native as N
from tests.parity import to_device_meta
wl = bench.make_workload(sys.argv[1] if len(sys.argv)>1 else "cfg2")
dev = torch.device("cuda")
meta = to_device_meta(wl["meta"], dev)
def T(f, *a, **k):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(*a, **k); torch.cuda.synchronize(); return r, (time.perf_counter()-t)*1e3
lay, t_lay = T(E.build_layout, wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
lay, t_lay = T(E.build_layout, wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
print("layout ms", t_lay)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
eng.run(); eng.run()
(tiers, big), t_plan = T(eng.plan)
w = lay.row_work
print("plan ms", t_plan, "rows", [t.numel() for t in tiers], big.numel(), "work share", [round(float(w[t.long()].sum())/float(w.sum()),3) for t in tiers], round(float(w[big.long()].sum())/float(w.sum()),3))
empty = torch.zeros(0, dtype=torch.int32, device=dev)
a0 = eng._args(0)
for rep in range(2):
    ts = []
    for i in range(len(tiers)):
        sel = [tiers[q] if q == i else empty for q in range(len(tiers))]
        _, t = T(eng._run_rows, a0, sel, empty); ts.append(round(t,2))
    _, tc = T(eng._run_rows, a0, [empty]*len(tiers), big)
    print("pass1 tiers ms", ts, "big %.2f ms" % tc)
for rep in range(2):
    _, t1 = T(eng.pass1); _, t2 = T(eng.pass2, eng.row_flags)
    print("pass1 total %.2f ms  pass2 %.2f ms" % (t1, t2))
