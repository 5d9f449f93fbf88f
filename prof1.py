import sys, time, json; sys.path.insert(0,".")
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
eng.run()
(t0,t1,big), t_plan = T(eng.plan)
w = lay.row_work
print("plan ms", t_plan, "rows", t0.numel(), t1.numel(), big.numel(), "work share", float(w[t0.long()].sum())/float(w.sum()), float(w[t1.long()].sum())/float(w.sum()), float(w[big.long()].sum())/float(w.sum()))
empty = torch.zeros(0, dtype=torch.int32, device=dev)
a0 = eng._args(0)
for rep in range(2):
    _, ta = T(eng._run_rows, a0, t0, empty, empty)
    _, tb = T(eng._run_rows, a0, empty, t1, empty)
    _, tc = T(eng._run_rows, a0, empty, empty, big)
    print("pass1 tier0 %.2f ms tier1 %.2f ms big %.2f ms" % (ta, tb, tc))
# split big into accumulate vs finalize using events inside: monkeypatch
L = N.lib()
acc_ms = [0.0]; fin_ms=[0.0]
orig_acc, orig_fin = L.xmap_sim_big_accumulate, L.xmap_sim_big_finalize; orig_sb = L.xmap_sim_big_scratch_bytes
class W:
    def __init__(s, f, store): s.f=f; s.store=store
    def __call__(s, *a):
        torch.cuda.synchronize(); t=time.perf_counter(); r=s.f(*a); torch.cuda.synchronize(); s.store[0]+= (time.perf_counter()-t)*1e3; return r
class LW:
    def __getattr__(s, n):
        if n=="xmap_sim_big_accumulate": return W(orig_acc, acc_ms)
        if n=="xmap_sim_big_finalize": return W(orig_fin, fin_ms)
        return getattr(L, n)
N._lib = LW()
_, tc = T(eng._run_rows, a0, empty, empty, big)
print("big total %.2f: accumulate %.2f finalize %.2f host/other %.2f" % (tc, acc_ms[0], fin_ms[0], tc-acc_ms[0]-fin_ms[0]))
N._lib = L
for budget in (1<<30, 16<<30):
    eng2 = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"], table_budget=budget)
    eng2._run_rows(eng2._args(0), empty, empty, big)
    _, tc = T(eng2._run_rows, eng2._args(0), empty, empty, big)
    print("budget", budget>>30, "GB big ms", tc, "launches", eng2.launches)
s, t2 = T(eng.pass2, eng.row_flags)
print("pass2 ms", t2, s)
bigw = w[big.long()]
print("big rows work quantiles", [int(x) for x in torch.quantile(bigw.double(), torch.tensor([0,.5,.9,.99,1.0], dtype=torch.float64, device=dev))])
print("npairs big", int(eng.row_npairs[big.long()].sum()), "t1", int(eng.row_npairs[t1.long()].sum()), "t0", int(eng.row_npairs[t0.long()].sum()))
