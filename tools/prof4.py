import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import bench
from xmap_b200 import engine as E, extend as X
from tests.parity import to_device_meta
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
tabs = eng.run()
w = lay.row_work
for th in (64, 170, 350, 700, 1400, 2800, 5600, 20000, 100000):
    m = (w <= th) & (w > 0)
    print("rows w<=%d: %d  work share %.3f" % (th, int(m.sum()), float(w[m].sum())/float(w.sum())))
print("BB items", int((tabs.row_flags&1).sum()), "of", wl["n_items"])
torch.cuda.synchronize(); t=time.perf_counter()
plan = X.build_plan(tabs, lay.item_stats[:,3].contiguous(), meta.has_S, meta.has_T)
torch.cuda.synchronize(); print("plan s", time.perf_counter()-t)
ub = plan.ub.double()
q = torch.tensor([0.5,0.9,0.99,0.999,1.0], dtype=torch.float64, device=dev)
print("starts", ub.numel(), "sum ub %.3e" % float(ub.sum()), "quantiles", [int(x) for x in torch.quantile(ub, q)])
for th in (1e5, 1e6, 1e7, 1e8):
    m = ub > th
    print(" starts with ub>%g: %d holding %.3f of combos" % (th, int(m.sum()), float(ub[m].sum()/ub.sum())))
lc = (plan.leg_ptr[1:]-plan.leg_ptr[:-1]).double()
print("legs total", int(lc.sum()), "legs/start quantiles", [int(x) for x in torch.quantile(lc, q)])
R = (plan.rs_ptr[1:]-plan.rs_ptr[:-1]).double(); pc = (plan.par_ptr[1:]-plan.par_ptr[:-1]).double()
print("R_s quantiles", [int(x) for x in torch.quantile(R, q)], "n_s", R.numel(), "partners/t quantiles", [int(x) for x in torch.quantile(pc, q)], "n_t", pc.numel())
# per-leg ub
rt_all = torch.zeros(plan.t_items.numel(), dtype=torch.float64, device=dev)
t_of_pair = torch.repeat_interleave(torch.arange(plan.t_items.numel(), device=dev), (plan.par_ptr[1:]-plan.par_ptr[:-1]))
rt_all.index_add_(0, t_of_pair, R[plan.par_s.long()])
print("RT_all per t quantiles", [int(x) for x in torch.quantile(rt_all, q)])
