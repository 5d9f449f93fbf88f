"""Distribution of X-SIM work per start at cfg2 (paths = ub, distinct ends = count)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xmap_b200 import engine as E, extend as X
from xmap_b200.engine import to_device_meta
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
tabs = eng.run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
res = xe.run()
ub = plan.ub.double(); cnt = res.count.double()
tot = float(ub.sum())
print("starts", ub.numel(), "paths", tot, "cells", float(cnt.sum()), "legs", plan.leg_t.numel(), "partners", plan.par_s.numel(),
      "rsegs", plan.rs_end.numel(), "n_s", plan.rs_ptr.numel() - 1)
for thr in (1e3, 2e3, 4e3, 8e3, 16e3, 32e3, 64e3, 128e3, 256e3, 1e6, 1e9):
    m = ub <= thr
    print("paths<=%g: starts %.3f paths %.3f cells %.3f" % (thr, float(m.double().mean()), float(ub[m].sum()) / tot, float(cnt[m].sum()) / float(cnt.sum())))
for thr in (1e3, 2e3, 4e3, 8e3, 16e3, 32e3, 64e3):
    m = cnt <= thr
    print("cells<=%g: starts %.3f paths %.3f" % (thr, float(m.double().mean()), float(ub[m].sum()) / tot))
rl = (plan.rs_ptr[1:] - plan.rs_ptr[:-1]).double()
print("rseg list len mean %.1f max %d ; path-weighted mean len ~ %.1f" % (float(rl.mean()), int(rl.max()), float((rl * rl).sum() / rl.sum())))
lc = (plan.leg_ptr[1:] - plan.leg_ptr[:-1]).double()
print("legs per start mean %.2f max %d" % (float(lc.mean()), int(lc.max())))
pc = (plan.par_ptr[1:] - plan.par_ptr[:-1]).double()
print("partners per t mean %.2f max %d" % (float(pc.mean()), int(pc.max())))
# duplicate ends inside one right-segment list, and inside the fan of one bridge target
s_id = torch.repeat_interleave(torch.arange(plan.rs_ptr.numel() - 1, device=dev), (plan.rs_ptr[1:] - plan.rs_ptr[:-1]))
key = s_id * plan.n_items + plan.rs_end.long()
print("rsegs %d distinct (s,end) %d" % (key.numel(), torch.unique(key).numel()))
w = torch.zeros(plan.rs_ptr.numel() - 1, dtype=torch.float64, device=dev)   # how often each s list is walked
t_of_par = torch.repeat_interleave(torch.arange(plan.par_ptr.numel() - 1, device=dev), (plan.par_ptr[1:] - plan.par_ptr[:-1]))
legs_into_t = torch.bincount(plan.leg_t.long(), minlength=plan.par_ptr.numel() - 1).double()
w.index_add_(0, plan.par_s.long(), legs_into_t[t_of_par])
uk, inv = torch.unique(key, return_inverse=True)
dist_per_s = torch.bincount(torch.div(uk, plan.n_items, rounding_mode="floor"), minlength=w.numel()).double()
print("walk-weighted: paths %.4g, after per-list dedupe %.4g" % (float((w * rl).sum()), float((w * dist_per_s).sum())))
# legs of one start that share a bridge target: their paths hit the same ends
leg_start = torch.repeat_interleave(torch.arange(plan.start_item.numel(), device=dev), (plan.leg_ptr[1:] - plan.leg_ptr[:-1]))
lk = leg_start * (plan.par_ptr.numel()) + plan.leg_t.long()
uq, cnts = torch.unique(lk, return_counts=True)
print("legs %d distinct (start,t) %d" % (lk.numel(), uq.numel()))
# path-weighted: paths per (start,t) group = group_size * fan(t); updates after grouping = fan(t)
fan = torch.zeros(plan.par_ptr.numel() - 1, dtype=torch.float64, device=dev)
fan.index_add_(0, t_of_par, rl[plan.par_s.long()])
t_of_group = (uq % plan.par_ptr.numel())
print("paths(all partners) %.4g -> grouped updates %.4g" % (float((cnts.double() * fan[t_of_group]).sum()), float(fan[t_of_group].sum())))
jo = plan.leg_joint_only.bool()
print("joint-only legs %d of %d" % (int(jo.sum()), jo.numel()))
