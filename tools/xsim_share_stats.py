"""How often the legs of one start walk the SAME list (legs (x, n, t) through different n reach the same bridge target t):
paths / (list entries that would be read if a list were walked once per start)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
dev = torch.device("cuda"); meta = E.to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
n_lists = int(xe.rs_ptr.numel()) - 1
L = (xe.rs_ptr[1:] - xe.rs_ptr[:-1])
pairs_x = xe.lp_ptr[plan.leg_ptr[1:]] - xe.lp_ptr[plan.leg_ptr[:-1]]
start_of_pair = torch.repeat_interleave(torch.arange(pairs_x.numel(), device=dev), pairs_x)
s = xe.pd_s.long()
key = start_of_pair * n_lists + s
uk, cnt = torch.unique(key, return_counts=True)
paths = float(L[s].sum()); loads = float(L[uk % n_lists].sum())
print("pairs %d, distinct (start, list) %d: %.2f legs per walked list on average" % (key.numel(), uk.numel(), key.numel() / uk.numel()))
print("paths %.4g, entries read if every (start, list) were walked once %.4g: %.2f paths per entry read" % (paths, loads, paths / loads))
w = L[uk % n_lists].double() * cnt.double()          # paths of the group
q = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.99], dtype=torch.float64, device=dev)
o = torch.argsort(cnt); cw = torch.cumsum(w[o], 0) / w.sum()
print("path-weighted quantiles of the group size:", [int(cnt[o][torch.searchsorted(cw, qq)].item()) for qq in q])
heavy = plan.ub > 1e6
hx = heavy[torch.div(uk, n_lists, rounding_mode="floor")]
print("heavy starts (> 1e6 paths): %.2f paths per entry read; others %.2f" % (
    float(w[hx].sum() / L[uk % n_lists][hx].double().sum()), float(w[~hx].sum() / L[uk % n_lists][~hx].double().sum())))
