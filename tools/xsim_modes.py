"""Both X-SIM kernels on one plan: per-start path counts against the plan's exact bound, distinct-end counts and
top-m against each other."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
dev = torch.device("cuda"); meta = E.to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
print("plan ub sum", int(plan.ub.sum()))
out = {}
for mode in ("hybrid", "warp", "cta"):
    xe = X.XsimEngine(plan, 10, mode=mode)
    torch.cuda.synchronize(); t = time.perf_counter()
    res = xe.run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    bad = (res.combos != plan.ub).nonzero().flatten()
    print(mode, "ms %.1f" % (dt * 1e3), "combos", int(res.combos.sum()), "cells", int(res.count.sum()), "starts with combos != ub:", int(bad.numel()))
    for x in bad[:5].tolist():
        print("   start", x, "ub", int(plan.ub[x]), "combos", int(res.combos[x]), "T", int(xe.T[x]), "units", int(xe.n_units_x[x]), "count", int(res.count[x]))
    out[mode] = res
a, b = out["hybrid"], out["cta"]
dc = (a.count != b.count).nonzero().flatten()
print("starts with different distinct-end counts:", int(dc.numel()))
for x in dc[:5].tolist():
    print("   start", x, "hybrid", int(a.count[x]), "cta", int(b.count[x]), "ub", int(plan.ub[x]))
print("top_end equal rows:", int((a.top_end == b.top_end).all(1).sum()), "of", a.top_end.shape[0],
      "max |xsim| diff", float((a.top_xsim - b.top_xsim).abs().max()))
