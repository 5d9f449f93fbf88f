"""X-SIM variants on ONE bench workload (built once): engine build and kernel time per configuration, path counts
against the plan's exact bound, distinct ends and top-m rows against the first configuration.

usage: xsim_sweep.py <workload> "<mode> <fuse> <cells_lg> <unit_lg> [rho] [load] [warps]" ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X

wl = bench.make_workload(sys.argv[1])
dev = torch.device("cuda"); meta = E.to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
torch.cuda.synchronize(); t0 = time.perf_counter()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
torch.cuda.synchronize()
print("plan %.1f ms: starts %d legs %d (joint-only %d) src pairs %d joint %d  t %d s %d  rs entries %d  paths %d" % (
    (time.perf_counter() - t0) * 1e3, plan.start_item.numel(), plan.leg_t.numel(), int(plan.leg_joint_only.sum()),
    plan.n_src, plan.n_joint, plan.t_items.numel(), plan.s_items.numel(), plan.rs_end.numel(), int(plan.ub.sum())), flush=True)
first = None
for spec in sys.argv[2:]:
    f = spec.split()
    mode, fuse, clg, ulg = f[0], f[1] != "0", int(f[2]), int(f[3])
    rho = float(f[4]) if len(f) > 4 else X.XSIM_RHO
    load = float(f[5]) if len(f) > 5 else X.XSIM_LOAD
    kw = {}
    if len(f) > 6:
        kw["warps"] = int(f[6])
    best = None
    for rep in range(2):
        torch.cuda.synchronize(); t1 = time.perf_counter()
        xe = X.XsimEngine(plan, 10, cells_lg=clg, unit_work=1 << ulg, rho=rho, load=load, mode=mode, fuse=fuse, **kw)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        res = xe.run()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        if best is None or t3 - t2 < best[1]:
            best = (t2 - t1, t3 - t2)
    bad = int((res.combos != plan.ub).sum())
    line = "%-28s engine %.1f ms  kernels %.1f ms  fused %d  lp pairs %d  units %d gb %d  bad-combos %d" % (
        spec, best[0] * 1e3, best[1] * 1e3, xe.fused_entries, int(xe.lp_ptr[-1]), xe.n_units, xe.gb, bad)
    if first is None:
        first = res
    else:
        line += "  count-diff %d  top-rows-equal %d/%d  max|dx| %.3g" % (
            int((res.count != first.count).sum()), int((res.top_end == first.top_end).all(1).sum()), res.top_end.shape[0],
            float((res.top_xsim - first.top_xsim).abs().max()))
    print(line, flush=True)
    del xe, res
    torch.cuda.empty_cache()
