"""Warm X-SIM timing on a bench workload: engine build (host/torch index work) and the kernels separately."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X, _native as N
from xmap_b200.engine import to_device_meta
print("lib", N.LIB_PATH)
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    xe = X.XsimEngine(plan, 10)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = xe.run()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("plan %.1f ms  engine %.1f ms  kernels %.1f ms   units %d gb %d Tmax %d" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, xe.n_units, xe.gb, int(xe.T.max())), flush=True)
combos = int(res.combos.sum())
print("paths %d  cells %d  paths/s %.3g" % (combos, int(res.count.sum()), combos / (t3 - t2)))
print("checksum", int(res.top_end.long().sum()), float(res.top_xsim.sum()))
