"""Warm X-SIM timing on cfg2 (second run of the engine)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X, _native as N
from tests.parity import to_device_meta
print("lib", N.LIB_PATH)
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    xe = X.XsimEngine(plan, 10); res = xe.run()
    torch.cuda.synchronize(); print("xsim %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    del xe
print("checksum", int(res.top_end.long().sum()), float(res.top_xsim.sum()))
