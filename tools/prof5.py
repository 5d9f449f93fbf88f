import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import bench
from xmap_b200 import engine as E, extend as X
from xmap_b200.engine import to_device_meta
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
tabs = eng.run()
plan = X.build_plan(tabs, lay.item_stats[:,3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
print("units", xe.n_units, "starts", plan.start_item.numel(), "table GB total", float(xe.start_bytes.sum())/2**30, "max G", int(xe.G.max()))
torch.cuda.synchronize(); t=time.perf_counter()
nb=0
for sel in xe._batches():
    nb+=1
print("batches", nb)
res = xe.run()
torch.cuda.synchronize(); print("xsim run s", time.perf_counter()-t, "launches", xe.launches)
