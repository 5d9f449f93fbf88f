"""One X-SIM run on a bench workload (for ncu: -k regex:xsim_tile)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X
from xmap_b200.engine import to_device_meta
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2_small")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
torch.cuda.synchronize(); t = time.perf_counter()
res = xe.run()
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("xsim %.1f ms, paths %d, cells %d, units %d, passes %d, gb %d" % (dt * 1e3, int(res.combos.sum()), int(res.count.sum()),
      xe.n_units, int(xe.T.sum()), xe.gb))
