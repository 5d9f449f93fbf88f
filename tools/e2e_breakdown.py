"""Where the end-to-end time of one similarity stage goes (host wall-clock with a sync per phase)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E
from xmap_b200.engine import to_device_meta
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
pin = lambda a: torch.from_numpy(a).pin_memory()
hu, hi, hr = pin(wl["user"]), pin(wl["item"]), pin(wl["rating"])
def T(label, fn, acc):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t) * 1e3; return r
for rep in range(3):
    acc = {}
    du, di, dr = T("h2d", lambda: (hu.to(dev, non_blocking=True), hi.to(dev, non_blocking=True), hr.to(dev, non_blocking=True)), acc)
    lay = T("build_layout", lambda: E.build_layout(du, di, dr, wl["n_users"], wl["n_items"], device=dev), acc)
    eng = T("engine_init(tri layout, lists)", lambda: E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]), acc)
    T("plan", lambda: eng.plan(), acc)
    tabs = T("stage", lambda: eng.run(), acc)
    out = T("d2h pinned", lambda: eng.tables_to_host(tabs, reuse=True), acc)
    print(rep, {k: round(v, 1) for k, v in acc.items()}, "total %.1f" % sum(acc.values()), flush=True)
    del lay, eng, tabs, out
