"""One GPU: the X-SIM kernels on every rank's share of the units of a `world`-way run, one after the other
(what a multi-GPU run would take per rank, without the collectives), then the whole plan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E, extend as X, _native as N
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda"); meta = E.to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
tabs = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"]).run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
keep = []
a = xe._args(keep)
nu, m = xe.n_units, xe.top_m
z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
bufs = [z(nu, torch.int32), z(nu, torch.int64), z((nu, m), torch.int32), z((nu, m), torch.float64), z(nu, torch.int32)]
a.unit_count, a.unit_combos, a.unit_top_end, a.unit_top_xsim, a.unit_top_len = [N.ptr(t) for t in bufs]
st = torch.cuda.current_stream().cuda_stream
print("mode %s: %d hot units (est. paths > %.3g), %d cold" % (xe.mode, xe.hot_order.numel(), xe.hot_paths, xe.cold_order.numel()))
times = []
for w_, ranks in ((world, range(world)), (1, [0])):
    for rank in ranks:
        for rep in range(2):
            for t in bufs: t.zero_()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            xe._launch_units(a, rank, w_, st)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print("share %d/%d: %.1f ms, paths %d" % (rank, w_, dt, int(bufs[1].sum())), flush=True)
        times.append(dt)
print("max over the %d shares %.1f ms, sum %.1f ms; whole plan on one GPU %.1f ms" % (world, max(times[:-1]), sum(times[:-1]), times[-1]))
