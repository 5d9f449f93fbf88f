"""One similarity stage on cfg2 (or the workload named on the command line), for ncu:
  ncu --set full --import-source on --clock-control none -k regex:'tri_|select_' -o gpurun_out/sim python tools/prof_sim.py
Prints the per-launch CUDA-event times of a second, unprofiled-looking pass (still under ncu when
run that way: never a bench number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xmap_b200 import engine as E
from xmap_b200.engine import to_device_meta

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = bench.make_workload(name)
dev = torch.device("cuda")
meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
for _ in range(reps):
    eng.enable_profile()
    torch.cuda.synchronize(); t = time.perf_counter()
    tabs = eng.run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("stage wall %.2f ms" % (dt * 1e3))
    for k, (n, ms) in eng.profile_ms().items():
        print("  %-28s %8.3f ms" % (k, ms))
print("pairs", tabs.n_pairs_total, "kept", int(tabs.row_nkept.sum()))
