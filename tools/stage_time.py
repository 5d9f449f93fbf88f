import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from xmap_b200 import engine as E
from xmap_b200.engine import to_device_meta
from xmap_b200 import _native as _N
print("lib", _N.LIB_PATH, flush=True)
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
for classes in ("256,512,1024,1536,2048,3072,4096,6144,8192,12288",):
    for ns in ("1", "3"):
        os.environ["XMAP_SIM_STREAMS"] = ns
        E.CELL_CLASSES = tuple(int(c) for c in classes.split(","))
        eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"], max_smem_cells=E.CELL_CLASSES[-1])
        for _ in range(3): eng.run()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): tabs = eng.run()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
        print("classes", classes, "streams", ns, "stage %.2f ms" % (dt * 1e3), "launches/stage", len(eng.plan()[0]), flush=True)
        if ns == "1":
            eng.enable_profile(); eng.run()
            print("   ", {k: round(v[1], 2) for k, v in eng.profile_ms().items()}); eng.profile = None
        del eng
