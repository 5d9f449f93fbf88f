import sys, time, ctypes; sys.path.insert(0,".")
import numpy as np, torch
import bench
from xmap_b200 import _native as N
N.LIB_PATH = N.LIB_PATH.replace("libxmap_b200.so", "libxmap_b200_dbg.so")
from xmap_b200 import engine as E
from tests.parity import to_device_meta
wl = bench.make_workload("cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
t0,t1,big = eng.plan()
empty = torch.zeros(0, dtype=torch.int32, device=dev)
L = N.lib(); L.xmap_debug_phase_cycles.argtypes=[ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong*8)()
a0 = eng._args(0)
for name, rows in (("tier0", t0), ("tier1", t1)):
    L.xmap_debug_phase_cycles(buf, 1)
    torch.cuda.synchronize(); t=time.perf_counter()
    eng._run_rows(a0, rows if name=="tier0" else empty, rows if name=="tier1" else empty, empty)
    torch.cuda.synchronize(); ms=(time.perf_counter()-t)*1e3
    L.xmap_debug_phase_cycles(buf, 1)
    n = rows.numel()
    print(name, "rows", n, "ms %.2f"%ms, "avg cycles/row by phase [zero, accumulate, compact, eval, finalize]:", [int(buf[i]/n) for i in range(5)])
