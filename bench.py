#!/usr/bin/env python
"""Benchmark of the AlterEgo-construction hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|...]

A "step" is one pass of the similarity stage (all item rows: co-rating SpGEMM +
epilogue + top-k selection, passes 1 and 2) over the synthetic two-domain
workload; `value` = directed co-rated item pairs evaluated per second with the
ratings layout already resident in HBM.  `e2e` is the same metric through the
host-facing call: pinned host triples -> H2D -> layout build -> similarity ->
D2H of the neighbour tables, all inside the timed region.  The line also
carries the pipeline wall-time (similarity + X-SIM extension + generation), the
roofline of the dominant kernel and a CPU baseline (the oracle port, timed on a
bounded sample on this box's host cores).

Under torchrun (N > 1) every rank holds the replicated layout, owns a
work-balanced block of item rows, and the ranks exchange BB flags and neighbour
tables with NCCL all-gathers; timing is max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (users, items/domain, draws, domains, overlap, k)
    "cfg2": (1_000_000, 200_000, 20_000_000, 2, 0.05, 10),
    "cfg2_small": (100_000, 20_000, 2_000_000, 2, 0.05, 10),
    "tiny": (20_000, 3_000, 400_000, 2, 0.05, 10),
}
METRIC = "item_pair_sims_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
# (profiles/), keyed by kernel group; None until captured.
NCU_TRAFFIC = {"tri_cta_kernel": 5.857e9 / 14}   # (2.883 GB read + 2.974 GB written) over the 14 launches of one stage,
                                                 # profiles/r1_ncu_full_cfg2.csv; per launch like `achieved`


def make_workload(name):
    from xmap_b200 import synth
    from xmap_b200.encode import item_codes
    nu, ni, nd, ndom, ov, k = WORKLOADS[name]
    sr = synth.make_ratings(nu, ni, nd, n_domains=ndom, overlap=ov)
    present = np.unique(sr.item)
    iids = np.array([synth.item_id(int(g), sr.n_items_per_domain, sr.labels) for g in present])
    order = np.argsort(iids)
    present, iids = present[order], iids[order]
    imap = np.full(int(sr.item.max()) + 1, -1, dtype=np.int64)
    imap[present] = np.arange(len(present))
    uu = np.unique(sr.user)
    umap = np.full(int(sr.user.max()) + 1, -1, dtype=np.int64)
    umap[uu] = np.arange(len(uu))
    pc, dc, ct, hs, ht = item_codes(iids)
    du = np.bincount(umap[sr.user]).astype(np.int64)
    return dict(name=name, user=umap[sr.user].astype(np.int32), item=imap[sr.item].astype(np.int32),
                rating=sr.rating.astype(np.float32), ts=sr.ts, n_users=len(uu), n_items=len(iids), k=k,
                meta=dict(prefix_code=pc, dom_code=dc, contains=ct, has_S=hs, has_T=ht),
                nnz=int(len(sr.user)), W=int((du * (du - 1)).sum()))


def workload_label(wl, method):
    return "%s: %d users, %d items, %d ratings (Zipf), W=%d products, top-k=%d, %s" % (
        wl["name"], wl["n_users"], wl["n_items"], wl["nnz"], wl["W"], wl["k"], method)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max, self.stop_flag = index, [], set(), None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        self.sm.append(float(f[0])); self.sm_max = float(f[1])
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                self._nvml_sample() if self.nvml else self._smi_sample()
            except Exception:
                pass
            time.sleep(0.005 if self.nvml else 0.1)

    def summary(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
# CPU baseline: the oracle port on a bounded sample of the same workload
# --------------------------------------------------------------------------
def _cpu_block(args):
    from oracle import restate as RS
    user, item, rating, nU, nI, pc, method, na = args
    t0 = time.perf_counter()
    P = RS.sim_pairs(user, item, rating, nU, nI, pc, method, na)
    return P["n_pairs_total"], len(P["i"]), time.perf_counter() - t0


def _sample(wl, phase, stride):
    """Every `stride`-th user starting at `phase` (a uniform sample of light and heavy users),
    with users and items renumbered densely."""
    m = (wl["user"] % stride) == phase
    user, item = wl["user"][m].astype(np.int64) // stride, wl["item"][m].astype(np.int64)
    n_u = int(user.max()) + 1 if len(user) else 1
    present = np.unique(item)
    imap = np.full(wl["n_items"], -1, dtype=np.int64); imap[present] = np.arange(len(present))
    return (user, imap[item], wl["rating"][m].astype(np.float64), n_u, len(present),
            wl["meta"]["prefix_code"][present]), int(m.sum())


def cpu_baseline(wl, method="adjust_cosine", num_atleast=50, target_ratings=250_000, cores=1):
    """oracle/restate.py (numpy/scipy restatement of baselinerSim.py:17-216) on a bounded sample of
    the workload: `cores` disjoint strided user samples of ~target_ratings ratings each, one process
    per sample (the reference's per-task arithmetic is single-threaded)."""
    stride = max(cores, int(round(wl["nnz"] / max(1, target_ratings))))
    jobs, nr = [], 0
    for c in range(cores):
        args, r = _sample(wl, c, stride)
        jobs.append(args + (method, num_atleast)); nr += r
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_block(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_block, jobs)
    dt = time.perf_counter() - t0
    n_pairs = sum(r[0] for r in res)
    return {"value": n_pairs / dt, "unit": "item pairs/s", "cores": cores, "kind": "port",
            "sample": "%d sample(s), every %d-th user of %s (%d ratings, %d co-rated pairs, %.1f s wall): "
                      "oracle/restate.py, the numpy/scipy restatement of the reference arithmetic "
                      "(no Spark/JVM/shuffle serialisation)" % (cores, stride, wl["name"], nr, n_pairs, dt)}


def run_reference_arm(args, emit):
    """--impl reference: the reference's CPU arithmetic for the path on this box's host cores.
    /root/reference (pure Python on Spark) does not travel to the GPU box and Spark is absent,
    so the oracle port (oracle/restate.py) is what is timed, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload)
    vals, secs = [], []
    cores = min(os.cpu_count() or 1, 32)
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(wl, method=args.method, target_ratings=150_000, cores=cores)
        if s >= args.warmup:
            vals.append(cb["value"]); secs.append(time.perf_counter() - t0)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "item pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(wl, args.method), "note": "each step = one bounded sample (see cpu_baseline.sample); "
                       "Spark is not installed on this image, /root/reference (pure Python on Spark RDDs) does "
                       "not travel to the GPU box, so the timed code is the oracle port of its arithmetic"},
            "cpu_baseline": dict(cb, value=v),
            "e2e": {"value": v, "unit": "item pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--method", default="adjust_cosine")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the extension/generation timing")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    # exactly one JSON line may reach stdout: libraries (NCCL's version banner) write there too, so
    # stdout is pointed at stderr for the duration of the run and restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args, emit)

    import torch
    import torch.distributed as dist
    from xmap_b200 import engine as E
    from xmap_b200 import extend as X
    from xmap_b200 import generate as G
    from xmap_b200 import multi as MG
    from xmap_b200.engine import to_device_meta

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    wl = make_workload(args.workload)
    k = wl["k"]
    meta = to_device_meta(wl["meta"], dev)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_user, h_item, h_rating = pin(wl["user"]), pin(wl["item"]), pin(wl["rating"])
    h2d_bytes = h_user.numel() * 4 * 3

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # resident layout for the device-timed metric
    lay = E.build_layout(h_user, h_item, h_rating, wl["n_users"], wl["n_items"], device=dev)
    eng = E.SimEngine(lay, meta, args.method, 50, k)
    shard = MG.similarity_shard(eng, rank, world)

    def sim_step(engine):
        return MG.similarity_step(engine, shard)

    for _ in range(args.warmup):
        sim_step(eng)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launches
    barrier()
    for s in range(args.steps):
        ev[s][0].record()
        tabs = sim_step(eng)
        ev[s][1].record()
    barrier()
    clocks = sampler.summary()
    eng._check_error()
    launches = (eng.launches - l0)
    # per-kernel CUDA-event times: a second set of steps with every launch serialised on one stream
    eng.enable_profile()
    for s in range(args.steps):
        sim_step(eng)
    prof = eng.profile_ms()
    eng.profile = None
    barrier()
    ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.mean())
    P_total = tabs.n_pairs_total
    P_kept = int(tabs.row_nkept.sum().item())
    value = P_total / (ms_step * 1e-3)

    # e2e: host triples -> layout -> similarity -> neighbour tables back on the host
    def e2e_step():
        lay2 = E.build_layout(h_user, h_item, h_rating, wl["n_users"], wl["n_items"], device=dev)
        eng2 = E.SimEngine(lay2, meta, args.method, 50, k)
        t = MG.similarity_step(eng2, MG.similarity_shard(eng2, rank, world))
        out = eng2.tables_to_host(t, reuse=True)                 # pinned host buffers, synchronises
        return sum(o.numel() * o.element_size() for o in out.values())
    e2e_step()
    n_e2e = max(3, min(args.steps, 7))
    e2e_times = []
    for _ in range(n_e2e):
        barrier()
        t0 = time.perf_counter()
        d2h_bytes = e2e_step()
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    # median of the per-step wall-clock times (the host is shared: a single descheduling would skew a mean)
    e2e_s = torch.tensor([float(np.median(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = P_total / float(e2e_s)

    # pipeline wall-time: similarity + X-SIM extension (sharded by start) + generation (sharded by user)
    pipe = None
    for timed in ((False, True, True) if not args.no_pipeline else ()):  # one untimed pass, then the faster of two
        barrier()
        t0 = time.perf_counter()
        plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        xe = X.XsimEngine(plan, 10)
        res = xe.run(rank, world)
        barrier(); t2 = time.perf_counter()
        ch = G.choose_mapping(res, "argmax", sim_method=args.method)
        mp = G.invert_mapping(res.start_item, ch, wl["n_items"])
        ou, oi, orr, ot = MG.build_alterego_sharded(lay, wl["ts"], mp, MG.UserShard(lay.csr_ptr, rank, world))
        n_rec = torch.tensor([ou.numel()], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n_rec)
        barrier(); t3 = time.perf_counter()
        combos = int(res.combos.sum().item())
        if not timed or (pipe is not None and pipe["alterego_pipeline_ms"] <= ms_step + (t3 - t0) * 1e3):
            del plan, xe, res
            continue
        pipe = {"similarity_ms": ms_step, "extend_plan_ms": (t1 - t0) * 1e3, "extend_kernel_ms": (t2 - t1) * 1e3,
                "generate_ms": (t3 - t2) * 1e3,
                "alterego_pipeline_ms": ms_step + (t3 - t0) * 1e3,
                "xsim_paths": combos, "xsim_paths_per_s": combos / max(t2 - t1, 1e-9),
                "xsim_starts": int(res.start_item.numel()), "xsim_pairs": int(res.count.sum().item()),
                "alterego_synthetic_records": int(n_rec.item()), "bridge_pairs": plan.n_src,
                "joint_pairs": plan.n_joint,
                "sharding": "X-SIM by start item x%d, generation by user x%d (host wall-clock of rank 0 between barriers, "
                            "faster of two passes after one untimed pass)" % (world, world)}
        del plan, xe, res

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # the stage evaluates every unordered pair once: W/2 products of 8 B, one pass over the CSC
        # (24 B per rating: entry + suffix extent + rater mean), 2 x 16 B records written and read back per kept pair,
        # and the tables out
        alg_bytes = 8.0 * (wl["W"] / 2) + 24.0 * wl["nnz"] + 32.0 * P_kept + 80.0 * k * wl["n_items"]
        stage_gbs = alg_bytes / (ms_step * 1e-3) / 1e9
        # kernels: the accumulate launches are one kernel function per group width
        launches_plan, _, split_plan = eng.plan(None if world == 1 else shard.rows(dev))
        fam_ms, fam_n, fam_bytes = {}, {}, {}
        for kk, (n, ms_k) in prof.items():
            if kk.startswith("accumulate_c") or kk.startswith("accumulate_g"):
                fam = "tri_warp_kernel" if kk.endswith("_t32") else ("tri_gmem_kernel" if "_g" in kk else "tri_cta_kernel")
            elif kk == "accumulate_split":
                fam = "tri_split_kernel"
            elif kk == "select_warp":
                fam = "select_warp_kernel"
            elif kk == "select_cta":
                fam = "select_cta_kernel"
            else:
                fam = kk
            fam_ms[fam] = fam_ms.get(fam, 0.0) + ms_k
            fam_n[fam] = fam_n.get(fam, 0) + n
        for r, cells_cap, threads, in_gmem, _hdr in launches_plan:
            fam = "tri_gmem_kernel" if in_gmem else ("tri_warp_kernel" if threads == 32 else "tri_cta_kernel")
            fam_bytes[fam] = fam_bytes.get(fam, 0.0) + 8.0 * float(eng.tri_work[r.long()].sum().item())
        if split_plan is not None:
            fam_bytes["tri_split_kernel"] = 8.0 * float(eng.tri_work[split_plan["rows"].long()].sum().item())
        long_rows = tabs.row_nkept > 8192
        fam_bytes["select_cta_kernel"] = 16.0 * float(tabs.row_nkept[long_rows].sum().item())
        fam_bytes["select_warp_kernel"] = 16.0 * float(tabs.row_nkept[~long_rows].sum().item())
        kinds = {kk: (fam_n[kk], fam_ms[kk]) for kk in fam_ms}
        dom = max(kinds, key=lambda kk: kinds[kk][1])
        dom_bytes_step = fam_bytes.get(dom, 0.0)
        dom_ms_step = kinds[dom][1] / args.steps
        achieved = dom_bytes_step / (dom_ms_step * 1e-3) / 1e9 if dom_ms_step > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "item pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 epilogue over i64 fixed-point accumulators", "data": "synthetic",
            "config": {"workload": workload_label(wl, args.method),
                       "pairs_evaluated": P_total, "pairs_kept": P_kept,
                       "l2_policy": "inputs (CSR+CSC+tables) larger than L2; no explicit flush",
                       "parallelism": "item row-blocks x%d, ratings replicated, neighbour records exchanged" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "item pairs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": float(e2e_s) * 1e3,
                    "steps": n_e2e, "statistic": "median of per-step host wall-clock, max over ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC.get(dom),
                         "peak_source": peak_src, "algorithmic_bytes_per_step": dom_bytes_step,
                         "kernel_ms_per_step": dom_ms_step, "launches_per_step": kinds[dom][0] / args.steps,
                         "kernel_share_of_step": dom_ms_step / max(sum(v[1] for v in kinds.values()) / args.steps, 1e-9),
                         "per_kernel_ms_per_step": {kk: v[1] / args.steps for kk, v in kinds.items()},
                         "stage_algorithmic_gbs": stage_gbs, "stage_frac": stage_gbs / peaks["hbm_gbs"],
                         "per_launch_ms_per_step": {kk: v[1] / args.steps for kk, v in prof.items()},
                         "timing": "per-kernel times come from %d extra steps with every launch serialised on one stream "
                                   "(CUDA events around each launch); the timed steps overlap the small launches on a "
                                   "side stream" % args.steps,
                         "note": "achieved = algorithmic bytes of the kernel (8 B per co-rating product for the "
                                 "accumulate kernels, 16 B per neighbour record for the selection kernels) / its "
                                 "CUDA-event time; stage figure = (8*W/2 + 24*nnz + 32*P_kept + 80*k*I) / step time"},
            "pipeline": pipe,
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(wl, args.method)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
