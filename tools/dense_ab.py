"""A/B for north_star (1) "dense blocks of heavy items on int8 tensor cores": the top-H most popular items of a
workload, method "cosine" (every numerator is an integer matrix product: M^T M co-counts, A^T A and A^T M for the
mutuality, R^T R for the rating dot products; baselinerSim.py:97-142).

  dense  : the four H x H products over the compacted user set as int8 GEMMs (torch._int_mm -> cuBLASLt int8, the
           library's tcgen05 path on sm_100a), inputs already resident as dense int8 matrices
  sparse : the product path -- the triangular rows of the same H items (tri_* kernels, shared-memory tables)

Both produce the co-count / mutuality / inner product of every heavy x heavy pair; the counts are cross-checked.
Also prints the measured int8 GEMM peak (8192^3).  One JSON line on stdout.

  python tools/dense_ab.py [workload] [H]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cuda_ms(fn, reps=5):
    import torch
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)


def main():
    import numpy as np
    import torch
    import bench
    from xmap_b200 import engine as E
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    wl = bench.make_workload(name)
    dev = torch.device("cuda")
    meta = E.to_device_meta(wl["meta"], dev)
    lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
    eng = E.SimEngine(lay, meta, "cosine", 50, wl["k"])
    # int8 GEMM peak (library)
    n = 8192
    a8 = torch.randint(-4, 5, (n, n), dtype=torch.int8, device=dev)
    b8 = torch.randint(-4, 5, (n, n), dtype=torch.int8, device=dev)
    ms_peak = cuda_ms(lambda: torch._int_mm(a8, b8))
    int8_tops = 2.0 * n ** 3 / (ms_peak * 1e-3) / 1e12
    del a8, b8
    # ---- the heavy block ---------------------------------------------------------------------------------
    count = lay.item_stats[:, 3]
    heavy = torch.argsort(eng.ord, descending=True)[:H]               # the H most popular items
    hpos = torch.full((wl["n_items"],), -1, dtype=torch.int64, device=dev); hpos[heavy] = torch.arange(H, device=dev)
    user = torch.as_tensor(wl["user"], device=dev).long(); item = torch.as_tensor(wl["item"], device=dev).long()
    rating = torch.as_tensor(wl["rating"], device=dev)
    m = hpos[item] >= 0
    hu, hi, hr = user[m], hpos[item[m]], rating[m]
    users = torch.unique(hu)
    upos = torch.full((wl["n_users"],), -1, dtype=torch.int64, device=dev); upos[users] = torch.arange(users.numel(), device=dev)
    Ub = int(users.numel())
    Ub_pad = (Ub + 63) // 64 * 64
    avg = lay.item_stats[:, 0][heavy]
    M = torch.zeros((Ub_pad, H), dtype=torch.int8, device=dev)
    A = torch.zeros_like(M); R = torch.zeros_like(M)
    M[upos[hu], hi] = 1
    R[upos[hu], hi] = hr.to(torch.int8)
    A[upos[hu], hi] = (hr.double() >= avg[hi]).to(torch.int8)
    Mt, At, Rt = M.t().contiguous(), A.t().contiguous(), R.t().contiguous()
    out = {}

    def dense():
        out["n"] = torch._int_mm(Mt, M); out["aa"] = torch._int_mm(At, A)
        out["am"] = torch._int_mm(At, M); out["rr"] = torch._int_mm(Rt, R)
    ms_dense = cuda_ms(dense)
    ms_dense_one = cuda_ms(lambda: torch._int_mm(Mt, M))
    dense_ops = 4 * 2.0 * H * H * Ub_pad
    # ---- the sparse product path on the same rows -----------------------------------------------------------
    rows = torch.sort(heavy).values.to(torch.int32)
    eng.reset(); eng.accumulate(rows); torch.cuda.synchronize()          # plan + warm-up

    def sparse():
        eng.reset(); eng._accumulate(rows)
    ms_sparse = cuda_ms(sparse)
    eng._check_error()
    prods = int(eng.tri_work[rows.long()].sum().item())
    # cross-check: co-counts and mutuality of the kept heavy x heavy pairs
    p = eng.emit_pairs(rows)
    hi_i, hj = hpos[p["i"]], hpos[p["j"]]
    ok = hj >= 0
    n_d = out["n"][hi_i[ok], hj[ok]]
    agree = 2 * out["aa"][hi_i[ok], hj[ok]] - out["am"][hi_i[ok], hj[ok]] - out["am"][hj[ok], hi_i[ok]] + n_d
    same_n = bool((n_d == p["n"][ok].to(n_d.dtype)).all()); same_m = bool((agree == p["mutu"][ok].to(agree.dtype)).all())
    dens = float((out["n"] > 0).double().mean())
    line = {"workload": bench.workload_label(wl, "cosine"), "H": H, "users_in_block": Ub, "ratings_in_block": int(m.sum().item()),
            "block_density_of_R": float(m.sum().item()) / (Ub * H), "corated_pair_density": dens,
            "int8_gemm_peak_tops_8192": int8_tops,
            "dense_ms_4_products": ms_dense, "dense_ms_1_product": ms_dense_one, "dense_int8_ops": dense_ops,
            "dense_achieved_tops": dense_ops / (ms_dense * 1e-3) / 1e12,
            "sparse_ms_same_rows": ms_sparse, "sparse_products": prods, "sparse_products_per_s": prods / (ms_sparse * 1e-3),
            "counts_equal": same_n, "mutuality_equal": same_m,
            "dense_over_sparse": ms_dense / ms_sparse,
            "note": "dense = 4 cuBLASLt int8 GEMMs (M^T M, A^T A, A^T M, R^T R) on resident dense int8 matrices, excluding the "
                    "densification and the epilogue; sparse = the tri_* kernels on the triangular rows of the same items, "
                    "including their fused epilogue (similarity, filter, records)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
