"""GPU parity: X-SIM extension + AlterEgo generation through the C ABI vs the oracle."""
import numpy as np
import pytest

from tests import parity as PT

pytestmark = pytest.mark.gpu


def _sim(g):
    meta = PT.golden_meta(g)
    nU, nI = len(g["uids"]), len(g["iids"])
    lay, eng, tabs, _ = PT.run_gpu_sim(g["user"], g["item"], g["rating"], nU, nI, meta,
                                       str(g["method"]), int(g["num_atleast"]), int(g["k"]), emit=False)
    return meta, lay, tabs


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_xsim_matches_reference_golden(name):
    g = PT.load_golden(name)
    meta, lay, tabs = _sim(g)
    plan, xe, res, (s, e, v) = PT.run_gpu_extend(tabs, lay, meta)
    PT.compare_xsim(s, e, v, g["xs_start"], g["xs_end"], g["xs_val"])
    assert int(res.count.sum()) == len(g["xs_val"])


@pytest.mark.parametrize("name", ["adj_low_overlap", "cos_half_ratings"])
def test_generation_matches_reference_golden(name):
    import torch
    from xmap_b200 import generate as G
    g = PT.load_golden(name)
    meta, lay, tabs = _sim(g)
    plan, xe, res, _ = PT.run_gpu_extend(tabs, lay, meta)
    nI = len(g["iids"])
    starts = res.start_item.cpu().numpy()
    # private mapping as shipped (argmax): identical (target, source) pairs
    ch = G.choose_mapping(res, "argmax", sim_method=str(g["method"]))
    assert np.array_equal(starts, g["priv_rows"])
    assert np.array_equal(ch.cpu().numpy(), g["priv_chosen"])
    mp = G.invert_mapping(res.start_item, ch, nI)
    ou, oi, orr, ot = G.build_alterego(lay, g["ts"], mp)
    hasT = torch.as_tensor(meta["has_T"])
    keep = hasT[torch.as_tensor(g["item"]).long()].numpy()
    U = np.concatenate([g["user"][keep], ou.cpu().numpy()]); I = np.concatenate([g["item"][keep], oi.cpu().numpy()])
    R = np.concatenate([g["rating"][keep], orr.cpu().numpy()]); T = np.concatenate([g["ts"][keep], ot.cpu().numpy()])
    o = np.lexsort((T, R, I, U))
    assert np.array_equal(U[o], g["priv_ae_user"]) and np.array_equal(I[o], g["priv_ae_item"])
    assert np.array_equal(R[o], g["priv_ae_rating"]) and np.array_equal(T[o], g["priv_ae_ts"])
    # non-private with the reference's own np.random draws replayed as uniforms
    ok = (res.count.cpu().numpy() >= 2)
    u = np.zeros(len(starts)); u[ok] = g["nonpriv_uniforms"]
    chn = G.choose_mapping(res, "nonprivate", uniforms=u).cpu().numpy()
    assert np.array_equal(starts[ok], g["nonpriv_rows"])
    assert np.array_equal(chn[ok], g["nonpriv_chosen"])


def test_xsim_vs_restatement_and_pass_splits():
    """A larger case against oracle/restate.py in both kernels and many pass plans: tiny tables, tiny units, a wildly
    optimistic pass estimate (device-side splits), tables in global memory, with and without the fused bridge lists.
    Every plan gives the same keys, distinct-end counts and path counts; the sums are formed in an order that depends
    on the plan (runs of equal ends inside a 32-path chunk are summed by a shuffle tree), so values agree to <= 1e-10
    relative between plans -- and bit for bit from run to run of the same plan."""
    from oracle import restate as RS
    case = PT.synth_case(4000, 900, 60000, 0.04, seed=21)
    out = PT.check_sim_against_restatement(case, "adjust_cosine", 50, 5)
    X = RS.xsim_extend(out["P"], out["knn"], case["n_items"], case["meta"]["has_S"], case["meta"]["has_T"])
    plan, xe, res, (s, e, v) = PT.run_gpu_extend(out["tabs"], out["lay"], case["meta"])          # default: hybrid, fused lists
    assert xe.mode == "hybrid" and xe.fused_entries > 0
    assert plan.n_src == X["n_src"] and plan.n_joint == X["n_joint"]
    PT.compare_xsim(s, e, v, X["start"], X["end"], X["xsim"])
    assert int(res.combos.sum()) == X["combos"]
    assert np.array_equal(np.bincount(np.searchsorted(res.start_item.cpu().numpy(), X["start"]),
                                      minlength=len(res.count)), res.count.cpu().numpy())
    W = dict(mode="warp")
    variants = dict(warp=dict(W), small_tables=dict(W, cells_lg=6, max_passes=10 ** 9), many_units=dict(W, unit_work=500),
                    global_tables=dict(W, cells_lg=6, max_passes=1), big_tables=dict(W, cells_lg=11, warps=4),
                    device_splits=dict(W, cells_lg=6, rho=1e9, max_passes=10 ** 9),
                    global_splits=dict(W, cells_lg=6, rho=1e9, max_passes=2),
                    both=dict(W, cells_lg=7, unit_work=300, rho=3.0, warps=5),
                    cta=dict(mode="cta"), cta_small=dict(mode="cta", cells_lg=9, unit_work=400),
                    cta_splits=dict(mode="cta", cells_lg=9, rho=1e9),
                    cta_unfused=dict(mode="cta", fuse=False), warp_unfused=dict(W, fuse=False),
                    hybrid_some_hot=dict(hot_paths=3000.0), hybrid_all_hot=dict(hot_paths=0.0),
                    hybrid_small=dict(hot_paths=800.0, cells_lg=6, unit_work=500, max_passes=10 ** 9),
                    uniform_cuts=dict(balance="uniform", unit_work=500), heat_cuts=dict(W, balance="heat", cells_lg=6, unit_work=300))
    for name, kw in variants.items():
        plan2, xe2, res2, (s2, e2, v2) = PT.run_gpu_extend(out["tabs"], out["lay"], case["meta"], **kw)
        if name == "small_tables":
            assert int(xe2.T.max()) > 1 and xe2.gws is None
        if name == "global_tables":
            assert xe2.gws is not None and int(xe2.T.max()) <= 2
        if name == "many_units":
            assert xe2.n_units > plan2.start_item.numel()
        if name.endswith("unfused"):
            assert xe2.fused_entries == 0
        if name.startswith("hybrid"):
            assert xe2.hot_order.numel() > 0 and (name == "hybrid_all_hot") == (xe2.cold_order.numel() == 0)
        assert np.array_equal(s, s2) and np.array_equal(e, e2), name
        np.testing.assert_allclose(v2, v, rtol=1e-10, atol=0, err_msg=name)      # observed: <= 1.4e-12 (a cancelling sum)
        for f in ("count", "combos", "top_len"):
            assert np.array_equal(getattr(res, f).cpu().numpy(), getattr(res2, f).cpu().numpy()), (name, f)
        _, _, res3, (s3, e3, v3) = PT.run_gpu_extend(out["tabs"], out["lay"], case["meta"], **kw)
        assert np.array_equal(v2, v3) and np.array_equal(res2.top_end.cpu().numpy(), res3.top_end.cpu().numpy()), name
        assert np.array_equal(res2.top_xsim.cpu().numpy(), res3.top_xsim.cpu().numpy()), name
    # top-m rows = first m of the full rows ordered by |xsim| desc, ties to smaller end
    rows, cands = RS.candidates(X["start"], X["end"], X["xsim"], 10)
    te, tl = res.top_end.cpu().numpy(), res.top_len.cpu().numpy()
    assert np.array_equal(rows, res.start_item.cpu().numpy())
    bad = sum(0 if np.array_equal(te[r, :tl[r]], cands[r][0]) else 1 for r in range(len(rows)))
    assert bad <= 0.002 * len(rows) + 1, "%d top-m rows differ (only near-ties in xsim may)" % bad
    # ... and exactly the top-m of the kernel's own full rows
    rows2, cands2 = RS.candidates(s, e, v, 10)
    tx = res.top_xsim.cpu().numpy()
    for r in range(len(rows2)):
        assert np.array_equal(te[r, :tl[r]], cands2[r][0]) and np.array_equal(tx[r, :tl[r]], cands2[r][1])


def test_multi_domain_shape():
    """BASELINE.json configs[3] in miniature: three domains ("S:1:", "S:2:", "T:"), one source -> target
    pipeline per source domain as multidomain_demo.py:101-128 runs them; similarity, neighbour lists,
    X-SIM and the argmax mapping of each against the restatement."""
    from oracle import restate as RS
    from xmap_b200 import generate as G
    cases = PT.multi_domain_cases(4500, 500, 60000, 0.08, seed=31)
    assert [lab for lab, _ in cases] == ["S:1:", "S:2:"]
    for lab, case in cases:
        assert any(lab in str(i) for i in case["iids"]) and any(str(i).endswith("T:") for i in case["iids"])
        out = PT.check_sim_against_restatement(case, "adjust_cosine", 50, 6)
        Xr = RS.xsim_extend(out["P"], out["knn"], case["n_items"], case["meta"]["has_S"], case["meta"]["has_T"])
        plan, xe, res, (s, e, v) = PT.run_gpu_extend(out["tabs"], out["lay"], case["meta"])
        assert len(s) > 0 and plan.n_src == Xr["n_src"]
        PT.compare_xsim(s, e, v, Xr["start"], Xr["end"], Xr["xsim"])
        rows, cands = RS.candidates(Xr["start"], Xr["end"], Xr["xsim"], 10)
        ch = G.choose_mapping(res, "argmax").cpu().numpy()
        bad = sum(0 if ch[r] == cands[r][0][0] else 1 for r in range(len(rows)))
        assert bad <= 0.002 * len(rows) + 1, "%d argmax mappings differ (only near-ties in xsim may)" % bad


def test_exponential_mechanism_injected_uniforms_and_philox():
    """Injected uniforms reproduce oracle/restate.choose exactly; the Philox stream matches its
    numpy port; sampled frequencies pass a chi-square test against exp(eps*xsim/(2*k*GS))."""
    import torch
    from oracle import restate as RS
    from xmap_b200 import generate as G
    from xmap_b200.extend import XsimResult
    rng = np.random.default_rng(4)
    n, m = 20000, 10
    xs = np.sort(rng.uniform(-1, 1, size=(1, m)))[:, ::-1].repeat(n, 0)
    xs = xs[:, np.argsort(-np.abs(xs[0]))].copy()
    ends = np.tile(np.arange(m, dtype=np.int32), (n, 1))
    res = XsimResult(torch.arange(n, dtype=torch.int32, device="cuda"),
                     torch.full((n,), m, dtype=torch.int32, device="cuda"),
                     torch.zeros(n, dtype=torch.int64, device="cuda"),
                     torch.as_tensor(ends, device="cuda"), torch.as_tensor(xs, device="cuda"),
                     torch.full((n,), m, dtype=torch.int32, device="cuda"), 0)
    u = rng.random(n)
    got = G.choose_mapping(res, "exp_mech", epsilon=0.6, mapping_range=1, sim_method="adjust_cosine",
                           uniforms=u).cpu().numpy()
    want, _ = RS.choose(np.arange(n), [(ends[r], xs[r]) for r in range(n)], "exp_mech", uniforms=u,
                        epsilon=0.6, mapping_range=1, gs=2)
    assert np.array_equal(got, want)
    # Philox path == injecting the numpy port's uniforms
    seed = 0x1234567890ABCDEF
    a = G.choose_mapping(res, "exp_mech", epsilon=3.0, seed=seed).cpu().numpy()
    b = G.choose_mapping(res, "exp_mech", epsilon=3.0, uniforms=PT.philox_uniforms(seed, n)).cpu().numpy()
    assert np.array_equal(a, b)
    # non-private mapping with another `topn` than the reference's default (generator.py:100: any topn is legal there)
    tl = rng.integers(1, m + 1, size=n).astype(np.int32)
    res_v = XsimResult(res.start_item, res.count, res.combos, res.top_end, res.top_xsim, torch.as_tensor(tl, device="cuda"), 0)
    for topn in (2, 4, 7):
        got = G.choose_mapping(res_v, "nonprivate", uniforms=u, topn=topn).cpu().numpy()
        want, _ = RS.choose(np.arange(n), [(ends[r, :min(tl[r], topn)], xs[r, :min(tl[r], topn)]) for r in range(n)],
                            "nonprivate", uniforms=u)
        assert np.array_equal(got, want), topn
    # chi-square: 9 dof, 99.9% quantile = 27.88
    w = np.exp(3.0 * xs[0] / (2 * 1 * 2)); p = w / w.sum()
    obs = np.bincount(a, minlength=m)
    chi2 = float(((obs - n * p) ** 2 / (n * p)).sum())
    assert chi2 < 27.88, chi2
