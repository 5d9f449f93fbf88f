"""GPU parity: RecommenderSim `cosine_item` + local sensitivity on the AlterEgo profile,
through the C ABI, against what the UNMODIFIED reference produced (tests/golden/*_recsim.npz, minted by
oracle/make_golden_recsim.py) and against the numpy restatement on a larger profile."""
import numpy as np
import pytest

from tests import parity as PT

pytestmark = pytest.mark.gpu
RTOL = 1e-5            # BASELINE.json north_star: similarity values within 1e-5 relative


def _compare(R, ref_i, ref_j, ref_sim, ref_ls, ref_n=None):
    i, j = R.i.cpu().numpy(), R.j.cpu().numpy()
    assert np.array_equal(i, ref_i) and np.array_equal(j, ref_j), "pair sets differ"
    if ref_n is not None:
        assert np.array_equal(R.n.cpu().numpy(), ref_n)
    np.testing.assert_allclose(R.sim.cpu().numpy(), ref_sim, rtol=RTOL, atol=0)
    ls = R.ls.cpu().numpy()
    assert np.array_equal(np.isnan(ls), np.isnan(ref_ls)), "NaN pattern of the local sensitivity differs"
    ok = ~np.isnan(ref_ls)
    # a sensitivity is a difference of similarities: 1e-5 relative to the similarity scale
    np.testing.assert_allclose(ls[ok], ref_ls[ok], rtol=RTOL, atol=1e-12)
    return float(np.max(np.abs(R.sim.cpu().numpy() - ref_sim) / np.maximum(np.abs(ref_sim), 1e-300))) if len(ref_sim) else 0.0


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_recsim_matches_reference_golden(name):
    from xmap_b200 import recsim
    g = PT.load_golden(name + "_recsim")
    nI = int(max(g["ae_item"].max(), g["rs_i"].max(), g["rs_j"].max())) + 1
    R = recsim.cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    rel = _compare(R, g["rs_i"], g["rs_j"], g["rs_sim"], g["rs_ls"])
    assert int((R.i == R.j).sum()) > 0                             # self pairs: a real next to a synthetic rating
    info = R.info.cpu().numpy()
    np.testing.assert_allclose(info[g["info_item"], 0], g["info_avg"], rtol=1e-13)
    np.testing.assert_allclose(info[g["info_item"], 1], g["info_norm2"], rtol=1e-13)
    assert np.array_equal(info[g["info_item"], 2], g["info_count"].astype(np.float64))
    assert R.n_entries == int(R.n.sum())
    print(name, "recsim pairs", len(g["rs_i"]), "sim max rel", rel)


def test_recsim_vs_restatement_with_duplicates():
    """A larger profile with duplicate (user, item) records, long pair lists (popular items) and half-star means."""
    from oracle import restate as RS
    from xmap_b200 import recsim
    rng = np.random.default_rng(17)
    nU, nI, n = 3000, 500, 60000
    p_item = 1.0 / np.arange(1, nI + 1); p_item /= p_item.sum()
    user = rng.integers(0, nU, n); item = rng.choice(nI, n, p=p_item)
    rating = rng.integers(1, 11, n) * 0.5
    # synthetic records next to real ones: repeat 5 % of the records with another rating (generator.py:156-157)
    dup = rng.choice(n, n // 20, replace=False)
    user = np.concatenate([user, user[dup]]); item = np.concatenate([item, item[dup]])
    rating = np.concatenate([rating, rng.integers(1, 11, len(dup)) * 0.5 + 0.25])
    o = np.argsort(user, kind="stable")
    user, item, rating = user[o], item[o], rating[o]
    P = RS.recommender_cosine_item(user, item, rating, nI, 50)
    R = recsim.cosine_item(user, item, rating, nI, 50)
    _compare(R, P["i"], P["j"], P["sim"], P["ls"], P["n"])
    assert int((P["i"] == P["j"]).sum()) > 100 and int(P["n"].max()) > 200
    np.testing.assert_allclose(R.info.cpu().numpy()[:, 1], P["norm2"], rtol=1e-13)


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_neighbors_and_prediction_match_reference_golden(name):
    """SURVEY.md 8(f) #2-#3 through the C ABI against the UNMODIFIED reference (tests/golden/*_recpred.npz):
    non-private neighbour tables (indices exact, order included), bounded predictions with and without temporal
    decay (exact: they are integers), both MAEs."""
    from xmap_b200 import recsim
    g = PT.load_golden(name + "_recpred")
    nI = int(max(g["ae_item"].max(), g["test_item"].max(), g["nb_idx"].max())) + 1
    nU = int(max(g["ae_user"].max(), g["test_user"].max())) + 1
    K = int(g["mapping_range"])
    R = recsim.cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    nb = recsim.neighbors(R, nI, K)
    ln, idx, sim = nb.len.cpu().numpy(), nb.idx.cpu().numpy(), nb.sim.cpu().numpy()
    has = np.flatnonzero(ln)
    assert np.array_equal(has, g["nb_item"]) and np.array_equal(ln[has], np.diff(g["nb_ptr"]))
    near = 0
    for q, it in enumerate(g["nb_item"]):
        a, b = g["nb_ptr"][q], g["nb_ptr"][q + 1]
        if not np.array_equal(idx[it, :ln[it]], g["nb_idx"][a:b]):
            assert np.allclose(np.abs(sim[it, :ln[it]]), np.abs(g["nb_sim"][a:b]), rtol=1e-9, atol=0), "neighbours of item %d differ" % it
            near += 1
        else:
            np.testing.assert_allclose(sim[it, :ln[it]], g["nb_sim"][a:b], rtol=RTOL)
    assert near <= 1
    p0, p1, m0, m1 = recsim.predict(g["ae_user"], g["ae_item"], g["ae_rating"], g["ae_ts"], nU, R, nb,
                                    g["test_user"], g["test_item"], g["test_rating"], float(g["alpha"]))
    p0, p1 = p0.cpu().numpy(), p1.cpu().numpy()
    # a bounded prediction is int(x + 0.5): it can differ from the reference only if x + 0.5 is an integer to ~1e-12
    assert int((p0 != g["pred_nodecay"]).sum()) <= (1 if near else 0) and int((p1 != g["pred_decay"]).sum()) <= (1 if near else 0)
    assert abs(m0 - float(g["mae_nodecay"])) <= 1e-2 * near + 1e-12 and abs(m1 - float(g["mae_decay"])) <= 1e-2 * near + 1e-12


@pytest.mark.parametrize("name", ["adj_low_overlap", "cos_half_ratings"])
def test_private_neighbors_match_reference_golden(name):
    """The PRIVATE branch of SURVEY.md 8(f) #2 through the C ABI against the UNMODIFIED reference
    (tests/golden/*_recpriv.npz: private_neighbor_selection + noise_perturbation with the reference's np.random draws
    logged): with the same uniforms injected every item draws the same neighbour; the noisy similarity agrees to 1e-9
    (the similarities and sensitivities it is built from come from the device kernels, device exp / log are not glibc's)."""
    from xmap_b200 import recsim
    g0, g = PT.load_golden(name), PT.load_golden(name + "_recpriv")
    nI = len(g0["iids"])
    R = recsim.cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    nb = recsim.private_neighbors(R, nI, int(g["mapping_range"]), float(g["epsilon"]), float(g["rpo"]), g["u_pick"], g["u_noise"])
    ln, idx, sim = nb.len.cpu().numpy(), nb.idx.cpu().numpy()[:, 0], nb.sim.cpu().numpy()[:, 0]
    has = np.flatnonzero(ln)
    assert np.array_equal(has, g["item"]) and (ln[has] == 1).all()
    assert np.array_equal(idx[has], g["chosen"])
    np.testing.assert_allclose(sim[has], g["noisy_sim"], rtol=1e-9, atol=1e-12)
    # Philox draws: reproducible for a seed, different for another one
    a = recsim.private_neighbors(R, nI, 10, 0.6, 0.1, seed=11)
    b = recsim.private_neighbors(R, nI, 10, 0.6, 0.1, seed=11)
    c = recsim.private_neighbors(R, nI, 10, 0.6, 0.1, seed=12)
    assert bool((a.idx == b.idx).all()) and bool((a.sim == b.sim).all()) and not bool((a.idx == c.idx).all())


def test_private_neighbors_vs_restatement_long_lists():
    """Long neighbour lists (up to 3 000: numpy's pairwise np.sum recurses above 128 elements, count > k exercises the w
    term), negative and tied similarities, tiny sensitivities: the kernel against oracle/restate.py with random uniforms."""
    import torch
    from oracle import restate as RS
    from xmap_b200 import recsim
    rng = np.random.default_rng(8)
    nI = 400
    lens = rng.integers(0, 60, nI); lens[:12] = [1, 2, 7, 8, 9, 10, 11, 128, 129, 1000, 2047, 3000]
    i = np.repeat(np.arange(nI), lens)
    j = np.concatenate([np.sort(rng.choice(4000, n, replace=False)) for n in lens]).astype(np.int64)
    sim = np.round(rng.normal(0, 0.2, len(i)), 3)                     # rounded: exact ties in |sim|
    ls = np.abs(rng.normal(0, 0.05, len(i))) + 1e-3
    P = dict(i=i.astype(np.int64), j=j, sim=sim, ls=ls)
    n_have = int((lens > 0).sum())
    up, un = rng.random(n_have), rng.random(n_have)
    it0, ch0, out0 = RS.recommender_private_neighbors(P, nI, 10, 0.6, 0.1, up, un)
    dev = torch.device("cuda")
    R = recsim.RecSim(torch.as_tensor(i, dtype=torch.int32, device=dev), torch.as_tensor(j, dtype=torch.int32, device=dev),
                      torch.ones(len(i), dtype=torch.int64, device=dev), torch.as_tensor(sim, device=dev),
                      torch.as_tensor(ls, device=dev), torch.zeros((nI, 3), dtype=torch.float64, device=dev), len(i))
    nb = recsim.private_neighbors(R, nI, 10, 0.6, 0.1, up, un)
    ln, idx, out = nb.len.cpu().numpy(), nb.idx.cpu().numpy()[:, 0], nb.sim.cpu().numpy()[:, 0]
    has = np.flatnonzero(ln)
    assert np.array_equal(has, it0)
    assert int((idx[has] != ch0).sum()) == 0
    np.testing.assert_allclose(out[has], out0, rtol=1e-12, atol=1e-15)
