"""CPU: the oracle restatement (oracle/restate.py) is pinned against the golden vectors that
the UNMODIFIED reference produced (oracle/make_golden.py), stage by stage."""
import os

import numpy as np
import pytest

from oracle import restate as RS
from tests import parity as PT


def _case(name):
    g = PT.load_golden(name)
    meta = PT.golden_meta(g)
    nU, nI = len(g["uids"]), len(g["iids"])
    P = RS.sim_pairs(g["user"].astype(np.int64), g["item"].astype(np.int64), g["rating"], nU, nI,
                     meta["prefix_code"], str(g["method"]), int(g["num_atleast"]))
    return g, meta, nU, nI, P


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_restatement_similarity_vs_reference(name):
    g, meta, nU, nI, P = _case(name)
    assert np.array_equal(P["i"], g["sim_i"]) and np.array_equal(P["j"], g["sim_j"])
    assert np.array_equal(P["mutu"], g["sim_mutu"]) and np.array_equal(P["frac"], g["sim_frac"])
    assert np.array_equal(P["label"], g["sim_label"])
    # bit-equal wherever numpy sums sequentially (n < 8), 1e-12 elsewhere
    small = P["n"] < 8
    assert np.array_equal(P["sim"][small], g["sim_val"][small])
    np.testing.assert_allclose(P["sim"], g["sim_val"], rtol=1e-10)
    st = P["stats"]
    assert np.array_equal(st["mu"], g["user_avg"])
    assert np.array_equal(np.stack([st["avg"], st["norm2"], st["adj_norm2"], st["count"]], 1), g["item_info"])


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_restatement_selection_extension_generation_vs_reference(name):
    g, meta, nU, nI, P = _case(name)
    k = int(g["k"])
    knn = RS.select_knn(P, nI, k, meta["dom_code"], meta["contains"])
    assert np.array_equal(knn["bb"], g["bb"]) and np.array_equal(knn["valid_nb"], g["valid_nb"])
    lists = PT.restate_lists(knn, P)
    for nm in ("BB_BB", "BB_NB", "NB_BB", "NB_NN"):
        assert np.array_equal(lists[nm][0], g[nm + "_ptr"]) and np.array_equal(lists[nm][1], g[nm + "_nbr"])
    X = RS.xsim_extend(P, knn, nI, meta["has_S"], meta["has_T"])
    PT.compare_xsim(X["start"], X["end"], X["xsim"], g["xs_start"], g["xs_end"], g["xs_val"], rtol=1e-9)
    if "priv_rows" not in g:
        return
    rows, cands = RS.candidates(X["start"], X["end"], X["xsim"], 10)
    ch, _ = RS.choose(rows, cands, "argmax")
    assert np.array_equal(rows, g["priv_rows"]) and np.array_equal(ch, g["priv_chosen"])
    mp = RS.invert_mapping(rows, ch, nI)
    ae = RS.build_alterego(g["user"], g["item"], g["rating"], g["ts"], mp, meta["has_T"])
    o = np.lexsort((ae["ts"], ae["rating"], ae["item"], ae["user"]))
    assert np.array_equal(ae["user"][o], g["priv_ae_user"]) and np.array_equal(ae["item"][o], g["priv_ae_item"])
    assert np.array_equal(ae["rating"][o], g["priv_ae_rating"]) and np.array_equal(ae["ts"][o], g["priv_ae_ts"])
    rows4, cands4 = RS.candidates(X["start"], X["end"], X["xsim"], 4)
    ok = np.array([len(c[0]) >= 2 for c in cands4])
    u = np.zeros(len(rows4)); u[ok] = g["nonpriv_uniforms"]
    chn, single = RS.choose(rows4, cands4, "nonprivate", uniforms=u)
    assert np.array_equal(rows4[ok], g["nonpriv_rows"]) and np.array_equal(chn[ok], g["nonpriv_chosen"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/code/xmap"), reason="reference tree not mounted")
def test_harness_runs_the_reference_and_agrees_with_golden():
    """Where /root/reference exists, re-run the reference itself on the smallest case."""
    import subprocess, sys
    code = ("import numpy as np;from oracle import make_golden as M;"
            "o=M.build_case('adj_all_bridge');g=np.load('tests/golden/adj_all_bridge.npz');"
            "assert all(np.array_equal(o[k],g[k]) for k in g.files if k!='ref_seconds');print('ok')")
    r = subprocess.run([sys.executable, "-c", code], cwd=PT.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_restatement_recommender_sim_vs_reference(name):
    """Next row of the scope table: RecommenderSim `cosine_item` + local sensitivity on the
    AlterEgo profile -- the restatement against what the unmodified reference produced
    (oracle/make_golden_recsim.py).  Pins duplicate (user, item) records, self pairs and the absence of a filter."""
    g = PT.load_golden(name + "_recsim")
    nI = int(max(g["ae_item"].max(), g["rs_i"].max(), g["rs_j"].max())) + 1
    P = RS.recommender_cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    assert np.array_equal(P["i"], g["rs_i"]) and np.array_equal(P["j"], g["rs_j"])
    assert int((P["i"] == P["j"]).sum()) > 0                       # a real next to a synthetic rating of one item
    np.testing.assert_allclose(P["sim"], g["rs_sim"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(P["ls"], g["rs_ls"], rtol=1e-9, atol=1e-13, equal_nan=True)
    np.testing.assert_allclose(P["norm2"][g["info_item"]], g["info_norm2"], rtol=1e-14)
    cnt = np.bincount(g["ae_item"], minlength=nI)
    assert np.array_equal(cnt[g["info_item"]], g["info_count"])


def test_sim_rows_restriction_equals_sim_pairs():
    """restate.sim_rows (the sampled-row oracle of the full-size GPU test) is bit-equal to sim_pairs on its rows."""
    from tests import parity as PT
    case = PT.synth_case(1500, 400, 20000, 0.05, seed=5)
    args = (case["user"], case["item"], case["rating"], case["n_users"], case["n_items"], case["meta"]["prefix_code"])
    for method in ("adjust_cosine", "cosine"):
        P = RS.sim_pairs(*args, method, 50)
        rows = np.array([0, 3, 17, 100, 250, 398, 399, 399])
        Q = RS.sim_rows(*args, rows, method, 50)
        m = np.isin(P["i"], rows)
        assert m.sum() > 100
        for key in ("i", "j", "sim", "mutu", "n", "frac", "label"):
            assert np.array_equal(P[key][m], Q[key]), key


def test_restatement_vs_reference_mid_size():
    """The mid-size golden (90 K ratings, 366 K kept pairs; the reference took 17 s for the similarity stage):
    similarity, mutuality, labels, BB set and all four neighbour tables of the restatement against the
    unmodified reference.  Lists of ~1e3 neighbours and rows with thousands of co-rating products."""
    g, meta, nU, nI, P = _case("adj_mid")
    assert len(g["user"]) > 85000 and len(g["sim_i"]) > 300000
    assert np.array_equal(P["i"], g["sim_i"]) and np.array_equal(P["j"], g["sim_j"])
    assert np.array_equal(P["mutu"], g["sim_mutu"]) and np.array_equal(P["label"], g["sim_label"])
    assert np.array_equal(P["frac"].astype(np.float32), g["sim_frac"])      # stored as its float32 image
    np.testing.assert_allclose(P["sim"], g["sim_val"], rtol=1e-10)
    assert np.array_equal(P["stats"]["mu"], g["user_avg"])
    knn = RS.select_knn(P, nI, int(g["k"]), meta["dom_code"], meta["contains"])
    assert np.array_equal(knn["bb"], g["bb"]) and np.array_equal(knn["valid_nb"], g["valid_nb"])
    lists = PT.restate_lists(knn, P)
    n_diff = 0
    for nm in ("BB_BB", "BB_NB", "NB_BB", "NB_NN"):
        assert np.array_equal(lists[nm][0], g[nm + "_ptr"])
        n_diff += int((lists[nm][1] != g[nm + "_nbr"]).sum())
    assert n_diff == 0, "%d neighbour entries differ from the reference" % n_diff


@pytest.mark.parametrize("name", PT.GOLDEN_CASES)
def test_restatement_neighbors_and_prediction_vs_reference(name):
    """non-private neighbour selection and item-based prediction + MAE on the AlterEgo
    profile -- the restatement against the unmodified reference (oracle/make_golden_recpred.py)."""
    g = PT.load_golden(name + "_recpred")
    nI = int(max(g["ae_item"].max(), g["test_item"].max(), g["nb_idx"].max())) + 1
    P = RS.recommender_cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    ptr, idx, val = RS.recommender_neighbors(P, nI, int(g["mapping_range"]))
    has = np.flatnonzero(np.diff(ptr))
    assert np.array_equal(has, g["nb_item"])
    assert np.array_equal(np.diff(ptr)[has], np.diff(g["nb_ptr"]))
    assert np.array_equal(idx, g["nb_idx"])
    np.testing.assert_allclose(val, g["nb_sim"], rtol=1e-12)
    cnt = np.bincount(g["ae_item"], minlength=nI)
    avg = np.bincount(g["ae_item"], weights=g["ae_rating"], minlength=nI) / np.maximum(cnt, 1)
    p0, p1, m0, m1 = RS.recommender_predict(g["ae_user"], g["ae_item"], g["ae_rating"], g["ae_ts"], ptr, idx, val, avg,
                                            g["test_user"], g["test_item"], g["test_rating"], float(g["alpha"]))
    assert np.array_equal(p0, g["pred_nodecay"]) and np.array_equal(p1, g["pred_decay"])
    assert abs(m0 - float(g["mae_nodecay"])) < 1e-12 and abs(m1 - float(g["mae_decay"])) < 1e-12


@pytest.mark.parametrize("name", ["adj_low_overlap", "cos_half_ratings"])
def test_restatement_private_neighbors_vs_reference(name):
    """The PRIVATE branch of SURVEY.md 8(f) #2: private_neighbor_selection + noise_perturbation run by the unmodified
    reference with its np.random draws logged (tests/golden/*_recpriv.npz, oracle/make_golden_recpriv.py); the restatement
    fed the same uniforms picks the same neighbour for every item and returns the same noisy similarity."""
    g0, g = PT.load_golden(name), PT.load_golden(name + "_recpriv")
    nI = len(g0["iids"])
    P = RS.recommender_cosine_item(g["ae_user"], g["ae_item"], g["ae_rating"], nI, int(g["num_atleast"]))
    it, ch, out = RS.recommender_private_neighbors(P, nI, int(g["mapping_range"]), float(g["epsilon"]), float(g["rpo"]),
                                                   g["u_pick"], g["u_noise"])
    assert len(it) > 100 and np.array_equal(it, g["item"]) and np.array_equal(ch, g["chosen"])
    np.testing.assert_allclose(out, g["noisy_sim"], rtol=1e-13, atol=0)
    # the draw is not the arg-max: the mechanism really samples
    nb = RS.recommender_neighbors(P, nI, 1)
    assert 0 < int((ch != nb[1]).sum())
