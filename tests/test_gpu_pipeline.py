"""GPU parity at the reference's own boundary: the five pipeline functions with the
reference's record shapes (string ids, datetimes), compared with what the unmodified
reference produced on the same records (tests/golden)."""
import calendar
from datetime import datetime

import numpy as np
import pytest

from tests import parity as PT

pytestmark = pytest.mark.gpu


def _train_records(g):
    recs = {}
    for u, i, r, t in zip(g["user"], g["item"], g["rating"], g["ts"]):
        recs.setdefault(str(g["uids"][u]), []).append((str(g["iids"][i]), float(r), datetime.utcfromtimestamp(int(t))))
    return list(recs.items())


@pytest.mark.parametrize("name", ["adj_low_overlap", "cos_half_ratings"])
def test_pipelines_match_reference(name):
    from xmap_b200.rdd import LocalRDD
    from xmap_b200.core import (BaselinerSim, ExtendSim, Generator, baseliner_calculate_sim_pipeline,
                                extender_pipeline, generator_pipeline)
    g = PT.load_golden(name)
    iids = [str(s) for s in g["iids"]]
    uids = [str(s) for s in g["uids"]]
    trainRDD = LocalRDD(_train_records(g))
    method = str(g["method"])
    sim_tool = BaselinerSim(method, int(g["num_atleast"]))
    simRDD = baseliner_calculate_sim_pipeline(None, sim_tool, trainRDD)
    ext_tool = ExtendSim(int(g["k"]))
    xsRDD = extender_pipeline(None, None, sim_tool, ext_tool, simRDD)

    # similarity records, exactly the reference's shape: ((iid1, iid2), (sim, mutu, frac, label))
    got = sorted(simRDD.collect(), key=lambda x: x[0])
    want_keys = [(iids[a], iids[b]) for a, b in zip(g["sim_i"], g["sim_j"])]
    order = sorted(range(len(want_keys)), key=lambda q: want_keys[q])
    assert [k for k, _ in got] == [want_keys[q] for q in order]
    gv = np.array([[v[0], v[1], v[2], v[3]] for _, v in got])
    o = np.array(order)
    assert np.array_equal(gv[:, 1], g["sim_mutu"][o]) and np.array_equal(gv[:, 2], g["sim_frac"][o])
    assert np.array_equal(gv[:, 3], g["sim_label"][o])
    np.testing.assert_allclose(gv[:, 0], g["sim_val"][o], rtol=PT.SIM_RTOL)

    # X-SIM rows: (iid_T, [(iid_S, xsim)*])
    rows = xsRDD.collect()
    flat = sorted((t, s, v) for t, lst in rows for s, v in lst)
    want = sorted((iids[a], iids[b], v) for a, b, v in zip(g["xs_start"], g["xs_end"], g["xs_val"]))
    assert [(a, b) for a, b, _ in flat] == [(a, b) for a, b, _ in want]
    np.testing.assert_allclose([v for _, _, v in flat], [v for _, _, v in want], rtol=PT.SIM_RTOL)
    assert all("T:" in t for t, _ in rows) and all("S:" in s for _, lst in rows for s, _ in lst)

    # AlterEgo profile, private flag on (= argmax under Python 3): identical records
    gen = Generator(1, 0.6, method, 0.1)
    alter = generator_pipeline(gen, trainRDD, xsRDD, True).collect()
    mine = sorted((u, i, float(r), calendar.timegm(t.timetuple())) for u, i, r, t in alter)
    ref = sorted((uids[u], iids[i], float(r), int(t)) for u, i, r, t in
                 zip(g["priv_ae_user"], g["priv_ae_item"], g["priv_ae_rating"], g["priv_ae_ts"]))
    assert mine == ref

    # non-private, the reference's np.random draws replayed as injected uniforms
    cnt = xsRDD.handle.result.count.cpu().numpy()
    u = np.zeros(len(cnt)); u[cnt >= 2] = g["nonpriv_uniforms"]
    gen2 = Generator(1, 0.6, method, 0.1, uniforms=u)
    pairs = gen2.cross_nonprivate_mapping(xsRDD).collect()
    st = xsRDD.handle.result.start_item.cpu().numpy()
    okrows = {iids[t] for t, c in zip(st, cnt) if c >= 2}
    got_pairs = [(t, s) for t, s in pairs if t in okrows]
    assert got_pairs == [(iids[t], iids[s]) for t, s in zip(g["nonpriv_rows"], g["nonpriv_chosen"])]
    assert gen2.single_candidate_rows == int((cnt == 1).sum())


def test_generator_accepts_foreign_xsim_rows_and_mapping_dict():
    """generator_pipeline fed plain (target, [(source, xsim)]) records (not our lazy RDD),
    and build_alterEgo fed a mapping dict, as the reference API allows."""
    from xmap_b200.rdd import LocalRDD
    from xmap_b200.core import Generator, generator_pipeline
    from xmap_b200.utils.assist import map_to_dict
    g = PT.load_golden("adj_low_overlap")
    iids = [str(s) for s in g["iids"]]
    trainRDD = LocalRDD(_train_records(g))
    rows = {}
    for a, b, v in zip(g["xs_start"], g["xs_end"], g["xs_val"]):
        rows.setdefault(iids[a], []).append((iids[b], float(v)))
    xs = LocalRDD(rows.items())
    gen = Generator(1, 0.6, "adjust_cosine", 0.1)
    alter = generator_pipeline(gen, trainRDD, xs, True).collect()
    assert len(alter) == len(g["priv_ae_user"])
    mapped = gen.cross_private_mapping(xs, session=trainRDD._xmap_session)
    alter2 = gen.build_alterEgo(trainRDD, map_to_dict(mapped)).collect()
    assert sorted(map(str, alter)) == sorted(map(str, alter2))


def test_facade_methods_match_reference_golden():
    """The reference-named stage methods a caller can use directly (baselinerSim.py:17-82, 218-244;
    extender.py:16-44; assist.py:105-133) against the golden vectors of the unmodified reference."""
    from xmap_b200.rdd import LocalRDD
    from xmap_b200.core import BaselinerSim, ExtendSim, baseliner_calculate_sim_pipeline
    from xmap_b200.utils.assist import extract_siminfo
    g = PT.load_golden("adj_low_overlap")
    iids = [str(s) for s in g["iids"]]
    uids = [str(s) for s in g["uids"]]
    ipos = {s: n for n, s in enumerate(iids)}
    trainRDD = LocalRDD(_train_records(g))
    tool = BaselinerSim(str(g["method"]), int(g["num_atleast"]))
    # get_universal_user_info / get_universal_item_info
    ui = dict(tool.get_universal_user_info(trainRDD).collect())
    assert np.array_equal(np.array([ui[u][0] for u in uids]), g["user_avg"])
    ii = dict(tool.get_universal_item_info(trainRDD).collect())
    info = np.array([ii[i] for i in iids])
    assert np.array_equal(info[:, 0], g["item_info"][:, 0]) and np.array_equal(info[:, 3], g["item_info"][:, 3])
    np.testing.assert_allclose(info[:, 1:3], g["item_info"][:, 1:3], rtol=1e-13)
    simRDD = baseliner_calculate_sim_pipeline(None, tool, trainRDD)
    # get_item_sim: adjacency rows (iid1, [(iid2, sim, mutu, frac)*])
    adj = dict(tool.get_item_sim(simRDD).collect())
    want_rows = {}
    for a, b, v, m, f in zip(g["sim_i"], g["sim_j"], g["sim_val"], g["sim_mutu"], g["sim_frac"]):
        want_rows.setdefault(iids[a], []).append((iids[b], v, float(m), f))
    assert set(adj) == set(want_rows)
    for a, lst in want_rows.items():
        got = sorted(adj[a])
        assert [(x[0], x[2], x[3]) for x in got] == [(x[0], x[2], x[3]) for x in sorted(lst)]
        np.testing.assert_allclose([x[1] for x in got], [x[1] for x in sorted(lst)], rtol=PT.SIM_RTOL)
    # build_sim_DF: the BB set is DISTINCT id1 WHERE label = 1 (assist.py:82-87)
    df = tool.build_sim_DF(simRDD)
    assert len(df) == len(g["sim_i"]) and set(df[0]) == {"id1", "id2", "sim", "mutu", "frac_mutu", "label"}
    bb = sorted({r["id1"] for r in df if r["label"] == 1})
    assert bb == [iids[q] for q in np.flatnonzero(g["bb"])]
    # find_knn_items + extract_siminfo: the four neighbour tables and the two dictionaries
    ext = ExtendSim(int(g["k"]))
    classified = ext.find_knn_items(simRDD)
    lists = {n: {} for n in ("BB_BB", "BB_NB", "NB_BB", "NB_NN")}
    for iid, B, N in classified.collect():
        if B is not None:
            lists["BB_BB"][iid], lists["BB_NB"][iid] = [x[0] for x in B[0]], [x[0] for x in B[1]]
        else:
            lists["NB_BB"][iid], lists["NB_NN"][iid] = [x[0] for x in N[0]], [x[0] for x in N[1]]
    for nm in lists:
        ptr, nbr = g[nm + "_ptr"], g[nm + "_nbr"]
        for it, iid in enumerate(iids):
            want = [iids[q] for q in nbr[ptr[it]:ptr[it + 1]]]
            assert lists[nm].get(iid, []) == want, (nm, iid)
    BB_info, NB_info, knn_BB, knn_NB = extract_siminfo(None, classified)
    assert sorted(knn_BB.value) == bb
    assert sorted(knn_NB.value) == [iids[q] for q in np.flatnonzero(g["valid_nb"])]
    some = bb[0]
    nb = set(lists["BB_BB"][some]) | set(lists["BB_NB"][some])
    assert set(knn_BB.value[some]) == nb and all(len(v) == 3 for v in knn_BB.value[some].values())
    # a neighbour's value tuple is (sim, mutu, frac) of the reference pair
    key = {(iids[a], iids[b]): (v, float(m), f) for a, b, v, m, f in
           zip(g["sim_i"], g["sim_j"], g["sim_val"], g["sim_mutu"], g["sim_frac"])}
    for nbr_id, (s_, m_, f_) in knn_BB.value[some].items():
        w = key[(some, nbr_id)]
        assert m_ == w[1] and f_ == w[2] and abs(s_ - w[0]) <= PT.SIM_RTOL * abs(w[0])


@pytest.mark.parametrize("name", ["adj_low_overlap", "adj_all_bridge"])
def test_recommender_pipelines_match_reference(name):
    """twodomain_demo.py:107-134 after the AlterEgo profile: recommender_calculate_sim_pipeline ->
    recommender_privacy_pipeline (non-private) -> recommender_prediction_pipeline, reference names and record
    shapes, against what the unmodified reference produced (tests/golden/*_recpred.npz)."""
    from xmap_b200.rdd import LocalRDD, Broadcast
    from xmap_b200.core import (RecommenderSim, RecommenderPrivacy, RecommenderPrediction,
                                recommender_calculate_sim_pipeline, recommender_privacy_pipeline,
                                recommender_prediction_pipeline)
    g0, g = PT.load_golden(name), PT.load_golden(name + "_recpred")
    iids, uids = [str(s) for s in g0["iids"]], [str(s) for s in g0["uids"]]
    profile = LocalRDD([(uids[u], iids[i], float(r), datetime.utcfromtimestamp(int(t)))
                        for u, i, r, t in zip(g["ae_user"], g["ae_item"], g["ae_rating"], g["ae_ts"])])
    # "adjust_cosine_item" runs the cosine_item branch in the reference (substring dispatch, recommenderSim.py:188)
    sim_tool = RecommenderSim("adjust_cosine_item" if name == "adj_all_bridge" else "cosine_item", int(g["num_atleast"]))
    out = recommender_calculate_sim_pipeline(None, sim_tool, profile)
    item_info = out[5].value
    assert len(out) == 7 and set(item_info) == {iids[i] for i in np.unique(g["ae_item"])}
    neigh = recommender_privacy_pipeline(RecommenderPrivacy(int(g["mapping_range"]), 0.6, 0.1), out[6], False).collectAsMap()
    for q, it in enumerate(g["nb_item"]):
        a, b = g["nb_ptr"][q], g["nb_ptr"][q + 1]
        got = neigh[iids[it]]
        if [x[0] for x in got] != [iids[j] for j in g["nb_idx"][a:b]]:        # only a near-tie may reorder a list
            assert np.allclose(np.abs([x[1] for x in got]), np.abs(g["nb_sim"][a:b]), rtol=1e-9, atol=0)
    # test records: one per user, in the golden's order
    test, cur = [], None
    for u, i, r in zip(g["test_user"], g["test_item"], g["test_rating"]):
        if cur is None or cur[0] != uids[u]:
            cur = (uids[u], []); test.append(cur)
        cur[1].append((iids[i], float(r), datetime(2013, 1, 1)))
    mae = recommender_prediction_pipeline(RecommenderPrediction(float(g["alpha"]), "cosine_item"), sim_tool, LocalRDD(test),
                                          Broadcast(neigh), out[2], out[3], out[4], out[5])
    m0, m1 = (float(x) for x in mae.split(";"))
    assert abs(m0 - float(g["mae_nodecay"])) < 1e-2 and abs(m1 - float(g["mae_decay"])) < 1e-2


def test_device_clean_matches_host_clean():
    """the clean stage's record-level rules on the device (xmap_clean_records) against the host class,
    which tests/test_host_cleansplit.py pins to the unmodified reference: same users, same (item, rating, time) lists in
    the same order -- with duplicates at other timestamps, exact timestamp ties and out-of-period years in the input."""
    from xmap_b200 import synth
    from xmap_b200.core import BaselinerClean
    from xmap_b200.rdd import LocalRDD
    sr = synth.make_ratings(600, 80, 14000, overlap=0.3, seed=9)
    lines = synth.to_text_lines(sr, 0)
    bump = lambda l, dt: l.rsplit("\t", 1)[0] + "\t" + str(int(l.rsplit("\t", 1)[1]) + dt)
    rerate = lambda l, r: "\t".join(l.split("\t")[:2] + [str(r), l.split("\t")[3]])
    extra = [bump(l, k * 40000000) for k, l in enumerate(lines[:300])]                 # later duplicates, some out of period
    extra += [bump(l, -3600) for l in lines[300:500]]                                   # earlier duplicates: must lose
    extra += [rerate(l, 1 + (k % 5)) for k, l in enumerate(lines[500:700])]             # exact ties: the first seen wins
    extra += [bump(l, -10 * 365 * 86400) for l in lines[700:800]]                       # out of period only
    rng = np.random.default_rng(2)
    lines = lines + extra
    lines = [lines[k] for k in rng.permutation(len(lines))]
    for atleast in (1, 5, 12):
        tool = BaselinerClean(atleast, 10 ** 9, 2012, 2013, "S:")
        want = tool.clean_data(tool.filter_data(tool.parse_data(LocalRDD(lines)))).collect()
        got = tool.device_pipeline(LocalRDD(lines)).collect()
        assert len(want) > 0 and [u for u, _ in got] == [u for u, _ in want], atleast
        for (u, a), (_, b) in zip(got, want):
            assert a == b, (atleast, u)
    assert len(BaselinerClean(10 ** 6, 1, 2012, 2013, "S:").device_pipeline(LocalRDD(lines)).collect()) == 0
    # the kernel's keep mask and per-user item counts against the array-level restatement, bit for bit
    from xmap_b200 import clean as CL
    rng = np.random.default_rng(3)
    n, nu, ni = 200000, 5000, 300
    user, item = rng.integers(0, nu, n).astype(np.int32), rng.integers(0, ni, n).astype(np.int32)
    ts = rng.integers(1.30e9, 1.40e9, n).astype(np.float64)
    ts[rng.integers(0, n, 20000)] = ts[rng.integers(0, n, 20000)]                    # exact ties across records
    t_lo, t_hi = CL.period_bounds(2012, 2013)
    for atleast in (1, 20, 38):
        keep, per_user = CL.clean_encoded(user, item, ts, nu, t_lo, t_hi, atleast)
        k0, p0 = PT.clean_encoded_numpy(user, item, ts, nu, t_lo, t_hi, atleast)
        assert np.array_equal(keep.cpu().numpy(), k0) and np.array_equal(per_user.cpu().numpy(), p0), atleast
        assert 0 < int(k0.sum()) < n


def test_private_recommender_pipeline_matches_reference():
    """recommender_privacy_pipeline(..., is_private=True) with the reference's draws injected: the rows the unmodified
    reference produced (one noisy neighbour per item), then the prediction stage runs on them."""
    from xmap_b200.rdd import LocalRDD, Broadcast
    from xmap_b200.core import (RecommenderSim, RecommenderPrivacy, RecommenderPrediction,
                                recommender_calculate_sim_pipeline, recommender_privacy_pipeline,
                                recommender_prediction_pipeline)
    name = "adj_low_overlap"
    g0, g, gp = PT.load_golden(name), PT.load_golden(name + "_recpriv"), PT.load_golden(name + "_recpred")
    iids, uids = [str(s) for s in g0["iids"]], [str(s) for s in g0["uids"]]
    profile = LocalRDD([(uids[u], iids[i], float(r), datetime.utcfromtimestamp(int(t)))
                        for u, i, r, t in zip(gp["ae_user"], gp["ae_item"], gp["ae_rating"], gp["ae_ts"])])
    sim_tool = RecommenderSim("cosine_item", int(g["num_atleast"]))
    out = recommender_calculate_sim_pipeline(None, sim_tool, profile)
    tool = RecommenderPrivacy(int(g["mapping_range"]), float(g["epsilon"]), float(g["rpo"]), uniforms=(g["u_pick"], g["u_noise"]))
    neigh = recommender_privacy_pipeline(tool, out[6], True).collectAsMap()
    assert len(neigh) == len(g["item"])
    for it, ch, ns in zip(g["item"], g["chosen"], g["noisy_sim"]):
        (nid, val), = neigh[iids[it]]
        assert nid == iids[ch] and abs(val - ns) <= 1e-9 * abs(ns) + 1e-12
    test, cur = [], None
    for u, i, r in zip(gp["test_user"], gp["test_item"], gp["test_rating"]):
        if cur is None or cur[0] != uids[u]:
            cur = (uids[u], []); test.append(cur)
        cur[1].append((iids[i], float(r), datetime(2013, 1, 1)))
    mae = recommender_prediction_pipeline(RecommenderPrediction(float(gp["alpha"]), "cosine_item"), sim_tool, LocalRDD(test),
                                          Broadcast(neigh), out[2], out[3], out[4], out[5])
    m0, m1 = (float(x) for x in mae.split(";"))
    assert 0.0 < m0 < 5.0 and 0.0 < m1 < 5.0
