"""Golden vectors for SURVEY.md 8(f) #2-#3: non-private neighbour selection (recommenderPrivacy.py:141-178)
and item-based prediction + MAE (recommenderPrediction.py:26-139) run by the UNMODIFIED reference on the
AlterEgo profile of every golden case, against a generated test set of hidden target-domain ratings.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python -m oracle.make_golden_recpred
Writes tests/golden/<case>_recpred.npz.
"""
import calendar
import os
import sys
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H          # noqa: E402

CASES = ("adj_low_overlap", "cos_half_ratings", "adj_all_bridge")
MAPPING_RANGE, ALPHA, NUM_ATLEAST = 10, 0.03, 50      # parameters.yaml:28-33


def build(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    uids, iids = [str(u) for u in g["uids"]], [str(i) for i in g["iids"]]
    ipos = {s: n for n, s in enumerate(iids)}
    upos = {s: n for n, s in enumerate(uids)}
    au, ai, ar, at = g["priv_ae_user"], g["priv_ae_item"], g["priv_ae_rating"], g["priv_ae_ts"]
    recs = [(uids[u], iids[i], float(r), datetime.utcfromtimestamp(int(t))) for u, i, r, t in zip(au, ai, ar, at)]
    # test set: hidden target ratings -- per profile user a few target items (half of them items the user's
    # profile does not hold), integer ratings, one record per user like the reference's testRDD
    rng = np.random.default_rng(2026)
    t_items = np.unique(ai)
    users = np.unique(au)
    test = []
    tu, ti, tr = [], [], []
    for u in users[rng.random(len(users)) < 0.5]:
        k = int(rng.integers(1, 5))
        its = rng.choice(t_items, size=min(k, len(t_items)), replace=False)
        lst = []
        for it in sorted(its):
            r = float(rng.integers(1, 6))
            ts = int(rng.integers(1325548800, 1388275200))
            lst.append((iids[it], r, datetime.utcfromtimestamp(ts)))
            tu.append(u); ti.append(it); tr.append(r)
        test.append((uids[u], lst))
    neigh, preds, mae = H.run_recommender_predict(recs, test, MAPPING_RANGE, ALPHA, "cosine_item", NUM_ATLEAST)
    # neighbour tables
    nb_item = sorted(neigh, key=lambda s: ipos[s])
    nb_ptr = np.zeros(len(nb_item) + 1, dtype=np.int32)
    nb_idx, nb_sim = [], []
    for q, s in enumerate(nb_item):
        nb_ptr[q + 1] = nb_ptr[q] + len(neigh[s])
        nb_idx += [ipos[x[0]] for x in neigh[s]]; nb_sim += [float(x[1]) for x in neigh[s]]
    # predictions in test order
    pu, pi, pr, p0, p1 = [], [], [], [], []
    for uid, lst in preds:
        for (iid, real, _t), p in zip(dict(test)[uid], lst):
            pu.append(upos[uid]); pi.append(ipos[iid]); pr.append(real)
            if p == ():
                p0.append(-1.0); p1.append(-1.0)
            else:
                assert p[0] == iid and p[1] == real
                p0.append(float(p[2])); p1.append(float(p[3]))
    m0, m1 = (float(x) for x in mae.split(";"))
    return dict(ae_user=au, ae_item=ai, ae_rating=ar, ae_ts=at,
                mapping_range=np.int64(MAPPING_RANGE), alpha=np.float64(ALPHA), num_atleast=np.int64(NUM_ATLEAST),
                nb_item=np.array([ipos[s] for s in nb_item], dtype=np.int32), nb_ptr=nb_ptr,
                nb_idx=np.array(nb_idx, dtype=np.int32), nb_sim=np.array(nb_sim),
                test_user=np.array(pu, dtype=np.int32), test_item=np.array(pi, dtype=np.int32), test_rating=np.array(pr),
                pred_nodecay=np.array(p0), pred_decay=np.array(p1), mae_nodecay=np.float64(m0), mae_decay=np.float64(m1))


def main():
    for name in CASES:
        out = build(name)
        path = os.path.join(ROOT, "tests", "golden", name + "_recpred.npz")
        np.savez_compressed(path, **out)
        print("%-18s items with neighbours=%d test pairs=%d predicted=%d mae=%.4f / %.4f -> %s (%d KB)" % (
            name, len(out["nb_item"]), len(out["test_user"]), int((out["pred_nodecay"] >= 0).sum()),
            out["mae_nodecay"], out["mae_decay"], path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
