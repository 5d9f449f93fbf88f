"""Golden vectors for the PRIVATE branch of SURVEY.md 8(f) #2: private_neighbor_selection + noise_perturbation
(recommenderPrivacy.py:35-139, 152-171) run by the UNMODIFIED reference on the RecommenderSim output of every golden
case, with the reference's np.random draws logged so that they can be replayed as injected uniforms.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python -m oracle.make_golden_recpriv
Writes tests/golden/<case>_recpriv.npz.
"""
import os
import sys
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H          # noqa: E402

CASES = ("adj_low_overlap", "cos_half_ratings")
MAPPING_RANGE, EPSILON, RPO, NUM_ATLEAST = 10, 0.6, 0.1, 50      # parameters.yaml:28-33


def build(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    uids, iids = [str(u) for u in g["uids"]], [str(i) for i in g["iids"]]
    ipos = {s: n for n, s in enumerate(iids)}
    au, ai, ar, at = g["priv_ae_user"], g["priv_ae_item"], g["priv_ae_rating"], g["priv_ae_ts"]
    recs = [(uids[u], iids[i], float(r), datetime.utcfromtimestamp(int(t))) for u, i, r, t in zip(au, ai, ar, at)]
    sims, _, _ = H.run_recommender_sim(recs, "cosine_item", NUM_ATLEAST)
    rows = H.run_recommender_private(sims, MAPPING_RANGE, EPSILON, RPO, np_seed=20261018)
    return dict(ae_user=au, ae_item=ai, ae_rating=ar, mapping_range=np.int64(MAPPING_RANGE), epsilon=np.float64(EPSILON),
                rpo=np.float64(RPO), num_atleast=np.int64(NUM_ATLEAST),
                item=np.array([ipos[r[0]] for r in rows], dtype=np.int32),
                chosen=np.array([ipos[r[1]] for r in rows], dtype=np.int32),
                noisy_sim=np.array([r[2] for r in rows]), u_pick=np.array([r[3] for r in rows]),
                u_noise=np.array([r[4] for r in rows]))


def main():
    for name in CASES:
        out = build(name)
        path = os.path.join(ROOT, "tests", "golden", name + "_recpriv.npz")
        np.savez_compressed(path, **out)
        print("%-18s items=%d -> %s (%d KB)" % (name, len(out["item"]), path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
