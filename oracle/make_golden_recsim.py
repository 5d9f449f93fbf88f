"""Golden vectors for the NEXT row of the scope table: RecommenderSim.calculate_sim
(`cosine_item`, recommenderSim.py:90-132,186-195) run by the UNMODIFIED reference on the AlterEgo profile of
every golden case (the private / argmax profile frozen in tests/golden/<case>.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python -m oracle.make_golden_recsim
Writes tests/golden/<case>_recsim.npz; the existing golden files are not touched.
"""
import os
import sys
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H          # noqa: E402

CASES = ("adj_low_overlap", "cos_half_ratings", "adj_all_bridge")
NUM_ATLEAST = 50                          # parameters.yaml:28 calculate_xmap_weighting


def build(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    uids, iids = [str(u) for u in g["uids"]], [str(i) for i in g["iids"]]
    ipos = {s: n for n, s in enumerate(iids)}
    recs = [(uids[u], iids[i], float(r), datetime.utcfromtimestamp(int(t)))
            for u, i, r, t in zip(g["priv_ae_user"], g["priv_ae_item"], g["priv_ae_rating"], g["priv_ae_ts"])]
    sims, info, dt = H.run_recommender_sim(recs, "cosine_item", NUM_ATLEAST)
    present = sorted(info, key=lambda s: ipos[s])
    return dict(ae_user=g["priv_ae_user"], ae_item=g["priv_ae_item"], ae_rating=g["priv_ae_rating"],
                num_atleast=np.int64(NUM_ATLEAST),
                rs_i=np.array([ipos[a] for (a, b), _ in sims], dtype=np.int32),
                rs_j=np.array([ipos[b] for (a, b), _ in sims], dtype=np.int32),
                rs_sim=np.array([v[0] for _, v in sims]), rs_ls=np.array([v[1] for _, v in sims]),
                info_item=np.array([ipos[s] for s in present], dtype=np.int32),
                info_avg=np.array([float(info[s][0]) for s in present]),
                info_norm2=np.array([float(info[s][1]) for s in present]),
                info_count=np.array([int(info[s][2]) for s in present], dtype=np.int64),
                ref_seconds=np.array([dt]))


def main():
    for name in CASES:
        out = build(name)
        path = os.path.join(ROOT, "tests", "golden", name + "_recsim.npz")
        np.savez_compressed(path, **out)
        print("%-18s profile=%d pairs=%d (self-pairs %d, nan ls %d) ref s=%.2f -> %s (%d KB)" % (
            name, len(out["ae_user"]), len(out["rs_i"]), int((out["rs_i"] == out["rs_j"]).sum()),
            int(np.isnan(out["rs_ls"]).sum()), out["ref_seconds"][0], path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
