"""The reference's OWN similarity stage (unmodified classes through oracle/fake_pyspark, one core) on the same
bounded sample bench.py's cpu_baseline times the numpy port on.  Needs /root/reference: runs in the build container
only; the result is recorded in profiles/ and quoted by bench.py as `cpu_baseline.reference_own_python`.

  python tools/ref_shim_time.py [workload=cfg2] [target_ratings=250000]
"""
import json, os, sys, time
from datetime import datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import harness as H

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
target = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000
wl = bench.make_workload(name)
stride = max(1, int(round(wl["nnz"] / max(1, target))))
(user, item, rating, n_u, n_i, pc), nr = bench._sample(wl, 0, stride)
o = np.lexsort((item, user))
user, item, rating = user[o], item[o], rating[o]
t0 = datetime(2013, 1, 1)
# ids: a domain-distinct 2-character prefix + the reference's suffix label, like the bench items
lab = np.where(pc == pc.min(), "S:", "T:")
pre = np.where(pc == pc.min(), "bk", "mv")
iid = np.array(["%s%07d%s" % (pre[q], q, lab[q]) for q in range(n_i)])
recs, cur = [], None
for u, i, r in zip(user, item, rating):
    if cur is None or cur[0] != u:
        cur = (u, []); recs.append(cur)
    cur[1].append((iid[i], float(r), t0))
recs = [("u%08d" % u, lst) for u, lst in recs]
tool, trainRDD, simRDD, dt = H.run_sim(recs, "adjust_cosine", 50)
kept = simRDD.collect()
du = np.bincount(user)
emissions = int((du * (du - 1)).sum())
# directed co-rated pairs (pre-filter) = the metric's unit: count them from the arrays
import scipy.sparse as sp
M = sp.csr_matrix((np.ones(len(user)), (user, item)), shape=(n_u, n_i))
N = (M.T @ M).tocoo()
pairs = int((N.row != N.col).sum())
print(json.dumps({"workload": bench.workload_label(wl, "adjust_cosine"), "sample": "every %d-th user: %d users, %d ratings" % (stride, len(recs), nr),
                  "emissions": emissions, "corated_pairs": pairs, "kept_pairs": len(kept), "seconds": dt, "cores": 1,
                  "pairs_per_s": pairs / dt, "what": "BaselinerSim (unmodified reference classes, oracle/fake_pyspark single-partition RDD shim, "
                  "no Spark/JVM/shuffle serialisation), this build container's CPU"}))
