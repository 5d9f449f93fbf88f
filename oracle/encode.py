"""String records -> canonical integer arrays for oracle/restate.py.

TEST INFRASTRUCTURE ONLY (kept separate from the product's own encoder so the
oracle does not depend on the code it checks).
"""
import calendar

import numpy as np


def encode_train(records, labels=("S:", "T:")):
    """records: iterable of (uid, [(iid, rating, datetime)]).  Canonical index =
    rank of the id string in sorted order."""
    uids = sorted({uid for uid, _ in records})
    iids = sorted({x[0] for _, lst in records for x in lst})
    upos = {u: k for k, u in enumerate(uids)}
    ipos = {i: k for k, i in enumerate(iids)}
    user, item, rating, ts = [], [], [], []
    for uid, lst in records:
        for iid, r, t in lst:
            user.append(upos[uid]); item.append(ipos[iid]); rating.append(r)
            ts.append(calendar.timegm(t.timetuple()))
    prefixes = sorted({i[:2] for i in iids})
    suffixes = sorted({i[-2:] for i in iids})
    meta = dict(
        uids=uids, iids=iids,
        prefix_code=np.array([prefixes.index(i[:2]) for i in iids], dtype=np.int32),
        dom_code=np.array([suffixes.index(i[-2:]) for i in iids], dtype=np.int32),
        contains=np.array([sum((1 << d) for d, s in enumerate(suffixes) if s in i)
                           for i in iids], dtype=np.int32),
        has_S=np.array(["S:" in i for i in iids], dtype=bool),
        has_T=np.array(["T:" in i for i in iids], dtype=bool))
    return (np.array(user, dtype=np.int64), np.array(item, dtype=np.int64),
            np.array(rating, dtype=np.float64), np.array(ts, dtype=np.int64), meta)
