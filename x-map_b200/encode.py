"""Host-side dictionary encoding: reference records <-> dense integer arrays.

The reference carries string ids and datetimes through every RDD
(baselinerClean.py:49-52).  The kernels work on dense int32 indices, so this
module numbers users and items by the rank of their id string in sorted order
(the canonical index of SURVEY.md App. A.6: "ties resolve to the smaller
index") and derives, per item, the four string tests the reference applies
downstream -- all of which must be honoured as written, bugs included:

  prefix_code  iid[:2]             baselinerSim.py:189-191 (cross-domain label)
  dom_code     iid[-2:]            extender.py:29
  contains     label in iid        extender.py:32,34 (substring test)
  has_S/has_T  "S:" in / "T:" in   extender.py:68,79,174-175; generator.py:157
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Encoded:
    uids: np.ndarray            # sorted unique user id strings
    iids: np.ndarray            # sorted unique item id strings
    user: np.ndarray            # int32 [nnz]
    item: np.ndarray            # int32 [nnz]
    rating: np.ndarray          # float64 [nnz] (checked to be exact in float32)
    times: list                 # original time objects, one per rating
    prefix_code: np.ndarray
    dom_code: np.ndarray
    contains: np.ndarray
    has_S: np.ndarray
    has_T: np.ndarray

    @property
    def n_users(self):
        return len(self.uids)

    @property
    def n_items(self):
        return len(self.iids)


def item_codes(iids):
    """The per-item codes listed in the module docstring, from id strings."""
    iids = [str(s) for s in iids]
    pre = np.array([s[:2] for s in iids])
    suf = np.array([s[-2:] for s in iids])
    _, prefix_code = np.unique(pre, return_inverse=True) if len(iids) else (None, np.zeros(0, int))
    labels, dom_code = np.unique(suf, return_inverse=True) if len(iids) else ([], np.zeros(0, int))
    if len(labels) > 8:
        raise ValueError("more than 8 distinct 2-character id suffixes; at most 8 domain labels are supported")
    contains = np.zeros(len(iids), dtype=np.uint8)
    for d, lab in enumerate(labels):
        contains |= (np.array([lab in s for s in iids], dtype=np.uint8) << d)
    has_S = np.array(["S:" in s for s in iids], dtype=bool)
    has_T = np.array(["T:" in s for s in iids], dtype=bool)
    return (prefix_code.astype(np.int32), dom_code.astype(np.uint8), contains, has_S, has_T)


def check_ratings_f32(rating):
    r = np.asarray(rating, dtype=np.float64)
    if not np.array_equal(r.astype(np.float32).astype(np.float64), r):
        raise ValueError("ratings must be exactly representable in binary32 "
                         "(integer / half-star scales are); got values that are not")
    return r


def encode_records(records):
    """records: iterable of (uid, [(iid, rating, time)]) as produced by the clean /
    split stages (baselinerClean.py:83, baselinerSplit.py:102-106)."""
    u_l, i_l, r_l, t_l = [], [], [], []
    for uid, lst in records:
        for iid, r, t in lst:
            u_l.append(uid); i_l.append(iid); r_l.append(r); t_l.append(t)
    if u_l:
        uids, user = np.unique(np.array(u_l), return_inverse=True)
        iids, item = np.unique(np.array(i_l), return_inverse=True)
    else:
        uids = iids = np.zeros(0, dtype="<U1")
        user = item = np.zeros(0, dtype=np.int64)
    rating = check_ratings_f32(np.array(r_l, dtype=np.float64))
    pc, dc, ct, hs, ht = item_codes(iids)
    return Encoded(uids, iids, user.astype(np.int32), item.astype(np.int32), rating, t_l,
                   pc, dc, ct, hs, ht)


def encode_arrays(user, item, rating, iids, uids=None, times=None):
    """Already-indexed triples plus the item id strings (synthetic / bench path)."""
    pc, dc, ct, hs, ht = item_codes(iids)
    n_users = int(user.max()) + 1 if len(user) else 0
    if uids is None:
        uids = np.arange(n_users)
    return Encoded(np.asarray(uids), np.asarray(iids), np.asarray(user, dtype=np.int32),
                   np.asarray(item, dtype=np.int32), check_ratings_f32(rating),
                   times if times is not None else list(range(len(user))), pc, dc, ct, hs, ht)
