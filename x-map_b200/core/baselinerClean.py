"""BaselinerClean -- host-side string parsing; the data-parallel part (period test, latest-rating dedupe, minimum
ratings per user) also runs on the device (`device_pipeline`, csrc/clean.cu).
Semantics of xmap/core/baselinerClean.py:7-97: whitespace-split 4-column lines, keep ratings whose
local-time year lies in [date_from, date_to], suffix the item id with the domain label, keep per
(user, item) the strictly latest rating, drop users with fewer than num_atleast_rating items."""
from datetime import datetime

from ..rdd import LocalRDD, records_of


class BaselinerClean(object):
    def __init__(self, num_atleast_rating, size_subset, date_from, date_to, domain_label):
        self.num_atleast_rating = num_atleast_rating
        self.size_subset = size_subset
        self.period = range(date_from, date_to + 1)
        self.label = domain_label

    def parse_time(self, s):
        return datetime.fromtimestamp(float(s))        # local timezone, as baselinerClean.py:36

    def parse_data(self, originalRDD):
        """(uid, (iid + label, rating, time))* for in-period lines (baselinerClean.py:38-56)."""
        out = []
        for line in records_of(originalRDD):
            f = line.split()
            if len(f) < 4:
                raise IndexError("expected 'uid iid rating timestamp', got %r" % (line,))
            t = self.parse_time(f[3])
            if t.year in self.period:
                out.append((f[0], (f[1] + self.label, float(f[2]), t)))
        return LocalRDD(out)

    def filter_data(self, dataRDD):
        """Group by user; per item keep the strictly latest rating, the first seen winning
        ties (baselinerClean.py:62-92)."""
        users = {}
        for uid, (iid, r, t) in records_of(dataRDD):
            per = users.setdefault(uid, {})
            old = per.get(iid)
            if old is None or t > old[2]:
                per[iid] = (iid, r, t)
        return LocalRDD((uid, list(per.values())) for uid, per in users.items())

    def clean_data(self, filteredRDD):
        return LocalRDD(rec for rec in records_of(filteredRDD) if len(rec[1]) >= self.num_atleast_rating)

    def take_partial_data(self, dataRDD):
        return records_of(dataRDD)[: self.size_subset]

    def device_pipeline(self, originalRDD):
        """parse_data -> filter_data -> clean_data with the record-level work on the GPU (xmap_clean_records): the lines
        are split and the ids numbered on the host, the period test, the per-(user, item) latest-rating rule and the
        minimum-ratings rule run on the device.  Same records as clean_data(filter_data(parse_data(rdd))); users in
        first-appearance order, a user's items in first-appearance order."""
        import numpy as np
        import torch
        from .. import clean as CL
        lines = records_of(originalRDD)
        uid, iid, rating, ts = [], [], [], []
        for line in lines:
            f = line.split()
            if len(f) < 4:
                raise IndexError("expected 'uid iid rating timestamp', got %r" % (line,))
            uid.append(f[0]); iid.append(f[1]); rating.append(float(f[2])); ts.append(float(f[3]))
        uids, user = np.unique(np.array(uid, dtype=object).astype(str), return_inverse=True) if uid else ([], np.zeros(0, np.int64))
        iids, item = np.unique(np.array(iid, dtype=object).astype(str), return_inverse=True) if iid else ([], np.zeros(0, np.int64))
        t_lo, t_hi = CL.period_bounds(self.period[0], self.period[-1])
        keep, _ = CL.clean_encoded(user.astype(np.int32), item.astype(np.int32), np.array(ts, dtype=np.float64), len(uids),
                                   t_lo, t_hi, self.num_atleast_rating)
        kept = torch.nonzero(keep).flatten().cpu().numpy()
        # the reference's per-user dict keeps a pair where it FIRST appeared among the in-period lines, and users come in
        # first-appearance order: order the output the same way
        tsa = np.array(ts, dtype=np.float64)
        first_pos, first_user = {}, {}
        for k in np.flatnonzero((tsa >= t_lo) & (tsa < t_hi)):
            first_pos.setdefault((user[k], item[k]), k)
            first_user.setdefault(uid[k], k)
        users = {}
        for k in sorted(kept, key=lambda k: first_pos[(user[k], item[k])]):
            users.setdefault(uid[k], []).append((iid[k] + self.label, rating[k], self.parse_time(str(ts[k]))))
        return LocalRDD(sorted(users.items(), key=lambda kv: first_user[kv[0]]))
