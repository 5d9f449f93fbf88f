"""BaselinerClean -- host-side (string parsing, O(nnz); not on the measured path).
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
