"""Synthetic Zipf-popularity rating data of the shapes BASELINE.json names.

The reference ships no data sample (code/data/README.md:1-9), so every
configuration runs on data generated here, in the reference's own record
shapes: text lines ``uid iid rating unix_ts`` (README.md:41-42) for the clean
stage, or ``(uid, [(iid+label, rating, datetime)])`` train records
(baselinerClean.py:49-52) for the hot path.

Recipe: per domain, item popularity ~ rank^-1.0 and
user activity ~ rank^-0.5 with ranks permuted independently per domain;
``n_draws / n_domains`` (user, item) draws per domain, deduplicated; a
fraction ``overlap`` of all users is shared by every domain, the rest are
split evenly; integer ratings 1..5 with P = (.05, .05, .10, .25, .55);
timestamps uniform in 2012-01-03 .. 2013-12-29 UTC; raw item ids carry a
domain-distinct 2-character prefix (what baselinerSim.py:191 compares) and the
reference's suffix labels (baselinerClean.py:51).
"""
from collections import namedtuple
from datetime import datetime

import numpy as np

DEFAULT_SEED = 20261018
RATING_P = np.array([0.05, 0.05, 0.10, 0.25, 0.55])
TS_LO = 1325548800   # 2012-01-03 00:00:00 UTC
TS_HI = 1388275200   # 2013-12-29 00:00:00 UTC
PREFIXES = ("bk", "mv", "mu", "ga", "tv", "ap", "to", "el")

SynthRatings = namedtuple(
    "SynthRatings",
    "user item domain rating ts n_users n_items_per_domain n_domains labels")


def _zipf_cdf(n, expo):
    w = np.arange(1, n + 1, dtype=np.float64) ** (-expo)
    c = np.cumsum(w)
    return c / c[-1]


def make_ratings(n_users, n_items_per_domain, n_draws, n_domains=2,
                 overlap=0.05, seed=DEFAULT_SEED, labels=None,
                 item_expo=1.0, user_expo=0.5):
    """Return deduplicated rating triples as flat numpy arrays.

    ``item`` is a global item number: domain ``d`` owns
    ``[d*n_items_per_domain, (d+1)*n_items_per_domain)``.
    Domain 0 is the source ("S:"), the last domain the target ("T:").
    """
    rng = np.random.default_rng(seed)
    if labels is None:
        if n_domains == 2:
            labels = ("S:", "T:")
        else:
            labels = tuple("S:%d:" % (d + 1) for d in range(n_domains - 1)) + ("T:",)
    n_ov = int(round(n_users * overlap))
    rest = n_users - n_ov
    per = rest // n_domains
    users, items, doms = [], [], []
    for d in range(n_domains):
        own = np.arange(n_ov + d * per, n_ov + (d + 1) * per if d < n_domains - 1
                        else n_users, dtype=np.int64)
        pool = np.concatenate([np.arange(n_ov, dtype=np.int64), own])
        pool = pool[rng.permutation(len(pool))]          # rank -> user
        iperm = rng.permutation(n_items_per_domain)      # rank -> item
        nd = n_draws // n_domains
        ucdf = _zipf_cdf(len(pool), user_expo)
        icdf = _zipf_cdf(n_items_per_domain, item_expo)
        u = pool[np.minimum(np.searchsorted(ucdf, rng.random(nd)), len(pool) - 1)]
        i = iperm[np.minimum(np.searchsorted(icdf, rng.random(nd)),
                             n_items_per_domain - 1)]
        key = np.sort(u * np.int64(n_items_per_domain) + i)   # sorted distinct keys (np.unique's hash path is 10x slower)
        key = key[np.concatenate([[True], key[1:] != key[:-1]])] if len(key) else key
        users.append(key // n_items_per_domain)
        items.append((key % n_items_per_domain + d * n_items_per_domain).astype(np.int64))
        doms.append(np.full(len(key), d, dtype=np.int8))
    user = np.concatenate(users)
    item = np.concatenate(items)
    dom = np.concatenate(doms)
    rating = (rng.choice(5, size=len(user), p=RATING_P) + 1).astype(np.int8)
    ts = rng.integers(TS_LO, TS_HI, size=len(user), dtype=np.int64)
    return SynthRatings(user, item, dom, rating, ts, n_users,
                        n_items_per_domain, n_domains, labels)


def raw_item_id(g, n_items_per_domain):
    d = g // n_items_per_domain
    return "%s%07d" % (PREFIXES[d], g % n_items_per_domain)


def item_id(g, n_items_per_domain, labels):
    """Full item id as it appears after the clean stage (raw id + suffix label)."""
    return raw_item_id(g, n_items_per_domain) + labels[g // n_items_per_domain]


def user_id(u):
    return "u%08d" % u


def to_text_lines(sr, domain):
    """Raw 4-column lines for one domain, as sc.textFile would yield them."""
    m = sr.domain == domain
    return ["%s\t%s\t%d\t%d" % (user_id(u), raw_item_id(i, sr.n_items_per_domain), r, t)
            for u, i, r, t in zip(sr.user[m], sr.item[m], sr.rating[m], sr.ts[m])]


def to_train_records(sr):
    """``(uid, [(iid+label, rating, datetime)])`` records, one per user.

    Users in ascending user number, each user's ratings in ascending item id
    string (the canonical order of SURVEY.md App. A.6 rule 1).
    """
    order = np.lexsort((sr.item, sr.user))
    recs = []
    cur_u, cur = None, None
    for k in order:
        u = int(sr.user[k])
        if u != cur_u:
            if cur is not None:
                recs.append((user_id(cur_u), cur))
            cur_u, cur = u, []
        cur.append((item_id(int(sr.item[k]), sr.n_items_per_domain, sr.labels),
                    float(sr.rating[k]),
                    datetime.utcfromtimestamp(int(sr.ts[k]))))
    if cur is not None:
        recs.append((user_id(cur_u), cur))
    for _, lst in recs:
        lst.sort(key=lambda x: x[0])
    return recs


def workload_stats(sr):
    """nnz, W = sum_u d_u (d_u - 1), max degrees -- the sizes SURVEY 8(d) quotes."""
    du = np.bincount(sr.user, minlength=sr.n_users).astype(np.int64)
    ci = np.bincount(sr.item, minlength=sr.n_items_per_domain * sr.n_domains)
    return {"nnz": int(len(sr.user)), "W": int((du * (du - 1)).sum()),
            "d_max": int(du.max()), "c_max": int(ci.max()),
            "n_users_active": int((du > 0).sum()), "n_items_active": int((ci > 0).sum())}
