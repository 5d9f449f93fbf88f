"""In-process stand-in for the slice of PySpark 1.6 the X-MAP hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): this lets the *unmodified*
reference classes under /root/reference/code/xmap run inside one Python
process so that their outputs can be frozen into tests/golden/.  Nothing in
the product package imports this module.

Semantics: one partition, records kept in a Python list, every transformation
evaluated eagerly, and every keyed operation preserves first-appearance order
of keys and encounter order of values.  That makes the reference deterministic
and lets the harness impose the canonical-order rules of SURVEY.md App. A.6 by
ordering the inputs it hands to each pipeline stage.

Surface covered: map, flatMap, filter, mapPartitions,
keys, values, cache, collect, collectAsMap, take, count, union, intersection,
reduceByKey, aggregateByKey, combineByKey, join, randomSplit, reduce, toDF on
RDD; broadcast / parallelize / textFile on SparkContext.
"""
import numpy as _np


class Broadcast(object):
    def __init__(self, value):
        self.value = value


class RDD(object):
    def __init__(self, records, ctx=None):
        self._d = list(records)
        self.ctx = ctx

    # -- bookkeeping -------------------------------------------------------
    def _new(self, records):
        return RDD(records, self.ctx)

    def cache(self):
        return self

    persist = cache

    def unpersist(self):
        return self

    # -- element-wise ------------------------------------------------------
    def map(self, f):
        return self._new(f(x) for x in self._d)

    def flatMap(self, f):
        out = []
        for x in self._d:
            out.extend(f(x))
        return self._new(out)

    def filter(self, f):
        return self._new(x for x in self._d if f(x))

    def mapPartitions(self, f):
        return self._new(f(iter(self._d)))

    def keys(self):
        return self._new(x[0] for x in self._d)

    def values(self):
        return self._new(x[1] for x in self._d)

    # -- actions -----------------------------------------------------------
    def collect(self):
        return list(self._d)

    def collectAsMap(self):
        return dict(self._d)

    def take(self, n):
        return list(self._d[:n])

    def count(self):
        return len(self._d)

    def first(self):
        return self._d[0]

    def reduce(self, f):
        it = iter(self._d)
        acc = next(it)
        for x in it:
            acc = f(acc, x)
        return acc

    # -- set-like ----------------------------------------------------------
    def union(self, other):
        return self._new(self._d + other._d)

    def intersection(self, other):
        theirs = set(other._d)
        seen = set()
        out = []
        for x in self._d:
            if x in theirs and x not in seen:
                seen.add(x)
                out.append(x)
        return self._new(out)

    def distinct(self):
        seen = set()
        out = []
        for x in self._d:
            if x not in seen:
                seen.add(x)
                out.append(x)
        return self._new(out)

    # -- keyed -------------------------------------------------------------
    def combineByKey(self, create, merge_value, merge_combiners):
        acc = {}
        for k, v in self._d:
            if k in acc:
                acc[k] = merge_value(acc[k], v)
            else:
                acc[k] = create(v)
        return self._new(acc.items())

    def reduceByKey(self, f):
        return self.combineByKey(lambda v: v, f, f)

    def aggregateByKey(self, zero, seq_op, comb_op):
        import copy
        acc = {}
        for k, v in self._d:
            if k not in acc:
                acc[k] = copy.deepcopy(zero)
            acc[k] = seq_op(acc[k], v)
        return self._new(acc.items())

    def groupByKey(self):
        acc = {}
        for k, v in self._d:
            acc.setdefault(k, []).append(v)
        return self._new(acc.items())

    def join(self, other):
        right = {}
        for k, v in other._d:
            right.setdefault(k, []).append(v)
        out = []
        for k, v in self._d:
            for w in right.get(k, ()):
                out.append((k, (v, w)))
        return self._new(out)

    # -- sampling ----------------------------------------------------------
    def randomSplit(self, weights, seed=None):
        """Per-record cell sampling like Spark's, but with numpy's generator.

        Spark's XORShift stream is not reproduced (the split is upstream of
        the hot path and only has to be *a* deterministic by-user split).
        """
        w = _np.asarray(weights, dtype=_np.float64)
        edges = _np.cumsum(w / w.sum())
        rng = _np.random.RandomState(seed if seed is None else seed % (2**32))
        draws = rng.random_sample(len(self._d))
        cell = _np.searchsorted(edges, draws, side="right")
        cell = _np.minimum(cell, len(w) - 1)
        return [self._new(x for x, c in zip(self._d, cell) if c == i)
                for i in range(len(w))]

    # -- SQL bridge --------------------------------------------------------
    def toDF(self):
        from pyspark.sql import DataFrame
        return DataFrame(self._d, self.ctx)


class SparkConf(object):
    def __init__(self):
        self._c = {}

    def setAppName(self, name):
        self._c["spark.app.name"] = name
        return self

    def setMaster(self, m):
        self._c["spark.master"] = m
        return self

    def set(self, k, v):
        self._c[k] = v
        return self


class SparkContext(object):
    def __init__(self, conf=None, **_):
        self.conf = conf

    def broadcast(self, value):
        return Broadcast(value)

    def parallelize(self, data, numSlices=None):
        return RDD(data, self)

    def textFile(self, path, minPartitions=None):
        if path.startswith("file:"):
            path = path[len("file:"):]
        with open(path) as f:
            return RDD((line.rstrip("\n") for line in f if line.strip()), self)

    def stop(self):
        pass
