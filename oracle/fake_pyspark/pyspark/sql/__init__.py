"""pyspark.sql stand-in: Row, DataFrame, SQLContext (TEST INFRASTRUCTURE ONLY).

The reference uses exactly one Row constructor (baselinerSim.py:240-243), one
registerTempTable and one literal query (assist.py:82-86).  Only that query is
understood; anything else raises.
"""
import re

from pyspark import RDD


class Row(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self._fields = tuple(kw)

    def asDict(self):
        return {k: self.__dict__[k] for k in self._fields}

    def __repr__(self):
        return "Row(%s)" % ", ".join(
            "%s=%r" % (k, self.__dict__[k]) for k in self._fields)


_TABLES = {}


class DataFrame(object):
    def __init__(self, rows, ctx=None):
        self._rows = list(rows)
        self.ctx = ctx

    def registerTempTable(self, name):
        _TABLES[name] = self

    def collect(self):
        return list(self._rows)

    def map(self, f):
        return RDD((f(r) for r in self._rows), self.ctx)

    @property
    def rdd(self):
        return RDD(self._rows, self.ctx)


_Q = re.compile(
    r"^\s*SELECT\s+DISTINCT\s+(\w+)\s+FROM\s+(\w+)\s+WHERE\s+(\w+)\s*=\s*(\d+)\s*$",
    re.I)


class SQLContext(object):
    def __init__(self, sc=None):
        self.sc = sc

    def sql(self, query):
        m = _Q.match(query)
        if not m:
            raise NotImplementedError("fake SQLContext: %r" % query)
        col, table, wcol, wval = m.group(1), m.group(2), m.group(3), int(m.group(4))
        seen = set()
        out = []
        for r in _TABLES[table]._rows:
            if getattr(r, wcol) == wval:
                v = getattr(r, col)
                if v not in seen:
                    seen.add(v)
                    out.append(Row(**{col: v}))
        return DataFrame(out, self.sc)
