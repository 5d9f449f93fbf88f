"""CPU: the host-side clean / split stages against the unmodified reference run through the
oracle shim (needs /root/reference; skipped where it is not mounted)."""
import os
import subprocess
import sys
import textwrap

import pytest

from tests import parity as PT

needs_ref = pytest.mark.skipif(not os.path.isdir("/root/reference/code/xmap"),
                               reason="reference tree not mounted")

_SCRIPT = textwrap.dedent("""
    import sys, json
    sys.path.insert(0, %(root)r)
    from xmap_b200 import synth
    from xmap_b200.core import (BaselinerClean, BaselinerSplit, baseliner_clean_data_pipeline,
                                baseliner_split_data_pipeline)
    sr = synth.make_ratings(400, 60, 9000, overlap=0.3, seed=4)
    paths = []
    for d, nm in ((0, "book"), (1, "movie")):
        lines = synth.to_text_lines(sr, d)
        # duplicates with other timestamps / out-of-period years exercise the filters
        extra = [l.rsplit("\\t", 1)[0] + "\\t" + str(int(l.rsplit("\\t", 1)[1]) + k * 40000000)
                 for k, l in enumerate(lines[:200])]
        p = %(tmp)r + "/" + nm + ".txt"
        open(p, "w").write("\\n".join(lines + extra) + "\\n")
        paths.append(p)
    def run(mods, sc):
        cs = mods["BaselinerClean"](5, 150, 2012, 2013, "S:"); ct = mods["BaselinerClean"](5, 150, 2012, 2013, "T:")
        sp = mods["BaselinerSplit"](0, 0.2, 0.8, 666666)
        out = {}
        for dbg in (True, False):
            s = mods["clean"](sc, cs, paths[0], dbg, 30); t = mods["clean"](sc, ct, paths[1], dbg, 30)
            tr, te = mods["split"](sc, sp, s, t)
            norm = lambda rdd: sorted((u, sorted((i, r, str(x)) for i, r, x in l)) for u, l in rdd.collect())
            out[str(dbg)] = [norm(s), norm(t), norm(tr), norm(te)]
        return out
    mine = run(dict(BaselinerClean=BaselinerClean, BaselinerSplit=BaselinerSplit,
                    clean=baseliner_clean_data_pipeline, split=baseliner_split_data_pipeline), None)
    from oracle import harness as H
    R = H.load_reference()
    ref = run(dict(BaselinerClean=R["BaselinerClean"], BaselinerSplit=R["BaselinerSplit"],
                   clean=R["assist"].baseliner_clean_data_pipeline,
                   split=R["assist"].baseliner_split_data_pipeline), R["sc"])
    for dbg in ("True", "False"):
        for k, (a, b) in enumerate(zip(mine[dbg], ref[dbg])):
            assert a == b, (dbg, k, len(a), len(b))
        assert len(mine[dbg][2]) > 0 and len(mine[dbg][3]) > 0
    print("ok")
""")


@needs_ref
def test_clean_and_split_match_reference(tmp_path):
    code = _SCRIPT % dict(root=PT.ROOT, tmp=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=PT.ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


def test_local_rdd_surface():
    from xmap_b200.rdd import LocalRDD, LazyRDD
    r = LocalRDD([("a", 1), ("b", 2), ("a", 3)])
    assert r.reduceByKey(lambda x, y: x + y).collectAsMap() == {"a": 4, "b": 2}
    assert r.map(lambda kv: kv[1]).filter(lambda v: v > 1).collect() == [2, 3]
    assert r.join(LocalRDD([("a", "x")])).collect() == [("a", (1, "x")), ("a", (3, "x"))]
    assert r.keys().distinct().collect() == ["a", "b"]
    calls = []
    lz = LazyRDD(lambda: calls.append(1) or [1, 2, 3])
    assert not calls and lz.count() == 3 and lz.collect() == [1, 2, 3] and len(calls) == 1
    parts = LocalRDD(range(1000)).randomSplit([0.2, 0.8], seed=7)
    assert sum(p.count() for p in parts) == 1000 and 120 < parts[0].count() < 280


def test_encoder_replicates_reference_string_tests():
    from xmap_b200.encode import encode_records, item_codes
    from datetime import datetime
    t = datetime(2012, 5, 1)
    recs = [("u2", [("mv01T:", 4.0, t), ("bkS:9S:", 2.5, t)]), ("u1", [("bk07S:", 5.0, t)])]
    e = encode_records(recs)
    assert list(e.uids) == ["u1", "u2"] and list(e.iids) == ["bk07S:", "bkS:9S:", "mv01T:"]
    assert list(e.prefix_code) == [0, 0, 1] and list(e.dom_code) == [0, 0, 1]
    assert list(e.has_S) == [True, True, False] and list(e.has_T) == [False, False, True]
    assert list(e.contains) == [1, 1, 2]
    with pytest.raises(ValueError):
        encode_records([("u", [("bkS:", 3.7000001, t)])])
    pc, dc, ct, hs, ht = item_codes(["xxT:S:"])       # a label inside the raw id still counts (substring test)
    assert ct[0] == 3 or ct[0] == 1


def test_clean_restatement_matches_host_class(monkeypatch):
    """The array-level restatement of the clean stage's record rules (tests/parity.clean_encoded_numpy -- what the device
    kernel is checked against) drives BaselinerClean.device_pipeline to exactly the records of the host class, which
    test_clean_and_split_match_reference pins to the unmodified reference."""
    import numpy as np
    import torch
    from xmap_b200 import synth, clean as CL
    from xmap_b200.core import BaselinerClean
    from xmap_b200.rdd import LocalRDD
    monkeypatch.setattr(CL, "clean_encoded", lambda u, i, t, nu, lo, hi, k, device="cpu":
                        tuple(torch.as_tensor(x) for x in PT.clean_encoded_numpy(u, i, t, nu, lo, hi, k)))
    sr = synth.make_ratings(300, 50, 6000, overlap=0.3, seed=5)
    lines = synth.to_text_lines(sr, 1)
    bump = lambda l, dt: l.rsplit("\t", 1)[0] + "\t" + str(int(l.rsplit("\t", 1)[1]) + dt)
    lines = lines + [bump(l, k * 40000000) for k, l in enumerate(lines[:150])] + [bump(l, -3600) for l in lines[150:250]] + \
        ["\t".join(l.split("\t")[:2] + ["1", l.split("\t")[3]]) for l in lines[250:350]]
    rng = np.random.default_rng(1)
    lines = [lines[k] for k in rng.permutation(len(lines))]
    assert CL.period_bounds(2012, 2013)[0] < 1.34e9 < CL.period_bounds(2012, 2013)[1]
    for atleast in (1, 4, 9):
        tool = BaselinerClean(atleast, 10 ** 9, 2012, 2013, "T:")
        want = tool.clean_data(tool.filter_data(tool.parse_data(LocalRDD(lines)))).collect()
        got = tool.device_pipeline(LocalRDD(lines)).collect()
        assert len(want) > 0 and got == want, atleast
