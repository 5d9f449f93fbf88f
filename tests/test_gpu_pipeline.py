"""GPU parity at the reference's own boundary: the five pipeline functions with the
reference's record shapes (string ids, datetimes), compared with what the unmodified
reference produced on the same records (tests/golden)."""
import calendar
from datetime import datetime

import numpy as np
import pytest

from tests import parity as PT

pytestmark = pytest.mark.gpu


def _train_records(g):
    recs = {}
    for u, i, r, t in zip(g["user"], g["item"], g["rating"], g["ts"]):
        recs.setdefault(str(g["uids"][u]), []).append((str(g["iids"][i]), float(r), datetime.utcfromtimestamp(int(t))))
    return list(recs.items())


@pytest.mark.parametrize("name", ["adj_low_overlap", "cos_half_ratings"])
def test_pipelines_match_reference(name):
    from xmap_b200.rdd import LocalRDD
    from xmap_b200.core import (BaselinerSim, ExtendSim, Generator, baseliner_calculate_sim_pipeline,
                                extender_pipeline, generator_pipeline)
    g = PT.load_golden(name)
    iids = [str(s) for s in g["iids"]]
    uids = [str(s) for s in g["uids"]]
    trainRDD = LocalRDD(_train_records(g))
    method = str(g["method"])
    sim_tool = BaselinerSim(method, int(g["num_atleast"]))
    simRDD = baseliner_calculate_sim_pipeline(None, sim_tool, trainRDD)
    ext_tool = ExtendSim(int(g["k"]))
    xsRDD = extender_pipeline(None, None, sim_tool, ext_tool, simRDD)

    # similarity records, exactly the reference's shape: ((iid1, iid2), (sim, mutu, frac, label))
    got = sorted(simRDD.collect(), key=lambda x: x[0])
    want_keys = [(iids[a], iids[b]) for a, b in zip(g["sim_i"], g["sim_j"])]
    order = sorted(range(len(want_keys)), key=lambda q: want_keys[q])
    assert [k for k, _ in got] == [want_keys[q] for q in order]
    gv = np.array([[v[0], v[1], v[2], v[3]] for _, v in got])
    o = np.array(order)
    assert np.array_equal(gv[:, 1], g["sim_mutu"][o]) and np.array_equal(gv[:, 2], g["sim_frac"][o])
    assert np.array_equal(gv[:, 3], g["sim_label"][o])
    np.testing.assert_allclose(gv[:, 0], g["sim_val"][o], rtol=PT.SIM_RTOL)

    # X-SIM rows: (iid_T, [(iid_S, xsim)*])
    rows = xsRDD.collect()
    flat = sorted((t, s, v) for t, lst in rows for s, v in lst)
    want = sorted((iids[a], iids[b], v) for a, b, v in zip(g["xs_start"], g["xs_end"], g["xs_val"]))
    assert [(a, b) for a, b, _ in flat] == [(a, b) for a, b, _ in want]
    np.testing.assert_allclose([v for _, _, v in flat], [v for _, _, v in want], rtol=PT.SIM_RTOL)
    assert all("T:" in t for t, _ in rows) and all("S:" in s for _, lst in rows for s, _ in lst)

    # AlterEgo profile, private flag on (= argmax under Python 3): identical records
    gen = Generator(1, 0.6, method, 0.1)
    alter = generator_pipeline(gen, trainRDD, xsRDD, True).collect()
    mine = sorted((u, i, float(r), calendar.timegm(t.timetuple())) for u, i, r, t in alter)
    ref = sorted((uids[u], iids[i], float(r), int(t)) for u, i, r, t in
                 zip(g["priv_ae_user"], g["priv_ae_item"], g["priv_ae_rating"], g["priv_ae_ts"]))
    assert mine == ref

    # non-private, the reference's np.random draws replayed as injected uniforms
    cnt = xsRDD.handle.result.count.cpu().numpy()
    u = np.zeros(len(cnt)); u[cnt >= 2] = g["nonpriv_uniforms"]
    gen2 = Generator(1, 0.6, method, 0.1, uniforms=u)
    pairs = gen2.cross_nonprivate_mapping(xsRDD).collect()
    st = xsRDD.handle.result.start_item.cpu().numpy()
    okrows = {iids[t] for t, c in zip(st, cnt) if c >= 2}
    got_pairs = [(t, s) for t, s in pairs if t in okrows]
    assert got_pairs == [(iids[t], iids[s]) for t, s in zip(g["nonpriv_rows"], g["nonpriv_chosen"])]
    assert gen2.single_candidate_rows == int((cnt == 1).sum())


def test_generator_accepts_foreign_xsim_rows_and_mapping_dict():
    """generator_pipeline fed plain (target, [(source, xsim)]) records (not our lazy RDD),
    and build_alterEgo fed a mapping dict, as the reference API allows."""
    from xmap_b200.rdd import LocalRDD
    from xmap_b200.core import Generator, generator_pipeline
    from xmap_b200.utils.assist import map_to_dict
    g = PT.load_golden("adj_low_overlap")
    iids = [str(s) for s in g["iids"]]
    trainRDD = LocalRDD(_train_records(g))
    rows = {}
    for a, b, v in zip(g["xs_start"], g["xs_end"], g["xs_val"]):
        rows.setdefault(iids[a], []).append((iids[b], float(v)))
    xs = LocalRDD(rows.items())
    gen = Generator(1, 0.6, "adjust_cosine", 0.1)
    alter = generator_pipeline(gen, trainRDD, xs, True).collect()
    assert len(alter) == len(g["priv_ae_user"])
    mapped = gen.cross_private_mapping(xs, session=trainRDD._xmap_session)
    alter2 = gen.build_alterEgo(trainRDD, map_to_dict(mapped)).collect()
    assert sorted(map(str, alter)) == sorted(map(str, alter2))
