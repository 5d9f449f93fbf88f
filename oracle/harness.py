"""Run the UNMODIFIED reference hot path in-process -- the primary oracle.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this; the product
package never does.  It needs /root/reference (absent on the GPU box), so the
GPU-side tests consume the golden vectors this module produced
(oracle/make_golden.py -> tests/golden/), not this module.

What it does: puts oracle/fake_pyspark (an in-process RDD shim) and
/root/reference/code on sys.path, imports the reference's own classes
(xmap/core/baselinerSim.py, extender.py, generator.py) and pipeline glue
(xmap/utils/assist.py:66-150), and drives them stage by stage.  Between stages
it re-orders the records it hands over so that the reference -- which is
order-nondeterministic under a real Spark shuffle -- follows the canonical
order of SURVEY.md App. A.6:

  1. users ascending, each user's ratings ascending by item id string;
  2. hence per-pair co-rating lists in ascending user order;
  3. sim pairs sorted by (item1, item2) before the extender, so every list the
     reference stable-sorts by |sim| is pre-ordered by neighbour id and ties
     resolve to the smaller id;
  4. X-SIM rows and their candidate lists sorted by id before the generator,
     so map_to_dict's "later wins" means "largest target id wins";
  5. "first" time in avoid_duplicate_ratings = smallest source item id.

The reference source is never edited or copied; parity is therefore pinned by
executing it, there being no tests or golden vectors in the reference itself
(SURVEY.md section 4).
"""
import os
import sys
import time
import warnings

import numpy as np

REFERENCE_CODE = os.environ.get("XMAP_REFERENCE_CODE", "/root/reference/code")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "fake_pyspark")

_ref = None


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_CODE, "xmap", "utils", "assist.py"))


def load_reference():
    """Import the reference modules; returns a dict of the names we drive."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_CODE)
    for name in list(sys.modules):
        if name == "xmap" or name.startswith("xmap.") or \
                name == "pyspark" or name.startswith("pyspark."):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not (f.startswith(REFERENCE_CODE) or f.startswith(_SHIM)):
                raise RuntimeError(
                    "module %s (%s) is already imported and would shadow the "
                    "reference; run the oracle in a fresh interpreter" % (name, f))
    sys.path.insert(0, REFERENCE_CODE)
    sys.path.insert(0, _SHIM)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # '\i' escapes, `is not ()`
            import pyspark
            from pyspark.sql import SQLContext
            from xmap.core.baselinerSim import BaselinerSim
            from xmap.core.extender import ExtendSim
            from xmap.core.generator import Generator
            from xmap.core.baselinerClean import BaselinerClean
            from xmap.core.baselinerSplit import BaselinerSplit
            from xmap.core.recommenderSim import RecommenderSim
            from xmap.core.recommenderPrivacy import RecommenderPrivacy
            from xmap.core.recommenderPrediction import RecommenderPrediction
            from xmap.utils import assist
    finally:
        sys.path.remove(REFERENCE_CODE)
        sys.path.remove(_SHIM)
    sc = pyspark.SparkContext()
    _ref = dict(pyspark=pyspark, sc=sc, sqlContext=SQLContext(sc),
                BaselinerSim=BaselinerSim, ExtendSim=ExtendSim,
                Generator=Generator, BaselinerClean=BaselinerClean,
                BaselinerSplit=BaselinerSplit, RecommenderSim=RecommenderSim,
                RecommenderPrivacy=RecommenderPrivacy, RecommenderPrediction=RecommenderPrediction, assist=assist)
    return _ref


def canonical_train(records):
    """Rule 1: users ascending by uid, ratings ascending by item id."""
    out = [(uid, sorted(lst, key=lambda x: x[0])) for uid, lst in records]
    out.sort(key=lambda x: x[0])
    return out


def run_sim(train_records, method="adjust_cosine", num_atleast=50):
    """assist.py:66-77 on canonical-ordered train records.

    Returns (tool, trainRDD, simRDD, seconds); simRDD records are
    ((iid1, iid2), (sim, mutu, frac_mutu, label)) sorted by (iid1, iid2).
    """
    R = load_reference()
    sc = R["sc"]
    tool = R["BaselinerSim"](method, num_atleast)
    trainRDD = sc.parallelize(canonical_train(train_records))
    t0 = time.perf_counter()
    simRDD = R["assist"].baseliner_calculate_sim_pipeline(sc, tool, trainRDD)
    dt = time.perf_counter() - t0
    simRDD = sc.parallelize(sorted(simRDD.collect(), key=lambda x: x[0]))   # rule 3
    return tool, trainRDD, simRDD, dt


def run_extend(tool, simRDD, top_k=10):
    """assist.py:80-102.  Returns (xsimRDD canonical-ordered, classified, seconds)."""
    R = load_reference()
    sc, sqlc = R["sc"], R["sqlContext"]
    ext = R["ExtendSim"](top_k)
    t0 = time.perf_counter()
    xs = R["assist"].extender_pipeline(sc, sqlc, tool, ext, simRDD)
    dt = time.perf_counter() - t0
    rows = [(t, sorted(lst, key=lambda x: x[0])) for t, lst in xs.collect()]
    rows.sort(key=lambda x: x[0])                                            # rule 4
    return sc.parallelize(rows), dt


def run_classify(tool, simRDD, top_k=10):
    """The selection half of extender_pipeline (assist.py:82-95), exposed so the
    BB set and the four neighbour lists can be frozen as golden vectors."""
    R = load_reference()
    sc, sqlc = R["sc"], R["sqlContext"]
    ext = R["ExtendSim"](top_k)
    df = tool.build_sim_DF(simRDD)
    df.registerTempTable("sim_table")
    bb = sqlc.sql("SELECT DISTINCT id1 FROM sim_table WHERE label = 1").map(
        lambda line: line.id1).collect()
    item_sim = tool.get_item_sim(simRDD)
    classified = ext.find_knn_items(item_sim, sc.broadcast(bb)).collect()
    return bb, classified


def run_generate(trainRDD, xsimRDD, private, method="adjust_cosine",
                 mapping_range=1, epsilon=0.6, rpo=0.1, np_seed=None):
    """assist.py:136-150.  Non-private mode draws from np.random (generator.py:110);
    seed it here so the draws can be replayed as injected uniforms."""
    R = load_reference()
    gen = R["Generator"](mapping_range, epsilon, method, rpo)
    if np_seed is not None:
        np.random.seed(np_seed)
    t0 = time.perf_counter()
    if private:
        mapped = gen.cross_private_mapping(xsimRDD)
    else:
        mapped = gen.cross_nonprivate_mapping(xsimRDD)
    pairs = [(t, str(s)) for t, s in mapped.collect()]
    mapping = R["assist"].map_to_dict(mapped)
    mapping = {str(k): v for k, v in mapping.items()}
    alter = gen.build_alterEgo(trainRDD, mapping).collect()
    dt = time.perf_counter() - t0
    return pairs, mapping, alter, dt


def run_recommender_sim(alter_records, method="cosine_item", num_atleast=50):
    """assist.py:153-175 (the next row of the scope table): item-item similarity and
    per-pair local sensitivity on the AlterEgo profile, flat (uid, iid, rating, time) records in the
    order given (the caller passes them sorted by (uid, iid, rating, time)).

    Returns (sim records ((iid1, iid2), [sim, local sensitivity]) sorted by key,
             item_info {iid: (average, norm2, count)}, seconds)."""
    R = load_reference()
    sc = R["sc"]
    tool = R["RecommenderSim"](method, num_atleast)
    profile = sc.parallelize(list(alter_records))
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                  # sqrt of a rounding-negative leave-one-out norm
        out = R["assist"].recommender_calculate_sim_pipeline(sc, tool, profile)
        sims = out[6].collect()
    dt = time.perf_counter() - t0
    item_info = dict(out[5].value)
    sims = sorted(((k, [float(v[0]), float(v[1])]) for k, v in sims), key=lambda x: x[0])
    return sims, item_info, dt


def run_recommender_predict(alter_records, test_records, mapping_range=10, alpha=0.03,
                            method="cosine_item", num_atleast=50, epsilon=0.6, rpo=0.1):
    """The rest of the demo after AlterEgo generation (twodomain_demo.py:119-134):
    recommender_calculate_sim_pipeline -> recommender_privacy_pipeline (non-private: top-`mapping_range`
    neighbours by |sim|, recommenderPrivacy.py:141-178) -> item-based prediction + MAE
    (recommenderPrediction.py:26-139), all by the unmodified reference.

    Canonical order (SURVEY.md App. A.6 rule 3): the similarity records enter the neighbour selection sorted
    by (id1, id2), so ties in |sim| resolve to the smaller neighbour id.

    Returns (neighbours {iid: [(nid, sim)*]}, predictions [(uid, [(iid, real, no_decay, decay) | ()]*)], mae string)."""
    R = load_reference()
    sc, assist = R["sc"], R["assist"]
    sim_tool = R["RecommenderSim"](method, num_atleast)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = assist.recommender_calculate_sim_pipeline(sc, sim_tool, sc.parallelize(list(alter_records)))
        sims = sorted(out[6].collect(), key=lambda x: x[0])
    user_dict_bd, item_dict_bd, user_info_bd, item_info_bd = out[2], out[3], out[4], out[5]
    priv_tool = R["RecommenderPrivacy"](mapping_range, epsilon, rpo)
    simpair = assist.recommender_privacy_pipeline(priv_tool, sc.parallelize(sims), False)
    neigh = simpair.collectAsMap()
    pred_tool = R["RecommenderPrediction"](alpha, method)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):          # calculate_mae prints rdd.take(1)
        testRDD = sc.parallelize(list(test_records))
        predicted = pred_tool.item_based_recommendation(testRDD, item_dict_bd, sc.broadcast(neigh), item_info_bd)
        preds = predicted.collect()
        mae = pred_tool.calculate_mae(sc.parallelize(preds))
    return neigh, preds, mae


def run_recommender_private(sim_records, mapping_range=10, epsilon=0.6, rpo=0.1, np_seed=7):
    """recommender_privacy_pipeline(policy_tool, alterEgo_sim, True) (assist.py:178-185) by the unmodified reference:
    private_neighbor_selection (recommenderPrivacy.py:70-139) + noise_perturbation (:152-171).  Under Python 3
    `np.count_nonzero(map(...))` (:81) counts the map OBJECT, so exactly one neighbour is drawn per item.

    The reference draws from the global np.random; every draw is logged here so that it can be replayed: the uniform
    behind the weighted pick (`np.random.rand(1)`, :118) and the uniform behind the Laplace noise (legacy
    `RandomState.laplace` consumes one double: the state is rewound and the double read).

    sim_records: ((iid1, iid2), [sim, local sensitivity]) sorted by (iid1, iid2).
    Returns [(iid, neighbour id, noisy sim, u_pick, u_noise)] in the reference's item order."""
    R = load_reference()
    sc, assist = R["sc"], R["assist"]
    tool = R["RecommenderPrivacy"](mapping_range, epsilon, rpo)
    log = {"pick": [], "noise": []}
    orig_rand, orig_lap = np.random.rand, np.random.laplace

    def rand(*a):
        v = orig_rand(*a)
        log["pick"].extend(np.atleast_1d(v).tolist())
        return v

    def laplace(loc, scale):
        st = np.random.get_state()
        v = orig_lap(loc, scale)
        after = np.random.get_state()
        np.random.set_state(st)
        log["noise"].append(float(np.random.random_sample()))
        np.random.set_state(after)
        return v
    np.random.seed(np_seed)
    np.random.rand, np.random.laplace = rand, laplace
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = assist.recommender_privacy_pipeline(tool, sc.parallelize(list(sim_records)), True).collect()
            res = [(iid, list(lst)) for iid, lst in res]
    finally:
        np.random.rand, np.random.laplace = orig_rand, orig_lap
    assert all(len(lst) == 1 for _, lst in res) and len(log["pick"]) == len(res) == len(log["noise"])
    return [(iid, lst[0][0], float(lst[0][1]), log["pick"][q], log["noise"][q]) for q, (iid, lst) in enumerate(res)]
