"""BaselinerSim facade -- same constructor and method names as the reference class
(xmap/core/baselinerSim.py:11-244); the work happens in the CUDA similarity stage."""
from ..rdd import LazyRDD, LocalRDD
from ..session import session_of


class SimHandle(object):
    """Device-resident result of calculate_item2item_sim, carried on the lazy RDD."""

    def __init__(self, session, method, num_atleast):
        self.session, self.method, self.num_atleast = session, method, num_atleast

    def pairs(self, k=10):
        s = self.session
        s.similarity(self.method, self.num_atleast, k)
        p = s.sim_engine.emit_pairs()
        iids = s.enc.iids
        i, j = p["i"].cpu().numpy(), p["j"].cpu().numpy()
        sim, mutu = p["sim"].cpu().numpy(), p["mutu"].cpu().numpy()
        frac, label = p["frac"].cpu().numpy(), p["label"].cpu().numpy()
        for a in range(len(i)):
            yield ((str(iids[i[a]]), str(iids[j[a]])),
                   (sim[a], float(mutu[a]), float(frac[a]), int(label[a])))


class BaselinerSim(object):
    def __init__(self, method, num_atleast):
        self.method = method
        self.num_atleast = num_atleast

    def get_universal_user_info(self, dataRDD):
        """(uid, (average, norm2))* -- baselinerSim.py:17-38."""
        return LocalRDD(session_of(dataRDD).user_info())

    def get_universal_item_info(self, dataRDD, user_info=None):
        """(iid, (average, norm2, adjusted norm2, count))* -- baselinerSim.py:40-82."""
        return LocalRDD(session_of(dataRDD).item_info())

    def calculate_item2item_sim(self, dataRDD, item_info=None, user_info=None):
        """((iid1, iid2), (sim, mutu, frac_mutu, label))*, both directions -- baselinerSim.py:187-216.
        Like the reference, an unknown method yields None (:213-216)."""
        if self.method not in ("cosine", "adjust_cosine"):
            return None
        h = SimHandle(session_of(dataRDD), self.method, self.num_atleast)
        return LazyRDD(lambda: h.pairs(), handle=h)

    def get_item_sim(self, dataRDD):
        """(iid1, [(iid2, sim, mutu, frac_mutu)*])* -- baselinerSim.py:218-233."""
        def build():
            rows = {}
            for (a, b), (sim, mutu, frac, _label) in dataRDD.collect():
                rows.setdefault(a, []).append((b, sim, mutu, frac))
            return rows.items()
        return LazyRDD(build, handle=getattr(dataRDD, "handle", None))

    def build_sim_DF(self, sim_pairsRDD):
        """Rows id1, id2, sim, mutu, frac_mutu, label -- baselinerSim.py:235-244 (a list of dicts
        stands in for the Spark DataFrame when no SQLContext is in play)."""
        return [dict(id1=k[0], id2=k[1], sim=float(v[0]), mutu=v[1], frac_mutu=float(v[2]), label=v[3])
                for k, v in sim_pairsRDD.collect()]
