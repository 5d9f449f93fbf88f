"""ExtendSim facade -- reference class xmap/core/extender.py:8-217."""
import numpy as np

from ..rdd import LazyRDD
from .. import extend as X


class XsimHandle(object):
    def __init__(self, session, plan, engine, result):
        self.session, self.plan, self.engine, self.result = session, plan, engine, result

    def rows(self):
        """(iid_T, [(iid_S, xsim)*])* in ascending (start, end) index order."""
        s, e, v = self.engine.emit(self.result)
        s, e, v = s.cpu().numpy(), e.cpu().numpy(), v.cpu().numpy()
        iids = self.session.enc.iids
        cut = np.flatnonzero(np.diff(s)) + 1
        for lo, hi in zip(np.r_[0, cut], np.r_[cut, len(s)]):
            if hi > lo:
                yield (str(iids[s[lo]]), [(str(iids[e[q]]), v[q]) for q in range(lo, hi)])


class ExtendSim(object):
    def __init__(self, top_k):
        self.top_k = top_k

    def find_knn_items(self, rdd, BB_items_bd=None):
        """iid, (BB_BB, BB_NB), None | iid, None, (NB_BB, NB_NN) -- extender.py:16-44.
        `rdd` must descend from BaselinerSim.calculate_item2item_sim (device handle)."""
        h = _sim_handle(rdd)
        sess = h.session
        t = sess.similarity(h.method, h.num_atleast, self.top_k)

        def build():
            iids = sess.enc.iids
            cnt = sess.layout.item_stats[:, 3].cpu().numpy()
            fl, ln = t.row_flags.cpu().numpy(), t.tab_len.cpu().numpy()
            ix, sm = t.tab_idx.cpu().numpy(), t.tab_sim.cpu().numpy()
            mu, nn = t.tab_mutu.cpu().numpy(), t.tab_n.cpu().numpy()

            def lst(i, slot):
                return [(str(iids[ix[i, slot, r]]), sm[i, slot, r], float(mu[i, slot, r]),
                         float(mu[i, slot, r]) / (cnt[i] + cnt[ix[i, slot, r]] - nn[i, slot, r]))
                        for r in range(ln[i, slot])]
            for i in range(len(iids)):
                if fl[i] & 1:
                    yield str(iids[i]), (lst(i, 0), lst(i, 1)), None
                elif ln[i, 0] > 0:
                    yield str(iids[i]), None, (lst(i, 0), lst(i, 1))
        return LazyRDD(build, handle=h)

    def extend(self, sim_handle, top_m=10):
        """sim_extend + get_final_extension (extender.py:46-217) on the device."""
        sess = sim_handle.session
        tabs = sess.similarity(sim_handle.method, sim_handle.num_atleast, self.top_k)
        plan = X.build_plan(tabs, sess.layout.item_stats[:, 3].contiguous(), sess.meta.has_S, sess.meta.has_T)
        eng = X.XsimEngine(plan, top_m)
        res = eng.run()
        h = XsimHandle(sess, plan, eng, res)
        sess.xsim = h
        return LazyRDD(lambda: h.rows(), handle=h)


def _sim_handle(rdd):
    h = getattr(rdd, "handle", None)
    if h is None or not hasattr(h, "session"):
        raise TypeError("expected the RDD returned by baseliner_calculate_sim_pipeline (it carries the "
                        "device-resident similarity state); got %r" % type(rdd).__name__)
    return h
