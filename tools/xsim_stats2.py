"""X-SIM work structure at a bench workload: (leg, partner) pair counts, paths and distinct ends per start,
and the lookup overhead an end-tiled on-chip accumulator would pay for a given table capacity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xmap_b200 import engine as E, extend as X
from xmap_b200.engine import to_device_meta
wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
dev = torch.device("cuda"); meta = to_device_meta(wl["meta"], dev)
lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
tabs = eng.run()
plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
xe = X.XsimEngine(plan, 10)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = xe.run()
torch.cuda.synchronize(); print("xsim run s", time.perf_counter() - t0)
ub = plan.ub.double(); cnt = res.count.double()
n = ub.numel()
par_cnt = (plan.par_ptr[1:] - plan.par_ptr[:-1])
jcnt = torch.zeros_like(par_cnt); tpar = torch.repeat_interleave(torch.arange(par_cnt.numel(), device=dev), par_cnt)
jcnt.index_add_(0, tpar, plan.par_joint.long())
leg_lp = torch.where(plan.leg_joint_only.bool(), jcnt[plan.leg_t.long()], par_cnt[plan.leg_t.long()])
lc = plan.leg_ptr[1:] - plan.leg_ptr[:-1]
leg_start = torch.repeat_interleave(torch.arange(n, device=dev), lc)
LP = torch.zeros(n, dtype=torch.int64, device=dev); LP.index_add_(0, leg_start, leg_lp)
LPd = LP.double()
print("starts %d legs %d partners %d rsegs %d n_s %d n_t %d" % (n, plan.leg_t.numel(), plan.par_s.numel(), plan.rs_end.numel(),
      plan.rs_ptr.numel() - 1, plan.par_ptr.numel() - 1))
print("paths %.4g cells %.4g LP %.4g  paths/LP %.2f paths/cell %.2f" % (float(ub.sum()), float(cnt.sum()), float(LPd.sum()),
      float(ub.sum() / LPd.sum()), float(ub.sum() / cnt.sum())))
q = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0], dtype=torch.float64, device=dev)
def qs(x): return [float(v) for v in torch.quantile(x[:min(x.numel(), 10_000_000)], q)]
print("quantiles 10/25/50/75/90/99/100")
print(" paths/start", qs(ub)); print(" ends/start", qs(cnt)); print(" LP/start", qs(LPd)); print(" legs/start", qs(lc.double()))
print(" paths/LP per start", qs(ub / LPd.clamp(min=1))); print(" paths/end per start", qs(ub / cnt.clamp(min=1)))
for cap in (700, 1000, 1400, 2800, 5600, 9000, 12000):
    T = torch.ceil(cnt / cap).clamp(min=1)
    look = float((T * LPd).sum())
    legl = float((T * lc.double()).sum())
    print("cap %5d: units %.4g  lookups(LP*T) %.4g = %.2f per path ; leg visits %.4g ; starts with T=1 hold %.3f of paths; T<=8: %.3f" % (
        cap, float(T.sum()), look, look / float(ub.sum()), legl, float(ub[T <= 1].sum() / ub.sum()), float(ub[T <= 8].sum() / ub.sum())))
# path-weighted ends histogram
for thr in (1e3, 3e3, 1e4, 3e4, 1e5, 3e5):
    m = cnt <= thr
    print("ends<=%g: starts %.3f paths %.3f cells %.3f LP %.3f" % (thr, float(m.double().mean()), float(ub[m].sum() / ub.sum()),
          float(cnt[m].sum() / cnt.sum()), float(LPd[m].sum() / LPd.sum())))
rl = (plan.rs_ptr[1:] - plan.rs_ptr[:-1]).double()
print("R_s quantiles", qs(rl), "mean", float(rl.mean()))
# walk-weighted R_s: how long is the list of a (leg, partner) visit on average / quantiles by visits
legs_into_t = torch.bincount(plan.leg_t.long(), minlength=par_cnt.numel()).double()
w = legs_into_t[tpar]          # visits of each partner entry (ignoring joint-only restriction)
rs_of_par = rl[plan.par_s.long()]
o = torch.argsort(rs_of_par); cw = torch.cumsum(w[o], 0) / w.sum()
for f in (0.1, 0.25, 0.5, 0.75, 0.9, 0.99):
    i = int(torch.searchsorted(cw, torch.tensor(f, dtype=torch.float64, device=dev)))
    print(" visit-weighted R_s q%.2f = %d" % (f, int(rs_of_par[o][min(i, o.numel() - 1)])))
# distinct ends overall / S items
print("distinct ends", int(torch.unique(plan.rs_end).numel()), "items", plan.n_items)
print("top_len hist", torch.bincount(res.top_len.long(), minlength=11).tolist())
