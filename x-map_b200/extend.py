"""X-SIM extension driver (C ABI section 3).

Turns the fixed-width neighbour tables of the similarity stage into the three
index structures the kernel walks, exactly following the reference's wiring:

  attach(b)    NB items n that list bridge item b in NB_BB(n), with NB_NN(n)
               (extender.py:48-59, reduceByKey :171-173)
  src pairs    (t, s): s a bridge item with "S:" in its id and a non-empty attach
               list, t any key of knn_BB[s] with "T:" in its id   (extender.py:61-70)
  tgt pairs    (t, s): t a bridge item with "T:" and attach, s a key of knn_BB[t]
               with "S:"                                           (extender.py:72-81)
  joint        src pairs that are also tgt pairs (the inner join, extender.py:178)
  right(s)     s ; (s,n) ; (s,n,x) for x in NB_NN(n)               (extender.py:134-138)
  left(t)      (n,t) ; (x,n,t)                                     (extender.py:160-167)

Edge values are symmetric in this implementation (sim(a,b) == sim(b,a)
bitwise), so the reference's 4-way lookup precedence (extender.py:100-112)
cannot change a value and every edge is read from the row that lists it.

All index building is torch (sort / cumsum / repeat_interleave) on whatever
device the tables live on; the path evaluation itself is the CUDA kernel.
"""
from dataclasses import dataclass

import torch

from . import _native as N

XSIM_HASH_BUDGET = 48 << 30     # bytes of per-unit hash tables per launch
XSIM_UNIT_COMBOS = 1 << 20      # a start with more paths than this is cut into leg slices


def _excl_cumsum(x):
    return torch.cumsum(x, 0) - x


def _segment_ids(lengths):
    return torch.repeat_interleave(torch.arange(lengths.numel(), device=lengths.device), lengths)


@dataclass
class XsimPlan:
    n_items: int
    start_item: torch.Tensor
    leg_ptr: torch.Tensor
    leg_t: torch.Tensor
    leg_joint_only: torch.Tensor
    leg_vals: tuple              # e1, m1, f1, e2, m2, f2
    par_ptr: torch.Tensor
    par_s: torch.Tensor
    par_joint: torch.Tensor
    par_vals: tuple              # e, m, f
    rs_ptr: torch.Tensor
    rs_end: torch.Tensor
    rs_vals: tuple               # e1, m1, f1, e2, m2, f2
    ub: torch.Tensor             # per start: combos (= reference paths) it will evaluate
    ub_leg: torch.Tensor         # per leg
    n_src: int
    n_joint: int
    t_items: torch.Tensor        # bridge targets, indexed by leg_t / par_ptr
    s_items: torch.Tensor        # bridge sources, indexed by par_s / rs_ptr


def build_plan(tabs, item_count, has_S, has_T):
    """tabs: engine.SimTables (global, i.e. holding every item's lists)."""
    dev = tabs.tab_idx.device
    I, k = tabs.n_items, tabs.k
    i64 = torch.int64
    bb = (tabs.row_flags & 1).bool()
    tlen = tabs.tab_len.long()
    valid_nb = (~bb) & (tlen[:, 0] > 0)
    r = torch.arange(k, device=dev)

    def flat(slot, rowmask):
        """All (row, nbr, e, m, f) entries of `slot` for rows in rowmask, row-major."""
        m = rowmask[:, None] & (r[None, :] < tlen[:, slot, None])
        rows, pos = torch.nonzero(m, as_tuple=True)
        nbr = tabs.tab_idx[rows, slot, pos].long()
        sim = tabs.tab_sim[rows, slot, pos]
        mutu = tabs.tab_mutu[rows, slot, pos].double()
        n = tabs.tab_n[rows, slot, pos].double()
        frac = 1.0 * mutu / (item_count[rows] + item_count[nbr] - n)     # baselinerSim.py:171-173
        return rows, nbr, sim * mutu, mutu, frac

    # ---- attach edges (b <- n), ordered by (b, n) -------------------------
    n_row, b_nbr, a_e, a_m, a_f = flat(0, valid_nb)
    o = torch.argsort(b_nbr * I + n_row)
    att_b, att_n = b_nbr[o], n_row[o]
    att_e, att_m, att_f = a_e[o], a_m[o], a_f[o]
    a_cnt = torch.bincount(att_b, minlength=I)
    nn_len = torch.where(valid_nb, tlen[:, 1], torch.zeros_like(tlen[:, 1]))
    # NB_NN entries, row-major; nn_ptr[n] indexes them
    nn_row, nn_x, nn_e, nn_m, nn_f = flat(1, valid_nb)
    nn_ptr = torch.zeros(I + 1, dtype=i64, device=dev)
    nn_ptr[1:] = torch.cumsum(nn_len, 0)

    # ---- bridge pairs -----------------------------------------------------
    kb_row0, kb_nbr0, kb_e0, kb_m0, kb_f0 = flat(0, bb)
    kb_row1, kb_nbr1, kb_e1, kb_m1, kb_f1 = flat(1, bb)
    kb_row = torch.cat([kb_row0, kb_row1]); kb_nbr = torch.cat([kb_nbr0, kb_nbr1])
    kb_e = torch.cat([kb_e0, kb_e1]); kb_m = torch.cat([kb_m0, kb_m1]); kb_f = torch.cat([kb_f0, kb_f1])
    has_att = a_cnt > 0
    src_m = has_S[kb_row] & has_att[kb_row] & has_T[kb_nbr]
    src_t, src_s = kb_nbr[src_m], kb_row[src_m]
    src_e, src_m_, src_f = kb_e[src_m], kb_m[src_m], kb_f[src_m]
    tgt_m = has_T[kb_row] & has_att[kb_row] & has_S[kb_nbr]
    tgt_key = torch.unique(kb_row[tgt_m] * I + kb_nbr[tgt_m])
    src_key = src_t * I + src_s
    o = torch.argsort(src_key)
    src_key, src_t, src_s = src_key[o], src_t[o], src_s[o]
    src_e, src_m_, src_f = src_e[o], src_m_[o], src_f[o]
    if tgt_key.numel() and src_key.numel():
        p = torch.searchsorted(tgt_key, src_key).clamp(max=tgt_key.numel() - 1)
        joint = tgt_key[p] == src_key
    else:
        joint = torch.zeros_like(src_key, dtype=torch.bool)

    # ---- right segments per bridge source s -------------------------------
    s_items = torch.unique(src_s)
    s_index = torch.full((I,), -1, dtype=i64, device=dev)
    s_index[s_items] = torch.arange(s_items.numel(), device=dev)
    am = s_index[att_b] >= 0
    ra_s, ra_n = s_index[att_b[am]], att_n[am]
    ra_e, ra_m, ra_f = att_e[am], att_m[am], att_f[am]
    blk = 1 + nn_len[ra_n]
    R = torch.ones(s_items.numel(), dtype=i64, device=dev)
    R.index_add_(0, ra_s, blk)
    rs_ptr = torch.zeros(s_items.numel() + 1, dtype=i64, device=dev)
    rs_ptr[1:] = torch.cumsum(R, 0)
    total_rs = int(rs_ptr[-1].item()) if s_items.numel() else 0
    rs_end = torch.empty(total_rs, dtype=torch.int32, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    rs_e1 = torch.zeros(total_rs, **f64); rs_m1 = torch.zeros(total_rs, **f64); rs_f1 = torch.ones(total_rs, **f64)
    rs_e2 = torch.zeros(total_rs, **f64); rs_m2 = torch.zeros(total_rs, **f64); rs_f2 = torch.ones(total_rs, **f64)
    if total_rs:
        rs_end[rs_ptr[:-1]] = s_items.to(torch.int32)                    # the bare (t, s) path ends at s
        # block start of attach edge a inside its s: 1 + blocks before it in the same s
        cs = _excl_cumsum(blk)
        first_of_s = torch.zeros(s_items.numel(), dtype=i64, device=dev)
        cnt_s = torch.bincount(ra_s, minlength=s_items.numel())
        first_edge = _excl_cumsum(cnt_s)
        nz = cnt_s > 0
        first_of_s[nz] = cs[first_edge[nz]]
        blk_start = rs_ptr[ra_s] + 1 + (cs - first_of_s[ra_s])
        a_of = _segment_ids(blk)
        q = torch.arange(a_of.numel(), device=dev) - _excl_cumsum(blk)[a_of]
        pos = blk_start[a_of] + q
        is_n = q == 0
        nnp = (nn_ptr[ra_n[a_of]] + q - 1).clamp(min=0)
        rs_end[pos] = torch.where(is_n, ra_n[a_of], nn_x[nnp] if nn_x.numel() else ra_n[a_of]).to(torch.int32)
        rs_e1[pos] = ra_e[a_of]; rs_m1[pos] = ra_m[a_of]; rs_f1[pos] = ra_f[a_of]
        if nn_x.numel():
            one = torch.ones_like(nn_f[nnp]); zero = torch.zeros_like(nn_e[nnp])
            rs_e2[pos] = torch.where(is_n, zero, nn_e[nnp])
            rs_m2[pos] = torch.where(is_n, zero, nn_m[nnp])
            rs_f2[pos] = torch.where(is_n, one, nn_f[nnp])

    # ---- partners per bridge target t -------------------------------------
    t_items = torch.unique(src_t)
    t_index = torch.full((I,), -1, dtype=i64, device=dev)
    t_index[t_items] = torch.arange(t_items.numel(), device=dev)
    par_cnt = torch.bincount(t_index[src_t], minlength=t_items.numel()) if src_t.numel() else \
        torch.zeros(0, dtype=i64, device=dev)
    par_ptr = torch.zeros(t_items.numel() + 1, dtype=i64, device=dev)
    par_ptr[1:] = torch.cumsum(par_cnt, 0)
    par_s = s_index[src_s].to(torch.int32)               # src pairs are already sorted by (t, s)
    Rs = R[s_index[src_s]] if src_s.numel() else torch.zeros(0, dtype=i64, device=dev)
    rt_all = torch.zeros(t_items.numel(), dtype=i64, device=dev)
    rt_joint = torch.zeros(t_items.numel(), dtype=i64, device=dev)
    if src_t.numel():
        rt_all.index_add_(0, t_index[src_t], Rs)
        rt_joint.index_add_(0, t_index[src_t], Rs * joint.long())

    # ---- legs per start ---------------------------------------------------
    zt = torch.zeros(t_items.numel(), **f64); ot = torch.ones(t_items.numel(), **f64)
    # type 0: the bridge target itself starts the non-joint paths (extender.py:180)
    L_start = [t_items]; L_t = [torch.arange(t_items.numel(), device=dev)]
    L_jo = [torch.zeros(t_items.numel(), dtype=torch.uint8, device=dev)]
    L_type = [torch.zeros(t_items.numel(), dtype=i64, device=dev)]
    L_vals = [[zt], [zt], [ot], [zt], [zt], [ot]]
    # types 1, 2: attach edges of targets that have at least one joint partner
    tj = torch.zeros(I, dtype=torch.bool, device=dev)
    if src_t.numel():
        tj[src_t[joint]] = True
    lm = tj[att_b]
    la_t, la_n = att_b[lm], att_n[lm]
    la_e, la_m, la_f = att_e[lm], att_m[lm], att_f[lm]
    na = la_t.numel()
    if na:
        L_start.append(la_n); L_t.append(t_index[la_t])
        L_jo.append(torch.ones(na, dtype=torch.uint8, device=dev))
        L_type.append(torch.ones(na, dtype=i64, device=dev))
        for lst, v in zip(L_vals, (la_e, la_m, la_f, torch.zeros(na, **f64), torch.zeros(na, **f64),
                                   torch.ones(na, **f64))):
            lst.append(v)
        ln = nn_len[la_n]
        a_of = _segment_ids(ln)
        if a_of.numel():
            q = torch.arange(a_of.numel(), device=dev) - _excl_cumsum(ln)[a_of]
            nnp = nn_ptr[la_n[a_of]] + q
            L_start.append(nn_x[nnp]); L_t.append(t_index[la_t[a_of]])
            L_jo.append(torch.ones(a_of.numel(), dtype=torch.uint8, device=dev))
            L_type.append(torch.full((a_of.numel(),), 2, dtype=i64, device=dev))
            # path (x, n, t): edge (x, n) first, then (n, t)
            for lst, v in zip(L_vals, (nn_e[nnp], nn_m[nnp], nn_f[nnp], la_e[a_of], la_m[a_of], la_f[a_of])):
                lst.append(v)
    lg_start = torch.cat(L_start); lg_t = torch.cat(L_t); lg_jo = torch.cat(L_jo); lg_type = torch.cat(L_type)
    lg_vals = [torch.cat(v) for v in L_vals]
    # stable sort by (start, type): construction order inside a type is already (t, n, q)
    o = torch.argsort(lg_start * 4 + lg_type, stable=True)
    lg_start, lg_t, lg_jo = lg_start[o], lg_t[o], lg_jo[o]
    lg_vals = [v[o] for v in lg_vals]
    start_item, leg_cnt = torch.unique_consecutive(lg_start, return_counts=True)
    leg_ptr = torch.zeros(start_item.numel() + 1, dtype=i64, device=dev)
    leg_ptr[1:] = torch.cumsum(leg_cnt, 0)
    ub_leg = torch.where(lg_jo.bool(), rt_joint[lg_t], rt_all[lg_t]) if lg_t.numel() else \
        torch.zeros(0, dtype=i64, device=dev)
    ub = torch.zeros(start_item.numel(), dtype=i64, device=dev)
    if lg_t.numel():
        ub.index_add_(0, _segment_ids(leg_cnt), ub_leg)
    return XsimPlan(
        n_items=I, start_item=start_item.to(torch.int32), leg_ptr=leg_ptr, leg_t=lg_t.to(torch.int32),
        leg_joint_only=lg_jo.contiguous(), leg_vals=tuple(v.contiguous() for v in lg_vals),
        par_ptr=par_ptr, par_s=par_s.contiguous(), par_joint=joint.to(torch.uint8).contiguous(),
        par_vals=(src_e.contiguous(), src_m_.contiguous(), src_f.contiguous()),
        rs_ptr=rs_ptr, rs_end=rs_end, rs_vals=(rs_e1, rs_m1, rs_f1, rs_e2, rs_m2, rs_f2),
        ub=ub, ub_leg=ub_leg, n_src=int(src_t.numel()), n_joint=int(joint.sum().item()) if joint.numel() else 0,
        t_items=t_items, s_items=s_items)


def _pow2_at_least(x):
    """Smallest power of two >= x, in exact integer arithmetic (float pow / log2 on the
    device are not exact and a table size of 2^k - 1 breaks the probe mask)."""
    x = torch.clamp(x.long(), min=1)
    p = torch.ones_like(x)
    for _ in range(62):
        p = torch.where(p < x, p * 2, p)
    return p


@dataclass
class XsimResult:
    start_item: torch.Tensor     # int32 [n_starts]
    count: torch.Tensor          # int32 distinct ends per start
    combos: torch.Tensor         # int64 paths evaluated per start
    top_end: torch.Tensor        # int32 [n_starts, top_m], |xsim| desc, ties to smaller end
    top_xsim: torch.Tensor
    top_len: torch.Tensor
    launches: int


class XsimEngine:
    """Runs the extension kernels over an XsimPlan.

    Work units: a start whose path count exceeds `unit_combos` is cut, at leg granularity, into
    slices that run as independent warps with private hash tables; the slices are then merged by
    a fixed binary tree (slice g absorbs g + 2^r in round r), so every cell's summation order is
    a function of the path structure only -- results do not depend on launch batching."""

    def __init__(self, plan, top_m=10, hash_budget=XSIM_HASH_BUDGET, unit_combos=XSIM_UNIT_COMBOS):
        if not (1 <= top_m <= N.KMAX):
            raise ValueError("top_m must be in [1, %d]" % N.KMAX)
        self.plan, self.top_m, self.hash_budget = plan, int(top_m), hash_budget
        p = plan
        dev = self.device = p.start_item.device
        self.launches = 0
        n = p.start_item.numel()
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        i64 = torch.int64
        lc = p.leg_ptr[1:] - p.leg_ptr[:-1]
        leg_start = _segment_ids(lc)
        # slice id of a leg = (paths of the start before this leg) // unit_combos
        before = torch.cumsum(p.ub_leg, 0) - p.ub_leg
        start_base = torch.zeros(n, dtype=i64, device=dev)
        if n:
            start_base = before[p.leg_ptr[:-1].clamp(max=max(before.numel() - 1, 0))] if before.numel() else start_base
        gid = (before - start_base[leg_start]) // int(unit_combos) if before.numel() else before
        ukey = leg_start * (1 << 24) + gid
        _, ucnt = torch.unique_consecutive(ukey, return_counts=True)
        self.unit_leg_hi = torch.cumsum(ucnt, 0)
        self.unit_leg_lo = self.unit_leg_hi - ucnt
        self.unit_start = leg_start[self.unit_leg_lo] if ucnt.numel() else leg_start
        n_units = int(ucnt.numel())
        unit_ub = torch.zeros(n_units, dtype=i64, device=dev)
        if n_units:
            unit_ub.index_add_(0, _segment_ids(ucnt), p.ub_leg)
        self.G = torch.bincount(self.unit_start, minlength=n) if n_units else torch.zeros(n, dtype=i64, device=dev)
        self.u0 = torch.cumsum(self.G, 0) - self.G
        g = torch.arange(n_units, device=dev) - self.u0[self.unit_start] if n_units else unit_ub
        self.unit_g = g
        # table of slice g must hold the union of its merge subtree [g, g + lowbit(g)) (all for g == 0)
        lowbit = torch.where(g > 0, g & (-g), self.G[self.unit_start] if n_units else g)
        hi = torch.minimum(g + lowbit, self.G[self.unit_start]) if n_units else g
        P = torch.zeros(n_units + 1, dtype=i64, device=dev)
        P[1:] = torch.cumsum(unit_ub, 0)
        base_u = self.u0[self.unit_start] if n_units else g
        sub_ub = P[base_u + hi] - P[base_u + g] if n_units else unit_ub
        end_cap = int(torch.unique(p.rs_end).numel()) if p.rs_end.numel() else 1
        # distinct ends <= min(paths, reachable ends): a table of >= 1.25x that bound never fills up
        # (worst-case load 0.8; typical load ~0.25 because several paths share an end)
        self.hsize = torch.clamp((5 * torch.clamp(sub_ub, max=end_cap) + 3) // 4, min=32)
        self.start_bytes = torch.zeros(n, dtype=i64, device=dev)
        if n_units:
            self.start_bytes.index_add_(0, self.unit_start, self.hsize * 32)
        # the two edges of a right segment folded into one (N, D, C) triple: 28 instead of 60 bytes per
        # path (the sums are reassociated by at most one rounding; D is an exact integer either way)
        e1, m1, f1, e2, m2, f2 = p.rs_vals
        # ... and every list ordered by end (stable), so that right segments of one source that end at the
        # same item sit in the same 32-path step and are combined before they reach the table
        rl = p.rs_ptr[1:] - p.rs_ptr[:-1]
        seg = _segment_ids(rl)
        perm = torch.argsort(seg * int(p.n_items) + p.rs_end.long(), stable=True) if seg.numel() else seg
        self.rs_end = p.rs_end[perm].contiguous()
        self.rs_ndc = ((e1 + e2)[perm].contiguous(), (m1 + m2)[perm].contiguous(), (f1 * f2)[perm].contiguous())
        self.order = torch.argsort(p.ub, descending=True, stable=True)
        self.n_units = n_units
        self._cells = None
        self.epoch = 0

    # ------------------------------------------------------------------
    def _workspace(self, n_cells):
        """Persistent cell workspace (32 B cells), zeroed once; launches are told apart by epoch."""
        need = n_cells * 4
        if self._cells is None or self._cells.numel() < need:
            self._cells = None
            self._cells = torch.zeros(need, dtype=torch.int64, device=self.device)
        return self._cells

    def _batches(self, rank=0, world=1):
        """Starts of this rank (every world-th start of the descending-work order, so the ranks'
        loads match), cut into launches bounded by the hash budget."""
        order = self.order[rank::world] if world > 1 else self.order
        if order.numel() == 0:
            return
        cum = torch.cumsum(self.start_bytes[order], 0).cpu()
        lo, n = 0, order.numel()
        while lo < n:
            base = int(cum[lo - 1]) if lo else 0
            hi = int(torch.searchsorted(cum, torch.tensor(base + self.hash_budget)))
            hi = min(max(hi, lo + 1), n)
            yield order[lo:hi]
            lo = hi

    def _launch(self, sel, mode, out, emit=None):
        """Run accumulate -> merge rounds -> finalize over the starts `sel` (plan indices)."""
        L = N.lib()
        p, dev = self.plan, self.device
        ns = int(sel.numel())
        G = self.G[sel]
        units = torch.repeat_interleave(self.u0[sel], G) + \
            (torch.arange(int(G.sum().item()), device=dev) - torch.repeat_interleave(torch.cumsum(G, 0) - G, G))
        nu = int(units.numel())
        first_unit = torch.cumsum(G, 0) - G                       # batch-local index of each start's slice 0
        hs = self.hsize[units]
        hoff = torch.cumsum(hs, 0) - hs
        total = int(hs.sum().item())
        cells = self._workspace(total)
        # merge tree
        g = self.unit_g[units]
        Gu = torch.repeat_interleave(G, G)
        loc = torch.arange(nu, device=dev)
        rounds, pd, ps = [0], [], []
        r, gmax = 0, int(G.max().item()) if ns else 1
        while (1 << r) < gmax:
            m = ((g % (1 << (r + 1))) == 0) & (g + (1 << r) < Gu)
            d = loc[m]
            pd.append(d); ps.append(d + (1 << r))
            rounds.append(rounds[-1] + int(d.numel()))
            r += 1
        pair_dst = torch.cat(pd).to(torch.int32) if pd else torch.zeros(0, dtype=torch.int32, device=dev)
        pair_src = torch.cat(ps).to(torch.int32) if ps else torch.zeros(0, dtype=torch.int32, device=dev)
        import ctypes as C
        round_ptr = (C.c_int32 * len(rounds))(*rounds)
        a = N.XsimArgs()
        keep = [round_ptr]

        def P(t):
            t = t.contiguous()
            keep.append(t)
            return N.ptr(t)
        a.n_starts, a.n_units = ns, nu
        a.start_item = P(p.start_item[sel]); a.start_unit = P(first_unit.to(torch.int32))
        a.unit_leg_lo, a.unit_leg_hi = P(self.unit_leg_lo[units]), P(self.unit_leg_hi[units])
        ucomb = torch.zeros(nu, dtype=torch.int64, device=dev)
        a.unit_combos = P(ucomb)
        a.leg_t, a.leg_joint_only = P(p.leg_t), P(p.leg_joint_only)
        (a.leg_e1, a.leg_m1, a.leg_f1, a.leg_e2, a.leg_m2, a.leg_f2) = [P(v) for v in p.leg_vals]
        a.par_ptr = P(p.par_ptr); a.par_s = P(p.par_s); a.par_joint = P(p.par_joint)
        a.par_e, a.par_m, a.par_f = [P(v) for v in p.par_vals]
        a.rs_ptr = P(p.rs_ptr); a.rs_end = P(self.rs_end)
        a.rs_n, a.rs_d, a.rs_c = [P(v) for v in self.rs_ndc]
        a.hash_off = P(hoff); a.hash_size = P(hs.to(torch.int32))
        a.hash_cells = N.ptr(cells)
        self.epoch += 1
        a.epoch = self.epoch
        a.n_rounds = len(rounds) - 1
        a.round_ptr_h = C.cast(round_ptr, C.c_void_p)
        a.pair_dst, a.pair_src = P(pair_dst), P(pair_src)
        a.top_m, a.mode = self.top_m, mode
        cnt = torch.zeros(ns, dtype=torch.int32, device=dev)
        te = torch.full((ns, self.top_m), -1, dtype=torch.int32, device=dev)
        tx = torch.zeros((ns, self.top_m), dtype=torch.float64, device=dev)
        tl = torch.zeros(ns, dtype=torch.int32, device=dev)
        a.out_count = P(cnt)
        a.top_end, a.top_xsim, a.top_len = P(te), P(tx), P(tl)
        if emit is not None:
            a.emit_ptr, a.emit_end, a.emit_xsim = P(emit[0]), P(emit[1]), P(emit[2])
        a.error_flag = N.ptr(self.error_flag)
        N.check(L.xmap_xsim_extend(a, torch.cuda.current_stream().cuda_stream), "xmap_xsim_extend")
        self.launches += 2 + a.n_rounds
        if mode == 0:
            comb = torch.zeros(ns, dtype=torch.int64, device=dev)
            comb.index_add_(0, torch.repeat_interleave(torch.arange(ns, device=dev), G), ucomb)
            out["count"][sel] = cnt; out["combos"][sel] = comb
            out["top_end"][sel] = te; out["top_xsim"][sel] = tx; out["top_len"][sel] = tl
        torch.cuda.current_stream().synchronize()      # the batch's temporaries die here
        if int(self.error_flag.item()):
            raise N.NativeError("X-SIM kernel error %d (2: hash overflow, 3: bad table size)"
                                % int(self.error_flag.item()))

    def run(self, rank=0, world=1):
        """count + top-m for every start of the plan (world > 1: only this rank's starts are filled in;
        multi.allreduce_xsim assembles the full result on every rank)."""
        p, dev, n = self.plan, self.device, self.plan.start_item.numel()
        out = dict(count=torch.zeros(n, dtype=torch.int32, device=dev),
                   combos=torch.zeros(n, dtype=torch.int64, device=dev),
                   top_end=torch.full((n, self.top_m), -1, dtype=torch.int32, device=dev),
                   top_xsim=torch.zeros((n, self.top_m), dtype=torch.float64, device=dev),
                   top_len=torch.zeros(n, dtype=torch.int32, device=dev))
        for sel in self._batches(rank, world):
            self._launch(sel, 0, out)
        return XsimResult(p.start_item, out["count"], out["combos"], out["top_end"], out["top_xsim"],
                          out["top_len"], self.launches)

    def emit(self, res):
        """Every (start, end, xsim), sorted by (start, end): the materialised return
        value of extender_pipeline (assist.py:80-102)."""
        p, dev = self.plan, self.device
        n = p.start_item.numel()
        ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(res.count.long(), 0)
        total = int(ptr[-1].item()) if n else 0
        e_end = torch.empty(total, dtype=torch.int32, device=dev)
        e_x = torch.empty(total, dtype=torch.float64, device=dev)
        for sel in self._batches():
            self._launch(sel, 2, None, emit=(ptr[:-1][sel].contiguous(), e_end, e_x))
        start = torch.repeat_interleave(p.start_item.long(), res.count.long())
        o = torch.argsort(start * p.n_items + e_end.long())
        return start[o], e_end[o].long(), e_x[o]
