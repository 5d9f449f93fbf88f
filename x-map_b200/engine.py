"""Array-level driver of the CUDA hot path (one process = one GPU).

Everything here works on integer-encoded tensors resident in HBM and calls the
C ABI (include/xmap_b200.h) through ctypes; torch supplies device memory,
streams and a few index-building primitives (sort / cumsum / repeat_interleave)
around the kernels.  The reference-shaped facades in core/ and utils/assist.py
sit on top of this module.  No CPU fallback exists: without a GPU and the built
extension these functions raise.
"""
import math
from dataclasses import dataclass, field

import torch

from . import _native as N

RATER_GROUP = 128                # raters per unit of heavy-row work (matches sim.cu)
BIG_TABLE_BUDGET = 8 << 30       # bytes of HBM for heavy-row tables per batch
BIG_CAND_CAPACITY = 96 << 20     # candidate records (21 B each) per heavy-row batch


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _dev(t, dtype, device):
    return torch.as_tensor(t, dtype=dtype).to(device, non_blocking=True).contiguous()


@dataclass
class ItemMeta:
    """Per-item codes the reference derives from id strings (see encode.py)."""
    prefix_code: torch.Tensor     # int32  iid[:2]
    dom_code: torch.Tensor        # uint8  iid[-2:]
    contains: torch.Tensor        # uint8  bit d: label d in iid
    has_S: torch.Tensor           # bool   "S:" in iid
    has_T: torch.Tensor           # bool   "T:" in iid


@dataclass
class Layout:
    n_users: int
    n_items: int
    nnz: int
    csr_ptr: torch.Tensor
    csr_ent: torch.Tensor
    csr_src: torch.Tensor
    csc_ptr: torch.Tensor
    csc_ent: torch.Tensor
    user_mu: torch.Tensor
    item_stats: torch.Tensor      # [n_items, 4] = (avg, norm2, adj_norm2, count)
    row_work: torch.Tensor        # int64 [n_items]
    r_min: float
    r_max: float


def build_layout(user, item, rating, n_users, n_items, device="cuda"):
    """CSR/CSC + user/item statistics (xmap_build_layout, xmap_row_work)."""
    L = N.lib()
    user = _dev(user, torch.int32, device)
    item = _dev(item, torch.int32, device)
    rating = _dev(rating, torch.float32, device)
    nnz = int(user.numel())
    i32 = dict(dtype=torch.int32, device=device)
    f64 = dict(dtype=torch.float64, device=device)
    csr_ptr = torch.empty(n_users + 1, **i32)
    csc_ptr = torch.empty(n_items + 1, **i32)
    csr_ent = torch.empty(nnz, dtype=torch.int64, device=device)
    csc_ent = torch.empty(nnz, dtype=torch.int64, device=device)
    csr_src = torch.empty(nnz, **i32)
    user_mu = torch.empty(n_users, **f64)
    item_stats = torch.empty(n_items, 4, **f64)
    ws_bytes = L.xmap_layout_workspace_bytes(nnz, n_users, n_items)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    N.check(L.xmap_build_layout(N.ptr(user), N.ptr(item), N.ptr(rating), nnz, n_users, n_items,
                                N.ptr(csr_ptr), N.ptr(csr_ent), N.ptr(csr_src),
                                N.ptr(csc_ptr), N.ptr(csc_ent), N.ptr(user_mu), N.ptr(item_stats),
                                N.ptr(ws), ws_bytes, _stream_ptr()), "xmap_build_layout")
    row_work = torch.empty(n_items, dtype=torch.int64, device=device)
    N.check(L.xmap_row_work(N.ptr(csr_ptr), N.ptr(csc_ptr), N.ptr(csc_ent), n_items,
                            N.ptr(row_work), _stream_ptr()), "xmap_row_work")
    if nnz:
        r_min, r_max = float(rating.min()), float(rating.max())
    else:
        r_min = r_max = 0.0
    del ws
    return Layout(n_users, n_items, nnz, csr_ptr, csr_ent, csr_src, csc_ptr, csc_ent,
                  user_mu, item_stats, row_work, r_min, r_max)


def r2_bits_for(method, r_min, r_max):
    """ceil(log2(max |product|)) + 1 guard bit for the fixed-point inner product."""
    if method == "adjust_cosine":
        span = r_max - r_min            # |r - mean_u| <= span
    else:
        span = max(abs(r_min), abs(r_max))
    r2 = span * span
    if r2 <= 0:
        return 0
    return int(math.ceil(math.log2(r2))) + 1


@dataclass
class SimTables:
    """Output of the similarity + selection stage for the rows this rank owns."""
    k: int
    n_items: int
    row_flags: torch.Tensor       # uint8 [I]  bit0: bridge (BB) item
    row_npairs: torch.Tensor      # int32 [I]  co-rated neighbours (pre-filter)
    row_nkept: torch.Tensor       # int32 [I]  neighbours surviving the filter
    tab_idx: torch.Tensor         # int32 [I,2,k]
    tab_sim: torch.Tensor         # f64   [I,2,k]
    tab_mutu: torch.Tensor        # int32 [I,2,k]
    tab_n: torch.Tensor           # int32 [I,2,k]
    tab_len: torch.Tensor         # int32 [I,2]
    launches: int = 0
    stats: dict = field(default_factory=dict)


class SimEngine:
    """Similarity + top-k selection over a Layout (C ABI section 2)."""

    def __init__(self, layout, meta, method="adjust_cosine", num_atleast=50, k=10,
                 table_budget=BIG_TABLE_BUDGET):
        if method not in N.METHODS:
            raise ValueError("unknown similarity method %r" % (method,))
        if not (1 <= k <= N.KMAX):
            raise ValueError("top_k must be in [1, %d]" % N.KMAX)
        self.lay, self.meta = layout, meta
        self.method, self.num_atleast, self.k = method, int(num_atleast), int(k)
        self.device = layout.csr_ptr.device
        self.r2_bits = r2_bits_for(method, layout.r_min, layout.r_max)
        self.table_budget = table_budget
        self.launches = 0
        I = layout.n_items
        dev = self.device
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.row_flags = torch.zeros(I, dtype=torch.uint8, device=dev)
        self.row_npairs = torch.zeros(I, dtype=torch.int32, device=dev)
        self.row_nkept = torch.zeros(I, dtype=torch.int32, device=dev)
        self.tab_idx = torch.full((I, 2, k), -1, dtype=torch.int32, device=dev)
        self.tab_sim = torch.zeros((I, 2, k), dtype=torch.float64, device=dev)
        self.tab_mutu = torch.zeros((I, 2, k), dtype=torch.int32, device=dev)
        self.tab_n = torch.zeros((I, 2, k), dtype=torch.int32, device=dev)
        self.tab_len = torch.zeros((I, 2), dtype=torch.int32, device=dev)
        self._big_ws = None
        self._tier_ws = None
        self.profile = None        # dict kind -> [(start_event, end_event)] when enabled

    # -- argument block ----------------------------------------------------
    def _args(self, mode, bb_in=None, emit=None):
        lay, m = self.lay, self.meta
        a = N.SimArgs()
        a.csr_ptr, a.csr_ent = N.ptr(lay.csr_ptr), N.ptr(lay.csr_ent)
        a.csc_ptr, a.csc_ent = N.ptr(lay.csc_ptr), N.ptr(lay.csc_ent)
        a.user_mu, a.item_stats = N.ptr(lay.user_mu), N.ptr(lay.item_stats)
        a.prefix_code, a.dom_code, a.contains = N.ptr(m.prefix_code), N.ptr(m.dom_code), N.ptr(m.contains)
        a.bb_in = N.ptr(bb_in)
        a.row_work = N.ptr(lay.row_work)
        a.n_items, a.method = lay.n_items, N.METHODS[self.method]
        a.num_atleast, a.k, a.r2_bits, a.mode = self.num_atleast, self.k, self.r2_bits, mode
        a.row_flags, a.row_npairs, a.row_nkept = N.ptr(self.row_flags), N.ptr(self.row_npairs), N.ptr(self.row_nkept)
        a.tab_idx, a.tab_sim = N.ptr(self.tab_idx), N.ptr(self.tab_sim)
        a.tab_mutu, a.tab_n, a.tab_len = N.ptr(self.tab_mutu), N.ptr(self.tab_n), N.ptr(self.tab_len)
        if emit is not None:
            a.emit_ptr, a.emit_j, a.emit_sim = N.ptr(emit["ptr"]), N.ptr(emit["j"]), N.ptr(emit["sim"])
            a.emit_mutu, a.emit_n, a.emit_cursor = N.ptr(emit["mutu"]), N.ptr(emit["n"]), N.ptr(emit["cursor"])
        a.error_flag = N.ptr(self.error_flag)
        return a

    # -- planning ----------------------------------------------------------
    def plan(self, rows=None):
        """Split rows by cost: four warp-per-row tiers (by table size) and the heavy tier.
        Each tier is ordered by descending work so long rows start first."""
        w = self.lay.row_work
        if rows is None:
            rows = torch.arange(self.lay.n_items, dtype=torch.int32, device=self.device)
        else:
            rows = rows.to(self.device, dtype=torch.int32)
        wr = w[rows.long()]
        order = torch.argsort(wr, descending=True, stable=True)
        rows, wr = rows[order], wr[order]
        tiers, lo = [], 0
        for hi in N.TIER_MAXWORK:
            tiers.append(rows[(wr > lo) & (wr <= hi)].contiguous())
            lo = hi
        big = rows[wr > lo].contiguous()
        return tiers, big

    def _big_batches(self, big):
        """Cut the heavy rows (already sorted by descending work) into batches bounded by the
        table budget and by the candidate-record capacity.  One host sync per plan."""
        I = self.lay.n_items
        per_row = I * 20 + 4
        b_max = int(max(1, self.table_budget // per_row))
        w = self.lay.row_work[big.long()]
        cap = torch.where(w >= 2 * I, torch.full_like(w, I), torch.clamp(w, max=I)).cpu().tolist()   # dense rows: I records
        batches, lo, acc = [], 0, 0
        for q, c in enumerate(cap):
            if q > lo and (q - lo >= b_max or acc + c > BIG_CAND_CAPACITY):
                batches.append((lo, q, acc)); lo, acc = q, 0
            acc += c
        if len(cap) > lo:
            batches.append((lo, len(cap), acc))
        return batches, b_max

    def _big_workspace(self, n_rows, capacity):
        I = self.lay.n_items
        L = N.lib()
        ws = self._big_ws
        need = L.xmap_sim_big_scratch_bytes(n_rows, capacity)
        if ws is None or ws["B"] < n_rows or ws["scratch"].numel() < need:
            dev = self.device
            B = max(n_rows, ws["B"] if ws else 0)
            need = max(need, ws["scratch"].numel() if ws else 0)
            ws = None
            self._big_ws = None
            ws = dict(B=B,
                      table=torch.zeros(B * I * 2, dtype=torch.int64, device=dev),
                      touched=torch.empty(B * I, dtype=torch.int32, device=dev),
                      touched_n=torch.zeros(B, dtype=torch.int32, device=dev),
                      counter=torch.zeros(1, dtype=torch.int32, device=dev),
                      scratch=torch.empty(need, dtype=torch.uint8, device=dev))
            self._big_ws = ws
        return ws

    def enable_profile(self):
        """Record a CUDA event pair around every kernel group (read with profile_ms())."""
        self.profile = {}

    def _timed(self, kind, fn):
        if self.profile is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.profile.setdefault(kind, []).append((a, b))
        return r

    def profile_ms(self):
        """{kind: (launch groups, total ms)}; synchronises."""
        torch.cuda.synchronize()
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.profile or {}).items()}
        self.profile = {}
        return out

    def _run_rows(self, args, tiers, big):
        L = N.lib()
        st = _stream_ptr()
        for tier, rows in enumerate(tiers):
            if not rows.numel():
                continue
            ws = None
            need = L.xmap_sim_rows_workspace_bytes(tier)
            if need:
                if self._tier_ws is None or self._tier_ws.numel() < need:
                    self._tier_ws = None
                    self._tier_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                ws = self._tier_ws
            self._timed("warp_tier%d" % tier, lambda: N.check(
                L.xmap_sim_rows(args, N.ptr(rows), rows.numel(), tier, N.ptr(ws), need, st),
                "xmap_sim_rows[%d]" % tier))
            self.launches += 1
        if big.numel():
            lay = self.lay
            batches, _ = self._big_batches(big)
            ws = self._big_workspace(max(hi - lo for lo, hi, _ in batches), max(c for _, _, c in batches))
            for lo_b, hi_b, capacity in batches:
                rows = big[lo_b:hi_b].contiguous()
                rl = rows.long()
                c = (lay.csc_ptr[rl + 1] - lay.csc_ptr[rl]).long()
                grp_off = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=self.device)
                grp_off[1:] = torch.cumsum((c + RATER_GROUP - 1) // RATER_GROUP, 0)
                ws["counter"].zero_()
                self._timed("big_accumulate", lambda: N.check(L.xmap_sim_big_accumulate(
                    args, N.ptr(rows), rows.numel(), N.ptr(grp_off), N.ptr(ws["table"]),
                    N.ptr(ws["touched"]), N.ptr(ws["touched_n"]), N.ptr(ws["counter"]), st),
                    "xmap_sim_big_accumulate"))
                self._timed("big_epilogue", lambda: N.check(L.xmap_sim_big_finalize(
                    args, N.ptr(rows), rows.numel(), N.ptr(ws["table"]), N.ptr(ws["touched"]),
                    N.ptr(ws["touched_n"]), capacity, N.ptr(ws["scratch"]), ws["scratch"].numel(), st),
                    "xmap_sim_big_finalize"))
                self.launches += 4

    def _check_error(self):
        if int(self.error_flag.item()) != 0:
            raise N.NativeError("similarity kernel error %d (1: hash overflow, 4: candidate capacity)" % int(self.error_flag.item()))

    # -- passes ------------------------------------------------------------
    def pass1(self, rows=None):
        """Similarity rows + BB flags + (BB_BB, BB_NB) / NB_NN tables for `rows`."""
        tiers, big = self.plan(rows)
        self._run_rows(self._args(0), tiers, big)
        return dict(tiers=[int(t.numel()) for t in tiers], big=int(big.numel()))

    def pass2(self, bb_all, rows=None):
        """NB_BB tables for the non-bridge rows among `rows` (needs every item's BB flag)."""
        if rows is None:
            rows = torch.arange(self.lay.n_items, dtype=torch.int32, device=self.device)
        rl = rows.long()
        nb = rows[(self.row_flags[rl] == 0) & (self.row_nkept[rl] > 0)]
        tiers, big = self.plan(nb)
        self._run_rows(self._args(1, bb_in=bb_all.to(torch.uint8).contiguous()), tiers, big)
        return dict(tiers=[int(t.numel()) for t in tiers], big=int(big.numel()))

    def run(self, rows=None):
        """Single-GPU convenience: pass 1, pass 2, error check."""
        s1 = self.pass1(rows)
        s2 = self.pass2(self.row_flags, rows)
        self._check_error()
        return self.tables(dict(pass1=s1, pass2=s2))

    def tables(self, stats=None):
        return SimTables(self.k, self.lay.n_items, self.row_flags, self.row_npairs, self.row_nkept,
                         self.tab_idx, self.tab_sim, self.tab_mutu, self.tab_n, self.tab_len,
                         self.launches, stats or {})

    def emit_pairs(self, rows=None):
        """Materialise every kept directed pair (the return value of
        baseliner_calculate_sim_pipeline, assist.py:66-77), sorted by (i, j).
        Needs pass1 to have filled row_nkept."""
        dev = self.device
        I = self.lay.n_items
        if rows is None:
            rows = torch.arange(I, dtype=torch.int32, device=dev)
        nk = torch.zeros(I, dtype=torch.int64, device=dev)
        nk[rows.long()] = self.row_nkept[rows.long()].long()
        ptr = torch.zeros(I + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(nk, 0)
        total = int(ptr[-1].item())
        emit = dict(ptr=ptr,
                    j=torch.empty(total, dtype=torch.int32, device=dev),
                    sim=torch.empty(total, dtype=torch.float64, device=dev),
                    mutu=torch.empty(total, dtype=torch.int32, device=dev),
                    n=torch.empty(total, dtype=torch.int32, device=dev),
                    cursor=torch.zeros(I, dtype=torch.int32, device=dev))
        tiers, big = self.plan(rows)
        self._run_rows(self._args(2, emit=emit), tiers, big)
        self._check_error()
        i = torch.repeat_interleave(torch.arange(I, device=dev), nk)
        key = i * I + emit["j"].long()
        order = torch.argsort(key)
        i, j = i[order], emit["j"][order].long()
        sim, mutu, n = emit["sim"][order], emit["mutu"][order], emit["n"][order]
        cnt = self.lay.item_stats[:, 3]
        frac = mutu.double() / (cnt[i] + cnt[j] - n.double())
        label = (self.meta.prefix_code[i] != self.meta.prefix_code[j]).to(torch.int32)
        return dict(i=i, j=j, sim=sim, mutu=mutu, n=n, frac=frac, label=label)
