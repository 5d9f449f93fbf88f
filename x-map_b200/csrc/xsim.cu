// X-SIM bridge extension: every (left segment, bridge pair, right segment)
// combination is one path of the reference's enumeration (extender.py:124-169);
// its similarity s_p = sum(sim*mutu)/sum(mutu) and certainty c_p = prod(frac)
// (extender.py:83-89) are accumulated per (start, end) into
// xsim = sum(s_p c_p) / sum(c_p) (extender.py:198-201) without ever storing a path.
//
// Design: the accumulator of a start never leaves the SM, and no two warps ever share a cell.
// One WARP owns one unit = (start item, a run of passes) at a time and a private hash table in shared
// memory; a pass covers a range of the hashed end axis pi(y) = y * 0x9E3779B1 (every right-segment list
// is stored sorted by pi, and a per-source pointer table gives the sub-range of a pass without a search),
// chosen so that the distinct ends of the pass fit the table.  Inside a pass the warp walks the
// (leg, partner) pairs of its start 32 at a time: each lane loads one pair descriptor {list, folded (N, D, C)
// of leg + bridge edge} and resolves the list's sub-range of the pass, a shuffle scan numbers the products
// of the batch and they are dealt to the lanes in chunks of 32 whatever the sub-range lengths.  A lane
// loads one 28-byte right segment (the next chunk's loads are issued before the current chunk is
// processed) and evaluates the path.  Lanes that hit the same end: adjacent ones (the lists are ordered
// by pi(end), so a run of equal ends is the rule in a fused list) are summed by a segmented shuffle scan,
// the heads of the runs are grouped (MATCH.ANY) and the group's lowest lane adds the run totals in lane
// (= path) order to the cell: plain shared-memory loads and stores, no atomics on values, no block
// barrier.  The summation order of a (start, end) cell is a function of the path structure and the pass
// plan only, so results are bit-identical from run to run and for any number of GPUs (every rank builds
// the same plan).
//
// A pass whose table overflows is split in two on the device (the tile range halves) and redone; the top-m
// of a unit is kept in shared memory across its passes and a small kernel merges the units of a start.
// Units are fetched from a global counter in descending-work order (results do not depend on which warp
// runs a unit).
//
// Bound: issue rate of the per-path instruction stream and the 28 B right-segment read per path from
// L2/HBM; accumulator traffic is shared memory only.
#include "common.cuh"

namespace xmap {

constexpr int XT_MAX = 640;                   // threads per CTA (a multiple of 32, chosen at launch)
constexpr unsigned XGOLD2 = 0x85EBCA6Bu;      // second multiplicative hash of the end: owner warp and home cell
                                              // (the first one, pi(y) = y * 0x9E3779B1, orders the lists: host side)
constexpr int XSTACK = 16;                    // pending halves of split passes (depth <= gb + 1)
constexpr int XSV = 64;                       // candidates of a top-m merge: survivors (<= 32) + running best (<= 32)

// Per-warp staging (shared memory).  `sv` doubles as the chunk staging of the in-order combine
// (accumulate phase) and as the survivor list of the top-m selection (finalize phase).
struct __align__(16) WarpScratch {
    union {
        struct { double c_num[32], c_den[32]; } acc;
        struct { unsigned long long key[XSV]; double x[XSV]; int end[XSV]; } sv;
    } u;
    unsigned long long best_key[XMAP_KMAX];    // running top-m of the unit over its finished passes
    double best_x[XMAP_KMAX];
    int best_end[XMAP_KMAX];
    int stack_g0[XSTACK], stack_g1[XSTACK];
};

struct Fetched {
    double N, D, C, rn, rd, rc;
    int y;
    bool valid;
};

// products [32 c, 32 c + 32) of the batch: the lane's descriptor by a shuffle search over the inclusive
// product counts (empty descriptors are skipped by construction), then the right-segment loads
__device__ __forceinline__ Fetched fetch_chunk(const xmap_xsim_args &a, int c, int total, int incl, int len,
                                               long long base, double dN, double dD, double dC, int lane) {
    Fetched f;
    const int p = (c << 5) + lane;
    int l = 0;                                              // smallest lane with incl > p
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, incl, l + step - 1);
        if (v <= p) l += step;
    }
    const int excl = __shfl_sync(0xffffffffu, incl - len, l);
    const long long b = __shfl_sync(0xffffffffu, base, l);
    f.N = __shfl_sync(0xffffffffu, dN, l);
    f.D = __shfl_sync(0xffffffffu, dD, l);
    f.C = __shfl_sync(0xffffffffu, dC, l);
    f.valid = p < total;
    f.y = 0; f.rn = f.rd = f.rc = 0.0;
    if (f.valid) {
        const long long r = b + (long long)(p - excl);
        f.y = __ldg(a.rs_end + r);
        f.rn = __ldg(a.rs_n + r); f.rd = __ldg(a.rs_d + r); f.rc = __ldg(a.rs_c + r);
    }
    return f;
}

struct PairDesc {
    double N, D, C;
    long long base;
    int len;
};

// pair q of the unit: its descriptor {list, folded (N, D, C)} (one coalesced load per lane) and the sub-range of the
// list that falls into the pass (empty beyond q_hi)
__device__ __forceinline__ PairDesc load_pair(const xmap_xsim_args &a, long long q, long long q_hi, bool whole,
                                              int g0, int g1, int G1) {
    PairDesc d;
    d.N = d.D = d.C = 0.0; d.base = 0; d.len = 0;
    if (q < q_hi) {
        const int s = __ldg(a.pd_s + q);
        d.N = __ldg(a.pd_n + q); d.D = __ldg(a.pd_d + q); d.C = __ldg(a.pd_c + q);
        const long long rb = __ldg(a.rs_ptr + s);
        if (whole) { d.base = rb; d.len = (int)(__ldg(a.rs_ptr + s + 1) - rb); }
        else {
            const int32_t *tp = a.tile_ptr + (size_t)s * G1;
            const int b0 = __ldg(tp + g0), b1 = __ldg(tp + g1);
            d.base = rb + b0; d.len = b1 - b0;
        }
    }
    return d;
}

// One warp = one unit at a time (fetched from a global counter in descending-work order); the warps of a
// CTA are independent: no block barrier anywhere.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) xsim_warp_kernel(xmap_xsim_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nw = blockDim.x >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CS = 1 << a.cells_lg;                        // cells of a warp's shared-memory table
    WarpScratch &W = reinterpret_cast<WarpScratch *>(smem_raw)[warp];
    double2 *s_vals = reinterpret_cast<double2 *>(smem_raw + (size_t)nw * sizeof(WarpScratch)) + (size_t)warp * CS;
    int *s_keys = reinterpret_cast<int *>(smem_raw + (size_t)nw * sizeof(WarpScratch) + (size_t)nw * CS * sizeof(double2)) +
                  (size_t)warp * CS;
    // units whose ends do not fit the shared-memory table use this warp's slot of the global workspace
    // (L2-resident while the unit runs): 2^gcells_lg cells, vals first
    unsigned char *gslot = a.gws ? reinterpret_cast<unsigned char *>(a.gws) +
                                       ((size_t)blockIdx.x * nw + warp) * ((size_t)20 << a.gcells_lg) : nullptr;
    const int G = 1 << a.gb, G1 = G + 1;
    const int M = a.top_m;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (;;) {
        int uq = 0;
        if (lane == 0) uq = atomicAdd(a.unit_counter, 1);
        uq = __shfl_sync(0xffffffffu, uq, 0);
        if (uq >= a.n_units) break;
        const int u = a.unit_order ? __ldg(a.unit_order + uq) : uq;
        const long long t_unit = a.unit_cycles ? clock64() : 0ll;
        const long long leg_lo = a.unit_leg_lo[u], leg_hi = a.unit_leg_hi[u];
        const long long q_lo = a.lp_ptr[leg_lo], q_hi = a.lp_ptr[leg_hi];
        const int ug0 = a.unit_g0[u], ug1 = a.unit_g1[u], unpass = a.unit_npass[u];
        long long combos = 0;
        int unit_count = 0, best_len = 0, stack_n = 0, next_pass = 0, status = 0, emitted = 0;
        const int clg = a.unit_clg ? min(a.unit_clg[u], gslot ? a.gcells_lg : a.cells_lg) : a.cells_lg;
        const bool in_smem = clg <= a.cells_lg;
        const int C = 1 << clg;
        double2 *vals = in_smem ? s_vals : reinterpret_cast<double2 *>(gslot);
        int *keys = in_smem ? s_keys : reinterpret_cast<int *>(gslot + ((size_t)16 << a.gcells_lg));

        for (;;) {
            // ---- next pass: a split half if one is pending, else the unit's next own pass (all warp-uniform) ----
            int g0, g1;
            if (stack_n > 0) { --stack_n; g0 = W.stack_g0[stack_n]; g1 = W.stack_g1[stack_n]; }
            else if (next_pass < unpass) {
                const long long w = ug1 - ug0;             // the unit's tile range cut into unpass equal passes
                g0 = ug0 + (int)(w * next_pass / unpass);
                g1 = ug0 + (int)(w * (next_pass + 1) / unpass);
                ++next_pass;
            } else break;
            const bool whole = g0 == 0 && g1 == G;
            for (int c = lane; c < C; c += 32) keys[c] = 0;
            __syncwarp();

            // =================== accumulate ======================================================
            long long q0 = q_lo, pass_combos = 0;
            bool ovf = false;
            int n_ins = 0;                                 // occupied cells of the table
            PairDesc nxd = load_pair(a, q0 + lane, q_hi, whole, g0, g1, G1);
            while (q0 < q_hi && !ovf) {
                const int nq = (int)min(32ll, q_hi - q0);
                const PairDesc pd = nxd;
                q0 += nq;
                nxd = load_pair(a, q0 + lane, q_hi, whole, g0, g1, G1);   // the next 32 pairs resolve while these are walked
                const int len = pd.len;
                const long long base = pd.base;
                const double dN = pd.N, dD = pd.D, dC = pd.C;
                int incl = len;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += t;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                pass_combos += total;

                const int nchunk = (total + 31) >> 5;
                Fetched nxt = fetch_chunk(a, 0, total, incl, len, base, dN, dD, dC, lane);
                for (int c = 0; c < nchunk; ++c) {
                    const Fetched cur = nxt;
                    nxt = fetch_chunk(a, c + 1, total, incl, len, base, dN, dD, dC, lane);
                    double num = 0.0, den = 0.0;
                    if (cur.valid) {
                        const double Nn = __dadd_rn(cur.N, cur.rn);
                        const double Dd = __dadd_rn(cur.D, cur.rd);
                        den = __dmul_rn(cur.C, cur.rc);
                        const double sp = (Dd != 0.0) ? __ddiv_rn(Nn, Dd) : 0.0;      // extender.py:88-89
                        num = __dmul_rn(sp, den);
                    }
                    // Lanes that hit the same end.  The lists are ordered by pi(end), so equal ends sit in ADJACENT lanes
                    // (a fused list holds an end once per way of reaching it): a run of equal ends is summed by a
                    // segmented suffix scan over the lanes (a fixed tree), and only the heads of the runs go on;
                    // heads with equal ends (a chunk that spans several lists) are grouped by MATCH.ANY below.
                    const int y = cur.valid ? cur.y : (-1 - lane);
                    const int y_prev = __shfl_up_sync(0xffffffffu, y, 1);
                    const bool head = lane == 0 || y != y_prev;
                    const unsigned heads = __ballot_sync(0xffffffffu, head);
                    if (heads != 0xffffffffu) {
                        const unsigned later = lane == 31 ? 0u : (heads >> (lane + 1));
                        const int run_end = later ? lane + __ffs(later) - 1 : 31;
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const double vn = __shfl_down_sync(0xffffffffu, num, off);
                            const double vd = __shfl_down_sync(0xffffffffu, den, off);
                            if (lane + off <= run_end) { num = __dadd_rn(num, vn); den = __dadd_rn(den, vd); }
                        }
                    }
                    W.u.acc.c_num[lane] = num; W.u.acc.c_den[lane] = den;
                    const unsigned grp = __match_any_sync(0xffffffffu, head ? y : (-1 - lane));
                    __syncwarp();
                    bool lane_ovf = false, inserted = false;
                    if (cur.valid && head && (__ffs(grp) - 1) == lane) {
                        // find-or-insert (only this warp touches the table; the CAS settles lanes racing for one empty cell)
                        const int key = y + 1;
                        int pos = (int)(((unsigned)y * XGOLD2) >> (32 - clg));
                        int slot = -1;
                        bool isnew = false;
                        for (int probes = 0; probes < C; ++probes) {
                            const int kcur = *(volatile int *)(keys + pos);
                            if (kcur == key) { slot = pos; break; }
                            if (kcur == 0) {
                                const int old = atomicCAS(keys + pos, 0, key);
                                if (old == 0) { slot = pos; isnew = true; break; }
                                if (old == key) { slot = pos; break; }
                            }
                            pos = (pos + 1) & (C - 1);
                        }
                        inserted = isnew;
                        if (slot < 0) lane_ovf = true;
                        else {
                            // the runs of this end, one by one in lane (= path) order
                            double an = 0.0, ad = 0.0;
                            if (!isnew) { const double2 v = vals[slot]; an = v.x; ad = v.y; }
                            unsigned rem = grp;
                            while (rem) {
                                const int b = __ffs(rem) - 1;
                                an = __dadd_rn(an, W.u.acc.c_num[b]); ad = __dadd_rn(ad, W.u.acc.c_den[b]);
                                rem &= rem - 1u;
                            }
                            vals[slot] = make_double2(an, ad);
                        }
                    }
                    __syncwarp();                                  // staging reads done before the next chunk's writes
                    n_ins += __popc(__ballot_sync(0xffffffffu, inserted));
                    ovf = __any_sync(0xffffffffu, lane_ovf) || n_ins > C - (C >> 3);   // linear probing degrades past 7/8
                    if (ovf) break;
                }
            }
            if (ovf) {
                // the pass does not fit: halve its tile range and redo both halves (nothing of it was published)
                if (g1 - g0 < 2 || stack_n + 2 > XSTACK) status = 1;
                else {
                    const int mid = (g0 + g1) >> 1;
                    if (lane == 0) {
                        W.stack_g0[stack_n] = mid; W.stack_g1[stack_n] = g1;
                        W.stack_g0[stack_n + 1] = g0; W.stack_g1[stack_n + 1] = mid;
                    }
                    stack_n += 2;
                }
                __syncwarp();
                continue;
            }
            combos += pass_combos;

            // =================== finalize the pass ==============================================
            // xsim of every cell (kept in vals[c].x), count, lane-local best, optional emit
            int mycnt = 0;
            unsigned long long lk = 0ull; int le = 0x7FFFFFFF;
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                const int kk = keys[c];
                const bool occ = kk != 0;
                double x = 0.0;
                if (occ) {
                    const double2 v = vals[c];
                    x = __ddiv_rn(v.x, v.y);                   // extender.py:198-201
                    vals[c].x = x;
                    ++mycnt;
                    const unsigned long long ak = abs_key(x);
                    if (mycnt == 1 || better(ak, kk - 1, lk, le)) { lk = ak; le = kk - 1; }
                }
                if (a.emit_ptr) {
                    const unsigned mo = __ballot_sync(0xffffffffu, occ);
                    if (occ) {
                        const long long o = a.emit_ptr[u] + emitted + __popc(mo & lt_mask);
                        a.emit_end[o] = kk - 1; a.emit_xsim[o] = x;
                    }
                    emitted += __popc(mo);
                }
            }
            const unsigned have = __ballot_sync(0xffffffffu, mycnt > 0);
            int cnt_pass = mycnt;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) cnt_pass += __shfl_xor_sync(0xffffffffu, cnt_pass, off);
            unit_count += cnt_pass;
            const int newlen = min(M, cnt_pass + best_len);
            __syncwarp();
            // threshold (tk, te): a candidate worse than it cannot enter the list
            unsigned long long tk = 0ull; int te = 0x7FFFFFFF;
            bool fast = M <= 32;
            if (fast && cnt_pass > XSV / 2 && __popc(have) >= M) {
                // the M-th best of the lane-local maxima bounds the M-th best cell from below
                const int want = M;
                int rank = 0;
                for (int t = 0; t < 32; ++t) {
                    const unsigned long long k2 = __shfl_sync(0xffffffffu, lk, t);
                    const int e2 = __shfl_sync(0xffffffffu, le, t);
                    if (((have >> t) & 1u) && better(k2, e2, lk, le)) ++rank;
                }
                const unsigned sel = __ballot_sync(0xffffffffu, mycnt > 0 && rank == want - 1);
                const int sl = __ffs(sel) - 1;
                tk = __shfl_sync(0xffffffffu, lk, sl); te = __shfl_sync(0xffffffffu, le, sl);
            }
            if (best_len == M) {
                const unsigned long long bk = W.best_key[M - 1]; const int be = W.best_end[M - 1];
                if (better(bk, be, tk, te)) { tk = bk; te = be; }
            }
            // survivors: cells not worse than the threshold
            int nsurv = 0;
            for (int c0 = 0; c0 < C && fast; c0 += 32) {
                const int c = c0 + lane;
                const int kk = keys[c];
                bool take = false;
                unsigned long long ak = 0ull; double x = 0.0;
                if (kk != 0) { x = vals[c].x; ak = abs_key(x); take = !better(tk, te, ak, kk - 1); }
                const unsigned mt = __ballot_sync(0xffffffffu, take);
                if (nsurv + __popc(mt) > XSV / 2) { fast = false; break; }
                if (take) { const int q = nsurv + __popc(mt & lt_mask); W.u.sv.key[q] = ak; W.u.sv.x[q] = x; W.u.sv.end[q] = kk - 1; }
                nsurv += __popc(mt);
            }
            if (fast) {
                // candidates = survivors + the running best; rank by counting, winners rewrite the list
                if (lane < best_len) { W.u.sv.key[nsurv + lane] = W.best_key[lane]; W.u.sv.x[nsurv + lane] = W.best_x[lane]; W.u.sv.end[nsurv + lane] = W.best_end[lane]; }
                __syncwarp();
                const int ncand = nsurv + best_len;                 // <= XSV
                for (int q = lane; q < ncand; q += 32) {
                    const unsigned long long mk = W.u.sv.key[q]; const int me = W.u.sv.end[q];
                    int rank = 0;
                    for (int t = 0; t < ncand; ++t) rank += better(W.u.sv.key[t], W.u.sv.end[t], mk, me) ? 1 : 0;
                    if (rank < newlen) { W.best_key[rank] = mk; W.best_x[rank] = W.u.sv.x[q]; W.best_end[rank] = me; }
                }
            } else {
                // slow path (top_m > 32, or one value repeated very often): newlen rounds of a warp arg-best over
                // every cell and the running best; each round takes the best candidate strictly after the last
                __syncwarp();
                unsigned long long last_k = ~0ull; int last_t = -1;
                for (int r = 0; r < newlen; ++r) {
                    unsigned long long bk = 0ull; int bt = 0x7FFFFFFF, bp = -1;
                    for (int c = lane; c < C + best_len; c += 32) {
                        unsigned long long ak; int ee;
                        if (c < C) { const int kk = keys[c]; if (kk == 0) continue; ak = abs_key(vals[c].x); ee = kk - 1; }
                        else { ak = W.best_key[c - C]; ee = W.best_end[c - C]; }
                        if (r > 0 && !better(last_k, last_t, ak, ee)) continue;
                        if (bp < 0 || better(ak, ee, bk, bt)) { bk = ak; bt = ee; bp = c; }
                    }
                    warp_argbest(bk, bt, bp);
                    if (lane == 0 && bp >= 0) {
                        // the new list grows in the key array of the candidates; positions < r are final
                        W.u.sv.key[r] = bk; W.u.sv.end[r] = bt;
                        W.u.sv.x[r] = bp < C ? vals[bp].x : W.best_x[bp - C];
                    }
                    last_k = bk; last_t = bt;
                }
                __syncwarp();
                for (int q = lane; q < newlen; q += 32) { W.best_key[q] = W.u.sv.key[q]; W.best_x[q] = W.u.sv.x[q]; W.best_end[q] = W.u.sv.end[q]; }
            }
            best_len = newlen;
            __syncwarp();
        }

        // ---- publish the unit --------------------------------------------------------------------
        if (lane == 0) {
            a.unit_count[u] = unit_count;
            a.unit_combos[u] = combos;
            a.unit_top_len[u] = best_len;
            if (status) atomicExch(a.error_flag, 2);
            if (a.unit_cycles) a.unit_cycles[u] = clock64() - t_unit;
        }
        for (int q = lane; q < best_len; q += 32) {
            a.unit_top_end[(size_t)u * M + q] = W.best_end[q];
            a.unit_top_xsim[(size_t)u * M + q] = W.best_x[q];
        }
        __syncwarp();
    }
}

// One warp per start: distinct ends and paths are the sums over its units, its top-m the best m of the units' lists
// (the units of a start cover disjoint ends).
__global__ void __launch_bounds__(256) xsim_merge_kernel(xmap_xsim_args a) {
    const int x = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= a.n_starts) return;
    const int u0 = a.start_unit_ptr[x], u1 = a.start_unit_ptr[x + 1];
    const int M = a.top_m;
    int cnt = 0; long long comb = 0;
    for (int u = u0 + lane; u < u1; u += 32) { cnt += a.unit_count[u]; comb += a.unit_combos[u]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        comb += __shfl_xor_sync(0xffffffffu, comb, off);
    }
    if (lane == 0) { a.out_count[x] = cnt; a.out_combos[x] = comb; }
    unsigned long long last_k = ~0ull;
    int last_t = -1, got = 0;
    const int ncand = (u1 - u0) * M;
    for (int r = 0; r < M; ++r) {
        unsigned long long bk = 0ull; int bt = 0x7FFFFFFF, bp = -1;
        for (int q = lane; q < ncand; q += 32) {
            const int u = u0 + q / M, p = q % M;
            if (p >= a.unit_top_len[u]) continue;
            const size_t o = (size_t)u * M + p;
            const unsigned long long ak = abs_key(a.unit_top_xsim[o]);
            const int ee = a.unit_top_end[o];
            if (r > 0 && !better(last_k, last_t, ak, ee)) continue;
            if (bp < 0 || better(ak, ee, bk, bt)) { bk = ak; bt = ee; bp = (int)(o - (size_t)u0 * M); }
        }
        warp_argbest(bk, bt, bp);
        if (bp < 0) break;
        if (lane == 0) {
            a.top_end[(size_t)x * M + r] = bt;
            a.top_xsim[(size_t)x * M + r] = a.unit_top_xsim[(size_t)u0 * M + bp];
        }
        last_k = bk; last_t = bt;
        got = r + 1;
    }
    if (lane == 0) a.top_len[x] = got;
}

}  // namespace xmap

using namespace xmap;

extern "C" int64_t xmap_xsim_smem_bytes(int32_t cells_lg, int32_t warps) {
    return (int64_t)warps * ((int64_t)sizeof(WarpScratch) + ((int64_t)20 << cells_lg));
}

extern "C" int xmap_xsim_extend(const xmap_xsim_args *args_h, void *stream_) {
    const xmap_xsim_args &a = *args_h;
    if (a.n_starts <= 0 || a.n_units <= 0) return 0;
    if (a.top_m < 1 || a.top_m > XMAP_KMAX) return fail_msg("xmap_xsim_extend: top_m out of range");
    if (a.cells_lg < 6 || a.cells_lg > XMAP_XSIM_MAX_CELLS_LG) return fail_msg("xmap_xsim_extend: cells_lg out of range");
    if (a.gws && (a.gcells_lg < a.cells_lg || a.gcells_lg > 24)) return fail_msg("xmap_xsim_extend: gcells_lg out of range");
    if (a.gb < 0 || a.gb > 16) return fail_msg("xmap_xsim_extend: gb out of range");
    if (a.warps < 1 || a.warps * 32 > XT_MAX) return fail_msg("xmap_xsim_extend: warps per CTA out of range");
    if (!a.unit_counter) return fail_msg("xmap_xsim_extend: unit_counter is required");
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t smem = (size_t)xmap_xsim_smem_bytes(a.cells_lg, a.warps);
    if (smem > 227 * 1024) return fail_msg("xmap_xsim_extend: warps x table exceed shared memory");
    // two register budgets: up to 16 warps per CTA get 128 registers per thread, up to 20 get 96
    auto kern = a.warps * 32 <= 512 ? xsim_warp_kernel<512> : xsim_warp_kernel<XT_MAX>;
    XMAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XMAP_CUDA(cudaMemsetAsync(a.unit_counter, 0, sizeof(int32_t), st));
    int dev = 0, sms = 148;
    XMAP_CUDA(cudaGetDevice(&dev));
    XMAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long ctas = min((long long)sms, ((long long)a.n_units + a.warps - 1) / a.warps);
    kern<<<(unsigned)ctas, a.warps * 32, smem, st>>>(a);
    XMAP_LAUNCH_CHECK();
    if (a.merge) {
        xsim_merge_kernel<<<(unsigned)((a.n_starts + 7) / 8), 256, 0, st>>>(a);
        XMAP_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int xmap_xsim_merge(const xmap_xsim_args *args_h, void *stream_) {
    const xmap_xsim_args &a = *args_h;
    if (a.n_starts <= 0) return 0;
    if (a.top_m < 1 || a.top_m > XMAP_KMAX) return fail_msg("xmap_xsim_merge: top_m out of range");
    xsim_merge_kernel<<<(unsigned)((a.n_starts + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(a);
    XMAP_LAUNCH_CHECK();
    return 0;
}
