// X-SIM bridge extension, CTA-cooperative variant (see xsim.cu for the problem statement and the
// warp-per-unit variant).  One CTA of 8 warps owns one unit = (start item, a run of passes) and ONE
// shared-memory hash table of 2^cells_lg cells (8x the per-warp table of the other variant, so a start
// needs 8x fewer passes: longer sub-ranges per (leg, partner) pair, fewer pair look-ups, fewer duplicates
// per step).  Two CTAs are resident per SM, so one runs while the other waits at its barrier.
//
//   prep      256 (leg, partner) pairs at a time: each thread resolves one pair to a descriptor; a block
//             scan numbers the products of the macro-batch and compacts the non-empty descriptors;
//   produce   the products are dealt to the 8 warps in chunks of 32 (chunk c -> warp c mod 8): a lane finds
//             its descriptor, loads one 28-byte right segment (the next chunk's loads are issued first),
//             evaluates the path; lanes of the chunk that hit the same end are summed in lane (= path)
//             order by the lowest one, which leaves (end, num, den) in its payload slot; per
//             (owner, producer) lane masks say which slots belong to which owner warp (3 bits of a hash);
//   consume   after ONE block barrier, warp o walks the slots addressed to it in (producer, lane) order,
//             combines slots of the same end in that order and updates its private region of the table
//             with plain loads and stores (the key is claimed with a CAS): no atomics on values.
//
// The summation order of a (start, end) cell is a function of the path structure and the pass plan only.
#include "common.cuh"

namespace xmap {

constexpr int XT = 256;                       // threads per CTA
constexpr int XW = XT / 32;                   // warps = owners
constexpr int XMB = 256;                      // (leg, partner) pairs per macro-batch
constexpr int XOB = 3;                        // log2(XW): owner bits
constexpr unsigned XGOLD2 = 0x85EBCA6Bu;      // second multiplicative hash of the end: owner warp and home cell
                                              // (the first one, pi(y) = y * 0x9E3779B1, orders the lists: host side)
constexpr int XSTACK = 24;
constexpr int XF_BINS = 256;
constexpr int XSURV = 256;

// 16 sub-bins per octave for |xsim| in [2^-16, 2) (see sim_bin in sim.cu)
__device__ __forceinline__ int cxsim_bin(unsigned long long key_bits) {
    const int hi = int((key_bits & 0x7FFFFFFFFFFFFFFFull) >> 48);
    const int base = (1023 - 16) << 4;
    return max(0, min(XF_BINS - 1, hi - base));
}

struct CShared {
    // fixed part (the table follows, sized at launch)
    double d_N[XMB], d_D[XMB], d_C[XMB];      // descriptors of the macro-batch (compacted, non-empty)
    long long d_base[XMB];
    double p_num[2][XW][32], p_den[2][XW][32];   // payload slots, double-buffered; finalize scratch aliases them
    int d_cum[XMB + 4];                        // exclusive product prefix, d_cum[n_desc] = total
    int p_y[2][XW][32];
    unsigned p_mask[2][XW][XW];                // [owner][producer]
    int c_loc[XW][32];                         // consumer: payload slot (producer << 5 | lane) of the 32 items in flight
    unsigned long long s_wsum[XW];
    unsigned long long best_key[XMAP_KMAX];    // running top-m of the unit over its finished passes
    double best_x[XMAP_KMAX];
    int best_end[XMAP_KMAX];
    int best_len;
    int stack_g0[XSTACK], stack_g1[XSTACK];
    int stack_n, next_pass, cur_g0, cur_g1;
    int s_nd, s_total, s_overflow;
    int s_ins[XW];                             // occupied cells per owner region
    int s_cnt, s_nsurv, s_bstar, s_emit;
    int status;
    unsigned long long r_key[XW]; int r_tie[XW], r_pos[XW];   // block arg-best exchange (fallback path)
};

struct CFetched {
    double N, D, C, rn, rd, rc;
    int y;
    bool valid;
};

// products [32 c, 32 c + 32) of the macro-batch: descriptor lookup + the right-segment loads
__device__ __forceinline__ CFetched cfetch_chunk(const xmap_xsim_args &a, const CShared &S, int c, int nd, int total,
                                               int lane) {
    CFetched f;
    f.valid = false; f.y = 0; f.N = f.D = f.C = f.rn = f.rd = f.rc = 0.0;
    const int base_p = c << 5;
    if (base_p >= total) return f;
    // descriptor holding product base_p: largest d with d_cum[d] <= base_p (a 32-way and an 8-way step)
    constexpr int GS = XMB / 32;                          // descriptors per coarse group
    const int i1 = lane * GS;
    const unsigned m1 = __ballot_sync(0xffffffffu, i1 < nd && S.d_cum[i1] <= base_p);
    const int coarse = (__popc(m1) - 1) * GS;
    const int i2 = coarse + (lane % GS);
    const unsigned m2 = __ballot_sync(0xffffffffu, lane < GS && i2 < nd && S.d_cum[i2] <= base_p);
    const int d0 = coarse + __popc(m2) - 1;
    // inclusive product ends of the 32 descriptors from d0 on (every descriptor holds >= 1 product)
    const int e = S.d_cum[min(d0 + 1 + lane, nd)];
    const int p = base_p + lane;
    int l = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, e, l + step - 1);
        if (v <= p) l += step;
    }
    const int eprev = __shfl_sync(0xffffffffu, e, (l + 31) & 31);
    const int d0cum = S.d_cum[d0];
    if (p < total) {
        const int di = d0 + l;
        const int excl = l == 0 ? d0cum : eprev;
        const long long r = S.d_base[di] + (long long)(p - excl);
        f.valid = true;
        f.N = S.d_N[di]; f.D = S.d_D[di]; f.C = S.d_C[di];
        f.y = __ldg(a.rs_end + r);
        f.rn = __ldg(a.rs_n + r); f.rd = __ldg(a.rs_d + r); f.rc = __ldg(a.rs_c + r);
    }
    return f;
}

__global__ void __launch_bounds__(XT, 2) xsim_cta_kernel(xmap_xsim_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CShared &S = *reinterpret_cast<CShared *>(smem_raw);
    const int C = 1 << a.cells_lg;
    const int RB = a.cells_lg - XOB;                       // log2 of an owner's region
    double2 *vals = reinterpret_cast<double2 *>(smem_raw + ((sizeof(CShared) + 15) & ~(size_t)15));
    int *keys = reinterpret_cast<int *>(vals + C);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = a.unit_order ? a.unit_order[blockIdx.x] : (int)blockIdx.x;
    const long long leg_lo = a.unit_leg_lo[u], leg_hi = a.unit_leg_hi[u];
    const long long q_lo = a.lp_ptr[leg_lo], q_hi = a.lp_ptr[leg_hi];
    const int G = 1 << a.gb, G1 = G + 1;
    const int ug0 = a.unit_g0[u], ug1 = a.unit_g1[u], unpass = a.unit_npass[u];
    const int M = a.top_m;
    long long combos = 0;                                  // identical in every thread
    int unit_count = 0;
    unsigned ss = 0;                                       // superstep counter (payload buffer parity)
    if (tid == 0) {
        S.best_len = 0; S.stack_n = 0; S.next_pass = 0; S.status = 0; S.s_emit = 0;
    }
    __syncthreads();

    for (;;) {
        // ---- next pass: a split half if one is pending, else the unit's next own pass --------------------
        if (tid == 0) {
            if (S.stack_n > 0) { --S.stack_n; S.cur_g0 = S.stack_g0[S.stack_n]; S.cur_g1 = S.stack_g1[S.stack_n]; }
            else if (S.next_pass < unpass) {
                const long long w = ug1 - ug0;             // the unit's tile range cut into unpass equal passes
                S.cur_g0 = ug0 + (int)(w * S.next_pass / unpass);
                S.cur_g1 = ug0 + (int)(w * (S.next_pass + 1) / unpass);
                ++S.next_pass;
            } else S.cur_g0 = -1;
            S.s_overflow = 0; S.s_cnt = 0; S.s_nsurv = 0;
        }
        if (tid < XW) S.s_ins[tid] = 0;
        for (int c = tid; c < C; c += XT) keys[c] = 0;
        __syncthreads();
        const int g0 = S.cur_g0, g1 = S.cur_g1;
        if (g0 < 0) break;
        const bool whole = g0 == 0 && g1 == G;

        // =================== accumulate ======================================================
        long long q0 = q_lo;
        long long pass_combos = 0;
        while (q0 < q_hi) {
            const int nq = (int)min((long long)XMB, q_hi - q0);
            __syncthreads();
            if (S.s_overflow) break;                       // uniform: nobody writes the flag before the next barrier
            int len = 0;
            long long b = 0;
            double Nm = 0.0, Dm = 0.0, Cm = 0.0;
            if (tid < nq) {
                const long long q = q0 + tid;              // pair descriptors: one coalesced load each
                const int s = __ldg(a.pd_s + q);
                Nm = __ldg(a.pd_n + q); Dm = __ldg(a.pd_d + q); Cm = __ldg(a.pd_c + q);
                const long long rb = __ldg(a.rs_ptr + s);
                if (whole) { b = rb; len = (int)(__ldg(a.rs_ptr + s + 1) - rb); }
                else {
                    const int32_t *tp = a.tile_ptr + (size_t)s * G1;
                    const int b0 = __ldg(tp + g0), b1 = __ldg(tp + g1);
                    b = rb + b0; len = b1 - b0;
                }
            }
            // block scan of (non-empty flag, len)
            const unsigned long long mine = (len > 0 ? (1ull << 40) : 0ull) | (unsigned long long)len;
            unsigned long long incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            if (lane == 31) S.s_wsum[warp] = incl;
            __syncthreads();
            unsigned long long ws = lane < XW ? S.s_wsum[lane] : 0ull, wi = ws;
#pragma unroll
            for (int off = 1; off < XW; off <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            const unsigned long long tot = __shfl_sync(0xffffffffu, wi, XW - 1);
            const unsigned long long woff = __shfl_sync(0xffffffffu, wi - ws, warp);
            const unsigned long long excl = woff + incl - mine;
            if (len > 0) {
                const int c = (int)(excl >> 40);
                S.d_cum[c] = (int)(excl & ((1ull << 40) - 1ull));
                S.d_base[c] = b; S.d_N[c] = Nm; S.d_D[c] = Dm; S.d_C[c] = Cm;
            }
            const int nd = (int)(tot >> 40), total = (int)(tot & ((1ull << 40) - 1ull));
            if (tid == 0) S.d_cum[nd] = total;
            __syncthreads();
            q0 += nq;
            pass_combos += total;

            // ---- supersteps: 16 chunks of 32 products, one per warp -------------------------------
            const int nchunk = (total + 31) >> 5;
            CFetched nxt = cfetch_chunk(a, S, warp, nd, total, lane);
            for (int k = 0; k * XW < nchunk; ++k) {
                const CFetched cur = nxt;
                nxt = cfetch_chunk(a, S, (k + 1) * XW + warp, nd, total, lane);
                const int buf = ss & 1u; ++ss;
                // produce: evaluate, then sum the lanes of this chunk that hit the same end (lane = path order)
                int y = -1 - lane;
                double num = 0.0, den = 0.0;
                if (cur.valid) {
                    const double Nn = __dadd_rn(cur.N, cur.rn);
                    const double Dd = __dadd_rn(cur.D, cur.rd);
                    den = __dmul_rn(cur.C, cur.rc);
                    const double sp = (Dd != 0.0) ? __ddiv_rn(Nn, Dd) : 0.0;      // extender.py:88-89
                    num = __dmul_rn(sp, den);
                    y = cur.y;
                }
                // runs of equal ends in adjacent lanes (the lists are ordered by pi(end)) are summed by a segmented
                // suffix scan; only the heads of the runs go on
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
                if (cur.valid) {
                    S.p_num[buf][warp][lane] = num;
                    S.p_den[buf][warp][lane] = den;
                    S.p_y[buf][warp][lane] = cur.y;
                }
                if (lane < XW) S.p_mask[buf][lane][warp] = 0u;
                const unsigned gy = __match_any_sync(0xffffffffu, head ? y : (-1 - lane));
                __syncwarp();
                const bool lead = cur.valid && head && (__ffs(gy) - 1) == lane;
                if (lead && (gy & (gy - 1u))) {
                    double an = num, ad = den;
                    unsigned rem = gy & ~(1u << lane);
                    while (rem) {
                        const int b = __ffs(rem) - 1;
                        an = __dadd_rn(an, S.p_num[buf][warp][b]); ad = __dadd_rn(ad, S.p_den[buf][warp][b]);
                        rem &= rem - 1u;
                    }
                    S.p_num[buf][warp][lane] = an; S.p_den[buf][warp][lane] = ad;
                }
                const int owner = lead ? (int)(((unsigned)y * XGOLD2) >> (32 - XOB)) : XW;
                const unsigned grp = __match_any_sync(0xffffffffu, owner);
                if (lead && (__ffs(grp) - 1) == lane) S.p_mask[buf][owner][warp] = grp;
                __syncthreads();
                // consume: the slots addressed to owner `warp`, in (producer, lane) order
                const unsigned pm = lane < XW ? S.p_mask[buf][warp][lane] : 0u;
                const int cntl = __popc(pm);
                int incl_i = cntl;
#pragma unroll
                for (int off = 1; off < XW; off <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl_i, off);
                    if (lane >= off) incl_i += t;
                }
                const int tot_o = __shfl_sync(0xffffffffu, incl_i, XW - 1);
                const int region = warp << RB;
                for (int i0 = 0; i0 < tot_o; i0 += 32) {
                    const int i = i0 + lane;
                    const bool valid = i < tot_o;
                    int pr = 0;                            // smallest producer with incl > i
#pragma unroll
                    for (int step = XW / 2; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(0xffffffffu, incl_i, pr + step - 1);
                        if (v <= i) pr += step;
                    }
                    pr = min(pr, XW - 1);
                    const int before = __shfl_sync(0xffffffffu, incl_i - cntl, pr);
                    const unsigned mk = __shfl_sync(0xffffffffu, pm, pr);
                    int yy = -1 - lane, loc = 0;
                    if (valid) {
                        loc = (pr << 5) | __fns(mk, 0, i - before + 1);
                        yy = (&S.p_y[buf][0][0])[loc];
                    }
                    S.c_loc[warp][lane] = loc;
                    const unsigned g2 = __match_any_sync(0xffffffffu, yy);
                    __syncwarp();
                    if (valid && (__ffs(g2) - 1) == lane) {
                        // find-or-insert in the owner's region (only this warp touches it; the CAS settles lanes
                        // racing for one empty cell), then the slots of this end one by one in (producer, lane) order
                        const int key = yy + 1;
                        int pos = (int)((((unsigned)yy * XGOLD2) << XOB) >> (32 - RB));
                        int slot = -1;
                        bool isnew = false;
                        for (int probes = 0; probes < (1 << RB); ++probes) {
                            const int kcur = *(volatile int *)(keys + region + pos);
                            if (kcur == key) { slot = region + pos; break; }
                            if (kcur == 0) {
                                const int old = atomicCAS(keys + region + pos, 0, key);
                                if (old == 0) { slot = region + pos; isnew = true; break; }
                                if (old == key) { slot = region + pos; break; }
                            }
                            pos = (pos + 1) & ((1 << RB) - 1);
                        }
                        if (slot < 0) *(volatile int *)&S.s_overflow = 1;
                        else {
                            double an = 0.0, ad = 0.0;
                            if (!isnew) { const double2 v = vals[slot]; an = v.x; ad = v.y; }
                            unsigned rem = g2;
                            while (rem) {
                                const int l2 = S.c_loc[warp][__ffs(rem) - 1];
                                an = __dadd_rn(an, (&S.p_num[buf][0][0])[l2]); ad = __dadd_rn(ad, (&S.p_den[buf][0][0])[l2]);
                                rem &= rem - 1u;
                            }
                            vals[slot] = make_double2(an, ad);
                            if (isnew) atomicAdd(&S.s_ins[warp], 1);
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0 && S.s_ins[warp] > (1 << RB) - (1 << (RB - 3))) *(volatile int *)&S.s_overflow = 1;   // region past 7/8
            }
        }
        __syncthreads();
        if (S.s_overflow) {
            // the pass does not fit: halve its hash range and redo both halves (nothing of it was published)
            if (tid == 0) {
                if (g1 - g0 < 2 || S.stack_n + 2 > XSTACK) S.status = 1;
                else {
                    const int mid = (g0 + g1) >> 1;
                    S.stack_g0[S.stack_n] = mid; S.stack_g1[S.stack_n] = g1; ++S.stack_n;
                    S.stack_g0[S.stack_n] = g0; S.stack_g1[S.stack_n] = mid; ++S.stack_n;
                }
            }
            __syncthreads();
            continue;
        }
        combos += pass_combos;

        // =================== finalize the pass ==============================================
        unsigned *hist = reinterpret_cast<unsigned *>(&S.p_num[0][0][0]);                       // 1 KB
        unsigned long long *sv_key = reinterpret_cast<unsigned long long *>(&S.p_den[0][0][0]); // 4 KB
        double *sv_x = reinterpret_cast<double *>(&S.p_den[1][0][0]);                           // 4 KB
        int *sv_end = reinterpret_cast<int *>(&S.p_num[1][0][0]);                               // 2 KB
        for (int bq = tid; bq < XF_BINS; bq += XT) hist[bq] = 0u;
        __syncthreads();
        int mycnt = 0;
        for (int c0 = warp * 32; c0 < C; c0 += XT) {
            const int c = c0 + lane;
            const int kk = keys[c];
            const bool occ = kk != 0;
            double x = 0.0;
            if (occ) {
                const double2 v = vals[c];
                x = __ddiv_rn(v.x, v.y);                   // extender.py:198-201
                vals[c].x = x;
                atomicAdd(&hist[cxsim_bin(abs_key(x))], 1u);
                ++mycnt;
            }
            if (a.emit_ptr) {
                const unsigned mo = __ballot_sync(0xffffffffu, occ);
                int base = 0;
                if (lane == 0 && mo) base = atomicAdd(&S.s_emit, __popc(mo));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (occ) {
                    const long long o = a.emit_ptr[u] + base + __popc(mo & ((1u << lane) - 1u));
                    a.emit_end[o] = kk - 1; a.emit_xsim[o] = x;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mycnt += __shfl_xor_sync(0xffffffffu, mycnt, off);
        if (lane == 0 && mycnt) atomicAdd(&S.s_cnt, mycnt);
        __syncthreads();
        const int cnt_pass = S.s_cnt;
        unit_count += cnt_pass;
        const int want = min(M, cnt_pass);
        if (warp == 0) {
            // largest bin b* such that #(bin >= b*) >= want
            int bstar = 0, run = 0;
            bool found = false;
            for (int hb = XF_BINS - 32; hb >= 0 && !found && want > 0; hb -= 32) {
                unsigned suf = hist[hb + lane];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned t = __shfl_down_sync(0xffffffffu, suf, off);
                    if (lane + off < 32) suf += t;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, run + (int)suf >= want);
                if (hit) { bstar = hb + (31 - __clz(hit)); found = true; }
                else run += (int)__shfl_sync(0xffffffffu, suf, 0);
            }
            if (lane == 0) S.s_bstar = bstar;
        }
        __syncthreads();
        const int bstar = S.s_bstar;
        // survivors: cells at or above the threshold bin, plus the running best of the earlier passes
        for (int c = tid; c < C && want > 0; c += XT) {
            const int kk = keys[c];
            if (kk == 0) continue;
            const double x = vals[c].x;
            const unsigned long long ak = abs_key(x);
            if (cxsim_bin(ak) < bstar) continue;
            const int pos = atomicAdd(&S.s_nsurv, 1);
            if (pos < XSURV) { sv_key[pos] = ak; sv_x[pos] = x; sv_end[pos] = kk - 1; }
        }
        const int nbest = S.best_len;
        __syncthreads();
        int nsurv = S.s_nsurv;
        const int newlen = min(M, cnt_pass + nbest);
        if (nsurv + nbest <= XSURV) {
            if (tid < nbest) { sv_key[nsurv + tid] = S.best_key[tid]; sv_x[nsurv + tid] = S.best_x[tid]; sv_end[nsurv + tid] = S.best_end[tid]; }
            __syncthreads();
            nsurv += nbest;
            if (tid < nsurv) {
                const unsigned long long mk = sv_key[tid];
                const int me = sv_end[tid];
                int rank = 0;
                for (int t = 0; t < nsurv; ++t) rank += better(sv_key[t], sv_end[t], mk, me) ? 1 : 0;
                if (rank < newlen) { S.best_key[rank] = mk; S.best_x[rank] = sv_x[tid]; S.best_end[rank] = me; }
            }
            if (tid == 0) S.best_len = newlen;
        } else {
            // one bin holds too many equal values: plain rounds over every cell and the running best
            // (new list built in sv_*; every round picks the best candidate strictly after the last)
            __syncthreads();
            unsigned long long last_k = ~0ull; int last_t = -1;
            for (int r = 0; r < newlen; ++r) {
                unsigned long long bk = 0ull; int bt = 0x7FFFFFFF, bp = -1;
                for (int c = tid; c < C + nbest; c += XT) {
                    unsigned long long ak; int ee;
                    if (c < C) { const int kk = keys[c]; if (kk == 0) continue; ak = abs_key(vals[c].x); ee = kk - 1; }
                    else { ak = S.best_key[c - C]; ee = S.best_end[c - C]; }
                    if (r > 0 && !better(last_k, last_t, ak, ee)) continue;
                    if (bp < 0 || better(ak, ee, bk, bt)) { bk = ak; bt = ee; bp = c; }
                }
                warp_argbest(bk, bt, bp);
                if (lane == 0) { S.r_key[warp] = bk; S.r_tie[warp] = bt; S.r_pos[warp] = bp; }
                __syncthreads();
                bk = lane < XW ? S.r_key[lane] : 0ull; bt = lane < XW ? S.r_tie[lane] : 0x7FFFFFFF; bp = lane < XW ? S.r_pos[lane] : -1;
                warp_argbest(bk, bt, bp);
                if (tid == 0 && bp >= 0) {
                    sv_key[r] = bk; sv_end[r] = bt;
                    sv_x[r] = bp < C ? vals[bp].x : S.best_x[bp - C];
                }
                last_k = bk; last_t = bt;
                __syncthreads();
            }
            if (tid < newlen) { S.best_key[tid] = sv_key[tid]; S.best_x[tid] = sv_x[tid]; S.best_end[tid] = sv_end[tid]; }
            if (tid == 0) S.best_len = newlen;
        }
        __syncthreads();
    }

    // ---- publish the unit ------------------------------------------------------------------------
    if (tid == 0) {
        a.unit_count[u] = unit_count;
        a.unit_combos[u] = combos;
        a.unit_top_len[u] = S.best_len;
        if (S.status) atomicExch(a.error_flag, 2);
    }
    if (tid < S.best_len) {
        a.unit_top_end[(size_t)u * M + tid] = S.best_end[tid];
        a.unit_top_xsim[(size_t)u * M + tid] = S.best_x[tid];
    }
}

}  // namespace xmap

using namespace xmap;

extern "C" int64_t xmap_xsim_cta_smem_bytes(int32_t cells_lg) {
    return (int64_t)((sizeof(CShared) + 15) & ~(size_t)15) + ((int64_t)20 << cells_lg);
}

extern "C" int xmap_xsim_extend_cta(const xmap_xsim_args *args_h, void *stream_) {
    const xmap_xsim_args &a = *args_h;
    if (a.n_starts <= 0 || a.n_units <= 0) return 0;
    if (a.top_m < 1 || a.top_m > XMAP_KMAX) return fail_msg("xmap_xsim_extend_cta: top_m out of range");
    if (a.cells_lg < 9 || a.cells_lg > XMAP_XSIM_MAX_CELLS_LG) return fail_msg("xmap_xsim_extend_cta: cells_lg out of range");
    if (a.gb < 0 || a.gb > 16) return fail_msg("xmap_xsim_extend_cta: gb out of range");
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t smem = (size_t)xmap_xsim_cta_smem_bytes(a.cells_lg);
    if (smem > 227 * 1024) return fail_msg("xmap_xsim_extend_cta: table exceeds shared memory");
    XMAP_CUDA(cudaFuncSetAttribute(xsim_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xsim_cta_kernel<<<(unsigned)a.n_units, XT, smem, st>>>(a);
    XMAP_LAUNCH_CHECK();
    if (a.merge) return xmap_xsim_merge(args_h, stream_);
    return 0;
}
