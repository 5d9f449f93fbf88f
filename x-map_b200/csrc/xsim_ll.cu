// X-SIM bridge extension, record-list variant (see xsim.cu for the problem statement, xsim_cta.cu for the
// routed CTA variant this one replaces as the default).
//
// One CTA owns one unit = (start item, a run of passes) and ONE shared-memory hash table of 2^cells_lg cells
// {key, num, den}.  What changed against xsim_cta.cu is how a path reaches its cell.  There, every 256 paths
// were routed through payload slots to the warp owning the cell's region (two MATCH.ANY, a block barrier and a
// search per 32 paths): 22 warp-instructions per path, a fifth of the stall samples at the barrier.  Here a
// BATCH of up to 2 048 paths is evaluated with no routing at all:
//
//   phase A   chunks of 32 consecutive products are dealt to the warps (the right-segment loads stay coalesced); a
//             lane evaluates its path and stores (num, den) in the batch's staging arrays; the lanes of a chunk that
//             hit the same end are summed in lane (= path) order by the lowest one (one MATCH.ANY), which finds or
//             inserts the end's key in the table (CAS on the key only) and pushes ONE record on the cell's list:
//             old = atomicExch(head[cell], idx); next[idx] = old.  The thread that found the list empty is the
//             cell's reducer for this batch.
//   barrier
//   phase B   the reducer walks the cell's list once and adds the records to the cell IN ASCENDING CHUNK (= path)
//             order, whatever order they arrived in: one record -> nothing to order; two -> a + b is commutative;
//             more -> filed by chunk (a chunk contributes at most one record per cell), then summed by chunk number.
//             Plain loads and stores; no atomics on values.
//   barrier
//
// Two block barriers per batch instead of one per 256 paths, one MATCH.ANY per 32 paths instead of three, no
// cross-warp search.  The sum of a
// (start, end) cell is formed batch by batch, inside a batch in path order: a function of the path structure
// and the pass plan only, so results are bit-identical from run to run and for any number of GPUs.  Which
// cell a key lands in depends on the race of the CAS; no result depends on the cell index.
//
// Pass planning, device-side pass splits, top-m selection and the unit outputs are those of xsim_cta.cu.
#include "common.cuh"

namespace xmap {

constexpr int LMB = 256;                      // (leg, partner) pairs per macro-batch
constexpr int LT_MAX = 512;                   // threads per CTA: 256 or 512
constexpr int LR_MAX = 8;                     // paths per thread and batch
constexpr unsigned LGOLD2 = 0x85EBCA6Bu;      // multiplicative hash of the end: home cell
constexpr int LSTACK = 24;
constexpr int LF_BINS = 256;
constexpr int LSURV = 256;
constexpr unsigned LNIL = 0xFFFFu;            // end of a cell's record list (16-bit links)
constexpr int LCHUNKS = 64;                   // chunks of 32 paths per batch (a batch holds <= 2 048 paths)

__device__ __forceinline__ int lxsim_bin(unsigned long long key_bits) {
    const int hi = int((key_bits & 0x7FFFFFFFFFFFFFFFull) >> 48);
    const int base = (1023 - 16) << 4;
    return max(0, min(LF_BINS - 1, hi - base));
}

struct LShared {
    // fixed part; the table, the record lists and the staging arrays follow, sized at launch
    double d_N[LMB], d_D[LMB], d_C[LMB];      // descriptors of the macro-batch (compacted, non-empty)
    long long d_base[LMB];
    int d_cum[LMB + 4];                        // exclusive product prefix, d_cum[n_desc] = total
    int s_lp[LMB + 4];
    unsigned long long s_wsum[32];
    unsigned long long best_key[XMAP_KMAX];    // running top-m of the unit over its finished passes
    double best_x[XMAP_KMAX];
    int best_end[XMAP_KMAX];
    int best_len;
    int stack_g0[LSTACK], stack_g1[LSTACK];
    int stack_n, next_pass, cur_g0, cur_g1;
    int s_overflow, s_ins;
    long long s_nextleg;
    int s_cnt, s_nsurv, s_bstar, s_emit;
    int status;
    unsigned long long r_key[32]; int r_tie[32], r_pos[32];   // block arg-best exchange (fallback path)
};

struct LFetched {
    double N, D, C, rn, rd, rc;
    int y;
    bool valid;
};

// products [32 c, 32 c + 32) of the macro-batch: descriptor lookup + the right-segment loads
__device__ __forceinline__ LFetched lfetch_chunk(const xmap_xsim_args &a, const LShared &S, int c, int nd, int total,
                                                 int lane) {
    LFetched f;
    f.valid = false; f.y = 0; f.N = f.D = f.C = f.rn = f.rd = f.rc = 0.0;
    const int base_p = c << 5;
    if (base_p >= total) return f;
    // descriptor holding product base_p: largest d with d_cum[d] <= base_p (a 32-way and an 8-way step)
    constexpr int GS = LMB / 32;
    const int i1 = lane * GS;
    const unsigned m1 = __ballot_sync(0xffffffffu, i1 < nd && S.d_cum[i1] <= base_p);
    const int coarse = (__popc(m1) - 1) * GS;
    const int i2 = coarse + (lane % GS);
    const unsigned m2 = __ballot_sync(0xffffffffu, lane < GS && i2 < nd && S.d_cum[i2] <= base_p);
    const int d0 = coarse + __popc(m2) - 1;
    const int p = base_p + lane;
    const int d0cum = S.d_cum[d0];
    const int e0 = S.d_cum[d0 + 1];                       // d0 < nd
    int di = d0, excl = d0cum;
    if (e0 < base_p + 32) {
        // the chunk runs over a descriptor boundary: inclusive product ends of the 32 descriptors from d0 on
        // (every descriptor holds >= 1 product), searched by shuffles
        const int e = S.d_cum[min(d0 + 1 + lane, nd)];
        int l = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, e, l + step - 1);
            if (v <= p) l += step;
        }
        const int eprev = __shfl_sync(0xffffffffu, e, (l + 31) & 31);
        di = d0 + l;
        excl = l == 0 ? d0cum : eprev;
    }
    if (p < total) {
        const long long r = S.d_base[di] + (long long)(p - excl);
        f.valid = true;
        f.N = S.d_N[di]; f.D = S.d_D[di]; f.C = S.d_C[di];
        f.y = __ldg(a.rs_end + r);
        f.rn = __ldg(a.rs_n + r); f.rd = __ldg(a.rs_d + r); f.rc = __ldg(a.rs_c + r);
    }
    return f;
}

// phase B: the records of one cell, added in ascending record index.  A chunk of 32 paths contributes at most
// one record per cell (its lanes were combined in phase A), so a cell has at most one record per chunk of the
// batch (<= 64): the list is walked once, every record filed under its chunk, and the chunks summed in order.
__device__ __forceinline__ void ll_reduce(int slot, unsigned myidx, double2 *vals, unsigned *head,
                                          const double *st_num, const double *st_den, const unsigned short *st_next) {
    const unsigned first = head[slot];                     // newest arrival; the list ends at myidx (the oldest)
    head[slot] = LNIL;
    double sn, sd;
    if (first == myidx) { sn = st_num[first]; sd = st_den[first]; }
    else {
        const unsigned second = st_next[first];
        if (second == myidx) {                             // two records: a + b == b + a
            sn = __dadd_rn(st_num[first], st_num[second]); sd = __dadd_rn(st_den[first], st_den[second]);
        } else {
            unsigned short loc[LCHUNKS];
            unsigned long long mask = 0ull;
            for (unsigned p = first; p != LNIL; p = st_next[p]) {
                const unsigned c = p >> 5;
                loc[c] = (unsigned short)p;
                mask |= 1ull << c;
            }
            unsigned p = loc[__ffsll((long long)mask) - 1];
            mask &= mask - 1ull;
            sn = st_num[p]; sd = st_den[p];
            while (mask) {
                p = loc[__ffsll((long long)mask) - 1];
                mask &= mask - 1ull;
                sn = __dadd_rn(sn, st_num[p]); sd = __dadd_rn(sd, st_den[p]);
            }
        }
    }
    double2 v = vals[slot];
    v.x = __dadd_rn(v.x, sn); v.y = __dadd_rn(v.y, sd);
    vals[slot] = v;
}

__global__ void __launch_bounds__(LT_MAX, 1) xsim_ll_kernel(xmap_xsim_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LShared &S = *reinterpret_cast<LShared *>(smem_raw);
    const int T = blockDim.x, NW = T >> 5;
    const int C = 1 << a.cells_lg, B = 1 << a.batch_lg;
    const int R = B / T;                                   // paths per thread and batch, 1 .. LR_MAX
    double2 *vals = reinterpret_cast<double2 *>(smem_raw + ((sizeof(LShared) + 15) & ~(size_t)15));
    double *st_num = reinterpret_cast<double *>(vals + C);
    double *st_den = st_num + B;
    int *keys = reinterpret_cast<int *>(st_den + B);
    unsigned *head = reinterpret_cast<unsigned *>(keys + C);
    unsigned short *st_next = reinterpret_cast<unsigned short *>(head + C);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = a.unit_order ? a.unit_order[blockIdx.x] : (int)blockIdx.x;
    const long long leg_lo = a.unit_leg_lo[u], leg_hi = a.unit_leg_hi[u];
    const long long q_lo = a.lp_ptr[leg_lo], q_hi = a.lp_ptr[leg_hi];
    const int G = 1 << a.gb, G1 = G + 1;
    const int ug0 = a.unit_g0[u], ug1 = a.unit_g1[u], unpass = a.unit_npass[u];
    const int M = a.top_m;
    const int hshift = 32 - a.cells_lg;
    long long combos = 0;                                  // identical in every thread
    int unit_count = 0;
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
            S.s_overflow = 0; S.s_ins = 0; S.s_cnt = 0; S.s_nsurv = 0;
        }
        for (int c = tid; c < C; c += T) { keys[c] = 0; head[c] = LNIL; }
        __syncthreads();
        const int g0 = S.cur_g0, g1 = S.cur_g1;
        if (g0 < 0) break;
        const bool whole = g0 == 0 && g1 == G;

        // =================== accumulate ======================================================
        long long q0 = q_lo, cur_leg = leg_lo;
        long long pass_combos = 0;
        bool ovf = false;                                  // block-uniform
        while (q0 < q_hi && !ovf) {
            const int nq = (int)min((long long)LMB, q_hi - q0);
            const int nl = (int)min((long long)LMB, leg_hi - cur_leg);
            if (tid < nl) {
                const long long v = __ldg(a.lp_ptr + cur_leg + tid) - q0;
                S.s_lp[tid] = (int)max(-(1ll << 30), min(1ll << 30, v));
            }
            __syncthreads();
            int len = 0;
            long long b = 0;
            double Nm = 0.0, Dm = 0.0, Cm = 0.0;
            if (tid < nq) {
                int lo = 0, hi = nl;                       // largest leg slot with first pair <= tid
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (S.s_lp[mid] <= tid) lo = mid; else hi = mid;
                }
                const long long L = cur_leg + lo;
                const int pidx = tid - S.s_lp[lo];
                const long long p = __ldg(a.leg_par_base + L) + pidx;
                const int s = __ldg(a.par_s + p);
                Nm = __dadd_rn(__ldg(a.leg_n + L), __ldg(a.par_e + p));      // sums in path order (extender.py:85-88)
                Dm = __dadd_rn(__ldg(a.leg_d + L), __ldg(a.par_m + p));
                Cm = __dmul_rn(__ldg(a.leg_c + L), __ldg(a.par_f + p));
                const long long rb = __ldg(a.rs_ptr + s);
                if (whole) { b = rb; len = (int)(__ldg(a.rs_ptr + s + 1) - rb); }
                else {
                    const int32_t *tp = a.tile_ptr + (size_t)s * G1;
                    const int b0 = __ldg(tp + g0), b1 = __ldg(tp + g1);
                    b = rb + b0; len = b1 - b0;
                }
                if (tid == nq - 1) S.s_nextleg = (pidx + 1 == __ldg(a.leg_npar + L)) ? L + 1 : L;
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
            unsigned long long ws = lane < NW ? S.s_wsum[lane] : 0ull, wi = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            const unsigned long long tot = __shfl_sync(0xffffffffu, wi, NW - 1);
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
            cur_leg = S.s_nextleg; q0 += nq;
            pass_combos += total;

            // ---- batches of R * NW chunks of 32 products --------------------------------------------
            const int nchunk = (total + 31) >> 5;
            for (int cb = 0; cb < nchunk; cb += R * NW) {
                int myslot[LR_MAX];
                int n_new = 0;
                LFetched nxt = lfetch_chunk(a, S, cb + warp, nd, total, lane);
#pragma unroll
                for (int k = 0; k < LR_MAX; ++k) {
                    myslot[k] = -1;
                    if (k < R) {
                        const LFetched cur = nxt;
                        if (k + 1 < R) nxt = lfetch_chunk(a, S, cb + (k + 1) * NW + warp, nd, total, lane);
                        const unsigned idx = (unsigned)(((k * NW + warp) << 5) | lane);       // record index in the batch
                        int ykey = -1 - lane;
                        if (cur.valid) {
                            const double Nn = __dadd_rn(cur.N, cur.rn);
                            const double Dd = __dadd_rn(cur.D, cur.rd);
                            const double cp = __dmul_rn(cur.C, cur.rc);
                            const double sp = (Dd != 0.0) ? __ddiv_rn(Nn, Dd) : 0.0;      // extender.py:88-89
                            st_num[idx] = __dmul_rn(sp, cp);
                            st_den[idx] = cp;
                            ykey = cur.y;
                        }
                        // the lanes of this chunk that hit the same end become ONE record, summed in lane (= path) order
                        const unsigned grp = __match_any_sync(0xffffffffu, ykey);
                        __syncwarp();
                        if (cur.valid && (__ffs(grp) - 1) == lane) {
                            if (grp & (grp - 1u)) {
                                double an = st_num[idx], ad = st_den[idx];
                                unsigned rem = grp & (grp - 1u);
                                const unsigned cbase = idx & ~31u;
                                while (rem) {
                                    const unsigned b = cbase | (unsigned)(__ffs(rem) - 1);
                                    an = __dadd_rn(an, st_num[b]); ad = __dadd_rn(ad, st_den[b]);
                                    rem &= rem - 1u;
                                }
                                st_num[idx] = an; st_den[idx] = ad;
                            }
                            // find or insert the end's key (any thread may claim any cell; values are not touched here)
                            const int key = cur.y + 1;
                            int pos = (int)(((unsigned)cur.y * LGOLD2) >> hshift);
                            int slot = -1;
                            for (int probes = 0; probes < C; ++probes) {
                                const int kcur = *(volatile int *)(keys + pos);
                                if (kcur == key) { slot = pos; break; }
                                if (kcur == 0) {
                                    const int old = atomicCAS(keys + pos, 0, key);
                                    if (old == 0) { slot = pos; vals[pos] = make_double2(0.0, 0.0); ++n_new; break; }
                                    if (old == key) { slot = pos; break; }
                                }
                                pos = (pos + 1) & (C - 1);
                            }
                            if (slot < 0) *(volatile int *)&S.s_overflow = 1;
                            else {
                                const unsigned old = atomicExch(head + slot, idx);
                                st_next[idx] = (unsigned short)old;
                                if (old == LNIL) myslot[k] = slot;             // first arrival: this thread reduces the cell
                            }
                        }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) n_new += __shfl_xor_sync(0xffffffffu, n_new, off);
                if (lane == 0 && n_new) atomicAdd(&S.s_ins, n_new);
                __syncthreads();
                ovf = S.s_overflow != 0 || S.s_ins > C - (C >> 3);             // linear probing degrades past 7/8
                if (ovf) break;
#pragma unroll
                for (int k = 0; k < LR_MAX; ++k)
                    if (k < R && myslot[k] >= 0)
                        ll_reduce(myslot[k], (unsigned)(((k * NW + warp) << 5) | lane), vals, head, st_num, st_den, st_next);
                __syncthreads();
            }
        }
        if (ovf) {
            // the pass does not fit: halve its hash range and redo both halves (nothing of it was published)
            if (tid == 0) {
                if (g1 - g0 < 2 || S.stack_n + 2 > LSTACK) S.status = 1;
                else {
                    const int mid = (g0 + g1) >> 1;
                    S.stack_g0[S.stack_n] = mid; S.stack_g1[S.stack_n] = g1; ++S.stack_n;
                    S.stack_g0[S.stack_n] = g0; S.stack_g1[S.stack_n] = mid; ++S.stack_n;
                }
            }
            __syncthreads();
            if (S.status) break;                            // cannot be split any further: reported through error_flag
            continue;
        }
        combos += pass_combos;

        // =================== finalize the pass ==============================================
        // scratch in the staging arrays (B >= 512): st_num <- histogram (1 KB) + survivor ends (1 KB),
        // st_den <- survivor keys (2 KB) + values (2 KB)
        unsigned *hist = reinterpret_cast<unsigned *>(st_num);
        int *sv_end = reinterpret_cast<int *>(st_num) + LF_BINS;
        unsigned long long *sv_key = reinterpret_cast<unsigned long long *>(st_den);
        double *sv_x = st_den + LSURV;
        for (int bq = tid; bq < LF_BINS; bq += T) hist[bq] = 0u;
        __syncthreads();
        int mycnt = 0;
        for (int c0 = warp * 32; c0 < C; c0 += T) {
            const int c = c0 + lane;
            const int kk = keys[c];
            const bool occ = kk != 0;
            double x = 0.0;
            if (occ) {
                const double2 v = vals[c];
                x = __ddiv_rn(v.x, v.y);                   // extender.py:198-201
                vals[c].x = x;
                atomicAdd(&hist[lxsim_bin(abs_key(x))], 1u);
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
            for (int hb = LF_BINS - 32; hb >= 0 && !found && want > 0; hb -= 32) {
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
        for (int c = tid; c < C && want > 0; c += T) {
            const int kk = keys[c];
            if (kk == 0) continue;
            const double x = vals[c].x;
            const unsigned long long ak = abs_key(x);
            if (lxsim_bin(ak) < bstar) continue;
            const int pos = atomicAdd(&S.s_nsurv, 1);
            if (pos < LSURV) { sv_key[pos] = ak; sv_x[pos] = x; sv_end[pos] = kk - 1; }
        }
        const int nbest = S.best_len;
        __syncthreads();
        int nsurv = S.s_nsurv;
        const int newlen = min(M, cnt_pass + nbest);
        if (nsurv + nbest <= LSURV) {
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
                for (int c = tid; c < C + nbest; c += T) {
                    unsigned long long ak; int ee;
                    if (c < C) { const int kk = keys[c]; if (kk == 0) continue; ak = abs_key(vals[c].x); ee = kk - 1; }
                    else { ak = S.best_key[c - C]; ee = S.best_end[c - C]; }
                    if (r > 0 && !better(last_k, last_t, ak, ee)) continue;
                    if (bp < 0 || better(ak, ee, bk, bt)) { bk = ak; bt = ee; bp = c; }
                }
                warp_argbest(bk, bt, bp);
                if (lane == 0) { S.r_key[warp] = bk; S.r_tie[warp] = bt; S.r_pos[warp] = bp; }
                __syncthreads();
                bk = lane < NW ? S.r_key[lane] : 0ull; bt = lane < NW ? S.r_tie[lane] : 0x7FFFFFFF; bp = lane < NW ? S.r_pos[lane] : -1;
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

extern "C" int64_t xmap_xsim_ll_smem_bytes(int32_t cells_lg, int32_t batch_lg) {
    return (int64_t)((sizeof(LShared) + 15) & ~(size_t)15) + ((int64_t)24 << cells_lg) + ((int64_t)18 << batch_lg);
}

extern "C" int xmap_xsim_extend_ll(const xmap_xsim_args *args_h, void *stream_) {
    const xmap_xsim_args &a = *args_h;
    if (a.n_starts <= 0 || a.n_units <= 0) return 0;
    if (a.top_m < 1 || a.top_m > XMAP_KMAX) return fail_msg("xmap_xsim_extend_ll: top_m out of range");
    if (a.cells_lg < 9 || a.cells_lg > XMAP_XSIM_MAX_CELLS_LG) return fail_msg("xmap_xsim_extend_ll: cells_lg out of range");
    if (a.gb < 0 || a.gb > 16) return fail_msg("xmap_xsim_extend_ll: gb out of range");
    if (a.warps != 8 && a.warps != 16) return fail_msg("xmap_xsim_extend_ll: warps per CTA must be 8 or 16");
    const int T = a.warps * 32, B = 1 << a.batch_lg;
    if (a.batch_lg < 9 || B > 32 * LCHUNKS || B < T || B > T * LR_MAX)
        return fail_msg("xmap_xsim_extend_ll: batch_lg out of range (512 .. 2048 paths, 1 .. 8 per thread)");
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t smem = (size_t)xmap_xsim_ll_smem_bytes(a.cells_lg, a.batch_lg);
    if (smem > 227 * 1024) return fail_msg("xmap_xsim_extend_ll: table + batch exceed shared memory");
    XMAP_CUDA(cudaFuncSetAttribute(xsim_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xsim_ll_kernel<<<(unsigned)a.n_units, T, smem, st>>>(a);
    XMAP_LAUNCH_CHECK();
    if (a.merge) return xmap_xsim_merge(args_h, stream_);
    return 0;
}
