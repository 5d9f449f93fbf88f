// Item-item co-rating similarity rows (sparse R^T R, Gustavson row-wise) with a
// fused epilogue (adjusted-cosine / cosine, significance weighting, mutuality,
// exact-zero filter, cross-domain label) and fused per-row top-k selection.
//
// Reference semantics: baselinerSim.py:84-216 (pairs, similarity, mutuality,
// filter, label), :218-233 (adjacency), assist.py:82-87 (BB set) and
// extender.py:16-44 (find_knn_items).
//
// Numerics.  Per directed pair (i, j) three exact integer accumulators are
// kept: n_ij, the mutuality count, and the inner product in 64-bit fixed point
// with 2^-q resolution, q = 62 - r2_bits - min(cls_i, cls_j), cls = ceil(log2
// (item count)).  n_ij <= min(count_i, count_j) <= 2^cls, so the sum cannot
// overflow, and because integer addition is associative the result does not
// depend on the order products arrive in: any split of a row over warps, CTAs
// or GPUs gives bit-identical output, and sim(i,j) == sim(j,i) bitwise.
// Absolute error of the inner product <= n_ij * 2^-(q+1)  (q >= 33 at
// count 2^24; q = 54 for a tail pair with ratings in 1..5).
//
// Memory.  The hot loop streams 8-byte CSR entries of each rater's row
// (coalesced, read-only path); algorithmic bytes = 8 * (W + nnz).
#include "common.cuh"

namespace xmap {

constexpr int KMAX = XMAP_KMAX;

// A row with at least 2 * n_items products touches a large share of the columns: it is accumulated
// with fire-and-forget reductions only (no first-touch list, so no warp ever waits for an atomic to
// return) and evaluated by a linear, coalesced scan of its dense table.
__host__ __device__ __forceinline__ bool row_is_dense(long long row_work, int n_items) { return row_work >= 2LL * n_items; }



// Candidate buffer of the per-row top-k selection.  A candidate belongs to exactly one of the
// row's two lists; it is stored as a 128-bit sortable record:
//   skey = (list 0 ? 1<<63 : 0) | bits(|sim|)        (0 = empty pad)
//   spay = sign(sim)<<63 | item<<16 | position        (position indexes c_mutu / c_n)
// A CTA-wide bitonic sort (descending skey, ties to the smaller item) then leaves list 0's best
// first and list 1's best right after list 0's block.
template <int THREADS, int CAP_>
struct Scratch {
    static constexpr int CAP = CAP_;
    unsigned long long skey[CAP];
    unsigned long long spay[CAP];
    double t_sim[2][KMAX];
    unsigned long long thr[2];
    int c_mutu[CAP], c_n[CAP];
    int t_j[2][KMAX], t_mutu[2][KMAX], t_n[2][KMAX];
    int t_len[2];
    int fin_len[2][THREADS / 32];
    int ncand, n0;
};

constexpr unsigned long long LIST0_BIT = 1ull << 63;
constexpr unsigned long long ABS_MASK = ~LIST0_BIT;

__device__ __forceinline__ int pay_item(unsigned long long pay) { return int((pay & ABS_MASK) >> 16); }
__device__ __forceinline__ bool ranks_before(unsigned long long ka, unsigned long long pa,
                                             unsigned long long kb, unsigned long long pb) {
    return ka > kb || (ka == kb && pay_item(pa) < pay_item(pb));
}

struct RowCtx {
    int row, dom_i, prefix_i, cls_i;
    double den_i;
};

__device__ __forceinline__ RowCtx make_ctx(const xmap_sim_args &a, int row) {
    RowCtx c;
    c.row = row;
    c.dom_i = a.dom_code[row];
    c.prefix_i = a.prefix_code[row];
    const double *si = a.item_stats + 4 * (size_t)row;
    c.den_i = (a.method == XMAP_METHOD_COSINE) ? si[1] : si[2];
    c.cls_i = ceil_log2_u32((uint32_t)si[3]);
    return c;
}

// Epilogue of one pair: baselinerSim.py:163-173 / :125-141, :89, :95, :191, :207.
__device__ __forceinline__ bool eval_pair(const xmap_sim_args &a, const RowCtx &c, int j, int n, int mutu,
                                          long long fx, double &sim, int &label) {
    const double *sj = a.item_stats + 4 * (size_t)j;
    double den_j = (a.method == XMAP_METHOD_COSINE) ? sj[1] : sj[2];
    int cls_j = ceil_log2_u32((uint32_t)sj[3]);
    int q = 62 - a.r2_bits - min(c.cls_i, cls_j);
    double inner = (double)fx * pow2d(-q);
    double dd = __dmul_rn(c.den_i, den_j);
    double cosv = (dd != 0.0) ? __ddiv_rn(inner, dd) : 0.0;
    int mn = min(n, a.num_atleast);
    sim = __ddiv_rn(__dmul_rn(cosv, (double)mn), (double)a.num_atleast);
    label = (a.prefix_code[j] != c.prefix_i) ? 1 : 0;
    return sim != 0.0 && mutu != 0;
}

// Walk raters [lo, hi) of `row` (CSC order); for every other item j in each rater's CSR row call
// add(valid, j, agree, fx) with ALL lanes converged.  The raters are dealt round-robin to the
// `nwarps` warps (a row with few raters still occupies every warp of its CTA) and each warp
// prefetches the CSR extents of 32 of its raters lane-parallel (one latency round per 32 raters).
template <class Add>
__device__ __forceinline__ void accumulate_raters(const xmap_sim_args &a, int row, int cls_i, int lo, int hi,
                                                  int warp, int nwarps, Add add) {
    const int lane = lane_id();
    const bool adj = (a.method == XMAP_METHOD_ADJUST_COSINE);
    const int qbase = 62 - a.r2_bits;
    for (int first = lo + warp; first < hi; first += nwarps * 32) {
        const int e = first + nwarps * lane;
        uint32_t ux = 0;
        float r_l = 0.f;
        int rb = 0, re = 0;
        double mu_l = 0.0;
        if (e < hi) {
            uint2 ce = ld_ent(a.csc_ent + e);
            ux = ce.x;
            r_l = __uint_as_float(ce.y);
            int u = int(ux & 0x7FFFFFFFu);
            rb = __ldg(a.csr_ptr + u);
            re = __ldg(a.csr_ptr + u + 1);
            if (adj) mu_l = __ldg(a.user_mu + u);
        }
        const int cnt = min(32, (hi - first + nwarps - 1) / nwarps);
        for (int t = 0; t < cnt; ++t) {
            const uint32_t ux_t = __shfl_sync(0xffffffffu, ux, t);
            const float r_t = __shfl_sync(0xffffffffu, r_l, t);
            const int rb_t = __shfl_sync(0xffffffffu, rb, t);
            const int re_t = __shfl_sync(0xffffffffu, re, t);
            const double mu_t = __shfl_sync(0xffffffffu, mu_l, t);
            const int ge_i = int(ux_t >> 31);
            const double c_i = (double)r_t - mu_t;
            for (int k0 = rb_t; k0 < re_t; k0 += 32) {
                const int k = k0 + lane;
                bool valid = k < re_t;
                int j = 0;
                unsigned agree = 0;
                long long fx = 0;
                if (valid) {
                    uint2 en = ld_ent(a.csr_ent + k);
                    j = ent_item(en.x);
                    valid = (j != row);
                    double c_j = (double)__uint_as_float(en.y) - mu_t;
                    double p = __dmul_rn(c_i, c_j);
                    int q = qbase - min(cls_i, ent_cls(en.x));
                    fx = __double2ll_rn(p * pow2d(q));
                    agree = (ent_ge(en.x) == ge_i) ? 1u : 0u;
                }
                add(valid, j, agree, fx);
            }
        }
    }
}

// --------------------------------------------------------------------------
// Selection: two-stage tournament over the candidate buffer (plus the running tops).
//   stage 1  every warp takes a contiguous slice of the buffer, holds it in registers (up to
//            PER_LANE records per lane) and extracts its own best K per list with K rounds of
//            a register-local max + a 5-step warp arg-best; no shared-memory traffic, no barrier;
//   stage 2  warp 0 (list 0) and warp 1 (list 1) pick the best K among the <= nwarps*K finalists.
// K is small (10 by default, <= 64) against hundreds or thousands of candidates, so this is far
// cheaper than sorting the buffer.
// --------------------------------------------------------------------------
template <int THREADS, int CAP>
__device__ void select_lists(Scratch<THREADS, CAP> &S, int K) {
    constexpr int NW = THREADS / 32;
    constexpr int PER_LANE = CAP / THREADS;               // slice = PER_LANE * 32 records per warp
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0t = S.t_len[0], n1t = S.t_len[1];
    const int nc = S.ncand;
    // running tops re-enter as candidates of their own list
    if (tid < n0t) {
        const int p = nc + tid;
        const double v = S.t_sim[0][tid];
        S.skey[p] = LIST0_BIT | abs_key(v);
        S.spay[p] = (v < 0.0 ? LIST0_BIT : 0ull) | ((unsigned long long)S.t_j[0][tid] << 16) | (unsigned)p;
        S.c_mutu[p] = S.t_mutu[0][tid]; S.c_n[p] = S.t_n[0][tid];
    } else if (tid >= KMAX && tid < KMAX + n1t) {
        const int q = tid - KMAX, p = nc + n0t + q;
        const double v = S.t_sim[1][q];
        S.skey[p] = abs_key(v);
        S.spay[p] = (v < 0.0 ? LIST0_BIT : 0ull) | ((unsigned long long)S.t_j[1][q] << 16) | (unsigned)p;
        S.c_mutu[p] = S.t_mutu[1][q]; S.c_n[p] = S.t_n[1][q];
    }
    __syncthreads();
    const int total = nc + n0t + n1t;
    // ---- stage 1: per-warp top-K of its slice, from registers --------------------------------
    unsigned long long rk[PER_LANE], rp[PER_LANE];
    const int base = warp * (PER_LANE * 32);
#pragma unroll
    for (int r = 0; r < PER_LANE; ++r) {
        const int p = base + r * 32 + lane;
        rk[r] = (p < total) ? S.skey[p] : 0ull;
        rp[r] = (p < total) ? S.spay[p] : 0ull;
    }
    __syncthreads();                                        // everybody has its slice; reuse skey/spay
    // finalists of warp w, list l live at skey[(l * NW + w) * K + r]   (needs 2*NW*K <= CAP)
    for (int list = 0; list < 2; ++list) {
        const unsigned long long want = (list == 0) ? LIST0_BIT : 0ull;
        int got = 0;
        for (int rr = 0; rr < K; ++rr) {
            unsigned long long bk = 0ull, bpay = 0ull;
            int bt = 0x7FFFFFFF, bp = -1;
#pragma unroll
            for (int r = 0; r < PER_LANE; ++r) {
                const unsigned long long key = rk[r];
                if (key == 0ull || (key & LIST0_BIT) != want) continue;
                const int tt = pay_item(rp[r]);
                if (bp < 0 || better(key, tt, bk, bt)) { bk = key; bt = tt; bp = r * 32 + lane; bpay = rp[r]; }
            }
            // arg-best across the warp, carrying the payload of the winner
            unsigned long long k2 = bk; int t2 = bt, p2 = bp;
            warp_argbest(k2, t2, p2);
            if (p2 < 0) break;
            const int owner = p2 & 31, slot = p2 >> 5;
            const unsigned long long wpay = __shfl_sync(0xffffffffu, bpay, owner);
            if (lane == owner) {
#pragma unroll
                for (int r = 0; r < PER_LANE; ++r) if (r == slot) rk[r] = 0ull;   // taken
            }
            if (lane == 0) {
                const int f = (list * NW + warp) * K + rr;
                S.skey[f] = k2; S.spay[f] = wpay;
            }
            got = rr + 1;
        }
        if (lane == 0) S.fin_len[list][warp] = got;
    }
    __syncthreads();
    // ---- stage 2: best K among the finalists, one warp per list --------------------------------
    if (warp < 2) {
        const int list = warp;
        int got = 0;
        unsigned long long last_k = ~0ull; int last_t = -1;
        for (int rr = 0; rr < K; ++rr) {
            unsigned long long bk = 0ull; int bt = 0x7FFFFFFF, bp = -1;
            for (int w = 0; w < NW; ++w) {
                const int fl = S.fin_len[list][w];
                for (int r = lane; r < fl; r += 32) {
                    const int f = (list * NW + w) * K + r;
                    const unsigned long long key = S.skey[f];
                    const int tt = pay_item(S.spay[f]);
                    if (rr > 0 && !better(last_k, last_t, key, tt)) continue;   // already taken
                    if (bp < 0 || better(key, tt, bk, bt)) { bk = key; bt = tt; bp = f; }
                }
            }
            warp_argbest(bk, bt, bp);
            if (bp < 0) break;
            if (lane == 0) {
                const unsigned long long pay = S.spay[bp];
                const double mag = __longlong_as_double((long long)(bk & ABS_MASK));
                const int pos = int(pay & 0xFFFFull);
                S.t_sim[list][rr] = (pay & LIST0_BIT) ? -mag : mag;
                S.t_j[list][rr] = bt; S.t_mutu[list][rr] = S.c_mutu[pos]; S.t_n[list][rr] = S.c_n[pos];
            }
            last_k = bk; last_t = bt;
            got = rr + 1;
        }
        if (lane == 0) {
            S.t_len[list] = got;
            S.thr[list] = (got == K) ? (last_k & ABS_MASK) : 0ull;
        }
    }
    __syncthreads();
    if (tid == 0) { S.ncand = 0; S.n0 = 0; }
    __syncthreads();
}

// --------------------------------------------------------------------------
// Warp tiers: every row whose products fit one hash table is owned by ONE WARP.
//   - private open-addressing table (key, n<<16|mutu, 64-bit fixed-point inner product);
//   - plain read-modify-write: within one rater's CSR row all columns are distinct, so lanes never
//     collide on a value; only the insertion of a new key uses a compare-and-swap;
//   - no CTA-wide barrier anywhere: rows are independent, tens of them are in flight per SM;
//   - epilogue (similarity, filter, label), BB detection and both top-k lists in the same warp.
// Table placement: shared memory for 512 / 1024 / 2048 slots (row_work <= 350 / 700 / 1400),
// a per-warp slice of a global workspace (L2) for 8192 slots (row_work <= 5600), where warps are
// persistent and fetch rows from a counter.
// --------------------------------------------------------------------------
// One 16-byte slot per column: AoS so a global-memory table costs one 32-byte sector per product.
//   accumulate phase : key = item + 1, cnt = n << 16 | mutu, inner = fixed-point inner product
//   after compaction : the occupied slots sit contiguously at the front of the same array, and the
//                      epilogue overwrites inner with the similarity bits (0 = filtered) and tags key
//                      with the label (bit 31) and the candidate's list (bits 28-29: 0 none, 1, 2)
struct __align__(16) Slot {
    unsigned key, cnt;
    unsigned long long inner;
};
constexpr unsigned SLOT_ITEM_MASK = 0x01FFFFFFu;     // item + 1 <= 2^24

__host__ __device__ constexpr size_t warp_tab_bytes(int slots) { return (size_t)slots * sizeof(Slot); }

constexpr int SEL_BINS = 256;
constexpr int SEL_BUF = 192;                 // survivors buffer (aliases the histogram): 192 * 12 B
constexpr int SEL_BYTES = SEL_BUF * 12;      // >= SEL_BINS * 4

// 16 sub-bins per octave for |sim| in [2^-16, 2): bits 62..48 of the double, offset so that
// 2^-16 maps to bin 0; smaller values share bin 0, values >= 1 clamp to the top bin.
__device__ __forceinline__ int sim_bin(unsigned long long key_bits) {
    const int hi = int((key_bits & 0x7FFFFFFFFFFFFFFFull) >> 48);       // 11 exponent bits + 4 mantissa bits
    const int base = (1023 - 16) << 4;
    return max(0, min(SEL_BINS - 1, hi - base));
}

template <int LOG2_SLOTS>
__device__ void warp_row(const xmap_sim_args &a, Slot *__restrict__ S, unsigned char *sel, int row) {
    const int lane = threadIdx.x & 31;
    const RowCtx c = make_ctx(a, row);
    const int lo = a.csc_ptr[row], hi = a.csc_ptr[row + 1];
    const long long w = a.row_work[row];
    int log2n = 5;
    while (log2n < LOG2_SLOTS && (1LL << log2n) < 2 * w) ++log2n;
    const int nslots = 1 << log2n;
    const unsigned mask = nslots - 1;
    const int shift = 32 - log2n;
    for (int s = lane; s < nslots; s += 32) S[s] = Slot{0u, 0u, 0ull};
    __syncwarp();

    auto add = [&](bool valid, int j, unsigned agree, long long fx) {
        if (valid) {
            const unsigned key = (unsigned)j + 1u;
            unsigned slot = ((unsigned)j * 2654435761u) >> shift;
            bool ok = false;
            for (int probe = 0; probe < nslots; ++probe) {
                unsigned cur = *(volatile unsigned *)&S[slot].key;
                if (cur != key) {
                    if (cur == 0u) cur = atomicCAS(&S[slot].key, 0u, key);
                    if (cur != 0u && cur != key) { slot = (slot + 1) & mask; continue; }
                }
                ok = true;
                break;
            }
            if (ok) {
                S[slot].cnt += (1u << 16) | agree;
                S[slot].inner += (unsigned long long)fx;
            } else {
                atomicExch(a.error_flag, 1);
            }
        }
        __syncwarp();
    };
    accumulate_raters(a, row, c.cls_i, lo, hi, 0, 1, add);

    // in-place compaction: a chunk of 32 slots is read into registers before anything is written,
    // and the write positions never run ahead of the chunk being read
    int n_ent = 0;
    for (int s0 = 0; s0 < nslots; s0 += 32) {
        const Slot v = S[s0 + lane];
        const bool occ = v.key != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        __syncwarp();
        if (occ) S[n_ent + __popc(m & ((1u << lane) - 1u))] = v;
        n_ent += __popc(m);
        __syncwarp();
    }
    // One pass over the compacted entries: similarity (replaces the accumulator), filter, label,
    // the candidate's class, and a 256-bin histogram of |sim| per class.  Classes: mode 0 ->
    // 1 = other-domain ("cross"), 2 = same-domain; mode 1 -> 1 = neighbour is a bridge item.
    unsigned *hist = reinterpret_cast<unsigned *>(sel);          // [2][SEL_BINS]
    for (int b = lane; b < 2 * SEL_BINS; b += 32) hist[b] = 0u;
    __syncwarp();
    int lk = 0, ll = 0, n1 = 0, n2 = 0;
    for (int q = lane; q < n_ent; q += 32) {
        Slot v = S[q];
        double sim; int label;
        const int j = int(v.key - 1u);
        const bool keep = eval_pair(a, c, j, int(v.cnt >> 16), int(v.cnt & 0xFFFFu), (long long)v.inner, sim, label);
        unsigned cls = 0u;
        if (keep) {
            ++lk; ll |= label;
            if (a.mode == 1) cls = (a.bb_in[j] != 0) ? 1u : 0u;
            else cls = ((a.contains[j] >> c.dom_i) & 1) ? 2u : 1u;
            v.inner = (unsigned long long)__double_as_longlong(sim);
            if (cls) {
                atomicAdd(&hist[(cls - 1u) * SEL_BINS + sim_bin(v.inner)], 1u);
                if (cls == 1u) ++n1; else ++n2;
            }
        } else {
            v.inner = 0ull;
        }
        v.key |= (label ? 0x80000000u : 0u) | (cls << 28);
        S[q] = v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lk += __shfl_xor_sync(0xffffffffu, lk, off);
        ll |= __shfl_xor_sync(0xffffffffu, ll, off);
        n1 += __shfl_xor_sync(0xffffffffu, n1, off);
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    }
    __syncwarp();

    if (a.mode == 2) {  // emit every kept pair (materialised sim RDD, assist.py:75-77)
        const int64_t base = a.emit_ptr[row];
        int written = 0;
        for (int q0 = 0; q0 < n_ent; q0 += 32) {
            const int q = q0 + lane;
            Slot v = Slot{0u, 0u, 0ull};
            if (q < n_ent) v = S[q];
            const unsigned m = __ballot_sync(0xffffffffu, v.inner != 0ull);
            if (v.inner != 0ull) {
                const int64_t o = base + written + __popc(m & ((1u << lane) - 1u));
                a.emit_j[o] = int((v.key & SLOT_ITEM_MASK) - 1u);
                a.emit_sim[o] = __longlong_as_double((long long)v.inner);
                a.emit_mutu[o] = int(v.cnt & 0xFFFFu); a.emit_n[o] = int(v.cnt >> 16);
            }
            written += __popc(m);
        }
        __syncwarp();
        return;
    }
    const bool bb = (a.mode == 0) && (ll != 0);          // bridge item: a kept cross-domain pair (assist.py:84-86)
    if (a.mode == 0 && lane == 0) {
        a.row_flags[row] = bb ? 1 : 0;
        a.row_npairs[row] = n_ent;
        a.row_nkept[row] = lk;
    }
    // Lists (extender.py:30-43):  BB row: slot 0 = class 1, slot 1 = class 2;
    // NB row (pass 1): slot 1 = every kept neighbour (classes 1 and 2); pass 2: slot 0 = class 1.
    const int K = a.k;
    int bstar[2], want[2];
    unsigned cmask[2];                                    // classes that belong to the list
    for (int list = 0; list < 2; ++list) {
        bool use; int n_list;
        if (a.mode == 1) { use = (list == 0); cmask[list] = 2u; n_list = n1; }             // bit c = class c
        else if (bb) { use = true; cmask[list] = (list == 0) ? 2u : 4u; n_list = (list == 0) ? n1 : n2; }
        else { use = (list == 1); cmask[list] = 6u; n_list = n1 + n2; }
        want[list] = use ? min(K, n_list) : 0;
        bstar[list] = 0;
        if (want[list] == 0) continue;
        // largest bin b* such that #(bin >= b*) >= want: scan the (summed) histogram from the top
        int run = 0;
        bool found = false;
        for (int hb = SEL_BINS - 32; hb >= 0 && !found; hb -= 32) {
            unsigned v = 0u;
            if (cmask[list] & 2u) v += hist[hb + lane];
            if (cmask[list] & 4u) v += hist[SEL_BINS + hb + lane];
            unsigned suf = v;                              // suffix sums, highest lane = highest bin
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_down_sync(0xffffffffu, suf, off);
                if (lane + off < 32) suf += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, run + (int)suf >= want[list]);
            if (hit) { bstar[list] = hb + (31 - __clz(hit)); found = true; }
            else run += (int)__shfl_sync(0xffffffffu, suf, 0);
        }
    }
    __syncwarp();                                          // the histogram is dead; its memory becomes the survivor buffer
    for (int list = 0; list < 2; ++list) {
        if (list == 1 && a.mode != 0) continue;           // pass 2 only writes slot 0
        int got = 0;
        const size_t o = ((size_t)row * 2 + list) * K;
        if (want[list] > 0) {
            unsigned long long *bkey = reinterpret_cast<unsigned long long *>(sel);
            int *bent = reinterpret_cast<int *>(bkey + SEL_BUF);
            int nb = 0;                                    // warp-uniform
            bool overflow = false;
            for (int q0 = 0; q0 < n_ent; q0 += 32) {
                const int q = q0 + lane;
                bool take = false;
                unsigned long long key = 0ull;
                if (q < n_ent) {
                    const Slot v = S[q];
                    if ((cmask[list] >> ((v.key >> 28) & 3u)) & 1u) {
                        key = v.inner & 0x7FFFFFFFFFFFFFFFull;
                        take = sim_bin(key) >= bstar[list];
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (nb + __popc(m) > SEL_BUF) { overflow = true; break; }
                if (take) {
                    const int pos = nb + __popc(m & ((1u << lane) - 1u));
                    bkey[pos] = key; bent[pos] = q;
                }
                nb += __popc(m);
            }
            __syncwarp();
            unsigned long long last_k = ~0ull; int last_t = -1;
            for (int rr = 0; rr < want[list]; ++rr) {
                unsigned long long bk = 0; int bt = 0x7FFFFFFF, bp = -1;
                if (!overflow) {
                    for (int q = lane; q < nb; q += 32) {
                        const unsigned long long kk = bkey[q];
                        const int e = bent[q];
                        const int tt = int((S[e].key & SLOT_ITEM_MASK) - 1u);
                        if (rr > 0 && !better(last_k, last_t, kk, tt)) continue;
                        if (bp < 0 || better(kk, tt, bk, bt)) { bk = kk; bt = tt; bp = e; }
                    }
                } else {   // many equal similarities in the threshold bin: rounds over all entries
                    for (int q = lane; q < n_ent; q += 32) {
                        const Slot v = S[q];
                        if (!((cmask[list] >> ((v.key >> 28) & 3u)) & 1u)) continue;
                        const unsigned long long kk = v.inner & 0x7FFFFFFFFFFFFFFFull;
                        const int tt = int((v.key & SLOT_ITEM_MASK) - 1u);
                        if (rr > 0 && !better(last_k, last_t, kk, tt)) continue;
                        if (bp < 0 || better(kk, tt, bk, bt)) { bk = kk; bt = tt; bp = q; }
                    }
                }
                warp_argbest(bk, bt, bp);
                if (bp < 0) break;
                if (lane == 0) {
                    const Slot v = S[bp];
                    a.tab_idx[o + rr] = bt;
                    a.tab_sim[o + rr] = __longlong_as_double((long long)v.inner);
                    a.tab_mutu[o + rr] = int(v.cnt & 0xFFFFu); a.tab_n[o + rr] = int(v.cnt >> 16);
                }
                last_k = bk; last_t = bt;
                got = rr + 1;
            }
            __syncwarp();
        }
        if (lane == 0) a.tab_len[(size_t)row * 2 + list] = got;
    }
    __syncwarp();
}

// shared-memory tables: WARPS rows per CTA, rows assigned statically (sorted by descending work)
template <int LOG2_SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sim_warp_smem_kernel(xmap_sim_args a, const int32_t *__restrict__ rows,
                                                                   int n_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    unsigned char *mine = smem_raw + (threadIdx.x >> 5) * (warp_tab_bytes(1 << LOG2_SLOTS) + SEL_BYTES);
    warp_row<LOG2_SLOTS>(a, reinterpret_cast<Slot *>(mine), mine + warp_tab_bytes(1 << LOG2_SLOTS), rows[r]);
}

// global (L2) tables: persistent warps, rows fetched from a counter
template <int LOG2_SLOTS>
__global__ void __launch_bounds__(128) sim_warp_gmem_kernel(xmap_sim_args a, const int32_t *__restrict__ rows,
                                                            int n_rows, unsigned char *__restrict__ workspace,
                                                            int32_t *__restrict__ counter) {
    __shared__ __align__(16) unsigned char s_sel[4][SEL_BYTES];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Slot *T = reinterpret_cast<Slot *>(workspace + (size_t)gw * warp_tab_bytes(1 << LOG2_SLOTS));
    while (true) {
        int r = 0;
        if (lane == 0) r = atomicAdd(counter, 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= n_rows) break;
        warp_row<LOG2_SLOTS>(a, T, s_sel[threadIdx.x >> 5], rows[r]);
    }
}

// --------------------------------------------------------------------------
// Heavy rows: chunks of raters -> dense per-row table with 64-bit atomics.
// --------------------------------------------------------------------------
constexpr int BIG_THREADS = 256;
constexpr int RATER_GROUP = 128;       // raters per unit of work (one warp fetches one group)

// Work = groups of RATER_GROUP consecutive raters, row-major (so concurrently running warps work
// on the same few rows and their tables stay L2-resident).  Every warp fetches its own groups
// from a global counter: no CTA-wide barrier anywhere.
__global__ void __launch_bounds__(BIG_THREADS) sim_big_accum_kernel(
    xmap_sim_args a, const int32_t *__restrict__ rows, int n_rows, const int64_t *__restrict__ grp_off,
    ulonglong2 *__restrict__ table, int32_t *__restrict__ touched, int32_t *__restrict__ touched_n,
    int32_t *__restrict__ work_counter) {
    __shared__ int s_stage[BIG_THREADS / 32][64];     // per-warp staging of first-touched columns
    const int lane = threadIdx.x & 31;
    int *stage = s_stage[threadIdx.x >> 5];
    const long long n_groups = grp_off[n_rows];
    while (true) {
        int g = 0;
        if (lane == 0) g = atomicAdd(work_counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_groups) break;
        int lo_b = 0, hi_b = n_rows;                  // largest b with grp_off[b] <= g
        while (hi_b - lo_b > 1) {
            const int mid = (lo_b + hi_b) >> 1;
            if (grp_off[mid] <= g) lo_b = mid; else hi_b = mid;
        }
        const int b = lo_b, row = rows[b];
        const int c_lo = a.csc_ptr[row], c_hi = a.csc_ptr[row + 1];
        const int lo = c_lo + (int)(g - grp_off[b]) * RATER_GROUP;
        const int hi = min(lo + RATER_GROUP, c_hi);
        const int cls_i = ceil_log2_u32((uint32_t)a.item_stats[4 * (size_t)row + 3]);
        ulonglong2 *T = table + (size_t)b * a.n_items;
        int32_t *tl = touched + (size_t)b * a.n_items;
        int32_t *tn = touched_n + b;
        const bool dense = row_is_dense(a.row_work[row], a.n_items);
        if (dense) {                                   // fire-and-forget: nothing to wait for
            auto add_red = [&](bool valid, int j, unsigned agree, long long fx) {
                if (valid) {
                    atomicAdd(&T[j].x, (1ull << 32) | (unsigned long long)agree);
                    atomicAdd(&T[j].y, (unsigned long long)fx);
                }
            };
            accumulate_raters(a, row, cls_i, lo, hi, 0, 1, add_red);
            continue;
        }
        int nstage = 0;                                // warp-uniform
        auto flush = [&](int count) {                  // move `count` staged columns to the row's list
            int basep = 0;
            if (lane == 0) basep = atomicAdd(tn, count);
            basep = __shfl_sync(0xffffffffu, basep, 0);
            for (int q = lane; q < count; q += 32) tl[basep + q] = stage[q];
            __syncwarp();
        };
        auto add = [&](bool valid, int j, unsigned agree, long long fx) {
            bool first = false;
            if (valid) {
                unsigned long long old = atomicAdd(&T[j].x, (1ull << 32) | (unsigned long long)agree);
                atomicAdd(&T[j].y, (unsigned long long)fx);
                first = (old == 0ull);
            }
            const unsigned m = __ballot_sync(0xffffffffu, first);
            if (m) {
                if (first) stage[nstage + __popc(m & ((1u << lane) - 1u))] = j;
                nstage += __popc(m);
                __syncwarp();
                if (nstage > 32) {                     // keep at most 32 staged so the next round fits
                    flush(nstage);
                    nstage = 0;
                }
            }
        };
        accumulate_raters(a, row, cls_i, lo, hi, 0, 1, add);
        if (nstage) flush(nstage);
    }
}

constexpr int FIN_THREADS = 256;
constexpr int EVAL_TILE = 2048;
constexpr int EVAL_UNROLL = 4;

// Scratch of the heavy-row epilogue (device memory, caller-provided):
struct BigScratch {
    long long *ent_off;     // [n_rows + 1] scan of per-row entry counts (n_items for dense rows)
    long long *tile_off;    // [n_rows + 1] scan of per-row tile counts
    int *row_kept;          // [n_rows]
    int *row_label;         // [n_rows]
    int *row_pairs;         // [n_rows] co-rated columns of the row
    int *tile_counter;      // [1]
    double *c_sim;          // [capacity] 0.0 = filtered / empty
    int *c_j, *c_mutu, *c_n;
    unsigned char *c_label;
    long long capacity;
};

__global__ void big_scan_kernel(const int32_t *__restrict__ touched_n, const int32_t *__restrict__ rows,
                                const int64_t *__restrict__ row_work, int n_rows, int n_items, BigScratch sc,
                                int32_t *error_flag) {
    __shared__ long long s_ent[1024], s_tile[1024];
    const int tid = threadIdx.x;
    const int per = (n_rows + 1023) / 1024;
    long long le = 0, lt = 0;
    for (int q = 0; q < per; ++q) {
        int b = tid * per + q;
        if (b < n_rows) {
            const long long len = row_is_dense(row_work[rows[b]], n_items) ? n_items : touched_n[b];
            le += len; lt += (len + EVAL_TILE - 1) / EVAL_TILE;
        }
    }
    s_ent[tid] = le; s_tile[tid] = lt;
    __syncthreads();
    if (tid == 0) {
        long long re = 0, rt = 0;
        for (int t = 0; t < 1024; ++t) {
            long long v = s_ent[t]; s_ent[t] = re; re += v;
            v = s_tile[t]; s_tile[t] = rt; rt += v;
        }
        sc.ent_off[n_rows] = re; sc.tile_off[n_rows] = rt;
        *sc.tile_counter = 0;
        if (re > sc.capacity) atomicExch(error_flag, 4);
    }
    __syncthreads();
    long long re = s_ent[tid], rt = s_tile[tid];
    for (int q = 0; q < per; ++q) {
        int b = tid * per + q;
        if (b < n_rows) {
            const bool dense = row_is_dense(row_work[rows[b]], n_items);
            const long long len = dense ? n_items : touched_n[b];
            sc.ent_off[b] = re; sc.tile_off[b] = rt;
            re += len; rt += (len + EVAL_TILE - 1) / EVAL_TILE;
            sc.row_kept[b] = 0; sc.row_label[b] = 0; sc.row_pairs[b] = dense ? 0 : touched_n[b];
        }
    }
}

// Grid-wide epilogue over every touched cell of the batch: similarity, filter, label -> compact
// candidate records; clears the table cell.  Work is cut into tiles of EVAL_TILE entries of ONE row,
// so a row with 400 K neighbours is spread over ~200 CTAs instead of being one CTA's long pole.
__global__ void __launch_bounds__(256) big_eval_kernel(xmap_sim_args a, const int32_t *__restrict__ rows, int n_rows,
                                                       ulonglong2 *__restrict__ table,
                                                       const int32_t *__restrict__ touched,
                                                       const int32_t *__restrict__ touched_n, BigScratch sc) {
    __shared__ int s_tile, s_b, s_kept, s_label, s_pairs;
    const int tid = threadIdx.x;
    if (sc.ent_off[n_rows] > sc.capacity) return;
    const long long n_tiles = sc.tile_off[n_rows];
    while (true) {
        if (tid == 0) {
            const int t = atomicAdd(sc.tile_counter, 1);
            s_tile = t; s_kept = 0; s_label = 0; s_pairs = 0;
            int lo = 0, hi = n_rows;                  // largest b with tile_off[b] <= t
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sc.tile_off[mid] <= t) lo = mid; else hi = mid;
            }
            s_b = lo;
        }
        __syncthreads();
        const long long t = s_tile;
        const int b = s_b;
        if (t >= n_tiles) break;
        const int row = rows[b];
        const RowCtx c = make_ctx(a, row);
        const bool dense = row_is_dense(a.row_work[row], a.n_items);
        const int len = dense ? a.n_items : touched_n[b];
        const int e0 = (int)(t - sc.tile_off[b]) * EVAL_TILE;
        const int e1 = min(e0 + EVAL_TILE, len);
        ulonglong2 *T = table + (size_t)b * a.n_items;
        const int32_t *tl = touched + (size_t)b * a.n_items;
        const long long goff = sc.ent_off[b];
        const bool cosine = (a.method == XMAP_METHOD_COSINE);
        int lk = 0, ll = 0, lp = 0;
        for (int eb = e0 + tid; eb < e1; eb += 256 * EVAL_UNROLL) {
            int j[EVAL_UNROLL];
            ulonglong2 cell[EVAL_UNROLL];
            double den_j[EVAL_UNROLL], cnt_j[EVAL_UNROLL];
            int pre_j[EVAL_UNROLL];
#pragma unroll
            for (int u = 0; u < EVAL_UNROLL; ++u) {
                const int e = eb + u * 256;
                j[u] = (e < e1) ? (dense ? e : tl[e]) : -1;
            }
#pragma unroll
            for (int u = 0; u < EVAL_UNROLL; ++u) {
                cell[u] = make_ulonglong2(0ull, 0ull);
                if (j[u] >= 0) cell[u] = T[j[u]];
            }
#pragma unroll
            for (int u = 0; u < EVAL_UNROLL; ++u) {
                if (j[u] >= 0 && cell[u].x != 0ull) {
                    const double *sj = a.item_stats + 4 * (size_t)j[u];
                    den_j[u] = cosine ? sj[1] : sj[2];
                    cnt_j[u] = sj[3];
                    pre_j[u] = a.prefix_code[j[u]];
                    T[j[u]] = make_ulonglong2(0ull, 0ull);
                }
            }
#pragma unroll
            for (int u = 0; u < EVAL_UNROLL; ++u) {
                const int e = eb + u * 256;
                if (j[u] < 0) continue;
                const long long g = goff + e;
                if (cell[u].x == 0ull) {               // dense scan: an untouched column
                    if (a.mode != 2) sc.c_sim[g] = 0.0;
                    continue;
                }
                ++lp;
                const int n = int(cell[u].x >> 32), mutu = int(cell[u].x & 0xFFFFFFFFull);
                const int q = 62 - a.r2_bits - min(c.cls_i, ceil_log2_u32((uint32_t)cnt_j[u]));
                const double inner = (double)(long long)cell[u].y * pow2d(-q);
                const double dd = __dmul_rn(c.den_i, den_j[u]);
                const double cosv = (dd != 0.0) ? __ddiv_rn(inner, dd) : 0.0;
                const double sim = __ddiv_rn(__dmul_rn(cosv, (double)min(n, a.num_atleast)), (double)a.num_atleast);
                const int label = (pre_j[u] != c.prefix_i) ? 1 : 0;
                const bool keep = sim != 0.0 && mutu != 0;
                if (a.mode == 2) {
                    if (keep) {
                        const int pos = atomicAdd(&a.emit_cursor[row], 1);
                        const int64_t o = a.emit_ptr[row] + pos;
                        a.emit_j[o] = j[u]; a.emit_sim[o] = sim; a.emit_mutu[o] = mutu; a.emit_n[o] = n;
                    }
                } else {
                    sc.c_sim[g] = keep ? sim : 0.0; sc.c_j[g] = j[u]; sc.c_mutu[g] = mutu; sc.c_n[g] = n;
                    sc.c_label[g] = (unsigned char)label;
                }
                if (keep) { ++lk; ll |= label; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lk += __shfl_xor_sync(0xffffffffu, lk, off);
            ll |= __shfl_xor_sync(0xffffffffu, ll, off);
            lp += __shfl_xor_sync(0xffffffffu, lp, off);
        }
        if ((tid & 31) == 0) {
            if (lk) atomicAdd(&s_kept, lk);
            if (ll) atomicOr(&s_label, 1);
            if (lp) atomicAdd(&s_pairs, lp);
        }
        __syncthreads();
        if (tid == 0) {
            if (s_kept) atomicAdd(&sc.row_kept[b], s_kept);
            if (s_label) atomicOr(&sc.row_label[b], 1);
            if (dense && s_pairs) atomicAdd(&sc.row_pairs[b], s_pairs);
        }
        __syncthreads();
    }
}

constexpr int SEL_UNROLL = 4;                 // candidate records per thread per step
constexpr int SEL_CAP = 8 * FIN_THREADS;      // candidate buffer of the heavy-row selection

// One CTA per heavy row: stream the row's candidate records (coalesced, SEL_UNROLL independent
// 8-byte loads per thread per step), drop everything below the current K-th best before touching
// the rest of the record, and keep the running top-K per list with the register tournament.
__global__ void __launch_bounds__(FIN_THREADS) big_select_kernel(xmap_sim_args a, const int32_t *__restrict__ rows,
                                                                 int n_rows, int32_t *__restrict__ touched_n,
                                                                 BigScratch sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using ScratchT = Scratch<FIN_THREADS, SEL_CAP>;
    ScratchT &S = *reinterpret_cast<ScratchT *>(smem_raw);
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int row = rows[b];
    if (sc.ent_off[n_rows] > sc.capacity) return;
    const long long off = sc.ent_off[b];
    const int n_entries = (int)(sc.ent_off[b + 1] - off);
    const int n_pairs = sc.row_pairs[b];
    if (tid == 0) touched_n[b] = 0;
    if (a.mode == 2) return;
    const RowCtx c = make_ctx(a, row);
    const int K = a.k;
    if (tid == 0) {
        S.ncand = 0; S.n0 = 0;
        S.t_len[0] = S.t_len[1] = 0; S.thr[0] = S.thr[1] = 0ull;
    }
    __syncthreads();
    bool bb = false;
    if (a.mode == 0) {
        bb = sc.row_label[b] != 0;                       // a kept cross-domain pair (assist.py:84-86)
        if (tid == 0) {
            a.row_flags[row] = bb ? 1 : 0;
            a.row_npairs[row] = n_pairs;
            a.row_nkept[row] = sc.row_kept[b];
        }
    }
    const bool use0 = (a.mode == 1) || bb;               // list definitions: extender.py:30-43
    const bool use1 = (a.mode == 0);
    for (int base = 0; base < n_entries; base += FIN_THREADS * SEL_UNROLL) {
        if (S.ncand + FIN_THREADS * SEL_UNROLL + 2 * KMAX > SEL_CAP) select_lists<FIN_THREADS, SEL_CAP>(S, K);
        // the smallest threshold any list in use still accepts
        unsigned long long floor_key = 0ull;
        {
            const unsigned long long f0 = use0 ? S.thr[0] : ~0ull, f1 = use1 ? S.thr[1] : ~0ull;
            floor_key = (f0 < f1) ? f0 : f1;
            if (floor_key == ~0ull) floor_key = 0ull;
        }
        double sv[SEL_UNROLL];
#pragma unroll
        for (int u = 0; u < SEL_UNROLL; ++u) {
            const int e = base + u * FIN_THREADS + tid;
            sv[u] = (e < n_entries) ? sc.c_sim[off + e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < SEL_UNROLL; ++u) {
            if (sv[u] == 0.0) continue;
            const unsigned long long key = abs_key(sv[u]);
            if (key < floor_key) continue;
            const long long g = off + base + u * FIN_THREADS + tid;
            const int j = sc.c_j[g];
            int lst;
            if (a.mode == 1) lst = (a.bb_in[j] != 0) ? 0 : -1;
            else if (bb) lst = ((a.contains[j] >> c.dom_i) & 1) ? 1 : 0;
            else lst = 1;
            if (lst >= 0 && S.thr[lst] && key < S.thr[lst]) lst = -1;
            if (lst >= 0) {
                const int p = atomicAdd(&S.ncand, 1);
                if (lst == 0) atomicAdd(&S.n0, 1);
                S.skey[p] = (lst == 0 ? LIST0_BIT : 0ull) | key;
                S.spay[p] = (sv[u] < 0.0 ? LIST0_BIT : 0ull) | ((unsigned long long)j << 16) | (unsigned)p;
                S.c_mutu[p] = sc.c_mutu[g]; S.c_n[p] = sc.c_n[g];
            }
        }
        __syncthreads();
    }
    select_lists<FIN_THREADS, SEL_CAP>(S, K);
    // write tables [n_items][2][K]
    for (int slot = 0; slot < 2; ++slot) {
        if (slot == 1 && a.mode != 0) continue;
        const int len = ((slot == 0) ? use0 : use1) ? S.t_len[slot] : 0;
        const size_t o = ((size_t)row * 2 + slot) * K;
        if (tid < len) {
            a.tab_idx[o + tid] = S.t_j[slot][tid]; a.tab_sim[o + tid] = S.t_sim[slot][tid];
            a.tab_mutu[o + tid] = S.t_mutu[slot][tid]; a.tab_n[o + tid] = S.t_n[slot][tid];
        }
        if (tid == 0) a.tab_len[(size_t)row * 2 + slot] = len;
    }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t big_scratch_layout(int n_rows, long long capacity, char *base, BigScratch *sc) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return base ? base + o : (char *)nullptr; };
    char *p_off = take(sizeof(long long) * ((size_t)n_rows + 1));
    char *p_toff = take(sizeof(long long) * ((size_t)n_rows + 1));
    char *p_kept = take(sizeof(int) * (size_t)n_rows);
    char *p_lab = take(sizeof(int) * (size_t)n_rows);
    char *p_pairs = take(sizeof(int) * (size_t)n_rows);
    char *p_cnt = take(256);
    char *p_sim = take(sizeof(double) * (size_t)capacity);
    char *p_j = take(sizeof(int) * (size_t)capacity);
    char *p_m = take(sizeof(int) * (size_t)capacity);
    char *p_n = take(sizeof(int) * (size_t)capacity);
    char *p_l = take((size_t)capacity);
    if (sc) {
        sc->ent_off = (long long *)p_off; sc->tile_off = (long long *)p_toff; sc->row_kept = (int *)p_kept; sc->row_label = (int *)p_lab; sc->row_pairs = (int *)p_pairs;
        sc->tile_counter = (int *)p_cnt; sc->c_sim = (double *)p_sim; sc->c_j = (int *)p_j;
        sc->c_mutu = (int *)p_m; sc->c_n = (int *)p_n; sc->c_label = (unsigned char *)p_l;
        sc->capacity = capacity;
    }
    return off;
}

template <int LOG2_SLOTS, int WARPS>
static int launch_warp_smem(const xmap_sim_args &a, const int32_t *rows, int n_rows, cudaStream_t st) {
    const size_t smem = WARPS * (warp_tab_bytes(1 << LOG2_SLOTS) + SEL_BYTES);
    auto kern = sim_warp_smem_kernel<LOG2_SLOTS, WARPS>;
    XMAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(n_rows + WARPS - 1) / WARPS, WARPS * 32, smem, st>>>(a, rows, n_rows);
    XMAP_LAUNCH_CHECK();
    return 0;
}

constexpr int GMEM_CTAS_PER_SM = 12;      // x 4 warps = 48 persistent warps per SM

static int check_args(const xmap_sim_args &a) {
    if (a.k < 1 || a.k > KMAX) return fail_msg("xmap_sim: k out of range [1, XMAP_KMAX]");
    if (a.mode < 0 || a.mode > 2) return fail_msg("xmap_sim: bad mode");
    if (a.method != XMAP_METHOD_ADJUST_COSINE && a.method != XMAP_METHOD_COSINE)
        return fail_msg("xmap_sim: bad method");
    if (a.num_atleast < 1) return fail_msg("xmap_sim: num_atleast must be >= 1");
    return 0;
}

}  // namespace xmap

using namespace xmap;

static int gmem_log2_slots(int tier) { return tier == 3 ? 13 : (tier == 4 ? 15 : 0); }

extern "C" size_t xmap_sim_rows_workspace_bytes(int32_t tier) {
    if (tier != 3 && tier != 4) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return (size_t)sms * GMEM_CTAS_PER_SM * 4 * warp_tab_bytes(1 << gmem_log2_slots(tier)) + 256;
}

extern "C" int xmap_sim_rows(const xmap_sim_args *args_h, const int32_t *rows, int32_t n_rows, int32_t tier,
                             void *workspace, size_t workspace_bytes, void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    if (tier == 0) return launch_warp_smem<9, 4>(*args_h, rows, n_rows, st);
    if (tier == 1) return launch_warp_smem<10, 2>(*args_h, rows, n_rows, st);
    if (tier == 2) return launch_warp_smem<11, 1>(*args_h, rows, n_rows, st);
    if (tier == 3 || tier == 4) {
        if (workspace_bytes < xmap_sim_rows_workspace_bytes(tier)) return fail_msg("xmap_sim_rows: workspace too small");
        int dev = 0, sms = 148;
        XMAP_CUDA(cudaGetDevice(&dev));
        XMAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int32_t *counter = reinterpret_cast<int32_t *>(workspace);
        XMAP_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
        if (tier == 3)
            sim_warp_gmem_kernel<13><<<sms * GMEM_CTAS_PER_SM, 128, 0, st>>>(
                *args_h, rows, n_rows, reinterpret_cast<unsigned char *>(workspace) + 256, counter);
        else
            sim_warp_gmem_kernel<15><<<sms * GMEM_CTAS_PER_SM, 128, 0, st>>>(
                *args_h, rows, n_rows, reinterpret_cast<unsigned char *>(workspace) + 256, counter);
        XMAP_LAUNCH_CHECK();
        return 0;
    }
    return fail_msg("xmap_sim_rows: bad tier");
}

extern "C" int xmap_sim_big_accumulate(const xmap_sim_args *args_h, const int32_t *rows, int32_t n_rows,
                                       const int64_t *grp_off, uint64_t *table, int32_t *touched,
                                       int32_t *touched_n, int32_t *work_counter, void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    int dev = 0, sms = 148;
    XMAP_CUDA(cudaGetDevice(&dev));
    XMAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = 2;
    XMAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sim_big_accum_kernel, BIG_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    sim_big_accum_kernel<<<sms * per_sm, BIG_THREADS, 0, st>>>(*args_h, rows, n_rows, grp_off,
                                                              reinterpret_cast<ulonglong2 *>(table), touched,
                                                              touched_n, work_counter);
    XMAP_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t xmap_sim_big_scratch_bytes(int32_t n_rows, int64_t capacity) {
    return big_scratch_layout(n_rows, capacity, nullptr, nullptr);
}

extern "C" int xmap_sim_big_finalize(const xmap_sim_args *args_h, const int32_t *rows, int32_t n_rows,
                                     uint64_t *table, int32_t *touched, int32_t *touched_n,
                                     int64_t capacity, void *scratch, size_t scratch_bytes, void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    BigScratch sc;
    if (big_scratch_layout(n_rows, capacity, (char *)scratch, &sc) > scratch_bytes)
        return fail_msg("xmap_sim_big_finalize: scratch too small");
    int dev = 0, sms = 148;
    XMAP_CUDA(cudaGetDevice(&dev));
    XMAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    big_scan_kernel<<<1, 1024, 0, st>>>(touched_n, rows, args_h->row_work, n_rows, args_h->n_items, sc,
                                      args_h->error_flag);
    XMAP_LAUNCH_CHECK();
    int eval_per_sm = 3;
    XMAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eval_per_sm, big_eval_kernel, 256, 0));
    if (eval_per_sm < 1) eval_per_sm = 1;
    big_eval_kernel<<<sms * eval_per_sm, 256, 0, st>>>(*args_h, rows, n_rows,
                                                      reinterpret_cast<ulonglong2 *>(table), touched, touched_n, sc);
    XMAP_LAUNCH_CHECK();
    const size_t sel_smem = sizeof(Scratch<FIN_THREADS, SEL_CAP>);
    XMAP_CUDA(cudaFuncSetAttribute(big_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    big_select_kernel<<<n_rows, FIN_THREADS, sel_smem, st>>>(*args_h, rows, n_rows, touched_n, sc);
    XMAP_LAUNCH_CHECK();
    return 0;
}
