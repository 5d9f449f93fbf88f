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

template <int THREADS>
struct Scratch {
    static constexpr int CAP = 2 * THREADS + 2 * KMAX;
    double c_sim[CAP];
    double t_sim[2][KMAX];
    unsigned long long thr[2];
    int c_j[CAP], c_mutu[CAP], c_n[CAP];
    int t_j[2][KMAX], t_mutu[2][KMAX], t_n[2][KMAX];
    int t_len[2];
    int ncand, n_pairs, n_kept, any_label;
    unsigned char c_l0[CAP], c_l1[CAP];
};

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

// Walk raters [lo, hi) of `row` (CSC order); for every other item j in each
// rater's CSR row call add(valid, j, agree, fx) with ALL lanes converged.
template <class Add>
__device__ __forceinline__ void accumulate_raters(const xmap_sim_args &a, int row, int cls_i, int lo, int hi,
                                                  int warp, int nwarps, Add add) {
    const int lane = lane_id();
    const bool adj = (a.method == XMAP_METHOD_ADJUST_COSINE);
    const int qbase = 62 - a.r2_bits;
    for (int base = lo + warp * 32; base < hi; base += nwarps * 32) {
        // lane-parallel prefetch of 32 raters' row extents (one latency round per 32 raters)
        int e = base + lane;
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
        const int cnt = min(32, hi - base);
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
// Selection: keep the best K of a candidate buffer per list, one warp per list.
// --------------------------------------------------------------------------
template <int THREADS>
__device__ void select_lists(Scratch<THREADS> &S, int K, bool use0, bool use1) {
    const int tid = threadIdx.x;
    const int n0 = S.t_len[0], n1 = S.t_len[1];
    const int nc = S.ncand;
    // running tops re-enter as candidates of their own list
    if (tid < n0) {
        int p = nc + tid;
        S.c_sim[p] = S.t_sim[0][tid]; S.c_j[p] = S.t_j[0][tid];
        S.c_mutu[p] = S.t_mutu[0][tid]; S.c_n[p] = S.t_n[0][tid];
        S.c_l0[p] = 1; S.c_l1[p] = 0;
    } else if (tid >= KMAX && tid < KMAX + n1) {
        int q = tid - KMAX, p = nc + n0 + q;
        S.c_sim[p] = S.t_sim[1][q]; S.c_j[p] = S.t_j[1][q];
        S.c_mutu[p] = S.t_mutu[1][q]; S.c_n[p] = S.t_n[1][q];
        S.c_l0[p] = 0; S.c_l1[p] = 1;
    }
    __syncthreads();
    const int total = nc + n0 + n1;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 2 && ((warp == 0) ? use0 : use1)) {
        unsigned char *flag = (warp == 0) ? S.c_l0 : S.c_l1;
        int got = 0;
        for (int r = 0; r < K; ++r) {
            unsigned long long bk = 0;
            int bt = 0x7FFFFFFF, bp = -1;
            for (int p = lane; p < total; p += 32) {
                if (flag[p]) {
                    unsigned long long kk = abs_key(S.c_sim[p]);
                    int tt = S.c_j[p];
                    if (bp < 0 || better(kk, tt, bk, bt)) { bk = kk; bt = tt; bp = p; }
                }
            }
            warp_argbest(bk, bt, bp);
            if (bp < 0) break;
            if (lane == 0) {
                S.t_sim[warp][r] = S.c_sim[bp]; S.t_j[warp][r] = S.c_j[bp];
                S.t_mutu[warp][r] = S.c_mutu[bp]; S.t_n[warp][r] = S.c_n[bp];
                flag[bp] = 0;
            }
            __syncwarp();
            got = r + 1;
        }
        if (lane == 0) {
            S.t_len[warp] = got;
            S.thr[warp] = (got == K) ? abs_key(S.t_sim[warp][K - 1]) : 0ull;
        }
    }
    __syncthreads();
    if (tid == 0) S.ncand = 0;
    __syncthreads();
}

// Fetch(e, j, n, mutu, fx, last_sweep) -> bool (false: empty entry)
template <int THREADS, class Fetch>
__device__ void finalize_row(const xmap_sim_args &a, const RowCtx &c, int n_entries, Fetch fetch,
                             Scratch<THREADS> &S) {
    const int tid = threadIdx.x;
    const int K = a.k;
    const int row = c.row;
    if (tid == 0) {
        S.ncand = 0; S.n_pairs = 0; S.n_kept = 0; S.any_label = 0;
        S.t_len[0] = S.t_len[1] = 0; S.thr[0] = S.thr[1] = 0ull;
    }
    __syncthreads();

    if (a.mode == 2) {  // emit every kept pair (materialised sim RDD, assist.py:75-77)
        const int64_t base = a.emit_ptr[row];
        for (int e = tid; e < n_entries; e += THREADS) {
            int j, n, mutu, label; long long fx; double sim;
            if (!fetch(e, j, n, mutu, fx, true)) continue;
            if (!eval_pair(a, c, j, n, mutu, fx, sim, label)) continue;
            int pos = atomicAdd(&S.n_kept, 1);
            a.emit_j[base + pos] = j; a.emit_sim[base + pos] = sim;
            a.emit_mutu[base + pos] = mutu; a.emit_n[base + pos] = n;
        }
        return;
    }

    bool bb = false;
    if (a.mode == 0) {  // sweep 1: counts + is this a bridge item (assist.py:84-86)
        int lp = 0, lk = 0, ll = 0;
        for (int e = tid; e < n_entries; e += THREADS) {
            int j, n, mutu, label; long long fx; double sim;
            if (!fetch(e, j, n, mutu, fx, false)) continue;
            ++lp;
            if (eval_pair(a, c, j, n, mutu, fx, sim, label)) { ++lk; ll |= label; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lp += __shfl_xor_sync(0xffffffffu, lp, off);
            lk += __shfl_xor_sync(0xffffffffu, lk, off);
            ll |= __shfl_xor_sync(0xffffffffu, ll, off);
        }
        if ((tid & 31) == 0) {
            if (lp) atomicAdd(&S.n_pairs, lp);
            if (lk) atomicAdd(&S.n_kept, lk);
            if (ll) atomicOr(&S.any_label, 1);
        }
        __syncthreads();
        bb = S.any_label != 0;
        if (tid == 0) {
            a.row_flags[row] = bb ? 1 : 0;
            a.row_npairs[row] = S.n_pairs;
            a.row_nkept[row] = S.n_kept;
        }
    }
    // list definitions (extender.py:30-43)
    const bool use0 = (a.mode == 1) || bb;
    const bool use1 = (a.mode == 0);
    for (int base = 0; base < n_entries; base += THREADS) {
        if (S.ncand > THREADS) select_lists<THREADS>(S, K, use0, use1);
        const int e = base + tid;
        int j, n, mutu, label; long long fx; double sim;
        if (e < n_entries && fetch(e, j, n, mutu, fx, true) && eval_pair(a, c, j, n, mutu, fx, sim, label)) {
            bool l0, l1;
            if (a.mode == 1) { l0 = a.bb_in[j] != 0; l1 = false; }
            else if (bb) { bool same = (a.contains[j] >> c.dom_i) & 1; l0 = !same; l1 = same; }
            else { l0 = false; l1 = true; }
            const unsigned long long key = abs_key(sim);
            if (l0 && S.thr[0] && key < S.thr[0]) l0 = false;
            if (l1 && S.thr[1] && key < S.thr[1]) l1 = false;
            if (l0 || l1) {
                int p = atomicAdd(&S.ncand, 1);
                S.c_sim[p] = sim; S.c_j[p] = j; S.c_mutu[p] = mutu; S.c_n[p] = n;
                S.c_l0[p] = l0; S.c_l1[p] = l1;
            }
        }
        __syncthreads();
    }
    select_lists<THREADS>(S, K, use0, use1);
    // write tables [n_items][2][K]
    for (int slot = 0; slot < 2; ++slot) {
        const bool wr = (slot == 0) ? (a.mode == 1 || a.mode == 0) : (a.mode == 0);
        if (!wr) continue;
        const int len = ((slot == 0) ? use0 : use1) ? S.t_len[slot] : 0;
        const size_t o = ((size_t)row * 2 + slot) * K;
        if (tid < len) {
            a.tab_idx[o + tid] = S.t_j[slot][tid]; a.tab_sim[o + tid] = S.t_sim[slot][tid];
            a.tab_mutu[o + tid] = S.t_mutu[slot][tid]; a.tab_n[o + tid] = S.t_n[slot][tid];
        }
        if (tid == 0) a.tab_len[(size_t)row * 2 + slot] = len;
    }
}

// --------------------------------------------------------------------------
// Tier 0 / 1: one CTA per row, accumulators in a shared-memory hash table.
// --------------------------------------------------------------------------
template <int LOG2_SLOTS, int THREADS>
__global__ void __launch_bounds__(THREADS) sim_hash_kernel(xmap_sim_args a, const int32_t *__restrict__ rows,
                                                            int n_rows) {
    constexpr int SLOTS = 1 << LOG2_SLOTS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *h_inner = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *h_key = reinterpret_cast<unsigned *>(h_inner + SLOTS);
    unsigned *h_cnt = h_key + SLOTS;
    Scratch<THREADS> &S = *reinterpret_cast<Scratch<THREADS> *>(h_cnt + SLOTS);

    const int tid = threadIdx.x;
    const int row = rows[blockIdx.x];
    const RowCtx c = make_ctx(a, row);
    const int lo = a.csc_ptr[row], hi = a.csc_ptr[row + 1];
    long long w = a.row_work[row];
    int log2n = 6;
    while (log2n < LOG2_SLOTS && (1LL << log2n) < 2 * w) ++log2n;
    const int nslots = 1 << log2n;
    const unsigned mask = nslots - 1;
    const int shift = 32 - log2n;
    for (int s = tid; s < nslots; s += THREADS) { h_key[s] = 0u; h_cnt[s] = 0u; h_inner[s] = 0ull; }
    __syncthreads();

    auto add = [&](bool valid, int j, unsigned agree, long long fx) {
        if (!valid) return;
        const unsigned key = (unsigned)j + 1u;
        unsigned slot = ((unsigned)j * 2654435761u) >> shift;
        for (int probe = 0; probe < nslots; ++probe) {
            unsigned cur = *(volatile unsigned *)&h_key[slot];
            if (cur != key) {
                if (cur == 0u) cur = atomicCAS(&h_key[slot], 0u, key);
                if (cur != 0u && cur != key) { slot = (slot + 1) & mask; continue; }
            }
            atomicAdd(&h_cnt[slot], (1u << 16) | agree);
            atomicAdd(&h_inner[slot], (unsigned long long)fx);
            return;
        }
        atomicExch(a.error_flag, 1);
    };
    accumulate_raters(a, row, c.cls_i, lo, hi, tid >> 5, THREADS >> 5, add);
    __syncthreads();

    auto fetch = [&](int e, int &j, int &n, int &mutu, long long &fx, bool) -> bool {
        unsigned key = h_key[e];
        if (key == 0u) return false;
        j = int(key - 1u);
        unsigned cn = h_cnt[e];
        n = int(cn >> 16); mutu = int(cn & 0xFFFFu);
        fx = (long long)h_inner[e];
        return true;
    };
    finalize_row<THREADS>(a, c, nslots, fetch, S);
}

// --------------------------------------------------------------------------
// Heavy rows: chunks of raters -> dense per-row table with 64-bit atomics.
// --------------------------------------------------------------------------
constexpr int BIG_THREADS = 512;

__global__ void __launch_bounds__(BIG_THREADS) sim_big_accum_kernel(
    xmap_sim_args a, const int32_t *__restrict__ chunk_slot, const int32_t *__restrict__ chunk_row,
    const int32_t *__restrict__ chunk_lo, const int32_t *__restrict__ chunk_hi, int n_chunks,
    ulonglong2 *__restrict__ table, int32_t *__restrict__ touched, int32_t *__restrict__ touched_n,
    int32_t *__restrict__ work_counter) {
    __shared__ int s_chunk;
    const int tid = threadIdx.x, lane = tid & 31;
    while (true) {
        if (tid == 0) s_chunk = atomicAdd(work_counter, 1);
        __syncthreads();
        const int cidx = s_chunk;
        __syncthreads();
        if (cidx >= n_chunks) break;
        const int b = chunk_slot[cidx], row = chunk_row[cidx];
        const int lo = chunk_lo[cidx], hi = chunk_hi[cidx];
        const int cls_i = ceil_log2_u32((uint32_t)a.item_stats[4 * (size_t)row + 3]);
        ulonglong2 *T = table + (size_t)b * a.n_items;
        int32_t *tl = touched + (size_t)b * a.n_items;
        int32_t *tn = touched_n + b;
        auto add = [&](bool valid, int j, unsigned agree, long long fx) {
            bool first = false;
            if (valid) {
                unsigned long long old = atomicAdd(&T[j].x, (1ull << 32) | (unsigned long long)agree);
                atomicAdd(&T[j].y, (unsigned long long)fx);
                first = (old == 0ull);
            }
            const unsigned m = __ballot_sync(0xffffffffu, first);
            if (m) {
                const int leader = __ffs(m) - 1;
                int basep = 0;
                if (lane == leader) basep = atomicAdd(tn, __popc(m));
                basep = __shfl_sync(0xffffffffu, basep, leader);
                if (first) tl[basep + __popc(m & ((1u << lane) - 1u))] = j;
            }
        };
        accumulate_raters(a, row, cls_i, lo, hi, tid >> 5, BIG_THREADS >> 5, add);
    }
}

constexpr int FIN_THREADS = 256;

__global__ void __launch_bounds__(FIN_THREADS) sim_big_finalize_kernel(
    xmap_sim_args a, const int32_t *__restrict__ rows, int n_rows, ulonglong2 *__restrict__ table,
    int32_t *__restrict__ touched, int32_t *__restrict__ touched_n) {
    __shared__ Scratch<FIN_THREADS> S;
    const int b = blockIdx.x;
    const int row = rows[b];
    const RowCtx c = make_ctx(a, row);
    ulonglong2 *T = table + (size_t)b * a.n_items;
    const int32_t *tl = touched + (size_t)b * a.n_items;
    const int n_entries = touched_n[b];
    auto fetch = [&](int e, int &j, int &n, int &mutu, long long &fx, bool last) -> bool {
        j = tl[e];
        ulonglong2 cell = T[j];
        n = int(cell.x >> 32); mutu = int(cell.x & 0xFFFFFFFFull);
        fx = (long long)cell.y;
        if (last) T[j] = make_ulonglong2(0ull, 0ull);
        return true;
    };
    finalize_row<FIN_THREADS>(a, c, n_entries, fetch, S);
    __syncthreads();
    if (threadIdx.x == 0) touched_n[b] = 0;
}

template <int LOG2_SLOTS, int THREADS>
static int launch_hash(const xmap_sim_args &a, const int32_t *rows, int n_rows, cudaStream_t st) {
    size_t smem = (size_t)(1 << LOG2_SLOTS) * 16 + sizeof(Scratch<THREADS>);
    auto kern = sim_hash_kernel<LOG2_SLOTS, THREADS>;
    XMAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_rows, THREADS, smem, st>>>(a, rows, n_rows);
    XMAP_LAUNCH_CHECK();
    return 0;
}

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

extern "C" int xmap_sim_rows_smem(const xmap_sim_args *args_h, const int32_t *rows, int32_t n_rows,
                                  int32_t tier, void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    if (tier == 0) return launch_hash<11, 256>(*args_h, rows, n_rows, st);
    if (tier == 1) return launch_hash<13, 1024>(*args_h, rows, n_rows, st);
    return fail_msg("xmap_sim_rows_smem: bad tier");
}

extern "C" int xmap_sim_big_accumulate(const xmap_sim_args *args_h, const int32_t *chunk_slot,
                                       const int32_t *chunk_row, const int32_t *chunk_lo,
                                       const int32_t *chunk_hi, int32_t n_chunks, uint64_t *table,
                                       int32_t *touched, int32_t *touched_n, int32_t *work_counter,
                                       void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_chunks <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    int dev = 0, sms = 148;
    XMAP_CUDA(cudaGetDevice(&dev));
    XMAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = 2;
    XMAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sim_big_accum_kernel, BIG_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > n_chunks) grid = n_chunks;
    sim_big_accum_kernel<<<grid, BIG_THREADS, 0, st>>>(*args_h, chunk_slot, chunk_row, chunk_lo, chunk_hi,
                                                      n_chunks, reinterpret_cast<ulonglong2 *>(table),
                                                      touched, touched_n, work_counter);
    XMAP_LAUNCH_CHECK();
    return 0;
}

extern "C" int xmap_sim_big_finalize(const xmap_sim_args *args_h, const int32_t *rows, int32_t n_rows,
                                     uint64_t *table, int32_t *touched, int32_t *touched_n, void *stream_) {
    if (int rc = check_args(*args_h)) return rc;
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    sim_big_finalize_kernel<<<n_rows, FIN_THREADS, 0, st>>>(*args_h, rows, n_rows,
                                                            reinterpret_cast<ulonglong2 *>(table), touched,
                                                            touched_n);
    XMAP_LAUNCH_CHECK();
    return 0;
}
