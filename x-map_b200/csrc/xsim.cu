// X-SIM bridge extension: every (left segment, bridge pair, right segment)
// combination is one path of the reference's enumeration (extender.py:124-169);
// its similarity s_p = sum(sim*mutu)/sum(mutu) and certainty c_p = prod(frac)
// (extender.py:83-89) are accumulated per (start, end) into
// xsim = sum(s_p c_p) / sum(c_p) (extender.py:198-201) without ever storing a
// path.  One warp owns one start item and walks its legs and partners in a
// fixed order, so a cell's sum depends only on the path structure (identical
// items get bit-identical X-SIM values, which keeps top-k ties deterministic).
//
// Bound by accumulator traffic: each combo reads 3 doubles + 1 int of the
// right-segment table (coalesced across the warp, L2-resident because a bridge
// source is shared by many starts) and does one hash-cell read-modify-write.
#include "common.cuh"

namespace xmap {

constexpr int XS_THREADS = 256;

// 32-byte accumulator cell, aligned to a DRAM sector, so a probe + update touches exactly one sector each
// way (a packed 24-byte cell straddles a sector boundary half of the time).  key = epoch << 32 |
// (end item + 1): a cell whose epoch differs from the launch's epoch is empty, so the workspace never has
// to be cleared between launches (it is zeroed once when allocated and every launch uses a fresh epoch).
struct __align__(32) XCell {
    unsigned long long key;
    double num, den;
    unsigned long long pad;
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// find-or-insert: returns the cell index or -1 (table full).  Only one warp (accumulate) or one
// CTA (merge) ever writes a given table; the CAS resolves lanes racing for one empty cell, and the
// thread whose CAS inserts the key initialises the values.
__device__ __forceinline__ int table_slot(XCell *tab, int hsize, unsigned long long key, unsigned epoch) {
    int slot = (int)__umulhi(((unsigned)key - 1u) * 2654435761u, (unsigned)hsize);
    for (int probe = 0; probe < hsize; ++probe) {
        unsigned long long cur = *(volatile unsigned long long *)&tab[slot].key;
        if (cur != key) {
            if ((unsigned)(cur >> 32) != epoch) {          // empty (stale epoch): try to claim it
                const unsigned long long old = atomicCAS(&tab[slot].key, cur, key);
                if (old == cur) { tab[slot].num = 0.0; tab[slot].den = 0.0; return slot; }
                cur = old;
                if (cur == key) return slot;
                if ((unsigned)(cur >> 32) != epoch) { --probe; continue; }   // changed under us but still empty: retry
            }
            slot = (slot + 1 == hsize) ? 0 : slot + 1;
            continue;
        }
        return slot;
    }
    return -1;
}

// One warp per work unit (a start item, or a slice of the legs of a heavy start).
__global__ void __launch_bounds__(XS_THREADS) xsim_accum_kernel(xmap_xsim_args a) {
    const int x = (blockIdx.x * XS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= a.n_units) return;
    const int hsize = a.hash_size[x];
    if (hsize < 32) {                                  // contract: at least 32 cells
        if (lane == 0) atomicExch(a.error_flag, 3);
        return;
    }
    XCell *tab = reinterpret_cast<XCell *>(a.hash_cells) + a.hash_off[x];
    const unsigned epoch = a.epoch;
    const unsigned long long ekey = (unsigned long long)epoch << 32;
    long long combos = 0;

    for (int64_t lg = a.unit_leg_lo[x]; lg < a.unit_leg_hi[x]; ++lg) {
        const int t = a.leg_t[lg];
        const bool joint_only = a.leg_joint_only[lg] != 0;
        // sums in path order (extender.py:85-88): left edges first
        const double Nl = __dadd_rn(a.leg_e1[lg], a.leg_e2[lg]);
        const double Dl = __dadd_rn(a.leg_m1[lg], a.leg_m2[lg]);
        const double Cl = __dmul_rn(a.leg_f1[lg], a.leg_f2[lg]);
        for (int64_t pp = a.par_ptr[t]; pp < a.par_ptr[t + 1]; ++pp) {
            if (joint_only && !a.par_joint[pp]) continue;
            const int s = a.par_s[pp];
            const double Nm = __dadd_rn(Nl, a.par_e[pp]);
            const double Dm = __dadd_rn(Dl, a.par_m[pp]);
            const double Cm = __dmul_rn(Cl, a.par_f[pp]);
            const int64_t rb = a.rs_ptr[s], re = a.rs_ptr[s + 1];
            // right-segment data of the next 32 paths is requested before the table update of the current ones
            int y_n = -1;
            double n_n = 0.0, d_n = 0.0, c_n = 0.0;
            if (rb + lane < re) {
                const int64_t r = rb + lane;
                y_n = a.rs_end[r];
                n_n = a.rs_n[r]; d_n = a.rs_d[r]; c_n = a.rs_c[r];
                prefetch_l2(&tab[__umulhi((unsigned)y_n * 2654435761u, (unsigned)hsize)]);
            }
            for (int64_t r0 = rb; r0 < re; r0 += 32) {
                const bool valid = r0 + lane < re;
                const int y = y_n;
                const double rn = n_n, rd = d_n, rc = c_n;
                y_n = -1;
                if (r0 + 32 + lane < re) {
                    const int64_t r = r0 + 32 + lane;
                    y_n = a.rs_end[r];
                    n_n = a.rs_n[r]; d_n = a.rs_d[r]; c_n = a.rs_c[r];
                    // pull the home cell of the next step's end towards L2 while this step updates the table
                    prefetch_l2(&tab[__umulhi((unsigned)y_n * 2654435761u, (unsigned)hsize)]);
                }
                double num = 0.0, den = 0.0;
                if (valid) {
                    const double Nn = __dadd_rn(Nm, rn);
                    const double Dd = __dadd_rn(Dm, rd);
                    const double cp = __dmul_rn(Cm, rc);
                    const double sp = (Dd != 0.0) ? __ddiv_rn(Nn, Dd) : 0.0;
                    num = __dmul_rn(sp, cp);
                    den = cp;
                    ++combos;
                }
                // lanes that hit the same end: the lowest lane adds the terms in lane order
                const unsigned grp = __match_any_sync(0xffffffffu, y);
                const bool leader = valid && ((__ffs(grp) - 1) == lane);
                unsigned rem = leader ? grp : 0u;
                double an = 0.0, ad = 0.0;
                while (__any_sync(0xffffffffu, rem != 0u)) {
                    const int src = rem ? (__ffs(rem) - 1) : lane;
                    const double n2 = __shfl_sync(0xffffffffu, num, src);
                    const double d2 = __shfl_sync(0xffffffffu, den, src);
                    if (rem) { an = __dadd_rn(an, n2); ad = __dadd_rn(ad, d2); rem &= rem - 1u; }
                }
                if (leader) {
                    const int slot = table_slot(tab, hsize, ekey | (unsigned)(y + 1), epoch);
                    if (slot < 0) atomicExch(a.error_flag, 2);
                    else { tab[slot].num = __dadd_rn(tab[slot].num, an); tab[slot].den = __dadd_rn(tab[slot].den, ad); }
                }
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) combos += __shfl_xor_sync(0xffffffffu, combos, off);
    if (lane == 0) a.unit_combos[x] = combos;
}

// One CTA per (dst, src) pair of a merge round: dst[y] += src[y] for every cell of src.
// The rounds form a fixed binary tree over the slices of a heavy start, so the summation order of
// every cell is a function of the path structure only.
__global__ void __launch_bounds__(XS_THREADS) xsim_merge_kernel(xmap_xsim_args a, const int32_t *__restrict__ pair_dst,
                                                                const int32_t *__restrict__ pair_src) {
    const int d = pair_dst[blockIdx.x], sx = pair_src[blockIdx.x];
    const int dsize = a.hash_size[d], ssize = a.hash_size[sx];
    XCell *dt = reinterpret_cast<XCell *>(a.hash_cells) + a.hash_off[d];
    const XCell *st = reinterpret_cast<const XCell *>(a.hash_cells) + a.hash_off[sx];
    const unsigned epoch = a.epoch;
    for (int q = threadIdx.x; q < ssize; q += XS_THREADS) {
        const unsigned long long key = st[q].key;
        if ((unsigned)(key >> 32) != epoch) continue;
        const int slot = table_slot(dt, dsize, key, epoch);
        if (slot < 0) { atomicExch(a.error_flag, 2); continue; }
        dt[slot].num = __dadd_rn(dt[slot].num, st[q].num);
        dt[slot].den = __dadd_rn(dt[slot].den, st[q].den);
    }
}

constexpr int XF_BINS = 256;
constexpr int XF_BUF = 224;                   // survivors; 224 * 12 B >= XF_BINS * 4 B

// 16 sub-bins per octave for |xsim| in [2^-16, 2) (see sim_bin in sim.cu)
__device__ __forceinline__ int xsim_bin(unsigned long long key_bits) {
    const int hi = int((key_bits & 0x7FFFFFFFFFFFFFFFull) >> 48);
    const int base = (1023 - 16) << 4;
    return max(0, min(XF_BINS - 1, hi - base));
}

// One warp per start: count the ends and pick the top-m (two passes over the table: a histogram of
// |xsim| to find the threshold bin, then the survivors), or emit every (end, xsim).
__global__ void __launch_bounds__(XS_THREADS) xsim_finalize_kernel(xmap_xsim_args a) {
    __shared__ __align__(16) unsigned char s_sel[XS_THREADS / 32][XF_BUF * 12];
    const int x = (blockIdx.x * XS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= a.n_starts) return;
    const int u = a.start_unit[x];                    // the unit whose table holds the merged result
    const int hsize = a.hash_size[u];
    const XCell *tab = reinterpret_cast<const XCell *>(a.hash_cells) + a.hash_off[u];
    const unsigned epoch = a.epoch;

    if (a.mode == 2) {  // emit every (end, xsim) of this start
        int64_t base = a.emit_ptr[x];
        int written = 0;
        for (int s0 = 0; s0 < hsize; s0 += 32) {
            XCell c{0ull, 0.0, 0.0, 0ull};
            if (s0 + lane < hsize) c = tab[s0 + lane];
            const bool occ = (unsigned)(c.key >> 32) == epoch;
            const unsigned m = __ballot_sync(0xffffffffu, occ);
            if (occ) {
                const int64_t p = base + written + __popc(m & ((1u << lane) - 1u));
                a.emit_end[p] = int((unsigned)c.key) - 1;
                a.emit_xsim[p] = __ddiv_rn(c.num, c.den);
            }
            written += __popc(m);
        }
        return;
    }

    // pass 1: count + histogram of |xsim|
    unsigned *hist = reinterpret_cast<unsigned *>(s_sel[threadIdx.x >> 5]);
    for (int b = lane; b < XF_BINS; b += 32) hist[b] = 0u;
    __syncwarp();
    int cnt = 0;
    for (int s = lane; s < hsize; s += 32) {
        const XCell c = tab[s];
        if ((unsigned)(c.key >> 32) != epoch) continue;
        ++cnt;
        atomicAdd(&hist[xsim_bin((unsigned long long)__double_as_longlong(__ddiv_rn(c.num, c.den)))], 1u);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    __syncwarp();
    if (lane == 0) a.out_count[x] = cnt;
    const int M = min(a.top_m, cnt);
    int bstar = 0;
    {
        int run = 0;
        bool found = false;
        for (int hb = XF_BINS - 32; hb >= 0 && !found && M > 0; hb -= 32) {
            unsigned suf = hist[hb + lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_down_sync(0xffffffffu, suf, off);
                if (lane + off < 32) suf += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, run + (int)suf >= M);
            if (hit) { bstar = hb + (31 - __clz(hit)); found = true; }
            else run += (int)__shfl_sync(0xffffffffu, suf, 0);
        }
    }
    __syncwarp();
    // pass 2: survivors (bin >= b*), or, if one bin holds too many equal values, plain rounds
    unsigned long long *bkey = reinterpret_cast<unsigned long long *>(s_sel[threadIdx.x >> 5]);
    int *bslot = reinterpret_cast<int *>(bkey + XF_BUF);
    int nb = 0;
    bool overflow = false;
    for (int s0 = 0; s0 < hsize && M > 0; s0 += 32) {
        XCell c{0ull, 0.0, 0.0, 0ull};
        if (s0 + lane < hsize) c = tab[s0 + lane];
        bool take = false;
        unsigned long long kk = 0ull;
        if ((unsigned)(c.key >> 32) == epoch) {
            kk = abs_key(__ddiv_rn(c.num, c.den));
            take = xsim_bin(kk) >= bstar;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (nb + __popc(m) > XF_BUF) { overflow = true; break; }
        if (take) {
            const int pos = nb + __popc(m & ((1u << lane) - 1u));
            bkey[pos] = kk; bslot[pos] = s0 + lane;
        }
        nb += __popc(m);
    }
    __syncwarp();
    unsigned long long last_k = ~0ull;
    int last_t = -1, got = 0;
    for (int r = 0; r < M; ++r) {
        unsigned long long bk = 0;
        int bt = 0x7FFFFFFF, bp = -1;
        if (!overflow) {
            for (int q = lane; q < nb; q += 32) {
                const unsigned long long kk = bkey[q];
                const int sl = bslot[q];
                const int tt = int((unsigned)tab[sl].key) - 1;
                if (r > 0 && !better(last_k, last_t, kk, tt)) continue;
                if (bp < 0 || better(kk, tt, bk, bt)) { bk = kk; bt = tt; bp = sl; }
            }
        } else {
            for (int s = lane; s < hsize; s += 32) {
                const XCell c = tab[s];
                if ((unsigned)(c.key >> 32) != epoch) continue;
                const unsigned long long kk = abs_key(__ddiv_rn(c.num, c.den));
                const int tt = int((unsigned)c.key) - 1;
                if (r > 0 && !better(last_k, last_t, kk, tt)) continue;    // strictly after the previous winner
                if (bp < 0 || better(kk, tt, bk, bt)) { bk = kk; bt = tt; bp = s; }
            }
        }
        warp_argbest(bk, bt, bp);
        if (bp < 0) break;
        if (lane == 0) {
            a.top_end[(size_t)x * a.top_m + r] = bt;
            a.top_xsim[(size_t)x * a.top_m + r] = __ddiv_rn(tab[bp].num, tab[bp].den);
        }
        last_k = bk; last_t = bt;
        got = r + 1;
    }
    if (lane == 0) a.top_len[x] = got;
}

}  // namespace xmap

using namespace xmap;

extern "C" int xmap_xsim_extend(const xmap_xsim_args *args_h, void *stream_) {
    const xmap_xsim_args &a = *args_h;
    if (a.n_starts <= 0 || a.n_units <= 0) return 0;
    if (a.top_m < 1 || a.top_m > XMAP_KMAX) return fail_msg("xmap_xsim_extend: top_m out of range");
    if (a.mode != 0 && a.mode != 2) return fail_msg("xmap_xsim_extend: bad mode");
    cudaStream_t st = (cudaStream_t)stream_;
    const int wpc = XS_THREADS / 32;
    xsim_accum_kernel<<<(unsigned)((a.n_units + wpc - 1) / wpc), XS_THREADS, 0, st>>>(a);
    XMAP_LAUNCH_CHECK();
    for (int r = 0; r < a.n_rounds; ++r) {
        const int lo = a.round_ptr_h[r], hi = a.round_ptr_h[r + 1];
        if (hi > lo) {
            xsim_merge_kernel<<<(unsigned)(hi - lo), XS_THREADS, 0, st>>>(a, a.pair_dst + lo, a.pair_src + lo);
            XMAP_LAUNCH_CHECK();
        }
    }
    xsim_finalize_kernel<<<(unsigned)((a.n_starts + wpc - 1) / wpc), XS_THREADS, 0, st>>>(a);
    XMAP_LAUNCH_CHECK();
    return 0;
}
