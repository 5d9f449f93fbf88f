// Clean stage on encoded records (the data-parallel part of BaselinerClean):
//   parse_data's period test   (baselinerClean.py:47-52)  -> ts in [t_lo, t_hi)  (the host turns the local-time
//                                                            years [date_from, date_to] into the two instants)
//   filter_data                (baselinerClean.py:62-92)  -> per (user, item) the STRICTLY latest rating, the first
//                                                            seen winning ties
//   clean_data                 (baselinerClean.py:94-97)  -> users with fewer than num_atleast items are dropped
// The records arrive with `order` = their indices sorted by (in-period first, user, item), stable, so a (user, item)
// group is a run of `order` in arrival order.  One thread per sorted position decides whether its record is the group's
// winner by walking the run (groups are tiny: a user re-rating an item); a second kernel applies the per-user count.
// HBM-bound streaming work: 20 B read per record in the first kernel, 9 B in the second.
#include "common.cuh"

namespace xmap {

__global__ void __launch_bounds__(256) clean_winner_kernel(const int32_t *__restrict__ user, const int32_t *__restrict__ item,
                                                           const double *__restrict__ ts, const int64_t *__restrict__ order,
                                                           long long n, double t_lo, double t_hi,
                                                           uint8_t *__restrict__ keep, int32_t *__restrict__ user_items) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long r = order[p];
    const double t = ts[r];
    bool win = t >= t_lo && t < t_hi;
    if (win) {
        const int u = user[r], i = item[r];
        // an earlier record of the group beats this one unless this one is strictly later
        for (long long q = p - 1; q >= 0 && win; --q) {
            const long long r2 = order[q];
            const double t2 = ts[r2];
            if (user[r2] != u || item[r2] != i || !(t2 >= t_lo && t2 < t_hi)) break;
            if (t2 >= t) win = false;
        }
        // a later record of the group beats this one only if it is strictly later
        for (long long q = p + 1; q < n && win; ++q) {
            const long long r2 = order[q];
            const double t2 = ts[r2];
            if (user[r2] != u || item[r2] != i || !(t2 >= t_lo && t2 < t_hi)) break;
            if (t2 > t) win = false;
        }
        if (win) atomicAdd(user_items + u, 1);
    }
    keep[r] = win ? 1 : 0;
}

__global__ void __launch_bounds__(256) clean_users_kernel(const int32_t *__restrict__ user, long long n,
                                                          const int32_t *__restrict__ user_items, int num_atleast,
                                                          uint8_t *__restrict__ keep) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (keep[r] && user_items[user[r]] < num_atleast) keep[r] = 0;
}

}  // namespace xmap

using namespace xmap;

extern "C" int xmap_clean_records(const int32_t *user, const int32_t *item, const double *ts, const int64_t *order,
                                  int64_t n, int32_t n_users, double t_lo, double t_hi, int32_t num_atleast,
                                  uint8_t *keep, int32_t *user_items, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_users < 0 || n < 0) return fail_msg("xmap_clean_records: negative size");
    if (n_users) XMAP_CUDA(cudaMemsetAsync(user_items, 0, sizeof(int32_t) * (size_t)n_users, st));
    if (n == 0) return 0;
    const unsigned g = (unsigned)((n + 255) / 256);
    clean_winner_kernel<<<g, 256, 0, st>>>(user, item, ts, order, (long long)n, t_lo, t_hi, keep, user_items);
    XMAP_LAUNCH_CHECK();
    clean_users_kernel<<<g, 256, 0, st>>>(user, (long long)n, user_items, num_atleast, keep);
    XMAP_LAUNCH_CHECK();
    return 0;
}
