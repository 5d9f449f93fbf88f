// Packing / unpacking of neighbour-record lists around the multi-GPU exchange
// (the reduceByKey shuffle of baselinerSim.py:210-211, 232-233 becomes one NCCL
// all-to-all): a segmented copy of 16-byte records.  HBM-bound streaming: one
// thread per record, the segment found by a binary search over the (L2-resident)
// exclusive offsets, so a 400 K-record list costs no more per record than a
// 3-record one.
#include "common.cuh"

namespace xmap {

template <class E>
__global__ void __launch_bounds__(256) segmented_copy_kernel(const E *__restrict__ src,
                                                             const int64_t *__restrict__ src_pos,
                                                             E *__restrict__ dst,
                                                             const int64_t *__restrict__ dst_pos,
                                                             const int64_t *__restrict__ seg_off, int32_t n_seg,
                                                             int64_t total) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int32_t lo = 0, hi = n_seg;                     // largest g with seg_off[g] <= t (empty segments are skipped)
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= t) lo = mid; else hi = mid;
    }
    const int64_t k = t - __ldg(seg_off + lo);
    dst[__ldg(dst_pos + lo) + k] = __ldg(src + __ldg(src_pos + lo) + k);
}

}  // namespace xmap

using namespace xmap;

extern "C" int xmap_segmented_copy16(const void *src, const int64_t *src_pos, void *dst, const int64_t *dst_pos,
                                     const int64_t *seg_off, int32_t n_seg, int64_t total, void *stream_) {
    if (total <= 0 || n_seg <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    const int T = 256;
    segmented_copy_kernel<uint4><<<(unsigned)((total + T - 1) / T), T, 0, st>>>(
        reinterpret_cast<const uint4 *>(src), src_pos, reinterpret_cast<uint4 *>(dst), dst_pos, seg_off, n_seg, total);
    XMAP_LAUNCH_CHECK();
    return 0;
}

extern "C" int xmap_segmented_copy4(const void *src, const int64_t *src_pos, void *dst, const int64_t *dst_pos,
                                    const int64_t *seg_off, int32_t n_seg, int64_t total, void *stream_) {
    if (total <= 0 || n_seg <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    const int T = 256;
    segmented_copy_kernel<int32_t><<<(unsigned)((total + T - 1) / T), T, 0, st>>>(
        reinterpret_cast<const int32_t *>(src), src_pos, reinterpret_cast<int32_t *>(dst), dst_pos, seg_off, n_seg, total);
    XMAP_LAUNCH_CHECK();
    return 0;
}
