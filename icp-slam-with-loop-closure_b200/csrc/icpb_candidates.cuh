// Loop-closure candidate generation on the GPU: the front half of the reference's
// detect_proximity (reference src/loop_closure_detection.py:12-25) and its "every pair within the
// radius" generalisation (BASELINE config 3).  The reference builds the S x S distance matrix with
// scipy's cdist (800 MB at S = 10,000) and walks it row by row in Python; here one warp owns a row
// and the matrix is never materialised.  All arithmetic is fp64 with the reference's rounding:
// sqrt((dx*dx) + (dy*dy)), products and sum rounded separately.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icpb {

__device__ __forceinline__ double pose_dist(double2 a, double2 b)
{
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// np.searchsorted(travelled, v, side="right"): first index whose value is > v
__device__ __forceinline__ int upper_bound(const double *travelled, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (travelled[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One warp per pose i: j* = first index of the smallest distance among j >= start(i)
// (np.argmin's rule, :21), reported if that distance is <= max_dist (:22), else -1.
__global__ void __launch_bounds__(256)
proximity_closest_kernel(const double2 *xy, const double *travelled, int n, double min_dist_along_path,
                         double max_dist, int32_t *closest, double *closest_dist)
{
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const int start = upper_bound(travelled, n, travelled[i] + min_dist_along_path);
    const double2 p = xy[i];
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int bj = 0x7fffffff;
    for (int j = start + lane; j < n; j += 32) {
        const double d = pose_dist(p, xy[j]);
        if (d < best) { best = d; bj = j; }                 // j ascends per lane: first index kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                       // lexicographic (distance, index) minimum
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) {
        const bool ok = start < n && best <= max_dist;
        closest[i] = ok ? bj : -1;
        closest_dist[i] = ok ? best : __longlong_as_double(0x7ff0000000000000LL);
    }
}

// Every pair (i, j), j >= start(i), with distance <= max_dist.  FILL = false: count per row;
// FILL = true: write (source = j, target = i) rows at row_offset[i], in ascending j.
template <bool FILL>
__global__ void __launch_bounds__(256)
proximity_pairs_kernel(const double2 *xy, const double *travelled, int n, double min_dist_along_path,
                       double max_dist, int64_t *row_count, const int64_t *row_offset, int32_t *pairs)
{
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const int start = upper_bound(travelled, n, travelled[i] + min_dist_along_path);
    const double2 p = xy[i];
    int64_t written = 0;
    for (int j0 = start; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool in = j < n && pose_dist(p, xy[j]) <= max_dist;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (FILL && in) {
            const int64_t slot = row_offset[i] + written + __popc(m & ((1u << lane) - 1u));
            pairs[2 * slot] = j;
            pairs[2 * slot + 1] = i;
        }
        written += __popc(m);
    }
    if (!FILL && lane == 0) row_count[i] = written;
}

}  // namespace icpb
