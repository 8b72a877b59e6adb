// SE(2) chain composition on the GPU: the serial prefix product of the odometry
// fan-out's results, pose_i = mat_to_pose(pose_to_mat(pose_{i-1}) @ T_{i-1})
// (reference scripts/main.py:249-256), as a parallel scan.  Composition of rigid transforms is
// associative, so M_i = pose_to_mat(pose_0) @ T_0 @ ... @ T_{i-1} can be built from segment products:
// every thread multiplies a contiguous run of transforms, one block-wide scan (Hillis-Steele over
// 2x3 matrices, order preserving) gives every run its prefix, and every thread walks its run again.
// The reference re-normalises the rotation at every step (atan2, then cos/sin); the scan does not,
// so the two differ by rounding that grows with the chain length (1e-12 at 5,000 steps, measured in
// tests/test_gpu_pipeline.py) -- far inside the 1e-4 m trajectory contract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icpb {

struct Se2 { double a, b, x, c, d, y; };          // [a b x; c d y; 0 0 1]

__device__ __forceinline__ Se2 se2_mul(const Se2 &L, const Se2 &R)
{
    Se2 o;
    o.a = L.a * R.a + L.b * R.c;  o.b = L.a * R.b + L.b * R.d;  o.x = L.a * R.x + L.b * R.y + L.x;
    o.c = L.c * R.a + L.d * R.c;  o.d = L.c * R.b + L.d * R.d;  o.y = L.c * R.x + L.d * R.y + L.y;
    return o;
}

constexpr int kComposeThreads = 512;

// poses_out has n + 1 rows (x, y, theta); row 0 is pose0.
__global__ void __launch_bounds__(kComposeThreads)
compose_chain_kernel(const double *T6, int64_t n, double p0x, double p0y, double p0t, double *poses_out)
{
    __shared__ Se2 s[kComposeThreads];
    const int tid = threadIdx.x;
    const int64_t L = (n + kComposeThreads - 1) / kComposeThreads;
    const int64_t i0 = min((int64_t)tid * L, n), i1 = min(i0 + L, n);
    const Se2 I = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    Se2 run = I;
    for (int64_t i = i0; i < i1; ++i) {
        const double *t = T6 + 6 * i;
        const Se2 m = {t[0], t[1], t[2], t[3], t[4], t[5]};
        run = se2_mul(run, m);
    }
    // inclusive scan of the run products, left to right
    s[tid] = run;
    __syncthreads();
    for (int o = 1; o < kComposeThreads; o <<= 1) {
        Se2 v = s[tid];
        if (tid >= o) v = se2_mul(s[tid - o], v);
        __syncthreads();
        s[tid] = v;
        __syncthreads();
    }
    double sn, cs;
    sincos(p0t, &sn, &cs);
    const Se2 P0 = {cs, -sn, p0x, sn, cs, p0y};
    Se2 acc = tid == 0 ? P0 : se2_mul(P0, s[tid - 1]);    // everything before this thread's run
    if (tid == 0) { poses_out[0] = p0x; poses_out[1] = p0y; poses_out[2] = p0t; }
    for (int64_t i = i0; i < i1; ++i) {
        const double *t = T6 + 6 * i;
        const Se2 m = {t[0], t[1], t[2], t[3], t[4], t[5]};
        acc = se2_mul(acc, m);
        double *q = poses_out + 3 * (i + 1);
        q[0] = acc.x; q[1] = acc.y; q[2] = atan2(acc.c, acc.a);
    }
}

}  // namespace icpb
