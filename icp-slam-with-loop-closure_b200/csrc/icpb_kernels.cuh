// Batched 2-D point-to-point ICP for sm_100a: one CTA per scan pair, the whole
// iterate-until-converged loop of the reference's icp() (reference src/icp.py:72-97) in-kernel.
//
// Numerical contract (DESIGN.md "Exactness"):
//   * the O(N1*N2) nearest-neighbour sweep runs in fp32 and only *filters*: per source point it
//     yields the best 16-target chunk, the best fp32 distance m1 and the best fp32 distance m2
//     found in any other chunk;
//   * every target whose fp32 distance is within a rigorous rounding bound of m1 is then
//     re-evaluated in fp64 with the reference's own arithmetic ((dx*dx)+(dy*dy), separately
//     rounded, src/icp.py:6) and the winner is the lexicographic (distance, index) minimum,
//     i.e. np.argmin's first-index rule (src/icp.py:7).  If another chunk is within the bound
//     (m2 <= m1 + tol) the whole target is re-filtered.  So the correspondences are those of an
//     all-fp64 search, and fp32 only decides how much fp64 work is needed;
//   * transform application, centroids, cross-covariance, error, composition and the stop
//     rules are fp64, reduced in a fixed order (deterministic).
//   * Exact pruning: 16-target chunks carry a bounding circle; a warp skips a chunk only when the
//     triangle inequality proves that every target in it is farther from every one of the warp's
//     128 source points than that point's current upper bound (its filter distance to the
//     previous pass's match) plus the rounding bound.  Skipped targets can therefore never be a
//     candidate of the exact decision, so the result is that of the exhaustive search.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "icpb.h"

namespace icpb {

constexpr int   kChunk    = 16;        // targets per bookkeeping chunk
constexpr float kPadCoord = 1.0e15f;   // coordinates of padding targets (distance ~2e30, never a candidate)
constexpr int   kMaxWarps = 32;
constexpr int   kNumSums  = 9;

struct KernelArgs {
    const double  *xy;        // scan table, (sum m_i, 2) fp64
    const int64_t *offsets;   // CSR offsets, n_scans + 1
    const int32_t *pairs;     // B x 2 or nullptr (pair_mode 1)
    const double  *init;      // B x 6 or nullptr
    int64_t        B;
    int64_t        n_scans;
    icpb_params    p;
    double        *T_out;     // B x 6
    double        *err_out;   // B
    int32_t       *passes_out;// B
    double        *hist;      // B x hist_cap x 6 or nullptr
    int32_t       *corr;      // B x corr_stride or nullptr
    unsigned long long *queue;// work-queue counter (zeroed before launch)
    int32_t        n2pad_cap; // floats per target coordinate array in shared memory
    int32_t        n1_cap;    // int32 slots for correspondences in shared memory
    int32_t        nchunk_cap;// chunk bounding circles in shared memory
    int32_t        ntile_cap; // source tiles (32*R points) whose partial sums live in shared memory
    unsigned long long *executed; // optional: += distance evaluations actually executed
    // streaming upload (icpb_align_host): pair b may start once *arrived > seg_of_pair[b], i.e. the
    // copy engine has delivered the scan-table segment holding the later of its two scans
    const int32_t *seg_of_pair;      // by queue position
    const volatile int32_t *arrived;
    const int32_t *order;            // queue position -> pair id (nullptr: identity)
    int32_t *upload_timeout;         // set to 1 if a wait on `arrived` gave up (the host then reruns the batch)
    // fused gather (multi-GPU): every finished pair's 8-double constraint record
    // [T(6), error, passes] is stored straight into every rank's gather buffer over NVLink peer
    // memory, at row rec_row0 + pair id; peers[r] is rank r's buffer as mapped in this process
    double *const *peers;
    int32_t n_peers;
    int64_t rec_row0;
};

// ---- fp32 filter distance: one definition, used by the sweep and by the refine step ----------
__device__ __forceinline__ float dist32(float px, float py, float qx, float qy)
{
    const float dx = __fsub_rn(qx, px);
    const float dy = __fsub_rn(qy, py);
    return __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
}

// ---- fp64 distance with the reference's rounding: (dx*dx) + (dy*dy), no contraction ----------
__device__ __forceinline__ double dist64(double px, double py, double qx, double qy)
{
    const double dx = __dsub_rn(qx, px);
    const double dy = __dsub_rn(qy, py);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// p' = T p, accumulated in the order a k-loop over the three homogeneous columns does
// (src/icp.py:62: np.dot(previous_transform, pc1.T).T).
__device__ __forceinline__ void apply_T(const double *T, double x, double y, double &ox, double &oy)
{
    ox = fma(T[1], y, T[0] * x) + T[2];
    oy = fma(T[4], y, T[3] * x) + T[5];
}

// Packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): two targets per instruction, so the sweep
// needs 2 issue slots per distance instead of 4.  Each half rounds exactly like the scalar
// __fsub_rn / __fmul_rn / __fmaf_rn in dist32 (IEEE round-to-nearest, no flush), so the sweep
// and the refine step see bit-identical filter distances.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// filter distances of one source point to four targets (x0..x3, y0..y3)
__device__ __forceinline__ void dist32x4(u64 PX, u64 PY, const float4 &X, const float4 &Y, float *d)
{
    const u64 dxa = sub2(pack2(X.x, X.y), PX), dya = sub2(pack2(Y.x, Y.y), PY);
    const u64 dxb = sub2(pack2(X.z, X.w), PX), dyb = sub2(pack2(Y.z, Y.w), PY);
    unpack2(fma2(dya, dya, mul2(dxa, dxa)), d[0], d[1]);
    unpack2(fma2(dyb, dyb, mul2(dxb, dxb)), d[2], d[3]);
}

__device__ __forceinline__ float min3f(float a, float b, float c)
{
    return fminf(fminf(a, b), c);
}

// MUFU square root (1-2 ulp); every use below carries a much larger safety factor
__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Bound on |fp32 filter distance - exact distance| doubled, as a function of the best filter
// distance m1.  e bounds the error of one coordinate difference: both inputs were rounded to
// fp32 (relative 2^-24 each) and the subtraction rounds once more.
__device__ __forceinline__ float filter_tol(float m1, float px, float py, float qmax)
{
    const float u = 5.9604645e-8f;                              // 2^-24
    const float e = 2.0f * u * (fmaxf(fabsf(px), fabsf(py)) + qmax) * 1.0001f;
    return 8.0f * sqrt_fast(m1) * e + 16.0f * e * e + 16.0f * u * m1;
}

// Rare path of the decision step: all targets j in [lo, hi) whose filter distance is <= thr are
// evaluated exactly; keeps the lexicographic (distance, index) minimum (j ascends, so strict <
// keeps the first index = np.argmin's rule).
__device__ __forceinline__ void exact_range(int lo, int hi, float thr, float px, float py, double Px, double Py,
                                            const float *tqx, const float *tqy, const double2 *dst,
                                            double &best, int &idx)
{
    for (int j = lo; j < hi; ++j) {
        const float d = dist32(px, py, tqx[j], tqy[j]);
        if (d <= thr) {
            const double2 q = dst[j];
            const double D = dist64(Px, Py, q.x, q.y);
            if (D < best) { best = D; idx = j; }
        }
    }
}

// Not inlined: these run for a fraction of a percent of the points and would otherwise be
// replicated per register-tiled point.
// (a) several candidates inside the best chunk: bit k of `cand` marks target j0 + k
__device__ __noinline__ int exact_decide_chunk(int j0, unsigned cand, double Px, double Py, const double2 *dst)
{
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int idx = j0;
    while (cand) {                                   // ascending k: strict < keeps the first index
        const int j = j0 + __ffs(cand) - 1;
        cand &= cand - 1;
        const double2 q = dst[j];
        const double D = dist64(Px, Py, q.x, q.y);
        if (D < best) { best = D; idx = j; }
    }
    return idx;
}
// (b) another chunk is within the bound: every chunk whose circle reaches within sqrt(thr) of the
// point may hold a candidate (chunks ascend, so the first-index rule still holds across chunks)
__device__ __noinline__ int exact_decide_all(int nchunks, int n2, int fallback, float thr, float px, float py,
                                             double Px, double Py, const float *tqx, const float *tqy,
                                             const float4 *cb, const double2 *dst)
{
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int idx = fallback;
    const float s = sqrt_fast(thr) * 1.0001f + 1e-30f;
    for (int c = 0; c < nchunks; ++c) {
        const float4 b = cb[c];
        const float lim = (s + b.z) * 1.0001f;
        if (dist32(px, py, b.x, b.y) <= lim * lim)
            exact_range(c * kChunk, min(c * kChunk + kChunk, n2), thr, px, py, Px, Py, tqx, tqy, dst, best, idx);
    }
    return idx;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// linear index over the strict upper triangle of an n x n matrix, row-major -> (i, j), i < j
__device__ __forceinline__ void decode_pair(int64_t k, int64_t n, int32_t &i_out, int32_t &j_out)
{
    const double nn = (double)n - 0.5;
    int64_t i = (int64_t)(nn - sqrt(nn * nn - 2.0 * (double)k));
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && i * (2 * n - i - 1) / 2 > k) --i;
    while ((i + 1) * (2 * n - i - 2) / 2 <= k) ++i;
    const int64_t start = i * (2 * n - i - 1) / 2;
    i_out = (int32_t)i;
    j_out = (int32_t)(k - start + i + 1);
}

// Warp-wide float max/min through the integer REDUX unit: one instruction instead of a 5-level
// shuffle tree.  Floats are mapped to unsigned keys whose order equals the float order.
__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ float warp_max(float v) { return key2f(__reduce_max_sync(0xffffffffu, f2key(v))); }
__device__ __forceinline__ float warp_min(float v) { return key2f(__reduce_min_sync(0xffffffffu, f2key(v))); }
// non-negative inputs: the bit pattern is already ordered
__device__ __forceinline__ float warp_max_nonneg(float v)
{
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

// Butterfly reduction of 8 doubles across the warp: each level halves the number of values a lane
// carries (lanes exchange the half they do not keep), then two plain levels finish.  9 double
// shuffles instead of 40.  Afterwards lane L holds the warp total of value
// id = 4*bit4(L) + 2*bit3(L) + bit2(L).  The addition order is fixed, so the result is
// deterministic.
__device__ __forceinline__ double warp_sum8(const double (&v)[8], int lane)
{
    double w4[4], w2[2], w1;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double send = up ? v[j] : v[j + 4];
            const double keep = up ? v[j + 4] : v[j];
            w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double send = up ? w4[j] : w4[j + 2];
            const double keep = up ? w4[j + 2] : w4[j];
            w2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
        const double send = up ? w2[0] : w2[1];
        const double keep = up ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    return w1;
}

// PRUNE = false: exhaustive sweep over every chunk (the reference's brute force; used for the
// FP32-pipe roofline characterisation).  PRUNE = true: exact chunk pruning (the product default).
//
// Shared memory: target SoA | chunk circles | correspondences | red[2][tiles][9] | Tw[warps][6]
// A tile = 32*R consecutive source points = one warp's register tile.  Warps pull tiles from a
// shared counter (tiles differ in how many chunks survive pruning); partial sums are stored per
// tile and folded in tile order, so the result does not depend on which warp ran which tile.
//
// CLUSTER = true (latency mode, few problems): one thread-block *cluster* of up to 8 CTAs per scan
// pair.  Every CTA stages the target itself; the source tiles are dealt round-robin to the CTAs
// (tile t belongs to CTA t mod cluster size); the per-tile partial sums stay in their owner's
// shared memory and every CTA folds them over distributed shared memory after the pass's
// cluster barrier -- in the same tile order as the single-CTA kernel, so the two give the same bits.
template <int R, bool PRUNE, bool CLUSTER>
#ifndef ICPB_MIN_CTAS
#define ICPB_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(256, (R >= 4 ? 2 : ICPB_MIN_CTAS))
icp_align_kernel(const KernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float  *tqx    = reinterpret_cast<float *>(smem_raw);
    float  *tqy    = tqx + a.n2pad_cap;
    float4 *cb     = reinterpret_cast<float4 *>(tqy + a.n2pad_cap);          // chunk circle (cx, cy, r, -)
    int    *corr_s = reinterpret_cast<int *>(cb + a.nchunk_cap);
    double *red    = reinterpret_cast<double *>(corr_s + ((a.n1_cap + 3) & ~3));   // [2][ntile_cap][9]
    double *Tw     = red + 2 * a.ntile_cap * kNumSums;                       // [kMaxWarps][6] per-warp copy of T
    __shared__ long long s_pid;
    __shared__ unsigned int s_qmax_bits;
    __shared__ int s_tile_ctr[2];

    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const float kInf = __int_as_float(0x7f800000);
    double *Tmine = Tw + warp * 6;
    unsigned int executed = 0;
    namespace cg = cooperative_groups;
    int crank = 0, csize = 1;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        crank = (int)cluster.block_rank();
        csize = (int)cluster.num_blocks();
    }
    // block-wide barrier, cluster-wide in latency mode
    auto sync_all = [&]() {
        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
    };

    for (;;) {
        // ---------------- pop a problem ----------------
        if (tid == 0) {
            if (!CLUSTER) {
                s_pid = (long long)atomicAdd(a.queue, 1ULL);
            } else if (crank == 0) {                       // one pop per cluster, broadcast over DSMEM
                const long long v = (long long)atomicAdd(a.queue, 1ULL);
                cg::cluster_group cluster = cg::this_cluster();
                for (int r = 0; r < csize; ++r) *cluster.map_shared_rank(&s_pid, r) = v;
            }
            s_qmax_bits = 0u;
            s_tile_ctr[0] = 0; s_tile_ctr[1] = 0;
        }
        sync_all();
        const int64_t qpos = s_pid;
        if (qpos >= a.B) break;
        const int64_t pid = a.order ? (int64_t)a.order[qpos] : qpos;
        if (a.arrived) {
            // wait for the copy engine (a DMA on another stream, not another kernel): the queue hands
            // pairs out in arrival order, so only the CTAs at the front of the upload ever wait
            if (tid == 0) {
                const int need = a.seg_of_pair[qpos];
                const long long t0 = clock64();
                while (*a.arrived <= need) {
                    __nanosleep(256);
                    // never hang the GPU on a copy that does not come (a profiler that serialises the
                    // kernel against the copy stream, a failed transfer): give up after ~0.5 s, once for
                    // the whole grid, and tell the host, which reruns the batch on the resident table
                    if (*(volatile int32_t *)a.upload_timeout) break;
                    if (clock64() - t0 > (1LL << 30)) { *(volatile int32_t *)a.upload_timeout = 1; break; }
                }
                __threadfence();                                 // the scans are read after the counter
            }
            __syncthreads();
        }

        int32_t sid, did;
        if (a.p.pair_mode == 1) {
            const int64_t k = a.p.k_first + (pid / a.p.k_block) * a.p.k_stride + (pid % a.p.k_block);
            int32_t i, j;
            decode_pair(k, a.n_scans, i, j);
            sid = j; did = i;
        } else {
            sid = a.pairs[2 * pid]; did = a.pairs[2 * pid + 1];
        }
        const int64_t so = a.offsets[sid], dof = a.offsets[did];
        const int n1 = (int)(a.offsets[sid + 1] - so);
        const int n2 = (int)(a.offsets[did + 1] - dof);
        const double2 *src = reinterpret_cast<const double2 *>(a.xy) + so;
        const double2 *dst = reinterpret_cast<const double2 *>(a.xy) + dof;
        const int n2pad = (n2 + kChunk - 1) / kChunk * kChunk;
        const int nchunks = n2pad / kChunk;
        const double inv_n = 1.0 / (double)n1;
        // reduction tiles are always 64 points (32 lanes x 2), whatever R: a warp work item covers
        // SUBT = R/2 consecutive reduction tiles, so kernels with different R add in the same order
        constexpr int SUBT = R / 2;
        static_assert(R % 2 == 0, "R must be even");
        const int ntiles = (n1 + 63) / 64;                 // reduction tiles
        const int nwork = (ntiles + SUBT - 1) / SUBT;            // warp work items

        // ---------------- stage the target in shared memory as fp32 SoA ----------------
        {
            float qm = 0.0f;
            for (int j = tid; j < n2pad; j += NT) {
                float x = kPadCoord, y = kPadCoord;
                if (j < n2) {
                    const double2 q = dst[j];
                    x = (float)q.x; y = (float)q.y;
                    qm = fmaxf(qm, fmaxf(fabsf(x), fabsf(y)));
                }
                tqx[j] = x; tqy[j] = y;
            }
            qm = warp_max_nonneg(qm);
            if (lane == 0) atomicMax(&s_qmax_bits, __float_as_uint(qm));   // qm >= 0: bit order = value order
        }
        if (lane < 6) {
            double v = a.init ? a.init[6 * pid + lane] : ((lane == 0 || lane == 4) ? 1.0 : 0.0);
            if (a.p.rotation_only && (lane == 2 || lane == 5)) v = 0.0;     // src/icp.py:60-61
            Tmine[lane] = v;
        }
        __syncthreads();
        const float qmax = __uint_as_float(s_qmax_bits);
        // ---------------- bounding circle of every 16-target chunk ----------------
        for (int c = tid; c < nchunks; c += NT) {
            const int j0 = c * kChunk, j1 = min(j0 + kChunk, n2);
            float lx = kInf, ly = kInf, hx = -kInf, hy = -kInf;
            for (int j = j0; j < j1; ++j) {
                lx = fminf(lx, tqx[j]); hx = fmaxf(hx, tqx[j]);
                ly = fminf(ly, tqy[j]); hy = fmaxf(hy, tqy[j]);
            }
            const float cx = 0.5f * (lx + hx), cy = 0.5f * (ly + hy);
            float r2 = 0.0f;
            for (int j = j0; j < j1; ++j) r2 = fmaxf(r2, dist32(cx, cy, tqx[j], tqy[j]));
            cb[c] = make_float4(cx, cy, sqrtf(r2) * 1.00001f, 0.0f);
        }
        const double2 g = dst[0];                 // shifts for the one-pass covariance sums
        const double2 s0 = src[0];
        __syncthreads();

        int passes = 0, iteration = 0;
        bool have_last = false;
        double last_err = 0.0, err = 0.0;

        for (;;) {
            double *redp = red + (passes & 1) * a.ntile_cap * kNumSums;

            for (;;) {
                int tile = 0;
                if (lane == 0) tile = atomicAdd(&s_tile_ctr[passes & 1], 1);
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (CLUSTER) tile = crank + tile * csize;        // this CTA owns items crank, crank + csize, ...
                if (tile >= nwork) break;
                // point r of this lane: reduction tile tile*SUBT + r/2, lane's pair of consecutive points
                auto pidx = [&](int r) { return ((tile * SUBT + (r >> 1)) * 32 + lane) * 2 + (r & 1); };
                // ---- transform, upper bounds, tile bounding circle ----
                float px[R], py[R];
                float ubmax = 0.0f, lx = kInf, ly = kInf, hx = -kInf, hy = -kInf;
                {
                    double T[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) T[k] = Tmine[k];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = min(pidx(r), n1 - 1);      // lanes past the end repeat the last point
                        const double2 s = src[i];
                        double X, Y;
                        apply_T(T, s.x, s.y, X, Y);
                        px[r] = (float)X; py[r] = (float)Y;
                    }
                }
                float tcx = 0.f, tcy = 0.f, reach = kInf;
                if (PRUNE) {
                    // bounding circle of the tile's points
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        lx = fminf(lx, px[r]); hx = fmaxf(hx, px[r]);
                        ly = fminf(ly, py[r]); hy = fmaxf(hy, py[r]);
                    }
                    lx = warp_min(lx); ly = warp_min(ly); hx = warp_max(hx); hy = warp_max(hy);
                    tcx = 0.5f * (lx + hx); tcy = 0.5f * (ly + hy);
                    float rho2 = 0.0f;
#pragma unroll
                    for (int r = 0; r < R; ++r) rho2 = fmaxf(rho2, dist32(tcx, tcy, px[r], py[r]));
                    rho2 = warp_max_nonneg(rho2);
                    // upper bound of every point's nearest-neighbour filter distance: its distance to
                    // a real target -- the previous pass's match, or on the first pass the best
                    // target in the chunks around the chunk centre nearest to the tile centre
                    float ub[R];
                    if (passes > 0) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int j = corr_s[min(pidx(r), n1 - 1)];
                            ub[r] = dist32(px[r], py[r], tqx[j], tqy[j]);
                        }
                    } else {
                        unsigned key = 0xffffffffu;
                        for (int c = lane; c < nchunks; c += 32) {
                            const float4 bc = cb[c];
                            const unsigned k = (__float_as_uint(dist32(tcx, tcy, bc.x, bc.y)) & 0xfffff000u) | (unsigned)min(c, 4095);
                            key = min(key, k);
                        }
                        const int cstar = (int)(__reduce_min_sync(0xffffffffu, key) & 0xfffu);
#pragma unroll
                        for (int r = 0; r < R; ++r) ub[r] = kInf;
                        const float4 *qx4 = reinterpret_cast<const float4 *>(tqx);
                        const float4 *qy4 = reinterpret_cast<const float4 *>(tqy);
                        for (int c = max(cstar - 1, 0); c <= min(cstar + 1, nchunks - 1); ++c) {
                            float4 X[4], Y[4];
#pragma unroll
                            for (int v = 0; v < 4; ++v) { X[v] = qx4[4 * c + v]; Y[v] = qy4[4 * c + v]; }
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                float d[16];
                                const u64 PX = pack2(px[r], px[r]), PY = pack2(py[r], py[r]);
#pragma unroll
                                for (int v = 0; v < 4; ++v) dist32x4(PX, PY, X[v], Y[v], d + 4 * v);
                                float cm = min3f(d[0], d[1], d[2]);
                                cm = min3f(cm, d[3], d[4]);   cm = min3f(cm, d[5], d[6]);
                                cm = min3f(cm, d[7], d[8]);   cm = min3f(cm, d[9], d[10]);
                                cm = min3f(cm, d[11], d[12]); cm = min3f(cm, d[13], d[14]);
                                ub[r] = fminf(ub[r], fminf(cm, d[15]));
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) ubmax = fmaxf(ubmax, ub[r]);
                    ubmax = warp_max_nonneg(ubmax);
                    // the decision threshold of any point of the tile is at most ubmax + tol(ubmax)
                    // (filter_tol grows with the distance and with the coordinate magnitudes)
                    const float pmax = fmaxf(fmaxf(fabsf(lx), fabsf(hx)), fmaxf(fabsf(ly), fabsf(hy)));
                    ubmax += filter_tol(ubmax, pmax, pmax, qmax);
                    // every target within sqrt(ubmax) of some point of the tile lies within `reach`
                    // of the tile centre; e covers the fp32 rounding of the centres and differences
                    const float e = 4.0f * 1.1920929e-7f * (pmax + qmax);
                    reach = (sqrt_fast(rho2) + sqrt_fast(ubmax)) * 1.0001f + e;
                }

                float m1[R], m2[R];
                int c1[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { m1[r] = kInf; m2[r] = kInf; c1[r] = 0; }

                const float4 *qx4 = reinterpret_cast<const float4 *>(tqx);
                const float4 *qy4 = reinterpret_cast<const float4 *>(tqy);
                for (int cbase = 0; cbase < nchunks; cbase += 32) {
                    unsigned need;
                    {
                        const int c = cbase + lane;
                        bool nd = c < nchunks;
                        if (PRUNE && nd) {
                            const float4 b = cb[c];
                            const float lim = (reach + b.z) * 1.00001f;
                            nd = dist32(tcx, tcy, b.x, b.y) <= lim * lim;
                        }
                        need = __ballot_sync(0xffffffffu, nd);
                    }
                    executed += (unsigned)__popc(need);
#pragma unroll 1
                    while (need) {
                        const int c = cbase + __ffs(need) - 1;
                        need &= need - 1;
                        float4 X[4], Y[4];
#pragma unroll
                        for (int v = 0; v < 4; ++v) { X[v] = qx4[4 * c + v]; Y[v] = qy4[4 * c + v]; }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            float d[16];
                            const u64 PX = pack2(px[r], px[r]), PY = pack2(py[r], py[r]);
#pragma unroll
                            for (int v = 0; v < 4; ++v) dist32x4(PX, PY, X[v], Y[v], d + 4 * v);
                            float cm = min3f(d[0], d[1], d[2]);
                            cm = min3f(cm, d[3], d[4]);
                            cm = min3f(cm, d[5], d[6]);
                            cm = min3f(cm, d[7], d[8]);
                            cm = min3f(cm, d[9], d[10]);
                            cm = min3f(cm, d[11], d[12]);
                            cm = min3f(cm, d[13], d[14]);
                            cm = fminf(cm, d[15]);
                            const bool better = cm < m1[r];
                            m2[r] = fminf(m2[r], better ? m1[r] : cm);
                            m1[r] = fminf(m1[r], cm);
                            c1[r] = better ? c : c1[r];
                        }
                    }
                }
                // ---- exact decision among the filter's candidates, then the fit sums ----
                {
                    double T[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) T[k] = Tmine[k];
                    double cx, cy;                               // shift = transformed first source point
                    apply_T(T, s0.x, s0.y, cx, cy);
#pragma unroll
                    for (int sub = 0; sub < SUBT; ++sub) {
                    double sum[kNumSums];
#pragma unroll
                    for (int k = 0; k < kNumSums; ++k) sum[k] = 0.0;
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = sub * 2 + rr;
                        const int i = pidx(r);
                        if (i < n1) {
                            const double2 s = src[i];
                            double Px, Py;
                            apply_T(T, s.x, s.y, Px, Py);
                            const float thr = m1[r] + filter_tol(m1[r], px[r], py[r], qmax);
                            const int j0 = c1[r] * kChunk;
                            // candidates inside the best chunk: how many, and where
                            float d[16];
                            {
                                const float4 *bx = reinterpret_cast<const float4 *>(tqx + j0);
                                const float4 *by = reinterpret_cast<const float4 *>(tqy + j0);
                                const u64 PX = pack2(px[r], px[r]), PY = pack2(py[r], py[r]);
#pragma unroll
                                for (int v = 0; v < 4; ++v) dist32x4(PX, PY, bx[v], by[v], d + 4 * v);
                            }
                            unsigned cand = 0;                   // bit k: target j0 + k is a candidate
                            // one FSETP + one predicated LOP per target
#define ICPB_CAND(K) asm("{ .reg .pred q; setp.le.f32 q, %1, %2; @q or.b32 %0, %0, %3; }" \
                         : "+r"(cand) : "f"(d[K]), "f"(thr), "n"(1 << K))
                            ICPB_CAND(0);  ICPB_CAND(1);  ICPB_CAND(2);  ICPB_CAND(3);
                            ICPB_CAND(4);  ICPB_CAND(5);  ICPB_CAND(6);  ICPB_CAND(7);
                            ICPB_CAND(8);  ICPB_CAND(9);  ICPB_CAND(10); ICPB_CAND(11);
                            ICPB_CAND(12); ICPB_CAND(13); ICPB_CAND(14); ICPB_CAND(15);
#undef ICPB_CAND
                            static_assert(kChunk == 16, "candidate count is written out for 16 targets");
                            int idx = j0 + __ffs(cand) - 1;      // unique candidate: no fp64 needed
                            if (m2[r] <= thr)                    // another chunk is within the bound
                                idx = exact_decide_all(nchunks, n2, j0, thr, px[r], py[r], Px, Py, tqx, tqy, cb, dst);
                            else if (cand & (cand - 1))          // more than one candidate in the chunk
                                idx = exact_decide_chunk(j0, cand, Px, Py, dst);
                            else if (cand == 0)                  // (non-finite input: keep a valid index)
                                idx = j0;
                            corr_s[i] = idx;
                            const double2 q = dst[idx];
                            const double ax = Px - cx, ay = Py - cy, bx = q.x - g.x, by = q.y - g.y;
                            sum[0] += ax; sum[1] += ay; sum[2] += bx; sum[3] += by;
                            sum[4] = fma(ax, bx, sum[4]); sum[5] = fma(ax, by, sum[5]);
                            sum[6] = fma(ay, bx, sum[6]); sum[7] = fma(ay, by, sum[7]);
                            const double ex = Px - q.x, ey = Py - q.y;
                            sum[8] += __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                        }
                    }
                    const int rt = tile * SUBT + sub;                           // reduction tile
                    if (rt < ntiles) {                                       // warp-uniform
                        const double (&s8)[8] = reinterpret_cast<const double (&)[8]>(sum);
                        const double tot = warp_sum8(s8, lane);              // sums 0..7, see warp_sum8
                        const double e8 = warp_sum(sum[8]);
                        if ((lane & 3) == 0)
                            redp[rt * kNumSums + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = tot;
                        if (lane == 1) redp[rt * kNumSums + 8] = e8;
                    }
                    }
                }
            }

            // =========== fit (src/icp.py:22-52): deterministic reduction, one barrier per pass ===========
            sync_all();
            if (tid == 0) s_tile_ctr[passes & 1] = 0;          // next used two passes from now
            // every warp folds the tile partials in the same fixed order and updates its own copy of T:
            // lane k + 9*part (part 0..2) adds column k over tiles = part (mod 3), in tile order
            double S[kNumSums];
            {
                double col = 0.0;
                if (lane < 3 * kNumSums) {
                    const int k = lane % kNumSums;
                    if (!CLUSTER) {
                        for (int t = lane / kNumSums; t < ntiles; t += 3) col += redp[t * kNumSums + k];
                    } else {                                     // tile t lives in CTA t mod csize
                        cg::cluster_group cluster = cg::this_cluster();
                        for (int t = lane / kNumSums; t < ntiles; t += 3)
                            col += cluster.map_shared_rank(redp, (t / SUBT) % csize)[t * kNumSums + k];
                    }
                }
#pragma unroll
                for (int k = 0; k < kNumSums; ++k)
                    S[k] = (__shfl_sync(0xffffffffu, col, k) + __shfl_sync(0xffffffffu, col, k + kNumSums))
                           + __shfl_sync(0xffffffffu, col, k + 2 * kNumSums);
            }
            {
                double T[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) T[k] = Tmine[k];
                double cx, cy;
                apply_T(T, s0.x, s0.y, cx, cy);
                const double ma_x = S[0] * inv_n, ma_y = S[1] * inv_n, mb_x = S[2] * inv_n, mb_y = S[3] * inv_n;
                // centred cross-covariance S = X Y^T (src/icp.py:29-32)
                const double s00 = S[4] - S[0] * mb_x, s01 = S[5] - S[0] * mb_y;
                const double s10 = S[6] - S[1] * mb_x, s11 = S[7] - S[1] * mb_y;
                // rotation maximising tr(R S): closed form of the SVD + det fix (src/icp.py:33-38)
                const double A = s00 + s11, Bv = s01 - s10;
                const double h2 = A * A + Bv * Bv;
                double c = 1.0, s = 0.0;
                if (h2 > 0.0 && h2 < 1e300) {
                    const double rh = rsqrt(h2);
                    c = A * rh; s = Bv * rh;
                } else if (h2 > 0.0) {                                   // huge coordinates: scale first
                    const double m = fmax(fabs(A), fabs(Bv));
                    const double an = A / m, bn = Bv / m;
                    const double rh = rsqrt(an * an + bn * bn);
                    c = an * rh; s = bn * rh;
                }
                const double xbar = cx + ma_x, ybar = cy + ma_y;       // mean of moved source
                const double qbx = g.x + mb_x, qby = g.y + mb_y;       // mean of matched target
                double tx = qbx - (c * xbar - s * ybar);               // src/icp.py:39
                double ty = qby - (s * xbar + c * ybar);
                if (a.p.rotation_only) { tx = 0.0; ty = 0.0; }         // src/icp.py:65-66
                double N[6];                                           // inc @ T (src/icp.py:67)
                N[0] = c * T[0] - s * T[3];
                N[1] = c * T[1] - s * T[4];
                N[2] = c * T[2] - s * T[5] + tx;
                N[3] = s * T[0] + c * T[3];
                N[4] = s * T[1] + c * T[4];
                N[5] = s * T[2] + c * T[5] + ty;
                err = S[8];
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) Tmine[k] = N[k];
                    if (warp == 0 && crank == 0 && a.hist && passes < a.p.hist_cap) {
                        double *hrow = a.hist + ((size_t)pid * a.p.hist_cap + passes) * 6;
#pragma unroll
                        for (int k = 0; k < 6; ++k) hrow[k] = N[k];
                    }
                }
                __syncwarp();
            }
            ++passes;
            // stop rules, in the reference's order (src/icp.py:86-95)
            bool done = err < a.p.epsilon;
            if (!done) done = iteration > a.p.max_iters;
            if (!done && have_last) done = fabs(last_err - err) < a.p.stopping_thresh;
            if (done) break;
            last_err = err; have_last = true;
            ++iteration;
        }
        if (tid == 0 && crank == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a.T_out[6 * pid + k] = Tmine[k];
            a.err_out[pid] = err;
            a.passes_out[pid] = passes;
        }
        if (a.n_peers > 0 && crank == 0 && tid < 8 * a.n_peers) {
            const int r = tid >> 3, k = tid & 7;                // 8 lanes per peer: one 64-byte record each
            // every warp reads its OWN copy of T (identical bits, ordered by its own __syncwarp):
            // with more than four peers the lanes of warp 1 store too, and warp 0's copy is not
            // ordered against them
            const double v = k < 6 ? Tmine[k] : (k == 6 ? err : (double)passes);
            a.peers[r][(a.rec_row0 + pid) * 8 + k] = v;
        }
        if (a.corr) {
            int32_t *crow = a.corr + (size_t)pid * a.p.corr_stride;
            const int lim = min(n1, a.p.corr_stride);
            for (int i = tid; i < lim; i += NT)
                if (!CLUSTER || (i / (32 * R)) % csize == crank) crow[i] = corr_s[i];   // own work items only
        }
        sync_all();             // smem is reused by the next problem (remote reads of red included)
    }
    if (a.executed) {
        // chunks processed by this warp x 16 targets x 32*R source-point slots
        const unsigned long long ex = (unsigned long long)executed * (unsigned long long)(kChunk * 32 * R);
        if (lane == 0 && ex) atomicAdd(a.executed, ex);
    }
}


// Rigid fit of n matched point pairs a[i] -> b[i] and their SSE: the reference's get_transform
// (src/icp.py:22-46) and get_error (src/icp.py:49-52) as stand-alone operations.  One CTA.
__global__ void __launch_bounds__(256)
fit_pairs_kernel(const double2 *a, const double2 *b, int n, double *T_out, double *err_out)
{
    __shared__ double red[8][kNumSums];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 a0 = a[0], b0 = b[0];
    double sum[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) sum[k] = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double2 p = a[i], q = b[i];
        const double ax = p.x - a0.x, ay = p.y - a0.y, bx = q.x - b0.x, by = q.y - b0.y;
        sum[0] += ax; sum[1] += ay; sum[2] += bx; sum[3] += by;
        sum[4] = fma(ax, bx, sum[4]); sum[5] = fma(ax, by, sum[5]);
        sum[6] = fma(ay, bx, sum[6]); sum[7] = fma(ay, by, sum[7]);
        const double ex = p.x - q.x, ey = p.y - q.y;
        sum[8] += __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    }
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) sum[k] = warp_sum(sum[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kNumSums; ++k) red[warp][k] = sum[k];
    }
    __syncthreads();
    if (tid == 0) {
        double S[kNumSums];
        for (int k = 0; k < kNumSums; ++k) {
            S[k] = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) S[k] += red[w][k];
        }
        const double inv_n = 1.0 / (double)n;
        const double ma_x = S[0] * inv_n, ma_y = S[1] * inv_n, mb_x = S[2] * inv_n, mb_y = S[3] * inv_n;
        const double s00 = S[4] - S[0] * mb_x, s01 = S[5] - S[0] * mb_y;
        const double s10 = S[6] - S[1] * mb_x, s11 = S[7] - S[1] * mb_y;
        const double A = s00 + s11, Bv = s01 - s10;
        const double h = hypot(A, Bv);
        double c = 1.0, s = 0.0;
        if (h > 0.0) { c = A / h; s = Bv / h; }
        const double xbar = a0.x + ma_x, ybar = a0.y + ma_y, qbx = b0.x + mb_x, qby = b0.y + mb_y;
        T_out[0] = c; T_out[1] = -s; T_out[2] = qbx - (c * xbar - s * ybar);
        T_out[3] = s; T_out[4] = c;  T_out[5] = qby - (s * xbar + c * ybar);
        *err_out = S[8];
    }
}

}  // namespace icpb
