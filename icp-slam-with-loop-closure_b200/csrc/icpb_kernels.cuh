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
//   * Exact pruning: 16-target chunks carry a bounding circle; a group of lanes (16 source points)
//     skips a chunk only when the triangle inequality proves that every target in it is farther
//     from every one of the group's points than that point's current upper bound (its filter
//     distance to the previous pass's match) plus the rounding bound.  Skipped targets can
//     therefore never be a candidate of the exact decision, so the result is that of the
//     exhaustive search.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "icpb.h"

namespace icpb {

constexpr int   kChunk    = 16;        // targets per bookkeeping chunk
constexpr float kPadCoord = 1.0e15f;   // coordinates of padding targets (distance ~2e30, never a candidate)
constexpr int   kMaxWarps = 32;
constexpr int   kNumSums  = 8;        // 7 fp64 sums per reduction tile (+1 pad): see kS* below
// columns of a tile's partial sums (a = moved source point - c, b = matched target - g):
enum { kSbx = 0, kSby = 1, kSaxbx = 2, kSaxby = 3, kSaybx = 4, kSayby = 5, kSerr = 6 };
constexpr int   kTw       = 8;        // doubles per warp-private transform record: T[6], sigma, -
// Pruning granularity: a warp's 64 source points are split into kGroups groups of consecutive lanes;
// every group has its own bounding circle, upper bound and list of target chunks, and in a sweep step
// each group reads its own next chunk (kGroups distinct shared-memory addresses per load).  Measured
// chunk steps per 64-point tile on the 1,024-beam chain: 6.3 with one group, 4.3 with two, 3.3 with
// four, 2.9 with eight (the step count is the maximum over the groups).
#ifndef ICPB_G
#define ICPB_G 4
#endif
#ifndef ICPB_EXH_UNROLL
#define ICPB_EXH_UNROLL 1
#endif
constexpr int   kExhUnroll = ICPB_EXH_UNROLL;     // chunk-loop unrolling of the exhaustive variant
constexpr int   kGroups   = ICPB_G;
constexpr int   kLpg      = 32 / kGroups;   // lanes per group
static_assert(kGroups == 1 || kGroups == 2 || kGroups == 4 || kGroups == 8, "kGroups must be 1, 2, 4 or 8");

// Dynamic shared memory of icp_align_kernel, in this order:
//   tqx, tqy   target SoA as fp32, n2pad floats each (whole chunks + one all-padding chunk)
//   cb         bounding circle per 16-target chunk, count rounded up to 32 (padding entries never pass)
//   tc         bounding circle per pruning group of every source tile (untransformed source)
//   mm         [2][tiles] (min, max) matched target index per tile
//   corr       current correspondences
//   red        [2][tiles][8] per-tile partial sums, double buffered by pass parity
//   tw         [warps][8] warp-private copy of T and its stretch
//   s0         sum of (source point - first source point)
//   scr        [warps][7 * 33] transpose scratch of the per-tile reduction
struct SmemLayout {
    int32_t n2pad, nchunk, ntile, n1c;
    int32_t o_tqy, o_cb, o_tc, o_mm, o_corr, o_red, o_tw, o_s0, o_scr, bytes;
};
constexpr int kScrStride = 33;                       // doubles per column of the transpose scratch
constexpr int kScrWarp   = 7 * kScrStride + 1;       // doubles per warp (kept even)
__host__ __device__ inline SmemLayout smem_layout(int64_t longest, int warps)
{
    SmemLayout L;
    const int64_t n2pad = (longest + 15) / 16 * 16 + 16;
    const int64_t nchunk = (n2pad / 16 + 31) / 32 * 32;
    const int64_t ntile = (longest + 63) / 64;
    const int64_t n1c = (longest + 3) & ~int64_t(3);
    int64_t o = 0;
    o += 4 * n2pad;                 L.o_tqy = (int32_t)o;
    o += 4 * n2pad;                 L.o_cb = (int32_t)o;
    o += 16 * nchunk;               L.o_tc = (int32_t)o;
    o += 16 * 8 * ntile;            L.o_mm = (int32_t)o;          // room for up to 8 groups per tile
    o += 8 * 2 * ntile;             L.o_corr = (int32_t)o;
    o += 4 * n1c;                   L.o_red = (int32_t)o;
    o += 8 * 2 * ntile * 8;         L.o_tw = (int32_t)o;
    o += 8 * 32 * 8;                L.o_s0 = (int32_t)o;
    o += 16;                        L.o_scr = (int32_t)o;
    o += 8 * (int64_t)warps * kScrWarp;
    L.n2pad = (int32_t)n2pad; L.nchunk = (int32_t)nchunk; L.ntile = (int32_t)ntile; L.n1c = (int32_t)n1c;
    L.bytes = o > 0x7fffffff ? 0x7fffffff : (int32_t)o;
    return L;
}

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
    // byte offsets of the arrays in dynamic shared memory (SmemLayout, computed once by the host):
    // a pointer is then one add away from the shared window, cheap to rematerialise
    int32_t        o_tqy, o_cb, o_tc, o_mm, o_corr, o_red, o_tw, o_s0, o_scr;
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
    // row = rec_row0 + (pid / rec_block) * rec_stride + pid % rec_block, so a rank that owns
    // interleaved blocks of the problem index space (dist.shard_indices) writes global rows
    double *const *peers;
    int32_t n_peers;
    int64_t rec_row0, rec_block, rec_stride;
    // acceptance epilogue (reference src/loop_closure_detection.py:35-39,155-159):
    // pairs with error < accept_thresh append their record [T(6), error, (row << 16) | passes] to
    // accept_rec (this rank's region of every peer's buffer when accept_peers is set) in completion
    // order; the last CTA to leave publishes the count
    double   accept_thresh;
    double  *accept_rec;              // local buffer, accept_cap rows of 8 doubles (or nullptr)
    double *const *accept_peers;      // n_peers buffers (each world * accept_cap rows) or nullptr
    unsigned long long *accept_ctr;   // [0] rows appended, [1] CTAs that have left
    long long *accept_count_out;      // local: final count; with peers: slot accept_rank of every peer's counts
    long long *const *accept_count_peers;
    int64_t  accept_cap;
    int32_t  accept_rank;
    // Per-scan staging images (the device form of a resident scan table, built once by a launch of this
    // same kernel with prep_out set): record s holds exactly what the staging phase below leaves in shared
    // memory for scan s -- target SoA, chunk circles, group circles -- plus S0 and the largest target
    // coordinate.  With prep_in a pair's staging is a copy: target arrays from the target scan's record,
    // group circles and S0 from the source scan's.
    unsigned char       *prep_out;   // != nullptr: problem b is scan b (source = target = b); build its record, no ICP
    const unsigned char *prep_in;
    int64_t              prep_stride;
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

// Packed or scalar?  Measured on B200 (profiles/r02_micro_pipes.log): scalar FADD/FMUL/FFMA issue at
// 3.9 warp-instructions per clock per SM, packed f32x2 at 1.98 (the same 126 lanes per clock), and a
// packed instruction does NOT overlap with an ALU-pipe instruction (FMNMX3, SEL, SHF, LOP3: 1.98 per
// clock on their own) -- FFMA2 and FMNMX3 alternating run at 0.99 + 0.99 -- while a scalar FFMA does
// (1.78 + 1.78).  So the all-packed sweep pays for its min tree in full: 4 x 2 + 1 x 2 = 10 cycles per
// 64 distances, an attainable 25.6 distances per clock per SM against the 32 of the FP32 pipe alone;
// the isolated loop reaches 98% of that (profiles/r02_micro_sweep.log).  Computing ICPB_MIX of every
// four targets with scalar instructions (same IEEE operation per element, bit-identical distances)
// should let the ALU work issue underneath them; in the kernel it did not pay (chain 1.51 ms either
// way, exhaustive variant 69% -> 66% with 2 of 4 scalar, 60% all scalar: more issue slots, and the
// kernel runs at 16-24 warps per SM where the mix is latency-bound), so the default stays packed.
#ifndef ICPB_MIX
#define ICPB_MIX 0
#endif
constexpr int kMixScalar = ICPB_MIX;      // 0, 2 or 4 of every 4 targets use scalar FADD/FMUL/FFMA

// filter distances of one source point to four targets (x0..x3, y0..y3)
__device__ __forceinline__ void dist32x4(float px, float py, const float4 &X, const float4 &Y, float *d)
{
    const u64 PX = pack2(px, px), PY = pack2(py, py);
    if (kMixScalar < 4) {
        const u64 dxa = sub2(pack2(X.x, X.y), PX), dya = sub2(pack2(Y.x, Y.y), PY);
        unpack2(fma2(dya, dya, mul2(dxa, dxa)), d[0], d[1]);
    } else {
        d[0] = dist32(px, py, X.x, Y.x); d[1] = dist32(px, py, X.y, Y.y);
    }
    if (kMixScalar < 2) {
        const u64 dxb = sub2(pack2(X.z, X.w), PX), dyb = sub2(pack2(Y.z, Y.w), PY);
        unpack2(fma2(dyb, dyb, mul2(dxb, dxb)), d[2], d[3]);
    } else {
        d[2] = dist32(px, py, X.z, Y.z); d[3] = dist32(px, py, X.w, Y.w);
    }
}

__device__ __forceinline__ float min3f(float a, float b, float c)
{
    return fminf(fminf(a, b), c);
}

// MUFU square root (1-2 ulp, denormal inputs flushed to zero).  Every use below carries a much larger
// safety factor; a flushed argument (< 1.2e-38) costs at most 8.7e-19 e in decision_thr, which its
// 16 e^2 + 1e-30 terms cover for every e, and the other uses take decision_thr's result (>= 1e-30).
__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Bound on |fp32 filter distance - exact distance|, doubled, as a function of the best filter
// distance m1: 8 sqrt(m1) e + 16 e^2 + 16 u m1, where e bounds the error of one coordinate difference
// (both inputs were rounded to fp32, relative 2^-24 each, and the subtraction rounds once more).
// Decision threshold of a point whose best filter distance is m1: m1 plus the doubled rounding bound,
// times (1 + 8u) for the sign-bit form of the candidate test (dist32x4_minus), plus an absolute 1e-30
// so that the threshold is positive (a zero filter distance is a candidate) and fp32 underflow of
// the squares (coordinates below ~1e-15) only ever adds candidates.  e is the bound on one
// coordinate difference (2u (|p|max + |q|max), u = 2^-24), taken once per tile from the tile's
// largest coordinate, so it is >= the bound of every point in it.
__device__ __forceinline__ float decision_thr(float m1, float e)
{
    return fmaf(m1, 1.000002f, fmaf(8.0f * sqrt_fast(m1), e, fmaf(16.0f * e, e, 1e-30f)));
}

// Rare path of the decision step: all targets j in [lo, hi) whose filter distance is <= thr are
// evaluated exactly; keeps the lexicographic (distance, index) minimum (j ascends, so strict <
// keeps the first index = np.argmin's rule).
extern __shared__ __align__(16) unsigned char smem_raw[];

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
// (the shared arrays are addressed by their byte offsets: a generic pointer argument would make the
// caller form -- and keep rematerialising -- 64-bit generic addresses of shared memory)
__device__ __noinline__ int exact_decide_all(int nchunks, int n2, int fallback, float thr, float px, float py,
                                             double Px, double Py, int o_tqy, int o_cb, const double2 *dst)
{
    const float *tqx = reinterpret_cast<const float *>(smem_raw);
    const float *tqy = reinterpret_cast<const float *>(smem_raw + o_tqy);
    const float4 *cb = reinterpret_cast<const float4 *>(smem_raw + o_cb);
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

// max / min over the lanes of one pruning group (kLpg consecutive lanes); every lane gets the result
__device__ __forceinline__ float group_max(float v)
{
#pragma unroll
    for (int o = kLpg / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group_min(float v)
{
#pragma unroll
    for (int o = kLpg / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// filter distance minus a threshold, for the decision step's candidate test: the sign bit of
// e = fl(dy*dy + fl(dx*dx - thr)) is set whenever the sweep's filter distance fl(dy*dy + fl(dx*dx))
// is <= thr / (1 + 8u) (u = 2^-24): with s = dx*dx + dy*dy <= thr (1 + 3u) / (1 + 8u) the exact value
// s - thr is below -4u thr, and the one rounding before the sign is taken moves it by at most
// u max(dx*dx, thr).  Same four packed instructions as a distance.
__device__ __forceinline__ float dist32_minus(float px, float py, float nthr, float qx, float qy)
{
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py);
    return __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, nthr));
}
__device__ __forceinline__ void dist32x4_minus(float px, float py, float nthr, const float4 &X, const float4 &Y, float *e)
{
    const u64 PX = pack2(px, px), PY = pack2(py, py), NTHR = pack2(nthr, nthr);
    if (kMixScalar < 4) {
        const u64 dxa = sub2(pack2(X.x, X.y), PX), dya = sub2(pack2(Y.x, Y.y), PY);
        unpack2(fma2(dya, dya, fma2(dxa, dxa, NTHR)), e[0], e[1]);
    } else {
        e[0] = dist32_minus(px, py, nthr, X.x, Y.x); e[1] = dist32_minus(px, py, nthr, X.y, Y.y);
    }
    if (kMixScalar < 2) {
        const u64 dxb = sub2(pack2(X.z, X.w), PX), dyb = sub2(pack2(Y.z, Y.w), PY);
        unpack2(fma2(dyb, dyb, fma2(dxb, dxb, NTHR)), e[2], e[3]);
    } else {
        e[2] = dist32_minus(px, py, nthr, X.z, Y.z); e[3] = dist32_minus(px, py, nthr, X.w, Y.w);
    }
}

// Largest singular value of the linear part of T (a bound on how much T stretches a distance), as
// a float with a safety factor.  1 for the rigid transforms ICP composes; the caller's initial
// guess may be any affine map with bottom row [0 0 1].
__device__ __forceinline__ double stretch_of(const double *T)
{
    const float a = (float)T[0], b = (float)T[1], c = (float)T[3], d = (float)T[4];
    const float f2 = a * a + b * b + c * c + d * d;
    const float det = a * d - b * c;
    const float disc = fmaxf(f2 * f2 - 4.0f * det * det, 0.0f);
    return (double)(sqrtf(0.5f * (f2 + sqrtf(disc))) * 1.00002f + 1e-30f);
}

// PRUNE = false: exhaustive sweep over every chunk (the reference's brute force; used for the
// FP32-pipe roofline characterisation).  PRUNE = true: exact chunk pruning (the product default).
//
// Shared memory: target SoA | chunk circles | tile circles | mm[2][tiles] | correspondences |
//                red[2][tiles][8] | Tw[warps][8] | S0[2]
// A tile = 32*R consecutive source points = one warp's register tile.  Warps pull tiles from a
// shared counter (tiles differ in how many chunks survive pruning); partial sums are stored per
// 64-point reduction tile and folded in tile order, so the result does not depend on which warp ran
// which tile.
//
// Per pair, once: the bounding circle of every tile's UNTRANSFORMED source points and
// S0 = sum(p_i - p_0).  A pass then gets the tile's circle as (T centre, stretch(T) radius) and the
// sum of the moved, shifted source points as A S0 (A = linear part of T) instead of reducing them
// again: only the largest upper bound of the tile and the 7 sums that depend on the matches are
// reduced per pass.
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
#ifndef ICPB_EXH_CTAS
#define ICPB_EXH_CTAS 2
#endif
__global__ void __launch_bounds__(256, (CLUSTER ? 1 : (R >= 4 ? ICPB_EXH_CTAS : ICPB_MIN_CTAS)))
icp_align_kernel(const KernelArgs a)
{
    float  *tqx    = reinterpret_cast<float *>(smem_raw);
    float  *tqy    = reinterpret_cast<float *>(smem_raw + a.o_tqy);
    float4 *cb     = reinterpret_cast<float4 *>(smem_raw + a.o_cb);          // chunk circle (cx, cy, r', -)
    float4 *tc     = reinterpret_cast<float4 *>(smem_raw + a.o_tc);          // group circles, untransformed source
    int2   *mm     = reinterpret_cast<int2 *>(smem_raw + a.o_mm);            // [2][ntile_cap] (min, max) matched index per tile
    int    *corr_s = reinterpret_cast<int *>(smem_raw + a.o_corr);
    double *red    = reinterpret_cast<double *>(smem_raw + a.o_red);         // [2][ntile_cap][8]
    double *Tw     = reinterpret_cast<double *>(smem_raw + a.o_tw);          // [kMaxWarps][8] per-warp copy of T
    double *S0     = reinterpret_cast<double *>(smem_raw + a.o_s0);          // sum of (source - first source point)
    __shared__ long long s_pid;
    __shared__ unsigned int s_qmax_bits;
    __shared__ int s_tile_ctr[2];

    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const float kInf = __int_as_float(0x7f800000);
    double *Tmine = Tw + warp * kTw;
    double *scr = reinterpret_cast<double *>(smem_raw + a.o_scr) + warp * kScrWarp;   // reduction scratch of this warp
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
        if (a.prep_out) {
            sid = did = (int32_t)pid;
        } else if (a.p.pair_mode == 1) {
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
        const double2 g = dst[0];                 // shifts for the one-pass covariance sums
        const double2 s0 = src[0];

        if (a.prep_in) {
            // ---------------- staging by copy from the per-scan images ----------------
            const unsigned char *imgT = a.prep_in + (size_t)did * a.prep_stride;
            const unsigned char *imgS = a.prep_in + (size_t)sid * a.prep_stride;
            {
                const float4 *gx = reinterpret_cast<const float4 *>(imgT), *gy = reinterpret_cast<const float4 *>(imgT + a.o_tqy);
                float4 *sx4 = reinterpret_cast<float4 *>(tqx), *sy4 = reinterpret_cast<float4 *>(tqy);
                for (int j = tid; j < (n2pad + kChunk) / 4; j += NT) { sx4[j] = gx[j]; sy4[j] = gy[j]; }
                const float4 *gc = reinterpret_cast<const float4 *>(imgT + a.o_cb);
                for (int c = tid; c < ((nchunks + 31) & ~31); c += NT) cb[c] = gc[c];
                if (PRUNE) {
                    const float4 *gt = reinterpret_cast<const float4 *>(imgS + a.o_tc);
                    for (int w = tid; w < nwork * kGroups; w += NT) tc[w] = gt[w];
                }
                if (tid == 0) {
                    const double *g0 = reinterpret_cast<const double *>(imgS + a.o_mm);
                    S0[0] = g0[0]; S0[1] = g0[1];
                    s_qmax_bits = *reinterpret_cast<const unsigned int *>(imgT + a.o_mm + 16);
                }
            }
            if (lane < 6) {
                double v = a.init ? a.init[6 * pid + lane] : ((lane == 0 || lane == 4) ? 1.0 : 0.0);
                if (a.p.rotation_only && (lane == 2 || lane == 5)) v = 0.0;     // src/icp.py:60-61
                Tmine[lane] = v;
            }
            __syncwarp();
            if (lane == 0) Tmine[6] = stretch_of(Tmine);
            __syncthreads();
        } else {
        // ---------------- stage the target in shared memory as fp32 SoA ----------------
        {
            float qm = 0.0f;
            for (int j = tid; j < n2pad + kChunk; j += NT) {     // + one all-padding chunk (index nchunks)
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
        // ---------------- per work item: circle of the untransformed source points, S0 partials -------
        // (the partials go to the second half of `red`, which the passes use from pass 1 on)
        for (int w = warp; w < nwork; w += nwarps) {
            float lx = kInf, ly = kInf, hx = -kInf, hy = -kInf;
            float sx[R], sy[R];
#pragma unroll
            for (int sub = 0; sub < SUBT; ++sub) {
                double ax = 0.0, ay = 0.0;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = sub * 2 + rr;
                    const int i = ((w * SUBT + sub) * 32 + lane) * 2 + rr;
                    const double2 s = src[min(i, n1 - 1)];       // lanes past the end repeat the last point
                    sx[r] = (float)s.x; sy[r] = (float)s.y;
                    if (i < n1) { ax += s.x - s0.x; ay += s.y - s0.y; }
                }
                const int rt = w * SUBT + sub;
                if (rt < ntiles) {                               // warp-uniform
                    ax = warp_sum(ax); ay = warp_sum(ay);
                    if (lane == 0) {
                        red[(a.ntile_cap + rt) * kNumSums] = ax;
                        red[(a.ntile_cap + rt) * kNumSums + 1] = ay;
                    }
                }
            }
            if (PRUNE) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    lx = fminf(lx, sx[r]); hx = fmaxf(hx, sx[r]);
                    ly = fminf(ly, sy[r]); hy = fmaxf(hy, sy[r]);
                }
                lx = group_min(lx); ly = group_min(ly); hx = group_max(hx); hy = group_max(hy);
                const float ccx = 0.5f * (lx + hx), ccy = 0.5f * (ly + hy);
                float r2 = 0.0f;
#pragma unroll
                for (int r = 0; r < R; ++r) r2 = fmaxf(r2, dist32(ccx, ccy, sx[r], sy[r]));
                r2 = group_max(r2);
                // fp32 rounding of the points and of the centre: a few ulps of the largest coordinate
                const float smax = fmaxf(fmaxf(fabsf(lx), fabsf(hx)), fmaxf(fabsf(ly), fabsf(hy)));
                if ((lane & (kLpg - 1)) == 0)
                    tc[w * kGroups + lane / kLpg] = make_float4(ccx, ccy, sqrtf(r2) * 1.0001f + 4.8e-7f * smax, 0.0f);
            }
        }
        __syncthreads();
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
            // radius with the safety factors of the chunk test folded in: (reach + r) 1.00001 there
            cb[c] = make_float4(cx, cy, sqrtf(r2) * 1.00001f * 1.00001f, 0.0f);
        }
        // padding entries up to a multiple of 32 chunks: so far away that no test passes
        for (int c = nchunks + tid; c < ((nchunks + 31) & ~31); c += NT)
            cb[c] = make_float4(kPadCoord, kPadCoord, 0.0f, 0.0f);
        if (tid == NT - 1) {                                      // fixed order: tile 0, 1, 2, ...
            double ax = 0.0, ay = 0.0;
            for (int t = 0; t < ntiles; ++t) {
                ax += red[(a.ntile_cap + t) * kNumSums];
                ay += red[(a.ntile_cap + t) * kNumSums + 1];
            }
            S0[0] = ax; S0[1] = ay;
        }
        if (lane == 0) Tmine[6] = stretch_of(Tmine);
        __syncthreads();
        }
        const float qmax = __uint_as_float(s_qmax_bits);
        if (a.prep_out) {
            // ---------------- prep mode: this scan's staging image goes to HBM, no ICP ----------------
            unsigned char *img = a.prep_out + (size_t)pid * a.prep_stride;
            float4 *gx = reinterpret_cast<float4 *>(img), *gy = reinterpret_cast<float4 *>(img + a.o_tqy);
            const float4 *sx4 = reinterpret_cast<const float4 *>(tqx), *sy4 = reinterpret_cast<const float4 *>(tqy);
            for (int j = tid; j < (n2pad + kChunk) / 4; j += NT) { gx[j] = sx4[j]; gy[j] = sy4[j]; }
            float4 *gc = reinterpret_cast<float4 *>(img + a.o_cb);
            for (int c = tid; c < ((nchunks + 31) & ~31); c += NT) gc[c] = cb[c];
            float4 *gt = reinterpret_cast<float4 *>(img + a.o_tc);
            for (int w = tid; w < nwork * kGroups; w += NT) gt[w] = tc[w];
            if (tid == 0) {
                double *g0 = reinterpret_cast<double *>(img + a.o_mm);
                g0[0] = S0[0]; g0[1] = S0[1];
                *reinterpret_cast<unsigned int *>(img + a.o_mm + 16) = s_qmax_bits;
            }
            sync_all();
            continue;
        }

        int passes = 0, iteration = 0;
        bool have_last = false;
        double last_err = 0.0, err = 0.0;

        for (;;) {
            double *redp = red + (passes & 1) * a.ntile_cap * kNumSums;
            int2 *mmp = mm + (passes & 1) * a.ntile_cap;

            for (;;) {
                int tile = 0;
                if (lane == 0) tile = atomicAdd(&s_tile_ctr[passes & 1], 1);
                tile = __shfl_sync(0xffffffffu, tile, 0);
                if (CLUSTER) tile = crank + tile * csize;        // this CTA owns items crank, crank + csize, ...
                if (tile >= nwork) break;
                // point r of this lane: reduction tile tile*SUBT + r/2, lane's pair of consecutive points
                auto pidx = [&](int r) { return ((tile * SUBT + (r >> 1)) * 32 + lane) * 2 + (r & 1); };
                // ---- transform ----
                float px[R], py[R];
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
                // ---- group circle, upper bounds, reach (per pruning group of kLpg lanes) ----
                float tol_e = 0.f;                               // per-coordinate rounding bound of the group
                float m1[R], m2[R];
                int c1[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { m1[r] = kInf; m2[r] = kInf; c1[r] = 0; }
                const float4 *qx4 = reinterpret_cast<const float4 *>(tqx);
                const float4 *qy4 = reinterpret_cast<const float4 *>(tqy);
                // smallest filter distance of point r to the 16 targets of a staged chunk
                auto chunk_min = [&](const float4 (&X)[4], const float4 (&Y)[4], int r) -> float {
                    float d[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) dist32x4(px[r], py[r], X[v], Y[v], d + 4 * v);
                    float cm = min3f(d[0], d[1], d[2]);
                    cm = min3f(cm, d[3], d[4]);
                    cm = min3f(cm, d[5], d[6]);
                    cm = min3f(cm, d[7], d[8]);
                    cm = min3f(cm, d[9], d[10]);
                    cm = min3f(cm, d[11], d[12]);
                    cm = min3f(cm, d[13], d[14]);
                    return fminf(cm, d[15]);
                };
                // one sweep step: chunk c (per lane: the lanes of a group agree, groups differ)
                auto sweep = [&](int c) {
                    float4 X[4], Y[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) { X[v] = qx4[4 * c + v]; Y[v] = qy4[4 * c + v]; }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float cm = chunk_min(X, Y, r);
                        const bool better = cm < m1[r];
                        m2[r] = fminf(m2[r], better ? m1[r] : cm);
                        m1[r] = fminf(m1[r], cm);
                        c1[r] = better ? c : c1[r];
                    }
                };
                if (PRUNE) {
                    const int grp = lane / kLpg, li = lane & (kLpg - 1);
                    const float4 c0 = tc[tile * kGroups + grp];
                    double X, Y;
                    apply_T(T, (double)c0.x, (double)c0.y, X, Y);
                    const float tcx = (float)X, tcy = (float)Y;
                    const float rho = (float)Tmine[6] * c0.z;
                    // largest coordinate magnitude of the group's moved points
                    const float pmax = fmaxf(fabsf(tcx), fabsf(tcy)) * 1.000001f + rho;
                    tol_e = 2.0f * 5.9604645e-8f * (pmax + qmax) * 1.0001f;
                    // upper bound of every point's nearest-neighbour filter distance: its distance to
                    // a real target -- the previous pass's match, or on the first pass the best
                    // target in the chunks around the chunk centre nearest to the group centre
                    float ub[R];
                    if (passes > 0) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int j = corr_s[min(pidx(r), n1 - 1)];
                            ub[r] = dist32(px[r], py[r], tqx[j], tqy[j]);
                        }
                    } else {
                        unsigned key = 0xffffffffu;
                        for (int c = li; c < nchunks; c += kLpg) {
                            const float4 bc = cb[c];
                            const unsigned k = (__float_as_uint(dist32(tcx, tcy, bc.x, bc.y)) & 0xfffff000u) | (unsigned)min(c, 4095);
                            key = min(key, k);
                        }
#pragma unroll
                        for (int o = kLpg / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
                        const int cstar = (int)(key & 0xfffu);
#pragma unroll
                        for (int r = 0; r < R; ++r) ub[r] = kInf;
                        for (int dc = -1; dc <= 1; ++dc) {
                            int c = cstar + dc;
                            if (c < 0 || c >= nchunks) c = nchunks;      // the all-padding chunk
                            float4 X4[4], Y4[4];
#pragma unroll
                            for (int v = 0; v < 4; ++v) { X4[v] = qx4[4 * c + v]; Y4[v] = qy4[4 * c + v]; }
#pragma unroll
                            for (int r = 0; r < R; ++r) ub[r] = fminf(ub[r], chunk_min(X4, Y4, r));
                        }
                    }
                    float ubmax = 0.0f;
#pragma unroll
                    for (int r = 0; r < R; ++r) ubmax = fmaxf(ubmax, ub[r]);
                    ubmax = group_max(ubmax);
                    // the decision threshold of any point of the group is at most ubmax + tol(ubmax)
                    // (the tolerance grows with the distance and with the coordinate magnitudes)
                    ubmax = decision_thr(ubmax, tol_e);
                    // every target within sqrt(ubmax) of some point of the group lies within `reach`
                    // of the group centre; e covers the fp32 rounding of the centres and differences
                    const float e = 4.0f * 1.1920929e-7f * (pmax + qmax);
                    const float reach = ((rho + sqrt_fast(ubmax)) * 1.0001f + 2.0f * e) * 1.00001f;
                    const u64 TC = pack2(tcx, tcy);

                    for (int cbase = 0; cbase < nchunks; cbase += 32) {
                        // chunks of this block that the group has to sweep: lane li of the group tests
                        // chunks cbase + li, cbase + kLpg + li, ... against the group's circle
                        // (padding circles beyond nchunks never pass: no bounds check)
                        unsigned mask = 0, bal[kGroups];
#pragma unroll
                        for (int j = 0; j < kGroups; ++j) {
                            const float4 b = cb[cbase + j * kLpg + li];
                            float dx2, dy2;
                            const u64 dd = sub2(pack2(b.x, b.y), TC);
                            unpack2(mul2(dd, dd), dx2, dy2);
                            const float lim = reach + b.z;
                            bal[j] = __ballot_sync(0xffffffffu, dx2 + dy2 <= lim * lim);
                        }
                        if (kGroups == 1) {
                            mask = bal[0];
                        } else if (kGroups == 4) {              // byte grp of every ballot, by byte permutes
                            const unsigned sel = (unsigned)grp | ((unsigned)(4 + grp) << 4);
                            mask = __byte_perm(__byte_perm(bal[0], bal[1], sel), __byte_perm(bal[2], bal[3], sel), 0x5410);
                        } else if (kGroups == 2) {              // 16-bit half grp of both ballots
                            const unsigned h = 2u * (unsigned)grp;
                            mask = __byte_perm(bal[0], bal[1], h | ((h + 1) << 4) | ((h + 4) << 8) | ((h + 5) << 12));
                        } else {
#pragma unroll
                            for (int j = 0; j < kGroups; ++j)
                                mask |= ((bal[j] >> (grp * kLpg)) & ((1u << (kLpg & 31)) - 1u)) << ((j * kLpg) & 31);
                        }
                        // the warp takes as many steps as its busiest group; a group that has run out of
                        // chunks sweeps the all-padding chunk (index nchunks), which changes nothing
                        const int steps = __reduce_max_sync(0xffffffffu, __popc(mask));
                        executed += (unsigned)steps;
#pragma unroll 1
                        for (int st = 0; st < steps; ++st) {
                            const int c = mask ? cbase + __ffs(mask) - 1 : nchunks;
                            mask &= mask - 1;
                            sweep(c);
                        }
                    }
                } else {
                    // no bound: the rounding bound of each point from the warp's largest coordinate
                    float pm = 0.0f;
#pragma unroll
                    for (int r = 0; r < R; ++r) pm = fmaxf(pm, fmaxf(fabsf(px[r]), fabsf(py[r])));
                    pm = warp_max_nonneg(pm);
                    tol_e = 2.0f * 5.9604645e-8f * (pm + qmax) * 1.0001f;
                    executed += (unsigned)nchunks;
#pragma unroll kExhUnroll
                    for (int c = 0; c < nchunks; ++c) sweep(c);
                }
                // ---- exact decision among the filter's candidates, then the fit sums ----
                {
                    double cx, cy;                               // shift = transformed first source point
                    apply_T(T, s0.x, s0.y, cx, cy);
#pragma unroll
                    for (int sub = 0; sub < SUBT; ++sub) {
                    double sum[kNumSums];
#pragma unroll
                    for (int k = 0; k < kNumSums; ++k) sum[k] = 0.0;
                    int imin = 0x7fffffff, imax = -1;            // range of this lane's matched target indices
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = sub * 2 + rr;
                        const int i = pidx(r);
                        if (i < n1) {
                            const double2 s = src[i];
                            double Px, Py;
                            apply_T(T, s.x, s.y, Px, Py);
                            const float thr = decision_thr(m1[r], tol_e);
                            const int j0 = c1[r] * kChunk;
                            // candidates inside the best chunk: sign bit of (distance - thr), shifted
                            // into a mask one target at a time (bit 15 - k <-> target j0 + k)
                            unsigned acc0 = 0, acc1 = 0;
                            {
                                const float4 *bx = reinterpret_cast<const float4 *>(tqx + j0);
                                const float4 *by = reinterpret_cast<const float4 *>(tqy + j0);
                                float e[16];
#pragma unroll
                                for (int v = 0; v < 4; ++v) dist32x4_minus(px[r], py[r], -thr, bx[v], by[v], e + 4 * v);
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    acc0 = __funnelshift_l(__float_as_uint(e[k]), acc0, 1);
                                    acc1 = __funnelshift_l(__float_as_uint(e[k + 8]), acc1, 1);
                                }
                            }
                            static_assert(kChunk == 16, "the candidate mask is written out for 16 targets");
                            const unsigned acc = (acc0 << 8) | acc1;
                            int idx = j0 + __clz((int)acc) - 16;  // unique candidate: no fp64 needed
                            if (m2[r] <= thr)                    // another chunk is within the bound
                                idx = exact_decide_all(nchunks, n2, j0, thr, px[r], py[r], Px, Py, a.o_tqy, a.o_cb, dst);
                            else if (acc & (acc - 1))            // more than one candidate in the chunk
                                idx = exact_decide_chunk(j0, __brev(acc) >> 16, Px, Py, dst);
                            else if (acc == 0)                   // (non-finite input: keep a valid index)
                                idx = j0;
                            corr_s[i] = idx;
                            imin = min(imin, idx); imax = max(imax, idx);
                            const double2 q = dst[idx];
                            const double ax = Px - cx, ay = Py - cy, bx = q.x - g.x, by = q.y - g.y;
                            sum[kSbx] += bx; sum[kSby] += by;
                            sum[kSaxbx] = fma(ax, bx, sum[kSaxbx]); sum[kSaxby] = fma(ax, by, sum[kSaxby]);
                            sum[kSaybx] = fma(ay, bx, sum[kSaybx]); sum[kSayby] = fma(ay, by, sum[kSayby]);
                            const double ex = Px - q.x, ey = Py - q.y;
                            sum[kSerr] += __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                        }
                    }
                    const int rt = tile * SUBT + sub;                           // reduction tile
                    if (rt < ntiles) {                                       // warp-uniform
                        // Warp total of the 7 sums through the transpose scratch: every lane stores its
                        // values (column k at k*33 + lane: conflict free), lane k + 8*part adds column k
                        // over lanes 8*part .. 8*part + 7 in lane order, two butterfly levels add the four
                        // parts.  Fixed order -> deterministic; ~37 instructions against ~77 for a shuffle
                        // butterfly with its register selects.
#pragma unroll
                        for (int k = 0; k < 7; ++k) scr[k * kScrStride + lane] = sum[k];
                        __syncwarp();
                        double tot = 0.0;
                        {
                            const double *col = scr + (lane & 7) * kScrStride + (lane >> 3) * 8;
                            if ((lane & 7) < 7) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) tot += col[j];
                            }
                        }
                        tot += __shfl_xor_sync(0xffffffffu, tot, 8);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 16);
                        if (lane < 7) redp[rt * kNumSums + lane] = tot;
                        __syncwarp();
                        imin = __reduce_min_sync(0xffffffffu, imin); imax = __reduce_max_sync(0xffffffffu, imax);
                        if (lane == 0) mmp[rt] = make_int2(imin, imax);
                    }
                    }
                }
            }

            // =========== fit (src/icp.py:22-52): deterministic reduction, one barrier per pass ===========
            sync_all();
            if (tid == 0) s_tile_ctr[passes & 1] = 0;          // next used two passes from now
            // every warp folds the tile partials in the same fixed order and updates its own copy of T:
            // lane k + 8*part (part 0..3) adds column k over tiles = part (mod 4), in tile order; two
            // butterfly levels add the four parts, then column k is broadcast from lane k
            double S[kNumSums];
            bool single_target;                                // every source point matched the same target
            {
                double col = 0.0;
                const int k = lane & 7;
                int mn = 0x7fffffff, mx = -1;
                if (!CLUSTER) {
                    for (int t = lane >> 3; t < ntiles; t += 4) col += redp[t * kNumSums + k];
                    for (int t = lane; t < ntiles; t += 32) { const int2 v = mmp[t]; mn = min(mn, v.x); mx = max(mx, v.y); }
                } else {                                     // tile t lives in CTA (t / SUBT) mod csize
                    cg::cluster_group cluster = cg::this_cluster();
                    for (int t = lane >> 3; t < ntiles; t += 4)
                        col += cluster.map_shared_rank(redp, (t / SUBT) % csize)[t * kNumSums + k];
                    for (int t = lane; t < ntiles; t += 32) {
                        const int2 v = cluster.map_shared_rank(mmp, (t / SUBT) % csize)[t];
                        mn = min(mn, v.x); mx = max(mx, v.y);
                    }
                }
                single_target = __reduce_min_sync(0xffffffffu, mn) == __reduce_max_sync(0xffffffffu, mx);
                col += __shfl_xor_sync(0xffffffffu, col, 8);
                col += __shfl_xor_sync(0xffffffffu, col, 16);
#pragma unroll
                for (int kk = 0; kk < 7; ++kk) S[kk] = __shfl_sync(0xffffffffu, col, kk);
            }
            {
                double T[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) T[k] = Tmine[k];
                double cx, cy;
                apply_T(T, s0.x, s0.y, cx, cy);
                // sum of the moved source points minus c: the linear part of T applied to S0
                const double Sax = fma(T[1], S0[1], T[0] * S0[0]), Say = fma(T[4], S0[1], T[3] * S0[0]);
                const double ma_x = Sax * inv_n, ma_y = Say * inv_n, mb_x = S[kSbx] * inv_n, mb_y = S[kSby] * inv_n;
                // centred cross-covariance S = X Y^T (src/icp.py:29-32)
                const double s00 = S[kSaxbx] - Sax * mb_x, s01 = S[kSaxby] - Sax * mb_y;
                const double s10 = S[kSaybx] - Say * mb_x, s11 = S[kSayby] - Say * mb_y;
                // rotation maximising tr(R S): closed form of the SVD + det fix (src/icp.py:33-38)
                const double A = s00 + s11, Bv = s01 - s10;
                const double h2 = A * A + Bv * Bv;
                // One distinct matched target: the centred target cloud is zero, the covariance is zero and
                // its SVD gives the identity rotation (SURVEY probe B5).  The reference only gets there when
                // its mean of N copies happens to round to the value itself; otherwise its rotation is the
                // rounding noise of that mean.  Here the case is decided exactly: identity.
                double c = 1.0, s = 0.0;
                if (single_target) {
                } else if (h2 > 0.0 && h2 < 1e300) {
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
                err = S[kSerr];
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) Tmine[k] = N[k];
                    Tmine[6] = stretch_of(N);
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
        // ---- epilogues: constraint records straight into the consumers' buffers ----
        if ((a.n_peers > 0 || a.accept_rec || a.accept_peers) && crank == 0) {
            const int64_t row = a.rec_row0 + (pid / a.rec_block) * a.rec_stride + pid % a.rec_block;
            // every warp reads its OWN copy of T (identical bits, ordered by its own __syncwarp)
            auto field = [&](int k) -> double {
                return k < 6 ? Tmine[k] : (k == 6 ? err : (double)passes);
            };
            if (a.peers)                                         // fused all-gather: one 64-byte record per peer
                for (int t = tid; t < 8 * a.n_peers; t += NT)
                    a.peers[t >> 3][row * 8 + (t & 7)] = field(t & 7);
            if ((a.accept_rec || a.accept_peers) && err < a.accept_thresh && warp == 0) {
                // acceptance test + compaction (src/loop_closure_detection.py:35-39): append in
                // completion order; the row id travels in the record, so the consumer can restore order
                long long slot = 0;
                if (lane == 0) slot = (long long)atomicAdd(a.accept_ctr, 1ULL);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot < a.accept_cap) {
                    const long long tag = (long long)((row << 16) | (int64_t)min(passes, 0xffff));
                    const int npeer = a.accept_peers ? a.n_peers : 1;
                    for (int t = lane; t < 8 * npeer; t += 32) {
                        const int k = t & 7;
                        double *base = a.accept_peers ? a.accept_peers[t >> 3] + (size_t)a.accept_rank * a.accept_cap * 8
                                                      : a.accept_rec;
                        base[slot * 8 + k] = k == 7 ? __longlong_as_double(tag) : field(k);
                    }
                }
            }
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
    if ((a.accept_rec || a.accept_peers) && tid == 0 && crank == 0) {
        // the last CTA (cluster) to leave publishes how many records were appended
        __threadfence();
        const unsigned long long left = atomicAdd(a.accept_ctr + 1, 1ULL) + 1;
        const unsigned long long total_ctas = (unsigned long long)(gridDim.x / csize);
        if (left == total_ctas) {
            __threadfence();
            long long n = (long long)*(volatile unsigned long long *)a.accept_ctr;
            if (n > a.accept_cap) n = -n;                        // overflow: the caller sees a negative count
            if (a.accept_count_peers)
                for (int r = 0; r < a.n_peers; ++r) a.accept_count_peers[r][a.accept_rank] = n;
            else if (a.accept_count_out)
                *a.accept_count_out = n;
        }
    }
}


// Rigid fit of n matched point pairs a[i] -> b[i] and their SSE: the reference's get_transform
// (src/icp.py:22-46) and get_error (src/icp.py:49-52) as stand-alone operations.  One CTA.
constexpr int kFitSums = 9;
__global__ void __launch_bounds__(256)
fit_pairs_kernel(const double2 *a, const double2 *b, int n, double *T_out, double *err_out)
{
    __shared__ double red[8][kFitSums];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 a0 = a[0], b0 = b[0];
    double sum[kFitSums];
#pragma unroll
    for (int k = 0; k < kFitSums; ++k) sum[k] = 0.0;
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
    for (int k = 0; k < kFitSums; ++k) sum[k] = warp_sum(sum[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kFitSums; ++k) red[warp][k] = sum[k];
    }
    __syncthreads();
    if (tid == 0) {
        double S[kFitSums];
        for (int k = 0; k < kFitSums; ++k) {
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
