// Pose-graph relaxation on the GPU: one pass of the reference's "modified SGD"
// (reference src/pose_graph_optimization.py:7-49, pose_graph_optimization_step_sgd), the consumer
// of the ICP path's constraints.
//
// The reference walks every loop-closure edge (a, b, tf) twice in pure Python:
//   (1) weights  :13-24  M[i] += diag(inv(R(theta_a) sigma R^T)) for i in (a, b], gamma = the
//                        smallest-norm diagonal seen (first one on ties);
//   (2) updates  :27-48  residual r of the edge under the CURRENT poses, d = 2 inv(R^T sigma R) r,
//                        per dof j a step beta_j spread over the nodes (a, b] in proportion to
//                        1/M[i,j], accumulated along the chain and added to every later node.
// Pass (2) is order dependent -- an edge reads poses that earlier edges moved -- so it stays a
// sequential chain over the edges, but the O(N) inner loops per edge are data parallel:
//   * sgd_weights_kernel: one thread per edge computes its diagonal (E sin/cos pairs);
//   * sgd_accumulate_kernel: one thread per node adds the diagonals of the edges covering it, in
//     edge order (the reference's order of additions, so M has the reference's bits given the
//     same diagonals); the edge list streams through shared memory;
//   * sgd_chain_kernel: ONE CTA walks the edges.  The prologue finds gamma (first minimum) and the
//     prefix sums P_j[i] = sum_{k<=i} 1/M[k,j]; per edge every thread evaluates the residual and
//     the clipped step redundantly (no broadcast barrier), then node i > a receives
//     beta_j/total_j * (P_j[min(i,b)] - P_j[a]) -- the reference's running sum `dpose` in closed
//     form -- and one barrier orders the edge against the next.  Poses live in shared memory when
//     they fit (24 B per node), else in global memory.
// All arithmetic is fp64; the results agree with the reference to rounding (the 3x3 inverses are
// evaluated in closed form, the running sums as prefix differences); tests/test_gpu_sgd.py pins
// them to 1e-9 against goldens of the unmodified reference.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace icpb {

struct SgdArgs {
    double        *poses;     // n x 3 (x, y, theta), updated in place
    const int32_t *edges;     // E x 2 (a, b) in the graph's iteration order
    const double  *tf;        // E x 6: top two rows of the edge's 3x3 transform
    int32_t        n, E;
    double         learning_rate, lcu;
    double        *dW;        // E x 4 scratch: diag(W) and its squared norm (+inf: edge adds no weight)
    double        *M;         // n x 3 scratch: weights
    double        *P;         // n x 3 scratch: inclusive prefix sums of 1/M
    double        *PB;        // E x 6 scratch: per edge P_j[a] and P_j[b] - P_j[a]
    int32_t        poses_in_smem;
    int32_t        slice;     // cluster launch: nodes per CTA
};

// the optimiser ignores odometry edges (src/pose_graph_optimization.py:14-16, :28-30)
__device__ __forceinline__ bool sgd_skipped(int a, int b) { return a - b == 1 || b - a == 1; }

// diag(inv(R sigma R^T)), R = rot(theta), sigma = lcu * I (:17-19), products in matmul order
__device__ __forceinline__ void sgd_diag(double theta, double lcu, double *w)
{
    double s, c;
    sincos(theta, &s, &c);
    const double cl = c * lcu, sl = s * lcu, nsl = -s * lcu;
    const double b00 = cl * c + nsl * -s, b01 = cl * s + nsl * c;
    const double b10 = sl * c + cl * -s,  b11 = sl * s + cl * c;
    const double det = b00 * b11 - b01 * b10;
    w[0] = b11 / det; w[1] = b00 / det; w[2] = 1.0 / lcu;
}

__global__ void __launch_bounds__(256)
sgd_weights_kernel(const SgdArgs a)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    const int ea = a.edges[2 * e], eb = a.edges[2 * e + 1];
    double w[3] = {0.0, 0.0, 0.0};
    double nrm = __longlong_as_double(0x7ff0000000000000LL);
    if (!sgd_skipped(ea, eb) && eb > ea) {                     // (the host already drops the others)
        sgd_diag(a.poses[3 * ea + 2], a.lcu, w);
        nrm = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    }
    a.dW[4 * e] = w[0]; a.dW[4 * e + 1] = w[1]; a.dW[4 * e + 2] = w[2]; a.dW[4 * e + 3] = nrm;
}

__global__ void __launch_bounds__(256)
sgd_accumulate_kernel(const SgdArgs a)
{
    __shared__ int2 s_ab[256];
    __shared__ double s_w[256][3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    for (int e0 = 0; e0 < a.E; e0 += 256) {
        const int e = e0 + threadIdx.x;
        if (e < a.E) {
            int ea = a.edges[2 * e], eb = a.edges[2 * e + 1];
            if (sgd_skipped(ea, eb)) eb = ea;                  // empty range
            s_ab[threadIdx.x] = make_int2(ea, eb);
            s_w[threadIdx.x][0] = a.dW[4 * e]; s_w[threadIdx.x][1] = a.dW[4 * e + 1]; s_w[threadIdx.x][2] = a.dW[4 * e + 2];
        }
        __syncthreads();
        const int cnt = min(256, a.E - e0);
        for (int k = 0; k < cnt; ++k) {                        // edge order = the reference's order of additions
            const int2 ab = s_ab[k];
            if (ab.x < i && i <= ab.y) { m0 += s_w[k][0]; m1 += s_w[k][1]; m2 += s_w[k][2]; }
        }
        __syncthreads();
    }
    if (i < a.n) { a.M[3 * i] = m0; a.M[3 * i + 1] = m1; a.M[3 * i + 2] = m2; }
}

// x mod m for m > 0 with the sign of m (np.remainder, :35).  x - floor(x/m)*m in one FMA: for
// |x| < 2^20 m the exact difference is a multiple of ulp(m) below m, hence representable, so this
// equals the exact fmod-based result except when x/m rounds across an integer (fixed up below).
__device__ __forceinline__ double mod_pos(double x, double m)
{
    const double q = floor(x / m);
    double r = fma(-q, m, x);
    if (r < 0.0) r += m;
    if (r >= m) r -= m;
    return r;
}

// what one edge needs that does not depend on the moving poses
struct SgdEdge {
    int ea, eb;
    double t0, t2, t3, t5;        // tf[0,0], tf[0,2], tf[1,0], tf[1,2]
    double base[3], total[3];     // P_j[a] and P_j[b] - P_j[a] = sum of 1/M over (a, b]  (:41)
};

__device__ __forceinline__ void sgd_load_edge(const SgdArgs &a, int e, SgdEdge &x)
{
    x.ea = a.edges[2 * e]; x.eb = a.edges[2 * e + 1];
    const double *T = a.tf + 6 * e;
    x.t0 = T[0]; x.t2 = T[2]; x.t3 = T[3]; x.t5 = T[5];
    const double *pb = a.PB + 6 * e;
#pragma unroll
    for (int j = 0; j < 3; ++j) { x.base[j] = pb[j]; x.total[j] = pb[3 + j]; }
}

// Residual and clipped step of one edge under the current poses (:33-44): beta_j / total_j and
// beta_j, the two factors the node updates need.
__device__ __forceinline__ void sgd_edge_step(const SgdEdge &x, const double *pose_a, const double *pose_b,
                                              const double *alpha, double lcu, double inv_lcu,
                                              double *coef, double *tot)
{
    const double two_pi = 6.283185307179586;                   // 2 * np.pi
    const double pax = pose_a[0], pay = pose_a[1], pat = pose_a[2];
    const double pbx = pose_b[0], pby = pose_b[1], pbt = pose_b[2];
    double s, c;
    sincos(pat, &s, &c);
    // Pb_new = pose_to_mat(poses[a]) @ tf (:33), r = mat_to_pose(Pb_new) - poses[b] (:34-35)
    const double m00 = c * x.t0 + -s * x.t3;
    const double m10 = s * x.t0 + c * x.t3;
    double r[3];
    r[0] = (c * x.t2 + -s * x.t5 + pax) - pbx;
    r[1] = (s * x.t2 + c * x.t5 + pay) - pby;
    r[2] = mod_pos(atan2(m10, m00) - pbt, two_pi);
    // d = 2 inv(R^T sigma R) r (:36); one reciprocal of the determinant instead of four divisions
    const double cl = c * lcu, sl = s * lcu, nsl = -s * lcu;
    const double a00 = cl * c + sl * s,   a01 = cl * -s + sl * c;
    const double a10 = nsl * c + cl * s,  a11 = nsl * -s + cl * c;
    const double rdet = 1.0 / (a00 * a11 - a01 * a10);
    double d[3];
    d[0] = 2.0 * ((a11 * rdet) * r[0] + (-a01 * rdet) * r[1]);
    d[1] = 2.0 * ((-a10 * rdet) * r[0] + (a00 * rdet) * r[1]);
    d[2] = 2.0 * (inv_lcu * r[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double beta = (double)(x.eb - x.ea) * d[j] * alpha[j];                   // :42
        if (fabs(beta) > fabs(r[j])) beta = r[j];                                // :43-44
        coef[j] = beta / x.total[j]; tot[j] = coef[j] * x.total[j];
    }
}

// The host passes only edges with b > a + 1 (the others move nothing, see icpb_pose_graph_sgd).
//
// Warp 0 is the *scalar warp*: while the other warps add edge e's update to the nodes, it brings
// the two endpoints of edge e+1 up to date itself (the sweeping warps leave those two nodes alone)
// and evaluates edge e+1's step from them -- the trigonometry of the next edge overlaps the sweep
// of this one, and a single barrier per edge orders both.
//
// Poses live in shared memory: 24 B per node, up to 9,600 nodes in one CTA.  CLUSTER: a thread-block
// cluster of up to 8 CTAs holds `slice` consecutive nodes per CTA; every CTA sweeps its own slice,
// the scalar warp (CTA 0) reads and updates the next edge's endpoints over distributed shared
// memory and stores each step into every CTA's shared memory; the per-edge barrier is the cluster
// barrier.  Beyond 8 x 9,600 nodes the poses stay in global memory (poses_in_smem = 0).
constexpr int kSgdThreads = 512;
template <bool CLUSTER>
__global__ void __launch_bounds__(kSgdThreads, 1)
sgd_chain_kernel(const SgdArgs a)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char sgd_smem[];   // scan scratch, then the poses
    __shared__ double s_best[32];
    __shared__ int    s_beste[32];
    __shared__ double s_gamma[3];
    __shared__ double s_step[2][6];                            // per edge parity: coef[3], tot[3]
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int crank = 0, csize = 1;
    if (CLUSTER) {
        crank = (int)cg::this_cluster().block_rank();
        csize = (int)cg::this_cluster().num_blocks();
    }
    auto sync_all = [&]() { if (CLUSTER) cg::this_cluster().sync(); else __syncthreads(); };
    const int slice = CLUSTER ? a.slice : n;                   // nodes per CTA
    const int lo = crank * slice, hi = min(lo + slice, n);     // this CTA's nodes

    // ---- gamma: the first edge with the smallest |diag(W)|^2 (strict > in :23) ----
    {
        double best = inf;
        int be = 0x7fffffff;
        for (int e = tid; e < a.E; e += NT) {
            const double v = a.dW[4 * e + 3];
            if (v < best) { best = v; be = e; }                // e ascends: first index kept
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oe = __shfl_xor_sync(0xffffffffu, be, o);
            if (ov < best || (ov == best && oe < be)) { best = ov; be = oe; }
        }
        if (lane == 0) { s_best[warp] = best; s_beste[warp] = be; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (NT >> 5); ++w)
                if (s_best[w] < best || (s_best[w] == best && s_beste[w] < be)) { best = s_best[w]; be = s_beste[w]; }
            const bool any = be != 0x7fffffff && best < inf;
            for (int j = 0; j < 3; ++j) s_gamma[j] = any ? a.dW[4 * be + j] : inf;
        }
    }
    if (crank == 0) {
        // ---- P_j[i] = sum_{k <= i} 1/M[k,j]: contiguous slices per thread, then the slice offsets ----
        double (*s_part)[kSgdThreads] = reinterpret_cast<double (*)[kSgdThreads]>(sgd_smem);     // [3][threads]
        const int L = (n + NT - 1) / NT;
        const int i0 = min(tid * L, n), i1 = min(i0 + L, n);
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = i0; i < i1; ++i)
            for (int j = 0; j < 3; ++j) {
                const double m = a.M[3 * i + j];
                acc[j] += m > 0.0 ? 1.0 / m : 0.0;             // uncovered nodes never enter a range
                a.P[3 * i + j] = acc[j];
            }
        for (int j = 0; j < 3; ++j) s_part[j][tid] = acc[j];
        __syncthreads();
        if (tid < 3) {                                         // exclusive scan of the slice totals
            double run = 0.0;
            for (int t = 0; t < NT; ++t) { const double v = s_part[tid][t]; s_part[tid][t] = run; run += v; }
        }
        __syncthreads();
        for (int i = i0; i < i1; ++i)
            for (int j = 0; j < 3; ++j) a.P[3 * i + j] += s_part[j][tid];
        __syncthreads();
        // ---- per edge: P_j[a] and the range total (independent of the poses) ----
        for (int e = tid; e < a.E; e += NT) {
            const int ea = a.edges[2 * e], eb = a.edges[2 * e + 1];
            for (int j = 0; j < 3; ++j) {
                const double pa = a.P[3 * ea + j];
                a.PB[6 * e + j] = pa; a.PB[6 * e + 3 + j] = a.P[3 * eb + j] - pa;
            }
        }
    }
    sync_all();                                                // P and PB are visible to every CTA
    // pz + 3 * (i - lo): this CTA's copy of node i (shared memory, or global memory when too large)
    double *pz = a.poses_in_smem ? reinterpret_cast<double *>(sgd_smem) : a.poses + 3 * lo;
    if (a.poses_in_smem)
        for (int k = tid; k < 3 * (hi - lo); k += NT) pz[k] = a.poses[3 * lo + k];
    // any node, wherever it lives
    auto node = [&](int i) -> double * {
        if (!CLUSTER) return pz + 3 * i;
        const int r = i / slice;
        return cg::this_cluster().map_shared_rank(pz, r) + 3 * (i - r * slice);
    };
    auto publish = [&](int parity, const double *coef, const double *tot) {     // scalar warp, all lanes
        if (!CLUSTER) {
            if (lane < 3) { s_step[parity][lane] = coef[lane]; s_step[parity][3 + lane] = tot[lane]; }
        } else {
            for (int r = 0; r < csize; ++r) {
                double *dst = cg::this_cluster().map_shared_rank(&s_step[0][0], r) + 6 * parity;
                if (lane < 3) { dst[lane] = coef[lane]; dst[3 + lane] = tot[lane]; }
            }
        }
    };
    sync_all();
    // step factors that do not depend on the moving poses: alpha_j = lr / gamma_j (:39-40)
    double alpha[3];
    for (int j = 0; j < 3; ++j) alpha[j] = (1.0 / s_gamma[j]) * a.learning_rate;
    const double inv_lcu = 1.0 / a.lcu;
    const double *__restrict__ P = a.P;

    // ---- the edges, in order (:27-48); every thread fetches the next edge's record during this one ----
    SgdEdge cur, nxt;
    if (a.E > 0) {
        sgd_load_edge(a, 0, cur);
        if (warp == 0 && crank == 0) {
            double coef[3], tot[3];
            sgd_edge_step(cur, node(cur.ea), node(cur.eb), alpha, a.lcu, inv_lcu, coef, tot);
            publish(0, coef, tot);
        }
    }
    sync_all();
    const int sweepers = NT - 32;
    for (int e = 0; e < a.E; ++e) {
        const int ea = cur.ea, eb = cur.eb;
        const bool more = e + 1 < a.E;
        if (more) sgd_load_edge(a, e + 1, nxt);
        double coef[3], tot[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { coef[j] = s_step[e & 1][j]; tot[j] = s_step[e & 1][3 + j]; }
        const int na = more ? nxt.ea : -1, nb = more ? nxt.eb : -1;
        // node i > a receives beta/total * sum_{k in (a, min(i,b)]} 1/M[k], the running `dpose` (:45-48)
        if (warp != 0) {
            for (int i0 = max(ea + 1, lo) + (tid - 32); i0 < hi; i0 += 2 * sweepers) {
                const int i1 = i0 + sweepers;
                const bool in1 = i1 < hi;
                const bool do0 = i0 != na && i0 != nb, do1 = in1 && i1 != na && i1 != nb;
                double inc0[3], inc1[3], v0[3], v1[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    inc0[j] = i0 <= eb ? coef[j] * (P[3 * i0 + j] - cur.base[j]) : tot[j];
                    inc1[j] = in1 && i1 <= eb ? coef[j] * (P[3 * i1 + j] - cur.base[j]) : tot[j];
                    v0[j] = pz[3 * (i0 - lo) + j];
                    v1[j] = in1 ? pz[3 * (i1 - lo) + j] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (do0) pz[3 * (i0 - lo) + j] = v0[j] + inc0[j];
                    if (do1) pz[3 * (i1 - lo) + j] = v1[j] + inc1[j];
                }
            }
        } else if (more && crank == 0) {
            // the next edge's endpoints first (lanes 0 and 1), then its step from the updated poses
            if (lane < 2) {
                const int i = lane == 0 ? na : nb;
                if (i > ea) {
                    double *q = node(i);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        q[j] += i <= eb ? coef[j] * (P[3 * i + j] - cur.base[j]) : tot[j];
                }
            }
            __syncwarp();
            double ncoef[3], ntot[3];
            sgd_edge_step(nxt, node(na), node(nb), alpha, a.lcu, inv_lcu, ncoef, ntot);
            publish((e + 1) & 1, ncoef, ntot);
        }
        sync_all();
        cur = nxt;
    }
    if (a.poses_in_smem)
        for (int k = tid; k < 3 * (hi - lo); k += NT) a.poses[3 * lo + k] = pz[k];
}

}  // namespace icpb
