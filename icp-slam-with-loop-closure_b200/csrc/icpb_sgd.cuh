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
//   * sgd_chain_kernel: ONE CTA walks the edges lazily: every edge leaves a record, the endpoints
//     of the next edge are evaluated from the start-of-pass poses and the records so far (see the
//     kernel's own comment); node i > a receives beta_j/total_j * (P_j[min(i,b)] - P_j[a]) -- the
//     reference's running sum `dpose` in closed form over the prefix sums P_j[i] = sum_{k<=i} 1/M[k,j];
//   * sgd_apply_kernel: one thread per node applies all records, in edge order.
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
    double        *PB;        // E x 10 scratch: per edge P_j[a], P_j[b] - P_j[a], its reciprocal, atan2(tf[1,0], tf[0,0])
    double        *REC;       // E x 10 scratch: the per-edge records of the lazy chain (SgdRecords)
    double        *ES;        // E x 23 scratch: the per-edge inputs of the chain (SgdEdgeFull)
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
// equals the exact fmod-based result except when x/m rounds across an integer (fixed up below; the
// quotient comes from a multiplication by 1/m, which can only move it by one in that same case).
__device__ __forceinline__ double mod_pos(double x, double m, double inv_m)
{
    const double q = floor(x * inv_m);
    double r = fma(-q, m, x);
    if (r < 0.0) r += m;
    if (r >= m) r -= m;
    return r;
}

// what one edge needs that does not depend on the moving poses
struct SgdEdge {
    int ea, eb;
    double t2, t5, phi;           // tf[0,2], tf[1,2], atan2(tf[1,0], tf[0,0])
    double base[3], total[3], itot[3];   // P_j[a], P_j[b] - P_j[a] = sum of 1/M over (a, b] (:41), 1 / that
};

__device__ __forceinline__ void sgd_load_edge(const SgdArgs &a, int e, SgdEdge &x)
{
    x.ea = a.edges[2 * e]; x.eb = a.edges[2 * e + 1];
    const double *T = a.tf + 6 * e;
    x.t2 = T[2]; x.t5 = T[5];
    const double *pb = a.PB + 10 * e;
#pragma unroll
    for (int j = 0; j < 3; ++j) { x.base[j] = pb[j]; x.total[j] = pb[3 + j]; x.itot[j] = pb[6 + j]; }
    x.phi = pb[9];
}

// Residual and clipped step of one edge under the current poses (:33-44): beta_j / total_j and
// beta_j, the two factors the node updates need.  This sits on the critical chain of the pass (edge
// e+1 reads what edge e wrote), so everything that does not depend on the moving poses has been taken
// off it, using two identities that hold to rounding (1e-16 relative; the contract is 1e-9, and the
// goldens of the unmodified reference agree to 1e-12, tests/test_gpu_sgd.py):
//   * heading of Pb_new = pose_to_mat(poses[a]) @ tf (:33-34): atan2 of a product of rotations is the
//     sum of their angles mod 2 pi, so atan2(tf[1,0], tf[0,0]) is computed once per edge beforehand;
//   * inv(R^T sigma R) with sigma = lcu I (:36) is I / lcu whatever R is.
// What remains per edge: one sincos, a handful of multiply-adds, the clip, three multiplications.
__device__ __forceinline__ void sgd_edge_step(const SgdEdge &x, const double *pose_a, const double *pose_b,
                                              const double *alpha, double inv_lcu, double *coef, double *tot)
{
    const double two_pi = 6.283185307179586, inv_two_pi = 0.15915494309189535;   // 2 * np.pi
    const double pax = pose_a[0], pay = pose_a[1], pat = pose_a[2];
    const double pbx = pose_b[0], pby = pose_b[1], pbt = pose_b[2];
    double s, c;
    sincos(pat, &s, &c);
    double r[3];
    r[0] = (c * x.t2 + -s * x.t5 + pax) - pbx;                 // translation of Pb_new minus poses[b]
    r[1] = (s * x.t2 + c * x.t5 + pay) - pby;
    r[2] = mod_pos((pat + x.phi) - pbt, two_pi, inv_two_pi);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double d = 2.0 * (inv_lcu * r[j]);                                 // :36
        double beta = (double)(x.eb - x.ea) * d * alpha[j];                      // :42
        if (fabs(beta) > fabs(r[j])) beta = r[j];                                // :43-44
        coef[j] = beta * x.itot[j]; tot[j] = coef[j] * x.total[j];
    }
}

// The host passes only edges with b > a + 1 (the others move nothing, see icpb_pose_graph_sgd).
//
// LAZY chain.  Edge e adds to node i > a_e the amount f_e(i) = coef_e (P[min(i, b_e)] - P[a_e]) (per
// dof).  The only poses the chain itself ever needs are the two endpoints of the edge it is about to
// evaluate, and pose(i) at that moment = pose0(i) + sum over the edges processed so far of f_e'(i).
// So nothing is swept per edge: every edge leaves a 80-byte record (a, b, coef, P[a], P[b]); the
// endpoints of the next edge are evaluated from pose0 and the records -- 480 threads take the
// records strided, a fixed-order tree adds the partial sums -- and a separate kernel
// (sgd_apply_kernel, one thread per node, all SMs) applies every record to every node at the end, in
// edge order per node, which is the reference's own order of additions.  The work per edge no
// longer depends on the number of poses, the poses never have to fit shared memory (no cluster
// path), and the pass costs O(E^2 / 480 + E) steps on one SM plus O(N E) fully parallel work instead
// of O(N E / 480) steps on the critical chain.
//
// Pipelining: warp 0 is the *scalar warp*.  While it evaluates edge e (sum of the partials the other
// warps produced during the previous edge + the one record, e - 1, they could not yet see; sincos;
// clip; record e), warps 1..15 already add up records 0..e-1 for the endpoints of edge e + 1.  One
// barrier per edge.
constexpr int kSgdThreads = 512;
constexpr int kSgdWarps = kSgdThreads / 32;

struct SgdRecords {          // structure of arrays, E entries each (global memory)
    int2   *ab;
    double *coef, *pa, *pb;  // [3][E]
};

__device__ __forceinline__ SgdRecords sgd_records(const SgdArgs &a)
{
    SgdRecords r;
    r.coef = a.REC; r.pa = a.REC + 3 * (size_t)a.E; r.pb = a.REC + 6 * (size_t)a.E;
    r.ab = reinterpret_cast<int2 *>(a.REC + 9 * (size_t)a.E);
    return r;
}

// f_e(i) for one record and one node, added to acc[3] (product and sum rounded separately, like the
// reference's `dpose += ...; poses[i] += dpose`)
__device__ __forceinline__ void sgd_add_term(int i, const double *Pi, int ra, int rb, const double *coef,
                                             const double *pa, const double *pb, double *acc)
{
    if (i > ra) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            acc[j] = __dadd_rn(acc[j], __dmul_rn(coef[j], (i <= rb ? Pi[j] : pb[j]) - pa[j]));
    }
}

// Everything edge e needs that does not depend on the moving poses, gathered once per pass into one
// contiguous 184-byte struct, so that the chain only ever issues one round of independent loads per
// edge -- and issues it one edge ahead, under the arithmetic of the current edge.
struct SgdEdgeFull {
    double ea, eb;                // node ids (exact in a double)
    double t2, t5, phi;
    double base[3], total[3], itot[3];
    double p0a[3], p0b[3];        // poses at the start of the pass
    double Pb[3];                 // P[b] (P[a] is base)
};
static_assert(sizeof(SgdEdgeFull) == 184, "SgdEdgeFull layout");

__global__ void __launch_bounds__(kSgdThreads, 1)
sgd_chain_kernel(const SgdArgs a)
{
    __shared__ double s_best[32];
    __shared__ int    s_beste[32];
    __shared__ double s_gamma[3];
    __shared__ double s_scan[3][kSgdThreads];
    __shared__ double s_part[2][kSgdWarps][6];                 // per edge parity: warp partial sums, 2 nodes x 3 dofs
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const SgdRecords R = sgd_records(a);
    SgdEdgeFull *ES = reinterpret_cast<SgdEdgeFull *>(a.ES);

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
    // ---- P_j[i] = sum_{k <= i} 1/M[k,j]: contiguous slices per thread, then the slice offsets ----
    {
        const int L = (n + NT - 1) / NT;
        const int i0 = min(tid * L, n), i1 = min(i0 + L, n);
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = i0; i < i1; ++i)
            for (int j = 0; j < 3; ++j) {
                const double m = a.M[3 * i + j];
                acc[j] += m > 0.0 ? 1.0 / m : 0.0;             // uncovered nodes never enter a range
                a.P[3 * i + j] = acc[j];
            }
        for (int j = 0; j < 3; ++j) s_scan[j][tid] = acc[j];
        __syncthreads();
        if (tid < 3) {                                         // exclusive scan of the slice totals
            double run = 0.0;
            for (int t = 0; t < NT; ++t) { const double v = s_scan[tid][t]; s_scan[tid][t] = run; run += v; }
        }
        __syncthreads();
        for (int i = i0; i < i1; ++i)
            for (int j = 0; j < 3; ++j) a.P[3 * i + j] += s_scan[j][tid];
        __syncthreads();
        // ---- per edge: everything the chain needs, in one struct; the static part of its record ----
        for (int e = tid; e < a.E; e += NT) {
            const int ea = a.edges[2 * e], eb = a.edges[2 * e + 1];
            SgdEdgeFull x;
            x.ea = (double)ea; x.eb = (double)eb;
            x.t2 = a.tf[6 * e + 2]; x.t5 = a.tf[6 * e + 5];
            x.phi = atan2(a.tf[6 * e + 3], a.tf[6 * e]);
            for (int j = 0; j < 3; ++j) {
                const double pa = a.P[3 * ea + j], pb = a.P[3 * eb + j], tot = pb - pa;
                x.base[j] = pa; x.total[j] = tot; x.itot[j] = 1.0 / tot; x.Pb[j] = pb;
                x.p0a[j] = a.poses[3 * ea + j]; x.p0b[j] = a.poses[3 * eb + j];
                R.pa[j * (size_t)a.E + e] = pa; R.pb[j * (size_t)a.E + e] = pb;
            }
            ES[e] = x;
            R.ab[e] = make_int2(ea, eb);
        }
    }
    __syncthreads();
    // step factors that do not depend on the moving poses: alpha_j = lr / gamma_j (:39-40)
    double alpha[3];
    for (int j = 0; j < 3; ++j) alpha[j] = (1.0 / s_gamma[j]) * a.learning_rate;
    const double inv_lcu = 1.0 / a.lcu;

    if (a.E == 0) return;
    // The edge structs reach the chain through a 4-slot ring in shared memory that the last warp (idle
    // otherwise) fills two edges ahead: lane k copies double k of the 184-byte struct, one coalesced
    // load per edge, off everybody's critical path.
    __shared__ SgdEdgeFull s_es[4];
    if (warp == kSgdWarps - 1 && lane < 23)
        for (int k = 0; k < 2 && k < a.E; ++k)
            reinterpret_cast<double *>(&s_es[k])[lane] = reinterpret_cast<const double *>(ES + k)[lane];
    __syncthreads();
    double last_coef[3] = {0.0, 0.0, 0.0};                     // warp 0: record e - 1, the one the partials lack
    double last_pa[3] = {0.0, 0.0, 0.0}, last_pb[3] = {0.0, 0.0, 0.0};
    int last_a = 0x7fffffff, last_b = 0;
    for (int e = 0; e < a.E; ++e) {
        // Worker warps: as many as keep every lane at <= 4 records (the partial sums over e records cost
        // one warp reduction per worker, and all 16 warps share one SM's issue slots with the scalar
        // warp: with 534 edges, 15 workers made the pass issue-bound at 2,700 cycles per edge)
        const int workers = min(kSgdWarps - 2, max(1, (e + 127) / 128));
        if (warp == kSgdWarps - 1) {
            if (lane < 23 && e + 2 < a.E)                        // slot (e + 2) % 4 was last read in iteration e - 1
                reinterpret_cast<double *>(&s_es[(e + 2) & 3])[lane] = reinterpret_cast<const double *>(ES + e + 2)[lane];
        } else if (warp == 0) {
            const SgdEdgeFull &cur = s_es[e & 3];
            // ---- evaluate edge e: its endpoints = start-of-pass poses + the partial sums over records
            // 0..e-2 (warps 1..15, previous iteration) + record e-1 (kept in registers) ----
            const int ea = (int)cur.ea, eb = (int)cur.eb;
            double pose_a[3], pose_b[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { pose_a[j] = cur.p0a[j]; pose_b[j] = cur.p0b[j]; }
            if (e > 0) {
                // second stage of the tree: lane k < 6 adds value k of the workers' partials, in warp order
                // (the workers of the previous iteration: the same formula with e - 1)
                const int prev_workers = min(kSgdWarps - 2, max(1, (e - 1 + 127) / 128));
                double v = 0.0;
                if (lane < 6)
                    for (int w = 1; w <= prev_workers; ++w) v += s_part[e & 1][w][lane];
                double sa[3], sb[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    sa[j] = __shfl_sync(0xffffffffu, v, j);
                    sb[j] = __shfl_sync(0xffffffffu, v, 3 + j);
                }
                sgd_add_term(ea, cur.base, last_a, last_b, last_coef, last_pa, last_pb, sa);
                sgd_add_term(eb, cur.Pb, last_a, last_b, last_coef, last_pa, last_pb, sb);
#pragma unroll
                for (int j = 0; j < 3; ++j) { pose_a[j] += sa[j]; pose_b[j] += sb[j]; }
            }
            SgdEdge x;
            x.ea = ea; x.eb = eb; x.t2 = cur.t2; x.t5 = cur.t5; x.phi = cur.phi;
#pragma unroll
            for (int j = 0; j < 3; ++j) { x.base[j] = cur.base[j]; x.total[j] = cur.total[j]; x.itot[j] = cur.itot[j]; }
            double coef[3], tot[3];
            sgd_edge_step(x, pose_a, pose_b, alpha, inv_lcu, coef, tot);
            if (lane < 3) R.coef[lane * (size_t)a.E + e] = coef[lane];
#pragma unroll
            for (int j = 0; j < 3; ++j) { last_coef[j] = coef[j]; last_pa[j] = cur.base[j]; last_pb[j] = cur.Pb[j]; }
            last_a = ea; last_b = eb;
        } else if (e + 1 < a.E && warp <= workers) {
            // ---- partial sums for the endpoints of edge e+1 over records 0..e-1 (all complete) ----
            const SgdEdgeFull &cur = s_es[(e + 1) & 3];
            const int na = (int)cur.ea, nb = (int)cur.eb;
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int r = tid - 32; r < e; r += 32 * workers) {
                const int2 ab = R.ab[r];
                double coef[3], pa[3], pb[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    coef[j] = R.coef[j * (size_t)a.E + r]; pa[j] = R.pa[j * (size_t)a.E + r]; pb[j] = R.pb[j * (size_t)a.E + r];
                }
                sgd_add_term(na, cur.base, ab.x, ab.y, coef, pa, pb, acc);
                sgd_add_term(nb, cur.Pb, ab.x, ab.y, coef, pa, pb, acc + 3);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k)
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            if (lane < 6) {
                double v = acc[0];
#pragma unroll
                for (int k = 1; k < 6; ++k) v = lane == k ? acc[k] : v;
                s_part[(e + 1) & 1][warp][lane] = v;
            }
        }
        __syncthreads();           // record e, the partials for edge e+1 and the struct of edge e+2 are visible
    }
}

// Every record applied to every node, in edge order per node: pose(i) += f_0(i), += f_1(i), ... --
// the reference's own sequence of additions for that node (:45-48).  One thread per node.
__global__ void __launch_bounds__(256)
sgd_apply_kernel(const SgdArgs a)
{
    __shared__ int2 s_ab[128];
    __shared__ double s_c[3][128], s_pa[3][128], s_pb[3][128];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const SgdRecords R = sgd_records(a);
    double p[3] = {0.0, 0.0, 0.0}, Pi[3] = {0.0, 0.0, 0.0};
    if (i < a.n) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { p[j] = a.poses[3 * i + j]; Pi[j] = a.P[3 * i + j]; }
    }
    for (int e0 = 0; e0 < a.E; e0 += 128) {                     // the records stream through shared memory
        const int cnt = min(128, a.E - e0);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) s_ab[k] = R.ab[e0 + k];
        for (int k = threadIdx.x; k < 3 * cnt; k += blockDim.x) {
            const int j = k / cnt, r = k - j * cnt;
            s_c[j][r] = R.coef[j * (size_t)a.E + e0 + r];
            s_pa[j][r] = R.pa[j * (size_t)a.E + e0 + r];
            s_pb[j][r] = R.pb[j * (size_t)a.E + e0 + r];
        }
        __syncthreads();
        for (int k = 0; k < cnt; ++k) {                         // edge order = the reference's order of additions
            const int2 ab = s_ab[k];
            if (i > ab.x) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double v = (i <= ab.y ? Pi[j] : s_pb[j][k]) - s_pa[j][k];
                    p[j] = __dadd_rn(p[j], __dmul_rn(s_c[j][k], v));
                }
            }
        }
        __syncthreads();
    }
    if (i < a.n) {
#pragma unroll
        for (int j = 0; j < 3; ++j) a.poses[3 * i + j] = p[j];
    }
}

}  // namespace icpb
