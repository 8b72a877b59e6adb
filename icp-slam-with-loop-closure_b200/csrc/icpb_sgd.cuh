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
//   * sgd_accumulate_kernel: one thread per (node, dof) adds the diagonals of the edges covering it, in
//     edge order (the reference's order of additions, so M has the reference's bits given the
//     same diagonals); the edge list streams through shared memory;
//   * sgd_chain_kernel: ONE CTA walks the edges lazily: every edge leaves a record, and only the
//     endpoints of the edges still to come are kept up to date (see the kernel's own comment);
//     node i > a receives beta_j/total_j * (P_j[min(i,b)] - P_j[a]) -- the
//     reference's running sum `dpose` in closed form over the prefix sums P_j[i] = sum_{k<=i} 1/M[k,j];
//   * sgd_apply_kernel: one thread per (node, dof) applies all records, in edge order.
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
    double        *REC;       // E x 10 scratch: the per-edge records of the lazy chain (SgdRecords)
    double        *RECA;      // E x 10 scratch: the same records as an array of structs (the chain's catch-up bursts)
    double        *ES;        // 23 x E scratch: the per-edge inputs of the chain (the fields of SgdEdgeFull, one array each)
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

// The per-node kernels (this one and sgd_apply_kernel) walk all E edges per node in edge order, so
// their parallelism is the number of nodes: small CTAs (kSgdNodeThreads) put the few thousand nodes
// of a map on as many SMs as possible, one or two warps per scheduler.
constexpr int kSgdNodeThreads = 64;
constexpr int kSgdNodeTile = 512;                          // edges per shared-memory tile (one global round trip each)

__global__ void __launch_bounds__(kSgdNodeThreads)
sgd_accumulate_kernel(const SgdArgs a)
{
    __shared__ int2 s_ab[kSgdNodeTile];
    __shared__ double s_w[kSgdNodeTile];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;                                  // dof: one (node, dof) per thread
    double m = 0.0;
    for (int e0 = 0; e0 < a.E; e0 += kSgdNodeTile) {
        for (int k = threadIdx.x; k < kSgdNodeTile && e0 + k < a.E; k += blockDim.x) {
            const int e = e0 + k;
            int ea = a.edges[2 * e], eb = a.edges[2 * e + 1];
            if (sgd_skipped(ea, eb)) eb = ea;                  // empty range
            s_ab[k] = make_int2(ea, eb);
            s_w[k] = a.dW[4 * e + j];
        }
        __syncthreads();
        const int cnt = min(kSgdNodeTile, a.E - e0);
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {                        // edge order = the reference's order of additions
            const int2 ab = s_ab[k];
            const double q = m + s_w[k];
            m = (ab.x < i && i <= ab.y) ? q : m;                 // select, not branch: the loads run ahead
        }
        __syncthreads();
    }
    if (i < a.n) a.M[3 * i + j] = m;
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

// The residual and clipped step of one edge under the current poses (:33-44) sit on the critical chain
// of the pass (edge e+1 reads what edge e wrote), so everything that does not depend on the moving
// poses is taken off it, using identities that hold to rounding (1e-16 relative; the contract is
// 1e-9, and the goldens of the unmodified reference agree to 1e-12, tests/test_gpu_sgd.py):
//   * heading of Pb_new = pose_to_mat(poses[a]) @ tf (:33-34): atan2 of a product of rotations is the
//     sum of their angles mod 2 pi, so atan2(tf[1,0], tf[0,0]) is computed once per edge beforehand;
//   * inv(R^T sigma R) with sigma = lcu I (:36) is I / lcu whatever R is.
// What remains per edge: one sincos, a dozen multiply-adds, the clip, three multiplications.  (The
// heading of a moves by whole radians within a pass -- the reference never wraps the residual's
// [0, 2 pi) back -- so there is no small-angle shortcut for the sincos.)

// The host passes only edges with b > a + 1 (the others move nothing, see icpb_pose_graph_sgd).
//
// LAZY chain.  Edge e adds to node i > a_e the amount f_e(i) = coef_e (P[min(i, b_e)] - P[a_e]) (per
// dof).  The only poses the chain itself ever needs are the two endpoints of the edge it is about to
// evaluate, and those are known before the pass starts: the 2E endpoints are the chain's *slots*.
// A slot carries the running pose of its node -- pose0(i) + f_0(i) + f_1(i) + ..., the reference's
// own sequence of additions for that node.  A separate kernel (sgd_apply_kernel, all SMs) applies
// every record to every node at the end, in the same order, so the chain's endpoint values are the
// bits the apply kernel produces for those nodes.  The work per edge does not depend on the number
// of poses, and neither the poses nor the slots have to fit shared memory.
//
// Slots live in REGISTERS.  The edges are cut into blocks of kSgdBlock = 240; a block has 480 slots,
// one per thread of warps 1..15, and a thread holds two: its slot of the block the chain is in and of
// the next one.  Every iteration the owners add the newest record to both (two slots x three dofs,
// no memory traffic but the record) and the owners of the next edge's endpoints publish them.  When
// the chain enters a new block, every owner loads its slot of the block after it and catches it up
// on all records so far in one burst -- 80-byte records stream through shared-memory tiles (one
// coalesced read per tile, fetched a tile ahead), the running pose stays in registers.  Total work
// O(E^2 / 480) like any lazy scheme, but no reductions and no per-edge memory round trips, any E.
// (Spreading the catch-up over the iterations of the block instead -- a few records per edge -- lost:
// every iteration then waits for an L2 round trip that the burst amortises over hundreds of records.)
//
// Pipelining: warp 0 is the *scalar warp*.  While it evaluates edge e (published endpoint poses +
// record e-1, which the slots did not have yet when they were published; sincos; clip; record e),
// warps 1..15 add record e-1 to their slots and publish the endpoints of edge e+1.  One barrier per edge.
//
// (Round-2 history, measured with a cycle probe: a first lazy version summed all earlier records per
// endpoint with a warp reduction per worker warp -- the busiest worker, not the scalar warp, set the
// pace at 1,780 of 1,910 cycles per edge; slots in shared / global memory: 1,100 cycles per edge up
// to 2,000 edges and 4.3 us per edge beyond.)
constexpr int kSgdThreads = 512;
constexpr int kSgdWarps = kSgdThreads / 32;
constexpr int kSgdOwners = kSgdThreads - 32;                   // threads that own slots (warps 1..15)
constexpr int kSgdBlock = kSgdOwners / 2;                      // edges per block: one slot per owner and block
constexpr int kSgdTile = kSgdOwners / 5;                       // records per catch-up tile: one 16-byte piece per owner

struct SgdRecords {          // structure of arrays, E entries each (global memory; read by sgd_apply_kernel)
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

// Everything edge e needs that does not depend on the moving poses, gathered once per pass into one
// 184-byte struct (stored field by field in global memory) and handed to the chain through a ring in
// shared memory.
struct SgdEdgeFull {
    int ea, eb;                   // node ids
    double span;                  // (double)(eb - ea)
    double t2, t5, phi;
    double base[3], total[3], itot[3];
    double p0a[3], p0b[3];        // poses at the start of the pass
    double Pb[3];                 // P[b] (P[a] is base)
};
constexpr int kSgdEdgeDoubles = 23;
static_assert(sizeof(SgdEdgeFull) == 8 * kSgdEdgeDoubles, "SgdEdgeFull layout");
// positions of the fields in the struct seen as an array of doubles (the scalar warp indexes by lane)
enum { kEsSpan = 1, kEsT2 = 2, kEsT5 = 3, kEsPhi = 4, kEsBase = 5, kEsTotal = 8, kEsItot = 11,
       kEsP0a = 14, kEsP0b = 17, kEsPb = 20 };


__global__ void __launch_bounds__(kSgdThreads, 1)
sgd_chain_kernel(const SgdArgs a)
{
    __shared__ double s_best[32];
    __shared__ int    s_beste[32];
    __shared__ double s_gamma[3];
    __shared__ double s_wtot[3][kSgdWarps];
    __shared__ double s_next[2][6];                            // per edge parity: published poses of its endpoints
    __shared__ double s_coef[2][4];                            // per edge parity: the coefficients of its record
    __shared__ double2 s_tile[2][kSgdOwners];                  // catch-up bursts: two tiles of kSgdTile 80-byte records
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const SgdRecords R = sgd_records(a);
    // field f of edge e (SgdEdgeFull seen as 23 doubles) lives at ES[f * E + e]: the threads of a warp fill
    // the structs of 32 consecutive edges with coalesced stores
    double *ES = a.ES;
    const size_t ES_stride = (size_t)a.E;

#ifdef ICPB_SGD_PROBE
    const long long pq0 = clock64();
#endif
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
#ifdef ICPB_SGD_PROBE
    const long long pq1 = clock64();
    long long pq3 = 0, pq4 = 0;
#endif
    // ---- P_j[i] = sum_{k <= i} 1/M[k,j]: chunks of NT consecutive nodes, one node per thread (coalesced
    // loads and stores; thread-contiguous slices cost a sector per value), an inclusive scan per chunk
    // -- shuffles within a warp, the warp totals through shared memory, both in a fixed order -- on top of
    // the running total of the chunks before it.  The next chunk's M is fetched before this one is scanned.
    {
        double carry[3] = {0.0, 0.0, 0.0};
        double mn[3];
        for (int j = 0; j < 3; ++j) mn[j] = tid < n ? __ldg(a.M + 3 * tid + j) : 0.0;
        for (int c0 = 0; c0 < n; c0 += NT) {
            const int i = c0 + tid;
            double v[3];
            for (int j = 0; j < 3; ++j) v[j] = mn[j] > 0.0 ? 1.0 / mn[j] : 0.0;   // uncovered nodes never enter a range
            if (c0 + NT < n)
                for (int j = 0; j < 3; ++j) mn[j] = i + NT < n ? __ldg(a.M + 3 * (i + NT) + j) : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double t = __shfl_up_sync(0xffffffffu, v[j], o);
                    if (lane >= o) v[j] += t;
                }
            if (lane == 31)
                for (int j = 0; j < 3; ++j) s_wtot[j][warp] = v[j];
            __syncthreads();
            double off[3], tot[3];
            for (int j = 0; j < 3; ++j) { off[j] = carry[j]; tot[j] = carry[j]; }
            for (int w = 0; w < kSgdWarps; ++w)
                for (int j = 0; j < 3; ++j) {
                    const double t = s_wtot[j][w];
                    tot[j] += t;
                    if (w < warp) off[j] += t;
                }
            if (i < n)
                for (int j = 0; j < 3; ++j) a.P[3 * i + j] = off[j] + v[j];
            for (int j = 0; j < 3; ++j) carry[j] = tot[j];
            __syncthreads();                                   // s_wtot is rewritten by the next chunk
        }
#ifdef ICPB_SGD_PROBE
        pq3 = clock64();
#endif
        // ---- per edge: everything the chain needs in one struct, the static part of its record ----
        for (int e = tid; e < a.E; e += NT) {
            // every load before the first store (the compiler has to assume that the outputs overlap the inputs)
            const int ea = __ldg(a.edges + 2 * e), eb = __ldg(a.edges + 2 * e + 1);
            const double t0 = __ldg(a.tf + 6 * e), t2 = __ldg(a.tf + 6 * e + 2), t3 = __ldg(a.tf + 6 * e + 3), t5 = __ldg(a.tf + 6 * e + 5);
            double pa[3], pb[3], q0a[3], q0b[3];
            for (int j = 0; j < 3; ++j) {
                pa[j] = a.P[3 * ea + j]; pb[j] = a.P[3 * eb + j];
                q0a[j] = a.poses[3 * ea + j]; q0b[j] = a.poses[3 * eb + j];
            }
            SgdEdgeFull x;
            x.ea = ea; x.eb = eb; x.span = (double)(eb - ea);
            x.t2 = t2; x.t5 = t5;
            x.phi = atan2(t3, t0);
            for (int j = 0; j < 3; ++j) {
                const double tot = pb[j] - pa[j];
                x.base[j] = pa[j]; x.total[j] = tot; x.itot[j] = 1.0 / tot; x.Pb[j] = pb[j];
                x.p0a[j] = q0a[j]; x.p0b[j] = q0b[j];
            }
            for (int j = 0; j < 3; ++j) { R.pa[j * (size_t)a.E + e] = pa[j]; R.pb[j * (size_t)a.E + e] = pb[j]; }
            {
                const double *xd = reinterpret_cast<const double *>(&x);
#pragma unroll
                for (int f = 0; f < kSgdEdgeDoubles; ++f) ES[f * ES_stride + e] = xd[f];
            }
            R.ab[e] = make_int2(ea, eb);
            double *ra = a.RECA + 10 * (size_t)e;                // {coef[3] (written by the chain), pa[3], pb[3], (a, b)}
            for (int j = 0; j < 3; ++j) { ra[3 + j] = x.base[j]; ra[6 + j] = x.Pb[j]; }
            reinterpret_cast<int2 *>(ra)[9] = make_int2(ea, eb);
            if (e == 0)
                for (int j = 0; j < 3; ++j) { s_next[0][j] = x.p0a[j]; s_next[0][3 + j] = x.p0b[j]; }
        }
    }
    __syncthreads();
#ifdef ICPB_SGD_PROBE
    pq4 = clock64();
#endif
    // step factors that do not depend on the moving poses: alpha_j = lr / gamma_j (:39-40)
    double alpha[3];
    for (int j = 0; j < 3; ++j) alpha[j] = (1.0 / s_gamma[j]) * a.learning_rate;
    const double inv_lcu = 1.0 / a.lcu;

    if (a.E == 0) return;
    // The edge structs reach the scalar warp through a 4-slot ring in shared memory that the last warp
    // fills two edges ahead: lane k copies double k of the 184-byte struct, one coalesced load per edge.
    // The load is issued one iteration before its store (an L2 round trip is longer than an edge).
    __shared__ SgdEdgeFull s_es[4];
    double es_staged = 0.0;                                    // double `lane` of the struct of edge e + 2
    if (warp == kSgdWarps - 1 && lane < kSgdEdgeDoubles) {
        for (int k = 0; k < 2 && k < a.E; ++k)
            reinterpret_cast<double *>(&s_es[k])[lane] = ES[lane * ES_stride + k];
        if (2 < a.E) es_staged = ES[lane * ES_stride + 2];
    }
    __syncthreads();
    // Scalar warp, lane-parallel: lane l < 6 carries dof j = l % 3 of endpoint side = l / 3 (0: a, 1: b);
    // lanes 0..2 go on to the residual and step of dof j.  Lanes >= 6 shadow lane 5.
    const int l6 = min(lane, 5), side = l6 / 3, dof = l6 - 3 * side;
    const double alpha_l = dof == 0 ? alpha[0] : (dof == 1 ? alpha[1] : alpha[2]);
    double last_coef = 0.0;                                    // record e - 1 (the one the slots lack), dof of this lane
    int last_a = 0x7fffffff, last_b = 0;
    // slot owners (warps 1..15): thread `own` holds endpoint `own & 1` of edge `own >> 1` of a block
    const int own = tid - 32, s_off = own >> 1, s_side = own & 1;
    int cur_node = -1, nxt_node = -1;                          // -1: no slot (block past the last edge)
    double cur_P[3], cur_acc[3], nxt_P[3], nxt_acc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { cur_P[j] = cur_acc[j] = nxt_P[j] = nxt_acc[j] = 0.0; }
    auto slot_load = [&](int edge, int &node, double *P, double *acc) {
        node = -1;
        if (warp >= 1 && edge < a.E) {
            const double ab = ES[edge];                                        // field 0: (ea, eb) as two ints
            node = s_side ? __double2hiint(ab) : __double2loint(ab);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                P[j] = ES[((s_side ? kEsPb : kEsBase) + j) * ES_stride + edge];
                acc[j] = ES[((s_side ? kEsP0b : kEsP0a) + j) * ES_stride + edge];
            }
        }
    };
    slot_load(s_off, cur_node, cur_P, cur_acc);
    slot_load(kSgdBlock + s_off, nxt_node, nxt_P, nxt_acc);
    int blk_first = 0;                                         // first edge of the block the chain is in
#ifdef ICPB_SGD_PROBE
    const long long pr_begin = clock64();
#endif
    for (int e = 0; e < a.E; ++e) {
        if (warp == 0) {
            const SgdEdgeFull &ce = s_es[e & 3];
            const double *cur = reinterpret_cast<const double *>(&ce);
            const double *prev = reinterpret_cast<const double *>(&s_es[(e + 3) & 3]);   // edge e - 1, still in the ring
            // ---- the endpoints of edge e: the published running poses (records 0..e-2) + record e-1 ----
            const int ea = ce.ea, eb = ce.eb;
            const int node = side ? eb : ea;
            const double Pn = cur[(side ? kEsPb : kEsBase) + dof];
            double pose = s_next[e & 1][l6];
            {
                const double v = (node <= last_b ? Pn : prev[kEsPb + dof]) - prev[kEsBase + dof];
                const double q = __dadd_rn(pose, __dmul_rn(last_coef, v));
                pose = node > last_a ? q : pose;
            }
            // ---- residual (:33-35) ----
            const double pat = __shfl_sync(0xffffffffu, pose, 2);          // heading of a
            const double pb_v = __shfl_down_sync(0xffffffffu, pose, 3);    // lane j < 3: dof j of b
            double sn, cs;
            sincos(pat, &sn, &cs);
            const double two_pi = 6.283185307179586, inv_two_pi = 0.15915494309189535;   // 2 * np.pi
            const double u = dof == 0 ? cs : sn, v = dof == 0 ? -sn : cs;
            const double r_lin = (u * cur[kEsT2] + v * cur[kEsT5] + pose) - pb_v;   // translation of Pb_new minus poses[b]
            const double r_ang = mod_pos((pose + cur[kEsPhi]) - pb_v, two_pi, inv_two_pi);
            const double r = dof == 2 ? r_ang : r_lin;
            // ---- clipped step (:36-44) ----
            const double dd = 2.0 * (inv_lcu * r);                                   // :36
            double beta = cur[kEsSpan] * dd * alpha_l;                               // :42
            if (fabs(beta) > fabs(r)) beta = r;                                      // :43-44
            const double coef = beta * cur[kEsItot + dof];
            if (lane < 3) {
                s_coef[e & 1][lane] = coef;
                R.coef[lane * (size_t)a.E + e] = coef;                   // for sgd_apply_kernel
                a.RECA[10 * (size_t)e + lane] = coef;                    // for the catch-up bursts
            }
            last_coef = __shfl_sync(0xffffffffu, coef, dof);
            last_a = ea; last_b = eb;
        } else {
            if (warp == kSgdWarps - 1 && lane < kSgdEdgeDoubles && e + 2 < a.E) {
                // ring slot (e + 2) % 4 was last read in iteration e - 1 (as the previous edge of e - 1's successor)
                reinterpret_cast<double *>(&s_es[(e + 2) & 3])[lane] = es_staged;
                if (e + 3 < a.E) es_staged = ES[lane * ES_stride + e + 3];
            }
            if (e == blk_first + kSgdBlock) {
                // ---- the chain enters the next block: its slots move up, the slots of the block after it
                // are loaded and caught up on the records 0..e-2 (record e-1 follows below, as for all) ----
                blk_first += kSgdBlock;
                cur_node = nxt_node;
#pragma unroll
                for (int j = 0; j < 3; ++j) { cur_P[j] = nxt_P[j]; cur_acc[j] = nxt_acc[j]; }
                slot_load(blk_first + kSgdBlock + s_off, nxt_node, nxt_P, nxt_acc);
                // The records stream through shared memory in tiles of kSgdTile: thread `own` fetches 16-byte
                // piece `own` of the tile (one coalesced 7.5 KB read per tile, issued one tile ahead), then
                // everybody reads the tile as broadcast loads.  One named barrier per tile among the owners.
                const int n_rec = e - 1;
                const double2 *src = reinterpret_cast<const double2 *>(a.RECA);
                const int n_piece = 5 * n_rec;                               // 16-byte pieces in records 0..n_rec-1
                double2 stage = own < n_piece ? src[own] : make_double2(0.0, 0.0);
                for (int t0 = 0, buf = 0; t0 < n_rec; t0 += kSgdTile, buf ^= 1) {
                    s_tile[buf][own] = stage;
                    asm volatile("bar.sync 1, %0;" ::"n"(kSgdOwners) : "memory");
                    const int nxt_piece = 5 * (t0 + kSgdTile) + own;
                    if (nxt_piece < n_piece) stage = src[nxt_piece];
                    const int cnt = min(kSgdTile, n_rec - t0);
                    if (nxt_node >= 0) {
#pragma unroll 2
                        for (int r = 0; r < cnt; ++r) {
                            const double2 *rec = &s_tile[buf][5 * r];
                            const double2 v0 = rec[0], v1 = rec[1], v2 = rec[2], v3 = rec[3], v4 = rec[4];
                            const double rc[3] = {v0.x, v0.y, v1.x}, rpa[3] = {v1.y, v2.x, v2.y}, rpb[3] = {v3.x, v3.y, v4.x};
                            const int ra = __double2loint(v4.y), rb = __double2hiint(v4.y);
                            const bool on = nxt_node > ra, inside = nxt_node <= rb;
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const double q = __dadd_rn(nxt_acc[j], __dmul_rn(rc[j], (inside ? nxt_P[j] : rpb[j]) - rpa[j]));
                                nxt_acc[j] = on ? q : nxt_acc[j];
                            }
                        }
                    }
                }
            }
            if (e + 1 < a.E) {
                // ---- record e-1 into both slots; the owners of the endpoints of edge e+1 publish them ----
                if (e >= 1) {
                    const SgdEdgeFull &re = s_es[(e + 3) & 3];     // edge e - 1: still in the ring
                    const int ra = re.ea, rb = re.eb;
                    const bool on_c = cur_node > ra, in_c = cur_node <= rb, on_n = nxt_node > ra, in_n = nxt_node <= rb;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double rc = s_coef[(e - 1) & 1][j], rpa = re.base[j], rpb = re.Pb[j];
                        const double qc = __dadd_rn(cur_acc[j], __dmul_rn(rc, (in_c ? cur_P[j] : rpb) - rpa));
                        const double qn = __dadd_rn(nxt_acc[j], __dmul_rn(rc, (in_n ? nxt_P[j] : rpb) - rpa));
                        cur_acc[j] = on_c ? qc : cur_acc[j];
                        nxt_acc[j] = on_n ? qn : nxt_acc[j];
                    }
                }
                const bool in_next = e + 1 >= blk_first + kSgdBlock;
                if (s_off == e + 1 - blk_first - (in_next ? kSgdBlock : 0)) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) s_next[(e + 1) & 1][3 * s_side + j] = in_next ? nxt_acc[j] : cur_acc[j];
                }
            }
        }
        __syncthreads();           // record e, the endpoints of edge e+1 and the struct of edge e+2 are visible
    }
#ifdef ICPB_SGD_PROBE
    if (tid == 0) printf("sgd probe: loop %lld cycles, E %d; prologue: gamma %lld, P %lld, edge structs %lld, ring %lld\n", clock64() - pr_begin, a.E, pq1 - pq0, pq3 - pq1, pq4 - pq3, pr_begin - pq4);
#endif
}

// Every record applied to every node, in edge order per node: pose(i) += f_0(i), += f_1(i), ... --
// the reference's own sequence of additions for that node (:45-48).  One thread per node.
__global__ void __launch_bounds__(kSgdNodeThreads)
sgd_apply_kernel(const SgdArgs a)
{
    __shared__ int2 s_ab[kSgdNodeTile];
    __shared__ double s_c[kSgdNodeTile], s_pa[kSgdNodeTile], s_pb[kSgdNodeTile];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;                                  // dof: one (node, dof) per thread
    const SgdRecords R = sgd_records(a);
    const double *Rc = R.coef + j * (size_t)a.E, *Rpa = R.pa + j * (size_t)a.E, *Rpb = R.pb + j * (size_t)a.E;
    double p = 0.0, Pi = 0.0;
    if (i < a.n) { p = a.poses[3 * i + j]; Pi = a.P[3 * i + j]; }
    for (int e0 = 0; e0 < a.E; e0 += kSgdNodeTile) {            // the records stream through shared memory
        const int cnt = min(kSgdNodeTile, a.E - e0);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            s_ab[k] = R.ab[e0 + k];
            s_c[k] = Rc[e0 + k]; s_pa[k] = Rpa[e0 + k]; s_pb[k] = Rpb[e0 + k];
        }
        __syncthreads();
        // edge order = the reference's order of additions.  Branch-free and unrolled: the terms do not
        // depend on p, so only one addition and one select per record sit on the dependent chain
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const int2 ab = s_ab[k];
            const double v = (i <= ab.y ? Pi : s_pb[k]) - s_pa[k];
            const double q = __dadd_rn(p, __dmul_rn(s_c[k], v));
            p = i > ab.x ? q : p;                                // nodes up to a are not touched
        }
        __syncthreads();
    }
    if (i < a.n) a.poses[3 * i + j] = p;
}

}  // namespace icpb
