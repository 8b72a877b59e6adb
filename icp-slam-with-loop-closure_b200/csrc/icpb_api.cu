// C ABI of the batched ICP library (include/icpb.h).  Plain CUDA runtime: no torch types, no
// C++ exceptions across the boundary.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <time.h>
#include <sched.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "icpb.h"
#include "icpb_kernels.cuh"
#include "icpb_candidates.cuh"
#include "icpb_sgd.cuh"
#include "icpb_grid.cuh"
#include "icpb_compose.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(g_err, sizeof g_err, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return (int)e_;                                                               \
        }                                                                                 \
    } while (0)

// Every entry point runs on the handle's device and puts the caller's current device back on the
// way out (a torch process with several devices must not find its current device changed).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) {
            const cudaError_t e = cudaSetDevice(device);
            if (e != cudaSuccess) {
                snprintf(g_err, sizeof g_err, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
                ok = false;
            }
        } else {
            prev = -1;                                        // nothing to restore
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(h)                                      \
    DeviceGuard guard_((h)->device);                      \
    if (!guard_.ok) return (int)cudaErrorInvalidDevice

#ifndef ICPB_R
#define ICPB_R 2
#endif
constexpr int kPointsPerThread = ICPB_R;
constexpr int kQueueRing = 64;


struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return (int)e; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// host clock in microseconds (ICPB_TRACE timings of the host entry points)
static double now_us()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e)); return (int)e; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int kMaxSegments = 64;        // pieces of the streaming scan-table upload (icpb_align_host)

struct LaunchCfg {
    int threads, smem, ctas_per_sm;
    icpb::SmemLayout L;
    int cluster;   // CTAs per problem (1 = ordinary launch)
};

typedef void (*kernel_fn)(const icpb::KernelArgs);
// The exhaustive variant (every source point sweeps every target) keeps 4 points per thread in
// 256-thread CTAs: with no pruning there is nothing to balance, and the wider register tile
// amortises the shared-memory loads and the per-chunk bookkeeping over twice the distances.
constexpr int kPointsExhaustive = 4;
bool is_exhaustive(const icpb_params *p) { return p && (p->flags & ICPB_FLAG_EXHAUSTIVE); }
int points_per_thread(const icpb_params *p) { return is_exhaustive(p) ? kPointsExhaustive : kPointsPerThread; }
kernel_fn pick_kernel(const icpb_params *p)
{
    return is_exhaustive(p) ? icpb::icp_align_kernel<kPointsExhaustive, false, false>
                            : icpb::icp_align_kernel<kPointsPerThread, true, false>;
}
kernel_fn cluster_kernel() { return icpb::icp_align_kernel<kPointsPerThread, true, true>; }

}  // namespace

namespace {

// A few host threads that copy scans from the caller's (pageable, scattered) arrays into the pinned
// staging table, one job = a run of consecutive scans.  Jobs are handed out in table order, so the
// first upload piece is complete first; `done[k]` counts the finished jobs of piece k and the
// dispatching thread enqueues the piece's copy as soon as it is full.
struct PackJob {
    const double *const *scan_xy;       // n_scans pointers to (m_i, 2) float64 rows
    const int64_t *offsets;             // CSR offsets of the packed table
    double *dst;                        // pinned staging
    const int64_t *job_first, *job_last;  // scans [first, last) of job j
    const int32_t *job_piece;           // upload piece of job j
    int64_t n_jobs;
    std::atomic<int64_t> next{0};
    std::atomic<int32_t> *done;         // per piece
    std::atomic<uint64_t> bad{0};       // non-finite coordinate seen
};

// One scan into the pinned staging table with NON-TEMPORAL stores, checking for inf / NaN on the way.
// Ordinary stores leave the freshly written lines dirty in the writing core's private cache; the
// copy engine's reads then have to snoop them out of cores that have already gone idle, and the
// last pieces of the table -- the ones still cache resident when the packing ends -- crawl at
// ~8 GB/s instead of 55 (measured: the last two 5 MB pieces took 0.43 and 0.85 ms against 0.10 ms
// for the others).  Streaming stores go to memory through the write-combining buffers and skip
// the read-for-ownership as well.
static inline uint64_t copy_scan(double *dst, const double *src, int64_t n_doubles)
{
    uint64_t bad = 0;
#if defined(__x86_64__)
    const __m128i expo = _mm_set_epi32(0x7ff00000, 0, 0x7ff00000, 0);       // exponent bits of both doubles
    __m128i acc = _mm_setzero_si128();
    int64_t k = 0;
    for (; k + 8 <= n_doubles; k += 8) {                                    // one 64-byte line per trip
        const __m128d a = _mm_loadu_pd(src + k), b = _mm_loadu_pd(src + k + 2);
        const __m128d c = _mm_loadu_pd(src + k + 4), d = _mm_loadu_pd(src + k + 6);
        _mm_stream_pd(dst + k, a); _mm_stream_pd(dst + k + 2, b);
        _mm_stream_pd(dst + k + 4, c); _mm_stream_pd(dst + k + 6, d);
        // all-ones exponent = inf or NaN: compare the high words, ignore the low ones
        const __m128i ea = _mm_cmpeq_epi32(_mm_and_si128(_mm_castpd_si128(a), expo), expo);
        const __m128i eb = _mm_cmpeq_epi32(_mm_and_si128(_mm_castpd_si128(b), expo), expo);
        const __m128i ec = _mm_cmpeq_epi32(_mm_and_si128(_mm_castpd_si128(c), expo), expo);
        const __m128i ed = _mm_cmpeq_epi32(_mm_and_si128(_mm_castpd_si128(d), expo), expo);
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(ea, eb), _mm_or_si128(ec, ed)));
    }
    for (; k + 2 <= n_doubles; k += 2) {
        const __m128d a = _mm_loadu_pd(src + k);
        _mm_stream_pd(dst + k, a);
        acc = _mm_or_si128(acc, _mm_cmpeq_epi32(_mm_and_si128(_mm_castpd_si128(a), expo), expo));
    }
    // the low words compare equal trivially (0 == 0): keep the high words only
    const __m128i hi = _mm_and_si128(acc, _mm_set_epi32(-1, 0, -1, 0));
    bad = (uint64_t)(_mm_movemask_epi8(hi) != 0);
#else
    memcpy(dst, src, sizeof(double) * (size_t)n_doubles);
    const uint64_t *u = (const uint64_t *)dst;
    for (int64_t k = 0; k < n_doubles; ++k)
        bad |= (uint64_t)((u[k] & 0x7ff0000000000000ULL) == 0x7ff0000000000000ULL);
#endif
    return bad;
}

static void pack_run(PackJob *job)
{
    for (;;) {
        const int64_t j = job->next.fetch_add(1, std::memory_order_relaxed);
        if (j >= job->n_jobs) break;
        uint64_t bad = 0;
        for (int64_t s = job->job_first[j]; s < job->job_last[j]; ++s) {
            const int64_t o = job->offsets[s], m = job->offsets[s + 1] - o;
            bad |= copy_scan(job->dst + 2 * o, job->scan_xy[s], 2 * m);
        }
#if defined(__x86_64__)
        _mm_sfence();                                   // the streamed lines are in memory before the piece is announced
#endif
        if (bad) job->bad.fetch_or(1, std::memory_order_relaxed);
        job->done[job->job_piece[j]].fetch_add(1, std::memory_order_release);
    }
}

struct PackPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    PackJob *job = nullptr;
    uint64_t generation = 0;
    int running = 0;
    bool stop = false;

    explicit PackPool(int n)
    {
        for (int t = 0; t < n; ++t)
            workers.emplace_back([this] {
                uint64_t seen = 0;
                for (;;) {
                    PackJob *j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || generation != seen; });
                        if (stop) return;
                        seen = generation;
                        j = job;
                    }
                    pack_run(j);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        --running;
                    }
                    cv.notify_all();
                }
            });
    }
    void start(PackJob *j)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = j; ++generation; running = (int)workers.size();
        }
        cv.notify_all();
    }
    void finish()                                       // the caller packs too, then waits for the workers
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return running == 0; });
        job = nullptr;
    }
    ~PackPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &w : workers) w.join();
    }
};

}  // namespace

struct icpb_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_limit = 0;                             // dynamic shared memory an alignment launch may ask for
    // scan table
    const double *xy = nullptr;
    const int64_t *offsets = nullptr;
    int64_t n_scans = 0, longest = 0;
    DevBuf own_xy, own_off;
    // queue counters
    unsigned long long *queue = nullptr;      // kQueueRing counters, then one work counter
    unsigned long long *executed = nullptr;   // non-null while work counting is enabled
    int64_t launches = 0;
    // scratch for host-pointer entry points
    DevBuf s_pairs, s_init, s_T, s_err, s_passes, s_hist, s_corr, s_pair_xy, s_pair_off;
    cudaStream_t stream = nullptr;
    cudaStream_t cstream[2] = {nullptr, nullptr};   // compute streams of the pipelined host entry point
    cudaEvent_t seg_ev[16] = {};                    // "scan segment k has arrived"
    cudaEvent_t done_ev[2] = {};
    int32_t *arrived_dev = nullptr;                 // streaming upload: segments delivered so far
    // cuStreamWriteValue32 (driver API, looked up at run time): a stream-ordered 4-byte store, cheaper
    // than a 4-byte DMA for the "segment k has arrived" counter; nullptr -> the counter is copied
    int (*write32)(cudaStream_t, unsigned long long, unsigned int, unsigned int) = nullptr;
    int32_t *seg_vals_pinned = nullptr;             // 0..kMaxSegments in pinned host memory (sources of the flag copies)
    DevBuf s_seg;
    PinnedBuf stage;                                // pinned staging of the small per-call arrays
    DevBuf s_sgd;                                   // pose-graph SGD: poses, edges, transforms, scratch
    DevBuf s_grid;                                  // occupancy grid: poses, per-cell words, the grid
    DevBuf s_accept;                                // acceptance epilogue: [appended, CTAs left] counters
    // per-scan staging images of the resident table (KernelArgs::prep_in), built on first use
    DevBuf prep;
    const double *prep_for = nullptr;               // the table (h->xy) the images were built from
    int64_t prep_stride = 0, prep_scans = 0, prep_longest = 0;
    PinnedBuf stage_xy;                             // pinned copy of a scan table packed from a list of arrays
    PackPool *pool = nullptr;                       // host threads that pack scans into stage_xy
    // tuning / test hooks (icpb_set_tuning); 0 or -1 = the library's own choice
    int tune_threads = 0, tune_cluster = -1, tune_segments = 0, tune_pack_threads = 0;
    bool tune_flag_copy = false, tune_drop_counter = false, tune_trace = false;
};

namespace {

int make_cfg(icpb_ctx *h, int64_t longest, int64_t B, LaunchCfg *c, kernel_fn fn, int R = kPointsPerThread)
{
    if (longest <= 0) return fail(ICPB_EINVAL, "empty scan table%s");
    const int64_t ntile = (longest + 63) / 64;              // 64-point reduction tiles (independent of R)
    const int64_t nwork = (ntile + R / 2 - 1) / (R / 2);    // warp work items of 32*R points
    // Warps per CTA.  Small CTAs keep more independent problems in flight per SM (less idling at the
    // per-pass barrier); large CTAs finish a problem sooner, which matters when the batch is only a few
    // problems per resident CTA (tail) or a single pair (latency).  Within the cap, the count that leaves
    // the fewest warps idle in the last round of a pass wins (360-beam scans have 6 tiles: 3 warps take
    // two each, 12.2 M pairs/s, where 4 warps left two idle every second round, 11.5 M), larger on ties.
    int max_warps = (B >= 2048 && R < 4) ? 4 : 8;
    if (h->tune_threads >= 32 && h->tune_threads <= 256 && h->tune_threads % 32 == 0)
        max_warps = h->tune_threads / 32;                 // icpb_set_tuning("threads")
    int warps = 1;
    {
        int64_t best_waste = -1;
        for (int w = max_warps < nwork ? max_warps : (int)nwork; w >= (max_warps >= 2 && nwork >= 2 ? 2 : 1); --w) {
            const int64_t waste = (nwork + w - 1) / w * w - nwork;
            if (best_waste < 0 || waste < best_waste) { best_waste = waste; warps = w; }
        }
        if (h->tune_threads > 0) warps = max_warps < nwork ? max_warps : (int)nwork;   // forced: no search
    }
    int threads = warps * 32;
    const icpb::SmemLayout L = icpb::smem_layout(longest, threads / 32);
    const int64_t smem = L.bytes;
    if (smem > h->smem_limit) {
        snprintf(g_err, sizeof g_err, "scan of %lld points needs %lld B of shared memory (limit %d)",
                 (long long)longest, (long long)smem, h->smem_limit);
        return ICPB_ETOOLONG;
    }
    c->threads = threads; c->smem = (int)smem; c->L = L;
    // Latency mode: when every CTA of every cluster can have an SM of its own and a CTA's eight warps
    // would each get more than one tile per pass, spread each problem over a thread-block cluster (a
    // power of two, at most 8 CTAs) so that every tile has a warp.  Measured with one cluster per
    // problem resident (tools/latency.py, tools/midsize_probe.py, round 2): one 1,024-point pair 0.165 ms
    // in a CTA, 0.153 in a cluster of two, 0.141 in eight; 40 such pairs 0.37 -> 0.31 ms with clusters of
    // two, 80 pairs 0.41 -> 0.38; from 120 pairs on (more CTAs than SMs) single CTAs win; one 4,096-point
    // pair 0.78 -> 0.41 ms in a cluster of eight.
    c->cluster = 1;
    if (fn == pick_kernel(nullptr) && nwork > 8) {
        int cl = 2;
        while (cl < 8 && (int64_t)cl * 8 < nwork) cl *= 2;
        while (cl > 1 && B * cl > h->sm_count) cl /= 2;
        c->cluster = cl;
    }
    if (h->tune_cluster >= 0) {                           // icpb_set_tuning("cluster"): force a size (0 = never)
        const int v = h->tune_cluster;
        if (v == 0 || v == 1) c->cluster = 1;
        else if ((v == 2 || v == 4 || v == 8) && fn == pick_kernel(nullptr)) c->cluster = v;
    }
    // (the kernels' dynamic shared-memory limit is raised to the device maximum once, in icpb_create: the
    // attribute belongs to the function, not to a handle, so it must not follow one handle's history)
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) return fail(ICPB_EINVAL, "kernel does not fit on an SM%s");
    c->ctas_per_sm = per_sm;
    return 0;
}

int check_params(const icpb_params *p, int64_t B)
{
    if (!p) return fail(ICPB_EINVAL, "params is null%s");
    if (B < 0) return fail(ICPB_EINVAL, "negative batch size%s");
    if (p->hist_cap < 0 || p->corr_stride < 0) return fail(ICPB_EINVAL, "negative hist_cap/corr_stride%s");
    if (p->pair_mode != 0 && p->pair_mode != 1) return fail(ICPB_EINVAL, "pair_mode must be 0 or 1%s");
    if (p->pair_mode == 1 && p->k_block < 1) return fail(ICPB_EINVAL, "pair_mode 1 needs k_block >= 1%s");
    if (isnan(p->epsilon) || isnan(p->stopping_thresh)) return fail(ICPB_EINVAL, "NaN epsilon/stopping_thresh%s");
    return 0;
}

int check_epilogue(const icpb_epilogue *ep, const icpb_params *p)
{
    if (!ep) return 0;
    const bool gather = ep->d_peer_ptrs != nullptr;
    const bool accept = ep->d_accept_rec != nullptr || ep->d_accept_peer_ptrs != nullptr;
    if ((gather || ep->d_accept_peer_ptrs) && (ep->n_peers < 1 || ep->n_peers > 64))
        return fail(ICPB_EINVAL, "epilogue: 1..64 peer buffers expected%s");
    if (ep->row0 < 0 || ep->row_stride < 0) return fail(ICPB_EINVAL, "epilogue: negative row0 / row_stride%s");
    if (accept) {
        if (ep->accept_cap < 1) return fail(ICPB_EINVAL, "epilogue: accept_cap must be >= 1%s");
        if (isnan(ep->accept_thresh)) return fail(ICPB_EINVAL, "epilogue: NaN accept_thresh%s");
        if (ep->d_accept_peer_ptrs && (!ep->d_accept_count_peer_ptrs || ep->rank < 0 || ep->rank >= ep->n_peers))
            return fail(ICPB_EINVAL, "epilogue: peer acceptance needs the peers' count arrays and a rank in range%s");
        if (!ep->d_accept_peer_ptrs && !ep->d_accept_count) return fail(ICPB_EINVAL, "epilogue: d_accept_count is null%s");
    }
    (void)p;
    return 0;
}

int ensure_prep(icpb_ctx *h, cudaStream_t stream);

int launch(icpb_ctx *h, const double *xy, const int64_t *offsets, int64_t n_scans, int64_t longest,
           const int32_t *d_pairs, const double *d_init, int64_t B, const icpb_params *p,
           double *d_T, double *d_err, int32_t *d_passes, double *d_hist, int32_t *d_corr,
           cudaStream_t stream, int64_t B_total = 0, const int32_t *d_seg_of_pair = nullptr,
           const int32_t *d_arrived = nullptr, const icpb_epilogue *ep = nullptr,
           const int32_t *d_order = nullptr, int32_t *d_upload_timeout = nullptr,
           unsigned char *prep_out = nullptr, int64_t prep_stride_out = 0)
{
    const bool prep_building = prep_out != nullptr;
    if (B == 0) return 0;
    LaunchCfg cfg;
    kernel_fn fn = pick_kernel(p);
    int rc = make_cfg(h, longest, B_total > B ? B_total : B, &cfg, fn, points_per_thread(p));
    if (rc) return rc;
    icpb::KernelArgs a;
    a.xy = xy; a.offsets = offsets; a.pairs = d_pairs; a.init = d_init; a.B = B; a.n_scans = n_scans;
    a.p = *p;
    a.T_out = d_T; a.err_out = d_err; a.passes_out = d_passes;
    a.hist = p->hist_cap > 0 ? d_hist : nullptr;
    a.corr = p->corr_stride > 0 ? d_corr : nullptr;
    a.queue = h->queue + (h->launches % kQueueRing);
    a.n2pad_cap = cfg.L.n2pad; a.n1_cap = cfg.L.n1c; a.nchunk_cap = cfg.L.nchunk; a.ntile_cap = cfg.L.ntile;
    a.o_tqy = cfg.L.o_tqy; a.o_cb = cfg.L.o_cb; a.o_tc = cfg.L.o_tc; a.o_mm = cfg.L.o_mm; a.o_corr = cfg.L.o_corr;
    a.o_red = cfg.L.o_red; a.o_tw = cfg.L.o_tw; a.o_s0 = cfg.L.o_s0; a.o_scr = cfg.L.o_scr;
    a.executed = h->executed;
    a.seg_of_pair = d_seg_of_pair; a.arrived = d_arrived; a.order = d_order; a.upload_timeout = d_upload_timeout;
    a.prep_out = nullptr; a.prep_in = nullptr; a.prep_stride = 0;
    if (d_arrived == nullptr && xy == h->xy && h->xy != nullptr && !prep_building) {
        // a resident table: stage every pair by copy from the per-scan images (built once per table)
        int rc2 = ensure_prep(h, stream);
        if (rc2) return rc2;
        if (h->prep_for == h->xy) { a.prep_in = (const unsigned char *)h->prep.p; a.prep_stride = h->prep_stride; }
    }
    if (prep_out) { a.prep_out = prep_out; a.prep_stride = prep_stride_out; }
    a.peers = nullptr; a.n_peers = 0; a.rec_row0 = 0; a.rec_block = B > 0 ? B : 1; a.rec_stride = 0;
    a.accept_thresh = 0.0; a.accept_rec = nullptr; a.accept_peers = nullptr; a.accept_ctr = nullptr;
    a.accept_count_out = nullptr; a.accept_count_peers = nullptr; a.accept_cap = 0; a.accept_rank = 0;
    if (ep) {
        a.peers = (double *const *)ep->d_peer_ptrs; a.n_peers = ep->n_peers;
        a.rec_row0 = ep->row0; a.rec_stride = ep->row_stride;
        if (ep->row_block > 0) a.rec_block = ep->row_block;
        if (ep->d_accept_rec || ep->d_accept_peer_ptrs) {
            int rc2;
            if ((rc2 = h->s_accept.reserve(2 * sizeof(unsigned long long)))) return rc2;
            a.accept_thresh = ep->accept_thresh; a.accept_cap = ep->accept_cap; a.accept_rank = ep->rank;
            a.accept_rec = ep->d_accept_peer_ptrs ? nullptr : ep->d_accept_rec;
            a.accept_peers = (double *const *)ep->d_accept_peer_ptrs;
            a.accept_count_peers = (long long *const *)ep->d_accept_count_peer_ptrs;
            a.accept_count_out = (long long *)ep->d_accept_count;
            a.accept_ctr = (unsigned long long *)h->s_accept.p;
            CU(cudaMemsetAsync(a.accept_ctr, 0, 2 * sizeof(unsigned long long), stream));
        }
    }
    CU(cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), stream));
    int64_t grid = (int64_t)cfg.ctas_per_sm * h->sm_count;
    if (grid > B) grid = B;
    if (cfg.cluster > 1) {
        cudaLaunchConfig_t lc = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cfg.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        lc.blockDim = dim3((unsigned)cfg.threads);
        lc.dynamicSmemBytes = (size_t)cfg.smem; lc.stream = stream; lc.attrs = attr; lc.numAttrs = 1;
        // one cluster per problem, as many as the device can co-schedule (the occupancy query needs a
        // grid that is a multiple of the cluster size; it answers in clusters)
        lc.gridDim = dim3((unsigned)cfg.cluster);
        int max_clusters = 0;
        CU(cudaOccupancyMaxActiveClusters(&max_clusters, cluster_kernel(), &lc));
        int64_t clusters = max_clusters > 0 ? max_clusters : 1;
        if (clusters > B) clusters = B;
        lc.gridDim = dim3((unsigned)(clusters * cfg.cluster));
        CU(cudaLaunchKernelEx(&lc, cluster_kernel(), a));
    } else {
        fn<<<(unsigned)grid, cfg.threads, cfg.smem, stream>>>(a);
    }
    CU(cudaGetLastError());
    h->launches++;
    return 0;
}

// The device form of a resident scan table: besides the fp64 points, one staging image per scan (fp32
// target arrays, chunk circles, group circles, S0, largest coordinate), built by one launch of the
// alignment kernel in prep mode the first time the table is used.  A pair's staging is then a 10 KB copy
// instead of ~11,000 warp instructions; scans that take part in many pairs (loop-closure candidates, all
// pairs) are prepared once instead of once per pair.  Tables whose images would exceed 8 GB are staged
// per pair as before.
int ensure_prep(icpb_ctx *h, cudaStream_t stream)
{
    if (h->prep_for == h->xy && h->prep_scans == h->n_scans && h->prep_longest == h->longest) return 0;
    h->prep_for = nullptr;
    const icpb::SmemLayout L = icpb::smem_layout(h->longest, 4);
    const int64_t stride = ((int64_t)L.o_mm + 32 + 15) & ~int64_t(15);
    if (stride * h->n_scans > (int64_t(8) << 30)) return 0;
    int rc;
    if ((rc = h->prep.reserve((size_t)(stride * h->n_scans)))) return rc;
    icpb_params p;
    icpb_default_params(&p);
    p.k_block = h->n_scans;
    rc = launch(h, h->xy, h->offsets, h->n_scans, h->longest, nullptr, nullptr, h->n_scans, &p,
                nullptr, nullptr, nullptr, nullptr, nullptr, stream, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                (unsigned char *)h->prep.p, stride);
    if (rc) return rc;
    h->prep_for = h->xy; h->prep_stride = stride; h->prep_scans = h->n_scans; h->prep_longest = h->longest;
    return 0;
}

}  // namespace

extern "C" {

void icpb_default_params(icpb_params *p)
{
    memset(p, 0, sizeof *p);
    p->epsilon = 0.01;
    p->stopping_thresh = 0.0001;
    p->max_iters = 100;
    p->k_block = 1;
}

int icpb_abi_version(void) { return ICPB_ABI_VERSION; }

const char *icpb_last_error(void) { return g_err; }

int icpb_create(int device, icpb_handle *out)
{
    if (!out) return fail(ICPB_EINVAL, "out is null%s");
    *out = nullptr;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(ICPB_EINVAL, "no such CUDA device%s");
    DeviceGuard guard_(device);
    if (!guard_.ok) return (int)cudaErrorInvalidDevice;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        snprintf(g_err, sizeof g_err, "device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
        return ICPB_EINVAL;
    }
    icpb_ctx *h = new (std::nothrow) icpb_ctx();
    if (!h) return fail(ICPB_EINVAL, "out of host memory%s");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaMalloc(&h->queue, sizeof(unsigned long long) * (kQueueRing + 1));
    if (e == cudaSuccess) e = cudaMemset(h->queue, 0, sizeof(unsigned long long) * (kQueueRing + 1));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(&h->cstream[k], cudaStreamNonBlocking);
    for (int k = 0; k < 16 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&h->seg_ev[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&h->done_ev[k], cudaEventDisableTiming);
    // dynamic shared memory up to the device's opt-in maximum minus each kernel's static part, set
    // once: the attribute belongs to the function, not to a handle
    h->smem_limit = (int)prop.sharedMemPerBlockOptin;
    for (kernel_fn fn : {pick_kernel(nullptr), cluster_kernel(),
                         (kernel_fn)icpb::icp_align_kernel<kPointsExhaustive, false, false>}) {
        cudaFuncAttributes fa;
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) break;
        const int lim = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        if (lim < h->smem_limit) h->smem_limit = lim;
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->arrived_dev, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaHostAlloc(&h->seg_vals_pinned, sizeof(int32_t) * (kMaxSegments + 1), cudaHostAllocDefault);
    if (e == cudaSuccess) for (int k = 0; k <= kMaxSegments; ++k) h->seg_vals_pinned[k] = k;
    if (e == cudaSuccess) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fp, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            h->write32 = (int (*)(cudaStream_t, unsigned long long, unsigned int, unsigned int))fp;
        else
            cudaGetLastError();
    }
    if (e != cudaSuccess) {
        icpb_destroy(h);                                      // frees whatever was created so far
        snprintf(g_err, sizeof g_err, "icpb_create: %s", cudaGetErrorString(e));
        return (int)e;
    }
    *out = h;
    return 0;
}

int icpb_set_tuning(icpb_handle h, const char *key, int64_t value)
{
    if (!h || !key) return fail(ICPB_EINVAL, "icpb_set_tuning: bad argument%s");
    const int v = (int)value;
    if (!strcmp(key, "threads")) h->tune_threads = v;
    else if (!strcmp(key, "cluster")) h->tune_cluster = v;
    else if (!strcmp(key, "segments")) h->tune_segments = v;
    else if (!strcmp(key, "pack_threads")) {
        if (h->pool && (int)h->pool->workers.size() != v - 1) { delete h->pool; h->pool = nullptr; }
        h->tune_pack_threads = v;
    }
    else if (!strcmp(key, "flag_copy")) h->tune_flag_copy = v != 0;
    else if (!strcmp(key, "drop_counter")) h->tune_drop_counter = v != 0;
    else if (!strcmp(key, "trace")) h->tune_trace = v != 0;
    else return fail(ICPB_EINVAL, "icpb_set_tuning: unknown key %s", key);
    return 0;
}

int icpb_destroy(icpb_handle h)
{
    if (!h) return 0;
    DeviceGuard guard_(h->device);
    delete h->pool;
    h->pool = nullptr;
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (int k = 0; k < 2; ++k) if (h->cstream[k]) { cudaStreamSynchronize(h->cstream[k]); cudaStreamDestroy(h->cstream[k]); }
    for (int k = 0; k < 16; ++k) if (h->seg_ev[k]) cudaEventDestroy(h->seg_ev[k]);
    for (int k = 0; k < 2; ++k) if (h->done_ev[k]) cudaEventDestroy(h->done_ev[k]);
    if (h->arrived_dev) cudaFree(h->arrived_dev);
    if (h->seg_vals_pinned) cudaFreeHost(h->seg_vals_pinned);
    h->s_seg.release(); h->stage.release(); h->s_sgd.release(); h->s_grid.release();
    h->s_accept.release(); h->stage_xy.release(); h->prep.release();
    h->own_xy.release(); h->own_off.release();
    h->s_pairs.release(); h->s_init.release(); h->s_T.release(); h->s_err.release();
    h->s_passes.release(); h->s_hist.release(); h->s_corr.release();
    h->s_pair_xy.release(); h->s_pair_off.release();
    if (h->queue) cudaFree(h->queue);
    delete h;
    return 0;
}

static int validate_offsets(const int64_t *off, int64_t n_scans, int64_t *longest)
{
    if (off[0] != 0) return fail(ICPB_EINVAL, "offsets[0] must be 0%s");
    int64_t L = 0;
    for (int64_t s = 0; s < n_scans; ++s) {
        const int64_t m = off[s + 1] - off[s];
        if (m <= 0) return fail(ICPB_EINVAL, "empty scan in the table (the reference's argmin raises on it)%s");
        if (m > L) L = m;
    }
    *longest = L;
    return 0;
}

int icpb_upload_scans(icpb_handle h, const double *h_xy, const int64_t *h_offsets, int64_t n_scans)
{
    if (!h || !h_xy || !h_offsets || n_scans <= 0) return fail(ICPB_EINVAL, "icpb_upload_scans: bad argument%s");
    int64_t longest = 0;
    int rc = validate_offsets(h_offsets, n_scans, &longest);
    if (rc) return rc;
    ON_DEVICE(h);
    const size_t nb_xy = sizeof(double) * 2 * (size_t)h_offsets[n_scans];
    const size_t nb_off = sizeof(int64_t) * (size_t)(n_scans + 1);
    if ((rc = h->own_xy.reserve(nb_xy))) return rc;
    if ((rc = h->own_off.reserve(nb_off))) return rc;
    CU(cudaMemcpyAsync(h->own_xy.p, h_xy, nb_xy, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->own_off.p, h_offsets, nb_off, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->prep_for = nullptr; h->xy = (const double *)h->own_xy.p;
    h->offsets = (const int64_t *)h->own_off.p;
    h->n_scans = n_scans;
    h->longest = longest;
    return 0;
}

int icpb_set_scans_device(icpb_handle h, const double *d_xy, const int64_t *d_offsets,
                          int64_t n_scans, int64_t longest_scan)
{
    if (!h || !d_xy || !d_offsets || n_scans <= 0 || longest_scan <= 0)
        return fail(ICPB_EINVAL, "icpb_set_scans_device: bad argument%s");
    h->prep_for = nullptr; h->xy = d_xy; h->offsets = d_offsets; h->n_scans = n_scans; h->longest = longest_scan;
    return 0;
}

int icpb_run_device(icpb_handle h, const int32_t *d_pairs, const double *d_init, int64_t B,
                    const icpb_params *p, double *d_T, double *d_err, int32_t *d_passes,
                    double *d_hist, int32_t *d_corr, void *stream)
{
    if (!h) return fail(ICPB_EINVAL, "handle is null%s");
    int rc = check_params(p, B);
    if (rc) return rc;
    if (!h->xy) return fail(ICPB_ENOSCANS, "no scan table set%s");
    if (B > 0 && (!d_T || !d_err || !d_passes)) return fail(ICPB_EINVAL, "output pointer is null%s");
    if (B > 0 && p->pair_mode == 0 && !d_pairs) return fail(ICPB_EINVAL, "pairs is null with pair_mode 0%s");
    if (p->hist_cap > 0 && !d_hist) return fail(ICPB_EINVAL, "hist_cap > 0 but hist is null%s");
    if (p->corr_stride > 0 && !d_corr) return fail(ICPB_EINVAL, "corr_stride > 0 but corr is null%s");
    if (p->corr_stride > 0 && p->corr_stride < h->longest)
        return fail(ICPB_EINVAL, "corr_stride is smaller than the longest scan%s");
    ON_DEVICE(h);
    return launch(h, h->xy, h->offsets, h->n_scans, h->longest, d_pairs, d_init, B, p,
                  d_T, d_err, d_passes, d_hist, d_corr, (cudaStream_t)stream);
}

int icpb_run_device_ex(icpb_handle h, const int32_t *d_pairs, const double *d_init, int64_t B,
                       const icpb_params *p, double *d_T, double *d_err, int32_t *d_passes,
                       const icpb_epilogue *ep, void *stream)
{
    if (!h) return fail(ICPB_EINVAL, "handle is null%s");
    int rc = check_params(p, B);
    if (rc) return rc;
    if ((rc = check_epilogue(ep, p))) return rc;
    if (!h->xy) return fail(ICPB_ENOSCANS, "no scan table set%s");
    if (B > 0 && (!d_T || !d_err || !d_passes)) return fail(ICPB_EINVAL, "output pointer is null%s");
    if (B > 0 && p->pair_mode == 0 && !d_pairs) return fail(ICPB_EINVAL, "pairs is null with pair_mode 0%s");
    if (p->hist_cap > 0 || p->corr_stride > 0) return fail(ICPB_EINVAL, "icpb_run_device_ex: no history/correspondences%s");
    ON_DEVICE(h);
    return launch(h, h->xy, h->offsets, h->n_scans, h->longest, d_pairs, d_init, B, p, d_T, d_err, d_passes,
                  nullptr, nullptr, (cudaStream_t)stream, 0, nullptr, nullptr, ep);
}

int icpb_run_device_gather(icpb_handle h, const int32_t *d_pairs, const double *d_init, int64_t B,
                           const icpb_params *p, double *d_T, double *d_err, int32_t *d_passes,
                           const uint64_t *d_peer_ptrs, int32_t n_peers, int64_t row0, void *stream)
{
    if (!d_peer_ptrs) return fail(ICPB_EINVAL, "icpb_run_device_gather: peer buffers expected%s");
    icpb_epilogue ep;
    memset(&ep, 0, sizeof ep);
    ep.d_peer_ptrs = d_peer_ptrs; ep.n_peers = n_peers; ep.row0 = row0;
    return icpb_run_device_ex(h, d_pairs, d_init, B, p, d_T, d_err, d_passes, &ep, stream);
}

static int run_host_common(icpb_handle h, const double *xy, const int64_t *offsets, int64_t n_scans,
                           int64_t longest, const int32_t *h_pairs, const double *h_init, int64_t B,
                           const icpb_params *p, double *h_T, double *h_err, int32_t *h_passes,
                           double *h_hist, int32_t *h_corr)
{
    int rc;
    cudaStream_t st = h->stream;
    const size_t nbP = sizeof(int32_t) * 2 * (size_t)B, nbI = sizeof(double) * 6 * (size_t)B;
    const size_t nbH = sizeof(double) * 6 * (size_t)B * (size_t)p->hist_cap;
    const size_t nbC = sizeof(int32_t) * (size_t)B * (size_t)p->corr_stride;
    if (p->pair_mode == 0) {
        if ((rc = h->s_pairs.reserve(nbP))) return rc;
        CU(cudaMemcpyAsync(h->s_pairs.p, h_pairs, nbP, cudaMemcpyHostToDevice, st));
    }
    if (h_init) {
        if ((rc = h->s_init.reserve(nbI))) return rc;
        CU(cudaMemcpyAsync(h->s_init.p, h_init, nbI, cudaMemcpyHostToDevice, st));
    }
    if ((rc = h->s_T.reserve(nbI))) return rc;
    if ((rc = h->s_err.reserve(sizeof(double) * (size_t)B))) return rc;
    if ((rc = h->s_passes.reserve(sizeof(int32_t) * (size_t)B))) return rc;
    if (nbH) { if ((rc = h->s_hist.reserve(nbH))) return rc; CU(cudaMemsetAsync(h->s_hist.p, 0, nbH, st)); }
    if (nbC) { if ((rc = h->s_corr.reserve(nbC))) return rc; CU(cudaMemsetAsync(h->s_corr.p, 0xff, nbC, st)); }
    rc = launch(h, xy, offsets, n_scans, longest,
                p->pair_mode == 0 ? (const int32_t *)h->s_pairs.p : nullptr,
                h_init ? (const double *)h->s_init.p : nullptr, B, p,
                (double *)h->s_T.p, (double *)h->s_err.p, (int32_t *)h->s_passes.p,
                (double *)h->s_hist.p, (int32_t *)h->s_corr.p, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_T, h->s_T.p, nbI, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_err, h->s_err.p, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_passes, h->s_passes.p, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, st));
    if (nbH) CU(cudaMemcpyAsync(h_hist, h->s_hist.p, nbH, cudaMemcpyDeviceToHost, st));
    if (nbC) CU(cudaMemcpyAsync(h_corr, h->s_corr.p, nbC, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int icpb_run_host(icpb_handle h, const int32_t *h_pairs, const double *h_init, int64_t B,
                  const icpb_params *p, double *h_T, double *h_err, int32_t *h_passes,
                  double *h_hist, int32_t *h_corr)
{
    if (!h) return fail(ICPB_EINVAL, "handle is null%s");
    int rc = check_params(p, B);
    if (rc) return rc;
    if (!h->xy) return fail(ICPB_ENOSCANS, "no scan table set%s");
    if (B == 0) return 0;
    if (!h_T || !h_err || !h_passes) return fail(ICPB_EINVAL, "output pointer is null%s");
    if (p->pair_mode == 0 && !h_pairs) return fail(ICPB_EINVAL, "pairs is null with pair_mode 0%s");
    if (p->hist_cap > 0 && !h_hist) return fail(ICPB_EINVAL, "hist_cap > 0 but hist is null%s");
    if (p->corr_stride > 0 && !h_corr) return fail(ICPB_EINVAL, "corr_stride > 0 but corr is null%s");
    if (p->corr_stride > 0 && p->corr_stride < h->longest)
        return fail(ICPB_EINVAL, "corr_stride is smaller than the longest scan%s");
    if (p->pair_mode == 0) {
        for (int64_t b = 0; b < 2 * B; ++b)
            if (h_pairs[b] < 0 || h_pairs[b] >= h->n_scans) return fail(ICPB_EINVAL, "pair index out of range%s");
    }
    ON_DEVICE(h);
    return run_host_common(h, h->xy, h->offsets, h->n_scans, h->longest, h_pairs, h_init, B, p,
                           h_T, h_err, h_passes, h_hist, h_corr);
}

/* Upload + align with the upload hidden behind the kernels: the scan table goes up in segments on
 * the copy stream; pairs are grouped by the segment that completes them (the larger of their two
 * scan ids) and every group is launched as soon as its segment has arrived. */
/* Pieces of the streaming scan-table upload (pure host logic, no CUDA call).
 * Equal pieces of about 4 MB, at most 16 by default (the first pairs start ~0.1 ms into the upload;
 * the kernel is launched before any piece has arrived and its CTAs wait on the counter).  Measured on
 * the 81 MB chain table with counter copies: 4 pieces 2.49 ms, 16 2.35 ms, 32 2.55 ms, 64 2.86 ms per
 * call; with cuStreamWriteValue32 16 and 32 pieces cost the same.
 * A piece must end on a 32-byte sector boundary (an even point offset): the kernel reads scans
 * through L1, and a sector that straddled two pieces could be cached while its second half had not
 * arrived yet.  Boundaries on a 128-byte line (offset % 8 == 0) are preferred when one is near; a
 * wanted boundary with no even offset within 64 scans is dropped (its piece merges with the next). */
int icpb_plan_upload(const int64_t *h_offsets, int64_t n_scans, int32_t pieces_wanted,
                     int64_t *piece_end, int32_t *n_pieces)
{
    if (!h_offsets || n_scans <= 0 || !piece_end || !n_pieces)
        return fail(ICPB_EINVAL, "icpb_plan_upload: bad argument%s");
    const int64_t total = h_offsets[n_scans];
    const size_t nb_xy = sizeof(double) * 2 * (size_t)total;
    int nseg = (int)(nb_xy / (4u << 20)) + 1;
    if (nseg > 16) nseg = 16;
    // Default plan: the first two and the last two pieces are a quarter and a half of a standard
    // piece, so the first bytes are on the wire (and the first pairs running) sooner and fewer pairs
    // are left waiting for the last piece.  An explicit piece count gives equal pieces.
    const bool graded = !(pieces_wanted >= 1 && pieces_wanted <= kMaxSegments) && nseg >= 8;
    if (graded) nseg += 2;
    if (pieces_wanted >= 1 && pieces_wanted <= kMaxSegments) nseg = pieces_wanted;
    if (nseg > n_scans) nseg = (int)n_scans;
    double cum[kMaxSegments + 1];
    cum[0] = 0.0;
    for (int k = 0; k < nseg; ++k) {
        double w = 1.0;
        if (graded && (k == 0 || k == nseg - 1)) w = 0.25;
        if (graded && (k == 1 || k == nseg - 2)) w = 0.5;
        cum[k + 1] = cum[k] + w;
    }
    int64_t s = 0;
    int made = 0;
    for (int k = 0; k < nseg - 1; ++k) {
        const int64_t want = (int64_t)((double)total * (cum[k + 1] / cum[nseg]));
        while (s < n_scans && h_offsets[s] < want) ++s;
        int64_t pick = -1;
        for (int64_t c = s; c < n_scans && c < s + 64; ++c) {
            if (h_offsets[c] % 8 == 0) { pick = c; break; }
            if (pick < 0 && h_offsets[c] % 2 == 0) pick = c;
        }
        if (pick <= 0 || pick >= n_scans || (made > 0 && pick <= piece_end[made - 1])) continue;   // merge with the next piece
        piece_end[made++] = pick;
        s = pick;
    }
    piece_end[made++] = n_scans;
    *n_pieces = made;
    return 0;
}

namespace {

// Where icpb_align_host*'s scan table comes from: one packed (sum m_i, 2) array with its offsets, or the
// reference's own `lidar_points` form -- a list of separate (m_i, 2) arrays in pageable memory.
struct TableSrc {
    const double *xy = nullptr;                 // packed form
    const int64_t *offsets = nullptr;           // n_scans + 1 (computed from the lengths for the list form)
    const double *const *scan_xy = nullptr;     // list form
};

// Everything that was enqueued is drained before an error is reported, and the handle is left without
// a table: its buffer may hold a mixture of the old and the new scans.
int align_abort(icpb_ctx *h, int rc)
{
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep);
    cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->cstream[0]); cudaStreamSynchronize(h->cstream[1]);
    cudaGetLastError();
    memcpy(g_err, keep, sizeof keep);
    h->prep_for = nullptr; h->xy = nullptr; h->offsets = nullptr; h->n_scans = 0; h->longest = 0;
    return rc;
}
#define CUA(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(g_err, sizeof g_err, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return align_abort(h, (int)e_);                                               \
        }                                                                                 \
    } while (0)

int host_threads_available()
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned n = std::thread::hardware_concurrency();
    return n > 0 ? (int)n : 1;
}

// icpb_align_host_accept: only the pairs that pass the acceptance test come back to the host
struct AcceptHost {
    double thresh;
    int64_t cap;
    int64_t *n_out, *rows;
    double *T;
    int32_t T_ld;
    double *err;
    int32_t *passes;
};

int align_impl(icpb_handle h, const TableSrc &src, int64_t n_scans, int64_t longest,
               const int32_t *h_pairs, const double *h_init, int32_t init_ld, int64_t B,
               const icpb_params *p, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes,
               const icpb_epilogue *ep, const AcceptHost *acc = nullptr)
{
    const bool trace = h->tune_trace;
    const double t_entry = trace ? now_us() : 0.0;
    const int64_t *h_offsets = src.offsets;
    int rc;
    ON_DEVICE(h);
    const size_t nb_xy = sizeof(double) * 2 * (size_t)h_offsets[n_scans];
    const size_t nb_off = sizeof(int64_t) * (size_t)(n_scans + 1);
    if ((rc = h->own_xy.reserve(nb_xy))) return rc;
    if ((rc = h->own_off.reserve(nb_off))) return rc;
    if (src.scan_xy && (rc = h->stage_xy.reserve(nb_xy))) return rc;
    int nseg = 0;
    int64_t seg_end[kMaxSegments];                            // exclusive scan id
    if ((rc = icpb_plan_upload(h_offsets, n_scans, B == 0 ? 1 : h->tune_segments, seg_end, &nseg))) return rc;
    // pinned staging, 8-byte members first:  up = [init | pairs | seg | order],  down = [T | err | passes]
    const size_t nbI = sizeof(double) * 6 * (size_t)B, nbE = sizeof(double) * (size_t)B, nb4 = sizeof(int32_t) * (size_t)B;
    const size_t nb_up = nbI + 4 * nb4, nb_down = nbI + nbE + nb4 + sizeof(int32_t);   // + the "upload timed out" word
    if ((rc = h->stage.reserve(nb_up + nb_down + 32))) return rc;
    if ((rc = h->s_init.reserve(nb_up))) return rc;
    if ((rc = h->s_T.reserve(nb_down))) return rc;
    icpb_epilogue ep_acc;
    if (acc) {
        // acceptance on the device: compact records + their count in device scratch, fetched at the end
        if ((rc = h->s_hist.reserve(sizeof(double) * 8 * (size_t)acc->cap + 16))) return rc;
        if (ep) ep_acc = *ep; else memset(&ep_acc, 0, sizeof ep_acc);
        ep_acc.accept_thresh = acc->thresh; ep_acc.accept_cap = acc->cap;
        ep_acc.d_accept_count = (int64_t *)h->s_hist.p;
        ep_acc.d_accept_rec = (double *)h->s_hist.p + 2;
        ep_acc.d_accept_peer_ptrs = nullptr; ep_acc.d_accept_count_peer_ptrs = nullptr;
        ep = &ep_acc;
    }
    double *pinit = (double *)h->stage.p;
    int32_t *ppairs = (int32_t *)(pinit + 6 * B), *pseg = ppairs + 2 * B, *porder = pseg + B;
    double *tT = (double *)((char *)h->stage.p + nb_up), *tE = tT + 6 * B;
    int32_t *tP = (int32_t *)(tE + B);
    double *d_init = (double *)h->s_init.p;
    int32_t *d_pairs = (int32_t *)(d_init + 6 * B), *d_seg = d_pairs + 2 * B, *d_order = d_seg + B;
    double *d_T = (double *)h->s_T.p, *d_err = d_T + 6 * B;
    int32_t *d_passes = (int32_t *)(d_err + B);
    const double *d_xy = (const double *)h->own_xy.p;
    const int64_t *d_off = (const int64_t *)h->own_off.p;

    // ---- list form: host threads pack the scans into pinned staging, piece by piece ----
    PackJob job;
    std::vector<int64_t> job_first, job_last;
    std::vector<int32_t> job_piece, jobs_of_piece;
    std::vector<std::atomic<int32_t>> piece_done(src.scan_xy ? (size_t)nseg : 0);
    const double *up_xy = src.xy;
    if (src.scan_xy) {
        int nthr = h->tune_pack_threads > 0 ? h->tune_pack_threads : host_threads_available();
        if (nthr > 32) nthr = 32;
        if (h->tune_pack_threads <= 0 && nthr > 8) nthr = 8;      // default: at most 8
        if (nthr < 1) nthr = 1;
        if (!h->pool && nthr > 1) h->pool = new (std::nothrow) PackPool(nthr - 1);
        jobs_of_piece.assign((size_t)nseg, 0);
        for (int k = 0, s0 = 0; k < nseg; ++k) {              // every piece split into nthr runs of scans
            const int64_t s1 = seg_end[k], n = s1 - s0;
            const int parts = (int)(n < nthr ? n : nthr);
            for (int q = 0; q < parts; ++q) {
                job_first.push_back(s0 + n * q / parts); job_last.push_back(s0 + n * (q + 1) / parts);
                job_piece.push_back(k);
            }
            jobs_of_piece[(size_t)k] = parts;
            piece_done[(size_t)k].store(0, std::memory_order_relaxed);
            s0 = (int)s1;
        }
        job.scan_xy = src.scan_xy; job.offsets = h_offsets; job.dst = (double *)h->stage_xy.p;
        job.job_first = job_first.data(); job.job_last = job_last.data(); job.job_piece = job_piece.data();
        job.n_jobs = (int64_t)job_first.size(); job.done = piece_done.data();
        if (h->pool) h->pool->start(&job);
        up_xy = (const double *)h->stage_xy.p;
    }
    // once the pool runs, every exit path must stop it first (the job lives on this stack frame)
    auto pack_finish = [&]() {
        if (!src.scan_xy) return;
        job.next.store(job.n_jobs, std::memory_order_relaxed);   // hand out no more jobs
        if (h->pool) h->pool->finish();
    };
#define CUP(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(g_err, sizeof g_err, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            pack_finish();                                                                \
            return align_abort(h, (int)e_);                                               \
        }                                                                                 \
    } while (0)

    // The packers (list input) are already at work on the pinned staging; the caller's pairs and initial
    // guesses are checked now, before any state of the handle changes: a rejected call leaves the old
    // table usable.
    {
        int32_t lo = 0, hi = 0;                               // branch-free range check
        for (int64_t b = 0; b < 2 * B; ++b) { lo = h_pairs[b] < lo ? h_pairs[b] : lo; hi = h_pairs[b] > hi ? h_pairs[b] : hi; }
        if (lo < 0 || hi >= n_scans) { pack_finish(); return fail(ICPB_EINVAL, "pair index out of range%s"); }
    }
    if (h_init && B > 0) {                                    // before any state of the handle changes
        for (int64_t b = 0; b < B && init_ld == 9; ++b) {
            const double *m = h_init + 9 * b;
            if (!(m[6] == 0.0 && m[7] == 0.0 && m[8] == 1.0))
            { pack_finish(); return fail(ICPB_EINVAL, "transform bottom row must be [0, 0, 1] (an SE(2) matrix)%s"); }
        }
        uint64_t bad = 0;                                     // all-ones exponent = inf or NaN; integer test, vectorises
        for (int64_t k = 0; k < (int64_t)init_ld * B; ++k) {
            uint64_t u;
            memcpy(&u, h_init + k, sizeof u);
            bad |= (uint64_t)((u & 0x7ff0000000000000ULL) == 0x7ff0000000000000ULL);
        }
        if (bad) { pack_finish(); return fail(ICPB_EINVAL, "transform holds non-finite values%s"); }
    }
    // ---- from here on the handle's table is being replaced: a failure leaves it without one ----
    h->prep_for = nullptr; h->xy = nullptr; h->offsets = nullptr; h->n_scans = 0; h->longest = 0;
    cudaStream_t cp = h->stream, cp2 = h->cstream[1], cs = h->cstream[0];
    if (B == 0) {                                             // upload only
        if (src.scan_xy) { pack_run(&job); pack_finish(); }
        if (src.scan_xy && job.bad.load()) return align_abort(h, fail(ICPB_EINVAL, "scan table holds non-finite coordinates%s"));
        CUA(cudaMemcpyAsync(h->own_xy.p, up_xy, nb_xy, cudaMemcpyHostToDevice, cp));
        CUA(cudaMemcpyAsync(h->own_off.p, h_offsets, nb_off, cudaMemcpyHostToDevice, cp));
        CUA(cudaStreamSynchronize(cp));
        h->prep_for = nullptr; h->xy = d_xy; h->offsets = d_off; h->n_scans = n_scans; h->longest = longest;
        return 0;
    }
    // The first pieces start NOW, before the per-pair staging below: the copy engine works while the
    // host sorts and packs (everything has been validated above; nothing below can fail on user input).
    int64_t s_prev = 0;
    auto enqueue_piece = [&](int k) -> cudaError_t {
        const int64_t o0 = h_offsets[s_prev], o1 = h_offsets[seg_end[k]];
        cudaError_t e = cudaSuccess;
        if (o1 > o0)
            e = cudaMemcpyAsync((double *)h->own_xy.p + 2 * o0, up_xy + 2 * o0, sizeof(double) * 2 * (size_t)(o1 - o0),
                                cudaMemcpyHostToDevice, cp);
        // stream order: the counter changes after the segment it announces has landed
        if (e != cudaSuccess || h->tune_drop_counter) {
            // (drop_counter, tests only: the kernel must time out and the batch be rerun)
        } else if (h->tune_flag_copy || !h->write32 ||
                   h->write32(cp, (unsigned long long)(uintptr_t)h->arrived_dev, (unsigned)(k + 1), 0) != 0) {
            e = cudaMemcpyAsync(h->arrived_dev, h->seg_vals_pinned + (k + 1), sizeof(int32_t), cudaMemcpyHostToDevice, cp);
        }
        s_prev = seg_end[k];
        return e;
    };
    CUP(cudaMemsetAsync(h->arrived_dev, 0, sizeof(int32_t), cp));
    CUP(cudaEventRecord(h->seg_ev[1], cp));                   // the counter is zero before the kernel may start
    int n_sent = 0;
    if (!src.scan_xy) {
        const int n_early = nseg < 4 ? nseg : 4;
        for (; n_sent < n_early; ++n_sent) CUP(enqueue_piece(n_sent));
    }
    // Queue order: pairs grouped by the segment that completes them; the pairs, initial guesses and
    // results themselves stay in the caller's order.  A batch that already comes in arrival order
    // (the odometry chain does) needs no permutation at all.
    bool in_order = true;
    {
        std::vector<int32_t> seg_of_scan((size_t)n_scans);
        for (int64_t sc = 0, k = 0; sc < n_scans; ++sc) { while (sc >= seg_end[k]) ++k; seg_of_scan[sc] = (int32_t)k; }
        int32_t *seg_of_pair = tP;                            // scratch: the download area is free until the end
        int32_t prev = 0, sorted = 1;
        for (int64_t b = 0; b < B; ++b) {
            const int32_t i = h_pairs[2 * b], j = h_pairs[2 * b + 1];
            const int32_t sg = seg_of_scan[i > j ? i : j];
            seg_of_pair[b] = sg;
            sorted &= (int32_t)(sg >= prev);
            prev = sg;
        }
        in_order = sorted != 0;
        if (in_order) {
            memcpy(pseg, seg_of_pair, nb4);
        } else {                                              // stable counting sort
            int64_t start[kMaxSegments + 1] = {0};
            for (int64_t b = 0; b < B; ++b) ++start[seg_of_pair[b] + 1];
            for (int k = 0; k < nseg; ++k) start[k + 1] += start[k];
            for (int64_t b = 0; b < B; ++b) {
                const int64_t q = start[seg_of_pair[b]]++;
                porder[q] = (int32_t)b; pseg[q] = seg_of_pair[b];
            }
        }
    }
    memcpy(ppairs, h_pairs, 2 * nb4);
    if (h_init) {
        if (init_ld == 6) {
            memcpy(pinit, h_init, nbI);
        } else {
            for (int64_t b = 0; b < B; ++b) memcpy(pinit + 6 * b, h_init + 9 * b, 6 * sizeof(double));
        }
    }
    const double t_prep = trace ? now_us() : 0.0;
    // the small per-pair block goes up on a second copy stream, beside the pieces already in flight
    CUP(cudaMemsetAsync(d_passes + B, 0, sizeof(int32_t), cp2));
    CUP(cudaMemcpyAsync(h->own_off.p, h_offsets, nb_off, cudaMemcpyHostToDevice, cp2));
    const size_t nb_idx = (in_order ? 3 : 4) * nb4;           // pairs, seg (, order)
    if (h_init) CUP(cudaMemcpyAsync(d_init, pinit, nbI + nb_idx, cudaMemcpyHostToDevice, cp2));
    else        CUP(cudaMemcpyAsync(d_pairs, ppairs, nb_idx, cudaMemcpyHostToDevice, cp2));
    CUP(cudaEventRecord(h->seg_ev[0], cp2));
    // ONE launch over all pairs, in arrival order; its CTAs wait on the segment counter
    CUP(cudaStreamWaitEvent(cs, h->seg_ev[0], 0));
    CUP(cudaStreamWaitEvent(cs, h->seg_ev[1], 0));
    rc = launch(h, d_xy, d_off, n_scans, longest, d_pairs, h_init ? d_init : nullptr, B, p,
                d_T, d_err, d_passes, nullptr, nullptr, cs, B, d_seg, h->arrived_dev, ep,
                in_order ? nullptr : d_order, d_passes + B);
    if (rc) { pack_finish(); return align_abort(h, rc); }
    bool bad_scans = false;
    double t_ready[kMaxSegments] = {0}, t_enqd[kMaxSegments] = {0};   // trace: piece k packed / enqueued (host clock)
    cudaEvent_t tev[kMaxSegments + 1] = {nullptr};            // trace: when piece k had landed (device clock)
    cudaEvent_t tev0[kMaxSegments] = {nullptr};               // trace: when piece k's copy could start
    if (trace) { cudaEventCreate(&tev[kMaxSegments]); cudaEventRecord(tev[kMaxSegments], cp); }
    for (int k = n_sent; k < nseg; ++k) {
        if (src.scan_xy) {
            // this thread packs too while a piece is incomplete (pack_run returns when no job is left)
            while (piece_done[(size_t)k].load(std::memory_order_acquire) < jobs_of_piece[(size_t)k]) {
                const int64_t j = job.next.fetch_add(1, std::memory_order_relaxed);
                if (j < job.n_jobs) {
                    PackJob one;                              // run exactly job j on this thread
                    one.scan_xy = job.scan_xy; one.offsets = job.offsets; one.dst = job.dst;
                    one.job_first = job.job_first + j; one.job_last = job.job_last + j; one.job_piece = job.job_piece + j;
                    one.n_jobs = 1; one.done = job.done;
                    pack_run(&one);
                    if (one.bad.load()) job.bad.fetch_or(1, std::memory_order_relaxed);
                }
            }
            if (job.bad.load(std::memory_order_relaxed)) { bad_scans = true; break; }
        }
        if (trace) { t_ready[k] = now_us() - t_entry; cudaEventCreate(&tev0[k]); cudaEventRecord(tev0[k], cp); }
        CUP(enqueue_piece(k));
        if (trace) { t_enqd[k] = now_us() - t_entry; cudaEventCreate(&tev[k]); cudaEventRecord(tev[k], cp); }
    }
    pack_finish();
    if (bad_scans) {
        // the kernel is waiting for pieces that will not come: raise its give-up flag, drain, report
        CUA(cudaMemcpyAsync(d_passes + B, h->seg_vals_pinned + 1, sizeof(int32_t), cudaMemcpyHostToDevice, cp));
        return align_abort(h, fail(ICPB_EINVAL, "scan table holds non-finite coordinates%s"));
    }
    CUA(cudaEventRecord(h->done_ev[0], cs));
    CUA(cudaStreamWaitEvent(cp, h->done_ev[0], 0));
    int64_t *acc_cnt = (int64_t *)(tP + B + 2);               // pinned: [timeout word | pad | accepted count]
    if (!acc) {
        CUA(cudaMemcpyAsync(tT, d_T, nb_down, cudaMemcpyDeviceToHost, cp));
    } else {                                                  // the timeout word and the count, not the results
        CUA(cudaMemcpyAsync(tP + B, d_passes + B, sizeof(int32_t), cudaMemcpyDeviceToHost, cp));
        CUA(cudaMemcpyAsync(acc_cnt, h->s_hist.p, sizeof(int64_t), cudaMemcpyDeviceToHost, cp));
    }
    const double t_enq = trace ? now_us() : 0.0;
    CUA(cudaStreamSynchronize(cp));
    if (tP[B] != 0) {
        // a CTA gave up waiting for its scans (see the kernel): by now the whole table is resident,
        // so run the batch again without the streaming protocol
        rc = launch(h, d_xy, d_off, n_scans, longest, d_pairs, h_init ? d_init : nullptr, B, p,
                    d_T, d_err, d_passes, nullptr, nullptr, cp, 0, nullptr, nullptr, ep);
        if (rc) return align_abort(h, rc);
        if (!acc) CUA(cudaMemcpyAsync(tT, d_T, nb_down, cudaMemcpyDeviceToHost, cp));
        else      CUA(cudaMemcpyAsync(acc_cnt, h->s_hist.p, sizeof(int64_t), cudaMemcpyDeviceToHost, cp));
        CUA(cudaStreamSynchronize(cp));
    }
    // only now does the handle own the new table
    h->prep_for = nullptr; h->xy = d_xy; h->offsets = d_off; h->n_scans = n_scans; h->longest = longest;
    const double t_sync = trace ? now_us() : 0.0;
    if (acc) {
        // the accepted records only: fetch, order by pair id (the kernel appended them as they finished)
        const int64_t n = *acc_cnt;
        *acc->n_out = n < 0 ? -n : n;
        if (n < 0) return fail(ICPB_EINVAL, "icpb_align_host_accept: more pairs accepted than the capacity given%s");
        std::vector<double> rec(8 * (size_t)n);
        if (n > 0) {
            CU(cudaMemcpyAsync(rec.data(), (double *)h->s_hist.p + 2, sizeof(double) * 8 * (size_t)n,
                               cudaMemcpyDeviceToHost, cp));
            CU(cudaStreamSynchronize(cp));
        }
        std::vector<int64_t> idx((size_t)n);
        for (int64_t k = 0; k < n; ++k) idx[(size_t)k] = k;
        auto tag_of = [&](int64_t k) { int64_t t; memcpy(&t, &rec[8 * (size_t)k + 7], sizeof t); return t; };
        std::sort(idx.begin(), idx.end(), [&](int64_t x, int64_t y) { return (tag_of(x) >> 16) < (tag_of(y) >> 16); });
        for (int64_t q = 0; q < n; ++q) {
            const double *r = &rec[8 * (size_t)idx[(size_t)q]];
            const int64_t tag = tag_of(idx[(size_t)q]);
            acc->rows[q] = tag >> 16;
            acc->passes[q] = (int32_t)(tag & 0xffff);
            acc->err[q] = r[6];
            double *m = acc->T + (size_t)acc->T_ld * q;
            memcpy(m, r, 6 * sizeof(double));
            if (acc->T_ld == 9) { m[6] = 0.0; m[7] = 0.0; m[8] = 1.0; }
        }
        return 0;
    }
    if (T_ld == 6) {
        memcpy(h_T, tT, nbI);
    } else {
        for (int64_t b = 0; b < B; ++b) {
            double *m = h_T + 9 * b;
            memcpy(m, tT + 6 * b, 6 * sizeof(double));
            m[6] = 0.0; m[7] = 0.0; m[8] = 1.0;
        }
    }
    memcpy(h_err, tE, nbE);
    memcpy(h_passes, tP, nb4);
    if (trace) {
        fprintf(stderr, "[icpb_align_host] prep %.0f us, enqueue %.0f us, wait %.0f us, scatter %.0f us (%d segments)\n",
                t_prep - t_entry, t_enq - t_prep, t_sync - t_enq, now_us() - t_sync, nseg);
        fprintf(stderr, "  piece: packed at / enqueued at (host us) / landed at (us after the copy stream's first event):");
        for (int k = 0; k < nseg; ++k) {
            float ms = 0.f, ms0 = 0.f;
            if (tev[k]) { cudaEventElapsedTime(&ms, tev[kMaxSegments], tev[k]); cudaEventDestroy(tev[k]); }
            if (tev0[k]) { cudaEventElapsedTime(&ms0, tev[kMaxSegments], tev0[k]); cudaEventDestroy(tev0[k]); }
            fprintf(stderr, " %d:%.0f/%.0f/%.0f-%.0f", k, t_ready[k], t_enqd[k], ms0 * 1e3, ms * 1e3);
        }
        fprintf(stderr, "\n");
        cudaEventDestroy(tev[kMaxSegments]);
    }
    return 0;
#undef CUP
}

int align_check(icpb_handle h, int64_t n_scans, const int32_t *h_pairs, int32_t init_ld, int64_t B,
                const icpb_params *p, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes,
                const icpb_epilogue *ep)
{
    if (!h || n_scans <= 0) return fail(ICPB_EINVAL, "icpb_align_host: bad argument%s");
    if ((init_ld != 6 && init_ld != 9) || (T_ld != 6 && T_ld != 9))
        return fail(ICPB_EINVAL, "icpb_align_host: init_ld / T_ld must be 6 (2x3 rows) or 9 (3x3)%s");
    int rc = check_params(p, B);
    if (rc) return rc;
    if ((rc = check_epilogue(ep, p))) return rc;
    if (p->pair_mode != 0 || p->hist_cap > 0 || p->corr_stride > 0)
        return fail(ICPB_EINVAL, "icpb_align_host: explicit pairs, no history/correspondences (use upload + run)%s");
    if (B >= (int64_t(1) << 31)) return fail(ICPB_EINVAL, "icpb_align_host: at most 2^31 - 1 pairs per call%s");
    if (B > 0 && (!h_pairs || !h_T || !h_err || !h_passes)) return fail(ICPB_EINVAL, "null pointer%s");
    return 0;
}

}  // namespace

int icpb_align_host_ex(icpb_handle h, const double *h_xy, const int64_t *h_offsets, int64_t n_scans,
                       const int32_t *h_pairs, const double *h_init, int32_t init_ld, int64_t B,
                       const icpb_params *p, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes,
                       const icpb_epilogue *ep)
{
    if (!h_xy || !h_offsets) return fail(ICPB_EINVAL, "icpb_align_host: bad argument%s");
    int rc = align_check(h, n_scans, h_pairs, init_ld, B, p, h_T, T_ld, h_err, h_passes, ep);
    if (rc) return rc;
    int64_t longest = 0;
    if ((rc = validate_offsets(h_offsets, n_scans, &longest))) return rc;
    TableSrc src;
    src.xy = h_xy; src.offsets = h_offsets;
    return align_impl(h, src, n_scans, longest, h_pairs, h_init, init_ld, B, p, h_T, T_ld, h_err, h_passes, ep);
}

int icpb_align_host_scans(icpb_handle h, const double *const *scan_xy, const int64_t *scan_len, int64_t n_scans,
                          const int32_t *h_pairs, const double *h_init, int32_t init_ld, int64_t B,
                          const icpb_params *p, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes,
                          const icpb_epilogue *ep)
{
    if (!scan_xy || !scan_len) return fail(ICPB_EINVAL, "icpb_align_host_scans: bad argument%s");
    int rc = align_check(h, n_scans, h_pairs, init_ld, B, p, h_T, T_ld, h_err, h_passes, ep);
    if (rc) return rc;
    std::vector<int64_t> offsets((size_t)n_scans + 1);
    offsets[0] = 0;
    int64_t longest = 0;
    for (int64_t s = 0; s < n_scans; ++s) {
        if (!scan_xy[s] || scan_len[s] <= 0)
            return fail(ICPB_EINVAL, "empty scan in the list (the reference's argmin raises on it)%s");
        offsets[(size_t)s + 1] = offsets[(size_t)s] + scan_len[s];
        if (scan_len[s] > longest) longest = scan_len[s];
    }
    TableSrc src;
    src.scan_xy = scan_xy; src.offsets = offsets.data();
    return align_impl(h, src, n_scans, longest, h_pairs, h_init, init_ld, B, p, h_T, T_ld, h_err, h_passes, ep);
}

int icpb_align_host_accept(icpb_handle h, const double *h_xy, const int64_t *h_offsets,
                           const double *const *scan_xy, const int64_t *scan_len, int64_t n_scans,
                           const int32_t *h_pairs, const double *h_init, int32_t init_ld, int64_t B,
                           const icpb_params *p, double accept_thresh, int64_t capacity, int64_t *n_accepted,
                           int64_t *h_rows, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes)
{
    if (!h || n_scans <= 0 || B <= 0 || !h_pairs) return fail(ICPB_EINVAL, "icpb_align_host_accept: bad argument%s");
    if (!((h_xy && h_offsets) || (scan_xy && scan_len))) return fail(ICPB_EINVAL, "icpb_align_host_accept: no scans%s");
    if ((init_ld != 6 && init_ld != 9) || (T_ld != 6 && T_ld != 9))
        return fail(ICPB_EINVAL, "icpb_align_host_accept: init_ld / T_ld must be 6 or 9%s");
    if (capacity < 1 || !n_accepted || !h_rows || !h_T || !h_err || !h_passes || isnan(accept_thresh))
        return fail(ICPB_EINVAL, "icpb_align_host_accept: bad output arguments%s");
    int rc = check_params(p, B);
    if (rc) return rc;
    if (p->pair_mode != 0 || p->hist_cap > 0 || p->corr_stride > 0)
        return fail(ICPB_EINVAL, "icpb_align_host_accept: explicit pairs, no history/correspondences%s");
    if (B >= (int64_t(1) << 31)) return fail(ICPB_EINVAL, "icpb_align_host_accept: at most 2^31 - 1 pairs per call%s");
    AcceptHost acc = {accept_thresh, capacity, n_accepted, h_rows, h_T, T_ld, h_err, h_passes};
    TableSrc src;
    std::vector<int64_t> offsets;
    int64_t longest = 0;
    if (h_xy) {
        if ((rc = validate_offsets(h_offsets, n_scans, &longest))) return rc;
        src.xy = h_xy; src.offsets = h_offsets;
    } else {
        offsets.resize((size_t)n_scans + 1);
        offsets[0] = 0;
        for (int64_t sc = 0; sc < n_scans; ++sc) {
            if (!scan_xy[sc] || scan_len[sc] <= 0)
                return fail(ICPB_EINVAL, "empty scan in the list (the reference's argmin raises on it)%s");
            offsets[(size_t)sc + 1] = offsets[(size_t)sc] + scan_len[sc];
            if (scan_len[sc] > longest) longest = scan_len[sc];
        }
        src.scan_xy = scan_xy; src.offsets = offsets.data();
    }
    return align_impl(h, src, n_scans, longest, h_pairs, h_init, init_ld, B, p, nullptr, T_ld, nullptr, nullptr,
                      nullptr, &acc);
}

int icpb_align_host_ld(icpb_handle h, const double *h_xy, const int64_t *h_offsets, int64_t n_scans,
                       const int32_t *h_pairs, const double *h_init, int32_t init_ld, int64_t B,
                       const icpb_params *p, double *h_T, int32_t T_ld, double *h_err, int32_t *h_passes)
{
    return icpb_align_host_ex(h, h_xy, h_offsets, n_scans, h_pairs, h_init, init_ld, B, p, h_T, T_ld, h_err, h_passes,
                              nullptr);
}

int icpb_align_host(icpb_handle h, const double *h_xy, const int64_t *h_offsets, int64_t n_scans,
                    const int32_t *h_pairs, const double *h_init, int64_t B, const icpb_params *p,
                    double *h_T, double *h_err, int32_t *h_passes)
{
    return icpb_align_host_ld(h, h_xy, h_offsets, n_scans, h_pairs, h_init, 6, B, p, h_T, 6, h_err, h_passes);
}

int icpb_icp_pair_host(icpb_handle h, const double *h_src_xy, int64_t n_src,
                       const double *h_dst_xy, int64_t n_dst, const double *h_init6,
                       const icpb_params *p, double *h_T6, double *h_err, int32_t *h_passes,
                       double *h_hist, int32_t *h_corr)
{
    if (!h) return fail(ICPB_EINVAL, "handle is null%s");
    if (!h_src_xy || !h_dst_xy || n_src <= 0 || n_dst <= 0)
        return fail(ICPB_EINVAL, "empty point cloud (the reference's argmin raises on it)%s");
    int rc = check_params(p, 1);
    if (rc) return rc;
    if (!h_T6 || !h_err || !h_passes) return fail(ICPB_EINVAL, "output pointer is null%s");
    if (p->hist_cap > 0 && !h_hist) return fail(ICPB_EINVAL, "hist_cap > 0 but hist is null%s");
    if (p->corr_stride > 0 && (!h_corr || p->corr_stride < n_src))
        return fail(ICPB_EINVAL, "corr buffer missing or shorter than the source cloud%s");
    icpb_params q = *p;
    q.pair_mode = 0; q.k_first = 0; q.k_block = 1; q.k_stride = 0;
    ON_DEVICE(h);
    const size_t nb = sizeof(double) * 2 * (size_t)(n_src + n_dst);
    if ((rc = h->s_pair_xy.reserve(nb))) return rc;
    if ((rc = h->s_pair_off.reserve(sizeof(int64_t) * 3))) return rc;
    const int64_t off[3] = {0, n_src, n_src + n_dst};
    const int32_t pair[2] = {0, 1};
    double *dxy = (double *)h->s_pair_xy.p;
    CU(cudaMemcpyAsync(dxy, h_src_xy, sizeof(double) * 2 * (size_t)n_src, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dxy + 2 * n_src, h_dst_xy, sizeof(double) * 2 * (size_t)n_dst, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->s_pair_off.p, off, sizeof off, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));       // `off` lives on this stack frame
    return run_host_common(h, dxy, (const int64_t *)h->s_pair_off.p, 2, n_src > n_dst ? n_src : n_dst,
                           pair, h_init6, 1, &q, h_T6, h_err, h_passes, h_hist, h_corr);
}

int icpb_fit_pairs_host(icpb_handle h, const double *h_a_xy, const double *h_b_xy, int64_t n,
                        double *h_T6, double *h_err)
{
    if (!h || !h_a_xy || !h_b_xy || n <= 0 || n > 0x7fffffff || !h_T6 || !h_err)
        return fail(ICPB_EINVAL, "icpb_fit_pairs_host: bad argument%s");
    ON_DEVICE(h);
    int rc;
    const size_t nb = sizeof(double) * 2 * (size_t)n;
    if ((rc = h->s_pair_xy.reserve(2 * nb))) return rc;
    if ((rc = h->s_T.reserve(sizeof(double) * 8))) return rc;
    double *da = (double *)h->s_pair_xy.p, *db = da + 2 * n;
    CU(cudaMemcpyAsync(da, h_a_xy, nb, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(db, h_b_xy, nb, cudaMemcpyHostToDevice, h->stream));
    icpb::fit_pairs_kernel<<<1, 256, 0, h->stream>>>((const double2 *)da, (const double2 *)db, (int)n,
                                                     (double *)h->s_T.p, (double *)h->s_T.p + 6);
    CU(cudaGetLastError());
    double out[7];
    CU(cudaMemcpyAsync(out, h->s_T.p, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(h_T6, out, 6 * sizeof(double));
    *h_err = out[6];
    return 0;
}

static int upload_poses(icpb_handle h, const double *h_xy, const double *h_travelled, int64_t n,
                        double **d_xy, double **d_trav)
{
    int rc;
    if ((rc = h->s_pair_xy.reserve(sizeof(double) * 3 * (size_t)n))) return rc;
    *d_xy = (double *)h->s_pair_xy.p;
    *d_trav = *d_xy + 2 * n;
    CU(cudaMemcpyAsync(*d_xy, h_xy, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(*d_trav, h_travelled, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int icpb_proximity_closest(icpb_handle h, const double *h_xy, const double *h_travelled, int64_t n,
                           double min_dist_along_path, double max_dist, int32_t *h_closest, double *h_dist)
{
    if (!h || !h_xy || !h_travelled || n <= 0 || n > 0x3fffffff || !h_closest || !h_dist)
        return fail(ICPB_EINVAL, "icpb_proximity_closest: bad argument%s");
    ON_DEVICE(h);
    double *d_xy, *d_trav;
    int rc = upload_poses(h, h_xy, h_travelled, n, &d_xy, &d_trav);
    if (rc) return rc;
    if ((rc = h->s_passes.reserve(sizeof(int32_t) * (size_t)n))) return rc;
    if ((rc = h->s_err.reserve(sizeof(double) * (size_t)n))) return rc;
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    icpb::proximity_closest_kernel<<<grid, 256, 0, h->stream>>>((const double2 *)d_xy, d_trav, (int)n,
                                                                min_dist_along_path, max_dist,
                                                                (int32_t *)h->s_passes.p, (double *)h->s_err.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_closest, h->s_passes.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h_dist, h->s_err.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int icpb_proximity_pairs(icpb_handle h, const double *h_xy, const double *h_travelled, int64_t n,
                         double min_dist_along_path, double max_dist, int64_t capacity,
                         int32_t *h_pairs, int64_t *n_pairs)
{
    if (!h || !h_xy || !h_travelled || n <= 0 || n > 0x3fffffff || !n_pairs || capacity < 0 ||
        (capacity > 0 && !h_pairs))
        return fail(ICPB_EINVAL, "icpb_proximity_pairs: bad argument%s");
    ON_DEVICE(h);
    double *d_xy, *d_trav;
    int rc = upload_poses(h, h_xy, h_travelled, n, &d_xy, &d_trav);
    if (rc) return rc;
    if ((rc = h->s_T.reserve(sizeof(int64_t) * 2 * (size_t)n))) return rc;
    int64_t *d_count = (int64_t *)h->s_T.p, *d_off = d_count + n;
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    icpb::proximity_pairs_kernel<false><<<grid, 256, 0, h->stream>>>((const double2 *)d_xy, d_trav, (int)n,
                                                                     min_dist_along_path, max_dist,
                                                                     d_count, nullptr, nullptr);
    CU(cudaGetLastError());
    std::vector<int64_t> cnt((size_t)n), off((size_t)n);
    CU(cudaMemcpyAsync(cnt.data(), d_count, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) { off[i] = total; total += cnt[i]; }      // S values: a host scan is enough
    *n_pairs = total;
    if (capacity == 0 || total == 0) return 0;                                // size query
    if (total > capacity) return fail(ICPB_EINVAL, "icpb_proximity_pairs: capacity too small%s");
    if ((rc = h->s_pairs.reserve(sizeof(int32_t) * 2 * (size_t)total))) return rc;
    CU(cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    icpb::proximity_pairs_kernel<true><<<grid, 256, 0, h->stream>>>((const double2 *)d_xy, d_trav, (int)n,
                                                                    min_dist_along_path, max_dist,
                                                                    nullptr, d_off, (int32_t *)h->s_pairs.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_pairs, h->s_pairs.p, sizeof(int32_t) * 2 * (size_t)total, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

/* Host epilogue of the odometry fan-out: the serial SE(2) prefix product of scripts/main.py:249-256,
 * pose_i = mat_to_pose(pose_to_mat(pose_{i-1}) @ T_{i-1}), with the reference's own sequence of
 * operations (cos/sin of the heading, 3x3 product, atan2) so the poses match it to rounding.  It
 * is a 5,000-step dependency chain of a few flops each: a host loop in C (0.3 ms) beats both the
 * reference's Python loop (~100 ms) and a kernel launch. */
int icpb_compose_chain(const double *pose0, const double *T6, int64_t n, double *poses_out)
{
    if (!pose0 || (n > 0 && !T6) || n < 0 || !poses_out) return fail(ICPB_EINVAL, "icpb_compose_chain: bad argument%s");
    poses_out[0] = pose0[0]; poses_out[1] = pose0[1]; poses_out[2] = pose0[2];
    for (int64_t i = 0; i < n; ++i) {
        const double *p = poses_out + 3 * i, *T = T6 + 6 * i;
        const double c = cos(p[2]), s = sin(p[2]);
        // rows of pose_to_mat(p) @ [T; 0 0 1], summed left to right like a 3-term dot product
        const double m00 = c * T[0] + -s * T[3];
        const double m10 = s * T[0] + c * T[3];
        const double x = c * T[2] + -s * T[5] + p[0];
        const double y = s * T[2] + c * T[5] + p[1];
        double *q = poses_out + 3 * (i + 1);
        q[0] = x; q[1] = y; q[2] = atan2(m10, m00);
    }
    return 0;
}

/* The same prefix product as a parallel scan on the device (csrc/icpb_compose.cuh).  d_T6 may be the
 * d_T output of icpb_run_device: the transforms then never visit the host. */
int icpb_compose_chain_device(icpb_handle h, const double *pose0, const double *d_T6, int64_t n,
                              double *d_poses_out, void *stream)
{
    if (!h || !pose0 || (n > 0 && !d_T6) || n < 0 || !d_poses_out)
        return fail(ICPB_EINVAL, "icpb_compose_chain_device: bad argument%s");
    ON_DEVICE(h);
    icpb::compose_chain_kernel<<<1, icpb::kComposeThreads, 0, (cudaStream_t)stream>>>(d_T6, n, pose0[0], pose0[1], pose0[2],
                                                                                      d_poses_out);
    CU(cudaGetLastError());
    return 0;
}

/* Host buffers in and out, through the device scan (for the comparison with the host loop). */
int icpb_compose_chain_gpu(icpb_handle h, const double *pose0, const double *h_T6, int64_t n, double *h_poses_out)
{
    if (!h || !pose0 || (n > 0 && !h_T6) || n < 0 || !h_poses_out)
        return fail(ICPB_EINVAL, "icpb_compose_chain_gpu: bad argument%s");
    ON_DEVICE(h);
    int rc;
    if ((rc = h->s_pair_xy.reserve(sizeof(double) * (6 * (size_t)n + 3 * (size_t)(n + 1))))) return rc;
    double *d_T = (double *)h->s_pair_xy.p, *d_P = d_T + 6 * n;
    if (n > 0) CU(cudaMemcpyAsync(d_T, h_T6, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if ((rc = icpb_compose_chain_device(h, pose0, d_T, n, d_P, h->stream))) return rc;
    CU(cudaMemcpyAsync(h_poses_out, d_P, sizeof(double) * 3 * (size_t)(n + 1), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

/* Pose-graph SGD (include/icpb.h): n_steps passes of the reference's
 * pose_graph_optimization_step_sgd over the same edge list, poses updated in place. */
int icpb_pose_graph_sgd(icpb_handle h, double *h_poses, int64_t n, const int32_t *h_edges,
                        const double *h_edge_T6, int64_t n_edges, const double *h_learning_rates,
                        int32_t n_steps, double loop_closure_uncertainty)
{
    if (!h || !h_poses || n <= 0 || n > 0x1fffffff || n_edges < 0 || n_edges > 0x1fffffff || n_steps < 0 ||
        (n_edges > 0 && (!h_edges || !h_edge_T6)) || (n_steps > 0 && !h_learning_rates))
        return fail(ICPB_EINVAL, "icpb_pose_graph_sgd: bad argument%s");
    if (!(loop_closure_uncertainty > 0.0)) return fail(ICPB_EINVAL, "icpb_pose_graph_sgd: loop_closure_uncertainty must be > 0%s");
    for (int64_t e = 0; e < 2 * n_edges; ++e)
        if (h_edges[e] < 0 || h_edges[e] >= n) return fail(ICPB_EINVAL, "icpb_pose_graph_sgd: edge endpoint out of range%s");
    // Edges the optimiser ignores (|a - b| == 1, :14-16 and :28-30) or that have an empty node range
    // (b <= a: `range(a+1, b+1)` :20 and `i <= b` :46 never hold) move nothing: they stay on the host.
    std::vector<int32_t> ve;
    std::vector<double> vt;
    for (int64_t e = 0; e < n_edges; ++e) {
        const int32_t ea = h_edges[2 * e], eb = h_edges[2 * e + 1];
        if (eb > ea + 1) {
            ve.push_back(ea); ve.push_back(eb);
            vt.insert(vt.end(), h_edge_T6 + 6 * e, h_edge_T6 + 6 * e + 6);
        }
    }
    if (ve.empty() || n_steps == 0) return 0;
    ON_DEVICE(h);
    const size_t E = ve.size() / 2, N = (size_t)n;
    // [RECA 10E | tf 6E | dW 4E | REC 10E | ES 23E | poses 3N | M 3N | P 3N] doubles, then [edges 2E] int32
    // (RECA first: its 80-byte records are read as 16-byte pairs)
    const size_t n_dbl = 10 * E + 6 * E + 4 * E + 10 * E + 23 * E + 3 * N + 3 * N + 3 * N;
    int rc;
    if ((rc = h->s_sgd.reserve(sizeof(double) * n_dbl + sizeof(int32_t) * 2 * E))) return rc;
    double *d_RECA = (double *)h->s_sgd.p, *d_tf = d_RECA + 10 * E, *d_dW = d_tf + 6 * E;
    double *d_REC = d_dW + 4 * E, *d_ES = d_REC + 10 * E, *d_poses = d_ES + 23 * E, *d_M = d_poses + 3 * N, *d_P = d_M + 3 * N;
    int32_t *d_edges = (int32_t *)(d_P + 3 * N);
    CU(cudaMemcpyAsync(d_poses, h_poses, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_tf, vt.data(), sizeof(double) * 6 * E, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_edges, ve.data(), sizeof(int32_t) * 2 * E, cudaMemcpyHostToDevice, h->stream));
    icpb::SgdArgs a;
    a.poses = d_poses; a.edges = d_edges; a.tf = d_tf; a.n = (int32_t)n; a.E = (int32_t)E;
    a.lcu = loop_closure_uncertainty; a.dW = d_dW; a.REC = d_REC; a.RECA = d_RECA; a.ES = d_ES; a.M = d_M; a.P = d_P;
    // per pass: weights of the edges, their sum per node, the lazy chain over the edges on one SM (the
    // only sequential part), then every record applied to every node on all SMs
    const dim3 node_grid((unsigned)((N + icpb::kSgdNodeThreads - 1) / icpb::kSgdNodeThreads), 3);   // (nodes, dof)
    for (int32_t k = 0; k < n_steps; ++k) {
        a.learning_rate = h_learning_rates[k];
        icpb::sgd_weights_kernel<<<(unsigned)((E + 255) / 256), 256, 0, h->stream>>>(a);
        icpb::sgd_accumulate_kernel<<<node_grid, icpb::kSgdNodeThreads, 0, h->stream>>>(a);
        icpb::sgd_chain_kernel<<<1, icpb::kSgdThreads, 0, h->stream>>>(a);
        icpb::sgd_apply_kernel<<<node_grid, icpb::kSgdNodeThreads, 0, h->stream>>>(a);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(h_poses, d_poses, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

/* ---- occupancy grid (include/icpb.h; reference src/produce_occupancy_grid.py) ---- */
namespace {

int grid_check(icpb_handle h, const double *h_poses, int64_t n, double cell_width, const char *who)
{
    if (!h || !h_poses || n <= 0 || !(cell_width > 0.0)) return fail(ICPB_EINVAL, "%s: bad argument", who);
    if (!h->xy) return fail(ICPB_ENOSCANS, "%s: no scan table set", who);
    if (n > h->n_scans) return fail(ICPB_EINVAL, "%s: more poses than scans in the table", who);
    return 0;
}

// (cos theta, sin theta, x, y) per pose: odom_change_to_mat's trigonometry (src/utils.py:8-9) in libm
std::vector<double> pose_records(const double *h_poses, int64_t n)
{
    std::vector<double> r(4 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        r[4 * i] = cos(h_poses[3 * i + 2]); r[4 * i + 1] = sin(h_poses[3 * i + 2]);
        r[4 * i + 2] = h_poses[3 * i]; r[4 * i + 3] = h_poses[3 * i + 1];
    }
    return r;
}

double key_to_double(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
}

}  // namespace

int icpb_occupancy_grid_bounds(icpb_handle h, const double *h_poses, int64_t n, double cell_width,
                               double min_width, double min_height, double *min_x_out, double *min_y_out,
                               int64_t *height, int64_t *width)
{
    int rc = grid_check(h, h_poses, n, cell_width, "icpb_occupancy_grid_bounds");
    if (rc) return rc;
    if (!min_x_out || !min_y_out || !height || !width) return fail(ICPB_EINVAL, "icpb_occupancy_grid_bounds: null output%s");
    ON_DEVICE(h);
    if ((rc = h->s_grid.reserve(sizeof(double) * 4 * (size_t)n + 4 * sizeof(unsigned long long)))) return rc;
    unsigned long long *d_mm = (unsigned long long *)h->s_grid.p;
    double *d_poses = (double *)(d_mm + 4);
    const unsigned long long init[4] = {~0ULL, ~0ULL, 0ULL, 0ULL};
    const std::vector<double> rec = pose_records(h_poses, n);
    CU(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_poses, rec.data(), sizeof(double) * 4 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    icpb::GridArgs a = {};
    a.xy = h->xy; a.offsets = h->offsets; a.poses = d_poses; a.n = (int32_t)n;
    icpb::grid_bounds_kernel<<<(unsigned)n, 256, 0, h->stream>>>(a, d_mm);
    CU(cudaGetLastError());
    unsigned long long mm[4];
    CU(cudaMemcpyAsync(mm, d_mm, sizeof mm, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (mm[0] == ~0ULL) return fail(ICPB_EINVAL, "icpb_occupancy_grid_bounds: the scans hold no points%s");
    // src/produce_occupancy_grid.py:30-51, operation for operation
    double min_x = key_to_double(mm[0]) - (cell_width / 2), max_x = key_to_double(mm[2]) + (cell_width / 2);
    double min_y = key_to_double(mm[1]) - (cell_width / 2), max_y = key_to_double(mm[3]) + (cell_width / 2);
    double width_dist = max_x - min_x, height_dist = max_y - min_y;
    if (width_dist < min_width) { const double off = (min_width - width_dist) / 2; min_x -= off; width_dist = min_width; }
    if (height_dist < min_height) { const double off = (min_height - height_dist) / 2; min_y -= off; height_dist = min_height; }
    *min_x_out = min_x; *min_y_out = min_y;
    *width = (int64_t)ceil(width_dist / cell_width);
    *height = (int64_t)ceil(height_dist / cell_width);
    return 0;
}

int icpb_occupancy_grid_update(icpb_handle h, const double *h_poses, int64_t n, int8_t *h_grid,
                               int64_t height, int64_t width, double min_x, double min_y,
                               double cell_width, int32_t k_hit, int32_t k_miss)
{
    int rc = grid_check(h, h_poses, n, cell_width, "icpb_occupancy_grid_update");
    if (rc) return rc;
    if (!h_grid || height <= 0 || width <= 0 || height > 0x7fffffff || width > 0x7fffffff ||
        height * width > ((int64_t)1 << 30))
        return fail(ICPB_EINVAL, "icpb_occupancy_grid_update: bad grid%s");
    // the per-cell closed form needs a miss to leave a cell negative and a hit to leave it positive
    if (k_hit < 1 || k_hit > 127 || k_miss < 1 || k_miss > 127)
        return fail(ICPB_EINVAL, "icpb_occupancy_grid_update: kHitOdds and kMissOdds must be integers in 1..127%s");
    ON_DEVICE(h);
    // the beam order keys are 32 bit: 2 * (beam index + 1) + 1 must fit (2^31 - 2 beams)
    int64_t n_points = 0;
    CU(cudaMemcpy(&n_points, h->offsets + n, sizeof n_points, cudaMemcpyDeviceToHost));
    if (n_points >= 0x7ffffffeLL) return fail(ICPB_EINVAL, "icpb_occupancy_grid_update: more than 2^31 - 2 beams in one call%s");
    const size_t cells = (size_t)height * (size_t)width;
    const size_t words = 4 * cells;
    if ((rc = h->s_grid.reserve(sizeof(double) * 4 * (size_t)n + sizeof(uint32_t) * words + cells + 64))) return rc;
    double *d_poses = (double *)h->s_grid.p;
    uint32_t *d_words = (uint32_t *)(d_poses + 4 * (size_t)n);
    int8_t *d_grid = (int8_t *)(d_words + words);
    const std::vector<double> rec = pose_records(h_poses, n);
    CU(cudaMemcpyAsync(d_poses, rec.data(), sizeof(double) * 4 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_grid, h_grid, cells, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(d_words, 0, sizeof(uint32_t) * words, h->stream));
    icpb::GridArgs a = {};
    a.xy = h->xy; a.offsets = h->offsets; a.poses = d_poses; a.n = (int32_t)n;
    a.min_x = min_x; a.min_y = min_y; a.cell = cell_width; a.h = (int32_t)height; a.w = (int32_t)width;
    a.last_hit = d_words; a.n_miss = d_words + cells; a.n_hit = d_words + 2 * cells; a.miss_after = d_words + 3 * cells;
    icpb::grid_hits_kernel<<<(unsigned)n, 256, 0, h->stream>>>(a);
    CU(cudaGetLastError());
    icpb::grid_misses_kernel<<<(unsigned)n, 256, 0, h->stream>>>(a);
    CU(cudaGetLastError());
    icpb::grid_finalize_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(a, d_grid, k_hit, k_miss);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_grid, d_grid, cells, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int icpb_get_kernel_info(icpb_handle h, int64_t B, icpb_kernel_info *out)
{
    if (!h || !out) return fail(ICPB_EINVAL, "icpb_get_kernel_info: bad argument%s");
    if (!h->xy) return fail(ICPB_ENOSCANS, "no scan table set%s");
    ON_DEVICE(h);
    LaunchCfg cfg;
    kernel_fn fn = pick_kernel(nullptr);
    int rc = make_cfg(h, h->longest, B > 0 ? B : (int64_t)1 << 40, &cfg, fn);
    if (rc) return rc;
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, fn));
    int64_t grid = (int64_t)cfg.ctas_per_sm * h->sm_count;
    if (B > 0 && grid > B) grid = B;
    out->threads_per_cta = cfg.threads; out->ctas_per_sm = cfg.ctas_per_sm; out->sm_count = h->sm_count;
    out->grid = (int32_t)grid; out->regs_per_thread = fa.numRegs; out->smem_bytes = cfg.smem;
    out->points_per_thread = kPointsPerThread; out->variant = 2;
    return 0;
}

int64_t icpb_launch_count(icpb_handle h) { return h ? h->launches : 0; }

int64_t icpb_scan_count(icpb_handle h) { return h && h->xy ? h->n_scans : 0; }

int icpb_count_work(icpb_handle h, int enable)
{
    if (!h) return fail(ICPB_EINVAL, "handle is null%s");
    ON_DEVICE(h);
    if (enable) {
        CU(cudaMemset(h->queue + kQueueRing, 0, sizeof(unsigned long long)));
        h->executed = h->queue + kQueueRing;
    } else {
        h->executed = nullptr;
    }
    return 0;
}

int icpb_read_work(icpb_handle h, uint64_t *executed_pde)
{
    if (!h || !executed_pde) return fail(ICPB_EINVAL, "icpb_read_work: bad argument%s");
    ON_DEVICE(h);
    CU(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, h->queue + kQueueRing, sizeof v, cudaMemcpyDeviceToHost));
    *executed_pde = v;
    return 0;
}

}  // extern "C"
