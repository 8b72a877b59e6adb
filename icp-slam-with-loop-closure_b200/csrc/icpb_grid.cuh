// Occupancy-grid update on the GPU: the reference's
// produce_occupancy_grid / update_occupancy_grid (reference src/produce_occupancy_grid.py:11-80),
// which walk one Bresenham line per LiDAR beam in pure Python -- every crossed cell gets a "miss"
// (:109-112), the last cell a "hit" (:128-131), saturating in int8.
//
// The update of a cell is order dependent (a hit after a miss is not a miss after a hit), and the
// reference applies the beams one after the other.  What makes the problem parallel is the
// reference's own int8 arithmetic: `-128 - grid[y, x]` and `127 - grid[y, x]` are evaluated in int8
// and wrap, so
//     miss(g) = g > 0 ? -128 : max(g - kMiss, -128)        hit(g) = g < 0 ? 127 : min(g + kHit, 127)
// After a miss the cell is negative, after a hit positive; a change of event type therefore throws
// the cell to the opposite rail, and the value after ANY sequence of events is a function of
//     the initial value, the number of misses, the number of hits, and the TYPE OF THE LAST EVENT
// (tests/test_grid_reference.py::test_per_cell_closed_form replays this against numpy int8 scalars).
// Counts add, and "which type came last" is a comparison of beam-order keys: pass 1 records every
// cell's hit count and the order key of its LAST hit (a maximum; hits need no walking, two atomics
// per beam), pass 2 walks every beam concurrently, counts the misses (one RED.ADD in L2 per visited
// cell) and flags the cells that see a miss after their last hit, and a final pass over the cells
// applies the closed form.  The grid that comes out is bit-identical to the
// reference's sequential loops.  HBM/L2-bound integer work: no tensor cores, no shared-memory
// staging (the per-cell words of a 20 m x 12 m map at 5 cm are 1.2 MB and live in L2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icpb {

struct GridArgs {
    const double  *xy;        // scan table (sum m_i, 2), points in each scan's local frame
    const int64_t *offsets;   // CSR offsets
    const double  *poses;     // n x 4 (cos theta, sin theta, x, y): the trigonometry is done on the host
                              // with libm, whose bits numpy's cos/sin match (src/utils.py:8-9); CUDA's
                              // sincos can differ in the last place, which would move cell boundaries
    int32_t        n;         // scans / poses
    double         min_x, min_y, cell;
    int32_t        h, w;      // grid size in cells
    uint32_t      *last_hit;  // h*w: max over the cell's hits of 2 * (beam order + 1) + 1; 0 = never hit
    uint32_t      *n_miss;    // h*w
    uint32_t      *n_hit;     // h*w
    uint32_t      *miss_after;// h*w: 1 if some miss on the cell comes after its last hit
};

// global = odom_change_to_mat(pose) @ [x, y, 1] (src/produce_occupancy_grid.py:90-93, src/utils.py:3-19)
// with the rounding of numpy's (3,3) @ (3,1) product on the reference's BLAS: fma(m00, x, m01*y) + m02
// (matched bit for bit by tests/test_grid_reference.py::test_global_points_bit_exact)
__device__ __forceinline__ void grid_global_point(double c, double s, double px, double py, double x, double y,
                                                  double &gx, double &gy)
{
    gx = fma(c, x, __dmul_rn(-s, y)) + px;
    gy = fma(s, x, __dmul_rn(c, y)) + py;
}

__device__ __forceinline__ unsigned long long grid_key(double v)       // order-preserving double -> u64
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// Bounding box of all global points (:29-33): one CTA per scan, four 64-bit atomics per CTA.
// mm = [min_x_key, min_y_key, max_x_key, max_y_key], initialised to [~0, ~0, 0, 0].
__global__ void __launch_bounds__(256)
grid_bounds_kernel(const GridArgs a, unsigned long long *mm)
{
    const int i = blockIdx.x;
    const double c = a.poses[4 * i], s = a.poses[4 * i + 1], px = a.poses[4 * i + 2], py = a.poses[4 * i + 3];
    unsigned long long lox = ~0ULL, loy = ~0ULL, hix = 0ULL, hiy = 0ULL;
    for (int64_t k = a.offsets[i] + threadIdx.x; k < a.offsets[i + 1]; k += blockDim.x) {
        double gx, gy;
        grid_global_point(c, s, px, py, a.xy[2 * k], a.xy[2 * k + 1], gx, gy);
        const unsigned long long kx = grid_key(gx), ky = grid_key(gy);
        lox = min(lox, kx); hix = max(hix, kx); loy = min(loy, ky); hiy = max(hiy, ky);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o)); loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o));
        hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o)); hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 0, lox); atomicMin(mm + 1, loy); atomicMax(mm + 2, hix); atomicMax(mm + 3, hiy);
    }
}

// floor((pos - min) / cell_width) (:133-138)
__device__ __forceinline__ long long grid_cell(double pos, double mn, double cell)
{
    return (long long)floor(__ddiv_rn(__dsub_rn(pos, mn), cell));
}

// The cells of one beam: start (robot) and end (beam end point) cell, and whether anything happens.
struct GridBeam {
    long long x0, y0, x1, y1;
};
__device__ __forceinline__ GridBeam grid_beam(const GridArgs &a, double c, double s, double px, double py, int64_t k)
{
    double gx, gy;
    grid_global_point(c, s, px, py, a.xy[2 * k], a.xy[2 * k + 1], gx, gy);
    GridBeam b;
    b.y0 = grid_cell(py, a.min_y, a.cell); b.x0 = grid_cell(px, a.min_x, a.cell);          // :98
    b.y1 = grid_cell(gy, a.min_y, a.cell); b.x1 = grid_cell(gx, a.min_x, a.cell);          // :99
    return b;
}

// Pass 1 -- the hits.  A beam that starts inside the grid and whose end cell is inside never leaves
// it (Bresenham stays in the bounding box of its two end cells), so its walk ends on the end cell
// and the hit (:127-131) lands there; a beam that starts outside does nothing (:106-107, :127), one
// that ends outside walks off the grid and has no hit.  No walking needed: two atomics per beam.
__global__ void __launch_bounds__(256)
grid_hits_kernel(const GridArgs a)
{
    const int i = blockIdx.x;
    const double c = a.poses[4 * i], s = a.poses[4 * i + 1], px = a.poses[4 * i + 2], py = a.poses[4 * i + 3];
    const long long W = a.w, H = a.h;
    for (int64_t k = a.offsets[i] + threadIdx.x; k < a.offsets[i + 1]; k += blockDim.x) {
        const GridBeam b = grid_beam(a, c, s, px, py, k);
        const bool start_in = b.x0 >= 0 && b.x0 < W && b.y0 >= 0 && b.y0 < H;
        const bool end_in = b.x1 >= 0 && b.x1 < W && b.y1 >= 0 && b.y1 < H;
        if (start_in && end_in) {
            const long long cell = b.y1 * W + b.x1;
            atomicAdd(a.n_hit + cell, 1u);
            atomicMax(a.last_hit + cell, 2u * (uint32_t)(k + 1) + 1u);      // sorts after the beam's own miss
        }
    }
}

// Pass 2 -- the misses.  One thread per beam, one warp per 32 consecutive beams of a scan, walking
// in lockstep (:105-124); every visited cell counts a miss, and a miss that comes after the cell's
// last hit (known from pass 1, a cached read) raises the cell's `miss_after` flag.  Near the robot
// the lanes of a warp cross the same cells: for the first kAggregateSteps steps lanes that stand on
// the same cell elect a leader that issues one atomic for the group.  I = int when every beam of the
// warp spans fewer than 2^28 cells (always, in practice), long long otherwise.
constexpr int kAggregateSteps = 24;
template <typename I>
__device__ __forceinline__ void grid_walk(const GridArgs &a, bool active, const GridBeam &b, uint32_t key_miss, int lane)
{
    const I W = (I)a.w, H = (I)a.h;
    I x0 = (I)b.x0, y0 = (I)b.y0;
    const I x1 = (I)b.x1, y1 = (I)b.y1;
    const I dx = x1 > x0 ? x1 - x0 : x0 - x1, dy = y1 > y0 ? y0 - y1 : y1 - y0;             // :100-101 (dy <= 0)
    const I sx = x1 > x0 ? 1 : -1, sy = y1 > y0 ? 1 : -1;                                   // :102-103
    I error = dx + dy;
    int step = 0;
    bool walking = active;
    while (__any_sync(0xffffffffu, walking)) {
        const bool inside = walking && !(x0 < 0 || x0 >= W || y0 < 0 || y0 >= H);           // :106-107
        if (walking && !inside) walking = false;
        const int cell = inside ? (int)(y0 * W + x0) : -1 - lane;
        if (step < kAggregateSteps) {
            const unsigned grp = __match_any_sync(0xffffffffu, cell);
            if (inside && (__ffs(grp) - 1) == lane) atomicAdd(a.n_miss + cell, (uint32_t)__popc(grp));   // :109-112
        } else if (inside) {
            atomicAdd(a.n_miss + cell, 1u);
        }
        if (inside) {
            const uint32_t lh = __ldg(a.last_hit + cell);
            if (lh != 0u && key_miss > lh) a.miss_after[cell] = 1u;
            const I e2 = error * 2;                                                         // :114-124
            bool stop = false;
            if (e2 >= dy) { if (x0 == x1) stop = true; else { error += dy; x0 += sx; } }
            if (!stop && e2 <= dx) { if (y0 == y1) stop = true; else { error += dx; y0 += sy; } }
            if (stop) walking = false;
        }
        ++step;
    }
}

__global__ void __launch_bounds__(256)
grid_misses_kernel(const GridArgs a)
{
    const int i = blockIdx.x;                                   // scan
    const int lane = threadIdx.x & 31;
    const double c = a.poses[4 * i], s = a.poses[4 * i + 1], px = a.poses[4 * i + 2], py = a.poses[4 * i + 3];
    const int64_t k_begin = a.offsets[i], k_end = a.offsets[i + 1];
    for (int64_t kw = k_begin + (threadIdx.x & ~31); kw < k_end; kw += blockDim.x) {
        const int64_t k = kw + lane;
        const bool active = k < k_end;
        GridBeam b = {0, 0, 0, 0};
        if (active) b = grid_beam(a, c, s, px, py, k);
        const uint32_t key_miss = 2u * (uint32_t)(k + 1);       // order of the beam
        const long long lim = 1LL << 28;
        const bool small = llabs(b.x0) < lim && llabs(b.y0) < lim && llabs(b.x1) < lim && llabs(b.y1) < lim;
        if (__all_sync(0xffffffffu, small)) grid_walk<int>(a, active, b, key_miss, lane);
        else                                grid_walk<long long>(a, active, b, key_miss, lane);
    }
}

// The closed form per cell (see the header comment); grid is int8, updated in place.
__global__ void __launch_bounds__(256)
grid_finalize_kernel(const GridArgs a, int8_t *grid, int k_hit, int k_miss)
{
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= (int64_t)a.h * a.w) return;
    const long long n_m = a.n_miss[cell], n_h = a.n_hit[cell];
    if (n_m == 0 && n_h == 0) return;                           // no beam touched the cell
    const int g0 = grid[cell];
    const bool last_is_hit = n_h > 0 && a.miss_after[cell] == 0u;
    int g;
    if (last_is_hit) g = (n_m > 0 || g0 < 0) ? 127 : (int)min((long long)g0 + n_h * k_hit, 127LL);
    else             g = (n_h > 0 || g0 > 0) ? -128 : (int)max((long long)g0 - n_m * k_miss, -128LL);
    grid[cell] = (int8_t)g;
}

}  // namespace icpb
