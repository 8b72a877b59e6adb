/*
 * Sequential restatement of the reference's occupancy-grid update -- TEST INFRASTRUCTURE, NOT
 * PRODUCT CODE (only tests/ and bench tools may call it, as the checker).
 *
 * Follows the reference's src/produce_occupancy_grid.py:
 *   construct_global_points  :84-94    global = odom_change_to_mat(pose) @ [x, y, 1]
 *   bresenham_update         :96-131   the beam's cells get a "miss", its last cell a "hit"
 *   global_position_to_grid_cell :133-138
 * including the int8 arithmetic of the saturation tests (:109, :128): `-128 - grid[y, x]` and
 * `127 - grid[y, x]` are evaluated in int8 and wrap, so a miss on a positive cell sets it to -128
 * and a hit on a negative cell sets it to 127.  Beams are applied one after the other in (scan,
 * beam) order, exactly like the reference's loops (:54-56, :76-78).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* numpy's (3,3) @ (3,1) product as this image's OpenBLAS evaluates it, found by matching bits
 * against numpy (tests/golden/make_grid_golden.py): fma(m00, x, m01*y) + m02. */
void grid_oracle_global_points(const double *poses, const double *xy, const int64_t *off, int64_t n,
                               double *gxy)
{
    for (int64_t i = 0; i < n; ++i) {
        const double c = cos(poses[3 * i + 2]), s = sin(poses[3 * i + 2]);
        for (int64_t k = off[i]; k < off[i + 1]; ++k) {
            const double x = xy[2 * k], y = xy[2 * k + 1];
            gxy[2 * k] = fma(c, x, -s * y) + poses[3 * i];
            gxy[2 * k + 1] = fma(s, x, c * y) + poses[3 * i + 1];
        }
    }
}

static int8_t wrap8(int v) { return (int8_t)(uint8_t)(v & 0xff); }

static void beam(int8_t *grid, int64_t h, int64_t w, double px, double py, double qx, double qy,
                 double min_x, double min_y, double cell, int k_hit, int k_miss)
{
    int64_t x0 = (int64_t)floor((px - min_x) / cell), y0 = (int64_t)floor((py - min_y) / cell);
    const int64_t x1 = (int64_t)floor((qx - min_x) / cell), y1 = (int64_t)floor((qy - min_y) / cell);
    const int64_t dx = llabs(x1 - x0), dy = -llabs(y1 - y0);
    const int64_t sx = x1 > x0 ? 1 : -1, sy = y1 > y0 ? 1 : -1;
    int64_t error = dx + dy;
    for (;;) {
        if (x0 < 0 || x0 >= w || y0 < 0 || y0 >= h) break;
        int8_t *g = grid + y0 * w + x0;
        if (wrap8(-128 - *g) < -k_miss) *g = wrap8(*g - k_miss); else *g = -128;       /* :109-112 */
        const int64_t e2 = error * 2;
        if (e2 >= dy) { if (x0 == x1) break; error += dy; x0 += sx; }
        if (e2 <= dx) { if (y0 == y1) break; error += dx; y0 += sy; }
    }
    if (x0 >= 0 && x0 < w && y0 >= 0 && y0 < h) {
        int8_t *g = grid + y0 * w + x0;
        if (wrap8(127 - *g) > k_hit) *g = wrap8(*g + k_hit); else *g = 127;             /* :128-131 */
    }
}

/* The double loop of produce_occupancy_grid (:54-56) / update_occupancy_grid (:76-78). */
void grid_oracle_update(int8_t *grid, int64_t h, int64_t w, const double *poses, const double *gxy,
                        const int64_t *off, int64_t n, double min_x, double min_y, double cell,
                        int k_hit, int k_miss)
{
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = off[i]; k < off[i + 1]; ++k)
            beam(grid, h, w, poses[3 * i], poses[3 * i + 1], gxy[2 * k], gxy[2 * k + 1], min_x, min_y, cell,
                 k_hit, k_miss);
}
