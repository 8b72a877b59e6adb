/*
 * CPU oracle for the ICP scan-matching path, plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  The product never links or calls it.
 *
 * Scalar float64 restatement of the reference's algorithm (reference src/icp.py:4-97); each
 * function names the lines it follows.  Parity status: pinned -- tests/test_oracle.py checks it
 * against golden outputs of the unmodified reference (tests/golden/, generator committed beside
 * them) and against the numpy restatement in oracle/icp_oracle.py.
 *
 * Compiled with -ffp-contract=off so every product and sum rounds separately, like numpy's
 * element-wise ufuncs do.
 *
 * Scans are held the way the callers hold them before homogenising: (m, 2) float64 rows
 * (reference src/dataloader.py:47-55).  The homogeneous third column the reference carries is
 * identically 1 on both clouds under every SE(2) transform, so its contribution to distances
 * and to the error is exactly 0.0 (src/icp.py:6, :52); it is not materialised here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    double epsilon;          /* src/icp.py:72 `epsilon`          */
    double stopping_thresh;  /* src/icp.py:72 `stopping_thresh`  */
    int32_t max_iters;       /* src/icp.py:72 `max_iters`        */
    int32_t rotation_only;   /* src/icp.py:72 `rotation_only`    */
} oracle_params;

/* src/icp.py:4-7 -- first index of the smallest squared distance. */
static int64_t closest_point(double px, double py, const double *pc, int64_t n)
{
    int64_t best = 0;
    double bestd = INFINITY;
    for (int64_t j = 0; j < n; ++j) {
        double dx = pc[2 * j] - px, dy = pc[2 * j + 1] - py;
        double d = dx * dx + dy * dy;     /* + 0.0 from the homogeneous column */
        if (d < bestd) { bestd = d; best = j; }
    }
    return best;
}

/* src/icp.py:22-46 with the SVD + reflection fix written in its closed form: for a 2x2
 * cross-covariance S the maximiser of tr(R S) over rotations is the angle
 * atan2(S01 - S10, S00 + S11) (identity when S = 0); SURVEY.md probe B5 measured <= 3.6e-15
 * against numpy's SVD route over 20k cases, and tests/test_oracle.py re-checks it. */
static void rigid_fit(const double *a, const double *b, int64_t n, double inc[6])
{
    double ax = 0, ay = 0, bx = 0, by = 0;
    for (int64_t i = 0; i < n; ++i) { ax += a[2 * i]; ay += a[2 * i + 1]; bx += b[2 * i]; by += b[2 * i + 1]; }
    ax /= (double)n; ay /= (double)n; bx /= (double)n; by /= (double)n;
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
    for (int64_t i = 0; i < n; ++i) {
        double x0 = a[2 * i] - ax, x1 = a[2 * i + 1] - ay;
        double y0 = b[2 * i] - bx, y1 = b[2 * i + 1] - by;
        s00 += x0 * y0; s01 += x0 * y1; s10 += x1 * y0; s11 += x1 * y1;
    }
    double th = atan2(s01 - s10, s00 + s11);
    double c = cos(th), s = sin(th);
    inc[0] = c; inc[1] = -s; inc[2] = bx - (c * ax - s * ay);
    inc[3] = s; inc[4] = c;  inc[5] = by - (s * ax + c * ay);
}

/*
 * src/icp.py:55-69 (one pass) inside src/icp.py:72-97 (loop and stop rules).
 * T is the 2x3 top of the 3x3 SE(2) matrix, row-major.  hist (optional) receives the
 * cumulative transform after every pass, corr (optional) the last pass's correspondences.
 * Returns the number of passes (= len(transforms) - 1 in the reference).
 */
int32_t icp_oracle_pair(const double *src, int64_t n1, const double *dst, int64_t n2,
                        const double init[6], const oracle_params *p,
                        double T_out[6], double *err_out, double *hist, int64_t hist_cap,
                        int32_t *corr)
{
    double T[6];
    memcpy(T, init, sizeof T);
    double *moved = (double *)malloc(sizeof(double) * 2 * (size_t)n1);
    double *match = (double *)malloc(sizeof(double) * 2 * (size_t)n1);
    int32_t passes = 0, iteration = 0;
    int have_last = 0;
    double last = 0.0, err = 0.0;
    for (;;) {
        if (p->rotation_only) { T[2] = 0.0; T[5] = 0.0; }                    /* :60-61 */
        for (int64_t i = 0; i < n1; ++i) {                                    /* :62    */
            double x = src[2 * i], y = src[2 * i + 1];
            moved[2 * i] = T[0] * x + T[1] * y + T[2];
            moved[2 * i + 1] = T[3] * x + T[4] * y + T[5];
        }
        err = 0.0;
        for (int64_t i = 0; i < n1; ++i) {                                    /* :63-64, :68 */
            int64_t j = closest_point(moved[2 * i], moved[2 * i + 1], dst, n2);
            if (corr) corr[i] = (int32_t)j;
            match[2 * i] = dst[2 * j]; match[2 * i + 1] = dst[2 * j + 1];
            double dx = moved[2 * i] - match[2 * i], dy = moved[2 * i + 1] - match[2 * i + 1];
            err += dx * dx + dy * dy;
        }
        double inc[6];
        rigid_fit(moved, match, n1, inc);                                     /* :64 */
        if (p->rotation_only) { inc[2] = 0.0; inc[5] = 0.0; }                 /* :65-66 */
        double N[6];                                                          /* :67 inc @ T */
        N[0] = inc[0] * T[0] + inc[1] * T[3];
        N[1] = inc[0] * T[1] + inc[1] * T[4];
        N[2] = inc[0] * T[2] + inc[1] * T[5] + inc[2];
        N[3] = inc[3] * T[0] + inc[4] * T[3];
        N[4] = inc[3] * T[1] + inc[4] * T[4];
        N[5] = inc[3] * T[2] + inc[4] * T[5] + inc[5];
        memcpy(T, N, sizeof T);
        if (hist && passes < hist_cap) memcpy(hist + 6 * (size_t)passes, T, sizeof T);
        ++passes;                                                             /* :84 */
        if (err < p->epsilon) break;                                          /* :86 */
        if (iteration > p->max_iters) break;                                  /* :88 */
        if (have_last && fabs(last - err) < p->stopping_thresh) break;        /* :91-95 */
        last = err; have_last = 1;
        ++iteration;                                                          /* :97 */
    }
    memcpy(T_out, T, sizeof T);
    *err_out = err;
    free(moved); free(match);
    return passes;
}

/*
 * The reference's only fan-out: one independent icp() per scan pair on a process pool
 * (scripts/main.py:240-247).  Here: worker threads pop pair ids from an atomic counter over a
 * CSR scan table.  pairs[b] = (src scan, dst scan); init is B x 6 or NULL for identity.
 */
typedef struct {
    const double *xy; const int64_t *offsets; const int32_t *pairs; const double *init;
    int64_t B; const oracle_params *p; double *T_out; double *err_out; int32_t *passes_out;
    atomic_llong next;
} batch_job;

static void *batch_worker(void *arg)
{
    static const double ident[6] = {1, 0, 0, 0, 1, 0};
    batch_job *j = (batch_job *)arg;
    for (;;) {
        int64_t b = atomic_fetch_add(&j->next, 1);
        if (b >= j->B) break;
        int32_t s = j->pairs[2 * b], d = j->pairs[2 * b + 1];
        j->passes_out[b] = icp_oracle_pair(j->xy + 2 * j->offsets[s], j->offsets[s + 1] - j->offsets[s],
                                           j->xy + 2 * j->offsets[d], j->offsets[d + 1] - j->offsets[d],
                                           j->init ? j->init + 6 * b : ident, j->p,
                                           j->T_out + 6 * b, j->err_out + b, NULL, 0, NULL);
    }
    return NULL;
}

int icp_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int icp_oracle_batch(const double *xy, const int64_t *offsets, const int32_t *pairs,
                     const double *init, int64_t B, const oracle_params *p, int n_threads,
                     double *T_out, double *err_out, int32_t *passes_out)
{
    batch_job job = {xy, offsets, pairs, init, B, p, T_out, err_out, passes_out, 0};
    if (n_threads <= 0) n_threads = icp_oracle_max_threads();
    if (n_threads > B) n_threads = (int)(B > 0 ? B : 1);
    if (n_threads > 1024) n_threads = 1024;
    pthread_t tid[1024];
    int started = 0;
    for (int t = 1; t < n_threads; ++t)
        if (pthread_create(&tid[started], NULL, batch_worker, &job) == 0) ++started;
    batch_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    return 0;
}
