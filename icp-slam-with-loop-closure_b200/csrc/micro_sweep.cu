// Micro-benchmark of candidate inner loops for the nearest-neighbour sweep (developer tool, not
// part of libicpb.so).  Measures point-pair distance evaluations (PDE) per second for
//   mode 0: difference form, scalar   (FADD, FADD, FMUL, FFMA per PDE)
//   mode 1: difference form, packed   (sub/sub/mul/fma .f32x2 per 2 PDE)
//   mode 2: expanded form, scalar     (FFMA, FFMA per PDE on precomputed |q|^2)
//   mode 3: expanded form, packed     (fma/fma .f32x2 per 2 PDE)
// with R source points per thread, including the min tree and the chunk bookkeeping.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o micro_sweep micro_sweep.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

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
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int MODE, int R, int CH, int VAR = 0>
__global__ void __launch_bounds__(256) sweep(const float *gx, const float *gy, int n2, int reps, float *out)
{
    extern __shared__ __align__(16) float sm[];
    float *tqx = sm, *tqy = sm + n2, *tqq = sm + 2 * n2;
    for (int j = threadIdx.x; j < n2; j += blockDim.x) {
        tqx[j] = gx[j]; tqy[j] = gy[j]; tqq[j] = gx[j] * gx[j] + gy[j] * gy[j];
    }
    __syncthreads();
    float px[R], py[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        px[r] = 0.01f * (threadIdx.x * R + r) + blockIdx.x * 1e-3f;
        py[r] = 0.02f * (threadIdx.x * R + r) - 3.0f;
    }
    float acc = 0.f;
    const int nchunks = n2 / CH;
    for (int rep = 0; rep < reps; ++rep) {
        float m1[R], m2[R]; int c1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { m1[r] = 3e38f; m2[r] = 3e38f; c1[r] = 0; }
        const float4 *qx4 = reinterpret_cast<const float4 *>(tqx);
        const float4 *qy4 = reinterpret_cast<const float4 *>(tqy);
        const float4 *qq4 = reinterpret_cast<const float4 *>(tqq);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            float4 X[CH / 4], Y[CH / 4], Q[CH / 4];
#pragma unroll
            for (int v = 0; v < CH / 4; ++v) {
                X[v] = qx4[(CH / 4) * c + v]; Y[v] = qy4[(CH / 4) * c + v];
                if (MODE >= 2) Q[v] = qq4[(CH / 4) * c + v];
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float d[CH];
                if (MODE == 0) {
#pragma unroll
                    for (int v = 0; v < CH / 4; ++v) {
                        const float xs[4] = {X[v].x, X[v].y, X[v].z, X[v].w};
                        const float ys[4] = {Y[v].x, Y[v].y, Y[v].z, Y[v].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float dx = __fsub_rn(xs[k], px[r]), dy = __fsub_rn(ys[k], py[r]);
                            d[4 * v + k] = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                        }
                    }
                } else if (MODE == 1) {
                    const u64 PX = pack2(px[r], px[r]), PY = pack2(py[r], py[r]);
#pragma unroll
                    for (int v = 0; v < CH / 4; ++v) {
                        const u64 xa = pack2(X[v].x, X[v].y), xb = pack2(X[v].z, X[v].w);
                        const u64 ya = pack2(Y[v].x, Y[v].y), yb = pack2(Y[v].z, Y[v].w);
                        const u64 dxa = sub2(xa, PX), dya = sub2(ya, PY);
                        const u64 dxb = sub2(xb, PX), dyb = sub2(yb, PY);
                        const u64 da = fma2(dya, dya, mul2(dxa, dxa));
                        const u64 db = fma2(dyb, dyb, mul2(dxb, dxb));
                        unpack2(da, d[4 * v + 0], d[4 * v + 1]);
                        unpack2(db, d[4 * v + 2], d[4 * v + 3]);
                    }
                } else if (MODE == 2) {
                    const float ax = -2.f * px[r], ay = -2.f * py[r];
#pragma unroll
                    for (int v = 0; v < CH / 4; ++v) {
                        d[4 * v + 0] = __fmaf_rn(ay, Y[v].x, __fmaf_rn(ax, X[v].x, Q[v].x));
                        d[4 * v + 1] = __fmaf_rn(ay, Y[v].y, __fmaf_rn(ax, X[v].y, Q[v].y));
                        d[4 * v + 2] = __fmaf_rn(ay, Y[v].z, __fmaf_rn(ax, X[v].z, Q[v].z));
                        d[4 * v + 3] = __fmaf_rn(ay, Y[v].w, __fmaf_rn(ax, X[v].w, Q[v].w));
                    }
                } else {
                    const u64 AX = pack2(-2.f * px[r], -2.f * px[r]), AY = pack2(-2.f * py[r], -2.f * py[r]);
#pragma unroll
                    for (int v = 0; v < CH / 4; ++v) {
                        const u64 xa = pack2(X[v].x, X[v].y), xb = pack2(X[v].z, X[v].w);
                        const u64 ya = pack2(Y[v].x, Y[v].y), yb = pack2(Y[v].z, Y[v].w);
                        const u64 qa = pack2(Q[v].x, Q[v].y), qb = pack2(Q[v].z, Q[v].w);
                        const u64 da = fma2(AY, ya, fma2(AX, xa, qa));
                        const u64 db = fma2(AY, yb, fma2(AX, xb, qb));
                        unpack2(da, d[4 * v + 0], d[4 * v + 1]);
                        unpack2(db, d[4 * v + 2], d[4 * v + 3]);
                    }
                }
                if (VAR == 2) {          // no min at all: fold the distances with packed adds
                    float sacc = 0.f;
#pragma unroll
                    for (int k = 0; k < CH; k += 4) sacc += (d[k] + d[k + 1]) + (d[k + 2] + d[k + 3]);
                    m1[r] += sacc;
                } else if (VAR == 3) {   // balanced min tree instead of a chain
                    float t[CH / 2];
#pragma unroll
                    for (int k = 0; k < CH / 2; ++k) t[k] = fminf(d[2 * k], d[2 * k + 1]);
                    float cm = min3f(min3f(t[0], t[1], t[2]), min3f(t[3], t[4], t[5]), fminf(t[6], t[7]));
                    const bool better = cm < m1[r];
                    m2[r] = fminf(m2[r], better ? m1[r] : cm);
                    m1[r] = fminf(m1[r], cm);
                    c1[r] = better ? c : c1[r];
                } else {
                    float cm;
                    if (VAR == 4) {      // two independent min3 chains
                        float ca = min3f(d[0], d[1], d[2]), cb2 = min3f(d[3], d[4], d[5]);
                        ca = min3f(ca, d[6], d[7]); cb2 = min3f(cb2, d[8], d[9]);
                        ca = min3f(ca, d[10], d[11]); cb2 = min3f(cb2, d[12], d[13]);
                        cm = min3f(ca, cb2, fminf(d[14], d[15]));
                    } else {
                        cm = d[0];
#pragma unroll
                        for (int k = 1; k + 1 < CH; k += 2) cm = min3f(cm, d[k], d[k + 1]);
                        cm = fminf(cm, d[CH - 1]);
                    }
                    if (VAR == 1) {      // min only, no bookkeeping
                        m1[r] = fminf(m1[r], cm);
                    } else {
                        const bool better = cm < m1[r];
                        m2[r] = fminf(m2[r], better ? m1[r] : cm);
                        m1[r] = fminf(m1[r], cm);
                        c1[r] = better ? c : c1[r];
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { acc += m1[r] + m2[r] + (float)c1[r]; px[r] += 1e-4f; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int R, int CH, int VAR = 0>
void run(const char *name, const float *gx, const float *gy, int n2, int sms, float *out, int threads)
{
    const int smem = 3 * n2 * sizeof(float);
    CK(cudaFuncSetAttribute(sweep<MODE, R, CH, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep<MODE, R, CH, VAR>, threads, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, sweep<MODE, R, CH, VAR>));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int per_sm = 1; per_sm <= occ && per_sm <= 8; per_sm *= 2) {
        const int grid = sms * per_sm;
        const int reps = 64;
        sweep<MODE, R, CH, VAR><<<grid, threads, smem>>>(gx, gy, n2, 4, out);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int t = 0; t < 3; ++t) {
            CK(cudaEventRecord(e0));
            sweep<MODE, R, CH, VAR><<<grid, threads, smem>>>(gx, gy, n2, reps, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double pde = (double)grid * threads * R * (double)n2 * reps;
        printf("%-22s R=%d CH=%2d thr=%3d regs=%3d ctas/sm=%d (occ %d): %7.3f ms  %.3f TPDE/s\n", name, R, CH,
               threads, fa.numRegs, per_sm, occ, best, pde / best * 1e-9);
    }
}

int main(int argc, char **argv)
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, clock %d kHz, scalar-slot roofline %.3f TPDE/s\n", sms, clk, sms * 128.0 * clk * 1e3 / 4 * 1e-12);
    const int n2 = 1024;
    std::vector<float> hx(n2), hy(n2);
    for (int j = 0; j < n2; ++j) { hx[j] = 5.f * cosf(j * 0.00613f); hy[j] = 5.f * sinf(j * 0.00613f); }
    float *gx, *gy, *out;
    CK(cudaMalloc(&gx, n2 * 4)); CK(cudaMalloc(&gy, n2 * 4)); CK(cudaMalloc(&out, 148 * 64 * 1024 * 4));
    CK(cudaMemcpy(gx, hx.data(), n2 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gy, hy.data(), n2 * 4, cudaMemcpyHostToDevice));
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    if (only < 0 || only == 0) run<0, 4, 16>("diff scalar", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 1) run<0, 8, 16>("diff scalar", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 2) run<0, 4, 8>("diff scalar", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 3) run<0, 2, 16>("diff scalar", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 4) run<1, 4, 16>("diff packed f32x2", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 5) run<1, 8, 16>("diff packed f32x2", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 6) run<1, 4, 8>("diff packed f32x2", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 7) run<1, 2, 16>("diff packed f32x2", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 8) run<2, 4, 16>("expanded scalar", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 9) run<2, 8, 16>("expanded scalar", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 10) run<3, 4, 16>("expanded packed f32x2", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 11) run<3, 8, 16>("expanded packed f32x2", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 12) run<3, 8, 32>("expanded packed f32x2", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 20) run<1, 2, 16, 0>("diffp R2 base", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 21) run<1, 2, 16, 1>("diffp R2 min only", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 22) run<1, 2, 16, 2>("diffp R2 no min", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 23) run<1, 2, 16, 3>("diffp R2 tree", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 24) run<1, 2, 16, 4>("diffp R2 2chains", gx, gy, n2, sms, out, 128);
    if (only < 0 || only == 25) run<1, 4, 16, 4>("diffp R4 2chains", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 26) run<1, 4, 16, 1>("diffp R4 min only", gx, gy, n2, sms, out, 256);
    if (only < 0 || only == 27) run<1, 4, 16, 2>("diffp R4 no min", gx, gy, n2, sms, out, 256);
    return 0;
}
