// Pipe-rate calibration on B200 (developer tool): warp-instructions per cycle per SM for the
// instructions the sweep uses, and the SM clock actually sustained under that load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_pipes micro_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

constexpr int NCH = 8;   // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(256) pipes(int iters, float *out, u64 *clk)
{
    float a[NCH], b[NCH], e[NCH], f[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { a[k] = threadIdx.x * 0.001f + k; b[k] = 1.0f + k * 1e-3f; e[k] = a[k] + 2.f; f[k] = b[k] + 3.f; }
    const float c0 = 0.999f + blockIdx.x * 1e-9f, c1 = 1e-3f;
    const u64 t0 = gtime();
    const long long k0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            if (OP == 0) a[k] = __fmaf_rn(a[k], c0, c1);                       // FFMA
            if (OP == 1) a[k] = __fadd_rn(a[k], c1);                           // FADD
            if (OP == 2) a[k] = __fmul_rn(a[k], c0);                           // FMUL
            if (OP == 3) a[k] = fminf(a[k], b[k] + 0.f * it);                  // FMNMX (+ cheap dep breaker)
            if (OP == 4) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0));  // FMNMX3
            if (OP == 5) {                                                     // FFMA2 on (a[k], b[k])
                u64 v, cc, dd;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[k]), "f"(b[k]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c0));
                asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(c1));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(cc), "l"(dd));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[k]), "=f"(b[k]) : "l"(v));
            }
            if (OP == 6) {                                                     // FADD2
                u64 v, dd;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[k]), "f"(b[k]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(c1));
                asm("add.f32x2 %0, %0, %1;" : "+l"(v) : "l"(dd));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[k]), "=f"(b[k]) : "l"(v));
            }
            if (OP == 8 || OP == 9) {                                          // FFMA2 + 1 (or 2) independent scalar FFMA
                u64 v, cc, dd;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[k]), "f"(b[k]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c0));
                asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(c1));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(cc), "l"(dd));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[k]), "=f"(b[k]) : "l"(v));
                e[k] = __fmaf_rn(e[k], c0, c1);
                if (OP == 9) f[k] = __fmaf_rn(f[k], c0, c1);
            }
            if (OP >= 10 && OP <= 13) {
                // the sweep's packed mix on independent chains: dx=q-p, dy=q-p, t=dx*dx, d=dy*dy+t
                u64 q, pp, dx, dy, t;
                asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(a[k]), "f"(b[k]));
                if (OP == 10 || OP == 12 || OP == 13) asm("mov.b64 %0, {%1, %1};" : "=l"(pp) : "f"(c0));   // broadcast scalar operand
                else asm("mov.b64 %0, {%1, %2};" : "=l"(pp) : "f"(e[k]), "f"(f[k]));     // two distinct halves
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(q), "l"(pp));
                if (OP >= 12) {
                    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(q), "l"(pp));
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dx));
                    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(dx) : "l"(dy), "l"(t));
                }
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(dx));
                if (OP == 13) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(e[k]) : "f"(lo), "f"(hi)); }
                else { a[k] = lo; b[k] = hi; }
            }
            if (OP == 7) {                                                     // FFMA + FMNMX3 interleaved 2:1
                a[k] = __fmaf_rn(a[k], c0, c1);
                b[k] = __fmaf_rn(b[k], c0, c1);
                if ((k & 1) == 0) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0));
            }
        }
    }
    const long long k1 = clock64();
    const u64 t1 = gtime();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) s += a[k] + b[k] + e[k] + f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = (u64)(k1 - k0); clk[1] = t1 - t0; }
}

template <int OP>
void run(const char *name, double warp_instr_per_iter, int sms, float *out, u64 *clk)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int per_sm = 2; per_sm <= 8; per_sm *= 2) {
        const int grid = sms * per_sm;
        pipes<OP><<<grid, 256>>>(100, out, clk);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        pipes<OP><<<grid, 256>>>(iters, out, clk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        u64 h[2];
        CK(cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost));
        const double mhz = (double)h[0] / (double)h[1] * 1e3;
        const double winstr = (double)grid * 8 * iters * warp_instr_per_iter;     // 8 warps per CTA
        const double per_cyc_sm = winstr / ((double)h[0]) / sms;                   // warp-instr / cycle / SM
        printf("%-16s warps/SM=%2d: %8.3f ms, SM clock %7.1f MHz, %.3f warp-instr/cycle/SM (%.1f lanes/clk/SM)\n",
               name, per_sm * 8, ms, mhz, per_cyc_sm, per_cyc_sm * 32);
    }
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *out; u64 *clk;
    CK(cudaMalloc(&out, sms * 8 * 256 * 4)); CK(cudaMalloc(&clk, 16));
    run<0>("FFMA", NCH, sms, out, clk);
    run<1>("FADD", NCH, sms, out, clk);
    run<2>("FMUL", NCH, sms, out, clk);
    run<3>("FMNMX(+FFMA)", 2 * NCH, sms, out, clk);
    run<4>("FMNMX3", NCH, sms, out, clk);
    run<5>("FFMA2", NCH, sms, out, clk);
    run<6>("FADD2", NCH, sms, out, clk);
    run<7>("2FFMA+.5FMNMX3", 2.5 * NCH, sms, out, clk);
    run<10>("FADD2 bcast", NCH, sms, out, clk);
    run<11>("FADD2 full", NCH, sms, out, clk);
    run<12>("2FADD2+FMUL2+FFMA2", 4 * NCH, sms, out, clk);
    run<13>("..+FMNMX3", 5 * NCH, sms, out, clk);
    run<8>("FFMA2+1FFMA", 2 * NCH, sms, out, clk);
    run<9>("FFMA2+2FFMA", 3 * NCH, sms, out, clk);
    return 0;
}
