// Pipe-rate calibration on B200 (developer tool): warp-instructions per cycle per SM for the
// instruction FORMS the sweep uses.  Round-2 rewrite: every rate is (instructions the whole grid
// executes, counted from the loop structure and checked against the SASS with cuobjdump) divided by
// (CUDA-event time x the SM clock measured in the same launch); values stay packed in 64-bit
// registers across iterations, so no MOV sits between the measured instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_pipes micro_pipes.cu
//   cuobjdump -sass micro_pipes | grep -c FFMA2     (static check of the unrolled bodies)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

constexpr int NCH = 8;      // independent chains per thread
constexpr int UNR = 4;      // loop body repeats

// OP: 0 FFMA a=a*b+c (three distinct registers)        1 FFMA a=a*a+c (repeated source)
//     2 FFMA a=a*imm+c                                  3 FADD a=a+c             4 FMUL a=a*a
//     5 FFMA2 v=v*w+z (three distinct pairs)            6 FFMA2 v=x*x+v (the sweep's form: repeated source + accumulator)
//     7 FADD2 v=v-bcast (the sweep's form: pair minus one broadcast scalar)    8 FADD2 v=v+w (two pairs)
//     9 FMUL2 v=x*x (the sweep's form)                 10 FMNMX3     11 FMNMX
//    12 the sweep's mix per target pair: FADD2 bcast, FADD2 bcast, FMUL2 x*x, FFMA2 y*y+t
//    13 the same + one FMNMX3 per pair (= the sweep's inner loop without loads)
//    14 FFMA2 + FMNMX3 on independent chains (1:1)     15 FFMA + FMNMX3 (1:1)      16 FFMA + LOP3 (1:1)
//    17 FFMA + IMAD (1:1)       18 FFMA2 + DFMA (1:1)      19 DFMA      20 FMNMX3 + LOP3 (1:1, both ALU pipe)
//    21 2 FFMA + FMNMX3 (2:1)   22 LOP3   23 FFMA2 + 2 FFMA
template <int OP>
__global__ void __launch_bounds__(256) pipes(int iters, float *out, u64 *clk)
{
    float a[NCH], b[NCH];
    u64 v[NCH], w[NCH];
    unsigned ia[NCH], ib[NCH];
    double da[NCH], db = 0.999 + threadIdx.x * 1e-12, dc = 1e-3 + blockIdx.x * 1e-12;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        a[k] = threadIdx.x * 0.001f + k; b[k] = 1.0f + k * 1e-3f + threadIdx.x * 1e-6f;
        v[k] = pack2(a[k], b[k]); w[k] = pack2(b[k], a[k] * 0.5f);
        ia[k] = threadIdx.x * 2654435761u + k; ib[k] = blockIdx.x * 40503u + 3 * k + 1; da[k] = a[k];
    }
    float c0 = 0.999f + blockIdx.x * 1e-9f, c1 = 1e-3f + threadIdx.x * 1e-9f, m[NCH];
    u64 z = pack2(c1, c0);
    // opaque to the optimiser: nothing below is rematerialised or folded
#pragma unroll
    for (int k = 0; k < NCH; ++k) { m[k] = 1e30f; asm volatile("" : "+l"(w[k]), "+l"(v[k]), "+f"(a[k]), "+f"(b[k]), "+f"(m[k])); }
    asm volatile("" : "+l"(z), "+f"(c0), "+f"(c1));
    const u64 t0 = gtime();
    const long long k0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c0), "f"(c1));
                if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(a[k]) : "f"(c1));
                if (OP == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %1;" : "+f"(a[k]) : "f"(c1));
                if (OP == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(c1));
                if (OP == 4) asm volatile("mul.rn.f32 %0, %0, %0;" : "+f"(a[k]));
                if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(w[k]), "l"(z));
                if (OP == 6) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(v[k]) : "l"(w[k]));
                if (OP == 7) { u64 bc; asm("mov.b64 %0, {%1, %1};" : "=l"(bc) : "f"(c1));
                               asm volatile("sub.rn.f32x2 %0, %0, %1;" : "+l"(v[k]) : "l"(bc)); }
                if (OP == 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[k]) : "l"(w[k]));
                if (OP == 9) asm volatile("mul.rn.f32x2 %0, %0, %0;" : "+l"(v[k]));
                if (OP == 10) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0));
                if (OP == 11) { asm volatile("min.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b[k])); asm volatile("" : "+f"(a[k])); }
                if (OP == 14) { asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(v[k]) : "l"(w[k]));
                                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0)); }
                if (OP == 15) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(c0), "f"(c1));
                                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0)); }
                if (OP == 16) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(c0), "f"(c1));
                                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[k]) : "r"(ib[k]), "r"(ib[(k + 1) % NCH])); }
                if (OP == 17) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(c0), "f"(c1));
                                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ia[k]) : "r"(ib[k]), "r"(ib[(k + 1) % NCH])); }
                if (OP == 18) { asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(v[k]) : "l"(w[k]));
                                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(da[k]) : "d"(db), "d"(dc)); }
                if (OP == 19) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(da[k]) : "d"(db), "d"(dc));
                if (OP == 20) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c0));
                                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[k]) : "r"(ib[k]), "r"(ib[(k + 1) % NCH])); }
                if (OP == 21) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(c0), "f"(c1));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[k]) : "f"(c0), "f"(c1));
                                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c1), "f"(c0)); }
                if (OP == 22) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[k]) : "r"(ib[k]), "r"(ib[(k + 1) % NCH]));
                if (OP == 23) { asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(v[k]) : "l"(w[k]));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(c0), "f"(c1));
                                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(c0), "f"(c1)); }
                if (OP == 12 || OP == 13) {
                    u64 bx, by, dx, dy, t, d;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(bx) : "f"(c0));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(by) : "f"(c1));
                    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(v[k]), "l"(bx));
                    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(w[k]), "l"(by));
                    asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dx));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dy), "l"(t));
                    if (OP == 13) {
                        float lo, hi;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
                        asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[k]) : "f"(lo), "f"(hi));
                    }
                    v[k] = d; w[k] = d;                     // the next step depends on this one
                }
            }
        }
    }
    const long long k1 = clock64();
    const u64 t1 = gtime();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) { s += m[k]; float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[k])); s += a[k] + lo + hi + (float)ia[k] + (float)da[k] + b[k]; }

    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = (u64)(k1 - k0); clk[1] = t1 - t0; }
}

template <int OP>
void run(const char *name, double instr_per_chain_step, double fma_pipe_instr, int sms, float *out, u64 *clk)
{
    const int iters = 4000;
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
        const double mhz = (double)h[0] / (double)h[1] * 1e3;                    // SM clock during the launch
        const double cycles = ms * 1e-3 * mhz * 1e6;                              // elapsed SM cycles (event time)
        const double per_sm_steps = (double)per_sm * 8 * iters * UNR * NCH;       // chain steps per SM (8 warps per CTA)
        const double all = per_sm_steps * instr_per_chain_step / cycles, fma = per_sm_steps * fma_pipe_instr / cycles;
        printf("%-34s warps/SM=%2d: %8.3f ms, %7.1f MHz, %.3f warp-instr/clk/SM, FMA-pipe instr %.3f /clk/SM\n",
               name, per_sm * 8, ms, mhz, all, fma);
    }
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *out; u64 *clk;
    CK(cudaMalloc(&out, sms * 8 * 256 * 4)); CK(cudaMalloc(&clk, 16));
    run<0>("FFMA a=a*b+c (3 regs)", 1, 1, sms, out, clk);
    run<1>("FFMA a=a*a+c (repeated src)", 1, 1, sms, out, clk);
    run<2>("FFMA a=a*imm+c", 1, 1, sms, out, clk);
    run<3>("FADD a=a+c", 1, 1, sms, out, clk);
    run<4>("FMUL a=a*a", 1, 1, sms, out, clk);
    run<5>("FFMA2 v=v*w+z (3 pairs)", 1, 1, sms, out, clk);
    run<6>("FFMA2 v=x*x+v (sweep form)", 1, 1, sms, out, clk);
    run<7>("FADD2 v=v-bcast (sweep form)", 1, 1, sms, out, clk);
    run<8>("FADD2 v=v+w (2 pairs)", 1, 1, sms, out, clk);
    run<9>("FMUL2 v=x*x (sweep form)", 1, 1, sms, out, clk);
    run<10>("FMNMX3", 1, 0, sms, out, clk);
    run<11>("FMNMX x2 (ptxas fuses into FMNMX3)", 0.5, 0, sms, out, clk);
    run<12>("sweep mix 2FADD2+FMUL2+FFMA2", 4, 4, sms, out, clk);
    run<13>("sweep mix + FMNMX3", 5, 4, sms, out, clk);
    run<14>("FFMA2 + FMNMX3 (1:1)", 2, 1, sms, out, clk);
    run<15>("FFMA + FMNMX3 (1:1)", 2, 1, sms, out, clk);
    run<21>("2 FFMA + FMNMX3 (2:1)", 3, 2, sms, out, clk);
    run<22>("LOP3", 1, 0, sms, out, clk);
    run<16>("FFMA + LOP3 (1:1)", 2, 1, sms, out, clk);
    run<20>("FMNMX3 + LOP3 (1:1)", 2, 0, sms, out, clk);
    run<17>("FFMA + IMAD (1:1)", 2, 2, sms, out, clk);
    run<19>("DFMA", 1, 0, sms, out, clk);
    run<18>("FFMA2 + DFMA (1:1)", 2, 1, sms, out, clk);
    run<23>("FFMA2 + 2 FFMA", 3, 3, sms, out, clk);
    return 0;
}
