// Dependent-chain latency of a few sm_100a instructions, one warp on one SM (developer tool):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/latency tools/micro/latency.cu && gpurun_out/latency
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP>
__global__ void chain(double x0, double y0, int *idx, long long *out, double *sink)
{
    __shared__ int s_next[256];
    __shared__ double s_d[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_next[i] = idx[i]; s_d[i] = x0 + i; }
    __syncthreads();
    double x = x0, y = y0;
    float f = (float)x0, g = (float)y0;
    int k = threadIdx.x & 255;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, y);
        if (OP == 1) x = __dadd_rn(x, y);
        if (OP == 2) x = __dmul_rn(x, y);
        if (OP == 3) f = fmaf(f, g, g);
        if (OP == 4) k = s_next[k];
        if (OP == 6) x = (double)(int)x + y;                      // F2I + I2F + DADD
        if (OP == 7) x = floor(x) + y;
        if (OP == 8) x = s_d[((int)__double2hiint(x)) & 255] + y;   // LDS.64 + DADD (address from the value)
        if (OP == 9) x = x > y ? x - y : x + y;                  // DSETP + select-ish
        if (OP == 10) { double s, c; sincos(x, &s, &c); x = s + c; }
        if (OP == 11) x = 1.0 / x + y;
        if (OP == 12) x = rsqrt(x) + y;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x + f + k;
}

template <int OP>
void run(const char *name, int *idx, long long *out, double *sink)
{
    chain<OP><<<1, 32>>>(1.0000001, 0.9999999, idx, out, sink);
    chain<OP><<<1, 32>>>(1.0000001, 0.9999999, idx, out, sink);
    long long c;
    cudaMemcpy(&c, out, sizeof c, cudaMemcpyDeviceToHost);
    printf("%-28s %7.1f cycles per dependent step\n", name, (double)c / N);
}

int main()
{
    int h[256];
    for (int i = 0; i < 256; ++i) h[i] = (i * 37 + 11) & 255;
    int *idx; long long *out; double *sink;
    cudaMalloc(&idx, sizeof h); cudaMalloc(&out, 8); cudaMalloc(&sink, 8 * 32);
    cudaMemcpy(idx, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>("DFMA", idx, out, sink);
    run<1>("DADD", idx, out, sink);
    run<2>("DMUL", idx, out, sink);
    run<3>("FFMA", idx, out, sink);
    run<4>("LDS.32 pointer chase", idx, out, sink);
    run<6>("F2I.F64 + I2F.F64 + DADD", idx, out, sink);
    run<7>("floor(double) + DADD", idx, out, sink);
    run<8>("LDS.64 + DADD", idx, out, sink);
    run<9>("DSETP + 2 DADD + select", idx, out, sink);
    run<10>("sincos(double) + DADD", idx, out, sink);
    run<11>("1.0 / x + DADD", idx, out, sink);
    run<12>("rsqrt(double) + DADD", idx, out, sink);
    return cudaDeviceSynchronize() != cudaSuccess;
}
