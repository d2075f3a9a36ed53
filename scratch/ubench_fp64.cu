// FP64 on B200: dependent-chain latency of DFMA / DADD and throughput with 8 independent accumulators per thread,
// one warp and eight warps per SM sub-partition (what bounds the complex128 trainers).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP>
__global__ void k(double *out, double a, double b, unsigned long long *cyc)
{
    double x = a + threadIdx.x, acc[8];
    for (int i = 0; i < 8; i++) acc[i] = b + i;
    unsigned long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        if (OP == 0) x = fma(x, a, b);
        if (OP == 1) x = x + a;
        if (OP == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
        }
        if (OP == 3) { long long v = __shfl_xor_sync(0xffffffffu, __double_as_longlong(x), 1); x = __longlong_as_double(v) + x; }
    }
    unsigned long long t1 = clock64();
    double s = x;
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, int threads, double per)
{
    double *out; unsigned long long *cyc, h;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    k<OP><<<148, threads>>>(out, 1.0000001, 0.5, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %7.2f cycles per %s\n", name, (double)h / ITERS / per, per == 1 ? "dependent op" : "warp instruction per sub-partition");
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("DFMA dependent chain (1 warp / SM)", 32, 1);
    run<1>("DADD dependent chain (1 warp / SM)", 32, 1);
    run<3>("64-bit SHFL.BFLY + DADD dependent chain", 32, 1);
    run<2>("DFMA x8 independent, 1 warp per sub-partition", 128, 8);
    run<2>("DFMA x8 independent, 8 warps per sub-partition", 1024, 8.0 * 8);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
