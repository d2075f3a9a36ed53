// Dependent-chain latencies (B200), one warp per SM: cycles per dependent instruction.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float *out, float a, float b, unsigned long long *cyc)
{
    float x = a + threadIdx.x, y = b;
    unsigned long long P = 0;
    asm("mov.b64 %0, {%1,%2};" : "=l"(P) : "f"(x), "f"(y));
    unsigned long long Q = P;
    unsigned long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < ITERS; it++) {
        if (OP == 0) x = fmaf(x, a, b);
        if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(P) : "l"(Q));
        if (OP == 2) x = __shfl_xor_sync(0xffffffffu, x, 1);
        if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + x;
        if (OP == 4) x = fminf(x, y) ;
        if (OP == 5) x = __fadd_rn(x, a);
        if (OP == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(P) : "l"(Q));
        if (OP == 7) { x = (x > y) ? a : x * b; }
        if (OP == 8) { asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(P) : "l"(Q)); }   // accumulate-only dependence
        if (OP == 9) { x = fmaf(a, b, x); }                                                     // accumulate-only dependence
    }
    unsigned long long t1 = clock64();
    float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(P));
    out[threadIdx.x] = x + r.x + r.y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name)
{
    float *out; unsigned long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    k<OP><<<1, 32>>>(out, 1.0001f, 0.5f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %6.2f cycles per dependent op\n", name, (double)h / ITERS);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("FFMA (dep on multiplicand)"); run<9>("FFMA (dep on addend)"); run<1>("FFMA2 (dep on multiplicand)");
    run<8>("FFMA2 (dep on addend)"); run<5>("FADD"); run<6>("FADD2"); run<4>("FMNMX"); run<2>("SHFL.BFLY");
    run<3>("SHFL.BFLY + FADD"); run<7>("FSETP+FMUL+FSEL");
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
