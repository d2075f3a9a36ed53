// Instruction-throughput microbenchmarks (B200): cycles per warp-instruction per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define NACC 8
template <int OP>
__global__ void k(float *out, float a, float b, unsigned long long *cyc)
{
    __shared__ float2 tab[64];
    __shared__ float ring[32 * 64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = make_float2(a * i, b * i);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) ring[i] = a * i;
    __syncthreads();
    float x[NACC], y[NACC];
#pragma unroll
    for (int j = 0; j < NACC; j++) { x[j] = a + j + threadIdx.x; y[j] = b - j; }
    unsigned u = threadIdx.x;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (OP == 0) x[j] = fmaf(x[j], a, b);                       // FFMA
            if (OP == 1) x[j] = __fadd_rn(x[j], a);                     // FADD
            if (OP == 2) x[j] = __fmul_rn(x[j], a);                     // FMUL
            if (OP == 3) x[j] = fminf(fabsf(x[j]), fabsf(y[j]));        // FMNMX
            if (OP == 4) {                                              // FMUL2
                asm volatile("{.reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mul.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra;}"
                             : "+f"(x[j]), "+f"(y[j]) : "f"(a), "f"(b));
            }
            if (OP == 5) {                                              // FADD2
                asm volatile("{.reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra;}"
                             : "+f"(x[j]), "+f"(y[j]) : "f"(a), "f"(b));
            }
            if (OP == 6) {                                              // FFMA2
                asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mov.b64 rc, {%3,%2}; fma.rn.f32x2 ra, ra, rb, rc; mov.b64 {%0,%1}, ra;}"
                             : "+f"(x[j]), "+f"(y[j]) : "f"(a), "f"(b));
            }
            if (OP == 7) { asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(a), "f"(b)); }  // FFMA.SAT
            if (OP == 8) { u = u * 8u + (unsigned)j; x[j] += __uint_as_float(u); }   // LEA/IMAD + FADD
            if (OP == 9) {                                              // LDS.64 table lookup (dependent index)
                const float2 l = tab[__float_as_uint(x[j]) & 7u];
                x[j] = l.x + l.y;
            }
            if (OP == 10) {                                             // CREDUX.MIN + broadcast back
                x[j] = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(x[j]) + j));
            }
            if (OP == 11) {                                             // VOTE.ballot on compare
                u += __ballot_sync(0xffffffffu, x[j] == y[j]); x[j] = __uint_as_float(u);
            }
            if (OP == 12) {                                             // FFMA + FMNMX mix 1:1 (two pipes)
                x[j] = fmaf(x[j], a, b); y[j] = fminf(fabsf(y[j]), x[j]);
            }
            if (OP == 13) {                                             // ring: LDS + STS same address, conflict-free
                const int ad = ((it * NACC + j) & 31) * 64 + (threadIdx.x & 31);
                const float o = ring[ad]; ring[ad] = x[j]; x[j] = __fadd_rn(x[j], o);
            }
            if (OP == 14) {                                             // SHFL idx
                x[j] = __shfl_sync(0xffffffffu, x[j], (threadIdx.x + j) & 31);
            }
            if (OP == 15) {                                             // FFMA 3 distinct regs + FADD2 alternating
                x[j] = fmaf(x[j], y[j], x[(j + 1) % NACC]);
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += x[j] + y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char *name, int ninstr)
{
    float *out; unsigned long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    printf("%-28s", name);
    for (int nw = 1; nw <= 16; nw *= 2) {   // warps per SMSP
        k<OP><<<148, 128 * nw>>>(out, 1.0001f, 0.5f, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (nw * 128 > 1024) break;
        printf("  w%d: %6.2f", nw, (double)h / ((double)ITERS * NACC * ninstr * nw));
    }
    printf("   (cyc per warp-instr per SMSP)\n");
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("FFMA", 1); run<1>("FADD", 1); run<2>("FMUL", 1); run<3>("FMNMX |a|,|b|", 1);
    run<4>("FMUL2", 1); run<5>("FADD2", 1); run<6>("FFMA2", 1); run<7>("FFMA.SAT", 1);
    run<8>("LEA/IMAD+FADD (2)", 2); run<9>("LOP+LDS.64 tab+FADD (3)", 3); run<10>("IADD+CREDUX+MOV (3)", 3);
    run<11>("FSETP+VOTE+IADD (3)", 3); run<12>("FFMA+FMNMX (2)", 2); run<13>("ring LDS+STS+FADD (3+addr)", 3);
    run<14>("SHFL idx", 1); run<15>("FFMA 3reg", 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
