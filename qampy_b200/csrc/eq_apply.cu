// eq_apply: static MIMO FIR + decimation (replaces apply_filter_to_signal,
// qampy/core/equalisation/pythran_equalisation.py:37-76).
//
//   out[s, j, i] = sum_k sum_t E[s, k, i*os + t] * wx[s, modes[j], k, t]        (no conjugate, :24-31)
//   N = (L - ntaps + 1) / os                                                     (:69)
//
// Two kernels:
//   * apply_generic_kernel: any nmodes / os / ntaps / dtype.  One CTA = one (segment, output mode,
//     tile of outputs).  The input window of the tile is staged to shared memory with TMA bulk
//     copies (cp.async.bulk + mbarrier) when the rows are 16-byte aligned, plain loads otherwise.
//   * apply_2x2_os2_kernel<float>: the north-star case (dual-pol in, both modes out, 2 samples per
//     symbol, complex64).  Register-tiled: every thread owns R consecutive output symbols of both
//     modes and slides a register window over 128-bit shared-memory loads, so the kernel is bound
//     by FP32 FMA issue (30 flop/byte at ntaps = 45) rather than by shared-memory bandwidth.
#include "qb_common.cuh"

namespace qb {

template <typename T>
struct ApplyParams {
    const cx<T> *E;
    const cx<T> *wx;
    cx<T> *out;
    long long seg_stride, row_stride;
    int nmodes, nsel, os, ntaps;
    long long N;  // outputs per (segment, mode)
    int R;        // outputs per thread (generic kernel)
    ModeList modes;
};

constexpr int APPLY_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(APPLY_THREADS) apply_generic_kernel(ApplyParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int j = blockIdx.y, s = blockIdx.z;
    const int TO = APPLY_THREADS * p.R;
    const long long i0 = (long long)blockIdx.x * TO;
    const int nout = (int)min((long long)TO, p.N - i0);
    const int tile_len = (nout - 1) * p.os + p.ntaps;        // samples of each input row needed
    const int row_pitch = ((TO - 1) * p.os + p.ntaps + 1) & ~1;  // even -> 16B aligned rows (c64)

    cx<T> *xs = reinterpret_cast<cx<T> *>(smem_raw);
    cx<T> *ws = xs + (size_t)p.nmodes * row_pitch;

    const cx<T> *Eseg = p.E + (long long)s * p.seg_stride + i0 * p.os;
    const size_t row_bytes = (size_t)tile_len * sizeof(cx<T>);
    const size_t bulk_bytes = row_bytes & ~(size_t)15;
    // TMA path needs 16B aligned global rows (shared rows are aligned by construction)
    bool tma_ok = bulk_bytes >= 16;
    for (int k = 0; k < p.nmodes; k++)
        tma_ok = tma_ok && ((reinterpret_cast<uintptr_t>(Eseg + (long long)k * p.row_stride) & 15) == 0);

    if (tma_ok) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar, (uint32_t)(bulk_bytes * p.nmodes));
            for (int k = 0; k < p.nmodes; k++)
                tma_bulk_g2s(xs + (size_t)k * row_pitch, Eseg + (long long)k * p.row_stride,
                             (uint32_t)bulk_bytes, &bar);
        }
        // tail (< 16 bytes per row) by ordinary loads
        const int done = (int)(bulk_bytes / sizeof(cx<T>));
        for (int c = tid; c < p.nmodes * (tile_len - done); c += APPLY_THREADS) {
            int k = c / (tile_len - done), o = done + c % (tile_len - done);
            xs[(size_t)k * row_pitch + o] = Eseg[(long long)k * p.row_stride + o];
        }
    } else {
        for (int c = tid; c < p.nmodes * tile_len; c += APPLY_THREADS) {
            int k = c / tile_len, o = c % tile_len;
            xs[(size_t)k * row_pitch + o] = Eseg[(long long)k * p.row_stride + o];
        }
    }
    const cx<T> *wsrc =
        p.wx + ((long long)s * p.nmodes + p.modes.m[j]) * (long long)p.nmodes * p.ntaps;
    for (int c = tid; c < p.nmodes * p.ntaps; c += APPLY_THREADS) ws[c] = wsrc[c];
    __syncthreads();
    if (tma_ok) mbar_wait(&bar, 0);

    T ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
    for (int k = 0; k < p.nmodes; k++) {
        const cx<T> *xr = xs + (size_t)k * row_pitch;
        const cx<T> *wr = ws + k * p.ntaps;
        for (int t = 0; t < p.ntaps; t++) {
            const cx<T> w = wr[t];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int il = tid + r * APPLY_THREADS;
                if (r < p.R && il < nout) {
                    const cx<T> x = xr[il * p.os + t];
                    ar[r] = fma(x.x, w.x, ar[r]);
                    ar[r] = fma(-x.y, w.y, ar[r]);
                    ai[r] = fma(x.x, w.y, ai[r]);
                    ai[r] = fma(x.y, w.x, ai[r]);
                }
            }
        }
    }
    cx<T> *o = p.out + ((long long)s * p.nsel + j) * p.N + i0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int il = tid + r * APPLY_THREADS;
        if (r < p.R && il < nout) o[il] = make_cx<T>(ar[r], ai[r]);
    }
}

// ------------------------------------------------------------------------------------------------
// 2 in x 2 out, os = 2, complex64: register-tiled sliding window
// ------------------------------------------------------------------------------------------------
constexpr int A22_R = 9;          // output symbols per thread; ODD so that the per-lane stride of the
                                  // 128-bit window loads (R*16 B) is conflict-free in shared memory
constexpr int A22_THREADS = 128;  // threads per CTA
constexpr int A22_TO = A22_R * A22_THREADS;

struct Apply22Params {
    const float2 *E;
    const float2 *wx;
    float2 *out;
    long long seg_stride, row_stride;
    long long N, L;
    int ntaps;
};

// Shared layout: x rows as float4 "sample pairs" (x[2q], x[2q+1]); weights as float4 pairs
// (w[2u], w[2u+1]) per (mode, pol), zero padded to an even tap count.
__global__ void __launch_bounds__(A22_THREADS) apply_2x2_os2_kernel(Apply22Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const long long i0 = (long long)blockIdx.x * A22_TO;
    const int nout = (int)min((long long)A22_TO, p.N - i0);
    const int npair_w = (p.ntaps + 1) >> 1;                  // tap pairs
    const int npair_x = A22_TO + npair_w;                    // sample pairs staged per row (max)
    float4 *xs = reinterpret_cast<float4 *>(smem_raw);       // [2][npair_x]
    float4 *ws = xs + 2 * (size_t)npair_x;                   // [2 modes][2 pols][npair_w]

    const float2 *Eseg = p.E + (long long)s * p.seg_stride + i0 * 2;
    // samples of this row that exist from i0*2 on; the tile needs 2*(nout-1) + ntaps of them, but we
    // stage whole pairs (zero filled past the end of the row so the padded odd tap reads a finite 0)
    const long long avail = p.L - i0 * 2;
    const int need = 2 * (nout - 1) + 2 * npair_w;           // even
    const int have = (int)min((long long)need, avail);
    const size_t bulk_bytes = ((size_t)have * sizeof(float2)) & ~(size_t)15;
    bool tma_ok = bulk_bytes >= 16 && ((reinterpret_cast<uintptr_t>(Eseg) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(Eseg + p.row_stride) & 15) == 0);
    float2 *xs2 = reinterpret_cast<float2 *>(xs);
    if (tma_ok) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar, (uint32_t)(2 * bulk_bytes));
            tma_bulk_g2s(xs2, Eseg, (uint32_t)bulk_bytes, &bar);
            tma_bulk_g2s(xs2 + 2 * (size_t)npair_x, Eseg + p.row_stride, (uint32_t)bulk_bytes, &bar);
        }
    }
    {
        const int done = tma_ok ? (int)(bulk_bytes / sizeof(float2)) : 0;
        for (int c = tid; c < 2 * (need - done); c += A22_THREADS) {
            const int k = c / (need - done), o = done + c % (need - done);
            xs2[(size_t)k * 2 * npair_x + o] =
                o < have ? Eseg[(long long)k * p.row_stride + o] : make_float2(0.f, 0.f);
        }
    }
    {
        const float2 *wsrc = p.wx + (long long)s * 4 * p.ntaps;  // (2, 2, ntaps)
        float2 *ws2 = reinterpret_cast<float2 *>(ws);
        for (int c = tid; c < 4 * 2 * npair_w; c += A22_THREADS) {
            const int mk = c / (2 * npair_w), t = c % (2 * npair_w);
            ws2[c] = t < p.ntaps ? wsrc[mk * p.ntaps + t] : make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    if (tma_ok) mbar_wait(&bar, 0);

    // thread owns outputs il = tid*R + r; for tap pair u it needs sample pair q = il + u
    float acc[A22_R][4];  // [r][m0.re, m0.im, m1.re, m1.im]
#pragma unroll
    for (int r = 0; r < A22_R; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const int base = tid * A22_R;
    // warps whose outputs all lie past the end of a ragged last tile skip the arithmetic (warp-uniform)
    const int kend = (base - (tid & 31) * A22_R) < nout ? 2 : 0;
#pragma unroll 1
    for (int k = 0; k < kend; k++) {
        const float4 *xr = xs + (size_t)k * npair_x + base;
        const float4 *w0 = ws + (0 * 2 + k) * npair_w;
        const float4 *w1 = ws + (1 * 2 + k) * npair_w;
        // circular register window: at tap pair u, win[(r + u) % R] holds sample pair base + r + u.  u0 advances
        // by R and uu is a compile-time index, so the rotation is pure register renaming (no moves).
        float4 win[A22_R];
#pragma unroll
        for (int r = 0; r < A22_R - 1; r++) win[r] = xr[r];  // pairs base .. base+R-2
        for (int u0 = 0; u0 < npair_w; u0 += A22_R) {
#pragma unroll
            for (int uu = 0; uu < A22_R; uu++) {
                const int u = u0 + uu;
                if (u < npair_w) {
                    win[(uu + A22_R - 1) % A22_R] = xr[u + A22_R - 1];   // replaces pair base + u - 1
                    const float4 a = w0[u], b = w1[u];
#pragma unroll
                    for (int r = 0; r < A22_R; r++) {
                        const float4 x = win[(r + uu) % A22_R];  // (x[2q].re, x[2q].im, x[2q+1].re, x[2q+1].im)
                        acc[r][0] = fmaf(x.x, a.x, acc[r][0]);
                        acc[r][0] = fmaf(-x.y, a.y, acc[r][0]);
                        acc[r][1] = fmaf(x.x, a.y, acc[r][1]);
                        acc[r][1] = fmaf(x.y, a.x, acc[r][1]);
                        acc[r][0] = fmaf(x.z, a.z, acc[r][0]);
                        acc[r][0] = fmaf(-x.w, a.w, acc[r][0]);
                        acc[r][1] = fmaf(x.z, a.w, acc[r][1]);
                        acc[r][1] = fmaf(x.w, a.z, acc[r][1]);
                        acc[r][2] = fmaf(x.x, b.x, acc[r][2]);
                        acc[r][2] = fmaf(-x.y, b.y, acc[r][2]);
                        acc[r][3] = fmaf(x.x, b.y, acc[r][3]);
                        acc[r][3] = fmaf(x.y, b.x, acc[r][3]);
                        acc[r][2] = fmaf(x.z, b.z, acc[r][2]);
                        acc[r][2] = fmaf(-x.w, b.w, acc[r][2]);
                        acc[r][3] = fmaf(x.z, b.w, acc[r][3]);
                        acc[r][3] = fmaf(x.w, b.z, acc[r][3]);
                    }
                }
            }
        }
    }
    // stage results through shared memory so the global stores are coalesced
    __syncthreads();
    float2 *os0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *os1 = os0 + A22_TO;
#pragma unroll
    for (int r = 0; r < A22_R; r++) {
        os0[base + r] = make_float2(acc[r][0], acc[r][1]);
        os1[base + r] = make_float2(acc[r][2], acc[r][3]);
    }
    __syncthreads();
    float2 *o0 = p.out + ((long long)s * 2 + 0) * p.N + i0;
    float2 *o1 = p.out + ((long long)s * 2 + 1) * p.N + i0;
    for (int c = tid; c < nout; c += A22_THREADS) {
        o0[c] = os0[c];
        o1[c] = os1[c];
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename T>
static int launch_apply(const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                        int64_t nmodes, int64_t L, int64_t os, const void *wx, int64_t ntaps,
                        const int64_t *modes, int64_t nsel, void *out, cudaStream_t st)
{
    const long long N = (L - ntaps + 1) / os;
    if (N <= 0 || nseg == 0 || nsel == 0) return QB_OK;

    bool identity = (nsel == nmodes);
    for (int j = 0; j < nsel; j++) identity = identity && (modes[j] == j);
    if (sizeof(T) == 4 && nmodes == 2 && identity && os == 2 && nseg <= 65535) {
        Apply22Params p;
        p.E = (const float2 *)E;
        p.wx = (const float2 *)wx;
        p.out = (float2 *)out;
        p.seg_stride = seg_stride;
        p.row_stride = row_stride;
        p.N = N;
        p.L = L;
        p.ntaps = (int)ntaps;
        const int npair_w = ((int)ntaps + 1) / 2;
        size_t smem = (size_t)(2 * (A22_TO + npair_w) + 4 * npair_w) * sizeof(float4);
        if (smem <= 200 * 1024) {
            // set on every launch: the attribute belongs to the device that is current, and it is cheap
            QB_CUDA_CHECK(cudaFuncSetAttribute(apply_2x2_os2_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
            dim3 grid((unsigned)((N + A22_TO - 1) / A22_TO), (unsigned)nseg);
            apply_2x2_os2_kernel<<<grid, A22_THREADS, smem, st>>>(p);
            count_launch();
            QB_CUDA_CHECK(cudaGetLastError());
            return QB_OK;
        }
    }

    ApplyParams<T> p;
    p.E = (const cx<T> *)E;
    p.wx = (const cx<T> *)wx;
    p.out = (cx<T> *)out;
    p.seg_stride = seg_stride;
    p.row_stride = row_stride;
    p.nmodes = (int)nmodes;
    p.nsel = (int)nsel;
    p.os = (int)os;
    p.ntaps = (int)ntaps;
    p.N = N;
    p.modes.n = (int)nsel;
    for (int j = 0; j < nsel; j++) p.modes.m[j] = (int)modes[j];
    size_t smem = 0;
    int R = 4;
    for (; R >= 1; R >>= 1) {
        const size_t pitch = (((size_t)APPLY_THREADS * R - 1) * os + ntaps + 1) & ~(size_t)1;
        smem = (nmodes * pitch + nmodes * ntaps) * sizeof(cx<T>);
        if (smem <= 200 * 1024) break;
    }
    if (R < 1)
        return set_error(QB_ERR_UNSUPPORTED, "apply_filter_to_signal: nmodes*ntaps*os too large for shared memory");
    p.R = R;
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(apply_generic_kernel<T>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const long long TO = (long long)APPLY_THREADS * R;
    const long long nblk = (N + TO - 1) / TO;
    if (nseg > 65535 || nblk > 2147483647LL)
        return set_error(QB_ERR_UNSUPPORTED, "apply_filter_to_signal: grid too large");
    dim3 grid((unsigned)nblk, (unsigned)nsel, (unsigned)nseg);
    apply_generic_kernel<T><<<grid, APPLY_THREADS, smem, st>>>(p);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int apply_dispatch(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                   int64_t nmodes, int64_t L, int64_t os, const void *wx, int64_t ntaps,
                   const int64_t *modes, int64_t nsel, void *out, cudaStream_t st)
{
    if (dtype == QB_C64)
        return launch_apply<float>(E, nseg, seg_stride, row_stride, nmodes, L, os, wx, ntaps, modes,
                                   nsel, out, st);
    return launch_apply<double>(E, nseg, seg_stride, row_stride, nmodes, L, os, wx, ntaps, modes, nsel,
                                out, st);
}

}  // namespace qb
