// Shared pieces of the equaliser-training kernels: parameter block and the error functions
// (qampy/core/equalisation/pythran_equalisation.py:4-16, 178-265).
#pragma once
#include "qb_common.cuh"

namespace qb {

constexpr int GRID_MAX_K = 256;   // largest alphabet the grid slicer of the fast kernels handles (one byte per point)
// levels per axis of the smallest square grid that holds K points (cross QAM: 32 -> 6, 128 -> 12)
__host__ __device__ __forceinline__ int grid_side(int K)
{
    int n = 2;
    while (n * n < K) n++;
    return n;
}

template <typename T>
struct TrainParams {
    const cx<T> *E;
    cx<T> *wx;
    const cx<T> *symbols;
    T *mu;
    cx<T> *err;
    long long seg_stride, row_stride;
    long long TrSyms;
    int nmodes, nsel, os, ntaps;
    int Niter, adaptive, method, K;
    int tile_syms, tile_pitch, nsym_smem;
    int nsym_pitch;      // fast kernels: float2 slots per lane group (constants + scratch of the grid slicer)
    long long L;         // samples per row that may be read (fast kernel: bounds of the padded window)
    long long nstreams;  // nseg * nsel
    ModeList modes;
};

// first strict minimum of |x - s_j|^2 over the alphabet, d0 = 1000, s = 1 (pythran_equalisation.py:240-265)
template <typename T>
__device__ __forceinline__ cx<T> det_symbol_warp(cx<T> x, const cx<T> *syms, int K, int lane)
{
    T best = (T)1000.;
    int bj = 0x7fffffff;
    for (int j = lane; j < K; j += 32) {
        const cx<T> s = syms[j];
        const T dr = x.x - s.x, di = x.y - s.y;
        const T d = dr * dr + di * di;
        if (d < best) {
            best = d;
            bj = j;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const T ob = shfl_xor(best, m);
        const int oj = shfl_xor(bj, m);
        if (ob < best || (ob == best && oj < bj)) {
            best = ob;
            bj = oj;
        }
    }
    if (bj == 0x7fffffff) return make_cx<T>((T)1, (T)0);
    return syms[bj];
}

// partition_value, pythran_equalisation.py:4-9 (part = 0: real tables, 1: imaginary tables)
template <typename T>
__device__ __forceinline__ T partition_value(T signal, const cx<T> *partitions, int np_,
                                             const cx<T> *codebook, int part)
{
    int index = 0;
    while (index < np_ && signal > (part ? partitions[index].y : partitions[index].x)) index++;
    return part ? codebook[index].y : codebook[index].x;
}

template <typename T>
__device__ __forceinline__ cx<T> error_fct(int method, cx<T> x, const cx<T> *syms, int K,
                                           const cx<T> *gsyms, long long i, int lane)
{
    switch (method) {
    case QB_CMA:
    case QB_SGNCMA: {
        const T d = syms[0].x - (x.x * x.x + x.y * x.y);
        return make_cx<T>(d * x.x, d * x.y);
    }
    case QB_CMA2: {
        const T dr = syms[0].x - (x.x * x.x - x.y * x.y);
        const T di = syms[0].y - (x.x * x.y + x.y * x.x);
        return make_cx<T>(dr * x.x - di * x.y, dr * x.y + di * x.x);
    }
    case QB_MCMA: {
        const T dr = syms[0].x - x.x * x.x;
        const T di = syms[0].y - x.y * x.y;
        return make_cx<T>(dr * x.x, di * x.y);
    }
    case QB_RDE: {
        const int nc = (K + 1) / 2;
        const T sq = x.x * x.x + x.y * x.y;
        const T r = partition_value<T>(sq, syms + nc, K - nc, syms, 0);
        const T d = r - sq;
        return make_cx<T>(x.x * d, x.y * d);
    }
    case QB_MRDE: {
        const int nc = (K + 1) / 2;
        const T sqr = x.x * x.x, sqi = x.y * x.y;
        const T rr = partition_value<T>(sqr, syms + nc, K - nc, syms, 0);
        const T ri = partition_value<T>(sqi, syms + nc, K - nc, syms, 1);
        return make_cx<T>((rr - sqr) * x.x, (ri - sqi) * x.y);
    }
    case QB_SBD: {
        const cx<T> s = det_symbol_warp<T>(x, syms, K, lane);
        return make_cx<T>((s.x - x.x) * fabs(s.x), (s.y - x.y) * fabs(s.y));
    }
    case QB_SBD_DATA: {
        const cx<T> s = gsyms[i];
        return make_cx<T>((s.x - x.x) * fabs(s.x), (s.y - x.y) * fabs(s.y));
    }
    case QB_MDDMA: {
        const cx<T> s = det_symbol_warp<T>(x, syms, K, lane);
        return make_cx<T>((s.x * s.x - x.x * x.x) * x.x, (s.y * s.y - x.y * x.y) * x.y);
    }
    case QB_CMA_REAL: {   // pythran_equalisation.py:113-115, real signal in the real part
        const T d = syms[0].x - x.x * x.x;
        return make_cx<T>(d * x.x, (T)0);
    }
    case QB_SGNCMA_REAL: {   // :117-119, np.sign(0) = 0
        const T v = syms[0].x - x.x * x.x;
        const T d = v > (T)0 ? (T)1 : (v < (T)0 ? (T)-1 : (T)0);
        const T sx = x.x > (T)0 ? (T)1 : (x.x < (T)0 ? (T)-1 : (T)0);
        return make_cx<T>(d * sx, (T)0);
    }
    case QB_DD_REAL: {   // :121-123 with det_symbol_argmin (:232-235): first minimum of |X - s|
        const cx<T> s = det_symbol_warp<T>(make_cx<T>(x.x, (T)0), syms, K, lane);
        return make_cx<T>((s.x - x.x) * fabs(s.x), (T)0);
    }
    case QB_DD_DATA_REAL: {   // :125-128
        const T s = gsyms[i].x;
        return make_cx<T>((s - x.x) * fabs(s), (T)0);
    }
    default: {
        const cx<T> s = det_symbol_warp<T>(x, syms, K, lane);
        return make_cx<T>(s.x - x.x, s.y - x.y);
    }
    }
}

// adapt_step(mu, err_p = e_i, err = e_{i-1}), pythran_equalisation.py:12-16 with the :172 call order
template <typename T>
__device__ __forceinline__ T adapt_step(T mu, cx<T> cur, cx<T> prev)
{
    if (prev.x * cur.x > 0 && prev.y * cur.y > 0) return mu;
    return mu / ((T)1 + mu * (prev.x * prev.x + prev.y * prev.y));
}
// adapt_step_real(mu, err_p = e_i, err = e_{i-1}), pythran_equalisation.py:18-22
template <typename T>
__device__ __forceinline__ T adapt_step_real(T mu, T cur, T prev)
{
    if (prev * cur > 0) return mu;
    return mu / ((T)1 + mu * (prev * prev));
}

}  // namespace qb
