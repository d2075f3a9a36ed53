// Parameter block, slicer, distance, unwrap step and rotation shared by the generic (any dtype, any alphabet) blind
// phase search kernels: bps.cu (one CTA per stream) and the phase-parallel form for few long streams (bps_par.cu).
#pragma once
#include "qb_common.cuh"

namespace qb {

constexpr int BPS_THREADS = 256;

template <typename T>
struct BpsParams {
    const cx<T> *E;
    const cx<T> *comp;
    const T *angles;
    const cx<T> *symbols;
    const T *lev_re, *lev_im;
    int32_t *idx;
    T *ph;
    cx<T> *Eout;
    long long stream_stride, L;
    int A, M, n_re, n_im, N;
    int tile_rows, ring_rows;
    int comp_rows;   // 0: one table comp[A]; 1: per-symbol table comp[L][A] per stream (two-stage BPS, :74-77)
    int windowed;    // QB_BPS_WINDOWED: window sums formed directly from the 2N distances (bps_kernel only)
};

// Per-axis slicer.  `pairs[f] = (lev[f], lev[f+1])` are the two levels bracketing a value whose
// (approximate) grid coordinate floors to f; the nearest level is always one of them (a coordinate
// that is off by rounding near an integer still yields a bracket that contains the nearest level),
// and because the subtraction uses the STORED level values the result is bit-identical to the
// brute-force minimum.  The coordinate is formed with ONE FMA whose addend already carries the
// 1.5*2^23 rounding constant, so floor() costs no conversion instruction: the integer sits in the
// low mantissa bits.  scale = 1/step, bias = -lev0/step - 0.5 + 1.5*2^23.
template <typename T>
struct AxisGrid {
    T scale, bias;
    int npair;
};
__device__ __forceinline__ float axis_min(float t, const float2 *pairs, const AxisGrid<float> &g)
{
    const float v = fmaf(t, g.scale, g.bias);
    int f = __float_as_int(v) - 0x4b400000;          // round-to-nearest integer of t*scale - lev0*scale - 0.5
    f = max(0, min(f, g.npair - 1));                 // also tames huge / NaN inputs
    const float2 l = pairs[f];
    return fminf(fabsf(__fsub_rn(t, l.x)), fabsf(__fsub_rn(t, l.y)));  // NaN only if t is NaN
}
__device__ __forceinline__ double axis_min(double t, const double2 *pairs, const AxisGrid<double> &g)
{
    double uf = floor(fma(t, g.scale, g.bias));
    uf = uf > 0. ? uf : 0.;
    uf = uf < (double)(g.npair - 1) ? uf : (double)(g.npair - 1);
    const double2 l = pairs[(int)uf];
    return fmin(fabs(__dsub_rn(t, l.x)), fabs(__dsub_rn(t, l.y)));
}
__device__ __forceinline__ AxisGrid<float> make_grid(const float *lev, int n)
{
    AxisGrid<float> g;
    g.npair = max(n - 1, 1);
    g.scale = n > 1 ? (float)(n - 1) / (lev[n - 1] - lev[0]) : 0.f;
    g.bias = -lev[0] * g.scale - 0.5f + 12582912.f;
    return g;
}
__device__ __forceinline__ AxisGrid<double> make_grid(const double *lev, int n)
{
    AxisGrid<double> g;
    g.npair = max(n - 1, 1);
    g.scale = n > 1 ? (double)(n - 1) / (lev[n - 1] - lev[0]) : 0.;
    g.bias = -lev[0] * g.scale;
    return g;
}

template <typename T>
struct Pi;
template <>
struct Pi<float> {
    static __device__ __forceinline__ float pi() { return 3.14159274101257324219f; }      // fl32(pi)
    static __device__ __forceinline__ float two_pi() { return 6.28318548202514648438f; }  // fl32(2 pi)
};
template <>
struct Pi<double> {
    static __device__ __forceinline__ double pi() { return 3.141592653589793115997963; }
    static __device__ __forceinline__ double two_pi() { return 6.283185307179586231995927; }
};

// one step of np.unwrap's correction (numpy/lib/_function_base_impl.py unwrap, default period/discont)
template <typename T>
__device__ __forceinline__ T unwrap_corr(T p, T pprev)
{
    const T PI = Pi<T>::pi(), TWO_PI = Pi<T>::two_pi();
    const T dd = sub_rn(p, pprev);
    T m = fmod(add_rn(dd, PI), TWO_PI);  // np.mod: python-style, divisor > 0
    if (m != (T)0 && m < (T)0) m = add_rn(m, TWO_PI);
    T ddmod = sub_rn(m, PI);
    if (ddmod == -PI && dd > (T)0) ddmod = PI;
    T corr = sub_rn(ddmod, dd);
    if (fabs(dd) < PI) corr = (T)0;
    return corr;
}

__device__ __forceinline__ void qb_sincos(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void qb_sincos(double x, double *s, double *c) { sincos(x, s, c); }

// E * exp(1j*ph), phaserecovery.py:157-159
template <typename T>
__device__ __forceinline__ cx<T> rotate(cx<T> e, T ph)
{
    T s, c;
    qb_sincos(ph, &s, &c);
    return make_cx<T>(e.x * c - e.y * s, e.x * s + e.y * c);
}

template <typename T>
__device__ __forceinline__ T min_distance(cx<T> e, cx<T> c, bool slicer, const cx<T> *pre, const cx<T> *pim,
                                          const AxisGrid<T> &gre, const AxisGrid<T> &gim, const cx<T> *syms,
                                          int M)
{
    const T tr = sub_rn(mul_rn(e.x, c.x), mul_rn(e.y, c.y));   // E[i]*comp[a], unfused (pythran_dsp.py:79)
    const T ti = add_rn(mul_rn(e.x, c.y), mul_rn(e.y, c.x));
    T d;
    if (slicer) {
        const T da = axis_min(tr, pre, gre);
        const T db = axis_min(ti, pim, gim);
        d = add_rn(mul_rn(da, da), mul_rn(db, db));
    } else {
        d = (T)1000.;
        for (int m = 0; m < M; m++) {
            const cx<T> sy = syms[m];
            const T dr = sub_rn(tr, sy.x), di = sub_rn(ti, sy.y);
            const T dd = add_rn(mul_rn(dr, dr), mul_rn(di, di));
            if (dd < d) d = dd;
        }
    }
    return d < (T)100. ? d : (T)100.;                          // :73, :81-82
}

}  // namespace qb
