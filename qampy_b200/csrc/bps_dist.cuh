// Distance evaluation, unwrap step and parameter block shared by the column-per-lane blind-phase-search kernels
// (bps_fast.cu: fused / producer-chain mappings; bps_par.cu: the phase-parallel form for few long streams).  One
// definition, so that every mapping performs the same operations on the same operands: indices and phases are
// bit-identical between them.
#pragma once
#include "qb_common.cuh"

namespace qb {

struct BpsFastParams {
    const float2 *E;
    const float2 *comp;
    const float *angles;
    const float *lev_re, *lev_im;
    int32_t *idx;
    float *ph;
    float2 *Eout;
    long long stream_stride, L;
    int A, n_re, n_im, N;
};

__device__ __forceinline__ float fma_sat(float a, float b, float c)
{
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// slicer table entry: read-only after initialisation, so the load is a pure function of the address
__device__ __forceinline__ f32x2 lds_tab(uint32_t addr)
{
    f32x2 r;
    asm("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(addr));
    return r;
}

// Slicer constants of one axis.  The table entry f holds the NEGATED bracket (-lev[f], -lev[min(f+1, n-1)]),
// f = 0 .. n-1, so that the two candidate differences are one FADD2.  f = floor((t - lev0)/step) clamped to
// [0, n-1] is computed as u = sat(t*s1 + b1) (u = (x - 1/2)/(n-1) clamped to [0, 1]) followed by
// w = u*(n-1) + 1.5*2^23, whose low mantissa bits are round(x - 1/2).  A coordinate that lands on the
// wrong side of an integer by rounding still selects a bracket with the nearest level as an end point,
// and the differences use the stored level values, so the minimum is bit-identical to the reference's
// search over all M symbols (IEEE rounding is monotone).
struct FastAxis {
    float s1, b1, nm1;
    uint32_t kaddr;   // table byte address - (0x4b400000 << 3)
};
__device__ __forceinline__ FastAxis make_fast_axis(const float *lev, int n, uint32_t tab_addr)
{
    FastAxis g;
    const float span = n > 1 ? lev[n - 1] - lev[0] : 1.f;
    const float step = n > 1 ? span / (float)(n - 1) : 1.f;
    g.nm1 = (float)(n > 1 ? n - 1 : 0);
    g.s1 = n > 1 ? 1.f / span : 0.f;
    g.b1 = n > 1 ? (-lev[0] / step - 0.5f) / (float)(n - 1) : 0.f;
    g.kaddr = tab_addr - (0x4b400000u << 3);
    return g;
}
// min over the levels of |t - lev| (bit-exact), one axis
__device__ __forceinline__ float axis_min_fast(float t, const FastAxis &g)
{
    const float u = fma_sat(t, g.s1, g.b1);
    const float w = fmaf(u, g.nm1, 12582912.f);
    const f32x2 nl = lds_tab((__float_as_uint(w) << 3) + g.kaddr);
    const float2 df = add2_bcast(t, nl);
    return fminf(fabsf(df.x), fabsf(df.y));
}

__device__ __forceinline__ float unwrap_corr_f(float p, float pprev)
{
    // one step of np.unwrap (default period / discont) in float32, op by op
    const float PI = 3.14159274101257324219f, TWO_PI = 6.28318548202514648438f;
    const float dd = __fsub_rn(p, pprev);
    float m = fmodf(__fadd_rn(dd, PI), TWO_PI);
    if (m != 0.f && m < 0.f) m = __fadd_rn(m, TWO_PI);
    float ddmod = __fsub_rn(m, PI);
    if (ddmod == -PI && dd > 0.f) ddmod = PI;
    float corr = __fsub_rn(ddmod, dd);
    if (fabsf(dd) < PI) corr = 0.f;
    return corr;
}
__device__ __forceinline__ float2 rotate_f(float2 e, float ph)
{
    float s, c;
    sincosf(ph, &s, &c);
    return make_float2(e.x * c - e.y * s, e.x * s + e.y * c);
}

// distance of one row to the nearest alphabet point after rotation by one test angle: c1 = (cr, ci), c2 = (-ci, cr)
// of that angle; unfused complex multiply (pythran_dsp.py:79), d = fl(fl(dr^2) + fl(di^2)) clamped at 100 (:73, :81-82)
__device__ __forceinline__ float fast_dist(float2 ev, f32x2 c1, f32x2 c2, const FastAxis &gre, const FastAxis &gim)
{
    const float2 pa = mul2_bcast(ev.x, c1), pb = mul2_bcast(ev.y, c2);
    const float tr = __fadd_rn(pa.x, pb.x);
    const float ti = __fadd_rn(pa.y, pb.y);
    float2 dm;
    dm.x = axis_min_fast(tr, gre);
    dm.y = axis_min_fast(ti, gim);
    const float2 sq = sqr2(dm);
    return fminf(__fadd_rn(sq.x, sq.y), 100.f);   // NaN -> 100
}

}  // namespace qb
