// eq_train "look-ahead" kernel (complex64, os = 2, fixed step size, cma/mcma/rde/mrde).
//
// The plain recurrence  y_i = x_i . w_i ;  w_{i+1} = w_i + c_i conj(x_i)  (pythran_equalisation.py:166-170)
// puts tap update -> tap dot -> shuffle all-reduce -> error function on ONE dependent chain per symbol
// (~400 cycles on B200 with one warp per SM sub-partition).  Substituting the last update,
//
//     y_i = x_i . w_{i-1}  +  c_{i-1} * G_i ,     G_i = sum_q x_i[q] conj(x_{i-1}[q])        (exact algebra)
//
// the expensive part (the two partial dots and their all-reduce) no longer depends on c_{i-1}: while
// the reductions of (P_i, G_i) are in flight the warp applies update i-1 and already accumulates the
// partial dots of symbol i+1; only  P + c*G -> error -> c  (~30 cycles) stays serial.  Same math, same
// results to rounding (the parity tests hold it to the same 1e-5 rms as the direct form), about half
// the time per trained symbol.  Data-only extra work: one more complex MAC per tap for G.
//
// Layout and staging are those of train_sub_kernel (eq_train_fast.cuh); the register window holds
// NQ + 2 samples because the previous symbol's window is needed too.
#pragma once
#include "eq_train_fast.cuh"

namespace qb {

template <int LPS, int NQ, int METHOD, int NVMIN>
__global__ void __launch_bounds__(32) train_la_kernel(TrainParams<float> p, FastGeom g)
{
    static_assert(NQ % 2 == 0, "NQ must be even (os = 2 window rotation)");
    constexpr int B = NQ + 2;      // circular window: previous + current symbol
    constexpr int U = B / 2;       // symbols per unrolled chunk (window rotation period)
    constexpr int GPW = 32 / LPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x, grp = lane / LPS, gl = lane % LPS;
    const long long stream0 = (long long)blockIdx.x * GPW;
    const bool active = stream0 + grp < p.nstreams;
    const long long stream = active ? stream0 + grp : p.nstreams - 1;
    const long long seg = stream / p.nsel;
    const int jsel = (int)(stream % p.nsel);
    const int mode = p.modes.m[jsel];
    const long long seg_first = stream0 / p.nsel;
    const long long seg_last = min(stream0 + GPW - 1, p.nstreams - 1) / p.nsel;
    const int nslots = (int)(seg_last - seg_first) + 1;
    const int slot = (int)(seg - seg_first);

    const int slot_samples = p.nmodes * g.pitch;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *tile1 = tile0 + g.nslots * slot_samples;
    float2 *errs = tile1 + g.nslots * slot_samples;  // [GPW][tile_syms]
    float2 *syms = errs + GPW * g.tile_syms;         // [GPW][nsym_smem]

    const float2 *gsyms = p.symbols + (long long)mode * p.K;
    float2 *mysyms = syms + grp * p.nsym_smem;
    for (int c = gl; c < p.nsym_smem; c += LPS) mysyms[c] = gsyms[c];

    const int k = gl / g.lpp, t0 = (gl % g.lpp) * NQ;
    float wr[NQ], wi[NQ];
    float2 *wg = p.wx + ((long long)seg * p.nmodes + mode) * (long long)(p.nmodes * p.ntaps) + k * p.ntaps;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const bool valid = t0 + q < p.ntaps;
        const float2 w = valid ? wg[t0 + q] : make_float2(0.f, 0.f);
        wr[q] = w.x;
        wi[q] = w.y;
    }
    const float mu = p.mu[stream];
    const int nvalid = min(max(p.ntaps - t0, 0), NQ);
    const uint32_t errs_addr = smem_u32(errs + grp * g.tile_syms);
    __syncwarp();
    const ErrConst ec = load_err_const<METHOD>(mysyms, p.nsym_smem);

    const long long ntiles_it = (p.TrSyms + g.tile_syms - 1) / g.tile_syms;
    const long long ntiles = ntiles_it * p.Niter;
    const long long Lread = (p.TrSyms - 1) * 2 + p.ntaps;
    const bool al16 = ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0) && (p.seg_stride % 2 == 0) &&
                      (p.row_stride % 2 == 0);

    auto load_tile = [&](long long gt, float2 *buf) {
        const long long i0 = (gt % ntiles_it) * g.tile_syms;
        const long long s0 = i0 * 2;
        const int have = (int)max(0LL, min((long long)g.pitch, Lread - s0));
        for (int sl = 0; sl < nslots; sl++) {
            for (int kk = 0; kk < p.nmodes; kk++) {
                const float2 *src = p.E + (seg_first + sl) * p.seg_stride + (long long)kk * p.row_stride + s0;
                float2 *dst = buf + sl * slot_samples + kk * g.pitch;
                if (al16) {
                    const int npair = have >> 1;
                    for (int c = lane; c < npair; c += 32) cp_async<16>(dst + 2 * c, src + 2 * c);
                    if ((have & 1) && lane == 0) cp_async<8>(dst + have - 1, src + have - 1);
                } else {
                    for (int c = lane; c < have; c += 32) cp_async<8>(dst + c, src + c);
                }
                for (int c = have + lane; c < g.pitch; c += 32) dst[c] = make_float2(0.f, 0.f);
            }
        }
        cp_async_commit();
    };

    float2 X[B];
#pragma unroll
    for (int j = 0; j < B; j++) X[j] = make_float2(0.f, 0.f);

    if (ntiles > 0) load_tile(0, tile0);
    for (long long gt = 0; gt < ntiles; gt++) {
        float2 *cur = (gt & 1) ? tile1 : tile0;
        if (gt + 1 < ntiles) {
            load_tile(gt + 1, (gt & 1) ? tile0 : tile1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const long long it = gt / ntiles_it;
        const long long i0 = (gt % ntiles_it) * g.tile_syms;
        const int n = (int)min((long long)g.tile_syms, p.TrSyms - i0);
        const float2 *xrow = cur + slot * slot_samples + k * g.pitch + t0;

        // tile prologue: window of symbol 0 (tile-local positions 0..NQ-1 -> slots 0..NQ-1), no pending
        // update (the previous tile flushed its last one), partial dot of symbol 0
#pragma unroll
        for (int q = 0; q < NQ; q += 2) {
            const float4 v = *reinterpret_cast<const float4 *>(xrow + q);
            X[q] = make_float2(v.x, v.y);
            X[q + 1] = make_float2(v.z, v.w);
        }
        float cpr = 0.f, cpi = 0.f;                   // c_{i-1} = mu * e_{i-1}
        float pr, pi, gr = 0.f, gi = 0.f;             // lane partials of P_i and G_i
        {
            float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                a1 = fmaf(X[q].x, wr[q], a1);
                a2 = fmaf(X[q].y, wi[q], a2);
                b1 = fmaf(X[q].x, wi[q], b1);
                b2 = fmaf(X[q].y, wr[q], b2);
            }
            pr = a1 - a2;
            pi = b1 + b2;
        }
#pragma unroll 1
        for (int il0 = 0; il0 < n; il0 += U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int il = il0 + u;
                const bool live = il < n;
                // (1) all-reduce of this symbol's partials; nothing below (2)-(4) depends on it
                const float Pr = group_sum<LPS>(pr), Pi = group_sum<LPS>(pi);
                const float Gr = group_sum<LPS>(gr), Gi = group_sum<LPS>(gi);
                // (2) apply the previous symbol's update: w_i = w_{i-1} + c_{i-1} conj(x_{i-1})
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float2 x = X[(2 * u + q - 2 + B) % B];
                    if (q < NVMIN || q < nvalid) {
                        wr[q] = fmaf(cpr, x.x, wr[q]);
                        wr[q] = fmaf(cpi, x.y, wr[q]);
                        wi[q] = fmaf(cpi, x.x, wi[q]);
                        wi[q] = fmaf(-cpr, x.y, wi[q]);
                    }
                }
                // (3) slide the window: two new samples replace the two oldest (symbol i-1's first two)
                {
                    const float4 v = *reinterpret_cast<const float4 *>(xrow + 2 * il + NQ);
                    X[(2 * u + NQ) % B] = make_float2(v.x, v.y);
                    X[(2 * u + NQ + 1) % B] = make_float2(v.z, v.w);
                }
                // (4) partial dots of the NEXT symbol with the weights just updated (= w_i) and the
                //     data-only cross term with the current symbol's window
                float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f, g1 = 0.f, g2 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float2 xn = X[(2 * u + 2 + q) % B];   // x_{i+1}[q]
                    const float2 xc = X[(2 * u + q) % B];       // x_i[q]
                    a1 = fmaf(xn.x, wr[q], a1);
                    a2 = fmaf(xn.y, wi[q], a2);
                    b1 = fmaf(xn.x, wi[q], b1);
                    b2 = fmaf(xn.y, wr[q], b2);
                    if (q < NVMIN || q < nvalid) {              // x_{i+1} conj(x_i) over the real taps only
                        g1 = fmaf(xn.x, xc.x, g1);
                        g2 = fmaf(xn.y, xc.y, g2);
                        h1 = fmaf(xn.y, xc.x, h1);
                        h2 = fmaf(xn.x, xc.y, h2);
                    }
                }
                // (5) the only serial part: y_i = P_i + c_{i-1} G_i -> error -> c_i
                float yr = fmaf(cpr, Gr, Pr);
                yr = fmaf(-cpi, Gi, yr);
                float yi = fmaf(cpr, Gi, Pi);
                yi = fmaf(cpi, Gr, yi);
                const float2 e = err_fast<METHOD, LPS>(p.method, make_float2(yr, yi), ec, mysyms, p.K, gsyms, 0, gl);
                if (gl == 0 && live)
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(errs_addr + 8u * il), "f"(e.x), "f"(e.y)
                                 : "memory");
                cpr = live ? mu * e.x : 0.f;
                cpi = live ? mu * e.y : 0.f;
                pr = a1 - a2;
                pi = b1 + b2;
                gr = g1 + g2;
                gi = h1 - h2;
            }
        }
        // tile epilogue: flush the last pending update (window of the last symbol = slots (q-2) mod B)
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const float2 x = X[(q - 2 + B) % B];
            if (q < NVMIN || q < nvalid) {
                wr[q] = fmaf(cpr, x.x, wr[q]);
                wr[q] = fmaf(cpi, x.y, wr[q]);
                wi[q] = fmaf(cpi, x.x, wi[q]);
                wi[q] = fmaf(-cpr, x.y, wi[q]);
            }
        }
        __syncwarp();
        if (p.err && active) {
            float2 *eg = p.err + ((long long)seg * p.nmodes + mode) * (p.TrSyms * p.Niter) + it * p.TrSyms + i0;
            for (int c = gl; c < n; c += LPS) eg[c] = errs[grp * g.tile_syms + c];
        }
        __syncwarp();
    }
    if (active) {
#pragma unroll
        for (int q = 0; q < NQ; q++)
            if (t0 + q < p.ntaps) wg[t0 + q] = make_float2(wr[q], wi[q]);
    }
}

template <int LPS, int NQ, int METHOD, int NVMIN>
static int launch_la(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    constexpr int GPW = 32 / LPS;
    static bool attr_done = false;
    if (!attr_done) {
        QB_CUDA_CHECK(cudaFuncSetAttribute(train_la_kernel<LPS, NQ, METHOD, NVMIN>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_done = true;
    }
    const long long nblk = (p.nstreams + GPW - 1) / GPW;
    train_la_kernel<LPS, NQ, METHOD, NVMIN><<<(unsigned)nblk, 32, smem, st>>>(p, g);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

// Returns 1 if launched, 0 if this problem is outside the look-ahead kernel (adaptive step, generic
// error functions, shapes that do not fit), < 0 on error.
template <int LPS, int NQ>
static int try_la(const TrainParams<float> &p, FastGeom g, cudaStream_t st)
{
    constexpr int GPW = 32 / LPS;
    constexpr int NVMIN = NQ > 4 ? NQ - 4 : 0;
    if (p.adaptive) return 0;
    int method = p.method == QB_SGNCMA ? (int)QB_CMA : p.method;
    if (!(method == QB_CMA || method == QB_MCMA || method == QB_RDE || method == QB_MRDE)) return 0;
    if ((method == QB_RDE || method == QB_MRDE) && (p.K + 1) / 2 > MAXC) return 0;
    // geometry: window period (NQ+2)/2, one more symbol of samples per tile
    constexpr int U = (NQ + 2) / 2;
    g.tile_syms = (64 / U) * U;
    g.pitch = (g.tile_syms * 2 + g.lpp * NQ + 2 + 1) & ~1;
    const size_t smem = ((size_t)2 * g.nslots * p.nmodes * g.pitch + (size_t)GPW * g.tile_syms +
                         (size_t)GPW * p.nsym_smem) * sizeof(float2);
    if (smem > 64 * 1024) return 0;
    int min_valid = NQ;
    for (int j = 0; j < g.lpp; j++) {
        const int nv = p.ntaps - j * NQ;
        if (nv > 0 && nv < min_valid) min_valid = nv;
    }
    const bool empty_lanes = (g.lpp - 1) * NQ >= p.ntaps;
    const bool fastpad = NVMIN > 0 && min_valid >= NVMIN && !empty_lanes;
    int rc;
#define QB_LA_CASE(M)                                                         \
    case M:                                                                   \
        rc = fastpad ? launch_la<LPS, NQ, M, NVMIN>(p, g, smem, st)           \
                     : launch_la<LPS, NQ, M, 0>(p, g, smem, st);              \
        break;
    switch (method) {
        QB_LA_CASE(QB_CMA)
        QB_LA_CASE(QB_MCMA)
        QB_LA_CASE(QB_RDE)
    default:
        rc = fastpad ? launch_la<LPS, NQ, QB_MRDE, NVMIN>(p, g, smem, st)
                     : launch_la<LPS, NQ, QB_MRDE, 0>(p, g, smem, st);
        break;
    }
#undef QB_LA_CASE
    return rc == QB_OK ? 1 : rc;
}

}  // namespace qb
