// eq_train look-ahead kernel (complex64, os = 2, fixed step size): the planar/packed layout of
// train_sub_kernel (eq_train_fast.cuh) with the serial chain cut down to the error function.
//
// The reference recurrence (pythran_equalisation.py:166-170)
//     y_i = X_i . W_i ,   c_i = mu * errfct(y_i) ,   W_{i+1} = W_i + c_i conj(X_i)
// puts tap update -> tap dot -> shuffle all-reduce -> error function on ONE dependent chain per symbol
// (~320 cycles on B200 with one warp per SM sub-partition, of which ~120 are pure latency during which
// the warp has nothing to issue).  Substituting the last update into the dot,
//     y_{i+1} = X_{i+1} . W_i  +  c_i * G_{i+1} ,      G_{i+1} = X_{i+1} . conj(X_i)        (exact algebra)
// the expensive part Q_{i+1} = X_{i+1} . W_i no longer depends on c_i, and G is a property of the
// SIGNAL alone (the lag-2 autocorrelation of the window, the same for every mode trained on the
// segment): the warp computes it for a whole staged tile at once from a running sum of lag-2 products
// (tile_gram, ~4 instructions per symbol) and reads it back one value per symbol.  Per symbol the warp
// then issues, in this order,
//     1. the all-reduce of the partial Q_i, its three shuffle hops interleaved with the tap update
//        W_i = W_{i-1} + c_{i-1} conj(X_{i-1})                        (24 FFMA2 fill the shuffle latency)
//     2. y_i = Q_i + c_{i-1} G_i -> error function -> c_i, interleaved with the partial dot
//        Q_{i+1} = X_{i+1} . W_i                                      (24 FFMA2 fill the error-function chain)
// so only  Q + c*G -> errfct -> c  (~35 cycles) is serial and the rest is bound by instruction issue.
// Same FMAs as the direct form plus 4 per symbol; results equal it to rounding (the parity tests hold
// both to the same 1e-5 rms against the oracle; a numpy model of this recursion differs from the strict
// oracle by 6e-7 rms on the error signal of a C3 segment).
//
// The register window holds NP + 2 pairs per plane because symbol i needs the windows of symbols i-1
// (update) and i+1 (dot); tiles are staged from one pair (2 samples) before the tile's first symbol.
#pragma once
#include <algorithm>
#include <type_traits>

#include "eq_train_fast.cuh"

namespace qb {

// Gram values of one staged tile, computed by the warp itself right after the tile has landed:
//     G_il = sum_k sum_{t < ntaps} x_k[2 il + 2 + t] conj(x_k[2 il + t])        (staged sample indices)
//          = S[2 il + ntaps] - S[2 il],      S[j] = sum_{m = 2 .. j + 1} p[m],   p[m] = sum_k x_k[m] conj(x_k[m - 2]).
// Every lane forms GRAM_CH consecutive lag-2 products from the planar tile (odd chunk length: conflict
// free), a warp scan turns them into the running sum S, and G is two loads and one subtraction per
// symbol.  About 4 instructions per trained symbol -- instead of a separate pass over the capture and
// 8 bytes of HBM per symbol.  G only ever enters multiplied by the step mu*e ~ 5e-4, so the rounding of
// a running sum over <= 416 products (<= 3e-5 absolute) is far below the fp32 resolution of y.
constexpr int GRAM_CH = 13;   // 32 * 13 = 416 >= 2 * 128 + 8 * 12 + 2 products per tile

// entries of the running sum a tile needs: S[2 il + ntaps] for il < tile_syms
__host__ __device__ __forceinline__ int gram_sum_len(int tile_syms, int ntaps) { return 2 * (tile_syms - 1) + ntaps + 1; }

__device__ __forceinline__ void tile_gram(const float *tile, int nslots, int nmodes, int pitch, int tile_syms,
                                          int ntaps, float2 *S, float2 *gbuf, int lane)
{
    const int slen = gram_sum_len(tile_syms, ntaps);
    const int row_floats = 2 * pitch, slot_floats = nmodes * row_floats;
    for (int sl = 0; sl < nslots; sl++) {
        const float *base = tile + sl * slot_floats;
        float2 *Ss = S + sl * slen;
        const int mb = 2 + lane * GRAM_CH;
        // products beyond mend read whatever follows in shared memory; running sums are causal, so they only
        // reach entries of S that nobody reads -- no bounds branch in here
        float pr[GRAM_CH], pi[GRAM_CH];
#pragma unroll
        for (int j = 0; j < GRAM_CH; j++) pr[j] = pi[j] = 0.f;
        for (int kk = 0; kk < nmodes; kk++) {
            const float *re = base + kk * row_floats + mb - 2, *im = re + pitch;
            float xr[GRAM_CH + 2], xi[GRAM_CH + 2];
#pragma unroll
            for (int j = 0; j < GRAM_CH + 2; j++) {
                xr[j] = re[j];
                xi[j] = im[j];
            }
#pragma unroll
            for (int j = 0; j < GRAM_CH; j++) {   // x[m] * conj(x[m-2])
                pr[j] = fmaf(xr[j + 2], xr[j], fmaf(xi[j + 2], xi[j], pr[j]));
                pi[j] = fmaf(xi[j + 2], xr[j], fmaf(-xr[j + 2], xi[j], pi[j]));
            }
        }
        float2 loc[GRAM_CH];
        float2 run = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < GRAM_CH; j++) {
            run.x += pr[j];
            run.y += pi[j];
            loc[j] = run;
        }
        // exclusive scan of the lane totals
        float2 off = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float ox = __shfl_up_sync(0xffffffffu, off.x, d), oy = __shfl_up_sync(0xffffffffu, off.y, d);
            if (lane >= d) {
                off.x += ox;
                off.y += oy;
            }
        }
        off.x -= run.x;
        off.y -= run.y;
        if (lane == 0) Ss[0] = make_float2(0.f, 0.f);
        // only S[0 .. 2 (tile_syms - 1) + ntaps] is ever read (and allocated: gram_sum_len)
#pragma unroll
        for (int j = 0; j < GRAM_CH; j++)
            if (mb + j - 1 < slen) Ss[mb + j - 1] = make_float2(off.x + loc[j].x, off.y + loc[j].y);
        __syncwarp();
        for (int il = lane; il < tile_syms; il += 32) {
            const float2 hi = Ss[2 * il + ntaps], lo = Ss[2 * il];
            gbuf[sl * tile_syms + il] = make_float2(hi.x - lo.x, hi.y - lo.y);
        }
        __syncwarp();
    }
}

// GRID (sbd / dd / mddma only): the decision of a searched alphabet is compiled in -- 1 grid slicer, 0 list search -- so that
// no branch sits in the symbol loop.  Whether an alphabet is a grid is found out in the kernel (detect_grid), so
// both instantiations are launched back to back and each one works on the streams whose alphabet is its kind (all
// or none of them in practice; the other launch returns after its prologue).  -1: decided at run time.
// ADAPT: the reference's adaptive step size (pythran_equalisation.py:12-16, :171-172).  The step size used for symbol
// i's update depends on the errors of symbols i-1 and i-2 only, so it is known before e_i is and stays off the serial
// chain of the look-ahead form: c_i = mu_i e_i, mu_{i+1} = adapt_step(mu_i, e_i, e_{i-1}).
template <int LPS, int NQ, int METHOD, int NMASK, int GRID = -1, bool ADAPT = false>
__global__ void __launch_bounds__(32 * TRAIN_WPB) train_la_kernel(TrainParams<float> p, FastGeom g, int warp_smem)
{
    static_assert(NQ % 2 == 0, "NQ must be even (os = 2: the window moves by one pair per symbol)");
    constexpr int NP = NQ / 2;     // pairs per lane
    constexpr int B = NP + 2;      // circular pair window: symbols i-1 .. i+1; also symbols per unrolled chunk
    constexpr int GPW = 32 / LPS;  // streams (lane groups) per warp
    extern __shared__ __align__(16) unsigned char smem_all[];
    // TRAIN_WPB independent warps per CTA, each with its own slice of shared memory and its own streams (no
    // block-level synchronisation anywhere): see launch geometry below
    const int wib = threadIdx.x >> 5;
    const long long wblk = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (wblk * GPW >= p.nstreams) return;
    unsigned char *smem_raw = smem_all + (size_t)wib * warp_smem;
    const int lane = threadIdx.x & 31, grp = lane / LPS, gl = lane % LPS;
    const long long stream0 = wblk * GPW;
    const bool active = stream0 + grp < p.nstreams;
    const long long stream = active ? stream0 + grp : p.nstreams - 1;
    const long long seg = stream / p.nsel;
    const int jsel = (int)(stream % p.nsel);
    const int mode = p.modes.m[jsel];
    const long long seg_first = stream0 / p.nsel;
    const long long seg_last = min(stream0 + GPW - 1, p.nstreams - 1) / p.nsel;
    const int nslots = (int)(seg_last - seg_first) + 1;
    const int slot = (int)(seg - seg_first);

    // tiles: [2][nslots][nmodes][2 planes][pitch] floats; Gram values [nslots][tile_syms] and their running
    // sums [nslots][32*GRAM_CH + 1]; errors; constants
    const int row_floats = 2 * g.pitch, slot_floats = p.nmodes * row_floats;
    float *tile0 = reinterpret_cast<float *>(smem_raw);
    float *tile1 = tile0 + g.nslots * slot_floats;
    float2 *gbuf = reinterpret_cast<float2 *>(tile1 + g.nslots * slot_floats);
    // The running sums live only inside tile_gram (start of a tile), the error buffer is filled by the symbol loop
    // and flushed at the end of the tile: they share one region, which keeps a warp's slice small enough for
    // 8 warps per SM when a launch holds more than one wave of streams (long captures).
    float2 *gsum = gbuf + g.nslots * g.tile_syms;
    // [GPW][tile_syms + 1]: the lane groups of a warp store the errors of their streams with ONE instruction, so the
    // rows sit one float2 apart in the banks (a pitch of tile_syms put all groups on one bank pair: 10 M of the 15 M
    // excessive shared-memory wavefronts of a C3 training launch)
    float2 *errs = gsum;
    const int shared_len = max(g.nslots * gram_sum_len(g.tile_syms, p.ntaps), GPW * (g.tile_syms + 1));
    float2 *syms = gsum + shared_len;                       // [GPW][nsym_smem]

    const float2 *gsyms = p.symbols + (long long)mode * p.K;
    float2 *mysyms = syms + grp * p.nsym_pitch;
    for (int c = gl; c < p.nsym_smem; c += LPS) mysyms[c] = gsyms[c];

    const int k = gl / g.lpp, t0 = (gl % g.lpp) * NQ;
    f32x2 PR[NP], PI[NP];   // taps: (re[2p], re[2p+1]) and (im[2p], im[2p+1])
    f32x2 MK[NMASK > 0 ? NMASK : 1];
    float2 *wg = p.wx + ((long long)seg * p.nmodes + mode) * (long long)(p.nmodes * p.ntaps) + k * p.ntaps;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const bool v0 = t0 + 2 * q < p.ntaps, v1 = t0 + 2 * q + 1 < p.ntaps;
        const float2 w0 = v0 ? wg[t0 + 2 * q] : make_float2(0.f, 0.f);
        const float2 w1 = v1 ? wg[t0 + 2 * q + 1] : make_float2(0.f, 0.f);
        PR[q] = pack2(w0.x, w1.x);
        PI[q] = pack2(w0.y, w1.y);
        if (q >= NP - NMASK) MK[q - (NP - NMASK)] = pack2(v0 ? 1.f : 0.f, v1 ? 1.f : 0.f);
    }
    float mu = p.mu[stream];
    float2 eprev = make_float2(0.f, 0.f);    // ADAPT: error of the symbol before
    const uint32_t errs_addr = smem_u32(errs + grp * (g.tile_syms + 1));
    __syncwarp();
    ErrConst ec = load_err_const<METHOD>(mysyms, p.nsym_smem);
    if (p.nsym_pitch > p.nsym_smem)   // searched alphabet: is it a square grid? (uniform)
        detect_grid<LPS>(ec, mysyms, p.nsym_smem, reinterpret_cast<float *>(mysyms + p.nsym_smem), gl);
    bool act = active;              // this launch records results for this lane's stream
    if (GRID >= 0) {
        const int kind = ec.gn == 0 ? 0 : (ec.gholes ? 2 : 1);     // list search, full grid, grid with empty cells
        const bool mine = kind == GRID;
        if (!__any_sync(0xffffffffu, mine && active)) return;
        act = active && mine;
    }

    const long long ntiles_it = (p.TrSyms + g.tile_syms - 1) / g.tile_syms;
    const long long ntiles = ntiles_it * p.Niter;
    const long long Lread = (p.TrSyms - 1) * 2 + p.ntaps;  // samples of a row the caller guarantees

    // stage tile gt: samples [2*i0 - 2, 2*i0 - 2 + pitch) of every row (zero outside [0, Lread))
    auto load_tile = [&](long long gt, float *buf) {
        const long long i0 = (gt % ntiles_it) * g.tile_syms;
        const long long s0 = i0 * 2 - 2;
        const int lo = s0 < 0 ? (int)(-s0) : 0;
        const int hi = (int)max((long long)lo, min((long long)g.pitch, Lread - s0));
        for (int sl = 0; sl < nslots; sl++) {
            for (int kk = 0; kk < p.nmodes; kk++) {
                const float *src = reinterpret_cast<const float *>(p.E + (seg_first + sl) * p.seg_stride +
                                                                   (long long)kk * p.row_stride) + 2 * s0;
                // float c of the staged row: even -> real plane, odd -> imaginary plane (c and lane share parity)
                float *dst = buf + sl * slot_floats + kk * row_floats + (lane & 1) * g.pitch + (lane >> 1);
                const float *s = src + lane;
                if (lo == 0 && hi == g.pitch) {          // interior tile: every staged sample exists
                    int c = lane;
#pragma unroll 4
                    for (; c < 2 * g.pitch; c += 32, dst += 16, s += 32) cp_async<4>(dst, s);
                } else {
                    for (int c = lane; c < 2 * g.pitch; c += 32, dst += 16, s += 32) {
                        const int m = c >> 1;
                        if (m >= lo && m < hi) cp_async<4>(dst, s);
                        else *dst = 0.f;
                    }
                }
            }
        }
        cp_async_commit();
    };

    float fx_prev = 0.f;            // latency layout: largest lane partial of the previous tile (0: none yet)
    float crp = 0.f, cip = 0.f;     // c_{i-1} = mu * e_{i-1}: the update that is still to be applied
    float pqr = 0.f, pqi = 0.f;     // this lane's partial of Q_i = X_i . W_{i-1}

    if (ntiles > 0) load_tile(0, tile0);
    for (long long gt = 0; gt < ntiles; gt++) {
        float *cur = (gt & 1) ? tile1 : tile0;
        if (gt + 1 < ntiles) {
            load_tile(gt + 1, (gt & 1) ? tile0 : tile1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        tile_gram(cur, nslots, p.nmodes, g.pitch, g.tile_syms, p.ntaps, gsum, gbuf, lane);
        const long long it = gt / ntiles_it;
        const long long tl = gt % ntiles_it;
        const long long i0 = tl * g.tile_syms;
        const int n = (int)min((long long)g.tile_syms, p.TrSyms - i0);
        const uint32_t xre = smem_u32(cur + slot * slot_floats + k * row_floats + t0);
        const uint32_t xim = xre + 4u * (uint32_t)g.pitch;
        const uint32_t gaddr = smem_u32(gbuf + slot * g.tile_syms);

        // staged pair j (samples 2j, 2j+1 of the staged row) lives in slot j % B; symbol il of the tile
        // has its window in pairs il+1 .. il+NP, symbol il-1 in pairs il .. il+NP-1, symbol il+1 in il+2 ..
        f32x2 XR[B], XI[B];
        // One pass over the tile's symbols.  FIXED (latency layout only): the 32 lane partials of the tap dot are
        // summed as integers by REDUX after scaling by `fx_scale`, a power of two chosen per tile (below); the
        // pass records the largest partial it has seen in `tmax`.  !FIXED: shuffle all-reduce, any magnitude.
        float tmax = 0.f;
        auto tile_pass = [&](auto fixed_tag, const float fx_scale, const float fx_inv) {
        constexpr bool FIXED = decltype(fixed_tag)::value;
#pragma unroll
        for (int q = 0; q <= NP; q++) {
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XR[q]) : "r"(xre + 8u * q));
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XI[q]) : "r"(xim + 8u * q));
        }
        if (tl == 0) {
            // start of a training iteration: nothing pending, Q_0 = X_0 . W_0 directly (pairs 1 .. NP)
            crp = cip = 0.f;
            f32x2 a1 = 0ull, a2 = 0ull, b1 = 0ull, b2 = 0ull;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const f32x2 xr = XR[(1 + q) % B], xi = XI[(1 + q) % B];
                a1 = fma2(xr, PR[q], a1);
                a2 = fma2(xi, PI[q], a2);
                b1 = fma2(xr, PI[q], b1);
                b2 = fma2(xi, PR[q], b2);
            }
            const float2 sa = unpack2(sub2(a1, a2)), sb = unpack2(add2(b1, b2));
            pqr = sa.x + sa.y;
            pqi = sb.x + sb.y;
        }
#pragma unroll 1
        for (int il0 = 0; il0 < n; il0 += B) {
#pragma unroll
            for (int u = 0; u < B; u++) {
                const int il = il0 + u;
                const bool live = il < n;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XR[(u + NP + 1) % B]) : "r"(xre + 8u * (il + NP + 1)));
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XI[(u + NP + 1) % B]) : "r"(xim + 8u * (il + NP + 1)));
                float2 Gi;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(Gi.x), "=f"(Gi.y) : "r"(gaddr + 8u * il));
                // ---- 1. all-reduce of the partial Q_il, hops interleaved with W += c_{il-1} conj(X_{il-1}) ----
                float qr = pqr, qi = pqi;
                const float ncr = -crp;
                if constexpr (FIXED) {
                    // one stream per warp (latency layout): the 32 partials are summed as block-floating fixed point
                    // by ONE warp-wide integer reduction each (REDUX) instead of five shuffle hops -- exact integer
                    // sum, so the result does not depend on the lane order; the tap update hides its latency
                    tmax = fmaxf(tmax, fmaxf(fabsf(pqr), fabsf(pqi)));
                    const int sr = __reduce_add_sync(0xffffffffu, __float2int_rn(pqr * fx_scale));
                    const int si = __reduce_add_sync(0xffffffffu, __float2int_rn(pqi * fx_scale));
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        f32x2 xr = XR[(u + q) % B], xi = XI[(u + q) % B];
                        if (q >= NP - NMASK) {
                            xr = mul2(xr, MK[q - (NP - NMASK)]);
                            xi = mul2(xi, MK[q - (NP - NMASK)]);
                        }
                        PR[q] = fma2_bcast(crp, xr, PR[q]);
                        PR[q] = fma2_bcast(cip, xi, PR[q]);
                        PI[q] = fma2_bcast(cip, xr, PI[q]);
                        PI[q] = fma2_bcast(ncr, xi, PI[q]);
                    }
                    // Q + c G with the Gram term formed first: only one FMA follows the reduction
                    qr = fmaf((float)sr, fx_inv, fmaf(-cip, Gi.y, crp * Gi.x));
                    qi = fmaf((float)si, fx_inv, fmaf(cip, Gi.x, crp * Gi.y));
                } else {
                    constexpr int NHOP = LPS == 8 ? 3 : (LPS == 16 ? 4 : 5);
                    if constexpr (LPS == 32) tmax = fmaxf(tmax, fmaxf(fabsf(pqr), fabsf(pqi)));
#pragma unroll
                    for (int h = 0; h < NHOP; h++) {
                        const int m = LPS >> (h + 1);
                        const float tr = __shfl_xor_sync(0xffffffffu, qr, m);
                        const float ti = __shfl_xor_sync(0xffffffffu, qi, m);
#pragma unroll
                        for (int q = (NP * h) / NHOP; q < (NP * (h + 1)) / NHOP; q++) {
                            f32x2 xr = XR[(u + q) % B], xi = XI[(u + q) % B];
                            if (q >= NP - NMASK) {   // taps past ntaps stay exactly zero
                                xr = mul2(xr, MK[q - (NP - NMASK)]);
                                xi = mul2(xi, MK[q - (NP - NMASK)]);
                            }
                            PR[q] = fma2_bcast(crp, xr, PR[q]);
                            PR[q] = fma2_bcast(cip, xi, PR[q]);
                            PI[q] = fma2_bcast(cip, xr, PI[q]);
                            PI[q] = fma2_bcast(ncr, xi, PI[q]);
                        }
                        qr += tr;
                        qi += ti;
                    }
                }
                // ---- 2. y = Q + c_{il-1} G -> error -> c_il, interleaved with the partial dot Q_{il+1} ---------
                const float yr = FIXED ? qr : fmaf(-cip, Gi.y, fmaf(crp, Gi.x, qr));
                const float yi = FIXED ? qi : fmaf(cip, Gi.x, fmaf(crp, Gi.y, qi));
                f32x2 a1 = 0ull, a2 = 0ull, b1 = 0ull, b2 = 0ull;
#pragma unroll
                for (int q = 0; q < NP; q++) {
                    const f32x2 xr = XR[(u + 2 + q) % B], xi = XI[(u + 2 + q) % B];
                    a1 = fma2(xr, PR[q], a1);
                    a2 = fma2(xi, PI[q], a2);
                    b1 = fma2(xr, PI[q], b1);
                    b2 = fma2(xi, PR[q], b2);
                }
                const long long i = i0 + il;
                const float2 e = err_fast<METHOD, LPS, GRID>(p.method, make_float2(yr, yi), ec, mysyms, p.K, gsyms,
                                                  live ? i : 0, gl);
                // every lane of the group stores the same value; symbols past n are never copied out
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(errs_addr + 8u * il), "f"(e.x), "f"(e.y)
                             : "memory");
                const float mu_l = live ? mu : 0.f;     // symbols past the end of the stream: zero step
                crp = mu_l * e.x;
                cip = mu_l * e.y;
                if constexpr (ADAPT) {                  // after the update of symbol i > 0 of an iteration (:171-172)
                    mu = adapt_step_sel(mu, e, eprev, live && i > 0);
                    eprev = live ? e : eprev;
                }
                const float2 sa = unpack2(sub2(a1, a2)), sb = unpack2(add2(b1, b2));
                pqr = sa.x + sa.y;
                pqi = sb.x + sb.y;
            }
        }
        };   // tile_pass
        if constexpr (LPS == 32) {
            // Block-floating scale of the integer reduction.  The reference does not normalise its input
            // (equalise_signal trains on whatever amplitude it is given), so the scale follows the data: the largest
            // lane partial of the PREVIOUS tile is mapped to [2^22, 2^23), which leaves 8x room before 32 lanes could
            // wrap an int32 and a quantum at least as fine as the fp32 rounding of the partials themselves.  The
            // pass records what it really saw; if that broke the bound (an amplitude step of more than 8x inside one
            // tile) the tile is redone from its saved start state with the shuffle reduction, as is the first tile
            // of a launch and any tile after a zero / non-finite one.
            const float prev = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fx_prev)));
            const int ex = (int)(__float_as_uint(prev) >> 23);               // biased exponent of the previous maximum
            const bool usable = ex >= 32 && ex <= 222;                        // finite, non-zero, scale representable
            bool redo = !usable;
            if (usable) {
                const float fx_scale = __uint_as_float((uint32_t)(127 + 22 + 127 - ex) << 23);   // 2^(22 - floor(log2 prev))
                const float fx_inv = __uint_as_float((uint32_t)(ex - 22) << 23);
                f32x2 sPR[NP], sPI[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) sPR[q] = PR[q], sPI[q] = PI[q];
                const float scr = crp, sci = cip, spr = pqr, spi = pqi, smu = mu;
                const float2 sep = eprev;
                tile_pass(std::true_type{}, fx_scale, fx_inv);
                const float seen = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(tmax)));
                redo = !(seen * fx_scale < 67108864.f);                       // 2^26 per lane; NaN -> redo
                if (redo) {
#pragma unroll
                    for (int q = 0; q < NP; q++) PR[q] = sPR[q], PI[q] = sPI[q];
                    crp = scr, cip = sci, pqr = spr, pqi = spi, mu = smu;
                    eprev = sep;
                    tmax = 0.f;
                }
            }
            if (redo) tile_pass(std::false_type{}, 0.f, 0.f);
            fx_prev = tmax;
        } else {
            tile_pass(std::false_type{}, 0.f, 0.f);
        }
        if (tl == ntiles_it - 1) {
            // end of a training iteration: apply the pending update c_{T-1} conj(X_{T-1}).  The loop ended
            // on a chunk boundary, so that window sits in slots 0 .. NP-1 (pairs il_end .. il_end+NP-1);
            // if the last chunk ran past the end of the stream the pending step is already zero.
            const float ncr = -crp;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                f32x2 xr = XR[q % B], xi = XI[q % B];
                if (q >= NP - NMASK) {
                    xr = mul2(xr, MK[q - (NP - NMASK)]);
                    xi = mul2(xi, MK[q - (NP - NMASK)]);
                }
                PR[q] = fma2_bcast(crp, xr, PR[q]);
                PR[q] = fma2_bcast(cip, xi, PR[q]);
                PI[q] = fma2_bcast(cip, xr, PI[q]);
                PI[q] = fma2_bcast(ncr, xi, PI[q]);
            }
            crp = cip = 0.f;
        }
        __syncwarp();
        if (p.err && act) {
            float2 *eg = p.err + ((long long)seg * p.nmodes + mode) * (p.TrSyms * p.Niter) + it * p.TrSyms + i0;
            for (int c = gl; c < n; c += LPS) eg[c] = errs[grp * (g.tile_syms + 1) + c];
        }
        __syncwarp();
    }
    if (act) {
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const float2 wr = unpack2(PR[q]), wi = unpack2(PI[q]);
            if (t0 + 2 * q < p.ntaps) wg[t0 + 2 * q] = make_float2(wr.x, wi.x);
            if (t0 + 2 * q + 1 < p.ntaps) wg[t0 + 2 * q + 1] = make_float2(wr.y, wi.y);
        }
        if (ADAPT && gl == 0) p.mu[stream] = mu;
    }
}

// geometry of the look-ahead layout; returns NQ or 0 if the shape does not fit
template <int LPS>
static int la_geometry(const TrainParams<float> &p, FastGeom &g, size_t &smem)
{
    constexpr int GPW = 32 / LPS;
    if (LPS % p.nmodes) return 0;
    g.lpp = LPS / p.nmodes;
    int nq = (p.ntaps + g.lpp - 1) / g.lpp;
    nq += nq & 1;
    // instantiated shapes: 6 / 12 taps per lane with 8 lanes per stream (ntaps 21 / 45 dual polarisation, ...),
    // 2 / 4 with one stream per warp (the latency layout)
    if (LPS == 32 ? (nq != 2 && nq != 4) : (nq != 6 && nq != 12)) return 0;
    const int B = nq / 2 + 2;
    // per-tile costs (loader set-up, window preload) amortised over 128 symbols; option LA_TILE overrides the target
    // for occupancy experiments: a shorter tile is a smaller shared-memory slice, so more warps fit an SM (measured:
    // shorter tiles are slower at every occupancy, profiles/README.md)
    const int tile_opt = option_int(OPT_LA_TILE, 0);
    const int tile_target = tile_opt >= B ? tile_opt : 128;
    g.tile_syms = (tile_target / B) * B;
    g.pitch = 2 * (g.tile_syms + 1) + g.lpp * nq;
    // Bank placement of the window loads (ld.shared.b64, one pair per lane): the lanes of a lane group that read input
    // polarisation k sit k * 2 * pitch floats apart, on top of their nq-float tap stride.  With 2 * pitch = 16 (mod 32)
    // the second polarisation's lanes fall into the banks the first one leaves free ({0, 12, 24, 4} + 16 for nq = 12,
    // {0, 6, 12, 18} + 16 for nq = 6); the un-padded 2 * pitch = 4 (mod 32) put lane (k=1, 0) on lane (k=0, 3)'s banks:
    // 32 % of the kernel's shared-memory wavefronts were conflict replays (profiles/r01_ncu_full_summary.txt).
    if (LPS == 8 && p.nmodes == 2) g.pitch += (8 - (g.pitch & 15)) & 15;
    g.nslots = (GPW % p.nsel == 0) ? GPW / p.nsel : (GPW / p.nsel + 2 < GPW ? GPW / p.nsel + 2 : GPW);
    if (g.nslots < 1) g.nslots = 1;
    if (2 * g.tile_syms + g.lpp * nq + 2 > 32 * GRAM_CH) return 0;
    const size_t shared_len = std::max((size_t)g.nslots * gram_sum_len(g.tile_syms, p.ntaps), (size_t)GPW * (g.tile_syms + 1));
    smem = ((size_t)2 * g.nslots * p.nmodes * g.pitch + (size_t)g.nslots * g.tile_syms + shared_len +
            (size_t)GPW * p.nsym_pitch) * sizeof(float2);
    if (smem > 56 * 1024) return 0;   // four warp slices per CTA must fit 227 kB
    return nq;
}

template <int LPS, int NQ, int METHOD, int NMASK, int GRID, bool ADAPT>
static int launch_la_one(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    constexpr int GPW = 32 / LPS;
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(train_la_kernel<LPS, NQ, METHOD, NMASK, GRID, ADAPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const long long nblk = (p.nstreams + GPW - 1) / GPW;
    const size_t wsm = (smem + 15) & ~(size_t)15;   // per-warp slice, 16-byte aligned
    const int wpb = (int)(train_warps_per_cta(nblk) < nblk ? train_warps_per_cta(nblk) : nblk);   // never more warps (or shared memory) than streams need
    // multi-warp CTAs (small launches) ask for more than half of an SM's shared memory so that ONE of them fits an
    // SM and concurrent launches spread over the machine (eq_train_fast.cuh, TRAIN_WPB)
    const size_t dyn = wpb > 1 ? std::max((size_t)wpb * wsm, (size_t)116 * 1024) : wsm;
    train_la_kernel<LPS, NQ, METHOD, NMASK, GRID, ADAPT><<<(unsigned)((nblk + wpb - 1) / wpb), 32 * wpb, dyn, st>>>(p, g, (int)wsm);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

template <int LPS, int NQ, int METHOD, int NMASK, bool ADAPT>
static int launch_la(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    if constexpr (METHOD == QB_SBD || METHOD == QB_DD || METHOD == QB_MDDMA) {
        if (p.nsym_pitch > p.nsym_smem) {    // grid scratch staged: one launch per decision kind (see the kernel)
            const int side = grid_side(p.K);
            const int rc = side * side == p.K ? launch_la_one<LPS, NQ, METHOD, NMASK, 1, ADAPT>(p, g, smem, st)
                                              : launch_la_one<LPS, NQ, METHOD, NMASK, 2, ADAPT>(p, g, smem, st);
            if (rc != QB_OK) return rc;
        }
        return launch_la_one<LPS, NQ, METHOD, NMASK, 0, ADAPT>(p, g, smem, st);
    } else {
        return launch_la_one<LPS, NQ, METHOD, NMASK, -1, ADAPT>(p, g, smem, st);
    }
}

template <int LPS, int NQ, int METHOD, bool ADAPT>
static int launch_la_pad(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    constexpr int NP = NQ / 2;
    constexpr int NMLO = NP >= 2 ? 2 : NP;
    const int valid_last = p.ntaps - (g.lpp - 1) * NQ;
    const int need = valid_last <= 0 ? NP : NP - valid_last / 2;
    if (need == 0) return launch_la<LPS, NQ, METHOD, 0, ADAPT>(p, g, smem, st);
    if (need <= NMLO) return launch_la<LPS, NQ, METHOD, NMLO, ADAPT>(p, g, smem, st);
    return launch_la<LPS, NQ, METHOD, NP, ADAPT>(p, g, smem, st);
}

template <int LPS, int NQ, bool ADAPT>
static int launch_la_method(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    switch (p.method) {
    case QB_CMA:
    case QB_SGNCMA:
        return launch_la_pad<LPS, NQ, QB_CMA, ADAPT>(p, g, smem, st);
    case QB_MCMA:
        return launch_la_pad<LPS, NQ, QB_MCMA, ADAPT>(p, g, smem, st);
    case QB_SBD:
        return launch_la_pad<LPS, NQ, QB_SBD, ADAPT>(p, g, smem, st);
    case QB_DD:
        return launch_la_pad<LPS, NQ, QB_DD, ADAPT>(p, g, smem, st);
    case QB_MDDMA:
        return launch_la_pad<LPS, NQ, QB_MDDMA, ADAPT>(p, g, smem, st);
    case QB_RDE:
        if ((p.K + 1) / 2 > MAXC) return launch_la_pad<LPS, NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
        if (p.K - (p.K + 1) / 2 <= 3) return launch_la_pad<LPS, NQ, METHOD_RDE3, ADAPT>(p, g, smem, st);
        return launch_la_pad<LPS, NQ, QB_RDE, ADAPT>(p, g, smem, st);
    case QB_MRDE:
        if ((p.K + 1) / 2 > MAXC) return launch_la_pad<LPS, NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
        if (p.K - (p.K + 1) / 2 <= 3) return launch_la_pad<LPS, NQ, METHOD_MRDE3, ADAPT>(p, g, smem, st);
        return launch_la_pad<LPS, NQ, QB_MRDE, ADAPT>(p, g, smem, st);
    default:
        return launch_la_pad<LPS, NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
    }
}

}  // namespace qb
