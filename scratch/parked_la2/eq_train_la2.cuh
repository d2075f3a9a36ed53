// eq_train, ONE stream per warp, look-ahead of depth TWO (complex64, os = 2): the single-stream kernel for calls
// whose time is the serial depth of a stream -- the reference's own call shape (train_equaliser on ONE capture,
// pythran_equalisation.py:162-172) and the tap acquisition of the segmented receiver.
//
// The depth-one form of eq_train_la.cuh,  y_{i+1} = X_{i+1} . W_i + c_i G1_{i+1},  still has the tap update ->
// tap dot -> lane reduction of symbol i+1 waiting for c_{i-1}: in the one-stream layout that chain (two conversions
// and a REDUX among it) is what a lone warp spends most of its time waiting for (profiles/r02_l32_hot_loop.txt).
// Substituting one more update,
//     y_i = X_i . V_i  +  c_{i-2} G2_i  +  c_{i-1} G1_i ,     V_i = taps with the updates of symbols <= i-3 applied,
//     G1_i = X_i . conj(X_{i-1}),  G2_i = X_i . conj(X_{i-2})                                   (exact algebra)
// the reduction of X_{i+1} . V_{i+1} depends on c_{i-2} only: it is issued at the TOP of iteration i and has the whole
// iteration to complete.  What is left on the dependent chain per symbol is  Q + c G -> error function -> c.
// G1 / G2 are properties of the signal alone (lag-2 and lag-4 autocorrelations of the window), formed per staged tile
// from running sums of lag products like the depth-one kernel does for G1.
//
// Per symbol i the warp issues, in this order:
//   1. W += c_{i-2} conj(X_{i-2});  partial dot of X_{i+1} with the updated taps;  scale, convert, REDUX (-> iteration i+1)
//   2. y_i = Q_i (the REDUX issued one iteration ago) + c_{i-2} G2_i + c_{i-1} G1_i;  error function;  c_i = mu_i e_i
// Everything else (tile staging with cp.async, block-floating scale of the integer reduction with a checked shuffle
// fallback, adaptive step size off the chain, padding masks, error functions) is eq_train_la.cuh's; results equal the
// direct recurrence to rounding and are held to the same 1e-5 against the oracle by the same tests.
#pragma once
#include "eq_train_la.cuh"

namespace qb {

// running sums of lag products of one staged tile (ONE slot): G[il] = sum_{t < ntaps} x[m0 + 2 il + t] conj(x[m0 + 2 il + t - LAG])
// with m0 = 4 (tiles are staged from two pairs before their first symbol); out[il * ostride] receives the value.
template <int LAG>
__device__ __forceinline__ void tile_gram_lag(const float *base, int nmodes, int pitch, int tile_syms, int ntaps,
                                              float2 *Ss, float2 *out, int ostride, int lane)
{
    const int slen = gram_sum_len(tile_syms, ntaps);
    const int row_floats = 2 * pitch;
    const int mb = 4 + lane * GRAM_CH;
    float pr[GRAM_CH], pi[GRAM_CH];
#pragma unroll
    for (int j = 0; j < GRAM_CH; j++) pr[j] = pi[j] = 0.f;
    for (int kk = 0; kk < nmodes; kk++) {
        const float *re = base + kk * row_floats + mb - LAG, *im = re + pitch;
        float xr[GRAM_CH + LAG], xi[GRAM_CH + LAG];
#pragma unroll
        for (int j = 0; j < GRAM_CH + LAG; j++) {
            xr[j] = re[j];
            xi[j] = im[j];
        }
#pragma unroll
        for (int j = 0; j < GRAM_CH; j++) {   // x[m] * conj(x[m - LAG])
            pr[j] = fmaf(xr[j + LAG], xr[j], fmaf(xi[j + LAG], xi[j], pr[j]));
            pi[j] = fmaf(xi[j + LAG], xr[j], fmaf(-xr[j + LAG], xi[j], pi[j]));
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
#pragma unroll
    for (int j = 0; j < GRAM_CH; j++) {      // S[k] = sum of the products m = 4 .. k + 3
        const int k = mb + j - 3;
        if (k < slen) Ss[k] = make_float2(off.x + loc[j].x, off.y + loc[j].y);
    }
    __syncwarp();
    for (int il = lane; il < tile_syms; il += 32) {
        const float2 hi = Ss[2 * il + ntaps], lo = Ss[2 * il];
        out[il * ostride] = make_float2(hi.x - lo.x, hi.y - lo.y);
    }
    __syncwarp();
}

template <int NQ, int METHOD, int NMASK, int GRID = -1, bool ADAPT = false>
__global__ void __launch_bounds__(32 * TRAIN_WPB) train_la2_kernel(TrainParams<float> p, FastGeom g, int warp_smem)
{
    static_assert(NQ % 2 == 0, "NQ must be even (os = 2: the window moves by one pair per symbol)");
    constexpr int NP = NQ / 2;     // pairs per lane
    constexpr int B = NP + 3;      // circular pair window: symbols i-2 .. i+1; also symbols per unrolled chunk
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int wib = threadIdx.x >> 5;
    const long long stream = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (stream >= p.nstreams) return;
    unsigned char *smem_raw = smem_all + (size_t)wib * warp_smem;
    const int lane = threadIdx.x & 31;
    const long long seg = stream / p.nsel;
    const int mode = p.modes.m[(int)(stream % p.nsel)];

    // tiles: [2][nmodes][2 planes][pitch] floats; Gram values [tile_syms] (G1, G2 interleaved); running sums / errors
    const int row_floats = 2 * g.pitch, slot_floats = p.nmodes * row_floats;
    float *tile0 = reinterpret_cast<float *>(smem_raw);
    float *tile1 = tile0 + slot_floats;
    float4 *gbuf = reinterpret_cast<float4 *>(tile1 + slot_floats);
    float2 *gsum = reinterpret_cast<float2 *>(gbuf + g.tile_syms);
    float2 *errs = gsum;                                    // the running sums are dead once G is formed
    const int shared_len = max(gram_sum_len(g.tile_syms, p.ntaps), g.tile_syms);
    float2 *mysyms = gsum + shared_len;

    const float2 *gsyms = p.symbols + (long long)mode * p.K;
    for (int c = lane; c < p.nsym_smem; c += 32) mysyms[c] = gsyms[c];

    const int k = lane / g.lpp, t0 = (lane % g.lpp) * NQ;
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
    const uint32_t errs_addr = smem_u32(errs);
    __syncwarp();
    ErrConst ec = load_err_const<METHOD>(mysyms, p.nsym_smem);
    if (p.nsym_pitch > p.nsym_smem)   // searched alphabet: is it a square grid? (uniform)
        detect_grid<32>(ec, mysyms, p.nsym_smem, reinterpret_cast<float *>(mysyms + p.nsym_smem), lane);
    if (GRID >= 0) {
        if ((ec.gn != 0) != (GRID == 1)) return;     // the other instantiation takes this stream
    }

    const long long ntiles_it = (p.TrSyms + g.tile_syms - 1) / g.tile_syms;
    const long long ntiles = ntiles_it * p.Niter;
    const long long Lread = (p.TrSyms - 1) * 2 + p.ntaps;  // samples of a row the caller guarantees

    // stage tile gt: samples [2*i0 - 4, 2*i0 - 4 + pitch) of every row (zero outside [0, Lread))
    auto load_tile = [&](long long gt, float *buf) {
        const long long i0 = (gt % ntiles_it) * g.tile_syms;
        const long long s0 = i0 * 2 - 4;
        const int lo = s0 < 0 ? (int)(-s0) : 0;
        const int hi = (int)max((long long)lo, min((long long)g.pitch, Lread - s0));
        for (int kk = 0; kk < p.nmodes; kk++) {
            const float *src = reinterpret_cast<const float *>(p.E + seg * p.seg_stride + (long long)kk * p.row_stride) + 2 * s0;
            // float c of the staged row: even -> real plane, odd -> imaginary plane (c and lane share parity)
            float *dst = buf + kk * row_floats + (lane & 1) * g.pitch + (lane >> 1);
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
        cp_async_commit();
    };

    float fx_prev = 0.f;            // largest lane partial of the previous tile (0: none yet)
    float c1r = 0.f, c1i = 0.f;     // c_{i-1} = mu e_{i-1}
    float c2r = 0.f, c2i = 0.f;     // c_{i-2}: the update applied in iteration i
    float pqr = 0.f, pqi = 0.f;     // this lane's partial of Q_i = X_i . V_i for the NEXT symbol to be decided

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
        tile_gram_lag<2>(cur, p.nmodes, g.pitch, g.tile_syms, p.ntaps, gsum, reinterpret_cast<float2 *>(gbuf), 2, lane);
        tile_gram_lag<4>(cur, p.nmodes, g.pitch, g.tile_syms, p.ntaps, gsum, reinterpret_cast<float2 *>(gbuf) + 1, 2, lane);
        const long long it = gt / ntiles_it;
        const long long tl = gt % ntiles_it;
        const long long i0 = tl * g.tile_syms;
        const int n = (int)min((long long)g.tile_syms, p.TrSyms - i0);
        const uint32_t xre = smem_u32(cur + k * row_floats + t0);
        const uint32_t xim = xre + 4u * (uint32_t)g.pitch;
        const uint32_t gaddr = smem_u32(gbuf);

        // staged pair j (samples 2j, 2j+1 of the staged row, relative to this lane's first tap) lives in slot j % B;
        // symbol il of the tile has its window in pairs il+2 .. il+NP+1, symbol il-2 in il .. il+NP-1, symbol il+1 in il+3 ..
        f32x2 XR[B], XI[B];
        float tmax = 0.f;
        auto dot_window = [&](int first_slot, float &dr, float &di) {   // sum over this lane's taps, window from slot first_slot
            f32x2 a1 = 0ull, a2 = 0ull, b1 = 0ull, b2 = 0ull;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const f32x2 xr = XR[(first_slot + q) % B], xi = XI[(first_slot + q) % B];
                a1 = fma2(xr, PR[q], a1);
                a2 = fma2(xi, PI[q], a2);
                b1 = fma2(xr, PI[q], b1);
                b2 = fma2(xi, PR[q], b2);
            }
            const float2 sa = unpack2(sub2(a1, a2)), sb = unpack2(add2(b1, b2));
            dr = sa.x + sa.y;
            di = sb.x + sb.y;
        };
        auto update_window = [&](int first_slot, float cr, float ci) {   // W += c conj(X), window from slot first_slot
            const float ncr = -cr;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                f32x2 xr = XR[(first_slot + q) % B], xi = XI[(first_slot + q) % B];
                if (q >= NP - NMASK) {   // taps past ntaps stay exactly zero
                    xr = mul2(xr, MK[q - (NP - NMASK)]);
                    xi = mul2(xi, MK[q - (NP - NMASK)]);
                }
                PR[q] = fma2_bcast(cr, xr, PR[q]);
                PR[q] = fma2_bcast(ci, xi, PR[q]);
                PI[q] = fma2_bcast(ci, xr, PI[q]);
                PI[q] = fma2_bcast(ncr, xi, PI[q]);
            }
        };
        // One pass over the tile's symbols.  FIXED: lane partials summed as block-floating integers by REDUX;
        // !FIXED: shuffle all-reduce, any magnitude (eq_train_la.cuh).
        auto tile_pass = [&](auto fixed_tag, const float fx_scale, const float fx_inv) {
            constexpr bool FIXED = decltype(fixed_tag)::value;
#pragma unroll
            for (int q = 0; q < NP + 2; q++) {
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XR[q]) : "r"(xre + 8u * q));
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XI[q]) : "r"(xim + 8u * q));
            }
            if (tl == 0) {
                // start of a training iteration: nothing pending, Q_0 = X_0 . W_0 directly (pairs 2 .. NP+1)
                c1r = c1i = c2r = c2i = 0.f;
                dot_window(2, pqr, pqi);
            }
            // the reduction of the tile's first symbol: its partial came over from the tile before (or from above)
            auto reduce_issue = [&](float vr, float vi, float &outr, float &outi) {
                tmax = fmaxf(tmax, fmaxf(fabsf(vr), fabsf(vi)));
                if constexpr (FIXED) {
                    const int sr = __reduce_add_sync(0xffffffffu, __float2int_rn(vr * fx_scale));
                    const int si = __reduce_add_sync(0xffffffffu, __float2int_rn(vi * fx_scale));
                    outr = (float)sr * fx_inv;
                    outi = (float)si * fx_inv;
                } else {
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) {
                        vr += __shfl_xor_sync(0xffffffffu, vr, m);
                        vi += __shfl_xor_sync(0xffffffffu, vi, m);
                    }
                    outr = vr;
                    outi = vi;
                }
            };
            float Qr, Qi;           // Q_il, reduced: ready one iteration after its partials were formed
            reduce_issue(pqr, pqi, Qr, Qi);
#pragma unroll 1
            for (int il0 = 0; il0 < n; il0 += B) {
#pragma unroll
                for (int u = 0; u < B; u++) {
                    const int il = il0 + u;
                    const bool live = il < n;
                    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XR[(u + NP + 2) % B]) : "r"(xre + 8u * (il + NP + 2)));
                    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(XI[(u + NP + 2) % B]) : "r"(xim + 8u * (il + NP + 2)));
                    float4 G;   // (G1.re, G1.im, G2.re, G2.im) of symbol il
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(G.x), "=f"(G.y), "=f"(G.z), "=f"(G.w)
                                 : "r"(gaddr + 16u * il));
                    // ---- 1. W += c_{il-2} conj(X_{il-2});  Q_{il+1} = X_{il+1} . W: needs nothing of this iteration ----
                    update_window(u, c2r, c2i);
                    float nr, ni, Qnr, Qni;
                    dot_window(u + 3, nr, ni);
                    reduce_issue(nr, ni, Qnr, Qni);
                    // ---- 2. y = Q + c_{il-2} G2 + c_{il-1} G1 -> error -> c_il ------------------------------------------
                    const float hr = fmaf(-c2i, G.w, fmaf(c2r, G.z, Qr));       // known before c_{il-1} is
                    const float hi = fmaf(c2i, G.z, fmaf(c2r, G.w, Qi));
                    const float yr = fmaf(-c1i, G.y, fmaf(c1r, G.x, hr));
                    const float yi = fmaf(c1i, G.x, fmaf(c1r, G.y, hi));
                    const long long i = i0 + il;
                    const float2 e = err_fast<METHOD, 32, GRID>(p.method, make_float2(yr, yi), ec, mysyms, p.K, gsyms,
                                                                live ? i : 0, lane);
                    // every lane stores the same value; symbols past n are never copied out
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(errs_addr + 8u * il), "f"(e.x), "f"(e.y)
                                 : "memory");
                    const float mu_l = live ? mu : 0.f;     // symbols past the end of the stream: zero step
                    c2r = c1r;
                    c2i = c1i;
                    c1r = mu_l * e.x;
                    c1i = mu_l * e.y;
                    if constexpr (ADAPT) {                  // after the update of symbol i > 0 of an iteration (:171-172)
                        mu = adapt_step_sel(mu, e, eprev, live && i > 0);
                        eprev = live ? e : eprev;
                    }
                    Qr = Qnr;
                    Qi = Qni;
                    pqr = nr;       // what goes over to the next tile is the un-reduced partial (its scale may differ)
                    pqi = ni;
                }
            }
        };   // tile_pass
        {
            // Block-floating scale of the integer reduction: eq_train_la.cuh (same bound, same checked fallback)
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
                const float s1r = c1r, s1i = c1i, s2r = c2r, s2i = c2i, spr = pqr, spi = pqi, smu = mu;
                const float2 sep = eprev;
                tile_pass(std::true_type{}, fx_scale, fx_inv);
                const float seen = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(tmax)));
                redo = !(seen * fx_scale < 67108864.f);                       // 2^26 per lane; NaN -> redo
                if (redo) {
#pragma unroll
                    for (int q = 0; q < NP; q++) PR[q] = sPR[q], PI[q] = sPI[q];
                    c1r = s1r, c1i = s1i, c2r = s2r, c2i = s2i, pqr = spr, pqi = spi, mu = smu;
                    eprev = sep;
                    tmax = 0.f;
                }
            }
            if (redo) tile_pass(std::false_type{}, 0.f, 0.f);
            fx_prev = tmax;
        }
        if (tl == ntiles_it - 1) {
            // end of a training iteration: apply the two pending updates.  The loop ended on a chunk boundary il_end
            // (a multiple of B), so X_{il_end-2} sits in slots 0 .. NP-1 and X_{il_end-1} in slots 1 .. NP; steps of
            // symbols past the end of the stream are zero.
            update_window(0, c2r, c2i);
            update_window(1, c1r, c1i);
            c1r = c1i = c2r = c2i = 0.f;
        }
        __syncwarp();
        if (p.err) {
            float2 *eg = p.err + ((long long)seg * p.nmodes + mode) * (p.TrSyms * p.Niter) + it * p.TrSyms + i0;
            for (int c = lane; c < n; c += 32) eg[c] = errs[c];
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const float2 wr = unpack2(PR[q]), wi = unpack2(PI[q]);
        if (t0 + 2 * q < p.ntaps) wg[t0 + 2 * q] = make_float2(wr.x, wi.x);
        if (t0 + 2 * q + 1 < p.ntaps) wg[t0 + 2 * q + 1] = make_float2(wr.y, wi.y);
    }
    if (ADAPT && lane == 0) p.mu[stream] = mu;
}

// geometry; returns NQ (2 or 4) or 0 if the shape does not fit
static int la2_geometry(const TrainParams<float> &p, FastGeom &g, size_t &smem)
{
    if (32 % p.nmodes) return 0;
    g.lpp = 32 / p.nmodes;
    int nq = (p.ntaps + g.lpp - 1) / g.lpp;
    nq += nq & 1;
    if (nq != 2 && nq != 4) return 0;
    const int B = nq / 2 + 3;
    g.tile_syms = (128 / B) * B;
    g.pitch = 2 * (g.tile_syms + 2) + g.lpp * nq;      // staged from two pairs before the first symbol
    g.nslots = 1;
    if (2 * g.tile_syms + g.lpp * nq + 4 > 32 * GRAM_CH) return 0;
    const size_t shared_len = std::max((size_t)gram_sum_len(g.tile_syms, p.ntaps), (size_t)g.tile_syms);
    smem = ((size_t)2 * p.nmodes * g.pitch + (size_t)2 * g.tile_syms + shared_len + (size_t)p.nsym_pitch) * sizeof(float2);
    if (smem > 56 * 1024) return 0;
    return nq;
}

template <int NQ, int METHOD, int NMASK, int GRID, bool ADAPT>
static int launch_la2_one(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    QB_CUDA_CHECK(cudaFuncSetAttribute(train_la2_kernel<NQ, METHOD, NMASK, GRID, ADAPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const long long nblk = p.nstreams;
    const size_t wsm = (smem + 15) & ~(size_t)15;
    const int wpb = (int)(train_warps_per_cta(nblk) < nblk ? train_warps_per_cta(nblk) : nblk);
    const size_t dyn = wpb > 1 ? std::max((size_t)wpb * wsm, (size_t)116 * 1024) : wsm;
    train_la2_kernel<NQ, METHOD, NMASK, GRID, ADAPT><<<(unsigned)((nblk + wpb - 1) / wpb), 32 * wpb, dyn, st>>>(p, g, (int)wsm);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

template <int NQ, int METHOD, int NMASK, bool ADAPT>
static int launch_la2(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    if constexpr (METHOD == QB_SBD || METHOD == QB_DD) {
        if (p.nsym_pitch > p.nsym_smem) {    // grid scratch staged: one launch per decision kind
            const int rc = launch_la2_one<NQ, METHOD, NMASK, 1, ADAPT>(p, g, smem, st);
            if (rc != QB_OK) return rc;
        }
        return launch_la2_one<NQ, METHOD, NMASK, 0, ADAPT>(p, g, smem, st);
    } else {
        return launch_la2_one<NQ, METHOD, NMASK, -1, ADAPT>(p, g, smem, st);
    }
}

template <int NQ, int METHOD, bool ADAPT>
static int launch_la2_pad(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    constexpr int NP = NQ / 2;
    const int valid_last = p.ntaps - (g.lpp - 1) * NQ;
    const int need = valid_last <= 0 ? NP : NP - valid_last / 2;
    if (need == 0) return launch_la2<NQ, METHOD, 0, ADAPT>(p, g, smem, st);
    return launch_la2<NQ, METHOD, NP, ADAPT>(p, g, smem, st);
}

template <int NQ, bool ADAPT>
static int launch_la2_method(const TrainParams<float> &p, const FastGeom &g, size_t smem, cudaStream_t st)
{
    switch (p.method) {
    case QB_CMA:
    case QB_SGNCMA: return launch_la2_pad<NQ, QB_CMA, ADAPT>(p, g, smem, st);
    case QB_MCMA: return launch_la2_pad<NQ, QB_MCMA, ADAPT>(p, g, smem, st);
    case QB_SBD: return launch_la2_pad<NQ, QB_SBD, ADAPT>(p, g, smem, st);
    case QB_DD: return launch_la2_pad<NQ, QB_DD, ADAPT>(p, g, smem, st);
    case QB_RDE:
        if ((p.K + 1) / 2 > MAXC) return launch_la2_pad<NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
        if (p.K - (p.K + 1) / 2 <= 3) return launch_la2_pad<NQ, METHOD_RDE3, ADAPT>(p, g, smem, st);
        return launch_la2_pad<NQ, QB_RDE, ADAPT>(p, g, smem, st);
    case QB_MRDE:
        if ((p.K + 1) / 2 > MAXC) return launch_la2_pad<NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
        if (p.K - (p.K + 1) / 2 <= 3) return launch_la2_pad<NQ, METHOD_MRDE3, ADAPT>(p, g, smem, st);
        return launch_la2_pad<NQ, QB_MRDE, ADAPT>(p, g, smem, st);
    default: return launch_la2_pad<NQ, METHOD_GENERIC, ADAPT>(p, g, smem, st);
    }
}

}  // namespace qb
