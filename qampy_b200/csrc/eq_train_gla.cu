// eq_train, generic look-ahead kernel: any dtype (complex64 / complex128), oversampling, number of modes and method,
// warp per stream, nmodes * ntaps <= 128.  It is what complex128 signals -- the reference's default dtype
// (signals.py:658), what its Scripts/*_equalisation.py run in -- and the shapes outside the packed fp32 kernels train
// with.
//
// Same recurrence as train_warp_kernel (eq_train.cu), re-associated like eq_train_la.cuh:
//     y_{i+1} = X_{i+1} . W_i  +  c_i G_{i+1} ,      G_{i+1} = X_{i+1} . conj(X_i)        (exact algebra)
// so that the tap dot of symbol i+1 and its five-hop shuffle reduction (64-bit shuffles for double: the longest chain
// of the direct form) no longer wait for c_i: per symbol the warp issues the reduction of Q_i interleaved with the tap
// update W_{i-1} + c_{i-1} conj(X_{i-1}), then  y_i = Q_i + c_{i-1} G_i -> error function -> c_i  next to the partial
// dot of X_{i+1}.  G is a property of the signal alone: per staged tile, running sums of the lag-os products (in the
// signal's precision), two loads and a subtraction per symbol.  The windows of symbols i-1, i, i+1 live in registers
// (NQ <= 4 taps per lane), so a tile is read once per symbol.  Adaptive step size as in the reference (:12-16,
// :171-172): known before e_i is, off the chain.  Results equal the direct form to rounding (1e-12 relative in
// complex128; the golden / oracle tests hold both to the same bounds).
#include <stdlib.h>

#include "eq_train_common.cuh"

namespace qb {

// METHOD >= 0: the error function is compiled in (the non-decision methods; no jump table per symbol); -1: run time
template <typename T, int NQ, int METHOD>
__global__ void __launch_bounds__(32) train_gla_kernel(TrainParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int stream = blockIdx.x;
    const int s = stream / p.nsel, jsel = stream % p.nsel;
    const int mode = p.modes.m[jsel];
    const int Ktot = p.nmodes * p.ntaps;
    const int os = p.os;
    const int tile_samples = p.nmodes * p.tile_pitch;
    const int nprod_max = p.tile_syms * os + p.ntaps;         // lag products of a full tile (+ one symbol of look-ahead)

    cx<T> *tile0 = reinterpret_cast<cx<T> *>(smem_raw);
    cx<T> *tile1 = tile0 + tile_samples;
    cx<T> *errs = tile1 + tile_samples;           // [tile_syms]
    cx<T> *gval = errs + p.tile_syms;             // [tile_syms + 1]  G of the tile's symbols
    cx<T> *ssum = gval + p.tile_syms + 1;         // [nprod_max + 1]  running sums of the lag products
    cx<T> *syms = ssum + nprod_max + 1;           // [nsym_smem]

    const cx<T> *Eseg = p.E + (long long)s * p.seg_stride;
    const cx<T> *gsyms = p.symbols + (long long)mode * p.K;
    for (int c = lane; c < p.nsym_smem; c += 32) syms[c] = gsyms[c];

    cx<T> *wg = p.wx + ((long long)s * p.nmodes + mode) * (long long)Ktot;
    T wr[NQ], wi[NQ];
    int off[NQ];
    bool val[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int c = lane + 32 * q;
        val[q] = c < Ktot;
        const int k = val[q] ? c / p.ntaps : 0, t = val[q] ? c % p.ntaps : 0;
        off[q] = k * p.tile_pitch + t;
        const cx<T> w = val[q] ? wg[c] : make_cx<T>(0, 0);
        wr[q] = w.x;
        wi[q] = w.y;
    }
    T mu = p.mu[stream];
    cx<T> eprev = make_cx<T>(0, 0);

    const long long ntiles_it = (p.TrSyms + p.tile_syms - 1) / p.tile_syms;
    const long long ntiles = ntiles_it * p.Niter;
    const long long Lread = (p.TrSyms - 1) * os + p.ntaps;     // samples of a row the caller guarantees

    // stage tile g: samples [(i0 - 1) os, (i0 + n) os + ntaps) of every row -- one symbol before the tile (for G of its
    // first symbol) and one after (the look-ahead dot of its last) -- zero outside [0, Lread)
    auto load_tile = [&](long long g, cx<T> *buf) {
        const long long i0 = (g % ntiles_it) * p.tile_syms;
        const long long s0 = (i0 - 1) * os;
        for (int k = 0; k < p.nmodes; k++) {
            const cx<T> *src = Eseg + (long long)k * p.row_stride + s0;
            cx<T> *dst = buf + k * p.tile_pitch;
            for (int c = lane; c < p.tile_pitch; c += 32) {
                const long long m = s0 + c;
                if (m >= 0 && m < Lread) cp_async<sizeof(cx<T>)>(dst + c, src + c);
                else dst[c] = make_cx<T>(0, 0);
            }
        }
        cp_async_commit();
    };

    T cr = 0, ci = 0;               // c_{i-1} = mu e_{i-1}: the update still to be applied
    T pr = 0, pi = 0;               // this lane's partial of Q_i = X_i . W_{i-1}
    cx<T> xp[NQ], xc[NQ];           // windows of symbols i-1 and i (this lane's taps)
#pragma unroll
    for (int q = 0; q < NQ; q++) xp[q] = xc[q] = make_cx<T>(0, 0);

    if (ntiles > 0) load_tile(0, tile0);
    for (long long g = 0; g < ntiles; g++) {
        cx<T> *cur = (g & 1) ? tile1 : tile0;
        if (g + 1 < ntiles) {
            load_tile(g + 1, (g & 1) ? tile0 : tile1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const long long it = g / ntiles_it, b = g % ntiles_it;
        const long long i0 = b * p.tile_syms;
        const int n = (int)min((long long)p.tile_syms, p.TrSyms - i0);

        // ---- G of the tile: lag products p[m] = sum_k x_k[m] conj(x_k[m - os]), m = os .. os + nprod - 1 (staged
        //      indices), running sum S[j] = sum of the first j of them, G_il = S[il os + ntaps] - S[il os] -------------
        {
            const int nprod = (n - 1) * os + p.ntaps;
            const int ch = (nprod + 31) / 32;                   // products per lane, consecutive
            const int m0 = os + lane * ch;
            T sr = 0, si = 0;
            for (int j = 0; j < ch; j++) {
                const int m = m0 + j;
                T ar = 0, ai = 0;
                if (m < os + nprod) {
                    for (int k = 0; k < p.nmodes; k++) {
                        const cx<T> a = cur[k * p.tile_pitch + m], c2 = cur[k * p.tile_pitch + m - os];
                        ar = fma(a.x, c2.x, fma(a.y, c2.y, ar));          // a conj(c2)
                        ai = fma(a.y, c2.x, fma(-a.x, c2.y, ai));
                    }
                }
                sr += ar;
                si += ai;
                if (m - os + 1 <= nprod) ssum[m - os + 1] = make_cx<T>(sr, si);   // local inclusive sums for now
            }
            T orr = sr, oi = si;                                // exclusive scan of the lane totals
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T a = shfl_up(orr, d), c2 = shfl_up(oi, d);
                if (lane >= d) {
                    orr += a;
                    oi += c2;
                }
            }
            orr -= sr;
            oi -= si;
            __syncwarp();
            for (int j = 0; j < ch; j++) {
                const int idx = lane * ch + j + 1;
                if (idx <= nprod) {
                    const cx<T> v = ssum[idx];
                    ssum[idx] = make_cx<T>(v.x + orr, v.y + oi);
                }
            }
            if (lane == 0) ssum[0] = make_cx<T>(0, 0);
            __syncwarp();
            for (int il = lane; il < n; il += 32) {
                const cx<T> hi = ssum[il * os + p.ntaps], lo = ssum[il * os];
                gval[il] = make_cx<T>(hi.x - lo.x, hi.y - lo.y);
            }
            __syncwarp();
        }

        if (b == 0) {
            // start of a training iteration: nothing pending, window of symbol 0 and Q_0 = X_0 . W_0 directly
            cr = ci = 0;
            pr = pi = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                xp[q] = make_cx<T>(0, 0);
                xc[q] = cur[os + off[q]];                       // staged from one symbol before the tile
                if (!val[q]) xc[q] = make_cx<T>(0, 0);          // lanes past the last tap carry zeros: no branches below
                pr = fma(xc[q].x, wr[q], pr);
                pr = fma(-xc[q].y, wi[q], pr);
                pi = fma(xc[q].x, wi[q], pi);
                pi = fma(xc[q].y, wr[q], pi);
            }
        }
        for (int il = 0; il < n; il++) {
            // ---- 1. all-reduce of the partial Q_il, hops interleaved with W += c_{il-1} conj(X_{il-1}) -------------
            T qr = pr, qi = pi;
#pragma unroll
            for (int h = 0; h < 5; h++) {
                const int m = 16 >> h;
                const T tr = shfl_xor(qr, m), ti = shfl_xor(qi, m);
                if (h < NQ) {                                   // one tap of the update per hop (NQ <= 4)
                    const int q = h;                            // (a lane without this tap holds x = 0: w stays 0)
                    wr[q] = fma(cr, xp[q].x, wr[q]);
                    wr[q] = fma(ci, xp[q].y, wr[q]);
                    wi[q] = fma(ci, xp[q].x, wi[q]);
                    wi[q] = fma(-cr, xp[q].y, wi[q]);
                }
                qr += tr;
                qi += ti;
            }
            // ---- 2. y = Q + c_{il-1} G -> error -> c_il, next to the partial dot of X_{il+1} with the updated taps ---
            const cx<T> G = gval[il];
            const T yr = fma(-ci, G.y, fma(cr, G.x, qr));
            const T yi = fma(ci, G.x, fma(cr, G.y, qi));
            const cx<T> *xb = cur + (il + 2) * os;              // window of symbol il + 1 (tile staged from symbol -1)
            cx<T> xn[NQ];
            T nr = 0, ni = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                xn[q] = xb[off[q]];
                if (!val[q]) xn[q] = make_cx<T>(0, 0);
                nr = fma(xn[q].x, wr[q], nr);
                nr = fma(-xn[q].y, wi[q], nr);
                ni = fma(xn[q].x, wi[q], ni);
                ni = fma(xn[q].y, wr[q], ni);
            }
            const long long i = i0 + il;
            const cx<T> e = error_fct<T>(METHOD >= 0 ? METHOD : p.method, make_cx<T>(yr, yi), syms, p.K, gsyms, i, lane);
            if (lane == 0) errs[il] = e;
            cr = mu * e.x;
            ci = mu * e.y;
            if (p.adaptive && i > 0) mu = adapt_step<T>(mu, e, eprev);
            eprev = e;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                xp[q] = xc[q];
                xc[q] = xn[q];
            }
            pr = nr;
            pi = ni;
        }
        if (b == ntiles_it - 1) {
            // end of a training iteration: the pending update c_{T-1} conj(X_{T-1}); xp holds that window now
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                wr[q] = fma(cr, xp[q].x, wr[q]);
                wr[q] = fma(ci, xp[q].y, wr[q]);
                wi[q] = fma(ci, xp[q].x, wi[q]);
                wi[q] = fma(-cr, xp[q].y, wi[q]);
            }
            cr = ci = 0;
        }
        __syncwarp();
        if (p.err) {
            cx<T> *eg = p.err + ((long long)s * p.nmodes + mode) * (p.TrSyms * p.Niter) + it * p.TrSyms + i0;
            for (int c = lane; c < n; c += 32) eg[c] = errs[c];
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int c = lane + 32 * q;
        if (c < Ktot) wg[c] = make_cx<T>(wr[q], wi[q]);
    }
    if (lane == 0) p.mu[stream] = mu;
}

template <typename T, int NQ, int METHOD>
static int launch_gla_m(const TrainParams<T> &p, size_t smem, cudaStream_t st)
{
    QB_CUDA_CHECK(cudaFuncSetAttribute(train_gla_kernel<T, NQ, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    train_gla_kernel<T, NQ, METHOD><<<(unsigned)p.nstreams, 32, smem, st>>>(p);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

template <typename T, int NQ>
static int launch_gla_nq(const TrainParams<T> &p, size_t smem, cudaStream_t st)
{
    switch (p.method) {
    case QB_CMA:
    case QB_SGNCMA: return launch_gla_m<T, NQ, QB_CMA>(p, smem, st);
    case QB_MCMA: return launch_gla_m<T, NQ, QB_MCMA>(p, smem, st);
    case QB_RDE: return launch_gla_m<T, NQ, QB_RDE>(p, smem, st);
    case QB_MRDE: return launch_gla_m<T, NQ, QB_MRDE>(p, smem, st);
    default: return launch_gla_m<T, NQ, -1>(p, smem, st);
    }
}

// Returns 1 if launched, 0 if the shape is outside this kernel (caller keeps the direct form), < 0 on error.
// Option TRAIN_GLA = 0 (qb_set_option) keeps the direct form.
template <typename T>
int train_gla_try(TrainParams<T> p, cudaStream_t st)
{
    if (option_char(OPT_TRAIN_GLA) == '0') return 0;
    if (p.method >= QB_CMA_REAL) return 0;                     // real-valued methods: REAL instantiation of the direct form
    const int Ktot = p.nmodes * p.ntaps;
    const int nq = (Ktot + 31) / 32;
    if (nq > 4) return 0;
    // tile: as many symbols as fit ~16 kB of double-buffered samples, at most 128
    int ts = 128;
    size_t smem = 0;
    for (; ts >= 8; ts >>= 1) {
        const int pitch = (ts + 1) * p.os + p.ntaps;
        smem = ((size_t)2 * p.nmodes * pitch + ts + (ts + 1) + ((size_t)ts * p.os + p.ntaps + 1) + p.nsym_smem) * sizeof(cx<T>);
        if (smem <= 24 * 1024 || (ts == 8 && smem <= 200 * 1024)) break;
    }
    if (ts < 8) return 0;
    p.tile_syms = ts;
    p.tile_pitch = (ts + 1) * p.os + p.ntaps;
    int rc;
    switch (nq) {
    case 1: rc = launch_gla_nq<T, 1>(p, smem, st); break;
    case 2: rc = launch_gla_nq<T, 2>(p, smem, st); break;
    case 3: rc = launch_gla_nq<T, 3>(p, smem, st); break;
    default: rc = launch_gla_nq<T, 4>(p, smem, st); break;
    }
    return rc == QB_OK ? 1 : rc;
}

template int train_gla_try<float>(TrainParams<float> p, cudaStream_t st);
template int train_gla_try<double>(TrainParams<double> p, cudaStream_t st);

}  // namespace qb
