// eq_train: stochastic-gradient training of the MIMO FIR equaliser (replaces train_equaliser,
// qampy/core/equalisation/pythran_equalisation.py:130-173).
//
// For every trained stream (segment s, output mode m), sequentially over iterations and symbols i:
//     X    = E[s, :, i*os : i*os+ntaps]                                  (:166)
//     Xest = sum_k sum_t X[k,t] * wx[s,m,k,t]            (no conjugate,    :24-31, :167)
//     e    = errorfct(Xest, symbols[m], i)                                (:168, :178-231)
//     wx[s,m] += (mu*e) * conj(X)                                         (:170)
//     mu   = adapt_step(mu, e_i, e_{i-1})  if adaptive and i > 0          (:12-16, :171-172)
//
// The recurrence is strictly serial in i; parallelism comes from independent streams
// (segments x modes).  Kernel v1 ("warp per stream"): the nmodes*ntaps taps of a stream live in
// the registers of one warp (NQ per lane), the sample window is read from a double-buffered
// shared-memory tile filled with cp.async while the previous tile is being consumed, the tap dot
// product is an in-register partial sum + butterfly warp-shuffle all-reduce, the error function is
// evaluated redundantly in every lane (decision-directed ones search the alphabet lane-parallel),
// and the tap update is in-register.
#include <stdlib.h>

#include "eq_train_common.cuh"

namespace qb {

int train_fast_try(TrainParams<float> p, cudaStream_t st);  // eq_train_fast.cu
template <typename T>
int train_gla_try(TrainParams<T> p, cudaStream_t st);       // eq_train_gla.cu: generic look-ahead form

// REAL: the real-valued methods (train_equaliser_realvalued, pythran_equalisation.py:80-128) run on data whose
// imaginary parts are exactly zero; the instantiation drops the imaginary half of the tap dot, of the reduction and of
// the update (bit-identical results: the dropped terms are products with +-0).
template <typename T, int NQ, bool REAL>
__global__ void __launch_bounds__(32) train_warp_kernel(TrainParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int stream = blockIdx.x;
    const int s = stream / p.nsel, jsel = stream % p.nsel;
    const int mode = p.modes.m[jsel];
    const int Ktot = p.nmodes * p.ntaps;
    const int tile_samples = p.nmodes * p.tile_pitch;

    cx<T> *tile0 = reinterpret_cast<cx<T> *>(smem_raw);
    cx<T> *tile1 = tile0 + tile_samples;
    cx<T> *errs = tile1 + tile_samples;           // [tile_syms]
    cx<T> *syms = errs + p.tile_syms;             // [nsym_smem]

    const cx<T> *Eseg = p.E + (long long)s * p.seg_stride;
    const cx<T> *gsyms = p.symbols + (long long)mode * p.K;
    for (int c = lane; c < p.nsym_smem; c += 32) syms[c] = gsyms[c];

    cx<T> *wg = p.wx + ((long long)s * p.nmodes + mode) * (long long)Ktot;
    T wr[NQ], wi[NQ];
    int off[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int c = lane + 32 * q;
        const bool valid = c < Ktot;
        const int k = valid ? c / p.ntaps : 0, t = valid ? c % p.ntaps : 0;
        off[q] = k * p.tile_pitch + t;
        const cx<T> w = valid ? wg[c] : make_cx<T>(0, 0);
        wr[q] = w.x;
        wi[q] = w.y;
    }
    T mu = p.mu[stream];
    cx<T> prev = make_cx<T>(0, 0);

    const long long ntiles_it = (p.TrSyms + p.tile_syms - 1) / p.tile_syms;
    const long long ntiles = ntiles_it * p.Niter;

    // issue the cp.async loads of tile `g` (global tile counter over all iterations) into buffer
    auto load_tile = [&](long long g, cx<T> *buf) {
        const long long b = g % ntiles_it;
        const long long i0 = b * p.tile_syms;
        const int n = (int)min((long long)p.tile_syms, p.TrSyms - i0);
        const int len = (n - 1) * p.os + p.ntaps;
        for (int k = 0; k < p.nmodes; k++) {
            const cx<T> *src = Eseg + (long long)k * p.row_stride + i0 * p.os;
            cx<T> *dst = buf + k * p.tile_pitch;
            for (int c = lane; c < len; c += 32) cp_async<sizeof(cx<T>)>(dst + c, src + c);
        }
        cp_async_commit();
    };

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
        for (int il = 0; il < n; il++) {
            const cx<T> *xb = cur + il * p.os;
            cx<T> x[NQ];
            T ar = 0, ai = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                x[q] = xb[off[q]];
                ar = fma(x[q].x, wr[q], ar);
                if (!REAL) {
                    ar = fma(-x[q].y, wi[q], ar);
                    ai = fma(x[q].x, wi[q], ai);
                    ai = fma(x[q].y, wr[q], ai);
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                ar += shfl_xor(ar, m);
                if (!REAL) ai += shfl_xor(ai, m);
            }
            const long long i = i0 + il;
            const cx<T> e = error_fct<T>(p.method, make_cx<T>(ar, ai), syms, p.K, gsyms, i, lane);
            if (lane == 0) errs[il] = e;
            const T cr = mu * e.x, ci = mu * e.y;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                if (lane + 32 * q < Ktot) {
                    // (cr + j ci) * (x.re - j x.im)
                    wr[q] = fma(cr, x[q].x, wr[q]);
                    if (!REAL) {
                        wr[q] = fma(ci, x[q].y, wr[q]);
                        wi[q] = fma(ci, x[q].x, wi[q]);
                        wi[q] = fma(-cr, x[q].y, wi[q]);
                    }
                }
            }
            if (p.adaptive && i > 0)
                mu = p.method >= QB_CMA_REAL ? adapt_step_real<T>(mu, e.x, prev.x) : adapt_step<T>(mu, e, prev);
            prev = e;
        }
        __syncwarp();
        if (p.err) {
            cx<T> *eg = p.err + ((long long)s * p.nmodes + mode) * (p.TrSyms * p.Niter) +
                        it * p.TrSyms + i0;
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

template <typename T, int NQ, bool REAL>
static int launch_train_nq_r(const TrainParams<T> &p, long long nstreams, size_t smem, cudaStream_t st)
{
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(train_warp_kernel<T, NQ, REAL>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    train_warp_kernel<T, NQ, REAL><<<(unsigned)nstreams, 32, smem, st>>>(p);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

template <typename T, int NQ>
static int launch_train_nq(const TrainParams<T> &p, long long nstreams, size_t smem, cudaStream_t st)
{
    return p.method >= QB_CMA_REAL ? launch_train_nq_r<T, NQ, true>(p, nstreams, smem, st)
                                   : launch_train_nq_r<T, NQ, false>(p, nstreams, smem, st);
}

template <typename T>
static int launch_train(const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                        int64_t nmodes, int64_t TrSyms, int64_t Niter, int64_t os, void *wx,
                        int64_t ntaps, const int64_t *modes, int64_t nsel, int adaptive,
                        const void *symbols, int64_t K, int method, void *mu, void *err,
                        cudaStream_t st)
{
    if (nseg == 0 || nsel == 0 || TrSyms <= 0 || Niter <= 0) return QB_OK;
    TrainParams<T> p;
    p.E = (const cx<T> *)E;
    p.wx = (cx<T> *)wx;
    p.symbols = (const cx<T> *)symbols;
    p.mu = (T *)mu;
    p.err = (cx<T> *)err;
    p.seg_stride = seg_stride;
    p.row_stride = row_stride;
    p.TrSyms = TrSyms;
    p.nmodes = (int)nmodes;
    p.nsel = (int)nsel;
    p.os = (int)os;
    p.ntaps = (int)ntaps;
    p.Niter = (int)Niter;
    p.adaptive = adaptive;
    p.method = method;
    p.K = (int)K;
    p.modes.n = (int)nsel;
    for (int j = 0; j < nsel; j++) p.modes.m[j] = (int)modes[j];
    p.nsym_smem = (method == QB_SBD_DATA || method == QB_DD_DATA_REAL) ? 0 : (int)K;
    p.L = 0;
    p.nstreams = nseg * nsel;
    if (p.nstreams > 2147483647LL) return set_error(QB_ERR_UNSUPPORTED, "train_equaliser: too many streams");
    if (sizeof(T) == 4) {
        // option TRAIN_KERNEL (qb_set_option) = warp forces the generic warp-per-stream kernel, = gla the generic
        // look-ahead kernel (used by the parity tests)
        const char force = option_char(OPT_TRAIN_KERNEL);
        // the real-valued error functions (QB_*_REAL) only exist in the generic kernels
        if (force != 'w' && force != 'g' && method < QB_CMA_REAL) {
            const int rc = train_fast_try(*reinterpret_cast<TrainParams<float> *>(&p), st);
            if (rc != 0) return rc < 0 ? rc : QB_OK;
        }
    }
    {
        // generic look-ahead kernel (any dtype / os / nmodes, nmodes*ntaps <= 128); TRAIN_KERNEL = warp keeps the
        // direct warp-per-stream form
        if (option_char(OPT_TRAIN_KERNEL) != 'w') {
            const int rc = train_gla_try<T>(p, st);
            if (rc != 0) return rc < 0 ? rc : QB_OK;
        }
    }
    // tile: as many symbols as fit a ~12 KB double buffer, at most 256
    int ts = 256;
    size_t smem = 0;
    for (; ts >= 8; ts >>= 1) {
        const int pitch = (ts - 1) * (int)os + (int)ntaps;
        smem = ((size_t)2 * nmodes * pitch + ts + p.nsym_smem) * sizeof(cx<T>);
        if (smem <= 12 * 1024 || (ts == 8 && smem <= 200 * 1024)) break;
    }
    if (ts < 8) return set_error(QB_ERR_UNSUPPORTED, "train_equaliser: window does not fit shared memory");
    p.tile_syms = ts;
    p.tile_pitch = (ts - 1) * (int)os + (int)ntaps;
    const long long nstreams = nseg * nsel;
    if (nstreams > 2147483647LL) return set_error(QB_ERR_UNSUPPORTED, "train_equaliser: too many streams");
    const int Ktot = (int)(nmodes * ntaps);
    const int nq = (Ktot + 31) / 32;
#define QB_NQ_CASE(N) \
    if (nq <= N) return launch_train_nq<T, N>(p, nstreams, smem, st)
    QB_NQ_CASE(1);
    QB_NQ_CASE(2);
    QB_NQ_CASE(3);
    QB_NQ_CASE(4);
    QB_NQ_CASE(6);
    QB_NQ_CASE(8);
    QB_NQ_CASE(12);
    QB_NQ_CASE(16);
#undef QB_NQ_CASE
    return set_error(QB_ERR_UNSUPPORTED, "train_equaliser: nmodes*ntaps = %d exceeds %d", Ktot,
                     QB_MAX_TAPDIM);
}

int train_dispatch(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                   int64_t nmodes, int64_t TrSyms, int64_t Niter, int64_t os, void *wx, int64_t ntaps,
                   const int64_t *modes, int64_t nsel, int adaptive, const void *symbols, int64_t K,
                   int method, void *mu, void *err, cudaStream_t st)
{
    if (dtype == QB_C64)
        return launch_train<float>(E, nseg, seg_stride, row_stride, nmodes, TrSyms, Niter, os, wx, ntaps,
                                   modes, nsel, adaptive, symbols, K, method, mu, err, st);
    return launch_train<double>(E, nseg, seg_stride, row_stride, nmodes, TrSyms, Niter, os, wx, ntaps,
                                modes, nsel, adaptive, symbols, K, method, mu, err, st);
}

}  // namespace qb
