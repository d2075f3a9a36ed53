// Decisions and quality metrics on the device: what scores a 1e7..1e9-symbol output without
// the CPU becoming the bottleneck.
//
//   make_decision                  qampy/core/equalisation/pythran_equalisation.py:306-334 (det_symbol_argmin :232-235)
//   soft_l_value_demapper          qampy/core/pythran_dsp.py:95-108 (cal_l_values :86-88)
//   soft_l_value_demapper_minmax   qampy/core/pythran_dsp.py:110-131 (find_minmax :110-121)
//   estimate_snr                   qampy/core/pythran_dsp.py:244-286
//
// All four are one pass (or two, for the SNR) over the symbols with the alphabet in shared memory.
#include "qb_common.cuh"

namespace qb {

constexpr int DEC_THREADS = 256;

// ---- make_decision ------------------------------------------------------------------------------------
// idx = np.argmin(np.abs(X - symbs)): FIRST minimum of the ROUNDED moduli.  The differences are formed in the
// signal dtype like NumPy does; the modulus is sqrt(dr^2 + di^2) evaluated in double and rounded to the signal
// dtype, i.e. a correctly rounded hypot -- what NumPy's abs (glibc hypot) returns -- so ties between moduli
// that round to the same value go to the lower index exactly as in the reference.
template <typename T>
__global__ void __launch_bounds__(DEC_THREADS) make_decision_kernel(const cx<T> *E, long long L, const cx<T> *symbols,
                                                                    int M, cx<T> *det, T *dist, int32_t *idx)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T> *sy = reinterpret_cast<cx<T> *>(smem_raw);
    for (int c = threadIdx.x; c < M; c += DEC_THREADS) sy[c] = symbols[c];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * DEC_THREADS + threadIdx.x; i < L; i += (long long)gridDim.x * DEC_THREADS) {
        const cx<T> x = E[i];
        T best = 0;
        int bj = 0;
        for (int j = 0; j < M; j++) {
            const T dr = x.x - sy[j].x, di = x.y - sy[j].y;
            const T d = (T)sqrt((double)dr * (double)dr + (double)di * (double)di);
            if (j == 0 || d < best) {   // NaN never wins: like np.argmin only if it comes first (j == 0)
                best = d;
                bj = j;
            }
        }
        if (det) det[i] = sy[bj];
        if (dist) dist[i] = best;
        if (idx) idx[i] = bj;
    }
}

// ---- soft demappers -----------------------------------------------------------------------------------
// bits_map (nbits_total, K, 2) complex: [bit][l][b] = l-th alphabet point whose `bit` equals b.
// minmax:  L = snr * (min_l |btx[l,0] - rx|^2 - min_l |btx[l,1] - rx|^2)          (:110-131; tmp = tmp2 = 10000 start)
// exact:   L = log(sum_l exp(-snr |btx[l,1] - rx|^2)) - log(sum_l exp(-snr |btx[l,0] - rx|^2))   (:86-88, :95-108)
template <typename T, bool MINMAX>
__global__ void __launch_bounds__(DEC_THREADS) demapper_kernel(const cx<T> *rx, long long N, int num_bits, T snr,
                                                               const cx<T> *bits_map, int K, double *Lv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T> *bm = reinterpret_cast<cx<T> *>(smem_raw);    // [num_bits][K][2]
    for (int c = threadIdx.x; c < num_bits * K * 2; c += DEC_THREADS) bm[c] = bits_map[c];
    __syncthreads();
    const long long total = N * num_bits;
    for (long long f = (long long)blockIdx.x * DEC_THREADS + threadIdx.x; f < total; f += (long long)gridDim.x * DEC_THREADS) {
        const long long s = f / num_bits;
        const int bit = (int)(f - s * num_bits);
        const cx<T> x = rx[s];
        const cx<T> *b = bm + (size_t)bit * K * 2;
        if (MINMAX) {
            T m1 = (T)10000., m0 = (T)10000.;
            for (int l = 0; l < K; l++) {
                const T r1 = b[2 * l + 1].x - x.x, i1 = b[2 * l + 1].y - x.y;
                const T r0 = b[2 * l].x - x.x, i0 = b[2 * l].y - x.y;
                const T a1 = (T)sqrt((double)r1 * r1 + (double)i1 * i1), a0 = (T)sqrt((double)r0 * r0 + (double)i0 * i0);
                const T d1 = a1 * a1, d0 = a0 * a0;          // abs(.)**2 in the signal dtype
                if (d1 < m1) m1 = d1;
                if (d0 < m0) m0 = d0;
            }
            Lv[f] = (double)(snr * (m0 - m1));
        } else {
            T s1 = 0, s0 = 0;
            for (int l = 0; l < K; l++) {
                const T r1 = b[2 * l + 1].x - x.x, i1 = b[2 * l + 1].y - x.y;
                const T r0 = b[2 * l].x - x.x, i0 = b[2 * l].y - x.y;
                const T a1 = (T)sqrt((double)r1 * r1 + (double)i1 * i1), a0 = (T)sqrt((double)r0 * r0 + (double)i0 * i0);
                s1 += exp(-snr * (a1 * a1));
                s0 += exp(-snr * (a0 * a0));
            }
            Lv[f] = (double)(log(s1) - log(s0));
        }
    }
}

// ---- estimate_snr -------------------------------------------------------------------------------------
// pass 1: per alphabet point count K_c and sum of the received symbols sent as that point; pass 2: sum of
// |x - mean_c|^2.  Accumulated in double (the reference sums in the signal dtype; the result agrees to rounding).
// A symbol belongs to class c when symbols_tx[i] == gray[c] exactly (:273).
template <typename T>
__global__ void __launch_bounds__(DEC_THREADS) snr_pass_kernel(const cx<T> *rx, const cx<T> *tx, long long n, const cx<T> *gray,
                                                               int Ncls, const double *means, double *acc, int pass)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T> *gs = reinterpret_cast<cx<T> *>(smem_raw);                      // [Ncls]
    double *loc = reinterpret_cast<double *>(gs + Ncls + (Ncls & 1));     // pass 1: [Ncls][3], pass 2: [Ncls]
    const int width = pass == 1 ? 3 : 1;
    for (int c = threadIdx.x; c < Ncls; c += DEC_THREADS) gs[c] = gray[c];
    for (int c = threadIdx.x; c < Ncls * width; c += DEC_THREADS) loc[c] = 0.;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * DEC_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * DEC_THREADS) {
        const cx<T> t = tx[i], x = rx[i];
        for (int c = 0; c < Ncls; c++) {
            if (t.x == gs[c].x && t.y == gs[c].y) {
                if (pass == 1) {
                    atomicAdd(&loc[3 * c], 1.0);
                    atomicAdd(&loc[3 * c + 1], (double)x.x);
                    atomicAdd(&loc[3 * c + 2], (double)x.y);
                } else {
                    const double dr = (double)x.x - means[2 * c], di = (double)x.y - means[2 * c + 1];
                    atomicAdd(&loc[c], dr * dr + di * di);
                }
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Ncls * width; c += DEC_THREADS)
        if (loc[c] != 0.) atomicAdd(&acc[c], loc[c]);
}

template <typename T>
static int launch_decision(const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist, int32_t *idx,
                           cudaStream_t st)
{
    if (L == 0) return QB_OK;
    const size_t smem = (size_t)M * sizeof(cx<T>);
    if (smem > 48 * 1024) return set_error(QB_ERR_UNSUPPORTED, "make_decision: alphabet too large for shared memory");
    const unsigned nb = (unsigned)min((long long)(148 * 8), (long long)((L + DEC_THREADS - 1) / DEC_THREADS));
    make_decision_kernel<T><<<nb, DEC_THREADS, smem, st>>>((const cx<T> *)E, L, (const cx<T> *)symbols, (int)M, (cx<T> *)det,
                                                           (T *)dist, idx);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int make_decision_dispatch(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                           int32_t *idx, cudaStream_t st)
{
    return dtype == QB_C64 ? launch_decision<float>(E, L, symbols, M, det, dist, idx, st)
                           : launch_decision<double>(E, L, symbols, M, det, dist, idx, st);
}

template <typename T>
static int launch_demapper(const void *rx, int64_t N, int64_t num_bits, double snr, const void *bits_map, int64_t K,
                           int minmax, double *Lv, cudaStream_t st)
{
    if (N == 0 || num_bits == 0) return QB_OK;
    const size_t smem = (size_t)num_bits * K * 2 * sizeof(cx<T>);
    if (smem > 48 * 1024) return set_error(QB_ERR_UNSUPPORTED, "demapper: bit map too large for shared memory");
    const long long total = N * num_bits;
    const unsigned nb = (unsigned)min((long long)(148 * 8), (long long)((total + DEC_THREADS - 1) / DEC_THREADS));
    if (minmax)
        demapper_kernel<T, true><<<nb, DEC_THREADS, smem, st>>>((const cx<T> *)rx, N, (int)num_bits, (T)snr,
                                                                (const cx<T> *)bits_map, (int)K, Lv);
    else
        demapper_kernel<T, false><<<nb, DEC_THREADS, smem, st>>>((const cx<T> *)rx, N, (int)num_bits, (T)snr,
                                                                 (const cx<T> *)bits_map, (int)K, Lv);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int demapper_dispatch(int dtype, const void *rx, int64_t N, int64_t num_bits, double snr, const void *bits_map, int64_t K,
                      int minmax, double *Lv, cudaStream_t st)
{
    return dtype == QB_C64 ? launch_demapper<float>(rx, N, num_bits, snr, bits_map, K, minmax, Lv, st)
                           : launch_demapper<double>(rx, N, num_bits, snr, bits_map, K, minmax, Lv, st);
}

template <typename T>
static int launch_snr_pass(const void *rx, const void *tx, int64_t n, const void *gray, int64_t Ncls, const double *means,
                           double *acc, int pass, cudaStream_t st)
{
    const size_t smem = (size_t)(Ncls + (Ncls & 1)) * sizeof(cx<T>) + (size_t)Ncls * 3 * sizeof(double);
    if (smem > 48 * 1024) return set_error(QB_ERR_UNSUPPORTED, "estimate_snr: alphabet too large for shared memory");
    const unsigned nb = (unsigned)min((long long)(148 * 4), (long long)((n + DEC_THREADS - 1) / DEC_THREADS));
    snr_pass_kernel<T><<<nb ? nb : 1, DEC_THREADS, smem, st>>>((const cx<T> *)rx, (const cx<T> *)tx, n, (const cx<T> *)gray,
                                                               (int)Ncls, means, acc, pass);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int snr_pass_dispatch(int dtype, const void *rx, const void *tx, int64_t n, const void *gray, int64_t Ncls,
                      const double *means, double *acc, int pass, cudaStream_t st)
{
    return dtype == QB_C64 ? launch_snr_pass<float>(rx, tx, n, gray, Ncls, means, acc, pass, st)
                           : launch_snr_pass<double>(rx, tx, n, gray, Ncls, means, acc, pass, st);
}

}  // namespace qb
