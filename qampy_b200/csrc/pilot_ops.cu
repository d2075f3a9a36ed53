// Element-wise stages of the pilot-based receiver (BASELINE config C4) that would otherwise
// stay O(L) NumPy on the host between the CUDA equaliser calls:
//
//   freq_shift   comp_freq_offset (qampy/core/phaserecovery.py:438-473):
//                out[r, t] = E[r, t] * exp(-2j pi (t + 1) f_r / os),  t = 0 .. L-1
//   pilot_cpe    pilot_based_cpe_new (qampy/core/pilotbased_receiver.py:258-327) for one frame per row:
//                residual phase at the pilots angle(conj(p) r) -> np.unwrap -> moving average over
//                num_average pilots -> np.interp to every symbol -> out = E * exp(-1j phase)
//
// The reference evaluates the ramp / interpolation in float64 and the products in complex128 before
// storing to the signal dtype; so do these kernels (phases in double, reduced to one turn before the
// sincos), which keeps them within rounding of the NumPy results (tests: <= 2e-6 rms in c64).
#include "qb_common.cuh"

namespace qb {

template <typename T>
__global__ void freq_shift_kernel(const cx<T> *E, long long row_stride, long long L, const double *freq, double inv_os,
                                  long long t0, cx<T> *out, long long out_stride)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (t >= L) return;
    const double turns = (double)(t0 + t + 1) * freq[r] * inv_os;     // lin_phase / (2 pi)  (:466-468)
    double s, c;
    sincospi(2.0 * (turns - floor(turns)), &s, &c);
    const cx<T> e = E[(long long)r * row_stride + t];
    // e * exp(-1j phase) in double, stored in the signal dtype
    out[(long long)r * out_stride + t] = make_cx<T>((T)((double)e.x * c + (double)e.y * s),
                                                    (T)((double)e.y * c - (double)e.x * s));
}

int freq_shift_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, const double *freq,
                        int64_t os, int64_t t0, void *out, int64_t out_stride, cudaStream_t st)
{
    if (nrows == 0 || L == 0) return QB_OK;
    dim3 grid((unsigned)((L + 255) / 256), (unsigned)nrows);
    if (dtype == QB_C64)
        freq_shift_kernel<float><<<grid, 256, 0, st>>>((const float2 *)E, row_stride, L, freq, 1.0 / (double)os, t0,
                                                       (float2 *)out, out_stride);
    else
        freq_shift_kernel<double><<<grid, 256, 0, st>>>((const double2 *)E, row_stride, L, freq, 1.0 / (double)os, t0,
                                                        (double2 *)out, out_stride);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

// One CTA per row (one frame of one mode).  Shared memory: unwrapped pilot phases [nph] (T), their running
// sum [nph + 1] (T), averaged phases [nph - navg + 1] (double).
constexpr int CPE_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(CPE_THREADS) pilot_cpe_kernel(const cx<T> *E, long long row_stride, long long nlen,
                                                                const long long *pidx, const cx<T> *pilots,
                                                                long long pilot_stride, int nph, int navg,
                                                                cx<T> *out, long long out_stride, T *trace,
                                                                long long trace_stride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *phase = reinterpret_cast<T *>(smem_raw);           // [nph]
    T *csum = phase + nph;                                // [nph + 1]
    double *avg = reinterpret_cast<double *>(smem_raw + (((size_t)(2 * nph + 1) * sizeof(T) + 7) & ~(size_t)7));
    const int tid = threadIdx.x;
    const long long row = blockIdx.x;
    const cx<T> *Er = E + row * row_stride;
    const cx<T> *pr = pilots + row * pilot_stride;
    // residual phase at the pilots: np.angle(conj(p) * r)  (:311)
    for (int k = tid; k < nph; k += CPE_THREADS) {
        const cx<T> p = pr[k], r = Er[pidx[k]];
        const T re = p.x * r.x + p.y * r.y, im = p.x * r.y - p.y * r.x;
        phase[k] = atan2(im, re);
    }
    __syncthreads();
    if (tid == 0) {
        // np.unwrap along the pilots (default period), then the running sum of moving_average (filter.py:233),
        // both sequential in the signal's real dtype like NumPy's cumsum
        const T PI = (T)3.141592653589793238462643383279502884, TWO_PI = (T)6.283185307179586476925286766559005768;
        T cum = 0, prev = nph > 0 ? phase[0] : (T)0;
        T run = 0;
        csum[0] = 0;
        for (int k = 0; k < nph; k++) {
            const T p = phase[k];
            if (k > 0) {
                const T dd = p - prev;
                T m = fmod(dd + PI, TWO_PI);
                if (m != (T)0 && m < (T)0) m += TWO_PI;
                T ddmod = m - PI;
                if (ddmod == -PI && dd > (T)0) ddmod = PI;
                T corr = ddmod - dd;
                if (fabs(dd) < PI) corr = 0;
                cum += corr;
            }
            prev = p;
            const T up = p + cum;
            phase[k] = up;
            run += up;
            csum[k + 1] = run;
        }
    }
    __syncthreads();
    const int navail = nph - navg + 1;                    // averaged phases, centred on pilots half .. nph-1-half
    const int half = (navg - 1) / 2;
    for (int k = tid; k < navail; k += CPE_THREADS) avg[k] = (double)((csum[k + navg] - csum[k]) / (T)navg);
    __syncthreads();
    // np.interp(arange(nlen), pidx[half : nph - half], avg) in double (:320), clamped at both ends
    const long long x0 = pidx[half], x1 = pidx[half + navail - 1];
    for (long long i = tid; i < nlen; i += CPE_THREADS) {
        double ph;
        if (i <= x0) ph = avg[0];
        else if (i >= x1) ph = avg[navail - 1];
        else {
            // pilots are sorted: binary search for the interval [pidx[half + j], pidx[half + j + 1]) holding i
            int lo = 0, hi = navail - 1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pidx[half + mid] <= i) lo = mid; else hi = mid;
            }
            const double xa = (double)pidx[half + lo], xb = (double)pidx[half + lo + 1];
            const double slope = (avg[lo + 1] - avg[lo]) / (xb - xa);
            ph = slope * ((double)i - xa) + avg[lo];
        }
        const T pht = (T)ph;                               // the reference stores the trace in the signal dtype (:317)
        if (trace) trace[row * trace_stride + i] = pht;
        double s, c;
        sincos((double)pht, &s, &c);
        const cx<T> e = Er[i];
        out[row * out_stride + i] = make_cx<T>((T)((double)e.x * c + (double)e.y * s), (T)((double)e.y * c - (double)e.x * s));
    }
}

int pilot_cpe_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t nlen, const int64_t *pidx,
                       const void *pilots, int64_t pilot_stride, int64_t nph, int64_t navg, void *out,
                       int64_t out_stride, void *trace, int64_t trace_stride, cudaStream_t st)
{
    if (nrows == 0 || nlen == 0) return QB_OK;
    const size_t ts = dtype == QB_C64 ? 4 : 8;
    const size_t smem = (((size_t)(2 * nph + 1) * ts + 7) & ~(size_t)7) + (size_t)(nph - navg + 1) * 8;
    if (smem > 200 * 1024) return set_error(QB_ERR_UNSUPPORTED, "pilot_cpe: too many pilots per row for shared memory");
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(pilot_cpe_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QB_CUDA_CHECK(cudaFuncSetAttribute(pilot_cpe_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dtype == QB_C64)
        pilot_cpe_kernel<float><<<(unsigned)nrows, CPE_THREADS, smem, st>>>(
            (const float2 *)E, row_stride, nlen, (const long long *)pidx, (const float2 *)pilots, pilot_stride, (int)nph,
            (int)navg, (float2 *)out, out_stride, (float *)trace, trace_stride);
    else
        pilot_cpe_kernel<double><<<(unsigned)nrows, CPE_THREADS, smem, st>>>(
            (const double2 *)E, row_stride, nlen, (const long long *)pidx, (const double2 *)pilots, pilot_stride, (int)nph,
            (int)navg, (double2 *)out, out_stride, (double *)trace, trace_stride);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb
