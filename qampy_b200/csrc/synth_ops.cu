// Signal synthesis on the device: the element-wise and spectral stages of the reference's generators,
// fused around cuFFT (which the host side calls through torch.fft -- the only library work in the chain):
//
//   pulse shaping    rrcos_resample / rrcos_pulseshaping  qampy/core/resample.py:73-126, core/filter.py:177-212
//                    zero insertion -> [FFT] -> multiply by the spectrum of the tap vector -> [IFFT] -> "same" crop,
//                    decimation, re-centring and re-scaling fused into ONE pass (moments by a reduction kernel)
//   PMD              apply_PMD_to_field / _applyPMD_dot   qampy/core/impairments.py:94-131
//                    [FFT] -> rotate(theta) -> e^{-+ i omega dgd/2} on the two axes -> rotate(-theta) -> [IFFT];
//                    the rotation pair and the phase ramp are one kernel over both polarisations, the ramp is evaluated
//                    in double from the bin index (the reference's fftshift / ifftshift pairs cancel around a
//                    bin-diagonal operator)
//   AWGN             add_awgn / change_snr                core/impairments.py:188-233
//   phase noise      phase_noise / apply_phase_noise      core/impairments.py:133-186  (Wiener walk = running sum)
//                    normal deviates from a counter-based generator in the kernel (Philox4x32-10 + Box-Muller: a
//                    sample depends on (seed, row, index) only, so a capture synthesised in blocks or on several GPUs
//                    is the same capture); the walk is a three-kernel scan in double; noise, walk and the final
//                    cast to the signal dtype are one pass
//
// All arithmetic is double (the reference generates complex128 and casts at the very end).
#include <algorithm>

#include "qb_common.cuh"

namespace qb {

// ---- Philox4x32-10 (Salmon et al., SC'11), restated ---------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// two independent standard normal deviates for (seed, stream, index): Box-Muller on two 53-bit uniforms
__device__ __forceinline__ double2 normal_pair(uint64_t seed, uint32_t stream, uint64_t index)
{
    const uint4 r = philox4x32_10(make_uint4((uint32_t)index, (uint32_t)(index >> 32), stream, 0x5eedu),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);     // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);             // [0, 1)
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    return make_double2(rad * c, rad * s);
}

// ---- zero insertion ------------------------------------------------------------------------------------------------------
__global__ void synth_upsample_kernel(const double2 *sym, long long n, int up, double2 *out, long long out_stride)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (t >= out_stride) return;
    double2 v = make_double2(0.0, 0.0);
    if (t < n * up && t % up == 0) v = sym[(long long)r * n + t / up];
    out[(long long)r * out_stride + t] = v;
}

// ---- spectrum times table (pulse shaping: table = spectrum of the zero-padded tap vector) ---------------------------------
__global__ void synth_specmul_kernel(double2 *X, long long nfft, const double2 *H)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nfft) return;
    double2 *x = X + (long long)blockIdx.y * nfft + k;
    const double2 a = *x, h = H[k];
    *x = make_double2(a.x * h.x - a.y * h.y, a.x * h.y + a.y * h.x);
}

// ---- moments of the "same" crop of a convolution, decimated: sum x, sum |x|^2 per row --------------------------------------
__global__ void synth_moments_kernel(const double2 *x, long long row_stride, long long first, int down, long long n,
                                     double *mom /* [rows][3]: sum re, sum im, sum |.|^2 */)
{
    __shared__ double red[3][8];
    const int r = blockIdx.y;
    double sr = 0.0, si = 0.0, sp = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const double2 v = x[(long long)r * row_stride + first + t * down];
        sr += v.x;
        si += v.y;
        sp += v.x * v.x + v.y * v.y;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, d);
        si += __shfl_xor_sync(0xffffffffu, si, d);
        sp += __shfl_xor_sync(0xffffffffu, sp, d);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[0][w] = sr, red[1][w] = si, red[2][w] = sp;
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) a += red[threadIdx.x][i];
        atomicAdd(&mom[r * 3 + threadIdx.x], a);
    }
}

// ---- crop + decimate + re-centre + re-scale: normalise_and_center(sig) * sqrt(p)  (resample.py:123-125) ---------------------
__global__ void synth_crop_norm_kernel(const double2 *x, long long row_stride, long long first, int down, long long n,
                                       const double *mom, const double *target_power, int renorm, double2 *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (t >= n) return;
    double2 v = x[(long long)r * row_stride + first + t * down];
    if (renorm) {
        const double mr = mom[r * 3] / (double)n, mi = mom[r * 3 + 1] / (double)n;
        // power of the centred signal: E|x - m|^2 = E|x|^2 - |m|^2
        const double p = mom[r * 3 + 2] / (double)n - (mr * mr + mi * mi);
        const double g = sqrt(target_power[r] / p);
        v = make_double2((v.x - mr) * g, (v.y - mi) * g);
    }
    out[(long long)r * n + t] = v;
}

// ---- first-order PMD on the spectra of both polarisations (FFT order) --------------------------------------------------------
__global__ void synth_pmd_kernel(double2 *S0, double2 *S1, long long n, double cth, double sth, double half_dgd_dw)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // omega_k t_dgd / 2 with omega on the reference's grid 2 pi linspace(-fs/2, fs/2, n, endpoint=False), which in FFT
    // order is 2 pi fs k' / n,  k' = k for k < ceil(n/2), k - n above (n even: bin n/2 carries -fs/2)
    // (odd n: the reference's linspace grid sits half a bin below the true bin frequencies; kept)
    const long long ks = (k < (n + 1) / 2) ? k : k - n;
    const double ph = ((double)ks - ((n & 1) ? 0.5 : 0.0)) * half_dgd_dw;
    double s, c;
    sincos(ph, &s, &c);
    const double2 x = S0[k], y = S1[k];
    // rotate_field(theta): [c -s; s c]
    const double2 a = make_double2(cth * x.x - sth * y.x, cth * x.y - sth * y.y);
    const double2 b = make_double2(sth * x.x + cth * y.x, sth * x.y + cth * y.y);
    // axis delays: a * exp(-i ph), b * exp(+i ph)
    const double2 ad = make_double2(a.x * c + a.y * s, a.y * c - a.x * s);
    const double2 bd = make_double2(b.x * c - b.y * s, b.y * c + b.x * s);
    // rotate_field(-theta): [c s; -s c]
    S0[k] = make_double2(cth * ad.x + sth * bd.x, cth * ad.y + sth * bd.y);
    S1[k] = make_double2(-sth * ad.x + cth * bd.x, -sth * ad.y + cth * bd.y);
}

// ---- Wiener phase walk: running sum of N(0, var) steps, three-kernel scan -----------------------------------------------------
constexpr int WALK_TILE = 2048;   // samples per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) synth_walk_tiles_kernel(long long n, double sigma, uint64_t seed, uint32_t stream0,
                                                                uint64_t index0, double *tile_sum, long long ntiles)
{
    __shared__ double red[8];
    const int r = blockIdx.y;
    const long long base = (long long)blockIdx.x * WALK_TILE + threadIdx.x * 8;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const long long t = base + j;       // even: one Philox call gives the steps of samples t and t + 1
        if (t < n) {
            const double2 g = normal_pair(seed, stream0 + r, (index0 + t) >> 1);
            s += g.x * sigma;
            if (t + 1 < n) s += g.y * sigma;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 8; i++) a += red[i];
        tile_sum[(long long)r * ntiles + blockIdx.x] = a;
    }
}

// exclusive scan of the tile sums of one row (one CTA per row; sequential over chunks of 256 tiles)
__global__ void __launch_bounds__(256) synth_walk_scan_kernel(double *tile_sum, long long ntiles, const double *phase0)
{
    __shared__ double buf[256];
    __shared__ double carry;
    double *row = tile_sum + (long long)blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry = phase0 ? phase0[blockIdx.x] : 0.0;
    __syncthreads();
    for (long long c0 = 0; c0 < ntiles; c0 += 256) {
        const long long i = c0 + threadIdx.x;
        const double v = i < ntiles ? row[i] : 0.0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const double add = threadIdx.x >= d ? buf[threadIdx.x - d] : 0.0;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        const double incl = buf[threadIdx.x], base = carry;
        __syncthreads();
        if (i < ntiles) row[i] = base + incl - v;
        if (threadIdx.x == 255) carry = base + incl;
        __syncthreads();
    }
}

// ---- the element-wise tail: (x + noise) * exp(i phase) -> signal dtype -------------------------------------------------------
// x: double complex rows; noise: sigma_n (N(0,1) + i N(0,1)) / sqrt(2) per sample (core/impairments.py:205); phase: the walk
// re-generated from the same counters (steps of a tile summed in order inside the CTA, tile offsets from the scan).
template <typename T>
__global__ void __launch_bounds__(256) synth_tail_kernel(const double2 *x, long long n, const double *noise_sigma,
                                                         double walk_sigma, uint64_t seed, uint32_t noise_stream0,
                                                         uint32_t walk_stream0, uint64_t index0, const double *tile_off,
                                                         long long ntiles, cx<T> *out, long long out_stride,
                                                         double *phase_out)
{
    __shared__ double wsum[8];
    const int r = blockIdx.y;
    const long long base = (long long)blockIdx.x * WALK_TILE + threadIdx.x * 8;
    double step[8];
#pragma unroll
    for (int j = 0; j < 8; j++) step[j] = 0.0;
    double run = 0.0;
    if (walk_sigma != 0.0) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const long long t = base + j;
            if (t < n) {
                const double2 g = normal_pair(seed, walk_stream0 + r, (index0 + t) >> 1);
                step[j] = g.x * walk_sigma;
                step[j + 1] = (t + 1 < n) ? g.y * walk_sigma : 0.0;
            }
        }
        // inclusive scan of the thread totals inside the CTA
        double tot = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) tot += step[j];
        double inc = tot;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        double woff = 0.0;
        for (int i = 0; i < w; i++) woff += wsum[i];
        run = tile_off[(long long)r * ntiles + blockIdx.x] + woff + inc - tot;     // walk before this thread's first sample
    }
    const double ns = noise_sigma ? noise_sigma[r] * 0.70710678118654752440 : 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const long long t = base + j;
        if (t >= n) break;
        run += step[j];
        double2 v = x[(long long)r * n + t];
        if (ns != 0.0) {
            const double2 g = normal_pair(seed, noise_stream0 + r, index0 + t);
            v.x += ns * g.x;
            v.y += ns * g.y;
        }
        if (walk_sigma != 0.0) {
            double s, c;
            sincos(run, &s, &c);
            v = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
            if (phase_out) phase_out[(long long)r * n + t] = run;
        }
        out[(long long)r * out_stride + t] = make_cx<T>((T)v.x, (T)v.y);
    }
}

// ---- host dispatch -------------------------------------------------------------------------------------------------------------
int synth_upsample_dispatch(const void *sym, int64_t rows, int64_t n, int64_t up, void *out, int64_t out_stride, cudaStream_t st)
{
    if (rows == 0 || out_stride == 0) return QB_OK;
    dim3 grid((unsigned)((out_stride + 255) / 256), (unsigned)rows);
    synth_upsample_kernel<<<grid, 256, 0, st>>>((const double2 *)sym, n, (int)up, (double2 *)out, out_stride);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int synth_specmul_dispatch(void *X, int64_t rows, int64_t nfft, const void *H, cudaStream_t st)
{
    if (rows == 0 || nfft == 0) return QB_OK;
    dim3 grid((unsigned)((nfft + 255) / 256), (unsigned)rows);
    synth_specmul_kernel<<<grid, 256, 0, st>>>((double2 *)X, nfft, (const double2 *)H);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int synth_crop_norm_dispatch(const void *x, int64_t rows, int64_t row_stride, int64_t first, int64_t down, int64_t n,
                             const double *target_power, int renorm, double *mom, void *out, cudaStream_t st)
{
    if (rows == 0 || n == 0) return QB_OK;
    if (renorm) {
        QB_CUDA_CHECK(cudaMemsetAsync(mom, 0, (size_t)rows * 3 * sizeof(double), st));
        dim3 g1((unsigned)std::min<int64_t>((n + 255) / 256, 1024), (unsigned)rows);
        synth_moments_kernel<<<g1, 256, 0, st>>>((const double2 *)x, row_stride, first, (int)down, n, mom);
        count_launch();
    }
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)rows);
    synth_crop_norm_kernel<<<grid, 256, 0, st>>>((const double2 *)x, row_stride, first, (int)down, n, mom, target_power,
                                                  renorm, (double2 *)out);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int synth_pmd_dispatch(void *S, int64_t n, double theta, double t_dgd, double fs, cudaStream_t st)
{
    if (n == 0) return QB_OK;
    double2 *S0 = (double2 *)S, *S1 = S0 + n;
    const double half_dgd_dw = 2.0 * 3.14159265358979323846 * fs / (double)n * t_dgd / 2.0;
    synth_pmd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S0, S1, n, cos(theta), sin(theta), half_dgd_dw);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int synth_tail_dispatch(int dtype, const void *x, int64_t rows, int64_t n, const double *noise_sigma, double walk_sigma,
                        uint64_t seed, int64_t row0, uint64_t index0, const double *phase0, double *tile_buf, void *out,
                        int64_t out_stride, double *phase_out, cudaStream_t st)
{
    if (rows == 0 || n == 0) return QB_OK;
    const long long ntiles = (n + WALK_TILE - 1) / WALK_TILE;
    dim3 grid((unsigned)ntiles, (unsigned)rows);
    // noise and walk draw from disjoint counter streams: 2 row + {0, 1}
    const uint32_t ns0 = (uint32_t)(2 * row0), ws0 = (uint32_t)(2 * row0) + 0x80000000u;
    if (walk_sigma != 0.0) {
        synth_walk_tiles_kernel<<<grid, 256, 0, st>>>(n, walk_sigma, seed, ws0, index0, tile_buf, ntiles);
        synth_walk_scan_kernel<<<(unsigned)rows, 256, 0, st>>>(tile_buf, ntiles, phase0);
        count_launch(2);
    }
    if (dtype == QB_C64)
        synth_tail_kernel<float><<<grid, 256, 0, st>>>((const double2 *)x, n, noise_sigma, walk_sigma, seed, ns0, ws0,
                                                       index0, tile_buf, ntiles, (float2 *)out, out_stride, phase_out);
    else
        synth_tail_kernel<double><<<grid, 256, 0, st>>>((const double2 *)x, n, noise_sigma, walk_sigma, seed, ns0, ws0,
                                                        index0, tile_buf, ntiles, (double2 *)out, out_stride, phase_out);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb
