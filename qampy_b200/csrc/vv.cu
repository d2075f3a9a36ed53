// Viterbi-Viterbi M-th power carrier recovery, qampy/core/phaserecovery.py:40-79, one signal row
// per grid row:
//
//   u[t]   = exp(1j*angle(E[t]))**M                                  (:63-64)
//   s[w]   = sum_{t = w .. w+N-1} u[t],    w = 0 .. L-N              (segment_axis(u, N, N-1), :65-66)
//   p[w]   = unwrap(angle(s[w]));  ph[w] = (p[w] - pi)/M             (:67-68)
//   out[o + w] = E[o + w] * exp(-1j*ph[w]),  o = (N-1)//2;  0 outside  (:69-72)
//
// The reference evaluates every step in the signal's own precision with libm transcendentals, so parity is a
// floating-point tolerance (tests: 2e-6 rad / 2e-6 rms in c64, 1e-12 in c128), not bit equality.  These kernels
// take the M-th power by repeated multiplication of the unit phasor and accumulate the window in double, round
// the wrapped phase to the signal's real dtype (what np.angle returns), and unwrap with integer turn counts:
//
//   vv_phase_kernel   wrapped phase per window + per-tile sum of the unwrap turns (tile = VV_TILE windows)
//   vv_scan_kernel    exclusive scan of the tile sums, one CTA per row
//   vv_apply_kernel   in-tile scan of the turns, ph = (p + 2 pi K - pi)/M, rotation of the symbols, zero head/tail
#include "qb_common.cuh"

namespace qb {

constexpr int VV_TILE = 2048;
constexpr int VV_THREADS = 256;
constexpr int VV_PER_THREAD = VV_TILE / VV_THREADS;
constexpr int VV_MAX_N = 512;

// unit phasor of z raised to the M-th power, in double (angle(0) = 0 -> 1)
template <typename T>
__device__ __forceinline__ double2 raise_unit(cx<T> z, int M)
{
    const double re = (double)z.x, im = (double)z.y;
    const double mag = hypot(re, im);
    double2 b = mag > 0.0 ? make_double2(re / mag, im / mag) : make_double2(1.0, 0.0);
    double2 r = make_double2(1.0, 0.0);
    for (int e = M; e > 0; e >>= 1) {
        if (e & 1) r = make_double2(r.x * b.x - r.y * b.y, r.x * b.y + r.y * b.x);
        b = make_double2(b.x * b.x - b.y * b.y, 2.0 * b.x * b.y);
    }
    return r;
}

// unwrap turn for the step prev -> cur (np.unwrap with the default period: a jump of more than pi is folded)
template <typename T>
__device__ __forceinline__ int unwrap_turn(T prev, T cur)
{
    const T PI = (T)3.141592653589793238462643383279502884;
    const T dd = cur - prev;
    return dd > PI ? -1 : (dd < -PI ? 1 : 0);
}

template <typename T>
__global__ void __launch_bounds__(VV_THREADS) vv_phase_kernel(const cx<T> *E, long long row_stride, long long L, int N,
                                                              int M, T *ph, long long ph_stride, int *tile_turns,
                                                              T *tile_prev, int ntiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *u = reinterpret_cast<double2 *>(smem_raw);                 // [VV_TILE + 1 + N - 1] raised symbols
    T *p = reinterpret_cast<T *>(u + VV_TILE + N);                     // [VV_TILE + 1] wrapped phases, p[0] = predecessor
    __shared__ int warp_sum[VV_THREADS / 32];
    const int tid = threadIdx.x;
    const long long row = blockIdx.y;
    const long long nwin = L - N + 1;
    const long long w0 = (long long)blockIdx.x * VV_TILE;              // first window of this tile
    const int nw = (int)min((long long)VV_TILE, nwin - w0);
    const int lead = w0 > 0 ? 1 : 0;                                   // also the window before the tile (for its phase)
    const cx<T> *Er = E + row * row_stride;
    const int nsym = nw + lead + N - 1;
    for (int k = tid; k < nsym; k += VV_THREADS) u[k] = raise_unit<T>(Er[w0 - lead + k], M);
    __syncthreads();
    for (int k = tid; k < nw + lead; k += VV_THREADS) {
        double sr = 0.0, si = 0.0;
        for (int j = 0; j < N; j++) {
            sr += u[k + j].x;
            si += u[k + j].y;
        }
        p[k + 1 - lead] = (T)atan2(si, sr);
    }
    __syncthreads();
    int turns = 0;
    for (int k = tid; k < nw; k += VV_THREADS) {
        const T cur = p[k + 1];
        ph[row * ph_stride + w0 + k] = cur;
        if (k + lead > 0) turns += unwrap_turn<T>(p[k], cur);
    }
    for (int o = 16; o > 0; o >>= 1) turns += __shfl_xor_sync(0xffffffffu, turns, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = turns;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int k = 0; k < VV_THREADS / 32; k++) s += warp_sum[k];
        tile_turns[row * ntiles + blockIdx.x] = s;
        tile_prev[row * ntiles + blockIdx.x] = lead ? p[0] : (T)0;
    }
}

// in place: tile_turns[row, b] <- sum of tile_turns[row, 0 .. b-1]
__global__ void __launch_bounds__(1024) vv_scan_kernel(int *tile_turns, int ntiles)
{
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    int *t = tile_turns + (long long)blockIdx.x * ntiles;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int k = base + tid;
        const int v = k < ntiles ? t[k] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            warp_tot[lane] = w;                                         // inclusive totals of the warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int before = carry + (wid > 0 ? warp_tot[wid - 1] : 0) + inc - v;
        if (k < ntiles) t[k] = before;
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(VV_THREADS) vv_apply_kernel(const cx<T> *E, long long row_stride, long long L, int N,
                                                              int M, T *ph, long long ph_stride, const int *tile_base,
                                                              const T *tile_prev, int ntiles, cx<T> *out,
                                                              long long out_stride)
{
    __shared__ int warp_tot[VV_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long row = blockIdx.y;
    const long long nwin = L - N + 1;
    const long long w0 = (long long)blockIdx.x * VV_TILE;
    const int nw = (int)min((long long)VV_TILE, nwin - w0);
    const long long off = (N - 1) / 2;
    T *pr = ph + row * ph_stride + w0;
    const cx<T> *Er = E + row * row_stride;
    cx<T> *orow = out + row * out_stride;
    // each thread owns VV_PER_THREAD consecutive windows
    const int k0 = tid * VV_PER_THREAD;
    T p[VV_PER_THREAD];
    int turn[VV_PER_THREAD];
    T prev = (k0 == 0) ? tile_prev[row * ntiles + blockIdx.x] : (k0 - 1 < nw ? pr[k0 - 1] : (T)0);
    int local = 0;
#pragma unroll
    for (int j = 0; j < VV_PER_THREAD; j++) {
        const int k = k0 + j;
        p[j] = k < nw ? pr[k] : (T)0;
        int t = 0;
        if (k < nw && (w0 + k) > 0) t = unwrap_turn<T>(prev, p[j]);
        local += t;
        turn[j] = local;                                               // inclusive within the thread
        prev = p[j];
    }
    int inc = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();                                                   // also: every pr[] read above precedes the writes below
    int before = tile_base[row * ntiles + blockIdx.x] + inc - local;
    for (int k = 0; k < wid; k++) before += warp_tot[k];
    const double TWO_PI = 6.283185307179586476925286766559005768;
    const T PI = (T)3.141592653589793238462643383279502884;
#pragma unroll
    for (int j = 0; j < VV_PER_THREAD; j++) {
        const int k = k0 + j;
        if (k >= nw) break;
        const T up = (T)((double)p[j] + TWO_PI * (double)(before + turn[j]));
        const T est = (up - PI) / (T)M;
        pr[k] = est;
        double s, c;
        sincos((double)est, &s, &c);
        const cx<T> e = Er[off + w0 + k];
        orow[off + w0 + k] = make_cx<T>((T)((double)e.x * c + (double)e.y * s), (T)((double)e.y * c - (double)e.x * s));
    }
    // symbols without a full window around them stay zero (:60, 70, 72)
    if (blockIdx.x == 0)
        for (long long k = tid; k < off; k += VV_THREADS) orow[k] = make_cx<T>((T)0, (T)0);
    if (blockIdx.x == gridDim.x - 1)
        for (long long k = off + nwin + tid; k < L; k += VV_THREADS) orow[k] = make_cx<T>((T)0, (T)0);
}

size_t vv_work_bytes(int dtype, int64_t nrows, int64_t L, int64_t N)
{
    const int64_t ntiles = (L - N + 1 + VV_TILE - 1) / VV_TILE;
    return (size_t)nrows * ntiles * (sizeof(int) + (dtype == QB_C64 ? 4 : 8));
}

template <typename T>
static int vv_launch(const void *E, int64_t nrows, int64_t row_stride, int64_t L, int N, int M, void *out,
                     int64_t out_stride, void *ph, int64_t ph_stride, void *work, cudaStream_t st)
{
    const int64_t nwin = L - N + 1;
    const int ntiles = (int)((nwin + VV_TILE - 1) / VV_TILE);
    T *tile_prev = reinterpret_cast<T *>(work);                        // T first: keeps doubles 8-byte aligned
    int *tile_turns = reinterpret_cast<int *>(tile_prev + (size_t)nrows * ntiles);
    const size_t smem = (size_t)(VV_TILE + N) * sizeof(double2) + (size_t)(VV_TILE + 1) * sizeof(T);
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(vv_phase_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((VV_TILE + VV_MAX_N) * sizeof(double2) + (VV_TILE + 1) * sizeof(T))));
    dim3 grid((unsigned)ntiles, (unsigned)nrows);
    vv_phase_kernel<T><<<grid, VV_THREADS, smem, st>>>((const cx<T> *)E, row_stride, L, N, M, (T *)ph, ph_stride,
                                                       tile_turns, tile_prev, ntiles);
    vv_scan_kernel<<<(unsigned)nrows, 1024, 0, st>>>(tile_turns, ntiles);
    vv_apply_kernel<T><<<grid, VV_THREADS, 0, st>>>((const cx<T> *)E, row_stride, L, N, M, (T *)ph, ph_stride,
                                                    tile_turns, tile_prev, ntiles, (cx<T> *)out, out_stride);
    count_launch(3);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int vv_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, int64_t N, int64_t M, void *out,
                int64_t out_stride, void *ph, int64_t ph_stride, void *work, cudaStream_t st)
{
    if (nrows == 0) return QB_OK;
    if (N > VV_MAX_N) return set_error(QB_ERR_UNSUPPORTED, "viterbiviterbi: N up to %d", VV_MAX_N);
    if (dtype == QB_C64)
        return vv_launch<float>(E, nrows, row_stride, L, (int)N, (int)M, out, out_stride, ph, ph_stride, work, st);
    return vv_launch<double>(E, nrows, row_stride, L, (int)N, (int)M, out, out_stride, ph, ph_stride, work, st);
}

}  // namespace qb
