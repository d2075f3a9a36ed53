// bps: blind phase search carrier recovery (replaces bps + select_angle_index + select_angles,
// qampy/core/pythran_dsp.py:47-85, 26-42, 137-153, and the L2 tail qampy/core/phaserecovery.py:150-159).
//
// Per 1-D stream E[0..L):
//   dists[i,a] = min(min_m |E[i]*comp[a] - s_m|^2, 100)                 (pythran_dsp.py:73-84)
//   csum[0,a]  = 0 ; csum[i,a] = csum[i-1,a] + dists[i,a]  (i >= 1)       (:30-37, SEQUENTIAL in fp)
//   for i >= 2N: idx[i-N] = first strict argmin_a (csum[i,a] - csum[i-2N,a]), dmin0 = 1000   (:38-41)
//   ph = angles[idx]; ph[N:L-N] = unwrap(4*ph[N:L-N])/4 ; Eout = E*exp(1j*ph)   (phaserecovery.py:150-159)
//
// The (L, A) distance and running-sum matrices of the reference (2 x 2.56 GB per polarisation at
// L = 1e7, A = 64) never exist: one CTA walks one stream tile by tile, keeps the last 2N+tile rows of
// the running sum in a shared-memory ring and fuses distance -> running sum -> window difference ->
// arg-min -> angle gather -> unwrap -> rotation.  The running sum is carried in one register per
// angle column and added in the reference's order, so the selected indices are bit-identical to
// the reference even where its fp32 sum has lost precision.
//
// Arithmetic contract of the distance: unfused complex multiply, |z|^2 = fl(fl(re*re) +
// fl(im*im)); for a full rectangular alphabet the per-axis slicer returns the same bits as the brute
// force search because IEEE rounding is monotone.
#include "qb_common.cuh"

namespace qb {

constexpr int BPS_THREADS = 256;

template <typename T>
struct BpsParams {
    const cx<T> *E;
    const cx<T> *comp;
    const T *angles;
    const cx<T> *symbols;
    const T *lev_re, *lev_im;
    int32_t *idx;
    T *ph;
    cx<T> *Eout;
    long long stream_stride, L;
    int A, M, n_re, n_im, N;
    int tile_rows, ring_rows;
};

// Per-axis slicer.  `pairs[f] = (lev[f], lev[f+1])` are the two levels bracketing a value whose
// (approximate) grid coordinate floors to f; the nearest level is always one of them, and because the
// subtraction uses the STORED level values the result is bit-identical to the brute-force minimum.
template <typename T>
__device__ __forceinline__ T axis_min(T t, const cx<T> *pairs, int npair, T lev0, T inv_step)
{
    T uf = floor((t - lev0) * inv_step);
    uf = uf > (T)0 ? uf : (T)0;  // NaN -> 0
    uf = uf < (T)(npair - 1) ? uf : (T)(npair - 1);
    const cx<T> l = pairs[(int)uf];
    return fmin(fabs(sub_rn(t, l.x)), fabs(sub_rn(t, l.y)));  // NaN only if t is NaN
}

template <typename T>
struct Pi;
template <>
struct Pi<float> {
    static __device__ __forceinline__ float pi() { return 3.14159274101257324219f; }      // fl32(pi)
    static __device__ __forceinline__ float two_pi() { return 6.28318548202514648438f; }  // fl32(2 pi)
};
template <>
struct Pi<double> {
    static __device__ __forceinline__ double pi() { return 3.141592653589793115997963; }
    static __device__ __forceinline__ double two_pi() { return 6.283185307179586231995927; }
};

// one step of np.unwrap's correction (numpy/lib/_function_base_impl.py unwrap, default period/discont)
template <typename T>
__device__ __forceinline__ T unwrap_corr(T p, T pprev)
{
    const T PI = Pi<T>::pi(), TWO_PI = Pi<T>::two_pi();
    const T dd = sub_rn(p, pprev);
    T m = fmod(add_rn(dd, PI), TWO_PI);  // np.mod: python-style, divisor > 0
    if (m != (T)0 && m < (T)0) m = add_rn(m, TWO_PI);
    T ddmod = sub_rn(m, PI);
    if (ddmod == -PI && dd > (T)0) ddmod = PI;
    T corr = sub_rn(ddmod, dd);
    if (fabs(dd) < PI) corr = (T)0;
    return corr;
}

__device__ __forceinline__ void qb_sincos(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void qb_sincos(double x, double *s, double *c) { sincos(x, s, c); }

// E * exp(1j*ph), phaserecovery.py:157-159
template <typename T>
__device__ __forceinline__ cx<T> rotate(cx<T> e, T ph)
{
    T s, c;
    qb_sincos(ph, &s, &c);
    return make_cx<T>(e.x * c - e.y * s, e.x * s + e.y * c);
}

template <typename T>
__global__ void __launch_bounds__(BPS_THREADS) bps_kernel(BpsParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.A, W = 2 * p.N, N = p.N, TR = p.tile_rows;
    const int RMASK = p.ring_rows - 1;  // ring_rows is a power of two >= TR + W
    const long long L = p.L;

    cx<T> *comp = reinterpret_cast<cx<T> *>(smem_raw);  // [A]
    cx<T> *syms = comp + A;                             // [M] (brute force) or level pairs (slicer)
    const bool slicer = p.n_re > 0;
    const int npr = slicer ? max(p.n_re - 1, 1) : 0, npi = slicer ? max(p.n_im - 1, 1) : 0;
    cx<T> *pre = syms, *pim = syms + npr;
    T *ring = reinterpret_cast<T *>(syms + (slicer ? npr + npi : p.M));  // [RR][A] running sums
    T *dt = ring + (size_t)p.ring_rows * A;             // [TR][A] window differences
    T *angs = dt + (size_t)TR * A;                      // [A]
    T *p4s = angs + A;                                  // [TR] 4*angle of the tile's output rows
    int *kidx = reinterpret_cast<int *>(p4s + TR);      // [TR]

    const cx<T> *E = p.E + (long long)blockIdx.x * p.stream_stride;
    int32_t *idx = p.idx ? p.idx + (long long)blockIdx.x * L : nullptr;
    T *ph = p.ph ? p.ph + (long long)blockIdx.x * L : nullptr;
    cx<T> *Eout = p.Eout ? p.Eout + (long long)blockIdx.x * L : nullptr;

    for (int c = tid; c < A; c += BPS_THREADS) {
        comp[c] = p.comp[c];
        angs[c] = p.angles ? p.angles[c] : (T)0;
    }
    T re0 = 0, rinv = 0, im0 = 0, iinv = 0;
    if (slicer) {
        for (int c = tid; c < npr; c += BPS_THREADS)
            pre[c] = make_cx<T>(p.lev_re[c], p.lev_re[min(c + 1, p.n_re - 1)]);
        for (int c = tid; c < npi; c += BPS_THREADS)
            pim[c] = make_cx<T>(p.lev_im[c], p.lev_im[min(c + 1, p.n_im - 1)]);
        re0 = p.lev_re[0];
        im0 = p.lev_im[0];
        rinv = p.n_re > 1 ? (T)(p.n_re - 1) / (p.lev_re[p.n_re - 1] - re0) : (T)0;
        iinv = p.n_im > 1 ? (T)(p.n_im - 1) / (p.lev_im[p.n_im - 1] - im0) : (T)0;
    } else {
        for (int c = tid; c < p.M; c += BPS_THREADS) syms[c] = p.symbols[c];
    }
    __syncthreads();

    // edges: idx = 0 -> ph = angles[0], not unwrapped (phaserecovery.py:155 touches [N:-N] only)
    const long long lo = N < L ? N : L;                  // rows [0, lo) are left edge
    const long long hi = (L - N > lo) ? L - N : lo;      // rows [hi, L) are right edge
    {
        const T a0 = angs[0];
        const long long nedge = lo + (L - hi);
        for (long long c = tid; c < nedge; c += BPS_THREADS) {
            const long long j = c < lo ? c : hi + (c - lo);
            if (idx) idx[j] = 0;
            if (ph) ph[j] = a0;
            if (Eout) Eout[j] = rotate<T>(E[j], a0);
        }
    }

    // phase-1 mapping: when A divides the block, a thread keeps one angle (its rotation in registers)
    const bool fixed = (BPS_THREADS % A) == 0;
    const int my_a = tid % A, my_r0 = tid / A, rstep = fixed ? BPS_THREADS / A : 1;
    const cx<T> my_c = comp[my_a];

    T csum = 0;                 // running column sum, owned by thread a < A
    T cum = 0, p4prev = 0;      // unwrap state, replicated in warp 0
    int slot0 = 0;              // ring slot of the tile's first row

    for (long long i0 = 0; i0 < L; i0 += TR, slot0 = (slot0 + TR) & RMASK) {
        const int nrows = (int)min((long long)TR, L - i0);
        // ---- phase 1: distances of the tile into the ring ---------------------------------------
        if (fixed) {
            for (int r = my_r0; r < nrows; r += rstep) {
                const cx<T> e = E[i0 + r];
                const T tr = sub_rn(mul_rn(e.x, my_c.x), mul_rn(e.y, my_c.y));
                const T ti = add_rn(mul_rn(e.x, my_c.y), mul_rn(e.y, my_c.x));
                T d;
                if (slicer) {
                    const T da = axis_min<T>(tr, pre, npr, re0, rinv);
                    const T db = axis_min<T>(ti, pim, npi, im0, iinv);
                    d = add_rn(mul_rn(da, da), mul_rn(db, db));
                } else {
                    d = (T)1000.;
                    for (int m = 0; m < p.M; m++) {
                        const cx<T> sy = syms[m];
                        const T dr = sub_rn(tr, sy.x), di = sub_rn(ti, sy.y);
                        const T dd = add_rn(mul_rn(dr, dr), mul_rn(di, di));
                        if (dd < d) d = dd;
                    }
                }
                ring[((slot0 + r) & RMASK) * A + my_a] = d < (T)100. ? d : (T)100.;
            }
        } else {
            for (int f = tid; f < nrows * A; f += BPS_THREADS) {
                const int r = f / A, a = f - r * A;
                const cx<T> e = E[i0 + r];
                const cx<T> c = comp[a];
                const T tr = sub_rn(mul_rn(e.x, c.x), mul_rn(e.y, c.y));
                const T ti = add_rn(mul_rn(e.x, c.y), mul_rn(e.y, c.x));
                T d;
                if (slicer) {
                    const T da = axis_min<T>(tr, pre, npr, re0, rinv);
                    const T db = axis_min<T>(ti, pim, npi, im0, iinv);
                    d = add_rn(mul_rn(da, da), mul_rn(db, db));
                } else {
                    d = (T)1000.;
                    for (int m = 0; m < p.M; m++) {
                        const cx<T> sy = syms[m];
                        const T dr = sub_rn(tr, sy.x), di = sub_rn(ti, sy.y);
                        const T dd = add_rn(mul_rn(dr, dr), mul_rn(di, di));
                        if (dd < d) d = dd;
                    }
                }
                ring[((slot0 + r) & RMASK) * A + a] = d < (T)100. ? d : (T)100.;
            }
        }
        __syncthreads();
        // ---- phase 2: sequential running sum per angle column + window difference ---------------
        if (tid < A) {
            const bool full = i0 >= W && i0 > 0;   // every row of the tile has a complete window
#pragma unroll 4
            for (int r = 0; r < nrows; r++) {
                T *slot = ring + ((slot0 + r) & RMASK) * A + tid;
                const T old = ring[((slot0 + r - W) & RMASK) * A + tid];   // csum[i - W] (garbage if i < W)
                csum = (i0 + r == 0) ? (T)0 : add_rn(csum, *slot);         // row 0 is never added (:30)
                *slot = csum;
                if (full || i0 + r >= W) dt[r * A + tid] = sub_rn(csum, old);
            }
        }
        __syncthreads();
        // ---- phase 3: first strict arg-min over angles per row ----------------------------------
        const int r_first = (int)max((long long)0, (long long)W - i0);
        for (int r = r_first + warp; r < nrows; r += BPS_THREADS / 32) {
            T best = (T)1000.;
            int bk = 0x7fffffff;
            for (int a = lane; a < A; a += 32) {
                const T v = dt[r * A + a];
                if (v < best) {
                    best = v;
                    bk = a;
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const T ob = shfl_xor(best, m);
                const int ok = shfl_xor(bk, m);
                if (ob < best || (ob == best && ok < bk)) {
                    best = ob;
                    bk = ok;
                }
            }
            if (lane == 0) {
                const int k = bk == 0x7fffffff ? 0 : bk;
                kidx[r] = k;
                p4s[r] = mul_rn(angs[k], (T)4);
            }
        }
        __syncthreads();
        // ---- phase 4: np.unwrap(4*ph)/4 over the output rows j = i - N, sequential in fp ----------
        // rows r >= r_first produce output j = i0 + r - N; the first output overall is j = N.
        if (warp == 0 && ph) {
            for (int rb = r_first; rb < nrows; rb += 32) {
                const int r = rb + lane;
                const bool valid = r < nrows;
                const long long j = i0 + r - N;
                const T p4 = valid ? p4s[r] : (T)0;
                T pp = p4prev;                                   // lane 0: carried from the previous chunk
                if (lane > 0 && valid) pp = p4s[r - 1];
                T corr = (T)0;
                if (valid && j > N) corr = unwrap_corr<T>(p4, pp);
                unsigned mask = __ballot_sync(0xffffffffu, corr != (T)0);
                T mycum = cum;
                while (mask) {
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const T ce = shfl_idx(corr, e);
                    cum = add_rn(cum, ce);
                    if (lane >= e) mycum = cum;
                }
                const T phv = add_rn(p4, mycum) / (T)4;
                if (valid) {
                    ph[j] = phv;
                    p4s[r] = phv;                                // phase 5 reads the final phase from here
                }
                const int nvalid = min(32, nrows - rb);
                p4prev = shfl_idx(p4, nvalid - 1);
            }
        }
        if (idx) {
            for (int r = r_first + tid; r < nrows; r += BPS_THREADS) idx[i0 + r - N] = kidx[r];
        }
        // ---- phase 5: rotate the tile's output rows ----------------------------------------------
        if (Eout) {
            __syncthreads();
            for (int r = r_first + tid; r < nrows; r += BPS_THREADS) {
                const long long j = i0 + r - N;
                Eout[j] = rotate<T>(E[j], p4s[r]);
            }
        }
        __syncthreads();
    }
}

template <typename T>
static int launch_bps(const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                      const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                      const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                      int32_t *idx, void *ph, void *Eout, cudaStream_t st)
{
    if (nstream == 0 || L == 0) return QB_OK;
    BpsParams<T> p;
    p.E = (const cx<T> *)E;
    p.comp = (const cx<T> *)comp;
    p.angles = (const T *)angles;
    p.symbols = (const cx<T> *)symbols;
    p.lev_re = (const T *)lev_re;
    p.lev_im = (const T *)lev_im;
    p.idx = idx;
    p.ph = (T *)ph;
    p.Eout = (cx<T> *)Eout;
    p.stream_stride = stream_stride;
    p.L = L;
    p.A = (int)A;
    p.M = (int)M;
    p.n_re = (int)n_re;
    p.n_im = (int)n_im;
    p.N = (int)N;
    const int W = 2 * (int)N;
    // tile rows: fill the power-of-two ring (>= TR + 2N rows) as well as possible, keep <= ~48 KB/CTA
    // so that several CTAs share an SM and hide each other's serial phases
    int TR = 0, RR = 0;
    size_t smem = 0;
    for (int rr = 32; rr <= 65536 && !TR; rr <<= 1) {
        if (rr <= W) continue;
        int tr = rr - W < 64 ? rr - W : 64;
        if (tr < 8 && rr < 65536) continue;
        const size_t need = ((size_t)rr * A + (size_t)tr * A) * sizeof(T) +
                            (A + (n_re ? n_re + n_im : M)) * sizeof(cx<T>) + (A + tr) * sizeof(T) +
                            tr * sizeof(int) + 64;
        if (need > 200 * 1024) break;
        TR = tr;
        RR = rr;
        smem = need;
    }
    if (!TR) return set_error(QB_ERR_UNSUPPORTED, "bps: 2N*A too large for the shared-memory ring");
    p.tile_rows = TR;
    p.ring_rows = RR;
    static bool attr_done[2] = {false, false};
    if (!attr_done[sizeof(T) == 8]) {
        QB_CUDA_CHECK(cudaFuncSetAttribute(bps_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
        attr_done[sizeof(T) == 8] = true;
    }
    if (nstream > 2147483647LL) return set_error(QB_ERR_UNSUPPORTED, "bps: too many streams");
    bps_kernel<T><<<(unsigned)nstream, BPS_THREADS, smem, st>>>(p);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int bps_dispatch(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                 const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                 const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                 int32_t *idx, void *ph, void *Eout, cudaStream_t st)
{
    if (dtype == QB_C64)
        return launch_bps<float>(E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re,
                                 lev_im, n_im, N, idx, ph, Eout, st);
    return launch_bps<double>(E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re,
                              lev_im, n_im, N, idx, ph, Eout, st);
}

// ---- select_angles -----------------------------------------------------------------------------------
template <typename T>
__global__ void select_angles_kernel(const T *angles, long long p, int A, const int64_t *idx, long long L,
                                     T *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) out[i] = angles[(p > 1 ? i : 0) * A + idx[i]];
}

int select_angles_dispatch(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx,
                           int64_t L, void *out, cudaStream_t st)
{
    if (L == 0) return QB_OK;
    const unsigned nb = (unsigned)((L + 255) / 256);
    if (dtype == QB_C64)
        select_angles_kernel<float><<<nb, 256, 0, st>>>((const float *)angles, p, (int)A, idx, L, (float *)out);
    else
        select_angles_kernel<double><<<nb, 256, 0, st>>>((const double *)angles, p, (int)A, idx, L, (double *)out);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb
