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
#include <stdlib.h>

#include "bps_generic.cuh"

namespace qb {

// One CTA walks one stream in tiles of TR = 32 rows:
//   phase 1 (all warps)  distances of the tile -> ring[row][angle]
//   phase 2 (A threads)  sequential running sum per angle column, window difference -> dt[angle][row]
//   phase 3 (warp 0)     lane = row: first strict arg-min over the angles, then in the same warp the
//                        sequential-order unwrap, the phase output and the rotation of the symbol
// Several CTAs share an SM, so the serial phases of one stream overlap the parallel phase of another.
constexpr int BPS_TR = 32;

template <typename T>
__global__ void __launch_bounds__(BPS_THREADS) bps_kernel(BpsParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.A, W = 2 * p.N, N = p.N;
    constexpr int TR = BPS_TR, TRP = BPS_TR + 1;
    const int RMASK = p.ring_rows - 1;  // ring_rows: power of two, multiple of TR, >= TR + W
    const long long L = p.L;

    cx<T> *comp = reinterpret_cast<cx<T> *>(smem_raw);  // [A]
    cx<T> *Et = comp + A;                               // [TR] the tile's input rows
    cx<T> *syms = Et + TR;                              // [M] (brute force) or level pairs (slicer)
    const bool slicer = p.n_re > 0;
    const int npr = slicer ? max(p.n_re - 1, 1) : 0, npi = slicer ? max(p.n_im - 1, 1) : 0;
    cx<T> *pre = syms, *pim = syms + npr;
    T *ring = reinterpret_cast<T *>(syms + (slicer ? npr + npi : p.M));  // [RR][A] running sums
    T *dt = ring + (size_t)p.ring_rows * A;             // [A][TRP] window differences (transposed)
    T *angs = dt + (size_t)A * TRP;                     // [A]

    const cx<T> *E = p.E + (long long)blockIdx.x * p.stream_stride;
    int32_t *idx = p.idx ? p.idx + (long long)blockIdx.x * L : nullptr;
    T *ph = p.ph ? p.ph + (long long)blockIdx.x * L : nullptr;
    cx<T> *Eout = p.Eout ? p.Eout + (long long)blockIdx.x * L : nullptr;

    for (int c = tid; c < A; c += BPS_THREADS) {
        comp[c] = p.comp[c];
        angs[c] = p.angles ? p.angles[c] : (T)0;
    }
    AxisGrid<T> gre, gim;
    gre.scale = gre.bias = gim.scale = gim.bias = (T)0;
    gre.npair = gim.npair = 1;
    if (slicer) {
        for (int c = tid; c < npr; c += BPS_THREADS)
            pre[c] = make_cx<T>(p.lev_re[c], p.lev_re[min(c + 1, p.n_re - 1)]);
        for (int c = tid; c < npi; c += BPS_THREADS)
            pim[c] = make_cx<T>(p.lev_im[c], p.lev_im[min(c + 1, p.n_im - 1)]);
        gre = make_grid(p.lev_re, p.n_re);
        gim = make_grid(p.lev_im, p.n_im);
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
    int slot0 = 0;              // ring slot of the tile's first row (multiple of TR: a tile never wraps)

    for (long long i0 = 0; i0 < L; i0 += TR, slot0 = (slot0 + TR) & RMASK) {
        const int nrows = (int)min((long long)TR, L - i0);
        if (tid < nrows) Et[tid] = E[i0 + tid];
        __syncthreads();
        // ---- phase 1: distances of the tile into the ring ---------------------------------------
        if (p.comp_rows) {
            // per-symbol test angles: comp[(i0 + r) * A + a] of this stream (pythran_dsp.py:74-77, ph_idx = i)
            const cx<T> *crow = p.comp + ((long long)blockIdx.x * L + i0) * A;
            for (int f = tid; f < nrows * A; f += BPS_THREADS) {
                const int r = f / A, a = f - r * A;
                ring[(size_t)(slot0 + r) * A + a] =
                    min_distance<T>(Et[r], crow[f], slicer, pre, pim, gre, gim, syms, p.M);
            }
        } else if (fixed) {
            T *dst = ring + (size_t)slot0 * A + my_a;
            for (int r = my_r0; r < nrows; r += rstep)
                dst[r * A] = min_distance<T>(Et[r], my_c, slicer, pre, pim, gre, gim, syms, p.M);
        } else {
            for (int f = tid; f < nrows * A; f += BPS_THREADS) {
                const int r = f / A, a = f - r * A;
                ring[(size_t)(slot0 + r) * A + a] =
                    min_distance<T>(Et[r], comp[a], slicer, pre, pim, gre, gim, syms, p.M);
            }
        }
        __syncthreads();
        // ---- phase 2: sequential running sum per angle column + window difference ---------------
        if (p.windowed) {
            // QB_BPS_WINDOWED: the ring keeps the distances themselves and every window sum is formed directly from
            // its 2N terms, accumulated in double -- no running sum whose magnitude (2e5 after 1e7 rows) swallows
            // the differences between neighbouring test angles.  Not the reference's bits: its exact fp32 running
            // sum is the default mode.  All threads share the (row, angle) pairs of the tile.
            for (int f = tid; f < nrows * A; f += BPS_THREADS) {
                const int r = f / A, a = f - r * A;
                if (i0 + r >= W) {
                    double acc = 0.0;
                    for (int w = W - 1; w >= 0; w--)          // oldest row first, like the reference's running sum
                        acc += (double)ring[(size_t)((slot0 + r - w) & RMASK) * A + a];
                    dt[(size_t)a * TRP + r] = (T)acc;
                }
            }
        } else if (tid < A) {
            T *xp = ring + (size_t)slot0 * A + tid;
            T *dp = dt + (size_t)tid * TRP;
#pragma unroll 4
            for (int r = 0; r < nrows; r++) {
                const T old = ring[(size_t)((slot0 + r - W) & RMASK) * A + tid];   // csum[i - W]; unused if i < W
                csum = (i0 + r == 0) ? (T)0 : add_rn(csum, xp[r * A]);             // row 0 is never added (:30)
                xp[r * A] = csum;
                dp[r] = sub_rn(csum, old);
            }
        }
        __syncthreads();
        // ---- phase 3 (warp 0): arg-min, unwrap, outputs -------------------------------------------
        // row r (lane) with i0 + r >= W produces output j = i0 + r - N; the first output overall is j = N
        if (warp == 0) {
            const int r = lane;
            const long long i = i0 + r, j = i - N;
            const bool valid = r < nrows && i >= W;
            T best = (T)1000.;
            int bk = 0;
#pragma unroll 8
            for (int a = 0; a < A; a++) {
                const T v = dt[a * TRP + r];
                if (v < best) {   // strict: first minimum, dmin0 = 1000 (:31, :39)
                    best = v;
                    bk = a;
                }
            }
            if (idx && valid) idx[j] = bk;
            if (ph) {
                // select_angles: angles[0, idx] or, with a per-symbol table, angles[j, idx] (pythran_dsp.py:137-153)
                const T ang = (p.comp_rows && valid) ? p.angles[((long long)blockIdx.x * L + j) * A + bk] : angs[bk];
                const T p4 = mul_rn(ang, (T)4);
                T pp = __shfl_up_sync(0xffffffffu, p4, 1);
                if (lane == 0) pp = p4prev;
                T corr = (T)0;
                if (valid && j > N) corr = unwrap_corr<T>(p4, pp);
                unsigned mask = __ballot_sync(0xffffffffu, corr != (T)0);
                T mycum = cum;
                while (mask) {   // fold the (rare) non-zero corrections in order: exact sequential fp sum
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const T ce = shfl_idx(corr, e);
                    cum = add_rn(cum, ce);
                    if (lane >= e) mycum = cum;
                }
                const T phv = add_rn(p4, mycum) / (T)4;
                if (valid) {
                    ph[j] = phv;
                    if (Eout) Eout[j] = rotate<T>(E[j], phv);
                }
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                if (vm) p4prev = shfl_idx(p4, 31 - __clz(vm));   // last valid row of this tile
            }
        }
        // no barrier needed here: the next tile's first barrier orders phase 3's reads of dt against
        // the next phase 2, and Et is only rewritten after every warp passed the phase-2 barrier
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised BPS kernel (default).  Same arithmetic as bps_kernel, but the parallel distance
// search and the serial tail run CONCURRENTLY inside a CTA:
//   producer warps (4)      distances of tile m+1 -> ring rows, one angle column per thread
//   consumer warps (A/32)   running sums + window differences of tile m (thread = angle column), then
//                           consumer warp 0: arg-min (16 rows x 2 angle halves), unwrap, phase output, rotation
// hand-off by named barriers (bar.arrive / bar.sync): FULL[m&1] producers -> consumers, EMPTY[m&1] back.
// A tile is 16 rows; the power-of-two ring holds >= 2N + 32 rows so the producers may run one tile
// ahead of the consumers without touching a row a window difference still needs.
// ------------------------------------------------------------------------------------------------
constexpr int WS_TR = 16;
constexpr int WS_PW = 4;   // producer warps

// barrier ids are immediates so that the kernel reserves 6 named barriers, not all 16
template <int ID>
__device__ __forceinline__ void named_bar_sync(int count)
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_bar_arrive(int count)
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID0>
__device__ __forceinline__ void named_bar_sync2(int parity, int count)
{
    if (parity) named_bar_sync<ID0 + 1>(count); else named_bar_sync<ID0>(count);
}
template <int ID0>
__device__ __forceinline__ void named_bar_arrive2(int parity, int count)
{
    if (parity) named_bar_arrive<ID0 + 1>(count); else named_bar_arrive<ID0>(count);
}

template <typename T>
__global__ void __launch_bounds__(32 * (WS_PW + 8)) bps_ws_kernel(BpsParams<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.A, W = 2 * p.N, N = p.N;
    constexpr int TR = WS_TR, TRP = WS_TR + 1;
    const int CW = (A + 31) >> 5;                     // consumer warps
    const int NT = (int)blockDim.x;                   // 32 * (CW + WS_PW)
    const int RMASK = p.ring_rows - 1;
    const long long L = p.L;
    const int Ah = (A + 1) >> 1;                      // first half of the angles (arg-min split)

    cx<T> *comp = reinterpret_cast<cx<T> *>(smem_raw);  // [A]
    cx<T> *syms = comp + A;
    const bool slicer = p.n_re > 0;
    const int npr = slicer ? max(p.n_re - 1, 1) : 0, npi = slicer ? max(p.n_im - 1, 1) : 0;
    cx<T> *pre = syms, *pim = syms + npr;
    T *ring = reinterpret_cast<T *>(syms + (slicer ? npr + npi : p.M));  // [RR][A]
    T *dt = ring + (size_t)p.ring_rows * A;             // [A][TRP], upper angle half shifted by 16 banks
    T *angs = dt + (size_t)A * TRP + 32;                // [A]

    const cx<T> *E = p.E + (long long)blockIdx.x * p.stream_stride;
    int32_t *idx = p.idx ? p.idx + (long long)blockIdx.x * L : nullptr;
    T *ph = p.ph ? p.ph + (long long)blockIdx.x * L : nullptr;
    cx<T> *Eout = p.Eout ? p.Eout + (long long)blockIdx.x * L : nullptr;

    for (int c = tid; c < A; c += NT) {
        comp[c] = p.comp[c];
        angs[c] = p.angles ? p.angles[c] : (T)0;
    }
    AxisGrid<T> gre, gim;
    gre.scale = gre.bias = gim.scale = gim.bias = (T)0;
    gre.npair = gim.npair = 1;
    if (slicer) {
        for (int c = tid; c < npr; c += NT) pre[c] = make_cx<T>(p.lev_re[c], p.lev_re[min(c + 1, p.n_re - 1)]);
        for (int c = tid; c < npi; c += NT) pim[c] = make_cx<T>(p.lev_im[c], p.lev_im[min(c + 1, p.n_im - 1)]);
        gre = make_grid(p.lev_re, p.n_re);
        gim = make_grid(p.lev_im, p.n_im);
    } else {
        for (int c = tid; c < p.M; c += NT) syms[c] = p.symbols[c];
    }
    // edges: idx = 0 -> ph = angles[0], not unwrapped (phaserecovery.py:155 touches [N:-N] only)
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    __syncthreads();
    {
        const T a0 = angs[0];
        const long long nedge = lo + (L - hi);
        for (long long c = tid; c < nedge; c += NT) {
            const long long j = c < lo ? c : hi + (c - lo);
            if (idx) idx[j] = 0;
            if (ph) ph[j] = a0;
            if (Eout) Eout[j] = rotate<T>(E[j], a0);
        }
    }
    const long long ntiles = (L + TR - 1) / TR;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_CONS = 5;

    if (warp >= CW) {
        // =========================== producers: distance search ===================================
        const int pt = tid - 32 * CW, NP = 32 * WS_PW;
        const bool fixed = (NP % A) == 0;
        const int my_a = pt % A, my_r0 = pt / A, rstep = fixed ? NP / A : 1;
        const cx<T> my_c = comp[my_a];
        for (long long m = 0; m < ntiles; m++) {
            const long long i0 = m * TR;
            const int nrows = (int)min((long long)TR, L - i0);
            const int slot0 = (int)(i0 & RMASK);
            if (m >= 2) named_bar_sync2<BAR_EMPTY>((int)(m & 1), NT);   // tile m-2 consumed: ring rows free
            if (fixed) {
                T *dst = ring + (size_t)slot0 * A + my_a;
#pragma unroll 4
                for (int r = my_r0; r < nrows; r += rstep)
                    dst[r * A] = min_distance<T>(E[i0 + r], my_c, slicer, pre, pim, gre, gim, syms, p.M);
            } else {
                for (int f = pt; f < nrows * A; f += NP) {
                    const int r = f / A, a = f - r * A;
                    ring[(size_t)(slot0 + r) * A + a] =
                        min_distance<T>(E[i0 + r], comp[a], slicer, pre, pim, gre, gim, syms, p.M);
                }
            }
            named_bar_arrive2<BAR_FULL>((int)(m & 1), NT);
        }
    } else {
        // =========================== consumers: serial tail ========================================
        T csum = 0;                 // running column sum of angle column `tid` (tid < A)
        T cum = 0, p4prev = 0;      // unwrap state (consumer warp 0)
        const int NC = 32 * CW;
        for (long long m = 0; m < ntiles; m++) {
            const long long i0 = m * TR;
            const int nrows = (int)min((long long)TR, L - i0);
            const int slot0 = (int)(i0 & RMASK);
            named_bar_sync2<BAR_FULL>((int)(m & 1), NT);
            if (tid < A) {
                T *xp = ring + (size_t)slot0 * A + tid;
                T *dp = dt + (size_t)tid * TRP + (tid >= Ah ? 16 : 0);
#pragma unroll 4
                for (int r = 0; r < nrows; r++) {
                    const T old = ring[(size_t)((slot0 + r - W) & RMASK) * A + tid];   // csum[i-W]; unused if i < W
                    csum = (i0 + r == 0) ? (T)0 : add_rn(csum, xp[r * A]);             // row 0 is never added (:30)
                    xp[r * A] = csum;
                    dp[r] = sub_rn(csum, old);
                }
            }
            named_bar_arrive2<BAR_EMPTY>((int)(m & 1), NT);     // ring rows of this tile are final
            named_bar_sync<BAR_CONS>(NC);                        // dt complete
            if (warp == 0) {
                const int r = lane & 15, h = lane >> 4;
                const long long i = i0 + r, j = i - N;
                const bool valid = r < nrows && i >= W && h == 0;
                const int a0 = h ? Ah : 0, a1 = h ? A : Ah;
                T best = (T)1000.;
                int bk = 0x7fffffff;
                const T *dr = dt + r + (h ? 16 : 0);
#pragma unroll 8
                for (int a = a0; a < a1; a++) {
                    const T v = dr[a * TRP];
                    if (v < best) {   // strict: first minimum, dmin0 = 1000 (:31, :39)
                        best = v;
                        bk = a;
                    }
                }
                {   // lower half wins ties (smaller angle index)
                    const T ob = shfl_xor(best, 16);
                    const int ok = shfl_xor(bk, 16);
                    if (ob < best || (ob == best && ok < bk)) {
                        best = ob;
                        bk = ok;
                    }
                }
                if (bk == 0x7fffffff) bk = 0;
                named_bar_sync<BAR_CONS>(NC);                    // dt may be rewritten by the next phase 2
                if (idx && valid) idx[j] = bk;
                if (ph) {
                    const T p4 = mul_rn(angs[bk], (T)4);
                    T pp = __shfl_up_sync(0xffffffffu, p4, 1);
                    if (r == 0) pp = p4prev;
                    T corr = (T)0;
                    if (valid && j > N) corr = unwrap_corr<T>(p4, pp);
                    unsigned mask = __ballot_sync(0xffffffffu, corr != (T)0);
                    T mycum = cum;
                    while (mask) {   // fold the (rare) non-zero corrections in row order: exact sequential sum
                        const int e = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const T ce = shfl_idx(corr, e);
                        cum = add_rn(cum, ce);
                        if (lane >= e) mycum = cum;
                    }
                    const T phv = add_rn(p4, mycum) / (T)4;
                    if (valid) {
                        ph[j] = phv;
                        if (Eout) Eout[j] = rotate<T>(E[j], phv);
                    }
                    const unsigned vm = __ballot_sync(0xffffffffu, valid);
                    if (vm) p4prev = shfl_idx(p4, 31 - __clz(vm));
                }
            } else {
                named_bar_sync<BAR_CONS>(NC);
            }
        }
    }
}

// bps_fast.cu: column-per-lane kernel (complex64, rectangular alphabet, A in {32,64,96,128});
// returns 1 when the problem is not covered.
bool bps_par_wanted(int64_t nstream, int64_t L, int64_t A, bool own_idx, int elem);               // bps_par.cu
template <typename T>
int bps_par_generic_launch(const BpsParams<T> &p, int64_t nstream, cudaStream_t st);

int bps_fast_dispatch(const void *E, int64_t nstream, int64_t stream_stride, int64_t L, const void *comp,
                      const void *angles, int64_t A, const void *lev_re, int64_t n_re, const void *lev_im,
                      int64_t n_im, int64_t N, int32_t *idx, void *ph, void *Eout, cudaStream_t st);

// Accumulation mode of the calling thread's bps launches (qb_set_bps_accumulation): 0 = exact (the reference's
// sequential fp32 / fp64 running sum, bit-exact indices), 1 = windowed (direct 2N-term sums, more accurate).
static thread_local int g_bps_accum = 0;
int bps_accumulation() { return g_bps_accum; }
int set_bps_accumulation(int mode)
{
    const int old = g_bps_accum;
    g_bps_accum = mode == 1 ? 1 : 0;
    return old;
}

template <typename T>
static int launch_bps(const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                      const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                      const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                      int32_t *idx, void *ph, void *Eout, int comp_rows, cudaStream_t st)
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
    p.comp_rows = comp_rows;
    p.windowed = bps_accumulation() == 1;
    const int W = 2 * (int)N;
    // ring: power of two >= TR + 2N rows (so a tile never wraps and slots are a mask away)
    int RR = 2 * BPS_TR;
    while (RR < BPS_TR + W) RR <<= 1;
    const size_t smem = ((size_t)RR * A + (size_t)A * (BPS_TR + 1) + A) * sizeof(T) +
                        (A + BPS_TR + (n_re ? n_re + n_im : M)) * sizeof(cx<T>) + 64;
    p.tile_rows = BPS_TR;
    p.ring_rows = RR;
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    QB_CUDA_CHECK(cudaFuncSetAttribute(bps_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    QB_CUDA_CHECK(cudaFuncSetAttribute(bps_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    if (nstream > 2147483647LL) return set_error(QB_ERR_UNSUPPORTED, "bps: too many streams");
    // default: column-per-lane kernel (bps_fast.cu) where it applies, else the warp-specialised tile kernel;
    // option BPS_KERNEL (qb_set_option) = ws / simple selects the tile kernels (tests run all three)
    char force = option_char(OPT_BPS_KERNEL);
    if (comp_rows || p.windowed) force = 's';   // per-symbol angle tables, windowed sums: phase-by-phase kernel only
    if (sizeof(T) == 4 && force != 's' && force != 'w') {
        const int rc = bps_fast_dispatch(E, nstream, stream_stride, L, comp, angles, A, lev_re, n_re, lev_im, n_im,
                                         N, idx, ph, Eout, st);
        if (rc <= 0) return rc;
    }
    if (force != 's' && force != 'w' && !comp_rows && !p.windowed &&
        bps_par_wanted(nstream, L, A, idx == nullptr, (int)sizeof(T))) {
        // few long streams, complex128 or an alphabet without a grid: the phase-parallel form with this file's distance
        const int rc = bps_par_generic_launch<T>(p, nstream, st);
        if (rc <= 0) return rc;
    }
    if (force != 's') {
        int RRw = 64;
        while (RRw < 2 * WS_TR + W) RRw <<= 1;
        const size_t smem_ws = ((size_t)RRw * A + (size_t)A * (WS_TR + 1) + 32 + A) * sizeof(T) +
                               (A + (n_re ? n_re + n_im : M)) * sizeof(cx<T>) + 64;
        if (smem_ws <= 200 * 1024) {
            p.tile_rows = WS_TR;
            p.ring_rows = RRw;
            const int threads = 32 * (WS_PW + (int)((A + 31) / 32));
            bps_ws_kernel<T><<<(unsigned)nstream, threads, smem_ws, st>>>(p);
            count_launch();
            QB_CUDA_CHECK(cudaGetLastError());
            return QB_OK;
        }
    }
    if (smem > 200 * 1024) {
        // the ring of 2N rows does not fit: the phase-parallel form keeps its history in HBM and has no such limit
        if (A <= 128) {
            const int rc = bps_par_generic_launch<T>(p, nstream, st);
            if (rc <= 0) return rc;
        }
        return set_error(QB_ERR_UNSUPPORTED, "bps: 2N*A too large for the shared-memory ring");
    }
    p.tile_rows = BPS_TR;
    p.ring_rows = RR;
    bps_kernel<T><<<(unsigned)nstream, BPS_THREADS, smem, st>>>(p);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

int bps_dispatch(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                 const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                 const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                 int32_t *idx, void *ph, void *Eout, int comp_rows, cudaStream_t st)
{
    if (dtype == QB_C64)
        return launch_bps<float>(E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re,
                                 lev_im, n_im, N, idx, ph, Eout, comp_rows, st);
    return launch_bps<double>(E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re,
                              lev_im, n_im, N, idx, ph, Eout, comp_rows, st);
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
