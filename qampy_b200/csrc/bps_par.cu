// bps_par: the blind phase search of FEW LONG streams (the reference's own call: phaserec.bps on ONE capture, one
// stream per polarisation), taken apart into phases so that only what IS serial runs serially.
//
// Same function, arithmetic contract and bits as bps_fast.cu (bps + select_angle_index + select_angles,
// qampy/core/pythran_dsp.py:47-85, 26-42, 137-153, and the L2 tail qampy/core/phaserecovery.py:150-159).  In the
// fused mappings a stream is one CTA from its first row to its last, so a call on one capture lasts
// rows x (79 .. 100 cycles).  What the reference's definition really chains from row to row is little:
//
//   A  distances d[a][i] of every row to the alphabet under every test angle       -- independent: whole GPU
//   B  running sums csum[a][i] = fl(csum[a][i-1] + d[a][i]) per angle (:33/:36)    -- serial in i: one FADD per row
//   C  window differences csum[i+N] - csum[i-N] and their first arg-min (:37-41)   -- independent: whole GPU
//   D  np.unwrap of 4 x phase: a running sum of the (rare) 2 pi corrections        -- serial in i, one warp, 32 rows
//                                                                                     per step
//   E  rotation of the input by the phase, edges (phaserecovery.py:155-159)        -- independent
//
// The matrix lives in HBM in TILES of 128 rows: D[stream][tile][angle][128 rows + 16 bytes of padding] (one float, or
// one double for complex128 signals, per entry).  Phase A's lane owns an angle and stores 8 consecutive rows as 16-byte
// vectors.  Phase B's lane owns an angle and walks down its column: a tile of 128 rows x 32 angles is one contiguous
// 16.5 kB block that comes and goes as ONE bulk copy each way (cp.async.bulk + mbarrier, seven tiles on their way in,
// up to four out); LDS.128 -> four dependent FADDs -> STS.128 in place, the memory instructions placed between the
// FADDs.  Measured 8.4 cycles per row = 1077 per tile (clock64 in the kernel): 683 in the adds (5.3 per row against the
// 4.4 of the bare FADD chain), 205 in lane 0's wait_group / expect_tx / bulk load and the mbarrier wait, 124 in the
// proxy fence and the bulk store.  Tried on the way: a plain column-major matrix (40 MB from one angle to the next at
// 1e7 rows: 33 cycles per row in TLB misses), 512-byte bulk copies per column (34 cycles per row).  Phase C's thread
// owns a ROW and walks over the angles (coalesced along rows, strict < in ascending angle order = the reference's first
// minimum; 5 TB/s of HBM reads).  Phase D evaluates a batch of 512 rows at once and enters the serial fold of
// np.unwrap's corrections only for a batch that has one (2.7 cycles per row).  1 kB of HBM traffic per row, which is
// why this form is for few streams only: many streams keep the fused kernel, whose distances never leave the lane.
//
// Any number of test angles up to 128 (the matrix has whole blocks of 32 columns; the padding columns are never read),
// and no limit on 2N x A: the history lives in HBM, not in a shared-memory ring -- shapes the tile kernels of bps.cu
// reject (complex128, 100 angles, N = 70) run here.  complex128 signals and alphabets without a rectangular grid take
// the same phases with the distance of bps.cu (generic slicer / search over the alphabet) in phase A and double-
// precision sums (DADD chain: 8.8 cycles per row).
//
// One capture of two polarisations, 64 angles (scratch/bps_par_time.py): 1e7 rows 63 ms against 408 ms in the producer /
// chain mapping (12.4 against 80 cycles per row), 1e6 rows 6.5 against 40 ms, 2^17 rows 1.2 against 5.3 ms; indices and
// phases identical.
#include <stdlib.h>

#include "bps_dist.cuh"
#include "bps_generic.cuh"

namespace qb {

constexpr int PAR_NR = 8;             // rows per group of phase A (whole 16-byte stores per lane)
constexpr int PAR_RC = 1024;          // rows per CTA of phase A
constexpr int PAR_TR = 128;           // rows per tile of the matrix
constexpr int PAR_UB = 512;           // rows per batch of phase D
constexpr int PAR_UST = 8;            // batches in the ring of phase D (6 in flight)

// T = float (complex64 signals) or double (complex128: the reference's default dtype)
template <typename T>
struct ParGeom {
    static constexpr int EV = 16 / sizeof(T);              // entries per 16-byte vector
    // pitch of an angle column inside a tile, in HBM and in shared memory alike: one vector of padding, so that the
    // 128-bit shared-memory accesses of the 32 lanes fall into disjoint banks and a tile is ONE bulk copy
    static constexpr int TRP = PAR_TR + EV;
    static constexpr int ST = sizeof(T) == 4 ? 12 : 6;     // stages of phase B's ring (203 kB / 200 kB)
    static constexpr int LD = sizeof(T) == 4 ? 7 : 3;      // tiles on their way in; ST - LD - 1 may be on their way out
};
// entries of one stream's matrix: Lp / 128 tiles of A columns
template <typename T>
__host__ __device__ __forceinline__ long long par_stream_elems(int A, long long Lp)
{
    return Lp / PAR_TR * A * ParGeom<T>::TRP;
}
// entry (angle a, row i) of a stream's matrix: ((tile i / 128) * A + a) * TRP + i % 128
template <typename T>
__device__ __forceinline__ long long par_at(int A, int a, long long i)
{
    return ((i / PAR_TR) * A + a) * ParGeom<T>::TRP + i % PAR_TR;
}

// what phases B .. E need to know about a call
template <typename T>
struct ParCall {
    const cx<T> *E;
    const T *angles;
    int32_t *idx;      // the caller's index array or nullptr
    T *ph;
    cx<T> *Eout;
    long long stream_stride, L;
    int A, N;
    int Ap;            // A rounded up to whole blocks of 32 angles: the matrix's column count (columns >= A are never read)
};

// ---- A (complex64, rectangular alphabet): distances with the packed slicer of bps_fast.cu -----------------------------
template <int NW>
__global__ void __launch_bounds__(64 * NW) bps_par_dist_kernel(BpsFastParams p, float *D, long long Lp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = 64 * NW;
    float2 *tabre = reinterpret_cast<float2 *>(smem_raw);                 // [n_re]
    float2 *tabim = tabre + p.n_re;                                        // [n_im]
    uint32_t *kst = reinterpret_cast<uint32_t *>(tabim + p.n_im);          // [2]
    float4 *bounce = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(kst + 2) + 15) & ~(uintptr_t)15);   // [NT]

    for (int c = tid; c < p.n_re; c += NT)
        tabre[c] = make_float2(-p.lev_re[c], -p.lev_re[min(c + 1, p.n_re - 1)]);
    for (int c = tid; c < p.n_im; c += NT)
        tabim[c] = make_float2(-p.lev_im[c], -p.lev_im[min(c + 1, p.n_im - 1)]);
    FastAxis gre = make_fast_axis(p.lev_re, p.n_re, smem_u32(tabre));
    FastAxis gim = make_fast_axis(p.lev_im, p.n_im, smem_u32(tabim));
    // address constants and rotation constants through shared memory, as in bps_fast_kernel (same instruction selection)
    if (tid == 0) {
        kst[0] = gre.kaddr;
        kst[1] = gim.kaddr;
    }
    const int colw = warp % NW, part = warp / NW;        // angle block, half of the CTA's rows
    const float2 cc = p.comp[colw * 32 + lane];
    bounce[tid] = make_float4(cc.x, cc.y, -cc.y, cc.x);
    __syncthreads();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(gre.kaddr) : "r"(smem_u32(kst)));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(gim.kaddr) : "r"(smem_u32(kst)));
    f32x2 c1, c2;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(c1), "=l"(c2) : "r"(smem_u32(bounce + tid)) : "memory");

    const float2 *E = p.E + (long long)blockIdx.y * p.stream_stride;
    float *mat = D + (long long)blockIdx.y * par_stream_elems<float>(p.A, Lp);
    const int a = colw * 32 + lane;
    const long long r0 = (long long)blockIdx.x * PAR_RC + (long long)part * (PAR_RC / 2);
    const long long last = p.L - 1;
#pragma unroll 1
    for (int g = 0; g < PAR_RC / 2 / PAR_NR; g++) {
        const long long i = r0 + (long long)g * PAR_NR;
        if (i >= p.L) break;                              // rows past the end are never read
        float2 e[PAR_NR];
#pragma unroll
        for (int u = 0; u < PAR_NR; u++) e[u] = __ldg(E + min(i + u, last));   // same address in every lane: one sector
        float c[PAR_NR];
#pragma unroll
        for (int u = 0; u < PAR_NR; u++) c[u] = fast_dist(e[u], c1, c2, gre, gim);
        float4 *dst = reinterpret_cast<float4 *>(mat + par_at<float>(p.A, a, i));
        dst[0] = make_float4(c[0], c[1], c[2], c[3]);
        dst[1] = make_float4(c[4], c[5], c[6], c[7]);
    }
}

// ---- A (any dtype, any alphabet): distances with the generic slicer / the search over the alphabet of bps.cu -----------
template <typename T>
__global__ void __launch_bounds__(256) bps_par_dist_generic_kernel(BpsParams<T> p, T *D, long long Lp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = p.A;
    cx<T> *comp = reinterpret_cast<cx<T> *>(smem_raw);  // [A]
    cx<T> *syms = comp + A;                             // [M] (search) or level pairs (slicer)
    const bool slicer = p.n_re > 0;
    const int npr = slicer ? max(p.n_re - 1, 1) : 0, npi = slicer ? max(p.n_im - 1, 1) : 0;
    cx<T> *pre = syms, *pim = syms + npr;
    for (int c = tid; c < A; c += 256) comp[c] = p.comp[c];
    AxisGrid<T> gre, gim;
    gre.scale = gre.bias = gim.scale = gim.bias = (T)0;
    gre.npair = gim.npair = 1;
    if (slicer) {
        for (int c = tid; c < npr; c += 256) pre[c] = make_cx<T>(p.lev_re[c], p.lev_re[min(c + 1, p.n_re - 1)]);
        for (int c = tid; c < npi; c += 256) pim[c] = make_cx<T>(p.lev_im[c], p.lev_im[min(c + 1, p.n_im - 1)]);
        gre = make_grid(p.lev_re, p.n_re);
        gim = make_grid(p.lev_im, p.n_im);
    } else {
        for (int c = tid; c < p.M; c += 256) syms[c] = p.symbols[c];
    }
    __syncthreads();
    const cx<T> *E = p.E + (long long)blockIdx.y * p.stream_stride;
    const int Ap = (A + 31) / 32 * 32;
    T *mat = D + (long long)blockIdx.y * par_stream_elems<T>(Ap, Lp);
    const int nw = Ap / 32;
    const long long last = p.L - 1;
    // units of work: (group of 8 rows, block of 32 angles); lane = angle
    for (int u = warp; u < nw * (PAR_RC / PAR_NR); u += 8) {
        const int cb = u % nw, g = u / nw;
        const long long i = (long long)blockIdx.x * PAR_RC + (long long)g * PAR_NR;
        if (i >= p.L) break;
        const int a = cb * 32 + lane;
        if (a >= A) continue;                              // padding columns of the last block: never read
        const cx<T> c = comp[a];
        T d[PAR_NR];
#pragma unroll
        for (int k = 0; k < PAR_NR; k++) d[k] = min_distance<T>(E[min(i + k, last)], c, slicer, pre, pim, gre, gim, syms, p.M);
        T *dst = mat + par_at<T>(Ap, a, i);
        if constexpr (sizeof(T) == 4) {
            reinterpret_cast<float4 *>(dst)[0] = make_float4(d[0], d[1], d[2], d[3]);
            reinterpret_cast<float4 *>(dst)[1] = make_float4(d[4], d[5], d[6], d[7]);
        } else {
#pragma unroll
            for (int k = 0; k < PAR_NR / 2; k++) reinterpret_cast<double2 *>(dst)[k] = make_double2(d[2 * k], d[2 * k + 1]);
        }
    }
}

// ---- B: running sums down the columns, in place ---------------------------------------------------------------------
// one warp per (stream, block of 32 angles); lane = angle.  A tile comes and goes as ONE bulk copy each way: the warp's
// own instructions per row are the add and a share of one 128-bit shared-memory load and store.
__device__ __forceinline__ void bulk_s2g(void *gmem, const void *smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(32) bps_par_csum_kernel(T *D, long long Lp, long long L, int A)
{
    using G = ParGeom<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tiles = reinterpret_cast<T *>(smem_raw);                            // [ST][32][TRP]
    __shared__ uint64_t full[G::ST];
    const int lane = threadIdx.x;
    const int nw = A / 32;
    const int s = blockIdx.x / nw, cb = blockIdx.x % nw;
    T *base = D + (long long)s * par_stream_elems<T>(A, Lp) + (long long)cb * 32 * G::TRP;   // tile t: + t * A * TRP
    const long long tstride = (long long)A * G::TRP;
    const long long ntiles = (L + PAR_TR - 1) / PAR_TR;
    constexpr uint32_t TILE_BYTES = 32 * G::TRP * sizeof(T);               // the 32 columns of this block: contiguous
    if (lane == 0) {
        for (int k = 0; k < G::ST; k++) mbar_init(&full[k], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto load = [&](long long t) {          // lane 0
        if (t < ntiles) {
            const int st = (int)(t % G::ST);
            mbar_arrive_expect_tx(&full[st], TILE_BYTES);
            tma_bulk_g2s(tiles + (size_t)st * 32 * G::TRP, base + t * tstride, TILE_BYTES, &full[st]);
        }
    };
    if (lane == 0)
        for (int t = 0; t < G::LD; t++) load(t);
    T csum = 0;
    for (long long t = 0; t < ntiles; t++) {
        if (lane == 0) {
            // tile t + LD lands in the stage of tile t + LD - ST, whose store must have read it: the stores of the
            // ST - LD - 1 tiles after that one may still be pending
            bulk_wait_read<G::ST - G::LD - 1>();
            load(t + G::LD);
        }
        mbar_wait(&full[t % G::ST], (uint32_t)((t / G::ST) & 1));
        T *stage = tiles + (size_t)(t % G::ST) * 32 * G::TRP;
        if constexpr (sizeof(T) == 4) {
            const uint32_t my = smem_u32(stage + lane * G::TRP);
            // Software pipeline in groups of 4 rows: the LDS.128 of group g + 4 and the STS.128 of group g - 1 sit
            // between the four dependent FADDs of group g (volatile, so that they stay where they are written)
            auto lds = [&](int g) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(my + 16u * g));
                return v;
            };
            auto sts = [&](int g, const float4 &v) {
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my + 16u * g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            };
            constexpr int NG = PAR_TR / 4, PF = 4;                         // groups per tile, groups loaded ahead
            float4 q[PF], done = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < PF; g++) q[g] = lds(g);
            if (t == 0) q[0].x = 0.f;                                      // row 0 is never added (pythran_dsp.py:30)
#pragma unroll
            for (int g = 0; g < NG; g++) {
                float4 v = q[g % PF];
                csum = __fadd_rn(csum, v.x);                               // :33/:36, sequential
                v.x = csum;
                if (g + PF < NG) q[g % PF] = lds(g + PF);
                csum = __fadd_rn(csum, v.y);
                v.y = csum;
                if (g > 0) sts(g - 1, done);
                csum = __fadd_rn(csum, v.z);
                v.z = csum;
                csum = __fadd_rn(csum, v.w);
                v.w = csum;
                done = v;
            }
            sts(NG - 1, done);
        } else {
            double2 *my = reinterpret_cast<double2 *>(stage + lane * G::TRP);
#pragma unroll 8
            for (int g = 0; g < PAR_TR / 2; g++) {
                double2 v = my[g];
                if (t == 0 && g == 0) v.x = 0.;                            // row 0 is never added (pythran_dsp.py:30)
                csum = __dadd_rn(csum, v.x);                               // :33/:36, sequential
                v.x = csum;
                csum = __dadd_rn(csum, v.y);
                v.y = csum;
                my[g] = v;
            }
        }
        fence_proxy_async();                                               // the sums above, seen by the copy engine
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(base + t * tstride, stage, TILE_BYTES);
            bulk_commit();
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    __syncwarp();
}

// ---- C: window differences and their first arg-min; thread = output row ----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bps_par_argmin_kernel(const T *C, long long Lp, long long L, int A, int Ap, int N,
                                                             int32_t *idx)
{
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const long long j = lo + (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= hi) return;
    const T *cs = C + (long long)blockIdx.y * par_stream_elems<T>(Ap, Lp);
    const T *up = cs + par_at<T>(Ap, 0, j + N), *dn = cs + par_at<T>(Ap, 0, j - N);   // csum[i], csum[i - 2N], i = j + N
    T best = (T)1000.;                                       // dmin0 = 1000 (:31): nothing below it -> idx stays 0
    int bk = 0;
#pragma unroll 8
    for (int a = 0; a < A; a++) {
        const T v = sub_rn(__ldg(up + a * ParGeom<T>::TRP), __ldg(dn + a * ParGeom<T>::TRP));
        if (v < best) {                                      // strict <, ascending angles: first minimum (:39)
            best = v;
            bk = a;
        }
    }
    idx[(long long)blockIdx.y * L + j] = bk;
}

// ---- D: phases and their unwrap; one warp per stream, lane = row -------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) bps_par_unwrap_kernel(const int32_t *idx, const T *angles, int A, long long L,
                                                            int N, T *ph)
{
    __shared__ T angs[128];
    __shared__ int32_t ring[PAR_UST][PAR_UB];
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const int32_t *ix = idx + (long long)blockIdx.x * L;
    T *pho = ph + (long long)blockIdx.x * L;
    for (int c = lane; c < A; c += 32) angs[c] = angles ? angles[c] : (T)0;
    const long long nb = (hi - lo + PAR_UB - 1) / PAR_UB;
    auto load = [&](long long b) {
        if (b < nb) {
#pragma unroll
            for (int k = 0; k < PAR_UB / 32; k++) {
                const long long j = lo + b * PAR_UB + 32 * k + lane;
                if (j < hi) cp_async<4>(&ring[b % PAR_UST][32 * k + lane], ix + j);
            }
        }
        cp_async_commit();
    };
    for (int b = 0; b < PAR_UST - 2; b++) load(b);
    __syncwarp();
    T cum = 0, p4prev = 0;
    constexpr int KB = PAR_UB / 32;
    for (long long b = 0; b < nb; b++) {
        load(b + PAR_UST - 2);
        cp_async_wait<PAR_UST - 2>();
        __syncwarp();
        // everything that does not depend on the running correction, for the 16 steps of the batch at once
        T p4[KB], pp[KB];
        bool valid[KB], cand[KB];
        unsigned anyc = 0u;
#pragma unroll
        for (int k = 0; k < KB; k++) {
            const long long j = lo + b * PAR_UB + 32 * k + lane;
            valid[k] = j < hi;
            const int bk = valid[k] ? ring[b % PAR_UST][32 * k + lane] : 0;
            p4[k] = mul_rn(angs[bk], (T)4);
        }
#pragma unroll
        for (int k = 0; k < KB; k++) {
            const long long j = lo + b * PAR_UB + 32 * k + lane;
            pp[k] = shfl_up(p4[k], 1);
            const T carry = k == 0 ? p4prev : shfl_idx(p4[k > 0 ? k - 1 : 0], 31);   // a step before the last is full
            if (lane == 0) pp[k] = carry;
            // np.unwrap only acts where |dd| >= pi
            cand[k] = valid[k] && j > lo && !(fabs(sub_rn(p4[k], pp[k])) < Pi<T>::pi());
            anyc |= __ballot_sync(FULL, cand[k]);
        }
        if (anyc == 0u) {                                   // the usual batch: no correction, the running sum stands
#pragma unroll
            for (int k = 0; k < KB; k++)
                if (valid[k]) pho[lo + b * PAR_UB + 32 * k + lane] = add_rn(p4[k], cum) / (T)4;
        } else {
#pragma unroll
            for (int k = 0; k < KB; k++) {
                const T corr = cand[k] ? unwrap_corr<T>(p4[k], pp[k]) : (T)0;
                unsigned mask = __ballot_sync(FULL, corr != (T)0);
                T mycum = cum;
                while (mask) {   // fold the (rare) non-zero corrections in row order: exact sequential sum
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const T ce = shfl_idx(corr, e);
                    cum = add_rn(cum, ce);
                    if (lane >= e) mycum = cum;
                }
                if (valid[k]) pho[lo + b * PAR_UB + 32 * k + lane] = add_rn(p4[k], mycum) / (T)4;
            }
        }
        // the last valid row of the batch (only the capture's last batch is not full)
        if (b + 1 < nb) {
            p4prev = shfl_idx(p4[KB - 1], 31);
        } else {
#pragma unroll
            for (int k = 0; k < KB; k++) {
                const unsigned vm = __ballot_sync(FULL, valid[k]);
                if (vm) p4prev = shfl_idx(p4[k], 31 - __clz(vm));
            }
        }
        __syncwarp();
    }
}

// ---- E: edges and rotation --------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 par_rotate(float2 e, float ph) { return rotate_f(e, ph); }          // as bps_fast.cu
__device__ __forceinline__ double2 par_rotate(double2 e, double ph) { return rotate<double>(e, ph); }  // as bps.cu

template <typename T>
__global__ void __launch_bounds__(256) bps_par_rotate_kernel(ParCall<T> p)
{
    const long long L = p.L;
    const int N = p.N;
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= L) return;
    const cx<T> *E = p.E + (long long)blockIdx.y * p.stream_stride;
    const long long o = (long long)blockIdx.y * L + j;
    if (j < lo || j >= hi) {
        // edges: idx = 0 -> ph = angles[0], not unwrapped (phaserecovery.py:155 touches [N:-N] only)
        const T a0 = p.angles ? p.angles[0] : (T)0;
        if (p.idx) p.idx[o] = 0;
        if (p.ph) p.ph[o] = a0;
        if (p.Eout) p.Eout[o] = par_rotate(E[j], a0);
    } else if (p.Eout && p.ph) {
        p.Eout[o] = par_rotate(E[j], p.ph[o]);
    }
}

static size_t par_dist_smem(int NW, int n_re, int n_im) { return (size_t)(n_re + n_im) * 8 + 8 + 16 + (size_t)64 * NW * 16; }

static long long par_padded_rows(long long L) { return (L + PAR_RC - 1) / PAR_RC * PAR_RC; }

// bytes of scratch the phase-parallel form needs (the distance / running-sum matrix and, unless the caller wants the
// indices, an index array); elem = 4 (complex64 signals) or 8
size_t bps_par_scratch_bytes(int64_t nstream, int64_t L, int64_t A, bool own_idx, int elem)
{
    const long long Lp = par_padded_rows(L);
    const int Ap = (int)((A + 31) / 32 * 32);
    const size_t mat = elem == 4 ? (size_t)par_stream_elems<float>(Ap, Lp) * 4 : (size_t)par_stream_elems<double>(Ap, Lp) * 8;
    return (size_t)nstream * mat + (own_idx ? (size_t)nstream * L * sizeof(int32_t) : 0);
}

// phases B .. E behind a phase A that `launch_a(D, Lp)` enqueues
template <typename T, typename LaunchA>
static int par_run(const ParCall<T> &c, int64_t nstream, cudaStream_t st, LaunchA launch_a)
{
    using G = ParGeom<T>;
    const long long L = c.L, Lp = par_padded_rows(L);
    const size_t mat = (size_t)nstream * par_stream_elems<T>(c.Ap, Lp) * sizeof(T);
    T *D = nullptr;
    int32_t *own_idx = nullptr;
    {
        // keep the scratch cached in the device's pool between calls (as the host entry points do): re-allocating 5 GB
        // per call costs more than the kernels
        int dev = 0;
        cudaMemPool_t pool;
        uint64_t thr = ~0ull;
        QB_CUDA_CHECK(cudaGetDevice(&dev));
        QB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        QB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    QB_CUDA_CHECK(cudaMallocAsync(&D, mat, st));
    int32_t *idx = c.idx;
    if (!idx) {
        cudaError_t e = cudaMallocAsync(&own_idx, (size_t)nstream * L * sizeof(int32_t), st);
        if (e != cudaSuccess) {
            cudaFreeAsync(D, st);
            return set_error(QB_ERR_CUDA, "bps: cudaMallocAsync failed: %s", cudaGetErrorString(e));
        }
        idx = own_idx;
    }
    int rc = QB_OK;
    do {
        const long long lo = c.N < L ? c.N : L, hi = (L - c.N > lo) ? L - c.N : lo;
        if (hi > lo) {
            launch_a(D, Lp);
            count_launch();
            const size_t smem_b = (size_t)G::ST * 32 * G::TRP * sizeof(T);   // + the static mbarriers
            if (cudaFuncSetAttribute(bps_par_csum_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) != cudaSuccess) {
                rc = set_error(QB_ERR_CUDA, "bps: cudaFuncSetAttribute failed");
                break;
            }
            bps_par_csum_kernel<T><<<(unsigned)(nstream * (c.Ap / 32)), 32, smem_b, st>>>(D, Lp, L, c.Ap);
            count_launch();
            bps_par_argmin_kernel<T><<<dim3((unsigned)((hi - lo + 255) / 256), (unsigned)nstream), 256, 0, st>>>(D, Lp, L, c.A, c.Ap, c.N, idx);
            count_launch();
            if (c.ph) {
                bps_par_unwrap_kernel<T><<<(unsigned)nstream, 32, 0, st>>>(idx, c.angles, c.A, L, c.N, c.ph);
                count_launch();
            }
        }
        if (L > 0) {
            bps_par_rotate_kernel<T><<<dim3((unsigned)((L + 255) / 256), (unsigned)nstream), 256, 0, st>>>(c);
            count_launch();
        }
        if (cudaGetLastError() != cudaSuccess) rc = set_error(QB_ERR_CUDA, "bps: launch of the phase-parallel kernels failed");
    } while (0);
    cudaFreeAsync(D, st);
    if (own_idx) cudaFreeAsync(own_idx, st);
    return rc;
}

// Is the phase-parallel form the one to take?  Few LONG streams (one capture): only the running sum and the unwrap stay
// serial.  It moves 1 kB (complex64) of HBM per row through a scratch matrix, so it needs room and few streams.  Option
// BPS_SPLIT = 2 (qb_set_option) forces it, 0 / 1 exclude it (tests run all mappings).
bool bps_par_wanted(int64_t nstream, int64_t L, int64_t A, bool own_idx, int elem)
{
    if (A < 1 || A > 128 || nstream < 1) return false;
    bool par = L >= 32768 && nstream * ((A + 31) / 32) <= 64;
    if (const char e = option_char(OPT_BPS_SPLIT)) par = e == '2' && L >= 1;
    if (par) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bps_par_scratch_bytes(nstream, L, A, own_idx, elem) > free_b / 3)
            par = false;
    }
    return par;
}

template <int NW>
static int launch_par_fast(const BpsFastParams &p, int64_t nstream, cudaStream_t st)
{
    ParCall<float> c{p.E, p.angles, p.idx, p.ph, p.Eout, p.stream_stride, p.L, p.A, p.N, p.A};   // A % 32 == 0 here
    return par_run<float>(c, nstream, st, [&](float *D, long long Lp) {
        bps_par_dist_kernel<NW><<<dim3((unsigned)(Lp / PAR_RC), (unsigned)nstream), 64 * NW, par_dist_smem(NW, p.n_re, p.n_im), st>>>(p, D, Lp);
    });
}

// complex64 signal on a rectangular alphabet (called from bps_fast.cu); same contract as launch_fast there
int bps_par_launch(const BpsFastParams &p, int64_t nstream, cudaStream_t st)
{
    switch (p.A / 32) {
    case 1: return launch_par_fast<1>(p, nstream, st);
    case 2: return launch_par_fast<2>(p, nstream, st);
    case 3: return launch_par_fast<3>(p, nstream, st);
    default: return launch_par_fast<4>(p, nstream, st);
    }
}

// any dtype, any alphabet (called from bps.cu); returns 1 if the shape is not covered (caller keeps its tile kernels)
template <typename T>
int bps_par_generic_launch(const BpsParams<T> &p, int64_t nstream, cudaStream_t st)
{
    const bool slicer = p.n_re > 0;
    const size_t smem_a = (size_t)(p.A + (slicer ? p.n_re + p.n_im : p.M)) * sizeof(cx<T>);
    if (p.comp_rows || p.windowed || smem_a > 48 * 1024) return 1;
    ParCall<T> c{p.E, p.angles, p.idx, p.ph, p.Eout, p.stream_stride, p.L, p.A, p.N, (p.A + 31) / 32 * 32};
    return par_run<T>(c, nstream, st, [&](T *D, long long Lp) {
        bps_par_dist_generic_kernel<T><<<dim3((unsigned)(Lp / PAR_RC), (unsigned)nstream), 256, smem_a, st>>>(p, D, Lp);
    });
}
template int bps_par_generic_launch<float>(const BpsParams<float> &, int64_t, cudaStream_t);
template int bps_par_generic_launch<double>(const BpsParams<double> &, int64_t, cudaStream_t);

}  // namespace qb
