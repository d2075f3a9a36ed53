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
// The matrix lives in HBM in TILES of 128 rows: D[stream][tile][angle][128 rows + 4 of padding] (one float per entry).
// Phase A's lane owns an angle and stores 8 consecutive rows as two 16-byte vectors.  Phase B's lane owns an angle and
// walks down its column: a tile of 128 rows x 32 angles is one contiguous 16.5 kB block that comes and goes as ONE bulk
// copy each way (cp.async.bulk + mbarrier, seven tiles on their way in, up to four out); LDS.128 -> four dependent
// FADDs -> STS.128 in place, the memory instructions placed between the FADDs.  Measured 8.6 cycles per row (the FADD
// chain alone is 4.4; a lone warp's issue cadence with the 128-bit shared-memory accesses makes up the rest; without
// the store 7.6).  What was tried on the way: a plain column-major matrix, 40 MB from one angle to the next at 1e7
// rows: 33 cycles per row in TLB misses; 512-byte bulk copies per column: 34 cycles per row.  Phase C's thread owns a
// ROW and walks over the angles (coalesced along rows, strict < in ascending angle order = the reference's first
// minimum).  Phase D evaluates a batch of 512 rows at once and enters the serial fold of np.unwrap's corrections only
// for a batch that has one (2.7 cycles per row).  1 kB of HBM traffic per row, which is why this form is for few
// streams only: many streams keep the fused kernel, whose distances never leave the lane.
//
// One capture of two polarisations, 64 angles (scratch/bps_par_time.py): 1e7 rows 63 ms against 408 ms in the producer /
// chain mapping (12.4 against 80 cycles per row), 1e6 rows 6.5 against 40 ms, 2^17 rows 1.2 against 5.3 ms; indices and
// phases identical.
#include <stdlib.h>

#include "bps_dist.cuh"

namespace qb {

constexpr int PAR_NR = 8;             // rows per group of phase A (two 16-byte stores per lane)
constexpr int PAR_RC = 1024;          // rows per CTA of phase A
constexpr int PAR_TR = 128;           // rows per staged tile of phase B
constexpr int PAR_TRP = PAR_TR + 4;   // pitch of an angle column inside a tile, in HBM and in shared memory alike: the
                                      // LDS.128 of the 32 lanes fall into disjoint banks, and a tile is ONE bulk copy
constexpr int PAR_ST = 12;            // stages of its ring
constexpr int PAR_LD = 7;             // tiles on their way in; the other PAR_ST - PAR_LD - 1 stages may still be on their way out
constexpr int PAR_UB = 512;           // rows per batch of phase D
constexpr int PAR_UST = 8;            // batches in the ring of phase D (6 in flight)

// floats of one stream's matrix: Lp / 128 tiles of A columns of 132
__host__ __device__ __forceinline__ long long par_stream_floats(int A, long long Lp) { return Lp / PAR_TR * A * PAR_TRP; }

// ---- A: distances ---------------------------------------------------------------------------------------------------
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
    // entry (angle a, row i) of this stream: ((tile i / 128) * A + a) * 132 + i % 128
    float *col = D + (long long)blockIdx.y * par_stream_floats(p.A, Lp) + (long long)(colw * 32 + lane) * PAR_TRP;
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
        float4 *dst = reinterpret_cast<float4 *>(col + (i / PAR_TR) * ((long long)p.A * PAR_TRP) + (i % PAR_TR));
        dst[0] = make_float4(c[0], c[1], c[2], c[3]);
        dst[1] = make_float4(c[4], c[5], c[6], c[7]);
    }
}

// ---- B: running sums down the columns, in place ---------------------------------------------------------------------
// one warp per (stream, block of 32 angles); lane = angle.  Tiles come and go as bulk copies (one 512-byte column piece
// per lane and tile each way): the warp's own instructions per row are 1/4 LDS.128 + FADD + 1/4 STS.128.
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

__global__ void __launch_bounds__(32) bps_par_csum_kernel(float *D, long long Lp, long long L, int A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                    // [PAR_ST][32][PAR_TRP]
    __shared__ uint64_t full[PAR_ST];
    const int lane = threadIdx.x;
    const int nw = A / 32;
    const int s = blockIdx.x / nw, cb = blockIdx.x % nw;
    float *base = D + (long long)s * par_stream_floats(A, Lp) + (long long)cb * 32 * PAR_TRP;   // tile t: + t * A * 132
    const long long tstride = (long long)A * PAR_TRP;
    const long long ntiles = (L + PAR_TR - 1) / PAR_TR;
    constexpr uint32_t TILE_BYTES = 32 * PAR_TRP * sizeof(float);          // 32 columns of this block: contiguous
    if (lane == 0) {
        for (int k = 0; k < PAR_ST; k++) mbar_init(&full[k], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto load = [&](long long t) {          // lane 0
        if (t < ntiles) {
            const int st = (int)(t % PAR_ST);
            mbar_arrive_expect_tx(&full[st], TILE_BYTES);
            tma_bulk_g2s(tiles + (size_t)st * 32 * PAR_TRP, base + t * tstride, TILE_BYTES, &full[st]);
        }
    };
    if (lane == 0)
        for (int t = 0; t < PAR_LD; t++) load(t);
    float csum = 0.f;
    for (long long t = 0; t < ntiles; t++) {
        if (lane == 0) {
            // tile t + PAR_LD lands in the stage of tile t + PAR_LD - PAR_ST, whose store must have read it: the stores of
            // the PAR_ST - PAR_LD - 1 tiles after that one may still be pending
            bulk_wait_read<PAR_ST - PAR_LD - 1>();
            load(t + PAR_LD);
        }
        mbar_wait(&full[t % PAR_ST], (uint32_t)((t / PAR_ST) & 1));
        float *stage = tiles + (size_t)(t % PAR_ST) * 32 * PAR_TRP;
        const uint32_t my = smem_u32(stage + lane * PAR_TRP);
        // Software pipeline in groups of 4 rows: the LDS.128 of group g + 2 and the STS.128 of group g - 1 sit between
        // the four dependent FADDs of group g (volatile, so that they stay where they are written): a lone warp
        // issues one instruction per ~2.5 cycles, the FADD chain needs one per 4.4 -- the memory instructions ride in
        // the chain's shadow instead of queueing before and after it.
        auto lds = [&](int g) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(my + 16u * g));
            return v;
        };
        auto sts = [&](int g, const float4 &v) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my + 16u * g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        constexpr int NG = PAR_TR / 4;
        float4 v0 = lds(0), v1 = lds(1), done = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t == 0) v0.x = 0.f;                                            // row 0 is never added (pythran_dsp.py:30)
#pragma unroll
        for (int g = 0; g < NG; g++) {
            float4 v2 = make_float4(0.f, 0.f, 0.f, 0.f);
            csum = __fadd_rn(csum, v0.x);                                  // :33/:36, sequential
            v0.x = csum;
            if (g + 2 < NG) v2 = lds(g + 2);
            csum = __fadd_rn(csum, v0.y);
            v0.y = csum;
            if (g > 0) sts(g - 1, done);
            csum = __fadd_rn(csum, v0.z);
            v0.z = csum;
            csum = __fadd_rn(csum, v0.w);
            v0.w = csum;
            done = v0;
            v0 = v1;
            v1 = v2;
        }
        sts(NG - 1, done);
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
__global__ void __launch_bounds__(256) bps_par_argmin_kernel(const float *C, long long Lp, long long L, int A, int N,
                                                             int32_t *idx)
{
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const long long j = lo + (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= hi) return;
    const float *cs = C + (long long)blockIdx.y * par_stream_floats(A, Lp);
    const long long iu = j + N, id = j - N;                  // csum[i], csum[i - 2N] with i = j + N
    const float *up = cs + (iu / PAR_TR) * ((long long)A * PAR_TRP) + iu % PAR_TR;
    const float *dn = cs + (id / PAR_TR) * ((long long)A * PAR_TRP) + id % PAR_TR;
    unsigned best = 0xffffffffu;
    int bk = 0;
#pragma unroll 8
    for (int a = 0; a < A; a++) {
        const unsigned db = __float_as_uint(__fsub_rn(__ldg(up + a * PAR_TRP), __ldg(dn + a * PAR_TRP)));
        if (db < best) {                                     // >= +0: orders like an unsigned; strict <: first minimum (:39)
            best = db;
            bk = a;
        }
    }
    if (best >= 0x447a0000u) bk = 0;                         // dmin0 = 1000 (:31): nothing below it -> idx stays 0
    idx[(long long)blockIdx.y * L + j] = bk;
}

// ---- D: phases and their unwrap; one warp per stream, lane = row -------------------------------------------------------
__global__ void __launch_bounds__(32) bps_par_unwrap_kernel(const int32_t *idx, const float *angles, int A, long long L,
                                                            int N, float *ph)
{
    __shared__ float angs[128];
    __shared__ int32_t ring[PAR_UST][PAR_UB];
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const int32_t *ix = idx + (long long)blockIdx.x * L;
    float *pho = ph + (long long)blockIdx.x * L;
    for (int c = lane; c < A; c += 32) angs[c] = angles ? angles[c] : 0.f;
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
    float cum = 0.f, p4prev = 0.f;
    constexpr int KB = PAR_UB / 32;
    for (long long b = 0; b < nb; b++) {
        load(b + PAR_UST - 2);
        cp_async_wait<PAR_UST - 2>();
        __syncwarp();
        // everything that does not depend on the running correction, for the 8 steps of the batch at once
        float p4[KB], pp[KB];
        bool valid[KB], cand[KB];
        unsigned anyc = 0u;
#pragma unroll
        for (int k = 0; k < KB; k++) {
            const long long j = lo + b * PAR_UB + 32 * k + lane;
            valid[k] = j < hi;
            const int bk = valid[k] ? ring[b % PAR_UST][32 * k + lane] : 0;
            p4[k] = __fmul_rn(angs[bk], 4.f);
        }
#pragma unroll
        for (int k = 0; k < KB; k++) {
            const long long j = lo + b * PAR_UB + 32 * k + lane;
            pp[k] = __shfl_up_sync(FULL, p4[k], 1);
            const float carry = k == 0 ? p4prev : __shfl_sync(FULL, p4[k > 0 ? k - 1 : 0], 31);   // a step before the last is full
            if (lane == 0) pp[k] = carry;
            // np.unwrap only acts where |dd| >= pi
            cand[k] = valid[k] && j > lo && !(fabsf(__fsub_rn(p4[k], pp[k])) < 3.14159274101257324219f);
            anyc |= __ballot_sync(FULL, cand[k]);
        }
        if (anyc == 0u) {                                   // the usual batch: no correction, the running sum stands
#pragma unroll
            for (int k = 0; k < KB; k++)
                if (valid[k]) pho[lo + b * PAR_UB + 32 * k + lane] = __fadd_rn(p4[k], cum) / 4.f;
        } else {
#pragma unroll
            for (int k = 0; k < KB; k++) {
                const float corr = cand[k] ? unwrap_corr_f(p4[k], pp[k]) : 0.f;
                unsigned mask = __ballot_sync(FULL, corr != 0.f);
                float mycum = cum;
                while (mask) {   // fold the (rare) non-zero corrections in row order: exact sequential sum
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float ce = __shfl_sync(FULL, corr, e);
                    cum = __fadd_rn(cum, ce);
                    if (lane >= e) mycum = cum;
                }
                if (valid[k]) pho[lo + b * PAR_UB + 32 * k + lane] = __fadd_rn(p4[k], mycum) / 4.f;
            }
        }
        // the last valid row of the batch (only the capture's last batch is not full)
        if (b + 1 < nb) {
            p4prev = __shfl_sync(FULL, p4[KB - 1], 31);
        } else {
#pragma unroll
            for (int k = 0; k < KB; k++) {
                const unsigned vm = __ballot_sync(FULL, valid[k]);
                if (vm) p4prev = __shfl_sync(FULL, p4[k], 31 - __clz(vm));
            }
        }
        __syncwarp();
    }
}

// ---- E: edges and rotation --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bps_par_rotate_kernel(BpsFastParams p)
{
    const long long L = p.L;
    const int N = p.N;
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= L) return;
    const float2 *E = p.E + (long long)blockIdx.y * p.stream_stride;
    const long long o = (long long)blockIdx.y * L + j;
    if (j < lo || j >= hi) {
        // edges: idx = 0 -> ph = angles[0], not unwrapped (phaserecovery.py:155 touches [N:-N] only)
        const float a0 = p.angles ? p.angles[0] : 0.f;
        if (p.idx) p.idx[o] = 0;
        if (p.ph) p.ph[o] = a0;
        if (p.Eout) p.Eout[o] = rotate_f(E[j], a0);
    } else if (p.Eout && p.ph) {
        p.Eout[o] = rotate_f(E[j], p.ph[o]);
    }
}

static size_t par_dist_smem(int NW, int n_re, int n_im) { return (size_t)(n_re + n_im) * 8 + 8 + 16 + (size_t)64 * NW * 16; }

// bytes of scratch the phase-parallel form needs (the distance / running-sum matrix and, unless the caller wants the
// indices, an index array)
size_t bps_par_scratch_bytes(int64_t nstream, int64_t L, int64_t A, bool own_idx)
{
    const long long Lp = (L + PAR_RC - 1) / PAR_RC * PAR_RC;
    return (size_t)nstream * par_stream_floats((int)A, Lp) * sizeof(float) + (own_idx ? (size_t)nstream * L * sizeof(int32_t) : 0);
}

template <int NW>
static int launch_par(const BpsFastParams &p, int64_t nstream, cudaStream_t st)
{
    const long long L = p.L, Lp = (L + PAR_RC - 1) / PAR_RC * PAR_RC;
    const size_t mat = (size_t)nstream * par_stream_floats(p.A, Lp) * sizeof(float);
    float *D = nullptr;
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
    int32_t *idx = p.idx;
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
        const long long lo = p.N < L ? p.N : L, hi = (L - p.N > lo) ? L - p.N : lo;
        if (hi > lo) {
            bps_par_dist_kernel<NW><<<dim3((unsigned)(Lp / PAR_RC), (unsigned)nstream), 64 * NW, par_dist_smem(NW, p.n_re, p.n_im), st>>>(p, D, Lp);
            count_launch();
            const size_t smem_b = (size_t)PAR_ST * 32 * PAR_TRP * sizeof(float);   // + the static mbarriers
            if (cudaFuncSetAttribute(bps_par_csum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) != cudaSuccess) {
                rc = set_error(QB_ERR_CUDA, "bps: cudaFuncSetAttribute failed");
                break;
            }
            bps_par_csum_kernel<<<(unsigned)(nstream * NW), 32, smem_b, st>>>(D, Lp, L, p.A);
            count_launch();
            bps_par_argmin_kernel<<<dim3((unsigned)((hi - lo + 255) / 256), (unsigned)nstream), 256, 0, st>>>(D, Lp, L, p.A, p.N, idx);
            count_launch();
            if (p.ph) {
                bps_par_unwrap_kernel<<<(unsigned)nstream, 32, 0, st>>>(idx, p.angles, p.A, L, p.N, p.ph);
                count_launch();
            }
        }
        if (L > 0) {
            bps_par_rotate_kernel<<<dim3((unsigned)((L + 255) / 256), (unsigned)nstream), 256, 0, st>>>(p);
            count_launch();
        }
        if (cudaGetLastError() != cudaSuccess) rc = set_error(QB_ERR_CUDA, "bps: launch of the phase-parallel kernels failed");
    } while (0);
    cudaFreeAsync(D, st);
    if (own_idx) cudaFreeAsync(own_idx, st);
    return rc;
}

// phase-parallel form; same contract as launch_fast in bps_fast.cu
int bps_par_launch(const BpsFastParams &p, int64_t nstream, cudaStream_t st)
{
    switch (p.A / 32) {
    case 1: return launch_par<1>(p, nstream, st);
    case 2: return launch_par<2>(p, nstream, st);
    case 3: return launch_par<3>(p, nstream, st);
    default: return launch_par<4>(p, nstream, st);
    }
}

}  // namespace qb
