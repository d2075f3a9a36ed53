// bps_fast: column-per-lane blind phase search (complex64, rectangular alphabet, A in {32,64,96,128}).
//
// Same function as bps_kernel in bps.cu (bps + select_angle_index + select_angles,
// qampy/core/pythran_dsp.py:47-85, 26-42, 137-153, and the L2 tail qampy/core/phaserecovery.py:150-159)
// and the same arithmetic contract: unfused complex multiply, d = fl(fl(dr^2)+fl(di^2)),
// clamp at 100, running column sum added in the reference's order, window difference, first strict
// arg-min with dmin0 = 1000.  What changes is the mapping, chosen to minimise issued instructions,
// which is what bounds this kernel (1.28e9 distance evaluations at C3, ~0.4 kB of HBM traffic each 1000):
//
//   * one lane owns one test-angle COLUMN of one stream for the whole stream: its rotation constants,
//     its running sum and its place in the ring live in registers; a warp owns 32 columns, a CTA of
//     A/32 warps owns one stream.  The distance, the running sum and the window difference of a
//     (row, angle) pair never leave the lane -- no distance matrix in shared memory, no phase barriers.
//   * the history csum[i-2N] is a ring of EXACTLY 2N rows per warp (one LDS + one STS at the same
//     address per row), 11.5 kB per warp at N = 45, so 16 warps are resident per SM.
//   * the arg-min over the angles is ONE warp reduction per row: window differences are non-negative
//     floats, so their bit patterns order like unsigned integers and CREDUX.MIN (redux.sync.min.u32)
//     gives the minimum; the first lane holding it comes from a ballot.  Rows are recorded as
//     (min bits, ballot) and turned into an index 32 rows at a time (lane = row) in the tail.
//   * packed fp32 (FMUL2 / FADD2, Blackwell) halves the issue slots of the rotation, the two level
//     differences per axis and the squares; the slicer coordinate is one FFMA.SAT (clamp for free) plus
//     one FFMA that leaves the bracket index in the low mantissa bits.  ptxas contracts
//     mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry .rn, so every addition that
//     consumes a packed product is a scalar add.rn (checked in SASS: the kernel contains no FFMA2).
//   * the tail (arg-min combine across warps, unwrap, phase output, rotation) runs once per 32 rows
//     with lane = row, on the warps of the CTA in turn so that none of them falls behind.
//
// Few streams (the reference's own call: bps on ONE capture, one stream per polarisation): the time of the call is the
// serial depth of a stream, 100 cycles per row in the fused mapping above because the lane that owns a column also
// evaluates its distances.  bps_fast_kernel<NW, NPG > 0> splits the two: NPG * NW PRODUCER warps evaluate the distances
// of a 32-row tile (lane = angle column, warp = a quarter of the rows) into a double-buffered tile in shared memory,
// the NW CHAIN warps only add, difference and reduce (one load replaces the 19-instruction evaluation); named barriers
// hand the tiles over.  Same operations on the same operands in the same order per (row, angle), so indices and phases
// are bit-identical to the fused mapping, which lets the launcher pick by the number of streams.
#include <stdlib.h>

#include "bps_dist.cuh"

namespace qb {

template <int ID>
__device__ __forceinline__ void fbar_sync(int count)
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void fbar_arrive(int count)
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}

int bps_par_launch(const BpsFastParams &p, int64_t nstream, cudaStream_t st);            // bps_par.cu
bool bps_par_wanted(int64_t nstream, int64_t L, int64_t A, bool own_idx, int elem);

constexpr int FAST_TR = 32;   // rows per tile (= lanes of the tail)
#ifndef QB_BPS_NR
#define QB_BPS_NR 8
#endif
constexpr int FAST_NR = QB_BPS_NR;   // rows per group (independent distance evaluations in flight per lane)

// NPG: producer row groups (0: fused mapping, every chain lane evaluates its own distances)
template <int NW, int NPG = 0>
__global__ void __launch_bounds__(32 * NW * (1 + NPG)) bps_fast_kernel(BpsFastParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = 32 * NW * (1 + NPG);       // all threads of the CTA
    constexpr int NC = 32 * NW;                   // chain threads
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_CHAIN = 5;   // named barriers (0 is __syncthreads)
    const int N = p.N, W = 2 * p.N;
    const long long L = p.L;
    const unsigned FULL = 0xffffffffu;

    // ---- shared memory ---------------------------------------------------------------------------
    float2 *tabre = reinterpret_cast<float2 *>(smem_raw);                 // [n_re]
    float2 *tabim = tabre + p.n_re;                                        // [n_im]
    float2 *stage = tabim + p.n_im;                                        // [NW][32] input rows of the tile
    uint2 *part = reinterpret_cast<uint2 *>(stage + NW * FAST_TR);         // [2][NW][32] hand-over to the tail
    float *angs = reinterpret_cast<float *>(part + 2 * NW * FAST_TR);      // [A]
    float *ust = angs + p.A;                                               // [2] unwrap state (cum, p4prev)
    float *ring = ust + 2;                                                 // [NW][W][32] running sums
    // NPG > 0: [2][32 rows][A] distances, 16-byte aligned (also the bounce buffer of the column constants below)
    float *dbuf = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ring + (size_t)NW * W * 32) + 15) & ~(uintptr_t)15);

    const float2 *E = p.E + (long long)blockIdx.x * p.stream_stride;
    int32_t *idx = p.idx ? p.idx + (long long)blockIdx.x * L : nullptr;
    float *ph = p.ph ? p.ph + (long long)blockIdx.x * L : nullptr;
    float2 *Eout = p.Eout ? p.Eout + (long long)blockIdx.x * L : nullptr;

    for (int c = tid; c < p.n_re; c += NT)
        tabre[c] = make_float2(-p.lev_re[c], -p.lev_re[min(c + 1, p.n_re - 1)]);
    for (int c = tid; c < p.n_im; c += NT)
        tabim[c] = make_float2(-p.lev_im[c], -p.lev_im[min(c + 1, p.n_im - 1)]);
    for (int c = tid; c < p.A; c += NT) angs[c] = p.angles ? p.angles[c] : 0.f;
    float *ring_w = ring + (size_t)(warp < NW ? warp : 0) * W * 32;
    if (warp < NW)
        for (int c = lane; c < W * 32; c += 32) ring_w[c] = 0.f;   // slot 0 = csum[0] = 0 (pythran_dsp.py:28)
    FastAxis gre = make_fast_axis(p.lev_re, p.n_re, smem_u32(tabre));
    FastAxis gim = make_fast_axis(p.lev_im, p.n_im, smem_u32(tabim));
    // bounce the address constants through shared memory: ptxas otherwise splits (bits(w) << 3) + kaddr
    // into LEA + IADD of the two halves of kaddr, one more instruction per axis and evaluation
    if (tid == 0) {
        reinterpret_cast<uint32_t *>(ust)[0] = gre.kaddr;
        reinterpret_cast<uint32_t *>(ust)[1] = gim.kaddr;
    }
    __syncthreads();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(gre.kaddr) : "r"(smem_u32(ust)));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(gim.kaddr) : "r"(smem_u32(ust)));
    __syncthreads();
    if (tid < 2) ust[tid] = 0.f;
    __syncthreads();

    // edges: idx = 0 -> ph = angles[0], not unwrapped (phaserecovery.py:155 touches [N:-N] only)
    const long long lo = N < L ? N : L;
    const long long hi = (L - N > lo) ? L - N : lo;
    {
        const float a0 = angs[0];
        const long long nedge = lo + (L - hi);
        for (long long c = tid; c < nedge; c += NT) {
            const long long j = c < lo ? c : hi + (c - lo);
            if (idx) idx[j] = 0;
            if (ph) ph[j] = a0;
            if (Eout) Eout[j] = rotate_f(E[j], a0);
        }
    }

    // ---- this lane's angle column ------------------------------------------------------------------
    const int colw = warp % NW;                    // angle block of this warp (chain warp w and its producers)
    const float2 cc = p.comp[colw * 32 + lane];
    // e.x * (cr, ci) and e.y * (-ci, cr) (negation commutes with rounding).  Both constants are read back
    // from shared memory as 64-bit values: a pair that ptxas can re-derive from cc is re-packed with
    // MOV + FADD in every group of the inner loop.
    f32x2 c1, c2;
    {
        // part[] is not in use yet (NPG > 0: the distance tile, which is large enough for every thread of the CTA)
        float4 *bounce = reinterpret_cast<float4 *>(NPG > 0 ? (void *)dbuf : (void *)part) + tid;
        *bounce = make_float4(cc.x, cc.y, -cc.y, cc.x);
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(c1), "=l"(c2) : "r"(smem_u32(bounce)) : "memory");
        __syncthreads();
    }
    float csum = 0.f;
    const uint32_t ring_lo = smem_u32(ring_w) + 4u * lane;
    const uint32_t stage_addr = smem_u32(stage + (warp < NW ? warp : 0) * FAST_TR);
    uint32_t dtile_addr = 0;                        // NPG > 0: this lane's column of the current distance tile
    float cum = 0.f, p4prev = 0.f;                  // unwrap state (NW == 1: registers; else via ust[])

    const long long ntiles = (L + FAST_TR - 1) / FAST_TR;
    float2 enext = lane < L ? E[lane] : make_float2(0.f, 0.f);
    uint2 mine = make_uint2(0xffffffffu, 0u);       // (min bits, ballot) of tile row `lane`

    // distance of one row to the nearest alphabet point after rotation by this lane's test angle
    auto dist = [&](float2 ev) { return fast_dist(ev, c1, c2, gre, gim); };

    if (NPG > 0 && warp >= NW) {
        // =========================== producers: distances of tile m into buffer m & 1 ===========================
        const int pg = (warp - NW) / NW;            // this warp's rows of a tile: pg, pg + NPG, ...
        constexpr int RPW = FAST_TR / (NPG > 0 ? NPG : 1);
        for (long long m = 0; m < ntiles; m++) {
            const long long i0 = m * FAST_TR;
            float2 ev[RPW];
#pragma unroll
            for (int k = 0; k < RPW; k++) {
                const long long i = i0 + pg + (long long)k * NPG;
                ev[k] = i < L ? __ldg(E + i) : make_float2(0.f, 0.f);   // rows past the end: masked in the tail
            }
            if (m >= 2) {                           // the chain has finished with tile m - 2: its buffer is free
                if (m & 1) fbar_sync<BAR_EMPTY + 1>(NT);
                else fbar_sync<BAR_EMPTY>(NT);
            }
            float *dst = dbuf + (size_t)(m & 1) * FAST_TR * p.A + colw * 32 + lane;
#pragma unroll
            for (int k = 0; k < RPW; k++) dst[(pg + k * NPG) * p.A] = dist(ev[k]);
            __syncwarp();
            if (m & 1) fbar_arrive<BAR_FULL + 1>(NT);
            else fbar_arrive<BAR_FULL>(NT);
        }
        return;
    }

    // FAST_NR consecutive rows of this lane's column: distance -> running sum -> window difference -> warp
    // arg-min record.  The distance evaluations are independent (ILP); only the running sum chains.
    // Rows 2k, 2k+1 of the group sit at ring address aP[k] (+128): the ring has an even number of rows and
    // groups start at even slots, so a pair never straddles the wrap.  Shared-memory accesses are volatile
    // asm in exactly the order wanted: the input rows, the old sums, the new sums.
    auto group = [&](bool first, int rbase, const uint32_t (&aP)[FAST_NR / 2]) {
        float old[FAST_NR], c[FAST_NR];
        if (NPG == 0) {
            float2 e[FAST_NR];
#pragma unroll
            for (int k = 0; k < FAST_NR / 2; k++)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(e[2 * k].x), "=f"(e[2 * k].y), "=f"(e[2 * k + 1].x), "=f"(e[2 * k + 1].y)
                             : "r"(stage_addr + 8u * rbase + 16u * k));
            // csum[i - 2N] (0 while i < 2N: unused)
#pragma unroll
            for (int k = 0; k < FAST_NR / 2; k++) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old[2 * k]) : "r"(aP[k]));
                asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(old[2 * k + 1]) : "r"(aP[k]));
            }
#pragma unroll
            for (int u = 0; u < FAST_NR; u++) c[u] = dist(e[u]);
        } else {
            // the producers have evaluated the distances: one load per row
            const uint32_t da = dtile_addr + 4u * (uint32_t)p.A * (uint32_t)rbase;
#pragma unroll
            for (int u = 0; u < FAST_NR; u++)
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c[u]) : "r"(da + 4u * (uint32_t)p.A * u));
#pragma unroll
            for (int k = 0; k < FAST_NR / 2; k++) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old[2 * k]) : "r"(aP[k]));
                asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(old[2 * k + 1]) : "r"(aP[k]));
            }
        }
        if (first) c[0] = 0.f;                                          // row 0 is never added (:30)
#pragma unroll
        for (int u = 0; u < FAST_NR; u++) {
            csum = __fadd_rn(csum, c[u]);                               // :33/:36, sequential
            c[u] = csum;
        }
#pragma unroll
        for (int k = 0; k < FAST_NR / 2; k++) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(aP[k]), "f"(c[2 * k]));
            asm volatile("st.shared.f32 [%0+128], %1;" ::"r"(aP[k]), "f"(c[2 * k + 1]));
        }
#pragma unroll
        for (int u = 0; u < FAST_NR; u++) {
            const unsigned db = __float_as_uint(__fsub_rn(c[u], old[u]));   // >= +0: orders like an unsigned
            const unsigned mn = __reduce_min_sync(FULL, db);
            const unsigned bal = __ballot_sync(FULL, db == mn);
            if (lane == rbase + u) mine = make_uint2(mn, bal);
        }
    };
    auto group_at = [&](bool first, int rbase, int slot) {   // slot: ring slot of the group's first row (even)
        uint32_t aP[FAST_NR / 2];
        if (slot + FAST_NR <= W) {                             // no wrap inside the group: immediate offsets
            const uint32_t aA = ring_lo + 128u * (uint32_t)slot;
#pragma unroll
            for (int k = 0; k < FAST_NR / 2; k++) aP[k] = aA + 256u * k;
            group(first, rbase, aP);
        } else {
#pragma unroll
            for (int k = 0; k < FAST_NR / 2; k++) {
                int sk = slot + 2 * k;
                if (sk >= W) sk -= W;
                aP[k] = ring_lo + 128u * (uint32_t)sk;
            }
            group(first, rbase, aP);
        }
    };

    int slot = 0;   // ring slot (= row index mod 2N) of the next group's first row; always even
    for (long long m = 0; m < ntiles; m++) {
        const long long i0 = m * FAST_TR;
        const int nrows = (int)min((long long)FAST_TR, L - i0);
        if (NPG == 0) {
            __syncwarp();
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stage_addr + 8u * lane), "f"(enext.x), "f"(enext.y)
                         : "memory");
            __syncwarp();
            if (i0 + FAST_TR + lane < L) enext = E[i0 + FAST_TR + lane];
        } else {
            if (m & 1) fbar_sync<BAR_FULL + 1>(NT);     // the distances of tile m are in buffer m & 1
            else fbar_sync<BAR_FULL>(NT);
            dtile_addr = smem_u32(dbuf + (size_t)(m & 1) * FAST_TR * p.A + colw * 32 + lane);
        }

        // rows past the end of the stream (last group only) run on stale inputs: they come after every
        // valid row in the running sums and are masked in the tail
        const int ngroups = (nrows + FAST_NR - 1) / FAST_NR;
#pragma unroll 1
        for (int g = 0; g < ngroups; g++) {
            group_at(m == 0 && g == 0, FAST_NR * g, slot);   // the stream's first row contributes nothing
            slot += FAST_NR;
            if (slot >= W) slot -= W;
        }
        __syncwarp();
        if (NPG > 0) {                                  // the producers may refill this buffer (tile m + 2)
            if (m & 1) fbar_arrive<BAR_EMPTY + 1>(NT);
            else fbar_arrive<BAR_EMPTY>(NT);
        }

        // ---- hand the 32 row records to the tail owner ----------------------------------------------
        const int owner = (int)(m % NW);
        uint2 *pbuf = part + (size_t)(m & 1) * NW * FAST_TR;
        if (NW > 1) {
            if (warp != owner) pbuf[warp * FAST_TR + lane] = mine;
            if (NPG > 0) fbar_sync<BAR_CHAIN>(NC);      // chain warps only: the producers are elsewhere
            else __syncthreads();
        }
        if (warp == owner) {
            // tail, lane = row: row i = i0 + lane with i >= 2N produces output j = i - N
            unsigned best = 0xffffffffu, bb = 1u;
            int bw = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {   // ascending angle blocks, strict <: first minimum (:39)
                const uint2 q = (w == owner || NW == 1) ? mine : pbuf[w * FAST_TR + lane];
                if (q.x < best) {
                    best = q.x;
                    bb = q.y;
                    bw = w;
                }
            }
            int bk = 32 * bw + __ffs(bb) - 1;
            const long long i = i0 + lane, j = i - N;
            const bool valid = lane < nrows && i >= W;
            if (best >= 0x447a0000u || !valid) bk = 0;   // dmin0 = 1000 (:31): nothing below it -> idx stays 0
            if (idx && valid) idx[j] = bk;
            if (ph) {
                if (NW > 1) {
                    cum = ust[0];
                    p4prev = ust[1];
                }
                const float p4 = __fmul_rn(angs[bk], 4.f);
                float pp = __shfl_up_sync(FULL, p4, 1);
                if (lane == 0) pp = p4prev;
                float corr = 0.f;
                // np.unwrap only acts where |dd| >= pi: skip its fmod arithmetic for the (usual) tile without such a row
                const bool cand = valid && j > N && !(fabsf(__fsub_rn(p4, pp)) < 3.14159274101257324219f);
                if (__any_sync(FULL, cand) && valid && j > N) corr = unwrap_corr_f(p4, pp);
                unsigned mask = __ballot_sync(FULL, corr != 0.f);
                float mycum = cum;
                while (mask) {   // fold the (rare) non-zero corrections in row order: exact sequential sum
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float ce = __shfl_sync(FULL, corr, e);
                    cum = __fadd_rn(cum, ce);
                    if (lane >= e) mycum = cum;
                }
                const float phv = __fadd_rn(p4, mycum) / 4.f;
                if (valid) {
                    ph[j] = phv;
                    if (Eout) Eout[j] = rotate_f(E[j], phv);
                }
                const unsigned vm = __ballot_sync(FULL, valid);
                if (vm) p4prev = __shfl_sync(FULL, p4, 31 - __clz(vm));
                if (NW > 1) {
                    __syncwarp();   // every lane has read the carried state before lane 0 replaces it
                    if (lane == 0) {
                        ust[0] = cum;
                        ust[1] = p4prev;
                    }
                }
            }
        }
    }
}

static size_t fast_smem_bytes(int NW, int A, int n_re, int n_im, int N, int npg)
{
    return (size_t)(n_re + n_im) * 8 + (size_t)NW * FAST_TR * 8 * 3 + (size_t)A * 4 + 8 +
           (size_t)NW * 2 * N * 32 * 4 + (npg > 0 ? (size_t)2 * FAST_TR * A * 4 + 16 : 0);
}

constexpr int FAST_NPG = 4;   // producer row groups of the few-streams mapping: 4 * NW producer warps per CTA

template <int NW>
static int launch_fast(const BpsFastParams &p, int64_t nstream, bool split, cudaStream_t st)
{
    const size_t smem = fast_smem_bytes(NW, p.A, p.n_re, p.n_im, p.N, split ? FAST_NPG : 0);
    // set on every launch: the attribute belongs to the device that is current, and it is cheap
    if (split) {
        QB_CUDA_CHECK(cudaFuncSetAttribute(bps_fast_kernel<NW, FAST_NPG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
        bps_fast_kernel<NW, FAST_NPG><<<(unsigned)nstream, 32 * NW * (1 + FAST_NPG), smem, st>>>(p);
    } else {
        QB_CUDA_CHECK(cudaFuncSetAttribute(bps_fast_kernel<NW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
        bps_fast_kernel<NW, 0><<<(unsigned)nstream, 32 * NW, smem, st>>>(p);
    }
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

// Returns QB_OK after launching, or 1 if this problem is not covered by the fast kernel (caller falls
// back to bps_kernel), or a negative error code.
int bps_fast_dispatch(const void *E, int64_t nstream, int64_t stream_stride, int64_t L, const void *comp,
                      const void *angles, int64_t A, const void *lev_re, int64_t n_re, const void *lev_im,
                      int64_t n_im, int64_t N, int32_t *idx, void *ph, void *Eout, cudaStream_t st)
{
    if (n_re < 1 || n_im < 1 || A % 32 != 0 || A > 128 || 2 * N < FAST_NR) return 1;   // a group must fit the ring
    const int NW = (int)(A / 32);
    if (fast_smem_bytes(NW, (int)A, (int)n_re, (int)n_im, (int)N, FAST_NPG) > 100 * 1024) return 1;
    // Few streams: a call lasts as long as one stream is deep -> producer / chain split (bit-identical results, so the
    // choice may follow the launch size).  Option BPS_SPLIT = 0 / 1 (qb_set_option) forces a mapping (tests run both).
    bool split = nstream <= 148 && A <= 64;      // at most one CTA per SM; 2 chain + 8 producer warps at A = 64
    if (const char e = option_char(OPT_BPS_SPLIT)) split = e == '1' && A <= 64;
    // few LONG streams (one capture): the phase-parallel form (bps_par.cu)
    const bool par = bps_par_wanted(nstream, L, A, idx == nullptr, 4);
    BpsFastParams p;
    p.E = (const float2 *)E;
    p.comp = (const float2 *)comp;
    p.angles = (const float *)angles;
    p.lev_re = (const float *)lev_re;
    p.lev_im = (const float *)lev_im;
    p.idx = idx;
    p.ph = (float *)ph;
    p.Eout = (float2 *)Eout;
    p.stream_stride = stream_stride;
    p.L = L;
    p.A = (int)A;
    p.n_re = (int)n_re;
    p.n_im = (int)n_im;
    p.N = (int)N;
    if (par) return bps_par_launch(p, nstream, st);
    switch (NW) {
    case 1: return launch_fast<1>(p, nstream, split, st);
    case 2: return launch_fast<2>(p, nstream, split, st);
    case 3: return launch_fast<3>(p, nstream, false, st);
    default: return launch_fast<4>(p, nstream, false, st);
    }
}

}  // namespace qb
