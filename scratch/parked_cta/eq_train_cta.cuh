// eq_train, ONE STREAM PER CTA (complex64, os = 2, fixed or adaptive step size): the kernel for calls whose time is the
// serial depth of a stream -- the reference's own call shape (train_equaliser on ONE capture trains one stream per
// mode, pythran_equalisation.py:162-172) and long, converged time segments.
//
// The recurrence  y_i = X_i . W_i ,  c_i = mu errfct(y_i) ,  W_{i+1} = W_i + c_i conj(X_i)  is rewritten block-exactly
// (Benesty & Duhamel's fast exact LMS): for any older tap state W_s, s <= i,
//
//     y_i = X_i . W_s  +  sum_{j = s}^{i-1} c_j G[i, j] ,      G[i, j] = sum_{k,t} x_k[2i + t] conj(x_k[2j + t])
//
// (exact algebra; G is a property of the signal alone).  Work is split over specialised warps of one CTA, blocks of
// CT_B = 8 symbols, connected by mbarrier-guarded rings in shared memory:
//
//   LD    stages the stream: TMA bulk copies (cp.async.bulk; plain loads at the ragged ends or for unaligned rows) of
//         64-sample-pair chunks, re-ordered into a ring of sample "quads" [re(2g), re(2g+1), im(2g), im(2g+1)] -- the
//         form every consumer multiplies pair-wise with FFMA2.
//   GA    two warps on alternate blocks, lane = lag l (0..31): lag products r_l[n] = sum_k x_k[n] conj(x_k[n - 2l]) of
//         one sample pair per step, handed on as (u, q) = (what enters the window, what the pair adds up to).
//   GB    lane = lag l: sliding window G[i, i-l] = G[i-1, i-1-l] + u - q_old (re-anchored to an exactly summed window
//         every 1024 symbols), written as ROW j = i - l of the lag matrix (what step j of the chain needs) at column
//         (i mod 32) + (j mod 32): conflict free for the writer (lanes = lags) and for the reader (lanes = symbols).
//   UPD   lane = two adjacent tap pairs (register window over the samples): W += c_j conj(X_j) for the 8 steps of a
//         finished block, then a snapshot of the taps.
//   BASE  8 lanes per symbol: base_i = X_i . S_n with the snapshot S_n = W_{8(n-3)} for the symbols of block n.
//   CHAIN the only serial part.  Lane L owns the symbols i = L (mod 32) and accumulates F_i = sum_{lags >= 2} c_j G[i,j]
//         (one complex FMA per step and lane); per step j:  c_j = mu errfct(y_j);  y_{j+1} = (F_{j+1} + base_{j+1}) +
//         c_j G[j+1, j]  in every lane redundantly (F_{j+1} broadcast by ONE shuffle pair issued a step ahead, so no
//         shuffle and no reduction sits on the dependent chain);  F += c_j G[., j].  Dependent chain per symbol:
//         one complex FMA + the error function.
//
// Symbol i = 8n + b uses the snapshot of three blocks ago, i.e. lags 1 .. 24 + b <= 31; GB writes zeros for the other
// lags.  A snapshot that old leaves 16 symbol times for UPD -> snapshot -> BASE, so the chain never waits for them.
// All arithmetic is fp32 FMA; results equal the direct recurrence to rounding (numpy model: scratch/cta_model.py,
// 7e-7 rms on the error signal; the parity tests hold every trainer to the same 1e-5 against the oracle).
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "eq_train_fast.cuh"

namespace qb {

constexpr int CT_B = 8;        // symbols per block
constexpr int CT_NS = 16;      // barrier slots per block ring
constexpr int CT_ROWS = 64;    // rows of the lag matrix in flight
constexpr int CT_QR = 512;     // sample quads (pairs of samples) per polarisation in the ring
constexpr int CT_QDUP = 32;    // head of the ring repeated behind its end: a window never wraps
constexpr int CT_PRE = 64;     // zero quads in front of sample 0 (the Gram warps reach back 31 symbols)
constexpr int CT_CH = 64;      // quads per loader chunk
constexpr int CT_NCH = CT_QR / CT_CH;
constexpr int CT_PR = 64;      // depth of the lag-product ring (steps)
constexpr int CT_MAXTP = 64;   // tap pairs (nmodes * ceil(ntaps/2)) this kernel holds: one per UPD lane
constexpr int CT_NWARPS = 8;
constexpr int CT_RA = 128;     // blocks between exact re-anchorings of the sliding lag sums
enum { W_CHAIN = 0, W_GA0 = 1, W_GB = 2, W_BASE0 = 3, W_LD = 4, W_UPD = 5, W_GA1 = 6, W_BASE1 = 7 };

template <int NM>
struct CtaSmem {
    float4 xq[NM][CT_QR + CT_QDUP];   // sample quads, 8.5 kB per polarisation
    float4 grow[CT_ROWS][64];         // 64 kB: (G.re, G.im, -G.im, G.re) at [j mod 64][(i mod 32) + (j mod 32)]
    float4 pring[CT_PR][32];          // 32 kB: (u.re, u.im, q.re, q.im) per lag, see GA
    float2 anchor[32];                // exactly summed window per lag (every CT_RA blocks)
    float4 snaps[4][CT_MAXTP];        // 4 kB: (wr[2p], wr[2p+1], wi[2p], wi[2p+1])
    float2 nbase[CT_ROWS];            // base_{j+2} of step j
    float4 ering[32];                 // (e.re, e.im, c.re, c.im)
    float4 stage[NM][CT_CH];          // raw TMA landing zone, one row per polarisation
    unsigned long long sfull[CT_NCH], gafull[CT_NS], gbdone[CT_NS], cin[CT_NS], efull[CT_NS], snapfull[4], tma;
    int chain_done;                   // chain blocks finished (polled by the loader)
};

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void warp_arrive(unsigned long long *bar, int lane)
{
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

struct CtaGeom {
    int npair;            // ceil(ntaps / 2)
    int tp;               // nmodes * npair
    int pb;               // warm-up blocks of the Gram warps = ceil((npair - 1) / 8)
    int nblk;             // chain blocks = ceil(TrSyms / 8)
    int nchunks;          // loader chunks
    long long err_off;    // iteration offset into the err rows
    int aligned16;        // every row of the stream starts 16-byte aligned (TMA allowed)
    long long *prof;      // QB_CTA_PROF=1: per warp of CTA 0 (total clocks, clocks spent in waits 0..3), else NULL
};

// ---- the serial chain (one warp) ------------------------------------------------------------------------------------
// Packed forms of the hot error functions: e and c = mu e as (re, im) pairs.
template <int METHOD>
__device__ __forceinline__ void err_pair(int method, f32x2 y2, const ErrConst &ec, const float2 *syms, int K,
                                         const float2 *gsyms, int j, int lane, f32x2 &e2)
{
    if (METHOD == QB_MCMA) {            // (R - y^2) y per axis, pythran_equalisation.py:190-194
        const f32x2 R2 = pack2(ec.Rr, ec.Ri);
        e2 = mul2(sub2(R2, mul2(y2, y2)), y2);
    } else {
        const float2 e = err_fast<METHOD, 32, -1>(method, unpack2(y2), ec, syms, K, gsyms, j, lane);
        e2 = pack2(e.x, e.y);
    }
}

template <int NM, int METHOD, bool ADAPT>
__device__ __forceinline__ void cta_chain(CtaSmem<NM> &sm, const TrainParams<float> &p, const CtaGeom &g,
                                          const float2 *syms_sm, int mode, long long stream, int lane, bool prof,
                                          long long *pw)
{
    const int T = (int)p.TrSyms, nblk = g.nblk;
    ErrConst ec = load_err_const<METHOD>(syms_sm, p.nsym_smem);
    float mu = p.mu[stream];
    const float2 *gsyms = p.symbols + (long long)mode * p.K;
    auto wait_cin = [&](int q, int slot) {
        uint64_t *bar = reinterpret_cast<uint64_t *>(&sm.cin[q & (CT_NS - 1)]);
        if (prof) {
            const long long t = clock64();
            mbar_wait(bar, (uint32_t)((q >> 4) & 1));
            pw[slot] += clock64() - t;
        } else {
            mbar_wait(bar, (uint32_t)((q >> 4) & 1));
        }
    };
    wait_cin(0, 0);
    const float2 y0 = sm.nbase[CT_ROWS - 2], v1 = sm.nbase[CT_ROWS - 1];
    f32x2 Y = pack2(y0.x, y0.y);        // base_0
    f32x2 vb = pack2(v1.x, v1.y);       // base_1
    f32x2 F = 0ull;
    float2 eprev = make_float2(0.f, 0.f);
    // operands of the step to come, loaded one step ahead (a block's first step: at the end of the block before -- the
    // rows and bases it needs are covered by that block's barrier): row entry of this lane, lag-1 entry, base
    float4 row = sm.grow[0][lane], ngv = sm.grow[0][1];
    float2 bs = sm.nbase[0];
    // one block of 8 steps; TAILCHK: the block may run past the end of the stream (steps >= T take a zero step)
    auto block = [&](auto tailchk, int q) {
        constexpr bool TAILCHK = decltype(tailchk)::value;
        const int j0 = 8 * q;
        const int r0 = j0 & (CT_ROWS - 1), c0 = j0 & 31;
        const int r1 = (r0 + 8) & (CT_ROWS - 1), c1 = (c0 + 8) & 31;
        const float4 *rowp = &sm.grow[r0][c0 + lane];
        const float4 *ngp = &sm.grow[r0][2 * c0 + 1];        // lag 1 of step j: column ((j+1) mod 32) + (j mod 32)
        const float2 *bsp = &sm.nbase[r0];
        float4 *erp = &sm.ering[c0];
        const int src = c0 + 2;
        const int own = lane - c0;          // == b at the step whose symbol this lane has just handed over
#pragma unroll
        for (int b = 0; b < CT_B; b++) {
            const int j = j0 + b;
            const bool live = !TAILCHK || j < T;
            // operands of step j + 1
            float4 row_n, ng_n;
            float2 bs_n;
            if (b < CT_B - 1) {
                row_n = rowp[(b + 1) * 65];
                // the one step whose successor's owner lane wraps: (j+1) mod 32 = 31 -> next lag-1 column is 31 + 0
                ng_n = (b == CT_B - 2) ? (c0 == 24 ? sm.grow[r0 + b + 1][31] : ngp[(b + 1) * 66]) : ngp[(b + 1) * 66];
                bs_n = bsp[b + 1];
            } else {
                row_n = sm.grow[r1][c1 + lane];
                ng_n = sm.grow[r1][2 * c1 + 1];
                bs_n = sm.nbase[r1];
            }
            f32x2 e2;
            err_pair<METHOD>(p.method, Y, ec, syms_sm, p.K, gsyms, live ? j : 0, lane, e2);
            const float mul = live ? mu : 0.f;
            const float2 c = mul2_bcast(mul, e2);
            const float2 e = unpack2(e2);
            Y = fma2_bcast(c.y, pack2(ngv.z, ngv.w), fma2_bcast(c.x, pack2(ngv.x, ngv.y), vb));
            F = fma2_bcast(c.y, pack2(row.z, row.w), fma2_bcast(c.x, pack2(row.x, row.y), F));
            // the lane whose symbol j has just been consumed starts over for symbol j + 32
            float2 f = unpack2(F);
            f.x = own == b ? 0.f : f.x;
            f.y = own == b ? 0.f : f.y;
            F = pack2(f.x, f.y);
            const float vx = __shfl_sync(0xffffffffu, f.x, src + b), vy = __shfl_sync(0xffffffffu, f.y, src + b);
            if (ADAPT) {   // pythran_equalisation.py:171-172: after symbol i > 0 of an iteration
                mu = adapt_step_sel(mu, e, eprev, live && j > 0);
                eprev = e;
            }
            if (lane == 0) erp[b] = make_float4(e.x, e.y, c.x, c.y);
            vb = add2(pack2(vx, vy), pack2(bs.x, bs.y));      // F_{j+2} + base_{j+2}
            row = row_n;
            ngv = ng_n;
            bs = bs_n;
        }
        if (lane == 0) {
            *reinterpret_cast<volatile int *>(&sm.chain_done) = q + 1;
            mbar_arrive(&sm.efull[q & (CT_NS - 1)]);
        }
    };
    bool ready = true;
    const int nfull = T / CT_B;
    for (int q = 0; q < nblk; q++) {
        if (!ready) wait_cin(q, 1);
        // the barrier of the next block is tested a block ahead: its latency never shows on the chain
        ready = q + 1 < nblk ? mbar_test(&sm.cin[(q + 1) & (CT_NS - 1)], (uint32_t)(((q + 1) >> 4) & 1)) : true;
        if (q < nfull) block(std::false_type{}, q);
        else block(std::true_type{}, q);
    }
    if (lane == 0) p.mu[stream] = mu;
}

template <int NM>
__global__ void __launch_bounds__(32 * CT_NWARPS, 1) train_cta_kernel(TrainParams<float> p, CtaGeom g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaSmem<NM> &sm = *reinterpret_cast<CtaSmem<NM> *>(smem_raw);
    float2 *syms_sm = reinterpret_cast<float2 *>(smem_raw + sizeof(CtaSmem<NM>));
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long stream = blockIdx.x;
    const long long seg = stream / p.nsel;
    const int mode = p.modes.m[(int)(stream % p.nsel)];
    const int ntaps = p.ntaps, NP = g.npair;
    const int T = (int)p.TrSyms, nblk = g.nblk;
    const float2 *Eseg = p.E + seg * p.seg_stride;
    const int Lvalid = (T - 1) * 2 + ntaps;          // samples of a row the caller guarantees

    if (threadIdx.x == 0) {
        for (int i = 0; i < CT_NCH; i++) mbar_init(reinterpret_cast<uint64_t *>(&sm.sfull[i]), 1);
        for (int i = 0; i < CT_NS; i++) {
            mbar_init(reinterpret_cast<uint64_t *>(&sm.gafull[i]), 1);
            mbar_init(reinterpret_cast<uint64_t *>(&sm.gbdone[i]), 1);
            mbar_init(reinterpret_cast<uint64_t *>(&sm.cin[i]), 3);        // GB + two BASE warps
            mbar_init(reinterpret_cast<uint64_t *>(&sm.efull[i]), 1);
        }
        for (int i = 0; i < 4; i++) mbar_init(reinterpret_cast<uint64_t *>(&sm.snapfull[i]), 1);
        mbar_init(reinterpret_cast<uint64_t *>(&sm.tma), 1);
        sm.chain_done = 0;
        mbar_fence_init();
    }
    {
        const float2 *gs = p.symbols + (long long)mode * p.K;
        for (int c = threadIdx.x; c < p.nsym_smem; c += blockDim.x) syms_sm[c] = gs[c];
        // products of sample pairs in front of the stream (all zero) are read as "old" terms by the first symbols
        float4 *pr = &sm.pring[0][0];
        for (int c = threadIdx.x; c < CT_PR * 32; c += blockDim.x) pr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // QB_CTA_PROF: where does every warp of CTA 0 spend its time?  (clock64 only around waits; nothing in the loops)
    const bool prof = g.prof != nullptr && blockIdx.x == 0;
    long long pw[4] = {0, 0, 0, 0};
    const long long pt0 = prof ? clock64() : 0;
    auto twait = [&](int slot, unsigned long long *bar, uint32_t parity) {
        if (prof) {
            const long long t = clock64();
            mbar_wait(reinterpret_cast<uint64_t *>(bar), parity);
            pw[slot] += clock64() - t;
        } else {
            mbar_wait(reinterpret_cast<uint64_t *>(bar), parity);
        }
    };
    // block-ring barrier of block index x (>= 0): slot x % CT_NS, phase parity (x / CT_NS) & 1
    auto rwait = [&](int slot, unsigned long long *bars, int x) {
        twait(slot, bars + (x & (CT_NS - 1)), (uint32_t)((x >> 4) & 1));
    };
    auto pdone = [&]() {
        if (prof && lane == 0) {
            long long *o = g.prof + wid * 8;
            o[0] = clock64() - pt0;
            for (int i = 0; i < 4; i++) o[1 + i] = pw[i];
        }
    };

    if (wid == W_LD) {
        // ---------------------------------------------------------------------------------------------------------
        // chunk c holds the ring positions a in [64c, 64c + 64), i.e. sample quads g = a - CT_PRE of the stream
        uint32_t tma_phase = 0;
        for (int c = 0; c < g.nchunks; c++) {
            if (c >= CT_NCH) {
                // The slots were last read by the update warps, and UPD(u) is over once the chain has finished block
                // u + 3.  The loader may be arbitrarily late relative to the chain, so it polls a monotone counter
                // instead of waiting on a phase of the 16-slot barrier ring.
                const int need = 8 * c - 62;
                if (need >= 0) {
                    const int want = (need < nblk ? need : nblk - 1) + 1;
                    const long long t = prof ? clock64() : 0;
                    while (*reinterpret_cast<volatile int *>(&sm.chain_done) < want) __nanosleep(64);
                    __threadfence_block();
                    if (prof) pw[0] += clock64() - t;
                }
            }
            const int g0 = c * CT_CH - CT_PRE;
            const bool interior = g0 >= 0 && 2 * (g0 + CT_CH) <= Lvalid;
            if (interior && g.aligned16) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(reinterpret_cast<uint64_t *>(&sm.tma), (uint32_t)(NM * CT_CH * 16));
#pragma unroll
                    for (int k = 0; k < NM; k++)
                        tma_bulk_g2s(&sm.stage[k][0], Eseg + (long long)k * p.row_stride + 2 * g0, CT_CH * 16,
                                     reinterpret_cast<uint64_t *>(&sm.tma));
                }
                twait(1, &sm.tma, tma_phase);
                tma_phase ^= 1;
#pragma unroll
                for (int k = 0; k < NM; k++) {
#pragma unroll
                    for (int r = 0; r < CT_CH / 32; r++) {
                        const int q = lane + 32 * r;
                        const float4 v = sm.stage[k][q];                       // (re0, im0, re1, im1)
                        const float4 o = make_float4(v.x, v.z, v.y, v.w);
                        const int slot = (c * CT_CH + q) & (CT_QR - 1);
                        sm.xq[k][slot] = o;
                        if (slot < CT_QDUP) sm.xq[k][CT_QR + slot] = o;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < NM; k++) {
                    const float2 *row = Eseg + (long long)k * p.row_stride;
#pragma unroll
                    for (int r = 0; r < CT_CH / 32; r++) {
                        const int q = lane + 32 * r;
                        const int n0 = 2 * (g0 + q);
                        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
                        if (n0 >= 0 && n0 < Lvalid) s0 = __ldg(row + n0);
                        if (n0 + 1 >= 0 && n0 + 1 < Lvalid) s1 = __ldg(row + n0 + 1);
                        const float4 o = make_float4(s0.x, s1.x, s0.y, s1.y);
                        const int slot = (c * CT_CH + q) & (CT_QR - 1);
                        sm.xq[k][slot] = o;
                        if (slot < CT_QDUP) sm.xq[k][CT_QR + slot] = o;
                    }
                }
            }
            warp_arrive(&sm.sfull[c & (CT_NCH - 1)], lane);
        }
        pdone();
        return;
    }

    // chunk that holds ring position a must have landed (chunks land in order; `have` = chunks known to be there)
    auto need_chunk = [&](int a, int &have) {
        const int c = a >> 6;
        while (have <= c) {
            twait(3, &sm.sfull[have & (CT_NCH - 1)], (uint32_t)((have >> 3) & 1));
            have++;
        }
    };

    if (wid == W_GA0 || wid == W_GA1) {
        // ---------------------------------------------------------------------------------------------------------
        // Gram block gb <-> symbols i = 8 (gb - pb) + b; step s = i + NP - 1 multiplies sample quad s (the newest pair
        // of the window of symbol i) with the quads s - l: r_lo = x[2s] conj(x[2s - 2l]), r_hi = x[2s+1] conj(x[2s+1 - 2l])
        // (summed over the polarisations).  Ring positions a = s + CT_PRE >= 1.  Handed to GB per step:
        //     q = r_lo + r_hi                      what quad s adds up to (leaves the window NP steps later)
        //     u = q                (ntaps even)     what enters the window of symbol i
        //       = r_hi[s-1] + r_lo[s]   (ntaps odd: the window ends in the middle of quad s)
        // The two GA warps take alternate blocks (the products of different steps are independent).
        const int gsel = wid == W_GA0 ? 0 : 1;
        constexpr int GS = NM <= 2 ? 4 : 2;      // steps whose operands are loaded before any of them is used
        int have = 0;
        const int ngb = g.pb + nblk + 4;
        const bool odd = ntaps & 1;
        auto prod = [&](const float4 (&A)[NM], const float4 (&Bq)[NM], float2 &rlo, float2 &rhi) {
            f32x2 re = 0ull, im1 = 0ull, im2 = 0ull;
#pragma unroll
            for (int k = 0; k < NM; k++) {
                const f32x2 ar = pack2(A[k].x, A[k].y), ai = pack2(A[k].z, A[k].w);
                const f32x2 br = pack2(Bq[k].x, Bq[k].y), bi = pack2(Bq[k].z, Bq[k].w);
                re = fma2(ar, br, re);       // a conj(b) = (ar br + ai bi) + j (ai br - ar bi)
                re = fma2(ai, bi, re);
                im1 = fma2(ai, br, im1);
                im2 = fma2(ar, bi, im2);
            }
            const float2 r = unpack2(re), q = unpack2(sub2(im1, im2));
            rlo = make_float2(r.x, q.x);
            rhi = make_float2(r.y, q.y);
        };
        for (int gb = gsel; gb < ngb; gb += 2) {
            if (gb - 4 >= g.pb) rwait(0, sm.gbdone, gb - 4);
            const int m = gb - g.pb;
            const int i0 = 8 * m;
            need_chunk(i0 + 7 + NP - 1 + CT_PRE, have);
            const int a0 = i0 + NP - 1 + CT_PRE;
            float4 *out = &sm.pring[i0 & (CT_PR - 1)][lane];
            float2 prev_hi = make_float2(0.f, 0.f);
            if (odd) {      // r_hi of the quad before this block's first one (the other warp's last step)
                float4 A[NM], Bq[NM];
                const int sa = (a0 - 1) & (CT_QR - 1), sb = (sa - lane) & (CT_QR - 1);
#pragma unroll
                for (int k = 0; k < NM; k++) A[k] = sm.xq[k][sa], Bq[k] = sm.xq[k][sb];
                float2 t;
                prod(A, Bq, t, prev_hi);
            }
#pragma unroll
            for (int b0 = 0; b0 < CT_B; b0 += GS) {
                float4 A[GS][NM], Bq[GS][NM];
#pragma unroll
                for (int t = 0; t < GS; t++) {
                    const int sa = (a0 + b0 + t) & (CT_QR - 1), sb = (sa - lane) & (CT_QR - 1);
#pragma unroll
                    for (int k = 0; k < NM; k++) A[t][k] = sm.xq[k][sa], Bq[t][k] = sm.xq[k][sb];
                }
#pragma unroll
                for (int t = 0; t < GS; t++) {
                    float2 rlo, rhi;
                    prod(A[t], Bq[t], rlo, rhi);
                    const float2 q = make_float2(rlo.x + rhi.x, rlo.y + rhi.y);
                    const float2 u = odd ? make_float2(prev_hi.x + rlo.x, prev_hi.y + rlo.y) : q;
                    prev_hi = rhi;
                    out[(b0 + t) * 32] = make_float4(u.x, u.y, q.x, q.y);
                }
            }
            if (m >= 0 && (m & (CT_RA - 1)) == 0) {
                // exact window of symbol i0: quads i0 .. i0 + NP - 1 (only the first half of the last one if ntaps is odd)
                float2 acc = make_float2(0.f, 0.f);
                for (int gq = 0; gq < NP; gq++) {
                    float4 A[NM], Bq[NM];
                    const int sa = (i0 + gq + CT_PRE) & (CT_QR - 1), sb = (sa - lane) & (CT_QR - 1);
#pragma unroll
                    for (int k = 0; k < NM; k++) A[k] = sm.xq[k][sa], Bq[k] = sm.xq[k][sb];
                    float2 rlo, rhi;
                    prod(A, Bq, rlo, rhi);
                    const bool half = odd && gq == NP - 1;
                    acc.x += rlo.x + (half ? 0.f : rhi.x);
                    acc.y += rlo.y + (half ? 0.f : rhi.y);
                }
                sm.anchor[lane] = acc;
            }
            warp_arrive(&sm.gafull[gb & (CT_NS - 1)], lane);
        }
        pdone();
        return;
    }

    if (wid == W_GB) {
        // ---------------------------------------------------------------------------------------------------------
        // lags this symbol may use: 1 .. 24 + b (the snapshot behind base_i already holds the older steps); lag 0 = 0
        float mk[CT_B];
#pragma unroll
        for (int b = 0; b < CT_B; b++) mk[b] = (lane >= 1 && lane <= 24 + b) ? 1.f : 0.f;
        float2 G = make_float2(0.f, 0.f);
        // byte offset of this lane's entry in the lag matrix, as (row << 10) | ((row & 31) << 4) for row = (i - lane)
        // mod 64; one step = + (1 << 10) + (1 << 4), both fields wrap by masking (bit 9 takes the column's carry)
        const int rw0 = (0 - lane) & (CT_ROWS - 1);
        uint32_t V = ((uint32_t)rw0 << 10) | (((uint32_t)rw0 & 31u) << 4);
        unsigned char *grow_b = reinterpret_cast<unsigned char *>(&sm.grow[0][0]);
        for (int gb = 0; gb < g.pb; gb++) rwait(0, sm.gafull, gb);       // products in front of symbol 0 (old terms)
        for (int m = 0; m < nblk + 4; m++) {
            rwait(0, sm.gafull, m + g.pb);
            if (m >= 8) rwait(1, sm.efull, m - 8);            // rows of chain block m - 8 have been read
            const int i0 = 8 * m;
            const float4 *in = &sm.pring[i0 & (CT_PR - 1)][lane];
            unsigned char *colb = grow_b + (i0 & 31) * 16;
            float2 u[CT_B], qo[CT_B];
#pragma unroll
            for (int b = 0; b < CT_B; b++) {
                u[b] = *reinterpret_cast<const float2 *>(&in[b * 32]);
                qo[b] = *(reinterpret_cast<const float2 *>(&sm.pring[(i0 + b - NP) & (CT_PR - 1)][lane]) + 1);
            }
            if ((m & (CT_RA - 1)) == 0) {     // exact window of symbol i0 (its sliding update is skipped)
                const float2 a = sm.anchor[lane];
                G = make_float2(a.x - u[0].x + qo[0].x, a.y - u[0].y + qo[0].y);
            }
#pragma unroll
            for (int b = 0; b < CT_B; b++) {
                G.x += u[b].x - qo[b].x;
                G.y += u[b].y - qo[b].y;
                const float gx = G.x * mk[b], gy = G.y * mk[b];
                *reinterpret_cast<float4 *>(colb + V + b * 16) = make_float4(gx, gy, -gy, gx);
                V = (V + 1040u) & 0xFDF0u;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.gbdone[(m + g.pb) & (CT_NS - 1)]);
                if (m >= 4 && m - 4 < nblk) mbar_arrive(&sm.cin[(m - 4) & (CT_NS - 1)]);
            }
        }
        pdone();
        return;
    }

    if (wid == W_BASE0 || wid == W_BASE1) {
        // ---------------------------------------------------------------------------------------------------------
        const int w = wid == W_BASE0 ? 0 : 1;
        const int bsym = 4 * w + (lane >> 3), h = lane & 7;
        int have = 0;
        for (int n = 0; n <= nblk; n++) {
            const int v = n > 3 ? n - 3 : 0;
            twait(0, &sm.snapfull[v & 3], (uint32_t)((v >> 2) & 1));
            need_chunk(8 * n + 7 + NP - 1 + CT_PRE, have);
            const int i = 8 * n + bsym;
            const int x0 = (i + CT_PRE) & (CT_QR - 1);
            f32x2 a1 = 0ull, a2 = 0ull, b1 = 0ull, b2 = 0ull;
#pragma unroll
            for (int k = 0; k < NM; k++) {
                const float4 *xr = &sm.xq[k][x0];
                const float4 *wr = &sm.snaps[v & 3][k * NP];
#pragma unroll 4
                for (int pq = h; pq < NP; pq += 8) {
                    const float4 x = xr[pq], wv = wr[pq];
                    const f32x2 xr2 = pack2(x.x, x.y), xi2 = pack2(x.z, x.w), wr2 = pack2(wv.x, wv.y), wi2 = pack2(wv.z, wv.w);
                    a1 = fma2(xr2, wr2, a1);
                    a2 = fma2(xi2, wi2, a2);
                    b1 = fma2(xr2, wi2, b1);
                    b2 = fma2(xi2, wr2, b2);
                }
            }
            const float2 sa = unpack2(sub2(a1, a2)), sb = unpack2(add2(b1, b2));
            float yr = sa.x + sa.y, yi = sb.x + sb.y;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                yr += __shfl_xor_sync(0xffffffffu, yr, d);
                yi += __shfl_xor_sync(0xffffffffu, yi, d);
            }
            if (h == 0) sm.nbase[(i - 2) & (CT_ROWS - 1)] = make_float2(yr, yi);
            __syncwarp();
            if (lane == 0 && n >= 1) mbar_arrive(&sm.cin[(n - 1) & (CT_NS - 1)]);
        }
        pdone();
        return;
    }

    if (wid == W_UPD) {
        // ---------------------------------------------------------------------------------------------------------
        // lane -> input polarisation k and the tap pairs pq0, pq0 + 1 (taps 2 pq0 .. 2 pq0 + 3): the window of a lane
        // moves by one sample quad per symbol, so one shared load per step feeds both pairs
        const int lpp = (NP + 1) / 2;                       // lanes per polarisation
        const bool act = lane < NM * lpp;
        const int k = act ? lane / lpp : 0, pq0 = act ? 2 * (lane % lpp) : 0;
        const bool hasp1 = act && pq0 + 1 < NP;
        float2 *wg = p.wx + ((long long)seg * NM + mode) * (long long)(NM * ntaps) + (long long)k * ntaps;
        bool vt[4];
        float2 w[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            vt[t] = act && 2 * pq0 + t < ntaps;
            w[t] = vt[t] ? wg[2 * pq0 + t] : make_float2(0.f, 0.f);
        }
        f32x2 WR0 = pack2(w[0].x, w[1].x), WI0 = pack2(w[0].y, w[1].y);
        f32x2 WR1 = pack2(w[2].x, w[3].x), WI1 = pack2(w[2].y, w[3].y);
        if (act) sm.snaps[0][k * NP + pq0] = make_float4(w[0].x, w[1].x, w[0].y, w[1].y);
        if (hasp1) sm.snaps[0][k * NP + pq0 + 1] = make_float4(w[2].x, w[3].x, w[2].y, w[3].y);
        warp_arrive(&sm.snapfull[0], lane);
        float2 *eg = p.err ? p.err + ((long long)seg * NM + mode) * ((long long)T * p.Niter) + g.err_off : nullptr;
        const float4 *xk = &sm.xq[k][pq0];
        for (int u = 0; u < nblk; u++) {
            rwait(0, sm.efull, u);
            const int i0 = 8 * u;
            const float4 *erp = &sm.ering[i0 & 31];
            const float4 *xr = xk + ((i0 + CT_PRE) & (CT_QR - 1));     // i0 + 8 stays inside the ring + its repeated head
            float4 ecs[CT_B], xs[CT_B + 1];
#pragma unroll
            for (int b = 0; b < CT_B; b++) ecs[b] = erp[b];
#pragma unroll
            for (int b = 0; b <= CT_B; b++) xs[b] = xr[b];
#pragma unroll
            for (int b = 0; b < CT_B; b++) {
                const float cr = ecs[b].z, ci = ecs[b].w;
                const f32x2 xr0 = pack2(xs[b].x, xs[b].y), xi0 = pack2(xs[b].z, xs[b].w);
                const f32x2 xr1 = pack2(xs[b + 1].x, xs[b + 1].y), xi1 = pack2(xs[b + 1].z, xs[b + 1].w);
                WR0 = fma2_bcast(ci, xi0, fma2_bcast(cr, xr0, WR0));          // (cr + j ci)(xr - j xi)
                WI0 = fma2_bcast(-cr, xi0, fma2_bcast(ci, xr0, WI0));
                WR1 = fma2_bcast(ci, xi1, fma2_bcast(cr, xr1, WR1));
                WI1 = fma2_bcast(-cr, xi1, fma2_bcast(ci, xr1, WI1));
            }
            if (eg && lane < CT_B && i0 + lane < T) eg[i0 + lane] = *reinterpret_cast<const float2 *>(&erp[lane]);
            // taps past ntaps stay exactly zero
            float2 r0 = unpack2(WR0), q0 = unpack2(WI0), r1 = unpack2(WR1), q1 = unpack2(WI1);
            if (!vt[1]) r0.y = q0.y = 0.f;
            if (!vt[2]) r1.x = q1.x = 0.f;
            if (!vt[3]) r1.y = q1.y = 0.f;
            WR0 = pack2(r0.x, r0.y), WI0 = pack2(q0.x, q0.y), WR1 = pack2(r1.x, r1.y), WI1 = pack2(q1.x, q1.y);
            if (act) sm.snaps[(u + 1) & 3][k * NP + pq0] = make_float4(r0.x, r0.y, q0.x, q0.y);
            if (hasp1) sm.snaps[(u + 1) & 3][k * NP + pq0 + 1] = make_float4(r1.x, r1.y, q1.x, q1.y);
            warp_arrive(&sm.snapfull[(u + 1) & 3], lane);
        }
        if (act) {
            const float2 r0 = unpack2(WR0), q0 = unpack2(WI0), r1 = unpack2(WR1), q1 = unpack2(WI1);
            if (vt[0]) wg[2 * pq0] = make_float2(r0.x, q0.x);
            if (vt[1]) wg[2 * pq0 + 1] = make_float2(r0.y, q0.y);
            if (vt[2]) wg[2 * pq0 + 2] = make_float2(r1.x, q1.x);
            if (vt[3]) wg[2 * pq0 + 3] = make_float2(r1.y, q1.y);
        }
        pdone();
        return;
    }

    // -------------------------------------------------------------------------------------------------------------
    // CHAIN: the error function and the step-size rule are compiled in per method (one switch outside the loop)
    if (nblk == 0) return;
#define QB_CTA_CHAIN(METHOD)                                                                                         \
    do {                                                                                                             \
        if (p.adaptive) cta_chain<NM, METHOD, true>(sm, p, g, syms_sm, mode, stream, lane, prof, pw);                \
        else cta_chain<NM, METHOD, false>(sm, p, g, syms_sm, mode, stream, lane, prof, pw);                          \
    } while (0)
    switch (p.method) {
    case QB_CMA:
    case QB_SGNCMA: QB_CTA_CHAIN(QB_CMA); break;
    case QB_MCMA: QB_CTA_CHAIN(QB_MCMA); break;
    case QB_RDE:
        if (p.K - (p.K + 1) / 2 <= 3) QB_CTA_CHAIN(METHOD_RDE3);
        else QB_CTA_CHAIN(QB_RDE);
        break;
    default:   // QB_MRDE (the host only launches the methods listed here)
        if (p.K - (p.K + 1) / 2 <= 3) QB_CTA_CHAIN(METHOD_MRDE3);
        else QB_CTA_CHAIN(QB_MRDE);
        break;
    }
#undef QB_CTA_CHAIN
    pdone();
}

// ---- host side ----------------------------------------------------------------------------------------------------
template <int NM>
static int launch_cta(const TrainParams<float> &p, const CtaGeom &g, cudaStream_t st)
{
    const size_t smem = sizeof(CtaSmem<NM>) + (size_t)p.nsym_smem * sizeof(float2);
    // set on every launch: the attribute belongs to the device that is current (cheap; no per-process flag to go stale)
    QB_CUDA_CHECK(cudaFuncSetAttribute(train_cta_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    train_cta_kernel<NM><<<(unsigned)p.nstreams, 32 * CT_NWARPS, smem, st>>>(p, g);
    count_launch();
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

// QB_CTA_PROF=1 (debugging): every launch is followed by a synchronise and a dump of CTA 0's wait clocks to stderr
static inline bool cta_prof()
{
    static const bool v = [] {
        const char *e = getenv("QB_CTA_PROF");
        return e && e[0] == '1';
    }();
    return v;
}

// shapes and methods the kernel holds; fills the geometry
static inline bool cta_geometry(const TrainParams<float> &p, CtaGeom &g)
{
    if (p.os != 2 || p.ntaps < 3 || p.ntaps > 64) return false;
    if (p.TrSyms > (1LL << 29)) return false;               // 32-bit symbol indices inside the kernel
    g.npair = (p.ntaps + 1) / 2;
    g.tp = p.nmodes * g.npair;
    if (g.tp > CT_MAXTP || g.npair < 2 || g.npair > 32) return false;
    switch (p.method) {
    case QB_CMA: case QB_SGNCMA: case QB_MCMA: break;
    case QB_RDE: case QB_MRDE:
        if ((p.K + 1) / 2 > MAXC) return false;
        break;
    default: return false;
    }
    g.pb = (g.npair - 1 + 7) / 8;
    g.nblk = (int)((p.TrSyms + CT_B - 1) / CT_B);
    const long long s_last = 8LL * (g.nblk + 3) + 7 + g.npair - 1;
    g.nchunks = (int)(((s_last + CT_PRE) >> 6) + 1);
    g.aligned16 = ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0) && (p.seg_stride % 2 == 0) && (p.row_stride % 2 == 0);
    g.err_off = 0;
    g.prof = nullptr;
    return true;
}

template <int NM>
static int train_cta_nm(TrainParams<float> p, cudaStream_t st)
{
    CtaGeom g;
    if (p.nmodes != NM || !cta_geometry(p, g)) return 0;
    if (p.TrSyms <= 0) return 1;
    p.nsym_smem = p.K;
    if (cta_prof()) {
        QB_CUDA_CHECK(cudaMalloc(&g.prof, CT_NWARPS * 8 * sizeof(long long)));
        QB_CUDA_CHECK(cudaMemset(g.prof, 0, CT_NWARPS * 8 * sizeof(long long)));
    }
    const int niter = p.Niter;
    for (int it = 0; it < niter; it++) {      // one launch per training iteration: taps and step size carry over in HBM
        g.err_off = (long long)it * p.TrSyms;
        const int rc = launch_cta<NM>(p, g, st);
        if (rc != QB_OK) return rc;
    }
    if (g.prof) {
        long long h[CT_NWARPS * 8];
        QB_CUDA_CHECK(cudaStreamSynchronize(st));
        QB_CUDA_CHECK(cudaMemcpy(h, g.prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(g.prof);
        static const char *names[CT_NWARPS] = {"CHAIN", "GA0", "GB", "BASE0", "LD", "UPD", "GA1", "BASE1"};
        for (int w = 0; w < CT_NWARPS; w++)
            fprintf(stderr, "[cta prof] %-6s clocks per block %7.0f   waits: %6.0f %6.0f %6.0f samples %6.0f\n", names[w],
                    (double)h[w * 8] / g.nblk, (double)h[w * 8 + 1] / g.nblk, (double)h[w * 8 + 2] / g.nblk,
                    (double)h[w * 8 + 3] / g.nblk, (double)h[w * 8 + 4] / g.nblk);
    }
    return 1;
}

}  // namespace qb
