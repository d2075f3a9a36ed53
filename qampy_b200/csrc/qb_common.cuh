// Shared device/host helpers for the qampy_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qampy_b200.h"

namespace qb {

// ---- error plumbing (host) -----------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);

// ---- kernel-selection overrides (tests, tuning) -----------------------------------------------
// Values are set through qb_set_option (include/qampy_b200.h); the environment (QB_<NAME>) is read ONCE, when the first
// option is looked up, never per launch.  option_char: first character of the value, 0 if unset; option_int: the
// value as an integer, `unset` if unset.
enum Option { OPT_TRAIN_KERNEL, OPT_TRAIN_LPS, OPT_TRAIN_GLA, OPT_LA_TILE, OPT_BPS_KERNEL, OPT_BPS_SPLIT, OPT_COUNT };
char option_char(Option o);
int option_int(Option o, int unset);

#define QB_CUDA_CHECK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return qb::set_error(QB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define QB_REQUIRE(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond))                                                                         \
            return qb::set_error(QB_ERR_ARG, __VA_ARGS__);                                   \
    } while (0)

// ---- complex scalar type -------------------------------------------------------------------
template <typename T>
struct cx_t;
template <>
struct cx_t<float> {
    using type = float2;
};
template <>
struct cx_t<double> {
    using type = double2;
};
template <typename T>
using cx = typename cx_t<T>::type;

template <typename T>
__host__ __device__ __forceinline__ cx<T> make_cx(T re, T im)
{
    cx<T> r;
    r.x = re;
    r.y = im;
    return r;
}

// Unfused IEEE ops: the BPS distance contract forbids FMA contraction.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template <typename T>
__device__ __forceinline__ T shfl_xor(T v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}
template <typename T>
__device__ __forceinline__ T shfl_up(T v, int d)
{
    return __shfl_up_sync(0xffffffffu, v, d);
}
template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int l)
{
    return __shfl_sync(0xffffffffu, v, l);
}

// ---- packed fp32 (Blackwell FMUL2 / FADD2 / FFMA2) -------------------------------------------------
// Pairs are carried as 64-bit values so that ptxas keeps them in aligned register pairs.  One packed
// instruction does two IEEE operations in one issue slot (the FP32 pipe still spends two cycles on it).
// NOTE: ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 although both carry .rn --
// where the unfused result matters (BPS), the addition after a packed product must be a scalar add.rn.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// same, but opaque to ptxas (it otherwise re-derives / duplicates a cheap pair at every use)
__device__ __forceinline__ f32x2 pack2_opaque(float lo, float hi)
{
    f32x2 r;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// (s*v.x, s*v.y), each product rounded once (FMUL2 with a broadcast scalar operand)
__device__ __forceinline__ float2 mul2_bcast(float s, f32x2 v)
{
    f32x2 r;
    asm("{.reg .b64 ra;\n\t"
        "mov.b64 ra, {%1, %1};\n\t"
        "mul.rn.f32x2 %0, ra, %2;}"
        : "=l"(r)
        : "f"(s), "l"(v));
    return unpack2(r);
}
// (s+v.x, s+v.y)  (FADD2).  Never feed it a packed product where the unfused sum is required.
__device__ __forceinline__ float2 add2_bcast(float s, f32x2 v)
{
    f32x2 r;
    asm("{.reg .b64 ra;\n\t"
        "mov.b64 ra, {%1, %1};\n\t"
        "add.rn.f32x2 %0, ra, %2;}"
        : "=l"(r)
        : "f"(s), "l"(v));
    return unpack2(r);
}
// (a.x*a.x, a.y*a.y)
__device__ __forceinline__ float2 sqr2(float2 a)
{
    f32x2 r;
    const f32x2 v = pack2(a.x, a.y);
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(v));
    return unpack2(r);
}
// acc + s*v: two fused multiply-adds in one FFMA2 (scalar s broadcast)
__device__ __forceinline__ f32x2 fma2_bcast(float s, f32x2 v, f32x2 acc)
{
    f32x2 r;
    asm("{.reg .b64 ra;\n\t"
        "mov.b64 ra, {%1, %1};\n\t"
        "fma.rn.f32x2 %0, ra, %2, %3;}"
        : "=l"(r)
        : "f"(s), "l"(v), "l"(acc));
    return r;
}
// element-wise a*b + c (FFMA2)
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- async copy (LDGSTS) and TMA bulk copy (UBLKCP) + mbarrier -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem)
{
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
                 "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA 1-D bulk copy global -> shared::cta, completion on an mbarrier.  Needs 16-byte aligned
// source, destination and size.
__device__ __forceinline__ void tma_bulk_g2s(void *smem, const void *gmem, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(smem)),
        "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct ModeList {
    int n;
    int m[QB_MAX_MODES];
};

}  // namespace qb
