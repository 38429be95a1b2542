// Shared helpers for libsydr_b200: error reporting, IQ sample decoding, sm_100a PTX wrappers
// (mbarrier, cp.async.bulk, st.async, cluster addressing).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sydr_b200.h"

namespace sydr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SYDR_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            sydr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                            cudaGetErrorString(e__));                                      \
            return SYDR_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define SYDR_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            sydr::set_error(__VA_ARGS__);                                                  \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

constexpr int kCodeChips = 1023;
constexpr int kPaddedChips = 1025;           // [c1022, c0..c1022, c0]  channel_l1ca_borre.py:173
constexpr int kCodeWords = 36;               // 1025 bits -> 33 words, padded to 36 (144 B)
constexpr int kMaxPrn = 37;
constexpr double kPi = 3.141592653589793;    // np.pi
constexpr double kGpsPi = 3.1415926535898;   // sydr/utils/constants.py:4
constexpr double kCodeFreq = 1.023e6;

// Device-resident code tables, built once per device by ensure_code_tables().
struct CodeTables {
    const uint32_t* padded_bits;   // [kMaxPrn][kCodeWords], bit k = (padded_code[k] > 0)
    const int8_t* chips;           // [kMaxPrn][1023], +-1
};
int ensure_code_tables(CodeTables* out);

// ------------------------------------------------------------------------------------------
// IQ decoding.  One 16-byte vector holds SPV complex samples.
// ------------------------------------------------------------------------------------------
template <int DT> struct IqTraits;
template <> struct IqTraits<SYDR_IQ_I8>  { static constexpr int BPS = 2, SPV = 8; };
template <> struct IqTraits<SYDR_IQ_I16> { static constexpr int BPS = 4, SPV = 4; };
template <> struct IqTraits<SYDR_IQ_F32> { static constexpr int BPS = 8, SPV = 2; };

template <int DT>
__device__ __forceinline__ void decode_vec(const uint4& v, float* re, float* im);

// Integer -> float without the conversion pipe (I2F issues once per ~8 clk per scheduler on
// sm_100a, profiles/r1_ubench_pipe_rates.txt): bias the integer to unsigned with one XOR per
// word, splice it into the mantissa of 1.5 * 2^23 with PRMT and subtract the magic constant.
// Exact for every int8 / int16 value.
template <>
__device__ __forceinline__ void decode_vec<SYDR_IQ_I8>(const uint4& v, float* re, float* im) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const float magic = 12582912.f + 128.f;                       // 0x4B400000 + bias
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w[k] ^ 0x80808080u;
        re[2 * k]     = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7640)) - magic;
        im[2 * k]     = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7641)) - magic;
        re[2 * k + 1] = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7642)) - magic;
        im[2 * k + 1] = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7643)) - magic;
    }
}
template <>
__device__ __forceinline__ void decode_vec<SYDR_IQ_I16>(const uint4& v, float* re, float* im) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const float magic = 12582912.f + 32768.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w[k] ^ 0x80008000u;
        re[k] = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7610)) - magic;
        im[k] = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7632)) - magic;
    }
}
template <>
__device__ __forceinline__ void decode_vec<SYDR_IQ_F32>(const uint4& v, float* re, float* im) {
    re[0] = __uint_as_float(v.x); im[0] = __uint_as_float(v.y);
    re[1] = __uint_as_float(v.z); im[1] = __uint_as_float(v.w);
}

// Scalar sample fetch (any alignment), used by the acquisition front end.
__device__ __forceinline__ float2 load_sample(const void* p, int dt, long long i) {
    if (dt == SYDR_IQ_I8) {
        const char2 c = reinterpret_cast<const char2*>(p)[i];
        return make_float2((float)c.x, (float)c.y);
    } else if (dt == SYDR_IQ_I16) {
        const short2 c = reinterpret_cast<const short2*>(p)[i];
        return make_float2((float)c.x, (float)c.y);
    } else if (dt == SYDR_IQ_F32) {
        return reinterpret_cast<const float2*>(p)[i];
    } else {
        const double2 c = reinterpret_cast<const double2*>(p)[i];
        return make_float2((float)c.x, (float)c.y);
    }
}

inline int iq_bytes_per_sample(int dt) {
    switch (dt) {
        case SYDR_IQ_I8: return 2;
        case SYDR_IQ_I16: return 4;
        case SYDR_IQ_F32: return 8;
        case SYDR_IQ_F64: return 16;
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------------
// PTX wrappers (sm_90+/sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Wait with cluster-scope acquire: data written by remote st.async is visible afterwards.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// TMA bulk copy global -> this CTA's shared memory, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// The same with shared-memory addresses already in the 32-bit shared window (hot loops).
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& a, uint32_t& b) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                     : "memory");
}
// Map a CTA-local shared address to the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Remote (DSMEM) 16-byte store that also signals complete_tx(16) on the destination CTA's mbarrier.
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster_addr, uint32_t bar_cluster_addr,
                                            float a, float b, float c, float d) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
            "r"(dst_cluster_addr),
        "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)),
        "r"(__float_as_uint(d)), "r"(bar_cluster_addr)
        : "memory");
}

__device__ __forceinline__ void st_async_v4u(uint32_t dst_cluster_addr, uint32_t bar_cluster_addr, uint32_t a, uint32_t b,
                                             uint32_t c, uint32_t d) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
            "r"(dst_cluster_addr),
        "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar_cluster_addr)
        : "memory");
}

// Remote (DSMEM) 4-byte store, complete_tx(4) on the destination CTA's mbarrier.
__device__ __forceinline__ void st_async_b32(uint32_t dst_cluster_addr, uint32_t bar_cluster_addr, float a) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster_addr),
                 "r"(__float_as_uint(a)), "r"(bar_cluster_addr)
                 : "memory");
}

__device__ __forceinline__ void st_async_b64(uint32_t dst_cluster_addr, uint32_t bar_cluster_addr, double a) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster_addr),
                 "l"(__double_as_longlong(a)), "r"(bar_cluster_addr)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sydr
