// Shared device/host helpers for the b200ssl kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#ifndef B200_API
#define B200_API extern "C" __attribute__((visibility("default")))
#endif

// ---------------------------------------------------------------- errors
enum {
    B200_OK = 0,
    B200_ERR_ARG = -1,      // bad pointer / shape / unsupported combination
    B200_ERR_CUDA = -2,     // a CUDA runtime call failed (message has the cudaError string)
    B200_ERR_WORKSPACE = -3 // caller workspace too small
};

void b200_set_error(const char* fmt, ...);

#define B200_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            b200_set_error(__VA_ARGS__);                                               \
            return B200_ERR_ARG;                                                       \
        }                                                                              \
    } while (0)

#define B200_CHECK_LAUNCH(name)                                                        \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            b200_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return B200_ERR_CUDA;                                                      \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------- fast division
// Unsigned division by a runtime constant via multiply-high (Granlund/Montgomery); valid for
// dividends < 2^31.
struct FastDiv {
    uint32_t d, mul, shr;
    __host__ void init(uint32_t div) {
        d = div ? div : 1;
        if (d == 1) { mul = 0; shr = 0; return; }
        uint32_t l = 0;
        while ((1u << l) < d) ++l;           // ceil(log2 d)
        uint64_t m = ((uint64_t(1) << 32) * ((uint64_t(1) << l) - d)) / d + 1;
        mul = (uint32_t)m;
        shr = l;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        if (d == 1) return n;
        uint32_t t = __umulhi(n, mul);
        return (t + ((n - t) >> 1)) >> (shr - 1);
    }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG: the same (seed, stream, counter) always gives the same 4 words, so dropout
// masks and noise are recomputed in backward instead of being stored.
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint32_t stream, uint64_t counter) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = stream, c3 = 0x5151B200u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

__host__ __device__ __forceinline__ float u32_to_unit(uint32_t v) {   // [0,1) with 24 bits
    return (float)(v >> 8) * (1.0f / 16777216.0f);
}

// keep-mask of 4 consecutive elements (element index 4*q .. 4*q+3) of dropout stream `stream`
__device__ __forceinline__ void dropout_keep4(uint64_t seed, uint32_t stream, uint64_t q, float p, bool (&keep)[4]) {
    Philox4 r = philox4x32_10(seed, stream, q);
    keep[0] = u32_to_unit(r.x) >= p;
    keep[1] = u32_to_unit(r.y) >= p;
    keep[2] = u32_to_unit(r.z) >= p;
    keep[3] = u32_to_unit(r.w) >= p;
}

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- streaming loads / stores
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

static inline int b200_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
