// Shared helpers for the sm_100a kernels behind include/dygb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dygb200.h"

void dyg_set_error(const char* fmt, ...);

#define DYG_CHECK_ARG(cond, ...)             \
    do {                                     \
        if (!(cond)) {                       \
            dyg_set_error(__VA_ARGS__);      \
            return 2;                        \
        }                                    \
    } while (0)

#define DYG_LAUNCH_CHECK(name)                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            dyg_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static inline cudaStream_t as_stream(dyg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int dyg_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cos of an fp32 argument of any magnitude, without the local-memory Payne-Hanek slow path of cosf():
// two-term Cody-Waite reduction by 2*pi in float64 (exact to ~1e-15 rad for |x| < 2^40), then the
// fast path of cosf on |r| <= pi.  The fp32 rounding of the *argument* is part of the reference's
// semantics; only the evaluation of cos differs from libm, by < 3e-7 absolute.
__device__ __forceinline__ float dyg_cosf(float x) {
    const double two_pi_hi = 6.283185307179586232e+00;
    const double two_pi_lo = 2.449293598294706414e-16;
    const double inv_two_pi = 1.591549430918953456e-01;
    double xd = (double)x;
    double q = rint(xd * inv_two_pi);
    double r = fma(-q, two_pi_hi, xd);
    r = fma(-q, two_pi_lo, r);
    return cosf((float)r);
}

// TimeEncoder element (models/modules.py:37): nn.Linear(1,T) on CPU is a single fp32 FMA, then cos.
__device__ __forceinline__ float dyg_time_enc(float dt, float w, float b) { return dyg_cosf(fmaf(dt, w, b)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 16-byte load that does not allocate in L1 (gathered rows are read once per kernel)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
