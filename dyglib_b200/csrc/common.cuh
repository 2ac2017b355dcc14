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

// cos of an fp32 argument of any magnitude (time deltas reach 1e8 rad), without the local-memory Payne-Hanek slow path
// of cosf(): the quadrant count and a two-term Cody-Waite reduction by pi/2 run in float64 (exact to ~1e-16 rad for
// |x| < 2^31), then one degree-3 minimax polynomial in r^2 (the cephes sinf / cosf kernels on |r| <= pi/4, coefficients
// selected by the quadrant's parity).  7 FP64 + ~14 FP32 instructions, branch-free; |error| < 1.2e-7 against the
// float64 cosine of the same fp32 argument.  The fp32 rounding of the *argument* is part of the reference's semantics
// (SURVEY.md 7.3(3)); only the evaluation of cos differs from libm.
__device__ __forceinline__ float dyg_cosf(float x) {
    if (fabsf(x) > 2.0e9f) return cosf(x);                       // quadrant count would overflow int32 (never hit by time deltas)
    const double xd = (double)x;
    // quadrant count by the magic-number trick: (x 2/pi + 1.5 2^52) has q = rint(x 2/pi) in its low mantissa bits (|q| < 2^31), so
    // neither FRND.F64 nor F2I.F64 is needed -- 64-bit conversions run at 16 lanes per clock and SM, eight cycles per warp instruction
    const double qs = fma(xd, 0.63661977236758134308, 6755399441055744.0);   // 2 / pi, 1.5 * 2^52
    const double q = qs - 6755399441055744.0;
    double r = fma(-q, 1.57079632679489655800e+00, xd);          // pi / 2, high part
    r = fma(-q, 6.12323399573676603587e-17, r);                  // pi / 2, low part
    const float rf = (float)r;                                   // |rf| <= pi / 4
    const int n = __double2loint(qs);
    const bool odd = n & 1;                                      // odd quadrants evaluate sin(r), even ones cos(r)
    const float r2 = rf * rf;
    float p = fmaf(odd ? -1.9515295891e-4f : 2.443315711809948e-5f, r2, odd ? 8.3321608736e-3f : -1.388731625493765e-3f);
    p = fmaf(p, r2, odd ? -1.6666654611e-1f : 4.166664568298827e-2f);
    const float res = (odd ? rf : 1.0f) + fmaf(odd ? rf * r2 : r2 * r2, p, odd ? 0.f : -0.5f * r2);
    return ((n + 1) & 2) ? -res : res;                           // cos(r + n pi/2): +cos, -sin, -cos, +sin
}

// sin and cos of the same fp32 argument with the reduction of dyg_cosf (used by the time encoder's backward:
// d cos(x) / dx = -sin(x) at the fp32 argument the forward pass used).
__device__ __forceinline__ void dyg_sincosf(float x, float* sn, float* cs) {
    if (fabsf(x) > 2.0e9f) { sincosf(x, sn, cs); return; }
    const double xd = (double)x;
    const double qs = fma(xd, 0.63661977236758134308, 6755399441055744.0);
    const double q = qs - 6755399441055744.0;
    double r = fma(-q, 1.57079632679489655800e+00, xd);
    r = fma(-q, 6.12323399573676603587e-17, r);
    const float rf = (float)r;
    const int n = __double2loint(qs);
    const float r2 = rf * rf;
    float ps = fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    const float sr = rf + rf * r2 * ps;                                   // sin(r)
    float pc = fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    const float cr = 1.0f + fmaf(r2 * r2, pc, -0.5f * r2);                // cos(r)
    // x = r + n pi/2:  n mod 4 = 0: (s, c) = (sr, cr); 1: (cr, -sr); 2: (-sr, -cr); 3: (-cr, sr)
    const float s0 = (n & 1) ? cr : sr, c0 = (n & 1) ? sr : cr;
    *sn = (n & 2) ? -s0 : s0;
    *cs = ((n + 1) & 2) ? -c0 : c0;
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
