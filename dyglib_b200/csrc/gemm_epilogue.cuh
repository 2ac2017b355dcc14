// Accumulator-tile epilogue shared by the tcgen05 GEMM kernels (gemm_tc.cu, attn_block.cu): bias / residual / activation and
// 256-bit stores of the fp32 result and / or its bf16 hi | mid operand planes.
#pragma once
#include <math.h>

#include "tc_common.cuh"

namespace {

struct GemmArgs {
    const float* bias;
    const float* residual;
    float* C;
    __nv_bfloat16* Chi;
    __nv_bfloat16* Cmid;
    int64_t M;
    int64_t m_tiles;
    int ldr, ldc, ldcs;
    int N, K, act;
    int NS;        // columns owned by one n tile (multiple of 16)
    int NT;        // MMA N of one n tile (== NS)
    int n_tiles;
    int stages;
    int resident;  // 1: A super tile stays in shared memory across the n tiles (K <= 224)
    int64_t m_super;   // 256-row super tiles (one per CTA pair and round)
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == DYG_ACT_RELU) return fmaxf(v, 0.f);
    if (act == DYG_ACT_GELU) {
        // exact-erf GELU (F.gelu default, models/DyGFormer.py:458) as max(v, 0) - 0.5 |v| erfc(|v| / sqrt 2) with
        // erfc(z) = 2^(-z Q(z)), Q a degree-6 minimax fit on [0, 6] (|erfc error| <= 1.2e-7 in fp32 arithmetic; GELU within
        // 4e-7 absolute, below the BF16x3 noise of the contraction feeding it): branch-free, ONE MUFU + 11 FP32 instructions.
        // (erff: ~25 divergent instructions; the Abramowitz-Stegun 7.1.26 form used before: two MUFU + 15.  The fused FFN's
        // epilogue warps are bound by the MUFU / conversion pipe, profiles/README.md.)
        const float az = fminf(fabsf(v) * 0.70710678118654752440f, 6.0f);
        float q = fmaf(-1.0021893831435591e-4f, az, 4.6156023745425045e-4f);
        q = fmaf(q, az, 2.302266424521804e-3f);
        q = fmaf(q, az, -2.9452543705701828e-2f);
        q = fmaf(q, az, 1.4896368980407715e-1f);
        q = fmaf(q, az, 9.183286428451538e-1f);
        q = fmaf(q, az, 1.6279137134552002f);
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-az * q));
        return fmaf(-0.5f * fabsf(v), e, fmaxf(v, 0.f));
    }
    if (act == DYG_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}
// One accumulator tile -> global memory: this thread owns TMEM lane `row` (output row m), this warp the 16-column
// chunks chunk0, chunk0 + step, ...  Warp-collective (tcgen05.ld): every lane runs the loop, stores are masked.
__device__ __forceinline__ void epilogue_tile(const GemmArgs& g, uint32_t taddr, int64_t m, int n0, int ncols, int chunk0, int step) {
    const bool rowok = m < g.M;
    const bool c_v8 = g.C && ((g.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 31u) == 0);
    const bool c_v4 = g.C && ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
    const bool s_v8 = g.Chi && ((g.ldcs & 15) == 0) && (((reinterpret_cast<uintptr_t>(g.Chi) | reinterpret_cast<uintptr_t>(g.Cmid)) & 31u) == 0);
    const bool s_v4 = g.Chi && ((g.ldcs & 7) == 0) && (((reinterpret_cast<uintptr_t>(g.Chi) | reinterpret_cast<uintptr_t>(g.Cmid)) & 15u) == 0);
    const bool r_v4 = g.residual && ((g.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.residual) & 15u) == 0);
    const bool r_v8 = g.residual && ((g.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(g.residual) & 31u) == 0);
    const bool b_v4 = g.bias && ((reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0);
    for (int col = 16 * chunk0; col < ncols; col += 16 * step) {
        const int n = n0 + col;
        const int valid = min(16, ncols - col);
        // the bias of this chunk does not depend on the accumulator: request it before the TMEM load so that its latency
        // hides behind tcgen05.ld + wait (27 % of the QKV GEMM's stall samples sat on the first add after the wait, profiles/)
        float4 bq[4];
        const bool bias_q = g.bias && b_v4 && valid == 16;
        if (bias_q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(g.bias + n) + j);
        }
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)col, r);
        if (rowok) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (valid == 16) {
                if (g.bias) {
                    if (b_v4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[4 * j] += bq[j].x; v[4 * j + 1] += bq[j].y; v[4 * j + 2] += bq[j].z; v[4 * j + 3] += bq[j].w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += __ldg(g.bias + n + j);
                    }
                }
                if (g.residual) {
                    const float* rp = g.residual + m * g.ldr + n;
                    if (r_v8) {
                        // one full 32-byte sector per lane and instruction: row-strided accesses cost one L1 wavefront
                        // per sector, so 256-bit loads halve the epilogue's load wavefronts
                        float q[16];
                        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(q[3]), "=f"(q[4]), "=f"(q[5]), "=f"(q[6]), "=f"(q[7])
                                     : "l"(rp));
                        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=f"(q[8]), "=f"(q[9]), "=f"(q[10]), "=f"(q[11]), "=f"(q[12]), "=f"(q[13]), "=f"(q[14]), "=f"(q[15])
                                     : "l"(rp + 8));
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += q[j];
                    } else if (r_v4) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += rp[j];
                    }
                }
                if (g.act != DYG_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = act_apply(v[j], g.act);
                }
                if (g.C) {
                    float* dst = g.C + m * g.ldc + n;
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(v[j]);
                    if (c_v8) {
                        st_v8(dst, o);
                        st_v8(dst + 8, o + 8);
                    } else if (c_v4) {
                        st_v4(dst, o); st_v4(dst + 4, o + 4); st_v4(dst + 8, o + 8); st_v4(dst + 12, o + 12);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) dst[j] = v[j];
                    }
                }
                if (g.Chi) {
                    uint32_t hi[8], mid[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split_pack(v[2 * j], v[2 * j + 1], hi[j], mid[j]);
                    __nv_bfloat16* dh = g.Chi + m * g.ldcs + n;
                    __nv_bfloat16* dm = g.Cmid + m * g.ldcs + n;
                    if (s_v8) {
                        st_v8(dh, hi);
                        st_v8(dm, mid);
                    } else if (s_v4) {
                        st_v4(dh, hi); st_v4(dh + 8, hi + 4);
                        st_v4(dm, mid); st_v4(dm + 8, mid + 4);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            *reinterpret_cast<uint32_t*>(dh + 2 * j) = hi[j];
                            *reinterpret_cast<uint32_t*>(dm + 2 * j) = mid[j];
                        }
                    }
                }
            } else {
                // ragged last chunk of the row (N not a multiple of 16)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j >= valid) break;
                    float x = v[j];
                    if (g.bias) x += __ldg(g.bias + n + j);
                    if (g.residual) x += g.residual[m * g.ldr + n + j];
                    x = act_apply(x, g.act);
                    if (g.C) g.C[m * g.ldc + n + j] = x;
                    if (g.Chi) {
                        const __nv_bfloat16 h = __float2bfloat16_rn(x);
                        g.Chi[m * g.ldcs + n + j] = h;
                        g.Cmid[m * g.ldcs + n + j] = __float2bfloat16_rn(x - __bfloat162float(h));
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace
