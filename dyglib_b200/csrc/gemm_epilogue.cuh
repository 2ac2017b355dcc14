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
            for (int j = 0; j < 4; ++j) bq[j] = reinterpret_cast<const float4*>(g.bias + n)[j];   // generic loads: the bias may sit in shared memory
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
                        for (int j = 0; j < 16; ++j) v[j] += g.bias[n + j];
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
                    if (g.bias) x += g.bias[n + j];
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

// planes-only output (bias, no residual / activation / fp32 copy, every chunk whole): eligible for the TMA-store epilogue below
__host__ __device__ __forceinline__ bool planes_fast_path(const GemmArgs& g) {
    return g.Chi && !g.C && !g.residual && g.bias && g.act == DYG_ACT_NONE && (g.N & 15) == 0 && ((g.ldcs & 15) == 0) &&
           (((reinterpret_cast<uintptr_t>(g.Chi) | reinterpret_cast<uintptr_t>(g.Cmid)) & 31u) == 0) &&
           ((reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0);
}
__device__ __forceinline__ void pin16(uint32_t (&r)[16]) {
    // orders every use of r after the tcgen05.wait::ld that precedes this statement
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
// one 16-column chunk of a row -> hi | mid planes with two 256-bit stores
__device__ __forceinline__ void planes_chunk(const GemmArgs& g, const uint32_t (&r)[16], const float4 (&bq)[4], int64_t m, int n, bool rowok) {
    if (!rowok) return;
    uint32_t hi[8], mid[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        split_pack(__uint_as_float(r[4 * j]) + bq[j].x, __uint_as_float(r[4 * j + 1]) + bq[j].y, hi[2 * j], mid[2 * j]);
        split_pack(__uint_as_float(r[4 * j + 2]) + bq[j].z, __uint_as_float(r[4 * j + 3]) + bq[j].w, hi[2 * j + 1], mid[2 * j + 1]);
    }
    st_v8(g.Chi + m * g.ldcs + n, hi);
    st_v8(g.Cmid + m * g.ldcs + n, mid);
}
// Planes epilogue through shared memory + TMA stores.  A row-strided st.global costs the LSU one wavefront per 32-byte sector
// (32 per 256-bit store instruction): globaltimer stamps put the plane stores of ln_gemm at a third of its epilogue time, and
// the LN warps' own loads / stores queue behind them.  Here a warp converts a unit of 32 columns (two TMEM chunks) of its 32
// rows, writes the hi | mid images (64-byte rows, SWIZZLE_64B: conflict-free 16-byte stores) into its 4 KB staging slot and
// one lane hands two 32 x 32 boxes to the TMA engine, which also clips the rows past M.  A trailing 16-column unit of the
// tile goes out with direct stores.  `stage` must be 1 KB aligned; the slot is reused once the previous boxes have been read.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, const uint32_t* r) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
template <typename Done>
__device__ __forceinline__ void epilogue_planes_tma_tile(const GemmArgs& g, const float* bias_s, const CUtensorMap* map_h, const CUtensorMap* map_m,
                                                         unsigned char* stage, uint32_t taddr, int lane, int64_t m, int n0, int ncols, int unit0,
                                                         int step, Done tmem_done) {
    const bool rowok = m < g.M;
    const int m_first = (int)(m - lane);
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
    const uint32_t row_h = smem_u32(stage) + (uint32_t)lane * 64u, row_m = row_h + 2048u;
    const uint32_t bias_a = smem_u32(bias_s);
    int col = 32 * unit0;
    if (col >= ncols) {
        tmem_done();
        return;
    }
    while (col < ncols) {
        const int next = col + 32 * step;
        const bool two = col + 32 <= ncols;
        uint32_t ra[16], rb[16];
        float4 ba[4], bb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ba[j] = lds_f4(bias_a + (uint32_t)(n0 + col) * 4u + 16u * j);
        tmem_ld16_nowait(taddr + (uint32_t)col, ra);
        if (two) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = lds_f4(bias_a + (uint32_t)(n0 + col + 16) * 4u + 16u * j);
            tmem_ld16_nowait(taddr + (uint32_t)(col + 16), rb);
        }
        tmem_ld_wait();
        pin16(ra);
        if (next >= ncols) tmem_done();
        if (!two) {
            planes_chunk(g, ra, ba, m, n0 + col, rowok);
            break;
        }
        pin16(rb);
        uint32_t hi[16], mid[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            split_pack(__uint_as_float(ra[4 * j]) + ba[j].x, __uint_as_float(ra[4 * j + 1]) + ba[j].y, hi[2 * j], mid[2 * j]);
            split_pack(__uint_as_float(ra[4 * j + 2]) + ba[j].z, __uint_as_float(ra[4 * j + 3]) + ba[j].w, hi[2 * j + 1], mid[2 * j + 1]);
            split_pack(__uint_as_float(rb[4 * j]) + bb[j].x, __uint_as_float(rb[4 * j + 1]) + bb[j].y, hi[8 + 2 * j], mid[8 + 2 * j]);
            split_pack(__uint_as_float(rb[4 * j + 2]) + bb[j].z, __uint_as_float(rb[4 * j + 3]) + bb[j].w, hi[8 + 2 * j + 1], mid[8 + 2 * j + 1]);
        }
        // the boxes of the previous unit must have left the staging slot
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t off = ((uint32_t)u ^ sw) << 4;
            sts_v4(row_h + off, hi + 4 * u);
            sts_v4(row_m + off, mid + 4 * u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m_first < g.M) {
            tma_store_2d(map_h, stage, n0 + col, m_first);
            tma_store_2d(map_m, stage + 2048, n0 + col, m_first);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        col = next;
    }
}

}  // namespace
