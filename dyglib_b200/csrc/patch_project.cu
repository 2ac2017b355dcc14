// dyg_patch_project: DyGFormer's feature gathering, time encoding, patching and the four channel projections
// (models/DyGFormer.py:247-306 get_features / get_patches, :148-174 projection_layer + stacking) as ONE kernel per batch.
//
// Token m of a side (src or dst sequences) covers P consecutive padded positions q = m*P + p.  Its row of the
// (B, S, 4C) transformer input is
//     X[row(m), ch*C:(ch+1)*C] = bias_ch + sum_p W_ch[:, p*F_ch:(p+1)*F_ch] feat_ch(q)
// with feat = node row, edge row, cos(fma(t_query - t_nbr, w, b)) (zero at padded ids) and the co-occurrence MLP rows
// LUT[count in src row] + LUT[count in dst row].  K is walked in "stages" of 32 columns, one (channel, p, 32-column
// block) at a time; each channel accumulates into its own 64-column TMEM accumulator (block-diagonal contraction).
//
// Gathers are pure data movement: the node / edge feature tables and the LUT are kept in HBM as BF16x3 operand planes
// (hi = bf16(x), mid = bf16(x - hi), the same 4 bytes per element as fp32), so 8 producer warps only compute row
// addresses and fire 16-byte cp.async copies straight into the SWIZZLE_64B operand tile; completion is counted on the
// stage's mbarrier (cp.async.mbarrier.arrive.noinc), so producers run several stages ahead of the MMAs.  Only the time
// encoding is computed (accurate cos on the fp32 FMA argument).  The packed weights arrive by TMA.
//   warps 0-7  producers, then epilogue (tcgen05.ld -> + bias -> X)
//   warp 8     TMA of the packed weight block of each stage
//   warp 9     TMEM allocation + single-thread tcgen05.mma issue (M=128, N=64, K=16, BF16x3 = three MMAs per k step)
// Two CTAs per SM (96 KB of shared memory, 256 TMEM columns each): one CTA's epilogue overlaps the other's gathers.
#include <math.h>
#include <string.h>

#include "tc_common.cuh"

namespace {

constexpr int PP_BM = 128;
constexpr int PP_PRODUCERS = 256;
constexpr int PP_THREADS = 320;
constexpr int PP_STAGES = 4;
constexpr int PP_A_PLANE = PP_BM * 64;    // 8 KB: 128 rows x 32 bf16
constexpr int PP_NT = 64;                 // accumulator columns per channel (C <= 64)
constexpr int PP_W_PLANE = PP_NT * 64;    // 4 KB
constexpr int PP_STAGE_BYTES = 2 * PP_A_PLANE + 2 * PP_W_PLANE;
constexpr int PP_TMEM_COLS = 4 * PP_NT;
constexpr int PP_MAX_STAGES = 512;        // stage table in shared memory (patch_size 16 needs 320)

struct ProjArgs {
    dyg_proj_side_t side[2];
    const __nv_bfloat16* tab_hi[5];   // per stage type: node, edge, (unused), lut, lut
    const __nv_bfloat16* tab_mid[5];
    int ld[5];
    int nblk[5];                      // 32-column blocks per (type, p)
    int w16[5];                       // valid width rounded up to 16
    int base[6];                      // first stage of each type; base[5] = number of stages
    const double* t_query;
    const float* tw;
    const float* tb;
    int T, P, C, S;
    const float* bias;
    float* X;
    int ldx;
    int64_t tiles0;                   // tiles of side 0 (side 1 tiles follow)
    int zero_rows;                    // bit u: row 0 of table u is all zero (padding id): skip its gather
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cta(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_cta(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct StageInfo {     // 16 bytes: one LDS.128 per stage
    int ty, p, blk, nk16;
};
__device__ __forceinline__ StageInfo decode_stage(const ProjArgs& a, int s) {
    StageInfo si;
    int ty = 0;
#pragma unroll
    for (int u = 1; u < 5; ++u)
        if (s >= a.base[u]) ty = u;
    const int local = s - a.base[ty];
    si.ty = ty;
    si.p = local / a.nblk[ty];
    si.blk = local - si.p * a.nblk[ty];
    const int rem16 = (a.w16[ty] - si.blk * 32) >> 4;
    si.nk16 = rem16 >= 2 ? 2 : rem16;
    return si;
}

__global__ void __launch_bounds__(PP_THREADS, 2) patch_project_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                     const __grid_constant__ CUtensorMap map_wm, const ProjArgs a) {
    extern __shared__ __align__(1024) unsigned char pp_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pp_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + PP_STAGES * PP_STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [PP_STAGES]  256 producer arrivals + the weight TMA (expect_tx)
    uint64_t* empty_bar = bars + PP_STAGES;       // [PP_STAGES]  one tcgen05.commit
    uint64_t* acc_bar = bars + 2 * PP_STAGES;     // accumulators complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PP_STAGES + 1);
    StageInfo* stage_tab = reinterpret_cast<StageInfo*>(bars + 2 * PP_STAGES + 2);   // [nst]
    // time-encoder weights / biases: read 32 times per thread and time stage; the gathers allocate in L1 (cp.async.ca) and evict them
    float* tw_s = reinterpret_cast<float*>(stage_tab + a.base[5]);                    // [T] | [T]
    float* tb_s = tw_s + a.T;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform: role branches stay on the uniform path
    const int sd = (int64_t)blockIdx.x >= a.tiles0 ? 1 : 0;
    const dyg_proj_side_t& side = a.side[sd];
    const int64_t m0 = ((int64_t)blockIdx.x - (sd ? a.tiles0 : 0)) * PP_BM;
    const int nst = a.base[5];

    for (int s = tid; s < nst; s += PP_THREADS) stage_tab[s] = decode_stage(a, s);
    for (int c = tid; c < a.T; c += PP_THREADS) {
        tw_s[c] = __ldg(a.tw + c);
        tb_s[c] = __ldg(a.tb + c);
    }
    if (tid == 0) {
        for (int s = 0; s < PP_STAGES; ++s) {
            mbar_init(full_bar + s, PP_PRODUCERS + 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)PP_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 8 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wm)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers
        // Time stages: thread = (tile row r, 16-column half).  Gather stages: four consecutive lanes copy the four 16-byte
        // chunks of one (row, plane) piece, so the 32 lanes of a cp.async instruction touch 8 contiguous 64-byte pieces
        // (8 L1 wavefronts) instead of 32 different rows (32 wavefronts): with one piece per thread the kernel was bound
        // by L1 wavefronts, 2.9 TB/s of L2-resident rows with DRAM at 6 % and L2 at 30 % (profiles/).  Instruction i of
        // warp w covers rows 16 w + 4 i + (lane >> 3), plane (lane >> 2) & 1, chunk lane & 3.
        const int r = tid >> 1;
        const int half = tid & 1;
        const int64_t m = m0 + r;
        const bool valid = m < side.tokens;
        const uint32_t row_off = (uint32_t)(r * 64);
        const uint32_t swz = (uint32_t)((r >> 1) & 3);
        const double tq = valid ? __ldg(a.t_query + m / side.ntok) : 0.0;
        const int g_row0 = 16 * warp + (lane >> 3);      // + 4 i
        const int g_plane = (lane >> 2) & 1, g_chunk = lane & 3;
        int64_t idx4[4] = {0, 0, 0, 0}, nidx4[4] = {0, 0, 0, 0};   // gathered rows of the current / next unit (per instruction)
        float dt = 0.f, ndt = 0.f;                                // time stages: delta of the current / next unit
        bool masked = true, nmasked = true;
        auto load_unit = [&](const StageInfo& u) {
            if (u.ty == 2) {
                nmasked = true;
                if (valid) {
                    const int64_t q = m * a.P + u.p;
                    nmasked = __ldg(side.ids + q) == 0;
                    ndt = (float)(tq - (double)__ldg(side.t_nbr + q));
                }
            } else {
                const int64_t* src = u.ty == 0 ? side.ids : (u.ty == 1 ? side.eids : (u.ty == 3 ? side.cnt_a : side.cnt_b));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t mm = m0 + g_row0 + 4 * i;
                    nidx4[i] = mm < side.tokens ? __ldg(src + mm * a.P + u.p) : 0;
                }
            }
        };
        load_unit(stage_tab[0]);
        for (int s = 0; s < nst; ++s) {
            const int slot = s % PP_STAGES;
            const uint32_t ph = (uint32_t)((s / PP_STAGES) & 1);
            const StageInfo si = stage_tab[s];
            if (si.blk == 0) {
                // the indices (and time) of this unit were requested one unit ago; request the next unit's now, so the
                // dependent load -> address -> cp.async chain never waits for memory
#pragma unroll
                for (int i = 0; i < 4; ++i) idx4[i] = nidx4[i];
                masked = nmasked;
                dt = ndt;
                const int s2 = s + a.nblk[si.ty];                   // first stage of the next unit
                if (s2 < nst) load_unit(stage_tab[s2]);
            }
            mbar_wait(empty_bar + slot, ph ^ 1u);
            unsigned char* st = base + slot * PP_STAGE_BYTES;
            if (si.ty != 2) {
                const __nv_bfloat16* tab = (g_plane ? a.tab_mid[si.ty] : a.tab_hi[si.ty]) + si.blk * 32 + g_chunk * 8;
                const uint32_t dst0 = smem_u32(st + g_plane * PP_A_PLANE);
                const bool zr = (a.zero_rows >> si.ty) & 1;
                const bool live = g_chunk < si.nk16 * 2;
                bool stored = false;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = g_row0 + 4 * i;
                    const uint32_t dst = dst0 + (uint32_t)row * 64u + ((((uint32_t)g_chunk) ^ ((uint32_t)(row >> 1) & 3u)) << 4);
                    if (!live) continue;
                    if (idx4[i] == 0 && zr) {
                        // padded position of a table whose row 0 is zero: no memory traffic (and no L2 hot spot on that row)
                        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
                        stored = true;
                    } else {
                        cp_async16(dst, tab + idx4[i] * a.ld[si.ty]);
                    }
                }
                if (stored) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                cp_async_arrive_noinc(full_bar + slot);
            } else {
                if (half < si.nk16) {
                    const int c0 = si.blk * 32 + half * 16;
                    uint32_t hi[8], mid[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + 2 * j;
                        float v0 = 0.f, v1 = 0.f;
                        if (!masked) {
                            if (c < a.T) v0 = dyg_time_enc(dt, tw_s[c], tb_s[c]);
                            if (c + 1 < a.T) v1 = dyg_time_enc(dt, tw_s[c + 1], tb_s[c + 1]);
                        }
                        split_pack(v0, v1, hi[j], mid[j]);
                    }
                    unsigned char* rowp = st + row_off;
                    const uint32_t c16 = (uint32_t)half * 2u;
                    *reinterpret_cast<uint4*>(rowp + ((c16 ^ swz) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(rowp + (((c16 + 1) ^ swz) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4*>(rowp + PP_A_PLANE + ((c16 ^ swz) << 4)) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
                    *reinterpret_cast<uint4*>(rowp + PP_A_PLANE + (((c16 + 1) ^ swz) << 4)) = make_uint4(mid[4], mid[5], mid[6], mid[7]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05 (async proxy)
                mbar_arrive_cta(full_bar + slot);
            }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        const int er = quarter * 32 + lane;                 // TMEM lane == tile row
        const int64_t em = m0 + er;
        const bool eok = em < side.tokens;
        const int64_t orow = eok ? (em / side.ntok) * a.S + side.tok_off + (em % side.ntok) : 0;
        const bool v2 = ((a.ldx & 1) == 0) && ((a.C & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 7u) == 0);
        // C == 50 (the reference's channel_embedding_dim): the two channels of this warp are 100 contiguous, 16-byte aligned floats of
        // the token's row, written with 25 128-bit stores instead of 50 64-bit ones (every row-strided store instruction costs the LSU
        // 32 wavefronts whatever its width; globaltimer stamps put this epilogue at 21 of the 88 us of a tile)
        const bool v4 = a.C == 50 && ((a.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(a.bias) & 15u) == 0);
        if (v4) {
            // Output floats o = 0 .. 99 of this warp's row segment: o < 50 is channel 2 pc (TMEM columns 0 .. 49 of its accumulator), o >= 50
            // channel 2 pc + 1.  The segment starts at column 100 pc of a 32-byte aligned row: pc = 0 stores twelve 256-bit groups and
            // one 128-bit tail, pc = 1 one 128-bit head (columns 100 .. 103) and twelve 256-bit groups from column 104.
            const int pc = warp >> 2;
            float* dst = a.X + orow * a.ldx + pc * 100;
            const float* bp = a.bias + pc * 100;
            const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * pc * PP_NT);
            auto put8 = [&](int o, const uint32_t* x) {                      // x[0..7] -> output floats o .. o + 7 (o % 4 == 0)
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + o)), b1 = __ldg(reinterpret_cast<const float4*>(bp + o + 4));
                uint32_t y[8];
                y[0] = __float_as_uint(__uint_as_float(x[0]) + b0.x); y[1] = __float_as_uint(__uint_as_float(x[1]) + b0.y);
                y[2] = __float_as_uint(__uint_as_float(x[2]) + b0.z); y[3] = __float_as_uint(__uint_as_float(x[3]) + b0.w);
                y[4] = __float_as_uint(__uint_as_float(x[4]) + b1.x); y[5] = __float_as_uint(__uint_as_float(x[5]) + b1.y);
                y[6] = __float_as_uint(__uint_as_float(x[6]) + b1.z); y[7] = __float_as_uint(__uint_as_float(x[7]) + b1.w);
                if (eok) st_v8(dst + o, y);
            };
            auto put4 = [&](int o, const uint32_t* x) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bp + o));
                if (eok)
                    *reinterpret_cast<float4*>(dst + o) = make_float4(__uint_as_float(x[0]) + b.x, __uint_as_float(x[1]) + b.y,
                                                                       __uint_as_float(x[2]) + b.z, __uint_as_float(x[3]) + b.w);
            };
            uint32_t w[48];                                                      // channel A columns 0 .. 47
            tmem_ld16(t0, *reinterpret_cast<uint32_t(*)[16]>(w));
            tmem_ld16(t0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(w + 16));
            tmem_ld16(t0 + 32u, *reinterpret_cast<uint32_t(*)[16]>(w + 32));
            uint32_t ta[16], u[40];
            if (pc == 0) {
#pragma unroll
                for (int g = 0; g < 6; ++g) put8(8 * g, w + 8 * g);             // o = 0 .. 47
                tmem_ld16(t0 + 48u, ta);                                         // A 48, 49
                tmem_ld16(t0 + PP_NT, *reinterpret_cast<uint32_t(*)[16]>(u + 2));       // B 0 .. 31
                tmem_ld16(t0 + PP_NT + 16u, *reinterpret_cast<uint32_t(*)[16]>(u + 18));
                u[0] = ta[0]; u[1] = ta[1];                                      // u = {A48, A49, B0 .. B31}: o = 48 .. 81
#pragma unroll
                for (int g = 0; g < 4; ++g) put8(48 + 8 * g, u + 8 * g);        // o = 48 .. 79
                u[0] = u[32]; u[1] = u[33];                                      // carry B30, B31
                tmem_ld16(t0 + PP_NT + 32u, *reinterpret_cast<uint32_t(*)[16]>(u + 2));       // B 32 .. 63 (50 used)
                tmem_ld16(t0 + PP_NT + 48u, *reinterpret_cast<uint32_t(*)[16]>(u + 18));
                put8(80, u);                                                     // B30 .. B37
                put8(88, u + 8);                                                 // B38 .. B45
                put4(96, u + 16);                                                // B46 .. B49
            } else {
                put4(0, w);                                                      // columns 100 .. 103
#pragma unroll
                for (int g = 0; g < 5; ++g) put8(4 + 8 * g, w + 4 + 8 * g);     // o = 4 .. 43
                tmem_ld16(t0 + 48u, ta);
                tmem_ld16(t0 + PP_NT, *reinterpret_cast<uint32_t(*)[16]>(u + 6));
                tmem_ld16(t0 + PP_NT + 16u, *reinterpret_cast<uint32_t(*)[16]>(u + 22));
                u[0] = w[44]; u[1] = w[45]; u[2] = w[46]; u[3] = w[47]; u[4] = ta[0]; u[5] = ta[1];   // u = {A44 .. A49, B0 .. B31}: o = 44 .. 81
#pragma unroll
                for (int g = 0; g < 4; ++g) put8(44 + 8 * g, u + 8 * g);        // o = 44 .. 75
#pragma unroll
                for (int q = 0; q < 6; ++q) u[q] = u[32 + q];                   // carry B26 .. B31
                tmem_ld16(t0 + PP_NT + 32u, *reinterpret_cast<uint32_t(*)[16]>(u + 6));
                tmem_ld16(t0 + PP_NT + 48u, *reinterpret_cast<uint32_t(*)[16]>(u + 22));
#pragma unroll
                for (int g = 0; g < 3; ++g) put8(76 + 8 * g, u + 8 * g);        // B26 .. B49
            }
            __syncwarp();
        } else
        for (int chn = (warp >> 2) * 2; chn < (warp >> 2) * 2 + 2; ++chn) {
            float* dst = a.X + orow * a.ldx + chn * a.C;
            const float* bp = a.bias + chn * a.C;
            for (int col = 0; col < a.C; col += 16) {
                uint32_t rr[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chn * PP_NT + col), rr);
                if (eok) {
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int c = col + j;
                        if (c + 1 < a.C && v2) {
                            *reinterpret_cast<float2*>(dst + c) = make_float2(__uint_as_float(rr[j]) + __ldg(bp + c),
                                                                               __uint_as_float(rr[j + 1]) + __ldg(bp + c + 1));
                        } else {
                            if (c < a.C) dst[c] = __uint_as_float(rr[j]) + __ldg(bp + c);
                            if (c + 1 < a.C) dst[c + 1] = __uint_as_float(rr[j + 1]) + __ldg(bp + c + 1);
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ packed weights by TMA
        if (lane == 0) {
            for (int s = 0; s < nst; ++s) {
                const int slot = s % PP_STAGES;
                const uint32_t ph = (uint32_t)((s / PP_STAGES) & 1);
                mbar_wait(empty_bar + slot, ph ^ 1u);
                unsigned char* st = base + slot * PP_STAGE_BYTES + 2 * PP_A_PLANE;
                mbar_expect_tx(full_bar + slot, 2u * PP_W_PLANE);
                tma_load_2d_cta(&map_wh, full_bar + slot, st, s * 32, 0);
                tma_load_2d_cta(&map_wm, full_bar + slot, st + PP_W_PLANE, s * 32, 0);
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (warp 9)
        // cute::UMMA::InstrDescriptor, kind::f16: c_format F32, a/b BF16, K-major, N >> 3 in [17,23), M >> 4 in [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PP_NT >> 3) << 17) | ((uint32_t)(PP_BM >> 4) << 24);
        for (int s = 0; s < nst; ++s) {
            const int slot = s % PP_STAGES;
            const uint32_t ph = (uint32_t)((s / PP_STAGES) & 1);
            const StageInfo si = stage_tab[s];
            mbar_wait_poll(full_bar + slot, ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // cp.async / st.shared data -> async proxy
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_h = smem_u32(base + slot * PP_STAGE_BYTES), a_m = a_h + PP_A_PLANE;
                const uint32_t b_h = a_h + 2 * PP_A_PLANE, b_m = b_h + PP_W_PLANE;
                const int chn = si.ty < 3 ? si.ty : 3;                          // both LUT gathers feed the co-occurrence channel
                const uint32_t tacc = tmem_base + (uint32_t)(chn * PP_NT);
                const bool fresh = (s == a.base[0]) || (s == a.base[1]) || (s == a.base[2]) || (s == a.base[3]);
                for (int kk = 0; kk < si.nk16; ++kk) {
                    const uint32_t o = (uint32_t)kk * 32u;
                    const uint64_t dah = make_desc_sw64(a_h + o), dam = make_desc_sw64(a_m + o);
                    const uint64_t dbh = make_desc_sw64(b_h + o), dbm = make_desc_sw64(b_m + o);
                    umma_bf16_cta(tacc, dah, dbh, idesc, !(fresh && kk == 0));
                    umma_bf16_cta(tacc, dah, dbm, idesc, 1);
                    umma_bf16_cta(tacc, dam, dbh, idesc, 1);
                }
                umma_commit_cta(empty_bar + slot);
                if (s == nst - 1) umma_commit_cta(acc_bar);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PP_TMEM_COLS) : "memory");
    }
}

void stage_layout(int F_node, int F_edge, int T, int F_lut, int P, int* nblk, int* w16, int* base) {
    const int w[5] = {F_node, F_edge, T, F_lut, F_lut};
    int s = 0;
    for (int u = 0; u < 5; ++u) {
        w16[u] = (w[u] + 15) / 16 * 16;
        nblk[u] = (w16[u] + 31) / 32;
        base[u] = s;
        s += P * nblk[u];
    }
    base[5] = s;
}

}  // namespace

extern "C" int dyg_patch_project_stages(int F_node, int F_edge, int T, int F_lut, int P, int32_t* nblk5) {
    int nblk[5], w16[5], base[6];
    stage_layout(F_node, F_edge, T, F_lut, P, nblk, w16, base);
    if (nblk5)
        for (int u = 0; u < 5; ++u) nblk5[u] = nblk[u];
    return base[5];
}

extern "C" int dyg_patch_project(const dyg_proj_side_t* sides_host, int nsides, const void* node_hi, const void* node_mid,
                                 int ld_node, int F_node, const void* edge_hi, const void* edge_mid, int ld_edge, int F_edge,
                                 const void* lut_hi, const void* lut_mid, int ld_lut, int F_lut, int zero_rows, const double* t_query,
                                 const float* tw, const float* tb, int T, const void* W_hi, const void* W_mid, int ldw,
                                 const float* bias, int P, int C, int S, float* X, int ldx, dyg_stream_t stream) {
    DYG_CHECK_ARG(nsides == 1 || nsides == 2, "dyg_patch_project: nsides=%d (1 or 2)", nsides);
    DYG_CHECK_ARG(P > 0 && C > 0 && C <= PP_NT && S > 0, "dyg_patch_project: patch_size=%d, channel dim=%d (max %d)", P, C, PP_NT);
    DYG_CHECK_ARG(F_node > 0 && F_edge > 0 && T > 0 && F_lut > 0, "dyg_patch_project: bad feature widths");
    DYG_CHECK_ARG(node_hi && node_mid && edge_hi && edge_mid && lut_hi && lut_mid && W_hi && W_mid && t_query && tw && tb && bias && X,
                  "dyg_patch_project: NULL pointer");
    ProjArgs a;
    memset(&a, 0, sizeof(a));
    stage_layout(F_node, F_edge, T, F_lut, P, a.nblk, a.w16, a.base);
    const int lds[5] = {ld_node, ld_edge, 0, ld_lut, ld_lut};
    const void* his[5] = {node_hi, edge_hi, nullptr, lut_hi, lut_hi};
    const void* mids[5] = {node_mid, edge_mid, nullptr, lut_mid, lut_mid};
    for (int u = 0; u < 5; ++u) {
        if (u == 2) continue;
        DYG_CHECK_ARG((lds[u] % 8) == 0 && lds[u] >= a.w16[u], "dyg_patch_project: table %d needs ld %% 8 == 0 and ld >= %d (got %d)", u, a.w16[u], lds[u]);
        DYG_CHECK_ARG(aligned16(his[u]) && aligned16(mids[u]), "dyg_patch_project: table %d planes must be 16-byte aligned", u);
        a.tab_hi[u] = reinterpret_cast<const __nv_bfloat16*>(his[u]);
        a.tab_mid[u] = reinterpret_cast<const __nv_bfloat16*>(mids[u]);
        a.ld[u] = lds[u];
    }
    const int nst = a.base[5];
    DYG_CHECK_ARG(ldw >= nst * 32 && (ldw % 8) == 0 && aligned16(W_hi) && aligned16(W_mid),
                  "dyg_patch_project: packed weights need ld >= %d (multiple of 8), 16-byte aligned planes", nst * 32);
    int64_t tiles = 0;
    for (int i = 0; i < nsides; ++i) {
        const dyg_proj_side_t& sd = sides_host[i];
        DYG_CHECK_ARG(sd.tokens >= 0 && sd.ntok > 0 && sd.tok_off >= 0 && sd.tok_off + sd.ntok <= S, "dyg_patch_project: bad side %d", i);
        DYG_CHECK_ARG(sd.tokens == 0 || (sd.ids && sd.eids && sd.t_nbr && sd.cnt_a && sd.cnt_b), "dyg_patch_project: side %d has NULL arrays", i);
        a.side[i] = sd;
        const int64_t t = (sd.tokens + PP_BM - 1) / PP_BM;
        if (i == 0) a.tiles0 = t;
        tiles += t;
    }
    if (tiles == 0) return 0;
    DYG_CHECK_ARG(tiles < ((int64_t)1 << 31), "dyg_patch_project: too many tiles");
    a.t_query = t_query; a.tw = tw; a.tb = tb; a.T = T; a.P = P; a.C = C; a.S = S;
    a.bias = bias; a.X = X; a.ldx = ldx;
    a.zero_rows = zero_rows & 3;
    CUtensorMap mwh, mwm;
    if (!dyg_tensor_map_bf16(W_hi, PP_NT, (uint64_t)nst * 32, (uint64_t)ldw, PP_NT, &mwh)) return 1;
    if (!dyg_tensor_map_bf16(W_mid, PP_NT, (uint64_t)nst * 32, (uint64_t)ldw, PP_NT, &mwm)) return 1;
    DYG_CHECK_ARG(nst <= PP_MAX_STAGES, "dyg_patch_project: %d stages exceed the stage table (%d)", nst, PP_MAX_STAGES);
    const size_t smem = (size_t)PP_STAGES * PP_STAGE_BYTES + 1024 + 128 + (size_t)nst * sizeof(StageInfo) + (size_t)2 * T * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(patch_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_patch_project: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    patch_project_kernel<<<(unsigned)tiles, PP_THREADS, smem, as_stream(stream)>>>(mwh, mwm, a);
    DYG_LAUNCH_CHECK("dyg_patch_project");
    return 0;
}
