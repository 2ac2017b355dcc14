// dyg_attn_block: the whole self-attention half of DyGFormer's TransformerEncoder (models/DyGFormer.py:442-455) in ONE kernel:
//
//   x1 = x + b_o + sum_h softmax(q_h k_h^T) v'_h,     [q | k | v'] = W_cat LayerNorm(x) + b_cat
//
// (W_cat / b_cat: ops.attn_fold_weights — q pre-scaled by log2(e)/sqrt(hd), the out-projection folded into v').  The projected
// rows never reach HBM: every CTA owns a 128-row tile (128 / SP sequences in slots of SP rows) and alternates two phases on it,
//
//   phase A  [q | k | v'] = LN(x) W_cat^T + b_cat   tcgen05 BF16x3, M = 128, N = 208 per n tile, A = LN(x) resident in shared
//            memory (written as an operand image by the IO warps during the previous tile's phase B, one bulk copy), W_cat
//            streamed by TMA, two accumulators in TMEM; the epilogue warps write bf16 hi | mid planes into the CTA's private
//            scratch (418 KB, reused every tile: it lives in L2)
//   phase B  attention over the tile (seq_attn_tc.cu's pipeline): K_h / V'_h back from the scratch by TMA into the shared
//            memory phase A no longer needs, Q and P as TMEM A operands, O' accumulated over the heads, + x + b_o -> x1
//
// so HBM sees x once in and x1 once out.  Warps: 0 TMA producer, 1 MMA issuer, 2-9 epilogue of phase A; in phase B warps 2-5
// normalise the NEXT tile's rows and run the softmax, warps 6-9 move Q rows into TMEM and drain O'.
//
// Measured on B200 (profiles/README.md, round 2): parity-green, but 2.7 - 3.6 ms per DyGFormer launch against 2.13 ms for the
// three-launch form (layernorm_split + projection GEMM + dyg_seq_attention_fold).  The phases of a tile cannot overlap (phase
// B's K | V' alias phase A's operands: BF16x3 planes leave no room for both), every CTA streams the whole W_cat (731 KB) per
// tile, and the scratch round trip moves the HBM traffic to L2, whose bandwidth is about the same: ~2.1 MB of L2 traffic per
// tile.  It is kept as the DYG_FUSED_ATTN=2 path; the default is the three-launch form.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace {

constexpr int AB_ROWS = 128;
constexpr int AB_KB = 7;                          // 32-column k blocks of the projection (D <= 224)
constexpr int AB_BLK = AB_ROWS * 64;              // one 32-column block of 128 rows, one plane: 8 KB
constexpr int AB_A_STAGE = 2 * AB_BLK;            // hi | mid
constexpr int AB_A_BYTES = AB_KB * AB_A_STAGE;    // 112 KB
constexpr int AB_NT = 208;                        // columns per n tile of the projection (MMA N)
constexpr int AB_W_PLANE = AB_NT * 64;
constexpr int AB_W_STAGE = 2 * AB_W_PLANE;        // 26 KB
constexpr int AB_WS = 4;
constexpr int AB_KSTEPS = 7;                      // k16 steps of Q K^T (head_dim <= 112)
constexpr int AB_KBLK = 4;
constexpr int AB_NV = 208;                        // MMA N of P V'
constexpr int AB_VCH = 7;
constexpr int AB_K_PLANE = AB_KBLK * AB_BLK;      // 32 KB
constexpr int AB_V_PLANE = AB_VCH * AB_BLK;       // 56 KB
constexpr int AB_THREADS = 320;
constexpr uint32_t AB_ACC_COLS = 256;             // accumulator buffer stride (phase A)
constexpr uint32_t AB_O_COL = 0, AB_S_COL = 256, AB_Q_COL = 384;   // phase B
constexpr int AB_SMEM = AB_A_BYTES + AB_WS * AB_W_STAGE;           // phase B's K | V' (176 KB) aliases A | W ring

struct BlockArgs {
    const float* x;
    const float* gamma;
    const float* beta;
    const float* bcat;
    const float* bout;
    float* out;
    __nv_bfloat16* sc_hi;           // scratch planes (gridDim.x * 128, ldp)
    __nv_bfloat16* sc_mid;
    unsigned char* img;             // gridDim.x operand images of LN(x)
    int64_t B;
    int64_t tiles;
    int ldx, ldo, ldp;
    int S, SP, sp_shift, NS;
    int H, hd, hdk, D, N, n_tiles;
    int q_col0, k_col0, v_col0;
    float eps;
};

__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(AB_THREADS, 1) attn_block_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                   const __grid_constant__ CUtensorMap map_wm,
                                                                   const __grid_constant__ CUtensorMap map_sh,
                                                                   const __grid_constant__ CUtensorMap map_sm, const BlockArgs a) {
    extern __shared__ __align__(1024) unsigned char ab_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ab_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_res = base;                               // phase A: LN(x) operand, 7 k blocks
    unsigned char* w_ring = base + AB_A_BYTES;                 // phase A: W_cat stages
    unsigned char* k_buf = base;                               // phase B: K_hi | K_mid
    unsigned char* v_buf = base + 2 * AB_K_PLANE;              // phase B: V'_hi | V'_mid
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + AB_SMEM);
    uint64_t* w_full = bars;            // [4]
    uint64_t* w_empty = bars + 4;       // [4]
    uint64_t* a_full = bars + 8;        // bulk copy of the image landed
    uint64_t* tfull = bars + 9;         // [2] accumulator complete
    uint64_t* tempty = bars + 11;       // [2] 8 epilogue warps drained it
    uint64_t* planes_ready = bars + 13; // 8 epilogue warps: the tile's [q | k | v'] planes are in the scratch
    uint64_t* img_ready = bars + 14;    // 4 IO warps: the next tile's LN(x) image is written
    uint64_t* k_full = bars + 15;
    uint64_t* k_empty = bars + 16;
    uint64_t* v_full = bars + 17;
    uint64_t* v_empty = bars + 18;
    uint64_t* q_full = bars + 19;       // 4 IO warps
    uint64_t* s_full = bars + 20;
    uint64_t* p_full = bars + 21;       // 4 softmax warps
    uint64_t* o_full = bars + 22;
    uint64_t* o_empty = bars + 23;      // 4 IO warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int nkb = (a.D + 31) >> 5;
    unsigned char* my_img = a.img + (size_t)blockIdx.x * AB_A_BYTES;
    const int64_t my_tiles = a.tiles > blockIdx.x ? (a.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // zero this CTA's operand image once: its K padding columns (D .. 32 nkb) are never written again
    for (int i = tid; i < AB_A_BYTES / 16; i += AB_THREADS) reinterpret_cast<uint4*>(my_img)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < AB_WS; ++s) {
            mbar_init(w_full + s, 1);
            mbar_init(w_empty + s, 1);
        }
        mbar_init(a_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 8);
        }
        mbar_init(planes_ready, 8);
        mbar_init(img_ready, 4);
        mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(q_full, 4); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wm)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_sh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_sm)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int64_t n = 0;
            const int row0 = (int)blockIdx.x * AB_ROWS;
            // contiguous rows of every sequence of a tile -> L2 (one bulk prefetch per sequence)
            auto prefetch_x = [&](int64_t tile) {
                if (tile >= a.tiles) return;
                for (int sl = 0; sl < a.NS; ++sl) {
                    const int64_t seq = tile * a.NS + sl;
                    if (seq >= a.B) break;
                    const char* p = reinterpret_cast<const char*>(a.x + seq * a.S * a.ldx);
                    const uint32_t bytes = (uint32_t)(a.S * a.ldx * 4) & ~15u;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
                }
            };
            for (int64_t it = 0; it < my_tiles; ++it) {
                // ---- phase A: the image of this tile, then the W_cat stages
                prefetch_x(blockIdx.x + (it + 1) * gridDim.x);                // the next tile's rows: LayerNorm reads them during this tile's phase B
                mbar_wait(img_ready, (uint32_t)(it & 1));
                if (it > 0) mbar_wait(o_full, (uint32_t)((it - 1) & 1));     // every MMA of the previous phase B retired: smem is free
                mbar_expect_tx(a_full, (uint32_t)AB_A_BYTES);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(a_res)),
                             "l"(my_img), "r"((uint32_t)AB_A_BYTES), "r"(smem_u32(a_full))
                             : "memory");
                for (int nt = 0; nt < a.n_tiles; ++nt) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(w_empty + s, ph ^ 1u);
                        mbar_expect_tx(w_full + s, (uint32_t)AB_W_STAGE);
                        unsigned char* st = w_ring + (size_t)s * AB_W_STAGE;
                        tma_load_2d_1cta(&map_wh, w_full + s, st, kb * 32, nt * AB_NT);
                        tma_load_2d_1cta(&map_wm, w_full + s, st + AB_W_PLANE, kb * 32, nt * AB_NT);
                        if (++s == AB_WS) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
                mbar_wait(planes_ready, (uint32_t)(it & 1));
                prefetch_x(blockIdx.x + it * gridDim.x);                      // this tile's rows again, for the residual of the final epilogue
                for (int h = 0; h < a.H; ++h, ++n) {
                    const uint32_t par = (uint32_t)(n & 1);
                    mbar_wait(k_empty, par ^ 1u);
                    mbar_expect_tx(k_full, 2u * AB_K_PLANE);
                    const int kc = a.k_col0 + h * a.hdk;
#pragma unroll
                    for (int j = 0; j < AB_KBLK; ++j) {
                        tma_load_2d_1cta(&map_sh, k_full, k_buf + j * AB_BLK, kc + 32 * j, row0);
                        tma_load_2d_1cta(&map_sm, k_full, k_buf + AB_K_PLANE + j * AB_BLK, kc + 32 * j, row0);
                    }
                    mbar_wait(v_empty, par ^ 1u);
                    mbar_expect_tx(v_full, 2u * AB_V_PLANE);
                    const int vc = a.v_col0 + h * a.D;
#pragma unroll
                    for (int j = 0; j < AB_VCH; ++j) {
                        tma_load_2d_1cta(&map_sh, v_full, v_buf + j * AB_BLK, vc + 32 * j, row0);
                        tma_load_2d_1cta(&map_sm, v_full, v_buf + AB_V_PLANE + j * AB_BLK, vc + 32 * j, row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, elected issue)
        const uint32_t idesc_pj = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AB_NT >> 3) << 17) | ((uint32_t)(AB_ROWS >> 4) << 24);
        const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AB_ROWS >> 3) << 17) | ((uint32_t)(AB_ROWS >> 4) << 24);
        const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AB_NV >> 3) << 17) | ((uint32_t)(AB_ROWS >> 4) << 24);
        const uint64_t ad_h = make_desc_sw64(smem_u32(a_res)), ad_m = ad_h + (uint64_t)(AB_BLK >> 4);
        const uint64_t kd_h = make_desc_sw64(smem_u32(k_buf)), kd_m = make_desc_sw64(smem_u32(k_buf + AB_K_PLANE));
        const uint64_t vd_h = make_desc_mn_sw64(smem_u32(v_buf), AB_BLK), vd_m = make_desc_mn_sw64(smem_u32(v_buf + AB_V_PLANE), AB_BLK);
        const uint32_t tq = tmem_base + AB_Q_COL, ts = tmem_base + AB_S_COL, to = tmem_base + AB_O_COL;
        int s = 0;
        uint32_t ph = 0;
        int64_t n = 0;
        uint32_t cnt0 = 0, cnt1 = 0;       // uses of the two accumulator buffers
        for (int64_t it = 0; it < my_tiles; ++it) {
            // ---- phase A
            mbar_wait(a_full, (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int nt = 0; nt < a.n_tiles; ++nt) {
                const int buf = (nt + 1) & 1;                   // buffer 1 first: buffer 0 is O', still drained by the previous tile's epilogue
                if (buf == 0 && nt == 1 && it > 0) mbar_wait(o_empty, (uint32_t)((it - 1) & 1));
                const uint32_t c = buf ? cnt1 : cnt0;
                mbar_wait(tempty + buf, (c & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)buf * AB_ACC_COLS;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(w_full + s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t wd_h = make_desc_sw64(smem_u32(w_ring + (size_t)s * AB_W_STAGE)), wd_m = wd_h + (uint64_t)(AB_W_PLANE >> 4);
                    const uint64_t ao = (uint64_t)((kb * AB_A_STAGE) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        if (kb * 32 + kk * 16 < a.D) {
                            const uint64_t o = (uint64_t)(kk * 2);
                            umma_ss_e(tacc, ad_h + ao + o, wd_h + o, idesc_pj, (kb | kk) != 0);
                            umma_ss_e(tacc, ad_h + ao + o, wd_m + o, idesc_pj, 1);
                            umma_ss_e(tacc, ad_m + ao + o, wd_h + o, idesc_pj, 1);
                        }
                    }
                    umma_commit_e(w_empty + s);
                    if (++s == AB_WS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit_e(tfull + buf);
                if (buf) ++cnt1; else ++cnt0;
            }
            // ---- phase B
            for (int h = 0; h < a.H; ++h, ++n) {
                const uint32_t par = (uint32_t)(n & 1);
                mbar_wait(q_full, par);
                mbar_wait(k_full, par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < AB_KSTEPS; ++kk) {
                    const uint64_t o = (uint64_t)(((kk >> 1) * AB_BLK + (kk & 1) * 32) >> 4);
                    umma_ts_e(ts, tq + 8 * kk, kd_h + o, idesc_qk, kk != 0);
                    umma_ts_e(ts, tq + 8 * kk, kd_m + o, idesc_qk, 1);
                    umma_ts_e(ts, tq + 56 + 8 * kk, kd_h + o, idesc_qk, 1);
                }
                umma_commit_e(k_empty);
                umma_commit_e(s_full);
                mbar_wait(p_full, par);
                mbar_wait(v_full, par);
                if (h == 0 && a.n_tiles < 2 && it > 0) mbar_wait(o_empty, (uint32_t)((it - 1) & 1));   // (phase A did not touch buffer 0)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < AB_ROWS / 16; ++kk) {
                    const uint64_t o = (uint64_t)((kk * 16 * 64) >> 4);
                    umma_ts_e(to, ts + 8 * kk, vd_h + o, idesc_pv, (h | kk) != 0);
                    umma_ts_e(to, ts + 8 * kk, vd_m + o, idesc_pv, 1);
                    umma_ts_e(to, ts + 64 + 8 * kk, vd_h + o, idesc_pv, 1);
                }
                umma_commit_e(v_empty);
                if (h == a.H - 1) umma_commit_e(o_full);
            }
        }
    } else {
        // ------------------------------------------------------------------ warps 2-9
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                       // tile row == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int slot = r >> a.sp_shift, i = r & (a.SP - 1);
        const int win = slot * a.SP;
        const bool io = warp >= 6;
        GemmArgs eg;
        eg.bias = a.bcat; eg.residual = nullptr; eg.C = nullptr;
        eg.Chi = a.sc_hi + (size_t)blockIdx.x * AB_ROWS * a.ldp;
        eg.Cmid = a.sc_mid + (size_t)blockIdx.x * AB_ROWS * a.ldp;
        eg.M = AB_ROWS; eg.ldr = 0; eg.ldc = 0; eg.ldcs = a.ldp; eg.N = a.N; eg.K = 0; eg.act = DYG_ACT_NONE;
        const __nv_bfloat16* my_qh = eg.Chi + (size_t)r * a.ldp + a.q_col0;
        const __nv_bfloat16* my_qm = eg.Cmid + (size_t)r * a.ldp + a.q_col0;

        // LayerNorm of the rows of tile `tile` owned by this softmax warp (32 rows, lane = 8 columns) -> operand image in global memory
        const int c0 = 8 * lane;
        const bool own = c0 < a.D;
        auto ln_image = [&](int64_t tile) {
            constexpr int RB = 8;
            const int kb = lane >> 2;
            const uint32_t chunk = (uint32_t)(lane & 3);
            const int rbase = quarter * 32;
            for (int i0 = 0; i0 < 32; i0 += RB) {
                // gamma / beta are re-read per batch (L1 / L2 hits) instead of living in registers across the whole tile loop:
                // with 223 KB of shared memory the L1 is too small to absorb spills
                float gm8[8], bt8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    gm8[j] = own ? __ldg(a.gamma + c0 + j) : 0.f;
                    bt8[j] = own ? __ldg(a.beta + c0 + j) : 0.f;
                }
                float4 v[RB][2];
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const int rr = rbase + i0 + k;
                    const int64_t seq = tile * a.NS + (rr >> a.sp_shift);
                    const int ii = rr & (a.SP - 1);
                    if (own && ii < a.S && seq < a.B) {
                        const float4* xp = reinterpret_cast<const float4*>(a.x + (seq * a.S + ii) * a.ldx + c0);
                        v[k][0] = xp[0];
                        v[k][1] = xp[1];
                    } else {
                        v[k][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                        v[k][1] = v[k][0];
                    }
                }
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const float xs[8] = {v[k][0].x, v[k][0].y, v[k][0].z, v[k][0].w, v[k][1].x, v[k][1].y, v[k][1].z, v[k][1].w};
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) sum += xs[j];
                    const float mean = warp_sum(sum) / (float)a.D;
                    float sq = 0.f;
                    if (own) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float d = xs[j] - mean;
                            sq += d * d;
                        }
                    }
                    const float rstd = rsqrtf(warp_sum(sq) / (float)a.D + a.eps);
                    if (own) {
                        const int rr = rbase + i0 + k;
                        const int64_t seq = tile * a.NS + (rr >> a.sp_shift);
                        const bool ok = (rr & (a.SP - 1)) < a.S && seq < a.B;
                        uint32_t hh[4], ll[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float y0 = ok ? (xs[2 * j] - mean) * rstd * gm8[2 * j] + bt8[2 * j] : 0.f;
                            const float y1 = ok ? (xs[2 * j + 1] - mean) * rstd * gm8[2 * j + 1] + bt8[2 * j + 1] : 0.f;
                            split_pack(y0, y1, hh[j], ll[j]);
                        }
                        unsigned char* p = my_img + kb * AB_A_STAGE + (uint32_t)((rr >> 3) * 512 + (rr & 7) * 64) + ((chunk ^ (uint32_t)((rr >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(p) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                        *reinterpret_cast<uint4*>(p + AB_BLK) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                    }
                }
            }
            asm volatile("fence.proxy.async;" ::: "memory");           // generic global writes -> visible to the bulk copy engine
            __syncwarp();
            if (lane == 0) mbar_arrive(img_ready);
        };
        // Q row of head h of this tile: scratch -> registers (L2 loads; the scratch is written by this CTA in the same launch)
        auto load_q = [&](int h, uint32_t (&qh)[56], uint32_t (&qm)[56]) {
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                uint4 vh = make_uint4(0u, 0u, 0u, 0u), vm = vh;
                if (8 * j < a.hdk) {
                    vh = ld_cg_v4(my_qh + h * a.hdk + 8 * j);
                    vm = ld_cg_v4(my_qm + h * a.hdk + 8 * j);
                }
                qh[4 * j] = vh.x; qh[4 * j + 1] = vh.y; qh[4 * j + 2] = vh.z; qh[4 * j + 3] = vh.w;
                qm[4 * j] = vm.x; qm[4 * j + 1] = vm.y; qm[4 * j + 2] = vm.z; qm[4 * j + 3] = vm.w;
            }
        };
        auto store_q = [&](const uint32_t (&qh)[56], const uint32_t (&qm)[56]) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                tmem_st8(lane_addr + AB_Q_COL + 8 * j, qh + 8 * j);
                tmem_st8(lane_addr + AB_Q_COL + 56 + 8 * j, qm + 8 * j);
            }
            tmem_st_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        };

        // the zero fill of the image (all threads, before the barrier above) precedes the first LayerNorm
        if (!io && my_tiles > 0) ln_image((int64_t)blockIdx.x);
        const int chunk0 = (warp - 2) >> 2;
        uint32_t ecnt0 = 0, ecnt1 = 0;
        int64_t n = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            // ---- phase A: accumulator tiles -> [q | k | v'] planes in the scratch (all 128 rows: rows of no token hold the bias)
            for (int nt = 0; nt < a.n_tiles; ++nt) {
                const int buf = (nt + 1) & 1;
                const uint32_t c = buf ? ecnt1 : ecnt0;
                mbar_wait(tfull + buf, c & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int n0 = nt * AB_NT;
                epilogue_tile(eg, lane_addr + (uint32_t)buf * AB_ACC_COLS, (int64_t)r, n0, min(AB_NT, a.N - n0), chunk0, 2);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + buf);
                if (buf) ++ecnt1; else ++ecnt0;
            }
            asm volatile("fence.proxy.async;" ::: "memory");           // the planes are read back by TMA
            __syncwarp();
            if (lane == 0) mbar_arrive(planes_ready);
            mbar_wait(planes_ready, (uint32_t)(it & 1));
            // ---- phase B
            const int64_t seq = tile * a.NS + slot;
            const bool rowok = i < a.S && seq < a.B;
            if (!io) {
                if (it + 1 < my_tiles) ln_image(tile + gridDim.x);   // the next tile's operand, while Q / K of this one are on their way
                for (int h = 0; h < a.H; ++h, ++n) {
                    mbar_wait(s_full, (uint32_t)(n & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t sv[64];
                    const uint32_t sa = lane_addr + AB_S_COL + (uint32_t)win;
                    tmem_ld16_nowait(sa, sv);
                    tmem_ld16_nowait(sa + 16, sv + 16);
                    if (a.SP == 64) {
                        tmem_ld16_nowait(sa + 32, sv + 32);
                        tmem_ld16_nowait(sa + 48, sv + 48);
                    }
                    tmem_ld_wait();
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 64; ++j)
                        if (j < a.S) mx = fmaxf(mx, __uint_as_float(sv[j]));
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const float p = (j < a.S) ? ex2_approx(__uint_as_float(sv[j]) - mx) : 0.f;
                        sum += p;
                        sv[j] = __float_as_uint(p);
                    }
                    const float inv = rowok ? 1.f / sum : 0.f;
                    const uint32_t pa = lane_addr + AB_S_COL;
                    const int wc = win >> 1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int rel = 16 * c - wc;
                        if (rel == 0 || (rel == 16 && a.SP == 64)) {
                            uint32_t ph[16], pm[16];
                            if (rel == 0) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) split_pack(__uint_as_float(sv[2 * j]) * inv, __uint_as_float(sv[2 * j + 1]) * inv, ph[j], pm[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    split_pack(__uint_as_float(sv[32 + 2 * j]) * inv, __uint_as_float(sv[33 + 2 * j]) * inv, ph[j], pm[j]);
                            }
                            tmem_st16(pa + 16 * c, ph);
                            tmem_st16(pa + 64 + 16 * c, pm);
                        } else {
                            tmem_st16_zero(pa + 16 * c);
                            tmem_st16_zero(pa + 64 + 16 * c);
                        }
                    }
                    tmem_st_wait();
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_full);
                }
            } else {
                {
                    uint32_t qh[56], qm[56];
                    load_q(0, qh, qm);
                    store_q(qh, qm);
                }
                for (int h = 0; h < a.H; ++h, ++n) {
                    uint32_t qh[56], qm[56];
                    load_q(h + 1 < a.H ? h + 1 : h, qh, qm);
                    mbar_wait(s_full, (uint32_t)(n & 1));            // Q K^T of this head has retired: the Q columns are free
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (h + 1 < a.H) store_q(qh, qm);
                }
                // ---- final epilogue of the tile: x1 = O' + x + b_o
                const int64_t row = rowok ? seq * a.S + i : 0;
                const float* xr = a.x + row * a.ldx;
                float* orow = a.out + row * a.ldo;
                float xa[16], xb[16], xc[16];
#define AB_LOAD_X(COL, XV)                                          \
    do {                                                             \
        _Pragma("unroll") for (int j = 0; j < 16; ++j) XV[j] = 0.f;  \
        if (rowok && (COL) < a.D) {                                  \
            ld_v8(xr + (COL), XV);                                   \
            if ((COL) + 8 < a.D) ld_v8(xr + (COL) + 8, XV + 8);      \
        }                                                            \
    } while (0)
#define AB_EMIT(COL, XV)                                                                              \
    do {                                                                                               \
        if ((COL) < a.D) {                                                                             \
            const bool wide = (COL) + 8 < a.D;                                                         \
            float bv[16];                                                                              \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                            \
                const float4 t = (j < 2 || wide) ? __ldg(reinterpret_cast<const float4*>(a.bout + (COL)) + j) \
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);                    \
                bv[4 * j] = t.x; bv[4 * j + 1] = t.y; bv[4 * j + 2] = t.z; bv[4 * j + 3] = t.w;        \
            }                                                                                          \
            uint32_t rr[16];                                                                           \
            tmem_ld16(lane_addr + AB_O_COL + (uint32_t)(COL), rr);                                     \
            if (rowok) {                                                                               \
                uint32_t o[16];                                                                        \
                _Pragma("unroll") for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(rr[j]) + XV[j] + bv[j]); \
                st_v8(orow + (COL), o);                                                                \
                if (wide) st_v8(orow + (COL) + 8, o + 8);                                              \
            }                                                                                          \
            __syncwarp();                                                                              \
        }                                                                                              \
    } while (0)
                AB_LOAD_X(0, xa);
                AB_LOAD_X(16, xb);
                mbar_wait(o_full, (uint32_t)(it & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int col = 0; col < a.D; col += 48) {
                    AB_LOAD_X(col + 32, xc);
                    AB_EMIT(col, xa);
                    AB_LOAD_X(col + 48, xa);
                    AB_EMIT(col + 16, xb);
                    AB_LOAD_X(col + 64, xb);
                    AB_EMIT(col + 32, xc);
                }
#undef AB_LOAD_X
#undef AB_EMIT
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int ab_ldp(int N) { return (N + 15) / 16 * 16; }

}  // namespace

extern "C" int64_t dyg_attn_block_workspace_bytes(int N) {
    return (int64_t)dyg_num_sms() * ((int64_t)2 * AB_ROWS * ab_ldp(N) * 2 + AB_A_BYTES);
}

extern "C" int dyg_attn_block(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W_hi, const void* W_mid,
                              int ldw, const float* bcat, int N, int q_col0, int k_col0, int v_col0, const float* bout, int64_t B, int S,
                              int H, int hd, int D, float* out, int ldo, void* workspace, int64_t workspace_bytes, dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && S > 0 && H > 0 && hd > 0 && D > 0 && N > 0, "dyg_attn_block: bad sizes");
    DYG_CHECK_ARG(S <= 64, "dyg_attn_block: S=%d unsupported (max 64)", S);
    DYG_CHECK_ARG(hd <= 112 && (hd % 4) == 0, "dyg_attn_block: head_dim=%d unsupported (multiple of 4, max 112)", hd);
    DYG_CHECK_ARG(D <= AB_NV && (D % 8) == 0, "dyg_attn_block: width %d unsupported (multiple of 8, max %d)", D, AB_NV);
    if (B == 0) return 0;
    DYG_CHECK_ARG(x && gamma && beta && W_hi && W_mid && bcat && bout && out && workspace, "dyg_attn_block: NULL pointer");
    const int hdk = (hd + 7) / 8 * 8;
    DYG_CHECK_ARG((q_col0 % 8) == 0 && (k_col0 % 8) == 0 && (v_col0 % 8) == 0 && q_col0 >= 0 && k_col0 >= q_col0 + H * hdk &&
                      v_col0 >= k_col0 + H * hdk && N == v_col0 + H * D,
                  "dyg_attn_block: segment offsets q=%d k=%d v=%d do not form [q | k | v'] of %d columns", q_col0, k_col0, v_col0, N);
    DYG_CHECK_ARG((ldw % 8) == 0 && ldw >= D && aligned16(W_hi) && aligned16(W_mid), "dyg_attn_block: weight planes must be 16-byte aligned, ldw %% 8 == 0");
    DYG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 31u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0 && (ldx % 8) == 0 && (ldo % 8) == 0 &&
                      aligned16(bout) && aligned16(bcat) && (reinterpret_cast<uintptr_t>(gamma) & 3u) == 0,
                  "dyg_attn_block: x / out must be 32-byte aligned with leading dimensions that are multiples of 8, bias vectors 16-byte aligned");
    DYG_CHECK_ARG(B * (int64_t)S < ((int64_t)1 << 31), "dyg_attn_block: too many tokens");
    DYG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023u) == 0 && workspace_bytes >= dyg_attn_block_workspace_bytes(N),
                  "dyg_attn_block: workspace of %lld bytes (1024-byte aligned) required", (long long)dyg_attn_block_workspace_bytes(N));
    BlockArgs a;
    memset(&a, 0, sizeof(a));
    const int sms = dyg_num_sms();
    a.x = x; a.gamma = gamma; a.beta = beta; a.bcat = bcat; a.bout = bout; a.out = out;
    a.ldx = ldx; a.ldo = ldo; a.ldp = ab_ldp(N);
    a.B = B; a.S = S; a.SP = S <= 32 ? 32 : 64; a.sp_shift = S <= 32 ? 5 : 6; a.NS = AB_ROWS / a.SP;
    a.tiles = (B + a.NS - 1) / a.NS;
    a.H = H; a.hd = hd; a.hdk = hdk; a.D = D; a.N = N; a.n_tiles = (N + AB_NT - 1) / AB_NT;
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0; a.eps = eps;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    a.img = ws;
    a.sc_hi = reinterpret_cast<__nv_bfloat16*>(ws + (size_t)sms * AB_A_BYTES);
    a.sc_mid = a.sc_hi + (size_t)sms * AB_ROWS * a.ldp;
    CUtensorMap mwh, mwm, msh, msm;
    if (!dyg_tensor_map_bf16(W_hi, (uint64_t)N, (uint64_t)D, (uint64_t)ldw, AB_NT, &mwh)) return 1;
    if (!dyg_tensor_map_bf16(W_mid, (uint64_t)N, (uint64_t)D, (uint64_t)ldw, AB_NT, &mwm)) return 1;
    if (!dyg_tensor_map_bf16(a.sc_hi, (uint64_t)sms * AB_ROWS, (uint64_t)N, (uint64_t)a.ldp, AB_ROWS, &msh)) return 1;
    if (!dyg_tensor_map_bf16(a.sc_mid, (uint64_t)sms * AB_ROWS, (uint64_t)N, (uint64_t)a.ldp, AB_ROWS, &msm)) return 1;
    const size_t smem = (size_t)AB_SMEM + 1024 + 512;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_attn_block: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = true;
    }
    const unsigned grid = (unsigned)(a.tiles < sms ? a.tiles : sms);
    attn_block_kernel<<<grid, AB_THREADS, smem, as_stream(stream)>>>(mwh, mwm, msh, msm, a);
    DYG_LAUNCH_CHECK("dyg_attn_block");
    return 0;
}
