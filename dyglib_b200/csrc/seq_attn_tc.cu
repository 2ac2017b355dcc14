// dyg_seq_attention_fold: DyGFormer's self-attention sub-block (models/DyGFormer.py:442-455, nn.MultiheadAttention without a
// mask) on tcgen05, with the output projection folded into the value projection:
//
//   x1 = x + b_o + sum_h softmax(q_h k_h^T) v'_h         v'_h = (W_o[:, h] W_v[h]) LN(x) + W_o[:, h] b_v[h]   (D wide per head)
//
// so the out-projection GEMM and the round trip of the attention output disappear (softmax rows sum to one, which is what
// lets the bias through).  q is pre-scaled by log2(e) / sqrt(hd) in the projection weights; the kernel uses exp2.
//
// Inputs are BF16x3 operand planes (hi | mid) written by the projection GEMM's epilogue, one row per token:
//   [ q_0 .. q_{H-1} | k_0 .. k_{H-1} (hdk each, hdk = hd rounded up to 8, zero padded) | v'_0 .. v'_{H-1} (D each) ]
//
// One CTA per SM (persistent), a tile = 128 rows = 128 / SP sequences in slots of SP = 32 or 64 rows (S <= SP valid):
//   warp 0     TMA producer: K_h (K-major SWIZZLE_64B blocks) and V'_h (the same boxes read as an MN-major operand) through
//              3-D tensor maps (column, token in sequence, sequence): rows past S are zero-filled, never read from HBM
//   warp 1     TMEM allocation + MMA issue (cta_group::1, M = 128): S = Q K^T with A = Q from TMEM (BF16x3: three MMAs per
//              k16 step), O' += P V'_h with A = P from TMEM and B = V' MN-major; O' accumulates over the heads
//   warps 2-5  thread = tile row: softmax of the row's slot window (exp2, masked), P as bf16 hi | mid written in place over S
//   warps 6-9  thread = tile row: Q row global -> registers -> TMEM one step ahead
//   both groups share the final epilogue O' + x + b_o -> x1 (even / odd 16-column chunks, the whole residual row requested up front)
// Keys of the other slots get P = 0, so one M=128 x K=128 product serves both sequences of a tile.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace {

constexpr int SA_ROWS = 128;
constexpr int SA_KSTEPS = 7;                      // k16 steps of Q K^T (head_dim <= 112)
constexpr int SA_KBLK = 4;                        // 32-column blocks of a K plane (the last one half used)
constexpr int SA_NV = 208;                        // MMA N of P V' (model width rounded up to 16)
constexpr int SA_VCH = 7;                         // 32-feature chunks of a V' plane
constexpr int SA_BLK = SA_ROWS * 64;              // one 32-column block of 128 rows: 8 KB
constexpr int SA_K_PLANE = SA_KBLK * SA_BLK;      // 32 KB
constexpr int SA_V_PLANE = SA_VCH * SA_BLK;       // 56 KB
constexpr int SA_THREADS = 320;                   // TMA, MMA, 4 softmax warps, 4 IO warps
constexpr uint32_t SA_O_COL = 0;                  // O' accumulator, 208 columns
constexpr uint32_t SA_S_COL = 256;                // scores (128 fp32 columns), then P_hi [0,64) | P_mid [64,128) in place
constexpr uint32_t SA_Q_COL = 384;                // Q_hi [0,56) | Q_mid [56,112)

struct AttnArgs {
    const __nv_bfloat16* q_hi;      // planes (M, ldp)
    const __nv_bfloat16* q_mid;
    const float* x;                 // (M, ldx) residual
    const float* bias;              // (D)
    float* out;                     // (M, ldo)
    int64_t B;                      // sequences
    int64_t tiles;
    int ldp, ldx, ldo;
    int S, SP, NS;                  // tokens per sequence, slot rows (32 | 64), slots per tile
    int H, hd, hdk, D;
    int q_col0, k_col0, v_col0;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Final epilogue of a tile, shared by the softmax and the IO warp of a TMEM lane quarter (16-column chunks `first`, `first + 2`, ...):
// x1 = O' + x + b_o.  ALL residual chunks of the row are requested before the accumulator is awaited (up to 7 x 16 registers): with
// two chunks in flight the epilogue was one HBM round trip per pair of chunks, 13 us per tile on the critical path of the single O'
// buffer (globaltimer stamps, profiles/README.md).
__device__ __forceinline__ void sa_tile_epilogue(const AttnArgs& a, const float* bias_s, uint32_t lane_addr, bool rowok, const float* xr, float* orow,
                                                 int first, uint64_t* o_full, uint32_t parity) {
    constexpr int MAXC = (SA_NV / 16 + 1) / 2;                  // chunks per warp
    float xv[MAXC][16];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int col = 16 * (first + 2 * c);
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[c][j] = 0.f;
        if (rowok && col < a.D) {
            ld_v8(xr + col, xv[c]);
            if (col + 8 < a.D) ld_v8(xr + col + 8, xv[c] + 8);
        }
    }
    mbar_wait(o_full, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int col = 16 * (first + 2 * c);
        if (col < a.D) {                                         // warp-uniform
            const bool wide = col + 8 < a.D;
            float bv[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = reinterpret_cast<const float4*>(bias_s + col)[j];
                bv[4 * j] = t.x; bv[4 * j + 1] = t.y; bv[4 * j + 2] = t.z; bv[4 * j + 3] = t.w;
            }
            uint32_t rr[16];
            tmem_ld16(lane_addr + SA_O_COL + (uint32_t)col, rr);
            if (rowok) {
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(rr[j]) + xv[c][j] + bv[j]);
                st_v8(orow + col, o);
                if (wide) st_v8(orow + col + 8, o + 8);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__global__ void __launch_bounds__(SA_THREADS, 1) seq_attention_fold_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                                           const __grid_constant__ CUtensorMap map_mid,
                                                                           const AttnArgs a) {
    extern __shared__ __align__(1024) unsigned char sa_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sa_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char* k_buf = base;                               // K_hi | K_mid
    unsigned char* v_buf = base + 2 * SA_K_PLANE;              // V'_hi | V'_mid
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_buf + 2 * SA_V_PLANE);
    uint64_t* k_full = bars;        // TMA landed
    uint64_t* k_empty = bars + 1;   // Q K^T retired
    uint64_t* v_full = bars + 2;
    uint64_t* v_empty = bars + 3;   // P V' retired
    uint64_t* q_full = bars + 4;    // 4 warps: Q of the step is in TMEM
    uint64_t* s_full = bars + 5;    // scores complete
    uint64_t* p_full = bars + 6;    // 4 warps: P written
    uint64_t* o_full = bars + 7;    // O' of the tile complete
    uint64_t* o_empty = bars + 8;   // 4 warps: O' drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* bias_s = reinterpret_cast<float*>(bars + 10);   // (SA_NV) b_o, zero padded: the L1 is swept by the Q / residual rows, so
                                                           // the epilogue's bias reads kept going to L2 on the critical path of the O' buffer

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    for (int c = tid; c < SA_NV; c += blockDim.x) bias_s[c] = c < a.D ? a.bias[c] : 0.f;
    if (tid == 0) {
        mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(q_full, 4); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_hi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_mid)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int64_t my_tiles = a.tiles > blockIdx.x ? (a.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t steps = my_tiles * a.H;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int64_t n = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int seq0 = (int)((blockIdx.x + it * gridDim.x) * a.NS);
                for (int h = 0; h < a.H; ++h, ++n) {
                    const uint32_t par = (uint32_t)(n & 1);
                    mbar_wait(k_empty, par ^ 1u);
                    mbar_expect_tx(k_full, 2u * SA_K_PLANE);
                    const int kc = a.k_col0 + h * a.hdk;
#pragma unroll
                    for (int j = 0; j < SA_KBLK; ++j) {
                        tma_load_3d(&map_hi, k_full, k_buf + j * SA_BLK, kc + 32 * j, 0, seq0);
                        tma_load_3d(&map_mid, k_full, k_buf + SA_K_PLANE + j * SA_BLK, kc + 32 * j, 0, seq0);
                    }
                    mbar_wait(v_empty, par ^ 1u);
                    mbar_expect_tx(v_full, 2u * SA_V_PLANE);
                    const int vc = a.v_col0 + h * a.D;
#pragma unroll
                    for (int j = 0; j < SA_VCH; ++j) {
                        tma_load_3d(&map_hi, v_full, v_buf + j * SA_BLK, vc + 32 * j, 0, seq0);
                        tma_load_3d(&map_mid, v_full, v_buf + SA_V_PLANE + j * SA_BLK, vc + 32 * j, 0, seq0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, elected issue)
        // instruction descriptors (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, N >> 3 at [17,23), M >> 4 at [24,29);
        // bit 16 = B operand MN-major (V' rows are keys)
        const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SA_ROWS >> 3) << 17) | ((uint32_t)(SA_ROWS >> 4) << 24);
        const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(SA_NV >> 3) << 17) | ((uint32_t)(SA_ROWS >> 4) << 24);
        const uint64_t kd_h = make_desc_sw64(smem_u32(k_buf)), kd_m = make_desc_sw64(smem_u32(k_buf + SA_K_PLANE));
        const uint64_t vd_h = make_desc_mn_sw64(smem_u32(v_buf), SA_BLK), vd_m = make_desc_mn_sw64(smem_u32(v_buf + SA_V_PLANE), SA_BLK);
        const uint32_t tq = tmem_base + SA_Q_COL, ts = tmem_base + SA_S_COL, to = tmem_base + SA_O_COL;
        int64_t n = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            for (int h = 0; h < a.H; ++h, ++n) {
                const uint32_t par = (uint32_t)(n & 1);
                mbar_wait(q_full, par);
                mbar_wait(k_full, par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < SA_KSTEPS; ++kk) {
                    const uint64_t o = (uint64_t)(((kk >> 1) * SA_BLK + (kk & 1) * 32) >> 4);
                    umma_ts_e(ts, tq + 8 * kk, kd_h + o, idesc_qk, kk != 0);
                    umma_ts_e(ts, tq + 8 * kk, kd_m + o, idesc_qk, 1);
                    umma_ts_e(ts, tq + 56 + 8 * kk, kd_h + o, idesc_qk, 1);
                }
                umma_commit_e(k_empty);
                umma_commit_e(s_full);
                mbar_wait(p_full, par);
                mbar_wait(v_full, par);
                if (h == 0) mbar_wait(o_empty, (uint32_t)((it & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < SA_ROWS / 16; ++kk) {
                    const uint64_t o = (uint64_t)((kk * 16 * 64) >> 4);                 // 16 keys = 16 rows of 64 bytes
                    umma_ts_e(to, ts + 8 * kk, vd_h + o, idesc_pv, (h | kk) != 0);
                    umma_ts_e(to, ts + 8 * kk, vd_m + o, idesc_pv, 1);
                    umma_ts_e(to, ts + 64 + 8 * kk, vd_h + o, idesc_pv, 1);
                }
                umma_commit_e(v_empty);
                if (h == a.H - 1) umma_commit_e(o_full);
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ softmax warps: thread = tile row
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int slot = r / a.SP, i = r - slot * a.SP;
        const int win = slot * a.SP;                            // first key column of this row's window (warp-uniform)
        int64_t n = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t seq = (blockIdx.x + it * gridDim.x) * a.NS + slot;
            const bool rowok = i < a.S && seq < a.B;
            for (int h = 0; h < a.H; ++h, ++n) {
                mbar_wait(s_full, (uint32_t)(n & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t sv[64];
                const uint32_t sa = lane_addr + SA_S_COL + (uint32_t)win;
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
                // P = p / sum as bf16 hi | mid over all 128 key columns (zero outside the window): 2 keys per 32-bit column,
                // 16 columns (32 keys) at a time
                const uint32_t pa = lane_addr + SA_S_COL;
                const int wc = win >> 1;                         // first column of the window inside a plane
#pragma unroll
                for (int c = 0; c < 4; ++c) {                    // 16-column pieces of a 64-column plane
                    const int rel = 16 * c - wc;                 // warp-uniform
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
            {
                const int64_t row = rowok ? seq * a.S + i : 0;
                sa_tile_epilogue(a, bias_s, lane_addr, rowok, a.x + row * a.ldx, a.out + row * a.ldo, 0, o_full, (uint32_t)(it & 1));
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
            }
        }
    } else {
        // ------------------------------------------------------------------ IO warps: Q rows -> TMEM, final epilogue of every tile
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int slot = r / a.SP, i = r - slot * a.SP;
        // Q row of step n (tile it, head h) -> registers: hdk bf16 per plane (zero padded by the projection), 16-byte loads
        auto load_q = [&](int64_t n, uint32_t (&qh)[56], uint32_t (&qm)[56]) {
            const int64_t it = n / a.H;
            const int h = (int)(n - it * a.H);
            const int64_t seq = (blockIdx.x + it * gridDim.x) * a.NS + slot;
            const bool ok = i < a.S && seq < a.B;
            const int64_t row = ok ? seq * a.S + i : 0;
            const uint4* ph = reinterpret_cast<const uint4*>(a.q_hi + row * a.ldp + a.q_col0 + h * a.hdk);
            const uint4* pm = reinterpret_cast<const uint4*>(a.q_mid + row * a.ldp + a.q_col0 + h * a.hdk);
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                uint4 vh = make_uint4(0u, 0u, 0u, 0u), vm = vh;
                if (ok && 8 * j < a.hdk) {
                    vh = __ldg(ph + j);
                    vm = __ldg(pm + j);
                }
                qh[4 * j] = vh.x; qh[4 * j + 1] = vh.y; qh[4 * j + 2] = vh.z; qh[4 * j + 3] = vh.w;
                qm[4 * j] = vm.x; qm[4 * j + 1] = vm.y; qm[4 * j + 2] = vm.z; qm[4 * j + 3] = vm.w;
            }
        };
        auto store_q = [&](const uint32_t (&qh)[56], const uint32_t (&qm)[56]) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                tmem_st8(lane_addr + SA_Q_COL + 8 * j, qh + 8 * j);
                tmem_st8(lane_addr + SA_Q_COL + 56 + 8 * j, qm + 8 * j);
            }
            tmem_st_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        };
        if (steps > 0) {
            uint32_t qh[56], qm[56];
            load_q(0, qh, qm);
            store_q(qh, qm);
        }
        int64_t n = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t seq = (blockIdx.x + it * gridDim.x) * a.NS + slot;
            const bool rowok = i < a.S && seq < a.B;
            const int64_t row = rowok ? seq * a.S + i : 0;
            const float* xr = a.x + row * a.ldx;
            float* orow = a.out + row * a.ldo;
            if (rowok) {
                for (int b = 0; b < a.D * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(xr) + b));
            }
            for (int h = 0; h < a.H; ++h, ++n) {
                uint32_t qh[56], qm[56];                         // scoped to the step: dead during the epilogue
                load_q(n + 1 < steps ? n + 1 : n, qh, qm);       // in flight while the scores of this step are computed
                mbar_wait(s_full, (uint32_t)(n & 1));            // Q K^T of this step has retired: the Q columns are free
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (n + 1 < steps) store_q(qh, qm);
            }
            // ---- final epilogue of the tile (the odd chunks; the softmax warp of this lane quarter takes the even ones)
            sa_tile_epilogue(a, bias_s, lane_addr, rowok, xr, orow, 1, o_full, (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (sequence, token, column) view of a (B * S, ld) bf16 plane; box = 32 columns x SP tokens x NS sequences, SWIZZLE_64B,
// tokens past S (and sequences past B) read as zeros
bool seq_tensor_map(const void* ptr, uint64_t cols, uint64_t ld, uint64_t S, uint64_t B, uint32_t SP, uint32_t NS, CUtensorMap* out) {
    static EncodeTiledFn3 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn3>(p);
    });
    if (!fn) {
        dyg_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return false;
    }
    const cuuint64_t dims[3] = {cols, S, B};
    const cuuint64_t strides[2] = {ld * 2, S * ld * 2};
    const cuuint32_t box[3] = {32, SP, NS};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dyg_set_error("cuTensorMapEncodeTiled (3-D) failed (%d): cols %llu, ld %llu, S %llu, B %llu", (int)r, (unsigned long long)cols,
                      (unsigned long long)ld, (unsigned long long)S, (unsigned long long)B);
        return false;
    }
    return true;
}

}  // namespace

extern "C" int dyg_seq_attention_fold(const void* planes_hi, const void* planes_mid, int ldp, int q_col0, int k_col0, int v_col0,
                                      int64_t B, int S, int H, int hd, int D, const float* x, int ldx, const float* bias, float* out,
                                      int ldo, dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && S > 0 && H > 0 && hd > 0 && D > 0, "dyg_seq_attention_fold: bad sizes");
    DYG_CHECK_ARG(S <= 64, "dyg_seq_attention_fold: S=%d unsupported (max 64)", S);
    DYG_CHECK_ARG(hd <= 112 && (hd % 4) == 0, "dyg_seq_attention_fold: head_dim=%d unsupported (multiple of 4, max 112)", hd);
    DYG_CHECK_ARG(D <= SA_NV && (D % 8) == 0, "dyg_seq_attention_fold: width %d unsupported (multiple of 8, max %d)", D, SA_NV);
    if (B == 0) return 0;
    DYG_CHECK_ARG(planes_hi && planes_mid && x && bias && out, "dyg_seq_attention_fold: NULL pointer");
    const int hdk = (hd + 7) / 8 * 8;
    const int cols = v_col0 + H * D;
    DYG_CHECK_ARG((ldp % 8) == 0 && ldp >= cols && aligned16(planes_hi) && aligned16(planes_mid),
                  "dyg_seq_attention_fold: planes must be 16-byte aligned with ldp %% 8 == 0 and ldp >= %d", cols);
    DYG_CHECK_ARG((q_col0 % 8) == 0 && (k_col0 % 8) == 0 && (v_col0 % 8) == 0 && q_col0 >= 0 && k_col0 >= q_col0 + H * hdk &&
                      v_col0 >= k_col0 + H * hdk,
                  "dyg_seq_attention_fold: segment offsets q=%d k=%d v=%d do not fit [q | k | v'] with 8-column alignment", q_col0, k_col0, v_col0);
    DYG_CHECK_ARG(B * (int64_t)S < ((int64_t)1 << 31), "dyg_seq_attention_fold: too many tokens");
    DYG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 31u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0 && (ldx % 8) == 0 && (ldo % 8) == 0 &&
                      aligned16(bias),
                  "dyg_seq_attention_fold: x / out must be 32-byte aligned with leading dimensions that are multiples of 8, bias 16-byte aligned");
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q_hi = reinterpret_cast<const __nv_bfloat16*>(planes_hi);
    a.q_mid = reinterpret_cast<const __nv_bfloat16*>(planes_mid);
    a.x = x; a.bias = bias; a.out = out;
    a.B = B; a.ldp = ldp; a.ldx = ldx; a.ldo = ldo;
    a.S = S; a.SP = S <= 32 ? 32 : 64; a.NS = SA_ROWS / a.SP;
    a.tiles = (B + a.NS - 1) / a.NS;
    a.H = H; a.hd = hd; a.hdk = hdk; a.D = D;
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0;
    CUtensorMap mh, mm;
    if (!seq_tensor_map(planes_hi, (uint64_t)cols, (uint64_t)ldp, (uint64_t)S, (uint64_t)B, (uint32_t)a.SP, (uint32_t)a.NS, &mh)) return 1;
    if (!seq_tensor_map(planes_mid, (uint64_t)cols, (uint64_t)ldp, (uint64_t)S, (uint64_t)B, (uint32_t)a.SP, (uint32_t)a.NS, &mm)) return 1;
    const size_t smem = (size_t)2 * SA_K_PLANE + 2 * SA_V_PLANE + 1024 + 256 + SA_NV * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(seq_attention_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_seq_attention_fold: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = true;
    }
    const int64_t sms = dyg_num_sms();
    const unsigned grid = (unsigned)(a.tiles < sms ? a.tiles : sms);
    seq_attention_fold_kernel<<<grid, SA_THREADS, smem, as_stream(stream)>>>(mh, mm, a);
    DYG_LAUNCH_CHECK("dyg_seq_attention_fold");
    return 0;
}
