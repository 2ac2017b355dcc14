// dyg_gemm_bf16x3: the dense contractions of DyGFormer's transformer (models/DyGFormer.py:442-461, 190-192) on tcgen05.
//
//   C[m, :N] = act(A W^T + bias + residual[m])      A (M,K), W (N,K), both handed over as BF16x3 operand pairs
//
// Precision: every fp32 value x travels as x = hi + mid, hi = bf16(x), mid = bf16(x - hi) (two bf16 planes, the same
// 4 bytes per element as fp32).  The product is A_hi W_hi + A_hi W_mid + A_mid W_hi with fp32 accumulation in TMEM;
// the dropped terms are O(2^-16) relative per product.  Producing kernels (LayerNorm, attention, the previous GEMM's
// epilogue) write the planes, so operands reach shared memory by TMA with no thread touching them.
//
// Structure: persistent CTA pairs (cluster of 2, tcgen05 cta_group::2), one CTA per SM, 320 threads each.
//   warp 0     TMA producer: per 32-wide k block A_hi | A_mid (this CTA's 128 rows) + W_hi | W_mid (this CTA's half of
//              the NT rows), 64-byte rows in SWIZZLE_64B layout, counted on the leader's full mbarrier (expect_tx)
//   warp 1     TMEM allocation (512 columns = two accumulator buffers); in the leader CTA one thread issues the
//              tcgen05.mma (M=256 over the pair, N=NT, K=16, kind::f16 bf16 -> fp32); tcgen05.commit (multicast to both
//              CTAs) frees the stage / publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld (32 lanes x 16 columns) -> bias / residual / activation -> 32-byte vector stores
//              of the fp32 result and / or its bf16 hi|mid planes; overlaps the next tile's MMAs (double-buffered TMEM)
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"
#include "gemm_epilogue.cuh"

namespace {

constexpr int G_BM = 128;            // rows per tile == TMEM lanes
constexpr int G_BK = 32;             // bf16 elements per stage along K == one 64-byte swizzle row
constexpr int G_EPI_WARPS = 8;         // two warps per TMEM lane quarter, interleaved over 16-column chunks
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;
constexpr int G_A_PLANE = G_BM * 64; // bytes of one A plane per stage
constexpr int G_MAX_STAGES = 8;
constexpr int G_TMEM_COLS = 512;
constexpr int G_BUF_COLS = 256;      // accumulator buffer stride in TMEM columns

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the CTA pair's leader is the even rank: clearing the peer bit of a shared::cluster address names the leader's copy
constexpr uint32_t G_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & G_PEER_MASK) : "memory");
}
// TMA load whose bytes are counted on the LEADER CTA's mbarrier (both CTAs of the pair feed one MMA)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & G_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// Warp-uniform issue: the whole (converged) warp runs the descriptor arithmetic, so the operands live in uniform registers,
// and elect.sync predicates the instruction itself onto one lane.  Issuing from inside `if (lane == 0)` instead makes the
// compiler move every operand with R2UR inside a divergence (ELECT / BRA.U.ANY) loop: ~15 extra instructions per MMA.
__device__ __forceinline__ void umma_bf16_pair_e(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair_e(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// Work decomposition: a CTA PAIR (cluster of 2, cta_group::2) owns 256-row super tiles; CTA r of the pair holds rows
// [256 ms + 128 r, +128) of A and rows [n0 + r NT/2, + NT/2) of W, the leader's thread issues M=256 MMAs that read both
// CTAs' shared memory and write each CTA's own TMEM.  Each W byte read from L2 thus serves 256 output rows.
// Two operand schedules:
//   resident (K <= 224, several n tiles): the pair keeps its A super tile in shared memory across the n tiles (per k
//             block slots with their own full / empty barriers, so the next super tile streams in behind the last n
//             tile's MMAs) and rings only W;
//   streamed: A and W blocks travel together through the ring.
__global__ void __launch_bounds__(G_THREADS, 1) gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_ah,
                                                                   const __grid_constant__ CUtensorMap map_am,
                                                                   const __grid_constant__ CUtensorMap map_wh,
                                                                   const __grid_constant__ CUtensorMap map_wm, const GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char gemm_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem) + 1023) & ~(uintptr_t)1023);
    const int nkb = (g.K + G_BK - 1) / G_BK;
    const bool resident = g.resident != 0;
    const int w_plane = (g.NT / 2) * 64;                       // this CTA's half of the W tile, one plane, one k block
    const int a_stage = 2 * G_A_PLANE;
    const int stage_bytes = (resident ? 0 : a_stage) + 2 * w_plane;
    unsigned char* a_res = base;                               // resident mode: nkb slots of [A_hi | A_mid]
    unsigned char* ring = base + (resident ? (size_t)nkb * a_stage : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)g.stages * stage_bytes);
    uint64_t* full_bar = bars;                                 // [stages]   leader: ring stage landed (both CTAs' bytes)
    uint64_t* empty_bar = bars + G_MAX_STAGES;                 // [stages]   both:   ring stage consumed
    uint64_t* afull_bar = bars + 2 * G_MAX_STAGES;             // [7]        leader: resident A slot landed
    uint64_t* aempty_bar = bars + 2 * G_MAX_STAGES + 8;        // [7]        both:   resident A slot consumed by the last n tile
    uint64_t* tfull_bar = bars + 2 * G_MAX_STAGES + 16;        // [2]        both:   accumulator buffer complete
    uint64_t* tempty_bar = bars + 2 * G_MAX_STAGES + 18;       // [2]        leader: accumulator buffer drained by both epilogues
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_MAX_STAGES + 20);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * G_MAX_STAGES + 22);   // the bias (when given): ~30 KB of L1 next to the ring cannot keep it

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform: role branches stay on the uniform path
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (g.bias)
        for (int c = tid; c < g.N; c += G_THREADS) bias_s[c] = g.bias[c];
    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int k = 0; k < 8; ++k) {
            mbar_init(afull_bar + k, 1);
            mbar_init(aempty_bar + k, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar + b, 1);
            mbar_init(tempty_bar + b, 2 * G_EPI_WARPS);        // one arrival per epilogue warp of either CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_am)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wm)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                         // barriers + TMEM of BOTH CTAs are ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one thread in each CTA)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, aph = 0;
            for (int64_t ms = pair; ms < g.m_super; ms += npairs, aph ^= 1u) {
                const int m0 = (int)ms * 2 * G_BM + (int)rank * G_BM;
                for (int nt = 0; nt < g.n_tiles; ++nt) {
                    const int n0 = nt * g.NS + (int)rank * (g.NT / 2);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (resident && nt == 0) {
                            mbar_wait(aempty_bar + kb, aph ^ 1u);
                            if (leader) mbar_expect_tx(afull_bar + kb, 2u * (uint32_t)a_stage);
                            unsigned char* as = a_res + (size_t)kb * a_stage;
                            tma_load_2d_pair(&map_ah, afull_bar + kb, as, kb * G_BK, m0);
                            tma_load_2d_pair(&map_am, afull_bar + kb, as + G_A_PLANE, kb * G_BK, m0);
                        }
                        mbar_wait(empty_bar + s, ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar + s, 2u * (uint32_t)stage_bytes);
                        unsigned char* st = ring + (size_t)s * stage_bytes;
                        if (!resident) {
                            tma_load_2d_pair(&map_ah, full_bar + s, st, kb * G_BK, m0);
                            tma_load_2d_pair(&map_am, full_bar + s, st + G_A_PLANE, kb * G_BK, m0);
                            st += a_stage;
                        }
                        tma_load_2d_pair(&map_wh, full_bar + s, st, kb * G_BK, n0);
                        tma_load_2d_pair(&map_wm, full_bar + s, st + w_plane, kb * G_BK, n0);
                        if (++s == g.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            // cute::UMMA::InstrDescriptor, kind::f16: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13), K-major A and B,
            // N >> 3 in [17,23), M >> 4 in [24,29) with M = 256 rows over the CTA pair
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.NT >> 3) << 17) | ((uint32_t)((2 * G_BM) >> 4) << 24);
            int s = 0;
            uint32_t ph = 0, aph = 0;
            int it = 0;
            for (int64_t ms = pair; ms < g.m_super; ms += npairs, aph ^= 1u) {
                for (int nt = 0; nt < g.n_tiles; ++nt, ++it) {
                    const int buf = it & 1;
                    const uint32_t bph = (uint32_t)((it >> 1) & 1);
                    mbar_wait(tempty_bar + buf, bph ^ 1u);              // both epilogues drained this accumulator buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tacc = tmem_base + (uint32_t)(buf * G_BUF_COLS);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (resident && nt == 0) mbar_wait(afull_bar + kb, aph);
                        mbar_wait(full_bar + s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        {
                            unsigned char* st = ring + (size_t)s * stage_bytes;
                            const uint64_t dah0 = make_desc_sw64(smem_u32(resident ? a_res + (size_t)kb * a_stage : st));
                            const uint64_t dam0 = dah0 + (uint64_t)(G_A_PLANE >> 4);
                            const uint64_t dbh0 = make_desc_sw64(smem_u32(st) + (resident ? 0u : (uint32_t)a_stage));
                            const uint64_t dbm0 = dbh0 + (uint64_t)(w_plane >> 4);
                            const int krem = g.K - kb * G_BK;
#pragma unroll
                            for (int kk = 0; kk < G_BK / 16; ++kk) {
                                if (kk * 16 < krem) {
                                    const uint64_t o = (uint64_t)(kk * 2);        // 16 bf16 = 32 bytes inside the swizzle row, >> 4
                                    umma_bf16_pair_e(tacc, dah0 + o, dbh0 + o, idesc, (kb | kk) != 0);
                                    umma_bf16_pair_e(tacc, dah0 + o, dbm0 + o, idesc, 1);
                                    umma_bf16_pair_e(tacc, dam0 + o, dbh0 + o, idesc, 1);
                                }
                            }
                            umma_commit_pair_e(empty_bar + s);                                   // ring stage reusable in both CTAs
                            if (resident && nt == g.n_tiles - 1) umma_commit_pair_e(aempty_bar + kb);   // A slot free for the next super tile
                            if (kb == nkb - 1) umma_commit_pair_e(tfull_bar + buf);              // accumulator complete in both CTAs
                        }
                        if (++s == g.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int quarter = warp & 3;                                      // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;
        GemmArgs eg = g;
        if (g.bias) eg.bias = bias_s;
        const int chunk0 = (warp - 2) >> 2;
        int it = 0;
        for (int64_t ms = pair; ms < g.m_super; ms += npairs) {
            const int64_t m = ms * 2 * G_BM + (int64_t)rank * G_BM + row;
            for (int nt = 0; nt < g.n_tiles; ++nt, ++it) {
                const int buf = it & 1;
                const uint32_t bph = (uint32_t)((it >> 1) & 1);
                const int n0 = nt * g.NS;
                const int ncols = min(g.NS, g.N - n0);
                if (g.residual && chunk0 == 0 && m < g.M) {
                    // this tile's MMAs are still running: pull its residual rows into L2 now, so the synchronous row
                    // reads of the epilogue see L2 latency instead of one DRAM round trip per 16-column chunk
                    const char* rp = reinterpret_cast<const char*>(g.residual + m * g.ldr + n0);
                    for (int b = 0; b < ncols * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + b));
                }
                mbar_wait(tfull_bar + buf, bph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * G_BUF_COLS);
                epilogue_tile(eg, taddr, m, n0, ncols, chunk0, G_EPI_WARPS / 4);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(tempty_bar + buf);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                         // no CTA leaves while its peer may still touch its smem / barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G_TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------ LayerNorm warps shared by the fused kernels
// The normalised rows of the NEXT tile are produced while the current tile computes: each of the four LN warps normalises
// 32 rows of its CTA's 128 (statistics and values from one read of x) and stores them as BF16x3 planes into the CTA's image
// of the A operand in global memory (L2 resident, already in the SWIZZLE_64B shared-memory layout).  When the MMA warp
// releases A, ONE bulk copy moves the 112 KB image into shared memory at engine speed; only that copy is exposed.
// Rows are handled eight at a time with the warp reductions of the eight rows interleaved step by step (a reduction is a
// chain of five dependent shuffles; one row after the other took ~25 us per tile with one LN warp per scheduler).
constexpr int LN_WARPS = 4;
constexpr int LN_KB = 7;                  // k blocks of the image (K <= 224)
struct LnSrc {
    const float* x;
    const float* gamma;
    const float* beta;
    int64_t M, m_super;
    int ldx, D;
    float eps;
};
__device__ __forceinline__ void ln_warps_loop(const LnSrc& f, int64_t pair, int64_t npairs, uint32_t rank, bool copier, int quarter, int lane,
                                              unsigned char* a_img, unsigned char* a_res, uint64_t* a_free, uint64_t* a_copy, uint64_t* a_full) {
    constexpr int RB = 8;                                               // rows per batch (16 float4 loads in flight per lane)
    constexpr int a_stage = 2 * G_A_PLANE;
    const int r0 = quarter * 32;                                        // this warp's 32 rows of the CTA's 128
    // lane l owns columns [8l, 8l+8) of a row: two 16-byte loads, and one 16-byte bf16 chunk per plane on the way out
    const int c0 = 8 * lane;
    const bool own = c0 < f.D;                                          // D % 8 == 0
    const int kb = lane >> 2;
    const uint32_t chunk = (uint32_t)(lane & 3);
    float gm8[8], bt8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gm8[j] = own ? __ldg(f.gamma + c0 + j) : 0.f;
        bt8[j] = own ? __ldg(f.beta + c0 + j) : 0.f;
    }
    const float invD = 1.f / (float)f.D;
    int it = 0;
    for (int64_t ms = pair; ms < f.m_super; ms += npairs, ++it) {
        const int64_t m0 = ms * 2 * G_BM + (int64_t)rank * G_BM;
        for (int i0 = 0; i0 < 32; i0 += RB) {
            float4 v[RB][2];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int64_t m = m0 + r0 + i0 + i;
                if (own && m < f.M) {
                    const float4* xp = reinterpret_cast<const float4*>(f.x + m * f.ldx + c0);
                    v[i][0] = xp[0];
                    v[i][1] = xp[1];
                } else {
                    v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    v[i][1] = v[i][0];
                }
            }
            float s[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) s[i] = ((v[i][0].x + v[i][0].y) + (v[i][0].z + v[i][0].w)) + ((v[i][1].x + v[i][1].y) + (v[i][1].z + v[i][1].w));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < RB; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            }
            float mean[RB], q[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                mean[i] = s[i] * invD;
                const float d0 = v[i][0].x - mean[i], d1 = v[i][0].y - mean[i], d2 = v[i][0].z - mean[i], d3 = v[i][0].w - mean[i];
                const float d4 = v[i][1].x - mean[i], d5 = v[i][1].y - mean[i], d6 = v[i][1].z - mean[i], d7 = v[i][1].w - mean[i];
                q[i] = own ? ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7)) : 0.f;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < RB; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
            }
            if (own) {
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const float rstd = rsqrtf(q[i] * invD + f.eps);
                    const float xs[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        split_pack((xs[2 * j] - mean[i]) * rstd * gm8[2 * j] + bt8[2 * j],
                                   (xs[2 * j + 1] - mean[i]) * rstd * gm8[2 * j + 1] + bt8[2 * j + 1], h[j], l[j]);
                    const int r = r0 + i0 + i;
                    unsigned char* p = a_img + kb * a_stage + (uint32_t)((r >> 3) * 512 + (r & 7) * 64) + ((chunk ^ (uint32_t)((r >> 1) & 3)) << 4);
                    *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(p + G_A_PLANE) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");               // generic global writes -> visible to the bulk copy engine
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LN_WARPS) : "memory");   // all four LN warps finished the image
        if (copier) {
            mbar_wait(a_free, (uint32_t)((it & 1) ^ 1));                 // last MMA reading A of the previous tile retired
            if (lane == 0) {
                mbar_expect_tx(a_copy, (uint32_t)(LN_KB * a_stage));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(a_res)),
                             "l"(a_img), "r"((uint32_t)(LN_KB * a_stage)), "r"(smem_u32(a_copy))
                             : "memory");
            }
            mbar_wait(a_copy, (uint32_t)(it & 1));
            if (lane == 0) mbar_arrive_leader(a_full);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LN_WARPS) : "memory");   // the image may be overwritten for the next tile
    }
}

// ====================================================================================================================
// dyg_ln_ffn_bf16x3: out = x + W2 gelu(W1 LayerNorm(x) + b1) + b2 — the feed-forward half of DyGFormer's transformer block
// (models/DyGFormer.py:456-461) as ONE kernel: the 4D-wide hidden activation never leaves the SM.
//
// Per CTA pair and 256-row super tile:
//   LN warps   LayerNorm of the pair's rows straight into the resident A operand (BF16x3 planes, SWIZZLE_64B k blocks):
//              row statistics are taken while the previous tile still computes, the normalised rows are written as soon
//              as the MMA warp releases A; the same warps run the final epilogue (+ b2 + x -> global)
//   per macro chunk of 128 hidden columns (7 for 800, the last one ragged):
//     GEMM1    acc1[mc&1] (TMEM, 128 cols) = LN(x) W1[mc]^T     (M=256, N=128, K=200: 13 k16 steps x 3 MMAs; W1 streamed
//              per 32-wide k block through the weight ring)
//     per 32-column sub chunk:
//       EPI1   8 warps (the two warps of a TMEM lane quarter take sub chunks round-robin): tcgen05.ld -> + b1 -> GELU ->
//              bf16 hi|mid -> h[.] in shared memory as the next A operand (3 buffers)
//       GEMM2  acc2 (TMEM, 208 cols) += h[.] W2[:, sub]^T        (M=256, N=208, K=32: 2 x 3 MMAs; W2 through the weight ring)
//   Weight ring: W1 k blocks (hi | mid, 8 KB) and the two planes of a W2 sub chunk (6.5 KB each) share ONE ring of eight 8 KB
//   slots, filled by the producer in the order the MMA warp consumes them.  With a 3-slot ring per weight the MMA-issuing thread
//   spent 15 % of the kernel waiting for W1 (a GEMM1 burst needs 7 k blocks, a slot turns around in ~2,000 cycles of TMA
//   latency against 384 cycles of MMAs) and 4 % for W2; the shared ring lets a burst take all the capacity: 1.56 -> 1.45 ms.
//   GEMM1 of macro chunk mc+1 is issued before the GEMM2s of mc, so the tensor pipe works while the GELU of mc runs.
// MMAs are sized so that each one is worth its issue cost (a first version with N=32 GEMM1s was bound by the issuing
// thread).  Barriers consumed by the leader's MMA warp live in the leader CTA (remote arrivals from the peer), barriers
// released by tcgen05.commit are multicast to both CTAs.
constexpr int F_SUB = 32;                 // hidden columns per sub chunk (K of one GEMM2 step)
constexpr int F_MC = 128;                 // hidden columns per macro chunk (N of GEMM1)
constexpr int F_KB1 = 7;                  // k blocks of GEMM1 (K <= 224)
constexpr int F_W1_PLANE = (F_MC / 2) * 64;
constexpr int F_W1_STAGE = 2 * F_W1_PLANE;   // one k block of this CTA's half of a macro chunk, hi | mid
constexpr int F_HB = 3;
constexpr int F_MAX_WS = 8;                // W1 k blocks and W2 sub chunks share ONE ring of equal slots, filled in the order the MMA warp consumes them
constexpr int F_EPI_WARPS = 8, F_LN_WARPS = 4;    // EPI warps per TMEM lane quarter: F_EPI_WARPS / 4, taking sub chunks round-robin
constexpr int F_THREADS = 64 + 32 * (F_EPI_WARPS + F_LN_WARPS);
constexpr int F_ACC1_COL = 256;           // acc2 at TMEM columns [0, 208), acc1 buffers at 256 and 384

struct FfnArgs {
    const float* x;        // (M, ldx) fp32 residual stream
    const float* gamma;
    const float* beta;
    const float* b1;
    const float* b2;
    float* out;            // (M, ldo)
    int64_t M;
    int64_t m_super;
    int ldx, ldo;
    int D;                 // model width (K of GEMM1, N of GEMM2), <= 208, even
    int Dff;               // hidden width, multiple of 32
    int NT2;               // MMA N of GEMM2 (D rounded up to 16)
    float eps;
    unsigned char* scratch;   // gridDim.x images of the A operand (F_KB1 * 16 KB each), L2 resident
    int wstages;              // slots of the shared W1 / W2 ring
};

__global__ void __launch_bounds__(F_THREADS, 1) ln_ffn_bf16x3_kernel(const __grid_constant__ CUtensorMap map_w1h,
                                                                     const __grid_constant__ CUtensorMap map_w1m,
                                                                     const __grid_constant__ CUtensorMap map_w2h,
                                                                     const __grid_constant__ CUtensorMap map_w2m, const FfnArgs f) {
    extern __shared__ __align__(1024) unsigned char ffn_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ffn_smem) + 1023) & ~(uintptr_t)1023);
    const int a_stage = 2 * G_A_PLANE;                         // one k block of A (or one h buffer): hi | mid
    const int w2_plane = (f.NT2 / 2) * 64;
    unsigned char* a_res = base;                               // [F_KB1] k blocks of LN(x)
    unsigned char* h_buf = a_res + F_KB1 * a_stage;            // [F_HB] hidden sub chunks
    constexpr int w_slot = F_W1_STAGE;                         // 8 KB: a W1 k block (hi | mid) or ONE plane of a W2 sub chunk (<= 6.5 KB)
    unsigned char* w_ring = h_buf + F_HB * a_stage;            // [f.wstages] slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + (size_t)f.wstages * w_slot);
    uint64_t* w_full = bars;               // [8] leader
    uint64_t* w_empty = bars + 8;          // [8] both
    uint64_t* h_full = bars + 16;          // [3] leader: h[b] written by both CTAs' EPI1 warps
    uint64_t* h_empty = bars + 19;         // [3] both:   GEMM2 finished reading h[b]
    uint64_t* acc1_full = bars + 22;       // [2] both
    uint64_t* acc1_empty = bars + 24;      // [2] leader: both CTAs' EPI1 warps drained acc1[b]
    uint64_t* a_full = bars + 26;          //     leader: LN(x) of the tile written by both CTAs
    uint64_t* a_free = bars + 27;          //     both:   last GEMM1 of the tile retired
    uint64_t* acc2_full = bars + 28;       //     both
    uint64_t* acc2_empty = bars + 29;      //     leader: final epilogue (EPI warps of both CTAs) drained acc2
    uint64_t* a_copy = bars + 30;          //     local:  bulk copy of the normalised tile into A landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);
    // b1 | b2 (b2 zero-padded to NT2): the epilogue warps read them per chunk, and the L1 (~30 KB next to 220 KB of shared memory) is
    // swept by the LN warps' row loads, so __ldg of the biases kept missing to L2 on the critical path of both epilogues
    float* b1_s = reinterpret_cast<float*>(bars + 32);
    float* b2_s = b1_s + f.Dff;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform: role branches stay on the uniform path
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nmacro = (f.Dff + F_MC - 1) / F_MC;
    const int nkb1 = (f.D + G_BK - 1) / G_BK;

    // zero this CTA's image of the A operand once: the K padding columns (D .. 32*nkb1) are never written again
    unsigned char* a_img = f.scratch + (size_t)blockIdx.x * (F_KB1 * a_stage);
    for (int i = tid; i < F_KB1 * a_stage / 16; i += F_THREADS) reinterpret_cast<uint4*>(a_img)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < f.Dff; i += F_THREADS) b1_s[i] = f.b1[i];
    for (int i = tid; i < f.NT2; i += F_THREADS) b2_s[i] = i < f.D ? f.b2[i] : 0.f;
    if (tid == 0) {
        for (int s = 0; s < F_MAX_WS; ++s) {
            mbar_init(w_full + s, 1);
            mbar_init(w_empty + s, 1);
        }
        for (int s = 0; s < 3; ++s) {
            mbar_init(h_full + s, 8); mbar_init(h_empty + s, 1);                 // one EPI warp per lane quarter and CTA writes a sub chunk
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc1_full + s, 1);
            mbar_init(acc1_empty + s, 2 * F_EPI_WARPS);
        }
        mbar_init(a_full, 2);                                   // one arrival per CTA, after its bulk copy landed
        mbar_init(a_copy, 1);
        mbar_init(a_free, 1);
        mbar_init(acc2_full, 1);
        mbar_init(acc2_empty, 2 * F_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w1h)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w1m)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w2h)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w2m)) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the zero fill above must be visible to the MMAs
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: W1 k blocks and W2 sub chunks
        if (lane == 0) {
            int ws = 0;
            uint32_t wp = 0;
            auto load_w1 = [&](int mc) {
                const int nmc = min(F_MC, f.Dff - mc * F_MC);
                const int row = mc * F_MC + (int)rank * (nmc / 2);
                for (int kb = 0; kb < nkb1; ++kb) {
                    mbar_wait(w_empty + ws, wp ^ 1u);
                    if (leader) mbar_expect_tx(w_full + ws, 2u * F_W1_STAGE);
                    unsigned char* d1 = w_ring + (size_t)ws * w_slot;
                    tma_load_2d_pair(&map_w1h, w_full + ws, d1, kb * G_BK, row);
                    tma_load_2d_pair(&map_w1m, w_full + ws, d1 + F_W1_PLANE, kb * G_BK, row);
                    if (++ws == f.wstages) { ws = 0; wp ^= 1u; }
                }
            };
            for (int64_t ms = pair; ms < f.m_super; ms += npairs) {
                load_w1(0);
                for (int mc = 0; mc < nmacro; ++mc) {
                    if (mc + 1 < nmacro) load_w1(mc + 1);                  // same order as the MMA warp consumes them
                    const int nsub = min(F_MC, f.Dff - mc * F_MC) / F_SUB;
                    for (int sub = 0; sub < nsub; ++sub) {
                        const int row2 = (int)rank * (f.NT2 / 2);
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {                   // hi plane, then mid plane: one slot each
                            mbar_wait(w_empty + ws, wp ^ 1u);
                            if (leader) mbar_expect_tx(w_full + ws, 2u * (uint32_t)w2_plane);
                            tma_load_2d_pair(pl ? &map_w2m : &map_w2h, w_full + ws, w_ring + (size_t)ws * w_slot, mc * F_MC + sub * F_SUB, row2);
                            if (++ws == f.wstages) { ws = 0; wp ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, elected issue)
        if (leader) {
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(f.NT2 >> 3) << 17) | ((uint32_t)((2 * G_BM) >> 4) << 24);
            const uint64_t ad_h = make_desc_sw64(smem_u32(a_res)), ad_m = ad_h + (uint64_t)(G_A_PLANE >> 4);
            int ws = 0;
            uint32_t wp = 0;
            int64_t gm = 0;      // global macro chunk counter: acc1 buffer gm & 1
            int64_t gs = 0;      // global sub chunk counter:   h buffer gs % F_HB
            int it = 0;
            auto gemm1 = [&](int64_t g1, int mc) {
                const int pb = (int)(g1 & 1);
                const int nmc = min(F_MC, f.Dff - mc * F_MC);
                const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nmc >> 3) << 17) | ((uint32_t)((2 * G_BM) >> 4) << 24);
                mbar_wait(acc1_empty + pb, (uint32_t)(((g1 >> 1) & 1) ^ 1));          // EPI1 drained the previous user of acc1[pb]
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(F_ACC1_COL + F_MC * pb);
                for (int kb = 0; kb < nkb1; ++kb) {
                    mbar_wait(w_full + ws, wp);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t wd_h = make_desc_sw64(smem_u32(w_ring + (size_t)ws * w_slot)), wd_m = wd_h + (uint64_t)(F_W1_PLANE >> 4);
                    const uint64_t ao = (uint64_t)((kb * a_stage) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        if (kb * G_BK + kk * 16 < f.D) {                               // this k16 step holds valid columns
                            const uint64_t o = (uint64_t)(kk * 2);
                            umma_bf16_pair_e(tacc, ad_h + ao + o, wd_h + o, idesc1, (kb | kk) != 0);
                            umma_bf16_pair_e(tacc, ad_h + ao + o, wd_m + o, idesc1, 1);
                            umma_bf16_pair_e(tacc, ad_m + ao + o, wd_h + o, idesc1, 1);
                        }
                    }
                    umma_commit_pair_e(w_empty + ws);
                    if (++ws == f.wstages) { ws = 0; wp ^= 1u; }
                }
                umma_commit_pair_e(acc1_full + pb);
            };
            for (int64_t ms = pair; ms < f.m_super; ms += npairs, ++it) {
                mbar_wait(a_full, (uint32_t)(it & 1));                                // LN(x) of this tile is in shared memory (both CTAs)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                gemm1(gm, 0);
                if (nmacro == 1) umma_commit_pair_e(a_free);
                for (int mc = 0; mc < nmacro; ++mc, ++gm) {
                    if (mc + 1 < nmacro) {
                        gemm1(gm + 1, mc + 1);
                        if (mc + 2 == nmacro) umma_commit_pair_e(a_free);             // A may be overwritten once the GEMM1s retire
                    }
                    const int nsub = min(F_MC, f.Dff - mc * F_MC) / F_SUB;
                    for (int sub = 0; sub < nsub; ++sub, ++gs) {
                        const int hb = (int)(gs % F_HB);
                        if (mc == 0 && sub == 0) mbar_wait(acc2_empty, (uint32_t)((it & 1) ^ 1));   // previous tile's final epilogue done
                        mbar_wait(h_full + hb, (uint32_t)((gs / F_HB) & 1));          // GELU(sub chunk) is in h[hb] in both CTAs
                        const int ws_h = ws;
                        mbar_wait(w_full + ws, wp);
                        if (++ws == f.wstages) { ws = 0; wp ^= 1u; }
                        const int ws_m = ws;
                        mbar_wait(w_full + ws, wp);
                        if (++ws == f.wstages) { ws = 0; wp ^= 1u; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t hd_h = make_desc_sw64(smem_u32(h_buf + hb * a_stage)), hd_m = hd_h + (uint64_t)(G_A_PLANE >> 4);
                        const uint64_t vd_h = make_desc_sw64(smem_u32(w_ring + (size_t)ws_h * w_slot)), vd_m = make_desc_sw64(smem_u32(w_ring + (size_t)ws_m * w_slot));
#pragma unroll
                        for (int kk = 0; kk < F_SUB / 16; ++kk) {
                            const uint64_t o = (uint64_t)(kk * 2);                    // 32 bytes >> 4
                            umma_bf16_pair_e(tmem_base, hd_h + o, vd_h + o, idesc2, (mc | sub | kk) != 0);
                            umma_bf16_pair_e(tmem_base, hd_h + o, vd_m + o, idesc2, 1);
                            umma_bf16_pair_e(tmem_base, hd_m + o, vd_h + o, idesc2, 1);
                        }
                        umma_commit_pair_e(w_empty + ws_h);
                        umma_commit_pair_e(w_empty + ws_m);
                        umma_commit_pair_e(h_empty + hb);
                        if (mc == nmacro - 1 && sub == nsub - 1) umma_commit_pair_e(acc2_full);
                    }
                }
            }
        }
    } else if (warp < 2 + F_EPI_WARPS) {
        // ------------------------------------------------------------------ EPI1: bias + GELU + split of every hidden sub chunk
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                                  // TMEM lane == tile row
        const int halfc = (warp - 2) >> 2;                                  // this warp's turn among the EPI warps of its lane quarter
        const uint32_t row_off = (uint32_t)((r >> 3) * 512 + (r & 7) * 64);
        const uint32_t swz = (uint32_t)((r >> 1) & 3);
        int64_t gm = 0, gs = 0;
        int it = 0;
        // 256-bit row accesses need 32-byte aligned rows (D % 8 == 0 holds): otherwise the generic tile epilogue is used
        const bool fast_out = ((f.ldx & 7) == 0) && ((f.ldo & 7) == 0) && (((reinterpret_cast<uintptr_t>(f.x) | reinterpret_cast<uintptr_t>(f.out)) & 31u) == 0) &&
                              ((reinterpret_cast<uintptr_t>(f.b2) & 3u) == 0);
        GemmArgs eg;
        eg.bias = f.b2; eg.residual = f.x; eg.C = f.out; eg.Chi = nullptr; eg.Cmid = nullptr;
        eg.M = f.M; eg.ldr = f.ldx; eg.ldc = f.ldo; eg.ldcs = 0; eg.N = f.D; eg.K = 0; eg.act = DYG_ACT_NONE;
        for (int64_t ms = pair; ms < f.m_super; ms += npairs, ++it) {
            for (int mc = 0; mc < nmacro; ++mc, ++gm) {
                const int pb = (int)(gm & 1);
                const int nsub = min(F_MC, f.Dff - mc * F_MC) / F_SUB;
                mbar_wait(acc1_full + pb, (uint32_t)((gm >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int sub = 0; sub < nsub; ++sub, ++gs) {
                    // the two warps of a TMEM lane quarter alternate over the sub chunks, each one converting all 32 columns
                    // of its 32 rows: half as many barrier waits, proxy fences and arrivals per converted element
                    if ((int)(gs % (F_EPI_WARPS / 4)) != halfc) continue;
                    const int hb = (int)(gs % F_HB);
                    uint32_t rr[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(F_ACC1_COL + F_MC * pb + sub * F_SUB), rr);
                    const float* bp = b1_s + mc * F_MC + sub * F_SUB;
                    uint32_t hi[16], mid[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 b2v = *reinterpret_cast<const float2*>(bp + 2 * j);
                        const float v0 = act_apply(__uint_as_float(rr[2 * j]) + b2v.x, DYG_ACT_GELU);
                        const float v1 = act_apply(__uint_as_float(rr[2 * j + 1]) + b2v.y, DYG_ACT_GELU);
                        split_pack(v0, v1, hi[j], mid[j]);
                    }
                    mbar_wait(h_empty + hb, (uint32_t)(((gs / F_HB) & 1) ^ 1));       // GEMM2 of sub chunk gs-3 no longer reads h[hb]
                    unsigned char* rowp = h_buf + hb * a_stage + row_off;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        *reinterpret_cast<uint4*>(rowp + (((uint32_t)c ^ swz) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                        *reinterpret_cast<uint4*>(rowp + G_A_PLANE + (((uint32_t)c ^ swz) << 4)) =
                            make_uint4(mid[4 * c], mid[4 * c + 1], mid[4 * c + 2], mid[4 * c + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(h_full + hb);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(acc1_empty + pb);
            }
            // final epilogue of the tile: these warps are idle until the next tile's first GEMM1 retires, the LN warps are
            // busy writing the next A operand
            {
                constexpr int STEP = F_EPI_WARPS / 4, PRE = 4;
                const int64_t m = ms * 2 * G_BM + (int64_t)rank * G_BM + r;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
                if (!fast_out) {
                    mbar_wait(acc2_full, (uint32_t)(it & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    epilogue_tile(eg, taddr, m, 0, f.D, halfc, STEP);
                } else {
                    // the residual rows do not depend on the MMAs: the first PRE chunks of this warp are requested before the
                    // accumulator is awaited, so their L2 / HBM latency hides behind the last GEMM2s
                    const bool rowok = m < f.M;
                    const float* xr = f.x + m * f.ldx;
                    float res[PRE][16];
#pragma unroll
                    for (int k = 0; k < PRE; ++k) {
                        const int col = 16 * (halfc + k * STEP);
#pragma unroll
                        for (int j = 0; j < 16; ++j) res[k][j] = 0.f;
                        if (rowok && col < f.D) {
                            ld_v8(xr + col, res[k]);
                            if (col + 8 < f.D) ld_v8(xr + col + 8, res[k] + 8);
                        }
                    }
                    mbar_wait(acc2_full, (uint32_t)(it & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    auto emit = [&](int col, const float (&rv)[16]) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + (uint32_t)col, rr);
                        if (rowok) {
                            uint32_t o[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                o[j] = __float_as_uint(__uint_as_float(rr[j]) + rv[j] + b2_s[col + j]);
                            float* dst = f.out + m * f.ldo + col;
                            st_v8(dst, o);
                            if (col + 8 < f.D) st_v8(dst + 8, o + 8);
                        }
                        __syncwarp();
                    };
#pragma unroll
                    for (int k = 0; k < PRE; ++k) {
                        const int col = 16 * (halfc + k * STEP);
                        if (col < f.D) emit(col, res[k]);                          // warp-uniform
                    }
                    for (int col = 16 * (halfc + PRE * STEP); col < f.D; col += 16 * STEP) {
                        float rv[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) rv[j] = 0.f;
                        if (rowok) {
                            ld_v8(xr + col, rv);
                            if (col + 8 < f.D) ld_v8(xr + col + 8, rv + 8);
                        }
                        emit(col, rv);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(acc2_empty);
        }
    } else {
        // ------------------------------------------------------------------ LN warps: LayerNorm -> resident A operand
        LnSrc ls;
        ls.x = f.x; ls.gamma = f.gamma; ls.beta = f.beta; ls.M = f.M; ls.m_super = f.m_super; ls.ldx = f.ldx; ls.D = f.D; ls.eps = f.eps;
        ln_warps_loop(ls, pair, npairs, rank, warp == 2 + F_EPI_WARPS, warp & 3, lane, a_img, a_res, a_free, a_copy, a_full);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G_TMEM_COLS) : "memory");
    }
}

// ====================================================================================================================
// dyg_ln_gemm_bf16x3: C = act(LayerNorm(x) W^T + bias) — the [q | k | v'] projection of DyGFormer's attention sub-block
// (models/DyGFormer.py:447-454) with its LayerNorm fused in: gemm_bf16x3_kernel's resident schedule (CTA pairs, the A super
// tile stays in shared memory across the n tiles, W streamed through the TMA ring, two accumulators in TMEM), but the A
// operand comes from the LN warps (ln_warps_loop) instead of a planes round trip through HBM.
constexpr int LG_THREADS = 64 + 32 * (G_EPI_WARPS + LN_WARPS);

struct LnGemmArgs {
    GemmArgs g;
    LnSrc ln;
    unsigned char* scratch;   // gridDim.x images of the A operand (LN_KB * 16 KB each), L2 resident
    int tma_out;              // planes leave through shared memory + TMA stores (epilogue_planes_tma_tile)
};
constexpr int LG_EPI_SLOT = 4096;
__host__ __device__ inline int lg_bias_bytes(int N) { return (N * 4 + 255) / 256 * 256; }

__global__ void __launch_bounds__(LG_THREADS, 1) ln_gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_wh,
                                                                       const __grid_constant__ CUtensorMap map_wm,
                                                                       const __grid_constant__ CUtensorMap map_ch,
                                                                       const __grid_constant__ CUtensorMap map_cm, const LnGemmArgs p) {
    extern __shared__ __align__(1024) unsigned char lg_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lg_smem) + 1023) & ~(uintptr_t)1023);
    const GemmArgs& g = p.g;
    const int nkb = (g.K + G_BK - 1) / G_BK;
    const int w_plane = (g.NT / 2) * 64;
    const int a_stage = 2 * G_A_PLANE;
    const int stage_bytes = 2 * w_plane;
    unsigned char* a_res = base;                               // LN_KB slots of [A_hi | A_mid]
    unsigned char* epi_stage = base + (size_t)LN_KB * a_stage;  // one 4 KB slot per epilogue warp (planes on their way to the TMA store)
    unsigned char* ring = epi_stage + (p.tma_out ? G_EPI_WARPS * LG_EPI_SLOT : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)g.stages * stage_bytes);
    uint64_t* full_bar = bars;                                 // [stages]   leader
    uint64_t* empty_bar = bars + G_MAX_STAGES;                 // [stages]   both
    uint64_t* tfull_bar = bars + 2 * G_MAX_STAGES;             // [2]        both
    uint64_t* tempty_bar = bars + 2 * G_MAX_STAGES + 2;        // [2]        leader
    uint64_t* a_full = bars + 2 * G_MAX_STAGES + 4;            //            leader: LN(x) of the tile is in both CTAs' shared memory
    uint64_t* a_free = bars + 2 * G_MAX_STAGES + 5;            //            both:   last MMA of the tile retired
    uint64_t* a_copy = bars + 2 * G_MAX_STAGES + 6;            //            local:  bulk copy of the image landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_MAX_STAGES + 7);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * G_MAX_STAGES + 8);   // the bias (tma_out): read per chunk by every epilogue warp
    static_assert(((2 * G_MAX_STAGES + 8) * 8) % 16 == 0 && (2 * G_MAX_STAGES + 8) * 8 <= 512, "barrier block layout: bias_s is read with 16-byte loads");

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    // zero this CTA's image of the A operand once: the K padding columns (D .. 32 nkb) are never written again
    unsigned char* a_img = p.scratch + (size_t)blockIdx.x * (LN_KB * a_stage);
    for (int i = tid; i < LN_KB * a_stage / 16; i += LG_THREADS) reinterpret_cast<uint4*>(a_img)[i] = make_uint4(0, 0, 0, 0);
    if (p.tma_out)
        for (int i = tid; i < g.N; i += LG_THREADS) bias_s[i] = g.bias[i];
    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar + b, 1);
            mbar_init(tempty_bar + b, 2 * G_EPI_WARPS);
        }
        mbar_init(a_full, 2);
        mbar_init(a_free, 1);
        mbar_init(a_copy, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wm)) : "memory");
        if (p.tma_out) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ch)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_cm)) : "memory");
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");            // the zero fill above must be visible to the bulk copy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: W k blocks of every n tile
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t ms = pair; ms < g.m_super; ms += npairs) {
                for (int nt = 0; nt < g.n_tiles; ++nt) {
                    const int n0 = nt * g.NS + (int)rank * (g.NT / 2);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(empty_bar + s, ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar + s, 2u * (uint32_t)stage_bytes);
                        unsigned char* st = ring + (size_t)s * stage_bytes;
                        tma_load_2d_pair(&map_wh, full_bar + s, st, kb * G_BK, n0);
                        tma_load_2d_pair(&map_wm, full_bar + s, st + w_plane, kb * G_BK, n0);
                        if (++s == g.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, elected issue)
        if (leader) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.NT >> 3) << 17) | ((uint32_t)((2 * G_BM) >> 4) << 24);
            const uint64_t ad_h = make_desc_sw64(smem_u32(a_res)), ad_m = ad_h + (uint64_t)(G_A_PLANE >> 4);
            int s = 0;
            uint32_t ph = 0;
            int it = 0, tile = 0;
            for (int64_t ms = pair; ms < g.m_super; ms += npairs, ++tile) {
                mbar_wait(a_full, (uint32_t)(tile & 1));                         // LN(x) of this super tile is in shared memory (both CTAs)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int nt = 0; nt < g.n_tiles; ++nt, ++it) {
                    const int buf = it & 1;
                    const uint32_t bph = (uint32_t)((it >> 1) & 1);
                    mbar_wait(tempty_bar + buf, bph ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tacc = tmem_base + (uint32_t)(buf * G_BUF_COLS);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(full_bar + s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        unsigned char* st = ring + (size_t)s * stage_bytes;
                        const uint64_t wd_h = make_desc_sw64(smem_u32(st)), wd_m = wd_h + (uint64_t)(w_plane >> 4);
                        const uint64_t ao = (uint64_t)((kb * a_stage) >> 4);
#pragma unroll
                        for (int kk = 0; kk < G_BK / 16; ++kk) {
                            if (kb * G_BK + kk * 16 < g.K) {
                                const uint64_t o = (uint64_t)(kk * 2);
                                umma_bf16_pair_e(tacc, ad_h + ao + o, wd_h + o, idesc, (kb | kk) != 0);
                                umma_bf16_pair_e(tacc, ad_h + ao + o, wd_m + o, idesc, 1);
                                umma_bf16_pair_e(tacc, ad_m + ao + o, wd_h + o, idesc, 1);
                            }
                        }
                        umma_commit_pair_e(empty_bar + s);
                        if (kb == nkb - 1) {
                            if (nt == g.n_tiles - 1) umma_commit_pair_e(a_free);     // A may be overwritten once these MMAs retire
                            umma_commit_pair_e(tfull_bar + buf);
                        }
                        if (++s == g.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp < 2 + G_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int chunk0 = (warp - 2) >> 2;
        int it = 0;
        for (int64_t ms = pair; ms < g.m_super; ms += npairs) {
            const int64_t m = ms * 2 * G_BM + (int64_t)rank * G_BM + row;
            for (int nt = 0; nt < g.n_tiles; ++nt, ++it) {
                const int buf = it & 1;
                const uint32_t bph = (uint32_t)((it >> 1) & 1);
                const int n0 = nt * g.NS;
                const int ncols = min(g.NS, g.N - n0);
                mbar_wait(tfull_bar + buf, bph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * G_BUF_COLS);
                if (p.tma_out) {
                    epilogue_planes_tma_tile(g, bias_s, &map_ch, &map_cm, epi_stage + (warp - 2) * LG_EPI_SLOT, taddr, lane, m, n0, ncols, chunk0,
                                             G_EPI_WARPS / 4, [&]() {
                                                 asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                                                 __syncwarp();
                                                 if (lane == 0) mbar_arrive_leader(tempty_bar + buf);
                                             });
                } else {
                    epilogue_tile(g, taddr, m, n0, ncols, chunk0, G_EPI_WARPS / 4);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty_bar + buf);
                }
            }
        }
        if (p.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // ------------------------------------------------------------------ LN warps: LayerNorm -> resident A operand
        ln_warps_loop(p.ln, pair, npairs, rank, warp == 2 + G_EPI_WARPS, warp & 3, lane, a_img, a_res, a_free, a_copy, a_full);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)G_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ fp32 -> bf16 hi | mid planes
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, int64_t M, int D, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ mid, int ld) {
    const int64_t pairs = (int64_t)(D + 1) / 2;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M * pairs) return;
    const int64_t m = i / pairs;
    const int c = (int)(i - m * pairs) * 2;
    const float a = x[m * ldx + c];
    const float b = (c + 1 < D) ? x[m * ldx + c + 1] : 0.f;
    uint32_t h, l;
    split_pack(a, b, h, l);
    if (c + 1 < D && ((ld & 1) == 0)) {
        *reinterpret_cast<uint32_t*>(hi + m * ld + c) = h;
        *reinterpret_cast<uint32_t*>(mid + m * ld + c) = l;
    } else {
        hi[m * ld + c] = __ushort_as_bfloat16((unsigned short)(h & 0xFFFF));
        mid[m * ld + c] = __ushort_as_bfloat16((unsigned short)(l & 0xFFFF));
        if (c + 1 < D) {
            hi[m * ld + c + 1] = __ushort_as_bfloat16((unsigned short)(h >> 16));
            mid[m * ld + c + 1] = __ushort_as_bfloat16((unsigned short)(l >> 16));
        }
    }
}

// ------------------------------------------------------------------ LayerNorm -> bf16 hi | mid planes
// One warp per row; lane l owns column pairs (2l, 2l+1), (2l+64, 2l+65), ...; two-pass mean / biased variance like
// torch.nn.functional.layer_norm; the fp32 result is split on the way out (and optionally also stored as fp32).
template <int MAXP>
__global__ void layernorm_split_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid, int ld, int64_t M, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    float2 v[MAXP];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int c = 2 * lane + 64 * i;
        float2 t = make_float2(0.f, 0.f);
        if (c < D) t = *reinterpret_cast<const float2*>(x + m * ldx + c);
        v[i] = t;
        sum += t.x + t.y;
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int c = 2 * lane + 64 * i;
        if (c < D) {
            const float dx = v[i].x - mean, dy = v[i].y - mean;
            sq += dx * dx + dy * dy;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int c = 2 * lane + 64 * i;
        if (c < D) {
            const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + c));
            const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + c));
            const float a = (v[i].x - mean) * rstd * gm.x + bt.x;
            const float b = (v[i].y - mean) * rstd * gm.y + bt.y;
            if (y) *reinterpret_cast<float2*>(y + m * ldy + c) = make_float2(a, b);
            uint32_t h, l;
            split_pack(a, b, h, l);
            *reinterpret_cast<uint32_t*>(hi + m * ld + c) = h;
            *reinterpret_cast<uint32_t*>(mid + m * ld + c) = l;
        }
    }
}

}  // namespace

namespace {

// ------------------------------------------------------------------ tensor maps (driver entry point fetched at run time:
// the library links against cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ k.rows;
        h = h * 1000003u ^ k.cols;
        h = h * 1000003u ^ k.ld;
        h = h * 1000003u ^ k.box_rows;
        return h;
    }
};

// (rows, cols) bf16 row-major with leading dimension ld elements; box = box_rows x 32 columns, SWIZZLE_64B, OOB -> 0.
// Encoding depends only on the key, so maps are cached (a descriptor holds the address, not the data).
}  // namespace

bool dyg_tensor_map_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, ld, box_rows};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return true;
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        dyg_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return false;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)G_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dyg_set_error("cuTensorMapEncodeTiled failed (%d) for %llu x %llu, ld %llu, box %u", (int)r,
                      (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
        return false;
    }
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}


extern "C" int dyg_gemm_bf16x3(const void* A_hi, const void* A_mid, int lda, const void* W_hi, const void* W_mid, int ldw,
                               const float* bias, const float* residual, int ldr, float* C, int ldc, void* C_hi, void* C_mid,
                               int ldcs, int64_t M, int N, int K, int act, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && K > 0, "dyg_gemm_bf16x3: bad sizes");
    DYG_CHECK_ARG(M < ((int64_t)1 << 31) - G_BM, "dyg_gemm_bf16x3: M too large");
    DYG_CHECK_ARG(act >= DYG_ACT_NONE && act <= DYG_ACT_SIGMOID, "dyg_gemm_bf16x3: unknown activation %d", act);
    if (M == 0) return 0;
    DYG_CHECK_ARG(A_hi && A_mid && W_hi && W_mid, "dyg_gemm_bf16x3: operand planes must not be NULL");
    DYG_CHECK_ARG(aligned16(A_hi) && aligned16(A_mid) && aligned16(W_hi) && aligned16(W_mid),
                  "dyg_gemm_bf16x3: operand planes must be 16-byte aligned");
    DYG_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && lda >= K && ldw >= K,
                  "dyg_gemm_bf16x3: lda=%d / ldw=%d must be multiples of 8 and >= K=%d", lda, ldw, K);
    DYG_CHECK_ARG(C || (C_hi && C_mid), "dyg_gemm_bf16x3: no output given");
    DYG_CHECK_ARG((C_hi == nullptr) == (C_mid == nullptr), "dyg_gemm_bf16x3: C_hi and C_mid go together");
    DYG_CHECK_ARG(!C_hi || (ldcs % 2) == 0, "dyg_gemm_bf16x3: ldcs must be even");
    if (M == 0) return 0;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.bias = bias; g.residual = residual; g.ldr = ldr;
    g.C = C; g.ldc = ldc;
    g.Chi = reinterpret_cast<__nv_bfloat16*>(C_hi); g.Cmid = reinterpret_cast<__nv_bfloat16*>(C_mid); g.ldcs = ldcs;
    g.M = M; g.N = N; g.K = K; g.act = act;
    g.m_tiles = (M + G_BM - 1) / G_BM;
    g.m_super = (M + 2 * G_BM - 1) / (2 * G_BM);
    g.n_tiles = (N + 207) / 208;                                   // <= 208 columns per tile: two tiles' accumulators fit TMEM
    g.NS = ((N + g.n_tiles - 1) / g.n_tiles + 15) / 16 * 16;       // tile starts stay 32-byte aligned in fp32 and in bf16 rows
    g.NT = g.NS;
    const int nkb = (K + G_BK - 1) / G_BK;
    g.resident = (nkb <= 7 && g.n_tiles > 1) ? 1 : 0;
    const int w_stage = 2 * (g.NT / 2) * 64;
    const int stage_bytes = (g.resident ? 0 : 2 * G_A_PLANE) + w_stage;
    const int fixed = (g.resident ? nkb * 2 * G_A_PLANE : 0) + 1024 + 512 + (bias ? (N * 4 + 15) / 16 * 16 : 0);
    const int max_smem = 227 * 1024;
    g.stages = (max_smem - fixed) / stage_bytes;
    if (g.stages > G_MAX_STAGES) g.stages = G_MAX_STAGES;
    DYG_CHECK_ARG(g.stages >= 2, "dyg_gemm_bf16x3: tile does not fit shared memory");
    const size_t smem = (size_t)g.stages * stage_bytes + fixed;
    CUtensorMap mah, mam, mwh, mwm;
    if (!dyg_tensor_map_bf16(A_hi, (uint64_t)M, (uint64_t)K, (uint64_t)lda, G_BM, &mah)) return 1;
    if (!dyg_tensor_map_bf16(A_mid, (uint64_t)M, (uint64_t)K, (uint64_t)lda, G_BM, &mam)) return 1;
    if (!dyg_tensor_map_bf16(W_hi, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(g.NT / 2), &mwh)) return 1;
    if (!dyg_tensor_map_bf16(W_mid, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(g.NT / 2), &mwm)) return 1;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_gemm_bf16x3: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    const int max_pairs = dyg_num_sms() / 2;
    const int pairs = (int)(g.m_super < max_pairs ? g.m_super : max_pairs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_bf16x3_kernel, mah, mam, mwh, mwm, g);
    if (le != cudaSuccess) {
        dyg_set_error("dyg_gemm_bf16x3: launch failed: %s", cudaGetErrorString(le));
        return 1;
    }
    DYG_LAUNCH_CHECK("dyg_gemm_bf16x3");
    return 0;
}

extern "C" int dyg_split_bf16(const float* x, int ldx, int64_t M, int D, void* hi, void* mid, int ld, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0 && ld >= D && ldx >= D, "dyg_split_bf16: bad sizes");
    if (M == 0) return 0;
    DYG_CHECK_ARG(x && hi && mid, "dyg_split_bf16: NULL pointer");
    const int64_t n = M * ((D + 1) / 2);
    split_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, M, D, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                             reinterpret_cast<__nv_bfloat16*>(mid), ld);
    DYG_LAUNCH_CHECK("dyg_split_bf16");
    return 0;
}

extern "C" int dyg_layernorm_split(const float* x, int ldx, const float* gamma, const float* beta, float eps, float* y, int ldy,
                                   void* hi, void* mid, int ld, int64_t M, int D, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0 && D <= 1024 && (D % 2) == 0, "dyg_layernorm_split: D=%d unsupported (even, max 1024)", D);
    DYG_CHECK_ARG((ldx % 2) == 0 && (ld % 2) == 0 && (!y || (ldy % 2) == 0), "dyg_layernorm_split: leading dims must be even");
    if (M == 0) return 0;
    DYG_CHECK_ARG(x && hi && mid && gamma && beta, "dyg_layernorm_split: NULL pointer");
    DYG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 7u) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 7u) == 0 &&
                      (reinterpret_cast<uintptr_t>(beta) & 7u) == 0 && (!y || (reinterpret_cast<uintptr_t>(y) & 7u) == 0),
                  "dyg_layernorm_split: fp32 pointers must be 8-byte aligned");
    if (M == 0) return 0;
    const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
    cudaStream_t s = as_stream(stream);
    __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
    __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(mid);
    if (D <= 256) layernorm_split_kernel<4><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, eps, y, ldy, h, l, ld, M, D);
    else if (D <= 512) layernorm_split_kernel<8><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, eps, y, ldy, h, l, ld, M, D);
    else layernorm_split_kernel<16><<<blocks, 256, 0, s>>>(x, ldx, gamma, beta, eps, y, ldy, h, l, ld, M, D);
    DYG_LAUNCH_CHECK("dyg_layernorm_split");
    return 0;
}

extern "C" int64_t dyg_ln_ffn_workspace_bytes(void) { return (int64_t)dyg_num_sms() * F_KB1 * 2 * G_A_PLANE; }

extern "C" int dyg_ln_ffn_bf16x3(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W1_hi,
                                 const void* W1_mid, int ldw1, const float* b1, const void* W2_hi, const void* W2_mid, int ldw2,
                                 const float* b2, float* out, int ldo, int64_t M, int D, int Dff, void* workspace,
                                 int64_t workspace_bytes, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0 && Dff > 0, "dyg_ln_ffn_bf16x3: bad sizes");
    DYG_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= dyg_ln_ffn_workspace_bytes(),
                  "dyg_ln_ffn_bf16x3: workspace of %lld bytes (16-byte aligned) required", (long long)dyg_ln_ffn_workspace_bytes());
    DYG_CHECK_ARG(D <= 208 && (D % 8) == 0, "dyg_ln_ffn_bf16x3: model width %d unsupported (multiple of 8, <= 208)", D);
    DYG_CHECK_ARG((Dff % F_SUB) == 0, "dyg_ln_ffn_bf16x3: hidden width %d must be a multiple of %d", Dff, F_SUB);
    if (M == 0) return 0;
    DYG_CHECK_ARG(x && gamma && beta && W1_hi && W1_mid && b1 && W2_hi && W2_mid && b2 && out, "dyg_ln_ffn_bf16x3: NULL pointer");
    DYG_CHECK_ARG((ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 7u) == 0 &&
                      (reinterpret_cast<uintptr_t>(beta) & 7u) == 0 && (reinterpret_cast<uintptr_t>(b1) & 7u) == 0,
                  "dyg_ln_ffn_bf16x3: x must be 16-byte aligned with ldx %% 4 == 0, the parameter vectors 8-byte aligned");
    DYG_CHECK_ARG((ldw1 % 8) == 0 && ldw1 >= D && (ldw2 % 8) == 0 && ldw2 >= Dff && aligned16(W1_hi) && aligned16(W1_mid) &&
                      aligned16(W2_hi) && aligned16(W2_mid),
                  "dyg_ln_ffn_bf16x3: weight planes must be 16-byte aligned with leading dimensions that are multiples of 8");
    DYG_CHECK_ARG(M < ((int64_t)1 << 31) - 2 * G_BM, "dyg_ln_ffn_bf16x3: M too large");
    if (M == 0) return 0;
    FfnArgs f;
    memset(&f, 0, sizeof(f));
    f.x = x; f.gamma = gamma; f.beta = beta; f.b1 = b1; f.b2 = b2; f.out = out;
    f.M = M; f.m_super = (M + 2 * G_BM - 1) / (2 * G_BM);
    f.ldx = ldx; f.ldo = ldo; f.D = D; f.Dff = Dff; f.NT2 = (D + 15) / 16 * 16; f.eps = eps;
    f.scratch = reinterpret_cast<unsigned char*>(workspace);
    CUtensorMap m1h, m1m, m2h, m2m;
    if (!dyg_tensor_map_bf16(W1_hi, (uint64_t)Dff, (uint64_t)D, (uint64_t)ldw1, F_MC / 2, &m1h)) return 1;
    if (!dyg_tensor_map_bf16(W1_mid, (uint64_t)Dff, (uint64_t)D, (uint64_t)ldw1, F_MC / 2, &m1m)) return 1;
    if (!dyg_tensor_map_bf16(W2_hi, (uint64_t)D, (uint64_t)Dff, (uint64_t)ldw2, (uint32_t)(f.NT2 / 2), &m2h)) return 1;
    if (!dyg_tensor_map_bf16(W2_mid, (uint64_t)D, (uint64_t)Dff, (uint64_t)ldw2, (uint32_t)(f.NT2 / 2), &m2m)) return 1;
    const int w_slot = F_W1_STAGE;
    const size_t fixed = (size_t)(F_KB1 + F_HB) * 2 * G_A_PLANE + 1024 + 512 + (size_t)(Dff + f.NT2) * 4;
    f.wstages = (int)((227 * 1024 - fixed) / w_slot);
    if (f.wstages > F_MAX_WS) f.wstages = F_MAX_WS;
    DYG_CHECK_ARG(f.wstages >= 3, "dyg_ln_ffn_bf16x3: weight ring does not fit shared memory");
    const size_t smem = fixed + (size_t)f.wstages * w_slot;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ln_ffn_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_ln_ffn_bf16x3: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    const int max_pairs = dyg_num_sms() / 2;
    const int pairs = (int)(f.m_super < max_pairs ? f.m_super : max_pairs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, ln_ffn_bf16x3_kernel, m1h, m1m, m2h, m2m, f);
    if (le != cudaSuccess) {
        dyg_set_error("dyg_ln_ffn_bf16x3: launch failed: %s", cudaGetErrorString(le));
        return 1;
    }
    return 0;
}

extern "C" int dyg_ln_gemm_bf16x3(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W_hi,
                                  const void* W_mid, int ldw, const float* bias, float* C, int ldc, void* C_hi, void* C_mid, int ldcs,
                                  int64_t M, int N, int D, int act, void* workspace, int64_t workspace_bytes, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && D > 0, "dyg_ln_gemm_bf16x3: bad sizes");
    DYG_CHECK_ARG(workspace && aligned16(workspace) && workspace_bytes >= dyg_ln_ffn_workspace_bytes(),
                  "dyg_ln_gemm_bf16x3: workspace of %lld bytes (16-byte aligned) required", (long long)dyg_ln_ffn_workspace_bytes());
    DYG_CHECK_ARG(D <= 224 && (D % 8) == 0, "dyg_ln_gemm_bf16x3: width %d unsupported (multiple of 8, <= 224)", D);
    DYG_CHECK_ARG(act >= DYG_ACT_NONE && act <= DYG_ACT_SIGMOID, "dyg_ln_gemm_bf16x3: unknown activation %d", act);
    DYG_CHECK_ARG(M < ((int64_t)1 << 31) - 2 * G_BM, "dyg_ln_gemm_bf16x3: M too large");
    if (M == 0) return 0;
    DYG_CHECK_ARG(x && gamma && beta && W_hi && W_mid, "dyg_ln_gemm_bf16x3: NULL pointer");
    DYG_CHECK_ARG((ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 3u) == 0 &&
                      (reinterpret_cast<uintptr_t>(beta) & 3u) == 0,
                  "dyg_ln_gemm_bf16x3: x must be 16-byte aligned with ldx %% 4 == 0");
    DYG_CHECK_ARG((ldw % 8) == 0 && ldw >= D && aligned16(W_hi) && aligned16(W_mid),
                  "dyg_ln_gemm_bf16x3: weight planes must be 16-byte aligned with ldw a multiple of 8");
    DYG_CHECK_ARG(C || (C_hi && C_mid), "dyg_ln_gemm_bf16x3: no output given");
    DYG_CHECK_ARG((C_hi == nullptr) == (C_mid == nullptr), "dyg_ln_gemm_bf16x3: C_hi and C_mid go together");
    DYG_CHECK_ARG(!C_hi || (ldcs % 2) == 0, "dyg_ln_gemm_bf16x3: ldcs must be even");
    LnGemmArgs p;
    memset(&p, 0, sizeof(p));
    GemmArgs& g = p.g;
    g.bias = bias; g.C = C; g.ldc = ldc;
    g.Chi = reinterpret_cast<__nv_bfloat16*>(C_hi); g.Cmid = reinterpret_cast<__nv_bfloat16*>(C_mid); g.ldcs = ldcs;
    g.M = M; g.N = N; g.K = D; g.act = act;
    g.m_tiles = (M + G_BM - 1) / G_BM;
    g.m_super = (M + 2 * G_BM - 1) / (2 * G_BM);
    g.n_tiles = (N + 207) / 208;
    g.NS = ((N + g.n_tiles - 1) / g.n_tiles + 15) / 16 * 16;
    g.NT = g.NS;
    g.resident = 1;
    const int stage_bytes = 2 * (g.NT / 2) * 64;
    p.tma_out = planes_fast_path(g);
    const int fixed = LN_KB * 2 * G_A_PLANE + (p.tma_out ? G_EPI_WARPS * LG_EPI_SLOT + lg_bias_bytes(N) : 0) + 1024 + 512;
    g.stages = (227 * 1024 - fixed) / stage_bytes;
    if (g.stages > G_MAX_STAGES) g.stages = G_MAX_STAGES;
    DYG_CHECK_ARG(g.stages >= 2, "dyg_ln_gemm_bf16x3: tile does not fit shared memory");
    p.ln.x = x; p.ln.gamma = gamma; p.ln.beta = beta; p.ln.M = M; p.ln.m_super = g.m_super; p.ln.ldx = ldx; p.ln.D = D; p.ln.eps = eps;
    p.scratch = reinterpret_cast<unsigned char*>(workspace);
    const size_t smem = (size_t)g.stages * stage_bytes + fixed;
    CUtensorMap mwh, mwm;
    if (!dyg_tensor_map_bf16(W_hi, (uint64_t)N, (uint64_t)D, (uint64_t)ldw, (uint32_t)(g.NT / 2), &mwh)) return 1;
    if (!dyg_tensor_map_bf16(W_mid, (uint64_t)N, (uint64_t)D, (uint64_t)ldw, (uint32_t)(g.NT / 2), &mwm)) return 1;
    CUtensorMap mch = mwh, mcm = mwm;
    if (p.tma_out) {
        if (!dyg_tensor_map_bf16(C_hi, (uint64_t)M, (uint64_t)N, (uint64_t)ldcs, 32u, &mch)) return 1;
        if (!dyg_tensor_map_bf16(C_mid, (uint64_t)M, (uint64_t)N, (uint64_t)ldcs, 32u, &mcm)) return 1;
    }
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ln_gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_ln_gemm_bf16x3: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    const int max_pairs = dyg_num_sms() / 2;
    const int pairs = (int)(g.m_super < max_pairs ? g.m_super : max_pairs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(LG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, ln_gemm_bf16x3_kernel, mwh, mwm, mch, mcm, p);
    if (le != cudaSuccess) {
        dyg_set_error("dyg_ln_gemm_bf16x3: launch failed: %s", cudaGetErrorString(le));
        return 1;
    }
    return 0;
}
