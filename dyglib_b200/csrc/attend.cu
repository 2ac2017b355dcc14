// Temporal attention over sampled neighbours and the small-sequence attention inside
// DyGFormer's transformer (row a16).
//
// dyg_temporal_attend is the folded form of MultiHeadAttention (models/modules.py:157-193): with
// qk_h = scaling * W_k,h^T W_q,h [x_i | cos(b)] precomputed per root, the kernel reads every neighbour row
// exactly once from HBM (node row, edge row) and computes the time encoding in registers, keeps a running
// (online) softmax per head and accumulates sum_j a_j x_ij.  The K/V projections of the reference then
// become two small dense GEMMs on (n, H*Dk) instead of two GEMMs on (n*k, Dk): HBM-bound by the gather.
#include <math.h>
#include "common.cuh"

template <int H, int R>
__global__ void __launch_bounds__(128, 3) temporal_attend_kernel(
    const float* __restrict__ qk, int ldq, int64_t n, int k,
    const float* __restrict__ node_tab, int ld_node, const float* __restrict__ node_tab2, int ld_node2,
    const int64_t* __restrict__ node_idx, int F4,
    const float* __restrict__ edge_tab, int ld_edge, const int64_t* __restrict__ edge_idx, int E4,
    const float* __restrict__ time_feat, const double* __restrict__ t_query, const float* __restrict__ t_nbr,
    const float* __restrict__ w, const float* __restrict__ b, int T4,
    const int64_t* __restrict__ mask_ids, float* __restrict__ out_s, int lds, float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int D4 = F4 + E4 + T4;
    const int Dk = D4 * 4;

    float4 q[H][R], acc[H][R];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = r * 32 + lane;
            q[h][r] = (c < D4) ? __ldg(reinterpret_cast<const float4*>(qk + i * ldq + h * Dk + 4 * c))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    // time-encoder parameters of the chunks this lane owns
    float4 tw[R], tb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int c = r * 32 + lane - F4 - E4;
        const bool mine = (c >= 0 && c < T4 && !time_feat);
        tw[r] = mine ? __ldg(reinterpret_cast<const float4*>(w + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        tb[r] = mine ? __ldg(reinterpret_cast<const float4*>(b + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const double tq = t_query ? __ldg(t_query + i) : 0.0;
    float mx[H], den[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { mx[h] = -INFINITY; den[h] = 0.f; }

    const int64_t base = i * (int64_t)k;
    for (int j0 = 0; j0 < k; j0 += 32) {
        // lane l stages the indices / time / mask of neighbour j0+l; broadcast by shuffle below
        const int jl = j0 + lane;
        int64_t my_n = 0, my_e = 0;
        float my_dt = 0.f;
        int my_masked = 0;
        if (jl < k) {
            my_n = node_idx ? __ldg(node_idx + base + jl) : base + jl;
            my_e = edge_idx ? __ldg(edge_idx + base + jl) : base + jl;
            if (!time_feat) my_dt = (float)(tq - (double)__ldg(t_nbr + base + jl));
            my_masked = mask_ids ? (__ldg(mask_ids + base + jl) == 0) : 0;
        }
        const int jn = (k - j0) < 32 ? (k - j0) : 32;

        auto load_x = [&](int jj, float4 (&x)[R]) {
            const int64_t rn = __shfl_sync(0xffffffffu, my_n, jj);
            const int64_t re = __shfl_sync(0xffffffffu, my_e, jj);
            const float dt = __shfl_sync(0xffffffffu, my_dt, jj);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int c = r * 32 + lane;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                // row 0 is the padding id: every padded neighbour of every root reads it, so it is kept in L1 (__ldg)
                // instead of hammering one L2 slice; real rows are read once and bypass L1
                if (c < F4) {
                    const float4* np = reinterpret_cast<const float4*>(node_tab + rn * ld_node) + c;
                    v = rn == 0 ? __ldg(np) : ldg_stream(np);
                    if (node_tab2) {
                        const float4* np2 = reinterpret_cast<const float4*>(node_tab2 + rn * ld_node2) + c;
                        const float4 u = rn == 0 ? __ldg(np2) : ldg_stream(np2);
                        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                    }
                } else if (c < F4 + E4) {
                    const float4* ep = reinterpret_cast<const float4*>(edge_tab + re * ld_edge) + (c - F4);
                    v = re == 0 ? __ldg(ep) : ldg_stream(ep);
                } else if (c < D4) {
                    if (time_feat) {
                        v = ldg_stream(reinterpret_cast<const float4*>(time_feat + (base + j0 + jj) * (int64_t)(T4 * 4)) + (c - F4 - E4));
                    } else {
                        v.x = dyg_time_enc(dt, tw[r].x, tb[r].x);
                        v.y = dyg_time_enc(dt, tw[r].y, tb[r].y);
                        v.z = dyg_time_enc(dt, tw[r].z, tb[r].z);
                        v.w = dyg_time_enc(dt, tw[r].w, tb[r].w);
                    }
                }
                x[r] = v;
            }
        };

        float4 xn[R];
        load_x(0, xn);
        for (int jj = 0; jj < jn; ++jj) {
            float4 x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = xn[r];
            if (jj + 1 < jn) load_x(jj + 1, xn);
            const int masked = __shfl_sync(0xffffffffu, my_masked, jj);
            float s[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float p = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    p = fmaf(q[h][r].x, x[r].x, p);
                    p = fmaf(q[h][r].y, x[r].y, p);
                    p = fmaf(q[h][r].z, x[r].z, p);
                    p = fmaf(q[h][r].w, x[r].w, p);
                }
                s[h] = p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int h = 0; h < H; ++h) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float sc = masked ? -1e10f : s[h];  // -1e10, not -inf (models/modules.py:184)
                if (out_scores && lane == 0) out_scores[(i * H + h) * (int64_t)k + j0 + jj] = sc;
                const float mnew = fmaxf(mx[h], sc);
                const float corr = expf(mx[h] - mnew);
                const float p = expf(sc - mnew);
                den[h] = den[h] * corr + p;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[h][r].x = fmaf(acc[h][r].x, corr, p * x[r].x);
                    acc[h][r].y = fmaf(acc[h][r].y, corr, p * x[r].y);
                    acc[h][r].z = fmaf(acc[h][r].z, corr, p * x[r].z);
                    acc[h][r].w = fmaf(acc[h][r].w, corr, p * x[r].w);
                }
                mx[h] = mnew;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.f / den[h];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = r * 32 + lane;
            if (c < D4) {
                float4 v = acc[h][r];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                *reinterpret_cast<float4*>(out_s + i * lds + h * Dk + 4 * c) = v;
            }
        }
    }
    if (out_scores) {
        __syncwarp();
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float inv = 1.f / den[h];
            for (int j = lane; j < k; j += 32) {
                float* p = out_scores + (i * H + h) * (int64_t)k + j;
                *p = expf(*p - mx[h]) * inv;
            }
        }
    }
}

// Leaner variant for F + E <= 384 and T <= 128 (TGAT / TGN: 344, 100).  Same math as temporal_attend_kernel with a
// different lane mapping: the node | edge part of a key row is owned as float4 chunks lane, lane+32, lane+64 (3 loads per
// neighbour), the T time features as scalars lane, lane+32, lane+64, lane+96, so the time encoding costs four cosines per
// lane and neighbour instead of eight half-empty ones; the running softmax rescales the accumulators only when the
// maximum actually moves.  ~2.4x fewer instructions per neighbour (the kernel is issue-bound, not HBM-bound: profiles/).
template <int H>
__global__ void __launch_bounds__(128, 4) temporal_attend_split_kernel(
    const float* __restrict__ qk, int ldq, int64_t n, int k,
    const float* __restrict__ node_tab, int ld_node, const float* __restrict__ node_tab2, int ld_node2,
    const int64_t* __restrict__ node_idx, int F4,
    const float* __restrict__ edge_tab, int ld_edge, const int64_t* __restrict__ edge_idx, int E4,
    const float* __restrict__ time_feat, const double* __restrict__ t_query, const float* __restrict__ t_nbr,
    const float* __restrict__ w, const float* __restrict__ b, int T,
    const int64_t* __restrict__ mask_ids, float* __restrict__ out_s, int lds, float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int NE4 = F4 + E4;                 // float4 chunks of the node | edge part
    const int Dk = NE4 * 4 + T;
    float4 q[H][3], acc[H][3];
    float qt[H][4], acct[H][4];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float* qrow = qk + i * ldq + h * Dk;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int c = r * 32 + lane;
            q[h][r] = (c < NE4) ? __ldg(reinterpret_cast<const float4*>(qrow) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int c = r * 32 + lane;
            qt[h][r] = (c < T) ? __ldg(qrow + NE4 * 4 + c) : 0.f;
            acct[h][r] = 0.f;
        }
    }
    float tw[4], tb[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = r * 32 + lane;
        const bool mine = c < T && !time_feat;
        tw[r] = mine ? __ldg(w + c) : 0.f;
        tb[r] = mine ? __ldg(b + c) : 0.f;
    }
    const double tq = t_query ? __ldg(t_query + i) : 0.0;
    float mx[H], den[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { mx[h] = -INFINITY; den[h] = 0.f; }

    const int64_t base = i * (int64_t)k;
    for (int j0 = 0; j0 < k; j0 += 32) {
        const int jl = j0 + lane;
        int64_t my_n = 0, my_e = 0;
        float my_dt = 0.f;
        int my_masked = 0;
        if (jl < k) {
            my_n = node_idx ? __ldg(node_idx + base + jl) : base + jl;
            my_e = edge_idx ? __ldg(edge_idx + base + jl) : base + jl;
            if (!time_feat) my_dt = (float)(tq - (double)__ldg(t_nbr + base + jl));
            my_masked = mask_ids ? (__ldg(mask_ids + base + jl) == 0) : 0;
        }
        const int jn = (k - j0) < 32 ? (k - j0) : 32;

        auto load_x = [&](int jj, float4 (&x)[3]) {
            const int64_t rn = __shfl_sync(0xffffffffu, my_n, jj);
            const int64_t re = __shfl_sync(0xffffffffu, my_e, jj);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int c = r * 32 + lane;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                // row 0 is the padding id: kept in L1 (__ldg) instead of hammering one L2 slice; real rows bypass L1
                if (c < F4) {
                    const float4* np = reinterpret_cast<const float4*>(node_tab + rn * ld_node) + c;
                    v = rn == 0 ? __ldg(np) : ldg_stream(np);
                    if (node_tab2) {
                        const float4* np2 = reinterpret_cast<const float4*>(node_tab2 + rn * ld_node2) + c;
                        const float4 u = rn == 0 ? __ldg(np2) : ldg_stream(np2);
                        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                    }
                } else if (c < NE4) {
                    const float4* ep = reinterpret_cast<const float4*>(edge_tab + re * ld_edge) + (c - F4);
                    v = re == 0 ? __ldg(ep) : ldg_stream(ep);
                }
                x[r] = v;
            }
        };

        float4 xn[3];
        load_x(0, xn);
        for (int jj = 0; jj < jn; ++jj) {
            float4 x[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) x[r] = xn[r];
            if (jj + 1 < jn) load_x(jj + 1, xn);
            // time features of this neighbour: 4 per lane
            float xt[4];
            if (time_feat) {
                const float* tf = time_feat + (base + j0 + jj) * (int64_t)T;
#pragma unroll
                for (int r = 0; r < 4; ++r) xt[r] = (r * 32 + lane < T) ? __ldg(tf + r * 32 + lane) : 0.f;
            } else {
                const float dt = __shfl_sync(0xffffffffu, my_dt, jj);
#pragma unroll
                for (int r = 0; r < 4; ++r) xt[r] = (r * 32 + lane < T) ? dyg_time_enc(dt, tw[r], tb[r]) : 0.f;
            }
            const int masked = __shfl_sync(0xffffffffu, my_masked, jj);
            float s[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float p = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    p = fmaf(q[h][r].x, x[r].x, p);
                    p = fmaf(q[h][r].y, x[r].y, p);
                    p = fmaf(q[h][r].z, x[r].z, p);
                    p = fmaf(q[h][r].w, x[r].w, p);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) p = fmaf(qt[h][r], xt[r], p);
                s[h] = p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int h = 0; h < H; ++h) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float sc = masked ? -1e10f : s[h];  // -1e10, not -inf (models/modules.py:184)
                if (out_scores && lane == 0) out_scores[(i * H + h) * (int64_t)k + j0 + jj] = sc;
                if (sc > mx[h]) {                         // warp-uniform: the maximum moves, rescale what was accumulated
                    const float corr = expf(mx[h] - sc);
                    den[h] *= corr;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        acc[h][r].x *= corr; acc[h][r].y *= corr; acc[h][r].z *= corr; acc[h][r].w *= corr;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) acct[h][r] *= corr;
                    mx[h] = sc;
                }
                const float p = expf(sc - mx[h]);
                den[h] += p;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    acc[h][r].x = fmaf(p, x[r].x, acc[h][r].x);
                    acc[h][r].y = fmaf(p, x[r].y, acc[h][r].y);
                    acc[h][r].z = fmaf(p, x[r].z, acc[h][r].z);
                    acc[h][r].w = fmaf(p, x[r].w, acc[h][r].w);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) acct[h][r] = fmaf(p, xt[r], acct[h][r]);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.f / den[h];
        float* orow = out_s + i * lds + h * Dk;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int c = r * 32 + lane;
            if (c < NE4) {
                float4 v = acc[h][r];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                *(reinterpret_cast<float4*>(orow) + c) = v;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int c = r * 32 + lane;
            if (c < T) orow[NE4 * 4 + c] = acct[h][r] * inv;
        }
    }
    if (out_scores) {
        __syncwarp();
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float inv = 1.f / den[h];
            for (int j = lane; j < k; j += 32) {
                float* p = out_scores + (i * H + h) * (int64_t)k + j;
                *p = expf(*p - mx[h]) * inv;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Ring variant of temporal_attend_split_kernel (same math, same lane mapping) for time encodings computed in the kernel:
// neighbour rows are moved by the bulk-copy engine (cp.async.bulk, completion on an mbarrier) into a per-warp ring of
// RING_DEPTH stages in shared memory instead of by register loads with one row of lookahead.  A warp owns a contiguous
// range of roots and streams their (root, neighbour) items through the ring, so RING_DEPTH rows per warp (16 warps per
// SM: ~180 KB) are in flight across root boundaries; the split kernel held <= 2 rows per warp and was latency-bound
// (DRAM 41 %, issue 35 %: profiles/).  Indices, time deltas and flags of 32 items at a time are loaded coalesced and
// parked in shared memory; the lane that loaded an item's indices issues its copies, nothing is shuffled.
constexpr int RING_DEPTH = 6;
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void ring_bar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
}
__device__ __forceinline__ void ring_bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ring_bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "RW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RW_DONE;\n\t"
        "bra RW_LOOP;\n\t"
        "RW_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

constexpr int RING_GROUP = 4;   // neighbours whose scores are reduced / soft-maxed together
__host__ __device__ constexpr int ring_stage_bytes(int F4, int E4, bool node2) { return (((F4 + E4 + (node2 ? F4 : 0)) * 16) + 127) & ~127; }
__host__ __device__ constexpr int ring_warp_bytes(int F4, int E4, bool node2) {
    return (RING_DEPTH * ring_stage_bytes(F4, E4, node2) + 64 * 24 + RING_DEPTH * 8 + RING_GROUP * 128 * 4 + 127) & ~127;
}

// All dimensions are template parameters (F4 / E4: float4 chunks of a node / edge row, T: time features): every bound in
// the lane loops folds at compile time.  The first version with run-time dimensions spent half of its 350 warp
// instructions per neighbour on predicates, selects and index arithmetic (ncu source page, profiles/).
// FULLG: k is a multiple of RING_GROUP, so every group is full and the per-neighbour "gi < g" tests fold away.
template <int H, bool NODE2, int F4, int E4, int T, bool FULLG>
__global__ void __launch_bounds__(128, NODE2 ? 3 : 4) temporal_attend_ring_kernel(
    const float* __restrict__ qk, int ldq, int64_t n, int k,
    const float* __restrict__ node_tab, int ld_node, const float* __restrict__ node_tab2, int ld_node2,
    const int64_t* __restrict__ node_idx, const float* __restrict__ edge_tab, int ld_edge, const int64_t* __restrict__ edge_idx,
    const double* __restrict__ t_query, const float* __restrict__ t_nbr,
    const float* __restrict__ w, const float* __restrict__ b,
    const int64_t* __restrict__ mask_ids, float* __restrict__ out_s, int lds, float* __restrict__ out_scores, int zero_row0,
    const float* __restrict__ prob_scale) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    constexpr int G = RING_GROUP, V = G * H;          // V partial scores per group: (neighbour g, head h) -> index g * H + h
    constexpr int SH = (V == 8) ? 2 : 3;              // after the transposing reduction lane l owns score index l >> SH
    constexpr int NE4 = F4 + E4, Dk = NE4 * 4 + T;
    constexpr int XR = (NE4 + 31) / 32;               // float4 rounds of the node | edge part
    constexpr int TF = T / 32, TR = T % 32;           // full time-feature rounds, remainder
    constexpr bool GROUPED_REM = TR > 0 && TR * G <= 32;   // remainder of the whole group computed in one round
    constexpr int ROUNDS = (TR > 0 && !GROUPED_REM) ? TF + 1 : TF;
    constexpr int TQ = (T + 31) / 32;
    constexpr int STAGE = ring_stage_bytes(F4, E4, NODE2), WARP_BYTES = ring_warp_bytes(F4, E4, NODE2);
    constexpr uint32_t NODE_BYTES = F4 * 16u, EDGE_BYTES = E4 * 16u;
    static_assert(XR <= 3 && TQ <= 4, "key row too wide for the lane mapping");
    const int lane = threadIdx.x & 31;
    const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned char* wbase = ring_smem + (size_t)(threadIdx.x >> 5) * WARP_BYTES;
    int64_t* s_rn = reinterpret_cast<int64_t*>(wbase + RING_DEPTH * STAGE);   // 2 x 32 items (two index chunks)
    int64_t* s_re = s_rn + 64;
    float* s_dt = reinterpret_cast<float*>(s_re + 64);
    int* s_fl = reinterpret_cast<int*>(s_dt + 64);    // bit 0: masked, bit 1: node row is the zero padding row, bit 2: edge row is
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_fl + 64);
    float* s_xt = reinterpret_cast<float*>(bars + RING_DEPTH);   // time encodings of the group's neighbours: G x 128
    if (lane < RING_DEPTH) ring_bar_init(bars + lane);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    const int64_t r0 = (n * gw) / nw, r1 = (n * (gw + 1)) / nw;
    const int items = (int)((r1 - r0) * k);
    if (items <= 0) return;
    const int64_t gi0 = r0 * k;

    // indices / time delta / flags of item 32 c + lane, loaded coalesced
    int64_t c_rn = 0, c_re = 0;
    float c_dt = 0.f;
    int c_fl = 0;
    auto load_chunk = [&](int c) {
        const int t = 32 * c + lane;
        c_rn = 0; c_re = 0; c_dt = 0.f; c_fl = 0;
        if (t < items) {
            const int64_t gi = gi0 + t;
            c_rn = node_idx ? __ldg(node_idx + gi) : gi;
            c_re = edge_idx ? __ldg(edge_idx + gi) : gi;
            c_dt = (float)(__ldg(t_query + gi / k) - (double)__ldg(t_nbr + gi));
            c_fl = (mask_ids ? (__ldg(mask_ids + gi) == 0) : 0) | ((node_idx && c_rn == 0 && (zero_row0 & 1)) ? 2 : 0) |
                   ((edge_idx && c_re == 0 && (zero_row0 & 2)) ? 4 : 0);
        }
    };
    auto park_chunk = [&](int c) {
        const int slot = (c & 1) * 32 + lane;
        s_rn[slot] = c_rn; s_re[slot] = c_re; s_dt[slot] = c_dt; s_fl[slot] = c_fl;
    };
    // Items t0 .. t0 + cnt - 1 (cnt <= RING_GROUP, consecutive ring stages starting at stage0) are issued in ONE pass: the lane
    // that parked item u's indices (lane u & 31) issues its copies.  Stage and phase of an item are carried as counters (the
    // ring depth is not a power of two: t % 6 and t / 6 cost a multiply-high sequence each, several times per neighbour).
    auto issue_group = [&](int t0, int cnt, int stage0) {
        const int off = (lane - t0) & 31;
        const int u = t0 + off;
        if (off < cnt && u < items) {
            int stg = stage0 + off;
            if (stg >= RING_DEPTH) stg -= RING_DEPTH;
            const int slot = u & 63;
            const int fl = s_fl[slot];
            const int64_t rn = s_rn[slot], re = s_re[slot];
            unsigned char* dst = wbase + stg * STAGE;
            uint64_t* bar = bars + stg;
            const uint32_t nb = (fl & 2) ? 0u : NODE_BYTES, eb = (fl & 4) ? 0u : EDGE_BYTES;
            ring_bar_expect(bar, nb * (NODE2 ? 2u : 1u) + eb);
            if (nb) {
                bulk_g2s(dst, node_tab + rn * ld_node, nb, bar);
                if (NODE2) bulk_g2s(dst + NODE_BYTES + EDGE_BYTES, node_tab2 + rn * ld_node2, nb, bar);
            }
            if (eb) bulk_g2s(dst + NODE_BYTES, edge_tab + re * ld_edge, eb, bar);
        }
    };
    // row of item t from its ring stage; fl (warp-uniform) says which parts were not copied because they are zero rows
    auto load_x = [&](int stg, int fl, float4 (&x)[XR]) {
        const float4* row = reinterpret_cast<const float4*>(wbase + stg * STAGE);
#pragma unroll
        for (int r = 0; r < XR; ++r) {
            const int c = r * 32 + lane;
            if ((r + 1) * 32 <= NE4 || c < NE4) {
                x[r] = row[c];
                if (NODE2 && (r * 32 < F4) && ((r + 1) * 32 <= F4 || c < F4)) {
                    const float4 u = row[NE4 + c];
                    x[r].x += u.x; x[r].y += u.y; x[r].z += u.z; x[r].w += u.w;
                }
            } else {
                x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (fl & 6) {   // rare, warp-uniform: the skipped parts hold stale bytes
#pragma unroll
            for (int r = 0; r < XR; ++r) {
                const int c = r * 32 + lane;
                if (fl & (c < F4 ? 2 : 4)) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };

    load_chunk(0);
    park_chunk(0);
    __syncwarp();
    issue_group(0, RING_DEPTH, 0);   // RING_DEPTH <= 32: one pass
    load_chunk(1);
    int parked = 0;                           // highest index chunk parked in shared memory; chunk parked + 1 sits in registers

    float tw[TQ], tb[TQ];
#pragma unroll
    for (int r = 0; r < TQ; ++r) {
        const int c = r * 32 + lane;
        tw[r] = c < T ? __ldg(w + c) : 0.f;
        tb[r] = c < T ? __ldg(b + c) : 0.f;
    }
    // remainder features of the time encoder (T % 32 <= 8): one round for the whole group, lane = neighbour * TR + feature
    const int rem_g = GROUPED_REM ? lane / (TR > 0 ? TR : 1) : G, rem_f = GROUPED_REM ? TF * 32 + lane % (TR > 0 ? TR : 1) : 0;
    const float rem_w = (GROUPED_REM && rem_g < G) ? __ldg(w + rem_f) : 0.f, rem_b = (GROUPED_REM && rem_g < G) ? __ldg(b + rem_f) : 0.f;

    float4 q[H][XR], acc[H][XR];
    float qt[H][TQ], acct[H][TQ];
    const int own = lane >> SH, own_g = own / H, own_h = own % H;
    float m_own = -INFINITY, den_own = 0.f;   // running softmax state of head own_h (identical in all lanes of that head)
    int j = 0;
    int64_t root = r0;
    int stage_t = 0;            // t % RING_DEPTH
    uint32_t phase_t = 0;       // (t / RING_DEPTH) & 1
#pragma unroll 1
    for (int t = 0; t < items;) {
        if (j == 0) {   // new root: its folded query, fresh accumulators; warm L2 with the next root's query row
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float* qrow = qk + root * ldq + h * Dk;
#pragma unroll
                for (int r = 0; r < XR; ++r) {
                    const int c = r * 32 + lane;
                    q[h][r] = (c < NE4) ? __ldg(reinterpret_cast<const float4*>(qrow) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int r = 0; r < TQ; ++r) {
                    const int c = r * 32 + lane;
                    qt[h][r] = (c < T) ? __ldg(qrow + NE4 * 4 + c) : 0.f;
                    acct[h][r] = 0.f;
                }
            }
            m_own = -INFINITY;
            den_own = 0.f;
            if (root + 1 < r1 && lane * 32 < H * Dk)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(qk + (root + 1) * ldq + lane * 32));
        }
        // chunk bookkeeping: chunk c+1 waits in registers while chunk c is consumed; it is parked (over chunk c-1, fully
        // consumed) once item 32 c + 8 is reached, before the first lookahead into it at item 32 (c+1) - RING_DEPTH
        if (parked == (t >> 5) && (t & 31) >= 8) {
            park_chunk(parked + 1);
            ++parked;
            __syncwarp();
            load_chunk(parked + 1);
        }
        const int g = FULLG ? G : ((k - j) < G ? (k - j) : G);

        // ---- pass 1: time encodings (parked in s_xt) and the partial scores of the group's neighbours
        float part[V];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
#pragma unroll
            for (int h = 0; h < H; ++h) part[gi * H + h] = 0.f;
            if (FULLG || gi < g) {
                const int tt = t + gi;
                int stg = stage_t + gi;
                uint32_t ph = phase_t;
                if (stg >= RING_DEPTH) { stg -= RING_DEPTH; ph ^= 1u; }
                const float dt = s_dt[tt & 63];
                float xt[TQ];
#pragma unroll
                for (int r = 0; r < TQ; ++r) {
                    xt[r] = 0.f;
                    if (r < ROUNDS) {
                        const int c = r * 32 + lane;
                        if ((r + 1) * 32 <= T || c < T) {
                            xt[r] = dyg_time_enc(dt, tw[r], tb[r]);
                            s_xt[gi * 128 + c] = xt[r];
                        }
                    }
                }
                ring_bar_wait(bars + stg, ph);
                float4 x[XR];
                load_x(stg, s_fl[tt & 63], x);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float p0 = 0.f, p1 = 0.f;   // two chains
#pragma unroll
                    for (int r = 0; r < XR; ++r) {
                        p0 = fmaf(q[h][r].x, x[r].x, p0);
                        p1 = fmaf(q[h][r].y, x[r].y, p1);
                        p0 = fmaf(q[h][r].z, x[r].z, p0);
                        p1 = fmaf(q[h][r].w, x[r].w, p1);
                    }
#pragma unroll
                    for (int r = 0; r < ROUNDS; ++r) p0 = fmaf(qt[h][r], xt[r], p0);
                    part[gi * H + h] = p0 + p1;
                }
            }
        }
        if (GROUPED_REM) {
            if (rem_g < g) s_xt[rem_g * 128 + rem_f] = dyg_time_enc(s_dt[(t + rem_g) & 63], rem_w, rem_b);
            __syncwarp();
            if (lane < TR) {
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    if (FULLG || gi < g) {
                        const float xr = s_xt[gi * 128 + TF * 32 + lane];
#pragma unroll
                        for (int h = 0; h < H; ++h) part[gi * H + h] = fmaf(qt[h][TQ - 1], xr, part[gi * H + h]);
                    }
                }
            }
        } else {
            __syncwarp();
        }
        // ---- transposing reduction: V values per lane -> lane l holds the warp total of value l >> SH.  Written out per
        // step with compile-time bounds: a loop over a shrinking count left `part` dynamically indexed (in local memory).
#define DYG_XSTEP(NV, WIDTH)                                                        \
    {                                                                               \
        const bool up = lane & (WIDTH);                                             \
        _Pragma("unroll") for (int i = 0; i < (NV) / 2; ++i) {                      \
            const float send = up ? part[i] : part[i + (NV) / 2];                   \
            const float keep = up ? part[i + (NV) / 2] : part[i];                   \
            part[i] = keep + __shfl_xor_sync(0xffffffffu, send, (WIDTH));           \
        }                                                                           \
    }
        if (V == 8) {
            DYG_XSTEP(8, 16) DYG_XSTEP(4, 8) DYG_XSTEP(2, 4)
        } else {
            DYG_XSTEP(4, 16) DYG_XSTEP(2, 8)
        }
#undef DYG_XSTEP
#pragma unroll
        for (int width = (1 << SH) >> 1; width > 0; width >>= 1) part[0] += __shfl_xor_sync(0xffffffffu, part[0], width);
        // ---- softmax update of the group (lane-parallel over (neighbour, head))
        float sc = -INFINITY;
        if (own_g < g) sc = (s_fl[(t + own_g) & 63] & 1) ? -1e10f : part[0];   // -1e10, not -inf (models/modules.py:184)
        if (out_scores && own_g < g && (lane & ((1 << SH) - 1)) == 0) out_scores[(root * H + own_h) * (int64_t)k + j + own_g] = sc;
        float gm = fmaxf(sc, __shfl_xor_sync(0xffffffffu, sc, 16));
        gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 8));
        const float new_m = fmaxf(m_own, gm);
        const float corr = __expf(m_own - new_m);
        const float p_own = __expf(sc - new_m);
        float ps = p_own + __shfl_xor_sync(0xffffffffu, p_own, 16);
        ps += __shfl_xor_sync(0xffffffffu, ps, 8);
        den_own = fmaf(den_own, corr, ps);
        m_own = new_m;
        // training: dropout multipliers of the attention probabilities (models/modules.py:187) weigh the sum, not the normaliser
        const float pw_own = (prob_scale && own_g < g) ? p_own * __ldg(prob_scale + (root * H + own_h) * (int64_t)k + j + own_g) : p_own;
        float corr_h[H];
        bool rescale = false;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            corr_h[h] = __shfl_sync(0xffffffffu, corr, h << SH);
            rescale |= corr_h[h] != 1.f;
        }
        if (rescale) {   // warp-uniform: a maximum moved
#pragma unroll
            for (int h = 0; h < H; ++h) {
#pragma unroll
                for (int r = 0; r < XR; ++r) {
                    acc[h][r].x *= corr_h[h]; acc[h][r].y *= corr_h[h]; acc[h][r].z *= corr_h[h]; acc[h][r].w *= corr_h[h];
                }
#pragma unroll
                for (int r = 0; r < TQ; ++r) acct[h][r] *= corr_h[h];
            }
        }
        // ---- pass 2: weighted sums (rows re-read from the ring, time encodings from s_xt)
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            if (FULLG || gi < g) {
                float pg[H];
#pragma unroll
                for (int h = 0; h < H; ++h) pg[h] = __shfl_sync(0xffffffffu, pw_own, (gi * H + h) << SH);
                int stg = stage_t + gi;
                if (stg >= RING_DEPTH) stg -= RING_DEPTH;
                float4 x[XR];
                load_x(stg, s_fl[(t + gi) & 63], x);
                float xt[TQ];
#pragma unroll
                for (int r = 0; r < TQ; ++r) xt[r] = ((r + 1) * 32 <= T || r * 32 + lane < T) ? s_xt[gi * 128 + r * 32 + lane] : 0.f;
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int r = 0; r < XR; ++r) {
                        acc[h][r].x = fmaf(pg[h], x[r].x, acc[h][r].x);
                        acc[h][r].y = fmaf(pg[h], x[r].y, acc[h][r].y);
                        acc[h][r].z = fmaf(pg[h], x[r].z, acc[h][r].z);
                        acc[h][r].w = fmaf(pg[h], x[r].w, acc[h][r].w);
                    }
#pragma unroll
                    for (int r = 0; r < TQ; ++r) acct[h][r] = fmaf(pg[h], xt[r], acct[h][r]);
                }
            }
        }
        __syncwarp();                         // every lane is done with the group's stages and s_xt: refill
        issue_group(t + RING_DEPTH, g, stage_t);   // item u + RING_DEPTH reuses the stage of item u
        t += g;
        j += g;
        stage_t += g;
        if (stage_t >= RING_DEPTH) { stage_t -= RING_DEPTH; phase_t ^= 1u; }
        if (j == k) {   // root finished: normalise and write
            float inv_h[H], mx_h[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                inv_h[h] = 1.f / __shfl_sync(0xffffffffu, den_own, h << SH);
                mx_h[h] = __shfl_sync(0xffffffffu, m_own, h << SH);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float inv = inv_h[h];
                float* orow = out_s + root * lds + h * Dk;
#pragma unroll
                for (int r = 0; r < XR; ++r) {
                    const int c = r * 32 + lane;
                    if (c < NE4) {
                        float4 v = acc[h][r];
                        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                        *(reinterpret_cast<float4*>(orow) + c) = v;
                    }
                }
#pragma unroll
                for (int r = 0; r < TQ; ++r) {
                    const int c = r * 32 + lane;
                    if (c < T) orow[NE4 * 4 + c] = acct[h][r] * inv;
                }
            }
            if (out_scores) {
                __syncwarp();
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    for (int jj = lane; jj < k; jj += 32) {
                        float* p = out_scores + (root * H + h) * (int64_t)k + jj;
                        *p = expf(*p - mx_h[h]) * inv_h[h];
                    }
                }
            }
            j = 0;
            ++root;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of dyg_temporal_attend (training path; the reference differentiates models/modules.py:157-193 with autograd).
// With probs a_hj = softmax_j(q_h . x_j) (before dropout), m_hj the dropout multipliers, s_h = sum_j a_hj m_hj x_j and
// g_h = dL/ds_h:   u_hj = g_h . x_j,  c_h = g_h . s_h,  dscore_hj = a_hj (m_hj u_hj - c_h)   (0 for masked neighbours:
// masked_fill cuts the gradient of the score, models/modules.py:184),
//   dL/dq_h = sum_j dscore_hj x_j,     dL/dx_j = sum_h (a_hj m_hj g_h + dscore_hj q_h).
// dL/dx_j's node part is written when the neighbour rows are a dense trainable tensor (deeper layers), its time part is
// chained through cos: dL/dw += -sin(arg) dt gx, dL/db += -sin(arg) gx, summed in registers and added once per warp.
// One warp per root, the lane mapping of temporal_attend_split_kernel; neighbours are read once.
template <int H>
__global__ void __launch_bounds__(128) temporal_attend_bwd_kernel(
    const float* __restrict__ qk, int ldq, int64_t n, int k,
    const float* __restrict__ node_tab, int ld_node, const float* __restrict__ node_tab2, int ld_node2,
    const int64_t* __restrict__ node_idx, int F4,
    const float* __restrict__ edge_tab, int ld_edge, const int64_t* __restrict__ edge_idx, int E4,
    const double* __restrict__ t_query, const float* __restrict__ t_nbr,
    const float* __restrict__ w, const float* __restrict__ b, int T, const int64_t* __restrict__ mask_ids,
    const float* __restrict__ probs, const float* __restrict__ prob_scale, const float* __restrict__ s_out, int lds,
    const float* __restrict__ gs, int ldg, float* __restrict__ gqk, int ldgq, float* __restrict__ g_nbr, int ld_gn,
    float* __restrict__ gw, float* __restrict__ gb) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int NE4 = F4 + E4;
    const int Dk = NE4 * 4 + T;
    float tw[4], tb[4], gw_acc[4], gb_acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = r * 32 + lane;
        tw[r] = c < T ? __ldg(w + c) : 0.f;
        tb[r] = c < T ? __ldg(b + c) : 0.f;
        gw_acc[r] = 0.f;
        gb_acc[r] = 0.f;
    }
    for (int64_t i = warp0; i < n; i += nwarps) {
        float4 q[H][3], g[H][3], gq[H][3];
        float qt[H][4], gt[H][4], gqt[H][4];
        float ch[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float* qrow = qk + i * ldq + h * Dk;
            const float* grow = gs + i * ldg + h * Dk;
            const float* srow = s_out + i * lds + h * Dk;
            float c = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int cc = r * 32 + lane;
                const bool on = cc < NE4;
                q[h][r] = on ? __ldg(reinterpret_cast<const float4*>(qrow) + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                g[h][r] = on ? __ldg(reinterpret_cast<const float4*>(grow) + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 sv = on ? __ldg(reinterpret_cast<const float4*>(srow) + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                c += g[h][r].x * sv.x + g[h][r].y * sv.y + g[h][r].z * sv.z + g[h][r].w * sv.w;
                gq[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int cc = r * 32 + lane;
                const bool on = cc < T;
                qt[h][r] = on ? __ldg(qrow + NE4 * 4 + cc) : 0.f;
                gt[h][r] = on ? __ldg(grow + NE4 * 4 + cc) : 0.f;
                c += gt[h][r] * (on ? __ldg(srow + NE4 * 4 + cc) : 0.f);
                gqt[h][r] = 0.f;
            }
            ch[h] = warp_sum(c);
        }
        const double tq = __ldg(t_query + i);
        const int64_t base = i * (int64_t)k;
        for (int j = 0; j < k; ++j) {
            const int64_t rn = node_idx ? __ldg(node_idx + base + j) : base + j;
            const int64_t re = edge_idx ? __ldg(edge_idx + base + j) : base + j;
            const float dt = (float)(tq - (double)__ldg(t_nbr + base + j));
            const bool masked = mask_ids ? (__ldg(mask_ids + base + j) == 0) : false;
            float4 x[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int c = r * 32 + lane;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < F4) {
                    v = __ldg(reinterpret_cast<const float4*>(node_tab + rn * ld_node) + c);
                    if (node_tab2) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(node_tab2 + rn * ld_node2) + c);
                        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                    }
                } else if (c < NE4) {
                    v = __ldg(reinterpret_cast<const float4*>(edge_tab + re * ld_edge) + (c - F4));
                }
                x[r] = v;
            }
            float xt[4], sn[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                xt[r] = 0.f;
                sn[r] = 0.f;
                if (r * 32 + lane < T) dyg_sincosf(fmaf(dt, tw[r], tb[r]), &sn[r], &xt[r]);
            }
            float u[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float p = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r) p += g[h][r].x * x[r].x + g[h][r].y * x[r].y + g[h][r].z * x[r].z + g[h][r].w * x[r].w;
#pragma unroll
                for (int r = 0; r < 4; ++r) p = fmaf(gt[h][r], xt[r], p);
                u[h] = warp_sum(p);
            }
            float am[H], ds[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float a = __ldg(probs + (i * H + h) * (int64_t)k + j);
                const float m = prob_scale ? __ldg(prob_scale + (i * H + h) * (int64_t)k + j) : 1.f;
                am[h] = a * m;
                ds[h] = masked ? 0.f : a * (m * u[h] - ch[h]);
            }
            float4 gx[3];
            float gxt[4];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                gx[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    gq[h][r].x = fmaf(ds[h], x[r].x, gq[h][r].x);
                    gq[h][r].y = fmaf(ds[h], x[r].y, gq[h][r].y);
                    gq[h][r].z = fmaf(ds[h], x[r].z, gq[h][r].z);
                    gq[h][r].w = fmaf(ds[h], x[r].w, gq[h][r].w);
                    gx[r].x += am[h] * g[h][r].x + ds[h] * q[h][r].x;
                    gx[r].y += am[h] * g[h][r].y + ds[h] * q[h][r].y;
                    gx[r].z += am[h] * g[h][r].z + ds[h] * q[h][r].z;
                    gx[r].w += am[h] * g[h][r].w + ds[h] * q[h][r].w;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                gxt[r] = 0.f;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    gqt[h][r] = fmaf(ds[h], xt[r], gqt[h][r]);
                    gxt[r] += am[h] * gt[h][r] + ds[h] * qt[h][r];
                }
                gb_acc[r] = fmaf(-sn[r], gxt[r], gb_acc[r]);
                gw_acc[r] = fmaf(-sn[r] * dt, gxt[r], gw_acc[r]);
            }
            if (g_nbr) {
                float* grow = g_nbr + (base + j) * (int64_t)ld_gn;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int c = r * 32 + lane;
                    if (c < F4) *(reinterpret_cast<float4*>(grow) + c) = gx[r];
                }
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float* orow = gqk + i * ldgq + h * Dk;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int c = r * 32 + lane;
                if (c < NE4) *(reinterpret_cast<float4*>(orow) + c) = gq[h][r];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = r * 32 + lane;
                if (c < T) orow[NE4 * 4 + c] = gqt[h][r];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = r * 32 + lane;
        if (c < T) {
            if (gw) atomicAdd(gw + c, gw_acc[r]);
            if (gb) atomicAdd(gb + c, gb_acc[r]);
        }
    }
}

extern "C" int dyg_temporal_attend_bwd(const float* qk, int ldq, int64_t n, int k, int H, const float* node_tab, int ld_node,
                                       const float* node_tab2, int ld_node2, const int64_t* node_idx, int F,
                                       const float* edge_tab, int ld_edge, const int64_t* edge_idx, int E,
                                       const double* t_query, const float* t_nbr, const float* w, const float* b, int T,
                                       const int64_t* mask_ids, const float* probs, const float* prob_scale, const float* s_out,
                                       int lds, const float* grad_s, int ldg, float* grad_qk, int ldgq, float* grad_nbr,
                                       int ld_gn, float* grad_w, float* grad_b, dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && k > 0, "dyg_temporal_attend_bwd: bad sizes");
    DYG_CHECK_ARG(H == 1 || H == 2, "dyg_temporal_attend_bwd: num_heads=%d unsupported (1 or 2)", H);
    DYG_CHECK_ARG((F % 4) == 0 && (E % 4) == 0 && F > 0 && E >= 0 && T >= 0 && F + E <= 384 && T <= 128,
                  "dyg_temporal_attend_bwd: unsupported feature dims %d / %d / %d", F, E, T);
    DYG_CHECK_ARG((ld_node % 4) == 0 && (ld_edge % 4) == 0 && (ldq % 4) == 0 && (lds % 4) == 0 && (ldg % 4) == 0 &&
                      (ldgq % 4) == 0 && (!node_tab2 || (ld_node2 % 4) == 0) && (!grad_nbr || (ld_gn % 4) == 0),
                  "dyg_temporal_attend_bwd: leading dims must be multiples of 4");
    DYG_CHECK_ARG(aligned16(qk) && aligned16(node_tab) && aligned16(edge_tab) && aligned16(s_out) && aligned16(grad_s) &&
                      aligned16(grad_qk) && (!node_tab2 || aligned16(node_tab2)) && (!grad_nbr || aligned16(grad_nbr)),
                  "dyg_temporal_attend_bwd: pointers must be 16-byte aligned");
    DYG_CHECK_ARG(t_query && t_nbr && w && b && probs, "dyg_temporal_attend_bwd: NULL pointer");
    if (n == 0) return 0;
    int64_t warps = (int64_t)dyg_num_sms() * 16;
    if (warps > n) warps = n;
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    cudaStream_t s = as_stream(stream);
#define ATTEND_BWD_ARGS qk, ldq, n, k, node_tab, ld_node, node_tab2, ld_node2, node_idx, F / 4, edge_tab, ld_edge, edge_idx, E / 4, \
                        t_query, t_nbr, w, b, T, mask_ids, probs, prob_scale, s_out, lds, grad_s, ldg, grad_qk, ldgq, grad_nbr,   \
                        ld_gn, grad_w, grad_b
    if (H == 2) temporal_attend_bwd_kernel<2><<<blocks, 128, 0, s>>>(ATTEND_BWD_ARGS);
    else temporal_attend_bwd_kernel<1><<<blocks, 128, 0, s>>>(ATTEND_BWD_ARGS);
#undef ATTEND_BWD_ARGS
    DYG_LAUNCH_CHECK("dyg_temporal_attend_bwd");
    return 0;
}

extern "C" int dyg_temporal_attend(const float* qk, int ldq, int64_t n, int k, int H, const float* node_tab, int ld_node,
                                   const float* node_tab2, int ld_node2, const int64_t* node_idx, int F,
                                   const float* edge_tab, int ld_edge, const int64_t* edge_idx, int E,
                                   const float* time_feat, const double* t_query, const float* t_nbr, const float* w,
                                   const float* b, int T, const int64_t* mask_ids, float* out_s, int lds,
                                   float* out_scores, int zero_row0, const float* prob_scale, dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && k > 0, "dyg_temporal_attend: bad sizes");
    DYG_CHECK_ARG(!prob_scale || (F == 172 && E == 172 && T == 100 && !time_feat),
                  "dyg_temporal_attend: prob_scale (training) needs the 172 / 172 / 100 feature widths and in-kernel time encoding");
    DYG_CHECK_ARG(H == 1 || H == 2, "dyg_temporal_attend: num_heads=%d unsupported (1 or 2)", H);
    DYG_CHECK_ARG((F % 4) == 0 && (E % 4) == 0 && (T % 4) == 0 && F > 0 && E >= 0 && T >= 0,
                  "dyg_temporal_attend: feature dims must be multiples of 4");
    DYG_CHECK_ARG((F + E + T) <= 512, "dyg_temporal_attend: key dim %d > 512 unsupported", F + E + T);
    DYG_CHECK_ARG((ld_node % 4) == 0 && (ld_edge % 4) == 0 && (ldq % 4) == 0 && (lds % 4) == 0 &&
                      (!node_tab2 || (ld_node2 % 4) == 0),
                  "dyg_temporal_attend: leading dims must be multiples of 4");
    DYG_CHECK_ARG(aligned16(qk) && aligned16(node_tab) && aligned16(edge_tab) && aligned16(out_s) &&
                      (!node_tab2 || aligned16(node_tab2)) && (!time_feat || aligned16(time_feat)) &&
                      (time_feat || (aligned16(w) && aligned16(b))),
                  "dyg_temporal_attend: pointers must be 16-byte aligned");
    DYG_CHECK_ARG(time_feat || (t_query && t_nbr && w && b), "dyg_temporal_attend: need time_feat or (t_query,t_nbr,w,b)");
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n * 32 + 127) / 128);
    cudaStream_t s = as_stream(stream);
#define ATTEND_ARGS qk, ldq, n, k, node_tab, ld_node, node_tab2, ld_node2, node_idx, F / 4, edge_tab, ld_edge, edge_idx, \
                    E / 4, time_feat, t_query, t_nbr, w, b, T / 4, mask_ids, out_s, lds, out_scores
    if (F == 172 && E == 172 && T == 100 && !time_feat) {
        // the reference's feature widths (172-d node / edge rows, 100 time features): ring kernel, rows by the bulk-copy
        // engine into a per-warp shared-memory ring
        const bool node2 = node_tab2 != nullptr;
        const int smem = 4 * ring_warp_bytes(43, 43, node2);
        int64_t warps = (int64_t)dyg_num_sms() * (node2 ? 12 : 16);
        if (warps > n) warps = n;
        const unsigned rblocks = (unsigned)((warps + 3) / 4);
#define ATTEND_RING_ARGS qk, ldq, n, k, node_tab, ld_node, node_tab2, ld_node2, node_idx, edge_tab, ld_edge, edge_idx, \
                         t_query, t_nbr, w, b, mask_ids, out_s, lds, out_scores, zero_row0, prob_scale
#define LAUNCH_RING(HH, N2)                                                                                                          \
    do {                                                                                                                             \
        static bool smem_set = false;                                                                                                \
        if (!smem_set) {                                                                                                             \
            cudaFuncSetAttribute(temporal_attend_ring_kernel<HH, N2, 43, 43, 100, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
            cudaFuncSetAttribute(temporal_attend_ring_kernel<HH, N2, 43, 43, 100, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            smem_set = true;                                                                                                         \
        }                                                                                                                            \
        if (k % RING_GROUP == 0) temporal_attend_ring_kernel<HH, N2, 43, 43, 100, true><<<rblocks, 128, smem, s>>>(ATTEND_RING_ARGS); \
        else temporal_attend_ring_kernel<HH, N2, 43, 43, 100, false><<<rblocks, 128, smem, s>>>(ATTEND_RING_ARGS);                    \
    } while (0)
        if (H == 2 && node2) LAUNCH_RING(2, true);
        else if (H == 2) LAUNCH_RING(2, false);
        else if (node2) LAUNCH_RING(1, true);
        else LAUNCH_RING(1, false);
#undef LAUNCH_RING
#undef ATTEND_RING_ARGS
    } else if (F + E <= 384 && T <= 128) {
#define ATTEND_SPLIT_ARGS qk, ldq, n, k, node_tab, ld_node, node_tab2, ld_node2, node_idx, F / 4, edge_tab, ld_edge, edge_idx, \
                          E / 4, time_feat, t_query, t_nbr, w, b, T, mask_ids, out_s, lds, out_scores
        if (H == 2) temporal_attend_split_kernel<2><<<blocks, 128, 0, s>>>(ATTEND_SPLIT_ARGS);
        else temporal_attend_split_kernel<1><<<blocks, 128, 0, s>>>(ATTEND_SPLIT_ARGS);
#undef ATTEND_SPLIT_ARGS
    } else if (H == 2) temporal_attend_kernel<2, 4><<<blocks, 128, 0, s>>>(ATTEND_ARGS);
    else temporal_attend_kernel<1, 4><<<blocks, 128, 0, s>>>(ATTEND_ARGS);
#undef ATTEND_ARGS
    DYG_LAUNCH_CHECK("dyg_temporal_attend");
    return 0;
}

// ------------------------------------------------------------------ DyGFormer sequence attention
// One CTA per (pair, head); K and V of the head in shared memory (row stride hd+1: conflict free for both
// the per-key dot products and the per-dim weighted sum); each warp owns query rows round-robin.
constexpr int SEQ_WARPS = 4;
constexpr int SEQ_MAXJ = 8;  // S <= 256
__global__ void __launch_bounds__(SEQ_WARPS * 32) seq_attention_kernel(const float* __restrict__ qkv, int ld_qkv, int S,
                                                                      int H, int hd, float* __restrict__ out, int ldo) {
    extern __shared__ float seq_smem[];
    const int st = hd + 1;
    float* Ks = seq_smem;
    float* Vs = Ks + (size_t)S * st;
    float* qrow = Vs + (size_t)S * st;          // SEQ_WARPS * hd
    float* prow = qrow + SEQ_WARPS * hd;        // SEQ_WARPS * S
    const int64_t b = blockIdx.x / H;
    const int h = blockIdx.x % H;
    const int D = H * hd;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = qkv + b * (int64_t)S * ld_qkv;
    for (int idx = threadIdx.x; idx < S * hd; idx += blockDim.x) {
        const int j = idx / hd, d = idx - j * hd;
        Ks[j * st + d] = base[(int64_t)j * ld_qkv + D + h * hd + d];
        Vs[j * st + d] = base[(int64_t)j * ld_qkv + 2 * D + h * hd + d];
    }
    __syncthreads();
    const float scale = rsqrtf((float)hd);
    float* myq = qrow + warp * hd;
    float* myp = prow + warp * S;
    for (int i = warp; i < S; i += SEQ_WARPS) {
        for (int d = lane; d < hd; d += 32) myq[d] = base[(int64_t)i * ld_qkv + h * hd + d] * scale;
        __syncwarp();
        float sc[SEQ_MAXJ];
        float m = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < SEQ_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float a = -INFINITY;
            if (j < S) {
                a = 0.f;
                const float* kr = Ks + j * st;
                for (int d = 0; d < hd; ++d) a = fmaf(myq[d], kr[d], a);
            }
            sc[jj] = a;
            m = fmaxf(m, a);
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < SEQ_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            const float p = (j < S) ? expf(sc[jj] - m) : 0.f;
            sc[jj] = p;
            sum += p;
        }
        const float inv = 1.f / warp_sum(sum);
#pragma unroll
        for (int jj = 0; jj < SEQ_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < S) myp[j] = sc[jj] * inv;
        }
        __syncwarp();
        for (int d = lane; d < hd; d += 32) {
            float a = 0.f;
            for (int j = 0; j < S; ++j) a = fmaf(myp[j], Vs[j * st + d], a);
            out[(b * S + i) * (int64_t)ldo + h * hd + d] = a;
        }
        __syncwarp();
    }
}

extern "C" int dyg_seq_attention(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, float* out, int ldo,
                                 dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && S > 0 && H > 0 && hd > 0, "dyg_seq_attention: bad sizes");
    DYG_CHECK_ARG(S <= 32 * SEQ_MAXJ, "dyg_seq_attention: sequence of %d tokens unsupported (max %d)", S, 32 * SEQ_MAXJ);
    const size_t smem = ((size_t)2 * S * (hd + 1) + (size_t)SEQ_WARPS * hd + (size_t)SEQ_WARPS * S) * sizeof(float);
    DYG_CHECK_ARG(smem <= 220 * 1024, "dyg_seq_attention: S=%d, head_dim=%d exceed shared memory", S, hd);
    if (B == 0) return 0;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(seq_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    seq_attention_kernel<<<(unsigned)(B * H), SEQ_WARPS * 32, smem, as_stream(stream)>>>(qkv, ld_qkv, S, H, hd, out, ldo);
    DYG_LAUNCH_CHECK("dyg_seq_attention");
    return 0;
}

// ------------------------------------------------------------------ DyGFormer sequence attention on tensor cores
// softmax(q k^T / sqrt(hd)) v for S <= 128 tokens, head_dim <= 128 (DyGFormer: S <= 64, hd = 100).  One CTA per
// (pair, head); K and V rows of the head are split into bf16 hi | mid on the way into shared memory; every warp owns
// 16-query blocks and runs both products as BF16x3 (hi*hi + hi*mid + mid*hi, fp32 accumulate) with register-resident
// m16n8k16 fragments: scores -> softmax in the accumulator layout -> probabilities reused in place as the A operand
// of P V (V read with ldmatrix.trans).  ~1e-5 relative, like the tcgen05 GEMMs around it.  The two products are 5 % of
// the transformer's flops on 64-token tiles; the dense projections run on tcgen05 (gemm_tc.cu).
#include <cuda_bf16.h>

namespace {

// (a, b) -> packed bf16 pairs hi = bf16(x), mid = bf16(x - hi): one packed convert per plane, the bf16 -> fp32 widening is a
// shift / mask (same values as two scalar conversions per element, 7 instructions instead of 12)
__device__ __forceinline__ void sa_split(float a, float b, uint32_t& hi, uint32_t& mid) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 m = __floats2bfloat162_rn(a - ah, b - bh);
    mid = *reinterpret_cast<const uint32_t*>(&m);
}
__device__ __forceinline__ void sa_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sa_ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}

// NKT: 8-key tiles (S <= 8*NKT, even), KDT: 16-wide head-dim steps (hd <= 16*KDT)
template <int NKT, int KDT>
__global__ void __launch_bounds__(128) seq_attention_mma_kernel(const float* __restrict__ qkv, int ld_qkv, int S, int H, int hd,
                                                                float* __restrict__ out, int ldo, __nv_bfloat16* __restrict__ ohi,
                                                                __nv_bfloat16* __restrict__ omid, int ldos) {
    constexpr int SK = NKT * 8, KD = KDT * 16, KS = KD + 8;   // KS/2 = 4 (mod 8): conflict-free fragment loads
    extern __shared__ __align__(16) unsigned char sa_smem[];
    __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(sa_smem);
    __nv_bfloat16* Km = Kh + SK * KS;
    __nv_bfloat16* Vh = Km + SK * KS;
    __nv_bfloat16* Vm = Vh + SK * KS;
    const int64_t b = blockIdx.x / H;
    const int h = blockIdx.x % H;
    const int D = H * hd;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const float* base = qkv + b * (int64_t)S * ld_qkv;
    // stage K and V: all loads of a round are issued before the first conversion (the loop is DRAM-latency bound otherwise)
    constexpr int TOT = SK * (KD / 2);
    constexpr int U = 14;
#pragma unroll 1
    for (int it0 = 0; it0 < TOT; it0 += 128 * U) {
        float2 kv[U], vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = it0 + u * 128 + tid;
            const int j = idx / (KD / 2), d = (idx - j * (KD / 2)) * 2;
            kv[u] = make_float2(0.f, 0.f);
            vv[u] = kv[u];
            if (idx < TOT && j < S && d < hd) {
                const float* rowp = base + (int64_t)j * ld_qkv + h * hd + d;
                kv[u] = *reinterpret_cast<const float2*>(rowp + D);
                vv[u] = *reinterpret_cast<const float2*>(rowp + 2 * D);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = it0 + u * 128 + tid;
            if (idx < TOT) {
                const int j = idx / (KD / 2), d = (idx - j * (KD / 2)) * 2;
                uint32_t hi, mid;
                sa_split(kv[u].x, kv[u].y, hi, mid);
                *reinterpret_cast<uint32_t*>(Kh + j * KS + d) = hi;
                *reinterpret_cast<uint32_t*>(Km + j * KS + d) = mid;
                sa_split(vv[u].x, vv[u].y, hi, mid);
                *reinterpret_cast<uint32_t*>(Vh + j * KS + d) = hi;
                *reinterpret_cast<uint32_t*>(Vm + j * KS + d) = mid;
            }
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)hd);
    for (int rb = warp; rb * 16 < S; rb += 4) {
        const int r0 = rb * 16 + g, r1 = r0 + 8;
        const float* q0 = base + (int64_t)min(r0, S - 1) * ld_qkv + h * hd;
        const float* q1 = base + (int64_t)min(r1, S - 1) * ld_qkv + h * hd;
        float sc[NKT][4];
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KDT; ++ks) {
            const int d0 = ks * 16 + 2 * t, d1 = d0 + 8;
            const float2 z = make_float2(0.f, 0.f);
            const float2 q00 = d0 < hd ? *reinterpret_cast<const float2*>(q0 + d0) : z;
            const float2 q10 = d0 < hd ? *reinterpret_cast<const float2*>(q1 + d0) : z;
            const float2 q01 = d1 < hd ? *reinterpret_cast<const float2*>(q0 + d1) : z;
            const float2 q11 = d1 < hd ? *reinterpret_cast<const float2*>(q1 + d1) : z;
            uint32_t ah[4], am[4];
            sa_split(q00.x, q00.y, ah[0], am[0]);
            sa_split(q10.x, q10.y, ah[1], am[1]);
            sa_split(q01.x, q01.y, ah[2], am[2]);
            sa_split(q11.x, q11.y, ah[3], am[3]);
#pragma unroll
            for (int nt = 0; nt < NKT; ++nt) {
                const int off = (nt * 8 + g) * KS + ks * 16 + 2 * t;
                const uint32_t b0h = *reinterpret_cast<const uint32_t*>(Kh + off), b1h = *reinterpret_cast<const uint32_t*>(Kh + off + 8);
                const uint32_t b0m = *reinterpret_cast<const uint32_t*>(Km + off), b1m = *reinterpret_cast<const uint32_t*>(Km + off + 8);
                sa_mma(sc[nt], ah, b0h, b1h);
                sa_mma(sc[nt], ah, b0m, b1m);
                sa_mma(sc[nt], am, b0h, b1h);
            }
        }
        // softmax over the keys: rows r0 (c0,c1) and r1 (c2,c3); the four lanes of a quad share a row
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
            const int key = nt * 8 + 2 * t;
            sc[nt][0] = key < S ? sc[nt][0] * scale : -INFINITY;
            sc[nt][1] = key + 1 < S ? sc[nt][1] * scale : -INFINITY;
            sc[nt][2] = key < S ? sc[nt][2] * scale : -INFINITY;
            sc[nt][3] = key + 1 < S ? sc[nt][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
            sc[nt][0] = expf(sc[nt][0] - mx0);
            sc[nt][1] = expf(sc[nt][1] - mx0);
            sc[nt][2] = expf(sc[nt][2] - mx1);
            sc[nt][3] = expf(sc[nt][3] - mx1);
            sum0 += sc[nt][0] + sc[nt][1];
            sum1 += sc[nt][2] + sc[nt][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        // P V: the score accumulators of two adjacent key tiles are exactly one A fragment (16 queries x 16 keys)
        float o[2 * KDT][4];
#pragma unroll
        for (int dt = 0; dt < 2 * KDT; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < NKT / 2; ++kk) {
            uint32_t ph[4], pm[4];
            sa_split(sc[2 * kk][0], sc[2 * kk][1], ph[0], pm[0]);
            sa_split(sc[2 * kk][2], sc[2 * kk][3], ph[1], pm[1]);
            sa_split(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pm[2]);
            sa_split(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pm[3]);
            const int voff = (kk * 16 + (lane & 15)) * KS;
#pragma unroll
            for (int dt = 0; dt < 2 * KDT; ++dt) {
                uint32_t b0h, b1h, b0m, b1m;
                sa_ldmatrix_x2_trans(b0h, b1h, Vh + voff + dt * 8);
                sa_ldmatrix_x2_trans(b0m, b1m, Vm + voff + dt * 8);
                sa_mma(o[dt], ph, b0h, b1h);
                sa_mma(o[dt], ph, b0m, b1m);
                sa_mma(o[dt], pm, b0h, b1h);
            }
        }
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
#pragma unroll
        for (int dt = 0; dt < 2 * KDT; ++dt) {
            const int d = dt * 8 + 2 * t;
            if (d >= hd) continue;
            const float a0 = o[dt][0] * inv0, a1 = o[dt][1] * inv0, c0 = o[dt][2] * inv1, c1 = o[dt][3] * inv1;
            if (r0 < S) {
                const int64_t row = b * S + r0;
                if (out) *reinterpret_cast<float2*>(out + row * ldo + h * hd + d) = make_float2(a0, a1);
                if (ohi) {
                    uint32_t hi, mid;
                    sa_split(a0, a1, hi, mid);
                    *reinterpret_cast<uint32_t*>(ohi + row * ldos + h * hd + d) = hi;
                    *reinterpret_cast<uint32_t*>(omid + row * ldos + h * hd + d) = mid;
                }
            }
            if (r1 < S) {
                const int64_t row = b * S + r1;
                if (out) *reinterpret_cast<float2*>(out + row * ldo + h * hd + d) = make_float2(c0, c1);
                if (ohi) {
                    uint32_t hi, mid;
                    sa_split(c0, c1, hi, mid);
                    *reinterpret_cast<uint32_t*>(ohi + row * ldos + h * hd + d) = hi;
                    *reinterpret_cast<uint32_t*>(omid + row * ldos + h * hd + d) = mid;
                }
            }
        }
    }
}

template <int NKT, int KDT>
int launch_seq_mma(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, float* out, int ldo, void* ohi, void* omid,
                   int ldos, cudaStream_t s) {
    const size_t smem = (size_t)4 * (NKT * 8) * (KDT * 16 + 8) * sizeof(__nv_bfloat16);
    static bool configured = false;
    if (!configured && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(seq_attention_mma_kernel<NKT, KDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_seq_attention_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = true;
    }
    seq_attention_mma_kernel<NKT, KDT><<<(unsigned)(B * H), 128, smem, s>>>(qkv, ld_qkv, S, H, hd, out, ldo,
                                                                           reinterpret_cast<__nv_bfloat16*>(ohi),
                                                                           reinterpret_cast<__nv_bfloat16*>(omid), ldos);
    return 0;
}

}  // namespace

extern "C" int dyg_seq_attention_tc(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, float* out, int ldo,
                                    void* out_hi, void* out_mid, int ldos, dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && S > 0 && H > 0 && hd > 0, "dyg_seq_attention_tc: bad sizes");
    DYG_CHECK_ARG(S <= 128 && hd <= 128 && (hd % 2) == 0, "dyg_seq_attention_tc: S=%d (max 128), head_dim=%d (even, max 128) unsupported", S, hd);
    if (B == 0) return 0;
    DYG_CHECK_ARG((ld_qkv % 2) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 7u) == 0, "dyg_seq_attention_tc: qkv must be 8-byte aligned with an even leading dimension");
    DYG_CHECK_ARG(out || (out_hi && out_mid), "dyg_seq_attention_tc: no output given");
    DYG_CHECK_ARG((out_hi == nullptr) == (out_mid == nullptr), "dyg_seq_attention_tc: out_hi and out_mid go together");
    DYG_CHECK_ARG(!out || ((ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0), "dyg_seq_attention_tc: out must be 8-byte aligned, ldo even");
    DYG_CHECK_ARG(!out_hi || (ldos % 2) == 0, "dyg_seq_attention_tc: ldos must be even");
    DYG_CHECK_ARG(B * H < ((int64_t)1 << 31), "dyg_seq_attention_tc: too many (pair, head) tiles");
    if (B == 0) return 0;
    cudaStream_t s = as_stream(stream);
    int rc;
#define SA_GO(NKT, KDT) rc = launch_seq_mma<NKT, KDT>(qkv, ld_qkv, B, S, H, hd, out, ldo, out_hi, out_mid, ldos, s)
#define SA_KD(NKT)                     \
    do {                               \
        if (hd <= 32) SA_GO(NKT, 2);   \
        else if (hd <= 64) SA_GO(NKT, 4); \
        else if (hd <= 112) SA_GO(NKT, 7); \
        else SA_GO(NKT, 8);            \
    } while (0)
    if (S <= 32) SA_KD(4);
    else if (S <= 64) SA_KD(8);
    else SA_KD(16);
#undef SA_KD
#undef SA_GO
    if (rc != 0) return rc;
    DYG_LAUNCH_CHECK("dyg_seq_attention_tc");
    return 0;
}
