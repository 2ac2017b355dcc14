// Backward kernels of DyGFormer's training path (train_link_prediction.py:230-257 differentiates the reference's eager
// modules; SURVEY.md 8(b): dyg_patch_project_bwd / dyg_tfm_block_bwd).  The forward of a training step runs on the same
// tcgen05 GEMMs as evaluation; what autograd needs on top is here, in fp32:
//   dyg_gemm_dw            dW += G^T X, db += column sums of G        (weight / bias gradient of a dense layer, split over the rows: FFMA tiles for
//                          short contractions, BF16x3 mma.sync with ldmatrix.trans operands from 512 rows on)
//   dyg_linear_bwd         dX, dW, db of a small layer in ONE launch, ReLU mask applied on the way in
//   dyg_gemm_dx            dX = (G * mask) W in one launch: BF16x3 mma.sync, W read with ldmatrix.trans (very large layers:
//                          dyg_gemm_bf16x3 on planes of G and W^T)
//   dyg_layernorm_bwd      dx, dgamma, dbeta of y = LayerNorm(x) gamma + beta                 (models/DyGFormer.py:447, 456)
//   dyg_gelu_fwd / _bwd    h = gelu(v) * mask -> operand planes, dv = dh * mask * gelu'(v)     (models/DyGFormer.py:458)
//   dyg_seq_attention_train_fwd / _bwd   softmax(q k^T / sqrt(hd)) (with dropout multipliers) v and its gradient (:454)
// These are small, launch-bound shapes (a 200-event batch = 25,600 tokens): plain shared-memory FFMA tiles, exact fp32.
#include <math.h>
#include <string.h>

#include "tc_common.cuh"

namespace {

// ------------------------------------------------------------------ dW += G^T X  (N x K), db += sum_m G
constexpr int DW_T = 64;        // output tile (n, k)
constexpr int DW_MC = 32;       // rows per shared-memory chunk
__global__ void __launch_bounds__(256) gemm_dw_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx,
                                                      int64_t M, int N, int K, float* __restrict__ dW, int ldw, float* __restrict__ db,
                                                      int64_t rows_per_cta) {
    __shared__ float sg[DW_MC][DW_T + 1];
    __shared__ float sx[DW_MC][DW_T + 1];
    const int n0 = blockIdx.x * DW_T, k0 = blockIdx.y * DW_T;
    const int64_t m_begin = blockIdx.z * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
    const int tid = threadIdx.x;
    const int tn = (tid >> 4) * 4, tk = (tid & 15) * 4;          // this thread's 4 x 4 outputs
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    float bsum = 0.f;                                             // column n0 + tid of G (threads 0..63, k tile 0 only)
    for (int64_t m0 = m_begin; m0 < m_end; m0 += DW_MC) {
        for (int i = tid; i < DW_MC * DW_T; i += 256) {
            const int r = i / DW_T, c = i - r * DW_T;
            const int64_t m = m0 + r;
            sg[r][c] = (m < m_end && n0 + c < N) ? G[m * ldg + n0 + c] : 0.f;
            sx[r][c] = (m < m_end && k0 + c < K) ? X[m * ldx + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < DW_MC; ++r) {
            float g[4], x[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                g[a] = sg[r][tn + a];
                x[a] = sx[r][tk + a];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(g[a], x[b], acc[a][b]);
        }
        if (db && blockIdx.y == 0 && tid < DW_T) {
#pragma unroll 8
            for (int r = 0; r < DW_MC; ++r) bsum += sg[r][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (n0 + tn + a < N && k0 + tk + b < K) atomicAdd(dW + (int64_t)(n0 + tn + a) * ldw + k0 + tk + b, acc[a][b]);
    if (db && blockIdx.y == 0 && tid < DW_T && n0 + tid < N) atomicAdd(db + n0 + tid, bsum);
}

// ------------------------------------------------------------------ dW on the warp-level tensor cores (long contractions)
// dW[n, k] += sum_m G[m, n] X[m, k] as BF16x3 mma.sync.m16n8k16: the "M" side of the MMA is n, its contraction is the row index m.
// G and X chunks (32 rows) are read as fp32, split into bf16 hi | mid on the way into shared memory ([m][n] / [m][k], the layout
// they have in memory), and BOTH operands are fetched with ldmatrix.trans (the contraction index is the slow one in both).
// CTA = 4 warps, 64 (n) x 64 (k) outputs, warp tile 32 x 32; rows split over blockIdx.z with atomic accumulation; the bias
// gradient (column sums of G) falls out of the loads.  ~5x the rate of the FFMA kernel above on 17 k-row layers.
constexpr int DM_T = 64, DM_MC = 32, DM_P = DM_T + 8;     // pitch 72 bf16 = 144 B: conflict-free ldmatrix rows
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128) gemm_dw_mma_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx,
                                                          int64_t M, int N, int K, float* __restrict__ dW, int ldw, float* __restrict__ db,
                                                          int64_t rows_per_cta) {
    __shared__ __align__(16) __nv_bfloat16 sg[2][DM_MC][DM_P];   // [plane][m][n]
    __shared__ __align__(16) __nv_bfloat16 sx[2][DM_MC][DM_P];   // [plane][m][k]
    const int n0 = blockIdx.x * DM_T, k0 = blockIdx.y * DM_T;
    const int64_t m_begin = blockIdx.z * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wn = (warp >> 1) * 32, wk = (warp & 1) * 32;        // this warp's 32 x 32 outputs
    // loads: thread t covers columns 4 (t % 16) .. + 3 of rows t / 16 + 8 i (i < 4) of both chunks
    const int lc = (tid & 15) * 4, lr = tid >> 4;
    const bool vec = ((ldg & 3) == 0) && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15u) == 0);
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    float4 gv[4], xv[4];
    auto fetch = [&](int64_t m0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + lr + 8 * i;
            gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            xv[i] = gv[i];
            if (m < m_end) {
                const float* gp = G + m * ldg + n0 + lc;
                const float* xp = X + m * ldx + k0 + lc;
                if (vec && n0 + lc + 3 < N) gv[i] = *reinterpret_cast<const float4*>(gp);
                else {
                    if (n0 + lc < N) gv[i].x = gp[0];
                    if (n0 + lc + 1 < N) gv[i].y = gp[1];
                    if (n0 + lc + 2 < N) gv[i].z = gp[2];
                    if (n0 + lc + 3 < N) gv[i].w = gp[3];
                }
                if (vec && k0 + lc + 3 < K) xv[i] = *reinterpret_cast<const float4*>(xp);
                else {
                    if (k0 + lc < K) xv[i].x = xp[0];
                    if (k0 + lc + 1 < K) xv[i].y = xp[1];
                    if (k0 + lc + 2 < K) xv[i].z = xp[2];
                    if (k0 + lc + 3 < K) xv[i].w = xp[3];
                }
            }
        }
    };
    fetch(m_begin);
    for (int64_t m0 = m_begin; m0 < m_end; m0 += DM_MC) {
        __syncthreads();                                          // the previous chunk's fragments have been read
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lr + 8 * i;
            uint32_t h0, l0, h1, l1;
            split_pack(gv[i].x, gv[i].y, h0, l0);
            split_pack(gv[i].z, gv[i].w, h1, l1);
            *reinterpret_cast<uint2*>(&sg[0][r][lc]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(&sg[1][r][lc]) = make_uint2(l0, l1);
            bs[0] += gv[i].x; bs[1] += gv[i].y; bs[2] += gv[i].z; bs[3] += gv[i].w;
            split_pack(xv[i].x, xv[i].y, h0, l0);
            split_pack(xv[i].z, xv[i].w, h1, l1);
            *reinterpret_cast<uint2*>(&sx[0][r][lc]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(&sx[1][r][lc]) = make_uint2(l0, l1);
        }
        __syncthreads();
        if (m0 + DM_MC < m_end) fetch(m0 + DM_MC);                // in flight while this chunk multiplies
#pragma unroll
        for (int ks = 0; ks < DM_MC / 16; ++ks) {
            const int mr = ks * 16;
            uint32_t ah[2][4], am[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                // A^T block: smem rows m, columns n; matrices (m 0-7, n 0-7), (m 0-7, n 8-15), (m 8-15, n 0-7), (m 8-15, n 8-15)
                const int row = mr + (lane & 7) + ((lane >> 4) & 1) * 8, col = wn + mi * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(ah[mi], &sg[0][row][col]);
                ldsm_x4_t(am[mi], &sg[1][row][col]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ni += 2) {
                // B block pair: matrices (m 0-7, k ni*8), (m 8-15, k ni*8), (m 0-7, k ni*8+8), (m 8-15, k ni*8+8)
                const int row = mr + (lane & 7) + ((lane >> 3) & 1) * 8, col = wk + ni * 8 + (lane >> 4) * 8;
                uint32_t bh[4], bm[4];
                ldsm_x4_t(bh, &sx[0][row][col]);
                ldsm_x4_t(bm, &sx[1][row][col]);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_bf16(acc[mi][ni], ah[mi], bh[0], bh[1]);
                    mma_bf16(acc[mi][ni], ah[mi], bm[0], bm[1]);
                    mma_bf16(acc[mi][ni], am[mi], bh[0], bh[1]);
                    mma_bf16(acc[mi][ni + 1], ah[mi], bh[2], bh[3]);
                    mma_bf16(acc[mi][ni + 1], ah[mi], bm[2], bm[3]);
                    mma_bf16(acc[mi][ni + 1], am[mi], bh[2], bh[3]);
                }
            }
        }
    }
    // accumulator layout of m16n8: c0,c1 -> row g, cols 2t, 2t+1; c2,c3 -> row g + 8
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = n0 + wn + mi * 16 + g + (c >> 1) * 8, k = k0 + wk + ni * 8 + 2 * t + (c & 1);
                if (n < N && k < K) atomicAdd(dW + (int64_t)n * ldw + k, acc[mi][ni][c]);
            }
    if (db && blockIdx.y == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n0 + lc + j < N) atomicAdd(db + n0 + lc + j, bs[j]);
    }
}

// ------------------------------------------------------------------ dX on the warp-level tensor cores (mid-size layers, one launch)
// dX[m, k] = sum_n G[m, n] W[n, k] as BF16x3 mma.sync: G rows are the A operand as they lie in memory (plain ldmatrix), W (N, K) is read
// with ldmatrix.trans (its row index is the contraction).  fp32 operands are split into bf16 hi | mid on the way into shared memory,
// so no operand planes and no transposed copy of W are needed; with `Y` the gradient is masked on the way in (ReLU layers).
constexpr int DXM_P = DM_MC + 8;                         // pitch of the G tile: 40 bf16 = 80 B
__device__ __forceinline__ void ldsm_x4_n(uint32_t (&r)[4], const __nv_bfloat16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__global__ void __launch_bounds__(128) gemm_dx_mma_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Y, int ldy,
                                                          const float* __restrict__ W, int ldw, int64_t M, int N, int K,
                                                          float* __restrict__ dX, int lddx) {
    __shared__ __align__(16) __nv_bfloat16 sg[2][DM_T][DXM_P];    // [plane][m][n chunk]
    __shared__ __align__(16) __nv_bfloat16 sw[2][DM_MC][DM_P];    // [plane][n chunk][k]
    const int64_t m0 = (int64_t)blockIdx.x * DM_T;
    const int k0 = blockIdx.y * DM_T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp >> 1) * 32, wk = (warp & 1) * 32;
    const int gc = (tid & 7) * 4, gr = tid >> 3;                  // G tile: columns 4 (t % 8) .. + 3 of rows t / 8 + 16 i
    const int wc = (tid & 15) * 4, wr = tid >> 4;                 // W chunk: columns 4 (t % 16) .. + 3 of rows t / 16 + 8 i
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    float gv[4][4], wv[4][4];
    auto fetch = [&](int n0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + gr + 16 * i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + gc + j;
                float g = 0.f;
                if (m < M && n < N) {
                    g = G[m * ldg + n];
                    if (Y && !(Y[m * ldy + n] > 0.f)) g = 0.f;
                }
                gv[i][j] = g;
                const int nn = n0 + wr + 8 * i, kk = k0 + wc + j;
                wv[i][j] = (nn < N && kk < K) ? W[(int64_t)nn * ldw + kk] : 0.f;
            }
        }
    };
    fetch(0);
    for (int n0 = 0; n0 < N; n0 += DM_MC) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t h0, l0, h1, l1;
            split_pack(gv[i][0], gv[i][1], h0, l0);
            split_pack(gv[i][2], gv[i][3], h1, l1);
            *reinterpret_cast<uint2*>(&sg[0][gr + 16 * i][gc]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(&sg[1][gr + 16 * i][gc]) = make_uint2(l0, l1);
            split_pack(wv[i][0], wv[i][1], h0, l0);
            split_pack(wv[i][2], wv[i][3], h1, l1);
            *reinterpret_cast<uint2*>(&sw[0][wr + 8 * i][wc]) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(&sw[1][wr + 8 * i][wc]) = make_uint2(l0, l1);
        }
        __syncthreads();
        if (n0 + DM_MC < N) fetch(n0 + DM_MC);
#pragma unroll
        for (int ks = 0; ks < DM_MC / 16; ++ks) {
            const int nr = ks * 16;
            uint32_t ah[2][4], am[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                // A block (16 m x 16 n): lanes 0-15 rows m at n, lanes 16-31 the same rows at n + 8
                const int row = wm + mi * 16 + (lane & 15), col = nr + (lane >> 4) * 8;
                ldsm_x4_n(ah[mi], &sg[0][row][col]);
                ldsm_x4_n(am[mi], &sg[1][row][col]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ni += 2) {
                const int row = nr + (lane & 7) + ((lane >> 3) & 1) * 8, col = wk + ni * 8 + (lane >> 4) * 8;
                uint32_t bh[4], bm[4];
                ldsm_x4_t(bh, &sw[0][row][col]);
                ldsm_x4_t(bm, &sw[1][row][col]);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_bf16(acc[mi][ni], ah[mi], bh[0], bh[1]);
                    mma_bf16(acc[mi][ni], ah[mi], bm[0], bm[1]);
                    mma_bf16(acc[mi][ni], am[mi], bh[0], bh[1]);
                    mma_bf16(acc[mi][ni + 1], ah[mi], bh[2], bh[3]);
                    mma_bf16(acc[mi][ni + 1], ah[mi], bm[2], bm[3]);
                    mma_bf16(acc[mi][ni + 1], am[mi], bh[2], bh[3]);
                }
            }
        }
    }
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t m = m0 + wm + mi * 16 + g + (c >> 1) * 8;
                const int k = k0 + wk + ni * 8 + 2 * t + (c & 1);
                if (m < M && k < K) dX[m * lddx + k] = acc[mi][ni][c];
            }
}

// ------------------------------------------------------------------ dX and dW (+ db) of a small dense layer in ONE launch
// Blocks [0, dx_tiles) compute 64 x 64 tiles of dX = G W, the others 64 x 64 tiles of dW += G^T X over a slice of the rows (and db).
// With `Y` the gradient is masked on the way in (ReLU layers: G * (Y > 0)), so the mask needs no launch of its own.
__device__ __forceinline__ float lb_g(const float* __restrict__ G, int ldg, const float* __restrict__ Y, int ldy, int64_t m, int n) {
    const float g = G[m * ldg + n];
    return (Y && !(Y[m * ldy + n] > 0.f)) ? 0.f : g;
}
__global__ void __launch_bounds__(256) linear_bwd_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Y, int ldy,
                                                         const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, int64_t M,
                                                         int N, int K, float* __restrict__ dX, int lddx, float* __restrict__ dW, int lddw,
                                                         float* __restrict__ db, int dx_tiles, int dx_ktiles, int dw_ktiles, int dw_tiles,
                                                         int64_t rows_per_cta) {
    __shared__ float sa[DW_T][DW_MC + 1];
    __shared__ float sb[DW_MC][DW_T + 1];
    const int tid = threadIdx.x;
    const int ta = (tid >> 4) * 4, tb = (tid & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    if ((int)blockIdx.x < dx_tiles) {
        // ---- dX tile (m0, k0): contraction over n
        const int64_t m0 = (int64_t)(blockIdx.x / dx_ktiles) * DW_T;
        const int k0 = (blockIdx.x % dx_ktiles) * DW_T;
        for (int n0 = 0; n0 < N; n0 += DW_MC) {
            for (int i = tid; i < DW_T * DW_MC; i += 256) {
                const int r = i / DW_MC, c = i - r * DW_MC;
                sa[r][c] = (m0 + r < M && n0 + c < N) ? lb_g(G, ldg, Y, ldy, m0 + r, n0 + c) : 0.f;
            }
            for (int i = tid; i < DW_MC * DW_T; i += 256) {
                const int r = i / DW_T, c = i - r * DW_T;
                sb[r][c] = (n0 + r < N && k0 + c < K) ? W[(int64_t)(n0 + r) * ldw + k0 + c] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int n = 0; n < DW_MC; ++n) {
                float g[4], w[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    g[a] = sa[ta + a][n];
                    w[a] = sb[n][tb + a];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(g[a], w[b], acc[a][b]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (m0 + ta + a < M && k0 + tb + b < K) dX[(m0 + ta + a) * lddx + k0 + tb + b] = acc[a][b];
    } else {
        // ---- dW tile (n0, k0) over rows [m_begin, m_end): contraction over m; sa holds G^T chunks as [n][m]
        const int t = (int)blockIdx.x - dx_tiles;
        const int tile = t % dw_tiles, split = t / dw_tiles;
        const int n0 = (tile / dw_ktiles) * DW_T, k0 = (tile % dw_ktiles) * DW_T;
        const int64_t m_begin = split * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
        float bsum = 0.f;
        for (int64_t m0 = m_begin; m0 < m_end; m0 += DW_MC) {
            for (int i = tid; i < DW_MC * DW_T; i += 256) {
                const int r = i / DW_T, c = i - r * DW_T;         // r: row m of the chunk, c: n / k of the tile
                const int64_t m = m0 + r;
                sa[c][r] = (m < m_end && n0 + c < N) ? lb_g(G, ldg, Y, ldy, m, n0 + c) : 0.f;
                sb[r][c] = (m < m_end && k0 + c < K) ? X[m * ldx + k0 + c] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int r = 0; r < DW_MC; ++r) {
                float g[4], x[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    g[a] = sa[ta + a][r];
                    x[a] = sb[r][tb + a];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(g[a], x[b], acc[a][b]);
            }
            if (db && k0 == 0 && tid < DW_T) {
#pragma unroll 8
                for (int r = 0; r < DW_MC; ++r) bsum += sa[tid][r];
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (n0 + ta + a < N && k0 + tb + b < K) atomicAdd(dW + (int64_t)(n0 + ta + a) * lddw + k0 + tb + b, acc[a][b]);
        if (db && k0 == 0 && tid < DW_T && n0 + tid < N) atomicAdd(db + n0 + tid, bsum);
    }
}

// ------------------------------------------------------------------ LayerNorm backward: one warp per row, rows strided over the grid
template <int MAXP>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, float eps,
                                                            const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int D) {
    extern __shared__ float ln_part[];                           // [2][D] block partials of dgamma | dbeta
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) ln_part[i] = 0.f;
    __syncthreads();
    float gsum[MAXP], bsum[MAXP], gm[MAXP];
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int c = lane + 32 * i;
        gsum[i] = 0.f;
        bsum[i] = 0.f;
        gm[i] = c < D ? __ldg(gamma + c) : 0.f;
    }
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; m < M; m += nw) {
        float xv[MAXP], gv[MAXP];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int c = lane + 32 * i;
            xv[i] = c < D ? x[m * ldx + c] : 0.f;
            gv[i] = c < D ? dy[m * lddy + c] : 0.f;
            s += xv[i];
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int c = lane + 32 * i;
            const float d = xv[i] - mean;
            q += c < D ? d * d : 0.f;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        float s1 = 0.f, s2 = 0.f;                                 // sum of g, sum of g * xhat with g = dy * gamma
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const float xh = (xv[i] - mean) * rstd;
            const float g = gv[i] * gm[i];
            s1 += g;
            s2 += g * xh;
            gsum[i] += gv[i] * xh;
            bsum[i] += gv[i];
            xv[i] = xh;
        }
        s1 = warp_sum(s1) / (float)D;
        s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int c = lane + 32 * i;
            if (c < D) dx[m * lddx + c] = rstd * (gv[i] * gm[i] - s1 - xv[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
            atomicAdd(ln_part + c, gsum[i]);
            atomicAdd(ln_part + D + c, bsum[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + i, ln_part[i]);
        if (dbeta) atomicAdd(dbeta + i, ln_part[D + i]);
    }
}

// ------------------------------------------------------------------ GELU (exact erf) forward into operand planes / backward
__global__ void gelu_fwd_kernel(const float* __restrict__ v, int ldv, const float* __restrict__ mask, int ldm, float* __restrict__ h, int ldh,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid, int lds, int64_t M, int N) {
    const int64_t pairs = (int64_t)(N / 2);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M * pairs) return;
    const int64_t m = i / pairs;
    const int c = (int)(i - m * pairs) * 2;
    const float2 a = *reinterpret_cast<const float2*>(v + m * ldv + c);
    float y0 = 0.5f * a.x * (1.f + erff(a.x * 0.70710678118654752440f));
    float y1 = 0.5f * a.y * (1.f + erff(a.y * 0.70710678118654752440f));
    if (mask) {
        const float2 k = *reinterpret_cast<const float2*>(mask + m * ldm + c);
        y0 *= k.x;
        y1 *= k.y;
    }
    if (h) *reinterpret_cast<float2*>(h + m * ldh + c) = make_float2(y0, y1);
    if (hi) {
        uint32_t ph, pm;
        split_pack(y0, y1, ph, pm);
        *reinterpret_cast<uint32_t*>(hi + m * lds + c) = ph;
        *reinterpret_cast<uint32_t*>(mid + m * lds + c) = pm;
    }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ v, int ldv, const float* __restrict__ mask, int ldm, const float* __restrict__ dh,
                                int lddh, float* __restrict__ dv, int lddv, int64_t M, int N) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N;
    const int c = (int)(i - m * N);
    const float a = v[m * ldv + c];
    // d/dv [v Phi(v)] = Phi(v) + v phi(v)
    const float cdf = 0.5f * (1.f + erff(a * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * a * a);
    float g = dh[m * lddh + c] * (cdf + a * pdf);
    if (mask) g *= mask[m * ldm + c];
    dv[m * lddv + c] = g;
}

// ------------------------------------------------------------------ sequence attention with saved probabilities (training)
// One CTA per (sequence, head); q, k, v rows of the head in shared memory (fp32); S <= 64, hd <= 128.
constexpr int TA_S = 64;
template <bool BWD>
__global__ void __launch_bounds__(256) seq_attention_train_kernel(const float* __restrict__ qkv, int ld_qkv, int S, int H, int hd,
                                                                  const float* __restrict__ pmask,   // (B, H, S, S) dropout multipliers or NULL
                                                                  float* __restrict__ probs,          // (B, H, S, S): written (fwd) / read (bwd)
                                                                  float* __restrict__ out, int ldo,  // fwd: (B*S, H*hd) attention output
                                                                  const float* __restrict__ dout, int lddo, float* __restrict__ dqkv, int lddq) {
    extern __shared__ float ta_smem[];
    const int HS = hd + 1;                                       // row stride: conflict-free column walks
    float* sq = ta_smem;
    float* sk = sq + TA_S * HS;
    float* sv = sk + TA_S * HS;
    float* sp = sv + TA_S * HS;                                  // [S][S+1] probabilities (fwd) / dS (bwd)
    float* sd = sp + TA_S * (TA_S + 1);                          // bwd: dOut rows of the head
    float* sm = sd + (BWD ? TA_S * HS : 0);                      // bwd: P * mask
    const int64_t b = blockIdx.x / H;
    const int h = blockIdx.x % H;
    const int D = H * hd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* base = qkv + b * (int64_t)S * ld_qkv + h * hd;
    for (int i = tid; i < S * hd; i += 256) {
        const int r = i / hd, d = i - r * hd;
        const float* rp = base + (int64_t)r * ld_qkv + d;
        sq[r * HS + d] = rp[0];
        sk[r * HS + d] = rp[D];
        sv[r * HS + d] = rp[2 * D];
        if (BWD) sd[r * HS + d] = dout[(b * S + r) * (int64_t)lddo + h * hd + d];
    }
    __syncthreads();
    const float scale = rsqrtf((float)hd);
    const int64_t pbase = (b * H + h) * (int64_t)S * S;
    if (!BWD) {
        for (int i = tid; i < S * S; i += 256) {
            const int r = i / S, c = i - r * S;
            float a = 0.f;
            for (int d = 0; d < hd; ++d) a = fmaf(sq[r * HS + d], sk[c * HS + d], a);
            sp[r * (TA_S + 1) + c] = a * scale;
        }
        __syncthreads();
        for (int r = warp; r < S; r += 8) {                      // softmax of row r by one warp
            float v0 = lane < S ? sp[r * (TA_S + 1) + lane] : -INFINITY;
            float v1 = lane + 32 < S ? sp[r * (TA_S + 1) + lane + 32] : -INFINITY;
            const float mx = warp_max(fmaxf(v0, v1));
            v0 = lane < S ? expf(v0 - mx) : 0.f;
            v1 = lane + 32 < S ? expf(v1 - mx) : 0.f;
            const float inv = 1.f / warp_sum(v0 + v1);
            v0 *= inv;
            v1 *= inv;
            if (lane < S) {
                probs[pbase + (int64_t)r * S + lane] = v0;
                sp[r * (TA_S + 1) + lane] = pmask ? v0 * pmask[pbase + (int64_t)r * S + lane] : v0;
            }
            if (lane + 32 < S) {
                probs[pbase + (int64_t)r * S + lane + 32] = v1;
                sp[r * (TA_S + 1) + lane + 32] = pmask ? v1 * pmask[pbase + (int64_t)r * S + lane + 32] : v1;
            }
        }
        __syncthreads();
        for (int i = tid; i < S * hd; i += 256) {
            const int r = i / hd, d = i - r * hd;
            float a = 0.f;
            for (int c = 0; c < S; ++c) a = fmaf(sp[r * (TA_S + 1) + c], sv[c * HS + d], a);
            out[(b * S + r) * (int64_t)ldo + h * hd + d] = a;
        }
    } else {
        // dPd = dOut V^T; dP = dPd * mask; r_i = sum_j dP_ij P_ij; dS = P (dP - r_i); sm = P * mask
        for (int i = tid; i < S * S; i += 256) {
            const int r = i / S, c = i - r * S;
            float a = 0.f;
            for (int d = 0; d < hd; ++d) a = fmaf(sd[r * HS + d], sv[c * HS + d], a);
            const float mk = pmask ? pmask[pbase + i] : 1.f;
            const float p = probs[pbase + i];
            sp[r * (TA_S + 1) + c] = a * mk;                     // dP
            sm[r * (TA_S + 1) + c] = p * mk;
        }
        __syncthreads();
        for (int r = warp; r < S; r += 8) {
            const float p0 = lane < S ? probs[pbase + (int64_t)r * S + lane] : 0.f;
            const float p1 = lane + 32 < S ? probs[pbase + (int64_t)r * S + lane + 32] : 0.f;
            const float d0 = lane < S ? sp[r * (TA_S + 1) + lane] : 0.f;
            const float d1 = lane + 32 < S ? sp[r * (TA_S + 1) + lane + 32] : 0.f;
            const float rs = warp_sum(d0 * p0 + d1 * p1);
            if (lane < S) sp[r * (TA_S + 1) + lane] = p0 * (d0 - rs) * scale;        // dS, with the 1 / sqrt(hd) of the scores
            if (lane + 32 < S) sp[r * (TA_S + 1) + lane + 32] = p1 * (d1 - rs) * scale;
        }
        __syncthreads();
        for (int i = tid; i < S * hd; i += 256) {
            const int r = i / hd, d = i - r * hd;
            float gq = 0.f, gk = 0.f, gv = 0.f;
            for (int c = 0; c < S; ++c) {
                gq = fmaf(sp[r * (TA_S + 1) + c], sk[c * HS + d], gq);               // dQ_r = sum_c dS_rc K_c
                gk = fmaf(sp[c * (TA_S + 1) + r], sq[c * HS + d], gk);               // dK_r = sum_c dS_cr Q_c
                gv = fmaf(sm[c * (TA_S + 1) + r], sd[c * HS + d], gv);               // dV_r = sum_c (P mask)_cr dOut_c
            }
            float* gp = dqkv + (b * S + r) * (int64_t)lddq + h * hd + d;
            gp[0] = gq;
            gp[D] = gk;
            gp[2 * D] = gv;
        }
    }
}

size_t ta_smem_bytes(int hd, bool bwd) {
    const size_t HS = hd + 1;
    return sizeof(float) * ((bwd ? 4 : 3) * TA_S * HS + (bwd ? 2 : 1) * TA_S * (TA_S + 1));
}

}  // namespace

extern "C" int dyg_gemm_dw(const float* G, int ldg, const float* X, int ldx, int64_t M, int N, int K, float* dW, int ldw, float* db,
                           dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldg >= N && ldx >= K && ldw >= K, "dyg_gemm_dw: bad sizes");
    if (M == 0) return 0;
    DYG_CHECK_ARG(G && X && dW, "dyg_gemm_dw: NULL pointer");
    const int gn = (N + DW_T - 1) / DW_T, gk = (K + DW_T - 1) / DW_T;
    int64_t splits = (4 * (int64_t)dyg_num_sms() + gn * gk - 1) / (gn * gk);
    const int64_t max_splits = (M + 4 * DW_MC - 1) / (4 * DW_MC);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int64_t rows = (M + splits - 1) / splits;
    rows = (rows + DW_MC - 1) / DW_MC * DW_MC;
    splits = (M + rows - 1) / rows;
    if (M >= 512) {
        // long contraction: BF16x3 on mma.sync (64 x 64 tiles, 32-row chunks)
        const int tn = (N + DM_T - 1) / DM_T, tk = (K + DM_T - 1) / DM_T;
        int64_t sp = (4 * (int64_t)dyg_num_sms() + tn * tk - 1) / (tn * tk);
        const int64_t max_sp = (M + 4 * DM_MC - 1) / (4 * DM_MC);
        if (sp > max_sp) sp = max_sp;
        if (sp < 1) sp = 1;
        if (sp > 65535) sp = 65535;
        int64_t r2 = (M + sp - 1) / sp;
        r2 = (r2 + DM_MC - 1) / DM_MC * DM_MC;
        sp = (M + r2 - 1) / r2;
        gemm_dw_mma_kernel<<<dim3((unsigned)tn, (unsigned)tk, (unsigned)sp), 128, 0, as_stream(stream)>>>(G, ldg, X, ldx, M, N, K, dW, ldw, db, r2);
        DYG_LAUNCH_CHECK("dyg_gemm_dw");
        return 0;
    }
    gemm_dw_kernel<<<dim3((unsigned)gn, (unsigned)gk, (unsigned)splits), 256, 0, as_stream(stream)>>>(G, ldg, X, ldx, M, N, K, dW, ldw, db, rows);
    DYG_LAUNCH_CHECK("dyg_gemm_dw");
    return 0;
}

extern "C" int dyg_gemm_dx(const float* G, int ldg, const float* Y, int ldy, const float* W, int ldw, int64_t M, int N, int K, float* dX,
                           int lddx, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldg >= N && ldw >= K && lddx >= K, "dyg_gemm_dx: bad sizes");
    if (M == 0) return 0;
    DYG_CHECK_ARG(G && W && dX, "dyg_gemm_dx: NULL pointer");
    const int64_t gm = (M + DM_T - 1) / DM_T;
    DYG_CHECK_ARG(gm < ((int64_t)1 << 31), "dyg_gemm_dx: M too large");
    gemm_dx_mma_kernel<<<dim3((unsigned)gm, (unsigned)((K + DM_T - 1) / DM_T)), 128, 0, as_stream(stream)>>>(G, ldg, Y, ldy, W, ldw, M, N, K, dX, lddx);
    DYG_LAUNCH_CHECK("dyg_gemm_dx");
    return 0;
}

extern "C" int dyg_linear_bwd(const float* G, int ldg, const float* Y, int ldy, const float* X, int ldx, const float* W, int ldw, int64_t M,
                              int N, int K, float* dX, int lddx, float* dW, int lddw, float* db, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldg >= N, "dyg_linear_bwd: bad sizes");
    if (M == 0) return 0;
    DYG_CHECK_ARG(G && (dX || dW), "dyg_linear_bwd: nothing to compute");
    DYG_CHECK_ARG(!dX || (W && ldw >= K && lddx >= K), "dyg_linear_bwd: dX needs W (N, K)");
    DYG_CHECK_ARG(!dW || (X && ldx >= K && lddw >= K), "dyg_linear_bwd: dW needs X (M, K)");
    DYG_CHECK_ARG(!db || dW, "dyg_linear_bwd: db is computed with dW");
    const int kt = (K + DW_T - 1) / DW_T;
    const int64_t dx_tiles64 = dX ? ((M + DW_T - 1) / DW_T) * kt : 0;
    DYG_CHECK_ARG(dx_tiles64 < ((int64_t)1 << 30), "dyg_linear_bwd: M too large");
    const int dw_tiles = dW ? ((N + DW_T - 1) / DW_T) * kt : 0;
    int64_t splits = 0, rows = 0;
    if (dW) {
        splits = (4 * (int64_t)dyg_num_sms() + dw_tiles - 1) / dw_tiles;
        const int64_t max_splits = (M + 4 * DW_MC - 1) / (4 * DW_MC);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        rows = (M + splits - 1) / splits;
        rows = (rows + DW_MC - 1) / DW_MC * DW_MC;
        splits = (M + rows - 1) / rows;
    }
    const int64_t grid = dx_tiles64 + dw_tiles * splits;
    DYG_CHECK_ARG(grid < ((int64_t)1 << 31), "dyg_linear_bwd: grid too large");
    linear_bwd_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(G, ldg, Y, ldy, X, ldx, W, ldw, M, N, K, dX, lddx, dW, lddw, db, (int)dx_tiles64,
                                                                    kt, kt, dw_tiles > 0 ? dw_tiles : 1, rows);
    DYG_LAUNCH_CHECK("dyg_linear_bwd");
    return 0;
}

extern "C" int dyg_layernorm_bwd(const float* x, int ldx, const float* gamma, float eps, const float* dy, int lddy, float* dx, int lddx,
                                 float* dgamma, float* dbeta, int64_t M, int D, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0 && D <= 512, "dyg_layernorm_bwd: D=%d unsupported (max 512)", D);
    if (M == 0) return 0;
    DYG_CHECK_ARG(x && gamma && dy && dx, "dyg_layernorm_bwd: NULL pointer");
    int64_t blocks = (M + 7) / 8;
    const int64_t cap = 4 * (int64_t)dyg_num_sms();
    if (blocks > cap) blocks = cap;
    const size_t smem = 2 * (size_t)D * sizeof(float);
    cudaStream_t s = as_stream(stream);
    if (D <= 128) layernorm_bwd_kernel<4><<<(unsigned)blocks, 256, smem, s>>>(x, ldx, gamma, eps, dy, lddy, dx, lddx, dgamma, dbeta, M, D);
    else if (D <= 256) layernorm_bwd_kernel<8><<<(unsigned)blocks, 256, smem, s>>>(x, ldx, gamma, eps, dy, lddy, dx, lddx, dgamma, dbeta, M, D);
    else layernorm_bwd_kernel<16><<<(unsigned)blocks, 256, smem, s>>>(x, ldx, gamma, eps, dy, lddy, dx, lddx, dgamma, dbeta, M, D);
    DYG_LAUNCH_CHECK("dyg_layernorm_bwd");
    return 0;
}

extern "C" int dyg_gelu_fwd(const float* v, int ldv, const float* mask, int ldm, float* h, int ldh, void* h_hi, void* h_mid, int lds,
                            int64_t M, int N, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0 && (N % 2) == 0 && (ldv % 2) == 0 && (!mask || (ldm % 2) == 0) && (!h || (ldh % 2) == 0) && (lds % 2) == 0,
                  "dyg_gelu_fwd: N and the leading dimensions must be even");
    DYG_CHECK_ARG((h_hi == nullptr) == (h_mid == nullptr) && (h || h_hi), "dyg_gelu_fwd: no output given");
    if (M == 0) return 0;
    DYG_CHECK_ARG(v && (reinterpret_cast<uintptr_t>(v) & 7u) == 0 && (!mask || (reinterpret_cast<uintptr_t>(mask) & 7u) == 0) &&
                      (!h || (reinterpret_cast<uintptr_t>(h) & 7u) == 0),
                  "dyg_gelu_fwd: fp32 pointers must be 8-byte aligned");
    const int64_t n = M * (N / 2);
    gelu_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(v, ldv, mask, ldm, h, ldh, reinterpret_cast<__nv_bfloat16*>(h_hi),
                                                                             reinterpret_cast<__nv_bfloat16*>(h_mid), lds, M, N);
    DYG_LAUNCH_CHECK("dyg_gelu_fwd");
    return 0;
}

extern "C" int dyg_gelu_bwd(const float* v, int ldv, const float* mask, int ldm, const float* dh, int lddh, float* dv, int lddv, int64_t M,
                            int N, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && N > 0, "dyg_gelu_bwd: bad sizes");
    if (M == 0) return 0;
    DYG_CHECK_ARG(v && dh && dv, "dyg_gelu_bwd: NULL pointer");
    const int64_t n = M * N;
    gelu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(v, ldv, mask, ldm, dh, lddh, dv, lddv, M, N);
    DYG_LAUNCH_CHECK("dyg_gelu_bwd");
    return 0;
}

static int ta_check(const char* name, int64_t B, int S, int H, int hd) {
    DYG_CHECK_ARG(B >= 0 && S > 0 && S <= TA_S && H > 0 && hd > 0 && hd <= 128, "%s: S=%d (max %d), head_dim=%d (max 128) unsupported", name, S,
                  TA_S, hd);
    DYG_CHECK_ARG(B * H < ((int64_t)1 << 31), "%s: too many (sequence, head) tiles", name);
    return 0;
}

extern "C" int dyg_seq_attention_train_fwd(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, const float* prob_mask, float* probs,
                                           float* out, int ldo, dyg_stream_t stream) {
    if (int rc = ta_check("dyg_seq_attention_train_fwd", B, S, H, hd)) return rc;
    if (B == 0) return 0;
    DYG_CHECK_ARG(qkv && probs && out, "dyg_seq_attention_train_fwd: NULL pointer");
    const size_t smem = ta_smem_bytes(hd, false);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(seq_attention_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_seq_attention_train_fwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    seq_attention_train_kernel<false><<<(unsigned)(B * H), 256, smem, as_stream(stream)>>>(qkv, ld_qkv, S, H, hd, prob_mask, probs, out, ldo, nullptr,
                                                                                         0, nullptr, 0);
    DYG_LAUNCH_CHECK("dyg_seq_attention_train_fwd");
    return 0;
}

extern "C" int dyg_seq_attention_train_bwd(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, const float* prob_mask,
                                           const float* probs, const float* dout, int lddo, float* dqkv, int lddq, dyg_stream_t stream) {
    if (int rc = ta_check("dyg_seq_attention_train_bwd", B, S, H, hd)) return rc;
    if (B == 0) return 0;
    DYG_CHECK_ARG(qkv && probs && dout && dqkv, "dyg_seq_attention_train_bwd: NULL pointer");
    const size_t smem = ta_smem_bytes(hd, true);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(seq_attention_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_seq_attention_train_bwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    seq_attention_train_kernel<true><<<(unsigned)(B * H), 256, smem, as_stream(stream)>>>(qkv, ld_qkv, S, H, hd, prob_mask, const_cast<float*>(probs),
                                                                                        nullptr, 0, dout, lddo, dqkv, lddq);
    DYG_LAUNCH_CHECK("dyg_seq_attention_train_bwd");
    return 0;
}
