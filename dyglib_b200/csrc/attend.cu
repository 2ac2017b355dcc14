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
                if (c < F4) {
                    v = ldg_stream(reinterpret_cast<const float4*>(node_tab + rn * ld_node) + c);
                    if (node_tab2) {
                        const float4 u = ldg_stream(reinterpret_cast<const float4*>(node_tab2 + rn * ld_node2) + c);
                        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                    }
                } else if (c < F4 + E4) {
                    v = ldg_stream(reinterpret_cast<const float4*>(edge_tab + re * ld_edge) + (c - F4));
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

extern "C" int dyg_temporal_attend(const float* qk, int ldq, int64_t n, int k, int H, const float* node_tab, int ld_node,
                                   const float* node_tab2, int ld_node2, const int64_t* node_idx, int F,
                                   const float* edge_tab, int ld_edge, const int64_t* edge_idx, int E,
                                   const float* time_feat, const double* t_query, const float* t_nbr, const float* w,
                                   const float* b, int T, const int64_t* mask_ids, float* out_s, int lds,
                                   float* out_scores, dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && k > 0, "dyg_temporal_attend: bad sizes");
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
    if (H == 2) temporal_attend_kernel<2, 4><<<blocks, 128, 0, s>>>(ATTEND_ARGS);
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
