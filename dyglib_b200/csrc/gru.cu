// Fused recurrent memory update (north-star kernel K9): nn.GRUCell / nn.RNNCell of
// models/MemoryModel.py:490-515 as ONE launch -- gathered message rows x W_ih, gathered memory rows x W_hh, the gate
// non-linearities and the scatter of the new memories -- instead of two gate GEMMs plus an element-wise kernel.
// The contraction is the fp32 FFMA tile of tile_gemm.cuh (the step is latency-bound at batch 200: ~0.3 GFLOP).
#include <math.h>
#include "tile_gemm.cuh"

namespace {

// A CTA owns TM*16 rows x 16 hidden units and, for every (row, unit), the G gate pre-activations.
template <int TM, int G>
__global__ void __launch_bounds__(tg::THREADS) gru_update_kernel(
    const float* __restrict__ msg, int64_t ldm, const int64_t* __restrict__ msg_idx, int msg_dim,
    const float* __restrict__ hid, int64_t ldh, const int64_t* __restrict__ hid_idx, int D,
    const float* __restrict__ w_ih, const float* __restrict__ b_ih, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
    const int32_t* __restrict__ winner, float* __restrict__ out, int64_t ldo, const int64_t* __restrict__ out_idx,
    float* __restrict__ gates, int64_t P) {
    extern __shared__ __align__(16) float smem[];   // tg::Tile<TM, G>::SMEM_FLOATS
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * (16 * TM);
    const int u0 = blockIdx.y * 16;
    float acc[TM][G];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int g = 0; g < G; ++g) acc[i][g] = 0.f;
    const tg::WGates wmap{u0, D, 16, G};
    tg::gemm_accum<TM, G>(acc, tg::ASeg{msg, msg_idx, ldm, msg_dim}, m0, P, w_ih, msg_dim, wmap, smem, t, 1);
    float in_n[TM];   // GRU: the candidate gate keeps its input and hidden halves apart (n = tanh(i_n + r * h_n))
    if (G == 3) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            in_n[i] = acc[i][G - 1];
            acc[i][G - 1] = 0.f;
        }
    }
    tg::gemm_accum<TM, G>(acc, tg::ASeg{hid, hid_idx, ldh, D}, m0, P, w_hh, D, wmap, smem, t, 1);
    const int u = u0 + tx;
    if (u >= D) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty + 16 * i;
        if (m >= P) continue;
        const int64_t hv = hid_idx ? hid_idx[m] : m;
        if (winner && winner[hv] != (int32_t)m) continue;
        float hn;
        if (G == 3) {   // nn.GRUCell gate order r, z, n
            const float r = tg::sigmoidf_(acc[i][0] + b_ih[u] + b_hh[u]);
            const float z = tg::sigmoidf_(acc[i][1] + b_ih[D + u] + b_hh[D + u]);
            const float hpre = acc[i][G - 1] + b_hh[2 * D + u];
            const float ng = tanhf(in_n[i] + b_ih[2 * D + u] + r * hpre);
            const float h = hid[hv * ldh + u];
            hn = ng + z * (h - ng);
            if (gates) {
                float* gp = gates + m * (4 * (int64_t)D);
                gp[u] = r;
                gp[D + u] = z;
                gp[2 * D + u] = ng;
                gp[3 * D + u] = hpre;
            }
        } else {        // nn.RNNCell (tanh)
            hn = tanhf(acc[i][0] + b_ih[u] + b_hh[u]);
            if (gates) gates[m * (int64_t)D + u] = hn;
        }
        out[(out_idx ? out_idx[m] : m) * ldo + u] = hn;
    }
}

// element-wise half of the cell's backward: gate gradients from the saved gates (the four weight / input products run on
// the tcgen05 GEMM, see dyglib_b200/autograd.py)
template <int G>
__global__ void gru_update_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ hid, int64_t ldh,
                                      const int64_t* __restrict__ hid_idx, const float* __restrict__ grad_out, int64_t ldg,
                                      float* __restrict__ d_gi, float* __restrict__ d_gh, float* __restrict__ d_h, int64_t P, int D) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= P * D) return;
    const int64_t m = e / D;
    const int u = (int)(e - m * D);
    const float g = grad_out[m * ldg + u];
    if (G == 3) {
        const float* gp = gates + m * (4 * (int64_t)D);
        const float r = gp[u], z = gp[D + u], n = gp[2 * D + u], hpre = gp[3 * D + u];
        const float h = hid[(hid_idx ? hid_idx[m] : m) * ldh + u];
        const float dn = g * (1.f - z) * (1.f - n * n);     // d pre-activation of the candidate gate
        const float dz = g * (h - n) * z * (1.f - z);
        const float dr = dn * hpre * r * (1.f - r);
        float* gi = d_gi + m * (3 * (int64_t)D);
        float* gh = d_gh + m * (3 * (int64_t)D);
        gi[u] = dr;
        gi[D + u] = dz;
        gi[2 * D + u] = dn;
        gh[u] = dr;
        gh[D + u] = dz;
        gh[2 * D + u] = dn * r;
        d_h[m * (int64_t)D + u] = g * z;
    } else {
        const float n = gates[m * (int64_t)D + u];
        const float dp = g * (1.f - n * n);
        d_gi[m * (int64_t)D + u] = dp;
        d_gh[m * (int64_t)D + u] = dp;
        d_h[m * (int64_t)D + u] = 0.f;
    }
}

}  // namespace

extern "C" int dyg_gru_update_fwd(const float* msg, int ldm, const int64_t* msg_idx, int msg_dim, const float* hid, int ldh,
                                  const int64_t* hid_idx, int D, const float* w_ih, const float* b_ih, const float* w_hh,
                                  const float* b_hh, int G, const int32_t* winner, float* out, int ldo, const int64_t* out_idx,
                                  float* gates, int64_t P, dyg_stream_t stream) {
    DYG_CHECK_ARG(G == 1 || G == 3, "dyg_gru_update_fwd: G must be 3 (GRU) or 1 (RNN)");
    DYG_CHECK_ARG(P >= 0 && D > 0 && msg_dim > 0, "dyg_gru_update_fwd: bad sizes");
    DYG_CHECK_ARG(msg_dim % 4 == 0 && D % 4 == 0 && ldm % 4 == 0 && ldh % 4 == 0, "dyg_gru_update_fwd: widths and leading dimensions must be multiples of 4");
    DYG_CHECK_ARG(aligned16(msg) && aligned16(hid) && aligned16(w_ih) && aligned16(w_hh), "dyg_gru_update_fwd: operands must be 16-byte aligned");
    DYG_CHECK_ARG(msg && hid && w_ih && b_ih && w_hh && b_hh && out, "dyg_gru_update_fwd: null operand");
    if (P == 0) return 0;
    constexpr int TM = 2;
    dim3 grid((unsigned)((P + 16 * TM - 1) / (16 * TM)), (unsigned)((D + 15) / 16));
    constexpr int smem3 = tg::Tile<TM, 3>::SMEM_FLOATS * 4, smem1 = tg::Tile<TM, 1>::SMEM_FLOATS * 4;
    {   // per device and cheap: set on every call (one process may drive several devices)
        cudaFuncSetAttribute(gru_update_kernel<TM, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
        cudaFuncSetAttribute(gru_update_kernel<TM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    }
    if (G == 3)
        gru_update_kernel<TM, 3><<<grid, tg::THREADS, smem3, as_stream(stream)>>>(msg, ldm, msg_idx, msg_dim, hid, ldh, hid_idx, D, w_ih, b_ih,
                                                                                   w_hh, b_hh, winner, out, ldo, out_idx, gates, P);
    else
        gru_update_kernel<TM, 1><<<grid, tg::THREADS, smem1, as_stream(stream)>>>(msg, ldm, msg_idx, msg_dim, hid, ldh, hid_idx, D, w_ih, b_ih,
                                                                                   w_hh, b_hh, winner, out, ldo, out_idx, gates, P);
    DYG_LAUNCH_CHECK("dyg_gru_update_fwd");
    return 0;
}

extern "C" int dyg_gru_update_bwd(const float* gates, const float* hid, int ldh, const int64_t* hid_idx, const float* grad_out,
                                  int ldg, int G, float* d_gi, float* d_gh, float* d_h, int64_t P, int D, dyg_stream_t stream) {
    DYG_CHECK_ARG(G == 1 || G == 3, "dyg_gru_update_bwd: G must be 3 (GRU) or 1 (RNN)");
    DYG_CHECK_ARG(P >= 0 && D > 0, "dyg_gru_update_bwd: bad sizes");
    DYG_CHECK_ARG(gates && hid && grad_out && d_gi && d_gh && d_h, "dyg_gru_update_bwd: null operand");
    if (P == 0) return 0;
    const unsigned blocks = (unsigned)((P * D + 255) / 256);
    if (G == 3)
        gru_update_bwd_kernel<3><<<blocks, 256, 0, as_stream(stream)>>>(gates, hid, ldh, hid_idx, grad_out, ldg, d_gi, d_gh, d_h, P, D);
    else
        gru_update_bwd_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(gates, hid, ldh, hid_idx, grad_out, ldg, d_gi, d_gh, d_h, P, D);
    DYG_LAUNCH_CHECK("dyg_gru_update_bwd");
    return 0;
}
