// dyg_linear: C = act(A W^T + bias + residual) in fp32 where A rows are assembled on the fly from gathered
// table rows, patch groups and time encodings (no (rows x K) staging buffer in HBM).
// Round-1 implementation: 128x64x16 shared-memory tiles, FFMA, register double buffering.  The tcgen05
// (TF32x3) replacement for the large-M call sites is the next optimisation step.
#include <math.h>
#include <string.h>
#include "common.cuh"

struct SegPack {
    dyg_seg_t s[DYG_MAX_SEGS];
    int koff[DYG_MAX_SEGS + 1];
    int nseg;
};

constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ int find_seg(const SegPack& sp, int k) {
    int s = 0;
#pragma unroll
    for (int i = 1; i < DYG_MAX_SEGS; ++i)
        if (i < sp.nseg && k >= sp.koff[i]) s = i;
    return s;
}

// 4 consecutive columns k..k+3 of A row m (all inside one sub-row because every width is a multiple of 4)
__device__ __forceinline__ float4 load_a4(const SegPack& sp, int64_t m, int k) {
    const int si = find_seg(sp, k);
    const dyg_seg_t& sg = sp.s[si];
    int c = k - sp.koff[si];
    int p = 0;
    if (sg.group > 1) {
        p = c / sg.width;
        c -= p * sg.width;
    }
    const int64_t r = m * sg.group + p;
    float4 v;
    if (sg.kind == 0) {
        const int64_t ri = sg.idx ? __ldg(sg.idx + r) : r;
        v = __ldg(reinterpret_cast<const float4*>(sg.ptr + ri * sg.ld + c));
        if (sg.ptr2) {
            const int64_t r2 = sg.idx2 ? __ldg(sg.idx2 + r) : ri;
            const float4 u = __ldg(reinterpret_cast<const float4*>(sg.ptr2 + r2 * sg.ld2 + c));
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
    } else {
        if (sg.mask_ids && __ldg(sg.mask_ids + r) == 0) {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float dt = __ldg(sg.dt + r);
            if (sg.t_query) dt = (float)(__ldg(sg.t_query + r / sg.tq_div) - (double)dt);
            const float4 w = __ldg(reinterpret_cast<const float4*>(sg.w + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(sg.b + c));
            v.x = dyg_time_enc(dt, w.x, b.x);
            v.y = dyg_time_enc(dt, w.y, b.y);
            v.z = dyg_time_enc(dt, w.z, b.z);
            v.w = dyg_time_enc(dt, w.w, b.w);
        }
    }
    return v;
}

__device__ __forceinline__ float load_a1(const SegPack& sp, int64_t m, int k) {
    const int si = find_seg(sp, k);
    const dyg_seg_t& sg = sp.s[si];
    int c = k - sp.koff[si];
    int p = 0;
    if (sg.group > 1) {
        p = c / sg.width;
        c -= p * sg.width;
    }
    const int64_t r = m * sg.group + p;
    if (sg.kind == 0) {
        const int64_t ri = sg.idx ? __ldg(sg.idx + r) : r;
        float v = __ldg(sg.ptr + ri * sg.ld + c);
        if (sg.ptr2) {
            const int64_t r2 = sg.idx2 ? __ldg(sg.idx2 + r) : ri;
            v += __ldg(sg.ptr2 + r2 * sg.ld2 + c);
        }
        return v;
    }
    if (sg.mask_ids && __ldg(sg.mask_ids + r) == 0) return 0.f;
    float dt = __ldg(sg.dt + r);
    if (sg.t_query) dt = (float)(__ldg(sg.t_query + r / sg.tq_div) - (double)dt);
    return dyg_time_enc(dt, __ldg(sg.w + c), __ldg(sg.b + c));
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == DYG_ACT_RELU) return fmaxf(v, 0.f);
    if (act == DYG_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    if (act == DYG_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

template <bool VEC>
__global__ void __launch_bounds__(256) linear_kernel(const SegPack sp, const float* __restrict__ W, int ldw,
                                                     const float* __restrict__ bias, const float* __restrict__ residual,
                                                     int ldr, float* __restrict__ C, int ldc, int64_t M, int N, int K,
                                                     int act, int c_group, int c_group_stride, int c_offset) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];
    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int lr = t >> 2;         // 0..63: tile row handled by the loaders
    const int lk = (t & 3) * 4;    // 0,4,8,12
    const int tx = t & 15, ty = t >> 4;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra0, ra1, rw;
    auto fetch = [&](int k0) {
        const int k = k0 + lk;
        const int64_t ma = m0 + lr, mb = m0 + lr + 64;
        const int n = n0 + lr;
        if (VEC) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            ra0 = (ma < M && k < K) ? load_a4(sp, ma, k) : z;
            ra1 = (mb < M && k < K) ? load_a4(sp, mb, k) : z;
            rw = (n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + k)) : z;
        } else {
            float a0[4], a1[4], w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a0[i] = (ma < M && k + i < K) ? load_a1(sp, ma, k + i) : 0.f;
                a1[i] = (mb < M && k + i < K) ? load_a1(sp, mb, k + i) : 0.f;
                w4[i] = (n < N && k + i < K) ? __ldg(W + (int64_t)n * ldw + k + i) : 0.f;
            }
            ra0 = make_float4(a0[0], a0[1], a0[2], a0[3]);
            ra1 = make_float4(a1[0], a1[1], a1[2], a1[3]);
            rw = make_float4(w4[0], w4[1], w4[2], w4[3]);
        }
    };
    auto stash = [&](int buf) {
        As[buf][lk + 0][lr] = ra0.x; As[buf][lk + 1][lr] = ra0.y; As[buf][lk + 2][lr] = ra0.z; As[buf][lk + 3][lr] = ra0.w;
        As[buf][lk + 0][lr + 64] = ra1.x; As[buf][lk + 1][lr + 64] = ra1.y; As[buf][lk + 2][lr + 64] = ra1.z; As[buf][lk + 3][lr + 64] = ra1.w;
        Ws[buf][lk + 0][lr] = rw.x; Ws[buf][lk + 1][lr] = rw.y; Ws[buf][lk + 2][lr] = rw.z; Ws[buf][lk + 3][lr] = rw.w;
    };

    const int nk = (K + BK - 1) / BK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) stash(buf ^ 1);
        __syncthreads();
    }

    const int col0 = n0 + tx * 4;
    if (col0 >= N) return;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (col0 + j < N) bv[j] = __ldg(bias + col0 + j);
    }
    const bool vec_out = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) && (col0 + 3 < N);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const int64_t crow = c_group > 0 ? (m / c_group) * c_group_stride + (m % c_group) + c_offset : m;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j] + bv[j];
            if (residual && col0 + j < N) v += __ldg(residual + crow * ldr + col0 + j);
            o[j] = apply_act(v, act);
        }
        float* dst = C + crow * ldc + col0;
        if (vec_out) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (col0 + j < N) dst[j] = o[j];
        }
    }
}

extern "C" int dyg_linear(const dyg_seg_t* segs, int nseg, const float* W, int ldw, const float* bias,
                          const float* residual, int ldr, float* C, int ldc, int64_t M, int N, int act, int c_group,
                          int c_group_stride, int c_offset, dyg_stream_t stream) {
    DYG_CHECK_ARG(nseg >= 1 && nseg <= DYG_MAX_SEGS, "dyg_linear: nseg=%d out of range", nseg);
    DYG_CHECK_ARG(M >= 0 && N > 0, "dyg_linear: bad sizes");
    DYG_CHECK_ARG(act >= DYG_ACT_NONE && act <= DYG_ACT_SIGMOID, "dyg_linear: unknown activation %d", act);
    if (M == 0) return 0;
    SegPack sp;
    memset(&sp, 0, sizeof(sp));
    sp.nseg = nseg;
    bool vec = ((ldw & 3) == 0) && aligned16(W);
    int K = 0;
    for (int i = 0; i < nseg; ++i) {
        const dyg_seg_t& s = segs[i];
        DYG_CHECK_ARG(s.width > 0 && s.group > 0, "dyg_linear: segment %d has empty shape", i);
        DYG_CHECK_ARG(s.kind == 0 || s.kind == 1, "dyg_linear: segment %d has unknown kind", i);
        if (s.kind == 0) {
            DYG_CHECK_ARG(s.ptr != nullptr, "dyg_linear: segment %d has no table", i);
            vec = vec && ((s.ld & 3) == 0) && aligned16(s.ptr);
            if (s.ptr2) vec = vec && ((s.ld2 & 3) == 0) && aligned16(s.ptr2);
        } else {
            DYG_CHECK_ARG(s.dt && s.w && s.b, "dyg_linear: time segment %d needs dt, w, b", i);
            DYG_CHECK_ARG(!s.t_query || s.tq_div > 0, "dyg_linear: time segment %d needs tq_div > 0", i);
            vec = vec && aligned16(s.w) && aligned16(s.b);
        }
        vec = vec && ((s.width & 3) == 0);
        sp.s[i] = s;
        sp.koff[i] = K;
        K += s.width * s.group;
    }
    for (int i = nseg; i <= DYG_MAX_SEGS; ++i) sp.koff[i] = K;
    DYG_CHECK_ARG(ldw >= K, "dyg_linear: ldw %d < K %d", ldw, K);
    const int64_t mt = (M + BM - 1) / BM;
    DYG_CHECK_ARG(mt < (1ll << 31), "dyg_linear: M too large");
    dim3 grid((unsigned)mt, (unsigned)((N + BN - 1) / BN));
    if (vec)
        linear_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(sp, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act,
                                                                 c_group, c_group_stride, c_offset);
    else
        linear_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(sp, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act,
                                                                  c_group, c_group_stride, c_offset);
    DYG_LAUNCH_CHECK("dyg_linear");
    return 0;
}
