// Co-occurrence counting, time encoding, row gathers, LayerNorm, token mean (parts of a9/a16).
// All HBM- or latency-bound element work; no tensor cores.
#include "common.cuh"

// ------------------------------------------------------------------ a13: co-occurrence counts
// One CTA per (src row, dst row) pair.  Both id rows are staged in shared memory; each thread owns one
// position and scans both rows (broadcast reads, conflict free): hash-free O(L^2) exact integer counting.
__global__ void cooc_count_kernel(const int64_t* __restrict__ src_ids, int ld_src, const int64_t* __restrict__ dst_ids,
                                  int ld_dst, int64_t B, int Ls, int Ld, float* __restrict__ out_src,
                                  float* __restrict__ out_dst, int64_t* __restrict__ cnt_src,
                                  int64_t* __restrict__ cnt_dst) {
    extern __shared__ long long cooc_smem[];
    long long* s = cooc_smem;
    long long* d = cooc_smem + Ls;
    const int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i < Ls; i += blockDim.x) s[i] = src_ids[b * ld_src + i];
    for (int i = threadIdx.x; i < Ld; i += blockDim.x) d[i] = dst_ids[b * ld_dst + i];
    __syncthreads();
    for (int i = threadIdx.x; i < Ls + Ld; i += blockDim.x) {
        const bool is_src = i < Ls;
        const long long id = is_src ? s[i] : d[i - Ls];
        int cs = 0, cd = 0;
        if (id != 0) {  // padded positions count as zero (models/DyGFormer.py:389-391)
            for (int j = 0; j < Ls; ++j) cs += (s[j] == id);
            for (int j = 0; j < Ld; ++j) cd += (d[j] == id);
        }
        if (is_src) {
            const int64_t o = (b * Ls + i) * 2;
            if (out_src) { out_src[o] = (float)cs; out_src[o + 1] = (float)cd; }
            if (cnt_src) { cnt_src[b * Ls + i] = cs; cnt_src[(B + b) * Ls + i] = cd; }
        } else {
            const int64_t o = (b * Ld + (i - Ls)) * 2;
            if (out_dst) { out_dst[o] = (float)cs; out_dst[o + 1] = (float)cd; }
            if (cnt_dst) { cnt_dst[b * Ld + (i - Ls)] = cs; cnt_dst[(B + b) * Ld + (i - Ls)] = cd; }
        }
    }
}

extern "C" int dyg_cooc_count(const int64_t* src_ids, int ld_src, const int64_t* dst_ids, int ld_dst, int64_t B, int Ls,
                              int Ld, float* out_src, float* out_dst, int64_t* cnt_src, int64_t* cnt_dst,
                              dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && Ls > 0 && Ld > 0 && ld_src >= Ls && ld_dst >= Ld, "dyg_cooc_count: bad sizes");
    DYG_CHECK_ARG((size_t)(Ls + Ld) * 8 <= 200 * 1024, "dyg_cooc_count: sequences too long for shared memory");
    if (B == 0) return 0;
    int threads = ((Ls + Ld + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = (size_t)(Ls + Ld) * 8;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(cooc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cooc_count_kernel<<<(unsigned)B, threads, smem, as_stream(stream)>>>(src_ids, ld_src, dst_ids, ld_dst, B, Ls, Ld,
                                                                        out_src, out_dst, cnt_src, cnt_dst);
    DYG_LAUNCH_CHECK("dyg_cooc_count");
    return 0;
}

// ------------------------------------------------------------------ a8: time encoder
__global__ void time_encode_kernel(const float* __restrict__ dt, int64_t n, const float* __restrict__ w,
                                   const float* __restrict__ b, int T, float* __restrict__ out) {
    const int64_t total = n * T;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / T;
        const int c = (int)(i - r * T);
        out[i] = dyg_time_enc(__ldg(dt + r), __ldg(w + c), __ldg(b + c));
    }
}
extern "C" int dyg_time_encode(const float* dt, int64_t n, const float* w, const float* b, int T, float* out,
                               dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && T > 0, "dyg_time_encode: bad sizes");
    if (n == 0) return 0;
    int64_t blocks = (n * T + 255) / 256;
    const int64_t cap = (int64_t)dyg_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    time_encode_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dt, n, w, b, T, out);
    DYG_LAUNCH_CHECK("dyg_time_encode");
    return 0;
}

// backward of the time encoder (training path; the reference differentiates models/modules.py:37 with autograd):
// out[i, c] = cos(a), a = fma(dt[i], w[c], b[c])  =>  gw[c] += sum_i g[i, c] (-sin a) dt[i],  gb[c] += sum_i g[i, c] (-sin a),
// with sin of the same fp32 argument the forward pass used.  A thread owns feature columns, a block a stripe of rows.
__global__ void time_encode_bwd_kernel(const float* __restrict__ dt, int64_t n, const float* __restrict__ w,
                                       const float* __restrict__ b, int T, const float* __restrict__ g, int64_t ldg,
                                       float* __restrict__ gw, float* __restrict__ gb) {
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t i0 = blockIdx.x * rows_per, i1 = i0 + rows_per < n ? i0 + rows_per : n;
    for (int c = threadIdx.x; c < T; c += blockDim.x) {
        const float wc = __ldg(w + c), bc = __ldg(b + c);
        float aw = 0.f, ab = 0.f;
        for (int64_t i = i0; i < i1; ++i) {
            const float d = __ldg(dt + i);
            float sn, cs;
            dyg_sincosf(fmaf(d, wc, bc), &sn, &cs);
            const float t = -__ldg(g + i * ldg + c) * sn;
            aw = fmaf(t, d, aw);
            ab += t;
        }
        atomicAdd(gw + c, aw);
        atomicAdd(gb + c, ab);
    }
}
extern "C" int dyg_time_encode_bwd(const float* dt, int64_t n, const float* w, const float* b, int T, const float* grad_out,
                                   int64_t ldg, float* grad_w, float* grad_b, dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && T > 0 && ldg >= T, "dyg_time_encode_bwd: bad sizes");
    if (n == 0) return 0;
    int64_t blocks = (n + 63) / 64;
    const int64_t cap = (int64_t)dyg_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    time_encode_bwd_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(dt, n, w, b, T, grad_out, ldg, grad_w, grad_b);
    DYG_LAUNCH_CHECK("dyg_time_encode_bwd");
    return 0;
}

// ------------------------------------------------------------------ row gather (+ add)
__global__ void gather_rows_kernel(const float* __restrict__ tab, int ld, const float* __restrict__ tab2, int ld2,
                                   const int64_t* __restrict__ idx, int64_t M, int D, float* __restrict__ out, int ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int64_t r = idx ? __ldg(idx + m) : m;
    const float* p = tab + r * ld;
    const float* p2 = tab2 ? tab2 + r * ld2 : nullptr;
    float* o = out + m * ldo;
    for (int c = lane; c < D; c += 32) o[c] = __ldg(p + c) + (p2 ? __ldg(p2 + c) : 0.f);
}
extern "C" int dyg_gather_rows(const float* tab, int ld, const float* tab2, int ld2, const int64_t* idx, int64_t M,
                               int D, float* out, int ldo, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0, "dyg_gather_rows: bad sizes");
    if (M == 0) return 0;
    gather_rows_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(tab, ld, tab2, ld2, idx, M, D, out, ldo);
    DYG_LAUNCH_CHECK("dyg_gather_rows");
    return 0;
}

// ------------------------------------------------------------------ LayerNorm(x + r)
// One warp per row, values kept in registers (D <= 32*MAXV), two-pass mean / biased variance like
// torch.nn.functional.layer_norm.
template <int MAXV>
__global__ void layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ r1, int ldr1, int F1,
                                 const float* __restrict__ rconst, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy, int64_t M,
                                 int D) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    float v[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        float t = 0.f;
        if (c < D) {
            t = x[m * ldx + c];
            if (c < F1) { if (r1) t += r1[m * ldr1 + c]; }
            else if (rconst) t += __ldg(rconst + (c - F1));
        }
        v[i] = t;
        sum += t;
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        const float d = (c < D) ? v[i] - mean : 0.f;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < D) y[m * ldy + c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
}
extern "C" int dyg_layernorm(const float* x, int ldx, const float* r1, int ldr1, int F1, const float* rconst,
                             const float* gamma, const float* beta, float eps, float* y, int ldy, int64_t M, int D,
                             dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0 && D <= 1024, "dyg_layernorm: D=%d unsupported (max 1024)", D);
    DYG_CHECK_ARG(F1 >= 0 && F1 <= D, "dyg_layernorm: bad residual split");
    if (M == 0) return 0;
    if (!r1 && !rconst) F1 = D;
    const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
    cudaStream_t s = as_stream(stream);
    if (D <= 256) layernorm_kernel<8><<<blocks, 256, 0, s>>>(x, ldx, r1, ldr1, F1, rconst, gamma, beta, eps, y, ldy, M, D);
    else if (D <= 512) layernorm_kernel<16><<<blocks, 256, 0, s>>>(x, ldx, r1, ldr1, F1, rconst, gamma, beta, eps, y, ldy, M, D);
    else layernorm_kernel<32><<<blocks, 256, 0, s>>>(x, ldx, r1, ldr1, F1, rconst, gamma, beta, eps, y, ldy, M, D);
    DYG_LAUNCH_CHECK("dyg_layernorm");
    return 0;
}

// ------------------------------------------------------------------ token mean (models/DyGFormer.py:185-187)
__global__ void mean_tokens_kernel(const float* __restrict__ x, int S, int D, int tok0, int cnt, float* __restrict__ out,
                                   int ldo) {
    const int64_t b = blockIdx.x;
    const float inv = 1.f / (float)cnt;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float acc = 0.f;
        const float* p = x + (b * S + tok0) * (int64_t)D + c;
        for (int t = 0; t < cnt; ++t) acc += p[(int64_t)t * D];
        out[b * ldo + c] = acc * inv;
    }
}
extern "C" int dyg_mean_tokens(const float* x, int64_t B, int S, int D, int tok0, int cnt, float* out, int ldo,
                               dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && cnt > 0 && tok0 >= 0 && tok0 + cnt <= S, "dyg_mean_tokens: bad token range");
    if (B == 0) return 0;
    mean_tokens_kernel<<<(unsigned)B, 128, 0, as_stream(stream)>>>(x, S, D, tok0, cnt, out, ldo);
    DYG_LAUNCH_CHECK("dyg_mean_tokens");
    return 0;
}

// ------------------------------------------------------------------ JODIE time projection (models/MemoryModel.py:114-118,543)
__global__ void jodie_project_kernel(const float* __restrict__ mem, int ld, const float* __restrict__ lu,
                                     const int64_t* __restrict__ ids, const double* __restrict__ t, int64_t M, int D,
                                     float mean, float stdv, const float* __restrict__ w, const float* __restrict__ b,
                                     float* __restrict__ out, int ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int64_t v = ids[m];
    const float iv = (((float)t[m] - lu[v]) - mean) / stdv;
    for (int c = lane; c < D; c += 32) out[m * ldo + c] = mem[v * ld + c] * (1.f + fmaf(iv, __ldg(w + c), __ldg(b + c)));
}
extern "C" int dyg_jodie_project(const float* mem, int ld, const float* lu, const int64_t* ids, const double* t,
                                 int64_t M, int D, float mean, float stdv, const float* w, const float* b, float* out,
                                 int ldo, dyg_stream_t stream) {
    DYG_CHECK_ARG(M >= 0 && D > 0, "dyg_jodie_project: bad sizes");
    if (M == 0) return 0;
    jodie_project_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(mem, ld, lu, ids, t, M, D, mean, stdv, w, b, out, ldo);
    DYG_LAUNCH_CHECK("dyg_jodie_project");
    return 0;
}
