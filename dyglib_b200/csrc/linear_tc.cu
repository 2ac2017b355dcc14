// dyg_linear_tc: the tensor-core version of dyg_linear for large M.
//
//   C[row(m), :N] = act(A W^T + bias + residual)        A assembled on the fly (gathers / patches / time encodings)
//
// Precision: "BF16x3".  Every fp32 operand x is split as x = hi + mid (+ lo dropped), hi = bf16(x),
// mid = bf16(x - hi); the product uses three tcgen05 MMAs  A_hi W_hi + A_hi W_mid + A_mid W_hi  with fp32
// accumulation in TMEM.  Dropped terms are O(2^-16) relative per product (~1e-5 observed on the models),
// far inside the 1e-3 parity budget, at 3 bf16 MMAs per tile (vs 6 for full fp32 emulation).
//
// Structure (one CTA per 128 x NT output tile, NT <= 256):
//   warps 0-3  producers: gather fp32 A rows (coalesced 32-byte chunks, 8 lanes per row), split to bf16 hi/mid
//              and write the K-major SWIZZLE_128B tile; weights (pre-split bf16) arrive by cp.async;
//              after the main loop the same warps run the epilogue (tcgen05.ld 32x32b -> bias/residual/act -> global)
//   warp 4     TMEM alloc + single-thread tcgen05.mma issue; tcgen05.commit releases smem stages / signals the epilogue
//   2-stage mbarrier ring (full: 128 producer arrivals after fence.proxy.async; empty: one tcgen05.commit arrival).
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>
#include "common.cuh"

struct SegPackTC {
    dyg_seg_t s[DYG_MAX_SEGS];
    int koff[DYG_MAX_SEGS + 1];
    int nseg;
};

namespace {

constexpr int TC_BM = 128;       // rows per CTA tile == TMEM lanes
constexpr int TC_BK = 64;        // bf16 elements per stage along K == one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_PRODUCERS = 128;
constexpr int TC_THREADS = 160;
constexpr int A_TILE_BYTES = TC_BM * 128;   // 16 KB per hi / mid tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major, canonical 1) in [16,30),
// SBO = 1024 B between 8-row groups in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13),
// K-major A and B (0) [15],[16], N>>3 in [17,23), M>>4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ int find_seg_tc(const SegPackTC& sp, int k) {
    int s = 0;
#pragma unroll
    for (int i = 1; i < DYG_MAX_SEGS; ++i)
        if (i < sp.nseg && k >= sp.koff[i]) s = i;
    return s;
}
__device__ __forceinline__ float4 load_a4_tc(const SegPackTC& sp, int64_t m, int k) {
    const int si = find_seg_tc(sp, k);
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

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& mid) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat16 am = __float2bfloat16_rn(a - __bfloat162float(ah)), bm = __float2bfloat16_rn(b - __bfloat162float(bh));
    hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
    mid = (uint32_t)__bfloat16_as_ushort(am) | ((uint32_t)__bfloat16_as_ushort(bm) << 16);
}

__device__ __forceinline__ float act_tc(float v, int act) {
    if (act == DYG_ACT_RELU) return fmaxf(v, 0.f);
    if (act == DYG_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    if (act == DYG_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

__global__ void __launch_bounds__(TC_THREADS, 1) linear_tc_kernel(
    const SegPackTC sp, const __nv_bfloat16* __restrict__ Wh, const __nv_bfloat16* __restrict__ Wm, int ldwp,
    const float* __restrict__ bias, const float* __restrict__ residual, int ldr, float* __restrict__ C, int ldc, int64_t M,
    int N, int K, int NT, int tmem_cols, int act, int c_group, int c_group_stride, int c_offset) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    // carve: stages of [A_hi | A_mid | B_hi | B_mid], all 1024-byte aligned (NT multiple of 16 -> NT*128 multiple of 1024... NT*128 = 2048*(NT/16))
    const int b_tile_bytes = NT * 128;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * stage_bytes);
    uint64_t* full_bar = bars;                 // [TC_STAGES]
    uint64_t* empty_bar = bars + TC_STAGES;    // [TC_STAGES]
    uint64_t* accum_bar = bars + 2 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * NT;
    const int nkb = (K + TC_BK - 1) / TC_BK;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full_bar + s, TC_PRODUCERS);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int c = tid & 7;          // 16-byte bf16 chunk (8 elements) inside the 128-byte row
        const int rsub = tid >> 3;      // 0..15
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % TC_STAGES;
            const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
            mbar_wait(empty_bar + s, ph ^ 1u);
            unsigned char* st = base + s * stage_bytes;
            // weights: pre-split bf16, zero padded to (Npad, Kpad) -> straight 16-byte async copies
            {
                const uint32_t bh = smem_u32(st + 2 * A_TILE_BYTES), bm = bh + b_tile_bytes;
                for (int j = tid; j < NT * 8; j += TC_PRODUCERS) {
                    const int n = j >> 3, cc = j & 7;
                    const size_t goff = (size_t)(n0 + n) * ldwp + (size_t)kb * TC_BK + cc * 8;
                    const uint32_t soff = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((cc ^ (n & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bh + soff), "l"(Wh + goff) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bm + soff), "l"(Wm + goff) : "memory");
                }
            }
            // activations: gather + split
            const int k = kb * TC_BK + c * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 16 + rsub;
                const int64_t m = m0 + r;
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (m < M) {
                    if (k < K) v0 = load_a4_tc(sp, m, k);
                    if (k + 4 < K) v1 = load_a4_tc(sp, m, k + 4);
                }
                uint4 hi, mid;
                split2(v0.x, v0.y, hi.x, mid.x);
                split2(v0.z, v0.w, hi.y, mid.y);
                split2(v1.x, v1.y, hi.z, mid.z);
                split2(v1.z, v1.w, hi.w, mid.w);
                const uint32_t soff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(st + soff) = hi;
                *reinterpret_cast<uint4*>(st + A_TILE_BYTES + soff) = mid;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05 (async proxy)
            mbar_arrive(full_bar + s);
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t m = m0 + tid;      // TMEM lane == tile row; warp w may only touch lanes 32w..32w+31
        const int64_t crow = (m < M) ? (c_group > 0 ? (m / c_group) * c_group_stride + (m % c_group) + c_offset : m) : 0;
        const bool vec_out = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
        for (int col = 0; col < NT; col += 16) {
            if (n0 + col >= N) break;                                   // warp-uniform
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, r);
            if (m < M) {
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + col + j;
                    float v = __uint_as_float(r[j]);
                    if (n < N) {
                        if (bias) v += __ldg(bias + n);
                        if (residual) v += __ldg(residual + crow * ldr + n);
                        v = act_tc(v, act);
                    }
                    o[j] = v;
                }
                float* dst = C + crow * ldc + n0 + col;
                if (vec_out && n0 + col + 15 < N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + col + j < N) dst[j] = o[j];
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (warp 4)
        const uint32_t idesc = make_idesc(NT);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % TC_STAGES;
            const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
            mbar_wait(full_bar + s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_h = smem_u32(base + s * stage_bytes), a_m = a_h + A_TILE_BYTES;
                const uint32_t b_h = a_h + 2 * A_TILE_BYTES, b_m = b_h + b_tile_bytes;
                const int krem = K - kb * TC_BK;
                const int nk16 = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
                for (int kk = 0; kk < nk16; ++kk) {
                    const uint32_t o = (uint32_t)kk * 32u;                 // 16 bf16 = 32 bytes inside the swizzle row
                    const uint64_t dah = make_desc(a_h + o), dam = make_desc(a_m + o);
                    const uint64_t dbh = make_desc(b_h + o), dbm = make_desc(b_m + o);
                    umma_bf16(tmem_base, dah, dbh, idesc, (kb | kk) != 0);
                    umma_bf16(tmem_base, dah, dbm, idesc, 1);
                    umma_bf16(tmem_base, dam, dbh, idesc, 1);
                }
                umma_commit(empty_bar + s);                                 // frees the stage when the MMAs retire
                if (kb == nkb - 1) umma_commit(accum_bar);                  // accumulator complete -> epilogue
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

}  // namespace

extern "C" int dyg_linear_tc_tile(int N) {
    // N tile: whole N when it fits one MMA (<= 256), else the even split rounded up to 16
    if (N <= 256) return (N + 15) / 16 * 16;
    const int parts = (N + 255) / 256;
    return ((N + parts - 1) / parts + 15) / 16 * 16;
}

extern "C" int dyg_linear_tc(const dyg_seg_t* segs, int nseg, const void* W_hi, const void* W_mid, int ldwp, int n_pad,
                             const float* bias, const float* residual, int ldr, float* C, int ldc, int64_t M, int N,
                             int act, int c_group, int c_group_stride, int c_offset, dyg_stream_t stream) {
    DYG_CHECK_ARG(nseg >= 1 && nseg <= DYG_MAX_SEGS, "dyg_linear_tc: nseg=%d out of range", nseg);
    DYG_CHECK_ARG(M >= 0 && N > 0, "dyg_linear_tc: bad sizes");
    DYG_CHECK_ARG(act >= DYG_ACT_NONE && act <= DYG_ACT_SIGMOID, "dyg_linear_tc: unknown activation %d", act);
    if (M == 0) return 0;
    SegPackTC sp;
    memset(&sp, 0, sizeof(sp));
    sp.nseg = nseg;
    int K = 0;
    for (int i = 0; i < nseg; ++i) {
        const dyg_seg_t& s = segs[i];
        DYG_CHECK_ARG(s.width > 0 && s.group > 0 && (s.width & 3) == 0, "dyg_linear_tc: segment %d width must be a positive multiple of 4", i);
        if (s.kind == 0) {
            DYG_CHECK_ARG(s.ptr && (s.ld & 3) == 0 && aligned16(s.ptr), "dyg_linear_tc: segment %d table must be 16-byte aligned, ld %% 4 == 0", i);
            DYG_CHECK_ARG(!s.ptr2 || ((s.ld2 & 3) == 0 && aligned16(s.ptr2)), "dyg_linear_tc: segment %d second table misaligned", i);
        } else {
            DYG_CHECK_ARG(s.kind == 1 && s.dt && s.w && s.b && aligned16(s.w) && aligned16(s.b), "dyg_linear_tc: bad time segment %d", i);
            DYG_CHECK_ARG(!s.t_query || s.tq_div > 0, "dyg_linear_tc: time segment %d needs tq_div > 0", i);
        }
        sp.s[i] = s;
        sp.koff[i] = K;
        K += s.width * s.group;
    }
    for (int i = nseg; i <= DYG_MAX_SEGS; ++i) sp.koff[i] = K;
    const int NT = dyg_linear_tc_tile(N);
    const int ntiles = (N + NT - 1) / NT;
    DYG_CHECK_ARG((ldwp % TC_BK) == 0 && ldwp >= K, "dyg_linear_tc: padded weight K stride %d must be a multiple of 64 and >= K=%d", ldwp, K);
    DYG_CHECK_ARG(n_pad >= ntiles * NT, "dyg_linear_tc: padded weight rows %d < %d", n_pad, ntiles * NT);
    DYG_CHECK_ARG(aligned16(W_hi) && aligned16(W_mid), "dyg_linear_tc: weights must be 16-byte aligned");
    int tmem_cols = 32;
    while (tmem_cols < NT) tmem_cols <<= 1;
    const size_t smem = (size_t)TC_STAGES * (2 * A_TILE_BYTES + 2 * NT * 128) + 1024 + 64;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            dyg_set_error("dyg_linear_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
            return 1;
        }
        configured = smem;
    }
    const int64_t mt = (M + TC_BM - 1) / TC_BM;
    DYG_CHECK_ARG(mt < (1ll << 31), "dyg_linear_tc: M too large");
    dim3 grid((unsigned)mt, (unsigned)ntiles);
    linear_tc_kernel<<<grid, TC_THREADS, smem, as_stream(stream)>>>(
        sp, reinterpret_cast<const __nv_bfloat16*>(W_hi), reinterpret_cast<const __nv_bfloat16*>(W_mid), ldwp, bias, residual,
        ldr, C, ldc, M, N, K, NT, tmem_cols, act, c_group, c_group_stride, c_offset);
    DYG_LAUNCH_CHECK("dyg_linear_tc");
    return 0;
}
