// BF16x3 tile contraction on the warp-level tensor-core path (mma.sync.m16n8k16) for the latency-bound small-batch kernels
// (the one-launch TGN step).
//
// Why not tcgen05 here: a 200-event step is ~1 GFLOP spread over a chain of dependent phases; what matters is the latency of a
// phase on 148 SMs, i.e. many small tiles (32 x 64), not the throughput of a 128-row UMMA tile.  Why not FFMA: the fp32 tile of
// tile_gemm.cuh is bound by shared-memory wavefronts (every LDS.128 costs four, profiles/r02_tile_gemm_probe.txt: "compute
// only" = 85 % of the full time).  Here operands travel as BF16x3 planes (x = hi + mid, csrc/gemm_tc.cu) that the PRODUCING phase
// writes in its epilogue, so a stage is plain 16-byte cp.async copies, fragments come from ldmatrix.x4, and a product is three
// MMAs (hi*hi + hi*mid + mid*hi, fp32 accumulate, ~1e-5 relative like every other dense contraction of the package):
// per 16 x 16 x 16 block a warp issues 4 ldmatrix + 6 mma instead of ~150 LDS / FFMA.
//
// Team = 256 threads = 8 warps arranged WM x WN; warp tile (16 MI) x (8 NI); stage = BK = 64 columns of (BM + BN) rows x 2 planes
// with a 72-element pitch (144-byte rows: the eight 16-byte rows of an ldmatrix phase fall on distinct bank groups).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace mt {

constexpr int BK = 64;
constexpr int PITCH = BK + 8;
#ifndef MT_STAGES
#define MT_STAGES 3
#endif
constexpr int STAGES = MT_STAGES;
constexpr int THREADS = 256;

typedef __nv_bfloat16 bf16;

// hi | mid planes of a (rows, ld) matrix; ld % 8 == 0, both planes 16-byte aligned, padding columns zero
struct Planes {
    const bf16* hi;
    const bf16* mid;
    int64_t ld;
};
// one K segment of A: columns [col0, col0 + width) of rows idx[m] (idx NULL: row m); col0 % 8 == 0, width % 8 == 0
struct ASeg {
    Planes p;
    const int64_t* idx;
    int col0;
    int width;
};

template <int BM, int BN>
struct Tile {
    static constexpr int STAGE_ELEMS = 2 * (BM + BN) * PITCH;
    static constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 2;
};

__device__ __forceinline__ void team_sync(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(THREADS) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (a, b) -> packed bf16 pairs hi = bf16(x), mid = bf16(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& mid) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 m = __floats2bfloat162_rn(a - ah, b - bh);
    mid = *reinterpret_cast<const uint32_t*>(&m);
}
// write (a, b) at columns c, c + 1 of row `row` of a plane pair (c even)
__device__ __forceinline__ void store_split2(bf16* hi, bf16* mid, int64_t ld, int64_t row, int c, float a, float b) {
    uint32_t h, m;
    split2(a, b, h, m);
    *reinterpret_cast<uint32_t*>(hi + row * ld + c) = h;
    *reinterpret_cast<uint32_t*>(mid + row * ld + c) = m;
}

// plain weight rows n0 .. n0 + BN - 1
struct WRows {
    int n0, N;
    __device__ __forceinline__ int operator()(int r) const { return n0 + r < N ? n0 + r : -1; }
};
// recurrent-cell weight rows (G*D, K), gate-major: the 8 NI columns of warp column wn are NI = G gate blocks of the same 8 hidden
// units, so a thread's accumulators hold every gate of its (row, unit): tile row r -> wn = r / (8 G), gate (r % (8 G)) / 8
struct WGates {
    int u0, D, G;
    __device__ __forceinline__ int operator()(int r) const {
        const int wn = r / (8 * G), g = (r % (8 * G)) / 8, u = u0 + wn * 8 + (r & 7);
        return u < D ? g * D + u : -1;
    }
};

// acc[mi][ni][4] += A_tile (rows m0 .., one K segment) x W_tile^T (rows wmap(r), columns [wcol0, wcol0 + seg.width) of W).
// All 256 threads of the team call it (t = thread index in the team, named barrier `bar`); smem = Tile<BM,BN>::SMEM_BYTES.
template <int WM, int WN, int MI, int NI, class WMap>
__device__ __forceinline__ void gemm_accum(float (&acc)[MI][NI][4], const ASeg& a, int64_t m0, int64_t M, const Planes& W, int wcol0,
                                           const WMap& wmap, bf16* __restrict__ smem, int t, int bar) {
    static_assert(WM * WN * 32 == THREADS, "8 warps per team");
    constexpr int BM = WM * MI * 16, BN = WN * NI * 8;
    constexpr int NROW = (BM + BN + 31) / 32;
    constexpr int STAGE = Tile<BM, BN>::STAGE_ELEMS;
    const int piece = t & 7, row0 = t >> 3;
    const bf16* src[NROW];        // hi-plane source of the staged rows this thread copies; the mid plane sits `delta` elements further
    int64_t delta[NROW];
#pragma unroll
    for (int l = 0; l < NROW; ++l) {
        const int row = row0 + 32 * l;
        src[l] = nullptr;
        delta[l] = 0;
        if (row < BM) {
            const int64_t m = m0 + row;
            if (m < M) {
                src[l] = a.p.hi + (a.idx ? __ldg(a.idx + m) : m) * a.p.ld + a.col0 + piece * 8;
                delta[l] = a.p.mid - a.p.hi;
            }
        } else if (row < BM + BN) {
            const int wr = wmap(row - BM);
            if (wr >= 0) {
                src[l] = W.hi + (int64_t)wr * W.ld + wcol0 + piece * 8;
                delta[l] = W.mid - W.hi;
            }
        }
    }
    const int nk = (a.width + BK - 1) / BK;
    auto issue = [&](int kt) {
        if (kt < nk) {
            bf16* dst = smem + (kt % STAGES) * STAGE + piece * 8;
            const int k0 = kt * BK;
            const bool in_k = k0 + piece * 8 < a.width;
#pragma unroll
            for (int l = 0; l < NROW; ++l) {
                const int row = row0 + 32 * l;
                if (row >= BM + BN) continue;
                const bool ok = src[l] != nullptr && in_k;
                // stage layout: [A hi | A mid | W hi | W mid], rows of PITCH elements
                bf16* d = dst + (row < BM ? row : 2 * BM + (row - BM)) * PITCH;
                const int plane = (row < BM ? BM : BN) * PITCH;
                cp_async16(d, ok ? src[l] + k0 : W.hi, ok ? 16 : 0);                        // src-size 0: zero fill
                cp_async16(d + plane, ok ? src[l] + delta[l] + k0 : W.hi, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };
    const int warp = t >> 5, lane = t & 31;
    const int wm = warp / WN, wn = warp % WN;
    // ldmatrix row addresses (see the fragment layouts of mma.m16n8k16): A: lanes 0-15 rows 0-15 at k, lanes 16-31 the same rows at k + 8;
    // W: lanes 0-7 rows n..n+7 at k, 8-15 same rows at k + 8, 16-23 rows n+8.. at k, 24-31 rows n+8.. at k + 8
    const int a_off = (wm * MI * 16 + (lane & 15)) * PITCH + (lane >> 4) * 8;
    const int w_off = (wn * NI * 8 + (lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 8;
    const int w_off2 = (wn * NI * 8 + (lane & 7)) * PITCH + ((lane >> 3) & 1) * 8;          // x2 form: one 8-column block
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        team_sync(bar);
        issue(kt + STAGES - 1);
        const bf16* Ah = smem + (kt % STAGES) * STAGE;
        const bf16* Am = Ah + BM * PITCH;
        const bf16* Wh = Am + BM * PITCH;
        const bf16* Wm = Wh + BN * PITCH;
        const int kleft = a.width - kt * BK;
        const int nks = kleft >= BK ? BK / 16 : (kleft + 15) / 16;
        for (int ks = 0; ks < nks; ++ks) {
            const int k0 = ks * 16;
            uint32_t ah[MI][4], am[MI][4];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                ldsm_x4(ah[mi], Ah + a_off + mi * 16 * PITCH + k0);
                ldsm_x4(am[mi], Am + a_off + mi * 16 * PITCH + k0);
            }
#pragma unroll
            for (int ni = 0; ni < NI; ni += 2) {
                uint32_t bh[4], bm[4];
                if (ni + 1 < NI) {
                    ldsm_x4(bh, Wh + w_off + ni * 8 * PITCH + k0);
                    ldsm_x4(bm, Wm + w_off + ni * 8 * PITCH + k0);
                } else {
                    ldsm_x2(bh[0], bh[1], Wh + w_off2 + ni * 8 * PITCH + k0);
                    ldsm_x2(bm[0], bm[1], Wm + w_off2 + ni * 8 * PITCH + k0);
                    bh[2] = bh[3] = bm[2] = bm[3] = 0;
                }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    mma16816(acc[mi][ni], ah[mi], bh[0], bh[1]);
                    mma16816(acc[mi][ni], ah[mi], bm[0], bm[1]);
                    mma16816(acc[mi][ni], am[mi], bh[0], bh[1]);
                    if (ni + 1 < NI) {
                        mma16816(acc[mi][ni + 1], ah[mi], bh[2], bh[3]);
                        mma16816(acc[mi][ni + 1], ah[mi], bm[2], bm[3]);
                        mma16816(acc[mi][ni + 1], am[mi], bh[2], bh[3]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    team_sync(bar);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace mt
