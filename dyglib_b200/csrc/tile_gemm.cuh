// fp32 tile contraction for the latency-bound small-batch kernels (fused GRU update, the one-launch TGN step).
//
// These steps move ~1 GFLOP through a chain of dependent phases at batch 200: what bounds them is the
// latency of each phase, not tensor throughput, so the contraction is plain FFMA in fp32 (exact parity with the reference's
// sgemm up to summation order) with operands staged by 16-byte cp.async straight from (gathered) rows:
//   acc[i][j] += sum_k A[row(m0 + ty + 16 i), k] * W[wrow(tx + 16 j), k]        256 threads = 16 x 16, ty = t >> 4, tx = t & 15
// A rows and W rows are both k-contiguous, so a stage is (BM + BN) rows x BK floats copied as 16-byte pieces into a
// (BK + 4)-float pitch: consecutive rows start 4 banks apart, which makes the 128-bit reads of 8 consecutive W rows conflict-free,
// and the two A rows a warp touches are broadcasts.  BK = 64: a first version with 16-column stages spent ~500 ns per stage
// whatever the tile size (L2 round trip + barrier per 16 columns, two warps per scheduler: profiles/r02_tgn_step.md); 64-column
// stages pay that once per 64 columns.  The 256 threads of a tile synchronise on a NAMED barrier, so two tile teams can share a CTA.
#pragma once
#include "common.cuh"

namespace tg {

constexpr int BK = 64;
constexpr int PITCH = BK + 4;
constexpr int STAGES = 3;
constexpr int THREADS = 256;     // threads of one tile team

__device__ __forceinline__ void team_sync(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(THREADS) : "memory"); }

// rows of A for one K segment: A[m, k] = tab[(idx ? idx[m] : m) * ld + k], k in [0, width); width % 4 == 0, rows 16-byte aligned
struct ASeg {
    const float* tab;
    const int64_t* idx;
    int64_t ld;
    int width;
};

template <int TM, int TN>
struct Tile {
    static constexpr int BM = 16 * TM, BN = 16 * TN;
    static constexpr int SMEM_FLOATS = STAGES * (BM + BN) * PITCH;
};

__device__ __forceinline__ void cp_async16(float* smem, const float* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// plain weight rows n0 .. n0 + BN - 1 of an (N, ldw) matrix
struct WRows {
    int n0, N;
    __device__ __forceinline__ int operator()(int r) const { return n0 + r < N ? n0 + r : -1; }
};
// gate-major rows of a recurrent cell's (G*D, K) weight: tile row r -> gate r / BU, unit u0 + r % BU
struct WGates {
    int u0, D, BU, G;
    __device__ __forceinline__ int operator()(int r) const {
        const int g = r / BU, u = u0 + r % BU;
        return (g < G && u < D) ? g * D + u : -1;
    }
};

// acc += A_tile (rows m0.., one K segment) x W_tile^T (rows wmap(r), columns [0, seg.width) of W, leading dimension ldw).
// All 256 threads of the tile team (t = thread index inside the team) must call it: it synchronises on named barrier `bar`.
// smem: Tile<TM,TN>::SMEM_FLOATS floats owned by the team, free again on return.
template <int TM, int TN, class WMap>
__device__ __forceinline__ void gemm_accum(float (&acc)[TM][TN], const ASeg& a, int64_t m0, int64_t M, const float* __restrict__ W,
                                           int64_t ldw, const WMap& wmap, float* __restrict__ smem, int t, int bar) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    constexpr int TPR = THREADS / 16;                 // a thread owns one 16-byte piece (of the BK / 4 = 16) of every TPR-th staged row
    constexpr int NROW = (BM + BN + TPR - 1) / TPR;
    const int tx = t & 15, ty = t >> 4;
    const int c4 = tx * 4;
    const float* src[NROW];
#pragma unroll
    for (int l = 0; l < NROW; ++l) {
        const int row = ty + l * TPR;
        src[l] = nullptr;
        if (row < BM) {
            const int64_t m = m0 + row;
            if (m < M) src[l] = a.tab + (a.idx ? __ldg(a.idx + m) : m) * a.ld + c4;
        } else if (row < BM + BN) {
            const int wr = wmap(row - BM);
            if (wr >= 0) src[l] = W + (int64_t)wr * ldw + c4;
        }
    }
    const int nk = (a.width + BK - 1) / BK;
    auto issue = [&](int kt) {
        if (kt < nk) {
            float* dst = smem + (kt % STAGES) * (BM + BN) * PITCH + c4;
            const int k0 = kt * BK;
            const bool in_k = k0 + c4 < a.width;
#pragma unroll
            for (int l = 0; l < NROW; ++l) {
                const int row = ty + l * TPR;
                if (row >= BM + BN) continue;
                const bool ok = src[l] != nullptr && in_k;
                cp_async16(dst + row * PITCH, ok ? src[l] + k0 : W, ok ? 16 : 0);     // src-size 0: zero fill
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        team_sync(bar);
        issue(kt + STAGES - 1);
        const float* As = smem + (kt % STAGES) * (BM + BN) * PITCH;
        const float* Ws = As + BM * PITCH;
        const int kmax = a.width - kt * BK < BK ? a.width - kt * BK : BK;     // multiple of 4: no FMAs on the zero fill
#pragma unroll 4
        for (int kk = 0; kk < kmax; kk += 4) {
            float4 av[TM], wv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * PITCH + kk);
#pragma unroll
            for (int j = 0; j < TN; ++j) wv[j] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * j) * PITCH + kk);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    float s = acc[i][j];
                    s = fmaf(av[i].x, wv[j].x, s);
                    s = fmaf(av[i].y, wv[j].y, s);
                    s = fmaf(av[i].z, wv[j].z, s);
                    s = fmaf(av[i].w, wv[j].w, s);
                    acc[i][j] = s;
                }
        }
    }
    cp_async_wait<0>();
    team_sync(bar);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace tg
