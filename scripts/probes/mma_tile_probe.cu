// Probe: memory pipeline of the BF16x3 mma.sync tile (csrc/mma_tile.cuh) on a persistent grid: one phase C = A W^T, one team per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dyglib_b200/csrc -DMT_STAGES=3 -o scripts/probes/bin/mma_tile_probe3 scripts/probes/mma_tile_probe.cu
#include <cstdio>
#include "mma_tile.cuh"
void dyg_set_error(const char*, ...) {}

template <int MI, int MODE>   // MODE 0 full, 1 loads + barriers only (no ldmatrix / mma)
__global__ void __launch_bounds__(256, 1) probe(mt::Planes A, mt::Planes W, float* C, int M, int N, int K, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    mt::bf16* smem = reinterpret_cast<mt::bf16*>(smem_raw);
    constexpr int WM = 2, WN = 4, NI = 2, BM = WM * MI * 16, BN = 64;
    const int ntn = (N + BN - 1) / BN;
    const int t = threadIdx.x;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m0 = (int64_t)(tile / ntn) * BM;
        const int n0 = (tile % ntn) * BN;
        float acc[MI][NI][4] = {};
        if (MODE == 0) {
            mt::gemm_accum<WM, WN, MI, NI>(acc, mt::ASeg{A, nullptr, 0, K}, m0, M, W, 0, mt::WRows{n0, N}, smem, t, 1);
        } else {
            constexpr int STAGE = mt::Tile<BM, BN>::STAGE_ELEMS;
            const int piece = t & 7, row0 = t >> 3;
            const int nk = (K + mt::BK - 1) / mt::BK;
            auto issue = [&](int kt) {
                if (kt < nk) {
                    mt::bf16* dst = smem + (kt % mt::STAGES) * STAGE + piece * 8;
                    for (int row = row0; row < BM + BN; row += 32) {
                        const mt::bf16* src = row < BM ? A.hi + (m0 + row < M ? m0 + row : 0) * A.ld : W.hi + (int64_t)(n0 + row - BM < N ? n0 + row - BM : 0) * W.ld;
                        const int64_t delta = row < BM ? A.mid - A.hi : W.mid - W.hi;
                        mt::bf16* d = dst + (row < BM ? row : 2 * BM + row - BM) * mt::PITCH;
                        mt::cp_async16(d, src + kt * mt::BK + piece * 8, 16);
                        mt::cp_async16(d + (row < BM ? BM : BN) * mt::PITCH, src + delta + kt * mt::BK + piece * 8, 16);
                    }
                }
                mt::cp_async_commit();
            };
            for (int s = 0; s < mt::STAGES - 1; ++s) issue(s);
            for (int kt = 0; kt < nk; ++kt) {
                mt::cp_async_wait<mt::STAGES - 2>();
                mt::team_sync(1);
                issue(kt + mt::STAGES - 1);
            }
            mt::cp_async_wait<0>();
            mt::team_sync(1);
        }
        const int warp = t >> 5, lane = t & 31, wm = warp / WN, wn = warp % WN, gq = lane >> 2, tq = lane & 3;
        for (int j = 0; j < NI; ++j) for (int i = 0; i < MI; ++i) for (int half = 0; half < 2; ++half) {
            const int64_t m = m0 + wm * MI * 16 + i * 16 + gq + 8 * half; const int n = n0 + wn * NI * 8 + j * 8 + 2 * tq;
            if (m < M && n < N) *reinterpret_cast<float2*>(C + m * N + n) = make_float2(acc[i][j][2 * half], acc[i][j][2 * half + 1]);
        }
    }
}

template <int MI, int MODE>
float run(mt::Planes A, mt::Planes W, float* C, int M, int N, int K, int grid) {
    const int ntiles = ((M + 32 * MI - 1) / (32 * MI)) * ((N + 63) / 64);
    const int smem = mt::Tile<32 * MI, 64>::SMEM_BYTES;
    cudaFuncSetAttribute(probe<MI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) probe<MI, MODE><<<grid, 256, smem>>>(A, W, C, M, N, K, ntiles);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) probe<MI, MODE><<<grid, 256, smem>>>(A, W, C, M, N, K, ntiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return ms * 1000.f / 20;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    struct Shape { int M, N, K; } shapes[] = {{600, 272, 896}, {600, 888, 176}, {600, 172, 448}, {600, 172, 176}, {400, 172, 352}};
    for (auto s : shapes) {
        mt::bf16 *A, *W; float* C;
        cudaMalloc(&A, (size_t)2 * s.M * s.K * 2); cudaMalloc(&W, (size_t)2 * s.N * s.K * 2); cudaMalloc(&C, (size_t)s.M * s.N * 4);
        cudaMemset(A, 0, (size_t)2 * s.M * s.K * 2); cudaMemset(W, 0, (size_t)2 * s.N * s.K * 2);
        mt::Planes Ap{A, A + (size_t)s.M * s.K, s.K}, Wp{W, W + (size_t)s.N * s.K, s.K};
        printf("M=%d N=%d K=%d (BK=%d, %d stages, grid %d x 256 threads)\n", s.M, s.N, s.K, mt::BK, mt::STAGES, sms);
        printf("  32x64 tiles=%3d: full %7.2f us | loads only %7.2f\n", ((s.M + 31) / 32) * ((s.N + 63) / 64), run<1, 0>(Ap, Wp, C, s.M, s.N, s.K, sms), run<1, 1>(Ap, Wp, C, s.M, s.N, s.K, sms));
        printf("  64x64 tiles=%3d: full %7.2f us | loads only %7.2f\n", ((s.M + 63) / 64) * ((s.N + 63) / 64), run<2, 0>(Ap, Wp, C, s.M, s.N, s.K, sms), run<2, 1>(Ap, Wp, C, s.M, s.N, s.K, sms));
        cudaFree(A); cudaFree(W); cudaFree(C);
    }
    // empty kernel launch overhead for reference
    return 0;
}
