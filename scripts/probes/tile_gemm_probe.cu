// Probe: where does a persistent-grid FFMA tile GEMM phase (csrc/tile_gemm.cuh) spend its time?  Variants of one phase
// C(M,N) = A(M,K) W(N,K)^T on a grid of one CTA per SM: full / loads only / compute only / W replicated per m-tile.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dyglib_b200/csrc -o scripts/probes/bin/tile_gemm_probe scripts/probes/tile_gemm_probe.cu
#include <cstdio>
#include <vector>
#include "tile_gemm.cuh"
void dyg_set_error(const char*, ...) {}

template <int TM, int MODE>   // MODE 0 full, 1 loads only, 2 compute only, 3 W replicated per m-tile
__global__ void __launch_bounds__(256, 1) probe(const float* A, const float* W, float* C, int M, int N, int K, int ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int ntn = (N + 63) / 64;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m0 = (int64_t)(tile / ntn) * (16 * TM);
        const int n0 = (tile % ntn) * 64;
        const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
        float acc[TM][4];
        for (int i = 0; i < TM; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const float* Wt = MODE == 3 ? W + (size_t)(tile / ntn) * N * K : W;
        if (MODE == 0 || MODE == 3) {
            tg::gemm_accum<TM, 4>(acc, tg::ASeg{A, nullptr, K, K}, m0, M, Wt, K, tg::WRows{n0, N}, smem, t, 1);
        } else if (MODE == 1) {
            const int nk = (K + tg::BK - 1) / tg::BK;
            for (int kt = 0; kt < nk; ++kt) {
                float* dst = smem + (kt % tg::STAGES) * (16 * TM + 64) * tg::PITCH + tx * 4;
                for (int row = ty; row < 16 * TM + 64; row += 16) {
                    const float* src = row < 16 * TM ? A + (m0 + row < M ? m0 + row : 0) * K : W + (size_t)(n0 + row - 16 * TM < N ? n0 + row - 16 * TM : 0) * K;
                    const bool ok = kt * tg::BK + tx * 4 < K;
                    tg::cp_async16(dst + row * tg::PITCH, ok ? src + kt * tg::BK + tx * 4 : W, ok ? 16 : 0);
                }
                tg::cp_async_commit();
                tg::cp_async_wait<tg::STAGES - 2>();
                tg::team_sync(1);
            }
            tg::cp_async_wait<0>();
        } else {
            const int nk = (K + tg::BK - 1) / tg::BK;
            for (int kt = 0; kt < nk; ++kt) {
                const float* As = smem + (kt % tg::STAGES) * (16 * TM + 64) * tg::PITCH;
                const float* Ws = As + 16 * TM * tg::PITCH;
                tg::team_sync(1);
#pragma unroll 4
                for (int kk = 0; kk < tg::BK; kk += 4) {
                    float4 av[TM], wv[4];
                    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * tg::PITCH + kk);
                    for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * j) * tg::PITCH + kk);
                    for (int i = 0; i < TM; ++i) for (int j = 0; j < 4; ++j) {
                        float s = acc[i][j];
                        s = fmaf(av[i].x, wv[j].x, s); s = fmaf(av[i].y, wv[j].y, s); s = fmaf(av[i].z, wv[j].z, s); s = fmaf(av[i].w, wv[j].w, s);
                        acc[i][j] = s;
                    }
                }
            }
        }
        for (int j = 0; j < 4; ++j) for (int i = 0; i < TM; ++i) {
            const int64_t m = m0 + ty + 16 * i; const int n = n0 + tx + 16 * j;
            if (m < M && n < N) C[m * N + n] = acc[i][j];
        }
    }
}

template <int TM, int MODE>
float run(const float* A, const float* W, float* C, int M, int N, int K, int grid) {
    const int ntiles = ((M + 16 * TM - 1) / (16 * TM)) * ((N + 63) / 64);
    const int smem = tg::Tile<TM, 4>::SMEM_FLOATS * 4;
    cudaFuncSetAttribute(probe<TM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) probe<TM, MODE><<<grid, 256, smem>>>(A, W, C, M, N, K, ntiles);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) probe<TM, MODE><<<grid, 256, smem>>>(A, W, C, M, N, K, ntiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return ms * 1000.f / 20;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    struct Shape { int M, N, K; } shapes[] = {{600, 272, 888}, {600, 888, 172}, {600, 172, 444}, {600, 172, 172}, {400, 172, 344}};
    for (auto s : shapes) {
        float *A, *W, *C;
        cudaMalloc(&A, (size_t)s.M * s.K * 4); cudaMalloc(&W, (size_t)64 * s.N * s.K * 4); cudaMalloc(&C, (size_t)s.M * s.N * 4);
        cudaMemset(A, 0, (size_t)s.M * s.K * 4); cudaMemset(W, 0, (size_t)64 * s.N * s.K * 4);
        printf("M=%d N=%d K=%d (BK=%d, %d stages, grid %d x 256 threads)\n", s.M, s.N, s.K, tg::BK, tg::STAGES, sms);
#define ROW(TM) printf("  TM=%d tiles=%3d: full %7.2f us | loads only %7.2f | compute only %7.2f | W replicated %7.2f\n", TM, \
        ((s.M + 16 * TM - 1) / (16 * TM)) * ((s.N + 63) / 64), run<TM, 0>(A, W, C, s.M, s.N, s.K, sms), run<TM, 1>(A, W, C, s.M, s.N, s.K, sms), \
        run<TM, 2>(A, W, C, s.M, s.N, s.K, sms), run<TM, 3>(A, W, C, s.M, s.N, s.K, sms));
        ROW(1) ROW(2) ROW(4)
        cudaFree(A); cudaFree(W); cudaFree(C);
    }
    return 0;
}
