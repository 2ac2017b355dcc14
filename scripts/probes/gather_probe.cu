// Probe: throughput of random row-piece gathers into shared memory with 16-byte cp.async, as patch_project_kernel issues them.
//   PIECE = 64:  4 lanes per 64-byte piece of a (row, plane), rows at a 352-byte pitch (current layout: pieces straddle 128-byte lines)
//   PIECE = 128: 8 lanes per 128-byte piece, rows at a 384-byte pitch (every piece is exactly one line)
// Grid = 2 CTAs per SM (PIECE 64) or 1 (PIECE 128) x 256 threads, DEPTH commit groups in flight; table of `rows` rows (L2 + HBM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/probes/bin/gather_probe scripts/probes/gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ void cp16(uint32_t s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
template <int PIECE, int DEPTH>
__global__ void __launch_bounds__(256) gather(const unsigned char* tab, int64_t pitch, const int* idx, int n_idx, int stages_per_cta, int blocks_per_row) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int LANES = PIECE / 16;                 // lanes per piece
    constexpr int ROWS_PER_INSTR = 256 / LANES;       // rows covered by one instruction of the CTA
    constexpr int INSTR = 128 / ROWS_PER_INSTR;       // instructions per thread for 128 rows
    constexpr int STAGE = 128 * PIECE;                // one plane
    const int t = threadIdx.x, chunk = t % LANES, r0 = t / LANES;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    int cursor = (blockIdx.x * 7919) % n_idx;
    for (int s = 0; s < stages_per_cta; ++s) {
        const int slot = s % DEPTH;
        const int blk = s % blocks_per_row;
#pragma unroll
        for (int i = 0; i < INSTR; ++i) {
            const int row = r0 + i * ROWS_PER_INSTR;
            const int64_t g = idx[(cursor + row) % n_idx];
            cp16(sbase + slot * STAGE + row * PIECE + chunk * 16, tab + g * pitch + blk * PIECE + chunk * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        if (blk == blocks_per_row - 1) cursor = (cursor + 128) % n_idx;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <int PIECE, int DEPTH>
void run(const unsigned char* tab, int64_t pitch, const int* idx, int n_idx, int ctas, const char* name) {
    const int stages = 4000, bpr = (PIECE == 64) ? 5 : 3;
    const size_t smem = (size_t)DEPTH * 128 * PIECE;
    cudaFuncSetAttribute(gather<PIECE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<PIECE, DEPTH><<<ctas, 256, smem>>>(tab, pitch, idx, n_idx, 200, bpr);
    cudaEventRecord(e0);
    gather<PIECE, DEPTH><<<ctas, 256, smem>>>(tab, pitch, idx, n_idx, stages, bpr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)ctas * stages * 128 * PIECE;
    printf("%-44s ctas %4d depth %d: %.3f ms, %.2f TB/s (%.1f GB/s per SM), err=%s\n", name, ctas, DEPTH, ms, bytes / ms / 1e9, bytes / ms / 1e6 / 148,
           cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 157475;     // edge table of the wiki config; 1300000 for lastfm
    const int n_idx = 1 << 22;
    unsigned char* tab; int* idx;
    cudaMalloc(&tab, rows * 384 + 4096); cudaMemset(tab, 1, rows * 384 + 4096);
    int* h = (int*)malloc(n_idx * sizeof(int));
    srand(1);
    for (int i = 0; i < n_idx; ++i) h[i] = (int)(((int64_t)rand() * 32768 + rand()) % rows);
    cudaMalloc(&idx, n_idx * sizeof(int)); cudaMemcpy(idx, h, n_idx * sizeof(int), cudaMemcpyHostToDevice);
    printf("table rows %lld\n", (long long)rows);
    run<64, 4>(tab, 352, idx, n_idx, 296, "64-byte pieces, 352-byte pitch, 2 CTAs/SM");
    run<64, 8>(tab, 352, idx, n_idx, 296, "64-byte pieces, 352-byte pitch, 2 CTAs/SM");
    run<64, 8>(tab, 384, idx, n_idx, 296, "64-byte pieces, 384-byte pitch, 2 CTAs/SM");
    run<128, 4>(tab, 384, idx, n_idx, 148, "128-byte pieces, 384-byte pitch, 1 CTA/SM");
    run<128, 4>(tab, 384, idx, n_idx, 296, "128-byte pieces, 384-byte pitch, 2 CTAs/SM");
    run<128, 8>(tab, 384, idx, n_idx, 148, "128-byte pieces, 384-byte pitch, 1 CTA/SM");
    run<128, 8>(tab, 384, idx, n_idx, 296, "128-byte pieces, 384-byte pitch, 2 CTAs/SM");
    return 0;
}
