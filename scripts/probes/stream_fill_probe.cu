// Probe: how fast can every SM stream a large array from HBM into shared memory (one CTA per SM, 128 KB stages, data used once)?
//   mode 0: cp.async.ca 16 B per thread (256 threads), commit groups        mode 1: cp.async.bulk of 8 KB pieces + mbarrier (one thread)
//   mode 2: cp.async.bulk of one 64 KB piece + mbarrier
// DEPTH stages of 64 KB in flight.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/probes/bin/stream_fill_probe ...
#include <cstdio>
#include <cstdint>
constexpr int STAGE = 64 * 1024;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_L:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_D;\n\tbra W_L;\n\tW_D:\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(256, 1) fill(const unsigned char* src, int64_t per_cta_bytes, float* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bars[DEPTH];
    const int t = threadIdx.x;
    const unsigned char* base = src + (int64_t)blockIdx.x * per_cta_bytes;
    const int nst = (int)(per_cta_bytes / STAGE);
    if (MODE != 0) {
        if (t == 0) for (int s = 0; s < DEPTH; ++s) bar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    float acc = 0.f;
    auto issue = [&](int st) {
        if (st >= nst) { if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory"); return; }
        unsigned char* d = sm + (st % DEPTH) * STAGE;
        const unsigned char* g = base + (int64_t)st * STAGE;
        if (MODE == 0) {
            for (int i = t; i < STAGE / 16; i += 256)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(d + i * 16)), "l"(g + i * 16) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else if (t == 0) {
            bar_expect(&bars[st % DEPTH], STAGE);
            if (MODE == 1) for (int i = 0; i < STAGE / 8192; ++i) bulk(d + i * 8192, g + i * 8192, 8192, &bars[st % DEPTH]);
            else bulk(d, g, STAGE, &bars[st % DEPTH]);
        }
    };
    for (int s = 0; s < DEPTH - 1; ++s) issue(s);
    for (int st = 0; st < nst; ++st) {
        issue(st + DEPTH - 1);
        if (MODE == 0) { asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"); }
        else bar_wait(&bars[st % DEPTH], (st / DEPTH) & 1);
        __syncthreads();
        acc += sm[(st % DEPTH) * STAGE + t * 16];      // touch
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}
template <int MODE, int DEPTH>
void run(const unsigned char* src, int64_t per, float* sink, const char* name) {
    const size_t smem = (size_t)DEPTH * STAGE;
    cudaFuncSetAttribute(fill<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fill<MODE, DEPTH><<<148, 256, smem>>>(src, per, sink);
    cudaEventRecord(e0);
    fill<MODE, DEPTH><<<148, 256, smem>>>(src, per, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s depth %d: %.3f ms, %.2f TB/s, %.1f GB/s per SM  (%s)\n", name, DEPTH, ms, 148.0 * per / ms / 1e9, per / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int64_t per = 32ll << 20;                       // 32 MB per CTA: 4.7 GB in total, far beyond L2
    unsigned char* src; float* sink;
    cudaMalloc(&src, 148 * per); cudaMemset(src, 1, 148 * per); cudaMalloc(&sink, 4);
    run<0, 1>(src, per, sink, "cp.async 16 B");
    run<0, 2>(src, per, sink, "cp.async 16 B");
    run<0, 3>(src, per, sink, "cp.async 16 B");
    run<1, 1>(src, per, sink, "bulk 8 KB pieces");
    run<1, 2>(src, per, sink, "bulk 8 KB pieces");
    run<1, 3>(src, per, sink, "bulk 8 KB pieces");
    run<2, 1>(src, per, sink, "bulk 64 KB");
    run<2, 2>(src, per, sink, "bulk 64 KB");
    run<2, 3>(src, per, sink, "bulk 64 KB");
    return 0;
}
