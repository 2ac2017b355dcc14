// Probe: how fast can one SM fill shared memory with gathered 128-byte row pieces (the stage copies of csrc/mma_tile.cuh)?
// Per "iteration" a team of 256 threads copies ROWS rows x 2 planes x 128 bytes; 3 buffers.  Variants:
//   0 cp.async.cg 16 B per thread (current)      1 ld.global.cg.v4 -> st.shared.v4 (register staged, next iteration prefetched)
//   2 cp.async.bulk 128 B per row-plane + mbarrier (one copy per thread)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/probes/bin/fill_probe scripts/probes/fill_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
typedef __nv_bfloat16 bf16;
constexpr int PITCH = 72, STAGES = 3;

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16z(void* s, const void* g, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(n)); }
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <int ROWS, int MODE>
__global__ void __launch_bounds__(256, 1) fill(const bf16* hi, const bf16* mid, int64_t ld, int nrows_total, int nk, int reps, float* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    bf16* smem = reinterpret_cast<bf16*>(raw);
    __shared__ uint64_t bars[STAGES];
    constexpr int STAGE = 2 * ROWS * PITCH;
    const int t = threadIdx.x;
    const int base_row = (blockIdx.x * ROWS) % (nrows_total - ROWS);
    float acc = 0.f;
    if (MODE == 2) {
        if (t == 0) for (int s = 0; s < STAGES; ++s) bar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    for (int rep = 0; rep < reps; ++rep) {
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            const int piece = t & 7, row0 = t >> 3;
            auto issue = [&](int kt) {
                if (kt < nk) {
                    bf16* dst = smem + (kt % STAGES) * STAGE + piece * 8;
                    for (int row = row0; row < ROWS; row += 32) {
                        const int grow = (MODE == 4 && row >= ROWS - 64) ? row : base_row + row;     // mode 4: the last 64 rows are the same for every CTA
                        const bf16* s = hi + (int64_t)grow * ld + kt * 64 + piece * 8;
                        if (MODE == 3) {
                            const int ok = (kt * 64 + piece * 8 < nk * 64) ? 16 : 0;
                            cp_async16z(dst + row * PITCH, s, ok);
                            cp_async16z(dst + (ROWS + row) * PITCH, s + (mid - hi), ok);
                        } else {
                            cp_async16(dst + row * PITCH, s);
                            cp_async16(dst + (ROWS + row) * PITCH, s + (mid - hi));
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            issue(0); issue(1);
            for (int kt = 0; kt < nk; ++kt) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncthreads();
                issue(kt + 2);
                acc += __bfloat162float(smem[(kt % STAGES) * STAGE + t]);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        } else if (MODE == 1) {
            const int piece = t & 7, row0 = t >> 3;
            constexpr int NL = ROWS / 32;
            uint4 rh[NL], rm[NL];
            auto load = [&](int kt) {
                for (int l = 0; l < NL; ++l) {
                    const bf16* s = hi + (int64_t)(base_row + row0 + 32 * l) * ld + kt * 64 + piece * 8;
                    rh[l] = __ldcg(reinterpret_cast<const uint4*>(s));
                    rm[l] = __ldcg(reinterpret_cast<const uint4*>(s + (mid - hi)));
                }
            };
            load(0);
            for (int kt = 0; kt < nk; ++kt) {
                bf16* dst = smem + (kt % 2) * STAGE + piece * 8;
                for (int l = 0; l < NL; ++l) {
                    *reinterpret_cast<uint4*>(dst + (row0 + 32 * l) * PITCH) = rh[l];
                    *reinterpret_cast<uint4*>(dst + (ROWS + row0 + 32 * l) * PITCH) = rm[l];
                }
                if (kt + 1 < nk) load(kt + 1);
                __syncthreads();
                acc += __bfloat162float(smem[(kt % 2) * STAGE + t]);
            }
        } else {
            // one 128-byte bulk copy per (row, plane): thread t < 2 * ROWS copies row t % ROWS of plane t / ROWS
            auto issue = [&](int kt) {
                if (kt < nk) {
                    uint64_t* bar = &bars[kt % STAGES];
                    if (t == 0) bar_expect(bar, 2 * ROWS * 128);
                    __syncwarp();
                    if (t < 2 * ROWS) {
                        const int row = t % ROWS, plane = t / ROWS;
                        const bf16* s = (plane ? mid : hi) + (int64_t)(base_row + row) * ld + kt * 64;
                        bulk_g2s(smem + (kt % STAGES) * STAGE + (plane * ROWS + row) * PITCH, s, 128, bar);
                    }
                }
            };
            const int use0 = rep * nk;
            // thread 0's expect must precede the copies of other warps: order with a barrier
            if (0 < nk) { if (t == 0) bar_expect(&bars[(use0 + 0) % STAGES], 2 * ROWS * 128); }
            if (1 < nk) { if (t == 0) bar_expect(&bars[(use0 + 1) % STAGES], 2 * ROWS * 128); }
            __syncthreads();
            for (int s2 = 0; s2 < 2 && s2 < nk; ++s2)
                if (t < 2 * ROWS) {
                    const int row = t % ROWS, plane = t / ROWS;
                    bulk_g2s(smem + ((use0 + s2) % STAGES) * STAGE + (plane * ROWS + row) * PITCH, (plane ? mid : hi) + (int64_t)(base_row + row) * ld + s2 * 64, 128,
                             &bars[(use0 + s2) % STAGES]);
                }
            for (int kt = 0; kt < nk; ++kt) {
                const int u = use0 + kt;
                bar_wait(&bars[u % STAGES], (u / STAGES) & 1);
                acc += __bfloat162float(smem[(u % STAGES) * STAGE + t]);
                __syncthreads();   // everyone is done with the buffer that the next issue overwrites
                if (kt + 2 < nk) {
                    const int un = u + 2;
                    if (t == 0) bar_expect(&bars[un % STAGES], 2 * ROWS * 128);
                    __syncthreads();
                    if (t < 2 * ROWS) {
                        const int row = t % ROWS, plane = t / ROWS;
                        bulk_g2s(smem + (un % STAGES) * STAGE + (plane * ROWS + row) * PITCH, (plane ? mid : hi) + (int64_t)(base_row + row) * ld + (kt + 2) * 64, 128,
                                 &bars[un % STAGES]);
                    }
                }
            }
            (void)issue;
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}

template <int ROWS, int MODE>
void run(const bf16* hi, const bf16* mid, int64_t ld, int nrows, int nk, int grid) {
    const int smem = STAGES * 2 * ROWS * PITCH * 2;
    float* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(fill<ROWS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 50;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fill<ROWS, MODE><<<grid, 256, smem>>>(hi, mid, ld, nrows, nk, 2, sink);
    cudaEventRecord(e0);
    fill<ROWS, MODE><<<grid, 256, smem>>>(hi, mid, ld, nrows, nk, reps, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    const double per_iter_us = ms * 1000.0 / (reps * nk);
    const double bytes = 2.0 * ROWS * 128;
    printf("  rows=%3d mode=%d grid=%3d: %6.3f us per iteration, %6.1f B/clk/SM (at 1.965 GHz), %5.2f TB/s aggregate %s\n", ROWS, MODE, grid, per_iter_us,
           bytes / (per_iter_us * 1965.0), bytes * grid / per_iter_us / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink);
}

int main() {
    const int nrows = 4096, K = 896; const int64_t ld = K;
    bf16* buf; cudaMalloc(&buf, (size_t)2 * nrows * ld * 2); cudaMemset(buf, 0, (size_t)2 * nrows * ld * 2);
    const bf16 *hi = buf, *mid = buf + (size_t)nrows * ld;
    for (int grid : {148, 95, 48}) {
        printf("grid %d CTAs x 256 threads, %d iterations of 64 columns per pass (L2-resident source)\n", grid, K / 64);
        run<96, 0>(hi, mid, ld, nrows, K / 64, grid); run<96, 3>(hi, mid, ld, nrows, K / 64, grid); run<96, 4>(hi, mid, ld, nrows, K / 64, grid);
        run<128, 0>(hi, mid, ld, nrows, K / 64, grid); run<128, 4>(hi, mid, ld, nrows, K / 64, grid);
    }
    return 0;
}
