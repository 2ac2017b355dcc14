// Probe: issue rate of legacy mma.sync.m16n8k16 bf16 on sm_100a (8 / 16 warps per SM, independent accumulators).
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k(float* out, int iters, long long* cyc) {
    float c[NACC][4] = {};
    uint32_t a[4] = {threadIdx.x, 2, 3, 4};
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma16816(c[j], a, i, j);
    const long long t1 = clock64();
    float s = 0; for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    for (int threads : {128, 256, 512, 1024}) {
        k<6><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
        k<6><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
        const double per_sm = (double)iters * 6 * (threads / 32);
        printf("warps/SM %2d: %lld cycles for %d x 6 MMAs per warp -> %.2f cycles per MMA per SM sub-partition, %.0f bf16 MAC/clk/SM\n", threads / 32, *cyc, iters,
               *cyc / (per_sm / 4), per_sm * 2048 / *cyc);
    }
    return 0;
}
