// Probe of cp.async.bulk.tensor.2d tile::gather4 semantics on sm_100a: which box shape the tensor map needs, where the four
// rows land in shared memory under SWIZZLE_64B, what out-of-bounds row indices do, how many bytes complete_tx counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu && ./gather4_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int col, int r0, int r1, int r2, int r3, uint32_t expect,
                      uint16_t* out, int* status) {
    __shared__ __align__(1024) uint16_t tile[8 * 32];
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) tile[i] = 0xFFFF;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&bar)), "r"(expect) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col), "r"(r0), "r"(r1), "r"(r2),
            "r"(r3), "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
        int ok = 0;
        for (int it = 0; it < 2000000 && !ok; ++it) {
            uint32_t p;
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(p) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
            ok = p;
        }
        *status = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) out[i] = tile[i];
}

int main() {
    const int R = 64, C = 64;   // bf16 table, element (r, c) = r * 100 + c encoded as an integer bit pattern
    uint16_t h[R * C];
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 100 + c);
    uint16_t* d;
    cudaMalloc(&d, sizeof(h));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    uint16_t* dout;
    int* dstat;
    cudaMalloc(&dout, 8 * 32 * 2);
    cudaMalloc(&dstat, 4);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
        for (int expect = 256; expect <= 256; expect += 256) {
            CUtensorMap m;
            cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 2};
            cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
            CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("box_rows=%d encode rc=%d\n", box_rows, (int)rc);
            if (rc != CUDA_SUCCESS) continue;
            cudaMemset(dstat, 0, 4);
            probe<<<1, 32>>>(m, 32, 5, 0, 17, R + 3, (uint32_t)expect, dout, dstat);
            cudaError_t e = cudaDeviceSynchronize();
            int st = -1;
            uint16_t o[8 * 32];
            cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
            printf("  expect=%d: err=%s barrier_completed=%d\n", expect, cudaGetErrorString(e), st);
            for (int r = 0; r < 8; ++r) {
                printf("  smem row %d:", r);
                for (int c = 0; c < 32; c += 8) printf(" [%5u %5u ..]", o[r * 32 + c], o[r * 32 + c + 1]);
                printf("\n");
            }
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
