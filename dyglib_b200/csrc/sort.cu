// Stable LSD radix sort of (key, value) pairs for the CSR build (utils/utils.py:96-103: per-node lists sorted by time with Python's
// stable sorted(); here: stable by time key, then stable by owner).  One 8-bit digit per pass:
//   dyg_radix_digit_hist   256-bin histogram of every digit of the keys in one read (the host skips passes whose digit is constant)
//   dyg_radix_sort_pass    block histograms -> digit-major exclusive scan -> stable scatter
// Stability inside a block: the 8 warps own consecutive 512-key segments of the block's 4096-key tile and walk them in order, 32
// consecutive keys per round; the rank of a key among the equal digits of its round comes from __match_any_sync, the ranks of
// earlier rounds / earlier warps / earlier blocks from running counters seeded by the scans.  No atomics decide an output position.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256, RS_WARPS = 8, RS_ROUNDS = 16;
constexpr int RS_TILE = RS_WARPS * RS_ROUNDS * 32;       // 4096 keys per block

template <typename K>
__device__ __forceinline__ unsigned digit_of(K key, int pass) {
    return (unsigned)((key >> (8 * pass)) & 0xFF);
}

template <typename K>
__global__ void __launch_bounds__(256) radix_digit_hist_kernel(const K* __restrict__ keys, int64_t n, unsigned long long* __restrict__ hist) {
    __shared__ unsigned sh[sizeof(K)][256];
    for (int i = threadIdx.x; i < (int)sizeof(K) * 256; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const K k = keys[i];
#pragma unroll
        for (int p = 0; p < (int)sizeof(K); ++p) atomicAdd(&sh[p][digit_of(k, p)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)sizeof(K) * 256; i += blockDim.x) {
        const unsigned v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, (unsigned long long)v);
    }
}

// block_hist[d * nblocks + b] = number of keys of block b's tile with digit d
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) radix_block_hist_kernel(const K* __restrict__ keys, int64_t n, int pass, unsigned* __restrict__ block_hist,
                                                                      int64_t nblocks) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const int64_t j = base + i;
        if (j < n) atomicAdd(&sh[digit_of(keys[j], pass)], 1u);
    }
    __syncthreads();
    block_hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// in-place exclusive scan of the digit-major array (256 rows of nblocks entries): block d scans row d, seeded with the number of keys whose
// digit is smaller (from the digit totals of dyg_radix_digit_hist), in chunks of 1024 with a carried total
__global__ void __launch_bounds__(1024) radix_scan_kernel(unsigned* __restrict__ a, int64_t nblocks, const unsigned long long* __restrict__ totals) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry_s;
    const int d = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        unsigned long long part = (int)threadIdx.x < d ? totals[threadIdx.x] : 0ull;      // d <= 255 < blockDim.x
        unsigned p = (unsigned)part;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0) warp_tot[warp] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < 32; ++w) t += warp_tot[w];
            carry_s = t;
        }
        __syncthreads();
    }
    unsigned* row = a + (int64_t)d * nblocks;
    for (int64_t c0 = 0; c0 < nblocks; c0 += 1024) {
        const int64_t i = c0 + threadIdx.x;
        const unsigned v = i < nblocks ? row[i] : 0u;
        unsigned s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        __syncthreads();                                          // warp_tot of the previous chunk has been read
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;                                   // inclusive over warps
        }
        __syncthreads();
        const unsigned before = carry_s + (warp ? warp_tot[warp - 1] : 0u);
        if (i < nblocks) row[i] = before + s - v;                 // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + s;
        __syncthreads();
    }
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const K* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                                   K* __restrict__ keys_out, unsigned* __restrict__ vals_out, int64_t n, int pass,
                                                                   const unsigned* __restrict__ block_base, int64_t nblocks) {
    __shared__ unsigned cnt[RS_WARPS][256];                      // warp histograms, then running output positions
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0u;
    __syncthreads();
    const int64_t seg = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ROUNDS * 32);
    K key[RS_ROUNDS];
    unsigned val[RS_ROUNDS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t j = seg + r * 32 + lane;
        const bool ok = j < n;
        key[r] = ok ? keys_in[j] : (K)0;
        val[r] = ok ? (vals_in ? vals_in[j] : (unsigned)j) : 0u;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned d = digit_of(key[r], pass);
            const unsigned peers = __match_any_sync(active, d);
            if ((peers & lt) == 0u) cnt[warp][d] += __popc(peers);                 // the lowest lane of a peer group adds its size
        }
        __syncwarp();
    }
    __syncthreads();
    {
        // digit d = thread d: exclusive prefix over the warps, seeded with this block's global base for the digit
        const int d = threadIdx.x;
        unsigned run = block_base[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const unsigned c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t j = seg + r * 32 + lane;
        const bool ok = j < n;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned d = digit_of(key[r], pass);
            const unsigned peers = __match_any_sync(active, d);
            const unsigned pos = cnt[warp][d] + __popc(peers & lt);
            __syncwarp(active);                                                    // every lane has read the running position
            if ((peers & lt) == 0u) cnt[warp][d] += __popc(peers);
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int dyg_radix_digit_hist(const void* keys, int key_bytes, int64_t n, unsigned long long* hist, dyg_stream_t stream) {
    DYG_CHECK_ARG((key_bytes == 4 || key_bytes == 8) && n >= 0 && hist, "dyg_radix_digit_hist: key_bytes must be 4 or 8");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(hist, 0, (size_t)key_bytes * 256 * sizeof(unsigned long long), s);
    if (n == 0) return 0;
    DYG_CHECK_ARG(keys, "dyg_radix_digit_hist: NULL keys");
    int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const int64_t cap = 8 * (int64_t)dyg_num_sms();
    if (blocks > cap) blocks = cap;
    if (key_bytes == 4) radix_digit_hist_kernel<uint32_t><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(keys), n, hist);
    else radix_digit_hist_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(keys), n, hist);
    DYG_LAUNCH_CHECK("dyg_radix_digit_hist");
    return 0;
}

extern "C" int64_t dyg_radix_sort_workspace_entries(int64_t n) { return 256 * ((n + RS_TILE - 1) / RS_TILE); }

extern "C" int dyg_radix_sort_pass(const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out, int key_bytes, int64_t n,
                                   int pass, const unsigned long long* hist, uint32_t* workspace, dyg_stream_t stream) {
    DYG_CHECK_ARG((key_bytes == 4 || key_bytes == 8) && n >= 0 && pass >= 0 && pass < key_bytes, "dyg_radix_sort_pass: bad key size / pass");
    DYG_CHECK_ARG(n < ((int64_t)1 << 32), "dyg_radix_sort_pass: more than 2^32 - 1 keys");
    if (n == 0) return 0;
    DYG_CHECK_ARG(keys_in && keys_out && vals_out && workspace && hist && keys_in != keys_out, "dyg_radix_sort_pass: NULL / aliased buffers");
    const int64_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    DYG_CHECK_ARG(nblocks < ((int64_t)1 << 31), "dyg_radix_sort_pass: too many tiles");
    cudaStream_t s = as_stream(stream);
    if (key_bytes == 4) {
        const uint32_t* ki = reinterpret_cast<const uint32_t*>(keys_in);
        radix_block_hist_kernel<uint32_t><<<(unsigned)nblocks, RS_THREADS, 0, s>>>(ki, n, pass, workspace, nblocks);
        radix_scan_kernel<<<256, 1024, 0, s>>>(workspace, nblocks, hist + 256 * pass);
        radix_scatter_kernel<uint32_t><<<(unsigned)nblocks, RS_THREADS, 0, s>>>(ki, vals_in, reinterpret_cast<uint32_t*>(keys_out), vals_out, n, pass,
                                                                                workspace, nblocks);
    } else {
        const unsigned long long* ki = reinterpret_cast<const unsigned long long*>(keys_in);
        radix_block_hist_kernel<unsigned long long><<<(unsigned)nblocks, RS_THREADS, 0, s>>>(ki, n, pass, workspace, nblocks);
        radix_scan_kernel<<<256, 1024, 0, s>>>(workspace, nblocks, hist + 256 * pass);
        radix_scatter_kernel<unsigned long long><<<(unsigned)nblocks, RS_THREADS, 0, s>>>(ki, vals_in, reinterpret_cast<unsigned long long*>(keys_out),
                                                                                           vals_out, n, pass, workspace, nblocks);
    }
    DYG_LAUNCH_CHECK("dyg_radix_sort_pass");
    return 0;
}
