// TGN / DyRep / JODIE memory path on device.
//
// Device state replaces the reference's python dict-of-lists (models/MemoryModel.py:304-407):
//   memory (N,D), last_update (N)            persisted state              (MemoryBank.node_memories / node_last_updated_times)
//   mem_view (N,D), lu_view (N), pending (N) look-ahead state == what get_updated_memories() would return for
//                                            every node (models/MemoryModel.py:461-487), maintained incrementally
//   msg_store (N,msg_dim), msg_time (N)      the last raw message of each node (node_raw_messages[v][-1])
// A node's pending message and its memory cannot change between the batch that stored the message and the
// batch that consumes it, so mem_view[v] = cell(msg_v, memory[v]) is computed once, at store time.
#include <math.h>
#include "common.cuh"

__global__ void tgn_persist_kernel(const int64_t* __restrict__ ids, int64_t n, float* __restrict__ memory,
                                   const float* __restrict__ mem_view, float* __restrict__ last_update,
                                   const float* __restrict__ lu_view, uint8_t* __restrict__ pending, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t e = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (e >= n) return;
    const int64_t v = ids[e];
    // duplicates of v copy identical values; the flag is cleared by a second kernel so that every duplicate sees it set
    if (!pending[v]) return;
    for (int c = lane; c < D; c += 32) memory[v * D + c] = mem_view[v * D + c];
    if (lane == 0) last_update[v] = lu_view[v];
}
__global__ void tgn_clear_pending_kernel(const int64_t* __restrict__ ids, int64_t n, uint8_t* __restrict__ pending) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n) pending[ids[e]] = 0;
}
extern "C" int dyg_tgn_persist(const int64_t* node_ids, int64_t n, float* memory, const float* mem_view,
                               float* last_update, const float* lu_view, uint8_t* pending, int D, dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0 && D > 0, "dyg_tgn_persist: bad sizes");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    tgn_persist_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(node_ids, n, memory, mem_view, last_update, lu_view, pending, D);
    tgn_clear_pending_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(node_ids, n, pending);
    DYG_LAUNCH_CHECK("dyg_tgn_persist");
    return 0;
}

__global__ void tgn_select_last_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t B,
                                       int32_t* __restrict__ winner) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= 2 * B) return;
    const int64_t v = c < B ? src[c] : dst[c - B];
    atomicMax(winner + v, (int32_t)c);
}
extern "C" int dyg_tgn_select_last(const int64_t* src, const int64_t* dst, int64_t B, int32_t* winner, dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && 2 * B < (1ll << 31), "dyg_tgn_select_last: bad batch size");
    if (B == 0) return 0;
    tgn_select_last_kernel<<<(unsigned)((2 * B + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, winner);
    DYG_LAUNCH_CHECK("dyg_tgn_select_last");
    return 0;
}

__global__ void tgn_build_messages_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                          const double* __restrict__ t, const int64_t* __restrict__ eid, int64_t B,
                                          const float* __restrict__ memory, const float* __restrict__ last_update, int D,
                                          const float* __restrict__ other_emb, int ld_other,
                                          const float* __restrict__ edge_tab, int ld_edge, int E,
                                          const float* __restrict__ w, const float* __restrict__ b, int T,
                                          float* __restrict__ msg, int ldm) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= 2 * B) return;
    const int64_t ev = c < B ? c : c - B;
    const int64_t owner = c < B ? src[ev] : dst[ev];
    const int64_t other = c < B ? dst[ev] : src[ev];
    float* o = msg + c * ldm;
    for (int j = lane; j < D; j += 32) o[j] = memory[owner * D + j];
    const float* op = other_emb ? other_emb + c * ld_other : memory + other * D;
    for (int j = lane; j < D; j += 32) o[D + j] = op[j];
    // float32(t) - float32 last update (models/MemoryModel.py:232-233)
    const float dt = (float)t[ev] - last_update[owner];
    for (int j = lane; j < T; j += 32) o[2 * D + j] = dyg_time_enc(dt, __ldg(w + j), __ldg(b + j));
    const float* ep = edge_tab + eid[ev] * ld_edge;
    for (int j = lane; j < E; j += 32) o[2 * D + T + j] = __ldg(ep + j);
}
extern "C" int dyg_tgn_build_messages(const int64_t* src, const int64_t* dst, const double* t, const int64_t* eid,
                                      int64_t B, const float* memory, const float* last_update, int D,
                                      const float* other_emb, int ld_other, const float* edge_tab, int ld_edge, int E,
                                      const float* w, const float* b, int T, float* msg, int ldm, dyg_stream_t stream) {
    DYG_CHECK_ARG(B >= 0 && D > 0 && E >= 0 && T >= 0 && ldm >= 2 * D + T + E, "dyg_tgn_build_messages: bad sizes");
    if (B == 0) return 0;
    tgn_build_messages_kernel<<<(unsigned)((2 * B * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        src, dst, t, eid, B, memory, last_update, D, other_emb, ld_other, edge_tab, ld_edge, E, w, b, T, msg, ldm);
    DYG_LAUNCH_CHECK("dyg_tgn_build_messages");
    return 0;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void tgn_cell_commit_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int G,
                                       const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                       const double* __restrict__ t, int64_t B, int32_t* __restrict__ winner,
                                       const float* __restrict__ memory, float* __restrict__ mem_view,
                                       float* __restrict__ lu_view, uint8_t* __restrict__ pending, int D,
                                       const float* __restrict__ msg, int ldm, int msg_dim,
                                       float* __restrict__ msg_store, double* __restrict__ msg_time) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= 2 * B) return;
    const int64_t ev = c < B ? c : c - B;
    const int64_t v = c < B ? src[ev] : dst[ev];
    if (winner[v] != (int32_t)c) return;  // warp-uniform
    const float* a = gi + c * (int64_t)(G * D);
    const float* hgt = gh + c * (int64_t)(G * D);
    for (int j = lane; gi && j < D; j += 32) {   // gi == NULL: dyg_gru_update_fwd already wrote mem_view[v]
        float hn;
        if (G == 3) {  // nn.GRUCell gate order r, z, n
            const float r = sigmoidf_(a[j] + hgt[j]);
            const float z = sigmoidf_(a[D + j] + hgt[D + j]);
            const float ng = tanhf(a[2 * D + j] + r * hgt[2 * D + j]);
            const float h = memory[v * D + j];
            hn = ng + z * (h - ng);
        } else {  // nn.RNNCell (tanh)
            hn = tanhf(a[j] + hgt[j]);
        }
        mem_view[v * D + j] = hn;
    }
    if (msg_store)
        for (int j = lane; j < msg_dim; j += 32) msg_store[v * (int64_t)msg_dim + j] = msg[c * ldm + j];
    __syncwarp();
    if (lane == 0) {
        lu_view[v] = (float)t[ev];
        if (msg_time) msg_time[v] = t[ev];
        pending[v] = 1;
        winner[v] = -1;  // leave the scratch table clean for the next batch
    }
}
extern "C" int dyg_tgn_cell_commit(const float* gi, const float* gh, int G, const int64_t* src, const int64_t* dst,
                                   const double* t, int64_t B, int32_t* winner, const float* memory, float* mem_view,
                                   float* lu_view, uint8_t* pending, int D, const float* msg, int ldm, int msg_dim,
                                   float* msg_store, double* msg_time, dyg_stream_t stream) {
    DYG_CHECK_ARG(G == 1 || G == 3, "dyg_tgn_cell_commit: G must be 3 (GRU) or 1 (RNN)");
    DYG_CHECK_ARG((gi == nullptr) == (gh == nullptr), "dyg_tgn_cell_commit: gi and gh must both be given or both be NULL");
    DYG_CHECK_ARG(B >= 0 && D > 0, "dyg_tgn_cell_commit: bad sizes");
    if (B == 0) return 0;
    tgn_cell_commit_kernel<<<(unsigned)((2 * B * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        gi, gh, G, src, dst, t, B, winner, memory, mem_view, lu_view, pending, D, msg, ldm, msg_dim, msg_store, msg_time);
    DYG_LAUNCH_CHECK("dyg_tgn_cell_commit");
    return 0;
}

__global__ void tgn_check_time_kernel(const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ last_update,
                                      const float* __restrict__ lu_view, const uint8_t* __restrict__ pending,
                                      int32_t* __restrict__ flag) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t v = ids[e];
    if (pending[v] && last_update[v] > lu_view[v]) atomicExch(flag, 1);
}
extern "C" int dyg_tgn_check_time(const int64_t* node_ids, int64_t n, const float* last_update, const float* lu_view,
                                  const uint8_t* pending, int32_t* flag, dyg_stream_t stream) {
    if (n <= 0) return 0;
    tgn_check_time_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(node_ids, n, last_update, lu_view, pending, flag);
    DYG_LAUNCH_CHECK("dyg_tgn_check_time");
    return 0;
}
