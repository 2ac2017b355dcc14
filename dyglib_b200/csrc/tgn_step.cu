// dyg_tgn_step: one 200-event batch of the TGN memory model in ONE cooperative launch.
//
// The reference's per-batch body (models/MemoryModel.py:87-168: look-ahead memories, 1-layer graph attention over 10 recent
// neighbours, MergeLayer, then update_memories / clear / new raw messages / store for the positive batch, and the link
// predictor of train_link_prediction.py:243-244) moves ~1 GFLOP and ~20 MB through a chain of dependent steps.  As separate
// kernels that chain was 25-29 launches at ~10 us each (round 1: 269 us per step, 1 % of its HBM time).  Here a persistent grid
// (one CTA per SM, cooperative launch) walks the chain as PHASES separated by a grid barrier; inside a phase every CTA takes
// tile tasks (BF16x3 mma.sync tiles of mma_tile.cuh, teams of 256 threads, two teams per CTA) and warp tasks (one warp per root /
// candidate message) from a static list.  Every operand of a dense phase is written as BF16x3 planes by the phase that produces it,
// so a tile stage is plain 16-byte cp.async copies.  The embedding chain and the memory-update chain of the same batch are independent until the commit, so they
// share phases:
//   P0  roots: lower bound on the CSR + recent-neighbour gather + layer-0 features (memory view + raw)
//       candidates: time-order check, persist the look-ahead memories of the batch's nodes, last-message election
//   P1  qk = feat Wqk^T + cq                               | candidates: build the 616-wide raw messages, clear pending
//   P2  roots: folded temporal attention (gather + time encoding + softmax + sum)   | GRU / RNN cell tiles of the candidates
//   P3  o = s Wvr^T + b                                    | winners: commit new look-ahead memories + message store
//   P4  y = LayerNorm(o + [feat | cos(b)])
//   P5  h = relu([y | feat] W1^T + b1)
//   P6  emb = h W2^T + b2                                  | ph = relu([h[a] | h[b]] (Wp1 W2)^T + const)   (link predictor folded onto h)
//   P8  prob = sigmoid(ph . wp2 + bp2)
// Reads of the look-ahead view (P0, P2) precede its update (P3); every scratch buffer is written in one phase and read in a
// later one through L2 (cp.async.cg / ld.global.cg), so no SM can hold a stale L1 line of it.
// History (profiles/r02_tgn_step.md): fp32 FFMA tiles, one CTA per SM: 164 us per step (shared-memory bound: every LDS.128 is four
// wavefronts; "compute only" 85 % of a phase); this version: see the phase table there.
#include <cooperative_groups.h>
#include <math.h>
#include "mma_tile.cuh"

namespace {

typedef dyg_tgn_step_t P;

__device__ __forceinline__ void stamp(const P& p, int slot) {
    if (p.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.phase_ns[slot] = t;
    }
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {   // relaxed polling (an acquire load per poll invalidates the L1 every time: CCTL.IVALL); one fence after the exit
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ mt::Planes planes(const dyg_planes_t& d) {
    return mt::Planes{reinterpret_cast<const mt::bf16*>(d.hi), reinterpret_cast<const mt::bf16*>(d.mid), d.ld};
}
__device__ __forceinline__ mt::bf16* hi_of(const dyg_planes_t& d) { return reinterpret_cast<mt::bf16*>(const_cast<void*>(d.hi)); }
__device__ __forceinline__ mt::bf16* mid_of(const dyg_planes_t& d) { return reinterpret_cast<mt::bf16*>(const_cast<void*>(d.mid)); }
// four consecutive values at columns c .. c + 3 (c % 4 == 0) of row `row`
__device__ __forceinline__ void store_planes4(const dyg_planes_t& d, int64_t row, int c, float4 v) {
    uint32_t h0, m0, h1, m1;
    mt::split2(v.x, v.y, h0, m0);
    mt::split2(v.z, v.w, h1, m1);
    *reinterpret_cast<uint2*>(hi_of(d) + row * d.ld + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(mid_of(d) + row * d.ld + c) = make_uint2(m0, m1);
}
// one value per lane at column c0 + lane (c0 even; `valid`: this lane's column exists): even lanes store the pair (own, next lane's)
__device__ __forceinline__ void store_planes_lane(const dyg_planes_t& d, int64_t row, int c0, int lane, float v, bool valid) {
    const float nxt = __shfl_down_sync(0xffffffffu, v, 1);
    const bool nvalid = __shfl_down_sync(0xffffffffu, (int)valid, 1) != 0;
    if (valid && !(lane & 1)) mt::store_split2(hi_of(d), mid_of(d), d.ld, row, c0 + lane, v, (lane < 31 && nvalid) ? nxt : 0.f);
}

// rows with t < tq among the deg records from a: the 32 lanes probe 32 evenly spaced records per round (33-ary search)
__device__ __forceinline__ int64_t warp_lower_bound(const dyg_halfedge_t* __restrict__ he, int64_t a, int64_t deg, double tq, int lane) {
    int64_t lo = 0, hi = deg;
    while (hi - lo > 32) {
        const int64_t len = hi - lo;
        const int64_t p = lo + ((int64_t)(lane + 1) * len) / 33;
        const bool pred = __ldg(reinterpret_cast<const double*>(he + a + p)) < tq;
        const int c = __popc(__ballot_sync(0xffffffffu, pred));
        const int64_t nlo = c > 0 ? lo + ((int64_t)c * len) / 33 + 1 : lo;
        const int64_t nhi = c < 32 ? lo + ((int64_t)(c + 1) * len) / 33 : hi;
        lo = nlo;
        hi = nhi;
    }
    const int64_t q = lo + lane;
    const bool pred = q < hi && __ldg(reinterpret_cast<const double*>(he + a + q)) < tq;
    return lo + __popc(__ballot_sync(0xffffffffu, pred));
}

// ------------------------------------------------------------------------------------------------ warp tasks
// roots == NULL: the batch's own node lists are the roots ([src | neg | dst], or [src | dst] without negatives) -- no host-side concatenation
__device__ __forceinline__ int64_t root_id(const P& p, int r) {
    if (p.roots) return p.roots[r];
    if (r < p.B) return p.src[r];
    if (p.neg) return r < 2 * p.B ? p.neg[r - p.B] : p.dst[r - 2 * p.B];
    return p.dst[r - p.B];
}
__device__ __forceinline__ void root_task(const P& p, int r, int lane) {
    const int64_t v = root_id(p, r);
    const double tq = p.t[r % p.B];
    int64_t a = 0, deg = 0;
    if (v >= 0 && v < p.num_nodes) {
        a = __ldg(p.indptr + v);
        deg = __ldg(p.indptr + v + 1) - a;
    }
    const int64_t cnt = warp_lower_bound(p.he, a, deg, tq, lane);
    const int k = p.k;
    const int pad = cnt < k ? (int)(k - cnt) : 0;
    for (int j = lane; j < k; j += 32) {
        int64_t nb = 0, ei = 0;
        float tf = 0.f;
        if (j >= pad) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(p.he + (a + cnt - k + j)));
            nb = rec.z;
            ei = rec.w;
            tf = (float)__hiloint2double(rec.y, rec.x);
        }
        p.nbr_ids[(int64_t)r * k + j] = nb;
        p.nbr_eids[(int64_t)r * k + j] = ei;
        p.nbr_t[(int64_t)r * k + j] = tf;
    }
    // layer-0 features: look-ahead memory + raw features (models/MemoryModel.py:609)
    const float* raw = p.node_raw + v * p.ld_node;
    const float* mem = p.mem_view + v * p.F;
    float* out = p.feat + (int64_t)r * p.F;
    for (int c = lane * 4; c < p.F; c += 128) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(raw + c));
        const float4 m = *reinterpret_cast<const float4*>(mem + c);
        const float4 f = make_float4(x.x + m.x, x.y + m.y, x.z + m.z, x.w + m.w);
        *reinterpret_cast<float4*>(out + c) = f;
        store_planes4(p.feat_pl, r, c, f);
    }
}

__device__ __forceinline__ int64_t cand_node(const P& p, int c) { return c < p.B ? p.src[c] : p.dst[c - p.B]; }

// update_memories for the batch's nodes = persist their look-ahead rows (models/MemoryModel.py:142, 435-459); last-message election
__device__ __forceinline__ void persist_task(const P& p, int c, int lane) {
    const int64_t v = cand_node(p, c);
    if (p.pending[v]) {   // cleared in P1, so every duplicate of v sees it and copies identical values
        if (p.check_time && lane == 0 && p.last_update[v] > p.lu_view[v]) atomicExch(p.flag, 1);
        for (int j = lane * 4; j < p.F; j += 128)
            *reinterpret_cast<float4*>(p.memory + v * p.F + j) = *reinterpret_cast<const float4*>(p.mem_view + v * p.F + j);
        if (lane == 0) p.last_update[v] = p.lu_view[v];
    }
    if (lane == 0) atomicMax(p.winner + v, (int32_t)c);
}

// compute_new_node_raw_messages (models/MemoryModel.py:212-251): [memory[owner] | memory[other] | time_enc(t - last_update[owner]) | edge]
__device__ __forceinline__ void message_task(const P& p, int c, int lane) {
    const int ev = c < p.B ? c : c - p.B;
    const int64_t owner = c < p.B ? p.src[ev] : p.dst[ev];
    const int64_t other = c < p.B ? p.dst[ev] : p.src[ev];
    const int D = p.F;
    float* o = p.msg + (int64_t)c * (2 * D + p.T + p.E);
    for (int j = lane * 4; j < D; j += 128) {
        const float4 a = ldcg4(p.memory + owner * D + j), b = ldcg4(p.memory + other * D + j);
        *reinterpret_cast<float4*>(o + j) = a;
        *reinterpret_cast<float4*>(o + D + j) = b;
        store_planes4(p.msg_pl, c, j, a);
        store_planes4(p.msg_pl, c, D + j, b);
    }
    const float dt = (float)p.t[ev] - __ldcg(p.last_update + owner);
    for (int j0 = 0; j0 < p.T; j0 += 32) {
        const int j = j0 + lane;
        const float v = j < p.T ? dyg_time_enc(dt, __ldg(p.time_w + j), __ldg(p.time_b + j)) : 0.f;
        if (j < p.T) o[2 * D + j] = v;
        store_planes_lane(p.msg_pl, c, 2 * D + j0, lane, v, j < p.T);
    }
    const float* ep = p.edge_raw + p.eid[ev] * p.ld_edge;
    for (int j = lane * 4; j < p.E; j += 128) {
        const float4 e4 = __ldg(reinterpret_cast<const float4*>(ep + j));
        *reinterpret_cast<float4*>(o + 2 * D + p.T + j) = e4;
        store_planes4(p.msg_pl, c, 2 * D + p.T + j, e4);
    }
    if (lane == 0) p.pending[owner] = 0;   // clear_node_raw_messages (:145); winners are set again by the commit
}

__device__ __forceinline__ void commit_task(const P& p, int c, int lane) {
    const int ev = c < p.B ? c : c - p.B;
    const int64_t v = cand_node(p, c);
    if (__ldcg(p.winner + v) != (int32_t)c) return;   // warp-uniform
    const int D = p.F, MD = 2 * D + p.T + p.E;
    for (int j = lane * 4; j < D; j += 128) *reinterpret_cast<float4*>(p.mem_view + v * D + j) = ldcg4(p.hnew + (int64_t)c * D + j);
    for (int j = lane * 4; j < MD; j += 128) *reinterpret_cast<float4*>(p.msg_store + v * MD + j) = ldcg4(p.msg + (int64_t)c * MD + j);
    __syncwarp();
    if (lane == 0) {
        p.lu_view[v] = (float)p.t[ev];
        p.msg_time[v] = p.t[ev];
        p.pending[v] = 1;
        p.winner[v] = -1;   // leave the election table clean for the next batch
    }
}

// Folded temporal attention of one root (models/modules.py:157-193 in the folded form of dyg_temporal_attend): the lane mapping of
// temporal_attend_split_kernel (node | edge part as float4 chunks lane, lane+32, lane+64; time features lane + 32 r).
template <int H>
__device__ __forceinline__ void attend_task(const P& p, int r, int lane) {
    const int F4 = p.F / 4, NE4 = (p.F + p.E) / 4, T = p.T, k = p.k;
    const int Dk = NE4 * 4 + T;
    float4 q[H][3], acc[H][3];
    float qt[H][4], acct[H][4];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float* qrow = p.qk + (int64_t)r * (H * Dk) + h * Dk;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int c = i * 32 + lane;
            q[h][i] = c < NE4 ? ldcg4(qrow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            qt[h][i] = c < T ? __ldcg(qrow + NE4 * 4 + c) : 0.f;
            acct[h][i] = 0.f;
        }
    }
    float tw[4], tb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 32 + lane;
        tw[i] = c < T ? __ldg(p.time_w + c) : 0.f;
        tb[i] = c < T ? __ldg(p.time_b + c) : 0.f;
    }
    const double tq = p.t[r % p.B];
    float mx[H], den[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        mx[h] = -INFINITY;
        den[h] = 0.f;
    }
    const int64_t base = (int64_t)r * k;
    for (int j0 = 0; j0 < k; j0 += 32) {
        const int jl = j0 + lane;
        int64_t my_n = 0, my_e = 0;
        float my_dt = 0.f;
        if (jl < k) {
            my_n = __ldcg(p.nbr_ids + base + jl);
            my_e = __ldcg(p.nbr_eids + base + jl);
            my_dt = (float)(tq - (double)__ldcg(p.nbr_t + base + jl));
        }
        const int jn = (k - j0) < 32 ? (k - j0) : 32;
        auto load_x = [&](int jj, float4(&x)[3]) {
            const int64_t rn = __shfl_sync(0xffffffffu, my_n, jj);
            const int64_t re = __shfl_sync(0xffffffffu, my_e, jj);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int c = i * 32 + lane;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < F4) {
                    v = __ldg(reinterpret_cast<const float4*>(p.node_raw + rn * p.ld_node) + c);
                    const float4 u = *(reinterpret_cast<const float4*>(p.mem_view + rn * p.F) + c);
                    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                } else if (c < NE4) {
                    v = __ldg(reinterpret_cast<const float4*>(p.edge_raw + re * p.ld_edge) + (c - F4));
                }
                x[i] = v;
            }
        };
        float4 xn[3];
        load_x(0, xn);
        for (int jj = 0; jj < jn; ++jj) {
            float4 x[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) x[i] = xn[i];
            if (jj + 1 < jn) load_x(jj + 1, xn);
            const float dt = __shfl_sync(0xffffffffu, my_dt, jj);
            const int masked = __shfl_sync(0xffffffffu, my_n, jj) == 0;
            float xt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xt[i] = (i * 32 + lane < T) ? dyg_time_enc(dt, tw[i], tb[i]) : 0.f;
            float s[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    d = fmaf(q[h][i].x, x[i].x, d);
                    d = fmaf(q[h][i].y, x[i].y, d);
                    d = fmaf(q[h][i].z, x[i].z, d);
                    d = fmaf(q[h][i].w, x[i].w, d);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) d = fmaf(qt[h][i], xt[i], d);
                s[h] = d;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int h = 0; h < H; ++h) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float sc = masked ? -1e10f : s[h];   // -1e10, not -inf (models/modules.py:184)
                if (sc > mx[h]) {
                    const float corr = expf(mx[h] - sc);
                    den[h] *= corr;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        acc[h][i].x *= corr; acc[h][i].y *= corr; acc[h][i].z *= corr; acc[h][i].w *= corr;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) acct[h][i] *= corr;
                    mx[h] = sc;
                }
                const float w = expf(sc - mx[h]);
                den[h] += w;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc[h][i].x = fmaf(w, x[i].x, acc[h][i].x);
                    acc[h][i].y = fmaf(w, x[i].y, acc[h][i].y);
                    acc[h][i].z = fmaf(w, x[i].z, acc[h][i].z);
                    acc[h][i].w = fmaf(w, x[i].w, acc[h][i].w);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) acct[h][i] = fmaf(w, xt[i], acct[h][i]);
            }
        }
    }
    // s[r, h, :] as operand planes of the next contraction (o = s Wvr^T)
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.f / den[h];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int c = i * 32 + lane;
            if (c < NE4) {
                float4 v = acc[h][i];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                store_planes4(p.s_pl, r, h * Dk + 4 * c, v);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) store_planes_lane(p.s_pl, r, h * Dk + NE4 * 4 + i * 32, lane, acct[h][i] * inv, i * 32 + lane < T);
    }
}

// y = LayerNorm(o + [feat | cos(b)]) (models/modules.py:199 with the residual of :155,:197); Dq = F + T <= 512.  Every load of the row
// (o, residual, gamma, beta) is issued before the first use: a first version that fetched gamma / beta inside the store loop spent
// ~6 us per row on nine exposed L2 round trips.
__device__ __forceinline__ void layernorm_task(const P& p, int r, int lane) {
    const int Dq = p.F + p.T;
    float v[16], gm[16], bt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = i * 32 + lane, cc = c < Dq ? c : Dq - 1;
        v[i] = gm[i] = bt[i] = 0.f;
        if (i * 32 < Dq) {   // warp-uniform
            const float* rp = cc < p.F ? p.feat + (int64_t)r * p.F + cc : p.t0 + (cc - p.F);
            const float a = __ldcg(p.o + (int64_t)r * Dq + cc), b = __ldcg(rp);
            gm[i] = __ldg(p.ln_g + cc);
            bt[i] = __ldg(p.ln_b + cc);
            v[i] = c < Dq ? a + b : 0.f;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += v[i];
    const float mean = warp_sum(sum) / Dq;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float d = (i * 32 + lane < Dq) ? v[i] - mean : 0.f;
        var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(warp_sum(var) / Dq + p.ln_eps);
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i * 32 < Dq) store_planes_lane(p.y_pl, r, i * 32, lane, (v[i] - mean) * rstd * gm[i] + bt[i], i * 32 + lane < Dq);
}

__device__ __forceinline__ void score_task(const P& p, int i, int lane) {
    float d = 0.f;
    for (int c = lane; c < p.F; c += 32) d = fmaf(__ldcg(p.ph + (int64_t)i * p.F + c), __ldg(p.p2_w + c), d);
    d = warp_sum(d);
    if (lane == 0) p.prob[i] = mt::sigmoidf_(d + __ldg(p.p2_b));
}

// ------------------------------------------------------------------------------------------------ tile tasks
struct Gemm {
    mt::ASeg seg[2];
    int wcol[2];          // first column of each segment inside the (segment-padded) weight planes
    int nseg;
    mt::Planes W;
    const float* bias;
    int relu;
    float* C;             // fp32 output (may be NULL)
    int64_t ldc;
    dyg_planes_t Cp;      // plane output (hi NULL: none)
    int64_t M;
    int N;
};

// 8 warps as 2 x 4; MI = 1: 32 x 64 tiles, MI = 2: 64 x 64 tiles
template <int MI>
__device__ __forceinline__ void gemm_tile(const Gemm& g, int tile, mt::bf16* smem, int t, int bar, unsigned long long* dbg = nullptr) {
    constexpr int WM = 2, WN = 4, NI = 2, BM = WM * MI * 16, BN = WN * NI * 8;
    auto tick = [&](int slot) {
        if (dbg && t == 0) {
            unsigned long long x;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(x));
            dbg[slot] = x;
        }
    };
    tick(9);
    const int ntn = (g.N + BN - 1) / BN;
    const int64_t m0 = (int64_t)(tile / ntn) * BM;
    const int n0 = (tile % ntn) * BN;
    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const mt::WRows wmap{n0, g.N};
    for (int sidx = 0; sidx < g.nseg; ++sidx) mt::gemm_accum<WM, WN, MI, NI>(acc, g.seg[sidx], m0, g.M, g.W, g.wcol[sidx], wmap, smem, t, bar);
    tick(10);
    const int warp = t >> 5, lane = t & 31, wm = warp / WN, wn = warp % WN, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        const int n = n0 + wn * NI * 8 + j * 8 + 2 * tq;
        if (n >= g.N) continue;      // N is even: n + 1 < N too
        const float b0 = g.bias ? __ldg(g.bias + n) : 0.f, b1 = g.bias ? __ldg(g.bias + n + 1) : 0.f;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int64_t m = m0 + wm * MI * 16 + i * 16 + gq + 8 * half;
                if (m >= g.M) continue;
                float v0 = acc[i][j][2 * half] + b0, v1 = acc[i][j][2 * half + 1] + b1;
                if (g.relu) {
                    v0 = fmaxf(v0, 0.f);
                    v1 = fmaxf(v1, 0.f);
                }
                if (g.C) *reinterpret_cast<float2*>(g.C + m * g.ldc + n) = make_float2(v0, v1);
                if (g.Cp.hi) mt::store_split2(hi_of(g.Cp), mid_of(g.Cp), g.Cp.ld, m, n, v0, v1);
            }
    }
    tick(11);
}

// recurrent cell of the candidates (nn.GRUCell / nn.RNNCell, models/MemoryModel.py:490-515): 64 candidates x 16 hidden units per tile,
// 8 warps as 4 x 2; the NI = G column blocks of a warp are the G gates of the same 8 units (mt::WGates)
template <int G>
__device__ __forceinline__ void cell_tile(const P& p, int tile, mt::bf16* smem, int t, int bar) {
    constexpr int WM = 4, WN = 2, MI = 1, NI = G;
    const int D = p.F, MD = 2 * D + p.T + p.E, Fp = (D + 7) & ~7, nu = (D + 15) / 16;
    const int64_t m0 = (int64_t)(tile / nu) * (WM * 16);
    const int u0 = (tile % nu) * 16;
    const int64_t Pn = 2 * (int64_t)p.B;
    float acc[MI][NI][4];
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[0][j][0] = acc[0][j][1] = acc[0][j][2] = acc[0][j][3] = 0.f;
    const mt::WGates wmap{u0, D, G};
    const mt::Planes msg = planes(p.msg_pl);
    mt::gemm_accum<WM, WN, MI, NI>(acc, mt::ASeg{msg, nullptr, 0, MD}, m0, Pn, planes(p.w_ih), 0, wmap, smem, t, bar);
    float in_n[4];   // GRU: the candidate gate keeps its input and hidden halves apart (n = tanh(i_n + r * h_n))
    if (G == 3) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            in_n[e] = acc[0][G - 1][e];
            acc[0][G - 1][e] = 0.f;
        }
    }
    // the hidden state of candidate c is memory[owner(c)] = the first F columns of its message (columns F .. Fp - 1 meet zero weights)
    mt::gemm_accum<WM, WN, MI, NI>(acc, mt::ASeg{msg, nullptr, 0, Fp}, m0, Pn, planes(p.w_hh), 0, wmap, smem, t, bar);
    const int warp = t >> 5, lane = t & 31, wm = warp / WN, wn = warp % WN, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t m = m0 + wm * 16 + gq + 8 * half;
        if (m >= Pn) continue;
        const int64_t v = cand_node(p, (int)m);
        if (__ldcg(p.winner + v) != (int32_t)m) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int u = u0 + wn * 8 + 2 * tq + e;
            if (u >= D) continue;
            const int ci = 2 * half + e;
            float hn;
            if (G == 3) {
                const float r = mt::sigmoidf_(acc[0][0][ci] + p.b_ih[u] + p.b_hh[u]);
                const float z = mt::sigmoidf_(acc[0][1][ci] + p.b_ih[D + u] + p.b_hh[D + u]);
                const float ng = tanhf(in_n[ci] + p.b_ih[2 * D + u] + r * (acc[0][G - 1][ci] + p.b_hh[2 * D + u]));
                const float h = __ldcg(p.memory + v * D + u);
                hn = ng + z * (h - ng);
            } else {
                hn = tanhf(acc[0][0][ci] + p.b_ih[u] + p.b_hh[u]);
            }
            p.hnew[m * D + u] = hn;
        }
    }
}

struct Plan {   // tile heights chosen on the host for the actual R / B (MI = 1: 32-row tiles, 2: 64-row tiles)
    int mi_qk, mi_o, mi_m1, mi_m2, mi_p1;
};

constexpr int TEAMS = 2;                       // tile teams (256 threads each) per CTA
constexpr int CTA_THREADS = TEAMS * mt::THREADS;
constexpr int TEAM_SMEM_BYTES = mt::Tile<64, 64>::SMEM_BYTES;

struct Team {
    int id, n;        // global team index, number of teams in the grid
    int t, bar;       // thread index inside the team, its named barrier
    mt::bf16* smem;
};

// warp tasks [0, ntasks) are handed out in groups of 8 (one per warp of a team); `task` counts groups from `first_task`
template <class F>
__device__ __forceinline__ void warp_tasks(const Team& tm, int first_task, int ntasks, int task, F f) {
    const int w = (task - first_task) * (mt::THREADS / 32) + (tm.t >> 5);
    if (w < ntasks) f(w, tm.t & 31);
}

__device__ __forceinline__ int tiles_of(const Gemm& g, int mi) { return (int)((g.M + 32 * mi - 1) / (32 * mi)) * ((g.N + 63) / 64); }
__device__ __forceinline__ void run_tile(const Gemm& g, int mi, int tile, const Team& team, unsigned long long* dbg = nullptr) {
    if (mi == 1) gemm_tile<1>(g, tile, team.smem, team.t, team.bar, dbg);
    else gemm_tile<2>(g, tile, team.smem, team.t, team.bar, dbg);
}

template <int H, int G>
__global__ void __launch_bounds__(CTA_THREADS, 1) tgn_step_kernel(const __grid_constant__ P p, const Plan plan) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    unsigned target = 0;
    Team team;
    const int local = threadIdx.x / mt::THREADS;
    team.id = local * gridDim.x + blockIdx.x;     // consecutive tasks land on different SMs (a phase with <= 148 tasks uses one team per SM:
    team.n = gridDim.x * TEAMS;                   // the L2 -> SM fill rate of an SM, not the tensor pipe, bounds a tile)
    team.t = threadIdx.x % mt::THREADS;
    team.bar = 1 + local;
    team.smem = reinterpret_cast<mt::bf16*>(smem_all + local * TEAM_SMEM_BYTES);
    constexpr int WPT = mt::THREADS / 32;
    const int R = p.R, C2 = 2 * p.B;
    const int Dk = p.F + p.E + p.T, Dq = p.F + p.T, Fp = (p.F + 7) & ~7;
    auto groups = [](int n) { return (n + WPT - 1) / WPT; };
    const dyg_planes_t none{nullptr, nullptr, 0};

    stamp(p, 0);
    // ---- P0
    {
        const int g0 = groups(R), g1 = groups(C2);
        for (int task = team.id; task < g0 + g1; task += team.n) {
            if (task < g0) warp_tasks(team, 0, R, task, [&](int w, int lane) { root_task(p, w, lane); });
            else warp_tasks(team, g0, C2, task, [&](int w, int lane) { persist_task(p, w, lane); });
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 1);
    // ---- P1
    {
        const Gemm g{{mt::ASeg{planes(p.feat_pl), nullptr, 0, Fp}, {}}, {0, 0}, 1, planes(p.wqk), p.cq, 0, p.qk, H * Dk, none, R, H * Dk};
        const int nt = tiles_of(g, plan.mi_qk), g1 = groups(C2);
        for (int task = team.id; task < nt + g1; task += team.n) {
            if (task < nt) run_tile(g, plan.mi_qk, task, team);
            else warp_tasks(team, nt, C2, task, [&](int w, int lane) { message_task(p, w, lane); });
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 2);
    // ---- P2: attention groups first (longer tasks), then the cell tiles
    {
        const int g1 = groups(R), nt = ((C2 + 63) / 64) * ((p.F + 15) / 16);
        for (int task = team.id; task < g1 + nt; task += team.n) {
            if (task < g1) warp_tasks(team, 0, R, task, [&](int w, int lane) { attend_task<H>(p, w, lane); });
            else cell_tile<G>(p, task - g1, team.smem, team.t, team.bar);
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 3);
    // ---- P3
    {
        const Gemm g{{mt::ASeg{planes(p.s_pl), nullptr, 0, H * Dk}, {}}, {0, 0}, 1, planes(p.wvr), p.rbias, 0, p.o, Dq, none, R, Dq};
        const int nt = tiles_of(g, plan.mi_o), g1 = groups(C2);
        for (int task = team.id; task < nt + g1; task += team.n) {
            if (task < nt) run_tile(g, plan.mi_o, task, team, (p.phase_ns && team.id == 0) ? p.phase_ns : nullptr);
            else warp_tasks(team, nt, C2, task, [&](int w, int lane) { commit_task(p, w, lane); });
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 4);
    // ---- P4
    for (int task = team.id; task < groups(R); task += team.n) warp_tasks(team, 0, R, task, [&](int w, int lane) { layernorm_task(p, w, lane); });
    grid_barrier(p.barrier, target);
    stamp(p, 5);
    // ---- P5
    {
        const Gemm g{{mt::ASeg{planes(p.y_pl), nullptr, 0, Dq}, mt::ASeg{planes(p.feat_pl), nullptr, 0, Fp}}, {0, Dq}, 2, planes(p.m1), p.m1_b, 1,
                     nullptr, 0, p.h1_pl, R, p.F};
        for (int task = team.id; task < tiles_of(g, plan.mi_m1); task += team.n) run_tile(g, plan.mi_m1, task, team);
    }
    grid_barrier(p.barrier, target);
    stamp(p, 6);
    // ---- P6: emb = h W2^T + b2, and the link predictor's first layer on the SAME input: fc1([emb_a | emb_b]) with emb = W2 h + b2 is
    //      (Wp1a W2) h_a + (Wp1b W2) h_b + const, folded on the host like the attention weights (exact algebra, different rounding)
    {
        const Gemm g{{mt::ASeg{planes(p.h1_pl), nullptr, 0, Fp}, {}}, {0, 0}, 1, planes(p.m2), p.m2_b, 0, p.emb, p.F, none, R, p.F};
        const mt::Planes h1 = planes(p.h1_pl);
        const Gemm gp{{mt::ASeg{h1, p.pair_a, 0, Fp}, mt::ASeg{h1, p.pair_b, 0, Fp}}, {0, Fp}, 2, planes(p.p1), p.p1_b, 1, p.ph, p.F, none, p.P, p.F};
        const int nt = tiles_of(g, plan.mi_m2), ntp = p.p1.hi ? tiles_of(gp, plan.mi_p1) : 0;
        for (int task = team.id; task < nt + ntp; task += team.n) {
            if (task < nt) run_tile(g, plan.mi_m2, task, team);
            else run_tile(gp, plan.mi_p1, task - nt, team);
        }
    }
    if (p.p1.hi) {
        grid_barrier(p.barrier, target);
        stamp(p, 7);
        stamp(p, 8);
        // ---- P8
        for (int task = team.id; task < groups(p.P); task += team.n) warp_tasks(team, 0, p.P, task, [&](int w, int lane) { score_task(p, w, lane); });
    }
    stamp(p, 15);
    // the last CTA out re-arms the barrier for the next launch (every CTA has passed every barrier before it counts itself out)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.barrier + 1, 1u) == (unsigned)gridDim.x - 1) {
            p.barrier[0] = 0;
            p.barrier[1] = 0;
        }
    }
}

int pick_mi(int64_t M, int N, int teams) {
    int best = 1;
    double cost = 1e30;
    for (int mi : {1, 2}) {
        const int64_t tiles = ((M + 32 * mi - 1) / (32 * mi)) * ((N + 63) / 64);
        const double c = (double)((tiles + teams - 1) / teams) * (mi + 1.0);   // waves x (A + W bytes per tile)
        if (c < cost) {
            cost = c;
            best = mi;
        }
    }
    return best;
}

}  // namespace

extern "C" int64_t dyg_tgn_step_sizeof(void) { return (int64_t)sizeof(dyg_tgn_step_t); }

extern "C" int dyg_tgn_step(const dyg_tgn_step_t* ph, dyg_stream_t stream) {
    DYG_CHECK_ARG(ph != nullptr, "dyg_tgn_step: null parameter block");
    const P& p = *ph;
    DYG_CHECK_ARG(p.B > 0 && p.R > 0 && p.k > 0, "dyg_tgn_step: empty batch");
    DYG_CHECK_ARG(p.H == 2 && (p.G == 1 || p.G == 3), "dyg_tgn_step: H must be 2, G 1 (RNN) or 3 (GRU)");
    DYG_CHECK_ARG(p.roots || p.R == (p.neg ? 3 : 2) * p.B, "dyg_tgn_step: without a root list R must be 3 B ([src | neg | dst]) or 2 B ([src | dst])");
    const int Dk = p.F + p.E + p.T, Dq = p.F + p.T, MD = 2 * p.F + p.T + p.E, Fp = (p.F + 7) & ~7;
    DYG_CHECK_ARG(p.F % 4 == 0 && p.E % 4 == 0 && p.T % 4 == 0 && p.ld_node % 4 == 0 && p.ld_edge % 4 == 0 && Dq % 8 == 0 && MD % 8 == 0 &&
                      (p.H * Dk) % 8 == 0,
                  "dyg_tgn_step: F, E, T must be multiples of 4 and F + T, 2F + T + E, H (F + E + T) multiples of 8");
    DYG_CHECK_ARG(p.F + p.E <= 384 && p.T <= 128 && Dq <= 512, "dyg_tgn_step: feature widths out of range");
    DYG_CHECK_ARG(p.he && p.indptr && p.src && p.dst && p.t && p.eid && p.node_raw && p.edge_raw && p.memory &&
                      p.last_update && p.mem_view && p.lu_view && p.pending && p.winner && p.msg_store && p.msg_time && p.flag && p.barrier,
                  "dyg_tgn_step: null state pointer");
    DYG_CHECK_ARG(p.nbr_ids && p.nbr_eids && p.nbr_t && p.feat && p.qk && p.o && p.msg && p.hnew && p.emb, "dyg_tgn_step: null scratch pointer");
    const dyg_planes_t* pls[] = {&p.wqk, &p.wvr, &p.m1, &p.m2, &p.w_ih, &p.w_hh, &p.feat_pl, &p.msg_pl, &p.s_pl, &p.y_pl, &p.h1_pl};
    const int64_t need[] = {Fp, p.H * Dk, Dq + Fp, Fp, MD, Fp, Fp, MD, p.H * Dk, Dq, Fp};
    for (int i = 0; i < 11; ++i)
        DYG_CHECK_ARG(pls[i]->hi && pls[i]->mid && pls[i]->ld >= need[i] && pls[i]->ld % 8 == 0 && aligned16(pls[i]->hi) && aligned16(pls[i]->mid),
                      "dyg_tgn_step: operand planes %d missing, misaligned or narrower than %lld columns", i, (long long)need[i]);
    if (p.p1.hi)
        DYG_CHECK_ARG(p.p1.mid && p.p1.ld >= 2 * Fp && p.p1.ld % 8 == 0 && p.p1_b && p.p2_w && p.p2_b && p.pair_a && p.pair_b && p.ph && p.prob && p.P > 0,
                      "dyg_tgn_step: incomplete link predictor");
    const int ctas = dyg_num_sms();
    const int teams = ctas * TEAMS;
    Plan plan;
    plan.mi_qk = pick_mi(p.R, p.H * Dk, teams);
    plan.mi_o = pick_mi(p.R, Dq, teams);
    plan.mi_m1 = pick_mi(p.R, p.F, teams);
    plan.mi_m2 = pick_mi(p.R, p.F, teams);
    plan.mi_p1 = pick_mi(p.P > 0 ? p.P : 1, p.F, teams);
    constexpr int smem_bytes = TEAMS * TEAM_SMEM_BYTES;
    static_assert(smem_bytes <= 227 * 1024, "two tile teams must fit one SM");
    static_assert(mt::Tile<64, 48>::SMEM_BYTES <= TEAM_SMEM_BYTES, "cell tile fits the team buffer");
    cudaFuncSetAttribute(tgn_step_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);   // per device, cheap
    cudaFuncSetAttribute(tgn_step_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    void* args[] = {(void*)&p, (void*)&plan};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(CTA_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of the whole grid: the phase barrier spins
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (p.G == 3) e = cudaLaunchKernelExC(&cfg, (const void*)tgn_step_kernel<2, 3>, args);
    else e = cudaLaunchKernelExC(&cfg, (const void*)tgn_step_kernel<2, 1>, args);
    if (e != cudaSuccess) {
        dyg_set_error("dyg_tgn_step: launch failed: %s", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
