// dyg_tgn_step: one 200-event batch of the TGN memory model in ONE cooperative launch.
//
// The reference's per-batch body (models/MemoryModel.py:87-168: look-ahead memories, 1-layer graph attention over 10 recent
// neighbours, MergeLayer, then update_memories / clear / new raw messages / store for the positive batch, and the link
// predictor of train_link_prediction.py:243-244) moves ~1 GFLOP and ~20 MB through a chain of dependent steps.  As separate
// kernels that chain was 25-29 launches at ~10 us each (round 1: 269 us per step, 1 % of its HBM time).  Here a persistent grid
// (one CTA per SM, cooperative launch) walks the chain as PHASES separated by a grid barrier; inside a phase every CTA takes
// tile tasks (fp32 FFMA tiles of tile_gemm.cuh, 256 threads) and warp tasks (one warp per root / candidate message) from a
// static list.  The embedding chain and the memory-update chain of the same batch are independent until the commit, so they
// share phases:
//   P0  roots: lower bound on the CSR + recent-neighbour gather + layer-0 features (memory view + raw)
//       candidates: time-order check, persist the look-ahead memories of the batch's nodes, last-message election
//   P1  qk = feat Wqk^T + cq                               | candidates: build the 616-wide raw messages, clear pending
//   P2  roots: folded temporal attention (gather + time encoding + softmax + sum)   | GRU / RNN cell tiles of the candidates
//   P3  o = s Wvr^T + b                                    | winners: commit new look-ahead memories + message store
//   P4  y = LayerNorm(o + [feat | cos(b)])
//   P5  h = relu([y | feat] W1^T + b1)       P6  emb = h W2^T + b2
//   P7  ph = relu([emb[a] | emb[b]] Wp1^T + bp1)           P8  prob = sigmoid(ph . wp2 + bp2)         (optional link predictor)
// Reads of the look-ahead view (P0, P2) precede its update (P3); every scratch buffer is written in one phase and read in a
// later one through L2 (cp.async.cg / ld.global.cg), so no SM can hold a stale L1 line of it.
#include <cooperative_groups.h>
#include <math.h>
#include "tile_gemm.cuh"

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
__device__ __forceinline__ void root_task(const P& p, int r, int lane) {
    const int64_t v = p.roots[r];
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
        *reinterpret_cast<float4*>(out + c) = make_float4(x.x + m.x, x.y + m.y, x.z + m.z, x.w + m.w);
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
        *reinterpret_cast<float4*>(o + j) = ldcg4(p.memory + owner * D + j);
        *reinterpret_cast<float4*>(o + D + j) = ldcg4(p.memory + other * D + j);
    }
    const float dt = (float)p.t[ev] - __ldcg(p.last_update + owner);
    for (int j = lane; j < p.T; j += 32) o[2 * D + j] = dyg_time_enc(dt, __ldg(p.time_w + j), __ldg(p.time_b + j));
    const float* ep = p.edge_raw + p.eid[ev] * p.ld_edge;
    for (int j = lane * 4; j < p.E; j += 128) *reinterpret_cast<float4*>(o + 2 * D + p.T + j) = __ldg(reinterpret_cast<const float4*>(ep + j));
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
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.f / den[h];
        float* orow = p.s + (int64_t)r * (H * Dk) + h * Dk;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int c = i * 32 + lane;
            if (c < NE4) {
                float4 v = acc[h][i];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                *(reinterpret_cast<float4*>(orow) + c) = v;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            if (c < T) orow[NE4 * 4 + c] = acct[h][i] * inv;
        }
    }
}

// y = LayerNorm(o + [feat | cos(b)]) (models/modules.py:199 with the residual of :155,:197); Dq = F + T <= 512
__device__ __forceinline__ void layernorm_task(const P& p, int r, int lane) {
    const int Dq = p.F + p.T;
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = i * 32 + lane;
        float x = 0.f;
        if (c < Dq) x = __ldcg(p.o + (int64_t)r * Dq + c) + (c < p.F ? __ldcg(p.feat + (int64_t)r * p.F + c) : __ldg(p.t0 + c - p.F));
        v[i] = x;
        sum += x;
    }
    const float mean = warp_sum(sum) / Dq;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = i * 32 + lane;
        const float d = c < Dq ? v[i] - mean : 0.f;
        var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(warp_sum(var) / Dq + p.ln_eps);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = i * 32 + lane;
        if (c < Dq) p.y[(int64_t)r * Dq + c] = (v[i] - mean) * rstd * __ldg(p.ln_g + c) + __ldg(p.ln_b + c);
    }
}

__device__ __forceinline__ void score_task(const P& p, int i, int lane) {
    float d = 0.f;
    for (int c = lane; c < p.F; c += 32) d = fmaf(__ldcg(p.ph + (int64_t)i * p.F + c), __ldg(p.p2_w + c), d);
    d = warp_sum(d);
    if (lane == 0) p.prob[i] = tg::sigmoidf_(d + __ldg(p.p2_b));
}

// ------------------------------------------------------------------------------------------------ tile tasks
struct Gemm {
    tg::ASeg seg[2];
    int nseg;
    const float* W;
    int64_t ldw;
    const float* bias;
    int relu;
    float* C;
    int64_t ldc;
    int64_t M;
    int N;
};
template <int TM>
__device__ __forceinline__ int gemm_tiles(const Gemm& g) { return (int)((g.M + 16 * TM - 1) / (16 * TM)) * ((g.N + 63) / 64); }

template <int TM>
__device__ __forceinline__ void gemm_tile(const Gemm& g, int tile, float* smem, int t, int bar) {
    constexpr int TN = 4;
    const int ntn = (g.N + 63) / 64;
    const int64_t m0 = (int64_t)(tile / ntn) * (16 * TM);
    const int n0 = (tile % ntn) * 64;
    const int tx = t & 15, ty = t >> 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const tg::WRows wmap{n0, g.N};
    int koff = 0;
    for (int sidx = 0; sidx < g.nseg; ++sidx) {
        tg::gemm_accum<TM, TN>(acc, g.seg[sidx], m0, g.M, g.W + koff, g.ldw, wmap, smem, t, bar);
        koff += g.seg[sidx].width;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx + 16 * j;
        if (n >= g.N) continue;
        const float b = g.bias ? __ldg(g.bias + n) : 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t m = m0 + ty + 16 * i;
            if (m >= g.M) continue;
            float v = acc[i][j] + b;
            if (g.relu) v = fmaxf(v, 0.f);
            g.C[m * g.ldc + n] = v;
        }
    }
}

// recurrent cell of the candidates (nn.GRUCell / nn.RNNCell, models/MemoryModel.py:490-515): 32 candidates x 16 hidden units per tile
template <int G>
__device__ __forceinline__ void cell_tile(const P& p, int tile, float* smem, int t, int bar) {
    constexpr int TM = 2;
    const int D = p.F, MD = 2 * D + p.T + p.E, nu = (D + 15) / 16;
    const int64_t m0 = (int64_t)(tile / nu) * (16 * TM);
    const int u0 = (tile % nu) * 16;
    const int tx = t & 15, ty = t >> 4;
    const int64_t Pn = 2 * (int64_t)p.B;
    float acc[TM][G];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int g = 0; g < G; ++g) acc[i][g] = 0.f;
    const tg::WGates wmap{u0, D, 16, G};
    tg::gemm_accum<TM, G>(acc, tg::ASeg{p.msg, nullptr, MD, MD}, m0, Pn, p.w_ih, MD, wmap, smem, t, bar);
    float in_n[TM];
    if (G == 3) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            in_n[i] = acc[i][G - 1];
            acc[i][G - 1] = 0.f;
        }
    }
    tg::gemm_accum<TM, G>(acc, tg::ASeg{p.memory, p.cand, D, D}, m0, Pn, p.w_hh, D, wmap, smem, t, bar);
    const int u = u0 + tx;
    if (u >= D) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty + 16 * i;
        if (m >= Pn) continue;
        const int64_t v = p.cand[m];
        if (__ldcg(p.winner + v) != (int32_t)m) continue;
        float hn;
        if (G == 3) {
            const float r = tg::sigmoidf_(acc[i][0] + p.b_ih[u] + p.b_hh[u]);
            const float z = tg::sigmoidf_(acc[i][1] + p.b_ih[D + u] + p.b_hh[D + u]);
            const float ng = tanhf(in_n[i] + p.b_ih[2 * D + u] + r * (acc[i][G - 1] + p.b_hh[2 * D + u]));
            const float h = __ldcg(p.memory + v * D + u);
            hn = ng + z * (h - ng);
        } else {
            hn = tanhf(acc[i][0] + p.b_ih[u] + p.b_hh[u]);
        }
        p.hnew[m * D + u] = hn;
    }
}

struct Plan {   // tile heights (TM) chosen on the host for the actual R / B so that every GEMM phase is ~one wave
    int tm_qk, tm_o, tm_m1, tm_m2, tm_p1;
};

constexpr int TEAMS = 2;                       // tile teams (256 threads each) per CTA: four warps per scheduler instead of two
constexpr int CTA_THREADS = TEAMS * tg::THREADS;
constexpr int TEAM_SMEM_FLOATS = tg::Tile<4, 4>::SMEM_FLOATS;

struct Team {
    int id, n;        // global team index, number of teams in the grid
    int t, bar;       // thread index inside the team, its named barrier
    float* smem;
};

// warp tasks [0, ntasks) are handed out in groups of 8 (one per warp of a team); `task` counts groups from `first_task`
template <class F>
__device__ __forceinline__ void warp_tasks(const Team& tm, int first_task, int ntasks, int task, F f) {
    const int w = (task - first_task) * (tg::THREADS / 32) + (tm.t >> 5);
    if (w < ntasks) f(w, tm.t & 31);
}

#define DYG_TM_SWITCH(tm, CALL) \
    switch (tm) {               \
        case 1: { constexpr int TM_ = 1; CALL; } break; \
        case 2: { constexpr int TM_ = 2; CALL; } break; \
        default: { constexpr int TM_ = 4; CALL; } break; \
    }

__device__ __forceinline__ int tiles_of(const Gemm& g, int tm) {
    const int h = tm == 1 ? 16 : tm == 2 ? 32 : 64;
    return (int)((g.M + h - 1) / h) * ((g.N + 63) / 64);
}
__device__ __forceinline__ void run_tile(const Gemm& g, int tm, int tile, const Team& team) {
    DYG_TM_SWITCH(tm, gemm_tile<TM_>(g, tile, team.smem, team.t, team.bar));
}

template <int H, int G>
__global__ void __launch_bounds__(CTA_THREADS, 1) tgn_step_kernel(const __grid_constant__ P p, const Plan plan) {
    extern __shared__ __align__(16) float smem_all[];
    unsigned target = 0;
    Team team;
    const int local = threadIdx.x / tg::THREADS;
    team.id = blockIdx.x * TEAMS + local;
    team.n = gridDim.x * TEAMS;
    team.t = threadIdx.x % tg::THREADS;
    team.bar = 1 + local;
    team.smem = smem_all + local * TEAM_SMEM_FLOATS;
    constexpr int WPT = tg::THREADS / 32;
    const int R = p.R, C2 = 2 * p.B;
    const int Dk = p.F + p.E + p.T, Dq = p.F + p.T;
    auto groups = [](int n) { return (n + WPT - 1) / WPT; };

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
        const Gemm g{{tg::ASeg{p.feat, nullptr, p.F, p.F}, {}}, 1, p.wqk, p.ld_wqk, p.cq, 0, p.qk, H * Dk, R, H * Dk};
        const int nt = tiles_of(g, plan.tm_qk), g1 = groups(C2);
        for (int task = team.id; task < nt + g1; task += team.n) {
            if (task < nt) run_tile(g, plan.tm_qk, task, team);
            else warp_tasks(team, nt, C2, task, [&](int w, int lane) { message_task(p, w, lane); });
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 2);
    // ---- P2
    {
        const int nt = ((C2 + 31) / 32) * ((p.F + 15) / 16), g1 = groups(R);
        for (int task = team.id; task < nt + g1; task += team.n) {
            if (task < nt) cell_tile<G>(p, task, team.smem, team.t, team.bar);
            else warp_tasks(team, nt, R, task, [&](int w, int lane) { attend_task<H>(p, w, lane); });
        }
    }
    grid_barrier(p.barrier, target);
    stamp(p, 3);
    // ---- P3
    {
        const Gemm g{{tg::ASeg{p.s, nullptr, H * Dk, H * Dk}, {}}, 1, p.wvr, H * Dk, p.rbias, 0, p.o, Dq, R, Dq};
        const int nt = tiles_of(g, plan.tm_o), g1 = groups(C2);
        for (int task = team.id; task < nt + g1; task += team.n) {
            if (task < nt) run_tile(g, plan.tm_o, task, team);
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
        const Gemm g{{tg::ASeg{p.y, nullptr, Dq, Dq}, tg::ASeg{p.feat, nullptr, p.F, p.F}}, 2, p.m1_w, Dq + p.F, p.m1_b, 1, p.h1, p.F, R, p.F};
        for (int task = team.id; task < tiles_of(g, plan.tm_m1); task += team.n) run_tile(g, plan.tm_m1, task, team);
    }
    grid_barrier(p.barrier, target);
    stamp(p, 6);
    // ---- P6
    {
        const Gemm g{{tg::ASeg{p.h1, nullptr, p.F, p.F}, {}}, 1, p.m2_w, p.F, p.m2_b, 0, p.emb, p.F, R, p.F};
        for (int task = team.id; task < tiles_of(g, plan.tm_m2); task += team.n) run_tile(g, plan.tm_m2, task, team);
    }
    if (p.p1_w) {
        grid_barrier(p.barrier, target);
        stamp(p, 7);
        // ---- P7
        {
            const Gemm g{{tg::ASeg{p.emb, p.pair_a, p.F, p.F}, tg::ASeg{p.emb, p.pair_b, p.F, p.F}}, 2, p.p1_w, 2 * p.F, p.p1_b, 1, p.ph, p.F, p.P, p.F};
            for (int task = team.id; task < tiles_of(g, plan.tm_p1); task += team.n) run_tile(g, plan.tm_p1, task, team);
        }
        grid_barrier(p.barrier, target);
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

int pick_tm(int64_t M, int N, int ctas) {
    int best = 4;
    double cost = 1e30;
    for (int tm : {1, 2, 4}) {
        const int64_t tiles = ((M + 16 * tm - 1) / (16 * tm)) * ((N + 63) / 64);
        const double c = (double)((tiles + ctas - 1) / ctas) * (tm + 0.35);   // waves x (work + fixed cost per tile)
        if (c < cost) {
            cost = c;
            best = tm;
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
    DYG_CHECK_ARG(p.F % 4 == 0 && p.E % 4 == 0 && p.T % 4 == 0 && p.ld_node % 4 == 0 && p.ld_edge % 4 == 0 && p.ld_wqk % 4 == 0,
                  "dyg_tgn_step: widths and leading dimensions must be multiples of 4");
    DYG_CHECK_ARG(p.F + p.E <= 384 && p.T <= 128 && p.F + p.T <= 512, "dyg_tgn_step: feature widths out of range");
    DYG_CHECK_ARG(p.he && p.indptr && p.src && p.dst && p.t && p.eid && p.roots && p.cand && p.node_raw && p.edge_raw && p.memory &&
                      p.last_update && p.mem_view && p.lu_view && p.pending && p.winner && p.msg_store && p.msg_time && p.flag && p.barrier,
                  "dyg_tgn_step: null state pointer");
    DYG_CHECK_ARG(p.nbr_ids && p.nbr_eids && p.nbr_t && p.feat && p.qk && p.s && p.o && p.y && p.h1 && p.emb && p.msg && p.hnew,
                  "dyg_tgn_step: null scratch pointer");
    DYG_CHECK_ARG(!p.p1_w || (p.p1_b && p.p2_w && p.p2_b && p.pair_a && p.pair_b && p.ph && p.prob && p.P > 0), "dyg_tgn_step: incomplete link predictor");
    const int ctas = dyg_num_sms();
    const int teams = ctas * TEAMS;
    const int Dk = p.F + p.E + p.T, Dq = p.F + p.T;
    Plan plan;
    plan.tm_qk = pick_tm(p.R, p.H * Dk, teams);
    plan.tm_o = pick_tm(p.R, Dq, teams);
    plan.tm_m1 = pick_tm(p.R, p.F, teams);
    plan.tm_m2 = pick_tm(p.R, p.F, teams);
    plan.tm_p1 = pick_tm(p.P > 0 ? p.P : 1, p.F, teams);
    constexpr int smem_bytes = TEAMS * TEAM_SMEM_FLOATS * 4;
    {   // per device and cheap: set on every call (one process may drive several devices)
        cudaFuncSetAttribute(tgn_step_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        cudaFuncSetAttribute(tgn_step_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    }
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
