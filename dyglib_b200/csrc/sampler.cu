// Temporal neighbour sampling over the device CSR.
// HBM-bound integer / float64-compare work: binary search on the 16-byte half-edge records,
// then a contiguous tail gather.  No tensor cores here by design.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

// ---------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void dyg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* dyg_last_error(void) { return g_err; }
extern "C" int dyg_abi_version(void) { return DYG_ABI_VERSION; }

// ---------------------------------------------------------------- device helpers
struct Rec {
    double t;
    int nbr;
    int eid;
};
__device__ __forceinline__ Rec load_rec(const dyg_halfedge_t* he, int64_t i) {
    const int4 r = __ldg(reinterpret_cast<const int4*>(he + i));
    Rec o;
    o.t = __hiloint2double(r.y, r.x);
    o.nbr = r.z;
    o.eid = r.w;
    return o;
}
__device__ __forceinline__ double load_time(const dyg_halfedge_t* he, int64_t i) {
    return __ldg(reinterpret_cast<const double*>(he + i));
}
// number of half-edges of [a, a+deg) with t < tq  (np.searchsorted side='left', utils/utils.py:141)
__device__ __forceinline__ int64_t lower_bound_time(const dyg_halfedge_t* he, int64_t a, int64_t deg, double tq) {
    int64_t lo = 0, hi = deg;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (load_time(he, a + mid) < tq) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ void node_range(const int64_t* indptr, int64_t num_nodes, int64_t v, int64_t& a, int64_t& deg) {
    a = 0;
    deg = 0;
    if (v >= 0 && v < num_nodes) {
        a = __ldg(indptr + v);
        deg = __ldg(indptr + v + 1) - a;
    }
}

// Cooperative (LANES+1)-ary version of lower_bound_time: the LANES lanes of a query group probe LANES evenly spaced
// records per round (independent loads, one round trip), a ballot counts how many are < tq, and the range shrinks by
// LANES+1; ranges of <= LANES records are finished with one coalesced read.  ceil(log_{LANES+1} deg) dependent round
// trips instead of ceil(log_2 deg).  EVERY lane of the warp must call it (groups without a query pass deg = 0).
template <int LANES>
__device__ __forceinline__ int64_t lower_bound_time_coop(const dyg_halfedge_t* he, int64_t a, int64_t deg, double tq, int lane) {
    const unsigned gshift = ((threadIdx.x & 31) / LANES) * LANES;
    const unsigned gmask = LANES >= 32 ? 0xffffffffu : ((1u << LANES) - 1u);
    int64_t lo = 0, hi = deg;   // answer in [lo, hi]: records below lo are < tq, records from hi on are >= tq
    while (true) {
        const int64_t len = hi - lo;
        const bool active = len > 0;
        if (!__any_sync(0xffffffffu, active)) break;
        const bool tail = len <= LANES;
        int64_t p = lo;
        bool inb = false;
        if (active) {
            if (tail) {
                p = lo + lane;
                inb = lane < len;
            } else {
                p = lo + ((int64_t)(lane + 1) * len) / (LANES + 1);
                inb = true;
            }
        }
        const bool pred = inb && (load_time(he, a + p) < tq);
        const int c = __popc((__ballot_sync(0xffffffffu, pred) >> gshift) & gmask);   // sorted -> the true lanes are a prefix
        if (active) {
            if (tail) {
                lo += c;
                hi = lo;
            } else {
                const int64_t nlo = c > 0 ? lo + ((int64_t)c * len) / (LANES + 1) + 1 : lo;
                const int64_t nhi = c < LANES ? lo + ((int64_t)(c + 1) * len) / (LANES + 1) : hi;
                lo = nlo;
                hi = nhi;
            }
        }
    }
    return lo;
}

// Fence index over the half-edge array (dyg_csr_fence_build): level l >= 1 holds, for every complete block of 16^l
// records (aligned globally, not per node), the time of the block's last record: lvl[l][i] = t[(i + 1) * 16^l - 1].
// Inside one node's run the records are time-sorted, so the fence entries of the complete blocks inside the run are
// sorted too and the lower bound descends level by level: one 128-byte line of 16 entries per level instead of one
// 32-byte sector per binary-search probe, and the top levels (1/256 of the records and less) stay in L2.
#define DYG_FENCE_MAX_LEVELS 8
struct Fence {
    const double* lvl[DYG_FENCE_MAX_LEVELS + 1];   // lvl[0] unused
    int nlev;
};
// Index arithmetic runs in uint32 when the half-edge array has fewer than 2^31 records (half the integer instructions).
// Entries < tq of fence level `f` in [lo, hi), counted by the LANES lanes of a query group (sorted -> a prefix): 16-byte
// loads of entry pairs, all loads of the unrolled rounds issued before the first compare; the group sum is an xor-shuffle
// butterfly, so EVERY lane of the warp must call it (groups with nothing to do pass lo == hi).  Pairs may start one entry
// before lo / end one entry after hi: both stay inside the level's padded allocation and are masked out.
template <int LANES, typename I>
__device__ __forceinline__ int count_fence(const double* __restrict__ f, I lo, I hi, double tq, int lane) {
    constexpr int UNROLL = (32 + 2 * LANES - 1) / (2 * LANES);
    const I i0 = (lo & ~(I)1) + 2 * (I)lane;
    double2 v[UNROLL];
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) {
        const I i = i0 + (I)(r * 2 * LANES);
        v[r] = make_double2(0.0, 0.0);
        if (i < hi) v[r] = __ldg(reinterpret_cast<const double2*>(f + i));
    }
    int c = 0;
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) {
        const I i = i0 + (I)(r * 2 * LANES);
        c += (i >= lo && i < hi && v[r].x < tq) + (i + 1 < hi && v[r].y < tq);
    }
#pragma unroll 1
    for (I i = i0 + (I)(UNROLL * 2 * LANES); i < hi; i += (I)(2 * LANES)) {
        const double2 w = __ldg(reinterpret_cast<const double2*>(f + i));
        c += (w.x < tq) + (i + 1 < hi && w.y < tq);
    }
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}
// same over the records' time fields (8-byte loads at a 16-byte stride)
template <int LANES, typename I>
__device__ __forceinline__ int count_records(const dyg_halfedge_t* __restrict__ he, I lo, I hi, double tq, int lane) {
    constexpr int UNROLL = (16 + LANES - 1) / LANES;
    const I i0 = lo + (I)lane;
    double v[UNROLL];
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) {
        const I i = i0 + (I)(r * LANES);
        v[r] = 0.0;
        if (i < hi) v[r] = load_time(he, (int64_t)i);
    }
    int c = 0;
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) c += (i0 + (I)(r * LANES) < hi) && (v[r] < tq);
#pragma unroll 1
    for (I i = i0 + (I)(UNROLL * LANES); i < hi; i += (I)LANES) c += load_time(he, (int64_t)i) < tq;
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// number of half-edges of [a, a+deg) with t < tq, through the fence index.  Warp-uniform control flow: every lane of
// the warp calls it (lanes without a query pass deg = 0) and the level loop runs from the highest level any group of
// the warp needs; groups that start lower idle through the upper iterations.  The start level is the first one with at
// most ~32 entries in the run; every level below scans one 16-entry block (up to 31 entries at the run's two ends).
template <int LANES, typename I>
__device__ __forceinline__ I lower_bound_time_fenced(const dyg_halfedge_t* __restrict__ he, const Fence& fx, I a, I deg,
                                                     double tq, int lane) {
    const I b = a + deg;
    int top = 0;
    for (I d = deg; d > 32 && top < fx.nlev; d >>= 4) ++top;
    I lo = (a + (((I)1 << (4 * top)) - 1)) >> (4 * top), hi = b >> (4 * top);
    if (hi < lo) hi = lo;
    for (int l = __reduce_max_sync(0xffffffffu, top); l > 0; --l) {   // l is warp-uniform
        const bool on = l <= top;
        const int c = count_fence<LANES, I>(fx.lvl[l], on ? lo : (I)0, on ? hi : (I)0, tq, lane);
        if (on) {
            const I pos = lo + (I)c;
            const I s_l = (a + (((I)1 << (4 * l)) - 1)) >> (4 * l);
            I e_l = b >> (4 * l);
            if (e_l < s_l) e_l = s_l;
            lo = pos == s_l ? (a + (((I)1 << (4 * (l - 1))) - 1)) >> (4 * (l - 1)) : pos << 4;
            hi = pos == e_l ? b >> (4 * (l - 1)) : (pos << 4) + 16;
        }
    }
    return lo + (I)count_records<LANES, I>(he, lo, hi, tq, lane) - a;
}

// Entries <= x of the sorted double array f in [lo, hi), by ONE lane: 16-byte loads of aligned entry pairs (a pair may start
// one entry before lo / end one after hi: both stay inside the allocation and are masked out).
template <typename I>
__device__ __forceinline__ int count_le(const double* __restrict__ f, I lo, I hi, double x) {
    if ((lo & 15) == 0 && hi == lo + 16) {   // the common case below the start level: one whole 128-byte line, no masks
        const double2* p = reinterpret_cast<const double2*>(f + lo);
        double2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(p + i);
        int c16 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) c16 += (v[i].x <= x) + (v[i].y <= x);
        return c16;
    }
    int c = 0;
#pragma unroll 4
    for (I i = lo & ~(I)1; i < hi; i += 2) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(f + i));
        c += (i >= lo && v.x <= x) + (i + 1 < hi && v.y <= x);
    }
    return c;
}
// searchsorted(cum[a : a + cnt], target, 'right') for the time_interval_aware prefix table (one lane per draw), descending the
// fence index of `cum` (same block structure as the CSR fence: level l holds cum[16^l i + 16^l - 1]).  cum is non-decreasing
// inside a node's run, so the complete blocks inside [a, a + cnt) have sorted fence entries.  A binary search costs
// ~log2(cnt) dependent probes, each its own DRAM access (profiles/: 10.4 KB read per query of 20 draws); here a draw reads
// one 128-byte line per level, and the levels above the first (1/256 of the table and less) stay in L2.
template <typename I>
__device__ __forceinline__ I upper_bound_cum_fenced(const double* __restrict__ cum, const Fence& fx, I a, I cnt, double target) {
    const I b = a + cnt;
    int top = 0;
    for (I d = cnt; d > 32 && top < fx.nlev; d >>= 4) ++top;
    I lo = (a + (((I)1 << (4 * top)) - 1)) >> (4 * top), hi = b >> (4 * top);
    if (hi < lo) hi = lo;
    for (int l = top; l > 0; --l) {
        const I pos = lo + (I)count_le<I>(fx.lvl[l], lo, hi, target);
        const I s_l = (a + (((I)1 << (4 * l)) - 1)) >> (4 * l);
        I e_l = b >> (4 * l);
        if (e_l < s_l) e_l = s_l;
        lo = pos == s_l ? (a + (((I)1 << (4 * (l - 1))) - 1)) >> (4 * (l - 1)) : pos << 4;
        hi = pos == e_l ? b >> (4 * (l - 1)) : (pos << 4) + 16;
    }
    return lo + (I)count_le<I>(cum, lo, hi, target) - a;
}

// Interpolation start for the same search.  The table's increments are exp(p) with p in [0, 1] (utils/utils.py:112-128 feeds
// probabilities, not logits, to the softmax of :183), i.e. within [1, e]: cum is close to linear inside a prefix and
// a linearly interpolated position g is within a few entries of the answer.  Read the 128-byte line around the guess, step to the neighbouring
// line while the target lies outside, and only then fall back to the fence descent on the remaining side.  The result is the
// exact searchsorted position in every case; in the common case a draw costs one line read instead of one per level.
template <typename I>
__device__ __forceinline__ I upper_bound_cum_interp(const double* __restrict__ cum, const Fence& fx, I a, I cnt, I g, double target) {
    const I b = a + cnt;
    if (g >= cnt) g = cnt - 1;
    I bs = (a + g) & ~(I)15;
    int dir = 0;   // direction of the previous step: a reversal means the answer is the boundary between the two lines
#pragma unroll 1
    for (int tries = 0; tries < 4; ++tries) {
        const I lo = bs > a ? bs : a, hi = bs + 16 < b ? bs + 16 : b;
        const I c = (I)count_le<I>(cum, lo, hi, target);
        if (c == 0 && lo > a) {
            if (dir > 0) return lo - a;
            if (tries == 3) return upper_bound_cum_fenced<I>(cum, fx, a, lo - a, target);
            bs -= 16;
            dir = -1;
        } else if (c == hi - lo && hi < b) {
            if (dir < 0) return hi - a;
            if (tries == 3) return (hi - a) + upper_bound_cum_fenced<I>(cum, fx, hi, b - hi, target);
            bs += 16;
            dir = 1;
        } else {
            return lo + c - a;
        }
    }
    return 0;   // not reached
}

// Secant refinement in front of it.  A divergent 16-byte load costs the L1 one wavefront per lane, so the 8 loads of a
// whole line per draw bound the kernel on L1 wavefronts (20 draws x 8 loads x 32 lanes; profiles/).  Here a candidate
// position s is checked by reading just cum[s-1] and cum[s] (one aligned pair when s is odd, two otherwise); a miss moves s
// by (target - value) / local increment.  With increments in [1, e] the second or third candidate is the answer; after four
// misses the line / fence search above takes over, so the result is exact for any table.
template <typename I>
__device__ __forceinline__ I upper_bound_cum_secant(const double* __restrict__ cum, const Fence& fx, I a, I cnt, I g, double target) {
    I s = g > cnt ? cnt : g;
#pragma unroll 1
    for (int tries = 0; tries < 4; ++tries) {
        double left, right;   // cum[a + s - 1], cum[a + s]
        const I i = a + s;
        if (s > 0 && s < cnt && (i & 1)) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(cum + i - 1));
            left = v.x;
            right = v.y;
        } else {
            left = s > 0 ? __ldg(cum + i - 1) : -INFINITY;
            right = s < cnt ? __ldg(cum + i) : INFINITY;
        }
        if (left <= target && target < right) return s;
        double m = (s > 0 && s < cnt) ? right - left : 1.0;
        if (!(m > 1e-300)) m = 1.0;
        double step;
        if (target >= right) step = 1.0 + floor((target - right) / m);
        else step = -1.0 - floor((left - target) / m);
        const double ns = (double)s + step;
        s = ns <= 0.0 ? (I)0 : (ns >= (double)cnt ? cnt : (I)ns);
    }
    return upper_bound_cum_interp<I>(cum, fx, a, cnt, g, target);
}

// MODE 0: no fence index (cooperative (LANES+1)-ary search over the records); 1: fence index, uint32 index arithmetic
// (fewer than 2^31 half-edges); 2: fence index, int64 index arithmetic.  Every lane of the warp must call it.
template <int LANES, int MODE>
__device__ __forceinline__ int64_t search_before(const dyg_halfedge_t* __restrict__ he, const Fence& fx, int64_t a, int64_t deg,
                                                 double tq, int lane) {
    if (MODE == 1) return (int64_t)lower_bound_time_fenced<LANES, uint32_t>(he, fx, (uint32_t)a, (uint32_t)deg, tq, lane);
    if (MODE == 2) return lower_bound_time_fenced<LANES, int64_t>(he, fx, a, deg, tq, lane);
    return lower_bound_time_coop<LANES>(he, a, deg, tq, lane);
}
static inline int search_mode(const double* fence, int64_t n_half) {
    if (!fence) return 0;
    if (getenv("DYG_FENCE_INDEX64")) return 2;   // test hook: exercise the int64 variant on small graphs
    return n_half < ((int64_t)1 << 31) - 64 ? 1 : 2;
}

// ---------------------------------------------------------------- CSR build
__global__ void csr_degrees_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                   int64_t num_nodes, unsigned long long* __restrict__ deg) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[e], d = dst[e];
        if (s >= 0 && s < num_nodes) atomicAdd(deg + s, 1ull);
        if (d >= 0 && d < num_nodes) atomicAdd(deg + d, 1ull);
    }
}
__global__ void csr_pack_kernel(const int64_t* __restrict__ order, const int64_t* __restrict__ src,
                                const int64_t* __restrict__ dst, const int64_t* __restrict__ eid,
                                const double* __restrict__ t, int64_t n_half, dyg_halfedge_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_half; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = order[i];
        const int64_t e = h >> 1;
        const int64_t nb = (h & 1) ? src[e] : dst[e];
        const double tt = t[e];
        int4 r;
        r.x = __double2loint(tt);
        r.y = __double2hiint(tt);
        r.z = (int)nb;
        r.w = (int)eid[e];
        reinterpret_cast<int4*>(out)[i] = r;
    }
}
// one fence level: out[i] = (level 1) time of record 16 i + 15, (level > 1) previous level's entry 16 i + 15
__global__ void csr_fence_kernel(const dyg_halfedge_t* __restrict__ he, const double* __restrict__ prev, int64_t n_out,
                                 double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = prev ? prev[16 * i + 15] : load_time(he, 16 * i + 15);
}
__global__ void csr_tia_kernel(const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr, int64_t num_nodes,
                               double tsf, double* __restrict__ prob, double* __restrict__ cum,
                               const double* __restrict__ prob_in) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < num_nodes; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = indptr[v], b = indptr[v + 1];
        if (a == b) continue;
        const double t_last = prob_in ? 0.0 : load_time(he, b - 1);
        double c = 0.0, s = 0.0;
        for (int64_t j = a; j < b; ++j) {
            double p;
            if (prob_in) {
                p = prob_in[j];
            } else {
                const double e = exp(tsf * (load_time(he, j) - t_last));
                c += e;
                p = e / c;
                if (isnan(p)) p = -1e10;
                if (prob) prob[j] = p;
            }
            if (cum) {
                s += exp(p);
                cum[j] = s;
            }
        }
    }
}

// ---------------------------------------------------------------- queries
// 8 lanes per query (search_before)
template <int MODE>
__global__ void __launch_bounds__(256) count_before_kernel(const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr,
                                    int64_t num_nodes, const __grid_constant__ Fence fx, const int64_t* __restrict__ node_ids,
                                    const double* __restrict__ times, int64_t n, int32_t* __restrict__ cnt) {
    const int lane = threadIdx.x % 8;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 8;
    const bool valid = q < n;
    int64_t a = 0, deg = 0;
    double tq = 0.0;
    if (valid) {
        node_range(indptr, num_nodes, __ldg(node_ids + q), a, deg);
        tq = __ldg(times + q);
    }
    const int64_t c = search_before<8, MODE>(he, fx, a, deg, tq, lane);
    if (valid && lane == 0) cnt[q] = (int32_t)c;
}

// LANES threads cooperate on one query: (LANES+1)-ary search (lower_bound_time_coop),
// per probe per group), then the lanes split the k-entry tail gather and the (n,k) row writes.
template <int LANES, int MODE>
__global__ void __launch_bounds__(256) sample_recent_kernel(
    const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr, int64_t num_nodes, const __grid_constant__ Fence fx,
    const int64_t* __restrict__ node_ids, const double* __restrict__ times, int64_t n, int k,
    int64_t* __restrict__ out_nbr, int64_t* __restrict__ out_eid, float* __restrict__ out_t,
    int32_t* __restrict__ cnt_out, bool pairs) {
    const int lane = threadIdx.x % LANES;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    const bool valid = q < n;
    int64_t a = 0, deg = 0;
    double tq = 0.0;
    if (valid) {
        node_range(indptr, num_nodes, __ldg(node_ids + q), a, deg);
        tq = __ldg(times + q);
    }
    const int64_t cnt = search_before<LANES, MODE>(he, fx, a, deg, tq, lane);
    if (!valid) return;
    if (cnt_out && lane == 0) cnt_out[q] = (int32_t)cnt;
    const int64_t base = q * (int64_t)k;
    const int64_t first = a + cnt - k;   // record of output column 0; columns j < k - cnt are the left padding (utils/utils.py:206-209)
    const int pad = cnt < k ? (int)(k - cnt) : 0;
    if (pairs) {
        // even k, aligned rows: a lane owns column pairs, two 16-byte record loads -> 16 + 16 + 8 byte row stores (rows are 16-byte aligned)
        for (int j = 2 * lane; j < k; j += 2 * LANES) {
            longlong2 nb = make_longlong2(0, 0), ei = make_longlong2(0, 0);
            float2 tf = make_float2(0.f, 0.f);
            if (j >= pad) {
                const Rec r = load_rec(he, first + j);
                nb.x = r.nbr;
                ei.x = r.eid;
                tf.x = (float)r.t;
            }
            if (j + 1 >= pad) {
                const Rec r = load_rec(he, first + j + 1);
                nb.y = r.nbr;
                ei.y = r.eid;
                tf.y = (float)r.t;
            }
            // streaming stores: the rows are written once and not read by this kernel; keeping them out of the way leaves L2
            // to the fence levels and the row pointers
            __stcs(reinterpret_cast<longlong2*>(out_nbr + base + j), nb);
            __stcs(reinterpret_cast<longlong2*>(out_eid + base + j), ei);
            __stcs(reinterpret_cast<float2*>(out_t + base + j), tf);
        }
    } else {
        for (int j = lane; j < k; j += LANES) {
            int64_t nb = 0, ei = 0;
            float tf = 0.f;
            if (j >= pad) {
                const Rec r = load_rec(he, first + j);
                nb = r.nbr;
                ei = r.eid;
                tf = (float)r.t;
            }
            out_nbr[base + j] = nb;
            out_eid[base + j] = ei;
            out_t[base + j] = tf;
        }
    }
}

// One warp per query: [self, last min(cnt, L-1) neighbours, zeros...] (models/DyGFormer.py:214-242).
template <int MODE>
__global__ void __launch_bounds__(256) first_hop_pad_kernel(
    const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr, int64_t num_nodes, const __grid_constant__ Fence fx,
    const int64_t* __restrict__ node_ids, const double* __restrict__ times, int64_t n, int L, int row_stride,
    int64_t* __restrict__ out_nbr, int64_t* __restrict__ out_eid, float* __restrict__ out_t,
    int32_t* __restrict__ out_len, int32_t* __restrict__ group_max, int group_size) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;                       // warp-uniform: one warp per query
    const int64_t v = __ldg(node_ids + q);
    const double tq = __ldg(times + q);
    int64_t a, deg;
    node_range(indptr, num_nodes, v, a, deg);
    const int64_t cnt = search_before<32, MODE>(he, fx, a, deg, tq, lane);
    const int m = (int)(cnt < (int64_t)(L - 1) ? cnt : (int64_t)(L - 1));
    if (lane == 0) {
        if (out_len) out_len[q] = m + 1;
        if (group_max) atomicMax(group_max + q / group_size, m + 1);
    }
    const int64_t base = q * (int64_t)row_stride;
    const int64_t first = a + cnt - m;
    for (int j = lane; j < row_stride; j += 32) {
        int64_t nb = 0, ei = 0;
        float tf = 0.f;
        if (j == 0) {
            nb = v;
            tf = (float)tq;
        } else if (j <= m) {
            const Rec r = load_rec(he, first + (j - 1));
            nb = r.nbr;
            ei = r.eid;
            tf = (float)r.t;
        }
        out_nbr[base + j] = nb;
        out_eid[base + j] = ei;
        out_t[base + j] = tf;
    }
}

// Gather at drawn positions, then order each row by float32 time (utils/utils.py:189-199).  Rank sort in
// shared memory; equal times keep draw order (the reference's argsort leaves that order unspecified).
template <int LANES>
__global__ void sample_indexed_kernel(const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr,
                                      const int64_t* __restrict__ node_ids, const int32_t* __restrict__ cnt,
                                      const int64_t* __restrict__ sel, int64_t n, int k,
                                      int64_t* __restrict__ out_nbr, int64_t* __restrict__ out_eid,
                                      float* __restrict__ out_t) {
    extern __shared__ unsigned char smem_raw[];
    const int groups = blockDim.x / LANES;
    const int g = threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    float* st = reinterpret_cast<float*>(smem_raw) + (size_t)g * k;
    int* sn = reinterpret_cast<int*>(smem_raw) + (size_t)groups * k + (size_t)g * k;
    int* se = reinterpret_cast<int*>(smem_raw) + (size_t)2 * groups * k + (size_t)g * k;
    const int64_t q = blockIdx.x * (int64_t)groups + g;
    const bool valid = q < n;
    int c = 0;
    int64_t a = 0;
    if (valid) {
        c = cnt[q];
        if (c > 0) a = __ldg(indptr + node_ids[q]);
    }
    const int64_t base = q * (int64_t)k;
    if (valid && c > 0) {
        for (int j = lane; j < k; j += LANES) {
            int64_t s = sel[base + j];
            s = s < 0 ? 0 : (s >= c ? c - 1 : s);
            const Rec r = load_rec(he, a + s);
            st[j] = (float)r.t;
            sn[j] = r.nbr;
            se[j] = r.eid;
        }
    }
    if (LANES >= 32) __syncthreads(); else __syncwarp();
    if (valid) {
        for (int j = lane; j < k; j += LANES) {
            if (c > 0) {
                const float tj = st[j];
                int rank = 0;
                for (int i = 0; i < k; ++i) {
                    const float ti = st[i];
                    rank += (ti < tj) || (ti == tj && i < j);
                }
                out_nbr[base + rank] = sn[j];
                out_eid[base + rank] = se[j];
                out_t[base + rank] = tj;
            } else {
                out_nbr[base + j] = 0;
                out_eid[base + j] = 0;
                out_t[base + j] = 0.f;
            }
        }
    }
}

__global__ void draw_uniform_kernel(const int32_t* __restrict__ cnt, const double* __restrict__ u, int64_t total, int k,
                                    int64_t* __restrict__ sel) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = cnt[i / k];
    int64_t s = 0;
    if (c > 0) {
        s = (int64_t)floor(u[i] * (double)c);
        if (s >= c) s = c - 1;
    }
    sel[i] = s;
}

// searchsorted(cdf, u, 'right') on cdf[j] = cum[a+j] / cum[a+cnt-1] without materialising the cdf.
__global__ void draw_tia_kernel(const double* __restrict__ cum, const int64_t* __restrict__ indptr,
                                const int64_t* __restrict__ node_ids, const int32_t* __restrict__ cnt,
                                const double* __restrict__ u, int64_t total, int k, int64_t* __restrict__ sel) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t q = i / k;
    const int c = cnt[q];
    int64_t s = 0;
    if (c > 0) {
        const int64_t a = __ldg(indptr + node_ids[q]);
        const double tot = __ldg(cum + a + c - 1);
        const double uu = u[i];
        if (!(tot > 0.0)) {
            s = (int64_t)floor(uu * (double)c);  // every weight underflowed: softmax of equal values is uniform
        } else {
            const double target = uu * tot;
            int64_t lo = 0, hi = c;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(cum + a + mid) <= target) lo = mid + 1; else hi = mid;
            }
            s = lo;
        }
        if (s >= c) s = c - 1;
    }
    sel[i] = s;
}

// Philox4x32-10 counter RNG -> float64 uniforms with numpy's 53-bit construction ((a>>5)*2^26+(b>>6))/2^53.
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
}
__global__ void philox_uniform_kernel(uint64_t seed, uint64_t offset, int64_t count, double* __restrict__ u) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // one counter -> two doubles
    if (2 * i >= count) return;
    const uint64_t ctr = offset / 2 + (uint64_t)i;
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const double inv = 1.0 / 9007199254740992.0;
    u[2 * i] = ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * inv;
    if (2 * i + 1 < count) u[2 * i + 1] = ((double)(c2 >> 5) * 67108864.0 + (double)(c3 >> 6)) * inv;
}

// Fused throughput path for the random strategies: search + counter-based draw + gather + time re-sort in one
// kernel (no (n,k) uniforms / positions round trip through HBM).  Draw (q, j) uses Philox counter
// offset + q*k + j, so results do not depend on the launch shape or on how queries are sharded over GPUs.
__device__ __forceinline__ double philox_u01(uint64_t seed, uint64_t ctr) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * (1.0 / 9007199254740992.0);
}

template <int LANES, bool TIA, int MODE>
__global__ void __launch_bounds__(256) sample_random_kernel(
    const dyg_halfedge_t* __restrict__ he, const int64_t* __restrict__ indptr, int64_t num_nodes, const __grid_constant__ Fence fx,
    const double* __restrict__ cum, const __grid_constant__ Fence cfx, const int64_t* __restrict__ node_ids,
    const double* __restrict__ times, int64_t n, int k, uint64_t seed, uint64_t offset, int64_t* __restrict__ out_nbr,
    int64_t* __restrict__ out_eid, float* __restrict__ out_t) {
    extern __shared__ unsigned char smem_raw[];
    const int groups = blockDim.x / LANES;
    const int g = threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    float* st = reinterpret_cast<float*>(smem_raw) + (size_t)g * k;
    int* sn = reinterpret_cast<int*>(smem_raw) + (size_t)groups * k + (size_t)g * k;
    int* se = reinterpret_cast<int*>(smem_raw) + (size_t)2 * groups * k + (size_t)g * k;
    const int64_t q = blockIdx.x * (int64_t)groups + g;
    const bool valid = q < n;
    int64_t a = 0, cnt = 0;
    {
        int64_t deg = 0;
        double tq = 0.0;
        if (valid) {
            node_range(indptr, num_nodes, __ldg(node_ids + q), a, deg);
            tq = __ldg(times + q);
        }
        cnt = search_before<LANES, MODE>(he, fx, a, deg, tq, lane);
    }
    const int64_t base = q * (int64_t)k;
    if (valid && cnt > 0) {
        double tot = 0.0;
        if (TIA) tot = __ldg(cum + a + cnt - 1);
        for (int j = lane; j < k; j += LANES) {
            const double u = philox_u01(seed, offset + (uint64_t)(base + j));
            int64_t s;
            if (TIA && tot > 0.0) {
                const double target = u * tot;
                const int64_t g = (int64_t)(u * (double)cnt);   // first guess: the table is close to linear (see upper_bound_cum_interp)
                if (cfx.nlev > 0 && MODE == 1) {
                    s = (int64_t)upper_bound_cum_secant<uint32_t>(cum, cfx, (uint32_t)a, (uint32_t)cnt, (uint32_t)g, target);
                } else if (cfx.nlev > 0) {
                    s = upper_bound_cum_secant<int64_t>(cum, cfx, a, cnt, g, target);
                } else {
                    int64_t lo = 0, hi = cnt;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (__ldg(cum + a + mid) <= target) lo = mid + 1; else hi = mid;
                    }
                    s = lo;
                }
            } else {
                s = (int64_t)floor(u * (double)cnt);
            }
            if (s >= cnt) s = cnt - 1;
            const Rec r = load_rec(he, a + s);
            st[j] = (float)r.t;
            sn[j] = r.nbr;
            se[j] = r.eid;
        }
    }
    if (LANES >= 32) __syncthreads(); else __syncwarp();
    if (valid) {
        for (int j = lane; j < k; j += LANES) {
            if (cnt > 0) {
                const float tj = st[j];
                int rank = 0;
                for (int i = 0; i < k; ++i) {
                    const float ti = st[i];
                    rank += (ti < tj) || (ti == tj && i < j);
                }
                out_nbr[base + rank] = sn[j];
                out_eid[base + rank] = se[j];
                out_t[base + rank] = tj;
            } else {
                out_nbr[base + j] = 0;
                out_eid[base + j] = 0;
                out_t[base + j] = 0.f;
            }
        }
    }
}

// ---------------------------------------------------------------- C ABI
#define DISPATCH_MODE(mode, LAUNCH) \
    do {                            \
        if ((mode) == 1) LAUNCH(1); \
        else if ((mode) == 2) LAUNCH(2); \
        else LAUNCH(0);             \
    } while (0)
static inline unsigned blocks_for(int64_t work, int threads, int64_t cap = (1ll << 30)) {
    int64_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (unsigned)b;
}

// level sizes / offsets of the fence buffer (every level starts on a 128-byte line)
static inline int fence_levels(int64_t n_half, int64_t* off, int64_t* cnt) {
    int nlev = 0;
    int64_t pos = 0, n = n_half >> 4;
    while (n > 0 && nlev < DYG_FENCE_MAX_LEVELS) {
        ++nlev;
        off[nlev] = pos;
        cnt[nlev] = n;
        pos += (n + 15) & ~(int64_t)15;
        n >>= 4;
    }
    off[0] = pos;   // total entries
    return nlev;
}
static inline Fence make_fence(const double* fence, int64_t n_half) {
    Fence fx;
    memset(&fx, 0, sizeof(fx));
    if (fence) {
        int64_t off[DYG_FENCE_MAX_LEVELS + 1], cnt[DYG_FENCE_MAX_LEVELS + 1];
        fx.nlev = fence_levels(n_half, off, cnt);
        for (int l = 1; l <= fx.nlev; ++l) fx.lvl[l] = fence + off[l];
    }
    return fx;
}

extern "C" int64_t dyg_csr_fence_entries(int64_t num_half_edges) {
    int64_t off[DYG_FENCE_MAX_LEVELS + 1], cnt[DYG_FENCE_MAX_LEVELS + 1];
    if (num_half_edges < 0) return 0;
    fence_levels(num_half_edges, off, cnt);
    return off[0];
}

extern "C" int dyg_csr_fence_build(const dyg_halfedge_t* he, int64_t num_half_edges, double* fence, dyg_stream_t stream) {
    DYG_CHECK_ARG(num_half_edges >= 0, "dyg_csr_fence_build: bad size");
    int64_t off[DYG_FENCE_MAX_LEVELS + 1], cnt[DYG_FENCE_MAX_LEVELS + 1];
    const int nlev = fence_levels(num_half_edges, off, cnt);
    if (nlev == 0) return 0;
    DYG_CHECK_ARG(he && fence && (reinterpret_cast<uintptr_t>(fence) & 127u) == 0, "dyg_csr_fence_build: fence must be 128-byte aligned");
    for (int l = 1; l <= nlev; ++l) {
        csr_fence_kernel<<<blocks_for(cnt[l], 256, 148 * 16), 256, 0, as_stream(stream)>>>(
            he, l == 1 ? nullptr : fence + off[l - 1], cnt[l], fence + off[l]);
        DYG_LAUNCH_CHECK("dyg_csr_fence_build");
    }
    return 0;
}

extern "C" int dyg_cum_fence_build(const double* cum, int64_t num_half_edges, double* fence, dyg_stream_t stream) {
    DYG_CHECK_ARG(num_half_edges >= 0, "dyg_cum_fence_build: bad size");
    int64_t off[DYG_FENCE_MAX_LEVELS + 1], cnt[DYG_FENCE_MAX_LEVELS + 1];
    const int nlev = fence_levels(num_half_edges, off, cnt);
    if (nlev == 0) return 0;
    DYG_CHECK_ARG(cum && fence && (reinterpret_cast<uintptr_t>(fence) & 127u) == 0 && (reinterpret_cast<uintptr_t>(cum) & 15u) == 0,
                  "dyg_cum_fence_build: fence must be 128-byte aligned, cum 16-byte aligned");
    for (int l = 1; l <= nlev; ++l) {
        csr_fence_kernel<<<blocks_for(cnt[l], 256, 148 * 16), 256, 0, as_stream(stream)>>>(
            nullptr, l == 1 ? cum : fence + off[l - 1], cnt[l], fence + off[l]);
        DYG_LAUNCH_CHECK("dyg_cum_fence_build");
    }
    return 0;
}

extern "C" int dyg_csr_degrees(const int64_t* src, const int64_t* dst, int64_t E, int64_t num_nodes, int64_t* deg,
                               dyg_stream_t stream) {
    DYG_CHECK_ARG(E >= 0 && num_nodes > 0, "dyg_csr_degrees: bad sizes");
    if (E == 0) return 0;
    csr_degrees_kernel<<<blocks_for(E, 256, 148 * 32), 256, 0, as_stream(stream)>>>(
        src, dst, E, num_nodes, reinterpret_cast<unsigned long long*>(deg));
    DYG_LAUNCH_CHECK("dyg_csr_degrees");
    return 0;
}

extern "C" int dyg_csr_pack(const int64_t* order, const int64_t* src, const int64_t* dst, const int64_t* eid,
                            const double* t, int64_t n_half, dyg_halfedge_t* out, dyg_stream_t stream) {
    DYG_CHECK_ARG(n_half >= 0, "dyg_csr_pack: bad size");
    DYG_CHECK_ARG(aligned16(out), "dyg_csr_pack: out must be 16-byte aligned");
    if (n_half == 0) return 0;
    csr_pack_kernel<<<blocks_for(n_half, 256, 148 * 32), 256, 0, as_stream(stream)>>>(order, src, dst, eid, t, n_half, out);
    DYG_LAUNCH_CHECK("dyg_csr_pack");
    return 0;
}

extern "C" int dyg_csr_tia_tables(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, double tsf,
                                  double* prob, double* cum, dyg_stream_t stream) {
    DYG_CHECK_ARG(num_nodes > 0, "dyg_csr_tia_tables: bad size");
    csr_tia_kernel<<<blocks_for(num_nodes, 128, 148 * 16), 128, 0, as_stream(stream)>>>(he, indptr, num_nodes, tsf, prob, cum, nullptr);
    DYG_LAUNCH_CHECK("dyg_csr_tia_tables");
    return 0;
}

extern "C" int dyg_csr_tia_cum(const double* prob, const int64_t* indptr, int64_t num_nodes, double* cum,
                               dyg_stream_t stream) {
    DYG_CHECK_ARG(num_nodes > 0 && prob && cum, "dyg_csr_tia_cum: bad args");
    csr_tia_kernel<<<blocks_for(num_nodes, 128, 148 * 16), 128, 0, as_stream(stream)>>>(nullptr, indptr, num_nodes, 0.0, nullptr, cum, prob);
    DYG_LAUNCH_CHECK("dyg_csr_tia_cum");
    return 0;
}

extern "C" int dyg_count_before(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                                int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int32_t* cnt,
                                dyg_stream_t stream) {
    DYG_CHECK_ARG(n >= 0, "dyg_count_before: bad size");
    if (n == 0) return 0;
    const Fence fx = make_fence(fence, num_half_edges);
#define LAUNCH_COUNT(MODE) count_before_kernel<MODE><<<blocks_for(n * 8, 256), 256, 0, as_stream(stream)>>>(he, indptr, num_nodes, fx, node_ids, times, n, cnt)
    DISPATCH_MODE(search_mode(fence, num_half_edges), LAUNCH_COUNT);
#undef LAUNCH_COUNT
    DYG_LAUNCH_CHECK("dyg_count_before");
    return 0;
}

extern "C" int dyg_sample_recent(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                                 int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int k,
                                 int64_t* out_nbr, int64_t* out_eid, float* out_t, int32_t* cnt, dyg_stream_t stream) {
    DYG_CHECK_ARG(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    DYG_CHECK_ARG(n >= 0, "dyg_sample_recent: bad size");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    const Fence fx = make_fence(fence, num_half_edges);
    const int mode = search_mode(fence, num_half_edges);
    const bool pairs = (k % 2) == 0 && aligned16(out_nbr) && aligned16(out_eid) && (reinterpret_cast<uintptr_t>(out_t) & 7u) == 0;
#define LAUNCH_RECENT_M(MODE) sample_recent_kernel<LR, MODE><<<blocks_for(n * LR, 256), 256, 0, s>>>(he, indptr, num_nodes, fx, node_ids, times, n, k, out_nbr, out_eid, out_t, cnt, pairs)
#define LAUNCH_RECENT(L)                          \
    do {                                          \
        constexpr int LR = L;                     \
        DISPATCH_MODE(mode, LAUNCH_RECENT_M);     \
    } while (0)
    // lanes per query: the search's control flow is per warp, so fewer lanes per query = more queries per warp instruction
    // (k = 20 on the sweep graph: 4 lanes 3.33 ms, 8 lanes 4.00 ms, 16 lanes 5.8 ms per 2^24 queries)
    if (k <= 8) LAUNCH_RECENT(2);
    else if (k <= 40) LAUNCH_RECENT(4);
    else if (k <= 96) LAUNCH_RECENT(8);
    else if (k <= 256) LAUNCH_RECENT(16);
    else LAUNCH_RECENT(32);
#undef LAUNCH_RECENT
#undef LAUNCH_RECENT_M
    DYG_LAUNCH_CHECK("dyg_sample_recent");
    return 0;
}

extern "C" int dyg_sample_indexed(const dyg_halfedge_t* he, const int64_t* indptr, const int64_t* node_ids,
                                  const int32_t* cnt, const int64_t* sel, int64_t n, int k, int64_t* out_nbr,
                                  int64_t* out_eid, float* out_t, dyg_stream_t stream) {
    DYG_CHECK_ARG(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    DYG_CHECK_ARG((size_t)k * 12 <= 40 * 1024, "dyg_sample_indexed: num_neighbors %d too large (max 3413)", k);
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    if (k <= 32) {
        const int lanes = 8, threads = 256, groups = threads / lanes;
        sample_indexed_kernel<8><<<blocks_for(n, groups), threads, (size_t)groups * k * 12, s>>>(
            he, indptr, node_ids, cnt, sel, n, k, out_nbr, out_eid, out_t);
    } else {
        int groups = (int)((40 * 1024) / ((size_t)k * 12));
        if (groups > 8) groups = 8;
        if (groups < 1) groups = 1;
        sample_indexed_kernel<32><<<blocks_for(n, groups), groups * 32, (size_t)groups * k * 12, s>>>(
            he, indptr, node_ids, cnt, sel, n, k, out_nbr, out_eid, out_t);
    }
    DYG_LAUNCH_CHECK("dyg_sample_indexed");
    return 0;
}

extern "C" int dyg_draw_uniform(const int32_t* cnt, const double* u, int64_t n, int k, int64_t* sel, dyg_stream_t stream) {
    DYG_CHECK_ARG(k > 0 && n >= 0, "dyg_draw_uniform: bad sizes");
    if (n == 0) return 0;
    draw_uniform_kernel<<<blocks_for(n * k, 256), 256, 0, as_stream(stream)>>>(cnt, u, n * k, k, sel);
    DYG_LAUNCH_CHECK("dyg_draw_uniform");
    return 0;
}

extern "C" int dyg_draw_tia(const double* cum, const int64_t* indptr, const int64_t* node_ids, const int32_t* cnt,
                            const double* u, int64_t n, int k, int64_t* sel, dyg_stream_t stream) {
    DYG_CHECK_ARG(k > 0 && n >= 0, "dyg_draw_tia: bad sizes");
    if (n == 0) return 0;
    draw_tia_kernel<<<blocks_for(n * k, 256), 256, 0, as_stream(stream)>>>(cum, indptr, node_ids, cnt, u, n * k, k, sel);
    DYG_LAUNCH_CHECK("dyg_draw_tia");
    return 0;
}

extern "C" int dyg_philox_uniform(uint64_t seed, uint64_t offset, int64_t count, double* u, dyg_stream_t stream) {
    DYG_CHECK_ARG(count >= 0 && (offset % 2) == 0, "dyg_philox_uniform: offset must be even");
    if (count == 0) return 0;
    philox_uniform_kernel<<<blocks_for((count + 1) / 2, 256), 256, 0, as_stream(stream)>>>(seed, offset, count, u);
    DYG_LAUNCH_CHECK("dyg_philox_uniform");
    return 0;
}

extern "C" int dyg_first_hop_pad(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                                 int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int L,
                                 int row_stride, int64_t* out_nbr, int64_t* out_eid, float* out_t, int32_t* out_len,
                                 int32_t* group_max, int group_size, dyg_stream_t stream) {
    DYG_CHECK_ARG(L - 1 > 0, "Maximal number of neighbors for each node should be greater than 1!");
    DYG_CHECK_ARG(row_stride >= L, "dyg_first_hop_pad: row_stride %d < max_input_sequence_length %d", row_stride, L);
    DYG_CHECK_ARG(!group_max || group_size > 0, "dyg_first_hop_pad: group_size must be positive");
    if (n == 0) return 0;
    const Fence fx = make_fence(fence, num_half_edges);
#define LAUNCH_PAD(MODE)                                                                                      \
    first_hop_pad_kernel<MODE><<<blocks_for(n * 32, 256), 256, 0, as_stream(stream)>>>(                       \
        he, indptr, num_nodes, fx, node_ids, times, n, L, row_stride, out_nbr, out_eid, out_t, out_len, group_max, \
        group_size > 0 ? group_size : 1)
    DISPATCH_MODE(search_mode(fence, num_half_edges), LAUNCH_PAD);
#undef LAUNCH_PAD
    DYG_LAUNCH_CHECK("dyg_first_hop_pad");
    return 0;
}

extern "C" int dyg_sample_random(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                                 int64_t num_half_edges, const double* cum, const double* cum_fence, const int64_t* node_ids,
                                 const double* times, int64_t n, int k, uint64_t seed, uint64_t offset, int64_t* out_nbr,
                                 int64_t* out_eid, float* out_t, dyg_stream_t stream) {
    DYG_CHECK_ARG(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    DYG_CHECK_ARG((size_t)k * 12 <= 40 * 1024, "dyg_sample_random: num_neighbors %d too large (max 3413)", k);
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    const Fence fx = make_fence(fence, num_half_edges);
    const Fence cfx = make_fence(cum ? cum_fence : nullptr, num_half_edges);
    const int mode = search_mode(fence, num_half_edges);
#define LAUNCH_RANDOM_T(MODE) sample_random_kernel<LR, true, MODE><<<grid_, block_, smem_, s>>>(he, indptr, num_nodes, fx, cum, cfx, node_ids, times, n, k, seed, offset, out_nbr, out_eid, out_t)
#define LAUNCH_RANDOM_F(MODE) sample_random_kernel<LR, false, MODE><<<grid_, block_, smem_, s>>>(he, indptr, num_nodes, fx, cum, cfx, node_ids, times, n, k, seed, offset, out_nbr, out_eid, out_t)
#define DISPATCH_RANDOM(L, GRID, BLOCK, SMEM)                        \
    do {                                                             \
        constexpr int LR = L;                                        \
        const unsigned grid_ = GRID, block_ = BLOCK;                 \
        const size_t smem_ = SMEM;                                   \
        if (cum) DISPATCH_MODE(mode, LAUNCH_RANDOM_T);               \
        else DISPATCH_MODE(mode, LAUNCH_RANDOM_F);                   \
    } while (0)
    if (k <= 32) {
        {   // 4 lanes per query: uniform 6.6 ms vs 7.3 ms with 8 on the sweep, time_interval_aware 12.3 vs 13.4 ms
            const int threads = 256, groups = threads / 4;
            DISPATCH_RANDOM(4, blocks_for(n, groups), threads, (size_t)groups * k * 12);
        }
    } else {
        int groups = (int)((40 * 1024) / ((size_t)k * 12));
        if (groups > 8) groups = 8;
        if (groups < 1) groups = 1;
        DISPATCH_RANDOM(32, blocks_for(n, groups), groups * 32, (size_t)groups * k * 12);
    }
#undef DISPATCH_RANDOM
#undef LAUNCH_RANDOM_T
#undef LAUNCH_RANDOM_F
    DYG_LAUNCH_CHECK("dyg_sample_random");
    return 0;
}
