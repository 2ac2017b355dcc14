/*
 * dygb200.h -- C ABI of libdygb200.so: B200 (sm_100a) kernels for DyGLib's temporal
 * neighbour-aggregation path.
 *
 * Boundary rules:
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - no allocation and no host synchronisation inside; the caller owns every buffer;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, non-zero = error, text via dyg_last_error() (thread-local);
 *   - ids / edge ids are int64 and sampled times float32 at this boundary because that is
 *     what the reference's numpy API returns (utils/utils.py:161-167).
 *
 * The reference has no FFI of its own (it is pure Python); each entry point names the
 * reference Python symbol it replaces.  INTEGRATION.md shows the ctypes binding.
 */
#ifndef DYGB200_H
#define DYGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYG_ABI_VERSION 13

typedef void* dyg_stream_t; /* cudaStream_t */

/* One directed half of an interaction as stored in the device CSR (16 bytes, one LDG.128).
 * Node v owns half-edges [indptr[v], indptr[v+1]) sorted by t (stable), exactly the order of
 * NeighborSampler.nodes_neighbor_{ids,edge_ids,times}[v] (utils/utils.py:96-103). */
typedef struct {
    double t;
    int32_t nbr;
    int32_t eid;
} dyg_halfedge_t;

const char* dyg_last_error(void);
int dyg_abi_version(void);

/* ---- a1: CSR construction (get_neighbor_sampler + NeighborSampler.__init__, utils/utils.py:283-302, 73-110) ---- */

/* deg[v] += number of half-edges owned by v.  deg must be zeroed by the caller (num_nodes entries). */
int dyg_csr_degrees(const int64_t* src, const int64_t* dst, int64_t num_events, int64_t num_nodes,
                    int64_t* deg, dyg_stream_t stream);
/* Stable LSD radix sort of (key, value) pairs for the CSR build (utils/utils.py:96-103: sorted(..., key=time) is stable; the build
 * sorts half-edges stably by time, then stably by owner).  Keys are 4- or 8-byte unsigned bit patterns, values uint32 positions.
 * dyg_radix_digit_hist: hist[p * 256 + d] = number of keys whose byte p equals d (the caller skips passes whose digit is constant).
 * dyg_radix_sort_pass: one 8-bit pass from (keys_in, vals_in) to (keys_out, vals_out); vals_in == NULL means vals = 0..n-1;
 * workspace: dyg_radix_sort_workspace_entries(n) uint32.  No atomics decide an output position: the result is deterministic. */
int dyg_radix_digit_hist(const void* keys, int key_bytes, int64_t n, unsigned long long* hist, dyg_stream_t stream);
int64_t dyg_radix_sort_workspace_entries(int64_t n);
int dyg_radix_sort_pass(const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out, int key_bytes, int64_t n,
                        int pass, const unsigned long long* hist, uint32_t* workspace, dyg_stream_t stream);
/* order[i] = index of the i-th half-edge in CSR order; half-edge h = 2*e + side (side 0: owner src[e],
 * neighbour dst[e]; side 1: owner dst[e], neighbour src[e]).  Packs the 16-byte records. */
int dyg_csr_pack(const int64_t* order, const int64_t* src, const int64_t* dst, const int64_t* eid,
                 const double* t, int64_t num_half_edges, dyg_halfedge_t* out, dyg_stream_t stream);
/* time_interval_aware tables (compute_sampled_probabilities, utils/utils.py:112-128), one thread per node,
 * sequential float64 like numpy: prob[j] = e_j / cumsum(e)_j (NaN -> -1e10), e_j = exp(tsf*(t_j - t_last));
 * cum[j] = sum_{l<=j} exp(prob[l]) (prefix table used by the device CDF search). Either output may be NULL. */
int dyg_csr_tia_tables(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, double time_scaling_factor,
                       double* prob, double* cum, dyg_stream_t stream);
/* cum from a given prob table (used when prob was computed on the host for bit-exactness). */
int dyg_csr_tia_cum(const double* prob, const int64_t* indptr, int64_t num_nodes, double* cum, dyg_stream_t stream);

/* Fence index of the CSR (no counterpart in the reference; it replaces the O(log2 deg) probes of np.searchsorted,
 * utils/utils.py:141, by O(log16 deg) 128-byte line reads).  Level l >= 1 holds, for every complete block of 16^l records of
 * the half-edge array, the time of its last record; levels are concatenated, each starting on a multiple of 16 entries.
 * dyg_csr_fence_entries: doubles to allocate (128-byte aligned); dyg_csr_fence_build fills them.  Every query entry point
 * takes `fence` (NULL: search the records directly) and `num_half_edges` (defines the level offsets). */
int64_t dyg_csr_fence_entries(int64_t num_half_edges);
int dyg_csr_fence_build(const dyg_halfedge_t* he, int64_t num_half_edges, double* fence, dyg_stream_t stream);
/* The same index over the time_interval_aware prefix table `cum` (non-decreasing inside a node's run): level l holds
 * cum[16^l i + 16^l - 1].  Same size as the CSR fence (dyg_csr_fence_entries).  dyg_sample_random descends it instead of
 * binary-searching cum (np.searchsorted(cdf, u, 'right') inside RandomState.choice, utils/utils.py:187). */
int dyg_cum_fence_build(const double* cum, int64_t num_half_edges, double* fence, dyg_stream_t stream);

/* ---- a2: find_neighbors_before (utils/utils.py:130-147): cnt[q] = #{j : t_j < times[q]} ---- */
int dyg_count_before(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                     int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int32_t* cnt, dyg_stream_t stream);

/* ---- a3: get_historical_neighbors, strategy 'recent' (utils/utils.py:149-175, 200-209) ----
 * last min(cnt,k) entries, left-padded with zeros. out_* are (n,k) row-major. cnt may be NULL. */
int dyg_sample_recent(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                      int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int k,
                      int64_t* out_nbr, int64_t* out_eid, float* out_t, int32_t* cnt, dyg_stream_t stream);

/* ---- a4/a5: 'uniform' and 'time_interval_aware' (utils/utils.py:176-199) ----
 * Gather at caller-supplied positions sel[q*k+j] in [0,cnt[q]) (the replayed RandomState draws), then
 * re-sort each row by float32 time (ties keep draw order).  Rows with cnt[q]==0 are all zero. */
int dyg_sample_indexed(const dyg_halfedge_t* he, const int64_t* indptr, const int64_t* node_ids,
                       const int32_t* cnt, const int64_t* sel, int64_t n, int k,
                       int64_t* out_nbr, int64_t* out_eid, float* out_t, dyg_stream_t stream);
/* Device draws.  u is (n*k) float64 in [0,1) (RandomState.random_sample replay, or any stream).
 * uniform: sel = floor(u*cnt) (throughput mode; NOT the reference's rejection-sampled stream).
 * tia:     sel = searchsorted(cum[:cnt]/cum[cnt-1], u, 'right') using the prefix table (utils/utils.py:183-187). */
int dyg_draw_uniform(const int32_t* cnt, const double* u, int64_t n, int k, int64_t* sel, dyg_stream_t stream);
int dyg_draw_tia(const double* cum, const int64_t* indptr, const int64_t* node_ids, const int32_t* cnt,
                 const double* u, int64_t n, int k, int64_t* sel, dyg_stream_t stream);
/* Counter-based uniforms for the sharded throughput mode (labelled non-parity): u[i] from (seed, offset+i). */
int dyg_philox_uniform(uint64_t seed, uint64_t offset, int64_t count, double* u, dyg_stream_t stream);
/* Fused throughput path (labelled non-parity): search + Philox draw (counter offset + q*k + j) + gather + time
 * re-sort in one kernel.  cum == NULL: uniform (floor(u*cnt)); cum != NULL: time_interval_aware prefix-CDF search, through
 * cum_fence (dyg_cum_fence_build) when it is not NULL. */
int dyg_sample_random(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                      int64_t num_half_edges, const double* cum, const double* cum_fence, const int64_t* node_ids, const double* times, int64_t n, int k, uint64_t seed, uint64_t offset,
                      int64_t* out_nbr, int64_t* out_eid, float* out_t, dyg_stream_t stream);

/* ---- a7 + a12: get_all_first_hop_neighbors + DyGFormer.pad_sequences (utils/utils.py:254-273,
 * models/DyGFormer.py:196-245) ----  row q = [node_ids[q], last min(cnt,L-1) neighbours..., 0...] over
 * row_stride columns (row_stride >= L); out_t[q,0] = (float)times[q]; out_len[q] = min(cnt,L-1)+1.
 * group_max[q / group_size] = max over the group of out_len (atomicMax; caller zeroes it; may be NULL). */
int dyg_first_hop_pad(const dyg_halfedge_t* he, const int64_t* indptr, int64_t num_nodes, const double* fence,
                      int64_t num_half_edges, const int64_t* node_ids, const double* times, int64_t n, int max_input_sequence_length,
                      int row_stride, int64_t* out_nbr, int64_t* out_eid, float* out_t, int32_t* out_len,
                      int32_t* group_max, int group_size, dyg_stream_t stream);

/* ---- a13: NeighborCooccurrenceEncoder.count_nodes_appearances (models/DyGFormer.py:337-393) ----
 * out_src (B,Ls,2), out_dst (B,Ld,2) float32 counts [in src row, in dst row], zero where id==0.
 * cnt_src (2,B,Ls) / cnt_dst (2,B,Ld) int64 hold the same counts as planes [in src row | in dst row], the
 * row indices of the count-LUT gather in the fused DyGFormer path; any output may be NULL. */
int dyg_cooc_count(const int64_t* src_ids, int ld_src, const int64_t* dst_ids, int ld_dst, int64_t B, int Ls, int Ld,
                   float* out_src, float* out_dst, int64_t* cnt_src, int64_t* cnt_dst, dyg_stream_t stream);

/* ---- a8: TimeEncoder.forward (models/modules.py:27-39): out[i,c] = cos(fma(dt[i], w[c], b[c])) ---- */
int dyg_time_encode(const float* dt, int64_t n, const float* w, const float* b, int T, float* out, dyg_stream_t stream);
/* Backward of dyg_time_encode for the training path (the reference differentiates models/modules.py:37 with autograd):
 * grad_w[c] += sum_i grad_out[i, c] * (-sin(fma(dt[i], w[c], b[c]))) * dt[i], grad_b[c] += the same without dt[i];
 * grad_out is (n, ldg) row-major.  Accumulates (the caller zeroes grad_w / grad_b). */
int dyg_time_encode_bwd(const float* dt, int64_t n, const float* w, const float* b, int T, const float* grad_out, int64_t ldg,
                        float* grad_w, float* grad_b, dyg_stream_t stream);

/* ---- dense contractions with fused gathers (nn.Linear call sites of models/modules.py:155-199,65-67;
 * models/DyGFormer.py:148-157,442-461; nn.GRUCell of models/MemoryModel.py:501) ----
 * A row m is the concatenation of up to DYG_MAX_SEGS segments; segment s contributes width*group columns:
 *   column (p*width + c) = ptr[idx[m*group+p]*ld + c] (+ ptr2[idx2[m*group+p]*ld2 + c])        kind 0
 *                        = mask_ids[m*group+p]==0 ? 0 : cos(fma(dt[m*group+p], w[c], b[c]))      kind 1
 *                          (with t_query set, dt[] holds float32 neighbour times and the delta is formed here)
 * idx NULL = identity; idx2 NULL = idx.  C[row(m), :N] = act(A W^T + bias + residual[row(m)]),
 * row(m) = c_group>0 ? (m / c_group)*c_group_stride + m % c_group + c_offset : m. */
#define DYG_MAX_SEGS 4
#define DYG_ACT_NONE 0
#define DYG_ACT_RELU 1
#define DYG_ACT_GELU 2
#define DYG_ACT_SIGMOID 3 /* link probability: model[1](...).sigmoid(), train_link_prediction.py:243-244 */
typedef struct {
    int32_t kind;
    int32_t width;
    int32_t group;
    int32_t ld;
    int32_t ld2;
    int32_t tq_div;          /* kind 1 with t_query: sub-row r belongs to query r / tq_div */
    const float* ptr;
    const int64_t* idx;
    const float* ptr2;
    const int64_t* idx2;
    const float* dt;
    const int64_t* mask_ids;
    const float* w;
    const float* b;
    const double* t_query;   /* kind 1: if set, dt = (float)(t_query[r / tq_div] - (double)dt[r]) (models/DyGFormer.py:263) */
} dyg_seg_t;

int dyg_linear(const dyg_seg_t* segs_host, int nseg, const float* W, int ldw, const float* bias,
               const float* residual, int ldr, float* C, int ldc, int64_t M, int N, int act,
               int c_group, int c_group_stride, int c_offset, dyg_stream_t stream);

/* Tensor-core variant (tcgen05, BF16x3: x = hi + mid, three bf16 MMAs, fp32 accumulation in TMEM; ~1e-5 relative).
 * W_hi / W_mid: bf16 (n_pad, ldwp) row-major copies of hi = bf16(W), mid = bf16(W - hi), zero padded so that
 * ldwp % 64 == 0, ldwp >= K and n_pad >= ceil(N / tile) * tile with tile = dyg_linear_tc_tile(N).
 * Same A-segment, epilogue and row-mapping semantics as dyg_linear; every segment width must be a multiple of 4. */
int dyg_linear_tc_tile(int N);
int dyg_linear_tc(const dyg_seg_t* segs_host, int nseg, const void* W_hi, const void* W_mid, int ldwp, int n_pad,
                  const float* bias, const float* residual, int ldr, float* C, int ldc, int64_t M, int N, int act,
                  int c_group, int c_group_stride, int c_offset, dyg_stream_t stream);

/* ---- a16 on tensor cores: DyGFormer's transformer / output GEMMs (models/DyGFormer.py:442-461, 190-192) ----
 * BF16x3 operand pairs: a float x is carried as hi = bf16(x), mid = bf16(x - hi) in two bf16 planes of identical layout.
 * dyg_gemm_bf16x3: C[m,:N] = act(A W^T + bias + residual[m]),  A (M,K) planes with leading dimension lda, W (N,K) planes
 * with leading dimension ldw (elements; multiples of 8, planes 16-byte aligned); operands are fetched by TMA, the
 * product runs as three tcgen05 bf16 MMAs with fp32 accumulation in TMEM (~1e-5 relative).  Outputs: C (M,ldc) fp32
 * and / or the bf16 planes C_hi | C_mid (M,ldcs) of the same fp32 result (either may be NULL, not both). */
int dyg_gemm_bf16x3(const void* A_hi, const void* A_mid, int lda, const void* W_hi, const void* W_mid, int ldw,
                    const float* bias, const float* residual, int ldr, float* C, int ldc, void* C_hi, void* C_mid,
                    int ldcs, int64_t M, int N, int K, int act, dyg_stream_t stream);
/* C = act(LayerNorm(x; gamma, beta, eps) W^T + bias): dyg_gemm_bf16x3 with the LayerNorm of models/DyGFormer.py:447 fused in
 * (the normalised rows go straight into the resident A operand; no planes round trip).  x (M, D) fp32, D <= 224 and a multiple
 * of 8; outputs as in dyg_gemm_bf16x3; workspace as for dyg_ln_ffn_bf16x3 (dyg_ln_ffn_workspace_bytes()). */
int dyg_ln_gemm_bf16x3(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W_hi,
                       const void* W_mid, int ldw, const float* bias, float* C, int ldc, void* C_hi, void* C_mid, int ldcs,
                       int64_t M, int N, int D, int act, void* workspace, int64_t workspace_bytes, dyg_stream_t stream);
/* Feed-forward half of the transformer block fused into one kernel (models/DyGFormer.py:456-461):
 *   out = x + W2 gelu(W1 LayerNorm(x) + b1) + b2   (LayerNorm with gamma / beta / eps, exact-erf GELU).
 * x, out: (M, ld) fp32.  W1_hi | W1_mid: (Dff, ldw1) operand planes of linear_layers.0.weight (Dff x D);
 * W2_hi | W2_mid: (D, ldw2) operand planes of linear_layers.1.weight (D x Dff).  D a multiple of 8, <= 208; Dff a multiple of 32.
 * The hidden activation stays in shared / tensor memory; the normalised rows pass through `workspace`
 * (dyg_ln_ffn_workspace_bytes() bytes, 16-byte aligned, L2 resident: one shared-memory image of the A operand per CTA). */
int64_t dyg_ln_ffn_workspace_bytes(void);
int dyg_ln_ffn_bf16x3(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W1_hi,
                      const void* W1_mid, int ldw1, const float* b1, const void* W2_hi, const void* W2_mid, int ldw2,
                      const float* b2, float* out, int ldo, int64_t M, int D, int Dff, void* workspace,
                      int64_t workspace_bytes, dyg_stream_t stream);
/* hi | mid planes (M,ld) of a fp32 matrix x (M,ldx), D columns. */
int dyg_split_bf16(const float* x, int ldx, int64_t M, int D, void* hi, void* mid, int ld, dyg_stream_t stream);
/* LayerNorm(x) * gamma + beta over D (even) columns written as hi | mid planes (and as fp32 y when y != NULL)
 * (norm_layers of models/DyGFormer.py:452,458 feeding the next GEMM). */
int dyg_layernorm_split(const float* x, int ldx, const float* gamma, const float* beta, float eps, float* y, int ldy,
                        void* hi, void* mid, int ld, int64_t M, int D, dyg_stream_t stream);

/* ---- a15 + first half of a16: get_features + get_patches + the four channel projections, fused
 * (models/DyGFormer.py:247-306, 148-174) ----
 * One side = the padded sequences of the src (or dst) nodes of a batch: B rows of Lp = ntok*P positions.  Token m covers
 * positions q = m*P .. m*P+P-1 (row-major over (B, Lp)); it is written to row (m / ntok)*S + tok_off + m % ntok of X
 * (B*S, ldx), columns [ch*C, (ch+1)*C) for channel ch = node, edge, time, co-occurrence:
 *   X = bias_ch + sum_p W_ch[:, p*F_ch:(p+1)*F_ch] feat_ch(q),
 *   feat = node_tab[ids[q]], edge_tab[eids[q]], (ids[q]==0 ? 0 : cos(fma((float)(t_query[m/ntok] - (double)t_nbr[q]), tw, tb))),
 *          lut[cnt_a[q]] + lut[cnt_b[q]]   (lut[c] = the co-occurrence MLP at count c, models/DyGFormer.py:409-411).
 * Tables are BF16x3 operand planes (hi | mid, see dyg_split_bf16) with ld % 8 == 0 and ld >= roundup(F, 16), padding
 * columns zero.  W_hi | W_mid: (64, ldw) planes of the packed weights: stage s (dyg_patch_project_stages enumerates
 * them: for type in node, edge, time, lut(cnt_a), lut(cnt_b): for p: for 32-column block) owns columns [32 s, 32 s + 32),
 * rows >= C and columns past the block's valid width zero.  zero_rows: bit 0 / bit 1 set when row 0 of the node / edge
 * table (the padding id) is all zero, so padded positions need no gather. */
typedef struct {
    const int64_t* ids;    /* (B, Lp) padded neighbour ids */
    const int64_t* eids;   /* (B, Lp) padded edge ids */
    const float* t_nbr;    /* (B, Lp) padded neighbour times */
    const int64_t* cnt_a;  /* (B, Lp) appearances in the src sequence */
    const int64_t* cnt_b;  /* (B, Lp) appearances in the dst sequence */
    int64_t tokens;        /* B * ntok */
    int32_t ntok;          /* Lp / P */
    int32_t tok_off;       /* first token of this side inside a pair's S tokens */
} dyg_proj_side_t;
/* number of stages for the given widths; nblk5 (may be NULL) receives the 32-column blocks per (type, p). */
int dyg_patch_project_stages(int F_node, int F_edge, int T, int F_lut, int P, int32_t* nblk5);
int dyg_patch_project(const dyg_proj_side_t* sides_host, int nsides, const void* node_hi, const void* node_mid, int ld_node,
                      int F_node, const void* edge_hi, const void* edge_mid, int ld_edge, int F_edge, const void* lut_hi,
                      const void* lut_mid, int ld_lut, int F_lut, int zero_rows, const double* t_query, const float* tw, const float* tb,
                      int T, const void* W_hi, const void* W_mid, int ldw, const float* bias, int P, int C, int S, float* X,
                      int ldx, dyg_stream_t stream);

/* y = LayerNorm(x + r) * gamma + beta over D columns; r row = [r1 row (F1 cols) | rconst (D-F1 cols)];
 * r1/rconst may be NULL (models/modules.py:199, models/DyGFormer.py:452,458). */
int dyg_layernorm(const float* x, int ldx, const float* r1, int ldr1, int F1, const float* rconst,
                  const float* gamma, const float* beta, float eps, float* y, int ldy, int64_t M, int D,
                  dyg_stream_t stream);

/* out[m,:] = tab[idx[m],:] (+ tab2[idx[m],:]) -- the advanced-index gathers of models/TGAT.py:84,
 * models/MemoryModel.py:609 (memory + raw features). */
int dyg_gather_rows(const float* tab, int ld, const float* tab2, int ld2, const int64_t* idx, int64_t M, int D,
                    float* out, int ldo, dyg_stream_t stream);

/* ---- a9: MultiHeadAttention.forward core (models/modules.py:157-193), folded form ----
 * For root i, head h:  score_j = qk[i,h,:] . x_ij ; masked (mask_ids==0) -> -1e10 ; a = softmax_j ;
 * out_s[i,h,:] = sum_j a_j x_ij, with x_ij = [node row | edge row | time enc] built on the fly:
 *   node row  = node_tab[r]  (+ node_tab2[r]),  r = node_idx ? node_idx[i*k+j] : i*k+j
 *   edge row  = edge_tab[e],                    e = edge_idx ? edge_idx[i*k+j] : i*k+j
 *   time enc  = time_feat ? time_feat[i*k+j,:] : cos(fma((float)(t_query[i]-(double)t_nbr[i*k+j]), w, b))
 * qk already holds scaling * W_k,h^T W_q,h [x_i | cos(b)] (H*Dk floats per root, Dk = F+E+T).
 * zero_row0: bit 0 = row 0 of node_tab (and node_tab2) is all zeros, bit 1 = row 0 of edge_tab is all zeros (the padding
 * rows of the reference's tables, preprocess_data/preprocess_data.py:101-108): the kernel then skips those reads.
 * prob_scale (n,H,k) or NULL: multipliers applied to the softmax probabilities before the weighted sum (training-mode
 * dropout of the attention scores, models/modules.py:187); out_scores stay the undropped probabilities. */
int dyg_temporal_attend(const float* qk, int ldq, int64_t n, int k, int H,
                        const float* node_tab, int ld_node, const float* node_tab2, int ld_node2,
                        const int64_t* node_idx, int F,
                        const float* edge_tab, int ld_edge, const int64_t* edge_idx, int E,
                        const float* time_feat, const double* t_query, const float* t_nbr,
                        const float* w, const float* b, int T,
                        const int64_t* mask_ids, float* out_s, int lds, float* out_scores, int zero_row0,
                        const float* prob_scale, dyg_stream_t stream);
/* Backward of dyg_temporal_attend for the training path (the reference differentiates models/modules.py:157-193 with
 * autograd).  probs (n,H,k): softmax probabilities of the forward pass (its out_scores); prob_scale (n,H,k) or NULL: the
 * dropout multipliers the forward pass applied (models/modules.py:187); s_out / grad_s: forward output and its gradient.
 * Writes grad_qk (n, H*Dk); grad_nbr (n*k, F) = gradient of the neighbour node rows (NULL when they are constants);
 * ADDS the time encoder's gradients into grad_w / grad_b (T floats each, caller zeroes them; may be NULL). */
int dyg_temporal_attend_bwd(const float* qk, int ldq, int64_t n, int k, int H,
                            const float* node_tab, int ld_node, const float* node_tab2, int ld_node2,
                            const int64_t* node_idx, int F,
                            const float* edge_tab, int ld_edge, const int64_t* edge_idx, int E,
                            const double* t_query, const float* t_nbr, const float* w, const float* b, int T,
                            const int64_t* mask_ids, const float* probs, const float* prob_scale,
                            const float* s_out, int lds, const float* grad_s, int ldg,
                            float* grad_qk, int ldgq, float* grad_nbr, int ld_gn, float* grad_w, float* grad_b,
                            dyg_stream_t stream);

/* ---- a16: nn.MultiheadAttention core inside DyGFormer's TransformerEncoder (models/DyGFormer.py:454) ----
 * qkv (B,S,3*H*hd) packed [q|k|v]; out (B,S,H*hd) = softmax(q k^T / sqrt(hd)) v per head; no mask. */
int dyg_seq_attention(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, float* out, int ldo,
                      dyg_stream_t stream);
/* Tensor-core variant for S <= 128, even hd <= 128: both products as BF16x3 m16n8k16 MMAs with fp32 accumulation
 * (~1e-5 relative); writes fp32 out (B*S, ldo) and / or the bf16 hi | mid planes (B*S, ldos) that feed dyg_gemm_bf16x3. */
int dyg_seq_attention_tc(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, float* out, int ldo,
                         void* out_hi, void* out_mid, int ldos, dyg_stream_t stream);
/* The whole self-attention sub-block of the TransformerEncoder (models/DyGFormer.py:442-455: multi_head_attention + residual)
 * on tcgen05, with the output projection folded into the value projection:
 *   out = x + bias + sum_h softmax(q_h k_h^T) v'_h,   v'_h = (W_o[:, h] W_v[h]) LN(x) + W_o[:, h] b_v[h]   (D wide per head)
 * planes_hi / planes_mid: bf16 hi | mid planes (B*S, ldp) written by the projection GEMM, one row per token laid out as
 * [q_0..q_{H-1} at q_col0 | k_0..k_{H-1} at k_col0 (hd rounded up to 8 columns each, zero padded) | v'_0..v'_{H-1} (D each) at v_col0];
 * q pre-scaled by log2(e) / sqrt(hd) (the kernel uses exp2).  S <= 64, hd <= 112 (multiple of 4), D <= 208 (multiple of 8). */
int dyg_seq_attention_fold(const void* planes_hi, const void* planes_mid, int ldp, int q_col0, int k_col0, int v_col0,
                           int64_t B, int S, int H, int hd, int D, const float* x, int ldx, const float* bias, float* out,
                           int ldo, dyg_stream_t stream);
/* The same sub-block INCLUDING its LayerNorm and the [q | k | v'] projection, one kernel (models/DyGFormer.py:442-455):
 *   out = x + bout + sum_h softmax(q_h k_h^T) v'_h,   [q | k | v'] = W LayerNorm(x; gamma, beta, eps) + bcat
 * W_hi / W_mid: bf16 planes of the (N, D) projection weight in the column layout of dyg_seq_attention_fold (N = v_col0 + H*D).
 * The projected rows stay in a per-CTA scratch inside `workspace` (dyg_attn_block_workspace_bytes(N) bytes, 1024-byte aligned,
 * L2 resident); HBM traffic is x in and out out. */
int64_t dyg_attn_block_workspace_bytes(int N);
int dyg_attn_block(const float* x, int ldx, const float* gamma, const float* beta, float eps, const void* W_hi, const void* W_mid,
                   int ldw, const float* bcat, int N, int q_col0, int k_col0, int v_col0, const float* bout, int64_t B, int S,
                   int H, int hd, int D, float* out, int ldo, void* workspace, int64_t workspace_bytes, dyg_stream_t stream);
/* ---- backward kernels of DyGFormer's training path (dyg_patch_project_bwd / dyg_tfm_block_bwd;
 * train_link_prediction.py:230-257 differentiates models/DyGFormer.py:148-192, 442-461).  The forward of a training step runs on
 * the GEMMs above; its backward is composed of dX = dY W on dyg_gemm_bf16x3 and the fp32 kernels below (dyglib_b200/autograd.py). ---- */
/* dW[n, k] += sum_m G[m, n] X[m, k], db[n] += sum_m G[m, n] (db may be NULL): weight / bias gradient of y = x W^T + b, split over the rows
 * with atomic accumulation (zero dW / db first).  For the patch projections X is the gathered patch matrix of a channel. */
int dyg_gemm_dw(const float* G, int ldg, const float* X, int ldx, int64_t M, int N, int K, float* dW, int ldw, float* db,
                dyg_stream_t stream);
/* dX[m, k] = sum_n G[m, n] W[n, k]: input gradient of y = act(x W^T + b) in one launch (BF16x3 on mma.sync; fp32 operands split on the
 * way into shared memory, W read transposed by ldmatrix).  Y (M, N), when given, is the layer's ReLU output and masks G.  Very large
 * layers use dyg_gemm_bf16x3 on the planes of G and W^T instead. */
int dyg_gemm_dx(const float* G, int ldg, const float* Y, int ldy, const float* W, int ldw, int64_t M, int N, int K, float* dX, int lddx,
                dyg_stream_t stream);
/* dyg_gemm_dx and dyg_gemm_dw (+ db) of one small layer in ONE launch (either output may be NULL); Y (M, N), when given, is the layer's
 * ReLU output and masks the gradient on the way in (G * (Y > 0)).  dW / db are accumulated (zero them first). */
int dyg_linear_bwd(const float* G, int ldg, const float* Y, int ldy, const float* X, int ldx, const float* W, int ldw, int64_t M,
                   int N, int K, float* dX, int lddx, float* dW, int lddw, float* db, dyg_stream_t stream);
/* y = LayerNorm(x) gamma + beta: dx (M, D) written, dgamma / dbeta (D) accumulated (zero them first; may be NULL). */
int dyg_layernorm_bwd(const float* x, int ldx, const float* gamma, float eps, const float* dy, int lddy, float* dx, int lddx,
                      float* dgamma, float* dbeta, int64_t M, int D, dyg_stream_t stream);
/* h = gelu(v) * mask (exact erf; mask = dropout multipliers or NULL) as fp32 and / or bf16 hi | mid planes; dv = dh * mask * gelu'(v). */
int dyg_gelu_fwd(const float* v, int ldv, const float* mask, int ldm, float* h, int ldh, void* h_hi, void* h_mid, int lds,
                 int64_t M, int N, dyg_stream_t stream);
int dyg_gelu_bwd(const float* v, int ldv, const float* mask, int ldm, const float* dh, int lddh, float* dv, int lddv, int64_t M,
                 int N, dyg_stream_t stream);
/* softmax(q k^T / sqrt(hd)) v per (sequence, head) with the probabilities kept for the backward pass: probs (B, H, S, S) holds the
 * softmax, prob_mask (same shape, may be NULL) the dropout multipliers applied before the product with v (nn.MultiheadAttention's
 * attention dropout).  S <= 64, hd <= 128.  _bwd writes dqkv (B*S, 3*H*hd) from dout (B*S, H*hd). */
int dyg_seq_attention_train_fwd(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, const float* prob_mask, float* probs,
                                float* out, int ldo, dyg_stream_t stream);
int dyg_seq_attention_train_bwd(const float* qkv, int ld_qkv, int64_t B, int S, int H, int hd, const float* prob_mask,
                                const float* probs, const float* dout, int lddo, float* dqkv, int lddq, dyg_stream_t stream);

/* out[b,:] = mean over tokens [tok0, tok0+cnt) of x[b,:,:] (models/DyGFormer.py:185-187). */
int dyg_mean_tokens(const float* x, int64_t B, int S, int D, int tok0, int cnt, float* out, int ldo, dyg_stream_t stream);

/* ---- a17-a19: TGN memory path (models/MemoryModel.py:139-161, 212-251, 275-300, 435-487) ---- */
/* persist look-ahead state of the batch's nodes: memory[v]=mem_view[v], last_update[v]=lu_view[v], pending[v]=0. */
int dyg_tgn_persist(const int64_t* node_ids, int64_t n, float* memory, const float* mem_view, float* last_update,
                    const float* lu_view, uint8_t* pending, int D, dyg_stream_t stream);
/* winner[v] = max over candidates c in [0,2B) owned by v of c (c<B: src role of event c; c>=B: dst role of
 * event c-B) == the message the reference's list order leaves last (src-role appends, then dst-role appends). */
int dyg_tgn_select_last(const int64_t* src, const int64_t* dst, int64_t B, int32_t* winner, dyg_stream_t stream);
/* msg[c,:] = [memory[owner] | other_feat | cos(fma((float)t - last_update[owner], w, b)) | edge_tab[eid]],
 * other_feat = other_emb ? other_emb[c] : memory[other]  (DyRep passes embeddings). msg is (2B, 2D+T+E). */
int dyg_tgn_build_messages(const int64_t* src, const int64_t* dst, const double* t, const int64_t* eid, int64_t B,
                           const float* memory, const float* last_update, int D, const float* other_emb, int ld_other,
                           const float* edge_tab, int ld_edge, int E, const float* w, const float* b, int T,
                           float* msg, int ldm, dyg_stream_t stream);
/* GRU / RNN cell epilogue + commit for winning candidates: gi=(2B,G*D) input gates incl. bias_ih, gh=(2B,G*D)
 * hidden gates incl. bias_hh (G=3 GRU: r,z,n order of nn.GRUCell; G=1 tanh RNNCell); gi = gh = NULL when
 * dyg_gru_update_fwd has already written mem_view (then only the bookkeeping below is done).
 * mem_view[v]=h', lu_view[v]=(float)t, pending[v]=1, msg_store[v]=msg[c], msg_time[v]=t, winner[v] reset to -1. */
int dyg_tgn_cell_commit(const float* gi, const float* gh, int G, const int64_t* src, const int64_t* dst,
                        const double* t, int64_t B, int32_t* winner, const float* memory, float* mem_view,
                        float* lu_view, uint8_t* pending, int D, const float* msg, int ldm, int msg_dim,
                        float* msg_store, double* msg_time, dyg_stream_t stream);
/* Fused recurrent cell (nn.GRUCell / nn.RNNCell of models/MemoryModel.py:490-515, called at :452-453 and :481-482): in ONE
 * launch, row m = cell(msg[msg_idx ? msg_idx[m] : m, :msg_dim], h = hid[hid_idx ? hid_idx[m] : m, :D]) with the gate GEMMs
 * (w_ih (G*D, msg_dim), w_hh (G*D, D), contiguous fp32), the gate non-linearities (G=3: r, z, n order; G=1: tanh) and the
 * scatter out[out_idx ? out_idx[m] : m, :D] = h'.  winner != NULL: only rows with winner[hid_idx[m]] == m are written (the
 * last-message choice of dyg_tgn_select_last).  gates != NULL: (P, 4D) for G=3 holding r | z | n | (W_hn h + b_hn), (P, D)
 * for G=1 holding h' -- what dyg_gru_update_bwd needs.  `out` must not alias `hid`.  Widths / leading dimensions % 4 == 0. */
int dyg_gru_update_fwd(const float* msg, int ldm, const int64_t* msg_idx, int msg_dim, const float* hid, int ldh,
                       const int64_t* hid_idx, int D, const float* w_ih, const float* b_ih, const float* w_hh,
                       const float* b_hh, int G, const int32_t* winner, float* out, int ldo, const int64_t* out_idx,
                       float* gates, int64_t P, dyg_stream_t stream);
/* Element-wise half of the cell's backward (the reference differentiates models/MemoryModel.py:481-482 with autograd):
 * from the saved gates and grad_out (P, ldg) it writes the gradients of the gate pre-activations d_gi (P, G*D) (input side)
 * and d_gh (P, G*D) (hidden side) and d_h (P, D) = the direct term grad_out * z.  The caller finishes with four dense
 * products: d_msg = d_gi W_ih, d_hid = d_h + d_gh W_hh, dW_ih = d_gi^T msg, dW_hh = d_gh^T h (bias gradients: column sums). */
int dyg_gru_update_bwd(const float* gates, const float* hid, int ldh, const int64_t* hid_idx, const float* grad_out, int ldg,
                       int G, float* d_gi, float* d_gh, float* d_h, int64_t P, int D, dyg_stream_t stream);
/* reference's "Trying to update memory to time in the past" assert (models/MemoryModel.py:448-449):
 * flag[0] set to 1 if any last_update[v] > (float)msg_time for pending v in node_ids. */
int dyg_tgn_check_time(const int64_t* node_ids, int64_t n, const float* last_update, const float* lu_view,
                       const uint8_t* pending, int32_t* flag, dyg_stream_t stream);

/* ---- a17-a20 in one launch: a whole batch of the TGN / DyRep-style memory model with a 1-layer graph-attention embedding
 * (MemoryModel.compute_src_dst_node_temporal_embeddings, models/MemoryModel.py:87-168, called for the negative and the positive
 * pair of a batch as train_link_prediction.py:236-247 does, plus the MergeLayer link predictor of :243-244) ----
 * One cooperative, persistent kernel walks the dependency chain as phases separated by a grid barrier (csrc/tgn_step.cu):
 * recent-neighbour search + gather for the R roots, folded attention on the look-ahead memories, LayerNorm, MergeLayer,
 * optional link probabilities; and for the 2B candidate messages of the positive batch: time-order check, persist, last-message
 * election, raw-message build, recurrent cell, commit into the look-ahead view / message store.  Semantics are those of the
 * separate entry points above (dyg_sample_recent, dyg_gather_rows, dyg_linear, dyg_temporal_attend, dyg_layernorm,
 * dyg_tgn_check_time / persist / select_last / build_messages, dyg_gru_update_fwd, dyg_tgn_cell_commit) run in that order.
 * Root r is embedded at time t[r % B]; candidate c is the src (c < B) or dst (c >= B) role of event c % B.  F, E, T % 4 == 0, (F + T) % 8 == 0, (F + E + T) % 4 == 0,
 * (2F + T + E) % 8 == 0, F + E <= 384, T <= 128, H = 2.  Dense contractions run as BF16x3 on mma.sync tiles (csrc/mma_tile.cuh).  `barrier`: 2 uint32, zeroed once by the caller and left zero by every launch. */
/* BF16x3 operand planes of a (rows, ld) matrix: x = hi + mid, bf16 each; ld % 8 == 0, both 16-byte aligned, padding columns zero */
typedef struct {
    const void* hi;
    const void* mid;
    int64_t ld;
} dyg_planes_t;
typedef struct {
    const dyg_halfedge_t* he; const int64_t* indptr; int64_t num_nodes;              /* device CSR */
    const int64_t* src; const int64_t* dst; const double* t; const int64_t* eid;      /* the positive batch (B events) */
    const int64_t* neg;                                                               /* (B) negative destinations, or NULL */
    const int64_t* roots;                                                             /* (R) node ids to embed; NULL: [src | neg | dst] (R = 3B) or, without neg, [src | dst] (R = 2B) */
    int32_t B, R, k, H, G, check_time;
    const float* node_raw; int32_t ld_node; const float* edge_raw; int32_t ld_edge; int32_t F, E, T;
    float* memory; float* last_update; float* mem_view; float* lu_view; uint8_t* pending; int32_t* winner;
    float* msg_store; double* msg_time; int32_t* flag;
    const float* time_w; const float* time_b; const float* t0;                        /* TimeEncoder w, b and cos(b) */
    /* dense weights as operand planes, every K segment padded with zero columns to a multiple of 8 (Fp = roundup(F, 8)):
     * wqk (H*Dk, Fp) folded query weights; wvr (Dq, H*Dk) folded value / residual_fc weights; m1 (F, Dq + Fp), m2 (F, Fp) MergeLayer;
     * w_ih (G*F, 2F+T+E), w_hh (G*F, Fp) recurrent cell; p1 (F, 2 Fp) = [Wp1[:, :F] W2 | Wp1[:, F:] W2], the link predictor's fc1 folded
     * onto the MergeLayer's hidden activation (hi NULL: no predictor), p1_b = its bias + Wp1[:, :F] b2 + Wp1[:, F:] b2 */
    dyg_planes_t wqk, wvr, m1, m2, w_ih, w_hh, p1;
    const float* cq; const float* rbias; const float* ln_g; const float* ln_b; float ln_eps;
    const float* m1_b; const float* m2_b; const float* b_ih; const float* b_hh;
    const float* p1_b; const float* p2_w; const float* p2_b;                          /* predictor fc1 bias, fc2 weight (F) and bias */
    const int64_t* pair_a; const int64_t* pair_b; int32_t P;                          /* prob[i] = predictor(emb[pair_a[i]], emb[pair_b[i]]) */
    int64_t* nbr_ids; int64_t* nbr_eids; float* nbr_t;                                /* scratch (R, k) */
    float* feat; float* qk; float* o; float* msg; float* hnew; float* ph;             /* scratch fp32 (R,F), (R,H*Dk), (R,Dq), (2B,MD), (2B,F), (P,F) */
    dyg_planes_t feat_pl, msg_pl, s_pl, y_pl, h1_pl;                                  /* scratch planes (R,Fp), (2B,MD), (R,H*Dk), (R,Dq), (R,Fp); padding zeroed by the caller */
    float* emb; float* prob;                                                          /* outputs (R, F), (P) */
    uint32_t* barrier;
    unsigned long long* phase_ns;   /* NULL, or 16 slots: globaltimer at the start, after every phase and at the end (profiling) */
} dyg_tgn_step_t;
int dyg_tgn_step(const dyg_tgn_step_t* params_host, dyg_stream_t stream);
int64_t dyg_tgn_step_sizeof(void); /* sizeof(dyg_tgn_step_t), for bindings to check their mirror of the struct */

/* JODIE TimeProjectionEmbedding (models/MemoryModel.py:534-545) fused with the time-shift normalisation
 * (models/MemoryModel.py:114-118): out[m,:] = mem[ids[m],:] * (1 + ((t[m]-lu[ids[m]]-mean)/std) * w + b). */
int dyg_jodie_project(const float* mem, int ld, const float* lu, const int64_t* ids, const double* t, int64_t M, int D,
                      float mean, float stdv, const float* w, const float* b, float* out, int ldo, dyg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DYGB200_H */
