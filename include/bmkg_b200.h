/* bmkg_b200 - C-ABI of the B200-native GCL training-step kernels.
 *
 * The reference (HySonLab/BioMedKG) has no FFI: its extension seam is the
 * torch.nn.Module surface (biomedkg/model/ (all files), biomedkg/utils/fusion.py,
 * biomedkg/gcl_module.py) and every kernel below replaces a third-party library
 * call reached from that surface.  Each entry point cites the reference call
 * site it serves.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless noted;
 *   - the caller allocates every output and workspace; kernels never allocate;
 *   - return 0 (BMKG_OK) or a negative error code; nothing throws;
 *   - stateless, re-entrant, stream-ordered on `stream` (a cudaStream_t), no
 *     internal synchronisation; callable from any host thread;
 *   - bf16 tensors are row-major, 16-byte aligned; index arrays are int32 on the
 *     device, int64 edge_index [2,E] at the boundary (as PyG / the reference).
 */
#ifndef BMKG_B200_H_
#define BMKG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMKG_ABI_VERSION 4

int bmkg_abi_version(void);
/* last CUresult (or 100000 + cudaError*100 + query status) seen while building a TMA tensor map; 0 = none */
int bmkg_last_driver_status(void);
const char* bmkg_error_string(int code);
/* Bind the calling host thread to CUDA device `device` (the library carries its own static cudart; a thread that never
 * touched CUDA - e.g. an autograd worker - has no current context and would otherwise default to device 0). */
int bmkg_bind_device(int device);

/* ---- G1: edge_index -> sorted parent graph ------------------------------------------
 * Replaces the per-call COO handling inside PyG GCNConv (biomedkg/model/encoder.py:155,160).
 * Stable sort of the E edges by key (major, minor); by_src = 0 sorts by destination (CSR
 * for the forward aggregation), by_src = 1 by source (CSC for the transposed backward).
 * Outputs: major_sorted/minor_sorted/perm_sorted [E], rowptr_raw [N+1], selfsplit [N]
 * (first sorted position of row i whose minor >= i). */
size_t bmkg_edge_sort_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int bmkg_edge_sort(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int by_src, int32_t* major_sorted,
                   int32_t* minor_sorted, int32_t* perm_sorted, int32_t* rowptr_raw, int32_t* selfsplit, void* ws,
                   size_t ws_bytes, void* stream);

/* ---- G2: per-view canonical CSR ----------------------------------------------------------
 * Replaces torch_geometric.utils.dropout_edge's edge_index[:, mask] (biomedkg/model/gcl.py:42-43,76)
 * followed by gcn_norm's remove/add self-loops and degree (PyG GCNConv).
 * keep: optional uint8 [E] mask in ORIGINAL edge order (NULL = keep all).  Outputs the canonical CSR
 * of the view's edge list ei' (kept non-self edges in order, then one self-loop per node):
 * rowptr [N+1], colind [E+N capacity], optional perm [E+N] (position in ei'; needs edge_index),
 * optional dis [N] = indegree^-1/2 (fp32), optional nnz_out (device int32), optional hub_rows_out = "hub info" (device int32
 * [bmkg_hub_info_len(N,E)]: [0] = number of rows longer than 1024 edges, [1+c] = row holding edge 512*c) which the aggregation
 * kernels take as `hub_rows` to run (or skip) their split-row pre-pass. Bit-exact. */
int64_t bmkg_hub_info_len(int64_t num_nodes, int64_t num_edges);
size_t bmkg_csr_filter_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int bmkg_csr_filter(const int32_t* major_sorted, const int32_t* minor_sorted, const int32_t* perm_sorted,
                    const int32_t* rowptr_raw, const int32_t* selfsplit, const uint8_t* keep, const int64_t* edge_index,
                    int64_t num_edges, int64_t num_nodes, int32_t* rowptr, int32_t* colind, int32_t* perm, float* dis,
                    int32_t* nnz_out, int32_t* hub_rows_out, void* ws, size_t ws_bytes, void* stream);

/* ---- A1/A2: GCN aggregation -----------------------------------------------------------------
 * Replaces GCNConv.propagate + bias (+ F.relu, F.dropout of encoder.py:155-158):
 *   out[r] = dropout(relu( dis[r] * sum_{k in row r} dis[colind[k]] * x[colind[k]] + bias ))
 * x bf16 [N,C]; out bf16 or fp32 [N,C]; C % 8 == 0, C <= 1024.  With a CSC graph and no
 * epilogue this is the transposed backward.  drop_keep: optional explicit uint8 [N,C] keep mask,
 * otherwise (drop_p > 0) the counter-based stream hash(drop_seed, r*C+c) decides.
 * hub_ws (optional, bmkg_gcn_aggregate_workspace_bytes(nnz_capacity, C) bytes; nnz_capacity = E + N): enables the split-row
 * path for power-law graphs - rows longer than 1024 edges are reduced chunk-wise by whole CTAs and combined in fixed order;
 * hub_rows (optional device int32 from bmkg_csr_filter) == 0 skips the pre-pass. */
size_t bmkg_gcn_aggregate_workspace_bytes(int64_t nnz_capacity, int channels);
int bmkg_gcn_aggregate(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* x_bf16, int64_t num_nodes,
                       int channels, const float* bias, int relu, float drop_p, uint64_t drop_seed, const uint8_t* drop_keep,
                       void* out, int out_is_fp32, int64_t nnz_capacity, const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes,
                       void* stream);

/* Row-sharded variant: aggregates destination rows [row_begin, row_begin + num_rows) of the graph (rowptr/dis indexed by the
 * global row, x_bf16 = all total_rows rows, e.g. all-gathered over NVLink) into a num_rows x C output. */
int bmkg_gcn_aggregate_rows(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* x_bf16, int64_t total_rows,
                            int64_t row_begin, int64_t num_rows, int channels, const float* bias, int relu, float drop_p,
                            uint64_t drop_seed, const uint8_t* drop_keep, void* out, int out_is_fp32, int64_t nnz_capacity,
                            const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes, void* stream);

/* 1-hop "star" aggregation for the embedding-export path: biomedkg/data/node.py:193-241 calls BaseGCL.forward
 * (biomedkg/gcl_module.py:55-58) once per seed node over NeighborLoader(num_neighbors=[-1]) batches
 * (biomedkg/data_module.py:71-79).  All N one-seed star graphs are evaluated in one pass over the full-graph CSR:
 *   out[s] = dis[s] * (sum_{j->s, j!=s} leaf[j] + dis[s] * seed[s]) + bias   (optionally ReLU),   dis = indeg^-1/2,
 * leaf/seed bf16 [N,C] = the layer's linear transform of the leaf chain and of the seed chain. */
int bmkg_gcn_star_aggregate(const int32_t* rowptr, const int32_t* colind, const float* dis, const void* leaf_bf16,
                            const void* seed_bf16, int64_t num_nodes, int channels, const float* bias, int relu, void* out,
                            int out_is_fp32, void* stream);

/* ---- S1: neighbour sampling (mini-batch regime) --------------------------------------------------
 * torch_geometric NeighborLoader as configured at biomedkg/data_module.py:71-79 (num_neighbors=[-1]) and :81-99
 * ([30]*3, batch_size seeds): per hop every node added in the previous hop draws min(in-degree, fanout) in-edges without
 * replacement; sampled edges are the batch's only edges; new sources are appended in order of first appearance.
 * Graph = raw CSR by destination of bmkg_edge_sort(by_src=0): rowptr_raw [N+1], colind = minor_sorted [E] (sources),
 * eperm = perm_sorted [E] (original edge ids).  One hop = count -> pick -> relabel; fanout in [1,32] or -1 (all).
 *   bmkg_sample_count:   off int32 [F+1] = exclusive prefix of min(deg, fanout) over the frontier; off[F] = T
 *   bmkg_sample_pick:    out_src int32 [T] global source ids, out_col int64 [T] = frontier_base + frontier index (the
 *                        target's batch-local id), out_eid int64 [T] original edge ids (or NULL); counter-based draws
 *                        keyed by (seed, hop, node): deterministic and independent of the batch composition
 *   bmkg_sample_relabel: local_id int32 [N] (-1 = not in the batch; seeds preset by bmkg_sample_set_ids) and first_pos
 *                        int32 [N] (all INT32_MAX) persist across hops; new_nodes int32 [<=T] receives the newly reached
 *                        nodes in order of first appearance (local ids n_before, n_before+1, ...), *new_count their number,
 *                        out_row int64 [T] the batch-local source ids
 *   bmkg_sample_set_ids: reset = 0: local_id[nodes[i]] = i (seeds);  reset = 1: local_id[nodes[i]] = -1 (end of batch) */
size_t bmkg_sample_workspace_bytes(int64_t max_entries);
int bmkg_sample_count(const int32_t* rowptr, const int32_t* frontier, int64_t frontier_len, int fanout, int32_t* off, void* ws,
                      size_t ws_bytes, void* stream);
int bmkg_sample_pick(const int32_t* rowptr, const int32_t* colind, const int32_t* eperm, const int32_t* frontier,
                     int64_t frontier_len, int fanout, const int32_t* off, uint64_t seed, int hop, int64_t frontier_base,
                     int32_t* out_src, int64_t* out_col, int64_t* out_eid, void* stream);
int bmkg_sample_relabel(const int32_t* src, int64_t num_entries, int64_t n_before, int32_t* local_id, int32_t* first_pos,
                        int32_t* new_nodes, int32_t* new_count, int64_t* out_row, void* ws, size_t ws_bytes, void* stream);
int bmkg_sample_set_ids(const int32_t* nodes, int64_t n, int32_t* local_id, int reset, void* stream);

/* ---- A3/A4: GAT aggregation (extension - BASELINE.json configs 2 and 5) ---------------------------
 * PyG GATConv(in, out, heads=H, concat=True, negative_slope, add_self_loops=True) semantics;
 * no reference call site on the GCL path (nearest: RGAT, biomedkg/model/encoder.py:62-121).
 * xh bf16 [N, H*C] = x W^T; a_src/a_dst fp32 [N,H]; H in {1,2,4}, C % 8 == 0, H*C <= 1024.
 * bmkg_gat_aggregate also writes rowmax/rowsum [N,H] (softmax statistics) for the backward, which recomputes
 * alpha from node arrays: csr pass -> d_adst, tsum_ws [N,H]; csc pass -> dxh bf16 [N,H*C], d_asrc. */
int bmkg_gat_scores(const void* xh_bf16, const float* att_src, const float* att_dst, int64_t num_nodes, int heads, int channels,
                    float* a_src, float* a_dst, void* stream);
/* hub_ws (optional, bmkg_gat_workspace_bytes(nnz_capacity, H, C)) + hub_rows (device int32 counts from bmkg_csr_filter) enable
 * the split-row path for rows longer than 1024 edges (softmax partials per 512-edge chunk, merged in chunk order). */
size_t bmkg_gat_workspace_bytes(int64_t nnz_capacity, int heads, int channels);
int bmkg_gat_aggregate(const int32_t* rowptr, const int32_t* colind, const void* xh_bf16, const float* a_src, const float* a_dst,
                       int64_t num_nodes, int heads, int channels, float negative_slope, const float* bias, int relu,
                       float drop_p, uint64_t drop_seed, const uint8_t* drop_keep, void* out, int out_is_fp32, float* rowmax,
                       float* rowsum, int64_t nnz_capacity, const int32_t* hub_rows, void* hub_ws, size_t hub_ws_bytes,
                       void* stream);
int bmkg_gat_aggregate_bwd(const int32_t* rowptr, const int32_t* colind, const int32_t* csc_rowptr, const int32_t* csc_colind,
                           const void* xh_bf16, const void* g_bf16, const float* a_src, const float* a_dst, const float* rowmax,
                           const float* rowsum, const float* att_src, const float* att_dst, int64_t num_nodes, int heads,
                           int channels, float negative_slope, void* dxh_bf16, float* d_asrc, float* d_adst, float* tsum_ws,
                           int64_t nnz_capacity, const int32_t* hub_rows_csr, const int32_t* hub_rows_csc, void* hub_ws,
                           size_t hub_ws_bytes, void* stream);

/* ---- elementwise / small reductions -------------------------------------------------------
 * bmkg_mask_cast: torch_geometric.utils.mask_feature(mode="all") (model/gcl.py:40-41,75) fused with the
 *   fp32->bf16 cast: x fp32 [n] -> x0 (plain), x1 (keep1), x2 (keep2), any of them NULL; n % 4 == 0.
 * bmkg_modality_mean: torch.mean(x, dim=1) of gcl_module.py:47-48; x fp32 [N,M,F].
 * bmkg_relu_dropout_bwd: backward of encoder.py:155-158, g_pre = (y>0) ? g_y*scale : 0, dbias = colsum(g_pre).
 * bmkg_colsum: deterministic (optionally row-weighted) column sums of fp32 [N,C].
 * bmkg_l2norm_colsum / bmkg_center_scale / bmkg_l2norm_scale_bwd: F.normalize of PyGCL's _similarity
 *   (GCL/losses/infonce.py, called from gcl_module.py:189) fused with the sqrt(log2e/tau) scale, in the centred
 *   representation the InfoNCE kernels consume:  z_u = mu + d_u  with a common fp32 vector mu [D] and bf16 deviations.
 *     l2norm_colsum: inv_norm[u] = 1/max(|h_u|, 1e-12), colsum[c] = sum_u h[u,c] * inv_norm[u]   (caller: mu = colsum * scale / rows)
 *     center_scale : d_u = bf16(h_u * inv_norm[u] * scale - mu), a[u] = mu . d_u (fp32)
 *   mu may be any vector (all zeros = the plain normalised rows); the column mean makes d small when embeddings are similar. */
int bmkg_mask_cast(const float* x, const uint8_t* keep1, const uint8_t* keep2, int64_t n, void* x0_bf16, void* x1_bf16,
                   void* x2_bf16, void* stream);
/* backward of bmkg_mask_cast: dx = g0 + keep1*g1 + keep2*g2 (bf16 grads, any NULL) -> fp32 */
int bmkg_mask_cast_bwd(const void* g0_bf16, const void* g1_bf16, const void* g2_bf16, const uint8_t* keep1, const uint8_t* keep2,
                       int64_t n, float* dx, void* stream);
/* deterministic column sums of a bf16 [N, C = H*Ch] matrix, optionally weighted per (row, head): out1 with w1 [N,H]
 * (NULL = 1), out2 with w2 (NULL = skipped).  GAT attention-vector gradients and bf16 bias gradients. ws >= 2x colsum ws. */
int bmkg_colsum_bf16(const void* x_bf16, const float* w1, const float* w2, int64_t num_rows, int channels, int heads, float* out1,
                     float* out2, void* ws, size_t ws_bytes, void* stream);
int bmkg_modality_mean(const float* x, int64_t num_nodes, int modalities, int features, float* out_f32, void* out_bf16,
                       void* stream);
size_t bmkg_colsum_workspace_bytes(int64_t num_rows, int channels);
int bmkg_relu_dropout_bwd(const void* gy_bf16, const void* y_bf16, float scale, int64_t num_rows, int channels,
                          void* gpre_bf16, float* dbias, void* ws, size_t ws_bytes, void* stream);
int bmkg_colsum(const float* z, const float* row_weight, int64_t num_rows, int channels, float* out, void* ws, size_t ws_bytes,
                void* stream);
/* out = bf16(x - m), x fp32 [N,C], m fp32 [C]: deviation operand of a centred Linear,  x W^T = (x - 1 m^T) W^T + 1 (W m)^T,
 * used for GRACE.project (model/gcl.py:49-51) so that bf16 GEMM rounding acts on deviations, not on the common part. */
int bmkg_center_cast(const float* x, const float* m, int64_t num_rows, int channels, void* out_bf16, void* stream);
int bmkg_l2norm_colsum(const float* h, int64_t num_rows, int dim, float* inv_norm, float* colsum, void* ws, size_t ws_bytes,
                       void* stream);   /* ws >= bmkg_colsum_workspace_bytes(num_rows, dim); dim <= 1024 */
int bmkg_center_scale(const float* h, const float* inv_norm, const float* mu, int64_t num_rows, int dim, float scale,
                      void* z_bf16, float* a, void* stream);
int bmkg_l2norm_scale_bwd(const float* h, const float* inv_norm, const float* dz, int64_t num_rows, int dim, float scale,
                          float* dh, void* stream);

/* ---- D1-D3: DGI / GGD heads ---------------------------------------------------------------
 * DGI.summary (model/gcl.py:19-21), SingleBranchContrast(JSD,"G2L") (gcl_module.py:127,142),
 * GGD head (model/gcl.py:87-91) and BCEWithLogits (gcl_module.py:231-233). */
int bmkg_colmean_sigmoid(const float* z, int64_t num_rows, int channels, float* summary, void* ws, size_t ws_bytes,
                         void* stream);
int bmkg_rowdot(const float* z, const float* v, int64_t num_rows, int channels, float* out, void* stream);
int bmkg_rowdot_bwd(const float* g, const float* v, int64_t num_rows, int channels, float* dz, void* stream);
size_t bmkg_softplus_pair_workspace_bytes(int64_t n);
int bmkg_softplus_pair_sum(const float* s_pos, const float* s_neg, int64_t n, float* out, void* ws, size_t ws_bytes,
                           void* stream);
int bmkg_softplus_pair_bwd(const float* s_pos, const float* s_neg, const float* gscale, int64_t n, float* d_pos, float* d_neg,
                           void* stream);

/* ---- F1: modality-fusion attention core ---------------------------------------------------
 * F.scaled_dot_product_attention over the modality axis + mean (biomedkg/utils/fusion.py:22-29).
 * qkv bf16 [N*M, 3E] (row = [q|k|v] of one (node, modality), bias-free x W^T); qkv_bias fp32 [3E] or NULL is
 * added on load; out fp32 [N,E]; probs fp32 [N,M,M]
 * saved for the backward; M <= 4, E % 8 == 0. */
int bmkg_fusion_attn_fwd(const void* qkv_bf16, const float* qkv_bias, int64_t num_nodes, int modalities, int embed, float* out,
                         float* probs, void* stream);
int bmkg_fusion_attn_bwd(const void* qkv_bf16, const float* qkv_bias, const float* probs, const float* dout, int64_t num_nodes,
                         int modalities, int embed, void* dqkv_bf16, void* stream);

/* ---- F2: ReDAF fusion epilogue -------------------------------------------------------------
 * biomedkg/utils/fusion.py:70-90 after the transform GEMM: out[n] = mean_m relu(dropout_p(relu(t[n,m] + bias) * gate[m])),
 * gate[m] = modal_weights[m] * sigmoid(relational_context_layer(0.2 * 1)) (fusion.py:54-56,82-84) computed by the caller.
 * t bf16 [N,M,E] = bias-free x W^T; bias fp32 [E]; gate fp32 [M,E]; out fp32 [N,E]; M <= 4, E % 8 == 0.
 * Dropout: drop_keep uint8 [N,M,E] if given, else a counter hash (element i is kept iff 16-bit half (i & 1) of
 * hash_u32(drop_seed, i >> 1) >= drop_p * 2^16), else none (drop_p = 0).
 * Backward: dt bf16 [N,M,E] (gradient w.r.t. t, hence also w.r.t. t + bias) and per-CTA partial sums
 * dgate_partial fp32 [bmkg_redaf_partial_rows(N,E), M, E] whose column sums (bmkg_colsum) are d gate. */
int64_t bmkg_redaf_partial_rows(int64_t num_nodes, int embed);
int bmkg_redaf_fwd(const void* t_bf16, const float* bias, const float* gate, int64_t num_nodes, int modalities, int embed,
                   float drop_p, uint64_t drop_seed, const uint8_t* drop_keep, float* out, void* stream);
int bmkg_redaf_bwd(const void* t_bf16, const float* bias, const float* gate, const float* dout, int64_t num_nodes, int modalities,
                   int embed, float drop_p, uint64_t drop_seed, const uint8_t* drop_keep, void* dt_bf16, float* dgate_partial,
                   void* stream);

/* ---- L1: Linear layers on tcgen05 (bf16 operands, fp32 accumulate in TMEM, TMA-fed) --------------------
 * torch.nn.Linear / PyG Linear y = x W^T + b at biomedkg/utils/fusion.py:18-20,70 (q/k/v and ReDAF projections),
 * biomedkg/model/encoder.py:138-143 (GCNConv.lin) and biomedkg/model/gcl.py:49-51 (GRACE.project).
 * bmkg_linear_nt: C[M,N] = A[M,K] B[N,K]^T (+ bias[N] fp32) (then ELU if elu); A, B bf16 row-major; C bf16 or fp32 (out_f32).
 *   Forward: B = W [out,in].  Input gradient: A = dY, B = W^T [in,out] (transposed by the caller).  N % 16 == 0, K % 64 == 0
 *   (bmkg_linear_supported).  Optional GAT epilogue (att_src != NULL; N = heads*channels <= 256, bf16 output): also writes
 *   a_src[m,h] = <bf16 row m, head h, att_src[h]>, a_dst likewise - the node scores of PyG GATConv (replaces bmkg_gat_scores).
 * bmkg_linear_tn: C[N,K] = sum_m G[m,N] X[m,K] (+ addend[N,K] fp32) - the weight gradient dW = dY^T X, fp32 out; the node range is
 *   split over CTAs and the partial tiles are added in fixed order (deterministic).  N % 8 == 0, K % 64 == 0. */
int bmkg_linear_supported(int64_t rows, int out_features, int in_features);
int bmkg_linear_nt(const void* a_bf16, const void* b_bf16, const float* bias, int64_t rows, int out_features, int in_features, int elu,
                   int out_f32, void* out, const float* att_src, const float* att_dst, int heads, float* a_src, float* a_dst,
                   void* stream);
size_t bmkg_linear_tn_workspace_bytes(int64_t rows, int out_features, int in_features);
int bmkg_linear_tn(const void* g_bf16, const void* x_bf16, const float* addend, int64_t rows, int out_features, int in_features,
                   float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- I1/I2: fused GRACE InfoNCE (tcgen05 / TMEM / TMA) -----------------------------------
 * PyGCL DualBranchContrast(InfoNCE(tau), "L2L", intraview_negs=True) (gcl_module.py:171-173,189).
 * Operand (bmkg_center_scale): z bf16 [R, D] holds the DEVIATIONS d_u of the normalised, sqrt(log2e/tau)-scaled rows from a
 * common fp32 vector mu [D]; a fp32 [P] = mu . d_u; xab bf16 [P, 32] (bmkg_infonce_ext) = 16 + 16 extra K columns per row that make
 * the forward's tensor-core pass add a_u + a_v to d_u . d_v (one extra K = 16 MMA step per tile).  D in {64,128,192,256}.
 * Block-interleaved stacked layout with view block B (B == N: one block, i.e. [h1 rows; h2 rows]; otherwise B % 128 == 0):
 *   rows [2kB, 2kB+B) = view 1 (h1) of nodes [kB, kB+B), rows [2kB+B, 2kB+2B) = view 2 (h2) of the same nodes, k < ceil(N/B);
 *   R = bmkg_infonce_stacked_rows(N, B) = 2 B ceil(N/B); rows of nodes >= N are ZERO in z and a; P = bmkg_infonce_padded_rows(N, B)
 *   (R rounded up to 128) and a is zero beyond R.  With B = the node block of one rank, a rank's rows of both views are ONE
 *   contiguous, 128-aligned range - the unit of the row-sharded multi-GPU path.
 * fwd writes the scalar loss and state fp32 [P][4] (32-byte aligned), an opaque hand-over to bwd: per four rows
 * (q0..q3, w0..w3, t0..t3, 0, 0, 0, 0) with t_u = 1 / R''_u, R''_u = sum_{v != u} 2^(d_u.d_v + a_u + a_v), w = 2^a, q = t w
 * (zeros for padding rows); rows [r0, r1) with r0, r1 multiples of 4 occupy floats [4 r0, 4 r1);
 * bwd writes dL/dz fp32 [R, D] (valid rows only) scaled by *gscale:
 *   dZ_u = gscale ln2/2N [ sum_v P_uv d_v + mu (sum_v P_uv - 2) - 2 d_pair(u) ],
 *   P_uv = 2^(d_u.d_v + a_u + a_v) (t_u + t_v) = 2^(d_u.d_v) (q_u w_v + q_v w_u).
 * Optional E store: pass e_store (bmkg_infonce_e_store_bytes bytes, 16-byte aligned; NULL = off) to fwd and the SAME buffer to
 * bwd: the forward then also writes E = 2^(d_u.d_v + a_u + a_v) as bf16 tiles and the backward streams them back instead of
 * recomputing the similarities (16 N^2 D -> 8 N^2 D executed; 8 N^2 bytes of HBM per full-range launch - the caller decides). */
int64_t bmkg_infonce_stacked_rows(int64_t num_nodes, int64_t view_block);
size_t bmkg_infonce_e_store_bytes(int64_t num_nodes, int64_t view_block, int64_t row_begin, int64_t row_end);
int64_t bmkg_infonce_padded_rows(int64_t num_nodes, int64_t view_block);
size_t bmkg_infonce_workspace_bytes(int64_t num_nodes, int dim);
/* xab[u] = [1,1,1,a_hi,a_mid,a_lo,0.. | a_hi,a_mid,a_lo,1,1,1,0..] (a split into three bf16 pieces) for valid rows, zeros otherwise */
int bmkg_infonce_ext(const float* a, int64_t num_nodes, int64_t view_block, void* xab_bf16, void* stream);
int bmkg_infonce_fwd(const void* z_bf16, const float* a, const void* xab_bf16, int64_t num_nodes, int dim, float* loss, float* state,
                     void* e_store, void* ws, size_t ws_bytes, void* stream);
/* bwd workspace (0 = none needed): when the stacked matrix does not fit L2, or the row blocks leave most of a wave idle, the
 * recompute backward walks the columns in phases and adds the phases' partial sums in a fixed order
 * (infonce_bwd_fixup_kernel); ws may be NULL - the launch then runs as one phase (same result up to fp32 summation order). */
size_t bmkg_infonce_bwd_workspace_bytes(int64_t num_nodes, int64_t view_block, int dim, int64_t row_begin, int64_t row_end);
/* tuning: bytes of Z one column phase may span (default 40 MB, sized for the L2 of one B200 die); returns the previous value,
 * bytes <= 0 only queries.  Process-wide; changes the sizes bmkg_infonce_bwd_workspace_bytes reports. */
int64_t bmkg_infonce_set_phase_bytes(int64_t bytes);
int bmkg_infonce_bwd(const void* z_bf16, const float* state, const float* mu, const float* gscale, const void* e_store,
                     int64_t num_nodes, int dim, float* dz, void* ws, size_t ws_bytes, void* stream);
/* Row-range variants: only rows [row_begin, row_end) of the stacked matrix are processed against ALL columns.
 * row_begin % 128 == 0; row_end % 128 == 0 or row_end == R.  fwd_rows writes this range's share of the loss (the shares of
 * all ranges add up to the loss) and state for the range; bwd_rows needs state for all rows (all-gathered) and writes dz rows of
 * the range (dz is addressed with the GLOBAL row index: pass the base of an [R, D] array, or a pointer offset by
 * -row_begin * D elements for a range-local buffer). */
size_t bmkg_infonce_workspace_bytes_rows(int64_t num_nodes, int64_t view_block, int dim, int64_t row_begin, int64_t row_end);
int bmkg_infonce_fwd_rows(const void* z_bf16, const float* a, const void* xab_bf16, int64_t num_nodes, int64_t view_block, int dim,
                          int64_t row_begin, int64_t row_end, float* loss, float* state, void* e_store, void* ws, size_t ws_bytes,
                          void* stream);
int bmkg_infonce_bwd_rows(const void* z_bf16, const float* state, const float* mu, const float* gscale, const void* e_store,
                          int64_t num_nodes, int64_t view_block, int dim, int64_t row_begin, int64_t row_end, float* dz,
                          void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BMKG_B200_H_ */
