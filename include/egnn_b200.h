/* egnn_b200.h -- C-ABI of the B200-native GNN message-passing hot path.
 *
 * The reference (Adredes-weslee/elliptic-gnn-project) has no FFI: its seam is the
 * torch_geometric class surface used by src/models/gnn.py.  Each entry point below
 * names the reference call it replaces (path:line relative to the reference root, or
 * the PyG 2.5.3 routine restated in SURVEY.md Appendix A).  The Python host layer
 * (elliptic-gnn-project_b200/*.py) binds these with ctypes and wraps them in
 * torch.autograd.Function; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no host
 *     synchronisation, CUDA-graph capturable) unless documented otherwise;
 *   - memory is owned by the caller (torch's caching allocator); nothing is freed here;
 *   - return value 0 = ok; nonzero = error, text via egnn_last_error();
 *   - there is NO CPU fallback: a build without a visible sm_100 device fails at launch.
 *   - dtype codes: EGNN_F32 = 0, EGNN_BF16 = 1.  Index arrays inside the library are
 *     int32 (N, E < 2^31); the boundary edge_index is int64 like the reference's
 *     (src/data/dataset_elliptic.py:245).
 */
#ifndef EGNN_B200_H
#define EGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGNN_ABI_VERSION 2

enum { EGNN_F32 = 0, EGNN_BF16 = 1, EGNN_F64 = 2 /* egnn_p2p_allreduce only */ };

/* graph-build flags */
enum {
  EGNN_G_SYMMETRIZE = 1, /* edge list := cat([ei, ei.flip(0)])   src/train_gnn.py:319-326 */
  EGNN_G_SELF_LOOPS = 2  /* PyG add_remaining_self_loops: drop loops, append (i,i) i<N   */
};

/* SpMM modes */
enum {
  EGNN_SPMM_SUM = 0,      /* out[i] = sum_p in[col[p]]                                    */
  EGNN_SPMM_MEAN = 1,     /* out[i] = (sum_p in[col[p]]) / max(rowlen(i),1)   (SAGE fwd)   */
  EGNN_SPMM_DIV_NBR = 2,  /* out[i] = sum_p in[col[p]] / nbr_cnt[col[p]]      (SAGE bwd)   */
  EGNN_SPMM_WEIGHTED = 3  /* out[i] = sum_p rn(w[p]*in[col[p]])               (GCN)        */
};

/* activation codes for the fused elementwise kernels */
enum { EGNN_ACT_NONE = 0, EGNN_ACT_RELU = 1, EGNN_ACT_ELU = 2 };

int egnn_abi_version(void);
const char* egnn_last_error(void);
/* number of kernel launches this library has enqueued since load (for bench.py's
 * `gpu_launches`); host counter, thread-safe. */
uint64_t egnn_launch_count(void);
/* How the fp32-operand tensor-core GEMMs (3xTF32: egnn_gemm / egnn_linear_tc with EGNN_F32 operands) accumulate.
 *   1 (default): exact -- every 8-wide k-step product leaves the tensor core in a fresh TMEM buffer and is added with IEEE
 *    round-to-nearest on the CUDA cores (the tensor core's own accumulator TRUNCATES, a bias that does not average out
 *     through a training step); fp32 logits AND gradients within 2e-6 of the fp32 oracle (products wider than 128
 *     output columns run as 128-column blocks).  What fp32 training uses (src/train_gnn.py:187-209 with amp off).
 *   0: accumulate in TMEM (~2e-6 per product, inside the 1e-5 logits bar): no-grad forwards (`eval_split`,
 *     src/train_gnn.py:248-257), ~15 % faster.
 * Process-wide host state read when a GEMM is enqueued (also under CUDA-graph capture); returns the previous value. */
int egnn_set_f32_tc_exact(int exact);

/* ---------------------------------------------------------------- K1: graph build ---- */
/* Replaces: edge symmetrisation src/train_gnn.py:319-326; PyG add_remaining_self_loops,
 * gcn_norm (GCNConv.forward, GATConv.forward: every call); the COO
 * gather/scatter order that PyG's scatter_add_ implies.
 *
 * Input  ei      int64 [2,E] row-major (row 0 = src, row 1 = dst), contiguous.
 * The expanded edge list ei2 has E2 <= cap = (SYMMETRIZE ? 2E : E) + (SELF_LOOPS ? N : 0).
 * Outputs (caller-allocated, capacities in elements):
 *   info      int32 [4]   : {E2, number of out-of-range node ids, len(csr_long), len(csc_long)}
 *   csr_ptr   int32 [N+1] : rows = destinations
 *   csr_src   int32 [cap] : source of each in-edge; within a row in ORIGINAL edge order
 *   csr_eid   int32 [cap] : index into ei2 of that edge
 *   csc_ptr   int32 [N+1] : rows = sources
 *   csc_dst   int32 [cap] : destination of each out-edge, original order within a row
 *   csc_pos   int32 [cap] : position of that edge in CSR order (addresses per-edge arrays)
 *   csr_long  int32 [cap/64+1] or NULL : destination rows with more than 64 in-edges (any order)
 *   csc_long  int32 [cap/64+1] or NULL : source rows with more than 64 out-edges (any order)
 *   csr_order int32 [N] or NULL : destination rows sorted by DESCENDING in-degree (stable); the
 *   csc_order int32 [N] or NULL   SpMM walks rows in this order (longest rows first: no tail)
 *   ei2       int64 [2,cap] or NULL : the expanded edge list itself (parity checks)
 *   dis       float [N]   or NULL : deg^-1/2 = rn(1/rn(sqrt(deg))), 0 where deg == 0
 *   w_edge    float [cap] or NULL : gcn_norm weights in ei2 order
 *   w_csr     float [cap] or NULL : same weights in CSR order
 *   w_csc     float [cap] or NULL : same weights in CSC order
 * The four float outputs are only written when EGNN_G_SELF_LOOPS is set or want_norm!=0.
 */
size_t egnn_graph_workspace_bytes(int64_t n_nodes, int64_t n_edges_in, int flags);
int egnn_graph_build(const int64_t* ei, int64_t n_edges_in, int64_t n_nodes, int flags,
                     int want_norm, int32_t* info, int32_t* csr_ptr, int32_t* csr_src,
                     int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_dst, int32_t* csc_pos,
                     int32_t* csr_long, int32_t* csc_long, int32_t* csr_order, int32_t* csc_order,
                     int64_t* ei2, float* dis, float* w_edge, float* w_csr, float* w_csc,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Hub ablation on the device (src/train_gnn.py:526-540, src/analysis/hub_ablation.py:56-71):
 *   deg = bincount(src) + bincount(dst); hubs = the num_hubs nodes of largest degree; kept = edges touching no hub,
 *   ORIGINAL order preserved.  ei int64 [2, E] contiguous; ei_out int64 [2, E] (row stride E, first info[0] columns
 *   valid); info int32 [2] = {edges kept, out-of-range node ids}; hub_mask uint8 [N]; deg_out int32 [N] or NULL.
 *   Ties at the k-th largest degree are resolved towards the LOWER node id (a stable descending sort); torch.topk
 *   leaves them unspecified, so the hub set equals the reference's whenever the k-th and (k+1)-th degrees differ.
 * egnn_edge_gather: out[:, j] = ei[:, idx[j]]  (random edge drop `edge_index[:, perm[drop_count:]]`,
 *   src/analysis/robustness.py:65-82; out int64 [2, n_idx]); n_bad counts indices outside [0, E). */
size_t egnn_hub_ablation_workspace_bytes(int64_t n_nodes, int64_t n_edges);
int egnn_hub_ablation(const int64_t* ei, int64_t n_edges, int64_t n_nodes, int64_t num_hubs, int64_t* ei_out,
                      int32_t* info, uint8_t* hub_mask, int32_t* deg_out, void* workspace, size_t workspace_bytes,
                      void* stream);
int egnn_edge_gather(const int64_t* ei, int64_t n_edges, const int64_t* idx, int64_t n_idx, int64_t* out,
                     int32_t* n_bad, void* stream);

/* ---------------------------------------------------------------- mini-batch sampling -- */
/* Replaces: torch_geometric.loader.NeighborLoader(data, num_neighbors=fanout, batch_size=..., input_nodes=...) as used by
 * src/train_gnn.py:329-348 and consumed by train_epoch_minibatch (:212-245): one sampled subgraph per batch of seed nodes.
 * Per hop h (fan-out fanouts[h], < 0 = all): every node added in the previous hop takes all its in-neighbours if it has
 * at most fanouts[h] of them, else fanouts[h] distinct ones (Robert Floyd's algorithm over the positions of its row of
 * the CSR-by-destination view, Philox4x32-10 keyed on (seed, batch_idx, hop, local id, draw)); sampled sources join the
 * node list in order of first appearance; local id = position in the node list (seeds first, in the given order).
 *   csr_ptr / csr_src / csr_eid  int32: the CSR-by-destination view of egnn_graph_build (stable: in-edges in original order)
 *   seeds int64 [batch] DISTINCT node ids (device);  fanouts int32 [n_hops] on the HOST, n_hops <= 8
 *   state int32 [2 * n_nodes]: set once by egnn_neighbor_sample_state_init, left clean by every call
 *   n_id int64 [cap_nodes], edge_index int64 [2, cap_edges] (row stride cap_edges; row 0 = source, row 1 = destination,
 *   LOCAL ids; edges ordered by hop, then frontier node, then draw), e_id int64 [cap_edges] (optional: original edge
 *   column of every sampled edge), capacities from egnn_neighbor_sample_caps;
 *   counts int32 [2 * n_hops + 2]: nodes after the seeds / every hop, then edges likewise; info int32 [2] = totals.
 * Everything stays on the device (no synchronisation); the caller reads `info` to slice the outputs.
 * egnn_gather_rows: out[i, :] = in[idx[i], :] for rows of row_bytes bytes (batch.x / y / timestep / masks). */
int egnn_neighbor_sample_caps(int64_t n_nodes, int64_t n_edges, int64_t batch, const int32_t* fanouts, int n_hops,
                              int64_t* cap_nodes, int64_t* cap_edges);
size_t egnn_neighbor_sample_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t batch, const int32_t* fanouts,
                                            int n_hops);
int egnn_neighbor_sample_state_init(int32_t* state, int64_t n_nodes, void* stream);
int egnn_neighbor_sample(const int32_t* csr_ptr, const int32_t* csr_src, const int32_t* csr_eid, int64_t n_nodes,
                         int64_t n_edges, const int64_t* seeds, int64_t batch, const int32_t* fanouts, int n_hops,
                         uint64_t seed, uint64_t batch_idx, int32_t* state, int64_t* n_id, int64_t cap_nodes,
                         int64_t* edge_index, int64_t* e_id, int64_t cap_edges, int32_t* counts, int32_t* info,
                         void* workspace, size_t workspace_bytes, void* stream);
int egnn_gather_rows(const void* in, int64_t ld_in_bytes, const int64_t* idx, int64_t n, int64_t row_bytes, void* out,
                     int64_t ld_out_bytes, void* stream);

/* ---------------------------------------------------------------- ingestion --------- */
/* The Elliptic tables -> the device edge list (src/data/dataset_elliptic.py:190-245; the reference walks a Python
 * dict per edge endpoint on the host):
 *   tx_ids int64 [N] in CSV row order (node i = row i); a txId occurring twice maps to its LAST row, like
 *   `{int(tx): i for i, tx in enumerate(tx_ids)}` (:195-196); timestep int64 [N];
 *   e_src_tx / e_dst_tx int64 [E_raw]: the raw txId pairs of elliptic_txs_edgelist.csv (:198-219).
 *   An edge survives iff both endpoints are known txIds (:221-232) and lie in the same timestep (:235-241);
 *   survivors keep their CSV order.  edge_index_out int64 [2, E_raw] (row stride E_raw, first info[0] columns valid),
 *   node indices; info int32 [3] = {kept_in_graph, mapped (both endpoints known), extra rows of duplicated txIds}.
 * egnn_temporal_masks: make_temporal_masks (:268-290); train_window_k < 0 = None; masks uint8 [N] (0/1). */
size_t egnn_txid_join_workspace_bytes(int64_t n_nodes, int64_t n_edges_raw);
int egnn_txid_join(const int64_t* tx_ids, const int64_t* timestep, int64_t n_nodes, const int64_t* e_src_tx,
                   const int64_t* e_dst_tx, int64_t n_edges_raw, int64_t* edge_index_out, int32_t* info,
                   void* workspace, size_t workspace_bytes, void* stream);
int egnn_temporal_masks(const int64_t* y, const int64_t* timestep, int64_t n_nodes, int64_t t_train_end,
                        int64_t t_val_end, int64_t train_window_k, uint8_t* train_mask, uint8_t* val_mask,
                        uint8_t* test_mask, void* stream);

/* *flag = 1 when the two device buffers differ in any byte, else 0 (both 16-byte aligned).  train.HostFeed uses it
 * to rebuild the sorted views only when a submitted edge_index differs from the one they were built from (the
 * reference builds nothing per step: `data.to(device)` once, src/train_gnn.py:350). */
int egnn_buffers_differ(const void* a, const void* b, int64_t n_bytes, int* flag, void* stream);

/* ---------------------------------------------------------------- K2/K3: SpMM --------- */
/* Deterministic segmented gather-reduce over a sorted view (CSR for forward, CSC for the
 * transposed backward).  One sub-warp per row; lanes split the FEATURE axis only; a row's
 * edges are accumulated sequentially in stored order in fp32, so fp32 results are bitwise
 * those of the CPU scatter_add_ path.  Replaces PyG MessagePassing.propagate for SAGEConv
 * (index_select + scatter mean, A.2), GCNConv (A.1) and their autograd backward.
 *   ptr [n_rows+1], col [nnz]; in [n_in_rows, ld_in]; out [n_rows, ld_out]
 *   w        float [nnz]   (WEIGHTED) else NULL
 *   nbr_cnt  int32 ptr array [n_in_rows+1] of the OTHER view (DIV_NBR: divides a gathered
 *            row c by max(nbr_ptr[c+1]-nbr_ptr[c],1)) else NULL
 *   long_rows/n_long  (both NULL or both set) list of rows with more than 64 entries and its
 *            device-side length, from egnn_graph_build; those rows take the whole-CTA path
 *   row_order int32 [n_rows] or NULL: processing order of the rows (egnn_graph_build's
 *            descending-degree order); results do not depend on it
 *   part/n_tasks  int32 [n_tasks+1] or NULL/0: cost-balanced row partition from egnn_spmm_partition;
 *            with it the streaming kernel runs (a lane group walks the entries of its consecutive rows
 *            as one stream with a shared-memory ring of cp.async copies in flight); results do not depend on it
 *   bias     float [F] or NULL, added after the reduction; act = EGNN_ACT_*
 *   accumulate != 0: out += result (one rounding in out dtype after an fp32 add);
 *   addend != NULL: out = addend + result instead (addend has the out dtype, leading dimension ld_addend)
 */
int egnn_spmm(int mode, const int32_t* ptr, const int32_t* col, const float* w,
              const int32_t* nbr_ptr, const int32_t* long_rows, const int32_t* n_long,
              const int32_t* row_order, const int32_t* part, int64_t n_tasks, const void* in,
              int in_dtype, int64_t ld_in, void* out, int out_dtype, int64_t ld_out, int64_t n_rows,
              int64_t n_feat, const float* bias, int act, int accumulate, const void* addend,
              int64_t ld_addend, void* stream);

/* Row partition for the streaming SpMM kernel: cost(r) = r + 2 * ptr[r]; part[k] = first row whose cost
 * prefix reaches 32 * k (k = 0..n_tasks; entries past the end of the matrix hold n_rows), so every task
 * [part[k], part[k+1]) carries about the same work whatever the degree skew (the reference's scatter has no
 * such notion; this replaces the dynamic load balancing of its atomics-based `scatter_add_`).
 * egnn_spmm_partition_tasks(n_rows, nnz_cap) = n_tasks to allocate for a matrix with at most nnz_cap entries. */
int64_t egnn_spmm_partition_tasks(int64_t n_rows, int64_t nnz_cap);
int egnn_spmm_partition(const int32_t* ptr, int64_t n_rows, int32_t* part, int64_t n_tasks, void* stream);

/* ---------------------------------------------------------------- K6: dense ----------- */
/* C[M,N] (+)= op(A)[M,K] . op(B)[K,N] (+ bias[N]).  Element (m,k) of op(A) is at
 * A[m*a_sm + k*a_sk], element (k,n) of op(B) at B[k*b_sk + n*b_sn]; C row-major with ld_c.
 * Covers Linear forward (x W^T), dgrad (g W) and wgrad (g^T x) of torch_geometric's
 * nn.dense.Linear / torch.nn.Linear (src/models/gnn.py:141-144 and the convs' lin*).
 * fp32 accumulate always.  row_div_ptr (optional CSR row pointer, int32 [M+1]): row m of the
 * result is divided by max(ptr[m+1]-ptr[m], 1) -- the 1/deg of SAGE's mean backward, fused; only
 * columns [0, row_div_cols) are divided when row_div_cols > 0 (the [dm | dx_root] concatenated dgrad).
 * impl: 0 = auto, 1 = SIMT FFMA (exact-fp32 parity path), 2 = tcgen05 only (bf16 operands, either
 * both contiguous along the contraction -- forward/dgrad -- or both contiguous along the other
 * dimension with fp32 dense C -- wgrad; error if the shape is not supported).
 * split_k > 1 (and every M <= 8 reduction) needs egnn_gemm_workspace_floats(...) floats of
 * workspace; partials are reduced in a fixed order (deterministic).
 */
size_t egnn_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int split_k);
int egnn_gemm(const void* A, int a_dtype, int64_t a_sm, int64_t a_sk, const void* B, int b_dtype,
              int64_t b_sk, int64_t b_sn, void* C, int c_dtype, int64_t ld_c, int64_t M,
              int64_t N, int64_t K, const float* bias, const int32_t* row_div_ptr, int64_t row_div_cols,
              int accumulate, int split_k, float* workspace, int impl, void* stream);

/* The layer GEMM of the fused SAGE / SAGE-ResBN step: C[M,N] = A[M,K] . W[N,K]^T on the tcgen05 kernel (bf16 A and W,
 * both contiguous along K, 16-byte rows; K <= 512, 8 <= N <= 256; C fp32 or bf16) with the epilogues the layers need:
 *   bias [N], row_div_ptr / row_div_cols and accumulate as in egnn_gemm;
 *   addend (same dtype as C, leading dimension ld_addend): C[m, addend_col0 + j] += addend[m, j] for the columns from
 *     addend_col0 (a multiple of 16) on -- the identity-residual gradient `+ h_in` of src/models/gnn.py:192 joining the
 *     root half of the concatenated dgrad [dm/deg | dx_root];
 *   colstats (float [egnn_linear_stats_parts(M), 2, colstats_cols], colstats_cols <= 64): per-CTA
 *     partial column sums and sums of squares of the values AS STORED in columns [0, colstats_cols) -- the batch
 *     statistics of nn.BatchNorm1d over all rows (src/models/gnn.py:134,189) without another pass over the layer
 *     output.  egnn_colstats_reduce sums the parts in a fixed order (float64 [2, F]; the multi-GPU path all-reduces
 *     that vector), egnn_bn_finalize_parts does the same and finalises mean / rstd / running buffers in one launch.
 * Compiled epilogue combinations: everything egnn_gemm accepts; addend (+ row_div); colstats (+ bias).  Any other
 * shape or combination is an error (the callers then use egnn_gemm + egnn_colreduce). */
int64_t egnn_linear_stats_parts(int64_t M);
/* ab_dtype = EGNN_F32: fp32 operands through the 3xTF32 kernel (a = a_hi + a_lo in tf32; a.w ~= a_lo.w_hi + a_hi.w_lo +
 * a_hi.w_hi with fp32 accumulate: ~2e-6 relative, inside the fp32 parity bar of 1e-5 that plain TF32 would break);
 * needs workspace = egnn_linear_tc_workspace_floats(EGNN_F32, N, K) floats (the hi / lo split of W).  K % 4 == 0. */
size_t egnn_linear_tc_workspace_floats(int ab_dtype, int64_t N, int64_t K);
int egnn_linear_tc(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int c_dtype, int64_t ld_c,
                   int64_t M, int64_t N, int64_t K, const float* bias, const int32_t* row_div_ptr,
                   int64_t row_div_cols, int accumulate, const void* addend, int64_t ld_addend,
                   int64_t addend_col0, float* colstats, int64_t colstats_cols, int ab_dtype, float* workspace,
                   void* stream);
/* Weight gradient of the concatenated SAGE GEMM on the tcgen05 wgrad kernel: dW[N_out, K_in] = G[M, N_out]^T . X[M, K_in]
 * (bf16 operands, rows 16-byte aligned, N_out <= 256, K_in <= 384; deterministic per-CTA partials + fixed-order sum),
 * written as TWO dense fp32 parameter gradients: columns [0, split_col) -> dst0, [split_col, K_in) -> dst1 (NULL: dropped),
 * each [N_out, valid_cols] with the zero-padding columns >= valid_cols dropped -- d lin_l.weight / d lin_r.weight of
 * `SAGEConv` (src/models/gnn.py:125-128) straight into their gradient buffers.
 * G2 (optional, [M, N2] bf16; then N_out % 64 == 0 and K_in = 2 * split_col): a second gradient matrix sharing the pass
 * over X -- the residual-projection gradient `res_projs[li]` of SAGE-ResBN (src/models/gnn.py:141-144,192), whose
 * weight gradient needs only the root half of X: dst2 [N2, valid_cols] = (G2^T X)[:, split_col:].
 * workspace: egnn_wgrad_tc_workspace_floats(N_out + N2, K_in) floats. */
size_t egnn_wgrad_tc_workspace_floats(int64_t N_out, int64_t K_in);
int egnn_wgrad_tc(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M, int64_t N_out, int64_t K_in,
                  float* dst0, float* dst1, int64_t split_col, int64_t valid_cols, const void* G2, int64_t ldg2,
                  int64_t N2, float* dst2, int dtype /* EGNN_BF16 | EGNN_F32 (3xTF32, G2 must be NULL) */,
                  float* workspace, void* stream);
int egnn_colstats_reduce(const float* parts, int64_t n_parts, int64_t n_feat, double* sums, void* stream);
int egnn_bn_finalize_parts(const float* parts, int64_t n_parts, int64_t n_feat, double count, float eps,
                           float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                           int64_t* num_batches_tracked /* optional device counter, += 1 */, void* stream);
/* out = [a | b] (fp32 vectors): [lin_l.weight ; lin_r.weight] of the logits layer, one launch */
int egnn_concat2_f32(const float* a, int64_t na, const float* b, int64_t nb, float* out, void* stream);
/* out[i] = (float)in[i]: the float64 statistics / reductions handed to fp32 parameter gradients */
int egnn_f64_to_f32(const double* in, float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------- narrow-output SAGEConv -- */
/* The `hidden -> num_classes` SAGEConv of every SAGE net (src/models/gnn.py:44,128: PyG
 * SAGEConv(hidden, 2), A.2) evaluated project-first -- aggregation and projection commute:
 *   p = h . [W_l ; W_r]^T  [N, P = 2C];  out_i = (mean_{j->i} p_j[0:C] + b) + p_i[C:2C].
 * egnn_skinny_project : out[r, 0:P] = sum_k a[r,k] * W[p,k]     W float [P, K] row-major, P in {2,4,8}
 * egnn_sage_out_fwd   : the width-C gather + combine over the CSR-by-destination view, C in {1,2,4}
 * egnn_sage_out_bwd   : dp[j] = [ sum_{j->i} dout_i / max(deg_in(i),1) | dout_j ]  (CSC view)
 *                       both gathers run as two uniform passes (one thread per entry, then one thread per
 *                       row summing its contiguous terms in stored order); edge_tmp = float [edge_cap, C]
 * egnn_skinny_wgrad   : dW[p,k] = sum_r dp[r,p]*a[r,k]; dsum[p] = sum_r dp[r,p] (may be NULL);
 *                       deterministic two-level reduction, workspace from ..._workspace_floats
 * egnn_skinny_dgrad   : dh[r,k] = sum_p dp[r,p] * W[p,k]
 * a / dh: dtype EGNN_F32 | EGNN_BF16, rows of K elements with leading dimension ld. */
int egnn_skinny_project(const void* a, int dtype, int64_t ld, int64_t n_rows, int64_t K,
                        const float* W, int P, float* out, void* stream);
int egnn_sage_out_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* p, const float* bias,
                      int C, float* out, int64_t n_rows, float* edge_tmp, int64_t edge_cap, void* stream);
int egnn_sage_out_bwd(const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csr_ptr,
                      const void* dout, int dtype, int C, float* dp, int64_t n_rows, float* edge_tmp,
                      int64_t edge_cap, void* stream);
/* The same project-first evaluation for the narrow GCNConv (`GCNConv(hidden, 2)`, src/models/gnn.py:23):
 *   p = h W^T [N, C];  out_i = sum_{j->i} rn(w_ji * p_j) + b  over the self-loop CSR with the gcn_norm weights;
 *   backward dp = [ sum_{j->i} rn(w_ji * dout_i) | dout_j ]  over the CSC view (weights in CSC order). */
int egnn_gcn_out_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* w_csr, const float* p,
                     const float* bias, int C, float* out, int64_t n_rows, float* edge_tmp, int64_t edge_cap,
                     void* stream);
int egnn_gcn_out_bwd(const int32_t* csc_ptr, const int32_t* csc_dst, const float* w_csc, const void* dout, int dtype,
                     int C, float* dp, int64_t n_rows, float* edge_tmp, int64_t edge_cap, void* stream);
size_t egnn_skinny_wgrad_workspace_floats(int64_t n_rows, int64_t K, int P);
int egnn_skinny_wgrad(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows,
                      int64_t K, float* dW, float* dsum, float* workspace, void* stream);
/* egnn_skinny_wgrad with the result written as the parameter gradients of the project-first logits layer: rows
 * [0, P/2) of dW -> dW_lo (d lin_l.weight), rows [P/2, P) -> dW_hi (d lin_r.weight), their dsum -> dsum_hi (the bias
 * gradient sum_r dout[r, :], optional). */
int egnn_skinny_wgrad_split(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows, int64_t K,
                            float* dW_lo, float* dW_hi, float* dsum_hi, float* workspace, void* stream);
int egnn_skinny_dgrad(const float* dp, const float* W, int P, void* dh, int dtype, int64_t ld,
                      int64_t n_rows, int64_t K, void* stream);

/* SAGE layer operand for the concatenated GEMM (ops.SageConvFn): one launch builds the bf16 matrix
 * [[W_l | W_r], [0 | W_res]] (each block zero-padded from K to K_padded columns; W_res optional, n_res rows)
 * and the matching bias [b_l | 0] from the fp32 parameters lin_l.weight / lin_r.weight / res_projs.weight
 * (src/models/gnn.py:125-128,141-144). */
int egnn_pack_sage_weights(const float* w_l, const float* w_r, const float* w_res, const float* b_l,
                           int64_t n_out, int64_t n_res, int64_t K, int64_t K_padded, void* out_bf16,
                           float* bias_out, void* out_t_bf16 /* optional [2*K_padded, n_out] = [W_l | W_r]^T */,
                           int out_dtype /* EGNN_BF16 | EGNN_F32: element type of the two packed matrices */,
                           void* stream);

/* cast / copy with optional column padding: out[r, 0:F] = in[r, 0:F], out[r, F:ld_out] = 0 */
int egnn_cast(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype,
              int64_t ld_out, int64_t n_rows, int64_t n_feat, void* stream);

/* SAGEResBNNet._inject_time (src/models/gnn.py:168-179) / scalar-time append
 * (src/train_gnn.py:314-317): out[r] = [x[r, 0:F] | table[clamp(t[r]-1, 0, T-1), 0:D] | 0-pad].
 * table is float [T, D] (sin/cos LUT computed by the host with the reference's own torch
 * ops, or the learned nn.Embedding weight).  out_f32 / out_bf16 may each be NULL; the bf16 copy has
 * its own leading dimension ld_out_bf16 (0 = ld_out). */
int egnn_inject_time(const float* x, int64_t ld_x, const int64_t* t, const float* table,
                     int64_t T, int64_t D, float* out_f32, void* out_bf16, int64_t ld_out,
                     int64_t ld_out_bf16, int64_t n_rows, int64_t n_feat, void* stream);

/* Gradient of the learned time-embedding table (`nn.Embedding(max_timestep, dim)` backward, src/models/gnn.py:152,
 * 172-176): dtab[r, d] = sum over rows n with clamp(t[n]-1, 0, T-1) == r of dout[n, col0 + d]  (float [T, D]).
 * Deterministic (fixed-order float64 partial sums; torch's index_add_ uses float atomics on CUDA).
 * workspace: egnn_embed_grad_workspace_bytes(n_rows, T, D). */
size_t egnn_embed_grad_workspace_bytes(int64_t n_rows, int64_t T, int64_t D);
int egnn_embed_grad(const void* dout, int dtype, int64_t ld, int64_t col0, int64_t D, const int64_t* t, int64_t T,
                    int64_t n_rows, float* dtab, void* workspace, void* stream);

/* column reductions over rows: sums[c] = sum_r a[r,c] (and sumsq[c] = sum_r a[r,c]^2 when
 * sumsq != NULL), deterministic two-level tree, fp64 combine.  Used for bias gradients and
 * BatchNorm batch statistics (nn.BatchNorm1d, src/models/gnn.py:134,189).
 * workspace: egnn_colreduce_workspace_bytes(n_feat). */
size_t egnn_colreduce_workspace_bytes(int64_t n_feat);
int egnn_colreduce(const void* a, int dtype, int64_t ld, int64_t n_rows, int64_t n_feat,
                   double* sums, double* sumsq, void* workspace, void* stream);

/* ---------------------------------------------------------------- K7: BN / act / dropout */
/* bn_finalize: from (sum, sumsq, count) -> mean, rstd = 1/sqrt(var_biased+eps); updates
 * running_mean/var (momentum, unbiased var) when they are non-NULL. count_total allows the
 * multi-GPU caller to pass globally reduced sums. */
int egnn_bn_finalize(const double* sums, const double* sumsq, double count, int64_t n_feat,
                     float eps, float momentum, float* mean, float* rstd, float* running_mean,
                     float* running_var, void* stream);

/* y = dropout(act(bn(z))) + res      -- src/models/gnn.py:186-192 (SAGE-ResBN hidden layer)
 *   bn(z) = (z-mean)*rstd*gamma+beta when mean != NULL, else z (+ nothing)
 *   dropout keep-mask = 16-bit lane (col % 8) of Philox4x32-10(seed; row0+r, col / 8, layer) >= floor(p * 2^16)
 *   (see egnn_dropout_mask; all 128 bits of a draw are used),
 *   kept values scaled by 1/(1-p); p == 0 disables it.  seed_off (device int64, may be NULL)
 *   is added to seed at run time so a captured CUDA graph draws a new mask on every replay
 *   (advance it with egnn_counter_add inside the graph).
 *   res may be NULL.  z/res/y share dtype `dtype`; z has leading dimension ld, res ld_res and y
 *   ld_y (0 = ld): the output may be the right half of a wider [h_agg | h] buffer.
 *   keep_bits (optional, uint8 [n_rows, n_feat/4], n_feat % 4 == 0): receives the dropout keep bits drawn
 *   (bit i of byte (r, c/4) = column c+i kept, low nibble; bit 4+i = ReLU gate of column c+i when act = ReLU); passing them to the backward kernels skips the Philox
 *   recomputation there.  The bits ARE the Philox mask of egnn_dropout_mask.
 *   proj_w (optional, float [4, n_feat]) + proj_out (float [n_rows, 4], 16-byte aligned): also emit
 *   proj_out[r, :] = y[r, :] . proj_w^T computed from y AS STORED -- the project-first evaluation of the logits layer
 *   `SAGEConv(hidden, 2)` (src/models/gnn.py:128,193; proj_w = [lin_l.weight ; lin_r.weight]) riding on the pass that
 *   produces its input (needs act = ReLU, n_feat / 8 a power of two <= 32 and 16-byte aligned rows). */
int egnn_bn_act_dropout_res_fwd(const void* z, const void* res, void* y, int dtype, int64_t ld,
                                int64_t n_rows, int64_t n_feat, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, int act,
                                float p, uint64_t seed, const int64_t* seed_off, uint32_t layer,
                                int64_t row0, int64_t ld_res, int64_t ld_y, uint8_t* keep_bits,
                                const float* proj_w, float* proj_out, void* stream);

/* backward stage 1: g = dy * keep/(1-p) * act'(.) ; sums[c] = sum g, sums_xhat[c] = sum g*xhat
 * (BatchNorm dbeta, dgamma).  stage 2: dz = gamma*rstd*(g - sum_g/n - xhat*sum_gx/n) (or g
 * when mean == NULL).  n_total = global row count (multi-GPU passes the global N).  dy and dz
 * have leading dimension ld, z has ld_z (0 = ld).  * dp (optional, float [n_rows, 4], 16-byte aligned) + dp_w (float [4, n_feat]): the incoming gradient is not read
 *   but computed, dy[r, :] = dp[r, :] . dp_w rounded to `dtype`, and WRITTEN to `dy` for the later consumers -- the
 *   input gradient of the project-first logits layer `SAGEConv(hidden, 2)` (src/models/gnn.py:128,193) folded into
 *   this pass (needs act = ReLU, keep_bits, n_feat / 8 a power of two <= 32, 16-byte aligned rows).
 * sum_g == sum_gx == NULL (same path): only the per-block partial rows are produced, workspace = double
 *   [egnn_bn_bwd_reduce_parts(n_rows, n_feat), 2, n_feat]; egnn_bn_bwd_sums_exchange reduces and all-reduces them.
 * sum_g_f32 / sum_gx_f32 (optional, same path): fp32 copies of the two sums = d beta / d gamma of BatchNorm, written
 *   where the caller keeps those parameter gradients.
 */
int64_t egnn_bn_bwd_reduce_parts(int64_t n_rows, int64_t n_feat);
int egnn_bn_act_dropout_bwd_reduce(const void* dy, const void* z, int dtype, int64_t ld,
                                   int64_t n_rows, int64_t n_feat, const float* mean,
                                   const float* rstd, const float* gamma, const float* beta,
                                   int act, float p, uint64_t seed, const int64_t* seed_off,
                                   uint32_t layer, int64_t row0, double* sum_g, double* sum_gx, void* workspace,
                                   int64_t ld_z, const uint8_t* keep_bits,
                                   const float* dp, const float* dp_w, float* sum_g_f32, float* sum_gx_f32,
                                   void* stream);
/* dz_colsum (float [n_feat], optional): column sums of the dz values written -- the gradient of the
 * conv bias that feeds the BatchNorm -- produced in the same pass; needs `workspace` of
 * egnn_colreduce_workspace_bytes(n_feat) + 8*8*n_feat bytes. */
int egnn_bn_act_dropout_bwd_apply(const void* dy, const void* z, void* dz, int dtype, int64_t ld,
                                  int64_t n_rows, int64_t n_feat, const float* mean,
                                  const float* rstd, const float* gamma, const float* beta,
                                  int act, float p, uint64_t seed, const int64_t* seed_off,
                                  uint32_t layer, int64_t row0, const double* sum_g, const double* sum_gx, double n_total,
                                  float* dz_colsum, void* workspace, int64_t ld_z, const uint8_t* keep_bits,
                                  void* stream);

/* *counter += inc on the device (one thread) */
int egnn_counter_add(int64_t* counter, int64_t inc, void* stream);

/* the keep-mask itself, uint8 [n_rows, n_feat] (for feeding the CPU oracle the same mask) */
int egnn_dropout_mask(uint8_t* mask, int64_t n_rows, int64_t n_feat, float p, uint64_t seed,
                      const int64_t* seed_off, uint32_t layer, int64_t row0, void* stream);

/* ---------------------------------------------------------------- K4/K5: GAT ---------- */
/* Fused PyG GATConv attention over the self-loop CSR:
 *   fwd: a_s[n,h] = <xs[n,h,:], att_src[h,:]>, a_d likewise (egnn_gat_scores), then per
 *        destination row: e = leaky_relu(a_s[src]+a_d[dst]), softmax over the row
 *        (max-subtracted, +1e-16), out[dst,h,:] = sum alpha*xs[src,h,:]; alpha [cap,H] is
 *        stored in CSR order for the backward.  concat=0 -> mean over heads. +bias.
 *   bwd: see SURVEY.md A.3 backward formulas.
 */
int egnn_gat_scores(const float* xs, int64_t n_rows, int H, int C, const float* att_src,
                    const float* att_dst, float* a_s, float* a_d, void* stream);
int egnn_gat_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* xs, const float* a_s,
                 const float* a_d, float negative_slope, int H, int C, int concat,
                 const float* bias, float* alpha, float* out, int64_t n_rows, void* stream);
/* bwd pass A (CSR rows = destinations): dpre[p,h] and da_d[n,h];
 * bwd pass B (CSC rows = sources): dxs[n,h,:] = sum alpha*do[dst] + da_s*att_src + da_d*att_dst,
 *            da_s[n,h]. */
int egnn_gat_bwd_dst(const int32_t* csr_ptr, const int32_t* csr_src, const float* xs,
                     const float* a_s, const float* a_d, const float* alpha, const float* dout,
                     float negative_slope, int H, int C, int concat, float* dpre, float* da_d,
                     int64_t n_rows, void* stream);
int egnn_gat_bwd_src(const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csc_pos,
                     const float* alpha, const float* dpre, const float* dout, const float* da_d,
                     const float* att_src, const float* att_dst, int H, int C, int concat,
                     float* dxs, float* da_s, int64_t n_rows, void* stream);

/* datt_src[h,c] = sum_n da_s[n,h]*xs[n,h,c]; datt_dst likewise (double [H*C] each).
 * workspace: egnn_colreduce_workspace_bytes(H*C). */
int egnn_gat_att_grad(const float* xs, const float* da_s, const float* da_d, int64_t n_rows, int H,
                      int C, double* datt_src, double* datt_dst, void* workspace, void* stream);

/* ---------------------------------------------------------------- multi-GPU ----------- */
/* One-shot all-reduce (sum) over NVLink peer memory of the small vectors the timestep-sharded step exchanges: the
 * BatchNorm statistics (nn.BatchNorm1d over ALL nodes, src/models/gnn.py:134,189; fp64) and the flat
 * weight-gradient buffer (fp32).  Each rank stores its vector into a slot of every peer's symmetric buffer,
 * publishes a system-scope flag per 2048-element chunk, waits for the peers' flags in its own memory and sums
 * the slots in rank order (bit-identical on every rank).  n <= n_max <= 131 072; dtype EGNN_F32 | EGNN_F64.
 * peer_bufs_dev: DEVICE array of `world` base pointers, entry r = rank r's buffer of
 * egnn_p2p_allreduce_buffer_bytes(world, n_max, dtype) bytes, zero-initialised, mapped into this process (the
 * Python layer gets them from torch.distributed._symmetric_memory).  epoch: device int64[2], local, zeros at start
 * ({calls completed, ticket counter}: the last block of a call advances the epoch; CUDA-graph replayable).  A peer that does not arrive within timeout_ms (<= 0: 2000; measured with %globaltimer)
 * makes the call fail LOUDLY: error_flag (optional, device int) is set to 1 and stays set, and the affected chunk of
 * `out` is filled with NaN, so nothing computed from the un-reduced vector can pass for a result.
 * in == out is allowed.  Every rank must make the same sequence of calls on a given buffer. */
size_t egnn_p2p_allreduce_buffer_bytes(int world, int64_t n_max, int dtype);
int egnn_p2p_allreduce(const void* in, void* out, int64_t n, int dtype, int64_t n_max, void* const* peer_bufs_dev,
                       int rank, int world, int64_t* epoch, int* error_flag, int64_t timeout_ms, void* stream);

/* The BatchNorm exchanges of the timestep-sharded SAGE-ResBN step as ONE single-block kernel each, producer and consumer
 * included (same buffers / flags / epoch protocol as egnn_p2p_allreduce on a float64 buffer with 2 * n_feat <= n_max):
 *   egnn_bn_stats_exchange   : reduce the layer GEMM's statistics parts (egnn_linear_tc colstats) -> push [sum, sumsq] to
 *                              every peer -> wait -> add the world's slots in rank order -> mean / rstd / running
 *                              buffers / num_batches_tracked (`count` = GLOBAL row count).  Replaces
 *                              egnn_colstats_reduce + egnn_p2p_allreduce + egnn_bn_finalize.
 *   egnn_bn_bwd_sums_exchange: reduce the backward partial rows -> exchange -> sums double [2, n_feat] = [sum g, sum g*xhat]
 *                              over ALL ranks; the fp32 outputs receive THIS rank's share (d beta / d gamma before
 *                              the weight-gradient all-reduce, which adds the ranks' shares like any other gradient).
 * A peer that does not arrive: NaN results and the sticky error flag, as egnn_p2p_allreduce.
 * `epoch` of all three entry points: device int64[2] = {calls completed, ticket of the running call}, both 0 at start. */
int egnn_bn_stats_exchange(const float* parts, int64_t n_parts, int64_t n_feat, double count, float eps, float momentum,
                           float* mean, float* rstd, float* running_mean, float* running_var,
                           int64_t* num_batches_tracked, int64_t n_max, void* const* peer_bufs_dev, int rank, int world,
                           int64_t* epoch, int* error_flag, int64_t timeout_ms, void* stream);
int egnn_bn_bwd_sums_exchange(const double* partial, int64_t n_parts, int64_t n_feat, double* sums, float* sum_g_f32,
                              float* sum_gx_f32, int64_t n_max, void* const* peer_bufs_dev, int rank, int world,
                              int64_t* epoch, int* error_flag, int64_t timeout_ms, void* stream);

/* ---------------------------------------------------------------- step tail ----------- */
/* Masked weighted cross-entropy over precomputed train-row indices:
 *   loss = (1/n_total) * sum_i cw[y_i] * (logsumexp(l_i) - l_i[y_i])      (2 classes)
 * i.e. F.cross_entropy(weight=cw, reduction='none').mean()  (src/train_gnn.py:159-176).
 * Writes dlogits [n_rows, 2] (zero outside the train rows) in `dtype`, and loss (float[1]).
 * workspace: float[ egnn_ce_workspace_floats(n_idx) ]. */
size_t egnn_ce_workspace_floats(int64_t n_idx);
int egnn_masked_ce(const void* logits, int dtype, int64_t n_rows, const int64_t* y,
                   const int64_t* idx, int64_t n_idx, const float* cw, double n_total,
                   float* loss, void* dlogits, float* workspace, void* stream);
/* `_make_loss_fn` in full (src/train_gnn.py:136-183): focal_gamma >= 0 selects the focal branch ((1-p_y)^gamma * CE,
 * no class weights, :153-158), < 0 the class-weighted CE; time_scheme 0 none / 1 linear / 2 sqrt multiplies each row's
 * loss by clamp(f((t - t_min) / max(t_max - t_min, 1)), 1e-3) (:165-174); idx NULL = rows 0..n_idx-1 (the reference
 * hands the loss already-masked logits).  Mean over n_total rows; dlogits as for egnn_masked_ce.
 * egnn_l2_mean_penalty: *loss += lambda * mean(w^2), grad += 2 lambda / n * w (learned time table L2, :178-180). */
int egnn_masked_loss(const void* logits, int dtype, int64_t n_rows, const int64_t* y, const int64_t* idx, int64_t n_idx,
                     const float* cw, double n_total, double focal_gamma, const int64_t* timestep, double t_min,
                     double t_max, int time_scheme, float* loss, void* dlogits, float* workspace, void* stream);
int egnn_l2_mean_penalty(const float* w, int64_t n, double lambda, float* loss, float* grad, void* stream);


/* Global-norm clip + Adam (coupled L2) over one flat fp32 parameter/grad buffer:
 * torch.nn.utils.clip_grad_norm_(params, max_norm) then torch.optim.Adam.step()
 * (src/train_gnn.py:203-207,357-359).  sqnorm_partial: float[egnn_adam_workspace_floats(n)].
 * step_count: device int64 scalar incremented by the kernel (graph-replay safe). */
size_t egnn_adam_workspace_floats(int64_t n);
int egnn_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                        int64_t n, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float max_norm, int64_t* step_count, float* grad_norm_out,
                        float* workspace, void* stream);

/* ---------------------------------------------------------------- epoch tail -- */
/* Validation PR-AUC on the device.  Replaces, per epoch, `eval_split`'s softmax -> .cpu().numpy() -> mask indexing
 * (src/train_gnn.py:248-257) and `pr_auc_illicit` = sklearn `average_precision_score` (src/utils/metrics.py:11-13):
 *   AP = sum over distinct score thresholds k, descending, of (R_k - R_{k-1}) * P_k  (ties share a threshold).
 * Exactly one of `logits` (float [n_rows, >=2], leading dimension ld_logits; score = softmax(row)[1]) or `scores`
 * (float [n_rows]) is given; y int64 [n_rows] (positive class: y == 1, as the caller's `(y_val == 1).astype(int)`,
 * src/train_gnn.py:391); mask uint8/bool [n_rows] or NULL (all rows).  scores_out (float [n_rows] or NULL) receives
 * the scores.  out double[8] (device) = {AP, selected rows, positives, distinct thresholds, ROC-AUC, 0, 0, 0}; no
 * selected row or no positive -> AP = 0.0 (the reference's `y_val.size == 0` guard / sklearn's no-positive
 * convention).  ROC-AUC = trapezoid area over the same thresholds (`roc_auc_illicit` = sklearn `roc_auc_score`,
 * src/utils/metrics.py:15-16, final metrics src/train_gnn.py:449-470); NaN when only one class is selected.
 * Counts are exact integers; AP is a fixed-order float64 sum (deterministic). */
size_t egnn_ap_workspace_bytes(int64_t n_rows);
int egnn_average_precision(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                           const uint8_t* mask, int64_t n_rows, float* scores_out, double* out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Final metrics of the reference's run tail (src/train_gnn.py:449-470) from the same sorted run:
 * out double[16]: [0..7] as egnn_average_precision, then
 *   [8]  max F1 over the precision-recall curve, [9] its threshold   (pick_threshold_max_f1, src/utils/metrics.py:22-27;
 *        ties -> the lowest threshold, as np.nanargmax over sklearn's ascending thresholds)
 *   [10] precision among the top_k scores, [15] = min(top_k, selected rows)   (precision_at_k, :39-41)
 *   [11] max recall with precision >= target_precision                        (recall_at_precision, :43-48)
 *   [12] lowest threshold with precision >= target_precision, 1.0 if none     (pick_threshold_for_precision, :29-37)
 *   [13] F1 of the prediction `score >= thr`, thr = *threshold_dev (a DEVICE double, e.g. &out_val[9] of a previous
 *        call on the validation rows) or this call's own [9] when threshold_dev is NULL   (f1_at_threshold, :18-20)
 *   [14] expected calibration error over ece_bins (<= 32) equal-width bins    (expected_calibration_error, :50-66)
 * Counts are exact integers; float64 everywhere the reference's numpy code is float64.  Same workspace as
 * egnn_average_precision. */
int egnn_ranking_metrics(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                         const uint8_t* mask, int64_t n_rows, int64_t top_k, double target_precision,
                         const double* threshold_dev, int ece_bins, float* scores_out, double* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Temperature scaling (TemperatureScaler.fit, src/utils/calibrate.py:8-30; call site src/train_gnn.py:424-429): the T
 * minimising CrossEntropyLoss(logits / T, y) over the rows with mask != 0 and y in {0, 1}.  The reference runs LBFGS
 * on T from 1.0; here one CTA runs a damped Newton iteration on 1/T (the objective is convex in it) with float64
 * fixed-order sums.  out double[5] (device): {T, mean NLL at T = 1, mean NLL at T, iterations, rows used}. */
int egnn_temperature_fit(const float* logits, int64_t ld_logits, const int64_t* y, const uint8_t* mask,
                         int64_t n_rows, int max_iter, double* out, void* stream);

/* Early-stopping bookkeeping without a host round trip (src/train_gnn.py:392-402: `if pr_val > best_val` ->
 * best_state = CPU clone of the state dict; else bad += 1).  state double[5] (device) = {best value, epochs since
 * the best, epoch of the best (1-based), epochs seen, improved flag}; initialise to {-1, 0, 0, 0, 0} (best_val = -1.0,
 * src/train_gnn.py:375).  When the value improves, params[0..n_params) is copied to best_params (both float,
 * 16-byte aligned, may be NULL together).  patience > 0: once state[1] >= patience (`if bad >= patience: break`,
 * src/train_gnn.py:411) the call is a no-op (improved flag cleared), so epochs run past the stopping point before
 * the host polls cannot change best value or snapshot; patience <= 0: never frozen. */
int egnn_early_stop_update(const double* ap, double* state, const float* params, float* best_params,
                           int64_t n_params, int64_t patience, void* stream);
/* The rest of `best_state` (BatchNorm running statistics and counters, any tensor outside the flat parameter
 * buffer): dst <- src when the LAST egnn_early_stop_update improved the best value (state[4] != 0).  Buffers are
 * 4-byte aligned, n_bytes a multiple of 4. */
int egnn_snapshot_if_improved(const double* state, const void* src, void* dst, int64_t n_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGNN_B200_H */
