/* phc_b200.h — C ABI of libphc_b200.so, the B200 (sm_100a) kernels behind the PHC-GNN
 * hypercomplex message-passing stack.
 *
 * The reference (bayer-science-for-a-better-life/phc-gnn) has NO native layer: its "operator API"
 * is the torch.nn.Module surface of phc/hypercomplex, and the device work is delegated to ATen,
 * torch_scatter and torch_geometric.  Each entry point below cites the reference call site whose
 * device work it replaces.  Conventions:
 *   - plain C symbols, raw DEVICE pointers + sizes + cudaStream_t; no torch types;
 *   - the library owns no memory: outputs and workspaces are allocated by the caller, every op with a
 *     workspace has a *_workspace_bytes() query;
 *   - returns 0 on success, non-zero on error (1 invalid argument, 2 unsupported, 3 CUDA error);
 *     phc_last_error() returns the thread-local message.  Never throws, never synchronises, never
 *     allocates, so every call is CUDA-graph capturable;
 *   - all feature matrices are fp32 row-major [rows, width]; a hypercomplex row stores its n
 *     components contiguously: component c occupies columns c*width/n .. (c+1)*width/n-1;
 *   - indices produced by the library are int32; indices taken from the framework (edge_index,
 *     batch, integer features) are int64 as PyTorch/PyG hand them over.
 */
#ifndef PHC_B200_H
#define PHC_B200_H

#include <stddef.h>
#include <stdint.h>

#include "phc_b200_layer.h" /* phc_conv_layer: descriptor of one whole message-passing layer */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* phc_stream_t; /* == cudaStream_t */

/* activation ids (reference phc/quaternion/activations.py:134-147 get_module_activation) */
#define PHC_ACT_IDENTITY 0
#define PHC_ACT_RELU 1
#define PHC_ACT_LRELU 2
#define PHC_ACT_ELU 3
#define PHC_ACT_SELU 4
#define PHC_ACT_SWISH 5
/* neighbour reducers (PyG aggr add/mean/max/min; softmax: messagepassing.py:297-300) */
#define PHC_RED_SUM 0
#define PHC_RED_MEAN 1
#define PHC_RED_MAX 2
#define PHC_RED_MIN 3
#define PHC_RED_SOFTMAX 4
/* PHMLinear arithmetic */
#define PHC_PREC_FP32 0   /* FFMA, bit-for-bit fp32 products */
#define PHC_PREC_TF32X3 1 /* tcgen05 kind::tf32, 3-term split: fp32-class accuracy on tensor cores */
#define PHC_PREC_BF16 2   /* tcgen05 kind::f16 (bf16 operands, fp32 accumulate) */

const char* phc_last_error(void);
int phc_version(void);

/* ---- graph structure ------------------------------------------------------------------------
 * Replaces the implicit structure of torch_scatter's atomic scatter used by PyG propagate
 * (messagepassing.py:136,221,306) and of global_add_pool (pooling.py:18).
 * edge_index: int64 [2,E] contiguous (row 0 source, row 1 target, any order).
 * Outputs (int32): rowptr[N+1], col[E] (source of each slot), perm[E] (edge id of each slot) sorted
 * stably by target; rowptr_t/col_t/perm_t the same keyed by source (col_t = target).  Bit-exact with
 * a stable argsort.  status: device int, bit0 = index out of range, bit1 = batch not ascending. */
size_t phc_csr_workspace_bytes(int num_nodes, int num_edges);
int phc_csr_build(const long long* edge_index, int num_edges, int num_nodes, int* rowptr, int* col, int* perm, int* rowptr_t,
                  int* col_t, int* perm_t, void* workspace, size_t workspace_bytes, int* status, phc_stream_t stream);
/* graph_ptr[B+1] from an ascending int64 batch vector (PyG Batch.batch). */
int phc_segment_ptr_build(const long long* batch, int num_nodes, int num_graphs, int* graph_ptr, int* status, phc_stream_t stream);
int phc_narrow_int64(const long long* in, int n, int* out, phc_stream_t stream);

/* ---- batch preparation: RemoveIsolatedNodes (train_hiv.py:171-173 `data = transform(data)`, :457; benchmarks/utils.py:39-49;
 * arithmetic in torch_geometric 1.6.1 utils/isolated.py remove_isolated_nodes, restated in oracle/phc_oracle.py) ------------
 * keep_mask[N] (bytes): node is an endpoint of a non-self-loop edge; assoc[N]: new id or -1; new_edge_index: int64 [2,E]
 * buffer (row stride E) holding the relabelled non-loop edges in original order followed by one self loop per kept node
 * that has any (ascending node id, last such edge wins); edge_order[E]: source edge id of every output edge (to gather
 * edge_attr); counts: DEVICE int[4] = kept nodes, kept non-loop edges, kept self loops, status (bit0 index out of range).
 * Bit-exact with the CPU restatement. */
size_t phc_isolated_workspace_bytes(int num_nodes, int num_edges);
int phc_remove_isolated_nodes(const long long* edge_index, int num_edges, int num_nodes, unsigned char* keep_mask, long long* assoc,
                              long long* new_edge_index, long long* edge_order, int* counts, void* workspace, size_t workspace_bytes,
                              phc_stream_t stream);

/* ---- batch preparation: Batch collate from a dataset resident in HBM (replaces the CPU DataLoader collate,
 * torch_geometric.data.DataLoader -> Batch.from_data_list, reference benchmarks/train_hiv.py:481-493 and the other train_*.py;
 * restated in oracle/phc_oracle.py::collate).  Store: graphs packed back to back — node_ptr / edge_ptr int64 [store_graphs+1],
 * edge_index int64 [2, store_edges] with node ids LOCAL to each graph, x rows per node, edge_attr rows per edge, y rows per graph
 * (row sizes in bytes, multiples of 4; pass 0 / NULL to skip a tensor).  Batch: graph_ids int64 [num_graphs] (any order,
 * repeats allowed), out_node_ptr / out_edge_ptr int64 [num_graphs+1] = exclusive prefix sums of the selected graphs' sizes
 * (checked on the device).  Writes out_edge_index [2, out_edges] (shifted by the batch node offsets), out_batch [out_nodes],
 * and the gathered rows.  status: DEVICE int, bit0 graph id out of range, bit1 prefix sums do not match.  Bit-exact. */
int phc_collate_batch(const long long* graph_ids, int num_graphs, int store_graphs, const long long* node_ptr, const long long* edge_ptr,
                      const long long* out_node_ptr, const long long* out_edge_ptr, const long long* edge_index, long long store_edges,
                      long long* out_edge_index, long long out_edges, long long out_nodes, long long* out_batch, const void* x, void* out_x,
                      int x_row_bytes, const void* edge_attr, void* out_edge_attr, int edge_attr_row_bytes, const void* y, void* out_y,
                      int y_row_bytes, int* status, phc_stream_t stream);

/* ---- fused neighbour aggregation (messagepassing.py:72-74,136-138,297-300) -------------------
 * out[i] = (self_loop ? x[i] : 0) + AGG_{e: dst(e)=i} act(x[src(e)] + ea[e])
 * aux_f: [2,N,F] (softmax only: log-sum-exp and aggregate), aux_i: [N,F] (max/min only: winning edge id).
 * beta: device pointer to the softmax inverse temperature (ignored otherwise). */
int phc_aggregate_fwd(const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int num_nodes, int width,
                      int reduce, int msg_act, const float* beta, int self_loop, float* out, float* aux_f, int* aux_i,
                      phc_stream_t stream);
size_t phc_aggregate_bwd_workspace_bytes(int num_nodes, int width);
int phc_aggregate_bwd(const float* gout, const float* x, const float* ea, const float* aux_f, const int* aux_i, const int* rowptr,
                      const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                      int width, int reduce, int msg_act, const float* beta, int self_loop, float* dx, float* dea, float* dbeta,
                      void* workspace, size_t workspace_bytes, phc_stream_t stream);

/* ---- graph pooling (pooling.py:10-25 global_add_pool; :57-66 soft attention) ------------------
 * gate_logits == NULL: plain sum.  Otherwise out[b,c,f] = sum_i sigmoid(gate_logits[i,f]) x[i,c,f],
 * gate_logits [N, width/phm_dim]. */
int phc_segment_pool_fwd(const float* x, const float* gate_logits, const int* graph_ptr, int num_graphs, int width, int phm_dim,
                         float* out, phc_stream_t stream);
int phc_segment_pool_bwd(const float* gout, const float* x, const float* gate_logits, const long long* batch, int num_nodes, int width,
                         int phm_dim, float* dx, float* dgate_logits, phc_stream_t stream);

/* ---- batch-norm + activation + dropout + skip add (norm.py:30-35, layers.py:31-55, models.py:206-215)
 * y = skip + dropout(act(gamma*(h-mean)*rstd+beta)); gamma/beta/running_* are flat [width] vectors (the
 * n per-component BatchNorm1d parameter blocks laid out back to back).  use_bn=0 skips the normalisation.
 * Dropout masks are a pure function of (seed, element index); drop_same=1 shares one mask between the
 * n components (layers.py:44-52). */
size_t phc_bn_workspace_bytes(int rows, int width);
int phc_bn_act_drop_skip_fwd(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                             int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                             unsigned long long seed, float* y, float* save_mean, float* save_rstd, void* workspace,
                             size_t workspace_bytes, phc_stream_t stream);
/* Training-mode forward with the chunk moments already produced by the kernel that wrote h (see phc_phm_linear_fwd_bnstats):
 * partials[ceil(rows/chunk_rows)][2][width] = (chunk mean, chunk M2). */
int phc_bn_act_drop_skip_fwd_partials(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                      long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                                      int training, float momentum, float eps, int act, float drop_p, int drop_same,
                                      unsigned long long seed, float* y, float* save_mean, float* save_rstd, const float* partials,
                                      int chunk_rows, phc_stream_t stream);
int phc_bn_act_drop_skip_bwd(const float* dy, const float* h, const float* gamma, const float* beta, const float* save_mean,
                             const float* save_rstd, int rows, int width, int phm_dim, int use_bn, int training, int act, float drop_p,
                             int drop_same, unsigned long long seed, float* dh, float* dgamma, float* dbeta, void* workspace,
                             size_t workspace_bytes, phc_stream_t stream);
/* The same pair with a ROW STRIDE on the forward output / the incoming gradient: they are the left column block of a wider
 * [rows, y_row_stride] buffer whose right block holds the skip features — PHMSkipConnectConcat's `torch.cat([h, x0], -1)`
 * (undirectional/models.py:467) without the concatenation pass: the norm writes straight into the layer's input buffer. */
int phc_bn_act_drop_skip_fwd_strided(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                                     int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                                     unsigned long long seed, float* y, int y_row_stride, float* save_mean, float* save_rstd,
                                     void* workspace, size_t workspace_bytes, phc_stream_t stream);
int phc_bn_act_drop_skip_bwd_strided(const float* dy, int dy_row_stride, const float* h, const float* gamma, const float* beta,
                                     const float* save_mean, const float* save_rstd, int rows, int width, int phm_dim, int use_bn,
                                     int training, int act, float drop_p, int drop_same, unsigned long long seed, float* dh,
                                     float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, phc_stream_t stream);


/* out = srcs[0] + srcs[1] + ... (list order) over `numel` floats; srcs is a HOST array of up to 16 device pointers.  Used for the
 * gradient of a skip connection that fans out to every layer (models.py:227-236, sc_type="first"). */
int phc_sum_tensors(const float* const* srcs, int count, long long numel, float* out, phc_stream_t stream);

/* ---- index validation: the embedding kernels clamp an index outside its table (nn.Embedding raises); this ORs bit 0 into the device
 * word *status for any idx[r, c] outside [0, vocab[c]) — read it back when you want to know (synchronising, debug / tests). ---------- */
int phc_index_check(const long long* idx, const int* vocab, int rows, int cols, int* status, phc_stream_t stream);

/* ---- encoders (encoder.py:31-34; quaternion/encoder.py:44-56) ---------------------------------
 * tables / weights / biases are HOST arrays of device pointers, ordered [component][column]. */
int phc_embed_sum_fwd(const long long* idx, const float* const* tables, const int* vocab, int rows, int cols, int phm_dim,
                      int width_per_component, float* out, phc_stream_t stream);
size_t phc_embed_bwd_workspace_bytes(int rows, int total_vocab, int width);
int phc_embed_sum_bwd(const float* gout, const long long* idx, float* const* dtables, const int* vocab, int rows, int cols, int phm_dim,
                      int width_per_component, void* workspace, size_t workspace_bytes, phc_stream_t stream);
int phc_linear_encoder_fwd(const float* feat, const float* const* weights, const float* const* biases, int rows, int in_dim, int phm_dim,
                           int width_per_component, float* out, phc_stream_t stream);
size_t phc_linear_encoder_bwd_workspace_bytes(int rows, int in_dim, int width);
int phc_linear_encoder_bwd(const float* gout, const float* feat, float* const* dweights, float* const* dbiases, int rows, int in_dim,
                           int phm_dim, int width_per_component, void* workspace, size_t workspace_bytes, phc_stream_t stream);

/* ---- PHMLinear (layers.py:198-219 matvec_product_new; kronecker.py:35-48) ---------------------
 * y = act(x * (sum_b A_b (x) W_b) + bias) + residual, without materialising the Kronecker weight.
 * phm_rule [n,n,n], W [n, in/n, out/n], bias [out] or NULL, residual [rows,out] or NULL.
 * bwd: dx [rows,in] (NULL to skip), d_rule [n,n,n] (NULL to skip), dW [n,in/n,out/n], dbias (NULL to skip). */
size_t phc_phm_linear_fwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision);
size_t phc_phm_linear_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision);
int phc_phm_linear_fwd(const float* x, const float* phm_rule, const float* W, const float* bias, const float* residual, float* y, int rows,
                       int in_features, int out_features, int phm_dim, int act, int precision, void* workspace, size_t workspace_bytes,
                       phc_stream_t stream);
/* As phc_phm_linear_fwd; additionally the epilogue of the tensor-core n = 4 kernel writes, for a batch-norm that follows
 * (norm.py:30-35 after layers.py:349-355), the per-32-row-chunk column moments of y (chunk mean, chunk M2) to
 * bn_partials[ceil(rows/32)][2][out_features]; *bn_produced = 1 when they were written (other paths: 0, use the plain
 * phc_bn_act_drop_skip_fwd).  Consume them with phc_bn_act_drop_skip_fwd_partials(..., bn_partials, 32, ...). */
int phc_phm_linear_fwd_bnstats(const float* x, const float* phm_rule, const float* W, const float* bias, const float* residual, float* y,
                               int rows, int in_features, int out_features, int phm_dim, int act, int precision, void* workspace,
                               size_t workspace_bytes, float* bn_partials, int* bn_produced, phc_stream_t stream);
int phc_phm_linear_bwd(const float* gy, const float* x, const float* phm_rule, const float* W, float* dx, float* d_rule, float* dW,
                       float* dbias, int rows, int in_features, int out_features, int phm_dim, int precision, void* workspace,
                       size_t workspace_bytes, const void* fwd_workspace, phc_stream_t stream);
/* fwd_workspace: optional — the workspace of the matching phc_phm_linear_fwd call, untouched since; lets the
 * tensor-core path reuse the operand packs written there instead of re-packing W (its scratch part is reused too). */

/* ---- aggregation with the edge encoder fused in (models.py:238-243 + messagepassing.py:72-74,136-138,297-300)
 * out[i] = (self_loop ? x[i] : 0) + AGG_e act(x[src(e)] + enc(edge_attr[e])); the [E,F] edge embedding is never
 * materialised.  enc_kind 0: n x Linear(enc_dim -> F/n), edge_attr float [E,enc_dim], params = HOST array of the n
 * weight pointers ([F/n, enc_dim]) followed by the n bias pointers.  enc_kind 1: integer features int64
 * [E,enc_dim], vocab[enc_dim], params = n*enc_dim embedding tables ordered [component][column].
 * table_rows = enc_dim+1 (Linear) or sum(vocab) (embeddings).  bwd writes the parameter gradients through
 * dparams (same order), dx and dbeta. */
int phc_conv_fused_supported(int width, int phm_dim, int enc_kind, int enc_dim, int table_rows);
int phc_conv_fused_fwd(const float* x, const void* edge_attr, int enc_kind, int enc_dim, const int* vocab, const float* const* params,
                       const int* rowptr, const int* col, const int* perm, int num_nodes, int width, int phm_dim, int reduce,
                       int msg_act, const float* beta, int self_loop, float* out, float* aux_f, int* aux_i, phc_stream_t stream);
/* sum / mean aggregation with the identity message: the (linear) encoder term leaves the edge loop,
 * out[i] = self*x[i] + scale_i * sum_e x[src(e)] + sum_r node_sums[i,r] * table[r]  (node_sums from phc_edge_feature_sums with
 * the same `mean` flag) — a pure row gather, ~3x faster than the per-edge kernel; same result up to fp32 summation order. */
size_t phc_conv_fused_fwd_sums_workspace_bytes(int width, int table_rows);
int phc_conv_fused_fwd_sums(const float* x, const float* node_sums, int enc_kind, int enc_dim, const int* vocab, const float* const* params,
                            const int* rowptr, const int* col, int num_nodes, int width, int phm_dim, int reduce, int self_loop,
                            float* out, void* workspace, size_t workspace_bytes, phc_stream_t stream);
size_t phc_conv_fused_bwd_workspace_bytes(int num_nodes, int width, int table_rows);
int phc_conv_fused_bwd(const float* gout, const float* x, const void* edge_attr, int enc_kind, int enc_dim, const int* vocab,
                       const float* const* params, float* const* dparams, const float* aux_f, const int* aux_i, const int* rowptr,
                       const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                       int width, int phm_dim, int reduce, int msg_act, const float* beta, int self_loop, const float* node_sums,
                       float* dx, float* dbeta, void* workspace, size_t workspace_bytes, phc_stream_t stream);
/* node_sums [N, table_rows] (optional, sum/mean + identity message only): per-node sums of the edge features
 * (Linear: raw features and in-degree; embeddings: value histograms), scaled by 1/deg for mean.  Depends only on
 * the batch, so it is computed once and shared by all layers; with it the encoder gradients need no edge loop. */
int phc_edge_feature_sums(const void* edge_attr, int enc_kind, int enc_dim, const int* vocab, const int* rowptr, const int* perm,
                          int num_nodes, int mean, float* node_sums, phc_stream_t stream);

/* ---- one whole message-passing layer per call (models.py:200-217; descriptor in phc_b200_layer.h) -----------
 * fwd: conv (aggregation + fused edge encoder) -> PHMLinear [-> BN -> act -> PHMLinear] -> BN -> act -> dropout -> + skip.
 * bwd: the gradients of all of it (dx, encoder / rule / W / bias / BN-affine gradients, d softmax beta).  Same kernels
 * and arithmetic as the per-operator entry points above, sequenced on `stream` by the library. */
size_t phc_conv_layer_desc_bytes(void);
size_t phc_conv_layer_workspace_bytes(int num_nodes, int width, int phm_dim, int table_rows, int precision);
int phc_conv_layer_fwd(const phc_conv_layer* layer, phc_stream_t stream);
int phc_conv_layer_bwd(const phc_conv_layer* layer, phc_stream_t stream);

/* ---- principal-neighbourhood aggregation (messagepassing.py:421-438 PHMPNAConvSimple.message/aggregate;
 * aggregator.py:70-93 aggregators, :112-135 scalers; utils.py:122-135 phm_cat) -------------------------------
 * out[i, c, s, t, j] = scale_s(deg_i) * AGG_t{act(x[src(e)] + ea[e]) : dst(e) = i}[c*F/n + j], out is [N, S*T*F].
 * aggr_code / scaler_code: the lists packed 4 bits per entry, first entry in the lowest nibble, 0 terminates.
 * aggregators: 1 sum, 2 mean, 3 min, 4 max, 5 var, 6 std.  scalers: 1 identity, 2 amplification, 3 attenuation,
 * 4 linear, 5 inverse_linear.  aux_f [2,N,F] (mean, var; needed by var/std), aux_i [2,N,F] (argmin, argmax edge
 * ids; needed by min/max) are written by fwd for bwd.  bwd writes dx [N,F] and dea [E,F]. */
int phc_pna_aggregate_fwd(const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int num_nodes, int width,
                          int phm_dim, int msg_act, unsigned long long aggr_code, unsigned long long scaler_code, float avg_deg_log,
                          float avg_deg_lin, float* out, float* aux_f, int* aux_i, phc_stream_t stream);
int phc_pna_aggregate_bwd(const float* gout, const float* x, const float* ea, const float* aux_f, const int* aux_i, const int* rowptr,
                          const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                          int width, int phm_dim, int msg_act, unsigned long long aggr_code, unsigned long long scaler_code,
                          float avg_deg_log, float avg_deg_lin, float* dx, float* dea, phc_stream_t stream);

/* ---- weight regulariser (regularization.py:15-23): out = sum_l mean_{k,p} ||W_l[:,k,p]||_2 ------------
 * weights / dweights: HOST arrays of device pointers to the [n_l, K_l, P_l] weight tensors; kp[l] = K_l*P_l. */
size_t phc_weight_reg_workspace_bytes(int count);
int phc_weight_reg_fwd(const float* const* weights, const int* phm_dims, const int* kp, int count, float* out, void* workspace,
                       size_t workspace_bytes, phc_stream_t stream);
int phc_weight_reg_bwd(const float* gout, const float* const* weights, float* const* dweights, const int* phm_dims, const int* kp,
                       int count, phc_stream_t stream);
/* dweights[l] += ... : the regulariser's gradient added onto gradients that are already in place (the flat gradient buffer), one
 * launch for all weight tensors instead of one autograd accumulation kernel per tensor. */
int phc_weight_reg_bwd_accumulate(const float* gout, const float* const* weights, float* const* dweights, const int* phm_dims,
                                  const int* kp, int count, phc_stream_t stream);

/* ---- task loss of the training step and its gradient in one launch (train_ppa.py / train_mnist.py:200 cross entropy;
 * train_hiv.py:174-178 / train_pcba.py BCE-with-logits over the labelled entries; train_zinc.py:192 mean absolute error) -------
 * kind 0: logits [rows, cols], targets int64 [rows];  kind 1: targets float [rows, cols], NaN = unlabelled;  kind 2: cols = 1,
 * targets float [rows].  loss: device scalar (mean);  dlogits [rows, cols]: d(loss)/d(logits). */
int phc_task_loss(int kind, const float* logits, const void* targets, int rows, int cols, float* loss, float* dlogits,
                  phc_stream_t stream);

/* ---- clip_grad_norm_(max_norm) + Adam.step() on one flat buffer (train_hiv.py:199-201) --------------------
 * coef = min(1, max_norm/(||grads||_2 + 1e-6)) (max_norm <= 0: no clipping); Adam with weight_decay 0;
 * bias_correction{1,2} = 1 - beta^t computed by the caller.  grad_norm_out (optional) receives ||grads||_2. */
size_t phc_adam_workspace_bytes(void);
int phc_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long numel, float lr, float beta1,
                       float beta2, float eps, float bias_correction1, float bias_correction2, float max_norm, float* grad_norm_out,
                       void* workspace, size_t workspace_bytes, phc_stream_t stream);
/* The same step with the learning rate and the step counter t in DEVICE memory (lr_dev[0], step_dev[0]; the call advances t and
 * derives the bias corrections from it): nothing the host passes by value changes from step to step, so the call can be recorded
 * into a CUDA graph and replayed (the reference's train() body as one graph: phc_gnn_b200/graphed.py). */
int phc_adam_clip_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long numel, const float* lr_dev,
                           float beta1, float beta2, float eps, int* step_dev, float max_norm, float* grad_norm_out, void* workspace,
                           size_t workspace_bytes, phc_stream_t stream);

/* ---- dropout epoch (graph replay) --------------------------------------------------------------------------
 * Dropout masks are a pure function of (seed, element index) with the seed passed by value — a recorded CUDA graph would replay
 * the same masks for ever.  With an epoch word registered (device memory, unsigned long long[1]; NULL unregisters) every dropout
 * kernel of this library keys its Philox stream with seed + epoch[0] * 0x9E3779B97F4A7C15, read on the device at run time;
 * phc_dropout_epoch_advance adds 1 to it on `stream` (recorded at the top of a captured step, so forward and backward of one
 * replay agree and consecutive replays differ).  Process-global, like torch's default generator. */
int phc_dropout_epoch_register(unsigned long long* epoch_dev);
int phc_dropout_epoch_advance(unsigned long long* epoch_dev, phc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PHC_B200_H */
