// One C-ABI call per message-passing layer (forward) and one for its gradients (backward).
//
// Replaces, per layer, the reference chain  conv.propagate (phc/hypercomplex/undirectional/messagepassing.py:55-70,
// 132-142) -> PHMLinear / PHMMLP (layers.py:284-299, 349-355) -> PHMNorm -> activation -> phm_dropout -> skip add
// (undirectional/models.py:200-217).  The kernels are those of the per-operator entry points (conv_fused.cu,
// phm_linear_*.cu, norm.cu); this file only sequences them on one stream, so that the host spends one foreign call per
// layer instead of six (forward) / seven (backward) — at ppa shape the Python-side cost of those calls had caught up
// with the GPU time of the kernels.
#include <cuda_runtime.h>
#include <stdlib.h>
#include "../../include/phc_b200.h"

void phc_set_error(const char* fmt, ...);

#define LAYER_REQUIRE(cond, ...)  \
  do {                            \
    if (!(cond)) {                \
      phc_set_error(__VA_ARGS__); \
      return 1;                   \
    }                             \
  } while (0)

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

extern "C" {

size_t phc_conv_layer_desc_bytes(void) { return sizeof(phc_conv_layer); }

size_t phc_conv_layer_workspace_bytes(int num_nodes, int width, int phm_dim, int table_rows, int precision) {
  size_t b = phc_bn_workspace_bytes(num_nodes, width);
  b = max_sz(b, phc_phm_linear_bwd_workspace_bytes(num_nodes, width, width, phm_dim, precision));
  b = max_sz(b, phc_conv_fused_bwd_workspace_bytes(num_nodes, width, table_rows));
  b = max_sz(b, phc_conv_fused_fwd_sums_workspace_bytes(width, table_rows));
  return b + 1024;
}

static int check_desc(const phc_conv_layer* L, const char* who) {
  LAYER_REQUIRE(L != nullptr, "%s: null descriptor", who);
  LAYER_REQUIRE(L->width > 0 && L->phm_dim > 0 && L->width % L->phm_dim == 0, "%s: width %d not divisible by phm_dim %d", who, L->width,
                L->phm_dim);
  LAYER_REQUIRE(L->x && L->agg && L->z && L->out && L->ws, "%s: x / agg / z / out / ws are required", who);
  LAYER_REQUIRE(!L->mlp || (L->y1 && L->a1 && L->rule2 && L->W2), "%s: mlp=1 needs y1, a1, rule2, W2", who);
  LAYER_REQUIRE(L->ws_bytes >= phc_conv_layer_workspace_bytes(L->num_nodes, L->width, L->phm_dim, L->table_rows, L->precision),
                "%s: workspace too small", who);
  return 0;
}

int phc_conv_layer_fwd(const phc_conv_layer* L, phc_stream_t stream) {
  int rc = check_desc(L, "phc_conv_layer_fwd");
  if (rc) return rc;
  const int N = L->num_nodes, F = L->width, n = L->phm_dim;
  if (N == 0) return 0;
  // 1. aggregation with the edge encoder fused in
  if (L->node_sums != nullptr && (L->reduce == PHC_RED_SUM || L->reduce == PHC_RED_MEAN) && L->msg_act == PHC_ACT_IDENTITY)
    rc = phc_conv_fused_fwd_sums(L->x, L->node_sums, L->enc_kind, L->enc_dim, L->vocab, L->enc_params, L->rowptr, L->col, N, F, n, L->reduce,
                                 L->self_loops && L->mlp, L->agg, L->ws, L->ws_bytes, stream);   // encoder term from the per-node sums
  else
    rc = phc_conv_fused_fwd(L->x, L->edge_attr, L->enc_kind, L->enc_dim, L->vocab, L->enc_params, L->rowptr, L->col, L->perm, N, F, n, L->reduce,
                            L->msg_act, L->softmax_beta, L->self_loops && L->mlp, L->agg, L->aux_f, L->aux_i, stream);
  if (rc) return rc;
  // 2.-4. PHM transform.  When a training-mode batch-norm follows a PHMLinear, the linear kernel's epilogue also emits the
  // chunk moments of its output (into the scratch workspace) and the norm only merges them.
  float* part = reinterpret_cast<float*>(L->ws);
  static const bool fuse_stats = getenv("PHC_NO_FUSED_BN_STATS") == nullptr;      // A/B switch
  const bool part_fits = fuse_stats && L->ws_bytes >= sizeof(float) * 2 * ((size_t)(N + 31) / 32) * F;
  int got1 = 0, got2 = 0;
  if (L->mlp) {
    const bool want1 = L->use_bn1 && L->training && part_fits;
    rc = phc_phm_linear_fwd_bnstats(L->agg, L->rule1, L->W1, L->b1, nullptr, L->y1, N, F, F, n, PHC_ACT_IDENTITY, L->precision, L->ws_lin1,
                                    L->ws_lin1_bytes, want1 ? part : nullptr, &got1, stream);
    if (rc) return rc;
    if (got1)
      rc = phc_bn_act_drop_skip_fwd_partials(L->y1, L->gamma1, L->beta1, L->running_mean1, L->running_var1, L->tracked1, L->n_tracked1, nullptr,
                                             N, F, n, 1, L->momentum1, L->eps1, L->act1, 0.f, 0, 0ull, L->a1, L->stats1, L->stats1 + F, part,
                                             32, stream);
    else
      rc = phc_bn_act_drop_skip_fwd(L->y1, L->gamma1, L->beta1, L->use_bn1 ? L->running_mean1 : nullptr,
                                    L->use_bn1 ? L->running_var1 : nullptr, (L->use_bn1 && L->training) ? L->tracked1 : nullptr,
                                    L->n_tracked1, nullptr, N, F, n, L->use_bn1, L->training, L->momentum1, L->eps1, L->act1, 0.f, 0, 0ull,
                                    L->a1, L->use_bn1 ? L->stats1 : nullptr, L->use_bn1 ? L->stats1 + F : nullptr, L->ws, L->ws_bytes, stream);
    if (rc) return rc;
  }
  const bool want2 = L->use_bn2 && L->training && part_fits;
  if (L->mlp)
    rc = phc_phm_linear_fwd_bnstats(L->a1, L->rule2, L->W2, L->b2, nullptr, L->z, N, F, F, n, PHC_ACT_IDENTITY, L->precision, L->ws_lin2,
                                    L->ws_lin2_bytes, want2 ? part : nullptr, &got2, stream);
  else
    rc = phc_phm_linear_fwd_bnstats(L->agg, L->rule1, L->W1, L->b1, L->self_loops ? L->x : nullptr, L->z, N, F, F, n, PHC_ACT_IDENTITY,
                                    L->precision, L->ws_lin1, L->ws_lin1_bytes, want2 ? part : nullptr, &got2, stream);
  if (rc) return rc;
  // 5. norm -> act -> dropout -> + skip
  if (got2)
    return phc_bn_act_drop_skip_fwd_partials(L->z, L->gamma2, L->beta2, L->running_mean2, L->running_var2, L->tracked2, L->n_tracked2, L->skip, N,
                                             F, n, 1, L->momentum2, L->eps2, L->act2, L->drop_p, L->drop_same, L->seed, L->out, L->stats2,
                                             L->stats2 + F, part, 32, stream);
  return phc_bn_act_drop_skip_fwd(L->z, L->gamma2, L->beta2, L->use_bn2 ? L->running_mean2 : nullptr, L->use_bn2 ? L->running_var2 : nullptr,
                                  (L->use_bn2 && L->training) ? L->tracked2 : nullptr, L->n_tracked2, L->skip, N, F, n, L->use_bn2, L->training,
                                  L->momentum2, L->eps2, L->act2, L->drop_p, L->drop_same, L->seed, L->out, L->use_bn2 ? L->stats2 : nullptr,
                                  L->use_bn2 ? L->stats2 + F : nullptr, L->ws, L->ws_bytes, stream);
}

int phc_conv_layer_bwd(const phc_conv_layer* L, phc_stream_t stream) {
  int rc = check_desc(L, "phc_conv_layer_bwd");
  if (rc) return rc;
  LAYER_REQUIRE(L->gout && L->tmp_a && L->tmp_b && L->dx && L->d_W1, "phc_conv_layer_bwd: gout / tmp_a / tmp_b / dx / d_W1 are required");
  LAYER_REQUIRE(!L->mlp || L->d_W2, "phc_conv_layer_bwd: mlp=1 needs d_W2");
  const int N = L->num_nodes, F = L->width, n = L->phm_dim;
  if (N == 0) return 0;
  float *dz = L->tmp_a, *dagg = L->tmp_b;
  // 5'
  rc = phc_bn_act_drop_skip_bwd(L->gout, L->z, L->gamma2, L->beta2, L->use_bn2 ? L->stats2 : nullptr, L->use_bn2 ? L->stats2 + F : nullptr, N, F,
                                n, L->use_bn2, L->training, L->act2, L->drop_p, L->drop_same, L->seed, dz, L->use_bn2 ? L->d_gb2 : nullptr,
                                L->use_bn2 ? L->d_gb2 + F : nullptr, L->ws, L->ws_bytes, stream);
  if (rc) return rc;
  // 4'-2'
  const void* pack1 = L->ws_lin1_bytes > 64 ? L->ws_lin1 : nullptr;     // tensor-core path: operand packs written by forward
  if (L->mlp) {
    const void* pack2 = L->ws_lin2_bytes > 64 ? L->ws_lin2 : nullptr;
    float *da1 = L->tmp_b, *dy1 = L->tmp_a;
    rc = phc_phm_linear_bwd(dz, L->a1, L->rule2, L->W2, da1, L->d_rule2, L->d_W2, L->d_b2, N, F, F, n, L->precision, L->ws, L->ws_bytes, pack2,
                            stream);
    if (rc) return rc;
    rc = phc_bn_act_drop_skip_bwd(da1, L->y1, L->gamma1, L->beta1, L->use_bn1 ? L->stats1 : nullptr, L->use_bn1 ? L->stats1 + F : nullptr, N, F,
                                  n, L->use_bn1, L->training, L->act1, 0.f, 0, 0ull, dy1, L->use_bn1 ? L->d_gb1 : nullptr,
                                  L->use_bn1 ? L->d_gb1 + F : nullptr, L->ws, L->ws_bytes, stream);
    if (rc) return rc;
    rc = phc_phm_linear_bwd(dy1, L->agg, L->rule1, L->W1, dagg, L->d_rule1, L->d_W1, L->d_b1, N, F, F, n, L->precision, L->ws, L->ws_bytes,
                            pack1, stream);
  } else {
    rc = phc_phm_linear_bwd(dz, L->agg, L->rule1, L->W1, dagg, L->d_rule1, L->d_W1, L->d_b1, N, F, F, n, L->precision, L->ws, L->ws_bytes, pack1,
                            stream);
  }
  if (rc) return rc;
  // 1'
  return phc_conv_fused_bwd(dagg, L->x, L->edge_attr, L->enc_kind, L->enc_dim, L->vocab, L->enc_params, L->d_enc_params, L->aux_f, L->aux_i,
                            L->rowptr, L->col, L->perm, L->rowptr_t, L->col_t, L->perm_t, N, F, n, L->reduce, L->msg_act, L->softmax_beta,
                            L->self_loops && L->mlp, L->node_sums, L->dx, L->d_softmax_beta, L->ws, L->ws_bytes, stream);
}

}  // extern "C"
