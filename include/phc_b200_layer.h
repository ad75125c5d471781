/* phc_b200_layer.h — descriptor of one whole message-passing layer for phc_conv_layer_fwd / phc_conv_layer_bwd
 * (declared in phc_b200.h).
 *
 * Replaces, per layer, the reference chain  conv.propagate (messagepassing.py:55-70,132-142) -> PHMLinear or
 * PHMMLP (layers.py:284-299,349-355) -> PHMNorm -> activation -> phm_dropout -> skip add (models.py:200-217).
 * Issuing these as separate calls costs more host time than the kernels take on the GPU, so the whole chain is
 * launched by ONE call (forward) and its gradients by ONE call (backward); the arithmetic is exactly that of the
 * per-operator entry points.  All pointers are DEVICE pointers unless stated otherwise; every buffer is owned by the
 * caller.  The same descriptor is passed to forward and then (with the "backward" fields filled) to backward.
 */
#ifndef PHC_B200_LAYER_H
#define PHC_B200_LAYER_H

#include <stddef.h>

typedef struct phc_conv_layer {
  /* shapes and switches */
  int num_nodes, width, phm_dim;
  int enc_kind, enc_dim, table_rows;       /* edge encoder: see phc_conv_fused_fwd */
  int reduce, msg_act, self_loops, mlp;    /* mlp=1: two PHMLinear with BN+act between (GINE); 0: one PHMLinear (+x if self_loops) */
  int act1, act2, use_bn1, use_bn2, training, drop_same, precision;
  int n_tracked1, n_tracked2;
  float drop_p, momentum1, eps1, momentum2, eps2;
  unsigned long long seed;
  /* graph structure (phc_csr_build) */
  const int *rowptr, *col, *perm, *rowptr_t, *col_t, *perm_t;
  /* inputs */
  const float* x;                          /* [N, width] */
  const float* skip;                       /* [N, width] or NULL */
  const void* edge_attr;                   /* raw edge features */
  const int* vocab;                        /* HOST array [enc_dim] (embeddings) or NULL */
  const float* const* enc_params;          /* HOST array of device pointers */
  const float* softmax_beta;               /* softmax inverse temperature or NULL */
  /* parameters */
  const float *rule1, *W1, *b1, *rule2, *W2, *b2;
  const float *gamma1, *beta1;
  float *running_mean1, *running_var1;
  long long* tracked1;
  const float *gamma2, *beta2;
  float *running_mean2, *running_var2;
  long long* tracked2;
  /* activations written by forward and read by backward, [N, width] each (y1, a1 unused when mlp=0) */
  float *agg, *y1, *a1, *z, *out;
  float *stats1, *stats2;                  /* [2, width] each: batch mean, rstd */
  float* aux_f;                            /* softmax: [2,N,width] */
  int* aux_i;                              /* max/min: [N,width] */
  void *ws_lin1, *ws_lin2;                 /* PHMLinear forward workspaces; kept untouched until backward (operand packs) */
  size_t ws_lin1_bytes, ws_lin2_bytes;
  void* ws;                                /* scratch, phc_conv_layer_workspace_bytes() */
  size_t ws_bytes;
  const float* node_sums;                  /* optional (sum / mean + identity message): phc_edge_feature_sums output, used by
                                              forward (phc_conv_fused_fwd_sums) and backward (phc_conv_fused_bwd) */
  /* backward only */
  const float* gout;                       /* [N, width] gradient of `out` */
  float *tmp_a, *tmp_b;                    /* two [N, width] scratch matrices; with mlp=0 tmp_a holds d(z) on return */
  float* dx;                               /* [N, width] */
  float* d_softmax_beta;                   /* scalar (zero-initialised by the caller) or NULL */
  float* const* d_enc_params;              /* HOST array of device pointers */
  float *d_rule1, *d_W1, *d_b1, *d_rule2, *d_W2, *d_b2;   /* d_rule*, d_b* may be NULL */
  float *d_gb1, *d_gb2;                    /* [2, width] each: dgamma then dbeta */
} phc_conv_layer;

#endif /* PHC_B200_LAYER_H */
