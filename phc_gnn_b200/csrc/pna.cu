// Principal-neighbourhood aggregation (PNA) for the hypercomplex convolution: every aggregator and every degree
// scaler in ONE segmented gather-reduce over the CSR-sorted in-edges, written straight into the component-wise
// concatenated layout the following PHMLinear(S*T*F -> F) expects.
//
//   m_e      = phi(x[src(e),:] + ea[e,:])
//   agg_t[i] = AGG_t {m_e : dst(e) = i},   t in {sum, mean, min, max, var, std}
//   out[i, c, s, t, j] = scale_s(deg_i) * agg_t[i, c*Fc + j]          (c: hypercomplex component, Fc = F/n)
//
// Replaces reference PHMPNAConvSimple.message/aggregate (phc/hypercomplex/undirectional/messagepassing.py:421-438):
// index_select + add + relu, then per aggregator a torch_scatter launch (aggregator.py:70-93; std = sqrt(relu(
// E[m^2]-E[m]^2)+1e-5)), phm_cat (utils.py:122-135), degree(), per scaler a multiply (aggregator.py:112-135) and a
// second phm_cat: ~25 launches, seven [E,F] / [N,4F] / [N,12F] temporaries and float atomics.  Here each edge row
// is read once, rows are reduced in CSR order by one thread per 4 features (deterministic), and the [N, S*T*F]
// result is written once.  Empty rows: every aggregator gives 0 (std: sqrt(1e-5)); attenuation / inverse_linear
// scale = 1 at degree 0 (aggregator.py:121-123,132-134).
// Roofline: HBM.  Algorithmic bytes fwd = 4F(N + E) + 4*S*T*F*N + 8E + 4(N+1).
#include "common.cuh"
#include <float.h>

int phc_aggregate_bwd_node_from_edges(const float* dea, const int* rowptr_t, const int* col_t, const int* perm_t, int N, int F, float* dx,
                                      cudaStream_t stream);   // aggregate.cu

namespace {

enum { PNA_SUM = 1, PNA_MEAN = 2, PNA_MIN = 3, PNA_MAX = 4, PNA_VAR = 5, PNA_STD = 6 };
enum { PNA_IDENTITY = 1, PNA_AMPLIFICATION = 2, PNA_ATTENUATION = 3, PNA_LINEAR = 4, PNA_INVERSE_LINEAR = 5 };
constexpr int PNA_MAX_LIST = 8;

struct PnaCfg {
  int T, S;                       // number of aggregators / scalers
  int aggr[PNA_MAX_LIST];
  int scaler[PNA_MAX_LIST];
  float avg_log, avg_lin;         // avg_deg['log'], avg_deg['lin'] (messagepassing.py:376-381)
};

__device__ __forceinline__ float pna_scale(int kind, int deg, const PnaCfg& c) {
  const float d = (float)deg;
  switch (kind) {
    case PNA_AMPLIFICATION: return logf(d + 1.f) / c.avg_log;
    case PNA_ATTENUATION: return deg == 0 ? 1.f : c.avg_log / logf(d + 1.f);
    case PNA_LINEAR: return d / c.avg_lin;
    case PNA_INVERSE_LINEAR: return deg == 0 ? 1.f : c.avg_lin / d;
    default: return 1.f;
  }
}

// column of (component-local feature j of component c, scaler s, aggregator t) in the [N, S*T*F] output
__device__ __forceinline__ size_t pna_col(int c, int j, int s, int t, int Fc, int S, int T) {
  return (size_t)c * S * T * Fc + (size_t)(s * T + t) * Fc + j;
}

template <int VEC>
__global__ void __launch_bounds__(256) pna_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ea,
                                                      const int* __restrict__ rowptr, const int* __restrict__ col,
                                                      const int* __restrict__ perm, int N, int F, int Fc, int act, PnaCfg cfg,
                                                      float* __restrict__ out, float* __restrict__ aux_f, int* __restrict__ aux_i) {
  pdl_begin();
  const int fv = F / VEC;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int i = (int)(t / fv);
  const int f = (int)(t % fv) * VEC;
  const int beg = rowptr[i], end = rowptr[i + 1];
  float s1[VEC], s2[VEC], mn[VEC], mx[VEC];
  int amn[VEC], amx[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { s1[q] = 0.f; s2[q] = 0.f; mn[q] = INFINITY; mx[q] = -INFINITY; amn[q] = -1; amx[q] = -1; }
  int k = beg;
  for (; k + 2 <= end; k += 2) {
    int j[2], e[2];
    Vec<VEC> xv[2], ev[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) { j[u] = __ldg(col + k + u); e[u] = __ldg(perm + k + u); }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      xv[u] = Vec<VEC>::load(x + (size_t)j[u] * F + f);
      ev[u] = Vec<VEC>::load_stream(ea + (size_t)e[u] * F + f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const float m = act_fwd_rt(act, xv[u].v[q] + ev[u].v[q]);
        s1[q] += m;
        s2[q] += m * m;
        if (m < mn[q]) { mn[q] = m; amn[q] = e[u]; }
        if (m > mx[q]) { mx[q] = m; amx[q] = e[u]; }
      }
    }
  }
  for (; k < end; ++k) {
    const int j = __ldg(col + k), e = __ldg(perm + k);
    const Vec<VEC> xv = Vec<VEC>::load(x + (size_t)j * F + f);
    const Vec<VEC> ev = Vec<VEC>::load_stream(ea + (size_t)e * F + f);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float m = act_fwd_rt(act, xv.v[q] + ev.v[q]);
      s1[q] += m;
      s2[q] += m * m;
      if (m < mn[q]) { mn[q] = m; amn[q] = e; }
      if (m > mx[q]) { mx[q] = m; amx[q] = e; }
    }
  }
  const int deg = end - beg;
  const float inv = 1.f / (float)max(deg, 1);
  float mean[VEC], var[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    mean[q] = s1[q] * inv;
    var[q] = s2[q] * inv - mean[q] * mean[q];          // aggregate_var: mean(m^2) - mean(m)^2 (aggregator.py:86-89)
    if (deg == 0) { mn[q] = 0.f; mx[q] = 0.f; }
  }
  const size_t off = (size_t)i * F + f;
  if (aux_f) {
    Vec<VEC> a, b;
#pragma unroll
    for (int q = 0; q < VEC; ++q) { a.v[q] = mean[q]; b.v[q] = var[q]; }
    a.store(aux_f + off);
    b.store(aux_f + (size_t)N * F + off);
  }
  if (aux_i) {
    IVec<VEC> a, b;
#pragma unroll
    for (int q = 0; q < VEC; ++q) { a.v[q] = amn[q]; b.v[q] = amx[q]; }
    a.store(aux_i + off);
    b.store(aux_i + (size_t)N * F + off);
  }
  float sc[PNA_MAX_LIST];
#pragma unroll
  for (int s = 0; s < PNA_MAX_LIST; ++s) sc[s] = s < cfg.S ? pna_scale(cfg.scaler[s], deg, cfg) : 0.f;
  float* orow = out + (size_t)i * cfg.S * cfg.T * F;
  const bool vec_store = VEC == 4 && (Fc & 3) == 0;
  const int c0 = f / Fc, j0 = f - c0 * Fc;
  for (int tt = 0; tt < cfg.T; ++tt) {
    float val[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      switch (cfg.aggr[tt]) {
        case PNA_SUM: val[q] = s1[q]; break;
        case PNA_MEAN: val[q] = mean[q]; break;
        case PNA_MIN: val[q] = mn[q]; break;
        case PNA_MAX: val[q] = mx[q]; break;
        case PNA_VAR: val[q] = var[q]; break;
        default: val[q] = sqrtf(fmaxf(var[q], 0.f) + 1e-5f); break;
      }
    }
    for (int s = 0; s < cfg.S; ++s) {
      if (vec_store) {
        Vec<VEC> o;
#pragma unroll
        for (int q = 0; q < VEC; ++q) o.v[q] = val[q] * sc[s];
        o.store(orow + pna_col(c0, j0, s, tt, Fc, cfg.S, cfg.T));
      } else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const int c = (f + q) / Fc, j = (f + q) - c * Fc;
          orow[pna_col(c, j, s, tt, Fc, cfg.S, cfg.T)] = val[q] * sc[s];
        }
      }
    }
  }
}

// per-edge gradient rows (original edge order), one thread per (target row, VEC features)
template <int VEC>
__global__ void __launch_bounds__(256) pna_bwd_edge_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ ea, const float* __restrict__ aux_f,
                                                           const int* __restrict__ aux_i, const int* __restrict__ rowptr,
                                                           const int* __restrict__ col, const int* __restrict__ perm, int N, int F, int Fc,
                                                           int act, PnaCfg cfg, float* __restrict__ dea) {
  pdl_begin();
  const int fv = F / VEC;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int i = (int)(t / fv);
  const int f = (int)(t % fv) * VEC;
  const int beg = rowptr[i], end = rowptr[i + 1];
  const int deg = end - beg;
  if (deg == 0) return;
  const size_t off = (size_t)i * F + f;
  float sc[PNA_MAX_LIST];
#pragma unroll
  for (int s = 0; s < PNA_MAX_LIST; ++s) sc[s] = s < cfg.S ? pna_scale(cfg.scaler[s], deg, cfg) : 0.f;
  // gradient with respect to each aggregator's value: sum over scalers
  float d_const[VEC], d_min[VEC], d_max[VEC], d_var[VEC], mean[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { d_const[q] = 0.f; d_min[q] = 0.f; d_max[q] = 0.f; d_var[q] = 0.f; mean[q] = 0.f; }
  const float inv = 1.f / (float)deg;
  const float* grow = g + (size_t)i * cfg.S * cfg.T * F;
  Vec<VEC> varv;
#pragma unroll
  for (int q = 0; q < VEC; ++q) varv.v[q] = 0.f;
  if (aux_f) {
    Vec<VEC> mv = Vec<VEC>::load(aux_f + off);
    varv = Vec<VEC>::load(aux_f + (size_t)N * F + off);
#pragma unroll
    for (int q = 0; q < VEC; ++q) mean[q] = mv.v[q];
  }
  for (int tt = 0; tt < cfg.T; ++tt) {
    float ga[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const int c = (f + q) / Fc, j = (f + q) - c * Fc;
      float a = 0.f;
      for (int s = 0; s < cfg.S; ++s) a += sc[s] * __ldg(grow + pna_col(c, j, s, tt, Fc, cfg.S, cfg.T));
      ga[q] = a;
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      switch (cfg.aggr[tt]) {
        case PNA_SUM: d_const[q] += ga[q]; break;
        case PNA_MEAN: d_const[q] += ga[q] * inv; break;
        case PNA_MIN: d_min[q] += ga[q]; break;
        case PNA_MAX: d_max[q] += ga[q]; break;
        case PNA_VAR: d_var[q] += ga[q]; break;
        default: {   // std = sqrt(relu(var) + 1e-5)
          const float v = varv.v[q];
          if (v > 0.f) d_var[q] += ga[q] * 0.5f / sqrtf(v + 1e-5f);
          break;
        }
      }
    }
  }
  IVec<VEC> amn, amx;
#pragma unroll
  for (int q = 0; q < VEC; ++q) { amn.v[q] = -1; amx.v[q] = -1; }
  if (aux_i) { amn = IVec<VEC>::load(aux_i + off); amx = IVec<VEC>::load(aux_i + (size_t)N * F + off); }
  for (int k = beg; k < end; ++k) {
    const int e = __ldg(perm + k), j = __ldg(col + k);
    const Vec<VEC> xv = Vec<VEC>::load(x + (size_t)j * F + f);
    const Vec<VEC> ev = Vec<VEC>::load_stream(ea + (size_t)e * F + f);
    Vec<VEC> d;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float pre = xv.v[q] + ev.v[q];
      const float m = act_fwd_rt(act, pre);
      float dm = d_const[q] + d_var[q] * 2.f * (m - mean[q]) * inv;
      if (amn.v[q] == e) dm += d_min[q];
      if (amx.v[q] == e) dm += d_max[q];
      d.v[q] = dm * act_bwd_rt(act, pre);
    }
    d.store_stream(dea + (size_t)e * F + f);
  }
}

int decode_list(unsigned long long code, int* out, int max_id, const char* what) {
  int n = 0;
  while (code & 15ull) {
    const int id = (int)(code & 15ull);
    if (id > max_id || n >= PNA_MAX_LIST) { phc_set_error("phc_pna: bad %s list code", what); return -1; }
    out[n++] = id;
    code >>= 4;
  }
  if (code != 0 || n == 0) { phc_set_error("phc_pna: bad %s list code", what); return -1; }
  return n;
}

int make_cfg(unsigned long long aggr_code, unsigned long long scaler_code, float avg_log, float avg_lin, PnaCfg& cfg) {
  for (int q = 0; q < PNA_MAX_LIST; ++q) { cfg.aggr[q] = 0; cfg.scaler[q] = 0; }
  cfg.T = decode_list(aggr_code, cfg.aggr, PNA_STD, "aggregator");
  cfg.S = decode_list(scaler_code, cfg.scaler, PNA_INVERSE_LINEAR, "scaler");
  if (cfg.T < 0 || cfg.S < 0) return PHC_ERR_INVALID;
  cfg.avg_log = avg_log;
  cfg.avg_lin = avg_lin;
  return PHC_OK;
}

}  // namespace

extern "C" {

int phc_pna_aggregate_fwd(const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int num_nodes, int width,
                          int phm_dim, int msg_act, unsigned long long aggr_code, unsigned long long scaler_code, float avg_deg_log,
                          float avg_deg_lin, float* out, float* aux_f, int* aux_i, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_pna_aggregate_fwd: width %d not divisible by phm_dim %d", width, phm_dim);
  PHC_REQUIRE(msg_act >= PHC_ACT_IDENTITY && msg_act <= PHC_ACT_SWISH, "phc_pna_aggregate_fwd: bad msg_act %d", msg_act);
  PnaCfg cfg;
  if (make_cfg(aggr_code, scaler_code, avg_deg_log, avg_deg_lin, cfg)) return PHC_ERR_INVALID;
  if (num_nodes == 0) return PHC_OK;
  const int N = num_nodes, F = width, Fc = width / phm_dim;
  const bool v4 = F % 4 == 0 && phc_aligned16(x) && phc_aligned16(ea) && phc_aligned16(out) && phc_aligned16(aux_f) && phc_aligned16(aux_i);
  if (v4)
    phc_launch(pna_fwd_kernel<4>, dim3(phc_div_up((long long)N * (F / 4), 256)), dim3(256), 0, stream, x, ea, rowptr, col, perm, N, F, Fc, msg_act, cfg, out, aux_f, aux_i);
  else
    phc_launch(pna_fwd_kernel<1>, dim3(phc_div_up((long long)N * F, 256)), dim3(256), 0, stream, x, ea, rowptr, col, perm, N, F, Fc, msg_act, cfg, out, aux_f, aux_i);
  return phc_check_launch("phc_pna_aggregate_fwd");
}

int phc_pna_aggregate_bwd(const float* gout, const float* x, const float* ea, const float* aux_f, const int* aux_i, const int* rowptr,
                          const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                          int width, int phm_dim, int msg_act, unsigned long long aggr_code, unsigned long long scaler_code,
                          float avg_deg_log, float avg_deg_lin, float* dx, float* dea, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_pna_aggregate_bwd: width %d not divisible by phm_dim %d", width, phm_dim);
  PnaCfg cfg;
  if (make_cfg(aggr_code, scaler_code, avg_deg_log, avg_deg_lin, cfg)) return PHC_ERR_INVALID;
  bool need_f = false, need_i = false;
  for (int t = 0; t < cfg.T; ++t) {
    need_f |= cfg.aggr[t] == PNA_VAR || cfg.aggr[t] == PNA_STD;
    need_i |= cfg.aggr[t] == PNA_MIN || cfg.aggr[t] == PNA_MAX;
  }
  PHC_REQUIRE(!need_f || aux_f, "phc_pna_aggregate_bwd: var/std need aux_f from the forward call");
  PHC_REQUIRE(!need_i || aux_i, "phc_pna_aggregate_bwd: min/max need aux_i from the forward call");
  if (num_nodes == 0) return PHC_OK;
  const int N = num_nodes, F = width, Fc = width / phm_dim;
  const bool v4 = F % 4 == 0 && phc_aligned16(x) && phc_aligned16(ea) && phc_aligned16(dea) && phc_aligned16(aux_f) && phc_aligned16(aux_i);
  if (v4)
    phc_launch(pna_bwd_edge_kernel<4>, dim3(phc_div_up((long long)N * (F / 4), 256)), dim3(256), 0, stream, gout, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, Fc, msg_act, cfg, dea);
  else
    phc_launch(pna_bwd_edge_kernel<1>, dim3(phc_div_up((long long)N * F, 256)), dim3(256), 0, stream, gout, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, Fc, msg_act, cfg, dea);
  int rc = phc_check_launch("phc_pna_aggregate_bwd(edge)");
  if (rc) return rc;
  return phc_aggregate_bwd_node_from_edges(dea, rowptr_t, col_t, perm_t, N, F, dx, stream);
}

}  // extern "C"
