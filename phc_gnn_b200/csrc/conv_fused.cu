// Neighbour aggregation with the per-layer bond/edge ENCODER fused in:
//
//   out[i,:] = (self ? x[i,:] : 0) + AGG_{e: dst(e)=i} phi( x[src(e),:] + enc(edge_attr[e]) )
//   enc = n x Linear(D -> F/n) (float edge features)  or  n x sum of per-column embeddings (integer features)
//
// Replaces reference phc/hypercomplex/undirectional/models.py:238-243 (bondencoders[i](edge_attr) -> [E,F] tensor
// -> conv) + messagepassing.py:72-74,136-138,297-300.  The [E,F] edge embedding is never written: each block keeps
// the encoder parameters as a small [R,F] table in shared memory (R = D+1 rows for Linear incl. bias, total
// vocabulary for embeddings) and rebuilds an edge's embedding from its D raw features (28 B instead of 2000 B per
// edge at ppa shape).  Backward never writes the [E,F] edge gradient either: encoder-parameter gradients are
// accumulated per block in shared memory (one private column per thread, one slot per concurrently processed row;
// no atomics) and block partials are summed in block order.
// Roofline: the x[src] row gather (E*F*4 bytes, served by L2 — x is N*F*4 = 31 MB) instead of HBM streaming.
#include "common.cuh"
#include <float.h>

// simple (sum/mean, identity message) input gradient lives in aggregate.cu
int phc_aggregate_bwd_node_simple(bool mean, const float* g, const int* rowptr, const int* rowptr_t, const int* col_t, const int* perm_t,
                                  int N, int F, int self_loop, float* dx, cudaStream_t stream);

#define ENC_LINEAR 0
#define ENC_EMBED 1
#define ENC_MAX_COLS 16
#define ENC_MAX_PTRS 128

struct EncDesc {
  int kind, D, R, Fc, n;            // D: Linear input width / number of integer columns; R: table rows
  int voff[ENC_MAX_COLS], vsz[ENC_MAX_COLS];
  const float* p[ENC_MAX_PTRS];     // Linear: weights[c] (c < n) then biases[n + c]; embed: tables[c*D + col]
  float* dp[ENC_MAX_PTRS];          // gradients, same order (backward only)
};

namespace {

template <int RED> struct Acc {
  float a, s, t;
  int arg;
  __device__ __forceinline__ void init() {
    a = (RED == PHC_RED_MAX || RED == PHC_RED_SOFTMAX) ? -INFINITY : (RED == PHC_RED_MIN ? INFINITY : 0.f);
    s = 0.f; t = 0.f; arg = -1;
  }
  __device__ __forceinline__ void push(float m, int e, float beta) {
    if (RED == PHC_RED_SUM || RED == PHC_RED_MEAN) a += m;
    else if (RED == PHC_RED_MAX) { if (m > a) { a = m; arg = e; } }
    else if (RED == PHC_RED_MIN) { if (m < a) { a = m; arg = e; } }
    else {
      const float sc = beta * m;
      if (sc > a) { const float r = expf(a - sc); s = s * r + 1.f; t = t * r + m; a = sc; }
      else { const float w = expf(sc - a); s += w; t += w * m; }
    }
  }
};

// cooperative load of the encoder table into shared memory: tab[r*F + f]
__device__ void load_table(const EncDesc& enc, int F, float* tab) {
  for (int idx = threadIdx.x; idx < enc.R * F; idx += blockDim.x) {
    const int r = idx / F, f = idx - r * F;
    const int c = f / enc.Fc, fp = f - c * enc.Fc;
    float v;
    if (enc.kind == ENC_LINEAR) {
      if (r < enc.D) v = __ldg(enc.p[c] + (size_t)fp * enc.D + r);
      else v = enc.p[enc.n + c] ? __ldg(enc.p[enc.n + c] + fp) : 0.f;
    } else {
      int col = 0;
      while (col + 1 < enc.D && enc.voff[col + 1] <= r) ++col;
      v = __ldg(enc.p[c * enc.D + col] + (size_t)(r - enc.voff[col]) * enc.Fc + fp);
    }
    tab[idx] = v;
  }
}

// edge embedding for 4 features starting at f
__device__ __forceinline__ void edge_embed(const EncDesc& enc, const float* tab, int F, const void* attr, int e, int f, float (&ev)[4]) {
  if (enc.kind == ENC_LINEAR) {
    const float4 b = *reinterpret_cast<const float4*>(tab + (size_t)enc.D * F + f);
    ev[0] = b.x; ev[1] = b.y; ev[2] = b.z; ev[3] = b.w;
    const float* a = reinterpret_cast<const float*>(attr) + (size_t)e * enc.D;
    for (int d = 0; d < enc.D; ++d) {
      const float av = __ldg(a + d);
      const float4 w = *reinterpret_cast<const float4*>(tab + (size_t)d * F + f);
      ev[0] += av * w.x; ev[1] += av * w.y; ev[2] += av * w.z; ev[3] += av * w.w;
    }
  } else {
    ev[0] = ev[1] = ev[2] = ev[3] = 0.f;
    const long long* a = reinterpret_cast<const long long*>(attr) + (size_t)e * enc.D;
    for (int col = 0; col < enc.D; ++col) {
      long long v = __ldg(a + col);
      v = v < 0 ? 0 : (v >= enc.vsz[col] ? enc.vsz[col] - 1 : v);
      const float4 w = *reinterpret_cast<const float4*>(tab + (size_t)(enc.voff[col] + (int)v) * F + f);
      ev[0] += w.x; ev[1] += w.y; ev[2] += w.z; ev[3] += w.w;
    }
  }
}

// Linear encoder with compile-time input width DT: the thread's 4 x DT weights and 4 biases live in registers
// (no shared-memory traffic per edge).  DT == 0 selects the generic shared-memory table path above.
template <int DT> struct LinRegs {
  float w[DT > 0 ? DT : 1][4], b[4];
  __device__ __forceinline__ void load(const float* tab, int F, int f) {
    if (DT > 0) {
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        const float4 t = *reinterpret_cast<const float4*>(tab + (size_t)d * F + f);
        w[d][0] = t.x; w[d][1] = t.y; w[d][2] = t.z; w[d][3] = t.w;
      }
      const float4 t = *reinterpret_cast<const float4*>(tab + (size_t)DT * F + f);
      b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
    }
  }
};
template <int DT>
__device__ __forceinline__ void edge_embed_t(const LinRegs<DT>& lr, const EncDesc& enc, const float* tab, int F, const void* attr, int e,
                                             int f, float (&ev)[4]) {
  if (DT > 0) {
    const float* a = reinterpret_cast<const float*>(attr) + (size_t)e * DT;
    float av[DT > 0 ? DT : 1];
#pragma unroll
    for (int d = 0; d < DT; ++d) av[d] = __ldg(a + d);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = lr.b[q];
#pragma unroll
      for (int d = 0; d < DT; ++d) v += av[d] * lr.w[d][q];
      ev[q] = v;
    }
  } else {
    edge_embed(enc, tab, F, attr, e, f, ev);
  }
}

template <int RED, int DT>
__global__ void __launch_bounds__(256, 4) conv_fwd_kernel(const EncDesc enc, const float* __restrict__ x, const void* __restrict__ attr, const int* __restrict__ rowptr,
                                const int* __restrict__ col, const int* __restrict__ perm, int N, int F, int rpi, int act,
                                const float* __restrict__ beta_ptr, int self_loop, float* __restrict__ out, float* __restrict__ aux_f,
                                int* __restrict__ aux_i) {
  pdl_begin();
  extern __shared__ __align__(16) float tab[];
  load_table(enc, F, tab);
  __syncthreads();
  const int fv = F / 4;
  const int slot = threadIdx.x / fv, f = (threadIdx.x - slot * fv) * 4;
  const float beta = (RED == PHC_RED_SOFTMAX) ? __ldg(beta_ptr) : 0.f;
  const bool has_act = act != PHC_ACT_IDENTITY;
  LinRegs<DT> lr;
  lr.load(tab, F, f);
  for (int i = blockIdx.x * rpi + slot; i < N; i += gridDim.x * rpi) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    Acc<RED> acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q].init();
    int k = beg;
    for (; k + 4 <= end; k += 4) {                 // four edges in flight
      int j[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { j[u] = __ldg(col + k + u); e[u] = __ldg(perm + k + u); }
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = *reinterpret_cast<const float4*>(x + (size_t)j[u] * F + f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float ev[4];
        edge_embed_t<DT>(lr, enc, tab, F, attr, e[u], f, ev);
        const float m[4] = {xv[u].x + ev[0], xv[u].y + ev[1], xv[u].z + ev[2], xv[u].w + ev[3]};
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q].push(has_act ? act_fwd_rt(act, m[q]) : m[q], e[u], beta);
      }
    }
    for (; k < end; ++k) {
      const int j0 = __ldg(col + k), e0 = __ldg(perm + k);
      const float4 x0 = *reinterpret_cast<const float4*>(x + (size_t)j0 * F + f);
      float v0[4];
      edge_embed_t<DT>(lr, enc, tab, F, attr, e0, f, v0);
      const float m0[4] = {x0.x + v0[0], x0.y + v0[1], x0.z + v0[2], x0.w + v0[3]};
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q].push(has_act ? act_fwd_rt(act, m0[q]) : m0[q], e0, beta);
    }
    const int deg = end - beg;
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (RED == PHC_RED_SUM) o[q] = acc[q].a;
      else if (RED == PHC_RED_MEAN) o[q] = acc[q].a / (float)max(deg, 1);
      else if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) o[q] = deg > 0 ? acc[q].a : 0.f;
      else o[q] = deg > 0 ? acc[q].t / (acc[q].s + 1e-12f) : 0.f;
    }
    const size_t off = (size_t)i * F + f;
    if (RED == PHC_RED_SOFTMAX) {
      float4 lse;
      lse.x = deg > 0 ? acc[0].a + logf(acc[0].s + 1e-12f) : 0.f; lse.y = deg > 0 ? acc[1].a + logf(acc[1].s + 1e-12f) : 0.f;
      lse.z = deg > 0 ? acc[2].a + logf(acc[2].s + 1e-12f) : 0.f; lse.w = deg > 0 ? acc[3].a + logf(acc[3].s + 1e-12f) : 0.f;
      *reinterpret_cast<float4*>(aux_f + off) = lse;
      *reinterpret_cast<float4*>(aux_f + (size_t)N * F + off) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (RED == PHC_RED_MAX || RED == PHC_RED_MIN)
      *reinterpret_cast<int4*>(aux_i + off) = make_int4(acc[0].arg, acc[1].arg, acc[2].arg, acc[3].arg);
    if (self_loop) {
      const float4 xi = *reinterpret_cast<const float4*>(x + off);
      o[0] += xi.x; o[1] += xi.y; o[2] += xi.z; o[3] += xi.w;
    }
    *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// per-edge d(pre) for 4 features (shared by both backward kernels)
template <int RED>
__device__ __forceinline__ void edge_dpre(int act, bool need_pre, float beta, const float (&pre)[4], const float (&g)[4],
                                          const float (&lse)[4], const float (&agg)[4], const int (&arg)[4], int e, float (&d)[4],
                                          float& dbeta) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float dm;
    if (RED == PHC_RED_SUM || RED == PHC_RED_MEAN) dm = g[q];
    else if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) dm = (arg[q] == e) ? g[q] : 0.f;
    else {
      const float m = act != PHC_ACT_IDENTITY ? act_fwd_rt(act, pre[q]) : pre[q];
      const float w = expf(beta * m - lse[q]);
      const float c = m - agg[q];
      dm = g[q] * w * (1.f + beta * c);
      dbeta += g[q] * w * m * c;
    }
    d[q] = (need_pre && act != PHC_ACT_IDENTITY) ? dm * act_bwd_rt(act, pre[q]) : dm;
  }
}

// encoder-parameter gradients + d(beta): by target rows, shared-memory accumulators, block partials
template <int RED, int DT>
__global__ void conv_bwd_param_kernel(const EncDesc enc, const float* __restrict__ g, const float* __restrict__ x,
                                      const void* __restrict__ attr, const float* __restrict__ aux_f, const int* __restrict__ aux_i,
                                      const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm, int N,
                                      int F, int rpi, int act, const float* __restrict__ beta_ptr, float* __restrict__ part,
                                      float* __restrict__ dbeta_part) {
  pdl_begin();
  extern __shared__ __align__(16) float smem[];
  float* tab = smem;                            // [R][F]
  float* accs = smem + (size_t)enc.R * F;       // [rpi][R][F]
  load_table(enc, F, tab);
  for (int idx = threadIdx.x; idx < rpi * enc.R * F; idx += blockDim.x) accs[idx] = 0.f;
  __syncthreads();
  const int fv = F / 4;
  const int slot = threadIdx.x / fv, f = (threadIdx.x - slot * fv) * 4;
  float* mine = accs + (size_t)slot * enc.R * F;
  const float beta = (RED == PHC_RED_SOFTMAX) ? __ldg(beta_ptr) : 0.f;
  const bool need_pre = act != PHC_ACT_IDENTITY || RED == PHC_RED_SOFTMAX;
  float db = 0.f;
  LinRegs<DT> lr;
  lr.load(tab, F, f);
  float racc[DT > 0 ? DT + 1 : 1][4];          // register accumulators for the small Linear encoders
#pragma unroll
  for (int d = 0; d < (DT > 0 ? DT + 1 : 1); ++d) racc[d][0] = racc[d][1] = racc[d][2] = racc[d][3] = 0.f;
  for (int i = blockIdx.x * rpi + slot; i < N; i += gridDim.x * rpi) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    const size_t off = (size_t)i * F + f;
    const float4 g4 = *reinterpret_cast<const float4*>(g + off);
    float gi[4] = {g4.x, g4.y, g4.z, g4.w};
    float lse[4] = {0.f, 0.f, 0.f, 0.f}, agg[4] = {0.f, 0.f, 0.f, 0.f};
    int arg[4] = {-1, -1, -1, -1};
    if (RED == PHC_RED_SOFTMAX) {
      const float4 l = *reinterpret_cast<const float4*>(aux_f + off), a = *reinterpret_cast<const float4*>(aux_f + (size_t)N * F + off);
      lse[0] = l.x; lse[1] = l.y; lse[2] = l.z; lse[3] = l.w;
      agg[0] = a.x; agg[1] = a.y; agg[2] = a.z; agg[3] = a.w;
    }
    if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) {
      const int4 a = *reinterpret_cast<const int4*>(aux_i + off);
      arg[0] = a.x; arg[1] = a.y; arg[2] = a.z; arg[3] = a.w;
    }
    if (RED == PHC_RED_MEAN) {
      const float inv = 1.f / (float)max(end - beg, 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) gi[q] *= inv;
    }
    for (int k = beg; k < end; ++k) {
      const int e = __ldg(perm + k);
      float pre[4] = {0.f, 0.f, 0.f, 0.f};
      if (need_pre) {
        const int j = __ldg(col + k);
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)j * F + f);
        float ev[4];
        edge_embed_t<DT>(lr, enc, tab, F, attr, e, f, ev);
        pre[0] = xv.x + ev[0]; pre[1] = xv.y + ev[1]; pre[2] = xv.z + ev[2]; pre[3] = xv.w + ev[3];
      }
      float d[4];
      edge_dpre<RED>(act, need_pre, beta, pre, gi, lse, agg, arg, e, d, db);
      if (DT > 0) {
        const float* a = reinterpret_cast<const float*>(attr) + (size_t)e * DT;
#pragma unroll
        for (int dd = 0; dd < DT; ++dd) {
          const float av = __ldg(a + dd);
#pragma unroll
          for (int q = 0; q < 4; ++q) racc[dd][q] += av * d[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) racc[DT > 0 ? DT : 0][q] += d[q];
      } else if (enc.kind == ENC_LINEAR) {
        const float* a = reinterpret_cast<const float*>(attr) + (size_t)e * enc.D;
        for (int dd = 0; dd < enc.D; ++dd) {
          const float av = __ldg(a + dd);
          float4* p = reinterpret_cast<float4*>(mine + (size_t)dd * F + f);
          float4 t = *p;
          t.x += av * d[0]; t.y += av * d[1]; t.z += av * d[2]; t.w += av * d[3];
          *p = t;
        }
        float4* p = reinterpret_cast<float4*>(mine + (size_t)enc.D * F + f);
        float4 t = *p;
        t.x += d[0]; t.y += d[1]; t.z += d[2]; t.w += d[3];
        *p = t;
      } else {
        const long long* a = reinterpret_cast<const long long*>(attr) + (size_t)e * enc.D;
        for (int cc = 0; cc < enc.D; ++cc) {
          long long v = __ldg(a + cc);
          v = v < 0 ? 0 : (v >= enc.vsz[cc] ? enc.vsz[cc] - 1 : v);
          float4* p = reinterpret_cast<float4*>(mine + (size_t)(enc.voff[cc] + (int)v) * F + f);
          float4 t = *p;
          t.x += d[0]; t.y += d[1]; t.z += d[2]; t.w += d[3];
          *p = t;
        }
      }
    }
  }
  if (DT > 0) {
#pragma unroll
    for (int d = 0; d <= DT; ++d) *reinterpret_cast<float4*>(mine + (size_t)d * F + f) = make_float4(racc[d][0], racc[d][1], racc[d][2], racc[d][3]);
  }
  __syncthreads();
  // block partial: slots summed in slot order
  float* dst = part + (size_t)blockIdx.x * enc.R * F;
  for (int idx = threadIdx.x; idx < enc.R * F; idx += blockDim.x) {
    float s = 0.f;
    for (int sl = 0; sl < rpi; ++sl) s += accs[(size_t)sl * enc.R * F + idx];
    dst[idx] = s;
  }
  if (RED == PHC_RED_SOFTMAX) {
    __syncthreads();
    float* red = accs;                 // reuse
    red[threadIdx.x] = db;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int t = 0; t < (int)blockDim.x; ++t) s += red[t];
      dbeta_part[blockIdx.x] = s;
    }
  }
}

// sum block partials and scatter into the individual parameter gradients.  Block = 32 consecutive table elements (lanes,
// coalesced 128-byte reads of part[b][r][f..f+31]) x 32 slices of the partial list (slice s takes blocks s, s+32, ... in
// order); the 32 slice sums are then added in slice order — a fixed association, hence deterministic.
__global__ void __launch_bounds__(1024) conv_bwd_param_final_kernel(const EncDesc enc, const float* __restrict__ part, int blocks, int F) {
  pdl_begin();
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long total = (long long)enc.R * F;
  const long long t = (long long)blockIdx.x * 32 + lane;
  float s = 0.f;
  if (t < total) {
#pragma unroll 4
    for (int b = slice; b < blocks; b += 32) s += part[(size_t)b * total + t];
  }
  red[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || t >= total) return;
  s = 0.f;
#pragma unroll
  for (int q = 0; q < 32; ++q) s += red[q][lane];
  const int r = (int)(t / F), f = (int)(t - (long long)r * F);
  const int c = f / enc.Fc, fp = f - c * enc.Fc;
  if (enc.kind == ENC_LINEAR) {
    if (r < enc.D) enc.dp[c][(size_t)fp * enc.D + r] = s;
    else if (enc.dp[enc.n + c]) enc.dp[enc.n + c][fp] = s;
  } else {
    int col = 0;
    while (col + 1 < enc.D && enc.voff[col + 1] <= r) ++col;
    enc.dp[c * enc.D + col][(size_t)(r - enc.voff[col]) * enc.Fc + fp] = s;
  }
}

// ---- sum / mean aggregation with identity message: d(pre_e) = g[dst(e)] * scale, so the encoder-parameter
// gradient collapses to  dTab[r,f] = sum_i S[i,r] * g[i,f]  with the per-NODE feature sums
//   S[i,r] = scale_i * sum_{e -> i} phi_r(edge_attr[e])    (phi = raw features + constant 1, or one-hot of the integer columns)
// S depends only on the batch (not on the layer), is N x R (R <= 16) and is computed once per batch.
__global__ void __launch_bounds__(256) edge_feature_sums_kernel(const EncDesc enc, const void* __restrict__ attr, const int* __restrict__ rowptr,
                                                                const int* __restrict__ perm, int N, int mean, float* __restrict__ S) {
  pdl_begin();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * enc.R) return;
  const int i = (int)(t / enc.R), r = (int)(t - (long long)i * enc.R);
  const int beg = rowptr[i], end = rowptr[i + 1];
  float s = 0.f;
  if (enc.kind == ENC_LINEAR) {
    if (r == enc.D) s = (float)(end - beg);
    else for (int k = beg; k < end; ++k) s += __ldg(reinterpret_cast<const float*>(attr) + (size_t)__ldg(perm + k) * enc.D + r);
  } else {
    int col = 0;
    while (col + 1 < enc.D && enc.voff[col + 1] <= r) ++col;
    const int v = r - enc.voff[col];
    for (int k = beg; k < end; ++k) {
      long long a = __ldg(reinterpret_cast<const long long*>(attr) + (size_t)__ldg(perm + k) * enc.D + col);
      a = a < 0 ? 0 : (a >= enc.vsz[col] ? enc.vsz[col] - 1 : a);
      s += (a == v) ? 1.f : 0.f;
    }
  }
  S[t] = mean ? s / (float)max(end - beg, 1) : s;
}

// Forward with the per-node feature sums: the encoder is linear in phi(edge_attr) (a Linear layer, or a sum of embedding
// rows = one-hot features), so for sum / mean aggregation with the identity message the edge term leaves the edge loop:
//   out[i,:] = self * x[i,:] + scale_i * sum_{e -> i} x[src(e),:]  +  sum_r S[i,r] * Tab[r,:]
// (S already carries the mean scale).  The kernel is then a pure row gather (same shape as the input-gradient kernel,
// which moves the same 4*E*F bytes through L2 in ~40 us where the per-edge kernel needs ~138 us for 7 FMAs, 7 feature
// loads and 28 weight registers per edge and thread) plus R FMAs per output element.
// the encoder parameters as one [R, F] table in global memory (16 KB at ppa shape: lives in L1 / L2)
__global__ void __launch_bounds__(256) enc_table_kernel(const EncDesc enc, int F, float* __restrict__ tab) {
  pdl_begin();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= enc.R * F) return;
  const int r = idx / F, f = idx - r * F;
  const int c = f / enc.Fc, fp = f - c * enc.Fc;
  float v;
  if (enc.kind == ENC_LINEAR) {
    if (r < enc.D) v = __ldg(enc.p[c] + (size_t)fp * enc.D + r);
    else v = enc.p[enc.n + c] ? __ldg(enc.p[enc.n + c] + fp) : 0.f;
  } else {
    int col = 0;
    while (col + 1 < enc.D && enc.voff[col + 1] <= r) ++col;
    v = __ldg(enc.p[c * enc.D + col] + (size_t)(r - enc.voff[col]) * enc.Fc + fp);
  }
  tab[idx] = v;
}

// one thread per (node, 4 features), flat grid, ~32 registers: full occupancy for the row gather
template <bool MEAN>
__global__ void __launch_bounds__(256) conv_fwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ S,
                                                            const float* __restrict__ tab, const int* __restrict__ rowptr,
                                                            const int* __restrict__ col, int N, int F, int R, int self_loop,
                                                            float* __restrict__ out) {
  pdl_begin();
  const int fv = F / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int i = (int)(t / fv), f = (int)(t - (long long)i * fv) * 4;
  const int beg = rowptr[i], end = rowptr[i + 1];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = beg;
  for (; k + 4 <= end; k += 4) {                 // four rows in flight
    int j[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) j[u] = __ldg(col + k + u);
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(x + (size_t)j[u] * F + f);
#pragma unroll
    for (int u = 0; u < 4; ++u) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
  }
  for (; k < end; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)__ldg(col + k) * F + f);
    a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
  }
  if (MEAN) {
    const float sc = 1.f / (float)max(end - beg, 1);
    a0 *= sc; a1 *= sc; a2 *= sc; a3 *= sc;
  }
  const float* si = S + (size_t)i * R;
  for (int r = 0; r < R; ++r) {
    const float sv = __ldg(si + r);
    const float4 w = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * F + f));
    a0 += sv * w.x; a1 += sv * w.y; a2 += sv * w.z; a3 += sv * w.w;
  }
  const size_t off = (size_t)i * F + f;
  if (self_loop) {
    const float4 xi = *reinterpret_cast<const float4*>(x + off);
    a0 += xi.x; a1 += xi.y; a2 += xi.z; a3 += xi.w;
  }
  *reinterpret_cast<float4*>(out + off) = make_float4(a0, a1, a2, a3);
}

template <int RT>
__global__ void __launch_bounds__(128) conv_bwd_param_simple_kernel(const float* __restrict__ g, const float* __restrict__ S, int N, int F,
                                                                    int rows_per_block, float* __restrict__ part) {
  pdl_begin();
  const int f = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (f >= F) return;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, N);
  float acc[RT][4];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  // eight gradient rows in flight per thread (a block walks only 16 rows: with two loads per trip the kernel ran at one memory latency
  // per pair of rows, 14 us for 31 MB); rows are still accumulated in order
  int i = r0;
  for (; i + 8 <= r1; i += 8) {
    float4 gv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) gv[u] = *reinterpret_cast<const float4*>(g + (size_t)(i + u) * F + f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float sv = __ldg(S + (size_t)(i + u) * RT + r);
        acc[r][0] += sv * gv[u].x; acc[r][1] += sv * gv[u].y; acc[r][2] += sv * gv[u].z; acc[r][3] += sv * gv[u].w;
      }
    }
  }
  for (; i < r1; ++i) {
    const float4 gv = *reinterpret_cast<const float4*>(g + (size_t)i * F + f);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float sv = __ldg(S + (size_t)i * RT + r);
      acc[r][0] += sv * gv.x; acc[r][1] += sv * gv.y; acc[r][2] += sv * gv.z; acc[r][3] += sv * gv.w;
    }
  }
#pragma unroll
  for (int r = 0; r < RT; ++r)
    *reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * RT + r) * F + f) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  pdl_begin();
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// input gradient, general case (message activation and/or max/min/softmax): by source rows, recomputes d(pre)
template <int RED>
__global__ void conv_bwd_node_kernel(const EncDesc enc, const float* __restrict__ g, const float* __restrict__ x,
                                     const void* __restrict__ attr, const float* __restrict__ aux_f, const int* __restrict__ aux_i,
                                     const int* __restrict__ rowptr, const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                     const int* __restrict__ perm_t, int N, int F, int rpi, int act, const float* __restrict__ beta_ptr,
                                     int self_loop, float* __restrict__ dx) {
  pdl_begin();
  extern __shared__ __align__(16) float tab[];
  load_table(enc, F, tab);
  __syncthreads();
  const int fv = F / 4;
  const int slot = threadIdx.x / fv, f = (threadIdx.x - slot * fv) * 4;
  const float beta = (RED == PHC_RED_SOFTMAX) ? __ldg(beta_ptr) : 0.f;
  const bool need_pre = act != PHC_ACT_IDENTITY || RED == PHC_RED_SOFTMAX;
  float unused = 0.f;
  for (int j = blockIdx.x * rpi + slot; j < N; j += gridDim.x * rpi) {
    const size_t offj = (size_t)j * F + f;
    const float4 xj = *reinterpret_cast<const float4*>(x + offj);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (self_loop) {
      const float4 gj = *reinterpret_cast<const float4*>(g + offj);
      acc[0] = gj.x; acc[1] = gj.y; acc[2] = gj.z; acc[3] = gj.w;
    }
    for (int k = rowptr_t[j]; k < rowptr_t[j + 1]; ++k) {
      const int e = __ldg(perm_t + k), i = __ldg(col_t + k);
      const size_t offi = (size_t)i * F + f;
      const float4 g4 = *reinterpret_cast<const float4*>(g + offi);
      float gi[4] = {g4.x, g4.y, g4.z, g4.w};
      float lse[4] = {0.f, 0.f, 0.f, 0.f}, agg[4] = {0.f, 0.f, 0.f, 0.f};
      int arg[4] = {-1, -1, -1, -1};
      if (RED == PHC_RED_SOFTMAX) {
        const float4 l = *reinterpret_cast<const float4*>(aux_f + offi), a = *reinterpret_cast<const float4*>(aux_f + (size_t)N * F + offi);
        lse[0] = l.x; lse[1] = l.y; lse[2] = l.z; lse[3] = l.w;
        agg[0] = a.x; agg[1] = a.y; agg[2] = a.z; agg[3] = a.w;
      }
      if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) {
        const int4 a = *reinterpret_cast<const int4*>(aux_i + offi);
        arg[0] = a.x; arg[1] = a.y; arg[2] = a.z; arg[3] = a.w;
      }
      if (RED == PHC_RED_MEAN) {
        const float inv = 1.f / (float)max(__ldg(rowptr + i + 1) - __ldg(rowptr + i), 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) gi[q] *= inv;
      }
      float pre[4] = {0.f, 0.f, 0.f, 0.f};
      if (need_pre) {
        float ev[4];
        edge_embed(enc, tab, F, attr, e, f, ev);
        pre[0] = xj.x + ev[0]; pre[1] = xj.y + ev[1]; pre[2] = xj.z + ev[2]; pre[3] = xj.w + ev[3];
      }
      float d[4];
      edge_dpre<RED>(act, need_pre, beta, pre, gi, lse, agg, arg, e, d, unused);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += d[q];
    }
    *reinterpret_cast<float4*>(dx + offj) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

int make_desc(EncDesc& d, int kind, int dim, const int* vocab, const float* const* params, float* const* dparams, int n, int F) {
  d.kind = kind; d.D = dim; d.n = n; d.Fc = F / n;
  const int nptr = kind == ENC_LINEAR ? 2 * n : n * dim;
  if (nptr > ENC_MAX_PTRS || dim > ENC_MAX_COLS || dim < 1) return -1;
  int r = 0;
  for (int c = 0; c < ENC_MAX_COLS; ++c) { d.voff[c] = 0; d.vsz[c] = 0; }
  if (kind == ENC_EMBED) {
    for (int c = 0; c < dim; ++c) { d.voff[c] = r; d.vsz[c] = vocab[c]; r += vocab[c]; }
  } else {
    r = dim + 1;
  }
  d.R = r;
  for (int i = 0; i < ENC_MAX_PTRS; ++i) { d.p[i] = i < nptr ? params[i] : nullptr; d.dp[i] = (dparams && i < nptr) ? dparams[i] : nullptr; }
  return 0;
}

struct Geometry { int rpi, threads, blocks; };
Geometry geometry(int N, int F) {
  Geometry g;
  const int fv = F / 4;
  g.rpi = 256 / fv; if (g.rpi < 1) g.rpi = 1;
  g.threads = g.rpi * fv;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int want = 4 * sms;
  const int groups = (N + g.rpi - 1) / g.rpi;
  g.blocks = groups < want ? (groups < 1 ? 1 : groups) : want;
  return g;
}

template <typename K>
bool ensure_smem(K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return true;
  if (bytes > 200 * 1024) return false;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
}

}  // namespace

extern "C" {

// 1 if the fused path can run this configuration (otherwise the caller uses encoder + phc_aggregate_*)
int phc_conv_fused_supported(int width, int phm_dim, int enc_kind, int enc_dim, int table_rows) {
  if (width % 4 != 0 || width / 4 > 256 || phm_dim < 1 || width % phm_dim != 0) return 0;
  if (enc_dim < 1 || enc_dim > ENC_MAX_COLS) return 0;
  if ((enc_kind == ENC_LINEAR ? 2 * phm_dim : phm_dim * enc_dim) > ENC_MAX_PTRS) return 0;
  int rpi = 256 / (width / 4); if (rpi < 1) rpi = 1;
  const size_t bwd = sizeof(float) * (size_t)(1 + rpi) * table_rows * width;
  return bwd <= 200 * 1024 ? 1 : 0;
}

int phc_conv_fused_fwd(const float* x, const void* edge_attr, int enc_kind, int enc_dim, const int* vocab, const float* const* params,
                       const int* rowptr, const int* col, const int* perm, int num_nodes, int width, int phm_dim, int reduce,
                       int msg_act, const float* beta, int self_loop, float* out, float* aux_f, int* aux_i, cudaStream_t stream) {
  PHC_REQUIRE(reduce >= PHC_RED_SUM && reduce <= PHC_RED_SOFTMAX, "phc_conv_fused_fwd: bad reduce %d", reduce);
  PHC_REQUIRE(msg_act >= PHC_ACT_IDENTITY && msg_act <= PHC_ACT_SWISH, "phc_conv_fused_fwd: bad msg_act %d", msg_act);
  EncDesc d;
  PHC_REQUIRE(make_desc(d, enc_kind, enc_dim, vocab, params, nullptr, phm_dim, width) == 0, "phc_conv_fused_fwd: unsupported encoder");
  PHC_REQUIRE(phc_conv_fused_supported(width, phm_dim, enc_kind, enc_dim, d.R), "phc_conv_fused_fwd: unsupported shape");
  if (num_nodes == 0) return PHC_OK;
  const Geometry g = geometry(num_nodes, width);
  const size_t smem = sizeof(float) * (size_t)d.R * width;
  const int dt = (enc_kind == ENC_LINEAR && enc_dim <= 8) ? enc_dim : 0;
#define PHC_LAUNCH2(RED, DT)                                                                                                     \
  { if (!ensure_smem(conv_fwd_kernel<RED, DT>, smem)) { phc_set_error("phc_conv_fused_fwd: shared memory"); return PHC_ERR_CUDA; } \
    phc_launch(conv_fwd_kernel<RED, DT>, dim3(g.blocks), dim3(g.threads), smem, stream, d, x, edge_attr, rowptr, col, perm, num_nodes, width, g.rpi, msg_act, \
                                                                    beta, self_loop, out, aux_f, aux_i); }
#define PHC_LAUNCH(RED)                                                                                                          \
  switch (dt) {                                                                                                                  \
    case 1: PHC_LAUNCH2(RED, 1) break; case 2: PHC_LAUNCH2(RED, 2) break; case 3: PHC_LAUNCH2(RED, 3) break;                     \
    case 4: PHC_LAUNCH2(RED, 4) break; case 5: PHC_LAUNCH2(RED, 5) break; case 6: PHC_LAUNCH2(RED, 6) break;                     \
    case 7: PHC_LAUNCH2(RED, 7) break; case 8: PHC_LAUNCH2(RED, 8) break; default: PHC_LAUNCH2(RED, 0) break; }
  switch (reduce) {
    case PHC_RED_SUM: PHC_LAUNCH(PHC_RED_SUM) break;
    case PHC_RED_MEAN: PHC_LAUNCH(PHC_RED_MEAN) break;
    case PHC_RED_MAX: PHC_LAUNCH(PHC_RED_MAX) break;
    case PHC_RED_MIN: PHC_LAUNCH(PHC_RED_MIN) break;
    default: PHC_LAUNCH(PHC_RED_SOFTMAX) break;
  }
#undef PHC_LAUNCH
#undef PHC_LAUNCH2
  return phc_check_launch("phc_conv_fused_fwd");
}

size_t phc_conv_fused_fwd_sums_workspace_bytes(int width, int table_rows) { return sizeof(float) * (size_t)table_rows * width + 16; }

int phc_conv_fused_fwd_sums(const float* x, const float* node_sums, int enc_kind, int enc_dim, const int* vocab, const float* const* params,
                            const int* rowptr, const int* col, int num_nodes, int width, int phm_dim, int reduce, int self_loop,
                            float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(reduce == PHC_RED_SUM || reduce == PHC_RED_MEAN, "phc_conv_fused_fwd_sums: sum / mean only (got reduce %d)", reduce);
  PHC_REQUIRE(node_sums != nullptr, "phc_conv_fused_fwd_sums: node_sums required");
  EncDesc d;
  PHC_REQUIRE(make_desc(d, enc_kind, enc_dim, vocab, params, nullptr, phm_dim, width) == 0, "phc_conv_fused_fwd_sums: unsupported encoder");
  PHC_REQUIRE(phc_conv_fused_supported(width, phm_dim, enc_kind, enc_dim, d.R), "phc_conv_fused_fwd_sums: unsupported shape");
  PHC_REQUIRE(workspace != nullptr && workspace_bytes >= phc_conv_fused_fwd_sums_workspace_bytes(width, d.R) && phc_aligned16(workspace),
              "phc_conv_fused_fwd_sums: workspace too small or misaligned");
  if (num_nodes == 0) return PHC_OK;
  float* tab = reinterpret_cast<float*>(workspace);
  phc_launch(enc_table_kernel, dim3(phc_div_up((long long)d.R * width, 256)), dim3(256), 0, stream, d, width, tab);
  const int grid = phc_div_up((long long)num_nodes * (width / 4), 256);
  if (reduce == PHC_RED_MEAN)
    phc_launch(conv_fwd_sums_kernel<true>, dim3(grid), dim3(256), 0, stream, x, node_sums, tab, rowptr, col, num_nodes, width, d.R, self_loop, out);
  else
    phc_launch(conv_fwd_sums_kernel<false>, dim3(grid), dim3(256), 0, stream, x, node_sums, tab, rowptr, col, num_nodes, width, d.R, self_loop, out);
  return phc_check_launch("phc_conv_fused_fwd_sums");
}

int phc_edge_feature_sums(const void* edge_attr, int enc_kind, int enc_dim, const int* vocab, const int* rowptr, const int* perm,
                          int num_nodes, int mean, float* node_sums, cudaStream_t stream) {
  EncDesc d;
  const float* none[ENC_MAX_PTRS] = {nullptr};
  PHC_REQUIRE(make_desc(d, enc_kind, enc_dim, vocab, none, nullptr, 1, 4) == 0, "phc_edge_feature_sums: unsupported encoder");
  if (num_nodes == 0) return PHC_OK;
  phc_launch(edge_feature_sums_kernel, dim3(phc_div_up((long long)num_nodes * d.R, 256)), dim3(256), 0, stream, d, edge_attr, rowptr, perm, num_nodes, mean,
                                                                                           node_sums);
  return phc_check_launch("phc_edge_feature_sums");
}

size_t phc_conv_fused_bwd_workspace_bytes(int num_nodes, int width, int table_rows) {
  const Geometry g = geometry(num_nodes, width);
  size_t blocks = (size_t)g.blocks;
  const size_t nb = (size_t)phc_div_up(num_nodes, 16);
  if (nb > blocks) blocks = nb;
  return sizeof(float) * (blocks * table_rows * width + blocks) + 64;
}

int phc_conv_fused_bwd(const float* gout, const float* x, const void* edge_attr, int enc_kind, int enc_dim, const int* vocab,
                       const float* const* params, float* const* dparams, const float* aux_f, const int* aux_i, const int* rowptr,
                       const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                       int width, int phm_dim, int reduce, int msg_act, const float* beta, int self_loop, const float* node_sums,
                       float* dx, float* dbeta, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  EncDesc d;
  PHC_REQUIRE(make_desc(d, enc_kind, enc_dim, vocab, params, dparams, phm_dim, width) == 0, "phc_conv_fused_bwd: unsupported encoder");
  PHC_REQUIRE(phc_conv_fused_supported(width, phm_dim, enc_kind, enc_dim, d.R), "phc_conv_fused_bwd: unsupported shape");
  PHC_REQUIRE(workspace_bytes >= phc_conv_fused_bwd_workspace_bytes(num_nodes, width, d.R), "phc_conv_fused_bwd: workspace too small");
  const int N = num_nodes, F = width;
  const Geometry g = geometry(N, F);
  float* part = reinterpret_cast<float*>(workspace);
  float* dbp = part + (size_t)g.blocks * d.R * F;
  const size_t smem_p = sizeof(float) * (size_t)(1 + g.rpi) * d.R * F;
  const size_t smem_n = sizeof(float) * (size_t)d.R * F;
  const bool simple = (reduce == PHC_RED_SUM || reduce == PHC_RED_MEAN) && msg_act == PHC_ACT_IDENTITY;
  if (simple && node_sums != nullptr && d.R <= 16 && N > 0) {
    // parameter gradients from the per-node feature sums: dTab = S^T g (reads g once, no edge loop)
    const int rpb = 16;
    const int nb = phc_div_up(N, rpb);
    PHC_REQUIRE((size_t)nb * d.R * F * sizeof(float) <= workspace_bytes, "phc_conv_fused_bwd: workspace too small for the node-sum path");
    dim3 grid(nb, phc_div_up(F / 4, 128));
    switch (d.R) {
#define PHC_CASE(RT) case RT: phc_launch(conv_bwd_param_simple_kernel<RT>, dim3(grid), dim3(128), 0, stream, gout, node_sums, N, F, rpb, part); break;
      PHC_CASE(1) PHC_CASE(2) PHC_CASE(3) PHC_CASE(4) PHC_CASE(5) PHC_CASE(6) PHC_CASE(7) PHC_CASE(8)
      PHC_CASE(9) PHC_CASE(10) PHC_CASE(11) PHC_CASE(12) PHC_CASE(13) PHC_CASE(14) PHC_CASE(15) PHC_CASE(16)
#undef PHC_CASE
    }
    phc_launch(conv_bwd_param_final_kernel, dim3(phc_div_up((long long)d.R * F, 32)), dim3(1024), 0, stream, d, part, nb, F);
    int rc = phc_check_launch("phc_conv_fused_bwd(node sums)");
    if (rc) return rc;
    if (dx) return phc_aggregate_bwd_node_simple(reduce == PHC_RED_MEAN, gout, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
    return PHC_OK;
  }
  const int dt = (enc_kind == ENC_LINEAR && enc_dim <= 8) ? enc_dim : 0;
#define PHC_PARAM(RED, DT)                                                                                                            \
  { if (!ensure_smem(conv_bwd_param_kernel<RED, DT>, smem_p)) { phc_set_error("phc_conv_fused_bwd: shared memory"); return PHC_ERR_CUDA; } \
    if (N > 0) phc_launch(conv_bwd_param_kernel<RED, DT>, dim3(g.blocks), dim3(g.threads), smem_p, stream, d, gout, x, edge_attr, aux_f, aux_i, rowptr, col, perm, \
                                                                                       N, F, g.rpi, msg_act, beta, part, dbp); }
#define PHC_LAUNCH(RED)                                                                                                               \
  if (!ensure_smem(conv_bwd_node_kernel<RED>, smem_n)) { phc_set_error("phc_conv_fused_bwd: shared memory"); return PHC_ERR_CUDA; }   \
  switch (dt) {                                                                                                                       \
    case 1: PHC_PARAM(RED, 1) break; case 2: PHC_PARAM(RED, 2) break; case 3: PHC_PARAM(RED, 3) break; case 4: PHC_PARAM(RED, 4) break; \
    case 5: PHC_PARAM(RED, 5) break; case 6: PHC_PARAM(RED, 6) break; case 7: PHC_PARAM(RED, 7) break; case 8: PHC_PARAM(RED, 8) break; \
    default: PHC_PARAM(RED, 0) break; }                                                                                               \
  if (N > 0 && dx && !simple) phc_launch(conv_bwd_node_kernel<RED>, dim3(g.blocks), dim3(g.threads), smem_n, stream, d, gout, x, edge_attr, aux_f, aux_i, rowptr,   \
                                                                                        rowptr_t, col_t, perm_t, N, F, g.rpi, msg_act, \
                                                                                        beta, self_loop, dx);
  switch (reduce) {
    case PHC_RED_SUM: PHC_LAUNCH(PHC_RED_SUM) break;
    case PHC_RED_MEAN: PHC_LAUNCH(PHC_RED_MEAN) break;
    case PHC_RED_MAX: PHC_LAUNCH(PHC_RED_MAX) break;
    case PHC_RED_MIN: PHC_LAUNCH(PHC_RED_MIN) break;
    default: PHC_LAUNCH(PHC_RED_SOFTMAX) break;
  }
#undef PHC_LAUNCH
#undef PHC_PARAM
  const int blocks_used = N > 0 ? g.blocks : 0;
  phc_launch(conv_bwd_param_final_kernel, dim3(phc_div_up((long long)d.R * F, 32)), dim3(1024), 0, stream, d, part, blocks_used, F);
  if (reduce == PHC_RED_SOFTMAX && dbeta) phc_launch(sum_partials_kernel, dim3(1), dim3(256), 0, stream, dbp, blocks_used, dbeta);
  int rc = phc_check_launch("phc_conv_fused_bwd");
  if (rc) return rc;
  if (dx && simple && N > 0)
    return phc_aggregate_bwd_node_simple(reduce == PHC_RED_MEAN, gout, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
  return PHC_OK;
}

}  // extern "C"
