// Fused neighbour aggregation: CSR-sorted segmented gather-reduce, forward and backward.
//
//   out[i,:] = (self_loop ? x[i,:] : 0) + AGG_{e : dst(e)=i} phi( x[src(e),:] + ea[e,:] )
//
// Replaces the reference's PyG propagate pipeline (phc/hypercomplex/undirectional/messagepassing.py
// :72-74 message, :136/:306 propagate, :297-300 softmax aggregate, :137-138 self-loop add):
// index_select -> add -> activation -> torch_scatter.scatter / scatter_softmax+scatter_sum, i.e.
// up to ten launches, six [E,F] temporaries and float atomics.  Here every edge row is read
// exactly once, rows are reduced in CSR order by a single thread per 4 features (no atomics, so
// the result is bit-reproducible), and the n hypercomplex components of a node live contiguously
// in one flat [*,F] row so one 128-bit load serves all of them.
//
// Data layout: x [N,F] fp32 row-major; ea [E,F] fp32 in ORIGINAL edge order (perm maps CSR slots
// to edge ids); rowptr/col/perm int32 from phc_csr_build.
// Roofline: HBM.  Algorithmic bytes fwd = 4F(2N+E) + 8E + 4(N+1) (+8NF for softmax aux).
#include "common.cuh"
#include <float.h>

namespace {

template <int RED> struct Acc {
  float a, s, t;  // a: sum / extremum / running max of beta*m ; s: softmax denominator ; t: softmax numerator
  int arg;
  __device__ __forceinline__ void init() {
    a = (RED == PHC_RED_MAX || RED == PHC_RED_SOFTMAX) ? -INFINITY : (RED == PHC_RED_MIN ? INFINITY : 0.f);
    s = 0.f; t = 0.f; arg = -1;
  }
  __device__ __forceinline__ void push(float m, int e, float beta) {
    if (RED == PHC_RED_SUM || RED == PHC_RED_MEAN) {
      a += m;
    } else if (RED == PHC_RED_MAX) {
      if (m > a) { a = m; arg = e; }
    } else if (RED == PHC_RED_MIN) {
      if (m < a) { a = m; arg = e; }
    } else {  // online softmax over scores beta*m, weighted sum of m
      float sc = beta * m;
      if (sc > a) {
        float r = expf(a - sc);  // exp(-inf) = 0 on the first edge
        s = s * r + 1.f;
        t = t * r + m;
        a = sc;
      } else {
        float w = expf(sc - a);
        s += w;
        t += w * m;
      }
    }
  }
};

template <int VEC, int RED, bool HAS_ACT>
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ea,
                                                            const int* __restrict__ rowptr, const int* __restrict__ col,
                                                            const int* __restrict__ perm, int N, int F, int act,
                                                            const float* __restrict__ beta_ptr, int self_loop, float* __restrict__ out,
                                                            float* __restrict__ aux_f, int* __restrict__ aux_i) {
  pdl_begin();
  const int fv = F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int i = (int)(t / fv);
  const int f = (int)(t % fv) * VEC;
  const int beg = rowptr[i], end = rowptr[i + 1];
  const float beta = (RED == PHC_RED_SOFTMAX) ? __ldg(beta_ptr) : 0.f;
  Acc<RED> acc[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc[q].init();

  int k = beg;
  for (; k + 4 <= end; k += 4) {
    int j[4], e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { j[u] = __ldg(col + k + u); e[u] = __ldg(perm + k + u); }
    Vec<VEC> xv[4], ev[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = Vec<VEC>::load(x + (size_t)j[u] * F + f);
      ev[u] = Vec<VEC>::load_stream(ea + (size_t)e[u] * F + f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float m = xv[u].v[q] + ev[u].v[q];
        if (HAS_ACT) m = act_fwd_rt(act, m);
        acc[q].push(m, e[u], beta);
      }
    }
  }
  for (; k < end; ++k) {
    int j = __ldg(col + k), e = __ldg(perm + k);
    Vec<VEC> xv = Vec<VEC>::load(x + (size_t)j * F + f);
    Vec<VEC> ev = Vec<VEC>::load_stream(ea + (size_t)e * F + f);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float m = xv.v[q] + ev.v[q];
      if (HAS_ACT) m = act_fwd_rt(act, m);
      acc[q].push(m, e, beta);
    }
  }

  const int deg = end - beg;
  Vec<VEC> o;
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    float r;
    if (RED == PHC_RED_SUM) r = acc[q].a;
    else if (RED == PHC_RED_MEAN) r = acc[q].a / (float)max(deg, 1);
    else if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) r = deg > 0 ? acc[q].a : 0.f;
    else r = deg > 0 ? acc[q].t / (acc[q].s + 1e-12f) : 0.f;
    o.v[q] = r;
  }
  const size_t off = (size_t)i * F + f;
  if (RED == PHC_RED_SOFTMAX) {
    Vec<VEC> lse;
#pragma unroll
    for (int q = 0; q < VEC; ++q) lse.v[q] = deg > 0 ? acc[q].a + logf(acc[q].s + 1e-12f) : 0.f;
    lse.store(aux_f + off);
    o.store(aux_f + (size_t)N * F + off);  // aggregated value without the self loop
  }
  if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) {
    IVec<VEC> a;
#pragma unroll
    for (int q = 0; q < VEC; ++q) a.v[q] = acc[q].arg;
    a.store(aux_i + off);
  }
  if (self_loop) {
    Vec<VEC> xi = Vec<VEC>::load(x + off);
#pragma unroll
    for (int q = 0; q < VEC; ++q) o.v[q] += xi.v[q];
  }
  o.store(out + off);
}

// d(ea): one thread per (target row, 4 features), walks the row's in-edges and writes the
// per-edge gradient rows in original edge order.  Also produces block partials of d(beta).
template <int VEC, int RED, bool HAS_ACT>
__global__ void __launch_bounds__(256) aggregate_bwd_edge_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                 const float* __restrict__ ea, const float* __restrict__ aux_f,
                                                                 const int* __restrict__ aux_i, const int* __restrict__ rowptr,
                                                                 const int* __restrict__ col, const int* __restrict__ perm, int N, int F,
                                                                 int act, const float* __restrict__ beta_ptr, float* __restrict__ dea,
                                                                 float* __restrict__ dbeta_part) {
  pdl_begin();
  constexpr bool NEED_PRE = HAS_ACT || RED == PHC_RED_SOFTMAX;
  const int fv = F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float db = 0.f;
  if (t < (long long)N * fv) {
    const int i = (int)(t / fv);
    const int f = (int)(t % fv) * VEC;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const size_t off = (size_t)i * F + f;
    const float beta = (RED == PHC_RED_SOFTMAX) ? __ldg(beta_ptr) : 0.f;
    Vec<VEC> gi = Vec<VEC>::load(g + off);
    Vec<VEC> lse, agg;
    IVec<VEC> arg;
    if (RED == PHC_RED_SOFTMAX) { lse = Vec<VEC>::load(aux_f + off); agg = Vec<VEC>::load(aux_f + (size_t)N * F + off); }
    if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) arg = IVec<VEC>::load(aux_i + off);
    if (RED == PHC_RED_MEAN) {
      float inv = 1.f / (float)max(end - beg, 1);
#pragma unroll
      for (int q = 0; q < VEC; ++q) gi.v[q] *= inv;
    }
    for (int k = beg; k < end; ++k) {
      const int e = __ldg(perm + k);
      Vec<VEC> pre;
      if (NEED_PRE) {
        const int j = __ldg(col + k);
        Vec<VEC> xv = Vec<VEC>::load(x + (size_t)j * F + f);
        Vec<VEC> ev = Vec<VEC>::load_stream(ea + (size_t)e * F + f);
#pragma unroll
        for (int q = 0; q < VEC; ++q) pre.v[q] = xv.v[q] + ev.v[q];
      }
      Vec<VEC> d;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float dm;
        if (RED == PHC_RED_SUM || RED == PHC_RED_MEAN) dm = gi.v[q];
        else if (RED == PHC_RED_MAX || RED == PHC_RED_MIN) dm = (arg.v[q] == e) ? gi.v[q] : 0.f;
        else {
          float m = HAS_ACT ? act_fwd_rt(act, pre.v[q]) : pre.v[q];
          float w = expf(beta * m - lse.v[q]);
          float c = m - agg.v[q];
          dm = gi.v[q] * w * (1.f + beta * c);
          db += gi.v[q] * w * m * c;
        }
        d.v[q] = HAS_ACT ? dm * act_bwd_rt(act, pre.v[q]) : dm;
      }
      d.store_stream(dea + (size_t)e * F + f);
    }
  }
  if (RED == PHC_RED_SOFTMAX) {
    // deterministic block reduction (fixed tree), one partial per block
    __shared__ float red[256];
    red[threadIdx.x] = db;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) dbeta_part[blockIdx.x] = red[0];
  }
}

// d(x): one thread per (source row, 4 features) over the transposed (by-source) structure.
//  SIMPLE (sum/mean without message activation): gathers g[dst] rows (L2-resident) — dea is not read.
//  otherwise: sums the already computed dea rows of the node's out-edges.
template <int VEC, bool SIMPLE, bool MEAN>
__global__ void __launch_bounds__(256) aggregate_bwd_node_kernel(const float* __restrict__ g, const float* __restrict__ dea,
                                                                 const int* __restrict__ rowptr, const int* __restrict__ rowptr_t,
                                                                 const int* __restrict__ col_t, const int* __restrict__ perm_t, int N,
                                                                 int F, int self_loop, float* __restrict__ dx) {
  pdl_begin();
  const int fv = F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int j = (int)(t / fv);
  const int f = (int)(t % fv) * VEC;
  const size_t off = (size_t)j * F + f;
  Vec<VEC> acc;
  if (self_loop) acc = Vec<VEC>::load(g + off);
  else {
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc.v[q] = 0.f;
  }
  const int beg = rowptr_t[j], end = rowptr_t[j + 1];
  int k = beg;
  for (; k + 4 <= end; k += 4) {
    Vec<VEC> v[4];
    float sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (SIMPLE) {
        const int i = __ldg(col_t + k + u);
        v[u] = Vec<VEC>::load(g + (size_t)i * F + f);
        sc[u] = MEAN ? 1.f / (float)max(__ldg(rowptr + i + 1) - __ldg(rowptr + i), 1) : 1.f;
      } else {
        const int e = __ldg(perm_t + k + u);
        v[u] = Vec<VEC>::load_stream(dea + (size_t)e * F + f);
        sc[u] = 1.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc.v[q] += MEAN ? v[u].v[q] * sc[u] : v[u].v[q];
    }
  }
  for (; k < end; ++k) {
    Vec<VEC> v;
    float sc = 1.f;
    if (SIMPLE) {
      const int i = __ldg(col_t + k);
      v = Vec<VEC>::load(g + (size_t)i * F + f);
      if (MEAN) sc = 1.f / (float)max(__ldg(rowptr + i + 1) - __ldg(rowptr + i), 1);
    } else {
      const int e = __ldg(perm_t + k);
      v = Vec<VEC>::load_stream(dea + (size_t)e * F + f);
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc.v[q] += MEAN ? v.v[q] * sc : v.v[q];
  }
  acc.store(dx + off);
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
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

template <int VEC, int RED>
int launch_fwd(const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int N, int F, int act,
               const float* beta, int self_loop, float* out, float* aux_f, int* aux_i, cudaStream_t st) {
  long long threads = (long long)N * (F / VEC);
  int grid = phc_div_up(threads, 256);
  if (act == PHC_ACT_IDENTITY)
    phc_launch(aggregate_fwd_kernel<VEC, RED, false>, dim3(grid), dim3(256), 0, st, x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i);
  else
    phc_launch(aggregate_fwd_kernel<VEC, RED, true>, dim3(grid), dim3(256), 0, st, x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i);
  return phc_check_launch("phc_aggregate_fwd");
}

template <int VEC, int RED>
int launch_bwd_edge(const float* g, const float* x, const float* ea, const float* aux_f, const int* aux_i, const int* rowptr,
                    const int* col, const int* perm, int N, int F, int act, const float* beta, float* dea, float* part, int grid,
                    cudaStream_t st) {
  if (act == PHC_ACT_IDENTITY)
    phc_launch(aggregate_bwd_edge_kernel<VEC, RED, false>, dim3(grid), dim3(256), 0, st, g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part);
  else
    phc_launch(aggregate_bwd_edge_kernel<VEC, RED, true>, dim3(grid), dim3(256), 0, st, g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part);
  return phc_check_launch("phc_aggregate_bwd(edge)");
}

template <int VEC>
int dispatch_fwd(int reduce, const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int N, int F, int act,
                 const float* beta, int self_loop, float* out, float* aux_f, int* aux_i, cudaStream_t st) {
  switch (reduce) {
    case PHC_RED_SUM: return launch_fwd<VEC, PHC_RED_SUM>(x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i, st);
    case PHC_RED_MEAN: return launch_fwd<VEC, PHC_RED_MEAN>(x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i, st);
    case PHC_RED_MAX: return launch_fwd<VEC, PHC_RED_MAX>(x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i, st);
    case PHC_RED_MIN: return launch_fwd<VEC, PHC_RED_MIN>(x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i, st);
    default: return launch_fwd<VEC, PHC_RED_SOFTMAX>(x, ea, rowptr, col, perm, N, F, act, beta, self_loop, out, aux_f, aux_i, st);
  }
}

template <int VEC>
int dispatch_bwd_edge(int reduce, const float* g, const float* x, const float* ea, const float* aux_f, const int* aux_i,
                      const int* rowptr, const int* col, const int* perm, int N, int F, int act, const float* beta, float* dea,
                      float* part, int grid, cudaStream_t st) {
  switch (reduce) {
    case PHC_RED_SUM: return launch_bwd_edge<VEC, PHC_RED_SUM>(g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part, grid, st);
    case PHC_RED_MEAN: return launch_bwd_edge<VEC, PHC_RED_MEAN>(g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part, grid, st);
    case PHC_RED_MAX: return launch_bwd_edge<VEC, PHC_RED_MAX>(g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part, grid, st);
    case PHC_RED_MIN: return launch_bwd_edge<VEC, PHC_RED_MIN>(g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part, grid, st);
    default: return launch_bwd_edge<VEC, PHC_RED_SOFTMAX>(g, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, act, beta, dea, part, grid, st);
  }
}

template <int VEC>
int dispatch_bwd_node(bool simple, bool mean, const float* g, const float* dea, const int* rowptr, const int* rowptr_t, const int* col_t,
                      const int* perm_t, int N, int F, int self_loop, float* dx, cudaStream_t st) {
  int grid = phc_div_up((long long)N * (F / VEC), 256);
  if (simple && mean)
    phc_launch(aggregate_bwd_node_kernel<VEC, true, true>, dim3(grid), dim3(256), 0, st, g, dea, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx);
  else if (simple)
    phc_launch(aggregate_bwd_node_kernel<VEC, true, false>, dim3(grid), dim3(256), 0, st, g, dea, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx);
  else
    phc_launch(aggregate_bwd_node_kernel<VEC, false, false>, dim3(grid), dim3(256), 0, st, g, dea, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx);
  return phc_check_launch("phc_aggregate_bwd(node)");
}

bool vec_ok(int F, const void* a, const void* b, const void* c, const void* d) {
  return F % 4 == 0 && phc_aligned16(a) && phc_aligned16(b) && phc_aligned16(c) && phc_aligned16(d);
}

}  // namespace

// input gradient of the sum/mean + identity-message case (no edge features needed) — shared with conv_fused.cu
int phc_aggregate_bwd_node_simple(bool mean, const float* g, const int* rowptr, const int* rowptr_t, const int* col_t, const int* perm_t,
                                  int N, int F, int self_loop, float* dx, cudaStream_t stream) {
  if (F % 4 == 0 && phc_aligned16(g) && phc_aligned16(dx))
    return dispatch_bwd_node<4>(true, mean, g, nullptr, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
  return dispatch_bwd_node<1>(true, mean, g, nullptr, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
}

// input gradient as the sum of already computed per-edge gradient rows over each node's out-edges — shared with pna.cu
int phc_aggregate_bwd_node_from_edges(const float* dea, const int* rowptr_t, const int* col_t, const int* perm_t, int N, int F, float* dx,
                                      cudaStream_t stream) {
  if (F % 4 == 0 && phc_aligned16(dea) && phc_aligned16(dx))
    return dispatch_bwd_node<4>(false, false, nullptr, dea, nullptr, rowptr_t, col_t, perm_t, N, F, 0, dx, stream);
  return dispatch_bwd_node<1>(false, false, nullptr, dea, nullptr, rowptr_t, col_t, perm_t, N, F, 0, dx, stream);
}

extern "C" {

int phc_aggregate_fwd(const float* x, const float* ea, const int* rowptr, const int* col, const int* perm, int num_nodes, int width,
                      int reduce, int msg_act, const float* beta, int self_loop, float* out, float* aux_f, int* aux_i,
                      cudaStream_t stream) {
  PHC_REQUIRE(reduce >= PHC_RED_SUM && reduce <= PHC_RED_SOFTMAX, "phc_aggregate_fwd: bad reduce %d", reduce);
  PHC_REQUIRE(msg_act >= PHC_ACT_IDENTITY && msg_act <= PHC_ACT_SWISH, "phc_aggregate_fwd: bad msg_act %d", msg_act);
  PHC_REQUIRE(width > 0 && num_nodes >= 0, "phc_aggregate_fwd: bad shape");
  PHC_REQUIRE(reduce != PHC_RED_SOFTMAX || (beta && aux_f), "phc_aggregate_fwd: softmax needs beta and aux_f");
  PHC_REQUIRE((reduce != PHC_RED_MAX && reduce != PHC_RED_MIN) || aux_i, "phc_aggregate_fwd: max/min need aux_i");
  if (num_nodes == 0) return PHC_OK;
  bool v4 = vec_ok(width, x, ea, out, aux_f) && phc_aligned16(aux_i);
  if (v4) return dispatch_fwd<4>(reduce, x, ea, rowptr, col, perm, num_nodes, width, msg_act, beta, self_loop, out, aux_f, aux_i, stream);
  return dispatch_fwd<1>(reduce, x, ea, rowptr, col, perm, num_nodes, width, msg_act, beta, self_loop, out, aux_f, aux_i, stream);
}

size_t phc_aggregate_bwd_workspace_bytes(int num_nodes, int width) {
  long long threads = (long long)num_nodes * width;  // upper bound on thread count (VEC=1)
  return sizeof(float) * (size_t)(phc_div_up(threads, 256) + 1);
}

int phc_aggregate_bwd(const float* gout, const float* x, const float* ea, const float* aux_f, const int* aux_i, const int* rowptr,
                      const int* col, const int* perm, const int* rowptr_t, const int* col_t, const int* perm_t, int num_nodes,
                      int width, int reduce, int msg_act, const float* beta, int self_loop, float* dx, float* dea, float* dbeta,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(reduce >= PHC_RED_SUM && reduce <= PHC_RED_SOFTMAX, "phc_aggregate_bwd: bad reduce %d", reduce);
  PHC_REQUIRE(msg_act >= PHC_ACT_IDENTITY && msg_act <= PHC_ACT_SWISH, "phc_aggregate_bwd: bad msg_act %d", msg_act);
  PHC_REQUIRE(workspace_bytes >= phc_aggregate_bwd_workspace_bytes(num_nodes, width), "phc_aggregate_bwd: workspace too small");
  if (num_nodes == 0) return PHC_OK;
  const int N = num_nodes, F = width;
  bool v4 = vec_ok(F, x, ea, gout, dx) && phc_aligned16(dea) && phc_aligned16(aux_f) && phc_aligned16(aux_i);
  float* part = reinterpret_cast<float*>(workspace);
  int rc;
  int grid = phc_div_up((long long)N * (v4 ? F / 4 : F), 256);
  if (v4) rc = dispatch_bwd_edge<4>(reduce, gout, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, msg_act, beta, dea, part, grid, stream);
  else rc = dispatch_bwd_edge<1>(reduce, gout, x, ea, aux_f, aux_i, rowptr, col, perm, N, F, msg_act, beta, dea, part, grid, stream);
  if (rc) return rc;
  if (reduce == PHC_RED_SOFTMAX && dbeta) {
    phc_launch(reduce_partials_kernel, dim3(1), dim3(256), 0, stream, part, grid, dbeta);
    rc = phc_check_launch("phc_aggregate_bwd(dbeta)");
    if (rc) return rc;
  }
  bool simple = (reduce == PHC_RED_SUM || reduce == PHC_RED_MEAN) && msg_act == PHC_ACT_IDENTITY;
  bool mean = reduce == PHC_RED_MEAN;
  if (v4) return dispatch_bwd_node<4>(simple, mean, gout, dea, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
  return dispatch_bwd_node<1>(simple, mean, gout, dea, rowptr, rowptr_t, col_t, perm_t, N, F, self_loop, dx, stream);
}

}  // extern "C"
