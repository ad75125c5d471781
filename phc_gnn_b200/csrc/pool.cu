// Graph-level pooling as a segmented reduction over the (ascending) batch vector.
//
//   plain : out[b,:]      = sum_{i in graph b} x[i,:]
//   gated : out[b,c,f']   = sum_{i in graph b} sigmoid(z[i,f']) * x[i,c,f']     (gate shared by the n components)
//
// Replaces torch_geometric.nn.global_add_pool (atomic scatter_add; reference
// phc/hypercomplex/pooling.py:10-25) and the sigmoid / broadcast-multiply / view chain of
// PHMSoftAttentionPooling.forward (phc/hypercomplex/pooling.py:57-66).  Deterministic: each output
// element is produced by a fixed-shape tree over node slices, no atomics.
// Roofline: HBM.  Algorithmic bytes fwd = 4F(N+B) + 4(B+1) (+4NF/n for the gate logits).
#include "common.cuh"

namespace {

constexpr int POOL_SLICES = 8;   // node slices per block
constexpr int POOL_LANES = 32;   // feature lanes per block (x VEC features each)

template <int VEC, bool GATED>
__global__ void __launch_bounds__(POOL_SLICES * POOL_LANES) pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                                           const int* __restrict__ gptr, int F, int Fc,
                                                                           float* __restrict__ out) {
  pdl_begin();
  const int b = blockIdx.x;
  const int lane = threadIdx.x % POOL_LANES;
  const int slice = threadIdx.x / POOL_LANES;
  const int f = (blockIdx.y * POOL_LANES + lane) * VEC;
  const bool active = f < F;
  const int beg = gptr[b], end = gptr[b + 1];
  Vec<VEC> acc;
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc.v[q] = 0.f;
  if (active) {
    // four rows in flight per thread (a block has only 8 row slices and a graph a few hundred rows: with one load per trip the loop
    // ran at one memory latency per row — 43 us for the 31 MB of a ppa batch); same summation order as the one-row loop
    int zoff[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) zoff[q] = (f + q) % Fc;
    int i = beg + slice;
    for (; i + 3 * POOL_SLICES < end; i += 4 * POOL_SLICES) {
      Vec<VEC> v[4];
      float zz[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = Vec<VEC>::load(x + (size_t)(i + u * POOL_SLICES) * F + f);
        if (GATED) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) zz[u][q] = __ldg(z + (size_t)(i + u * POOL_SLICES) * Fc + zoff[q]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          float gq = 1.f;
          if (GATED) gq = 1.f / (1.f + expf(-zz[u][q]));
          acc.v[q] += gq * v[u].v[q];
        }
      }
    }
    for (; i < end; i += POOL_SLICES) {
      Vec<VEC> v = Vec<VEC>::load(x + (size_t)i * F + f);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float gq = 1.f;
        if (GATED) gq = 1.f / (1.f + expf(-__ldg(z + (size_t)i * Fc + zoff[q])));
        acc.v[q] += gq * v.v[q];
      }
    }
  }
  __shared__ float red[POOL_SLICES][POOL_LANES * VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) red[slice][lane * VEC + q] = acc.v[q];
  __syncthreads();
  for (int s = POOL_SLICES / 2; s > 0; s >>= 1) {
    if (slice < s) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) red[slice][lane * VEC + q] += red[slice + s][lane * VEC + q];
    }
    __syncthreads();
  }
  if (slice == 0 && active) {
    Vec<VEC> o;
#pragma unroll
    for (int q = 0; q < VEC; ++q) o.v[q] = red[0][lane * VEC + q];
    o.store(out + (size_t)b * F + f);
  }
}

// backward: dx[i,c,f'] = sig(z[i,f']) * g[batch[i],c,f'] ; dz[i,f'] = sig'(z) * sum_c x[i,c,f'] g[b,c,f']
// one thread per (node, f') for the gated form (loops the n components), per (node, VEC feats) for plain.
__global__ void __launch_bounds__(256) pool_bwd_gated_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                             const float* __restrict__ z, const long long* __restrict__ batch, int N,
                                                             int F, int Fc, float* __restrict__ dx, float* __restrict__ dz) {
  pdl_begin();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * Fc) return;
  const int i = (int)(t / Fc), fp = (int)(t % Fc);
  const int b = (int)batch[i];
  const float s = 1.f / (1.f + expf(-z[(size_t)i * Fc + fp]));
  float dot = 0.f;
  const int n = F / Fc;
  for (int c = 0; c < n; ++c) {
    const float gv = g[(size_t)b * F + c * Fc + fp];
    dot += x[(size_t)i * F + c * Fc + fp] * gv;
    dx[(size_t)i * F + c * Fc + fp] = s * gv;
  }
  dz[(size_t)i * Fc + fp] = dot * s * (1.f - s);
}

template <int VEC>
__global__ void __launch_bounds__(256) pool_bwd_plain_kernel(const float* __restrict__ g, const long long* __restrict__ batch, int N, int F,
                                                             float* __restrict__ dx) {
  pdl_begin();
  const int fv = F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * fv) return;
  const int i = (int)(t / fv), f = (int)(t % fv) * VEC;
  const int b = (int)batch[i];
  Vec<VEC>::load(g + (size_t)b * F + f).store(dx + (size_t)i * F + f);
}

}  // namespace

extern "C" {

int phc_segment_pool_fwd(const float* x, const float* gate_logits, const int* graph_ptr, int num_graphs, int width, int phm_dim,
                         float* out, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_segment_pool_fwd: width %d not divisible by phm_dim %d", width, phm_dim);
  if (num_graphs == 0) return PHC_OK;
  const int Fc = width / phm_dim;
  const bool v4 = width % 4 == 0 && phc_aligned16(x) && phc_aligned16(out);
  dim3 grid(num_graphs, phc_div_up(width, POOL_LANES * (v4 ? 4 : 1)));
  const int threads = POOL_SLICES * POOL_LANES;
  if (gate_logits) {
    if (v4) phc_launch(pool_fwd_kernel<4, true>, dim3(grid), dim3(threads), 0, stream, x, gate_logits, graph_ptr, width, Fc, out);
    else phc_launch(pool_fwd_kernel<1, true>, dim3(grid), dim3(threads), 0, stream, x, gate_logits, graph_ptr, width, Fc, out);
  } else {
    if (v4) phc_launch(pool_fwd_kernel<4, false>, dim3(grid), dim3(threads), 0, stream, x, gate_logits, graph_ptr, width, Fc, out);
    else phc_launch(pool_fwd_kernel<1, false>, dim3(grid), dim3(threads), 0, stream, x, gate_logits, graph_ptr, width, Fc, out);
  }
  return phc_check_launch("phc_segment_pool_fwd");
}

int phc_segment_pool_bwd(const float* gout, const float* x, const float* gate_logits, const long long* batch, int num_nodes, int width,
                         int phm_dim, float* dx, float* dgate_logits, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_segment_pool_bwd: width %d not divisible by phm_dim %d", width, phm_dim);
  if (num_nodes == 0) return PHC_OK;
  const int Fc = width / phm_dim;
  if (gate_logits) {
    phc_launch(pool_bwd_gated_kernel, dim3(phc_div_up((long long)num_nodes * Fc, 256)), dim3(256), 0, stream, gout, x, gate_logits, batch, num_nodes, width, Fc,
                                                                                         dx, dgate_logits);
  } else {
    const bool v4 = width % 4 == 0 && phc_aligned16(gout) && phc_aligned16(dx);
    if (v4) phc_launch(pool_bwd_plain_kernel<4>, dim3(phc_div_up((long long)num_nodes * (width / 4), 256)), dim3(256), 0, stream, gout, batch, num_nodes, width, dx);
    else phc_launch(pool_bwd_plain_kernel<1>, dim3(phc_div_up((long long)num_nodes * width, 256)), dim3(256), 0, stream, gout, batch, num_nodes, width, dx);
  }
  return phc_check_launch("phc_segment_pool_bwd");
}

}  // extern "C"
