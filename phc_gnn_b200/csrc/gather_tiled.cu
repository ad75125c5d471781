// Graph-tiled neighbour gather: the sum / mean aggregation with identity message (forward, and its input gradient) with the
// gathered rows staged in SHARED memory, one CTA per (graph of the mini-batch, 64-feature slice).
//
// Replaces, where the batch's graph boundaries are known (SegmentStructure.graph_ptr, built from the `batch` vector the reference's
// pooling uses, models.py:219-232), conv_fwd_sums_kernel (conv_fused.cu) and aggregate_bwd_node_kernel<SIMPLE> (aggregate.cu).  Those
// kernels gather every neighbour row from L2: 290k edges x 2 000 B = 580 MB per layer at the ppa shape, 52.7 us = 11 TB/s — the L2
// request rate, not HBM, bounds them (ncu: profiles/r01_conv_fwd_sums_ppa_ncu_full.txt).  A PyG mini-batch is block diagonal: the
// neighbours of a node lie in its own graph, whose nodes are contiguous rows.  A CTA therefore copies the graph's rows of its feature
// slice into shared memory once (coalesced, every row read once from L2 / HBM: 31 MB per layer) and serves the ~19 gathers per row from
// there (a half warp reads the 256 contiguous bytes of one staged row: conflict free).  Shared-memory bandwidth (148 x 128 B/clk) is
// 3.5x the measured L2 gather rate.
//   * Arithmetic and summation order are exactly those of the kernels replaced (edges in CSR order, one accumulator per feature), so
//     results are bit-identical to them: determinism and every parity test carry over.
//   * Nothing is assumed about the input: a neighbour outside the CTA's graph (not a PyG batch) or a graph larger than the staging
//     buffer is read from global memory as before.
#include "common.cuh"

namespace {

constexpr int GT_SLICE = 64;        // floats of a feature slice: 256-byte rows in shared memory
constexpr int GT_THREADS = 256;
constexpr int GT_MAX_ROWS = 384;    // 96 KiB of staged rows: two CTAs per SM

// one (graph, slice) work item
struct GtItem {
  int n0, cnt, f0, fq;              // first node, node count, first feature, float4 per row
  bool staged;
};
__device__ __forceinline__ GtItem gt_item(int w, int nsl, const int* __restrict__ graph_ptr, int F, int cap_rows) {
  GtItem it;
  const int b = w / nsl, s = w - b * nsl;
  it.n0 = __ldg(graph_ptr + b);
  it.cnt = __ldg(graph_ptr + b + 1) - it.n0;
  it.f0 = s * GT_SLICE;
  it.fq = min(GT_SLICE, F - it.f0) >> 2;
  it.staged = it.cnt <= cap_rows;
  return it;
}
__device__ __forceinline__ void gt_stage(float* __restrict__ tile, const float* __restrict__ src, const GtItem& it, int F) {
  if (it.staged) {
    for (int t = threadIdx.x; t < it.cnt * it.fq; t += GT_THREADS) {
      const int r = t / it.fq, q = t - r * it.fq;
      *reinterpret_cast<float4*>(tile + r * GT_SLICE + q * 4) =
          *reinterpret_cast<const float4*>(src + (size_t)(it.n0 + r) * F + it.f0 + q * 4);
    }
  }
  __syncthreads();
}
// row j of the gathered matrix, this thread's four features: from the staged tile when j is one of the graph's nodes
__device__ __forceinline__ float4 gt_row(const float* __restrict__ tile, const float* __restrict__ src, const GtItem& it, int j, int q,
                                         int F) {
  const unsigned lj = (unsigned)(j - it.n0);
  if (it.staged && lj < (unsigned)it.cnt) return *reinterpret_cast<const float4*>(tile + lj * GT_SLICE + q * 4);
  return *reinterpret_cast<const float4*>(src + (size_t)j * F + it.f0 + q * 4);
}

// forward: out[i] = [x[i] +] (sum | mean)_{j in N(i)} x[j] + sum_r S[i,r] tab[r]     (conv_fwd_sums_kernel, conv_fused.cu)
template <bool MEAN>
__global__ void __launch_bounds__(GT_THREADS) conv_fwd_sums_tiled_kernel(const float* __restrict__ x, const float* __restrict__ S,
                                                                         const float* __restrict__ tab, const int* __restrict__ rowptr,
                                                                         const int* __restrict__ col, const int* __restrict__ graph_ptr,
                                                                         int num_graphs, int F, int R, int self_loop, int cap_rows,
                                                                         float* __restrict__ out) {
  pdl_begin();
  extern __shared__ __align__(16) float tile[];
  const int nsl = (F + GT_SLICE - 1) / GT_SLICE;
  for (int w = blockIdx.x; w < num_graphs * nsl; w += gridDim.x) {
    const GtItem it = gt_item(w, nsl, graph_ptr, F, cap_rows);
    gt_stage(tile, x, it, F);
    for (int t = threadIdx.x; t < it.cnt * it.fq; t += GT_THREADS) {
      const int r = t / it.fq, q = t - r * it.fq;
      const int i = it.n0 + r, f = it.f0 + q * 4;
      const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = beg;
      for (; k + 4 <= end; k += 4) {                 // four rows in flight
        int j[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) j[u] = __ldg(col + k + u);
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = gt_row(tile, x, it, j[u], q, F);
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
      }
      for (; k < end; ++k) {
        const float4 v = gt_row(tile, x, it, __ldg(col + k), q, F);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      }
      if (MEAN) {
        const float sc = 1.f / (float)max(end - beg, 1);
        a0 *= sc; a1 *= sc; a2 *= sc; a3 *= sc;
      }
      const float* si = S + (size_t)i * R;
      for (int rr = 0; rr < R; ++rr) {
        const float sv = __ldg(si + rr);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(tab + (size_t)rr * F + f));
        a0 += sv * wv.x; a1 += sv * wv.y; a2 += sv * wv.z; a3 += sv * wv.w;
      }
      if (self_loop) {
        const float4 xi = gt_row(tile, x, it, i, q, F);
        a0 += xi.x; a1 += xi.y; a2 += xi.z; a3 += xi.w;
      }
      *reinterpret_cast<float4*>(out + (size_t)i * F + f) = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();                                 // the tile is overwritten by the next item
  }
}

// input gradient: dx[j] = [g[j] +] sum_{i : j -> i} g[i] (/ deg(i) for the mean)     (aggregate_bwd_node_kernel<4, SIMPLE>, aggregate.cu)
template <bool MEAN>
__global__ void __launch_bounds__(GT_THREADS) aggregate_bwd_node_tiled_kernel(const float* __restrict__ g, const int* __restrict__ rowptr,
                                                                              const int* __restrict__ rowptr_t,
                                                                              const int* __restrict__ col_t,
                                                                              const int* __restrict__ graph_ptr, int num_graphs, int F,
                                                                              int self_loop, int cap_rows, float* __restrict__ dx) {
  pdl_begin();
  extern __shared__ __align__(16) float tile[];
  const int nsl = (F + GT_SLICE - 1) / GT_SLICE;
  for (int w = blockIdx.x; w < num_graphs * nsl; w += gridDim.x) {
    const GtItem it = gt_item(w, nsl, graph_ptr, F, cap_rows);
    gt_stage(tile, g, it, F);
    for (int t = threadIdx.x; t < it.cnt * it.fq; t += GT_THREADS) {
      const int r = t / it.fq, q = t - r * it.fq;
      const int j = it.n0 + r, f = it.f0 + q * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (self_loop) acc = gt_row(tile, g, it, j, q, F);
      const int beg = __ldg(rowptr_t + j), end = __ldg(rowptr_t + j + 1);
      int k = beg;
      for (; k + 4 <= end; k += 4) {
        float4 v[4];
        float sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = __ldg(col_t + k + u);
          v[u] = gt_row(tile, g, it, i, q, F);
          sc[u] = MEAN ? 1.f / (float)max(__ldg(rowptr + i + 1) - __ldg(rowptr + i), 1) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc.x += MEAN ? v[u].x * sc[u] : v[u].x;
          acc.y += MEAN ? v[u].y * sc[u] : v[u].y;
          acc.z += MEAN ? v[u].z * sc[u] : v[u].z;
          acc.w += MEAN ? v[u].w * sc[u] : v[u].w;
        }
      }
      for (; k < end; ++k) {
        const int i = __ldg(col_t + k);
        const float4 v = gt_row(tile, g, it, i, q, F);
        const float sc = MEAN ? 1.f / (float)max(__ldg(rowptr + i + 1) - __ldg(rowptr + i), 1) : 1.f;
        acc.x += MEAN ? v.x * sc : v.x;
        acc.y += MEAN ? v.y * sc : v.y;
        acc.z += MEAN ? v.z * sc : v.z;
        acc.w += MEAN ? v.w * sc : v.w;
      }
      *reinterpret_cast<float4*>(dx + (size_t)j * F + f) = acc;
    }
    __syncthreads();
  }
}

// staged rows per CTA: 2.5x the mean graph size (a graph above it gathers from global memory), at least 64, at most 96 KiB
int gt_cap_rows(int num_nodes, int num_graphs) {
  long long cap = ((long long)num_nodes * 5 / (2 * (long long)num_graphs) + 31) / 32 * 32;
  if (cap < 64) cap = 64;
  if (cap > GT_MAX_ROWS) cap = GT_MAX_ROWS;
  return (int)cap;
}

// the opt-in to more than 48 KiB of dynamic shared memory is per kernel and sticky: raise it to the maximum once
template <auto Kern>
bool gt_ensure_smem() {
  static const bool ok = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_MAX_ROWS * GT_SLICE * 4) == cudaSuccess;
  if (!ok) cudaGetLastError();
  return ok;
}

}  // namespace

// true when the tiled kernels apply: vectorised rows and a graph table
bool phc_gather_tiled_ok(const int* graph_ptr, int num_graphs, int num_nodes, int width, const void* a, const void* b) {
  return graph_ptr != nullptr && num_graphs > 0 && num_nodes > 0 && width % 4 == 0 && phc_aligned16(a) && phc_aligned16(b);
}

int phc_conv_fwd_sums_tiled_launch(bool mean, const float* x, const float* node_sums, const float* tab, const int* rowptr, const int* col,
                                   const int* graph_ptr, int num_graphs, int num_nodes, int width, int table_rows, int self_loop,
                                   float* out, cudaStream_t stream) {
  const int cap = gt_cap_rows(num_nodes, num_graphs);
  const size_t smem = (size_t)cap * GT_SLICE * 4;
  const int items = num_graphs * phc_div_up(width, GT_SLICE);
  if (mean) {
    if (!gt_ensure_smem<conv_fwd_sums_tiled_kernel<true>>()) { phc_set_error("phc_conv_fused_fwd_sums: shared memory"); return PHC_ERR_CUDA; }
    phc_launch(conv_fwd_sums_tiled_kernel<true>, dim3(items), dim3(GT_THREADS), smem, stream, x, node_sums, tab, rowptr, col, graph_ptr,
               num_graphs, width, table_rows, self_loop, cap, out);
  } else {
    if (!gt_ensure_smem<conv_fwd_sums_tiled_kernel<false>>()) { phc_set_error("phc_conv_fused_fwd_sums: shared memory"); return PHC_ERR_CUDA; }
    phc_launch(conv_fwd_sums_tiled_kernel<false>, dim3(items), dim3(GT_THREADS), smem, stream, x, node_sums, tab, rowptr, col, graph_ptr,
               num_graphs, width, table_rows, self_loop, cap, out);
  }
  return phc_check_launch("phc_conv_fused_fwd_sums(tiled)");
}

int phc_aggregate_bwd_node_tiled_launch(bool mean, const float* g, const int* rowptr, const int* rowptr_t, const int* col_t,
                                        const int* graph_ptr, int num_graphs, int num_nodes, int width, int self_loop, float* dx,
                                        cudaStream_t stream) {
  const int cap = gt_cap_rows(num_nodes, num_graphs);
  const size_t smem = (size_t)cap * GT_SLICE * 4;
  const int items = num_graphs * phc_div_up(width, GT_SLICE);
  if (mean) {
    if (!gt_ensure_smem<aggregate_bwd_node_tiled_kernel<true>>()) { phc_set_error("phc_aggregate_bwd(node, tiled): shared memory"); return PHC_ERR_CUDA; }
    phc_launch(aggregate_bwd_node_tiled_kernel<true>, dim3(items), dim3(GT_THREADS), smem, stream, g, rowptr, rowptr_t, col_t, graph_ptr,
               num_graphs, width, self_loop, cap, dx);
  } else {
    if (!gt_ensure_smem<aggregate_bwd_node_tiled_kernel<false>>()) { phc_set_error("phc_aggregate_bwd(node, tiled): shared memory"); return PHC_ERR_CUDA; }
    phc_launch(aggregate_bwd_node_tiled_kernel<false>, dim3(items), dim3(GT_THREADS), smem, stream, g, rowptr, rowptr_t, col_t, graph_ptr,
               num_graphs, width, self_loop, cap, dx);
  }
  return phc_check_launch("phc_aggregate_bwd(node, tiled)");
}
