// Graph structure kernels: CSR (by target) and CSC (by source) construction from a PyG-style
// edge_index, and graph segment offsets from the batch vector.
//
// Replaces, on the hot path, the implicit "structure" of torch_scatter's atomic scatter
// (reference phc/hypercomplex/undirectional/messagepassing.py:136,221,306 -> PyG propagate ->
// torch_scatter.scatter) and torch_geometric.nn.global_add_pool (phc/hypercomplex/pooling.py:18).
// Result is bit-exact with a stable argsort of the keys (oracle: oracle/phc_oracle.py csr_by_target):
//   perm   = stable_argsort(key)           (edge ids, ascending inside every row)
//   col    = other_end[perm]
//   rowptr = exclusive_cumsum(bincount(key))
// Algorithm: integer histogram (atomics on ints are order-independent) -> tiled single-block scan ->
// ticket scatter (arbitrary order inside a row) -> per-row sort by edge id (restores the unique
// stable order).  No floating point is involved, so the result is deterministic.
#include "common.cuh"
#include <limits.h>

namespace {

__global__ void hist_kernel(const long long* __restrict__ ei, int E, int N, int* __restrict__ cnt_dst,
                            int* __restrict__ cnt_src, int* __restrict__ status) {
  pdl_begin();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  long long s = ei[e], d = ei[(size_t)E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) { atomicOr(status, 1); return; }
  atomicAdd(&cnt_dst[d], 1);
  atomicAdd(&cnt_src[s], 1);
}

// One block per key array (blockIdx.x = 0: target keys, 1: source keys). Exclusive scan of cnt[0..N)
// into rowptr[0..N] and a copy into cursor[0..N) for the ticket scatter.  The block walks the array in tiles of 4096 counts with
// a running carry: coalesced loads (4 consecutive counts per thread), warp-shuffle scans, one 32-entry scan of the warp totals
// per tile (the first version gave each thread one contiguous chunk — uncoalesced — and ran a 10-round Hillis-Steele scan over
// 1024 partials: 24 us at N = 15.6k, ppa; now ~5 us).
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ cnt_all, int N, int* __restrict__ rowptr_dst,
                                                    int* __restrict__ rowptr_src, int* __restrict__ cursor_all) {
  pdl_begin();
  const int* cnt = cnt_all + (size_t)blockIdx.x * N;
  int* rowptr = blockIdx.x == 0 ? rowptr_dst : rowptr_src;
  int* cursor = cursor_all + (size_t)blockIdx.x * N;
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 4096) {
    const int i0 = base + t * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i0 + j < N) ? cnt[i0 + j] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int inc = mine;                                            // inclusive scan of the thread totals inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += n;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {                                              // scan of the 32 warp totals
      int x = warp_tot[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += n;
      }
      warp_tot[lane] = x;                                      // inclusive
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + (w > 0 ? warp_tot[w - 1] : 0) + inc - mine;     // exclusive prefix of this thread's first count
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < N) {
        rowptr[i0 + j] = run;
        cursor[i0 + j] = run;
      }
      run += v[j];
    }
    __syncthreads();                                           // everyone has read carry_s and warp_tot
    if (t == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (t == 0) rowptr[N] = carry_s;
}

__global__ void ticket_kernel(const long long* __restrict__ ei, int E, int N, int* __restrict__ cursor_all,
                              int* __restrict__ slot_dst, int* __restrict__ slot_src) {
  pdl_begin();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  long long s = ei[e], d = ei[(size_t)E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) return;
  slot_dst[atomicAdd(&cursor_all[d], 1)] = e;
  slot_src[atomicAdd(&cursor_all[(size_t)N + s], 1)] = e;
}

// Sort every row's edge ids ascending, then emit perm and col. One warp per row.
//  deg <= 32 : bitonic network over the lanes (shuffles)
//  deg  > 32 : rank by counting (each id is unique, so rank = #smaller), O(deg^2/32) per warp —
//              only hub rows take this path.
__global__ void __launch_bounds__(256) rowsort_kernel(const long long* __restrict__ ei, int E, int N,
                                                      const int* __restrict__ rowptr_dst, const int* __restrict__ rowptr_src,
                                                      const int* __restrict__ slot_dst, const int* __restrict__ slot_src,
                                                      int* __restrict__ perm_dst, int* __restrict__ col_dst,
                                                      int* __restrict__ perm_src, int* __restrict__ col_src) {
  pdl_begin();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= 2 * N) return;
  bool by_src = warp >= N;
  int row = by_src ? warp - N : warp;
  const int* rowptr = by_src ? rowptr_src : rowptr_dst;
  const int* slot = by_src ? slot_src : slot_dst;
  int* perm = by_src ? perm_src : perm_dst;
  int* col = by_src ? col_src : col_dst;
  // other end of the edge: rows keyed by target store the source and vice versa
  const long long* other = by_src ? ei + (size_t)E : ei;
  int beg = rowptr[row], end = rowptr[row + 1];
  int deg = end - beg;
  if (deg <= 0) return;
  if (deg <= 32) {
    int v = lane < deg ? slot[beg + lane] : INT_MAX;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        int o = __shfl_xor_sync(0xffffffffu, v, j);
        bool up = ((lane & k) == 0);
        bool lower = ((lane & j) == 0);
        v = (lower == up) ? min(v, o) : max(v, o);
      }
    }
    if (lane < deg) {
      perm[beg + lane] = v;
      col[beg + lane] = (int)other[v];
    }
  } else {
    for (int i = lane; i < deg; i += 32) {
      int v = slot[beg + i];
      int rank = 0;
      for (int j = 0; j < deg; ++j) rank += (slot[beg + j] < v) ? 1 : 0;
      perm[beg + rank] = v;
      col[beg + rank] = (int)other[v];
    }
  }
}

// graph_ptr from an ascending batch vector: ptr[b] = first node of graph b (empty graphs allowed).
__global__ void segptr_kernel(const long long* __restrict__ batch, int N, int B, int* __restrict__ ptr, int* __restrict__ status) {
  pdl_begin();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > N) return;
  if (N == 0) { if (i == 0) for (int b = 0; b <= B; ++b) ptr[b] = 0; return; }
  if (i == N) {
    long long last = batch[N - 1];
    if (last < 0 || last >= B) { atomicOr(status, 1); return; }
    for (long long b = last + 1; b <= B; ++b) ptr[b] = N;
    return;
  }
  long long cur = batch[i];
  if (cur < 0 || cur >= B) { atomicOr(status, 1); return; }
  long long prev = i == 0 ? -1 : batch[i - 1];
  if (prev > cur) { atomicOr(status, 2); return; }  // not ascending
  for (long long b = prev + 1; b <= cur; ++b) ptr[b] = i;
}

__global__ void narrow_i64_kernel(const long long* __restrict__ in, int n, int* __restrict__ out) {
  pdl_begin();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

}  // namespace

extern "C" {

size_t phc_csr_workspace_bytes(int num_nodes, int num_edges) {
  // cnt[2N] + cursor[2N] + slot_dst[E] + slot_src[E] + status[1]
  return sizeof(int) * (4 * (size_t)num_nodes + 2 * (size_t)num_edges + 4);
}

int phc_csr_build(const long long* edge_index, int num_edges, int num_nodes, int* rowptr, int* col, int* perm, int* rowptr_t,
                  int* col_t, int* perm_t, void* workspace, size_t workspace_bytes, int* status, cudaStream_t stream) {
  PHC_REQUIRE(num_edges >= 0 && num_nodes >= 0, "phc_csr_build: negative size");
  PHC_REQUIRE(workspace_bytes >= phc_csr_workspace_bytes(num_nodes, num_edges), "phc_csr_build: workspace too small");
  const int N = num_nodes, E = num_edges;
  int* cnt = reinterpret_cast<int*>(workspace);
  int* cursor = cnt + 2 * (size_t)N;
  int* slot_dst = cursor + 2 * (size_t)N;
  int* slot_src = slot_dst + E;
  cudaMemsetAsync(cnt, 0, sizeof(int) * 2 * (size_t)N, stream);
  cudaMemsetAsync(status, 0, sizeof(int), stream);
  if (E > 0) phc_launch(hist_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, N, cnt, cnt + N, status);
  phc_launch(scan_kernel, dim3(2), dim3(1024), 0, stream, cnt, N, rowptr, rowptr_t, cursor);
  if (E > 0 && N > 0) {
    phc_launch(ticket_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, N, cursor, slot_dst, slot_src);
    phc_launch(rowsort_kernel, dim3(phc_div_up(2LL * N * 32, 256)), dim3(256), 0, stream, edge_index, E, N, rowptr, rowptr_t, slot_dst, slot_src, perm, col,
                                                                    perm_t, col_t);
  }
  return phc_check_launch("phc_csr_build");
}

int phc_segment_ptr_build(const long long* batch, int num_nodes, int num_graphs, int* graph_ptr, int* status, cudaStream_t stream) {
  PHC_REQUIRE(num_nodes >= 0 && num_graphs >= 0, "phc_segment_ptr_build: negative size");
  cudaMemsetAsync(status, 0, sizeof(int), stream);
  phc_launch(segptr_kernel, dim3(phc_div_up(num_nodes + 1, 256)), dim3(256), 0, stream, batch, num_nodes, num_graphs, graph_ptr, status);
  return phc_check_launch("phc_segment_ptr_build");
}

int phc_narrow_int64(const long long* in, int n, int* out, cudaStream_t stream) {
  if (n > 0) phc_launch(narrow_i64_kernel, dim3(phc_div_up(n, 256)), dim3(256), 0, stream, in, n, out);
  return phc_check_launch("phc_narrow_int64");
}

}  // extern "C"
