// Batch preparation on the device: the per-batch `data = transform(data)` of the training loops
// (reference benchmarks/train_hiv.py:171-173, :233-234; transform = torch_geometric.transforms.RemoveIsolatedNodes,
// :457; benchmarks/utils.py:39-49 CustomRemoveIsolatedNodes).  The arithmetic lives in the third-party
// torch_geometric 1.6.1 (utils/isolated.py remove_isolated_nodes + utils/loop.py segregate_self_loops), restated in
// oracle/phc_oracle.py::remove_isolated_nodes:
//   keep[i]   = node i is an endpoint of at least one edge that is not a self loop
//   assoc[i]  = rank of i among the kept nodes (-1 if removed)
//   edges out = the non-loop edges in their original order, relabelled, followed by one self loop per kept node that has
//               any (ascending node id; of several self loops on one node the LAST in edge order survives)
// All integer work: bit-exact with the CPU restatement, no float atomics, the only atomics are byte flags and an
// integer max (order independent).
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(256) prep_mark_kernel(const long long* __restrict__ ei, int E, int N, unsigned char* __restrict__ keep,
                                                        int* __restrict__ loop_last, int* __restrict__ status) {
  pdl_begin();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long s = ei[e], d = ei[(size_t)E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) { atomicOr(status, 1); return; }
  if (s != d) { keep[s] = 1; keep[d] = 1; }
  else atomicMax(loop_last + s, e);
}

// Exclusive scan of flag(i) over n items by ONE block (each thread owns a contiguous slice): pos[i], total -> *count.
// which = 0: flag = keep[i];  1: flag = edge i is not a self loop;  2: flag = keep[i] && loop_last[i] >= 0
__global__ void __launch_bounds__(SCAN_THREADS) prep_scan_kernel(int which, int n, const unsigned char* __restrict__ keep,
                                                                 const int* __restrict__ loop_last, const long long* __restrict__ ei, int E,
                                                                 int* __restrict__ pos, int* __restrict__ count) {
  pdl_begin();
  __shared__ int sums[SCAN_THREADS];
  const int t = threadIdx.x;
  const int per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
  const int b = min(t * per, n), e = min(b + per, n);
  auto flag = [&](int i) -> int {
    if (which == 0) return keep[i] ? 1 : 0;
    if (which == 1) return ei[i] != ei[(size_t)E + i] ? 1 : 0;
    return (keep[i] && loop_last[i] >= 0) ? 1 : 0;
  };
  int c = 0;
  for (int i = b; i < e; ++i) c += flag(i);
  sums[t] = c;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {            // Hillis-Steele inclusive scan of the slice sums
    const int v = t >= o ? sums[t - o] : 0;
    __syncthreads();
    sums[t] += v;
    __syncthreads();
  }
  int run = sums[t] - c;
  for (int i = b; i < e; ++i) {
    pos[i] = run;
    run += flag(i);
  }
  if (t == SCAN_THREADS - 1) *count = sums[t];
}

__global__ void __launch_bounds__(256) prep_assoc_kernel(int N, const unsigned char* __restrict__ keep, const int* __restrict__ node_pos,
                                                         long long* __restrict__ assoc) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) assoc[i] = keep[i] ? (long long)node_pos[i] : -1LL;
}

// out_cap = E: row 0 of the relabelled edge list at out[0..], row 1 at out[out_cap..]
__global__ void __launch_bounds__(256) prep_edges_kernel(const long long* __restrict__ ei, int E, const int* __restrict__ node_pos,
                                                         const int* __restrict__ edge_pos, long long* __restrict__ out,
                                                         long long* __restrict__ order) {
  pdl_begin();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long s = ei[e], d = ei[(size_t)E + e];
  if (s == d) return;
  const int p = edge_pos[e];
  out[p] = node_pos[s];
  out[(size_t)E + p] = node_pos[d];
  order[p] = e;
}

__global__ void __launch_bounds__(256) prep_loops_kernel(int N, int E, const unsigned char* __restrict__ keep, const int* __restrict__ loop_last,
                                                         const int* __restrict__ node_pos, const int* __restrict__ loop_pos,
                                                         const int* __restrict__ counts, long long* __restrict__ out,
                                                         long long* __restrict__ order) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || !keep[i] || loop_last[i] < 0) return;
  const int p = counts[1] + loop_pos[i];                 // after the non-loop edges
  out[p] = node_pos[i];
  out[(size_t)E + p] = node_pos[i];
  order[p] = loop_last[i];
}

}  // namespace

extern "C" {

size_t phc_isolated_workspace_bytes(int num_nodes, int num_edges) {
  // keep flags live in the caller's mask; here: loop_last[N] + node_pos[N] + loop_pos[N] + edge_pos[E]
  return sizeof(int) * (3 * (size_t)num_nodes + (size_t)num_edges + 4);
}

// counts (device int[4]): [0] kept nodes, [1] kept non-loop edges, [2] kept self loops, [3] status (bit0: index out of range)
int phc_remove_isolated_nodes(const long long* edge_index, int num_edges, int num_nodes, unsigned char* keep_mask, long long* assoc,
                              long long* new_edge_index, long long* edge_order, int* counts, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
  PHC_REQUIRE(num_edges >= 0 && num_nodes >= 0, "phc_remove_isolated_nodes: negative size");
  PHC_REQUIRE(workspace_bytes >= phc_isolated_workspace_bytes(num_nodes, num_edges), "phc_remove_isolated_nodes: workspace too small");
  const int N = num_nodes, E = num_edges;
  int* loop_last = reinterpret_cast<int*>(workspace);
  int* node_pos = loop_last + N;
  int* loop_pos = node_pos + N;
  int* edge_pos = loop_pos + N;
  cudaMemsetAsync(counts, 0, sizeof(int) * 4, stream);
  if (N > 0) {
    cudaMemsetAsync(keep_mask, 0, (size_t)N, stream);
    cudaMemsetAsync(loop_last, 0xFF, sizeof(int) * (size_t)N, stream);       // -1
  }
  if (E > 0) phc_launch(prep_mark_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, N, keep_mask, loop_last, counts + 3);
  if (N > 0) {
    phc_launch(prep_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, stream, 0, N, keep_mask, loop_last, edge_index, E, node_pos, counts + 0);
    phc_launch(prep_assoc_kernel, dim3(phc_div_up(N, 256)), dim3(256), 0, stream, N, keep_mask, node_pos, assoc);
  }
  if (E > 0) {
    phc_launch(prep_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, stream, 1, E, keep_mask, loop_last, edge_index, E, edge_pos, counts + 1);
    phc_launch(prep_edges_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, node_pos, edge_pos, new_edge_index,
               edge_order);
  }
  if (N > 0 && E > 0) {
    phc_launch(prep_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, stream, 2, N, keep_mask, loop_last, edge_index, E, loop_pos, counts + 2);
    phc_launch(prep_loops_kernel, dim3(phc_div_up(N, 256)), dim3(256), 0, stream, N, E, keep_mask, loop_last, node_pos, loop_pos, counts,
               new_edge_index, edge_order);
  }
  return phc_check_launch("phc_remove_isolated_nodes");
}

}  // extern "C"
