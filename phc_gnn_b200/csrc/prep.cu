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

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                               // consecutive items per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;        // items per block

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

// Exclusive scan of flag(i) over n items -> pos[i], total -> *count, in three launches (tile counts, scan of the tile
// counts by one block, positions), every tile of SCAN_TILE items owned by one block and every thread by SCAN_ITEMS
// consecutive items.  Integer sums in a fixed order: deterministic and bit-exact.
// which = 0: flag = keep[i];  1: flag = edge i is not a self loop;  2: flag = keep[i] && loop_last[i] >= 0
__device__ __forceinline__ int prep_flag(int which, int i, const unsigned char* __restrict__ keep, const int* __restrict__ loop_last,
                                         const long long* __restrict__ ei, int E) {
  if (which == 0) return keep[i] ? 1 : 0;
  if (which == 1) return ei[i] != ei[(size_t)E + i] ? 1 : 0;
  return (keep[i] && loop_last[i] >= 0) ? 1 : 0;
}

// exclusive prefix of v over the block's threads (thread order); *total = block sum
__device__ __forceinline__ int prep_block_exclusive(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += up;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const int ws = warp_sums[w];
    if (w < warp) before += ws;
    all += ws;
  }
  __syncthreads();                                          // warp_sums may be reused by the caller's next call
  *total = all;
  return before + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) prep_tile_count_kernel(int which, int n, const unsigned char* __restrict__ keep,
                                                                      const int* __restrict__ loop_last,
                                                                      const long long* __restrict__ ei, int E,
                                                                      int* __restrict__ tile_sums) {
  pdl_begin();
  const int first = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (first + k < n) c += prep_flag(which, first + k, keep, loop_last, ei, E);
  int total;
  prep_block_exclusive(c, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: tile_sums[0..tiles) -> exclusive offsets in place, grand total -> *count
__global__ void __launch_bounds__(SCAN_THREADS) prep_tile_scan_kernel(int tiles, int* __restrict__ tile_sums, int* __restrict__ count) {
  pdl_begin();
  int carry = 0;
  for (int base = 0; base < tiles; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < tiles ? tile_sums[i] : 0;
    int total;
    const int ex = prep_block_exclusive(v, &total);
    if (i < tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) prep_tile_pos_kernel(int which, int n, const unsigned char* __restrict__ keep,
                                                                    const int* __restrict__ loop_last, const long long* __restrict__ ei,
                                                                    int E, const int* __restrict__ tile_offsets,
                                                                    int* __restrict__ pos) {
  pdl_begin();
  const int first = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int f[SCAN_ITEMS];
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    f[k] = (first + k < n) ? prep_flag(which, first + k, keep, loop_last, ei, E) : 0;
    c += f[k];
  }
  int total;
  int run = tile_offsets[blockIdx.x] + prep_block_exclusive(c, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (first + k < n) pos[first + k] = run;
    run += f[k];
  }
}

// the three launches of one scan; tile_sums: scratch of prep_tiles(n) ints
inline int prep_tiles(int n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

inline void prep_scan(int which, int n, const unsigned char* keep, const int* loop_last, const long long* ei, int E, int* pos,
                      int* count, int* tile_sums, cudaStream_t stream) {
  const int tiles = prep_tiles(n);
  phc_launch(prep_tile_count_kernel, dim3(tiles), dim3(SCAN_THREADS), 0, stream, which, n, keep, loop_last, ei, E, tile_sums);
  phc_launch(prep_tile_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, stream, tiles, tile_sums, count);
  phc_launch(prep_tile_pos_kernel, dim3(tiles), dim3(SCAN_THREADS), 0, stream, which, n, keep, loop_last, ei, E,
             (const int*)tile_sums, pos);
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
  // keep flags live in the caller's mask; here: loop_last[N] + node_pos[N] + loop_pos[N] + edge_pos[E] + the scans' tile sums
  const int n = num_nodes > num_edges ? num_nodes : num_edges;
  return sizeof(int) * (3 * (size_t)num_nodes + (size_t)num_edges + (size_t)prep_tiles(n) + 8);
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
  int* tile_sums = edge_pos + E;
  cudaMemsetAsync(counts, 0, sizeof(int) * 4, stream);
  if (N > 0) {
    cudaMemsetAsync(keep_mask, 0, (size_t)N, stream);
    cudaMemsetAsync(loop_last, 0xFF, sizeof(int) * (size_t)N, stream);       // -1
  }
  if (E > 0) phc_launch(prep_mark_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, N, keep_mask, loop_last, counts + 3);
  if (N > 0) {
    prep_scan(0, N, keep_mask, loop_last, edge_index, E, node_pos, counts + 0, tile_sums, stream);
    phc_launch(prep_assoc_kernel, dim3(phc_div_up(N, 256)), dim3(256), 0, stream, N, keep_mask, node_pos, assoc);
  }
  if (E > 0) {
    prep_scan(1, E, keep_mask, loop_last, edge_index, E, edge_pos, counts + 1, tile_sums, stream);
    phc_launch(prep_edges_kernel, dim3(phc_div_up(E, 256)), dim3(256), 0, stream, edge_index, E, node_pos, edge_pos, new_edge_index,
               edge_order);
  }
  if (N > 0 && E > 0) {
    prep_scan(2, N, keep_mask, loop_last, edge_index, E, loop_pos, counts + 2, tile_sums, stream);
    phc_launch(prep_loops_kernel, dim3(phc_div_up(N, 256)), dim3(256), 0, stream, N, E, keep_mask, loop_last, node_pos, loop_pos, counts,
               new_edge_index, edge_order);
  }
  return phc_check_launch("phc_remove_isolated_nodes");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Batch collate on the device: what the reference's DataLoader does on the CPU for every step
// (torch_geometric.data.DataLoader -> Batch.from_data_list, used at benchmarks/train_hiv.py:481-493 and the other
// train_*.py): node / edge tensors of the selected graphs concatenated in batch order, each graph's edge_index shifted
// by the number of nodes before it, `batch` = graph position repeated per node, graph-level targets stacked.
// Here the dataset is RESIDENT IN HBM as one packed store (graphs back to back, edge ids local to their graph), so a
// mini-batch is B contiguous segment copies per tensor — one launch, one pass over the bytes, no host involvement
// beyond the [B] graph ids and their size prefix sums.  Integer / byte work only: bit-exact with the CPU restatement
// (oracle/phc_oracle.py::collate).
namespace {

struct CollateRows {
  const char* src;
  char* dst;
  long long row_bytes;        // multiple of 4; 0 = tensor absent
};

// One contiguous segment (a multiple of 4 bytes, both ends 4-byte aligned) with 16-byte STORES whatever the relative phase of source
// and destination: 28-byte edge_attr rows put a graph's segment at any multiple of 4 bytes in both tensors, which used to force
// 4-byte words (28 % of the HBM peak at ppa size).  Destination words are written as aligned uint4; each is assembled from the two
// aligned source uint4 that cover it (the second one is the neighbour thread's first: an L1 hit), selected by the word shift k.
__device__ __forceinline__ void collate_copy_shifted(const char* src, char* dst, long long bytes, long long tid, long long nth) {
  const unsigned int* sw = reinterpret_cast<const unsigned int*>(src);
  unsigned int* dw = reinterpret_cast<unsigned int*>(dst);
  long long head = (long long)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2;      // words until dst is 16-byte aligned
  const long long words = bytes >> 2;
  if (head > words) head = words;
  for (long long j = tid; j < head; j += nth) dw[j] = sw[j];
  const long long n16 = (words - head) >> 2;
  const int k = (int)((reinterpret_cast<uintptr_t>(sw + head) & 15) >> 2);                       // source phase, in words
  uint4* d4 = reinterpret_cast<uint4*>(dw + head);
  const uint4* s4 = reinterpret_cast<const uint4*>(sw + head - k);                                // aligned down
  // the last destination uint4 would need source words beyond the segment when k != 0: it goes with the scalar tail
  const long long fast = k == 0 ? n16 : (n16 > 0 ? n16 - 1 : 0);
  if (k == 0) {
    for (long long j = tid; j < fast; j += nth) d4[j] = s4[j];
  } else {
    for (long long j = tid; j < fast; j += nth) {
      const uint4 a = s4[j], b = s4[j + 1];
      uint4 r;
      if (k == 1) r = make_uint4(a.y, a.z, a.w, b.x);
      else if (k == 2) r = make_uint4(a.z, a.w, b.x, b.y);
      else r = make_uint4(a.w, b.x, b.y, b.z);
      d4[j] = r;
    }
  }
  for (long long j = head + 4 * fast + tid; j < words; j += nth) dw[j] = sw[j];
}

// rows [first, first+count) of the store -> rows [out_first, ...) of the batch
__device__ __forceinline__ void collate_copy_rows(const CollateRows& r, long long first, long long count, long long out_first,
                                                  long long tid, long long nth) {
  if (r.row_bytes == 0 || count <= 0) return;
  collate_copy_shifted(r.src + first * r.row_bytes, r.dst + out_first * r.row_bytes, count * r.row_bytes, tid, nth);
}

// grid = (slices, B): blockIdx.y = position of the graph in the batch, the slices of one row of blocks share its elements
__global__ void __launch_bounds__(256) collate_kernel(const long long* __restrict__ ids, int B, int G,
                                                      const long long* __restrict__ node_ptr, const long long* __restrict__ edge_ptr,
                                                      const long long* __restrict__ out_node_ptr,
                                                      const long long* __restrict__ out_edge_ptr,
                                                      const long long* __restrict__ ei, long long store_edges,
                                                      long long* __restrict__ out_ei, long long out_edges,
                                                      long long* __restrict__ out_batch, CollateRows xr, CollateRows er, CollateRows yr,
                                                      int* __restrict__ status) {
  pdl_begin();
  const int b = blockIdx.y;
  const long long g = ids[b];
  if (g < 0 || g >= G) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, 1);
    return;
  }
  const long long n0 = node_ptr[g], nn = node_ptr[g + 1] - n0;
  const long long e0 = edge_ptr[g], ne = edge_ptr[g + 1] - e0;
  const long long on0 = out_node_ptr[b], oe0 = out_edge_ptr[b];
  if (nn < 0 || ne < 0 || nn != out_node_ptr[b + 1] - on0 || ne != out_edge_ptr[b + 1] - oe0) {   // offsets not of THESE graphs
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, 2);
    return;
  }
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long j = tid; j < nn; j += nth) out_batch[on0 + j] = b;
#pragma unroll
  for (int half = 0; half < 2; ++half) {               // sources, then targets
    const long long* s = ei + half * store_edges + e0;
    long long* d = out_ei + half * out_edges + oe0;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
      const longlong2* s2 = reinterpret_cast<const longlong2*>(s);
      longlong2* d2 = reinterpret_cast<longlong2*>(d);
      for (long long j = tid; j < (ne >> 1); j += nth) {
        longlong2 v = s2[j];
        v.x += on0; v.y += on0;
        d2[j] = v;
      }
      if ((ne & 1) && tid == 0) d[ne - 1] = s[ne - 1] + on0;
    } else {
      for (long long j = tid; j < ne; j += nth) d[j] = s[j] + on0;
    }
  }
  collate_copy_rows(xr, n0, nn, on0, tid, nth);
  collate_copy_rows(er, e0, ne, oe0, tid, nth);
  if (blockIdx.x == 0) collate_copy_rows(yr, g, 1, b, threadIdx.x, blockDim.x);
}

}  // namespace

extern "C" {

// status (device int, zeroed here): bit0 = graph id outside [0, store_graphs), bit1 = out_*_ptr are not the prefix sums of the
// selected graphs' sizes.  A tensor is skipped when its row size is 0.
int phc_collate_batch(const long long* graph_ids, int num_graphs, int store_graphs, const long long* node_ptr, const long long* edge_ptr,
                      const long long* out_node_ptr, const long long* out_edge_ptr, const long long* edge_index, long long store_edges,
                      long long* out_edge_index, long long out_edges, long long out_nodes, long long* out_batch, const void* x, void* out_x,
                      int x_row_bytes, const void* edge_attr, void* out_edge_attr, int edge_attr_row_bytes, const void* y, void* out_y,
                      int y_row_bytes, int* status, cudaStream_t stream) {
  PHC_REQUIRE(num_graphs >= 0 && store_graphs >= 0 && store_edges >= 0 && out_edges >= 0 && out_nodes >= 0,
              "phc_collate_batch: negative size");
  PHC_REQUIRE(x_row_bytes >= 0 && edge_attr_row_bytes >= 0 && y_row_bytes >= 0 && x_row_bytes % 4 == 0 &&
                  edge_attr_row_bytes % 4 == 0 && y_row_bytes % 4 == 0,
              "phc_collate_batch: row sizes must be non-negative multiples of 4 bytes");
  PHC_REQUIRE(num_graphs <= 65535, "phc_collate_batch: at most 65535 graphs per batch");
  cudaMemsetAsync(status, 0, sizeof(int), stream);
  if (num_graphs == 0) return phc_check_launch("phc_collate_batch");
  CollateRows xr{static_cast<const char*>(x), static_cast<char*>(out_x), (x && out_x) ? (long long)x_row_bytes : 0};
  CollateRows er{static_cast<const char*>(edge_attr), static_cast<char*>(out_edge_attr),
                 (edge_attr && out_edge_attr) ? (long long)edge_attr_row_bytes : 0};
  CollateRows yr{static_cast<const char*>(y), static_cast<char*>(out_y), (y && out_y) ? (long long)y_row_bytes : 0};
  // bytes one graph moves on average -> slices of 256 threads x 64 B each, at most 16 per graph
  const long long per_graph = (out_edges * (16 + er.row_bytes) + out_nodes * (8 + xr.row_bytes)) / num_graphs;
  long long slices = (per_graph + 256 * 64 - 1) / (256 * 64);
  slices = slices < 1 ? 1 : (slices > 16 ? 16 : slices);
  phc_launch(collate_kernel, dim3((unsigned)slices, (unsigned)num_graphs), dim3(256), 0, stream, graph_ids, num_graphs, store_graphs,
             node_ptr, edge_ptr, out_node_ptr, out_edge_ptr, edge_index, store_edges, out_edge_index, out_edges, out_batch, xr, er, yr,
             status);
  return phc_check_launch("phc_collate_batch");
}

}  // extern "C"
