// Hypercomplex feature encoders (the step feeding the message-passing path).
//
//  * integer encoder: out[r, c*Fc + f] = sum_col table[c][col][ idx[r,col], f ]
//      reference: PHMEncoder = n x IntegerEncoder, each a sum of per-column nn.Embedding lookups,
//      stacked on a new component axis (phc/hypercomplex/encoder.py:31-34,
//      phc/quaternion/encoder.py:44-56): n*#cols embedding launches + adds + stack; backward is
//      n*#cols embedding_dense_backward launches (15 % of the reference's CPU step).
//  * linear encoder: out[r, c*Fc + f] = sum_d feat[r,d] * W[c][f,d] + b[c][f]
//      reference: n x nn.Linear(D -> Fc) + stack (phc/hypercomplex/encoder.py:17-20).
//
// The n*#cols parameter tensors stay separate nn.Parameters (state-dict compatibility); the kernels
// take a small by-value table of device pointers.  Backward is deterministic: row chunks are
// accumulated in row order by the thread that owns a feature column, chunk partials are summed in
// chunk order.
#include "common.cuh"

#define PHC_MAX_TABLES 128

struct PtrTable {
  const float* p[PHC_MAX_TABLES];
};
struct MutPtrTable {
  float* p[PHC_MAX_TABLES];
};
struct IntTable {
  int v[PHC_MAX_TABLES];
};

namespace {

template <int VEC>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const long long* __restrict__ idx, PtrTable tables, IntTable vocab, int R, int C,
                                                        int n, int Fc, float* __restrict__ out) {
  pdl_begin();
  const int fcv = Fc / VEC;
  const int F = n * Fc;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * n * fcv) return;
  const int r = (int)(t / (n * fcv));
  const int rem = (int)(t % (n * fcv));
  const int c = rem / fcv, f = (rem % fcv) * VEC;
  Vec<VEC> acc;
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc.v[q] = 0.f;
  for (int col = 0; col < C; ++col) {
    long long v = __ldg(idx + (size_t)r * C + col);
    v = v < 0 ? 0 : (v >= vocab.v[col] ? vocab.v[col] - 1 : v);
    Vec<VEC> w = Vec<VEC>::load(tables.p[c * C + col] + (size_t)v * Fc + f);
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc.v[q] += w.v[q];
  }
  acc.store(out + (size_t)r * F + c * Fc + f);
}

// Component width not a multiple of 4 (ppa: Fc = 125) but flat rows 16-byte aligned: a thread owns 4 consecutive FLAT features
// (they may straddle a component boundary), gathers them with scalar table loads and writes one 128-bit word
// (the per-element kernel above: 7.8 M threads and 41 us for a ppa batch; this one: a quarter of the threads, vector stores).
__global__ void __launch_bounds__(256) embed_fwd_flat4_kernel(const long long* __restrict__ idx, PtrTable tables, IntTable vocab, int R, int C,
                                                              int n, int Fc, float* __restrict__ out) {
  pdl_begin();
  const int F = n * Fc, fq = F / 4;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * fq) return;
  const int r = (int)(t / fq), f = (int)(t % fq) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int c[4], fp[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { c[q] = (f + q) / Fc; fp[q] = (f + q) - c[q] * Fc; }
  for (int col = 0; col < C; ++col) {
    long long v = __ldg(idx + (size_t)r * C + col);
    v = v < 0 ? 0 : (v >= vocab.v[col] ? vocab.v[col] - 1 : v);
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += __ldg(tables.p[c[q] * C + col] + (size_t)v * Fc + fp[q]);
  }
  *reinterpret_cast<float4*>(out + (size_t)r * F + f) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// grid (row chunks, C, feature tiles of blockDim.x flat features). dynamic smem: vocab[col] * blockDim.x floats.
__global__ void embed_bwd_partial_kernel(const float* __restrict__ g, const long long* __restrict__ idx, IntTable vocab, IntTable voff,
                                         int R, int C, int F, int rows_per_chunk, int vtot, float* __restrict__ part) {
  pdl_begin();
  extern __shared__ float acc[];
  const int col = blockIdx.y;
  const int V = vocab.v[col];
  const int ft = blockDim.x;
  const int f = blockIdx.z * ft + threadIdx.x;
  const bool active = f < F;
  for (int v = 0; v < V; ++v) acc[v * ft + threadIdx.x] = 0.f;
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(r0 + rows_per_chunk, R);
  if (active) {
    int r = r0;
    for (; r + 4 <= r1; r += 4) {                     // four rows in flight: the loads are issued before the (ordered) accumulation
      long long v[4];
      float gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = __ldg(idx + (size_t)(r + u) * C + col);
        gv[u] = g[(size_t)(r + u) * F + f];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = (int)(v[u] < 0 ? 0 : (v[u] >= V ? V - 1 : v[u]));
        acc[vi * ft + threadIdx.x] += gv[u];
      }
    }
    for (; r < r1; ++r) {
      long long v = __ldg(idx + (size_t)r * C + col);
      v = v < 0 ? 0 : (v >= V ? V - 1 : v);
      acc[(int)v * ft + threadIdx.x] += g[(size_t)r * F + f];
    }
    float* dst = part + ((size_t)blockIdx.x * vtot + voff.v[col]) * F + f;
    for (int v = 0; v < V; ++v) dst[(size_t)v * F] = acc[v * ft + threadIdx.x];
  }
}

// dtable[c][col][v, f'] = sum over chunks of part[chunk][voff[col]+v][c*Fc+f'].  Block = 32 consecutive (vocab row, feature)
// elements x 8 chunk groups: group y adds chunks y, y+8, ... (coalesced over the 32 elements), the 8 group sums are added in
// group order — a fixed tree, so the result is reproducible.  (One thread per element walking all chunks: 22 us at 244 chunks.)
__global__ void __launch_bounds__(256) embed_bwd_final_kernel(const float* __restrict__ part, MutPtrTable dtables, IntTable vocab,
                                                              IntTable voff, int chunks, int C, int n, int Fc, int vtot) {
  pdl_begin();
  const int F = n * Fc;
  __shared__ float red[8][32];
  const int lx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const long long t = (long long)blockIdx.x * 32 + lx;
  const bool on = t < (long long)vtot * F;
  float s = 0.f;
  if (on)
    for (int k = gy; k < chunks; k += 8) s += part[(size_t)k * vtot * F + t];
  red[gy][lx] = s;
  __syncthreads();
  if (gy == 0 && on) {
    float tot = red[0][lx];
#pragma unroll
    for (int y = 1; y < 8; ++y) tot += red[y][lx];
    const int vr = (int)(t / F), f = (int)(t % F);
    int col = 0;
    while (col + 1 < C && voff.v[col + 1] <= vr) ++col;
    const int v = vr - voff.v[col];
    const int c = f / Fc, fp = f % Fc;
    dtables.p[c * C + col][(size_t)v * Fc + fp] = tot;
  }
}

// ---- linear encoder ------------------------------------------------------------------------
constexpr int LIN_MAX_D = 16;
constexpr int LIN_ROWS = 32;     // rows per block in forward

// One thread owns 4 consecutive flat features: its 4 x D weights and 4 biases live in registers and are
// reused for LIN_ROWS rows; per row it reads D broadcast features and writes one 128-bit result.
template <int D>
__global__ void __launch_bounds__(128) linenc_fwd_kernel(const float* __restrict__ feat, PtrTable weights, PtrTable biases, int R, int Fc,
                                                         int F, float* __restrict__ out) {
  pdl_begin();
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (f >= F) return;
  float w[4][D], bq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = (f + q) / Fc, fp = (f + q) - c * Fc;
    bq[q] = biases.p[c] ? __ldg(biases.p[c] + fp) : 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) w[q][d] = __ldg(weights.p[c] + (size_t)fp * D + d);
  }
  const int r0 = blockIdx.y * LIN_ROWS, r1 = min(r0 + LIN_ROWS, R);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    float a[D];
#pragma unroll
    for (int d = 0; d < D; ++d) a[d] = __ldg(feat + (size_t)r * D + d);
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = bq[q];
#pragma unroll
      for (int d = 0; d < D; ++d) v += a[d] * w[q][d];
      o[q] = v;
    }
    Vec<4> t; t.v[0] = o[0]; t.v[1] = o[1]; t.v[2] = o[2]; t.v[3] = o[3];
    t.store_stream(out + (size_t)r * F + f);
  }
}

// generic fallback (F % 4 != 0 or unaligned): one thread per output element
__global__ void __launch_bounds__(256) linenc_fwd_scalar_kernel(const float* __restrict__ feat, PtrTable weights, PtrTable biases, int R,
                                                                int D, int Fc, int F, float* __restrict__ out) {
  pdl_begin();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * F) return;
  const int r = (int)(t / F), f = (int)(t % F);
  const int c = f / Fc, fp = f - c * Fc;
  float v = biases.p[c] ? __ldg(biases.p[c] + fp) : 0.f;
  for (int d = 0; d < D; ++d) v += __ldg(feat + (size_t)r * D + d) * __ldg(weights.p[c] + (size_t)fp * D + d);
  out[(size_t)r * F + f] = v;
}

// part[chunk][f][0..D] : d = 0..D-1 weight grads, d = D bias grad. grid (feature quads / 128, row chunks).
// One thread owns 4 features (one 128-bit load of g per row) and 4 x (D+1) accumulators.
template <int D>
__global__ void __launch_bounds__(128) linenc_bwd_partial_kernel(const float* __restrict__ g, const float* __restrict__ feat, int R, int F,
                                                                 int rows_per_chunk, float* __restrict__ part) {
  pdl_begin();
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (f >= F) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, R);
  float acc[4][D + 1];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int d = 0; d <= D; ++d) acc[q][d] = 0.f;
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const Vec<4> gv = Vec<4>::load_stream(g + (size_t)r * F + f);
    float a[D];
#pragma unroll
    for (int d = 0; d < D; ++d) a[d] = __ldg(feat + (size_t)r * D + d);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int d = 0; d < D; ++d) acc[q][d] += gv.v[q] * a[d];
      acc[q][D] += gv.v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float* dst = part + ((size_t)blockIdx.y * F + f + q) * (D + 1);
#pragma unroll
    for (int d = 0; d <= D; ++d) dst[d] = acc[q][d];
  }
}

__global__ void __launch_bounds__(128) linenc_bwd_partial_scalar_kernel(const float* __restrict__ g, const float* __restrict__ feat, int R,
                                                                        int D, int F, int rows_per_chunk, float* __restrict__ part) {
  pdl_begin();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, R);
  float acc[LIN_MAX_D + 1];
#pragma unroll
  for (int d = 0; d <= LIN_MAX_D; ++d) acc[d] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float gv = g[(size_t)r * F + f];
#pragma unroll
    for (int d = 0; d < LIN_MAX_D; ++d)
      if (d < D) acc[d] += gv * __ldg(feat + (size_t)r * D + d);
    acc[LIN_MAX_D] += gv;
  }
  float* dst = part + ((size_t)blockIdx.y * F + f) * (D + 1);
#pragma unroll
  for (int d = 0; d < LIN_MAX_D; ++d)
    if (d < D) dst[d] = acc[d];
  dst[D] = acc[LIN_MAX_D];
}

__global__ void __launch_bounds__(256) linenc_bwd_final_kernel(const float* __restrict__ part, MutPtrTable dweights, MutPtrTable dbiases,
                                                               int chunks, int D, int n, int Fc) {
  pdl_begin();
  const int F = n * Fc;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)F * (D + 1)) return;
  const int f = (int)(t / (D + 1)), d = (int)(t % (D + 1));
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[((size_t)k * F + f) * (D + 1) + d];
  const int c = f / Fc, fp = f % Fc;
  if (d < D) dweights.p[c][(size_t)fp * D + d] = s;
  else if (dbiases.p[c]) dbiases.p[c][fp] = s;
}

template <int D>
void launch_linenc_fwd(const float* feat, const PtrTable& w, const PtrTable& b, int R, int Fc, int F, float* out, cudaStream_t st) {
  dim3 grid(phc_div_up(F / 4, 128), phc_div_up(R, LIN_ROWS));
  phc_launch(linenc_fwd_kernel<D>, dim3(grid), dim3(128), 0, st, feat, w, b, R, Fc, F, out);
}
template <int D>
void launch_linenc_bwd(const float* g, const float* feat, int R, int F, int rpc, int chunks, float* part, cudaStream_t st) {
  dim3 grid(phc_div_up(F / 4, 128), chunks);
  phc_launch(linenc_bwd_partial_kernel<D>, dim3(grid), dim3(128), 0, st, g, feat, R, F, rpc, part);
}

// row chunks of the embedding backward: 64 rows each (a chunk is one thread's serial walk; round 1 used 512-row chunks, i.e. 124
// blocks for a ppa batch — fewer than SMs — and 55 us), bounded so that the [chunks][vocab][F] partials stay below 32 MiB
int embed_chunks(int R, int vtot, int F) {
  long long chunks = phc_div_up(R, 64);
  const long long cap = (32ll << 20) / (4ll * (vtot > 0 ? vtot : 1) * (F > 0 ? F : 1));
  if (chunks > cap) chunks = cap;
  if (chunks > 1024) chunks = 1024;
  return chunks < 1 ? 1 : (int)chunks;
}
int linenc_chunks(int R) {
  int chunks = phc_div_up(R, 128);
  return chunks < 1 ? 1 : (chunks > 1184 ? 1184 : chunks);
}

}  // namespace

namespace {
// The embedding kernels (here and the fused edge encoder in conv_fused.cu) CLAMP an index outside its table, where nn.Embedding raises
// (a device-side assertion that would poison the context).  This check is the reporting side: it ORs bit 0 into *status for any
// index outside [0, vocab[col]); the host reads the word when it wants to know (ops.validate_indices — debug / tests, synchronising).
__global__ void __launch_bounds__(256) index_check_kernel(const long long* __restrict__ idx, IntTable vocab, long long total, int C,
                                                          int* __restrict__ status) {
  pdl_begin();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long v = __ldg(idx + t);
  if (v < 0 || v >= vocab.v[(int)(t % C)]) atomicOr(status, 1);
}
}  // namespace

extern "C" {

int phc_index_check(const long long* idx, const int* vocab, int rows, int cols, int* status, cudaStream_t stream) {
  PHC_REQUIRE(cols > 0 && cols <= PHC_MAX_TABLES && status != nullptr, "phc_index_check: cols=%d (max %d), status required", cols, PHC_MAX_TABLES);
  if (rows == 0) return PHC_OK;
  IntTable vc;
  for (int i = 0; i < cols; ++i) vc.v[i] = vocab[i];
  const long long total = (long long)rows * cols;
  phc_launch(index_check_kernel, dim3(phc_div_up(total, 256)), dim3(256), 0, stream, idx, vc, total, cols, status);
  return phc_check_launch("phc_index_check");
}

size_t phc_embed_bwd_workspace_bytes(int rows, int total_vocab, int width) {
  return sizeof(float) * (size_t)embed_chunks(rows, total_vocab, width) * total_vocab * width;
}

int phc_embed_sum_fwd(const long long* idx, const float* const* tables, const int* vocab, int rows, int cols, int phm_dim,
                      int width_per_component, float* out, cudaStream_t stream) {
  PHC_REQUIRE(cols > 0 && phm_dim > 0 && phm_dim * cols <= PHC_MAX_TABLES, "phc_embed_sum_fwd: phm_dim*cols=%d exceeds %d tables",
              phm_dim * cols, PHC_MAX_TABLES);
  if (rows == 0) return PHC_OK;
  PtrTable tb; IntTable vc;
  bool v4 = width_per_component % 4 == 0 && phc_aligned16(out);
  for (int i = 0; i < phm_dim * cols; ++i) { tb.p[i] = tables[i]; v4 = v4 && phc_aligned16(tables[i]); }
  for (int i = 0; i < cols; ++i) vc.v[i] = vocab[i];
  const int n = phm_dim, Fc = width_per_component;
  if (v4) phc_launch(embed_fwd_kernel<4>, dim3(phc_div_up((long long)rows * n * (Fc / 4), 256)), dim3(256), 0, stream, idx, tb, vc, rows, cols, n, Fc, out);
  else if ((n * Fc) % 4 == 0 && phc_aligned16(out))
    phc_launch(embed_fwd_flat4_kernel, dim3(phc_div_up((long long)rows * (n * Fc / 4), 256)), dim3(256), 0, stream, idx, tb, vc, rows, cols, n, Fc, out);
  else phc_launch(embed_fwd_kernel<1>, dim3(phc_div_up((long long)rows * n * Fc, 256)), dim3(256), 0, stream, idx, tb, vc, rows, cols, n, Fc, out);
  return phc_check_launch("phc_embed_sum_fwd");
}

int phc_embed_sum_bwd(const float* gout, const long long* idx, float* const* dtables, const int* vocab, int rows, int cols, int phm_dim,
                      int width_per_component, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(cols > 0 && phm_dim > 0 && phm_dim * cols <= PHC_MAX_TABLES, "phc_embed_sum_bwd: too many tables");
  const int n = phm_dim, Fc = width_per_component, F = n * Fc;
  MutPtrTable dt; IntTable vc, vo;
  int vtot = 0, vmax = 0;
  for (int i = 0; i < cols; ++i) { vc.v[i] = vocab[i]; vo.v[i] = vtot; vtot += vocab[i]; vmax = vocab[i] > vmax ? vocab[i] : vmax; }
  for (int i = 0; i < n * cols; ++i) dt.p[i] = dtables[i];
  PHC_REQUIRE(workspace_bytes >= phc_embed_bwd_workspace_bytes(rows, vtot, F), "phc_embed_sum_bwd: workspace too small");
  const int ft = 128;
  const size_t smem = sizeof(float) * (size_t)vmax * ft;
  PHC_REQUIRE(smem <= 200 * 1024, "phc_embed_sum_bwd: vocabulary %d too large for the shared-memory accumulator", vmax);
  const int chunks = embed_chunks(rows, vtot, F);
  const int rpc = phc_div_up(rows > 0 ? rows : 1, chunks);
  float* part = reinterpret_cast<float*>(workspace);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(embed_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid(chunks, cols, phc_div_up(F, ft));
  phc_launch(embed_bwd_partial_kernel, dim3(grid), dim3(ft), smem, stream, gout, idx, vc, vo, rows, cols, F, rpc, vtot, part);
  phc_launch(embed_bwd_final_kernel, dim3(phc_div_up((long long)vtot * F, 32)), dim3(256), 0, stream, part, dt, vc, vo, chunks, cols, n, Fc, vtot);
  return phc_check_launch("phc_embed_sum_bwd");
}

size_t phc_linear_encoder_bwd_workspace_bytes(int rows, int in_dim, int width) {
  return sizeof(float) * (size_t)linenc_chunks(rows) * width * (in_dim + 1);
}

int phc_linear_encoder_fwd(const float* feat, const float* const* weights, const float* const* biases, int rows, int in_dim, int phm_dim,
                           int width_per_component, float* out, cudaStream_t stream) {
  PHC_REQUIRE(phm_dim > 0 && phm_dim <= PHC_MAX_TABLES, "phc_linear_encoder_fwd: bad phm_dim");
  PHC_REQUIRE(in_dim > 0 && in_dim <= LIN_MAX_D, "phc_linear_encoder_fwd: in_dim %d not in 1..%d", in_dim, LIN_MAX_D);
  if (rows == 0) return PHC_OK;
  PtrTable w, b;
  for (int c = 0; c < phm_dim; ++c) { w.p[c] = weights[c]; b.p[c] = biases ? biases[c] : nullptr; }
  const int n = phm_dim, Fc = width_per_component, F = n * Fc;
  const bool v4 = F % 4 == 0 && phc_aligned16(out);
  if (v4) {
    switch (in_dim) {
#define PHC_CASE(D) case D: launch_linenc_fwd<D>(feat, w, b, rows, Fc, F, out, stream); break;
      PHC_CASE(1) PHC_CASE(2) PHC_CASE(3) PHC_CASE(4) PHC_CASE(5) PHC_CASE(6) PHC_CASE(7) PHC_CASE(8)
      PHC_CASE(9) PHC_CASE(10) PHC_CASE(11) PHC_CASE(12) PHC_CASE(13) PHC_CASE(14) PHC_CASE(15) PHC_CASE(16)
#undef PHC_CASE
    }
  } else {
    phc_launch(linenc_fwd_scalar_kernel, dim3(phc_div_up((long long)rows * F, 256)), dim3(256), 0, stream, feat, w, b, rows, in_dim, Fc, F, out);
  }
  return phc_check_launch("phc_linear_encoder_fwd");
}

int phc_linear_encoder_bwd(const float* gout, const float* feat, float* const* dweights, float* const* dbiases, int rows, int in_dim,
                           int phm_dim, int width_per_component, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(phm_dim > 0 && phm_dim <= PHC_MAX_TABLES, "phc_linear_encoder_bwd: bad phm_dim");
  PHC_REQUIRE(in_dim > 0 && in_dim <= LIN_MAX_D, "phc_linear_encoder_bwd: in_dim %d not in 1..%d", in_dim, LIN_MAX_D);
  const int n = phm_dim, Fc = width_per_component, F = n * Fc;
  PHC_REQUIRE(workspace_bytes >= phc_linear_encoder_bwd_workspace_bytes(rows, in_dim, F), "phc_linear_encoder_bwd: workspace too small");
  MutPtrTable dw, db;
  for (int c = 0; c < n; ++c) { dw.p[c] = dweights[c]; db.p[c] = dbiases ? dbiases[c] : nullptr; }
  const int chunks = linenc_chunks(rows);
  const int rpc = phc_div_up(rows > 0 ? rows : 1, chunks);
  float* part = reinterpret_cast<float*>(workspace);
  if (F % 4 == 0 && phc_aligned16(gout)) {
    switch (in_dim) {
#define PHC_CASE(D) case D: launch_linenc_bwd<D>(gout, feat, rows, F, rpc, chunks, part, stream); break;
      PHC_CASE(1) PHC_CASE(2) PHC_CASE(3) PHC_CASE(4) PHC_CASE(5) PHC_CASE(6) PHC_CASE(7) PHC_CASE(8)
      PHC_CASE(9) PHC_CASE(10) PHC_CASE(11) PHC_CASE(12) PHC_CASE(13) PHC_CASE(14) PHC_CASE(15) PHC_CASE(16)
#undef PHC_CASE
    }
  } else {
    dim3 grid(phc_div_up(F, 128), chunks);
    phc_launch(linenc_bwd_partial_scalar_kernel, dim3(grid), dim3(128), 0, stream, gout, feat, rows, in_dim, F, rpc, part);
  }
  phc_launch(linenc_bwd_final_kernel, dim3(phc_div_up((long long)F * (in_dim + 1), 256)), dim3(256), 0, stream, part, dw, db, chunks, in_dim, n, Fc);
  return phc_check_launch("phc_linear_encoder_bwd");
}

}  // extern "C"
