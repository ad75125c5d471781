// PHMLinear, exact-fp32 path on the FFMA pipe ("fp32" precision mode).
//
//   y[m, c*P+p] = sum_{a,k} x[m, a*K+k] * H[(a,k),(c,p)] + b[c*P+p],   H[(a,k),(c,p)] = sum_b A[b,a,c] W[b,k,p]
//
// Replaces reference phc/hypercomplex/layers.py:198-219 (matvec_product_new): an einsum that
// materialises the [n,in,out] Kronecker stack, a sum over it, a cuBLAS SGEMM and a bias kernel —
// again in backward.  Here H never exists in memory: every CTA builds the [BK,BN] tile of H it
// needs in shared memory from the n^3 rule scalars and the n small W blocks (n FMAs per element).
//
// This file is the strict-fp32 mode used for rtol-1e-4 parity checks of everything else; the
// tensor-core path (tcgen05, phm_linear_tc.cu) is the production mode.
//
// Backward:  dX = dY * H^T (same kernel, transposed tile builder),  dH = X^T dY via split-M partial
// tiles reduced in a fixed order (no atomics), then  dW[b,k,p] = sum_{a,c} A[b,a,c] dH[(a,k),(c,p)],
// dA[b,a,c] = sum_{k,p} W[b,k,p] dH[(a,k),(c,p)],  db = column sums of dY.
#include "common.cuh"

namespace {

constexpr int TM = 128, TN = 64, TK = 16;   // forward / dX tile
constexpr int MAXN = 16;                    // largest supported phm_dim

// Build element (r,q) of H (TRANS=false: r indexes `in`, q indexes `out`) or of H^T (TRANS=true:
// r indexes `out`, q indexes `in`).
template <bool TRANS>
__device__ __forceinline__ float h_element(const float* __restrict__ As, const float* __restrict__ W, int n, int K, int P, int r, int q) {
  int a, k, c, p;
  if (!TRANS) { a = r / K; k = r % K; c = q / P; p = q % P; }
  else { c = r / P; p = r % P; a = q / K; k = q % K; }
  float h = 0.f;
  for (int b = 0; b < n; ++b) h += As[(b * n + a) * n + c] * __ldg(W + ((size_t)b * K + k) * P + p);
  return h;
}

// C[M, Nout] = act(X[M, Kred] * B[Kred, Nout] + bias) + residual, with B built on the fly.
template <bool TRANS>
__global__ void __launch_bounds__(256) phm_gemm_kernel(const float* __restrict__ X, const float* __restrict__ A, const float* __restrict__ W,
                                                       const float* __restrict__ bias, const float* __restrict__ residual,
                                                       float* __restrict__ Y, int M, int Kred, int Nout, int n, int K, int P, int act) {
  pdl_begin();
  __shared__ float Xs[TK][TM + 4];
  __shared__ float Hs[TK][TN + 4];
  extern __shared__ float As[];  // n^3 rule scalars
  const int tid = threadIdx.x;
  for (int i = tid; i < n * n * n; i += 256) As[i] = A[i];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int ty = tid / 16, tx = tid % 16;  // 16x16 threads, each 8 rows x 4 cols
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  __syncthreads();

  for (int k0 = 0; k0 < Kred; k0 += TK) {
    // X tile: 128 rows x 16 k  -> 2048 elements, 8 per thread (coalesced along k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int lin = tid + i * 256;
      int mm = lin / TK, kk = lin % TK;
      int gm = m0 + mm, gk = k0 + kk;
      Xs[kk][mm] = (gm < M && gk < Kred) ? X[(size_t)gm * Kred + gk] : 0.f;
    }
    // H tile: 16 x 64 -> 1024 elements, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int lin = tid + i * 256;
      int kk = lin / TN, nn = lin % TN;
      int gr = k0 + kk, gq = n0 + nn;
      Hs[kk][nn] = (gr < Kred && gq < Nout) ? h_element<TRANS>(As, W, n, K, P, gr, gq) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float xa[8], hb[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) xa[i] = Xs[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) hb[j] = Hs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += xa[i] * hb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gq = n0 + tx * 4 + j;
      if (gq >= Nout) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + gq);
      v = act_fwd_rt(act, v);
      if (residual) v += residual[(size_t)gm * Nout + gq];
      Y[(size_t)gm * Nout + gq] = v;
    }
  }
}

// dH partials: part[z][i][o] = sum_{m in split z} x[m,i] * dy[m,o].  64x64 tile, 256 threads x (4x4).
__global__ void __launch_bounds__(256) phm_dh_kernel(const float* __restrict__ X, const float* __restrict__ G, int M, int In, int Out,
                                                     int rows_per_split, float* __restrict__ part) {
  pdl_begin();
  __shared__ float Xs[16][64 + 4];
  __shared__ float Gs[16][64 + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int i0 = blockIdx.y * 64, o0 = blockIdx.x * 64;
  const int mbeg = blockIdx.z * rows_per_split, mend = min(mbeg + rows_per_split, M);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m0 = mbeg; m0 < mend; m0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int lin = tid + i * 256;
      int mm = lin / 64, cc = lin % 64;
      int gm = m0 + mm;
      Xs[mm][cc] = (gm < mend && i0 + cc < In) ? X[(size_t)gm * In + i0 + cc] : 0.f;
      Gs[mm][cc] = (gm < mend && o0 + cc < Out) ? G[(size_t)gm * Out + o0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      float xa[4], gb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = Xs[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) gb[j] = Gs[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += xa[i] * gb[j];
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.z * In * Out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gi = i0 + ty * 4 + i;
    if (gi >= In) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int go = o0 + tx * 4 + j;
      if (go < Out) dst[(size_t)gi * Out + go] = acc[i][j];
    }
  }
}

// One thread per (k,p): folds the split partials, contracts with A for dW and produces block
// partials of dA[b,a,c] (reduced in fixed order: warp shuffle tree, then warps in order).
__global__ void __launch_bounds__(256) phm_contract_kernel(const float* __restrict__ part, int splits, const float* __restrict__ A,
                                                           const float* __restrict__ W, int n, int K, int P, float* __restrict__ dW,
                                                           float* __restrict__ dA_part) {
  pdl_begin();
  extern __shared__ float sm[];
  float* As = sm;                 // n^3
  float* wred = sm + n * n * n;   // 8 warps
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n * n * n; i += 256) As[i] = A[i];
  __syncthreads();
  const int In = n * K, Out = n * P;
  const long long t = (long long)blockIdx.x * 256 + tid;
  const bool active = t < (long long)K * P;
  const int k = active ? (int)(t / P) : 0, p = active ? (int)(t % P) : 0;
  float dw[MAXN];
#pragma unroll
  for (int b = 0; b < MAXN; ++b) dw[b] = 0.f;
  for (int a = 0; a < n; ++a) {
    for (int c = 0; c < n; ++c) {
      float dh = 0.f;
      if (active) {
        const size_t off = (size_t)(a * K + k) * Out + (c * P + p);
        for (int s = 0; s < splits; ++s) dh += part[(size_t)s * In * Out + off];
      }
#pragma unroll
      for (int b = 0; b < MAXN; ++b) {
        if (b < n) {
          dw[b] += As[(b * n + a) * n + c] * dh;
          float v = active ? __ldg(W + ((size_t)b * K + k) * P + p) * dh : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) wred[warp] = v;
          __syncthreads();
          if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += wred[w];
            dA_part[(size_t)blockIdx.x * n * n * n + (b * n + a) * n + c] = s;
          }
          __syncthreads();
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int b = 0; b < MAXN; ++b)
      if (b < n) dW[((size_t)b * K + k) * P + p] = dw[b];
  }
}

// Last step of the backward: ordered sums of the per-block dA partials (one warp per rule entry: lane l takes blocks
// l, l+32, ..., then a fixed butterfly) and, when the dH kernel produced column-sum partials of dy, of the bias gradient
// (one thread per column, partials in order).
__global__ void __launch_bounds__(256) phm_bwd_final_kernel(const float* __restrict__ dA_part, int blocks, int n3, float* __restrict__ dA,
                                                            const float* __restrict__ db_part, int db_parts, int Out,
                                                            float* __restrict__ db) {
  pdl_begin();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (t < n3 * 32) {
    const int i = t >> 5;
    float s = 0.f;
#pragma unroll 4
    for (int b = lane; b < blocks; b += 32) s += dA_part[(size_t)b * n3 + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dA[i] = s;
    return;
  }
  const int f = t - n3 * 32;
  if (db_part != nullptr && f < Out) {
    float s = 0.f;
#pragma unroll 4
    for (int q = 0; q < db_parts; ++q) s += db_part[(size_t)q * Out + f];
    db[f] = s;
  }
}

// column sums of G (bias gradient): chunk partials then ordered final sum.
// block = 32 feature lanes (x4 floats) x 8 row slices; slices reduced in fixed order through shared memory.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ G, int M, int F, int rows_per_chunk,
                                                             float* __restrict__ part) {
  pdl_begin();
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, M);
  __shared__ float red[8][128];
  const bool vec = (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(G) & 15u) == 0);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const int f = (blockIdx.x * 32 + lane) * 4;
  if (f < F) {
#pragma unroll 4
    for (int r = r0 + slice; r < r1; r += 8) {
      if (vec) {
        const float4 v = *reinterpret_cast<const float4*>(G + (size_t)r * F + f);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (f + q < F) s[q] += G[(size_t)r * F + f + q];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) red[slice][lane * 4 + q] = s[q];
  __syncthreads();
  if (slice == 0 && f < F) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = 0.f;
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) t += red[sl][lane * 4 + q];
      if (f + q < F) part[(size_t)blockIdx.y * F + f + q] = t;
    }
  }
}
// one warp per column: lane l sums chunks l, l+32, ... in order, then a fixed butterfly
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int chunks, int F, float* __restrict__ out) {
  pdl_begin();
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += part[(size_t)c * F + f];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[f] = s;
}

// Fast contraction for small phm_dim (N <= 4).  Thread (a, j): input component a of the block's j-th (k,p) pair
// (64 pairs per block, 64*N threads, the two warps of one `a` are adjacent).  The thread folds the split partials of
// its N dH entries (a; c = 0..N-1) with independent loads, contributes sum_c A[b,a,c] dh[c] to dW[b,k,p] (summed over a
// through shared memory, a = 0..N-1 in order) and w[b] dh[c] to dA[b,a,c] (warp butterfly, then the a's two warps).
// CG thread groups share the split list (group g folds splits g, g + CG, ...; the group sums are added in group order):
// with the 18 short splits of the wide dH kernel one group spends its time in dependent L2 round trips.
constexpr int CJ = 64, CG = 3;
template <int N>
__global__ void __launch_bounds__(CG * CJ * N) phm_contract_small_kernel(const float* __restrict__ part, int splits, const float* __restrict__ A,
                                                                         const float* __restrict__ W, int K, int P, float* __restrict__ dW,
                                                                         float* __restrict__ dA_part) {
  pdl_begin();
  constexpr int N2 = N * N, N3 = N * N * N;
  __shared__ float As[N3];
  __shared__ float dws[N][CJ][N];
  __shared__ float red[2 * N][N2];
  __shared__ float dhs[CG - 1][CJ * N][N];
  const int grp = threadIdx.x / (CJ * N);
  const int tid = threadIdx.x - grp * (CJ * N), lane = tid & 31, warp = tid >> 5;
  const int a = tid / CJ, j = tid % CJ;
  for (int i = threadIdx.x; i < N3; i += CG * CJ * N) As[i] = A[i];
  __syncthreads();
  const int In = N * K, Out = N * P;
  const long long t = (long long)blockIdx.x * CJ + j;
  const bool active = t < (long long)K * P;
  const int k = active ? (int)(t / P) : 0, p = active ? (int)(t % P) : 0;
  float dh[N], w[N];
#pragma unroll
  for (int c = 0; c < N; ++c) dh[c] = 0.f;
#pragma unroll
  for (int b = 0; b < N; ++b) w[b] = active ? __ldg(W + ((size_t)b * K + k) * P + p) : 0.f;
  if (active) {
    const float* src = part + (size_t)(a * K + k) * Out + p;
    const size_t stride = (size_t)In * Out;
#pragma unroll 4
    for (int s = grp; s < splits; s += CG)
#pragma unroll
      for (int c = 0; c < N; ++c) dh[c] += src[(size_t)s * stride + c * P];
  }
  if (grp > 0) {
#pragma unroll
    for (int c = 0; c < N; ++c) dhs[grp - 1][tid][c] = dh[c];
  }
  __syncthreads();
  if (grp > 0) return;                       // group 0 carries on with the folded values
#pragma unroll
  for (int g2 = 0; g2 < CG - 1; ++g2)
#pragma unroll
    for (int c = 0; c < N; ++c) dh[c] += dhs[g2][tid][c];
#pragma unroll
  for (int b = 0; b < N; ++b) {
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < N; ++c) v += As[(b * N + a) * N + c] * dh[c];
    dws[a][j][b] = v;
  }
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      float v = w[b] * dh[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][b * N + c] = v;
    }
  asm volatile("bar.sync 1, %0;" ::"r"(CJ * N) : "memory");      // group 0 only (the other groups have left)
  if (a == 0 && active) {
#pragma unroll
    for (int b = 0; b < N; ++b) {
      float v = 0.f;
#pragma unroll
      for (int aa = 0; aa < N; ++aa) v += dws[aa][j][b];
      dW[((size_t)b * K + k) * P + p] = v;
    }
  }
  if (tid < N3) {
    const int b = tid / N2, aa = (tid / N) % N, c = tid % N;
    dA_part[(size_t)blockIdx.x * N3 + tid] = red[2 * aa][b * N + c] + red[2 * aa + 1][b * N + c];
  }
}

int dh_splits(int M, int In, int Out) {
  long long tiles = (long long)phc_div_up(In, 64) * phc_div_up(Out, 64);
  int s = (int)((2 * 148 + tiles - 1) / tiles);
  int maxs = phc_div_up(M, 64);
  s = s < 1 ? 1 : s;
  s = s > maxs ? maxs : s;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}
int colsum_chunks(int M) {
  int c = phc_div_up(M, 64);
  return c < 1 ? 1 : (c > 1024 ? 1024 : c);
}

}  // namespace

// ---- shared with phm_linear_tc.cu: fold dH partials into dW / dA, and the bias gradient ----------
static int contract_blocks(int n, int K, int P) { return phc_div_up((long long)K * P, n <= 4 ? CJ : 256); }
static int bias_part_rows(int rows) { int c = colsum_chunks(rows); return c > 2 * 64 ? c : 2 * 64; }   // colsum chunks or 2 x dH splits

size_t phm_contract_scratch_floats(int rows, int in_features, int out_features, int phm_dim) {
  const int n = phm_dim, K = in_features / n, P = out_features / n;
  return (size_t)contract_blocks(n, K, P) * n * n * n + (size_t)bias_part_rows(rows) * out_features + 16;
}

// the bias-gradient partials live behind the dA partials in the scratch area
float* phm_contract_bias_partials(float* scratch, int in_features, int out_features, int phm_dim) {
  const int n = phm_dim;
  return scratch + (size_t)contract_blocks(n, in_features / n, out_features / n) * n * n * n;
}

// db_parts > 0: the dH kernel already wrote db_parts column-sum partials of gy to phm_contract_bias_partials(scratch, ...)
int phm_contract_and_bias(const float* part, int splits, const float* gy, const float* A, const float* W, float* dA, float* dW, float* db,
                          int rows, int in_features, int out_features, int phm_dim, float* scratch, int db_parts, cudaStream_t stream) {
  const int n = phm_dim, K = in_features / n, P = out_features / n, M = rows, n3 = n * n * n;
  const int cblocks = contract_blocks(n, K, P);
  float* da_part = scratch;
  float* cs_part = phm_contract_bias_partials(scratch, in_features, out_features, n);
  switch (n) {
    case 1: phc_launch(phm_contract_small_kernel<1>, dim3(cblocks), dim3(CG * CJ * 1), 0, stream, part, splits, A, W, K, P, dW, da_part); break;
    case 2: phc_launch(phm_contract_small_kernel<2>, dim3(cblocks), dim3(CG * CJ * 2), 0, stream, part, splits, A, W, K, P, dW, da_part); break;
    case 3: phc_launch(phm_contract_small_kernel<3>, dim3(cblocks), dim3(CG * CJ * 3), 0, stream, part, splits, A, W, K, P, dW, da_part); break;
    case 4: phc_launch(phm_contract_small_kernel<4>, dim3(cblocks), dim3(CG * CJ * 4), 0, stream, part, splits, A, W, K, P, dW, da_part); break;
    default: phc_launch(phm_contract_kernel, dim3(cblocks), dim3(256), sizeof(float) * (n3 + 8), stream, part, splits, A, W, n, K, P, dW, da_part);
  }
  const bool fused_bias = db != nullptr && db_parts > 0;
  if (db && !fused_bias) {
    const int chunks = colsum_chunks(M);
    const int rpc = phc_div_up(M > 0 ? M : 1, chunks);
    dim3 g3(phc_div_up(out_features, 128), chunks);
    phc_launch(colsum_partial_kernel, dim3(g3), dim3(256), 0, stream, gy, M, out_features, rpc, cs_part);
    phc_launch(colsum_final_kernel, dim3(phc_div_up((long long)out_features * 32, 256)), dim3(256), 0, stream, cs_part, chunks, out_features, db);
  }
  if (dA || fused_bias) {
    const int na = dA ? n3 : 0;
    phc_launch(phm_bwd_final_kernel, dim3(phc_div_up((long long)na * 32 + (fused_bias ? out_features : 0), 256)), dim3(256), 0, stream, da_part, cblocks, na, dA, fused_bias ? cs_part : nullptr, db_parts, out_features, db);
  }
  return phc_check_launch("phm_contract_and_bias");
}

size_t phm_simt_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim) {
  size_t dh = (size_t)dh_splits(rows, in_features, out_features) * in_features * out_features;
  return sizeof(float) * (dh + phm_contract_scratch_floats(rows, in_features, out_features, phm_dim)) + 64;
}

int phm_simt_fwd(const float* x, const float* A, const float* W, const float* bias, const float* residual, float* y, int rows,
                 int in_features, int out_features, int phm_dim, int act, cudaStream_t stream) {
  const int n = phm_dim, K = in_features / n, P = out_features / n;
  dim3 grid(phc_div_up(out_features, TN), phc_div_up(rows, TM));
  phc_launch(phm_gemm_kernel<false>, dim3(grid), dim3(256), sizeof(float) * n * n * n, stream, x, A, W, bias, residual, y, rows, in_features, out_features, n, K,
                                                                           P, act);
  return phc_check_launch("phc_phm_linear_fwd(simt)");
}

int phm_simt_bwd(const float* gy, const float* x, const float* A, const float* W, float* dx, float* dA, float* dW, float* db, int rows,
                 int in_features, int out_features, int phm_dim, void* workspace, cudaStream_t stream) {
  const int n = phm_dim, K = in_features / n, P = out_features / n, M = rows;
  const int n3 = n * n * n;
  if (dx) {
    dim3 grid(phc_div_up(in_features, TN), phc_div_up(M, TM));
    phc_launch(phm_gemm_kernel<true>, dim3(grid), dim3(256), sizeof(float) * n3, stream, gy, A, W, nullptr, nullptr, dx, M, out_features, in_features, n, K, P,
                                                                      PHC_ACT_IDENTITY);
  }
  const int splits = dh_splits(M, in_features, out_features);
  const int rps = phc_div_up(phc_div_up(M > 0 ? M : 1, splits), 16) * 16;
  float* part = reinterpret_cast<float*>(workspace);
  dim3 g2(phc_div_up(out_features, 64), phc_div_up(in_features, 64), splits);
  phc_launch(phm_dh_kernel, dim3(g2), dim3(256), 0, stream, x, gy, M, in_features, out_features, rps, part);
  return phm_contract_and_bias(part, splits, gy, A, W, dA, dW, db, M, in_features, out_features, n,
                               part + (size_t)splits * in_features * out_features, 0, stream);
}
