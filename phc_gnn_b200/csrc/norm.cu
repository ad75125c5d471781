// Per-component hypercomplex batch-norm fused with activation, (hypercomplex) dropout and the
// skip-connection add:    y = skip + dropout( act( gamma * (h - mean) * rstd + beta ) )
//
// Replaces, per layer, the reference chain  PHMNorm (n x BatchNorm1d on column blocks + permute +
// cat; phc/hypercomplex/norm.py:30-35) -> activation -> phm_dropout (phc/hypercomplex/layers.py:31-55)
// -> "x + tmp[1]" (phc/hypercomplex/undirectional/models.py:206-215).  n independent BatchNorm1d's on
// contiguous column blocks are exactly one per-column batch-norm over the flat [M,F] matrix, so the
// n components are handled by one kernel over flat columns.
//
// Statistics: shifted-data chunk moments (chunk mean, M2) merged in row order with Chan's update in
// double precision -> deterministic and free of the E[x^2]-E[x]^2 cancellation.
// Dropout masks are a pure function of (seed, element index) (Philox4x32-10), so backward regenerates
// them instead of storing an [M,F] mask.
// Roofline: HBM.  Algorithmic bytes fwd (training) = 4F(2M [stats+apply read] + M [skip] + M [write]).
#include "common.cuh"

namespace {

// Work decomposition shared by every kernel of this file: a block is 32 column lanes (x VEC columns each) by 8 row
// slices; blockIdx.x = row chunk, blockIdx.y = column group.  A thread keeps its per-column parameters in registers
// and walks its rows four at a time with all loads of the four rows issued before any arithmetic (the kernels are
// latency-bound otherwise: measured 1.5 TB/s with two loads in flight per thread against 4+ TB/s with eight).
// The chunk height is chosen on the host so that all blocks of a launch are co-resident in one balanced wave.
constexpr int BN_SLICES = 8;   // row slices per block
constexpr int BN_LANES = 32;   // column lanes per block
constexpr int BN_BATCH = 4;    // rows in flight per thread
constexpr int BN_THREADS = BN_SLICES * BN_LANES;
constexpr int BN_MIN_ROWS = 32;  // smallest chunk height (bounds the workspace: 2 * ceil(M/32) * F floats)

struct EwParams {
  int M, F, Fc;            // rows, width, width per component
  int rpc;                 // rows per chunk (multiple of BN_SLICES * BN_BATCH)
  int use_bn, act;
  int drop_on, drop_same;  // dropout active / one mask shared by the n components
  unsigned int keep_thr;   // keep iff philox < keep_thr
  float drop_scale;        // 1/(1-p)
  unsigned long long seed;
  const unsigned long long* epoch;   // optional device word added (times an odd constant) to the seed at run time: CUDA-graph replays
  int ldy;                 // row stride (floats) of the forward OUTPUT and of the incoming gradient in backward: F, or the width of the
                           // [M, F + F0] buffer whose left column block they are (PHMSkipConnectConcat: no torch.cat)
};

// registered by phc_dropout_epoch_register (process-global, like a default generator)
static unsigned long long* g_dropout_epoch = nullptr;

__device__ __forceinline__ void fold_epoch(EwParams& p) {
  if (p.drop_on && p.epoch != nullptr) p.seed += *p.epoch * 0x9E3779B97F4A7C15ULL;
}

__device__ __forceinline__ bool drop_keep_bit(const EwParams& p, int row, int f) {
  unsigned long long idx = p.drop_same ? (unsigned long long)row * p.Fc + (f % p.Fc) : (unsigned long long)row * p.F + f;
  return dropout_keep(p.seed, idx, p.keep_thr);
}

// keep bits of VEC consecutive features starting at f (f % 4 == 0 when VEC == 4): one Philox call yields the four
// uniforms of the aligned element quad, exactly the values drop_factor() produces one by one.  Computed for the
// whole row batch while its loads are in flight, so the generator's temporaries are dead when the arithmetic starts.
template <int VEC, bool DROP>
__device__ __forceinline__ uint32_t keep_bits(const EwParams& p, int row, int f) {
  if (!DROP) return 0xFu;
  uint32_t bits = 0;
  if (VEC == 4 && !p.drop_same && (p.F & 3) == 0) {
    const unsigned long long idx = (unsigned long long)row * p.F + f;      // multiple of 4
    uint32_t r[4];
    philox4x32_10((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
#pragma unroll
    for (int q = 0; q < VEC; ++q) bits |= (r[q] < p.keep_thr ? 1u : 0u) << q;
    return bits;
  }
#pragma unroll
  for (int q = 0; q < VEC; ++q) bits |= (drop_keep_bit(p, row, f + q) ? 1u : 0u) << q;
  return bits;
}
template <int VEC, bool DROP>
__device__ __forceinline__ uint32_t keep_bits_batch(const EwParams& p, int rb, int r1, int f) {
  uint32_t m = 0;
  if (DROP) {
#pragma unroll
    for (int j = 0; j < BN_BATCH; ++j) {
      const int r = rb + j * BN_SLICES;
      if (r < r1) m |= keep_bits<VEC, DROP>(p, r, f) << (4 * j);
    }
  }
  return m;
}
template <bool DROP>
__device__ __forceinline__ float drop_of(const EwParams& p, uint32_t mask, int j, int q) {
  if (!DROP) return 1.f;
  return (mask >> (4 * j + q)) & 1u ? p.drop_scale : 0.f;
}

// ACT >= 0: activation fixed at compile time (identity / relu, the common cases); ACT < 0: runtime switch
template <int ACT> __device__ __forceinline__ float actf(int rt, float v) {
  if constexpr (ACT >= 0) return act_fwd<ACT>(v);
  else return act_fwd_rt(rt, v);
}
template <int ACT> __device__ __forceinline__ float actb(int rt, float v) {
  if constexpr (ACT >= 0) return act_bwd<ACT>(v);
  else return act_bwd_rt(rt, v);
}

// per-column batch-norm parameters of a thread's VEC columns; without batch-norm (or without affine) the neutral
// values make  pre = (h - 0) * 1 * 1 + 0 = h  exactly, so one code path serves every configuration
template <int VEC>
struct Cols {
  float mean[VEC], rstd[VEC], gamma[VEC], beta[VEC];
  __device__ __forceinline__ void load(const EwParams& p, int f, const float* __restrict__ g, const float* __restrict__ b,
                                       const float* __restrict__ m, const float* __restrict__ r) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      mean[q] = p.use_bn ? __ldg(m + f + q) : 0.f;
      rstd[q] = p.use_bn ? __ldg(r + f + q) : 1.f;
      gamma[q] = (p.use_bn && g) ? __ldg(g + f + q) : 1.f;
      beta[q] = (p.use_bn && g) ? __ldg(b + f + q) : 0.f;
    }
  }
};

// reduce the 8 row slices of a block: red[which][slice][col] -> slice 0
template <int VEC>
__device__ __forceinline__ void slice_reduce(float (&red)[2][BN_SLICES][BN_LANES * VEC], int slice, int lane) {
  __syncthreads();
  for (int s = BN_SLICES / 2; s > 0; s >>= 1) {
    if (slice < s) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        red[0][slice][lane * VEC + q] += red[0][slice + s][lane * VEC + q];
        red[1][slice][lane * VEC + q] += red[1][slice + s][lane * VEC + q];
      }
    }
    __syncthreads();
  }
}

// ---- chunk statistics: part[chunk][0][f] = chunk mean, part[chunk][1][f] = chunk M2 ----------
template <int VEC>
__global__ void __launch_bounds__(BN_THREADS, 4) bn_chunk_stats_kernel(const float* __restrict__ h, int M, int F, int rpc,
                                                                       float* __restrict__ part) {
  pdl_begin();
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.y * BN_LANES + lane) * VEC;
  const int r0 = blockIdx.x * rpc;
  const int r1 = min(r0 + rpc, M);
  const bool active = f < F;
  float sh[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { sh[q] = 0.f; s1[q] = 0.f; s2[q] = 0.f; }
  if (active) {
    Vec<VEC> k = Vec<VEC>::load(h + (size_t)r0 * F + f);   // shift = first row of the chunk
#pragma unroll
    for (int q = 0; q < VEC; ++q) sh[q] = k.v[q];
    for (int rb = r0 + slice; rb < r1; rb += BN_SLICES * BN_BATCH) {
      Vec<VEC> v[BN_BATCH];
#pragma unroll
      for (int j = 0; j < BN_BATCH; ++j) {
        const int r = rb + j * BN_SLICES;
        if (r < r1) v[j] = Vec<VEC>::load(h + (size_t)r * F + f);
        else v[j] = k;                                       // contributes d = 0
      }
#pragma unroll
      for (int j = 0; j < BN_BATCH; ++j)
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          float d = v[j].v[q] - sh[q];
          s1[q] += d;
          s2[q] += d * d;
        }
    }
  }
  __shared__ float red[2][BN_SLICES][BN_LANES * VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { red[0][slice][lane * VEC + q] = s1[q]; red[1][slice][lane * VEC + q] = s2[q]; }
  slice_reduce<VEC>(red, slice, lane);
  if (slice == 0 && active) {
    const float cnt = (float)(r1 - r0);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float a = red[0][0][lane * VEC + q], b = red[1][0][lane * VEC + q];
      float mean = sh[q] + a / cnt;
      float m2 = fmaxf(b - a * a / cnt, 0.f);
      part[((size_t)blockIdx.x * 2 + 0) * F + f + q] = mean;
      part[((size_t)blockIdx.x * 2 + 1) * F + f + q] = m2;
    }
  }
}

// ---- finalize: merge the chunk moments, produce mean / rstd, update running statistics ----------
// One warp per feature, two passes over the (L2-resident) partials, all in double and in a fixed order:
//   mean = sum_c n_c mean_c / M ;   M2 = sum_c [ M2_c + n_c (mean_c - mean)^2 ]      (Chan's pairwise update, unrolled:
// every term is non-negative, so there is no cancellation).  Lane l takes chunks l, l+32, ...; the 32 lane sums are
// combined by a fixed butterfly — deterministic.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Merge of the chunk moments.  A block owns BNF_FL adjacent features and splits the chunks over BNF_CS slices: a warp reads four
// chunk rows x eight features = four fully used 32-byte sectors per request (one block per feature read 4 bytes of every sector it
// touched), and with 128 slices the 488 chunk rows of the ppa shape are one round of four independent loads per thread and pass.
// Two passes in double as before (mean, then M2 about that mean); slices are combined in a fixed order (xor-shuffles inside the
// warp, warps in index order), so the result is reproducible.
constexpr int BNF_FL = 8, BNF_CS = 128, BNF_WARPS = BNF_FL * BNF_CS / 32;
__device__ __forceinline__ double bnf_block_sum(double v, double (&sh)[BNF_WARPS][BNF_FL], int fl) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);            // the four slices of this warp (lane = slice-in-warp * 8 + feature lane)
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                                    // sh may still be read from the previous use
  if (lane < BNF_FL) sh[warp][lane] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll 8
  for (int w = 0; w < BNF_WARPS; ++w) t += sh[w][fl];
  return t;
}
__global__ void __launch_bounds__(BNF_FL * BNF_CS) bn_finalize_kernel(const float* __restrict__ part, int chunks, int rpc, int M, int F, float eps,
                                                                      float momentum, float* __restrict__ running_mean,
                                                                      float* __restrict__ running_var, float* __restrict__ save_mean,
                                                                      float* __restrict__ save_rstd, long long* __restrict__ tracked, int n_tracked) {
  pdl_begin();
  __shared__ double sh[BNF_WARPS][BNF_FL];
  const int fl = threadIdx.x & (BNF_FL - 1), cs = threadIdx.x / BNF_FL;
  const int f = blockIdx.x * BNF_FL + fl;
  const bool on = f < F;
  // num_batches_tracked counters: one thread each, the load up here and the store at the very end (four serial read-modify-writes
  // on one thread used to hold the whole block back by four memory latencies)
  const bool tracker = blockIdx.x == 0 && tracked != nullptr && (int)threadIdx.x >= BNF_FL * BNF_CS - n_tracked;
  long long* tslot = tracker ? tracked + (BNF_FL * BNF_CS - 1 - (int)threadIdx.x) : nullptr;
  const long long tval = tracker ? *tslot : 0;
  // running statistics: loaded up front, their latency overlaps the merge
  const bool writer = cs == 0 && on;
  const float rm_old = (writer && running_mean) ? running_mean[f] : 0.f;
  const float rv_old = (writer && running_var) ? running_var[f] : 0.f;
  double s = 0.0, mean, m2 = 0.0;
  if (chunks <= BNF_CS * 4) {
    // common case (<= 512 chunks: 16k rows of 32): every partial of this thread is loaded once, in ONE round of independent loads,
    // and both passes run on registers
    float mc[4], qc[4];
    int nc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = cs + u * BNF_CS;
      const bool valid = on && c < chunks;
      mc[u] = valid ? part[((size_t)c * 2 + 0) * F + f] : 0.f;
      qc[u] = valid ? part[((size_t)c * 2 + 1) * F + f] : 0.f;
      nc[u] = valid ? min(rpc, M - c * rpc) : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) s += (double)nc[u] * (double)mc[u];
    mean = bnf_block_sum(s, sh, fl) / (double)M;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double d = (double)mc[u] - mean;
      if (nc[u] > 0) m2 += (double)qc[u] + (double)nc[u] * d * d;
    }
  } else {
    if (on) {
#pragma unroll 4
      for (int c = cs; c < chunks; c += BNF_CS) s += (double)min(rpc, M - c * rpc) * (double)part[((size_t)c * 2 + 0) * F + f];
    }
    mean = bnf_block_sum(s, sh, fl) / (double)M;
    if (on) {
#pragma unroll 4
      for (int c = cs; c < chunks; c += BNF_CS) {
        const double d = (double)part[((size_t)c * 2 + 0) * F + f] - mean;
        m2 += (double)part[((size_t)c * 2 + 1) * F + f] + (double)min(rpc, M - c * rpc) * d * d;
      }
    }
  }
  m2 = bnf_block_sum(m2, sh, fl);
  if (tracker) *tslot = tval + 1;
  if (cs != 0 || !on) return;
  const double var = m2 / (double)M;
  save_mean[f] = (float)mean;
  save_rstd[f] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[f] = (1.f - momentum) * rm_old + momentum * (float)mean;
  if (running_var) running_var[f] = (1.f - momentum) * rv_old + momentum * (float)(m2 / (double)(M - 1));
}

__global__ void __launch_bounds__(128) bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                                            int F, float eps, float* __restrict__ save_mean,
                                                            float* __restrict__ save_rstd) {
  pdl_begin();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  save_mean[f] = running_mean[f];
  save_rstd[f] = 1.f / sqrtf(running_var[f] + eps);
}

// ---- apply -------------------------------------------------------------------------------------
template <int VEC, int ACT, bool DROP>
__global__ void __launch_bounds__(BN_THREADS, 3) bn_apply_fwd_kernel(EwParams p, const float* __restrict__ h, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, const float* __restrict__ skip,
                                                                     float* __restrict__ y) {
  pdl_begin();
  fold_epoch(p);
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.y * BN_LANES + lane) * VEC;
  if (f >= p.F) return;
  const int r0 = blockIdx.x * p.rpc, r1 = min(r0 + p.rpc, p.M);
  Cols<VEC> c;
  c.load(p, f, gamma, beta, mean, rstd);
  for (int rb = r0 + slice; rb < r1; rb += BN_SLICES * BN_BATCH) {
    Vec<VEC> v[BN_BATCH], s[BN_BATCH];
#pragma unroll
    for (int j = 0; j < BN_BATCH; ++j) {
      const int r = rb + j * BN_SLICES;
      if (r < r1) {
        v[j] = Vec<VEC>::load(h + (size_t)r * p.F + f);
        if (skip) s[j] = Vec<VEC>::load(skip + (size_t)r * p.F + f);
      }
    }
    const uint32_t keep = keep_bits_batch<VEC, DROP>(p, rb, r1, f);
#pragma unroll
    for (int j = 0; j < BN_BATCH; ++j) {
      const int r = rb + j * BN_SLICES;
      if (r >= r1) break;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float a = (v[j].v[q] - c.mean[q]) * c.rstd[q];
        a = a * c.gamma[q] + c.beta[q];
        a = actf<ACT>(p.act, a) * drop_of<DROP>(p, keep, j, q);
        v[j].v[q] = skip ? a + s[j].v[q] : a;
      }
      v[j].store(y + (size_t)r * p.ldy + f);
    }
  }
}

// d(act input) and xhat for one element, shared by the reduction and the apply pass of backward
template <int ACT>
__device__ __forceinline__ float bn_da(int act, float dy, float hv, float dfac, float mean, float rstd, float gamma, float beta, float* xhat_out) {
  const float xh = (hv - mean) * rstd;
  *xhat_out = xh;
  return dy * dfac * actb<ACT>(act, xh * gamma + beta);
}

// part[chunk][0][f] = sum da ; part[chunk][1][f] = sum da * xhat
template <int VEC, int ACT, bool DROP>
__global__ void __launch_bounds__(BN_THREADS, 3) bn_bwd_reduce_kernel(EwParams p, const float* __restrict__ dy, const float* __restrict__ h,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      float* __restrict__ part) {
  pdl_begin();
  fold_epoch(p);
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.y * BN_LANES + lane) * VEC;
  const int r0 = blockIdx.x * p.rpc, r1 = min(r0 + p.rpc, p.M);
  const bool active = f < p.F;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
  if (active) {
    Cols<VEC> c;
    c.load(p, f, gamma, beta, mean, rstd);
    for (int rb = r0 + slice; rb < r1; rb += BN_SLICES * BN_BATCH) {
      Vec<VEC> g[BN_BATCH], v[BN_BATCH];
#pragma unroll
      for (int j = 0; j < BN_BATCH; ++j) {
        const int r = rb + j * BN_SLICES;
        if (r < r1) {
          g[j] = Vec<VEC>::load(dy + (size_t)r * p.ldy + f);
          v[j] = Vec<VEC>::load(h + (size_t)r * p.F + f);
        }
      }
      const uint32_t keep = keep_bits_batch<VEC, DROP>(p, rb, r1, f);
#pragma unroll
      for (int j = 0; j < BN_BATCH; ++j) {
        const int r = rb + j * BN_SLICES;
        if (r >= r1) break;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          float xh;
          const float da = bn_da<ACT>(p.act, g[j].v[q], v[j].v[q], drop_of<DROP>(p, keep, j, q), c.mean[q], c.rstd[q], c.gamma[q], c.beta[q], &xh);
          s1[q] += da;
          s2[q] += da * xh;
        }
      }
    }
  }
  __shared__ float red[2][BN_SLICES][BN_LANES * VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { red[0][slice][lane * VEC + q] = s1[q]; red[1][slice][lane * VEC + q] = s2[q]; }
  slice_reduce<VEC>(red, slice, lane);
  if (slice == 0 && active) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      part[((size_t)blockIdx.x * 2 + 0) * p.F + f + q] = red[0][0][lane * VEC + q];
      part[((size_t)blockIdx.x * 2 + 1) * p.F + f + q] = red[1][0][lane * VEC + q];
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ part, int chunks, int F, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
  pdl_begin();
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  double a = 0.0, b = 0.0;
#pragma unroll 4
  for (int c = lane; c < chunks; c += 32) {
    a += part[((size_t)c * 2 + 0) * F + f];
    b += part[((size_t)c * 2 + 1) * F + f];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    dbeta[f] = (float)a;
    dgamma[f] = (float)b;
  }
}

// dh = gamma * rstd * (da - sum_da/M - xhat * sum_da_xhat/M)   [training]   or gamma * rstd * da   [eval]
template <int VEC, int ACT, bool DROP>
__global__ void __launch_bounds__(BN_THREADS, 3) bn_apply_bwd_kernel(EwParams p, int training, const float* __restrict__ dy,
                                                                     const float* __restrict__ h, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, const float* __restrict__ sum_da,
                                                                     const float* __restrict__ sum_da_xhat, float* __restrict__ dh) {
  pdl_begin();
  fold_epoch(p);
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.y * BN_LANES + lane) * VEC;
  if (f >= p.F) return;
  const int r0 = blockIdx.x * p.rpc, r1 = min(r0 + p.rpc, p.M);
  Cols<VEC> c;
  c.load(p, f, gamma, beta, mean, rstd);
  const float invM = 1.f / (float)p.M;
  float scq[VEC], c1[VEC], c2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    scq[q] = c.rstd[q] * c.gamma[q];
    c1[q] = 0.f; c2[q] = 0.f;
    if (p.use_bn && training) { c1[q] = __ldg(sum_da + f + q) * invM; c2[q] = __ldg(sum_da_xhat + f + q) * invM; }
  }
  for (int rb = r0 + slice; rb < r1; rb += BN_SLICES * BN_BATCH) {
    Vec<VEC> g[BN_BATCH], v[BN_BATCH];
#pragma unroll
    for (int j = 0; j < BN_BATCH; ++j) {
      const int r = rb + j * BN_SLICES;
      if (r < r1) {
        g[j] = Vec<VEC>::load(dy + (size_t)r * p.ldy + f);
        v[j] = Vec<VEC>::load(h + (size_t)r * p.F + f);
      }
    }
    const uint32_t keep = keep_bits_batch<VEC, DROP>(p, rb, r1, f);
#pragma unroll
    for (int j = 0; j < BN_BATCH; ++j) {
      const int r = rb + j * BN_SLICES;
      if (r >= r1) break;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float xh;
        float da = bn_da<ACT>(p.act, g[j].v[q], v[j].v[q], drop_of<DROP>(p, keep, j, q), c.mean[q], c.rstd[q], c.gamma[q], c.beta[q], &xh);
        if (p.use_bn) da = (da - c1[q] - xh * c2[q]) * scq[q];
        v[j].v[q] = da;
      }
      v[j].store(dh + (size_t)r * p.F + f);
    }
  }
}

int bn_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    else sms = 148;
  }
  return sms;
}

// chunk height so that (column groups) x (row chunks) blocks fill the GPU once, `per_sm` blocks per SM
int bn_rows_per_chunk(int M, int colgroups, int per_sm) {
  const int unit = BN_SLICES * BN_BATCH;
  int want = (bn_num_sms() * per_sm) / (colgroups > 0 ? colgroups : 1);
  if (want < 1) want = 1;
  long long rpc = ((long long)M + want - 1) / want;
  rpc = (rpc + unit - 1) / unit * unit;
  if (rpc < BN_MIN_ROWS) rpc = BN_MIN_ROWS;
  return (int)rpc;
}

EwParams make_params(int M, int F, int n, int use_bn, int act, float drop_p, int drop_same, int training, unsigned long long seed) {
  EwParams p;
  p.M = M; p.F = F; p.Fc = F / n; p.use_bn = use_bn; p.act = act; p.rpc = BN_MIN_ROWS;
  p.drop_on = (training && drop_p > 0.f) ? 1 : 0;
  p.drop_same = drop_same;
  double keep = 1.0 - (double)drop_p;
  double thr = keep * 4294967296.0;
  p.keep_thr = thr >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned int)thr;
  p.drop_scale = keep > 0.0 ? (float)(1.0 / keep) : 0.f;
  p.seed = seed;
  p.epoch = g_dropout_epoch;
  p.ldy = F;
  return p;
}

// compile-time activation for the common cases, runtime switch (-1) otherwise
#define BN_DISPATCH(KERNEL, V4, P, GRID, STREAM, ...)                                                                   \
  do {                                                                                                                  \
    const int a_ = (P).act == PHC_ACT_IDENTITY ? 0 : ((P).act == PHC_ACT_RELU ? 1 : 2);                                  \
    const int key_ = ((V4) ? 6 : 0) + a_ * 2 + ((P).drop_on ? 1 : 0);                                                    \
    switch (key_) {                                                                                                     \
      case 0: phc_launch(KERNEL<1, PHC_ACT_IDENTITY, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                  \
      case 1: phc_launch(KERNEL<1, PHC_ACT_IDENTITY, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                   \
      case 2: phc_launch(KERNEL<1, PHC_ACT_RELU, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                      \
      case 3: phc_launch(KERNEL<1, PHC_ACT_RELU, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                       \
      case 4: phc_launch(KERNEL<1, -1, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                                \
      case 5: phc_launch(KERNEL<1, -1, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                                 \
      case 6: phc_launch(KERNEL<4, PHC_ACT_IDENTITY, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                  \
      case 7: phc_launch(KERNEL<4, PHC_ACT_IDENTITY, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                   \
      case 8: phc_launch(KERNEL<4, PHC_ACT_RELU, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                      \
      case 9: phc_launch(KERNEL<4, PHC_ACT_RELU, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                       \
      case 10: phc_launch(KERNEL<4, -1, false>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                               \
      default: phc_launch(KERNEL<4, -1, true>, dim3(GRID), dim3(BN_THREADS), 0, STREAM, __VA_ARGS__); break;                                \
    }                                                                                                                   \
  } while (0)

}  // namespace

extern "C" {

size_t phc_bn_workspace_bytes(int rows, int width) { return sizeof(float) * 2 * ((size_t)phc_div_up(rows, BN_MIN_ROWS) + 1) * width + 16; }

static int bn_fwd_impl(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                       int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                       unsigned long long seed, float* y, float* save_mean, float* save_rstd, void* workspace,
                       size_t workspace_bytes, const float* pre_partials, int pre_chunk_rows, cudaStream_t stream, int ldy = 0);

int phc_bn_act_drop_skip_fwd(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                             int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                             unsigned long long seed, float* y, float* save_mean, float* save_rstd, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  return bn_fwd_impl(h, gamma, beta, running_mean, running_var, num_batches_tracked, n_tracked, skip, rows, width, phm_dim, use_bn, training,
                     momentum, eps, act, drop_p, drop_same, seed, y, save_mean, save_rstd, workspace, workspace_bytes, nullptr, 0, stream);
}

int phc_bn_act_drop_skip_fwd_strided(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                                     int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                                     unsigned long long seed, float* y, int y_row_stride, float* save_mean, float* save_rstd,
                                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return bn_fwd_impl(h, gamma, beta, running_mean, running_var, num_batches_tracked, n_tracked, skip, rows, width, phm_dim, use_bn, training,
                     momentum, eps, act, drop_p, drop_same, seed, y, save_mean, save_rstd, workspace, workspace_bytes, nullptr, 0, stream,
                     y_row_stride);
}

int phc_bn_act_drop_skip_fwd_partials(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                      long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                                      int training, float momentum, float eps, int act, float drop_p, int drop_same,
                                      unsigned long long seed, float* y, float* save_mean, float* save_rstd, const float* partials,
                                      int chunk_rows, cudaStream_t stream) {
  PHC_REQUIRE(partials != nullptr && chunk_rows > 0, "phc_bn_act_drop_skip_fwd_partials: partials / chunk_rows required");
  PHC_REQUIRE(training, "phc_bn_act_drop_skip_fwd_partials: batch statistics are only used in training mode");
  return bn_fwd_impl(h, gamma, beta, running_mean, running_var, num_batches_tracked, n_tracked, skip, rows, width, phm_dim, 1, 1, momentum,
                     eps, act, drop_p, drop_same, seed, y, save_mean, save_rstd, nullptr, 0, partials, chunk_rows, stream);
}

}  // extern "C"

static int bn_fwd_impl(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                       int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                       unsigned long long seed, float* y, float* save_mean, float* save_rstd, void* workspace,
                       size_t workspace_bytes, const float* pre_partials, int pre_chunk_rows, cudaStream_t stream, int ldy) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_bn_act_drop_skip_fwd: width %d not divisible by phm_dim %d", width, phm_dim);
  PHC_REQUIRE(ldy == 0 || ldy >= width, "phc_bn_act_drop_skip_fwd: output row stride %d smaller than the width %d", ldy, width);
  PHC_REQUIRE(act >= PHC_ACT_IDENTITY && act <= PHC_ACT_SWISH, "phc_bn_act_drop_skip_fwd: bad act %d", act);
  PHC_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, "phc_bn_act_drop_skip_fwd: dropout rate %f outside [0,1]", drop_p);
  PHC_REQUIRE((gamma == nullptr) == (beta == nullptr), "phc_bn_act_drop_skip_fwd: gamma/beta must both be given or both null");
  if (rows == 0) return PHC_OK;
  const int M = rows, F = width;
  if (use_bn) {
    PHC_REQUIRE(save_mean && save_rstd, "phc_bn_act_drop_skip_fwd: save_mean/save_rstd required with batch-norm");
    if (training && pre_partials != nullptr) {
      // chunk moments already produced by the kernel that wrote h (PHMLinear epilogue): only merge them
      PHC_REQUIRE(M > 1, "phc_bn_act_drop_skip_fwd: batch-norm in training mode needs more than 1 row");
      phc_launch(bn_finalize_kernel, dim3(phc_div_up(F, BNF_FL)), dim3(BNF_FL * BNF_CS), 0, stream, pre_partials,
                 phc_div_up(M, pre_chunk_rows), pre_chunk_rows, M, F, eps, momentum, running_mean, running_var, save_mean, save_rstd,
                 num_batches_tracked, n_tracked);
    } else if (training) {
      PHC_REQUIRE(M > 1, "phc_bn_act_drop_skip_fwd: batch-norm in training mode needs more than 1 row");
      PHC_REQUIRE(workspace_bytes >= phc_bn_workspace_bytes(M, F), "phc_bn_act_drop_skip_fwd: workspace too small");
      float* part = reinterpret_cast<float*>(workspace);
      const bool v4s = F % 4 == 0 && phc_aligned16(h);
      const int colgroups = phc_div_up(F, BN_LANES * (v4s ? 4 : 1));
      const int rpc = bn_rows_per_chunk(M, colgroups, 4);
      const int chunks = phc_div_up(M, rpc);
      dim3 grid(chunks, colgroups);
      if (v4s) phc_launch(bn_chunk_stats_kernel<4>, dim3(grid), dim3(BN_THREADS), 0, stream, h, M, F, rpc, part);
      else phc_launch(bn_chunk_stats_kernel<1>, dim3(grid), dim3(BN_THREADS), 0, stream, h, M, F, rpc, part);
      phc_launch(bn_finalize_kernel, dim3(phc_div_up(F, BNF_FL)), dim3(BNF_FL * BNF_CS), 0, stream, part, chunks, rpc, M, F, eps, momentum, running_mean,
                                                                                running_var, save_mean, save_rstd, num_batches_tracked,
                                                                                n_tracked);
    } else {
      PHC_REQUIRE(running_mean && running_var, "phc_bn_act_drop_skip_fwd: eval mode needs running statistics");
      phc_launch(bn_eval_stats_kernel, dim3(phc_div_up(F, 128)), dim3(128), 0, stream, running_mean, running_var, F, eps, save_mean, save_rstd);
    }
  }
  EwParams p = make_params(M, F, phm_dim, use_bn, act, drop_p, drop_same, training, seed);
  if (ldy > 0) p.ldy = ldy;
  const bool v4 = F % 4 == 0 && p.ldy % 4 == 0 && phc_aligned16(h) && phc_aligned16(y) && phc_aligned16(skip);
  const int colgroups = phc_div_up(F, BN_LANES * (v4 ? 4 : 1));
  p.rpc = bn_rows_per_chunk(M, colgroups, 3);
  dim3 grid(phc_div_up(M, p.rpc), colgroups);
  BN_DISPATCH(bn_apply_fwd_kernel, v4, p, grid, stream, p, h, gamma, beta, save_mean, save_rstd, skip, y);
  return phc_check_launch("phc_bn_act_drop_skip_fwd");
}

extern "C" {

int phc_bn_act_drop_skip_bwd_strided(const float* dy, int dy_row_stride, const float* h, const float* gamma, const float* beta,
                                     const float* save_mean, const float* save_rstd, int rows, int width, int phm_dim, int use_bn,
                                     int training, int act, float drop_p, int drop_same, unsigned long long seed, float* dh,
                                     float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, cudaStream_t stream);

int phc_bn_act_drop_skip_bwd(const float* dy, const float* h, const float* gamma, const float* beta, const float* save_mean,
                             const float* save_rstd, int rows, int width, int phm_dim, int use_bn, int training, int act, float drop_p,
                             int drop_same, unsigned long long seed, float* dh, float* dgamma, float* dbeta, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  return phc_bn_act_drop_skip_bwd_strided(dy, width, h, gamma, beta, save_mean, save_rstd, rows, width, phm_dim, use_bn, training, act,
                                          drop_p, drop_same, seed, dh, dgamma, dbeta, workspace, workspace_bytes, stream);
}

int phc_bn_act_drop_skip_bwd_strided(const float* dy, int dy_row_stride, const float* h, const float* gamma, const float* beta,
                                     const float* save_mean, const float* save_rstd, int rows, int width, int phm_dim, int use_bn,
                                     int training, int act, float drop_p, int drop_same, unsigned long long seed, float* dh,
                                     float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_bn_act_drop_skip_bwd: width %d not divisible by phm_dim %d", width, phm_dim);
  PHC_REQUIRE(dy_row_stride >= width, "phc_bn_act_drop_skip_bwd: gradient row stride %d smaller than the width %d", dy_row_stride, width);
  if (rows == 0) return PHC_OK;
  const int M = rows, F = width;
  EwParams p = make_params(M, F, phm_dim, use_bn, act, drop_p, drop_same, training, seed);
  p.ldy = dy_row_stride;
  const bool v4 = F % 4 == 0 && dy_row_stride % 4 == 0 && phc_aligned16(h) && phc_aligned16(dy) && phc_aligned16(dh);
  const int colgroups = phc_div_up(F, BN_LANES * (v4 ? 4 : 1));
  float* sum_da = dbeta;
  float* sum_da_xhat = dgamma;
  if (use_bn) {
    PHC_REQUIRE(dgamma && dbeta, "phc_bn_act_drop_skip_bwd: dgamma/dbeta buffers required with batch-norm");
    PHC_REQUIRE(workspace_bytes >= phc_bn_workspace_bytes(M, F), "phc_bn_act_drop_skip_bwd: workspace too small");
    float* part = reinterpret_cast<float*>(workspace);
    p.rpc = bn_rows_per_chunk(M, colgroups, 3);
    const int chunks = phc_div_up(M, p.rpc);
    dim3 grid(chunks, colgroups);
    BN_DISPATCH(bn_bwd_reduce_kernel, v4, p, grid, stream, p, dy, h, gamma, beta, save_mean, save_rstd, part);
    phc_launch(bn_bwd_finalize_kernel, dim3(phc_div_up((long long)F * 32, 256)), dim3(256), 0, stream, part, chunks, F, dgamma, dbeta);
  }
  p.rpc = bn_rows_per_chunk(M, colgroups, 3);
  dim3 grid(phc_div_up(M, p.rpc), colgroups);
  BN_DISPATCH(bn_apply_bwd_kernel, v4, p, grid, stream, p, training, dy, h, gamma, beta, save_mean, save_rstd, sum_da, sum_da_xhat, dh);
  return phc_check_launch("phc_bn_act_drop_skip_bwd");
}

// ---- out = sum of `count` equally shaped tensors, in list order (fixed association) ----------------------------------
// The gradient of a skip connection that fans out to L layers (models.py:227-236, sc_type="first": every layer adds the
// same h0) is the sum of L [N,F] tensors; autograd adds them pairwise (L-1 kernels, 3 tensors of traffic each), this is
// one pass: L reads + 1 write.
#define PHC_SUM_MAX 16
struct SumTable { const float* src[PHC_SUM_MAX]; };
__global__ void __launch_bounds__(256) sum_tensors_kernel(SumTable t, int count, long long n4, long long numel, float* __restrict__ out) {
  pdl_begin();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(t.src[0]) + i);
    for (int k = 1; k < count; ++k) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(t.src[k]) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  } else {
    const long long j = n4 * 4 + (i - n4);                        // scalar tail
    if (j < numel) {
      float a = t.src[0][j];
      for (int k = 1; k < count; ++k) a += t.src[k][j];
      out[j] = a;
    }
  }
}

int phc_sum_tensors(const float* const* srcs, int count, long long numel, float* out, cudaStream_t stream) {
  PHC_REQUIRE(count >= 1 && count <= PHC_SUM_MAX, "phc_sum_tensors: count %d not in 1..%d", count, PHC_SUM_MAX);
  PHC_REQUIRE(numel >= 0 && out != nullptr, "phc_sum_tensors: bad arguments");
  if (numel == 0) return PHC_OK;
  SumTable t;
  bool vec = phc_aligned16(out);
  for (int k = 0; k < count; ++k) {
    PHC_REQUIRE(srcs[k] != nullptr, "phc_sum_tensors: null source %d", k);
    t.src[k] = srcs[k];
    vec = vec && phc_aligned16(srcs[k]);
  }
  const long long n4 = vec ? numel / 4 : 0;
  const long long threads = n4 + (numel - 4 * n4);
  phc_launch(sum_tensors_kernel, dim3(phc_div_up(threads, 256)), dim3(256), 0, stream, t, count, n4, numel, out);
  return phc_check_launch("phc_sum_tensors");
}

}  // extern "C"

namespace {
__global__ void dropout_epoch_advance_kernel(unsigned long long* e) {
  pdl_begin();
  if (threadIdx.x == 0 && blockIdx.x == 0) *e += 1ULL;
}
}  // namespace

extern "C" int phc_dropout_epoch_register(unsigned long long* epoch_dev) {
  g_dropout_epoch = epoch_dev;
  return PHC_OK;
}

extern "C" int phc_dropout_epoch_advance(unsigned long long* epoch_dev, cudaStream_t stream) {
  PHC_REQUIRE(epoch_dev != nullptr, "phc_dropout_epoch_advance: null epoch word");
  phc_launch(dropout_epoch_advance_kernel, dim3(1), dim3(32), 0, stream, epoch_dev);
  return phc_check_launch("phc_dropout_epoch_advance");
}
