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

constexpr int BN_ROWS = 64;    // rows per statistics chunk
constexpr int BN_SLICES = 8;   // row slices per block
constexpr int BN_LANES = 32;   // feature lanes per block

struct EwParams {
  int M, F, Fc;            // rows, width, width per component
  int use_bn, act;
  int drop_on, drop_same;  // dropout active / one mask shared by the n components
  unsigned int keep_thr;   // keep iff philox < keep_thr
  float drop_scale;        // 1/(1-p)
  unsigned long long seed;
};

__device__ __forceinline__ float drop_factor(const EwParams& p, int row, int f) {
  if (!p.drop_on) return 1.f;
  unsigned long long idx = p.drop_same ? (unsigned long long)row * p.Fc + (f % p.Fc) : (unsigned long long)row * p.F + f;
  return dropout_keep(p.seed, idx, p.keep_thr) ? p.drop_scale : 0.f;
}

// dropout factors for VEC consecutive features starting at f (f % 4 == 0 when VEC == 4): one Philox call yields
// the four uniforms of the aligned element quad, exactly the values drop_factor() produces one by one.
template <int VEC>
__device__ __forceinline__ void drop_factors(const EwParams& p, int row, int f, float (&out)[VEC]) {
  if (!p.drop_on) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) out[q] = 1.f;
    return;
  }
  if (VEC == 4 && !p.drop_same && (p.F & 3) == 0) {
    const unsigned long long idx = (unsigned long long)row * p.F + f;      // multiple of 4
    uint32_t r[4];
    philox4x32_10((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
#pragma unroll
    for (int q = 0; q < VEC; ++q) out[q] = r[q] < p.keep_thr ? p.drop_scale : 0.f;
    return;
  }
#pragma unroll
  for (int q = 0; q < VEC; ++q) out[q] = drop_factor(p, row, f + q);
}

// ---- chunk statistics: part[chunk][0][f] = chunk mean, part[chunk][1][f] = chunk M2 ----------
template <int VEC>
__global__ void __launch_bounds__(BN_SLICES * BN_LANES) bn_chunk_stats_kernel(const float* __restrict__ h, int M, int F,
                                                                             float* __restrict__ part) {
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.x * BN_LANES + lane) * VEC;
  const int r0 = blockIdx.y * BN_ROWS;
  const int r1 = min(r0 + BN_ROWS, M);
  const bool active = f < F;
  float sh[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { sh[q] = 0.f; s1[q] = 0.f; s2[q] = 0.f; }
  if (active) {
    Vec<VEC> k = Vec<VEC>::load(h + (size_t)r0 * F + f);   // shift = first row of the chunk
#pragma unroll
    for (int q = 0; q < VEC; ++q) sh[q] = k.v[q];
#pragma unroll 4
    for (int r = r0 + slice; r < r1; r += BN_SLICES) {
      Vec<VEC> v = Vec<VEC>::load(h + (size_t)r * F + f);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float d = v.v[q] - sh[q];
        s1[q] += d;
        s2[q] += d * d;
      }
    }
  }
  __shared__ float red[2][BN_SLICES][BN_LANES * VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { red[0][slice][lane * VEC + q] = s1[q]; red[1][slice][lane * VEC + q] = s2[q]; }
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
  if (slice == 0 && active) {
    const float cnt = (float)(r1 - r0);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float a = red[0][0][lane * VEC + q], b = red[1][0][lane * VEC + q];
      float mean = sh[q] + a / cnt;
      float m2 = fmaxf(b - a * a / cnt, 0.f);
      part[((size_t)blockIdx.y * 2 + 0) * F + f + q] = mean;
      part[((size_t)blockIdx.y * 2 + 1) * F + f + q] = m2;
    }
  }
}

// ---- finalize: merge chunks (Chan), produce mean / rstd, update running statistics ----------
// One warp per feature: lane l merges chunks l, l+32, ... in order, then the 32 lane results are merged by a
// fixed butterfly (xor 16, 8, 4, 2, 1) — a fixed association order, hence deterministic.
__device__ __forceinline__ void chan_merge(double& cnt, double& mean, double& m2, double nb, double mb, double m2b) {
  if (nb == 0.0) return;
  const double tot = cnt + nb, delta = mb - mean;
  mean += delta * nb / tot;
  m2 += m2b + delta * delta * cnt * nb / tot;
  cnt = tot;
}
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ part, int chunks, int M, int F, float eps,
                                                          float momentum, float* __restrict__ running_mean,
                                                          float* __restrict__ running_var, float* __restrict__ save_mean,
                                                          float* __restrict__ save_rstd, long long* __restrict__ tracked, int n_tracked) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && tracked) for (int c = 0; c < n_tracked; ++c) tracked[c] += 1;
  if (f >= F) return;
  double mean = 0.0, m2 = 0.0, cnt = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    const double nb = (double)min(BN_ROWS, M - c * BN_ROWS);
    chan_merge(cnt, mean, m2, nb, (double)part[((size_t)c * 2 + 0) * F + f], (double)part[((size_t)c * 2 + 1) * F + f]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ocnt = __shfl_xor_sync(0xffffffffu, cnt, o), omean = __shfl_xor_sync(0xffffffffu, mean, o),
                 om2 = __shfl_xor_sync(0xffffffffu, m2, o);
    // merge (lower lane's partial first so both partners compute the identical result)
    double c0 = (lane & o) ? ocnt : cnt, me0 = (lane & o) ? omean : mean, q0 = (lane & o) ? om2 : m2;
    const double c1 = (lane & o) ? cnt : ocnt, me1 = (lane & o) ? mean : omean, q1 = (lane & o) ? m2 : om2;
    chan_merge(c0, me0, q0, c1, me1, q1);
    cnt = c0; mean = me0; m2 = q0;
  }
  if (lane != 0) return;
  const double var = m2 / (double)M;
  save_mean[f] = (float)mean;
  save_rstd[f] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * (float)mean;
  if (running_var) running_var[f] = (1.f - momentum) * running_var[f] + momentum * (float)(m2 / (double)(M - 1));
}

__global__ void __launch_bounds__(128) bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                                            int F, float eps, float* __restrict__ save_mean,
                                                            float* __restrict__ save_rstd) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  save_mean[f] = running_mean[f];
  save_rstd[f] = 1.f / sqrtf(running_var[f] + eps);
}

// ---- apply -------------------------------------------------------------------------------------
constexpr int EW_ROWS = 4;   // rows per thread in the apply kernels (independent 128-bit loads in flight)

template <int VEC>
__global__ void __launch_bounds__(256) bn_apply_fwd_kernel(EwParams p, const float* __restrict__ h, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ skip,
                                                           float* __restrict__ y) {
  const int fv = p.F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rgroups = (p.M + EW_ROWS - 1) / EW_ROWS;
  if (t >= (long long)rgroups * fv) return;
  const int rg = (int)(t / fv), f = (int)(t % fv) * VEC;
  Vec<VEC> v[EW_ROWS], s[EW_ROWS];
#pragma unroll
  for (int j = 0; j < EW_ROWS; ++j) {
    const int r = rg * EW_ROWS + j;
    if (r < p.M) {
      v[j] = Vec<VEC>::load(h + (size_t)r * p.F + f);
      if (skip) s[j] = Vec<VEC>::load(skip + (size_t)r * p.F + f);
    }
  }
#pragma unroll
  for (int j = 0; j < EW_ROWS; ++j) {
    const int r = rg * EW_ROWS + j;
    if (r >= p.M) break;
    float df[VEC];
    drop_factors<VEC>(p, r, f, df);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float a = p.use_bn ? (v[j].v[q] - __ldg(mean + f + q)) * __ldg(rstd + f + q) : v[j].v[q];
      if (p.use_bn && gamma) a = a * __ldg(gamma + f + q) + __ldg(beta + f + q);
      a = act_fwd_rt(p.act, a) * df[q];
      v[j].v[q] = skip ? a + s[j].v[q] : a;
    }
    v[j].store(y + (size_t)r * p.F + f);
  }
}

// d(act input) for one element, shared by the reduction and the apply pass of backward
__device__ __forceinline__ float bn_da(const EwParams& p, float dy, float hv, float dfac, int f, const float* gamma, const float* beta,
                                       const float* mean, const float* rstd, float* xhat_out) {
  float xh = hv, pre = hv;
  if (p.use_bn) {
    xh = (hv - __ldg(mean + f)) * __ldg(rstd + f);
    pre = gamma ? xh * __ldg(gamma + f) + __ldg(beta + f) : xh;
  }
  *xhat_out = xh;
  return dy * dfac * act_bwd_rt(p.act, pre);
}

// part[chunk][0][f] = sum da ; part[chunk][1][f] = sum da * xhat
template <int VEC>
__global__ void __launch_bounds__(BN_SLICES * BN_LANES) bn_bwd_reduce_kernel(EwParams p, const float* __restrict__ dy,
                                                                            const float* __restrict__ h, const float* __restrict__ gamma,
                                                                            const float* __restrict__ beta, const float* __restrict__ mean,
                                                                            const float* __restrict__ rstd, float* __restrict__ part) {
  const int lane = threadIdx.x % BN_LANES, slice = threadIdx.x / BN_LANES;
  const int f = (blockIdx.x * BN_LANES + lane) * VEC;
  const int r0 = blockIdx.y * BN_ROWS;
  const int r1 = min(r0 + BN_ROWS, p.M);
  const bool active = f < p.F;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
  if (active) {
#pragma unroll 4
    for (int r = r0 + slice; r < r1; r += BN_SLICES) {
      Vec<VEC> g = Vec<VEC>::load(dy + (size_t)r * p.F + f);
      Vec<VEC> v = Vec<VEC>::load(h + (size_t)r * p.F + f);
      float df[VEC];
      drop_factors<VEC>(p, r, f, df);
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        float xh;
        float da = bn_da(p, g.v[q], v.v[q], df[q], f + q, gamma, beta, mean, rstd, &xh);
        s1[q] += da;
        s2[q] += da * xh;
      }
    }
  }
  __shared__ float red[2][BN_SLICES][BN_LANES * VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { red[0][slice][lane * VEC + q] = s1[q]; red[1][slice][lane * VEC + q] = s2[q]; }
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
  if (slice == 0 && active) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      part[((size_t)blockIdx.y * 2 + 0) * p.F + f + q] = red[0][0][lane * VEC + q];
      part[((size_t)blockIdx.y * 2 + 1) * p.F + f + q] = red[1][0][lane * VEC + q];
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ part, int chunks, int F, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  double a = 0.0, b = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    a += part[((size_t)c * 2 + 0) * F + f];
    b += part[((size_t)c * 2 + 1) * F + f];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    dbeta[f] = (float)a;
    dgamma[f] = (float)b;
  }
}

// dh = gamma * rstd * (da - sum_da/M - xhat * sum_da_xhat/M)   [training]   or gamma * rstd * da   [eval]
template <int VEC>
__global__ void __launch_bounds__(256) bn_apply_bwd_kernel(EwParams p, int training, const float* __restrict__ dy, const float* __restrict__ h,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ sum_da, const float* __restrict__ sum_da_xhat,
                                                           float* __restrict__ dh) {
  const int fv = p.F / VEC;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rgroups = (p.M + EW_ROWS - 1) / EW_ROWS;
  if (t >= (long long)rgroups * fv) return;
  const int rg = (int)(t / fv), f = (int)(t % fv) * VEC;
  const float invM = 1.f / (float)p.M;
  float scq[VEC], c1[VEC], c2[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    scq[q] = 1.f; c1[q] = 0.f; c2[q] = 0.f;
    if (p.use_bn) {
      scq[q] = __ldg(rstd + f + q) * (gamma ? __ldg(gamma + f + q) : 1.f);
      if (training) { c1[q] = __ldg(sum_da + f + q) * invM; c2[q] = __ldg(sum_da_xhat + f + q) * invM; }
    }
  }
  Vec<VEC> g[EW_ROWS], v[EW_ROWS];
#pragma unroll
  for (int j = 0; j < EW_ROWS; ++j) {
    const int r = rg * EW_ROWS + j;
    if (r < p.M) {
      g[j] = Vec<VEC>::load(dy + (size_t)r * p.F + f);
      v[j] = Vec<VEC>::load(h + (size_t)r * p.F + f);
    }
  }
#pragma unroll
  for (int j = 0; j < EW_ROWS; ++j) {
    const int r = rg * EW_ROWS + j;
    if (r >= p.M) break;
    float df[VEC];
    drop_factors<VEC>(p, r, f, df);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float xh;
      float da = bn_da(p, g[j].v[q], v[j].v[q], df[q], f + q, gamma, beta, mean, rstd, &xh);
      if (p.use_bn) da = (da - c1[q] - xh * c2[q]) * scq[q];
      v[j].v[q] = da;
    }
    v[j].store(dh + (size_t)r * p.F + f);
  }
}

EwParams make_params(int M, int F, int n, int use_bn, int act, float drop_p, int drop_same, int training, unsigned long long seed) {
  EwParams p;
  p.M = M; p.F = F; p.Fc = F / n; p.use_bn = use_bn; p.act = act;
  p.drop_on = (training && drop_p > 0.f) ? 1 : 0;
  p.drop_same = drop_same;
  double keep = 1.0 - (double)drop_p;
  double thr = keep * 4294967296.0;
  p.keep_thr = thr >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned int)thr;
  p.drop_scale = keep > 0.0 ? (float)(1.0 / keep) : 0.f;
  p.seed = seed;
  return p;
}

}  // namespace

extern "C" {

size_t phc_bn_workspace_bytes(int rows, int width) { return sizeof(float) * 2 * (size_t)phc_div_up(rows, BN_ROWS) * width + 16; }

int phc_bn_act_drop_skip_fwd(const float* h, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             long long* num_batches_tracked, int n_tracked, const float* skip, int rows, int width, int phm_dim,
                             int use_bn, int training, float momentum, float eps, int act, float drop_p, int drop_same,
                             unsigned long long seed, float* y, float* save_mean, float* save_rstd, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_bn_act_drop_skip_fwd: width %d not divisible by phm_dim %d", width, phm_dim);
  PHC_REQUIRE(act >= PHC_ACT_IDENTITY && act <= PHC_ACT_SWISH, "phc_bn_act_drop_skip_fwd: bad act %d", act);
  PHC_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, "phc_bn_act_drop_skip_fwd: dropout rate %f outside [0,1]", drop_p);
  PHC_REQUIRE((gamma == nullptr) == (beta == nullptr), "phc_bn_act_drop_skip_fwd: gamma/beta must both be given or both null");
  if (rows == 0) return PHC_OK;
  const int M = rows, F = width;
  if (use_bn) {
    PHC_REQUIRE(save_mean && save_rstd, "phc_bn_act_drop_skip_fwd: save_mean/save_rstd required with batch-norm");
    if (training) {
      PHC_REQUIRE(M > 1, "phc_bn_act_drop_skip_fwd: batch-norm in training mode needs more than 1 row");
      PHC_REQUIRE(workspace_bytes >= phc_bn_workspace_bytes(M, F), "phc_bn_act_drop_skip_fwd: workspace too small");
      const int chunks = phc_div_up(M, BN_ROWS);
      float* part = reinterpret_cast<float*>(workspace);
      const bool v4s = F % 4 == 0 && phc_aligned16(h);
      dim3 grid(phc_div_up(F, BN_LANES * (v4s ? 4 : 1)), chunks);
      if (v4s) bn_chunk_stats_kernel<4><<<grid, BN_SLICES * BN_LANES, 0, stream>>>(h, M, F, part);
      else bn_chunk_stats_kernel<1><<<grid, BN_SLICES * BN_LANES, 0, stream>>>(h, M, F, part);
      bn_finalize_kernel<<<phc_div_up((long long)F * 32, 256), 256, 0, stream>>>(part, chunks, M, F, eps, momentum, running_mean, running_var, save_mean,
                                                                save_rstd, num_batches_tracked, n_tracked);
    } else {
      PHC_REQUIRE(running_mean && running_var, "phc_bn_act_drop_skip_fwd: eval mode needs running statistics");
      bn_eval_stats_kernel<<<phc_div_up(F, 128), 128, 0, stream>>>(running_mean, running_var, F, eps, save_mean, save_rstd);
    }
  }
  EwParams p = make_params(M, F, phm_dim, use_bn, act, drop_p, drop_same, training, seed);
  const bool v4 = F % 4 == 0 && phc_aligned16(h) && phc_aligned16(y) && phc_aligned16(skip);
  const long long rgroups = (M + EW_ROWS - 1) / EW_ROWS;
  if (v4) bn_apply_fwd_kernel<4><<<phc_div_up(rgroups * (F / 4), 256), 256, 0, stream>>>(p, h, gamma, beta, save_mean, save_rstd, skip, y);
  else bn_apply_fwd_kernel<1><<<phc_div_up(rgroups * F, 256), 256, 0, stream>>>(p, h, gamma, beta, save_mean, save_rstd, skip, y);
  return phc_check_launch("phc_bn_act_drop_skip_fwd");
}

int phc_bn_act_drop_skip_bwd(const float* dy, const float* h, const float* gamma, const float* beta, const float* save_mean,
                             const float* save_rstd, int rows, int width, int phm_dim, int use_bn, int training, int act, float drop_p,
                             int drop_same, unsigned long long seed, float* dh, float* dgamma, float* dbeta, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(width > 0 && phm_dim > 0 && width % phm_dim == 0, "phc_bn_act_drop_skip_bwd: width %d not divisible by phm_dim %d", width, phm_dim);
  if (rows == 0) return PHC_OK;
  const int M = rows, F = width;
  EwParams p = make_params(M, F, phm_dim, use_bn, act, drop_p, drop_same, training, seed);
  const bool v4 = F % 4 == 0 && phc_aligned16(h) && phc_aligned16(dy) && phc_aligned16(dh);
  float* sum_da = dbeta;
  float* sum_da_xhat = dgamma;
  if (use_bn) {
    PHC_REQUIRE(dgamma && dbeta, "phc_bn_act_drop_skip_bwd: dgamma/dbeta buffers required with batch-norm");
    PHC_REQUIRE(workspace_bytes >= phc_bn_workspace_bytes(M, F), "phc_bn_act_drop_skip_bwd: workspace too small");
    const int chunks = phc_div_up(M, BN_ROWS);
    float* part = reinterpret_cast<float*>(workspace);
    dim3 grid(phc_div_up(F, BN_LANES * (v4 ? 4 : 1)), chunks);
    if (v4) bn_bwd_reduce_kernel<4><<<grid, BN_SLICES * BN_LANES, 0, stream>>>(p, dy, h, gamma, beta, save_mean, save_rstd, part);
    else bn_bwd_reduce_kernel<1><<<grid, BN_SLICES * BN_LANES, 0, stream>>>(p, dy, h, gamma, beta, save_mean, save_rstd, part);
    bn_bwd_finalize_kernel<<<phc_div_up((long long)F * 32, 256), 256, 0, stream>>>(part, chunks, F, dgamma, dbeta);
  }
  const long long rgroups = (M + EW_ROWS - 1) / EW_ROWS;
  if (v4)
    bn_apply_bwd_kernel<4><<<phc_div_up(rgroups * (F / 4), 256), 256, 0, stream>>>(p, training, dy, h, gamma, beta, save_mean, save_rstd,
                                                                                 sum_da, sum_da_xhat, dh);
  else
    bn_apply_bwd_kernel<1><<<phc_div_up(rgroups * F, 256), 256, 0, stream>>>(p, training, dy, h, gamma, beta, save_mean, save_rstd,
                                                                           sum_da, sum_da_xhat, dh);
  return phc_check_launch("phc_bn_act_drop_skip_bwd");
}

}  // extern "C"
