// Task loss of the training step and its gradient in ONE launch.
//
// Replaces the eager chains of the reference's train() bodies — cross entropy (benchmarks/train_ppa.py / train_mnist.py:200:
// log_softmax, nll_loss and their backward kernels), masked BCE-with-logits over the labelled entries (train_hiv.py:174-178,
// train_pcba.py: isnan / where / binary_cross_entropy_with_logits / mul / sum / div: ~20 launches with backward) and mean absolute
// error (train_zinc.py:192).  These tensors are a few kilobytes, so each of those launches is pure latency: ~40 us of the 0.95 ms
// hiv step.  One block computes the mean loss and d(loss)/d(logits) for a unit upstream gradient; sums run in a fixed order
// (thread-strided partials, then a fixed tree), so the result is reproducible run to run.
//   kind 0 "ce"  : logits [B, C], targets int64 [B]            loss = mean_b (logsumexp(l_b) - l_b[y_b])
//   kind 1 "bce" : logits [B, T], targets float [B, T], NaN = unlabelled (skipped); loss = sum(per) / #labelled,
//                  per = max(l, 0) - l*y + log1p(exp(-|l|))   (binary_cross_entropy_with_logits, stable form)
//   kind 2 "l1"  : logits [B] (or [B, 1]), targets float [B]    loss = mean |l - y|
#include "common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ double block_sum_fixed(double v, double* sh) {
  // fixed-shape tree: lane shuffles, then the eight warp sums in warp order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < LOSS_THREADS / 32; ++w) t += sh[w];
  return t;
}

__global__ void __launch_bounds__(LOSS_THREADS) task_loss_kernel(int kind, const float* __restrict__ logits, const void* __restrict__ targets,
                                                                 int B, int C, float* __restrict__ loss, float* __restrict__ dlogits) {
  pdl_begin();
  __shared__ double sh[LOSS_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (kind == 0) {
    // one warp per row: max, sum of exponentials, gradient softmax - onehot
    const long long* y = reinterpret_cast<const long long*>(targets);
    double part = 0.0;
    const float invB = 1.f / (float)B;
    for (int r = warp; r < B; r += LOSS_THREADS / 32) {
      const float* l = logits + (size_t)r * C;
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, l[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += expf(l[c] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float lse = mx + logf(se);
      const int yr = (int)y[r];
      if (lane == 0) part += (unsigned)yr < (unsigned)C ? (double)(lse - l[yr]) : (double)NAN;    // class id out of range: NaN, never a wild read
      const float inv = 1.f / se;
      for (int c = lane; c < C; c += 32) dlogits[(size_t)r * C + c] = (expf(l[c] - mx) * inv - (c == yr ? 1.f : 0.f)) * invB;
    }
    const double tot = block_sum_fixed(part, sh);
    if (threadIdx.x == 0) *loss = (float)(tot / (double)B);
  } else if (kind == 1) {
    const float* y = reinterpret_cast<const float*>(targets);
    const long long n = (long long)B * C;
    double part = 0.0, cnt = 0.0;
    for (long long i = threadIdx.x; i < n; i += LOSS_THREADS) {
      const float yv = y[i];
      if (yv == yv) {                                        // labelled
        const float l = logits[i];
        part += (double)(fmaxf(l, 0.f) - l * yv + log1pf(expf(-fabsf(l))));
        cnt += 1.0;
      }
    }
    const double tot = block_sum_fixed(part, sh);
    const double m = block_sum_fixed(cnt, sh);
    const float invm = (float)(1.0 / m);                     // no labelled entry: inf / nan, as the eager expression gives
    for (long long i = threadIdx.x; i < n; i += LOSS_THREADS) {
      const float yv = y[i];
      const float l = logits[i];
      dlogits[i] = (yv == yv) ? (1.f / (1.f + expf(-l)) - yv) * invm : 0.f;
    }
    if (threadIdx.x == 0) *loss = (float)(tot / m);
  } else {
    const float* y = reinterpret_cast<const float*>(targets);
    double part = 0.0;
    const float invB = 1.f / (float)B;
    for (int i = threadIdx.x; i < B; i += LOSS_THREADS) {
      const float d = logits[i] - y[i];
      part += (double)fabsf(d);
      dlogits[i] = d > 0.f ? invB : (d < 0.f ? -invB : 0.f);
    }
    const double tot = block_sum_fixed(part, sh);
    if (threadIdx.x == 0) *loss = (float)(tot / (double)B);
  }
}

}  // namespace

extern "C" int phc_task_loss(int kind, const float* logits, const void* targets, int rows, int cols, float* loss, float* dlogits,
                             cudaStream_t stream) {
  PHC_REQUIRE(kind >= 0 && kind <= 2, "phc_task_loss: kind %d (0 ce, 1 masked bce, 2 l1)", kind);
  PHC_REQUIRE(rows > 0 && cols > 0, "phc_task_loss: empty logits [%d, %d]", rows, cols);
  PHC_REQUIRE(kind != 2 || cols == 1, "phc_task_loss: l1 takes one logit per row (got %d)", cols);
  phc_launch(task_loss_kernel, dim3(1), dim3(LOSS_THREADS), 0, stream, kind, logits, targets, rows, cols, loss, dlogits);
  return phc_check_launch("phc_task_loss");
}
