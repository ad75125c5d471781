// Global-norm gradient clipping + Adam on ONE flat parameter buffer: two launches per training step.
// Replaces torch.nn.utils.clip_grad_norm_(params, 2.0) + torch.optim.Adam.step() of the reference's train()
// (benchmarks/train_hiv.py:199-201), which issue dozens of multi-tensor launches and a Python loop over ~250
// parameter tensors.  Same arithmetic: coef = min(1, max_norm / (||g||_2 + 1e-6)); Adam with bias correction,
// weight_decay = 0 (as the scripts configure it).  Deterministic (fixed-order partial sums).
#include "common.cuh"

namespace {

constexpr int OPT_BLOCKS = 296;

// step_dev (optional): the optimizer's step counter in device memory, advanced here — by the FIRST kernel of the pair, so that every
// block of the second one reads the same value — which makes the step replayable from a CUDA graph (no host scalar changes per step)
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part,
                                                            int* __restrict__ step_dev) {
  pdl_begin();
  if (step_dev != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) { const float v = g[i]; s += v * v; }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                        float bc1, float bc2, float max_norm, const float* __restrict__ part, int nparts,
                                                        float* __restrict__ norm_out, const float* __restrict__ lr_dev,
                                                        const int* __restrict__ step_dev) {
  pdl_begin();
  __shared__ float coef_s, bc_s[2];
  if (threadIdx.x == 0) {
    if (step_dev != nullptr) {      // bias corrections 1 - beta^t from the device-resident step count (double, as the host computes them)
      const double t = (double)*step_dev;
      bc_s[0] = (float)(1.0 - pow((double)b1, t));
      bc_s[1] = (float)(1.0 - pow((double)b2, t));
    } else {
      bc_s[0] = bc1;
      bc_s[1] = bc2;
    }
    float coef = 1.f;
    if (max_norm > 0.f) {
      float s = 0.f;
      for (int i = 0; i < nparts; ++i) s += part[i];          // same order in every block -> identical coefficient
      const float norm = sqrtf(s);
      coef = fminf(1.f, max_norm / (norm + 1e-6f));
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
    coef_s = coef;
  }
  __syncthreads();
  const float coef = coef_s;
  if (lr_dev != nullptr) lr = *lr_dev;
  const float step = lr / bc_s[0], isq = rsqrtf(bc_s[1]);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * isq + eps);
  }
}

}  // namespace

extern "C" {

size_t phc_adam_workspace_bytes(void) { return sizeof(float) * OPT_BLOCKS + 16; }

int phc_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long numel, float lr, float beta1,
                       float beta2, float eps, float bias_correction1, float bias_correction2, float max_norm, float* grad_norm_out,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(numel >= 0, "phc_adam_clip_step: negative size");
  PHC_REQUIRE(workspace_bytes >= phc_adam_workspace_bytes(), "phc_adam_clip_step: workspace too small");
  if (numel == 0) return PHC_OK;
  float* part = reinterpret_cast<float*>(workspace);
  if (max_norm > 0.f) phc_launch(sumsq_partial_kernel, dim3(OPT_BLOCKS), dim3(256), 0, stream, grads, numel, part, (int*)nullptr);
  phc_launch(adam_clip_kernel, dim3(OPT_BLOCKS), dim3(256), 0, stream, params, grads, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, bias_correction1,
                                                   bias_correction2, max_norm, part, OPT_BLOCKS, grad_norm_out, (const float*)nullptr, (const int*)nullptr);
  return phc_check_launch("phc_adam_clip_step");
}

int phc_adam_clip_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long numel, const float* lr_dev,
                           float beta1, float beta2, float eps, int* step_dev, float max_norm, float* grad_norm_out, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(numel >= 0, "phc_adam_clip_step_dev: negative size");
  PHC_REQUIRE(lr_dev != nullptr && step_dev != nullptr, "phc_adam_clip_step_dev: lr and step count must be device pointers");
  PHC_REQUIRE(workspace_bytes >= phc_adam_workspace_bytes(), "phc_adam_clip_step_dev: workspace too small");
  if (numel == 0) return PHC_OK;
  float* part = reinterpret_cast<float*>(workspace);
  // always two launches: the first one also advances the step counter
  phc_launch(sumsq_partial_kernel, dim3(OPT_BLOCKS), dim3(256), 0, stream, grads, max_norm > 0.f ? numel : 0LL, part, step_dev);
  phc_launch(adam_clip_kernel, dim3(OPT_BLOCKS), dim3(256), 0, stream, params, grads, exp_avg, exp_avg_sq, numel, 0.f, beta1, beta2, eps, 1.f, 1.f,
                                                   max_norm, part, OPT_BLOCKS, grad_norm_out, lr_dev, (const int*)step_dev);
  return phc_check_launch("phc_adam_clip_step_dev");
}

}  // extern "C"
