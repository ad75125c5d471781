// phm_weight_regularization in one launch pair:  reg = sum_l mean_{k,p} || W_l[:,k,p] ||_2
// (reference phc/hypercomplex/regularization.py:15-23 — one norm + one mean kernel per PHMLinear, and
// their autograd, ~7 launches x #layers per step).  Deterministic: block partials in fixed order.
#include "common.cuh"

#define PHC_REG_MAX 256
#define PHC_REG_BLOCKS 16   // blocks per weight tensor

struct RegTable {
  const float* w[PHC_REG_MAX];
  float* dw[PHC_REG_MAX];
  int n[PHC_REG_MAX];
  int kp[PHC_REG_MAX];
};

namespace {

__global__ void __launch_bounds__(256) reg_fwd_kernel(RegTable t, float* __restrict__ part) {
  pdl_begin();
  const int l = blockIdx.y;
  const int n = t.n[l], kp = t.kp[l];
  const float* w = t.w[l];
  float s = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < kp; i += PHC_REG_BLOCKS * 256) {
    float q = 0.f;
    for (int b = 0; b < n; ++b) { const float v = w[(size_t)b * kp + i]; q += v * v; }
    s += sqrtf(q);
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[l * PHC_REG_BLOCKS + blockIdx.x] = red[0] / (float)kp;
}

__global__ void reg_final_kernel(const float* __restrict__ part, int count, float* __restrict__ out) {
  pdl_begin();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < count * PHC_REG_BLOCKS; ++i) s += part[i];
    *out = s;
  }
}

// dW[b,i] = g * W[b,i] / (||W[:,i]|| * kp)
__global__ void __launch_bounds__(256) reg_bwd_kernel(RegTable t, const float* __restrict__ gout, int accumulate) {
  pdl_begin();
  const int l = blockIdx.y;
  const int n = t.n[l], kp = t.kp[l];
  const float* w = t.w[l];
  float* dw = t.dw[l];
  const float g = __ldg(gout) / (float)kp;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < kp; i += PHC_REG_BLOCKS * 256) {
    float q = 0.f;
    for (int b = 0; b < n; ++b) { const float v = w[(size_t)b * kp + i]; q += v * v; }
    const float inv = q > 0.f ? g * rsqrtf(q) : 0.f;
    if (accumulate) {
      for (int b = 0; b < n; ++b) dw[(size_t)b * kp + i] += w[(size_t)b * kp + i] * inv;
    } else {
      for (int b = 0; b < n; ++b) dw[(size_t)b * kp + i] = w[(size_t)b * kp + i] * inv;
    }
  }
}

int fill(RegTable& t, const float* const* w, float* const* dw, const int* n, const int* kp, int count) {
  for (int i = 0; i < count; ++i) {
    t.w[i] = w[i];
    t.dw[i] = dw ? dw[i] : nullptr;
    t.n[i] = n[i];
    t.kp[i] = kp[i];
  }
  return 0;
}

}  // namespace

extern "C" {

size_t phc_weight_reg_workspace_bytes(int count) { return sizeof(float) * (size_t)count * PHC_REG_BLOCKS + 16; }

int phc_weight_reg_fwd(const float* const* weights, const int* phm_dims, const int* kp, int count, float* out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  PHC_REQUIRE(count >= 0 && count <= PHC_REG_MAX, "phc_weight_reg_fwd: %d weight tensors (max %d)", count, PHC_REG_MAX);
  PHC_REQUIRE(workspace_bytes >= phc_weight_reg_workspace_bytes(count), "phc_weight_reg_fwd: workspace too small");
  if (count == 0) { cudaMemsetAsync(out, 0, sizeof(float), stream); return PHC_OK; }
  RegTable t;
  fill(t, weights, nullptr, phm_dims, kp, count);
  float* part = reinterpret_cast<float*>(workspace);
  phc_launch(reg_fwd_kernel, dim3(dim3(PHC_REG_BLOCKS, count)), dim3(256), 0, stream, t, part);
  phc_launch(reg_final_kernel, dim3(1), dim3(32), 0, stream, part, count, out);
  return phc_check_launch("phc_weight_reg_fwd");
}

static int reg_bwd(const float* gout, const float* const* weights, float* const* dweights, const int* phm_dims, const int* kp, int count,
                   int accumulate, cudaStream_t stream) {
  PHC_REQUIRE(count >= 0 && count <= PHC_REG_MAX, "phc_weight_reg_bwd: %d weight tensors (max %d)", count, PHC_REG_MAX);
  if (count == 0) return PHC_OK;
  RegTable t;
  fill(t, weights, dweights, phm_dims, kp, count);
  phc_launch(reg_bwd_kernel, dim3(dim3(PHC_REG_BLOCKS, count)), dim3(256), 0, stream, t, gout, accumulate);
  return phc_check_launch("phc_weight_reg_bwd");
}

int phc_weight_reg_bwd(const float* gout, const float* const* weights, float* const* dweights, const int* phm_dims, const int* kp,
                       int count, cudaStream_t stream) {
  return reg_bwd(gout, weights, dweights, phm_dims, kp, count, 0, stream);
}

int phc_weight_reg_bwd_accumulate(const float* gout, const float* const* weights, float* const* dweights, const int* phm_dims,
                                  const int* kp, int count, cudaStream_t stream) {
  return reg_bwd(gout, weights, dweights, phm_dims, kp, count, 1, stream);
}

}  // extern "C"
