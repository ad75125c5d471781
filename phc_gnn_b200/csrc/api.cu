// libphc_b200.so: error reporting, version, and the PHMLinear precision-mode dispatch.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void phc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool phc_pdl_enabled() {
  static const bool on = getenv("PHC_NO_PDL") == nullptr;
  return on;
}

int phc_check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    phc_set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return PHC_ERR_CUDA;
  }
  return PHC_OK;
}

// fp32 FFMA path (phm_linear_simt.cu)
size_t phm_simt_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim);
int phm_simt_fwd(const float* x, const float* A, const float* W, const float* bias, const float* residual, float* y, int rows,
                 int in_features, int out_features, int phm_dim, int act, cudaStream_t stream);
int phm_simt_bwd(const float* gy, const float* x, const float* A, const float* W, float* dx, float* dA, float* dW, float* db, int rows,
                 int in_features, int out_features, int phm_dim, void* workspace, cudaStream_t stream);
// tensor-core path (phm_linear_tc.cu)
int phm_tc_supported(int rows, int in_features, int out_features, int phm_dim, int precision);
size_t phm_tc_fwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision);
size_t phm_tc_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision);
int phm_tc_fwd(const float* x, const float* A, const float* W, const float* bias, const float* residual, float* y, int rows,
               int in_features, int out_features, int phm_dim, int act, int precision, void* workspace, float* bn_partials,
               int* bn_produced, cudaStream_t stream);
int phm_tc_bwd(const float* gy, const float* x, const float* A, const float* W, float* dx, float* dA, float* dW, float* db, int rows,
               int in_features, int out_features, int phm_dim, int precision, void* workspace, const void* fwd_pack, cudaStream_t stream);

enum { PHC_PREC_FP32 = 0, PHC_PREC_TF32X3 = 1, PHC_PREC_BF16 = 2 };

static int check_linear_shape(const char* who, int rows, int in_features, int out_features, int phm_dim, int precision) {
  PHC_REQUIRE(phm_dim >= 1 && phm_dim <= 16, "%s: phm_dim %d not in 1..16", who, phm_dim);
  PHC_REQUIRE(in_features > 0 && in_features % phm_dim == 0, "%s: in_features=%d is not divisible by phm_dim=%d", who, in_features, phm_dim);
  PHC_REQUIRE(out_features > 0 && out_features % phm_dim == 0, "%s: out_features=%d is not divisible by phm_dim=%d", who, out_features, phm_dim);
  PHC_REQUIRE(rows >= 0, "%s: negative row count", who);
  PHC_REQUIRE(precision >= PHC_PREC_FP32 && precision <= PHC_PREC_BF16, "%s: unknown precision mode %d", who, precision);
  return PHC_OK;
}

extern "C" {

const char* phc_last_error(void) { return g_err; }
int phc_version(void) { return 100; }

size_t phc_phm_linear_fwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision) {
  if (precision != PHC_PREC_FP32 && phm_tc_supported(rows, in_features, out_features, phm_dim, precision))
    return phm_tc_fwd_workspace_bytes(rows, in_features, out_features, phm_dim, precision);
  return 16;
}

size_t phc_phm_linear_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int precision) {
  if (precision != PHC_PREC_FP32 && phm_tc_supported(rows, in_features, out_features, phm_dim, precision))
    return phm_tc_bwd_workspace_bytes(rows, in_features, out_features, phm_dim, precision);
  return phm_simt_bwd_workspace_bytes(rows, in_features, out_features, phm_dim);
}

int phc_phm_linear_fwd_bnstats(const float* x, const float* phm_rule, const float* W, const float* bias, const float* residual, float* y,
                               int rows, int in_features, int out_features, int phm_dim, int act, int precision, void* workspace,
                               size_t workspace_bytes, float* bn_partials, int* bn_produced, cudaStream_t stream) {
  if (bn_produced) *bn_produced = 0;
  int rc = check_linear_shape("phc_phm_linear_fwd", rows, in_features, out_features, phm_dim, precision);
  if (rc) return rc;
  PHC_REQUIRE(act >= PHC_ACT_IDENTITY && act <= PHC_ACT_SWISH, "phc_phm_linear_fwd: bad act %d", act);
  PHC_REQUIRE(workspace_bytes >= phc_phm_linear_fwd_workspace_bytes(rows, in_features, out_features, phm_dim, precision),
              "phc_phm_linear_fwd: workspace too small");
  if (rows == 0) return PHC_OK;
  if (precision != PHC_PREC_FP32 && phm_tc_supported(rows, in_features, out_features, phm_dim, precision))
    return phm_tc_fwd(x, phm_rule, W, bias, residual, y, rows, in_features, out_features, phm_dim, act, precision, workspace, bn_partials,
                      bn_produced, stream);
  return phm_simt_fwd(x, phm_rule, W, bias, residual, y, rows, in_features, out_features, phm_dim, act, stream);
}

int phc_phm_linear_fwd(const float* x, const float* phm_rule, const float* W, const float* bias, const float* residual, float* y, int rows,
                       int in_features, int out_features, int phm_dim, int act, int precision, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  return phc_phm_linear_fwd_bnstats(x, phm_rule, W, bias, residual, y, rows, in_features, out_features, phm_dim, act, precision, workspace,
                                    workspace_bytes, nullptr, nullptr, stream);
}

int phc_phm_linear_bwd(const float* gy, const float* x, const float* phm_rule, const float* W, float* dx, float* d_rule, float* dW,
                       float* dbias, int rows, int in_features, int out_features, int phm_dim, int precision, void* workspace,
                       size_t workspace_bytes, const void* fwd_workspace, cudaStream_t stream) {
  int rc = check_linear_shape("phc_phm_linear_bwd", rows, in_features, out_features, phm_dim, precision);
  if (rc) return rc;
  PHC_REQUIRE(dW != nullptr, "phc_phm_linear_bwd: dW is required");
  PHC_REQUIRE(workspace_bytes >= phc_phm_linear_bwd_workspace_bytes(rows, in_features, out_features, phm_dim, precision),
              "phc_phm_linear_bwd: workspace too small");
  if (precision != PHC_PREC_FP32 && phm_tc_supported(rows, in_features, out_features, phm_dim, precision))
    return phm_tc_bwd(gy, x, phm_rule, W, dx, d_rule, dW, dbias, rows, in_features, out_features, phm_dim, precision, workspace,
                      fwd_workspace, stream);
  return phm_simt_bwd(gy, x, phm_rule, W, dx, d_rule, dW, dbias, rows, in_features, out_features, phm_dim, workspace, stream);
}

}  // extern "C"
