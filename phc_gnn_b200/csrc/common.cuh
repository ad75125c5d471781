// Shared device/host helpers for libphc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#define PHC_OK 0
#define PHC_ERR_INVALID 1
#define PHC_ERR_UNSUPPORTED 2
#define PHC_ERR_CUDA 3

// ---- enums shared with include/phc_b200.h -------------------------------------------------
enum { PHC_ACT_IDENTITY = 0, PHC_ACT_RELU = 1, PHC_ACT_LRELU = 2, PHC_ACT_ELU = 3, PHC_ACT_SELU = 4, PHC_ACT_SWISH = 5 };
enum { PHC_RED_SUM = 0, PHC_RED_MEAN = 1, PHC_RED_MAX = 2, PHC_RED_MIN = 3, PHC_RED_SOFTMAX = 4 };

// ---- error reporting (thread-local message, never throws, never syncs) --------------------
void phc_set_error(const char* fmt, ...);
int phc_check_launch(const char* what);

#define PHC_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      phc_set_error(__VA_ARGS__);              \
      return PHC_ERR_INVALID;                  \
    }                                          \
  } while (0)

static inline int phc_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A training step is ~300 short kernels on one stream; the drain + launch gap between stream-ordered kernels adds up
// to ~8 % of the ppa step.  Kernels launched through phc_launch() carry the programmatic-stream-serialization
// attribute, so the launch of kernel i+1 is armed while kernel i still runs, and they block in griddepcontrol.wait
// until kernel i has completed and flushed its memory.  EVERY kernel launched this way must call pdl_begin() before its
// first global-memory access; without the attribute (PHC_NO_PDL=1) the instruction is a no-op.
// Measured on B200 (ppa step): no PDL 6.57 ms, PDL with the implicit trigger at kernel exit 6.02 ms, PDL with an explicit
// griddepcontrol.launch_dependents at the top of every kernel 6.93 ms (the early-resident CTAs of the next kernels
// get in the way of the running one) — hence PHC_PDL_TRIGGER defaults to 0.
#ifdef __CUDACC__
#ifndef PHC_PDL_TRIGGER
#define PHC_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if PHC_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_begin() { pdl_trigger(); pdl_wait(); }
#endif
bool phc_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t phc_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = phc_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- activations ---------------------------------------------------------------------------
#define PHC_SELU_ALPHA 1.6732632423543772848170429916717f
#define PHC_SELU_SCALE 1.0507009873554804934193349852946f

template <int ACT>
__device__ __forceinline__ float act_fwd(float v) {
  if (ACT == PHC_ACT_RELU) return v > 0.f ? v : 0.f;
  if (ACT == PHC_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  if (ACT == PHC_ACT_ELU) return v > 0.f ? v : expm1f(v);
  if (ACT == PHC_ACT_SELU) return PHC_SELU_SCALE * (v > 0.f ? v : PHC_SELU_ALPHA * expm1f(v));
  if (ACT == PHC_ACT_SWISH) return v / (1.f + expf(-v));
  return v;
}
// derivative with respect to the pre-activation v
template <int ACT>
__device__ __forceinline__ float act_bwd(float v) {
  if (ACT == PHC_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (ACT == PHC_ACT_LRELU) return v > 0.f ? 1.f : 0.01f;
  if (ACT == PHC_ACT_ELU) return v > 0.f ? 1.f : expf(v);
  if (ACT == PHC_ACT_SELU) return PHC_SELU_SCALE * (v > 0.f ? 1.f : PHC_SELU_ALPHA * expf(v));
  if (ACT == PHC_ACT_SWISH) {
    float s = 1.f / (1.f + expf(-v));
    return s * (1.f + v * (1.f - s));
  }
  return 1.f;
}
__device__ __forceinline__ float act_fwd_rt(int act, float v) {
  switch (act) {
    case PHC_ACT_RELU: return act_fwd<PHC_ACT_RELU>(v);
    case PHC_ACT_LRELU: return act_fwd<PHC_ACT_LRELU>(v);
    case PHC_ACT_ELU: return act_fwd<PHC_ACT_ELU>(v);
    case PHC_ACT_SELU: return act_fwd<PHC_ACT_SELU>(v);
    case PHC_ACT_SWISH: return act_fwd<PHC_ACT_SWISH>(v);
    default: return v;
  }
}
__device__ __forceinline__ float act_bwd_rt(int act, float v) {
  switch (act) {
    case PHC_ACT_RELU: return act_bwd<PHC_ACT_RELU>(v);
    case PHC_ACT_LRELU: return act_bwd<PHC_ACT_LRELU>(v);
    case PHC_ACT_ELU: return act_bwd<PHC_ACT_ELU>(v);
    case PHC_ACT_SELU: return act_bwd<PHC_ACT_SELU>(v);
    case PHC_ACT_SWISH: return act_bwd<PHC_ACT_SWISH>(v);
    default: return 1.f;
  }
}

// ---- vector access -------------------------------------------------------------------------
// VEC = 4 -> 128-bit accesses (requires F % 4 == 0 and 16-byte aligned bases); VEC = 1 -> scalar.
template <int VEC> struct Vec;
template <> struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec<4> load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  // streaming read-once data (edge embeddings): bypass L1 allocation
  __device__ __forceinline__ static Vec<4> load_stream(const float* p) {
    float4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
    Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ void store_stream(float* p) const {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  }
};
template <> struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec<1> load(const float* p) { Vec<1> r; r.v[0] = *p; return r; }
  __device__ __forceinline__ static Vec<1> load_stream(const float* p) { Vec<1> r; r.v[0] = __ldg(p); return r; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
  __device__ __forceinline__ void store_stream(float* p) const { *p = v[0]; }
};
template <int VEC> struct IVec;
template <> struct IVec<4> {
  int v[4];
  __device__ __forceinline__ static IVec<4> load(const int* p) {
    int4 t = *reinterpret_cast<const int4*>(p);
    IVec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  __device__ __forceinline__ void store(int* p) const { *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]); }
};
template <> struct IVec<1> {
  int v[1];
  __device__ __forceinline__ static IVec<1> load(const int* p) { IVec<1> r; r.v[0] = *p; return r; }
  __device__ __forceinline__ void store(int* p) const { *p = v[0]; }
};

static inline bool phc_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- counter-based RNG for dropout (stateless: mask is a pure function of seed and index) --
// Philox4x32-10 keyed by the 64-bit seed, counter = element-group index; 4 uniforms per call.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// keep-mask bit for flat element index `idx` under keep-probability threshold `thr` (uint32 scale)
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, uint32_t thr) {
  uint32_t r[4];
  philox4x32_10((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  return r[idx & 3] < thr;
}
