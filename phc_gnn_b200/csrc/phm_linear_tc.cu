// PHMLinear on the 5th-generation tensor cores: tcgen05.mma (kind::tf32), TMEM accumulators, TMA-fed weights.
//
//   y[m,(c,p)] = sum_{a,k} x[m,(a,k)] * H[(a,k),(c,p)] + b,     H[(a,k),(c,p)] = sum_b A[b,a,c] W[b,k,p]
//
// Replaces reference phc/hypercomplex/layers.py:198-219 (einsum Kronecker stack + sum + cuBLAS SGEMM
// + bias kernel, fp32) and its autograd backward.  The Kronecker weight H is NEVER formed — not in HBM
// and not in shared memory.  The contraction is re-associated so that the learned rule A mixes the
// ACTIVATIONS inside the operand producer and the tensor cores multiply by the small W blocks directly:
//
//   y_c[m,p] = sum_{(k,b)}  Xmix_c[m,(k,b)] * W[b,k,p],      Xmix_c[m,(k,b)] = sum_a A[b,a,c] x[m,(a,k)]
//
// i.e. per output component c one GEMM  [M, n*K] x [n*K, P]  whose B operand is W itself (identical for
// every c and every row tile).  Same FLOPs as the dense product (2*M*in*out), n-times less weight traffic.
// The input gradient is the mirror image:  dx_a[m,k] = sum_{(p,b)} Gmix_a[m,(p,b)] W[b,k,p],
// Gmix_a[m,(p,b)] = sum_c A[b,a,c] dy[m,(c,p)].
//
// Precision "tf32x3": every fp32 operand v is split into big = rna_tf32(v), small = rna_tf32(v - big) and
// the tensor cores accumulate  A_small*B_big + A_big*B_small + A_big*B_big  in fp32 (TMEM): fp32-class
// accuracy (error ~2^-21 per product, what the reference's fp32 SGEMM gives) at tensor-core speed.
//
// Kernel anatomy (persistent, warp-specialised, one CTA per SM, 672 threads):
//   warps 16-19 epilogue : tcgen05.ld TMEM -> registers -> smem transpose -> coalesced stores (+bias, act, residual)
//   warp  20    MMA      : TMEM alloc/dealloc; one elected lane issues tcgen05.mma / tcgen05.commit
//   warps 0-15  producers: build the A operand (rule-mixed activations, split to tf32 big/small) straight into the
//                          UMMA canonical layout (K-major, 128-byte swizzle); global loads for chunk i+1 are issued
//                          before chunk i is processed (register prefetch).  Producer thread 0 also issues the TMA
//                          bulk copy (cp.async.bulk -> mbarrier complete_tx) that brings the pre-split, pre-swizzled
//                          W chunk (32 KiB) into the stage.
//   pipeline: 3-stage smem ring (full/empty mbarriers) + 2 TMEM accumulator buffers (tmem_full/tmem_empty), so the
//             epilogue of tile i overlaps the MMAs of tile i+1.
// A small pack kernel per call writes W in both operand orders as tf32 big/small tile images (n*K*P elements,
// L2-resident) together with the rule re-indexed per output component.
//
// Weight gradients:  dH = X^T dY  with the same pipeline (both operands transposed 4x4 in registers, reduction over
// rows split across CTAs, partial tiles reduced in a fixed order), then
//   dW[b,k,p] = sum_{a,c} A[b,a,c] dH[(a,k),(c,p)],   dA[b,a,c] = sum_{k,p} W[b,k,p] dH[(a,k),(c,p)].
#include "common.cuh"
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda.h>   // CUtensorMap (encoded through the runtime's driver entry point; libcuda is not linked)

// helpers implemented in phm_linear_simt.cu (fixed-order reductions shared by both paths)
int phm_contract_and_bias(const float* part, int splits, const float* gy, const float* A, const float* W, float* dA, float* dW, float* db,
                          int rows, int in_features, int out_features, int phm_dim, float* scratch, int db_parts, cudaStream_t stream);
float* phm_contract_bias_partials(float* scratch, int in_features, int out_features, int phm_dim);
size_t phm_contract_scratch_floats(int rows, int in_features, int out_features, int phm_dim);

long long* g_phm_tc_prof = nullptr;   // set via phc_debug_set_tc_profile (debug only)

namespace tc {

#ifndef PHC_SMEM_SPACE
#define PHC_SMEM_SPACE 1
#endif
constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 elements = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
// Warp order matters: the SM's issue arbiter favours the highest warp id, so the latency-critical roles
// (MMA issuer, then epilogue) get the top ids and the throughput-bound producers the low ones.  The epilogue
// warps must start at a multiple of 4 (a warp may only touch TMEM lanes 32*(warp_id % 4) .. +31).
constexpr int PROD_WARP0 = 0;                        // warps 0..15
constexpr int EPI_WARP0 = PROD_WARPS;                // warps 16..19
constexpr int MMA_WARP = PROD_WARPS + EPI_WARPS;     // warp 20
static_assert(EPI_WARP0 % 4 == 0, "epilogue warps must be aligned to the TMEM lane quarters");
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KiB (A and B tiles have the same size, BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A_big, A_small, B_big, B_small
constexpr int TMEM_COLS = 2 * BN;                    // two accumulator buffers
constexpr int EPI_SCRATCH = EPI_WARPS * 32 * 33 * 4; // per-warp 32x33 transpose tiles
constexpr int UNITS = (BM * 8) / PROD_THREADS;       // 16-byte K-units per producer thread per chunk (4)

struct MixParams {          // y = act(mix GEMM + bias) + residual   (FWD and DX)
  const float* X;           // [M, n*Kin]
  const float* coef;        // [n_out_comp t][b][u]  rule re-indexed for this direction
  const uint8_t* Bpack;     // [ptiles][chunks][big|small][128 x 128 B swizzled]
  const float* bias;
  const float* residual;
  float* C;                 // [M, n*Pout]
  int M, n, Kin, Pout, S, chunks, ptiles, act, num_tiles;
  int single;               // 1: "bf16" mode — operands rounded to bf16, one tensor-core pass (no big/small split)
  long long* prof;          // optional per-CTA role timers (debug), 8 slots per CTA
  float* stat_part;         // v3: optional [ceil(M/32)][2][n*Pout] per-32-row-chunk column moments of the OUTPUT (chunk mean, chunk
                            //     M2) for a batch-norm that follows (norm.cu merges them) — the drain owns one column per lane anyway
  float* sk_part;           // v3 stream-K: per-CTA partial accumulators [grid][16 warps][32 rows][64 cols] (NULL: whole units only)
  unsigned int* sk_flags;   // v3 stream-K: [grid][16] "partial written" flags, zero between launches
  int ablate;               // debug (env PHC_TC_ABLATE): 1 no MMAs | 2 no mixing / tcgen05.st | 4 no activation TMA | 8 no drain | 16 no W TMA | 32 math without tcgen05.st | 64 tcgen05.st without the rule FMAs
};

struct DhParams {           // partial dH tiles
  const float* X;           // [M, In]
  const float* G;           // [M, Out]
  float* C;                 // [splits, In, Out]
  float* db_part;           // optional [2 * splits, Out]: column sums of G per split (bias gradient), TMA kernel only
  int M, In, Out, tiles_m, tiles_n, splits, rows_per_split, num_tiles;
  int single;
};

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (error code to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// Spinning wait on the non-blocking test_wait: try_wait may park the warp and wake it late; the MMA issuer has nothing else to do
// and every cycle between "operands ready" and the next tcgen05.mma is a tensor-pipe bubble.  One warp only — producers keep try_wait
// (sixteen spinning warps would take the issue slots of the ones that work).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  if (mbar_test_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_test_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// One lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows the guarded region runs on a single lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// split form for software-pipelined drains: issue the load of the NEXT column block, work on the current one, then wait.  The wait
// names the destination registers as in/out operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, N=BN, M=BM.
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void split_tf32(float v, float& big, float& small) {
  uint32_t b, s;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  big = __uint_as_float(b);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(s) : "f"(v - big));
  small = __uint_as_float(s);
}
__device__ __forceinline__ int swz(int row, int ku) { return row * 128 + ((ku ^ (row & 7)) << 4); }

// write one 16-byte K-unit (4 consecutive k of one row) of an operand tile, big and small copies
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ void store_unit(uint8_t* tile_big, uint8_t* tile_small, int row, int ku, const float (&v)[4], int single = 0) {
  const int off = swz(row, ku);
  if (single) {     // bf16-precision operands (exactly representable in tf32), no correction term
    *reinterpret_cast<float4*>(tile_big + off) = make_float4(round_bf16(v[0]), round_bf16(v[1]), round_bf16(v[2]), round_bf16(v[3]));
    return;
  }
  float b[4], s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_tf32(v[j], b[j], s[j]);
  *reinterpret_cast<float4*>(tile_big + off) = make_float4(b[0], b[1], b[2], b[3]);
  *reinterpret_cast<float4*>(tile_small + off) = make_float4(s[0], s[1], s[2], s[3]);
}

// ---------------------------------------------------------------------------- shared CTA scaffolding
struct Smem {
  uint8_t* stages;
  float* scratch;
  float* coef;
  uint64_t *full_bar, *empty_bar, *tfull_bar, *tempty_bar;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* raw, int coef_floats) {
  Smem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  s.stages = base;
  s.scratch = reinterpret_cast<float*>(base + STAGES * STAGE_BYTES);
  s.coef = s.scratch + EPI_WARPS * 32 * 33;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.coef + ((coef_floats + 3) & ~3));
  s.full_bar = bars;
  s.empty_bar = bars + STAGES;
  s.tfull_bar = bars + 2 * STAGES;
  s.tempty_bar = bars + 2 * STAGES + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  return s;
}

__device__ __forceinline__ uint32_t cta_prologue(const Smem& s, int full_count) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&s.full_bar[i]), full_count);
      mbar_init(smem_u32(&s.empty_bar[i]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&s.tfull_bar[b]), 1);
      mbar_init(smem_u32(&s.tempty_bar[b]), EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(s.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void cta_epilogue(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// MMA issuer for one tile: `chunks` pipeline stages, 4 K-steps x 3 split products each.
template <int NS = STAGES>
__device__ __forceinline__ void mma_tile(const Smem& s, uint32_t tmem_base, int it, int chunks, int& stage, int& phase,
                                         long long* prof = nullptr, int single = 0) {   // prof: thread-local accumulators
  const int lane = threadIdx.x & 31;
  const int buf = it & 1;
  const uint32_t d_tmem = tmem_base + buf * BN;
  long long t0 = prof ? clock64() : 0;
  mbar_wait(smem_u32(&s.tempty_bar[buf]), ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
  if (prof && lane == 0) prof[2] += clock64() - t0;
  tc_fence_after();
  uint32_t accum = 0;
  for (int c = 0; c < chunks; ++c) {
    t0 = prof ? clock64() : 0;
    mbar_wait(smem_u32(&s.full_bar[stage]), phase);
    if (prof && lane == 0) prof[3] += clock64() - t0;
    tc_fence_after();
    if (elect_one()) {                  // elect.sync: descriptors and tensor-memory addresses stay in uniform registers
      const uint32_t sa = smem_u32(s.stages + stage * STAGE_BYTES);
      const uint64_t a_big = make_desc(sa), a_small = make_desc(sa + TILE_BYTES);
      const uint64_t b_big = make_desc(sa + 2 * TILE_BYTES), b_small = make_desc(sa + 3 * TILE_BYTES);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint64_t adv = (uint64_t)((ks * 32) >> 4);            // 8 tf32 = 32 bytes inside the swizzle row
        if (single) {
          umma_tf32(d_tmem, a_big + adv, b_big + adv, IDESC_TF32, accum);
        } else {
          umma_tf32(d_tmem, a_small + adv, b_big + adv, IDESC_TF32, accum);
          umma_tf32(d_tmem, a_big + adv, b_small + adv, IDESC_TF32, 1u);
          umma_tf32(d_tmem, a_big + adv, b_big + adv, IDESC_TF32, 1u);
        }
        accum = 1u;
      }
      umma_commit(smem_u32(&s.empty_bar[stage]));                    // smem slot reusable when these MMAs retire
    }
    __syncwarp();
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
  if (elect_one()) umma_commit(smem_u32(&s.tfull_bar[buf]));        // accumulator complete -> epilogue
  __syncwarp();
}

// Epilogue for one 128 x 128 tile: rows [row0,row0+128) x cols [col0, col0+ncols) of a row-major matrix with
// leading dimension ldc; `col_limit` bounds valid columns of the tile (component / matrix edge).
template <bool FUSE>
__device__ __forceinline__ void epilogue_tile(const Smem& s, uint32_t tmem_base, int it, float* __restrict__ C, int ldc, int rows_c,
                                              int row0, int col0, int ncols, const float* __restrict__ bias,
                                              const float* __restrict__ residual, int act, long long* prof = nullptr) {
  const int warp = (threadIdx.x >> 5) - EPI_WARP0, lane = threadIdx.x & 31;     // == TMEM lane quarter
  float* my = s.scratch + warp * 32 * 33;
  const int buf = it & 1;
  long long t0 = prof ? clock64() : 0;
  mbar_wait(smem_u32(&s.tfull_bar[buf]), (it >> 1) & 1);
  if (prof && threadIdx.x == EPI_WARP0 * 32) { prof[4] += clock64() - t0; t0 = clock64(); }
  tc_fence_after();
  const int r0 = row0 + warp * 32;
  const bool plain = !FUSE || (act == PHC_ACT_IDENTITY);
#pragma unroll 1
  for (int cc = 0; cc < BN / 32; ++cc) {
    if (cc * 32 >= ncols) break;
    uint32_t v[32];
    long long t1 = prof ? clock64() : 0;
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN + cc * 32), v);
    if (prof && threadIdx.x == EPI_WARP0 * 32) prof[7] += clock64() - t1;
#pragma unroll
    for (int j = 0; j < 32; ++j) my[lane * 33 + j] = __uint_as_float(v[j]);
    __syncwarp();
    const int lc = cc * 32 + lane;
    const bool col_ok = lc < ncols;
    const int col = col0 + lc;
    float bv = 0.f;
    if (FUSE && bias != nullptr && col_ok) bv = __ldg(bias + col);
    const int nrows = min(32, rows_c - r0);
    if (col_ok) {
      float* dst = C + (size_t)r0 * ldc + col;
      if (plain && (!FUSE || residual == nullptr)) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr)
          if (rr < nrows) dst[(size_t)rr * ldc] = my[rr * 33 + lane] + bv;
      } else {
        const float* res = (FUSE && residual != nullptr) ? residual + (size_t)r0 * ldc + col : nullptr;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < nrows) {
            float o = my[rr * 33 + lane] + bv;
            if (!plain) o = act_fwd_rt(act, o);
            if (res != nullptr) o += res[(size_t)rr * ldc];
            dst[(size_t)rr * ldc] = o;
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  mbar_arrive(smem_u32(&s.tempty_bar[buf]));
  if (prof && threadIdx.x == EPI_WARP0 * 32) prof[5] += clock64() - t0;
}

// ---------------------------------------------------------------------------- mix kernel (FWD / DX)
// NT = phm_dim at compile time (1, 2, 4 fast paths with register prefetch) or 0 = any n (direct loads).
template <int NT> struct MixLoad { float v[NT == 0 ? 1 : 4 * UNITS]; };

// tile t -> (m-tile, output component, p-tile); the component varies fastest so that concurrently running CTAs
// share the same activation rows in L2.
__device__ __forceinline__ void mix_tile(const MixParams& p, int t, int& m0, int& comp, int& pt) {
  const int per_m = p.n * p.ptiles;
  m0 = (t / per_m) * BM;
  const int r = t % per_m;
  comp = r % p.n;
  pt = r / p.n;
}

template <int NT>
__device__ __forceinline__ void mix_issue_loads(const MixParams& p, int m0, int chunk, int ptid, MixLoad<NT>& L) {
  if (NT == 0) return;
  const int ld = p.n * p.Kin;
#pragma unroll
  for (int i = 0; i < UNITS; ++i) {
    const int u = i * PROD_THREADS + ptid;
    const int row = u >> 3, ku = u & 7;
    const int gm = m0 + row;
    const int s0 = chunk * BK + ku * 4;
    const float* src = p.X + (size_t)gm * ld;
    const bool ok = gm < p.M;
    if (NT == 4) {                    // unit = one k, b = 0..3 : needs x[row, u*Kin + k] for u = 0..3
      const int k = s0 >> 2;
      const bool okk = ok && k < p.Kin;
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) L.v[i * 4 + uu] = okk ? __ldg(src + uu * p.Kin + k) : 0.f;
    } else if (NT == 2) {             // unit = k, k+1 ; b = 0,1
      const int k = s0 >> 1;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) L.v[i * 4 + kk * 2 + uu] = (ok && k + kk < p.Kin) ? __ldg(src + uu * p.Kin + k + kk) : 0.f;
    } else {                          // NT == 1: unit = k..k+3
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) L.v[i * 4 + kk] = (ok && s0 + kk < p.Kin) ? __ldg(src + s0 + kk) : 0.f;
    }
  }
}

template <int NT>
__device__ __forceinline__ void mix_produce(const MixParams& p, const float* __restrict__ cf, int m0, int chunk, int ptid,
                                            const MixLoad<NT>& L, uint8_t* big, uint8_t* small) {
  // cf = coef for this tile's output component: [b][u] (registers for the NT fast paths, smem otherwise)
#pragma unroll
  for (int i = 0; i < UNITS; ++i) {
    const int u = i * PROD_THREADS + ptid;
    const int row = u >> 3, ku = u & 7;
    float v[4];
    if (NT == 4) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float h = 0.f;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) h += cf[b * 4 + uu] * L.v[i * 4 + uu];
        v[b] = h;
      }
    } else if (NT == 2) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int b = 0; b < 2; ++b) v[kk * 2 + b] = cf[b * 2] * L.v[i * 4 + kk * 2] + cf[b * 2 + 1] * L.v[i * 4 + kk * 2 + 1];
    } else if (NT == 1) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) v[kk] = cf[0] * L.v[i * 4 + kk];
    } else {
      const int n = p.n, ld = n * p.Kin;
      const int gm = m0 + row;
      const int s0 = chunk * BK + ku * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sidx = s0 + j;
        float h = 0.f;
        if (gm < p.M && sidx < p.S) {
          const int k = sidx / n, b = sidx - k * n;
          const float* src = p.X + (size_t)gm * ld + k;
          for (int uu = 0; uu < n; ++uu) h += cf[b * n + uu] * __ldg(src + uu * p.Kin);
        }
        v[j] = h;
      }
    }
    store_unit(big, small, row, ku, v, p.single);
  }
}

template <int NT>
__global__ void __launch_bounds__(THREADS, 1) phm_tc_mix_kernel(const MixParams p) {
  pdl_begin();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int n3 = p.n * p.n * p.n;
  const Smem s = carve(smem_raw, n3);
  for (int i = threadIdx.x; i < n3; i += THREADS) s.coef[i] = p.coef[i];
  const uint32_t tmem_base = cta_prologue(s, PROD_WARPS + 1);
  const int warp = threadIdx.x >> 5;
  const long long tk0 = p.prof ? clock64() : 0;

  if (warp < EPI_WARP0) {
    // ===================================================================== producers
    const int ptid = threadIdx.x - PROD_WARP0 * 32;
    int stage = 0, phase = 0;
    // Register ring of PF chunk loads: the global loads of chunk g+PF-1 are issued before chunk g is processed, so
    // ~3 chunk times of L2/HBM latency are hidden (all producer warps work on the same chunk, nothing else hides it).
    constexpr int PF = 4;
    MixLoad<NT> ring[PF];
    const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * p.chunks;
    auto locate = [&](int g, int& m0, int& comp, int& pt, int& c) {
      const int ti = g / p.chunks;
      c = g - ti * p.chunks;
      mix_tile(p, blockIdx.x + ti * gridDim.x, m0, comp, pt);
    };
#pragma unroll
    for (int j = 0; j < PF - 1; ++j) {
      if (j < total) {
        int m0, comp, pt, c;
        locate(j, m0, comp, pt, c);
        mix_issue_loads<NT>(p, m0, c, ptid, ring[j]);
      }
    }
    float cfr[NT == 0 ? 1 : NT * NT];
    int cur_comp = -1;
    long long acc_wait = 0, acc_prod = 0;
    for (int g0 = 0; g0 < total; g0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int g = g0 + j;
        if (g < total) {
          if (g + PF - 1 < total) {
            int m0n, compn, ptn, cn;
            locate(g + PF - 1, m0n, compn, ptn, cn);
            mix_issue_loads<NT>(p, m0n, cn, ptid, ring[(j + PF - 1) % PF]);
          }
          int m0, comp, pt, c;
          locate(g, m0, comp, pt, c);
          const float* cfs = s.coef + comp * p.n * p.n;
          if (NT != 0 && comp != cur_comp) {
#pragma unroll
            for (int i = 0; i < NT * NT; ++i) cfr[i] = cfs[i];
            cur_comp = comp;
          }
          const float* cf = NT != 0 ? cfr : cfs;
          long long tw0 = 0;
          if (p.prof && ptid == 0) tw0 = clock64();
          mbar_wait(smem_u32(&s.empty_bar[stage]), phase ^ 1);
          if (p.prof && ptid == 0) { const long long tn = clock64(); acc_wait += tn - tw0; tw0 = tn; }
          uint8_t* sb = s.stages + stage * STAGE_BYTES;
          if (ptid == 0) {                      // TMA: pre-split W chunk (big+small, already swizzled) -> B_big|B_small
            const uint32_t fb = smem_u32(&s.full_bar[stage]);
            const uint32_t bbytes = p.single ? TILE_BYTES : 2 * TILE_BYTES;
            mbar_arrive_expect_tx(fb, bbytes);
            tma_bulk_load(smem_u32(sb + 2 * TILE_BYTES), p.Bpack + ((size_t)pt * p.chunks + c) * (2 * TILE_BYTES), bbytes, fb);
          }
          mix_produce<NT>(p, cf, m0, c, ptid, ring[j], sb, sb + TILE_BYTES);
          fence_proxy_async();                  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if ((ptid & 31) == 0) mbar_arrive(smem_u32(&s.full_bar[stage]));   // one arrival per producer warp
          if (p.prof && ptid == 0) acc_prod += clock64() - tw0;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.prof && ptid == 0) { p.prof[blockIdx.x * 8 + 0] = acc_wait; p.prof[blockIdx.x * 8 + 1] = acc_prod; }
  } else if (warp == MMA_WARP) {
    int stage = 0, phase = 0, it = 0;
    long long loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it)
      mma_tile(s, tmem_base, it, p.chunks, stage, phase, p.prof ? loc : nullptr, p.single);
    if (p.prof && (threadIdx.x & 31) == 0) { p.prof[blockIdx.x * 8 + 2] = loc[2]; p.prof[blockIdx.x * 8 + 3] = loc[3]; }
  } else {
    int it = 0;
    const int ldc = p.n * p.Pout;
    long long loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int m0, comp, pt;
      mix_tile(p, t, m0, comp, pt);
      const int ncols = min(BN, p.Pout - pt * BN);
      epilogue_tile<true>(s, tmem_base, it, p.C, ldc, p.M, m0, comp * p.Pout + pt * BN, ncols, p.bias, p.residual, p.act,
                          p.prof ? loc : nullptr);
    }
    if (p.prof && threadIdx.x == EPI_WARP0 * 32) {
      p.prof[blockIdx.x * 8 + 4] = loc[4]; p.prof[blockIdx.x * 8 + 5] = loc[5]; p.prof[blockIdx.x * 8 + 7] = loc[7];
      p.prof[blockIdx.x * 8 + 6] = clock64() - tk0;
    }
  }
  cta_epilogue(tmem_base);
}

// ---------------------------------------------------------------------------- mix kernel, TMA-staged activations
// Same contraction as phm_tc_mix_kernel, but the raw activation tiles are fetched by the TMA engine
// (cp.async.bulk.tensor.2d, one [128 rows x 32/n floats] box per input component and K chunk, zero-filled
// outside the matrix) into a 4-deep shared-memory ring, so the producer warps never wait on global memory:
// they read the raw tile from smem, mix it with the rule, split to tf32 big/small and write the UMMA operand.
//   warps 0-15 producers | 16-19 epilogue | 20 MMA issuer | 21 TMA(x) | 22 TMA(W pack)
constexpr int TMA_OP_STAGES = 2, TMA_RAW_STAGES = 3;
constexpr int TMA_X_WARP = MMA_WARP + 1, TMA_B_WARP = MMA_WARP + 2;
constexpr int TMA_THREADS = (MMA_WARP + 3) * 32;
// Raw stage: n boxes of 128 rows x (32/n [+4]) floats.  TMA needs a 16-byte aligned start column; when the
// component width is not a multiple of 4 floats the box starts at the aligned-down column and is 4 floats
// wider, and the producers skip the (constant per component) 0..3 leading floats.
constexpr int RAW_BYTES = BM * (BK + 16) * 4;   // 24 KiB covers the padded n = 4 case

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int NT>
__device__ __forceinline__ void mix_produce_raw(const float* __restrict__ raw, const float (&cf)[NT * NT], int c, int Kin, int pitch,
                                                int ptid, uint8_t* big, uint8_t* small, int single) {
  constexpr int KQ = BK / NT;            // k values per chunk
  // component uu's box: BM rows of `pitch` floats, first needed float at offset (uu*Kin) & 3 when padded
  const int pad = pitch != KQ;
#pragma unroll
  for (int i = 0; i < UNITS; ++i) {
    const int u = i * PROD_THREADS + ptid;
    const int row = u >> 3, ku = u & 7;
    float v[4];
    if (NT == 4) {
      const bool ok = c * KQ + ku < Kin;
      float xv[4];
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) xv[uu] = ok ? raw[uu * (BM * pitch) + row * pitch + (pad ? ((uu * Kin) & 3) : 0) + ku] : 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = cf[b * 4] * xv[0] + cf[b * 4 + 1] * xv[1] + cf[b * 4 + 2] * xv[2] + cf[b * 4 + 3] * xv[3];
    } else if (NT == 2) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const bool ok = c * KQ + 2 * ku + kk < Kin;
        const float x0 = ok ? raw[row * pitch + 2 * ku + kk] : 0.f;
        const float x1 = ok ? raw[BM * pitch + row * pitch + (pad ? (Kin & 3) : 0) + 2 * ku + kk] : 0.f;
        v[kk * 2 + 0] = cf[0] * x0 + cf[1] * x1;
        v[kk * 2 + 1] = cf[2] * x0 + cf[3] * x1;
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) v[kk] = (c * KQ + ku * 4 + kk < Kin) ? cf[0] * raw[row * pitch + ku * 4 + kk] : 0.f;
    }
    store_unit(big, small, row, ku, v, single);
  }
}

template <int NT>
__global__ void __launch_bounds__(TMA_THREADS, 1) phm_tc_mix_tma_kernel(const MixParams p, const __grid_constant__ CUtensorMap tmapX) {
  pdl_begin();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.stages = base;                                                          // [2][A_big|A_small|B_big|B_small]
  uint8_t* rawbuf = base + TMA_OP_STAGES * STAGE_BYTES;                     // [4][16 KiB]
  s.scratch = reinterpret_cast<float*>(rawbuf + TMA_RAW_STAGES * RAW_BYTES);
  s.coef = s.scratch + EPI_WARPS * 32 * 33;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.coef + ((NT * NT * NT + 3) & ~3));
  s.full_bar = bars;                         // [2]
  s.empty_bar = bars + 2;                    // [2]
  uint64_t* rfull_bar = bars + 4;            // [4]
  uint64_t* rempty_bar = bars + 8;           // [4]
  s.tfull_bar = bars + 12;                   // [2]
  s.tempty_bar = bars + 14;                  // [2]
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NT * NT * NT; i += TMA_THREADS) s.coef[i] = p.coef[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < TMA_OP_STAGES; ++i) {
      mbar_init(smem_u32(&s.full_bar[i]), PROD_WARPS + 1);
      mbar_init(smem_u32(&s.empty_bar[i]), 1);
    }
    for (int i = 0; i < TMA_RAW_STAGES; ++i) {
      mbar_init(smem_u32(&rfull_bar[i]), 1);
      mbar_init(smem_u32(&rempty_bar[i]), PROD_WARPS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&s.tfull_bar[b]), 1);
      mbar_init(smem_u32(&s.tempty_bar[b]), EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(s.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;

  constexpr int KQ = BK / NT;
  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int pitch = (p.Kin & 3) ? KQ + 4 : KQ;           // must match the tensor map's box width

  if (warp < EPI_WARP0) {
    // ===================================================================== producers
    const int ptid = threadIdx.x;
    float cfr[NT * NT];
    int g = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      int m0, comp, pt;
      mix_tile(p, blockIdx.x + ti * gridDim.x, m0, comp, pt);
#pragma unroll
      for (int i = 0; i < NT * NT; ++i) cfr[i] = s.coef[comp * NT * NT + i];
      for (int c = 0; c < p.chunks; ++c, ++g) {
        const int rs = g % TMA_RAW_STAGES, os = g % TMA_OP_STAGES;
        mbar_wait(smem_u32(&rfull_bar[rs]), (g / TMA_RAW_STAGES) & 1);                 // raw x tile landed (TMA)
        mbar_wait(smem_u32(&s.empty_bar[os]), ((g / TMA_OP_STAGES) & 1) ^ 1);           // operand slot free (MMA done)
        uint8_t* sb = s.stages + os * STAGE_BYTES;
        mix_produce_raw<NT>(reinterpret_cast<const float*>(rawbuf + rs * RAW_BYTES), cfr, c, p.Kin, pitch, ptid, sb, sb + TILE_BYTES, p.single);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&s.full_bar[os]));
          mbar_arrive(smem_u32(&rempty_bar[rs]));
        }
      }
    }
  } else if (warp == TMA_X_WARP) {
    // ===================================================================== TMA: raw activation boxes
    if (elect_one()) {
      int g = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        int m0, comp, pt;
        mix_tile(p, blockIdx.x + ti * gridDim.x, m0, comp, pt);
        for (int c = 0; c < p.chunks; ++c, ++g) {
          const int rs = g % TMA_RAW_STAGES;
          mbar_wait(smem_u32(&rempty_bar[rs]), ((g / TMA_RAW_STAGES) & 1) ^ 1);
          const uint32_t bar = smem_u32(&rfull_bar[rs]);
          mbar_arrive_expect_tx(bar, NT * BM * pitch * 4);
#pragma unroll
          for (int uu = 0; uu < NT; ++uu)
            tma_load_2d(smem_u32(rawbuf + rs * RAW_BYTES + uu * (BM * pitch * 4)), &tmapX, (uu * p.Kin + c * KQ) & ~3, m0, bar);
        }
      }
    }
  } else if (warp == TMA_B_WARP) {
    // ===================================================================== TMA: pre-split W chunks
    if (elect_one()) {
      int g = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        int m0, comp, pt;
        mix_tile(p, blockIdx.x + ti * gridDim.x, m0, comp, pt);
        for (int c = 0; c < p.chunks; ++c, ++g) {
          const int os = g % TMA_OP_STAGES;
          mbar_wait(smem_u32(&s.empty_bar[os]), ((g / TMA_OP_STAGES) & 1) ^ 1);
          const uint32_t fb = smem_u32(&s.full_bar[os]);
          const uint32_t bbytes = p.single ? TILE_BYTES : 2 * TILE_BYTES;     // bf16 mode needs the "big" image only
          mbar_arrive_expect_tx(fb, bbytes);
          tma_bulk_load(smem_u32(s.stages + os * STAGE_BYTES + 2 * TILE_BYTES),
                        p.Bpack + ((size_t)pt * p.chunks + c) * (2 * TILE_BYTES), bbytes, fb);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    int stage = 0, phase = 0;
    for (int it = 0; it < my_tiles; ++it) mma_tile<TMA_OP_STAGES>(s, tmem_base, it, p.chunks, stage, phase, nullptr, p.single);
  } else {
    const int ldc = p.n * p.Pout;
    for (int it = 0; it < my_tiles; ++it) {
      int m0, comp, pt;
      mix_tile(p, blockIdx.x + it * gridDim.x, m0, comp, pt);
      const int ncols = min(BN, p.Pout - pt * BN);
      epilogue_tile<true>(s, tmem_base, it, p.C, ldc, p.M, m0, comp * p.Pout + pt * BN, ncols, p.bias, p.residual, p.act);
    }
  }
  cta_epilogue(tmem_base);
}

// ---------------------------------------------------------------------------- mix kernel v3 (n = 4): A operand in TENSOR MEMORY,
// two output components per CTA.
// Measured on B200 (in-kernel role timers, tools/tc_bench.py), the smem-staged kernels above are NOT bound by the
// tensor pipe: (1) the operand producers are instruction-issue bound (cvt.rna.tf32.f32 is emulated by 4 SASS
// instructions on sm_100a, ~800 instructions per thread and chunk), (2) every tile pulls 32 KiB of W pack + 24 KiB of raw
// activations per 32-wide K chunk through L2 (~42 B/clk/SM => ~1300 cycles) and (3) moves ~210 KiB through shared
// memory (~1640 cycles) against 768 tensor cycles (12 MMAs).  This kernel attacks all three:
//  * one CTA computes TWO output components of the same 128-row tile, so each W-pack chunk and each raw activation
//    box is fetched once for two tiles;
//  * the rule-mixed, tf32-split A operand never touches shared memory: a producer thread owns ONE row of the tile
//    (TMEM lane == row), mixes its 32 K-columns in registers and writes them with tcgen05.st into a TMEM operand
//    slot; the MMAs take A from TMEM (tcgen05.mma [d], [a], b_desc), only the W pack is read from shared memory;
//  * the fp32 -> tf32 big/small split is a mask and a subtract (big = v & ~0x1fff, small = v - big, exact; the tensor
//    core ignores the low 13 bits of `small`, an error of 2^-21 |v|): 2 instructions instead of 9.
//   TMEM (512 columns): 2 accumulators x 128 | 2 chunk parities x 2 components x (32 big + 32 small) operand columns
//   warps 0-15 producers: (lane quarter q, component h of the pair, chunk parity g) = (w & 3, (w >> 2) & 1, w >> 3);
//   they also drain the accumulators (warp (q,h,g): rows of quarter q, accumulator h, column half g) — the tensor
//   pipe is idle then anyway, both accumulators being busy.  16 MMA | 17 TMA(x) | 18 TMA(W pack)
constexpr int V3_NB = 3;                        // W-pack stages (big|small, 32 KiB each)
constexpr int V3_NR = 2;                        // chunk parities (operand slots in tensor memory)
constexpr int V3_NRAW = 3;                      // raw activation stages: deeper than the parities, the boxes come from L2 / HBM
constexpr int V3_RAW_BYTES = BM * 12 * 4 * 4;   // 4 input components x 128 rows x (8 + 4 pad) floats = 24 KiB
constexpr int V3_A_COL0 = 2 * BN;               // first operand-slot column
constexpr int V3_TMEM_COLS = 512;
// PHC_V3_SPLIT_H: operand-slot hand-over per (chunk parity, component) — a commit behind each component's twelve MMAs and a wait
// just before them, so that the four slots rotate (busy 768 tensor cycles, free 2 304) instead of being handed over in pairs (free
// 1 536, about the producers' turn-around).  With the warp-wide issue loop of rounds 1-2 this was SLOWER (52.9 us against 48.1 us:
// every commit / wait boundary cost the tensor pipe ~200 idle cycles and this doubled them); issue modes 1 / 2 only.
#ifndef PHC_V3_SPLIT_H
#define PHC_V3_SPLIT_H 0
#endif
constexpr int V3_AH = PHC_V3_SPLIT_H ? 2 : 1;   // operand barriers per chunk parity
constexpr int V3_BARS = 2 * V3_NB + 2 * V3_NR * V3_AH + 2 * V3_NRAW + 2;
// MMA issue loop: 0 = the whole warp waits on the barriers, lane 0 issues (rounds 1-2); 1 = lane 0 runs the loop alone; 2 = lane 0
// alone, and the barriers of the NEXT chunk are probed (mbarrier.test_wait, no spin) before the last PHC_V3_HOLD MMAs of the current
// chunk are issued, so that a chunk boundary carries no barrier round trips when the operands are already there.
#ifndef PHC_V3_ISSUE_MODE
#define PHC_V3_ISSUE_MODE 2
#endif
#ifndef PHC_V3_HOLD
#define PHC_V3_HOLD 4
#endif
constexpr int V3_MMA_WARP = PROD_WARPS, V3_TMA_X_WARP = PROD_WARPS + 1, V3_TMA_B_WARP = PROD_WARPS + 2;
constexpr int V3_THREADS = (PROD_WARPS + 3) * 32;
constexpr int V3_SPITCH = 20;                   // floats per scratch row (16 columns per drain pass): 16-byte aligned rows, conflict-free writes
constexpr int V3_SCRATCH = PROD_WARPS * 32 * V3_SPITCH * 4;
#ifndef PHC_TC_PROF
#define PHC_TC_PROF 0
#endif
#ifndef PHC_TC_SPIN
#define PHC_TC_SPIN 1
#endif
#if PHC_TC_SPIN
#define MMA_WAIT(bar, parity) mbar_spin(bar, parity)
#else
#define MMA_WAIT(bar, parity) mbar_wait(bar, parity)
#endif
// Role-ablation switches (MixParams::ablate, env PHC_TC_ABLATE) are compiled in only on request: as run-time tests they cost ~10
// instructions per column pair in the producers' hot loop.  Build with -DPHC_TC_ABLATE_SWITCHES=1 for tools/tc_ablate.sh.
#ifndef PHC_TC_ABLATE_SWITCHES
#define PHC_TC_ABLATE_SWITCHES 0
#endif
#define V3_ABL(mask) (PHC_TC_ABLATE_SWITCHES && (p.ablate & (mask)))
// (Tried: component boxes at their exact, not 16-byte aligned column instead of the aligned column below it plus a 4-float pad — 16
// instead of 24 KiB per raw stage.  cp.async.bulk.tensor rejects such coordinates on sm_100a: illegal instruction.)
#ifndef PHC_V3_HOIST
#define PHC_V3_HOIST 0
#endif
constexpr int V3_HOIST = PHC_V3_HOIST;          // column pairs (of 4 per chunk) mixed before the producers wait for their operand slot

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                 "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// packed fp32x2 arithmetic (PTX fma.rn.f32x2 -> SASS FFMA2, sm_100+); a pair lives in a 64-bit register, lane 0 = low word
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t bcast_f32x2(float v) { return pack_f32x2(v, v); }
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// explicit shared-space accesses (a generic pointer into dynamic shared memory compiles to LD.E / ST.E, which resolve the address
// space per access)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// unit u -> (m-tile, p-tile, component pair)
__device__ __forceinline__ void v3_unit(const MixParams& p, int u, int& m0, int& pair, int& pt) {
  const int per_m = 2 * p.ptiles;
  m0 = (u / per_m) * BM;
  const int r = u % per_m;
  pair = r & 1;
  pt = r >> 1;
}

// Stream-K work split: the (unit, K chunk) list is linearised and CTA b owns the contiguous range [gb, ge) of it, so every
// CTA runs the same number of chunk-steps (+-1) whatever the number of units (244 units on 148 CTAs used to cost two
// full units on 96 CTAs and one on the rest).  A unit that straddles a boundary is computed by two CTAs: the CTA holding
// its LAST chunks processes them first and parks the partial accumulators in global memory (16 warp regions + flags); the
// CTA holding its FIRST chunks processes them last, adds the partial (long since written) and stores the result.  The
// split points and the order of the addition are fixed: results are reproducible run to run.
#ifndef PHC_TC_STREAM_K
#define PHC_TC_STREAM_K 0      // compile-time: the split costs registers in the producers' hot loop even when it is not used
#endif
struct V3Seg { int u, c0, c1; };
__device__ __forceinline__ int v3_range(const MixParams& p, long long& gb, long long& ge) {
  if (PHC_TC_STREAM_K && p.sk_part != nullptr) {
    const long long G = (long long)p.num_tiles * p.chunks;
    gb = G * blockIdx.x / gridDim.x;
    ge = G * (blockIdx.x + 1) / gridDim.x;
  } else {                                                         // whole units only
    gb = ((long long)p.num_tiles * blockIdx.x / gridDim.x) * p.chunks;
    ge = ((long long)p.num_tiles * (blockIdx.x + 1) / gridDim.x) * p.chunks;
  }
  return ge > gb ? (int)((ge - 1) / p.chunks - gb / p.chunks + 1) : 0;
}
__device__ __forceinline__ V3Seg v3_seg(const MixParams& p, long long gb, long long ge, int i) {
  V3Seg sg;
  sg.u = (int)(gb / p.chunks) + i;
  const long long base = (long long)sg.u * p.chunks;
  sg.c0 = gb > base ? (int)(gb - base) : 0;
  sg.c1 = ge - base < p.chunks ? (int)(ge - base) : p.chunks;
  return sg;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* ptr) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* ptr, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

#ifndef PHC_V3_DRAIN_PIPE
#define PHC_V3_DRAIN_PIPE 1
#endif
// One warp drains 32 rows x 64 columns of an accumulator, 16 columns per pass: TMEM -> registers -> smem transpose ->
// row stores (each half warp writes 64 contiguous bytes of one row; the component offset c*P makes rows only 4-byte aligned).
// mode 0: whole unit; 1: park the partial accumulators in `part` and raise `flag`; 2: add the partner's partial first
template <bool STATS>
__device__ __forceinline__ void v3_drain(float* __restrict__ my, uint32_t taddr, float* __restrict__ C, int ldc, int nrows, int r0,
                                         int col0, int ncols, const float* __restrict__ bias, const float* __restrict__ residual, int act,
                                         int mode, float* __restrict__ part, unsigned int* __restrict__ flag,
                                         float* __restrict__ stat) {
  const int lane = threadIdx.x & 31;
  const int rsel = lane >> 4, lc16 = lane & 15;
  const bool plain = act == PHC_ACT_IDENTITY && residual == nullptr;
  if (mode == 2) {
    if (lane == 0 && ld_acquire_u32(flag) == 0u) {
      const long long t0 = clock64();
      while (ld_acquire_u32(flag) == 0u) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) __trap();               // bounded: a protocol bug must trap, never hang
      }
    }
    __syncwarp();
  }
  // Software pipeline (PHC_V3_DRAIN_PIPE): the tensor-memory load of column block cc + 1 is in flight while block cc goes through the
  // transpose and the row stores (tensor memory reads run at 64 B/clk per SM: 2 048 cycles for the 128 KiB of a unit, the exposed
  // load latency of four serial passes came on top).
  // The block's registers are dead once it sits in the transpose scratch, so the next load reuses them: no extra registers.
  uint32_t v[16];
#if PHC_V3_DRAIN_PIPE
  if (ncols > 0) tmem_ld16_issue(taddr, v);
#endif
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    if (cc * 16 >= ncols) break;
#if PHC_V3_DRAIN_PIPE
    tmem_ld16_wait(v);
#else
    tmem_ld16(taddr + (uint32_t)(cc * 16), v);
#endif
    if (mode != 0) {
      float4* pp = reinterpret_cast<float4*>(part + lane * 64 + cc * 16);
      if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcg(pp + j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                     __uint_as_float(v[4 * j + 3])));
#if PHC_V3_DRAIN_PIPE
        if (cc + 1 < 4 && (cc + 1) * 16 < ncols) tmem_ld16_issue(taddr + (uint32_t)((cc + 1) * 16), v);
#endif
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldcg(pp + j);                           // written by another SM: bypass L1
        v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + t.x);
        v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + t.y);
        v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + t.z);
        v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + t.w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(my + lane * V3_SPITCH + j * 4) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
#if PHC_V3_DRAIN_PIPE
    if (cc + 1 < 4 && (cc + 1) * 16 < ncols && mode != 1) tmem_ld16_issue(taddr + (uint32_t)((cc + 1) * 16), v);
#endif
    const int lc = cc * 16 + lc16;
    const bool on = lc < ncols;
    const int col = col0 + lc;
    const float bv = (on && bias != nullptr) ? __ldg(bias + col) : 0.f;
    float* dst = C + (size_t)r0 * ldc + col;
    const float* src = my + lc16;
    if (STATS) {
      // batch-norm statistics of the stored values for free: shifted moments of this warp's 32-row chunk (shift = the
      // chunk's first row); the two half warps hold the even / odd rows of a column and are combined by one shuffle
      const float* res = residual != nullptr ? residual + (size_t)r0 * ldc + col : nullptr;
      float sh = 0.f, s1 = 0.f, s2 = 0.f;
      if (on && plain && nrows == 32) {                            // the common case, fully unrolled
        sh = src[0] + bv;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
          const float o = src[(2 * rr + rsel) * V3_SPITCH] + bv;
          dst[(2 * rr + rsel) * ldc] = o;
          const float d = o - sh;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
      } else if (on) {
        sh = act_fwd_rt(act, src[0] + bv);
        if (res != nullptr) sh += res[0];
        for (int rr = rsel; rr < nrows; rr += 2) {
          float o = act_fwd_rt(act, src[rr * V3_SPITCH] + bv);
          if (res != nullptr) o += res[rr * ldc];
          dst[rr * ldc] = o;
          const float d = o - sh;
          s1 += d;
          s2 += d * d;
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      if (on && rsel == 0) {
        const float cnt = (float)nrows;
        float* sp = stat + ((size_t)(r0 >> 5) * 2) * ldc + col;
        sp[0] = sh + s1 / cnt;
        sp[ldc] = fmaxf(s2 - s1 * s1 / cnt, 0.f);
      }
    } else if (on) {
      if (plain && nrows == 32) {
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) dst[(2 * rr + rsel) * ldc] = src[(2 * rr + rsel) * V3_SPITCH] + bv;
      } else if (plain) {
        for (int rr = rsel; rr < nrows; rr += 2) dst[rr * ldc] = src[rr * V3_SPITCH] + bv;
      } else {
        const float* res = residual != nullptr ? residual + (size_t)r0 * ldc + col : nullptr;
        for (int rr = rsel; rr < nrows; rr += 2) {
          float o = act_fwd_rt(act, src[rr * V3_SPITCH] + bv);
          if (res != nullptr) o += res[rr * ldc];
          dst[rr * ldc] = o;
        }
      }
    }
    __syncwarp();
  }
  if (mode == 1) {
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_u32(flag, 1u);
  } else if (mode == 2) {
    __syncwarp();
    if (lane == 0) *flag = 0u;                                     // consumed: the flags are zero again for the next launch
  }
}

// R = Kin & 3: component uu's box starts (uu*R)&3 floats before its first needed float (TMA boxes start 16-byte aligned)
template <int R, bool SINGLE, bool STATS>
__global__ void __launch_bounds__(V3_THREADS, 1) phm_tc_mix_v3_kernel(const MixParams p, const __grid_constant__ CUtensorMap tmapX) {
  pdl_begin();
  constexpr int NT = 4, KQ = BK / NT;                       // 8 k values per chunk
  constexpr int PITCH = R == 0 ? KQ : KQ + 4;               // floats per raw row (must match the tensor map's box)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
#if PHC_SMEM_SPACE
  // 1 KiB alignment by POINTER arithmetic on the shared array: the detour through uintptr_t made every pointer derived from `base`
  // generic, and the drain's transpose scratch, the rule table and the barrier words were read with LD.E / ST.E instead of LDS / STS
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#else
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#endif
  uint8_t* bstage = base;                                                   // [V3_NB][B_big | B_small]
  uint8_t* rawbuf = base + V3_NB * 2 * TILE_BYTES;                          // [V3_NRAW][4][128][PITCH]
  float* scratch = reinterpret_cast<float*>(rawbuf + V3_NRAW * V3_RAW_BYTES);  // [16 warps][32][V3_SPITCH]
  float* coef = scratch + PROD_WARPS * 32 * V3_SPITCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(coef + NT * NT * NT);
  uint64_t* bfull = bars;                                  // [V3_NB]
  uint64_t* bempty = bars + V3_NB;                         // [V3_NB]
  uint64_t* afull = bars + 2 * V3_NB;                      // [V3_NR][V3_AH]
  uint64_t* aempty = afull + V3_NR * V3_AH;                // [V3_NR][V3_AH]
  uint64_t* rfull = aempty + V3_NR * V3_AH;                // [V3_NRAW]
  uint64_t* rempty = rfull + V3_NRAW;                      // [V3_NRAW]
  uint64_t* tfull = rempty + V3_NRAW;                      // accumulators complete -> drain
  uint64_t* tempty = tfull + 1;                            // accumulators drained  -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + V3_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NT * NT * NT; i += V3_THREADS) coef[i] = p.coef[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < V3_NB; ++i) { mbar_init(smem_u32(&bfull[i]), 1); mbar_init(smem_u32(&bempty[i]), 1); }
    for (int i = 0; i < V3_NR * V3_AH; ++i) {
      mbar_init(smem_u32(&afull[i]), 8 / V3_AH);   // the producer warps of the slot(s): (2 components x) 4 lane quarters
      mbar_init(smem_u32(&aempty[i]), 1);          // tcgen05.commit
    }
    for (int i = 0; i < V3_NRAW; ++i) {
      mbar_init(smem_u32(&rfull[i]), 1);      // TMA transaction
      mbar_init(smem_u32(&rempty[i]), 8);     // the eight producer warps that consume the chunk
    }
    mbar_init(smem_u32(tfull), 1);
    mbar_init(smem_u32(tempty), PROD_WARPS);
    fence_barrier_init();
  }
  if (warp == V3_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), V3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long gb, ge;
  const int nseg = v3_range(p, gb, ge);                              // num_tiles = #units; this CTA's share of the (unit, chunk) list

  if (warp < PROD_WARPS) {
    // ===================================================================== producers (+ accumulator drain)
    const int q = warp & 3, h = (warp >> 2) & 1, g = warp >> 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_slot = tmem_base + lane_base + (uint32_t)(V3_A_COL0 + (g * 2 + h) * 64);
    const int ab = PHC_V3_SPLIT_H ? g * 2 + h : g;               // this warp's operand barrier pair
    const uint32_t raw0 = smem_u32(rawbuf) + (uint32_t)(row * PITCH * 4);
    float* my = scratch + warp * 32 * V3_SPITCH;
    const int ldc = p.n * p.Pout;
    uint64_t cf2[NT][2];                                         // rule coefficients as packed pairs: [uu][b pair] = (A'[2bp][uu], A'[2bp+1][uu])
    int cur_comp = -1;
    // producer-side role timers only in -DPHC_TC_PROF=1 builds: five 64-bit counters live across the whole loop cost the registers
    // the hoisted mixing needs (the kernel sits at its 96-register cap); the MMA issuer's timers (waits, kernel total) stay
    const bool prof = PHC_TC_PROF && p.prof != nullptr && warp == 0 && lane == 0;
    long long t_rfull = 0, t_aempty = 0, t_work = 0, t_drain = 0, tp = 0;

    auto produce = [&](int gi, int c, int comp) {
      if (comp != cur_comp) {
#pragma unroll
        for (int uu = 0; uu < NT; ++uu)
#pragma unroll
          for (int bp = 0; bp < 2; ++bp)                                          // smem table is [b][uu]
            cf2[uu][bp] = pack_f32x2(coef[comp * NT * NT + (2 * bp) * NT + uu], coef[comp * NT * NT + (2 * bp + 1) * NT + uu]);
        cur_comp = comp;
      }
      const int use = gi >> 1;                                   // how often this parity's operand slot has been used
      const int rs = gi % V3_NRAW;                               // raw stage of this chunk
      const uint32_t rawg = raw0 + (uint32_t)(rs * V3_RAW_BYTES);
      if (prof) tp = clock64();
      if (!V3_ABL(4)) mbar_wait(smem_u32(&rfull[rs]), (gi / V3_NRAW) & 1);      // raw boxes landed (TMA)
      if (prof) { const long long tn = clock64(); t_rfull += tn - tp; tp = tn; }
      float xr[NT][KQ];
      uint32_t landed = 0;                                       // data dependence on every load (see the release below)
#pragma unroll
      for (int uu = 0; uu < NT; ++uu) {
        constexpr int NV = PITCH / 4;
        float t[NV * 4];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(t[4 * v]), "=f"(t[4 * v + 1]), "=f"(t[4 * v + 2]), "=f"(t[4 * v + 3])
                       : "r"(rawg + (uint32_t)((uu * (BM * PITCH) + v * 4) * 4)));
          landed |= __float_as_uint(t[4 * v]);
        }
        const int off = (uu * R) & 3;                            // compile-time after unrolling
#pragma unroll
        for (int k = 0; k < KQ; ++k) xr[uu][k] = t[off + k];
      }
      // Early release: the row is in registers, let TMA refill the stage while we mix.  The arrive must not be performed
      // before the loads have RETURNED (issuing them is not enough: measured as a rare corruption with the short
      // R = 0 rows).  A shared-memory store of a value that depends on every load cannot issue before they return, and
      // the arrive (release) stays behind the store.
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(my + lane)), "r"(landed) : "memory");
      __syncwarp();
      if (lane == 0 && !V3_ABL(4)) mbar_arrive(smem_u32(&rempty[rs]));
      if (c == p.chunks - 1) {                                   // K tail: columns past the component belong to its neighbour
#pragma unroll
        for (int k = 0; k < KQ; ++k)
          if (c * KQ + k >= p.Kin) {
#pragma unroll
            for (int uu = 0; uu < NT; ++uu) xr[uu][k] = 0.f;
          }
      }
      // mixing of one column pair (2 k values = 8 operand columns, column = k*4 + b) with packed fp32x2 FMAs (FFMA2, sm_100: one
      // instruction, two products) — the producers' FMA issue is what stretches the MMAs from 68 to 78-87 cycles
      // (profiles/r02_mix_v3_ablation.md).  The two lanes of a packed operation are two RULE ROWS b, b+1 of the same k: the raw
      // activation is the scalar-broadcast operand (a plain register, whatever its alignment in the staged row), the coefficient
      // pair is loop-invariant, and the results (k: b0 b1 | b2 b3) are already in operand-column order, so the eight registers of a
      // tcgen05.st need no shuffling.  (Pairing two k values instead cost 16 register moves per column pair to transpose the
      // results and up to 19 more to assemble unaligned input pairs when Kin % 4 != 0: 142 moves against 80 packed FMAs per chunk.)
      auto mix_pair = [&](int kp, float (&big)[8], float (&small)[8]) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = kp + kk;
#pragma unroll
          for (int bp = 0; bp < 2; ++bp) {
            uint64_t v2;
            if (V3_ABL(64)) {
              v2 = pack_f32x2(xr[2 * bp][k], xr[2 * bp + 1][k]);
            } else {
              v2 = mul_f32x2(bcast_f32x2(xr[0][k]), cf2[0][bp]);
#pragma unroll
              for (int uu = 1; uu < NT; ++uu) v2 = fma_f32x2(bcast_f32x2(xr[uu][k]), cf2[uu][bp], v2);
            }
            const int o = kk * 4 + bp * 2;
            if (SINGLE) {                                          // bf16 operands: round to nearest (ties away) by add + mask
              const uint64_t r = ((v2 & 0xFFFFFFFFull) + 0x8000ull) & 0xFFFF0000ull;
              const uint64_t h = (((v2 >> 32) + 0x8000ull) & 0xFFFF0000ull) << 32;
              unpack_f32x2(r | h, big[o], big[o + 1]);
            } else {
              const uint64_t b2 = v2 & 0xFFFFE000FFFFE000ull;      // tf32 "big" parts of both lanes
              const uint64_t s2 = fma_f32x2(b2, bcast_f32x2(-1.f), v2);   // v - big, exact
              unpack_f32x2(b2, big[o], big[o + 1]);
              unpack_f32x2(s2, small[o], small[o + 1]);
            }
          }
        }
      };
      auto store_pair = [&](int kp, const float (&big)[8], const float (&small)[8]) {
        if (V3_ABL(32)) {                                     // keep the arithmetic alive without the tensor-memory stores
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += big[j] + (SINGLE ? 0.f : small[j]);
          if (acc == 1.2345e-30f) my[lane] = acc;
        } else {
          tmem_st8(a_slot + kp * 4, big);
          if (!SINGLE) tmem_st8(a_slot + 32 + kp * 4, small);
        }
      };
      // V3_HOIST column pairs of the chunk can be mixed BEFORE waiting for the operand slot (the producers' turn-around, slot free ->
      // operand written, is ~1 500 cycles against ~1 750 tensor cycles per chunk, and the MMA issuer waits for operands 12 % of the
      // kernel).  Measured: one hoisted pair spills (the kernel sits at its 96-register cap) and is 3 % SLOWER (83.3k vs 80.4k
      // cycles), two pairs spill 300 bytes per thread — the default is 0 (profiles/r02_mix_v3_ablation.md).
      float bigA[V3_HOIST > 0 ? V3_HOIST : 1][8], smallA[V3_HOIST > 0 ? V3_HOIST : 1][8];
      if (!V3_ABL(2)) {
#pragma unroll
        for (int h2 = 0; h2 < V3_HOIST; ++h2) mix_pair(2 * h2, bigA[h2], smallA[h2]);
      }
      if (prof) { const long long tn = clock64(); t_work += tn - tp; tp = tn; }
      mbar_wait(smem_u32(&aempty[ab]), (use & 1) ^ 1);         // the MMAs that read this operand slot have retired
      if (prof) { const long long tn = clock64(); t_aempty += tn - tp; tp = tn; }
      tc_fence_after();
      if (!V3_ABL(2)) {
#pragma unroll
        for (int h2 = 0; h2 < V3_HOIST; ++h2) store_pair(2 * h2, bigA[h2], smallA[h2]);
#pragma unroll
        for (int kp = 2 * V3_HOIST; kp < KQ; kp += 2) {
          float big[8], small[8];
          mix_pair(kp, big, small);
          store_pair(kp, big, small);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&afull[ab]));
      if (prof) t_work += clock64() - tp;
    };
    auto drain = [&](int si) {
      const V3Seg sg = v3_seg(p, gb, ge, si);
      int m0, pair, pt;
      v3_unit(p, sg.u, m0, pair, pt);
      const int ncols = min(BN, p.Pout - pt * BN) - g * 64;     // valid columns of this warp's half
      const long long td0 = prof ? clock64() : 0;
      mbar_wait(smem_u32(tfull), si & 1);
      tc_fence_after();
      const int r0 = m0 + q * 32;
      // split unit: its last chunks (c0 > 0) are parked in this CTA's slot; its first chunks (c1 < chunks) belong to the
      // owner, which adds the partial parked by the NEXT CTA (whose first segment is the rest of this unit)
      const int mode = PHC_TC_STREAM_K ? (sg.c0 > 0 ? 1 : (sg.c1 < p.chunks ? 2 : 0)) : 0;
      const int slot = (int)blockIdx.x + (mode == 2 ? 1 : 0);
      if (ncols > 0 && r0 < p.M && !V3_ABL(8))
        v3_drain<STATS>(my, tmem_base + lane_base + (uint32_t)(h * BN + g * 64), p.C, ldc, min(32, p.M - r0), r0,
                 (2 * pair + h) * p.Pout + pt * BN + g * 64, ncols, p.bias, p.residual, p.act, mode,
                 mode ? p.sk_part + ((size_t)slot * PROD_WARPS + warp) * (32 * 64) : nullptr,
                 mode ? p.sk_flags + slot * PROD_WARPS + warp : nullptr, p.stat_part);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty));
      if (prof) t_drain += clock64() - td0;
    };

    int gi0 = 0;                                                 // CTA-local index of the segment's first chunk
    for (int si = 0; si < nseg; ++si) {
      const V3Seg sg = v3_seg(p, gb, ge, si);
      int m0, pair, pt;
      v3_unit(p, sg.u, m0, pair, pt);
      const int comp = 2 * pair + h;
      int c = sg.c0 + ((gi0 & 1) == g ? 0 : 1);                  // this parity's first chunk of the segment
      if (c < sg.c1) { produce(gi0 + c - sg.c0, c, comp); c += 2; }   // ... goes in BEFORE draining the previous one, so the
      if (si > 0) drain(si - 1);                                 //     tensor pipe restarts the moment the accumulators are free
      for (; c < sg.c1; c += 2) produce(gi0 + c - sg.c0, c, comp);
      gi0 += sg.c1 - sg.c0;
    }
    if (nseg > 0) drain(nseg - 1);
    if (prof) {
      p.prof[blockIdx.x * 8 + 0] = t_rfull; p.prof[blockIdx.x * 8 + 1] = t_work; p.prof[blockIdx.x * 8 + 7] = t_aempty;
      p.prof[blockIdx.x * 8 + 5] = t_drain;
    }
  } else if (warp == V3_TMA_X_WARP) {
    // ===================================================================== TMA: raw activation boxes
    if (elect_one() && !V3_ABL(4)) {
      int gi = 0;
      for (int si = 0; si < nseg; ++si) {
        const V3Seg sg = v3_seg(p, gb, ge, si);
        int m0, pair, pt;
        v3_unit(p, sg.u, m0, pair, pt);
        for (int c = sg.c0; c < sg.c1; ++c, ++gi) {
          const int rs = gi % V3_NRAW;
          mbar_wait(smem_u32(&rempty[rs]), ((gi / V3_NRAW) & 1) ^ 1);
          const uint32_t bar = smem_u32(&rfull[rs]);
          mbar_arrive_expect_tx(bar, NT * BM * PITCH * 4);
#pragma unroll
          for (int uu = 0; uu < NT; ++uu)
            tma_load_2d(smem_u32(rawbuf + rs * V3_RAW_BYTES + uu * (BM * PITCH * 4)), &tmapX, (uu * p.Kin + c * KQ) & ~3, m0, bar);
        }
      }
    }
  } else if (warp == V3_TMA_B_WARP) {
    // ===================================================================== TMA: pre-split W chunks
    if (elect_one() && !V3_ABL(16)) {
      int gi = 0;
      for (int si = 0; si < nseg; ++si) {
        const V3Seg sg = v3_seg(p, gb, ge, si);
        int m0, pair, pt;
        v3_unit(p, sg.u, m0, pair, pt);
        for (int c = sg.c0; c < sg.c1; ++c, ++gi) {
          const int bs = gi % V3_NB;
          mbar_wait(smem_u32(&bempty[bs]), ((gi / V3_NB) & 1) ^ 1);
          const uint32_t fb = smem_u32(&bfull[bs]);
          const uint32_t bbytes = SINGLE ? TILE_BYTES : 2 * TILE_BYTES;
          mbar_arrive_expect_tx(fb, bbytes);
          tma_bulk_load(smem_u32(bstage + bs * 2 * TILE_BYTES), p.Bpack + ((size_t)pt * p.chunks + c) * (2 * TILE_BYTES), bbytes, fb);
        }
      }
    }
  } else {
    // ===================================================================== MMA issuer: A from TMEM, B from smem
    const bool prof = p.prof != nullptr && lane == 0;
    long long t_tempty = 0, t_afull = 0, t_bfull = 0, tp = 0;
    const long long tk0 = prof ? clock64() : 0;
    // the 24 (SINGLE: 8) MMAs of a chunk in issue order; `first` clears the accumulators at the start of a unit
    auto issue = [&](int m, int g, uint64_t b_big, uint64_t b_small, uint32_t first) {
      constexpr int PER_H = SINGLE ? BK / 8 : 3 * (BK / 8);
      const int h = m / PER_H, r = m % PER_H;
      const int ks = SINGLE ? r : r / 3, pass = SINGLE ? 2 : r % 3;
      const uint32_t d_tmem = tmem_base + h * BN;
      const uint32_t a_big = tmem_base + (uint32_t)(V3_A_COL0 + (g * 2 + h) * 64) + ks * 8, a_small = a_big + 32;
      const uint64_t adv = (uint64_t)((ks * 32) >> 4);
      const uint32_t acc = r == 0 ? first : 1u;
      if (SINGLE) umma_tf32_ts(d_tmem, a_big, b_big + adv, IDESC_TF32, acc);
      else if (pass == 0) umma_tf32_ts(d_tmem, a_small, b_big + adv, IDESC_TF32, acc);
      else if (pass == 1) umma_tf32_ts(d_tmem, a_big, b_small + adv, IDESC_TF32, 1u);
      else umma_tf32_ts(d_tmem, a_big, b_big + adv, IDESC_TF32, 1u);
    };
    constexpr int NMMA = SINGLE ? 2 * (BK / 8) : 6 * (BK / 8);
    if (PHC_V3_ISSUE_MODE == 0) {
      int gi = 0;
      for (int si = 0; si < nseg; ++si) {
        const V3Seg sg = v3_seg(p, gb, ge, si);
        if (prof) tp = clock64();
        MMA_WAIT(smem_u32(tempty), (si & 1) ^ 1);                // both accumulators drained
        if (prof) t_tempty += clock64() - tp;
        tc_fence_after();
        uint32_t accum = 0;
        for (int c = sg.c0; c < sg.c1; ++c, ++gi) {
          const int g = gi & 1, bs = gi % V3_NB;
          if (prof) tp = clock64();
          MMA_WAIT(smem_u32(&afull[g]), (gi >> 1) & 1);
          if (prof) { const long long tn = clock64(); t_afull += tn - tp; tp = tn; }
          if (!V3_ABL(16)) MMA_WAIT(smem_u32(&bfull[bs]), (gi / V3_NB) & 1);
          if (prof) t_bfull += clock64() - tp;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_u32(bstage + bs * 2 * TILE_BYTES);
            const uint64_t b_big = make_desc(sb), b_small = make_desc(sb + TILE_BYTES);
            if (!V3_ABL(1)) {
#pragma unroll
              for (int m = 0; m < NMMA; ++m) issue(m, g, b_big, b_small, accum);
            }
            umma_commit(smem_u32(&aempty[g]));
            if (!V3_ABL(16)) umma_commit(smem_u32(&bempty[bs]));
          }
          accum = 1u;
          __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(tfull));
        __syncwarp();
      }
    } else if (elect_one()) {
      // One thread runs the whole issue loop.  Measured (round 2): the tensor pipe idles ~200 cycles at every chunk boundary — the
      // queue behind tcgen05.mma is shallow, and commit + two barrier round trips + fence sit between the last MMA of a chunk and
      // the first of the next.  Mode 2 moves the round trips in front of the chunk's last MMAs.
      static_assert(PHC_V3_ISSUE_MODE != 0 || !PHC_V3_SPLIT_H, "PHC_V3_SPLIT_H needs the single-thread issue loop");
      constexpr int NGRP = V3_AH, PER = NMMA / NGRP;             // issue groups per chunk (one per operand barrier), MMAs per group
      constexpr int HOLD = PHC_V3_HOLD < PER ? PHC_V3_HOLD : PER - 1;
      int gi = 0;
      bool ready = false;                                        // the next group's barriers were already seen complete
      for (int si = 0; si < nseg; ++si) {
        const V3Seg sg = v3_seg(p, gb, ge, si);
        if (prof) tp = clock64();
        MMA_WAIT(smem_u32(tempty), (si & 1) ^ 1);                // both accumulators drained
        if (prof) t_tempty += clock64() - tp;
        uint32_t accum = 0;
        for (int c = sg.c0; c < sg.c1; ++c, ++gi) {
          const int g = gi & 1, bs = gi % V3_NB;
          const uint32_t sb = smem_u32(bstage + bs * 2 * TILE_BYTES);
          const uint64_t b_big = make_desc(sb), b_small = make_desc(sb + TILE_BYTES);
#pragma unroll
          for (int hg = 0; hg < NGRP; ++hg) {
            if (!ready) {
              if (prof) tp = clock64();
              MMA_WAIT(smem_u32(&afull[g * NGRP + hg]), (gi >> 1) & 1);
              if (prof) { const long long tn = clock64(); t_afull += tn - tp; tp = tn; }
              if (hg == 0 && !V3_ABL(16)) MMA_WAIT(smem_u32(&bfull[bs]), (gi / V3_NB) & 1);
              if (prof) t_bfull += clock64() - tp;
            }
            tc_fence_after();
            ready = false;
            if (!V3_ABL(1)) {
#pragma unroll
              for (int m = hg * PER; m < (hg + 1) * PER - HOLD; ++m) issue(m, g, b_big, b_small, accum);
            }
            if (PHC_V3_ISSUE_MODE == 2) {
              if (hg + 1 < NGRP) {                               // the other component of this chunk
                ready = mbar_test_wait(smem_u32(&afull[g * NGRP + hg + 1]), (gi >> 1) & 1);
              } else if (c + 1 < sg.c1 || si + 1 < nseg) {       // the next chunk of this CTA, whatever its unit
                const int gn = gi + 1;
                ready = mbar_test_wait(smem_u32(&afull[(gn & 1) * NGRP]), (gn >> 1) & 1) &&
                        (V3_ABL(16) || mbar_test_wait(smem_u32(&bfull[gn % V3_NB]), (gn / V3_NB) & 1));
              }
            }
            if (!V3_ABL(1)) {
#pragma unroll
              for (int m = (hg + 1) * PER - HOLD; m < (hg + 1) * PER; ++m) issue(m, g, b_big, b_small, accum);
            }
            umma_commit(smem_u32(&aempty[g * NGRP + hg]));       // the slot(s) are free once these MMAs have retired
          }
          if (!V3_ABL(16)) umma_commit(smem_u32(&bempty[bs]));
          accum = 1u;
        }
        umma_commit(smem_u32(tfull));
      }
    }
    if (prof) {
      p.prof[blockIdx.x * 8 + 2] = t_tempty; p.prof[blockIdx.x * 8 + 3] = t_afull; p.prof[blockIdx.x * 8 + 4] = t_bfull;
      p.prof[blockIdx.x * 8 + 6] = clock64() - tk0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == V3_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, V3_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------- dH kernel
// operands: tile row = feature index f0.., k = sample index m; global [M, ld] row-major is read with 128-bit
// loads along the feature axis and transposed 4x4 in registers.
struct DhLoad { float t[4][4]; };

__device__ __forceinline__ void dh_issue_loads(const float* __restrict__ X, int ld, bool vec, int f0, int mk0, int mend, int ptid,
                                               DhLoad& L) {
  const int ku = ptid & 7, fg = ptid >> 3;      // 8 k-units x 32 feature groups of 4
  const int gf = f0 + fg * 4;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int gm = mk0 + ku * 4 + jj;
#pragma unroll
    for (int c = 0; c < 4; ++c) L.t[jj][c] = 0.f;
    if (gm < mend) {
      const float* src = X + (size_t)gm * ld + gf;
      if (vec && gf + 3 < ld) {
        float4 q = __ldg(reinterpret_cast<const float4*>(src));
        L.t[jj][0] = q.x; L.t[jj][1] = q.y; L.t[jj][2] = q.z; L.t[jj][3] = q.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gf + c < ld) L.t[jj][c] = __ldg(src + c);
      }
    }
  }
}
__device__ __forceinline__ void dh_produce(const DhLoad& L, int ptid, uint8_t* big, uint8_t* small, int single) {
  const int ku = ptid & 7, fg = ptid >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float v[4] = {L.t[0][c], L.t[1][c], L.t[2][c], L.t[3][c]};
    store_unit(big, small, fg * 4 + c, ku, v, single);
  }
}
__device__ __forceinline__ void dh_tile(const DhParams& p, int t, int& i0, int& o0, int& split, int& kbeg, int& kend) {
  const int per = p.tiles_m * p.tiles_n;
  split = t / per;
  const int r = t % per;
  i0 = (r / p.tiles_n) * BM;
  o0 = (r % p.tiles_n) * BN;
  kbeg = split * p.rows_per_split;
  kend = min(kbeg + p.rows_per_split, p.M);
}

__global__ void __launch_bounds__(THREADS, 1) phm_tc_dh_kernel(const DhParams p) {
  pdl_begin();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const Smem s = carve(smem_raw, 0);
  const uint32_t tmem_base = cta_prologue(s, PROD_WARPS);
  const int warp = threadIdx.x >> 5;

  if (warp < EPI_WARP0) {
    // 512 producer threads: the first 256 build the A operand (x^T), the other 256 the B operand (dy^T)
    static_assert(PROD_THREADS == 512, "dH producer mapping assumes 16 producer warps");
    const int ptid = threadIdx.x - PROD_WARP0 * 32;
    const int half = ptid >> 8, q = ptid & 255;
    int stage = 0, phase = 0;
    const float* src = half == 0 ? p.X : p.G;
    const int ld = half == 0 ? p.In : p.Out;
    const bool vec = ld % 4 == 0 && aligned16(src);
    DhLoad cur, nxt;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      const int f0 = half == 0 ? i0 : o0;
      dh_issue_loads(src, ld, vec, f0, kbeg, kend, q, nxt);
      for (int k0 = kbeg; k0 < kend; k0 += BK) {
        cur = nxt;
        if (k0 + BK < kend) dh_issue_loads(src, ld, vec, f0, k0 + BK, kend, q, nxt);
        mbar_wait(smem_u32(&s.empty_bar[stage]), phase ^ 1);
        uint8_t* sb = s.stages + stage * STAGE_BYTES + half * 2 * TILE_BYTES;
        dh_produce(cur, q, sb, sb + TILE_BYTES, p.single);
        fence_proxy_async();
        __syncwarp();
        if ((ptid & 31) == 0) mbar_arrive(smem_u32(&s.full_bar[stage]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      mma_tile(s, tmem_base, it, (kend - kbeg + BK - 1) / BK, stage, phase, nullptr, p.single);
    }
  } else {
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      epilogue_tile<false>(s, tmem_base, it, p.C + (size_t)split * p.In * p.Out, p.Out, p.In, i0, o0, min(BN, p.Out - o0), nullptr,
                           nullptr, 0);
    }
  }
  cta_epilogue(tmem_base);
}

// ---------------------------------------------------------------------------- dH kernel, TMA-staged
// Raw [32 samples x 128 features] boxes of x and dy arrive by TMA; the producers transpose them out of shared
// memory (4 strided LDS per 16-byte K-unit, conflict-free) into the K-major UMMA operands.
constexpr int DH_RAW_STAGES = 2;
constexpr int DH_RAW_BYTES = 2 * BK * BM * 4;     // x box + dy box = 32 KiB
constexpr int DH_TMA_THREADS = (MMA_WARP + 2) * 32;

__global__ void __launch_bounds__(DH_TMA_THREADS, 1) phm_tc_dh_tma_kernel(const DhParams p, const __grid_constant__ CUtensorMap tmapX,
                                                                         const __grid_constant__ CUtensorMap tmapG) {
  pdl_begin();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.stages = base;
  uint8_t* rawbuf = base + TMA_OP_STAGES * STAGE_BYTES;
  s.scratch = reinterpret_cast<float*>(rawbuf + DH_RAW_STAGES * DH_RAW_BYTES);
  s.coef = s.scratch + EPI_WARPS * 32 * 33;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.coef);
  s.full_bar = bars;                         // [2]
  s.empty_bar = bars + 2;                    // [2]
  uint64_t* rfull_bar = bars + 4;            // [2]
  uint64_t* rempty_bar = bars + 6;           // [2]
  s.tfull_bar = bars + 8;
  s.tempty_bar = bars + 10;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TMA_OP_STAGES; ++i) {
      mbar_init(smem_u32(&s.full_bar[i]), PROD_WARPS);
      mbar_init(smem_u32(&s.empty_bar[i]), 1);
    }
    for (int i = 0; i < DH_RAW_STAGES; ++i) {
      mbar_init(smem_u32(&rfull_bar[i]), 1);
      mbar_init(smem_u32(&rempty_bar[i]), PROD_WARPS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&s.tfull_bar[b]), 1);
      mbar_init(smem_u32(&s.tempty_bar[b]), EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(s.tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;

  if (warp < EPI_WARP0) {
    const int ptid = threadIdx.x;
    const int half = ptid >> 8, q = ptid & 255;          // first 256 threads: A (x^T), others: B (dy^T)
    int g = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      float colsum = 0.f;      // dy producers: this thread's share of the column sum of feature o0 + (q & 127) (bias gradient)
      for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
        const int rs = g % DH_RAW_STAGES, os = g % TMA_OP_STAGES;
        mbar_wait(smem_u32(&rfull_bar[rs]), (g / DH_RAW_STAGES) & 1);
        mbar_wait(smem_u32(&s.empty_bar[os]), ((g / TMA_OP_STAGES) & 1) ^ 1);
        const float* raw = reinterpret_cast<const float*>(rawbuf + rs * DH_RAW_BYTES + half * (BK * BM * 4));
        uint8_t* sb = s.stages + os * STAGE_BYTES + half * 2 * TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = i * 256 + q;
          const int f = u & (BM - 1), ku = u >> 7;
          const float v[4] = {raw[(ku * 4 + 0) * BM + f], raw[(ku * 4 + 1) * BM + f], raw[(ku * 4 + 2) * BM + f],
                              raw[(ku * 4 + 3) * BM + f]};
          store_unit(sb, sb + TILE_BYTES, f, ku, v, p.single);
          colsum += (v[0] + v[1]) + (v[2] + v[3]);              // rows past M are zero-filled by TMA
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&s.full_bar[os]));
          mbar_arrive(smem_u32(&rempty_bar[rs]));
        }
      }
      // bias gradient for free: the dy tile passes through these threads anyway.  One i-tile per (o-tile, split) writes;
      // threads q and q + 128 hold the two halves of a column's sum (k-units 2i and 2i + 1), summed later in fixed order.
      if (half == 1 && p.db_part != nullptr && i0 == 0 && o0 + (q & (BN - 1)) < p.Out)
        p.db_part[((size_t)split * 2 + (q >> 7)) * p.Out + o0 + (q & (BN - 1))] = colsum;
    }
  } else if (warp == MMA_WARP + 1) {
    if (elect_one()) {
      int g = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int i0, o0, split, kbeg, kend;
        dh_tile(p, t, i0, o0, split, kbeg, kend);
        for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
          const int rs = g % DH_RAW_STAGES;
          mbar_wait(smem_u32(&rempty_bar[rs]), ((g / DH_RAW_STAGES) & 1) ^ 1);
          const uint32_t bar = smem_u32(&rfull_bar[rs]);
          mbar_arrive_expect_tx(bar, DH_RAW_BYTES);
          tma_load_2d(smem_u32(rawbuf + rs * DH_RAW_BYTES), &tmapX, i0, k0, bar);
          tma_load_2d(smem_u32(rawbuf + rs * DH_RAW_BYTES + BK * BM * 4), &tmapG, o0, k0, bar);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      mma_tile<TMA_OP_STAGES>(s, tmem_base, it, (kend - kbeg + BK - 1) / BK, stage, phase, nullptr, p.single);
    }
  } else {
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int i0, o0, split, kbeg, kend;
      dh_tile(p, t, i0, o0, split, kbeg, kend);
      epilogue_tile<false>(s, tmem_base, it, p.C + (size_t)split * p.In * p.Out, p.Out, p.In, i0, o0, min(BN, p.Out - o0), nullptr,
                           nullptr, 0);
    }
  }
  cta_epilogue(tmem_base);
}

// ---------------------------------------------------------------------------- dH kernel v2: x^T operand in TENSOR MEMORY,
// two dy tiles per x tile.
// The TMA-staged kernel above sits at its SHARED-MEMORY roofline: both operands are transposed through shared memory
// (per 32-sample chunk: 32 KiB TMA in, 32 KiB read, 64 KiB operand tiles written, 96 KiB read by the 12 MMAs = 292 B/clk
// against 128 B/clk).  Here one CTA computes a 128 x 256 block of dH: the x^T operand (tile row = feature = TMEM lane)
// is written by 128 producer threads straight into tensor memory (tcgen05.st) and serves TWO 128-column dy tiles; only
// the dy^T operand goes through shared memory: 48 + 48 + 64 + 96 KiB per 24 MMAs = 167 B/clk.  A producer thread owns
// one feature row, so the transpose is a conflict-free column read of the raw box.
// (Tried and rejected: loading the operands straight from global memory with a one-chunk register prefetch instead of
//  TMA staging — 57 us against 45 us, the prefetch is too shallow for DRAM latency.)
//   warps 0-3 x^T -> TMEM | 4-11 dy^T -> smem (and the column sums of dy: bias gradient) | 12-15 epilogue | 16 MMA | 17 TMA
//   TMEM: 2 accumulators x 128 columns | 2 chunk parities x (32 big + 32 small) operand columns
constexpr int DH2_BN = 2 * BN;
constexpr int DH2_RAW_X = BK * BM * 4, DH2_RAW_G = BK * DH2_BN * 4;      // 16 KiB + 32 KiB per raw stage
constexpr int DH2_RAW_BYTES = DH2_RAW_X + DH2_RAW_G;
constexpr int DH2_B_STAGE = 4 * TILE_BYTES;                              // two dy tiles x (big, small) = 64 KiB
constexpr int DH2_A_WARPS = 4, DH2_B_WARPS = 8, DH2_EPI_WARP0 = 12, DH2_MMA_WARP = 16, DH2_TMA_WARP = 17;
constexpr int DH2_THREADS = 18 * 32;
constexpr int DH2_A_COL0 = 2 * BN;
constexpr int DH2_BARS = 14;
static_assert(DH2_EPI_WARP0 % 4 == 0, "epilogue warps must be aligned to the TMEM lane quarters");

__device__ __forceinline__ void dh2_tile(const DhParams& p, int t, int& i0, int& o0, int& split, int& kbeg, int& kend) {
  const int per = p.tiles_m * p.tiles_n;                 // tiles_n counts 256-column tiles here
  split = t / per;
  const int r = t % per;
  i0 = (r / p.tiles_n) * BM;
  o0 = (r % p.tiles_n) * DH2_BN;
  kbeg = split * p.rows_per_split;
  kend = min(kbeg + p.rows_per_split, p.M);
}

__global__ void __launch_bounds__(DH2_THREADS, 1) phm_tc_dh_v2_kernel(const DhParams p, const __grid_constant__ CUtensorMap tmapX,
                                                                     const __grid_constant__ CUtensorMap tmapG) {
  pdl_begin();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
#if PHC_SMEM_SPACE
  // 1 KiB alignment by POINTER arithmetic on the shared array: the detour through uintptr_t made every pointer derived from `base`
  // generic, and the drain's transpose scratch, the rule table and the barrier words were read with LD.E / ST.E instead of LDS / STS
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#else
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#endif
  uint8_t* bstage = base;                                        // [2][tile0 big | tile0 small | tile1 big | tile1 small]
  uint8_t* rawbuf = base + 2 * DH2_B_STAGE;                      // [2][x box 32x128 | dy box 32x256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawbuf + 2 * DH2_RAW_BYTES);
  uint64_t* rfull = bars;           // [2] TMA transaction
  uint64_t* rempty = bars + 2;      // [2] 12 producer warps
  uint64_t* afull = bars + 4;       // [2] 4 x^T warps
  uint64_t* aempty = bars + 6;      // [2] tcgen05.commit
  uint64_t* bfull = bars + 8;       // [2] 8 dy^T warps
  uint64_t* bempty = bars + 10;     // [2] tcgen05.commit
  uint64_t* tfull = bars + 12;
  uint64_t* tempty = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + DH2_BARS);
  uint32_t* landed = tmem_slot + 4;                              // [12 producer warps][32]: see the early release below
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&rfull[i]), 1);
      mbar_init(smem_u32(&rempty[i]), DH2_A_WARPS + DH2_B_WARPS);
      mbar_init(smem_u32(&afull[i]), DH2_A_WARPS);
      mbar_init(smem_u32(&aempty[i]), 1);
      mbar_init(smem_u32(&bfull[i]), DH2_B_WARPS);
      mbar_init(smem_u32(&bempty[i]), 1);
    }
    mbar_init(smem_u32(tfull), 1);
    mbar_init(smem_u32(tempty), 4 * 32);
    fence_barrier_init();
  }
  if (warp == DH2_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < DH2_A_WARPS) {
    // ===================================================================== x^T -> tensor memory (thread = feature row = lane)
    const int f = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int g = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int i0, o0, split, kbeg, kend;
      dh2_tile(p, t, i0, o0, split, kbeg, kend);
      for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
        const int rs = g & 1, ph = (g >> 1) & 1;
        mbar_wait(smem_u32(&rfull[rs]), ph);
        const uint32_t raw = smem_u32(rawbuf + rs * DH2_RAW_BYTES) + (uint32_t)f * 4u;
        float v[BK];
        uint32_t lor = 0;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) { v[kk] = lds_f32(raw + (uint32_t)(kk * BM * 4)); lor |= __float_as_uint(v[kk]); }
        // Early release: the column is in registers, let TMA refill the raw stage while we wait for the operand slot.
        // The arrive must not overtake the loads: a shared-memory store of a value that depends on every load cannot
        // issue before they have returned, and the arrive (release) stays behind the store.
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(landed + warp * 32 + lane)), "r"(lor) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&rempty[rs]));
        mbar_wait(smem_u32(&aempty[rs]), ph ^ 1);                 // the MMAs that read this operand slot have retired
        tc_fence_after();
        const uint32_t slot = tmem_base + lane_base + (uint32_t)(DH2_A_COL0 + rs * 64);
#pragma unroll
        for (int c8 = 0; c8 < BK; c8 += 8) {
          float big[8], small[8];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            if (p.single) {
              big[j] = __uint_as_float((__float_as_uint(v[c8 + j]) + 0x8000u) & 0xFFFF0000u);
              big[j + 1] = __uint_as_float((__float_as_uint(v[c8 + j + 1]) + 0x8000u) & 0xFFFF0000u);
              small[j] = small[j + 1] = 0.f;
            } else {                                   // exact split, two elements per instruction (64-bit mask, FFMA2): the tensor
              const uint64_t v2 = pack_f32x2(v[c8 + j], v[c8 + j + 1]);               // core ignores the low 13 bits of `small`
              const uint64_t b2 = v2 & 0xFFFFE000FFFFE000ull;
              unpack_f32x2(b2, big[j], big[j + 1]);
              unpack_f32x2(fma_f32x2(b2, bcast_f32x2(-1.f), v2), small[j], small[j + 1]);
            }
          }
          tmem_st8(slot + c8, big);
          tmem_st8(slot + 32 + c8, small);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&afull[rs]));
      }
    }
  } else if (warp < DH2_A_WARPS + DH2_B_WARPS) {
    // ===================================================================== dy^T -> shared memory (thread = feature row of 256)
    const int q = threadIdx.x - DH2_A_WARPS * 32;                 // 0..255
    const int tsel = q >> 7, row = q & (BN - 1);
    int g = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int i0, o0, split, kbeg, kend;
      dh2_tile(p, t, i0, o0, split, kbeg, kend);
      float colsum = 0.f;                                          // column sum of feature o0 + q over the split: bias gradient
      for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
        const int rs = g & 1, ph = (g >> 1) & 1;
        mbar_wait(smem_u32(&rfull[rs]), ph);
        const uint32_t raw = smem_u32(rawbuf + rs * DH2_RAW_BYTES + DH2_RAW_X) + (uint32_t)q * 4u;
        float v[BK];
        uint32_t lor = 0;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) { v[kk] = lds_f32(raw + (uint32_t)(kk * DH2_BN * 4)); lor |= __float_as_uint(v[kk]); }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(landed + warp * 32 + lane)), "r"(lor) : "memory");   // early release, as above
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&rempty[rs]));
        mbar_wait(smem_u32(&bempty[rs]), ph ^ 1);
        const uint32_t tb = smem_u32(bstage + rs * DH2_B_STAGE + tsel * 2 * TILE_BYTES);
#pragma unroll
        for (int ku = 0; ku < 8; ++ku) {
          float b[4], sm[4];
          colsum += (v[ku * 4] + v[ku * 4 + 1]) + (v[ku * 4 + 2] + v[ku * 4 + 3]);     // rows past M are zero-filled by TMA
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            if (p.single) {
              b[j] = __uint_as_float((__float_as_uint(v[ku * 4 + j]) + 0x8000u) & 0xFFFF0000u);
              b[j + 1] = __uint_as_float((__float_as_uint(v[ku * 4 + j + 1]) + 0x8000u) & 0xFFFF0000u);
              sm[j] = sm[j + 1] = 0.f;
            } else {
              const uint64_t v2 = pack_f32x2(v[ku * 4 + j], v[ku * 4 + j + 1]);
              const uint64_t b2 = v2 & 0xFFFFE000FFFFE000ull;
              unpack_f32x2(b2, b[j], b[j + 1]);
              unpack_f32x2(fma_f32x2(b2, bcast_f32x2(-1.f), v2), sm[j], sm[j + 1]);
            }
          }
          const uint32_t off = (uint32_t)swz(row, ku);
          sts_v4(tb + off, b[0], b[1], b[2], b[3]);
          sts_v4(tb + TILE_BYTES + off, sm[0], sm[1], sm[2], sm[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bfull[rs]));
      }
      if (p.db_part != nullptr && i0 == 0 && o0 + q < p.Out) p.db_part[(size_t)split * p.Out + o0 + q] = colsum;
    }
  } else if (warp == DH2_TMA_WARP) {
    if (elect_one()) {
      int g = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int i0, o0, split, kbeg, kend;
        dh2_tile(p, t, i0, o0, split, kbeg, kend);
        for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
          const int rs = g & 1;
          mbar_wait(smem_u32(&rempty[rs]), ((g >> 1) & 1) ^ 1);
          const uint32_t bar = smem_u32(&rfull[rs]);
          mbar_arrive_expect_tx(bar, DH2_RAW_BYTES);
          tma_load_2d(smem_u32(rawbuf + rs * DH2_RAW_BYTES), &tmapX, i0, k0, bar);
          tma_load_2d(smem_u32(rawbuf + rs * DH2_RAW_BYTES + DH2_RAW_X), &tmapG, o0, k0, bar);
        }
      }
    }
  } else if (warp == DH2_MMA_WARP) {
    // One elected thread runs the whole issue loop (elect.sync, not `lane == 0`: the compiler then keeps descriptors and tensor-memory
    // addresses in uniform registers instead of a per-MMA broadcast loop), spins on test_wait, and probes the next chunk's
    // barriers before the last MMAs of the current one — see the mix kernel's issue loop.
    if (elect_one()) {
      int g = 0, it = 0;
      bool ready = false;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        int i0, o0, split, kbeg, kend;
        dh2_tile(p, t, i0, o0, split, kbeg, kend);
        mbar_spin(smem_u32(tempty), (it & 1) ^ 1);                // accumulators drained
        uint32_t accum = 0;
        for (int k0 = kbeg; k0 < kend; k0 += BK, ++g) {
          const int rs = g & 1, ph = (g >> 1) & 1;
          if (!ready) {
            mbar_spin(smem_u32(&afull[rs]), ph);
            mbar_spin(smem_u32(&bfull[rs]), ph);
          }
          tc_fence_after();
          ready = false;
          const uint32_t a_big = tmem_base + (uint32_t)(DH2_A_COL0 + rs * 64), a_small = a_big + 32;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t sb = smem_u32(bstage + rs * DH2_B_STAGE + h * 2 * TILE_BYTES);
            const uint64_t b_big = make_desc(sb), b_small = make_desc(sb + TILE_BYTES);
            const uint32_t d_tmem = tmem_base + h * BN;
            uint32_t acc = accum;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);
              if (h == 1 && ks == BK / 8 - 1 && (k0 + BK < kend || t + (int)gridDim.x < p.num_tiles)) {
                const int gn = g + 1;                            // next chunk of this CTA (this tile's or the next one's)
                ready = mbar_test_wait(smem_u32(&afull[gn & 1]), (gn >> 1) & 1) && mbar_test_wait(smem_u32(&bfull[gn & 1]), (gn >> 1) & 1);
              }
              if (p.single) {
                umma_tf32_ts(d_tmem, a_big + ks * 8, b_big + adv, IDESC_TF32, acc);
              } else {
                umma_tf32_ts(d_tmem, a_small + ks * 8, b_big + adv, IDESC_TF32, acc);
                umma_tf32_ts(d_tmem, a_big + ks * 8, b_small + adv, IDESC_TF32, 1u);
                umma_tf32_ts(d_tmem, a_big + ks * 8, b_big + adv, IDESC_TF32, 1u);
              }
              acc = 1u;
            }
          }
          accum = 1u;
          umma_commit(smem_u32(&aempty[rs]));
          umma_commit(smem_u32(&bempty[rs]));
        }
        umma_commit(smem_u32(tfull));
      }
    }
  } else {
    // ===================================================================== epilogue: TMEM -> registers -> 16-byte row stores
    const int e = warp - DH2_EPI_WARP0;                            // TMEM lane quarter
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int i0, o0, split, kbeg, kend;
      dh2_tile(p, t, i0, o0, split, kbeg, kend);
      mbar_wait(smem_u32(tfull), it & 1);
      tc_fence_after();
      const int rowi = i0 + e * 32 + lane;
      float* dst = p.C + (size_t)split * p.In * p.Out + (size_t)rowi * p.Out + o0;
#pragma unroll 1
      for (int cc = 0; cc < DH2_BN / 32; ++cc) {
        if (o0 + cc * 32 >= p.Out) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(cc * 32), v);
        if (rowi < p.In) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (o0 + cc * 32 + j * 4 < p.Out)                      // Out % 4 == 0 on this path: whole float4 in or out
              *reinterpret_cast<float4*>(dst + cc * 32 + j * 4) =
                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(tempty));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DH2_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------- pack kernel
// Writes, for one direction, the rule re-indexed per output component and the W operand as tf32 big/small
// tile images:  Bpack[pt][chunk][half][q][32]  with element (q, s) = Wsrc(b = s % n, kk = s / n, p = pt*128 + q).
//   dir 0 (FWD): reduction index kk = k, output index q = p : value W[b,k,p];  coef[c][b][a] = A[b,a,c]
//   dir 1 (DX) : reduction index kk = p, output index q = k : value W[b,k,p];  coef[a][b][c] = A[b,a,c]
__global__ void __launch_bounds__(256) phm_pack_kernel(const float* __restrict__ A, const float* __restrict__ W, int n, int K, int P,
                                                       uint8_t* __restrict__ pack_fwd, float* __restrict__ coef_fwd,
                                                       uint8_t* __restrict__ pack_dx, float* __restrict__ coef_dx, int units_fwd,
                                                       int units_dx, int single, unsigned int* __restrict__ sk_flags, int n_flags) {
  pdl_begin();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n3 = n * n * n;
  if (sk_flags != nullptr && t < n_flags) sk_flags[t] = 0u;       // stream-K "partial written" flags of the mix kernel
  if (t < n3) {
    const int b = t / (n * n), a = (t / n) % n, c = t % n;
    const float v = A[t];
    coef_fwd[(c * n + b) * n + a] = v;
    coef_dx[(a * n + b) * n + c] = v;
  }
  for (int dir = 0; dir < 2; ++dir) {
    const int units = dir == 0 ? units_fwd : units_dx;
    if (t >= units) continue;
    const int Kin = dir == 0 ? K : P, Pout = dir == 0 ? P : K;
    const int S = n * Kin, chunks = (S + BK - 1) / BK;
    // unit id -> (pt, chunk, q, ku)
    const int ku = t & 7, q = (t >> 3) & (BN - 1);
    const int rest = t >> 10;
    const int chunk = rest % chunks, pt = rest / chunks;
    const int po = pt * BN + q;
    float big[4], small[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sidx = chunk * BK + ku * 4 + j;
      float v = 0.f;
      if (po < Pout && sidx < S) {
        const int kk = sidx / n, b = sidx - kk * n;
        v = dir == 0 ? W[((size_t)b * K + kk) * P + po] : W[((size_t)b * K + po) * P + kk];
      }
      if (single) { big[j] = round_bf16(v); small[j] = 0.f; }
      else split_tf32(v, big[j], small[j]);
    }
    uint8_t* dst = (dir == 0 ? pack_fwd : pack_dx) + ((size_t)pt * chunks + chunk) * (2 * TILE_BYTES);
    const int off = swz(q, ku);
    *reinterpret_cast<float4*>(dst + off) = make_float4(big[0], big[1], big[2], big[3]);
    *reinterpret_cast<float4*>(dst + TILE_BYTES + off) = make_float4(small[0], small[1], small[2], small[3]);
  }
}

// ---------------------------------------------------------------------------- host side
size_t smem_bytes(int coef_floats) {
  return 1024 + (size_t)STAGES * STAGE_BYTES + EPI_SCRATCH + sizeof(float) * ((coef_floats + 3) & ~3) + 8 * (2 * STAGES + 4) + 16;
}

// Stream-K split of the n = 4 mix kernel's unit list (v3_range): balances the CTAs (244 units on 148 CTAs at ppa shape) but
// every split unit costs two extra accumulator drains (park + add), and the drain is not overlapped with the MMAs because
// tensor memory is full.  Measured on B200 at ppa shape: 61.7 us with the split against 50.1 us without — opt-in only
// (build with -DPHC_TC_STREAM_K=1 and set PHC_TC_STREAM_K=1 in the environment).
bool stream_k_enabled() {
  static const bool on = PHC_TC_STREAM_K && getenv("PHC_TC_STREAM_K") != nullptr;
  return on;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int dh_splits(int M, int In, int Out, int* rows_per_split) {
  const int tiles = phc_div_up(In, BM) * phc_div_up(Out, BN);
  int s = num_sms() / tiles;
  const int maxs = phc_div_up(M, 4 * BK);
  s = s > maxs ? maxs : s;
  s = s < 1 ? 1 : (s > 64 ? 64 : s);
  const int rps = phc_div_up(phc_div_up(M, s), BK) * BK;
  if (rows_per_split) *rows_per_split = rps;
  return phc_div_up(M, rps);               // every split owns at least one K chunk
}

int dh2_splits(int M, int In, int Out, int* rows_per_split) {
  const int tiles = phc_div_up(In, BM) * phc_div_up(Out, DH2_BN);
  int s = num_sms() / tiles;
  const int maxs = phc_div_up(M, 4 * BK);
  s = s > maxs ? maxs : s;
  s = s < 1 ? 1 : (s > 64 ? 64 : s);
  const int rps = phc_div_up(phc_div_up(M, s), BK) * BK;
  if (rows_per_split) *rows_per_split = rps;
  return phc_div_up(M, rps);
}

struct PackLayout {
  size_t coef_fwd, coef_dx, pack_fwd, pack_dx, total;   // byte offsets into the pack buffer
  size_t sk_flags, sk_part;                             // stream-K scratch of the n = 4 mix kernel (0: none), inside `total`
  int chunks_fwd, chunks_dx, pt_fwd, pt_dx;
};
constexpr size_t SK_PART_BYTES = (size_t)PROD_WARPS * 32 * 64 * 4;     // one CTA's parked accumulators (128 KiB)

PackLayout pack_layout(int n, int K, int P) {
  PackLayout L;
  const size_t n3b = ((size_t)n * n * n * 4 + 1023) & ~(size_t)1023;
  L.chunks_fwd = phc_div_up(n * K, BK); L.pt_fwd = phc_div_up(P, BN);
  L.chunks_dx = phc_div_up(n * P, BK);  L.pt_dx = phc_div_up(K, BN);
  L.coef_fwd = 0;
  L.coef_dx = n3b;
  L.pack_fwd = 2 * n3b;
  L.pack_dx = L.pack_fwd + (size_t)L.pt_fwd * L.chunks_fwd * 2 * TILE_BYTES;
  L.total = L.pack_dx + (size_t)L.pt_dx * L.chunks_dx * 2 * TILE_BYTES;
  L.sk_flags = L.sk_part = 0;
  if (n == 4 && stream_k_enabled()) {
    const size_t ctas = (size_t)num_sms() + 1;
    L.sk_flags = L.total;
    L.sk_part = (L.sk_flags + ctas * PROD_WARPS * 4 + 1023) & ~(size_t)1023;
    L.total = L.sk_part + ctas * SK_PART_BYTES;
  }
  return L;
}

template <typename Kern>
int set_smem(Kern kern, bool* done) {
  if (*done) return PHC_OK;
  const size_t smem = smem_bytes(16 * 16 * 16);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    phc_set_error("phm_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PHC_ERR_CUDA;
  }
  *done = true;
  return PHC_OK;
}

template <int NT>
int launch_mix_nt(const MixParams& p, cudaStream_t stream) {
  static bool configured = false;
  int rc = set_smem(phm_tc_mix_kernel<NT>, &configured);
  if (rc) return rc;
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  phc_launch(phm_tc_mix_kernel<NT>, dim3(grid), dim3(THREADS), smem_bytes(p.n * p.n * p.n), stream, p);
  return phc_check_launch("phm_tc_mix_kernel");
}

int launch_mix(const MixParams& p, cudaStream_t stream, int* v3_used = nullptr);
int launch_mix_v2(const MixParams& p, cudaStream_t stream) {
  switch (p.n) {
    case 1: return launch_mix_nt<1>(p, stream);
    case 2: return launch_mix_nt<2>(p, stream);
    case 4: return launch_mix_nt<4>(p, stream);
    default: return launch_mix_nt<0>(p, stream);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

size_t smem_bytes_tma(int nt) {
  return 1024 + (size_t)TMA_OP_STAGES * STAGE_BYTES + (size_t)TMA_RAW_STAGES * RAW_BYTES + EPI_SCRATCH +
         sizeof(float) * ((nt * nt * nt + 3) & ~3) + 8 * 16 + 16;
}

template <int NT>
int launch_mix_tma_nt(const MixParams& p, const CUtensorMap& tmap, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(phm_tc_mix_tma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_tma(NT));
    if (e != cudaSuccess) {
      phc_set_error("phm_tc: cannot reserve %zu bytes of shared memory: %s", smem_bytes_tma(NT), cudaGetErrorString(e));
      return PHC_ERR_CUDA;
    }
    configured = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  phc_launch(phm_tc_mix_tma_kernel<NT>, dim3(grid), dim3(TMA_THREADS), smem_bytes_tma(NT), stream, p, tmap);
  return phc_check_launch("phm_tc_mix_tma_kernel");
}

// TMA path: n in {1,2,4}, 16-byte aligned rows.  Returns -1 when not applicable (caller falls back).
int try_launch_mix_tma(const MixParams& p, cudaStream_t stream) {
  const int n = p.n, Fin = p.n * p.Kin;
  if (!(n == 1 || n == 2 || n == 4) || Fin % 4 != 0 || (reinterpret_cast<uintptr_t>(p.X) & 15u) != 0) return -1;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return -1;
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)Fin, (cuuint64_t)p.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)Fin * 4};
  const cuuint32_t box[2] = {(cuuint32_t)(BK / n + ((p.Kin & 3) ? 4 : 0)), (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.X), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -1;
  switch (n) {
    case 1: return launch_mix_tma_nt<1>(p, tmap, stream);
    case 2: return launch_mix_tma_nt<2>(p, tmap, stream);
    default: return launch_mix_tma_nt<4>(p, tmap, stream);
  }
}

size_t smem_bytes_v3() {
  return 1024 + (size_t)V3_NB * 2 * TILE_BYTES + (size_t)V3_NRAW * V3_RAW_BYTES + V3_SCRATCH + sizeof(float) * 64 + 8 * V3_BARS + 16;
}

template <int R, bool SINGLE, bool STATS>
int launch_mix_v3_rs(const MixParams& p, const CUtensorMap& tmap, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(phm_tc_mix_v3_kernel<R, SINGLE, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_v3());
    if (e != cudaSuccess) {
      phc_set_error("phm_tc: cannot reserve %zu bytes of shared memory: %s", smem_bytes_v3(), cudaGetErrorString(e));
      return PHC_ERR_CUDA;
    }
    configured = true;
  }
  MixParams q = p;
  q.num_tiles = phc_div_up(p.M, BM) * 2 * p.ptiles;           // units: (m-tile, p-tile, component pair)
  const int grid = q.num_tiles < num_sms() ? q.num_tiles : num_sms();
  phc_launch(phm_tc_mix_v3_kernel<R, SINGLE, STATS>, dim3(grid), dim3(V3_THREADS), smem_bytes_v3(), stream, q, tmap);
  return phc_check_launch("phm_tc_mix_v3_kernel");
}

template <int R>
int launch_mix_v3_r(const MixParams& p, const CUtensorMap& tmap, cudaStream_t stream) {
  if (p.stat_part != nullptr)
    return p.single ? launch_mix_v3_rs<R, true, true>(p, tmap, stream) : launch_mix_v3_rs<R, false, true>(p, tmap, stream);
  return p.single ? launch_mix_v3_rs<R, true, false>(p, tmap, stream) : launch_mix_v3_rs<R, false, false>(p, tmap, stream);
}

// TMEM-operand path: n == 4, 16-byte aligned rows.  Returns -1 when not applicable (caller falls back).
int try_launch_mix_v3(const MixParams& p, cudaStream_t stream) {
  static const bool use_v3 = getenv("PHC_TC_NO_TMEM_A") == nullptr;
  static const int ablate = getenv("PHC_TC_ABLATE") ? atoi(getenv("PHC_TC_ABLATE")) : 0;
  const_cast<MixParams&>(p).ablate = ablate;
  const int Fin = p.n * p.Kin;
  if (!use_v3 || p.n != 4 || Fin % 4 != 0 || (reinterpret_cast<uintptr_t>(p.X) & 15u) != 0) return -1;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return -1;
  CUtensorMap tmap;
  const int R = p.Kin & 3;
  const cuuint64_t gdim[2] = {(cuuint64_t)Fin, (cuuint64_t)p.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)Fin * 4};
  const cuuint32_t box[2] = {(cuuint32_t)(BK / 4 + (R ? 4 : 0)), (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.X), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -1;
  switch (R) {
    case 0: return launch_mix_v3_r<0>(p, tmap, stream);
    case 1: return launch_mix_v3_r<1>(p, tmap, stream);
    case 2: return launch_mix_v3_r<2>(p, tmap, stream);
    default: return launch_mix_v3_r<3>(p, tmap, stream);
  }
}

int launch_pack(const float* A, const float* W, int n, int K, int P, uint8_t* buf, int single, cudaStream_t stream) {
  const PackLayout L = pack_layout(n, K, P);
  const int units_fwd = L.pt_fwd * L.chunks_fwd * BN * 8, units_dx = L.pt_dx * L.chunks_dx * BN * 8;
  int threads = units_fwd > units_dx ? units_fwd : units_dx;
  if (threads < n * n * n) threads = n * n * n;
  const int n_flags = L.sk_part ? (num_sms() + 1) * PROD_WARPS : 0;
  if (threads < n_flags) threads = n_flags;
  phc_launch(phm_pack_kernel, dim3(phc_div_up(threads, 256)), dim3(256), 0, stream, A, W, n, K, P, buf + L.pack_fwd, reinterpret_cast<float*>(buf + L.coef_fwd),
                                                                buf + L.pack_dx, reinterpret_cast<float*>(buf + L.coef_dx), units_fwd,
                                                                units_dx, single, L.sk_part ? reinterpret_cast<unsigned int*>(buf + L.sk_flags) : nullptr, n_flags);
  return phc_check_launch("phm_pack_kernel");
}

bool encode_2d(CUtensorMap* tmap, const float* ptr, int rows, int cols, int box_cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || cols % 4 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t smem_bytes_dh_tma() {
  return 1024 + (size_t)TMA_OP_STAGES * STAGE_BYTES + (size_t)DH_RAW_STAGES * DH_RAW_BYTES + EPI_SCRATCH + 8 * 16 + 16;
}

size_t smem_bytes_dh_v2() { return 1024 + 2 * (size_t)DH2_B_STAGE + 2 * (size_t)DH2_RAW_BYTES + 8 * DH2_BARS + 16 + 12 * 32 * 4; }

// TMEM-operand dH kernel; `d` must carry the wide tiling (tiles_n = 256-column tiles, dh2_splits).  -1: not applicable.
int try_launch_dh_v2(const DhParams& d, cudaStream_t stream) {
  static const bool use_v2 = getenv("PHC_TC_NO_DH_V2") == nullptr && getenv("PHC_TC_NO_TMA") == nullptr;
  if (!use_v2 || d.Out % 4 != 0 || (reinterpret_cast<uintptr_t>(d.C) & 15u) != 0) return -1;
  CUtensorMap tx, tg;
  if (!encode_2d(&tx, d.X, d.M, d.In, BM, BK) || !encode_2d(&tg, d.G, d.M, d.Out, DH2_BN, BK)) return -1;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(phm_tc_dh_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_dh_v2()) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    configured = true;
  }
  const int grid = d.num_tiles < num_sms() ? d.num_tiles : num_sms();
  phc_launch(phm_tc_dh_v2_kernel, dim3(grid), dim3(DH2_THREADS), smem_bytes_dh_v2(), stream, d, tx, tg);
  return phc_check_launch("phm_tc_dh_v2_kernel");
}

// returns -1 when the TMA path is not applicable
int try_launch_dh_tma(const DhParams& d, cudaStream_t stream) {
  static const bool use_tma = getenv("PHC_TC_NO_TMA") == nullptr;
  if (!use_tma) return -1;
  CUtensorMap tx, tg;
  if (!encode_2d(&tx, d.X, d.M, d.In, BM, BK) || !encode_2d(&tg, d.G, d.M, d.Out, BN, BK)) return -1;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(phm_tc_dh_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_dh_tma()) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    configured = true;
  }
  const int grid = d.num_tiles < num_sms() ? d.num_tiles : num_sms();
  phc_launch(phm_tc_dh_tma_kernel, dim3(grid), dim3(DH_TMA_THREADS), smem_bytes_dh_tma(), stream, d, tx, tg);
  return phc_check_launch("phm_tc_dh_tma_kernel");
}

int launch_mix(const MixParams& p, cudaStream_t stream, int* v3_used) {
  static const bool use_tma = getenv("PHC_TC_NO_TMA") == nullptr;       // debug switch: force the register-prefetch kernel
  if (v3_used) *v3_used = 0;
  if (use_tma) {
    int rc = try_launch_mix_v3(p, stream);
    if (rc >= 0) {
      if (v3_used) *v3_used = rc == 0;
      return rc;
    }
    rc = p.prof == nullptr ? try_launch_mix_tma(p, stream) : -1;
    if (rc >= 0) return rc;
  }
  return launch_mix_v2(p, stream);
}

}  // namespace tc

// ---------------------------------------------------------------------------- host entry points (used by api.cu)
int phm_tc_supported(int rows, int in_features, int out_features, int phm_dim, int precision) {
  // tf32x3 only for now; small problems (head layers, M = graphs per batch) stay on the exact FFMA path
  // (tiny M is latency-bound either way; the tensor-core kernel needs fewer serial K iterations than the FFMA tile)
  static const int min_rows = getenv("PHC_TC_MIN_ROWS") ? atoi(getenv("PHC_TC_MIN_ROWS")) : 32;     // debug knob
  return (precision == 1 || precision == 2) && rows >= min_rows && in_features >= 32 && out_features >= 32 && phm_dim <= 16;
}

size_t phm_tc_fwd_workspace_bytes(int, int in_features, int out_features, int phm_dim, int) {
  return tc::pack_layout(phm_dim, in_features / phm_dim, out_features / phm_dim).total + 1024;
}

size_t phm_tc_bwd_workspace_bytes(int rows, int in_features, int out_features, int phm_dim, int) {
  const int s1 = tc::dh_splits(rows, in_features, out_features, nullptr), s2 = tc::dh2_splits(rows, in_features, out_features, nullptr);
  const size_t part = (size_t)(s1 > s2 ? s1 : s2) * in_features * out_features;
  return tc::pack_layout(phm_dim, in_features / phm_dim, out_features / phm_dim).total + 1024 +
         sizeof(float) * (part + phm_contract_scratch_floats(rows, in_features, out_features, phm_dim)) + 64;
}

static uint8_t* align1k(void* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); }

int phm_tc_fwd(const float* x, const float* A, const float* W, const float* bias, const float* residual, float* y, int rows,
               int in_features, int out_features, int phm_dim, int act, int precision, void* workspace, float* bn_partials,
               int* bn_produced, cudaStream_t stream) {
  const int n = phm_dim, K = in_features / n, P = out_features / n;
  const int single = precision == 2;
  uint8_t* buf = align1k(workspace);
  int rc = tc::launch_pack(A, W, n, K, P, buf, single, stream);
  if (rc) return rc;
  const tc::PackLayout L = tc::pack_layout(n, K, P);
  tc::MixParams p{};
  p.X = x; p.coef = reinterpret_cast<const float*>(buf + L.coef_fwd); p.Bpack = buf + L.pack_fwd;
  p.bias = bias; p.residual = residual; p.C = y;
  p.M = rows; p.n = n; p.Kin = K; p.Pout = P; p.S = n * K; p.chunks = L.chunks_fwd; p.ptiles = L.pt_fwd; p.act = act;
  p.num_tiles = phc_div_up(rows, tc::BM) * n * p.ptiles;
  p.single = single;
  p.prof = g_phm_tc_prof;
  if (L.sk_part && tc::stream_k_enabled()) {
    p.sk_part = reinterpret_cast<float*>(buf + L.sk_part);
    p.sk_flags = reinterpret_cast<unsigned int*>(buf + L.sk_flags);
  }
  p.stat_part = bn_partials;
  int v3 = 0;
  rc = tc::launch_mix(p, stream, &v3);
  if (bn_produced) *bn_produced = (v3 && bn_partials != nullptr) ? 1 : 0;
  return rc;
}

int phm_tc_bwd(const float* gy, const float* x, const float* A, const float* W, float* dx, float* dA, float* dW, float* db, int rows,
               int in_features, int out_features, int phm_dim, int precision, void* workspace, const void* fwd_pack, cudaStream_t stream) {
  const int n = phm_dim, K = in_features / n, P = out_features / n;
  const int single = precision == 2;
  uint8_t* buf = align1k(workspace);
  const tc::PackLayout L = tc::pack_layout(n, K, P);
  if (dx) {
    // the forward call already wrote both operand packs (same A, W); reuse them when the caller kept that workspace
    const uint8_t* pk = fwd_pack ? align1k(const_cast<void*>(fwd_pack)) : buf;
    if (!fwd_pack) {
      int rc0 = tc::launch_pack(A, W, n, K, P, buf, single, stream);
      if (rc0) return rc0;
    }
    int rc = PHC_OK;
    tc::MixParams p{};
    p.X = gy; p.coef = reinterpret_cast<const float*>(pk + L.coef_dx); p.Bpack = pk + L.pack_dx;
    p.C = dx;
    p.M = rows; p.n = n; p.Kin = P; p.Pout = K; p.S = n * P; p.chunks = L.chunks_dx; p.ptiles = L.pt_dx; p.act = PHC_ACT_IDENTITY;
    p.num_tiles = phc_div_up(rows, tc::BM) * n * p.ptiles;
    p.single = single;
    if (L.sk_part && tc::stream_k_enabled()) {          // scratch region of the pack workspace (the forward call's when its packs are reused)
      uint8_t* wpk = const_cast<uint8_t*>(pk);
      p.sk_part = reinterpret_cast<float*>(wpk + L.sk_part);
      p.sk_flags = reinterpret_cast<unsigned int*>(wpk + L.sk_flags);
    }
    rc = tc::launch_mix(p, stream);
    if (rc) return rc;
  }
  float* part = reinterpret_cast<float*>(buf + L.total);
  tc::DhParams d{};
  d.X = x; d.G = gy; d.C = part;
  d.M = rows; d.In = in_features; d.Out = out_features;
  d.tiles_m = phc_div_up(in_features, tc::BM); d.tiles_n = phc_div_up(out_features, tc::BN);
  d.splits = tc::dh_splits(rows, in_features, out_features, &d.rows_per_split);
  d.num_tiles = d.tiles_m * d.tiles_n * d.splits;
  d.single = single;
  // first choice: x^T operand in tensor memory, 128 x 256 blocks (more, shorter splits)
  tc::DhParams w = d;
  w.tiles_n = phc_div_up(out_features, tc::DH2_BN);
  w.splits = tc::dh2_splits(rows, in_features, out_features, &w.rows_per_split);
  w.num_tiles = w.tiles_m * w.tiles_n * w.splits;
  float* wscratch = part + (size_t)w.splits * in_features * out_features;
  w.db_part = db ? phm_contract_bias_partials(wscratch, in_features, out_features, n) : nullptr;
  int rc = tc::try_launch_dh_v2(w, stream);
  if (rc >= 0) {
    if (rc) return rc;
    return phm_contract_and_bias(part, w.splits, gy, A, W, dA, dW, db, rows, in_features, out_features, n, wscratch, db ? w.splits : 0,
                                 stream);
  }
  float* scratch = part + (size_t)d.splits * in_features * out_features;
  d.db_part = db ? phm_contract_bias_partials(scratch, in_features, out_features, n) : nullptr;
  int db_parts = db ? 2 * d.splits : 0;            // the TMA kernel also emits the column sums of gy (bias gradient)
  rc = tc::try_launch_dh_tma(d, stream);
  if (rc < 0) {
    db_parts = 0;
    static bool configured = false;
    rc = tc::set_smem(tc::phm_tc_dh_kernel, &configured);
    if (rc) return rc;
    const int grid = d.num_tiles < tc::num_sms() ? d.num_tiles : tc::num_sms();
    phc_launch(tc::phm_tc_dh_kernel, dim3(grid), dim3(tc::THREADS), tc::smem_bytes(0), stream, d);
    rc = phc_check_launch("phm_tc_dh_kernel");
  }
  if (rc) return rc;
  return phm_contract_and_bias(part, d.splits, gy, A, W, dA, dW, db, rows, in_features, out_features, n, scratch, db_parts, stream);
}

extern "C" void phc_debug_set_tc_profile(long long* device_buffer) { g_phm_tc_prof = device_buffer; }
