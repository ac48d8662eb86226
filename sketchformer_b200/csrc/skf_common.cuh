// Shared host/device helpers for libskf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/skf.h"

namespace skf {

// ---- thread-local error message --------------------------------------------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
extern thread_local int64_t g_launch_count;   // kernels launched by this thread

#define SKF_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return skf::fail(SKF_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,            \
                       cudaGetErrorString(_e));                                            \
  } while (0)

#define SKF_LAUNCH_CHECK()                                                                 \
  do {                                                                                     \
    skf::g_launch_count++;                                                                 \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess)                                                                 \
      return skf::fail(SKF_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,               \
                       cudaGetErrorString(_e));                                            \
  } while (0)

#define SKF_TRY(expr)                                                                      \
  do {                                                                                     \
    int _s = (expr);                                                                       \
    if (_s != 0) return _s;                                                                \
  } while (0)

#define SKF_REQUIRE(cond, code, ...)                                                       \
  do {                                                                                     \
    if (!(cond)) return skf::fail(code, __VA_ARGS__);                                      \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// ---- programmatic dependent launch ---------------------------------------------------------
// A kernel launched through launch_k(..., pdl = true) may become resident while the previous kernel of its stream is
// still running: its prologue (barrier init, tensor-memory allocation, tensor-map prefetch, block scheduling)
// overlaps that kernel's tail.  It must not touch global memory before pdl_wait() returns — the previous grid has
// then completed and its writes are visible (this also covers write-after-read on reused buffers).  A kernel lets its
// successor start early with pdl_trigger(), issued by every CTA AFTER it has taken the resources it needs (tensor
// memory above all: a successor that allocated first would hold them while it waits for this grid).  Both are no-ops
// when the launch carried no programmatic attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

// skf_set_option("pdl", mask): which launches carry the attribute — 1 the SIMT kernels (LayerNorm, losses, embeddings,
// folded cross attention ...), 2 the tcgen05 GEMMs, 4 the tcgen05 attention kernels; 8: the tcgen05 kernels trigger
// LATE (their TMA producer, once it has issued its last load) instead of right after their prologue, so that a
// successor does not sit on SMs the low-priority weight-gradient stream could have used during the tail.
extern int g_pdl;
constexpr int PDL_SIMT = 1, PDL_GEMM = 2, PDL_ATTN = 4, PDL_LATE = 8;
// <<<grid, block, smem, st>>> with the programmatic-stream-serialisation attribute when the option has the bit `pdl`.
// Only kernels that call pdl_wait() / pdl_enter() before their first global access may pass a non-zero class.
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl & g_pdl) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // the error is picked up by SKF_LAUNCH_CHECK()
}

// ---- warp helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Round-to-nearest fp32 -> tf32 (kept in an fp32 container).  tcgen05 kind::tf32 ignores the
// 13 low mantissa bits, i.e. truncates; operands written through this are unbiased.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// Counter-based dropout mask (tf.keras.layers.Dropout, builders/layers/transformer.py:217,221,250,255,259,296,334):
// keep(element) is a pure function of (seed, site, optimizer step, element index), so the backward pass and a
// CUDA-graph replay regenerate the mask instead of storing it, and the oracle reproduces it bit-exactly
// (oracle/dropout_rng.py).  Two murmur3 finalisers; the top 24 bits are compared with rate * 2^24.
struct DropSpec {
  float scale;              // 1 / (1 - rate); rate == 0 <=> thresh == 0 (no dropout)
  uint32_t thresh;          // (uint32)(rate * 2^24)
  uint32_t seed;
  uint32_t site;
  const int64_t* step;      // device pointer to optimizer.iterations
};
__device__ __forceinline__ uint32_t skf_fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// key = per-launch constant skf_drop_key(); returns x kept (scaled) or 0
__device__ __forceinline__ uint32_t skf_drop_key(const DropSpec& d) {
  return d.seed ^ ((uint32_t)(*d.step) * 0x632BE5ABu);
}
__device__ __forceinline__ float skf_drop_apply(float x, uint32_t key, uint32_t site, uint32_t idx, uint32_t thresh,
                                                float scale) {
  const uint32_t h = skf_fmix32(skf_fmix32(idx + 0x9E3779B9u * (site + 1u)) ^ key);
  return (h >> 8) >= thresh ? x * scale : 0.f;
}
#endif

// ---- internal launchers shared between the C ABI and the plan ------------------------------
int gemm_simt(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
              int64_t sb_n, float* Y, int64_t ldy, int M, int N, int K,
              const skf_gemm_epilogue* epi, cudaStream_t st);

}  // namespace skf
