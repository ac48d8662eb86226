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
#endif

// ---- internal launchers shared between the C ABI and the plan ------------------------------
int gemm_simt(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
              int64_t sb_n, float* Y, int64_t ldy, int M, int N, int K,
              const skf_gemm_epilogue* epi, cudaStream_t st);
int gemm_tc(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
            int64_t sb_n, float* Y, int64_t ldy, int M, int N, int K,
            const skf_gemm_epilogue* epi, cudaStream_t st);
// returns 1 when the tcgen05 path can take this problem (alignment / stride constraints)
int gemm_tc_supported(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
                      int64_t sb_n, const float* Y, int64_t ldy, int M, int N, int K,
                      const skf_gemm_epilogue* epi);

}  // namespace skf
