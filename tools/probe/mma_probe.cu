// Micro-probe for the attention redesign (not part of the product library): validates the tcgen05 operand layouts the
// new kernels rely on and times chains of small MMAs on a B200.
//   1. SS  kind::f16 bf16 x bf16, K-major A [128 x 32] and B [N x 32] tiles (64-byte rows, SWIZZLE_64B), N = 208
//   2. TS  kind::f16 with A in TMEM (packed 16-bit pairs) and B MN-major [64 x 32] (+ concat N = 64): bf16 x bf16,
//      f16 x bf16 (mixed operand formats) and f16 x f16
//   3. timing: dependent chains of N = 32 / 64 TS MMAs, and the same MMAs round-robin over 2 / 4 accumulators
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../sketchformer_b200/csrc/skf_ptx.cuh"
using namespace skf::ptx;

__device__ __forceinline__ uint64_t desc16_k(uint32_t addr) { return make_smem_desc(addr, 16, 512, 4); }
__device__ __forceinline__ uint64_t desc16_mn(uint32_t addr, uint32_t chunk) { return make_smem_desc(addr, chunk, 512, 4); }
// a_fmt / b_fmt: 0 = f16, 1 = bf16
__host__ __device__ constexpr uint32_t idesc16(int M, int N, int a_fmt, int b_fmt, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
// bf16 tile [rows x 32] with 64-byte rows under the 64-byte swizzle: byte offset of element (r, c)
__host__ __device__ inline uint32_t sw64(int r, int c) { return r * 64 + ((((c >> 3) ^ ((r >> 1) & 3)) << 4) | ((c & 7) << 1)); }

struct Args {
  const float* Q;   // [128, 32]
  const float* Kt;  // [208, 32]
  const float* P;   // [128, 64]
  const float* V;   // [64, 64]  (columns 0..31 "hi", 32..63 "lo")
  float* S;         // [128, 208]
  float* O;         // [3][128, 64]  variant: 0 bf16 x bf16, 1 f16(A) x bf16(B), 2 f16 x f16
  long long* clk;   // timings
  int variants;     // bit v: run TS variant v
};

__global__ void __launch_bounds__(128, 1) probe_k(Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // tiles: Q 8 KB | K 16 KB (256 rows) | Vb (bf16 hi|lo chunks: 2 x 4 KB) | Vh (f16: 2 x 4 KB) | barriers
  const uint32_t sQ = base, sK = sQ + 8192, sVb = sK + 16384, sVh = sVb + 8192, bar = sVh + 8192, slot = bar + 64;
  uint8_t *gQ = gen, *gK = gQ + 8192, *gVb = gK + 16384, *gVh = gVb + 8192;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 32; i += 128) {
    const int r = i >> 5, c = i & 31;
    *reinterpret_cast<__nv_bfloat16*>(gQ + sw64(r, c)) = __float2bfloat16_rn(a.Q[i]);
  }
  for (int i = tid; i < 256 * 32; i += 128) {
    const int r = i >> 5, c = i & 31;
    *reinterpret_cast<__nv_bfloat16*>(gK + sw64(r, c)) = __float2bfloat16_rn(r < 208 ? a.Kt[i] : 0.f);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i >> 6, c = i & 63, ch = c >> 5;       // chunk ch = [64 rows x 32] at + ch * 4096
    *reinterpret_cast<__nv_bfloat16*>(gVb + ch * 4096 + sw64(r, c & 31)) = __float2bfloat16_rn(a.V[i]);
    *reinterpret_cast<__half*>(gVh + ch * 4096 + sw64(r, c & 31)) = __float2half_rn(a.V[i]);
  }
  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(bar + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t ph = 0;

  // ---- 1. S = Q K^T, N = 208, two K = 16 steps ---------------------------------------------
  if (tid == 0) {
    const uint32_t id = idesc16(128, 208, 1, 1, 0, 0);
    mma16_ss(tm, desc16_k(sQ), desc16_k(sK), id, 0);
    mma16_ss(tm, desc16_k(sQ + 32), desc16_k(sK + 32), id, 1);
    tc_commit(bar);
  }
  mbar_wait(bar, ph); ph ^= 1;
  tc_fence_after();
  for (int c0 = 0; c0 < 208; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tm + lane_off + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++)
      if (c0 + j < 208) a.S[tid * 208 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();

  // ---- 2. O = P V with P in TMEM (columns 256..), three operand-format variants ---------------------
  for (int var = 0; var < 3; var++) {
    if (!((a.variants >> var) & 1)) continue;
    const int afmt = var == 0 ? 1 : 0, bfmt = var == 2 ? 0 : 1;
    uint32_t pk[32];
    for (int j = 0; j < 32; j++) {
      const float p0 = a.P[tid * 64 + 2 * j], p1 = a.P[tid * 64 + 2 * j + 1];
      if (afmt == 1) {
        __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
        pk[j] = *reinterpret_cast<uint32_t*>(&t);
      } else {
        __half2 t = __floats2half2_rn(p0, p1);
        pk[j] = *reinterpret_cast<uint32_t*>(&t);
      }
    }
    tmem_st32(tm + lane_off + 256, pk);      // 64 K elements = 32 columns
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id = idesc16(128, 64, afmt, bfmt, 0, 1);      // B MN-major, N = 64 = two 32-wide chunks 4096 B apart
      const uint32_t sV = bfmt == 1 ? sVb : sVh;
      for (int k = 0; k < 4; k++) mma16_ts(tm + 320, tm + 256 + 8 * k, desc16_mn(sV + k * 1024, 4096), id, k > 0);
      tc_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1;
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tm + lane_off + 320 + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; j++) a.O[(var * 128 + tid) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---- 3. timing ---------------------------------------------------------------------------------------
  // t[0]: 64 dependent N=32 TS MMAs; t[1]: 64 dependent N=64; t[2]: 64 N=32 over 2 accumulators; t[3]: over 4;
  // t[4]: 64 N=64 over 2 accumulators; t[5]: 12 SS N=208; t[6]: 64 dependent SS N=32 (K-major B); t[7]: empty commit
  for (int test = 0; test < 8; test++) {
    __syncthreads();
    long long t0 = 0;
    if (tid == 0) {
      tc_fence_after();
      t0 = clock64();
      const uint32_t id32 = idesc16(128, 32, 1, 1, 0, 1), id64 = idesc16(128, 64, 1, 1, 0, 1);
      const uint32_t idS = idesc16(128, 208, 1, 1, 0, 0), idk32 = idesc16(128, 32, 1, 1, 0, 0);
      if (test == 0) for (int i = 0; i < 64; i++) mma16_ts(tm + 320, tm + 256 + 8 * (i & 3), desc16_mn(sVb + (i & 3) * 1024, 4096), id32, 1);
      if (test == 1) for (int i = 0; i < 64; i++) mma16_ts(tm + 320, tm + 256 + 8 * (i & 3), desc16_mn(sVb + (i & 3) * 1024, 4096), id64, 1);
      if (test == 2) for (int i = 0; i < 64; i++) mma16_ts(tm + 320 + 32 * (i & 1), tm + 256 + 8 * (i & 3), desc16_mn(sVb + (i & 3) * 1024, 4096), id32, 1);
      if (test == 3) for (int i = 0; i < 64; i++) mma16_ts(tm + 320 + 32 * (i & 3), tm + 256 + 8 * (i & 3), desc16_mn(sVb + (i & 3) * 1024, 4096), id32, 1);
      if (test == 4) for (int i = 0; i < 64; i++) mma16_ts(tm + 320 + 64 * (i & 1), tm + 256 + 8 * (i & 3), desc16_mn(sVb + (i & 3) * 1024, 4096), id64, 1);
      if (test == 5) for (int i = 0; i < 12; i++) mma16_ss(tm, desc16_k(sQ + 32 * (i & 1)), desc16_k(sK + 32 * (i & 1)), idS, 1);
      if (test == 6) for (int i = 0; i < 64; i++) mma16_ss(tm + 320, desc16_k(sQ + 32 * (i & 1)), desc16_k(sK + 32 * (i & 1)), idk32, 1);
      tc_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1;
    if (tid == 0) a.clk[test] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float hf(float x) { return __half2float(__float2half_rn(x)); }

int main(int argc, char** argv) {
  std::vector<float> Q(128 * 32), K(208 * 32), P(128 * 64), V(64 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : Q) x = rnd();
  for (auto& x : K) x = rnd();
  for (auto& x : P) x = fabsf(rnd());
  for (auto& x : V) x = rnd() * 1e-3f;          // small magnitudes, like gradients (exercises the f16 range)
  Args a;
  float *dQ, *dK, *dP, *dV, *dS, *dO;
  long long* dclk;
  cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dK, K.size() * 4); cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dV, V.size() * 4);
  cudaMalloc(&dS, 128 * 208 * 4); cudaMalloc(&dO, 3 * 128 * 64 * 4); cudaMalloc(&dclk, 8 * 8);
  cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dK, K.data(), K.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice);
  a.variants = argc > 1 ? atoi(argv[1]) : 7;
  cudaMemset(dO, 0, 3 * 128 * 64 * 4);
  a.Q = dQ; a.Kt = dK; a.P = dP; a.V = dV; a.S = dS; a.O = dO; a.clk = dclk;
  const int smem = 8192 + 16384 + 8192 + 8192 + 1024 + 256;
  cudaFuncSetAttribute(probe_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_k<<<1, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> S(128 * 208), O(3 * 128 * 64);
  long long clk[8];
  cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
  double num = 0, den = 0;
  for (int i = 0; i < 128; i++)
    for (int j = 0; j < 208; j++) {
      double r = 0;
      for (int k = 0; k < 32; k++) r += (double)bf(Q[i * 32 + k]) * bf(K[j * 32 + k]);
      num += (S[i * 208 + j] - r) * (S[i * 208 + j] - r); den += r * r;
    }
  printf("SS  bf16 K-major N=208      : rel err %.3e\n", sqrt(num / den));
  const char* names[3] = {"TS  bf16 x bf16 (MN-major B)", "TS  f16(A) x bf16(B) mixed  ", "TS  f16 x f16               "};
  for (int var = 0; var < 3; var++) {
    num = den = 0;
    for (int i = 0; i < 128; i++)
      for (int j = 0; j < 64; j++) {
        double r = 0;
        for (int k = 0; k < 64; k++) {
          float p = var == 0 ? bf(P[i * 64 + k]) : hf(P[i * 64 + k]);
          float v = var == 2 ? hf(V[k * 64 + j]) : bf(V[k * 64 + j]);
          r += (double)p * v;
        }
        double g = O[(var * 128 + i) * 64 + j];
        num += (g - r) * (g - r); den += r * r;
      }
    printf("%s: rel err %.3e\n", names[var], sqrt(num / den));
  }
  const char* tn[8] = {"64 dependent TS N=32", "64 dependent TS N=64", "64 TS N=32 over 2 accumulators", "64 TS N=32 over 4 accumulators",
                       "64 TS N=64 over 2 accumulators", "12 SS N=208 (dependent)", "64 dependent SS N=32", "commit only"};
  for (int t = 0; t < 8; t++) printf("%-34s: %lld clk\n", tn[t], clk[t]);
  return 0;
}
