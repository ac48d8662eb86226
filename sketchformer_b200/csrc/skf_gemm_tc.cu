// K2/K4/K5/K9/K13: the Dense layers of the hot path (tf.keras.layers.Dense at
// builders/layers/transformer.py:154-158, 196-197; models/sketchformer.py:86-100) and their
// dgrad / wgrad, as ONE warp-specialised tcgen05 kernel:
//
//   warp 0 (1 lane)  TMA producer : cp.async.bulk.tensor 2-D boxes -> 128B-swizzled smem ring
//   warp 1 (1 lane)  MMA issuer   : tcgen05.mma.cta_group::1.kind::tf32, 128 x BN x 8 per instr,
//                                   fp32 accumulator in TMEM; tcgen05.commit frees ring slots
//   warps 2..5       epilogue     : tcgen05.ld (thread == output row) -> bias / act / ReLU-mask /
//                                   residual / tf32 rounding -> 128-bit global stores, or
//                                   red.add for split-K weight gradients
//
// Operands are fp32 in HBM and go through the tensor cores as TF32 (the reference is true fp32;
// bf16 operands miss the 1e-3 parity bar).  Either operand may be K-major or
// MN-major, selected from the element strides, so forward (X*W), dgrad (dY*W^T) and wgrad
// (X^T*dY) all run here without any transpose kernel.  M/N/K tails rely on TMA zero fill.
#include "skf_common.cuh"
#include "skf_ptx.cuh"

namespace skf {

using namespace ptx;

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;              // fp32 elements = 128 bytes = one SW128 atom row
constexpr int TC_A_BYTES = TC_BM * 128;
constexpr int TC_SMEM_LIMIT = 200 * 1024;

template <int BN>
struct TcCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (TC_SMEM_LIMIT / STAGE) > 8 ? 8 : (TC_SMEM_LIMIT / STAGE);
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

// 1 (default): tensor maps are typed CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, i.e. the TMA unit rounds fp32 -> tf32
// (to nearest) on the way into shared memory; measured on B200 (tools/tf32_probe.py): unbiased, rel. error
// 1.5e-5 vs 7.4e-4 (biased) when kind::tf32 truncates raw fp32 bits.  0: plain FLOAT32 maps.
int g_tma_tf32_type = 1;

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
          float* __restrict__ Y, int64_t ldy, int M, int N, int K, int k_chunk, int nseg,
          skf_gemm_epilogue e) {
  using C = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar0 = base + C::STAGES * C::STAGE;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 16u * C::STAGES;
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + C::STAGES * C::STAGE + 16 * C::STAGES + 8);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM;
  const int n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk;
  const int ke = min(K, kb + k_chunk);
  const int nkb = (ke - kb + TC_BK - 1) / TC_BK;
  // nseg == 3: error-compensated "3xTF32" product  A*B ~= Ah*Bh + Ah*Bl + Al*Bh  (fp32-accurate);
  // the three passes accumulate into the same TMEM tile, so they are just a 3x longer K loop.
  const int niter = nseg * nkb;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (nseg > 1) {
      prefetch_tmap(&tmAlo);
      prefetch_tmap(&tmBlo);
    }
    for (int s = 0; s < C::STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < niter; i++) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        const int seg = i / nkb;
        const CUtensorMap* ma = seg == 2 ? &tmAlo : &tmA;
        const CUtensorMap* mb = seg == 1 ? &tmBlo : &tmB;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), C::STAGE);
        const uint32_t a_s = base + s * C::STAGE;
        const uint32_t b_s = a_s + TC_A_BYTES;
        const int k0 = kb + (i - seg * nkb) * TC_BK;
        if (!A_MN) {
          tma_load_2d(a_s, ma, full_bar(s), k0, m0);                // box {32 k, 128 m}
        } else {
#pragma unroll
          for (int j = 0; j < TC_BM / 32; j++)                       // box {32 m, 32 k}
            tma_load_2d(a_s + j * 4096, ma, full_bar(s), m0 + 32 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(b_s, mb, full_bar(s), k0, n0);                // box {32 k, BN n}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; j++)                          // box {32 n, 32 k}
            tma_load_2d(b_s + j * 4096, mb, full_bar(s), n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < niter; i++) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_s = base + s * C::STAGE;
        const uint32_t b_s = a_s + TC_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ks++) {                     // UMMA_K = 8 for tf32
          const uint64_t ad = A_MN ? make_desc_mnmajor(a_s + ks * 1024, 4096) : make_desc_kmajor(a_s + ks * 32);
          const uint64_t bd = B_MN ? make_desc_mnmajor(b_s + ks * 1024, 4096) : make_desc_kmajor(b_s + ks * 32);
          mma_tf32_ss(tmem_base, ad, bd, idesc, (i | ks) != 0 ? 1u : 0u);
        }
        tc_commit(empty_bar(s));     // slot reusable once these MMAs have read it
      }
      tc_commit(tmem_full_bar);      // accumulator complete
    }
  } else {
    // ---- epilogue: 4 warps, warp (w % 4) owns TMEM lanes [32*(w%4), +32) -------------------
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = m < M;
    const bool vec_ok = ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; c++) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (!row_ok || nb >= N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
      float* yrow = Y + (int64_t)m * ldy + nb;
      if (e.accumulate) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (nb + j < N) atomicAdd(yrow + j, v[j]);
        continue;
      }
      const int nval = min(32, N - nb);
      if (e.bias) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < nval) v[j] += __ldg(e.bias + nb + j);
      }
      if (e.act == SKF_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
      } else if (e.act == SKF_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = tanhf(v[j]);
      }
      if (e.mask) {
        const float* mrow = e.mask + (int64_t)m * e.ld_mask + nb;
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < nval) v[j] = mrow[j] > 0.f ? v[j] : 0.f;
      }
      if (e.resid) {
        const float* rrow = e.resid + (int64_t)m * e.ld_resid + nb;
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < nval) v[j] += rrow[j];
      }
      if (e.round_tf32) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = round_tf32(v[j]);
      }
      if (vec_ok && nval == 32) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          reinterpret_cast<float4*>(yrow)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < nval) yrow[j] = v[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart.
int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert);
int make_tmap_2d(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                 uint32_t box_inner, uint32_t box_outer, bool mn_major) {
  return make_tmap_2d_typed(out, ptr, inner, outer, ld, box_inner, box_outer, mn_major, g_tma_tf32_type != 0);
}
int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert) {
  EncodeTiledFn enc = get_encode();
  SKF_REQUIRE(enc != nullptr, SKF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = tf32_convert ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SKF_REQUIRE(r == CUDA_SUCCESS, SKF_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r,
              (const void*)ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld,
              box_inner, box_outer);
  return SKF_OK;
}

// 2-D bf16 tensor map with 64-byte rows in shared memory (box_inner = 32 elements) under the 64-byte swizzle: the
// layout of the pair GEMM's bf16 operand tiles (skf_gemm_tc3.cu).
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  SKF_REQUIRE(enc != nullptr, SKF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SKF_REQUIRE(r == CUDA_SUCCESS, SKF_ERR_CUDA,
              "cuTensorMapEncodeTiled (bf16) failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, ptr,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  return SKF_OK;
}

int gemm_tc_supported(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
                      int64_t sb_n, const float* Y, int64_t ldy, int M, int N, int K,
                      const skf_gemm_epilogue* epi) {
  (void)Y; (void)ldy; (void)epi;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 0;
  bool a_k = (sa_k == 1 && sa_m % 4 == 0 && sa_m >= K) || (sa_k == 1 && M == 1);
  bool a_mn = (sa_m == 1 && sa_k % 4 == 0 && sa_k >= M);
  bool b_k = (sb_k == 1 && sb_n % 4 == 0 && sb_n >= K);
  bool b_mn = (sb_n == 1 && sb_k % 4 == 0 && sb_k >= N);
  if (!(a_k || a_mn) || !(b_k || b_mn)) return 0;
  return 1;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap* tm, int nseg, float* Y, int64_t ldy, int M, int N, int K,
                     int k_chunk, int split, const skf_gemm_epilogue& e, cudaStream_t st) {
  using C = TcCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    SKF_CUDA(cudaFuncSetAttribute(gemm_tc_k<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done = true;
  }
  dim3 grid(cdiv(N, BN), cdiv(M, TC_BM), split);
  gemm_tc_k<BN, A_MN, B_MN><<<grid, 192, C::SMEM, st>>>(tm[0], tm[1], tm[2], tm[3], Y, ldy, M, N, K, k_chunk, nseg, e);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap* tm, int nseg, float* Y,
                          int64_t ldy, int M, int N, int K, int k_chunk, int split, const skf_gemm_epilogue& e,
                          cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_tc<BN, false, false>(tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true>(tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
  if (a_mn && !b_mn) return launch_tc<BN, true, false>(tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
  return launch_tc<BN, true, true>(tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
}

int gemm_tc(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
            float* Y, int64_t ldy, int M, int N, int K, const skf_gemm_epilogue* epi, cudaStream_t st) {
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  if (epi) e = *epi;
  SKF_REQUIRE(gemm_tc_supported(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e), SKF_ERR_BAD_ALIGN,
              "gemm_tc: operand strides/alignment not TMA-compatible (sa=%lld,%lld sb=%lld,%lld M=%d N=%d K=%d)",
              (long long)sa_m, (long long)sa_k, (long long)sb_k, (long long)sb_n, M, N, K);
  int split = e.split_k > 1 ? e.split_k : 1;
  SKF_REQUIRE(split == 1 || e.accumulate, SKF_ERR_BAD_ARG, "gemm: split_k needs accumulate");
  SKF_REQUIRE(!(e.accumulate && (e.bias || e.resid || e.mask || e.act)), SKF_ERR_BAD_ARG,
              "gemm: accumulate excludes other epilogue terms");
  int k_chunk = cdiv(cdiv(K, split), TC_BK) * TC_BK;
  split = cdiv(K, k_chunk);
  const bool a_mn = !(sa_k == 1);
  const bool b_mn = !(sb_k == 1);
  int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  if (BN == 256) {
    // prefer 128-wide tiles when 256-wide ones leave the 148 SMs under two waves
    long tiles256 = (long)cdiv(M, TC_BM) * cdiv(N, 256) * split;
    if (tiles256 < 2 * 148) BN = 128;
  }
  const int nseg = (e.a_lo && e.b_lo) ? 3 : 1;
  SKF_REQUIRE(nseg == 1 || (((reinterpret_cast<uintptr_t>(e.a_lo) | reinterpret_cast<uintptr_t>(e.b_lo)) & 15) == 0),
              SKF_ERR_BAD_ALIGN, "gemm_tc: a_lo/b_lo must be 16-byte aligned");
  CUtensorMap tm[4];  // A, B, A_lo, B_lo  (the lo tensors share the strides of A / B)
  for (int i = 0; i < (nseg == 3 ? 2 : 1); i++) {
    const float* a = i == 0 ? A : e.a_lo;
    const float* b = i == 0 ? B : e.b_lo;
    if (!a_mn) SKF_TRY(make_tmap_2d(&tm[2 * i], a, (uint64_t)K, (uint64_t)M, (uint64_t)(M == 1 ? K : sa_m), TC_BK, TC_BM, false));
    else SKF_TRY(make_tmap_2d(&tm[2 * i], a, (uint64_t)M, (uint64_t)K, (uint64_t)sa_k, 32, TC_BK, true));
    if (!b_mn) SKF_TRY(make_tmap_2d(&tm[2 * i + 1], b, (uint64_t)K, (uint64_t)N, (uint64_t)sb_n, TC_BK, (uint32_t)BN, false));
    else SKF_TRY(make_tmap_2d(&tm[2 * i + 1], b, (uint64_t)N, (uint64_t)K, (uint64_t)sb_k, 32, TC_BK, true));
  }
  if (nseg == 1) { tm[2] = tm[0]; tm[3] = tm[1]; }
  switch (BN) {
    case 64: return dispatch_major<64>(a_mn, b_mn, tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
    case 128: return dispatch_major<128>(a_mn, b_mn, tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
    default: return dispatch_major<256>(a_mn, b_mn, tm, nseg, Y, ldy, M, N, K, k_chunk, split, e, st);
  }
}

}  // namespace skf
