// Dense layers of the hot path (tf.keras.layers.Dense, builders/layers/transformer.py:154-158, 196-197;
// models/sketchformer.py:86-100) and their dgrad / wgrad as ONE persistent, warp-specialised tcgen05
// kernel.  One CTA per SM walks a static tile schedule (n fastest, so the CTAs running at the same
// time share A rows and the whole weight matrix in L2):
//
//   warp 0 (1 lane)   TMA producer : cp.async.bulk.tensor boxes -> 128B-swizzled smem ring
//   warps 6..9 (X3)   splitter     : lo = x - trunc_tf32(x) for both operand tiles, in shared memory
//   warp 1 (1 lane)   MMA issuer   : tcgen05.mma.cta_group::1.kind::tf32, 128 x BN x 8 per instruction,
//                                    fp32 accumulators DOUBLE-BUFFERED in TMEM (2 x BN columns)
//   warps 2..5        epilogue     : tcgen05.ld (thread == output row) -> bias / act / ReLU-mask /
//                                    residual -> 128B-swizzled smem tile -> TMA store (or TMA
//                                    reduce-add for split-K weight gradients); overlaps the next
//                                    tile's main loop
//
// Precision.  The reference is true fp32.  kind::tf32 reads only the top 19 bits of each fp32 operand
// (truncation), so X3 = true runs the error-compensated product  A*B ~= Ah*Bh + Ah*Bl + Al*Bh  with
// Xh = trunc_tf32(X) (what the tensor core sees of the raw tile) and Xl = X - Xh (exact in fp32) built
// by the splitter warps: three MMAs per k-step into the same accumulator.  X3 = false is the
// single-pass TF32 product (operands rounded to nearest by the TMA unit, TFLOAT32 tensor maps).
//
// Either operand may be K-major or MN-major (selected from the element strides), so forward (X*W),
// dgrad (dY*W^T) and wgrad (X^T*dY) all run here without a transpose.  M/N/K tails: TMA zero-fills
// loads and clips stores.
#include "skf_common.cuh"
#include "skf_ptx.cuh"

#ifndef SKF_SPLIT_UNROLL_N
#define SKF_SPLIT_UNROLL_N 8
#endif
constexpr int kSplitUnroll = SKF_SPLIT_UNROLL_N;   // float4s in flight per splitter thread

namespace skf {

using namespace ptx;

extern int g_tma_tf32_type;
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert);

constexpr int T2_BM = 128;
constexpr int T2_BK = 32;                   // fp32 elements = 128 bytes = one SW128 row
constexpr int T2_A_BYTES = T2_BM * 128;     // 16 KB
constexpr int T2_EPI_TILE = 32 * 128;       // 32 rows x 32 fp32: one TMA-store box per epilogue warp
constexpr int T2_SMEM_BUDGET = 227 * 1024 - 2048;

template <int BN, bool X3>
struct Tc2Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int RAW = T2_A_BYTES + B_BYTES;                 // bytes the TMA delivers per stage
  static constexpr int STAGE = X3 ? 2 * RAW : RAW;                 // + lo tiles
  static constexpr int EPI_BUFS = 2;
  static constexpr int EPI_BYTES = 4 * EPI_BUFS * T2_EPI_TILE;     // 32 KB
  static constexpr int STAGES_RAW = (T2_SMEM_BUDGET - EPI_BYTES - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;      // 128 / 256 / 512
  static constexpr int THREADS = X3 ? 320 : 192;
  static_assert(STAGES >= 2, "need at least a double-buffered smem ring");
};

struct Gemm2Params {
  int M, N, K;
  int k_chunk;          // K range of one split
  int mt, nt, splits;   // tile grid
  skf_gemm_epilogue e;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN, bool A_MN, bool B_MN, bool X3>
__global__ void __launch_bounds__(Tc2Cfg<BN, X3>::THREADS, 1)
gemm2_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
        const __grid_constant__ CUtensorMap tmY, const Gemm2Params p) {
  using C = Tc2Cfg<BN, X3>;
  constexpr int S = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t epi0 = base + S * C::STAGE;
  const uint32_t bar0 = epi0 + C::EPI_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto xfull_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * S + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (3 * S + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + S * C::STAGE + C::EPI_BYTES + 8 * (3 * S + 4));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int ntiles = p.mt * p.nt * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    for (int s = 0; s < S; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(xfull_bar(s), 128);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_enter();      // tensor memory is ours: the next kernel may become resident; nothing global was touched yet

  // tile t -> (split z, row block, column block), n fastest
  auto tile_coords = [&](int t, int& m0, int& n0, int& kb, int& nkb) {
    const int n = t % p.nt;
    const int r = t / p.nt;
    const int m = r % p.mt;
    const int z = r / p.mt;
    m0 = m * T2_BM;
    n0 = n * BN;
    kb = z * p.k_chunk;
    const int ke = min(p.K, kb + p.k_chunk);
    nkb = (ke - kb + T2_BK - 1) / T2_BK;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0, kb, nkb;
        tile_coords(t, m0, n0, kb, nkb);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), C::RAW);
          const uint32_t a_s = base + s * C::STAGE;
          const uint32_t b_s = a_s + T2_A_BYTES;
          const int k0 = kb + i * T2_BK;
          if (!A_MN) {
            tma_load_2d(a_s, &tmA, full_bar(s), k0, m0);                 // box {32 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < T2_BM / 32; j++)                          // box {32 m, 32 k}
              tma_load_2d(a_s + j * 4096, &tmA, full_bar(s), m0 + 32 * j, k0);
          }
          if (!B_MN) {
            tma_load_2d(b_s, &tmB, full_bar(s), k0, n0);                 // box {32 k, BN n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; j++)                             // box {32 n, 32 k}
              tma_load_2d(b_s + j * 4096, &tmB, full_bar(s), n0 + 32 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(T2_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0;
      int ti = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ti++) {
        int m0, n0, kb, nkb;
        tile_coords(t, m0, n0, kb, nkb);
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1u;
        mbar_wait(tempty_bar(a), aph ^ 1u);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(a * BN);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(X3 ? xfull_bar(s) : full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_s = base + s * C::STAGE;
          const uint32_t b_s = a_s + T2_A_BYTES;
          const uint32_t al_s = a_s + C::RAW;
          const uint32_t bl_s = al_s + T2_A_BYTES;
#pragma unroll
          for (int ks = 0; ks < T2_BK / 8; ks++) {                       // UMMA_K = 8 for tf32
            const uint32_t ao = A_MN ? ks * 1024 : ks * 32;
            const uint32_t bo = B_MN ? ks * 1024 : ks * 32;
            const uint64_t ad = A_MN ? make_desc_mnmajor(a_s + ao, 4096) : make_desc_kmajor(a_s + ao);
            const uint64_t bd = B_MN ? make_desc_mnmajor(b_s + bo, 4096) : make_desc_kmajor(b_s + bo);
            if (X3) {
              const uint64_t ald = A_MN ? make_desc_mnmajor(al_s + ao, 4096) : make_desc_kmajor(al_s + ao);
              const uint64_t bld = B_MN ? make_desc_mnmajor(bl_s + bo, 4096) : make_desc_kmajor(bl_s + bo);
              mma_tf32_ss(acc, ad, bld, idesc, (i | ks) != 0 ? 1u : 0u);   // small terms first
              mma_tf32_ss(acc, ald, bd, idesc, 1u);
              mma_tf32_ss(acc, ad, bd, idesc, 1u);
            } else {
              mma_tf32_ss(acc, ad, bd, idesc, (i | ks) != 0 ? 1u : 0u);
            }
          }
          tc_commit(empty_bar(s));       // slot reusable once these MMAs have read it
        }
        tc_commit(tfull_bar(a));         // accumulator complete
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int ew = warp - 2;
    const uint32_t stg0 = epi0 + (uint32_t)(ew * C::EPI_BUFS) * T2_EPI_TILE;
    uint8_t* stg0_gen = gen + S * C::STAGE + (ew * C::EPI_BUFS) * T2_EPI_TILE;
    const skf_gemm_epilogue& e = p.e;
    int ti = 0;
    uint32_t nstore = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ti++) {
      int m0, n0, kb, nkb;
      tile_coords(t, m0, n0, kb, nkb);
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1u;
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      const int mrow0 = m0 + q * 32;
      const int m = mrow0 + lane;
      const bool row_ok = m < p.M;
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
      constexpr int NCH = BN / 32;
#pragma unroll 1
      for (int c = 0; c < NCH; c++) {
        const int nb = n0 + c * 32;
        const bool live = nb < p.N && mrow0 < p.M;      // warp-uniform
        uint32_t r[32];
        if (live) {
          tmem_ld32(acc + (uint32_t)(c * 32), r);
          tmem_ld_wait();
        }
        if (c == NCH - 1) {
          // every column of this accumulator has been read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(a));
        }
        if (!live) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
        if (!e.accumulate) {
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (nb + 4 * j < p.N) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + nb) + j);
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
          }
          if (e.act == SKF_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
          } else if (e.act == SKF_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = tanhf(v[j]);
          }
          if (e.mask && row_ok) {
            const float4* mrow = reinterpret_cast<const float4*>(e.mask + (int64_t)m * e.ld_mask + nb);
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (nb + 4 * j < p.N) {
                const float4 k4 = mrow[j];
                v[4 * j] = k4.x > 0.f ? v[4 * j] : 0.f;
                v[4 * j + 1] = k4.y > 0.f ? v[4 * j + 1] : 0.f;
                v[4 * j + 2] = k4.z > 0.f ? v[4 * j + 2] : 0.f;
                v[4 * j + 3] = k4.w > 0.f ? v[4 * j + 3] : 0.f;
              }
          }
          if (e.resid && row_ok) {
            const float4* rrow = reinterpret_cast<const float4*>(e.resid + (int64_t)m * e.ld_resid + nb);
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (nb + 4 * j < p.N) {
                const float4 r4 = rrow[j];
                v[4 * j] += r4.x; v[4 * j + 1] += r4.y; v[4 * j + 2] += r4.z; v[4 * j + 3] += r4.w;
              }
          }
          if (e.round_tf32) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = round_tf32(v[j]);
          }
        }
        // registers (thread == row) -> 128B-swizzled staging tile -> TMA store
        const uint32_t buf = nstore % C::EPI_BUFS;
        if (lane == 0) bulk_wait_read<C::EPI_BUFS - 1>();     // the store that last used `buf` has read it
        __syncwarp();
        uint8_t* srow = stg0_gen + buf * T2_EPI_TILE + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<float4*>(srow + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const uint32_t src = stg0 + buf * T2_EPI_TILE;
          if (e.accumulate) tma_reduce_add_2d(&tmY, src, nb, mrow0);
          else tma_store_2d(&tmY, src, nb, mrow0);
          bulk_commit();
        }
        nstore++;
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  } else {
    // ------------------------------------------------------------------ splitter (X3 only, 4 warps)
    if (X3) {
      const int tid = threadIdx.x - 192;           // 0..127
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0, kb, nkb;
        tile_coords(t, m0, n0, kb, nkb);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(full_bar(s), ph);
          const float4* src = reinterpret_cast<const float4*>(gen + s * C::STAGE);
          float4* dst = reinterpret_cast<float4*>(gen + s * C::STAGE + C::RAW);
          // elementwise, so the (identical) swizzled layouts of the raw and lo tiles do not matter
#pragma unroll kSplitUnroll
          for (int k = tid; k < C::RAW / 16; k += 128) {
            const float4 x = src[k];
            float4 l;
            l.x = x.x - trunc_tf32(x.x); l.y = x.y - trunc_tf32(x.y);
            l.z = x.z - trunc_tf32(x.z); l.w = x.w - trunc_tf32(x.w);
            dst[k] = l;
          }
          fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
          mbar_arrive(xfull_bar(s));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------------
static int make_tmap_store(CUtensorMap* out, float* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
  // plain FLOAT32 (no tf32 rounding of stored values), box {32 cols, 32 rows}, 128B swizzle
  return make_tmap_2d_typed(out, ptr, inner, outer, ld, 32, 32, false, false);
}

int gemm_tc2_supported(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
                       const float* Y, int64_t ldy, int M, int N, int K, const skf_gemm_epilogue* e) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(Y) & 15))
    return 0;
  if ((ldy & 3) || (N & 3) || ldy < N) return 0;
  const bool a_k = (sa_k == 1 && sa_m % 4 == 0 && sa_m >= K) || (sa_k == 1 && M == 1);
  const bool a_mn = (sa_m == 1 && sa_k % 4 == 0 && sa_k >= M);
  const bool b_k = (sb_k == 1 && sb_n % 4 == 0 && sb_n >= K);
  const bool b_mn = (sb_n == 1 && sb_k % 4 == 0 && sb_k >= N);
  if (!(a_k || a_mn) || !(b_k || b_mn)) return 0;
  if (e) {
    if (e->bias && (reinterpret_cast<uintptr_t>(e->bias) & 15)) return 0;
    if (e->mask && ((reinterpret_cast<uintptr_t>(e->mask) & 15) || (e->ld_mask & 3))) return 0;
    if (e->resid && ((reinterpret_cast<uintptr_t>(e->resid) & 15) || (e->ld_resid & 3))) return 0;
  }
  return 1;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN, bool X3>
static int launch_tc2(const CUtensorMap* tm, const Gemm2Params& p, cudaStream_t st) {
  using C = Tc2Cfg<BN, X3>;
  static bool attr_done = false;
  if (!attr_done) {
    SKF_CUDA(cudaFuncSetAttribute(gemm2_k<BN, A_MN, B_MN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done = true;
  }
  const int ntiles = p.mt * p.nt * p.splits;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  launch_k(gemm2_k<BN, A_MN, B_MN, X3>, dim3(grid), dim3(C::THREADS), C::SMEM, st, PDL_GEMM, tm[0], tm[1], tm[2], p);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

template <int BN, bool X3>
static int dispatch_major2(bool a_mn, bool b_mn, const CUtensorMap* tm, const Gemm2Params& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_tc2<BN, false, false, X3>(tm, p, st);
  if (!a_mn && b_mn) return launch_tc2<BN, false, true, X3>(tm, p, st);
  if (a_mn && !b_mn) return launch_tc2<BN, true, false, X3>(tm, p, st);
  return launch_tc2<BN, true, true, X3>(tm, p, st);
}

// x3 != 0: error-compensated 3xTF32 product (fp32-grade), operands split in shared memory.
int gemm_tc2(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* Y,
             int64_t ldy, int M, int N, int K, const skf_gemm_epilogue* epi, int x3, cudaStream_t st) {
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  if (epi) e = *epi;
  SKF_REQUIRE(gemm_tc2_supported(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e), SKF_ERR_BAD_ALIGN,
              "gemm_tc2: operand strides/alignment not TMA-compatible (sa=%lld,%lld sb=%lld,%lld ldy=%lld M=%d N=%d K=%d)",
              (long long)sa_m, (long long)sa_k, (long long)sb_k, (long long)sb_n, (long long)ldy, M, N, K);
  int split = e.split_k > 1 ? e.split_k : 1;
  SKF_REQUIRE(split == 1 || e.accumulate, SKF_ERR_BAD_ARG, "gemm: split_k needs accumulate");
  SKF_REQUIRE(!(e.accumulate && (e.bias || e.resid || e.mask || e.act)), SKF_ERR_BAD_ARG,
              "gemm: accumulate excludes other epilogue terms");
  const bool a_mn = !(sa_k == 1);
  const bool b_mn = !(sb_k == 1);
  // 3xTF32 with a single 256-column tile per row block (N <= 256, e.g. the d_model-wide projections): 128-column
  // tiles double the tile count (better wave quantisation on 148 SMs) and buy a third smem stage; measured
  // 8-18 % faster on B200 (FFN-2 25600x1024x256: 91.8 -> 74.9 us).  Split-K weight gradients keep 256.
  const int BN = N <= 64 ? 64 : ((N <= 128 || (x3 && N <= 256 && !e.accumulate)) ? 128 : 256);
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.k_chunk = cdiv(cdiv(K, split), T2_BK) * T2_BK;
  p.splits = cdiv(K, p.k_chunk);
  p.mt = cdiv(M, T2_BM);
  p.nt = cdiv(N, BN);
  p.e = e;
  p.e.a_lo = p.e.b_lo = nullptr;
  const bool conv = !x3 && g_tma_tf32_type != 0;     // X3 needs the raw fp32 bits in shared memory
  CUtensorMap tm[3];
  if (!a_mn) SKF_TRY(make_tmap_2d_typed(&tm[0], A, (uint64_t)K, (uint64_t)M, (uint64_t)(M == 1 ? K : sa_m), T2_BK, T2_BM, false, conv));
  else SKF_TRY(make_tmap_2d_typed(&tm[0], A, (uint64_t)M, (uint64_t)K, (uint64_t)sa_k, 32, T2_BK, true, conv));
  if (!b_mn) SKF_TRY(make_tmap_2d_typed(&tm[1], B, (uint64_t)K, (uint64_t)N, (uint64_t)sb_n, T2_BK, (uint32_t)BN, false, conv));
  else SKF_TRY(make_tmap_2d_typed(&tm[1], B, (uint64_t)N, (uint64_t)K, (uint64_t)sb_k, 32, T2_BK, true, conv));
  SKF_TRY(make_tmap_store(&tm[2], Y, (uint64_t)N, (uint64_t)M, (uint64_t)ldy));
  if (x3) {
    switch (BN) {
      case 64: return dispatch_major2<64, true>(a_mn, b_mn, tm, p, st);
      case 128: return dispatch_major2<128, true>(a_mn, b_mn, tm, p, st);
      default: return dispatch_major2<256, true>(a_mn, b_mn, tm, p, st);
    }
  }
  switch (BN) {
    case 64: return dispatch_major2<64, false>(a_mn, b_mn, tm, p, st);
    case 128: return dispatch_major2<128, false>(a_mn, b_mn, tm, p, st);
    default: return dispatch_major2<256, false>(a_mn, b_mn, tm, p, st);
  }
}

}  // namespace skf
