// 2-CTA (cta_group::2) variant of the persistent 3xTF32 tcgen05 GEMM in skf_gemm_tc2.cu.
//
// A CTA pair shares the B operand: each CTA stages only HALF of the 256-column B tile (its own 128 columns)
// and the pair's tensor cores exchange the halves, so per CTA and 32-deep k-block the shared-memory traffic
// drops from 288 KB (48 TMA + 96 hi/lo splitter + 12 MMAs x 12 KB operand reads) to 192 KB and the L2 -> SM
// traffic per FLOP by a third; the smaller stage (64 KB instead of 96 KB) buys a third pipeline stage.
// Measured on B200 (FFN-1, 25600 x 256 x 1024, 3xTF32): 79.4 -> 61.4 us.  With the TMA loads and the splitter
// disabled the same kernel takes 56 us: what remains is the tensor pipe itself at the power-limited clock
// (~1.3 GHz under tensor load, the clock MEASURED_PEAKS.json's cuBLAS figure was taken at), i.e. the k-loop
// runs at ~90 % of the pipe and the whole kernel at 79 % of the measured TF32 peak in MMA passes.
//
// Things tried here and dropped because they measured no better (round-1 notes, profiles/):
//   * mbarrier.arrive.release.cluster for the cross-CTA "tiles ready" signal: a cluster-scope MEMBAR per
//     arrive, 37 % of the splitter warps' samples (ncu) -> plain arrive, as CUTLASS's ClusterBarrier does;
//   * decoupled raw / lo rings (4 x 32 KB of TMA in flight, 2 lo slots): 64 us (lo ring serialises the splitter);
//   * A operand from tensor memory (tcgen05.mma TS form, 256 x 128 cluster tiles): correct, 70 us on FFN-1
//     (more L2 traffic per FLOP), 1-3 % faster only for N = 256;
//   * 8 splitter warps, TMA L2 prefetch of the next tile: no change.
//   * (round 2) plain epilogue with coalesced st.global read back from the staging tile (8 lanes per 128-byte row
//     segment, no fence.proxy.async, no wait for the TMA engine): correct, 46.8 -> 50.9 us on FFN-1, every shape
//     slower (profiles/r2_gemm_bound_ab.txt) -> the TMA store stays.
//   * (round 2) what bounds the forward / dgrad kernel, from timing builds with parts switched off (same file): no
//     epilogue stores 47.0 -> 36.2 us on FFN-1 (N = 1024 outputs), 46.6 -> 42.5 on FFN-2; no A loads 42.1 / 41.5; neither
//     33.9 / 39.6; 3 stages instead of 4: 50.5 / 50.1, 2 stages 62.2 / 67.1; no A split 46.0 / 44.0.
//   * bf16 hi / lo images of the weights prepared once per step and TMA-loaded into the stage's bf16 B tiles
//     (the splitter then converts only A): with the B conversion merely skipped the kernel is 8-13 % faster,
//     but the extra 16 KB of TMA writes per stage give it back (FFN-1 58.4 -> 58.1 us, dgrads 2-7 %, step 0.5 %).
//
//   cluster tile = 256 (M) x 256 (N); CTA r of the pair owns rows [128 r, 128 r + 128) of A and of the
//   accumulator (its own TMEM) and stages columns [128 r, 128 r + 128) of B.
//   warp 0   TMA producer (each CTA: its A rows, its B half -> local full barrier)
//   warps 6-9 splitter   (each CTA: lo = x - trunc_tf32(x) for its tiles; one arrive per warp on the
//                         LEADER's xfull barrier, remote for CTA 1)
//   warp 1   MMA issuer  (leader CTA only): tcgen05.mma.cta_group::2 256x256x8, commits multicast to
//                         both CTAs' empty / tfull barriers
//   warps 2-5 epilogue   (each CTA drains its 128 TMEM lanes; tempty arrives go to the leader)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cuda_bf16.h>

#include "skf_common.cuh"
#include "skf_ptx.cuh"

#ifndef SKF_SPLIT_UNROLL_N
#define SKF_SPLIT_UNROLL_N 8
#endif
constexpr int kSplitUnroll3 = SKF_SPLIT_UNROLL_N;   // float4s in flight per splitter thread

namespace skf {

using namespace ptx;

int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert);
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);

constexpr int T3_BM = 128;                 // rows per CTA (256 per cluster)
constexpr int T3_BK = 32;
constexpr int T3_A_BYTES = T3_BM * 128;    // 16 KB
constexpr int T3_EPI_TILE = 32 * 128;
constexpr int T3_EPI_WARPS = 8;            // two warps per TMEM lane quarter: even / odd 32-column chunks
#ifndef SKF_T3_EPI_BUFS
#define SKF_T3_EPI_BUFS 1
#endif
#ifndef SKF_T3_SPLIT_WARPS
#define SKF_T3_SPLIT_WARPS 4
#endif
constexpr int T3_EPI_BUFS = SKF_T3_EPI_BUFS;
constexpr int T3_SPLIT_WARPS = SKF_T3_SPLIT_WARPS;     // splitter warps per CTA (the fused column sum assumes 4)
constexpr int T3_EPI_BYTES = T3_EPI_WARPS * T3_EPI_BUFS * T3_EPI_TILE;   // 32 KB
constexpr int T3_SPLIT_T0 = 32 * (2 + T3_EPI_WARPS);   // first splitter thread
constexpr int T3_THREADS = T3_SPLIT_T0 + 32 * T3_SPLIT_WARPS;   // 448
#ifndef SKF_T3_MAX_STAGES
#define SKF_T3_MAX_STAGES 4
#endif
constexpr int T3_MAX_STAGES = SKF_T3_MAX_STAGES;

// BN = columns per cluster tile (256, or 128 for N <= 256: twice as many tiles for the 74 clusters and a
// fourth pipeline stage); each CTA stages BN / 2 columns of B.
// WS (weights pre-split, bf16x3 only): the B operand arrives as bf16 hi / lo tiles straight from its HBM images, so
// a stage holds A raw | A hi | A lo | B hi | B lo = 48 KB / 40 KB (four stages) and the splitter converts only A.
template <int BN, bool WS = false>
struct Tc3Cfg {
  static constexpr int BNH = BN / 2;
  static constexpr int B_BYTES = BNH * 128;
  static constexpr int RAW = T3_A_BYTES + B_BYTES;      // bytes the TMA delivers per stage (WS: B hi + B lo = B_BYTES)
  static constexpr int STAGE = WS ? 2 * T3_A_BYTES + B_BYTES : 2 * RAW;   // raw + lo tiles: 64 KB / 48 KB
  static constexpr int STAGES_RAW = (227 * 1024 - 2048 - T3_EPI_BYTES - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > T3_MAX_STAGES ? T3_MAX_STAGES : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + T3_EPI_BYTES + 1024 + 256 + 512 /* column sums */;
  static constexpr int TMEM_COLS = 2 * BN;              // two accumulators
};

struct Gemm3Params {
  int M, N, K;
  int k_chunk;
  int mt, nt, splits;   // cluster-tile grid (mt counts 256-row blocks)
  int pdl_late;         // the TMA producer releases the next kernel after its last load instead of after the prologue
  // Tail split: the last `ntiles - tail_first` cluster tiles are done as two half-width units each (columns
  // [n0, n0 + BN/2) and [n0 + BN/2, n0 + BN)), so that e.g. 100 tiles on 74 clusters are 74 full units and one SHORT
  // wave of 52 half units instead of two full waves; units = tail_first + 2 (ntiles - tail_first).
  int tail_first, nunits;
  skf_gemm_epilogue e;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Plain (release.cta) arrive on a barrier of another CTA of the cluster, as cutlass::arch::ClusterBarrier::arrive
// does.  The data being published was written to THIS CTA's shared memory and made visible to the async proxy by
// fence.proxy.async; a .release.cluster qualifier here costs a cluster-scope MEMBAR per arrive (measured: 37 % of
// the splitter warps' time).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 x bf16 -> fp32 product of the mixed scheme (kind::f16), same accumulator as the tf32 MMAs
__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor for kind::f16 with bf16 operands, fp32 accumulate (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// bf16 operand tiles written by the splitter: 64-byte rows (32 elements) under the 64-byte swizzle
// (cute Swizzle<2,4,3>: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)).
//   K-major : rows = M/N index, 8-row groups 512 B apart (SBO), one MMA K = 16 elements = 32 bytes
//   MN-major: rows = k, 32 MN elements per row, MN chunks of 32 `chunk` bytes apart (LBO), 8-k groups 512 B apart
__device__ __forceinline__ uint64_t make_desc16_kmajor(uint32_t addr) { return make_smem_desc(addr, 16, 512, 4); }
__device__ __forceinline__ uint64_t make_desc16_mnmajor(uint32_t addr, uint32_t chunk) {
  return make_smem_desc(addr, chunk, 512, 4);
}
// 8 consecutive fp32 values -> 8 bf16 of x (hi) and 8 bf16 of the residual (lo), 16 bytes each.
//   BF3 = false: lo = x - trunc_tf32(x)   (the main term is a TF32 MMA on the raw tile)
//   BF3 = true : lo = x - bf16(x)         (the main term is the bf16 MMA hi * hi)
template <bool BF3>
__device__ __forceinline__ void t3_to_bf16_pair(const float4& u, const float4& v, uint4& hi, uint4& lo) {
  auto pk = [](float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
  };
  hi = make_uint4(pk(u.x, u.y), pk(u.z, u.w), pk(v.x, v.y), pk(v.z, v.w));
  if (BF3) {
    // a bf16 is the top half of an fp32: the residuals x - float(bf16(x)) are exact in fp32
    auto l2 = [&](float a, float b, uint32_t h) {
      return pk(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xFFFF0000u));
    };
    lo = make_uint4(l2(u.x, u.y, hi.x), l2(u.z, u.w, hi.y), l2(v.x, v.y, hi.z), l2(v.z, v.w, hi.w));
  } else {
    auto l = [](float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); };
    lo = make_uint4(pk(l(u.x), l(u.y)), pk(l(u.z), l(u.w)), pk(l(v.x), l(v.y)), pk(l(v.z), l(v.w)));
  }
}
// One operand tile of the mixed scheme: fp32 tile `src` (K-major: ROWS rows of 128 B under the 128B swizzle;
// MN-major: ROWS/32 boxes of 32 k-rows under the 32B-atom swizzle) -> bf16 hi / lo tiles with 64-byte rows.
// 32 * T3_SPLIT_WARPS threads, thread `tid`; tasks = ROWS * 4 groups of 8 elements.
template <int ROWS, bool MN, bool BF3>
__device__ __forceinline__ void t3_split_bf16(const uint8_t* src, uint8_t* hi16, uint8_t* lo16, int tid) {
#pragma unroll
  for (int task = tid; task < ROWS * 4; task += 32 * T3_SPLIT_WARPS) {
    const int r = task >> 2, g = task & 3;            // r: row inside the tile (box-major for MN), g: 8-element group
    const int rs = MN ? (r & 31) : r;                 // row index the swizzles see (boxes are 4 KB / 2 KB aligned)
    const uint32_t sx = MN ? (uint32_t)((rs & 3) << 1) : (uint32_t)(rs & 7);
    const uint8_t* rowp = src + r * 128;
    // the 4 lanes of a row read its 4 even chunks, then its 4 odd ones; under the 128B swizzle consecutive rows
    // flip parity, but the 32B-atom swizzle keeps it, so there odd rows read their odd chunks first (two rows of a
    // quarter-warp then never meet in a bank)
    const uint32_t first = (MN && (r & 1)) ? 1u : 0u;
    const float4 w0 = *reinterpret_cast<const float4*>(rowp + (((2 * g + first) ^ sx) << 4));
    const float4 w1 = *reinterpret_cast<const float4*>(rowp + (((2 * g + 1 - first) ^ sx) << 4));
    const float4 u = first ? w1 : w0, v = first ? w0 : w1;
    uint4 h, l;
    t3_to_bf16_pair<BF3>(u, v, h, l);
    const uint32_t dof = r * 64 + ((g ^ ((rs >> 1) & 3)) << 4);
    *reinterpret_cast<uint4*>(hi16 + dof) = h;
    *reinterpret_cast<uint4*>(lo16 + dof) = l;
  }
}
// arrives (once the MMAs issued so far have completed) on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void t3_tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void t3_tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void t3_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t3_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float t3_trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// MIX: the two small cross terms of the 3xTF32 product run as bf16 MMAs (kind::f16, twice the TF32 rate):
//   A*B ~= tf32(A)*tf32(B) [kind::tf32 on the raw tiles] + bf16(A)*bf16(Bl) + bf16(Al)*bf16(B) [kind::f16],
// Xl = X - trunc_tf32(X).  A cross term is ~2^-11 of the product, so bf16's 2^-9 on its factors costs 2^-20:
// the result keeps 3xTF32 accuracy for two TF32-pass equivalents instead of three.  The splitter writes four
// bf16 tiles (hi/lo of A and of this CTA's B half, 32 KB) where the plain scheme keeps two fp32 lo tiles.
// MIX = 2 (bf16x3): all three terms as bf16 MMAs on the splitter's hi / lo tiles,
//   A*B ~= bf16(A)*bf16(B) + bf16(A)*bf16(Bl) + bf16(Al)*bf16(B),  Xl = X - bf16(X)  (exact in fp32).
// Dropped: Al*Bl and the bf16 rounding of the two residuals, each <= 2^-18 of the product (3 * 2^-18 = 1.1e-5 worst
// case per product, random sign) -- far below the 1.4e-4 end-to-end floor set by attention's TF32-rounded P -- for
// three bf16-rate passes instead of one TF32 + two bf16 (= four bf16-rate passes), and the tensor cores read only
// the 32 KB of bf16 tiles per stage (the raw fp32 tiles feed the splitter alone).
template <int T3_BN, bool A_MN, bool B_MN, int MIX, bool WS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
gemm3_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
        const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmY, const Gemm3Params p) {
  static_assert(!WS || MIX == 2, "pre-split weights are bf16 images: bf16x3 mode only");
  using C = Tc3Cfg<T3_BN, WS>;
  // offset of the bf16 tiles (A hi | A lo | B hi | B lo) inside a stage
  constexpr int T3_B16 = WS ? T3_A_BYTES : C::RAW;
  constexpr int S = C::STAGES;
  constexpr int T3_BNH = C::BNH, T3_RAW = C::RAW, T3_STAGE = C::STAGE, T3_TMEM_COLS = C::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;      // identical in both CTAs (same kernel, same layout)
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t epi0 = base + S * T3_STAGE;
  const uint32_t bar0 = epi0 + T3_EPI_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };               // local: TMA -> splitter
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };        // local: MMA commit (multicast) -> TMA
  auto xfull_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };    // leader: 8 splitter warps -> MMA
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * S + a); };    // local: MMA commit (multicast) -> epilogue
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * S + 2 + a); };   // leader: 8 epilogue warps -> MMA
  const uint32_t tmem_slot = bar0 + 8u * (3 * S + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + S * T3_STAGE + T3_EPI_BYTES + 8 * (3 * S + 4));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;
  const int ntiles = p.nunits;          // work units: whole cluster tiles, then the half-width units of the tail

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (WS) prefetch_tmap(&tmBl);
    prefetch_tmap(&tmY);
    for (int s = 0; s < S; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(xfull_bar(s), 2 * T3_SPLIT_WARPS);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * T3_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, T3_TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // tensor memory is ours: the next kernel of the stream may become resident (programmatic dependent launch); this
  // one touches global memory only from here on
  if (!p.pdl_late) pdl_trigger();
  pdl_wait();

  // unit u -> rows, first column, k range and width (BN, or BN / 2 for a unit of the split tail)
  auto tile_coords = [&](int u, int& m0, int& n0, int& kb, int& nkb, int& ncols) {
    const int half = u < p.tail_first ? -1 : ((u - p.tail_first) & 1);
    const int t = u < p.tail_first ? u : p.tail_first + ((u - p.tail_first) >> 1);
    const int n = t % p.nt;
    const int r = t / p.nt;
    const int m = r % p.mt;
    const int z = r / p.mt;
    m0 = m * (2 * T3_BM) + (int)rank * T3_BM;     // this CTA's rows
    n0 = n * T3_BN + (half > 0 ? T3_BN / 2 : 0);
    ncols = half < 0 ? T3_BN : T3_BN / 2;
    kb = z * p.k_chunk;
    const int ke = min(p.K, kb + p.k_chunk);
    nkb = (ke - kb + T3_BK - 1) / T3_BK;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cluster_id; t < ntiles; t += nclusters) {
        int m0, n0, kb, nkb, ncols;
        tile_coords(t, m0, n0, kb, nkb, ncols);
        // this CTA's half of the unit's B columns; a half-width unit still loads whole boxes (the MMA reads the
        // first ncols / 2 rows of each CTA's tile)
        const int nh = n0 + (int)rank * (ncols >> 1);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
#ifdef SKF_T3_DEBUG_SKIP_A_LOAD    // timing experiment only (wrong results): the A tile is never loaded
          mbar_arrive_expect_tx(full_bar(s), T3_RAW - T3_A_BYTES);
#else
          mbar_arrive_expect_tx(full_bar(s), T3_RAW);
#endif
          const uint32_t a_s = base + s * T3_STAGE;
          const uint32_t b_s = a_s + T3_A_BYTES;
          const int k0 = kb + i * T3_BK;
#ifndef SKF_T3_DEBUG_SKIP_A_LOAD
          if (!A_MN) {
            tma_load_2d(a_s, &tmA, full_bar(s), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < T3_BM / 32; j++) tma_load_2d(a_s + j * 4096, &tmA, full_bar(s), m0 + 32 * j, k0);
          }
#endif
          if (WS) {
            // bf16 images of the weight: hi / lo tiles with 64-byte rows (64B swizzle), where the splitter would
            // have written them
            const uint32_t bh16 = a_s + T3_B16 + T3_A_BYTES, bl16 = bh16 + C::B_BYTES / 2;
            if (!B_MN) {
              tma_load_2d(bh16, &tmB, full_bar(s), k0, nh);               // box {32 k, BNH n}
              tma_load_2d(bl16, &tmBl, full_bar(s), k0, nh);
            } else {
#pragma unroll
              for (int j = 0; j < T3_BNH / 32; j++) {                     // boxes {32 n, 32 k}, 2 KB each
                tma_load_2d(bh16 + j * 2048, &tmB, full_bar(s), nh + 32 * j, k0);
                tma_load_2d(bl16 + j * 2048, &tmBl, full_bar(s), nh + 32 * j, k0);
              }
            }
          } else if (!B_MN) {
            tma_load_2d(b_s, &tmB, full_bar(s), k0, nh);
          } else {
#pragma unroll
            for (int j = 0; j < T3_BNH / 32; j++) tma_load_2d(b_s + j * 4096, &tmB, full_bar(s), nh + 32 * j, k0);
          }
        }
      }
      if (p.pdl_late) pdl_trigger();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc_full = make_idesc_tf32(2 * T3_BM, T3_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_half = make_idesc_tf32(2 * T3_BM, T3_BN / 2, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0;
      int ti = 0;
      for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
        int m0, n0, kb, nkb, ncols;
        tile_coords(t, m0, n0, kb, nkb, ncols);
        const bool halfw = ncols != T3_BN;
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1u;
        mbar_wait_cluster(tempty_bar(a), aph ^ 1u);     // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(a * T3_BN);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait_cluster(xfull_bar(s), ph);          // both CTAs' raw + lo tiles are in place
          tc_fence_after();
          const uint32_t a_s = base + s * T3_STAGE;
          const uint32_t b_s = a_s + T3_A_BYTES;
          const uint32_t al_s = a_s + T3_RAW;
          const uint32_t bl_s = al_s + T3_A_BYTES;
          const uint32_t idesc = halfw ? idesc_half : idesc_full;
          if (MIX) {
            const uint32_t idesc16 = halfw ? make_idesc_bf16(2 * T3_BM, T3_BN / 2, A_MN ? 1 : 0, B_MN ? 1 : 0)
                                           : make_idesc_bf16(2 * T3_BM, T3_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            const uint32_t ah16 = a_s + T3_B16, al16 = ah16 + T3_A_BYTES / 2;
            const uint32_t bh16 = al16 + T3_A_BYTES / 2, bl16 = bh16 + C::B_BYTES / 2;
#pragma unroll
            for (int k2 = 0; k2 < T3_BK / 16; k2++) {                      // cross terms: bf16, K = 16 per MMA
              const uint32_t ao = A_MN ? k2 * 1024 : k2 * 32;
              const uint32_t bo = B_MN ? k2 * 1024 : k2 * 32;
              const uint64_t ahd = A_MN ? make_desc16_mnmajor(ah16 + ao, 2048) : make_desc16_kmajor(ah16 + ao);
              const uint64_t ald = A_MN ? make_desc16_mnmajor(al16 + ao, 2048) : make_desc16_kmajor(al16 + ao);
              const uint64_t bhd = B_MN ? make_desc16_mnmajor(bh16 + bo, 2048) : make_desc16_kmajor(bh16 + bo);
              const uint64_t bld = B_MN ? make_desc16_mnmajor(bl16 + bo, 2048) : make_desc16_kmajor(bl16 + bo);
              mma2_bf16_ss(acc, ahd, bld, idesc16, (i | k2) != 0 ? 1u : 0u);
              mma2_bf16_ss(acc, ald, bhd, idesc16, 1u);
              if (MIX == 2) mma2_bf16_ss(acc, ahd, bhd, idesc16, 1u);      // bf16x3: main term on the hi tiles
            }
#pragma unroll
            for (int ks = 0; ks < (MIX == 2 ? 0 : T3_BK / 8); ks++) {      // main term: tf32 on the raw tiles
              const uint32_t ao = A_MN ? ks * 1024 : ks * 32;
              const uint32_t bo = B_MN ? ks * 1024 : ks * 32;
              const uint64_t ad = A_MN ? make_desc_mnmajor(a_s + ao, 4096) : make_desc_kmajor(a_s + ao);
              const uint64_t bd = B_MN ? make_desc_mnmajor(b_s + bo, 4096) : make_desc_kmajor(b_s + bo);
              mma2_tf32_ss(acc, ad, bd, idesc, 1u);
            }
          } else {
#pragma unroll
          for (int ks = 0; ks < T3_BK / 8; ks++) {
            const uint32_t ao = A_MN ? ks * 1024 : ks * 32;
            const uint32_t bo = B_MN ? ks * 1024 : ks * 32;
            const uint64_t ad = A_MN ? make_desc_mnmajor(a_s + ao, 4096) : make_desc_kmajor(a_s + ao);
            const uint64_t bd = B_MN ? make_desc_mnmajor(b_s + bo, 4096) : make_desc_kmajor(b_s + bo);
            const uint64_t ald = A_MN ? make_desc_mnmajor(al_s + ao, 4096) : make_desc_kmajor(al_s + ao);
            const uint64_t bld = B_MN ? make_desc_mnmajor(bl_s + bo, 4096) : make_desc_kmajor(bl_s + bo);
            mma2_tf32_ss(acc, ad, bld, idesc, (i | ks) != 0 ? 1u : 0u);   // small terms first
            mma2_tf32_ss(acc, ald, bd, idesc, 1u);
            mma2_tf32_ss(acc, ad, bd, idesc, 1u);
          }
          }
          tc_commit2(empty_bar(s));
        }
        tc_commit2(tfull_bar(a));
      }
    }
  } else if (warp < 2 + T3_EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (8 warps, both CTAs)
    // With K = 256 a tile's k-loop (8 k-blocks) is shorter than draining its 128 x 256 accumulator through four
    // warps, so two warps share each TMEM lane quarter and take alternate 32-column chunks.
    const int q = warp & 3;
    const int ew = warp - 2;
    const int chalf = ew >> 2;             // warps 2..5 take even chunks, 6..9 odd chunks
    const uint32_t stg0 = epi0 + (uint32_t)(ew * T3_EPI_BUFS) * T3_EPI_TILE;
    uint8_t* stg0_gen = gen + S * T3_STAGE + (ew * T3_EPI_BUFS) * T3_EPI_TILE;
    const skf_gemm_epilogue& e = p.e;
    int ti = 0;
    uint32_t nstore = 0;
    for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
      int m0, n0, kb, nkb, ncols;
      tile_coords(t, m0, n0, kb, nkb, ncols);
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1u;
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      const int mrow0 = m0 + q * 32;
      const int m = mrow0 + lane;
      const bool row_ok = m < p.M;
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * T3_BN);
      const int NCH = ncols / 32;
#pragma unroll 1
      for (int c = chalf; c < NCH; c += 2) {
        const int nb = n0 + c * 32;
        const bool live = nb < p.N && mrow0 < p.M;
        uint32_t r[32];
        if (live) {
          tmem_ld32(acc + (uint32_t)(c * 32), r);
          tmem_ld_wait();
        }
        if (c >= NCH - 2) {                // this warp's last chunk of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(a), 0));
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
#ifdef SKF_T3_DEBUG_SKIP_EPI_STORE   // timing experiment only (no output): the tile is read from tensor memory and dropped
        if (v[0] == 123.456f && v[31] == 654.321f) asm volatile("trap;");
        continue;
#endif
        const uint32_t buf = nstore % T3_EPI_BUFS;
        if (lane == 0) t3_bulk_wait_read<T3_EPI_BUFS - 1>();
        __syncwarp();
        uint8_t* srow = stg0_gen + buf * T3_EPI_TILE + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<float4*>(srow + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const uint32_t src = stg0 + buf * T3_EPI_TILE;
          if (e.accumulate) t3_tma_reduce_add_2d(&tmY, src, nb, mrow0);
          else t3_tma_store_2d(&tmY, src, nb, mrow0);
          t3_bulk_commit();
        }
        nstore++;
      }
    }
    if (lane == 0) t3_bulk_wait_read<0>();
  } else {
    // ------------------------------------------------------------------ splitter (4 warps, both CTAs)
    const int tid = threadIdx.x - T3_SPLIT_T0;
    uint32_t it = 0;
    uint32_t xfull_leader[S];
#pragma unroll
    for (int s = 0; s < S; s++) xfull_leader[s] = map_to_cta(xfull_bar(s), 0);
    // Fused bias gradient (weight-gradient products, B = dY row-major => MN-major tiles): the splitter reads every
    // element of the B tile anyway.  Thread `tid` always meets the same four columns of each 32-column box (chunk
    // tid + 128 i' of the B region is row tid/8 + 16 (i' & 1), box i' / 2, logical 16-byte chunk
    // q = (tid & 7) ^ (((tid >> 3) & 3) << 1) under the 32B-atom swizzle), so it keeps one float4 per box.
    constexpr int NBOX = T3_BNH / 32;
    float4 csum[NBOX];
#pragma unroll
    for (int j = 0; j < NBOX; j++) csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float* colsum_s = reinterpret_cast<float*>(gen + S * T3_STAGE + T3_EPI_BYTES + 256);
    const int cq = (tid & 7) ^ (((tid >> 3) & 3) << 1);
    for (int t = cluster_id; t < ntiles; t += nclusters) {
      int m0, n0, kb, nkb, ncols;
      tile_coords(t, m0, n0, kb, nkb, ncols);
      const bool do_colsum = B_MN && p.e.colsum != nullptr && m0 == (int)rank * T3_BM;   // row-block 0 only
      for (int i = 0; i < nkb; i++, it++) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(full_bar(s), ph);
        const float4* src = reinterpret_cast<const float4*>(gen + s * T3_STAGE);
        float4* dst = reinterpret_cast<float4*>(gen + s * T3_STAGE + T3_RAW);
        if (MIX) {
          uint8_t* st8 = gen + s * T3_STAGE;
          uint8_t* ah16 = st8 + T3_B16;
          uint8_t* al16 = ah16 + T3_A_BYTES / 2;
          uint8_t* bh16 = al16 + T3_A_BYTES / 2;
          uint8_t* bl16 = bh16 + C::B_BYTES / 2;
#ifndef SKF_T3_DEBUG_SKIP_ASPLIT   // timing experiment only (wrong results): upper bound of what activation images
                                   // written by the producing kernels would buy
          t3_split_bf16<T3_BM, A_MN, MIX == 2>(st8, ah16, al16, tid);
#endif
          if (!WS) t3_split_bf16<T3_BNH, B_MN, MIX == 2>(st8 + T3_A_BYTES, bh16, bl16, tid);
          if (!WS && B_MN && do_colsum) {
#pragma unroll
            for (int ib = 0; ib < 2 * NBOX; ib++) {
              const float4 x = src[T3_A_BYTES / 16 + tid + 128 * ib];
              csum[ib >> 1].x += x.x; csum[ib >> 1].y += x.y; csum[ib >> 1].z += x.z; csum[ib >> 1].w += x.w;
            }
          }
        } else {
#pragma unroll kSplitUnroll3
        for (int k = tid; k < T3_A_BYTES / 16; k += 128) {
          const float4 x = src[k];
          float4 l;
          l.x = x.x - t3_trunc_tf32(x.x); l.y = x.y - t3_trunc_tf32(x.y);
          l.z = x.z - t3_trunc_tf32(x.z); l.w = x.w - t3_trunc_tf32(x.w);
          dst[k] = l;
        }
#pragma unroll
        for (int ib = 0; ib < 2 * NBOX; ib++) {
          const int k = T3_A_BYTES / 16 + tid + 128 * ib;
          const float4 x = src[k];
          float4 l;
          l.x = x.x - t3_trunc_tf32(x.x); l.y = x.y - t3_trunc_tf32(x.y);
          l.z = x.z - t3_trunc_tf32(x.z); l.w = x.w - t3_trunc_tf32(x.w);
          dst[k] = l;
          if (B_MN && do_colsum) {
            csum[ib >> 1].x += x.x; csum[ib >> 1].y += x.y; csum[ib >> 1].z += x.z; csum[ib >> 1].w += x.w;
          }
        }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          uint32_t target = xfull_leader[0];
#pragma unroll
          for (int ss = 1; ss < S; ss++) target = (s == ss) ? xfull_leader[ss] : target;
          mbar_arrive_cluster(target);
        }
      }
      if (B_MN && do_colsum) {
        // 16 threads share each column: reduce in shared memory, then one global atomic per column and CTA
        if (tid < T3_BNH) colsum_s[tid] = 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < NBOX; j++) {
          float* cs = colsum_s + j * 32 + cq * 4;
          atomicAdd(cs, csum[j].x); atomicAdd(cs + 1, csum[j].y); atomicAdd(cs + 2, csum[j].z); atomicAdd(cs + 3, csum[j].w);
          csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int col = n0 + (int)rank * T3_BNH + tid;
        if (tid < T3_BNH && col < p.N) atomicAdd(p.e.colsum + col, colsum_s[tid]);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();          // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) tmem_dealloc2(tmem_base, T3_TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------------
static int g3_clusters = 0;
int g_gemm_mix = 2;       // 1: cross terms of the 3xTF32 product as bf16 MMAs, 2: all three terms bf16 (see gemm3_k MIX)

int g_gemm_wsplit = 1;    // 1: the plan hands forward / dgrad products bf16 images of their weights (b_hi16 / b_lo16)

int g_t3_oversubscribe = 0;
int g_t3_tail_split = 1;  // 1: the tiles of a partial last wave are done as two half-width units each (Gemm3Params::tail_first)

template <int BN, bool A_MN, bool B_MN, int MIX, bool WS = false>
static int launch_tc3(const CUtensorMap* tm, const Gemm3Params& p, cudaStream_t st) {
  constexpr int T3_SMEM = Tc3Cfg<BN, WS>::SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    SKF_CUDA(cudaFuncSetAttribute(gemm3_k<BN, A_MN, B_MN, MIX, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM));
    attr_done = true;
  }
  if (!g3_clusters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t qc;
    memset(&qc, 0, sizeof(qc));
    qc.gridDim = dim3(sms > 1 ? (sms & ~1) : 2);
    qc.blockDim = dim3(T3_THREADS);
    qc.dynamicSmemBytes = T3_SMEM;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, gemm3_k<BN, A_MN, B_MN, MIX, WS>, &qc) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = sms / 2;
    }
    g3_clusters = nc < 1 ? 1 : nc;
  }
  int ntiles = p.mt * p.nt * p.splits;
  Gemm3Params q = p;
  q.tail_first = ntiles;
  // a partial last wave of whole tiles costs a full tile time: when its tiles, halved, still fit one wave, halve them
  // (forward / dgrad products only: split-K and the fused column sum keep whole tiles)
  if (g_t3_tail_split && !g_t3_oversubscribe && p.splits == 1 && !p.e.accumulate && ntiles > g3_clusters) {
    const int rem = ntiles % g3_clusters;
    if (rem > 0 && 2 * rem <= g3_clusters) q.tail_first = ntiles - rem;
  }
  q.nunits = q.tail_first + 2 * (ntiles - q.tail_first);
  ntiles = q.nunits;
  // g_t3_oversubscribe (set by the plan around launches on its low-priority side stream): one cluster per work
  // unit instead of a persistent grid, so that the units are short and a higher-priority kernel waiting for SMs
  // gets them after one unit, not after a cluster's whole share
  const int ncl = (ntiles < g3_clusters || g_t3_oversubscribe) ? ntiles : g3_clusters;
  launch_k(gemm3_k<BN, A_MN, B_MN, MIX, WS>, dim3(2 * ncl), dim3(T3_THREADS), T3_SMEM, st, g_t3_oversubscribe ? 0 : PDL_GEMM, tm[0], tm[1], tm[3], tm[2], q);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int gemm_tc3_clusters() { return g3_clusters; }

// The pair kernel needs more than one 128-row block and more than 64 columns; everything else stays on the
// single-CTA kernel.  Measured on B200 (tools/prof_ops.py, M = 25600, 3xTF32): FFN-1 79 -> 61 us, FFN-2 dgrad
// 76 -> 60 us, QKV wgrad 54 -> 44 us.
int gemm_tc3_wanted(int M, int N, int K) { (void)K; return N > 64 && M > 128; }
// 128-column cluster tiles only when the problem has no more than 128 columns: for N = 256 the 256-wide tile
// (100 tiles on 74 clusters) measured 3 % faster than 2 x 128 (FFN-2 25600x1024x256: 69.6 vs 71.6 us).
int g_t3_bn128_max = 128;
int gemm_tc3_bn(int N, int accumulate) { (void)accumulate; return N <= g_t3_bn128_max ? 128 : 256; }

int gemm_tc3(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* Y,
             int64_t ldy, int M, int N, int K, const skf_gemm_epilogue* epi, cudaStream_t st) {
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  if (epi) e = *epi;
  int split = e.split_k > 1 ? e.split_k : 1;
  SKF_REQUIRE(split == 1 || e.accumulate, SKF_ERR_BAD_ARG, "gemm: split_k needs accumulate");
  SKF_REQUIRE(!(e.accumulate && (e.bias || e.resid || e.mask || e.act)), SKF_ERR_BAD_ARG,
              "gemm: accumulate excludes other epilogue terms");
  const bool a_mn = !(sa_k == 1);
  const bool b_mn = !(sb_k == 1);
  Gemm3Params p;
  p.M = M; p.N = N; p.K = K;
  p.k_chunk = cdiv(cdiv(K, split), T3_BK) * T3_BK;
  p.splits = cdiv(K, p.k_chunk);
  const int BN = gemm_tc3_bn(N, e.accumulate);
  p.mt = cdiv(M, 2 * T3_BM);
  p.nt = cdiv(N, BN);
  SKF_REQUIRE(!e.colsum || (e.accumulate && b_mn), SKF_ERR_BAD_ARG,
              "gemm: the fused column sum needs accumulate=1 and a row-major [K, N] B operand");
  p.e = e;
  p.e.a_lo = p.e.b_lo = nullptr;
  p.pdl_late = (g_pdl & PDL_LATE) != 0;
  CUtensorMap tm[4];   // A, B (or B hi), Y, B lo
  // pre-split weights: bf16 images of B with B's strides (bf16x3 mode, rows a multiple of 8 elements apart)
  const bool ws = g_gemm_mix == 2 && e.b_hi16 && e.b_lo16 && !e.accumulate &&
                  (((reinterpret_cast<uintptr_t>(e.b_hi16) | reinterpret_cast<uintptr_t>(e.b_lo16)) & 15) == 0) &&
                  ((b_mn ? sb_k : sb_n) % 8 == 0);
  // raw fp32 bits in shared memory (the splitter needs them): plain FLOAT32 maps
  if (!a_mn) SKF_TRY(make_tmap_2d_typed(&tm[0], A, (uint64_t)K, (uint64_t)M, (uint64_t)(M == 1 ? K : sa_m), T3_BK, T3_BM, false, false));
  else SKF_TRY(make_tmap_2d_typed(&tm[0], A, (uint64_t)M, (uint64_t)K, (uint64_t)sa_k, 32, T3_BK, true, false));
  if (ws) {
    if (!b_mn) {
      SKF_TRY(make_tmap_2d_bf16(&tm[1], e.b_hi16, (uint64_t)K, (uint64_t)N, (uint64_t)sb_n, T3_BK, (uint32_t)(BN / 2)));
      SKF_TRY(make_tmap_2d_bf16(&tm[3], e.b_lo16, (uint64_t)K, (uint64_t)N, (uint64_t)sb_n, T3_BK, (uint32_t)(BN / 2)));
    } else {
      SKF_TRY(make_tmap_2d_bf16(&tm[1], e.b_hi16, (uint64_t)N, (uint64_t)K, (uint64_t)sb_k, 32, T3_BK));
      SKF_TRY(make_tmap_2d_bf16(&tm[3], e.b_lo16, (uint64_t)N, (uint64_t)K, (uint64_t)sb_k, 32, T3_BK));
    }
  } else {
    if (!b_mn) SKF_TRY(make_tmap_2d_typed(&tm[1], B, (uint64_t)K, (uint64_t)N, (uint64_t)sb_n, T3_BK, (uint32_t)(BN / 2), false, false));
    else SKF_TRY(make_tmap_2d_typed(&tm[1], B, (uint64_t)N, (uint64_t)K, (uint64_t)sb_k, 32, T3_BK, true, false));
    tm[3] = tm[1];
  }
  SKF_TRY(make_tmap_2d_typed(&tm[2], Y, (uint64_t)N, (uint64_t)M, (uint64_t)ldy, 32, 32, false, false));
#define SKF_T3_DISPATCH(BNV, MIXV)                                                          \
  do {                                                                                     \
    if (!a_mn && !b_mn) return launch_tc3<BNV, false, false, MIXV>(tm, p, st);             \
    if (!a_mn && b_mn) return launch_tc3<BNV, false, true, MIXV>(tm, p, st);               \
    if (a_mn && !b_mn) return launch_tc3<BNV, true, false, MIXV>(tm, p, st);               \
    return launch_tc3<BNV, true, true, MIXV>(tm, p, st);                                   \
  } while (0)
  if (ws) {
#define SKF_T3_DISPATCH_WS(BNV)                                                            \
  do {                                                                                     \
    if (!a_mn && !b_mn) return launch_tc3<BNV, false, false, 2, true>(tm, p, st);          \
    if (!a_mn && b_mn) return launch_tc3<BNV, false, true, 2, true>(tm, p, st);            \
    if (a_mn && !b_mn) return launch_tc3<BNV, true, false, 2, true>(tm, p, st);            \
    return launch_tc3<BNV, true, true, 2, true>(tm, p, st);                                \
  } while (0)
    if (BN == 128) SKF_T3_DISPATCH_WS(128);
    SKF_T3_DISPATCH_WS(256);
#undef SKF_T3_DISPATCH_WS
  }
  if (g_gemm_mix == 2) {
    if (BN == 128) SKF_T3_DISPATCH(128, 2);
    SKF_T3_DISPATCH(256, 2);
  }
  if (g_gemm_mix) {
    if (BN == 128) SKF_T3_DISPATCH(128, 1);
    SKF_T3_DISPATCH(256, 1);
  }
  if (BN == 128) SKF_T3_DISPATCH(128, 0);
  SKF_T3_DISPATCH(256, 0);
#undef SKF_T3_DISPATCH
}

}  // namespace skf
