// Scaled-dot-product attention (builders/utils.py:71-105 with split_heads, transformer.py:160-165) and its backward
// for the shapes of the benchmarked configuration — head dim 32, sequences up to 224 (cfg2: 200 / 199) — as
// warp-specialised persistent tcgen05 kernels.  A work item is one (sample, head) pair; with T <= 224 the WHOLE score
// row of a query fits in tensor memory, so there is no online softmax, no rescaling of the output accumulator and no
// key-block loop.
//
//   warp 0        TMA producer   fp32 [64 rows x 32] boxes of Q / K / V (/ dO / O) -> a 4-stage ring (3-D tensor maps
//                                [B][T][H*dh]: rows past the sequence are zero-filled, never the next sample's rows)
//   warps 2..5    converters     ring stage -> bf16 hi / lo operand tiles (64-byte rows, 64B swizzle); operands whose
//                                last reader runs late in an item are double-buffered, so the next item's operands
//                                land while this item's math runs
//   warp 1        MMA issuer     tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), SS form for the score products
//                                (A, B K-major tiles), TS form for the output products (A = P / dS from tensor memory,
//                                B the MN-major view of the same tiles)
//   warps 8..15   math           two warpgroups, one thread per TMEM lane: mask in registers, exp2, row statistics,
//                                P / dS written back over the scores as packed bf16 hi / lo, epilogues
//
//   fwd : lanes = queries; warpgroup t owns query tile t (rows 128 t ..): S_t = Q_t K^T (N = T), softmax,
//         O_t = P_t V.  The two tiles ping-pong and the next item's S MMAs are issued between this item's P V products.
//   dq  : lanes = queries; warpgroup w owns key half w (112 columns) of every query tile: S, dP = dO V^T,
//         dS = P (dP - delta), dQ += dS K (one accumulator fed by both halves); also delta = rowsum(dO * O).
//   dkv : lanes = keys; warpgroup w owns query half w of every key tile: S^T = K Q^T, dP^T = V dO^T,
//         dV += P^T dO, dK += dS^T Q.
//
// Measured on B200 (tools/probe/mma_probe.cu, profiles/r2_mma_probe.log): a tcgen05.mma costs ~48 clk to issue
// whatever its size below N = 96 and ~135 clk at N = 208, so the kernels minimise the NUMBER of MMAs: where tensor
// memory allows, [Bh | Bl] are ONE N = 64 MN-major operand whose two accumulator halves the epilogue sums.
//
// Precision: every product is the error-compensated bf16x3 sum (hi*hi + hi*lo + lo*hi with lo = bf16(x - bf16(x)),
// 16 mantissa bits per operand, fp32 accumulation) — the scheme of the GEMMs (skf_gemm_tc3.cu).  Masks follow the
// reference exactly: -1e9 is ADDED to the scaled logit of a masked key (builders/utils.py:93-97).
#include <cuda_bf16.h>

#include "skf_common.cuh"
#include "skf_ptx.cuh"

namespace skf {
using namespace ptx;

int make_tmap_3d_f32(CUtensorMap* out, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                     uint64_t ld2, uint32_t b0, uint32_t b1);

namespace a3 {

constexpr int DH = 32;
constexpr int MAXT = 224;                 // 2 S tiles of MAXT columns + 64 accumulator columns = 512 TMEM columns
constexpr int HALF = MAXT / 2;            // backward: columns per warpgroup
constexpr int NST = 8;                    // ring stages at most (barrier slots); Args::nst are used
constexpr int STAGE_BYTES = 64 * 128;     // 64 rows x 32 fp32
constexpr int NCONV = 4;                  // converter warps (warps 2..5); 256 % (32 NCONV) == 0
constexpr int MATH0 = 8;                  // first math warp (a multiple of 4: warp % 4 == TMEM lane quarter); 6, 7 idle
constexpr int THREADS = 32 * (MATH0 + 8);
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASK_NEG = -1e9f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t desc16_k(uint32_t addr) { return make_smem_desc(addr, 16, 512, 4); }
__device__ __forceinline__ uint64_t desc16_mn(uint32_t addr, uint32_t chunk) { return make_smem_desc(addr, chunk, 512, 4); }
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
// 16 columns into the first half of a 32-register buffer (no address-of: the buffer stays in registers)
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// hi = bf16(x) pairs, lo = bf16(x - hi) pairs (a bf16 is the top half of an fp32, so x - hi is exact in fp32)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}
// The same split on the integer pipe for the math warps (F2FP shares the quarter-rate conversion / MUFU path with ex2,
// which bounds those warps): hi = x truncated to bf16 (top 16 bits), lo = bf16(x - hi) rounded half-up; x - hi is
// exact and has the sign of x, so hi + lo = x (1 +- 2^-17), unbiased.
__device__ __forceinline__ void split_pair_int(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  const float la = a - __uint_as_float(ua & 0xFFFF0000u), lb = b - __uint_as_float(ub & 0xFFFF0000u);
  lo = __byte_perm(__float_as_uint(la) + 0x8000u, __float_as_uint(lb) + 0x8000u, 0x7632);
}
// Operand tiles (converter warps): the same on the integer pipe with hi ROUNDED (half-up on the magnitude), so that the
// residual is unbiased and at most 2^-9 |x| and the dropped lo * lo term stays at 2^-18 of a product.
__device__ __forceinline__ void split_pair_rn(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
  hi = __byte_perm(ua, ub, 0x7632);
  const float la = a - __uint_as_float(ua & 0xFFFF0000u), lb = b - __uint_as_float(ub & 0xFFFF0000u);
  lo = __byte_perm(__float_as_uint(la) + 0x8000u, __float_as_uint(lb) + 0x8000u, 0x7632);
}
// 16 fp32 values -> 8 packed hi words followed by 8 packed lo words (the A operand of one K = 16 MMA step, twice)
__device__ __forceinline__ void split16(const float (&p)[16], uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 8; j++) split_pair_int(p[2 * j], p[2 * j + 1], w[j], w[8 + j]);
}

// One ring stage (64 rows x 32 fp32, 128-byte rows under the TMA's 128B swizzle) -> rows [row0, row0 + 64) of a bf16
// hi / lo operand tile pair (64-byte rows under the 64B swizzle: 16-byte chunk g of row R at chunk g ^ ((R >> 1) & 3)).
// Rows >= rows_alloc are not stored.  `ct` = thread index among the NCONV * 32 converter threads.
// WITH_DOT: other[round][0..1] hold this thread's 8 values of O for the task of each round (prefetched from HBM by
// prefetch_rows(): a quad of lanes covers one 128-byte row segment); writes rowsum(x * O) of every row into dot_out
// (shared, indexed by sequence row) and dot_glob (global, may be nullptr; rows >= rows_valid are not written).
constexpr int CONV_ROUNDS = 256 / (NCONV * 32);
template <bool WITH_DOT>
__device__ __forceinline__ void convert_stage(const uint8_t* stage, uint8_t* hi, uint8_t* lo, int row0, int rows_alloc,
                                              int ct, const float4 (*other)[2] = nullptr, float* dot_out = nullptr,
                                              float* dot_glob = nullptr, int rows_valid = 0) {
  // 256 tasks = 64 rows x 4 groups of 8 columns, NCONV * 32 per round; the 4 lanes of a quad share a row
  static_assert(256 % (NCONV * 32) == 0, "converter rounds");
#pragma unroll
  for (int round = 0; round < CONV_ROUNDS; round++) {
    const int task = ct + round * NCONV * 32;
    const int r = task >> 2, g = task & 3, R = row0 + r;
    float dot = 0.f;
    if (R < rows_alloc) {
      const uint32_t sx = (uint32_t)(r & 7);
      const uint32_t o0 = r * 128 + (((2 * g) ^ sx) << 4), o1 = r * 128 + (((2 * g + 1) ^ sx) << 4);
      const float4 u = *reinterpret_cast<const float4*>(stage + o0);
      const float4 v = *reinterpret_cast<const float4*>(stage + o1);
      if (WITH_DOT) {
        const float4 p = other[round][0], q = other[round][1];
        dot = (u.x * p.x + u.y * p.y) + (u.z * p.z + u.w * p.w) + (v.x * q.x + v.y * q.y) + (v.z * q.z + v.w * q.w);
      }
      uint4 h, l;
      split_pair_rn(u.x, u.y, h.x, l.x);
      split_pair_rn(u.z, u.w, h.y, l.y);
      split_pair_rn(v.x, v.y, h.z, l.z);
      split_pair_rn(v.z, v.w, h.w, l.w);
      const uint32_t dof = R * 64 + ((g ^ ((R >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(hi + dof) = h;
      *reinterpret_cast<uint4*>(lo + dof) = l;
    }
    if (WITH_DOT) {
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      if (g == 0 && R < rows_alloc) {
        dot_out[R] = dot;
        if (dot_glob && R < rows_valid) dot_glob[R] = dot;
      }
    }
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void conv_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONV * 32) : "memory"); }

// Optional phase trace (skf_set_option("attn3_trace", 1)): CTA 0 stamps clock64() at the phase boundaries of its first
// 8 items, role by role; tools/attn_trace.py prints the deltas.  [role][item][16 stamps]
__device__ long long g_trace[4 * 8 * 16];
__device__ __forceinline__ void stamp(int on, int role, uint32_t li, int k) {
  if (on && blockIdx.x == 0 && li < 8) g_trace[(role * 8 + li) * 16 + k] = clock64();
}

// This thread's O values for the conversion tasks of rows [row0, row0 + 64) (zeros past rows_valid)
__device__ __forceinline__ void prefetch_rows(float4 (*dst)[2], const float* src, int64_t ld, int row0, int rows_valid,
                                              int ct) {
#pragma unroll
  for (int round = 0; round < CONV_ROUNDS; round++) {
    const int task = ct + round * NCONV * 32;
    const int R = row0 + (task >> 2), g = task & 3;
    dst[round][0] = dst[round][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (R < rows_valid) {
      const float4* p = reinterpret_cast<const float4*>(src + (int64_t)R * ld + 8 * g);
      dst[round][0] = __ldg(p);
      dst[round][1] = __ldg(p + 1);
    }
  }
}

// Effective key length of an item: 1 + the index of its last unpadded key (Tk without a padding mask, or when every
// key is padded — then the reference's softmax over equally shifted logits is uniform and no column may be dropped).
// Keys past it get exactly zero weight in fp32 (exp(x - 1e9 - max) with a finite max), so the kernels skip their
// columns (fwd / dq) and their lane tiles (dkv) altogether; padded keys BEFORE it keep the additive -1e9.
// Computed redundantly by every converter warp (a handful of byte loads per lane); returns the warp-uniform value.
__device__ __forceinline__ int effective_klen(const uint8_t* key_pad, int mode, int64_t row_off, int Tk, int lane) {
  if (mode < 1 || key_pad == nullptr) return Tk;
  int last = 0;
  for (int j = lane; j < Tk; j += 32)
    if (!key_pad[row_off + j]) last = j + 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  return last == 0 ? Tk : last;
}

struct Args {
  int trace;
  int B, H, Tq, Tk, mode;
  int nst;                    // ring stages in use (<= NST): as many as shared memory leaves room for
  float scale;
  const uint8_t* key_pad;     // [B, Tk] or nullptr
  float* out0;                // fwd: O      dq: dQ     dkv: dK
  int64_t ld0;
  float* out1;                // fwd: lse    dq: delta  dkv: dV
  int64_t ld1;
  const float* O;             // dq: forward output [B*Tq, ldO] (for delta = rowsum(dO * O))
  int64_t ldO;
  const float* lse;           // dq / dkv: logsumexp of the forward pass [B, H, Tq]
  const float* delta;         // dkv: rowsum(dO * O) [B, H, Tq] (written by dq)
  // bf16 hi / lo operand images in HBM, one contiguous [TP x 32] tile per (operand, hi|lo, item): the forward kernel
  // writes those of Q, K, V from its shared-memory tiles (img != 0), dq writes those of dO (img_do != 0), and the
  // backward kernels TMA-load them straight into their operand slots instead of pulling fp32 rows through the ring and
  // the converter warps again (direct != 0).  Tile row of (op, hl, item) in the image map: ((2 op + hl) n_items + item) TP
  int img, img_do, direct;
  int pdl_late;               // the TMA producer releases the next kernel after its last load instead of after the prologue
};

// D[128 x N] = Ah Bh^T + Ah Bl^T + Al Bh^T: A rows [a_row0, +128) of the K-major tiles (ah, al), B rows [b_row0, +N)
__device__ __forceinline__ void mma_scores(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int a_row0,
                                           int b_row0, uint32_t idesc) {
  const uint32_t ao = (uint32_t)a_row0 * 64, bo = (uint32_t)b_row0 * 64;
  mma_ss(d, desc16_k(ah + ao), desc16_k(bh + bo), idesc, 0);
  mma_ss(d, desc16_k(ah + ao + 32), desc16_k(bh + bo + 32), idesc, 1);
  mma_ss(d, desc16_k(ah + ao), desc16_k(bl + bo), idesc, 1);
  mma_ss(d, desc16_k(ah + ao + 32), desc16_k(bl + bo + 32), idesc, 1);
  mma_ss(d, desc16_k(al + ao), desc16_k(bh + bo), idesc, 1);
  mma_ss(d, desc16_k(al + ao + 32), desc16_k(bh + bo + 32), idesc, 1);
}
// D[128 x 64] (+)= sum over k-steps of (Ah_s + Al_s) [Bh | Bl]: A in tensor memory, 16 columns per k-step (8 packed hi,
// 8 packed lo), B rows [b_row0 + 16 s, +16) of the MN-major tile pair (bl = bh + chunk bytes); the Al Bl quarter of the
// second MMA is 2^-18 of the product and harmless
__device__ __forceinline__ void mma_av_cat(uint32_t d, uint32_t a_tmem, uint32_t bh, uint32_t chunk, int b_row0,
                                           int ksteps, uint32_t idesc64, uint32_t accumulate) {
  // the start-address field of a descriptor counts 16-byte units: one k-step (16 rows x 64 B) further = + 64
  uint64_t bd = desc16_mn(bh + (uint32_t)b_row0 * 64, chunk);
  uint32_t acc = accumulate, ta = a_tmem;
  for (int s = 0; s < ksteps; s++, bd += 64, ta += 16) {
    mma_ts(d, ta, bd, idesc64, acc);
    mma_ts(d, ta + 8, bd, idesc64, 1);
    acc = 1;
  }
}
// D[128 x 32] (+)= sum over k-steps of Ah_s Bh + Al_s Bh + Ah_s Bl (32 accumulator columns: no room for the
// concatenated form)
__device__ __forceinline__ void mma_av_32(uint32_t d, uint32_t a_tmem, uint32_t bh, uint32_t bl, int b_row0, int ksteps,
                                          uint32_t idesc32, uint32_t accumulate) {
  uint64_t bhd = desc16_mn(bh + (uint32_t)b_row0 * 64, 0), bld = desc16_mn(bl + (uint32_t)b_row0 * 64, 0);
  uint32_t acc = accumulate, ta = a_tmem;
  for (int s = 0; s < ksteps; s++, bhd += 64, bld += 64, ta += 16) {
    mma_ts(d, ta, bhd, idesc32, acc);
    mma_ts(d, ta + 8, bhd, idesc32, 1);
    mma_ts(d, ta, bld, idesc32, 1);
    acc = 1;
  }
}

// shared-memory plan (offsets from the 1024-aligned base): ring | ntiles operand tiles | 1 KB zero pad | vectors | barriers
struct Smem {
  uint32_t ring, tiles, vec, meta, bars;
  uint32_t tile_bytes;          // one bf16 tile: TP x 64 bytes
  uint32_t total;
};
__host__ __device__ inline Smem smem_plan(int TP, int ntiles, int nvec, int nst) {
  Smem s;
  s.ring = 0;
  s.tiles = nst * STAGE_BYTES;
  s.tile_bytes = (uint32_t)TP * 64;
  s.vec = s.tiles + s.tile_bytes * ntiles + 4096;   // 4 KB pad: M = 128 / N reads past the last tile stay inside
  s.meta = s.vec + nvec * 256 * 4;                  // per-buffer item facts (effective key length)
  s.bars = s.meta + 64;
  s.total = s.bars + 64 * 8 + 16;
  return s;
}

struct Ctx {
  uint32_t base;
  uint8_t* gen;
  uint32_t tmem;
  int tid, warp, lane;
};
__device__ __forceinline__ void prologue(Ctx& c, uint8_t* smem_raw) {
  const uint32_t raw = smem_u32(smem_raw);
  c.base = (raw + 1023u) & ~1023u;
  c.gen = smem_raw + (c.base - raw);
  c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
}
__device__ __forceinline__ void alloc_tmem(Ctx& c, uint32_t slot_off, int pdl_late) {
  if (c.warp == 1) {
    tmem_alloc(c.base + slot_off, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *reinterpret_cast<volatile uint32_t*>(c.gen + slot_off);
  // tensor memory is ours: the next kernel of the stream may become resident (programmatic dependent launch); this
  // one touches global memory only from here on
  if (!pdl_late) pdl_trigger();
  pdl_wait();
}
__device__ __forceinline__ void epilogue_free(const Ctx& c) {
  tc_fence_before();
  __syncthreads();
  if (c.warp == 1) tmem_dealloc(c.tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// forward.  tiles: 2 buffers x (Qh Ql Kh Kl Vh Vl).   tmem: S0 [0, 224)  S1 [224, 448)  O [448, 512)
// ---------------------------------------------------------------------------------------------------
namespace fw {
enum { RING_FULL = 0, RING_EMPTY = NST, OP_FULL = 2 * NST /* [2 bufs][4] */, OP_EMPTY = 2 * NST + 8 /* [2] */,
       S_FULL = 2 * NST + 10 /* [2] */, P_FULL = 2 * NST + 12 /* [2 tiles][2 column parts] */, O_FULL = 2 * NST + 16 /* [2] */,
       O_FREE = 2 * NST + 18, COUNT = 2 * NST + 19 };
}

template <bool CAUSAL>
__global__ void __launch_bounds__(THREADS, 1)
attn3_fwd_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
            const __grid_constant__ CUtensorMap mV, const __grid_constant__ CUtensorMap mImg, const Args a) {
  using namespace fw;
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  prologue(c, smem_raw);
  const int QP = (a.Tq + 15) & ~15, KP = (a.Tk + 15) & ~15, TP = QP > KP ? QP : KP;
  const Smem L = smem_plan(TP, 12, 2, a.nst);
  const uint32_t buf_bytes = 6 * L.tile_bytes;
  const int tid = c.tid, warp = c.warp, lane = c.lane;
  const uint32_t base = c.base;
  uint8_t* gen = c.gen;
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int nq_ch = (a.Tq + 63) >> 6, nk_ch = (a.Tk + 63) >> 6;
  const int ntiles = a.Tq > 128 ? 2 : 1;
  const int n_items = a.B * a.H;

  if (tid == 0) {
    for (int i = 0; i < NST; i++) {
      mbar_init(bar(RING_FULL + i), 1);
      mbar_init(bar(RING_EMPTY + i), NCONV);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(bar(OP_FULL + 4 * b + 0), NCONV * nq_ch);   // Q
      mbar_init(bar(OP_FULL + 4 * b + 1), NCONV * nk_ch);   // K
      mbar_init(bar(OP_FULL + 4 * b + 2), NCONV * nk_ch);   // V
      mbar_init(bar(OP_FULL + 4 * b + 3), 1);
      mbar_init(bar(OP_EMPTY + b), 1);
      mbar_init(bar(S_FULL + b), 1);
      mbar_init(bar(P_FULL + 2 * b), 4);
      mbar_init(bar(P_FULL + 2 * b + 1), 4);
      mbar_init(bar(O_FULL + b), 1);
    }
    mbar_init(bar(O_FREE), 4);
    fence_barrier_init();
  }
  for (int i = tid; i < 1024; i += THREADS) reinterpret_cast<uint32_t*>(gen + L.vec - 4096)[i] = 0u;
  alloc_tmem(c, L.bars + 8u * COUNT, a.pdl_late);
  const uint32_t tmem = c.tmem;
  const uint32_t tS[2] = {tmem, tmem + MAXT}, tO = tmem + 2 * MAXT;
  // key columns a query tile needs: causal tiles never look past their last row
  volatile int* klen_s = reinterpret_cast<volatile int*>(gen + L.meta);      // [2 buffers]
  auto ncols = [&](int t, int klen) {
    int n = klen;
    if (CAUSAL) n = min(n, 128 * t + 128);
    return (n + 15) & ~15;
  };
  // P is handed to the tensor core in two column parts so that the first part's P V MMAs run under the softmax of
  // the second: part 0 = columns [0, split(nc)), a multiple of 32
  auto split = [&](int nc) { return min(nc, ((nc + 63) >> 6) << 5); };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      prefetch_tmap(&mQ); prefetch_tmap(&mK); prefetch_tmap(&mV);
      uint32_t n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int b = it / a.H, h = it % a.H;
        for (int op = 0; op < 3; op++) {
          const CUtensorMap* m = op == 0 ? &mQ : (op == 1 ? &mK : &mV);
          const int nch = op == 0 ? nq_ch : nk_ch;
          for (int ch = 0; ch < nch; ch++, n++) {
            const int st = n % a.nst;
            mbar_wait_sleep(bar(RING_EMPTY + st), ((n / a.nst) & 1u) ^ 1u);
            mbar_arrive_expect_tx(bar(RING_FULL + st), STAGE_BYTES);
            tma_load_3d(base + L.ring + st * STAGE_BYTES, m, bar(RING_FULL + st), h * DH, 64 * ch, b);
          }
        }
      }
      if (a.pdl_late) pdl_trigger();
    }
  } else if (warp >= 2 && warp < 2 + NCONV) {
    // ================= converters =================
    const int ct = tid - 64;
    uint32_t n = 0, li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, li++) {
      const int b = it / a.H, buf = li & 1;
      uint8_t* tiles = gen + L.tiles + buf * buf_bytes;
      // the MMAs of the item that used this buffer two items ago have completed
      mbar_wait_sleep(bar(OP_EMPTY + buf), ((li >> 1) & 1u) ^ 1u);
      if (a.img) {                 // ... and the image stores of the earlier items have read their tiles
        if (ct == 0) bulk_wait_read0();
        conv_bar_sync();
      }
      if (ct == 0) stamp(a.trace, 3, li, 0);
      {
        const int kl = effective_klen(a.key_pad, a.mode, (int64_t)b * a.Tk, a.Tk, lane);
        if (ct == 0) klen_s[buf] = kl;
      }
      // key bias of this item (log2 domain): -inf past the sequence, the reference's additive -1e9 on padded keys
      float* bias_s = reinterpret_cast<float*>(gen + L.vec) + buf * 256;
      for (int j = ct; j < 256; j += NCONV * 32) {
        float bj = 0.f;
        if (j >= a.Tk) bj = -INFINITY;
        else if (a.mode >= 1 && a.key_pad && a.key_pad[(int64_t)b * a.Tk + j]) bj = MASK_NEG * LOG2E;
        bias_s[j] = bj;
      }
      for (int op = 0; op < 3; op++) {
        const int nch = op == 0 ? nq_ch : nk_ch;
        uint8_t* hi = tiles + (2 * op) * L.tile_bytes;
        uint8_t* lo = hi + L.tile_bytes;
        for (int ch = 0; ch < nch; ch++, n++) {
          const int st = n % a.nst;
          mbar_wait_sleep(bar(RING_FULL + st), (n / a.nst) & 1u);
#ifndef SKF_A3_DEBUG_SKIP_CONVERT    // timing experiment only (wrong results): what operand images made by the producer
                                     // of Q / K / V would save the forward kernel
          convert_stage<false>(gen + L.ring + st * STAGE_BYTES, hi, lo, 64 * ch, TP, ct);
#endif
          fence_proxy_async();             // the tensor core reads the tiles through the async proxy
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(OP_FULL + 4 * buf + op));
            mbar_arrive(bar(RING_EMPTY + st));
          }
        }
        if (ct == 0) stamp(a.trace, 3, li, 1 + op);
        if (a.img && ct == 0) {
          // every converter warp has finished (and fenced) this operand's tiles: write them out for the backward pass
          mbar_wait_sleep(bar(OP_FULL + 4 * buf + op), (li >> 1) & 1u);
          const int fop = op == 0 ? 1 : (op == 1 ? 0 : 2);          // image order K, Q, V (the backward kernels' op index)
          tma_store_2d(&mImg, smem_u32(hi), 0, ((2 * fop) * n_items + it) * TP);
          tma_store_2d(&mImg, smem_u32(lo), 0, ((2 * fop + 1) * n_items + it) * TP);
          bulk_commit();
        }
      }
    }
    if (a.img && ct == 0) bulk_wait0();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc64 = idesc_bf16(128, 64, 1);
      uint32_t li = 0, uses = 0;
      auto issue_scores = [&](uint32_t item_li, int t) {
        const uint32_t tiles = base + L.tiles + (item_li & 1) * buf_bytes;
        mma_scores(tS[t], tiles, tiles + L.tile_bytes, tiles + 2 * L.tile_bytes, tiles + 3 * L.tile_bytes, 128 * t, 0,
                   idesc_bf16(128, ncols(t, klen_s[item_li & 1]), 0));
        tc_commit(bar(S_FULL + t));
      };
      const int first = blockIdx.x;
      if (first < n_items) {
        mbar_wait_sleep(bar(OP_FULL + 0), 0);
        mbar_wait_sleep(bar(OP_FULL + 1), 0);
        tc_fence_after();
        for (int t = 0; t < ntiles; t++) issue_scores(0, t);
      }
      for (int it = first; it < n_items; it += gridDim.x, li++) {
        const int buf = li & 1;
        const uint32_t par = (li >> 1) & 1u;
        const uint32_t tiles = base + L.tiles + buf * buf_bytes;
        const bool more = it + (int)gridDim.x < n_items;
        mbar_wait_sleep(bar(OP_FULL + 4 * buf + 2), par);                     // V of this item
        stamp(a.trace, 0, li, 0);
        const int klen = klen_s[buf];
        for (int t = 0; t < ntiles; t++, uses++) {
          const int k1 = split(ncols(t, klen)) >> 4, k2 = ncols(t, klen) >> 4;
          mbar_wait_sleep(bar(P_FULL + 2 * t), li & 1u);
          stamp(a.trace, 0, li, 1 + 6 * t);
          if (uses > 0) mbar_wait_sleep(bar(O_FREE), (uses - 1) & 1u);        // the previous tile's epilogue has drained O
          stamp(a.trace, 0, li, 2 + 6 * t);
          tc_fence_after();
          mma_av_cat(tO, tS[t], tiles + 4 * L.tile_bytes, L.tile_bytes, 0, k1, idesc64, 0);
          stamp(a.trace, 0, li, 3 + 6 * t);
          mbar_wait_sleep(bar(P_FULL + 2 * t + 1), li & 1u);
          stamp(a.trace, 0, li, 4 + 6 * t);
          tc_fence_after();
          if (k2 > k1) mma_av_cat(tO, tS[t] + 16 * k1, tiles + 4 * L.tile_bytes, L.tile_bytes, 16 * k1, k2 - k1, idesc64, 1);
          tc_commit(bar(O_FULL + t));
          stamp(a.trace, 0, li, 5 + 6 * t);
          if (more) {
            // the S tile is free as soon as the product above has consumed P (in-order execution): next item's scores
            if (t == 0) {
              const int nb = (li + 1) & 1;
              const uint32_t npar = ((li + 1) >> 1) & 1u;
              mbar_wait_sleep(bar(OP_FULL + 4 * nb + 0), npar);
              mbar_wait_sleep(bar(OP_FULL + 4 * nb + 1), npar);
              tc_fence_after();
            }
            issue_scores(li + 1, t);
          }
          stamp(a.trace, 0, li, 6 + 6 * t);
        }
        tc_commit(bar(OP_EMPTY + buf));        // every MMA that reads this item's operand buffer has been issued
      }
    }
  } else if (warp >= MATH0) {
    // ================= softmax + epilogue: warpgroup t owns query tile t =================
    const int t = (warp - MATH0) >> 2;
    const int row = tid - 32 * MATH0 - 128 * t;     // TMEM lane
    const int qi = 128 * t + row;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float c1 = a.scale * LOG2E;
    const int qw = 128 * t + (warp & 3) * 32;       // first query row of this warp
    const bool idle = qw >= a.Tq;                   // no valid row in this warp: only the barrier protocol
    if (t < ntiles) {
      uint32_t li = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, li++) {
        const int b = it / a.H, h = it % a.H, buf = li & 1;
        const float* bias_s = reinterpret_cast<const float*>(gen + L.vec) + buf * 256;
        mbar_wait_sleep(bar(OP_FULL + 4 * buf + 1), (li >> 1) & 1u);     // K converted => key bias and length are written
        const int nc = ncols(t, klen_s[buf]);
        mbar_wait_sleep(bar(S_FULL + t), li & 1u);
        tc_fence_after();
        if (row == 0) stamp(a.trace, 1 + t, li, 0);
        float m = -INFINITY, l = 0.f;
        if (!idle) {
          // TMEM -> register reads run at 64 B/clk per SM (B300_MICROARCH.md "LDTM throughput"; confirmed by the phase
          // trace profiles/r2_a8_trace.log: 32-column loads with two in flight took exactly as long as 16-column ones),
          // i.e. 1.7 k clk per 128 x 208 fp32 tile and pass whatever the access pattern: the passes below are bound by
          // that, not by latency or issue slots.  16-column chunks, two register buffers alternating.
          const uint32_t tRow = tS[t] + lane_off;
          uint32_t ra[16], rb[16];
          // pass 1: row maximum of the masked, scaled logits (log2 domain)
          auto max16 = [&](const uint32_t (&r)[16], int c0) {
            if (CAUSAL && c0 + 15 > qw) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                float v = fmaf(__uint_as_float(r[j]), c1, bias_s[c0 + j]);
                if (c0 + j > qi) v += MASK_NEG * LOG2E;
                m = fmaxf(m, v);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j++) m = fmaxf(m, fmaf(__uint_as_float(r[j]), c1, bias_s[c0 + j]));
            }
          };
          tmem_ld16(tRow, ra);
          for (int c0 = 0; c0 < nc; c0 += 32) {
            tmem_ld_wait();
            if (c0 + 16 < nc) tmem_ld16(tRow + c0 + 16, rb);
            max16(ra, c0);
            if (c0 + 16 < nc) {
              tmem_ld_wait();
              if (c0 + 32 < nc) tmem_ld16(tRow + c0 + 32, ra);
              max16(rb, c0 + 16);
            }
          }
          if (row == 0) stamp(a.trace, 1 + t, li, 1);
          // pass 2: p = 2^(t - m), row sum, P -> packed bf16 hi / lo over the S columns of the same 16 keys
          auto soft16 = [&](const uint32_t (&r)[16], int c0) {
            float p[16];
            uint32_t w[16];
            if (CAUSAL && c0 + 15 > qw) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                float v = fmaf(__uint_as_float(r[j]), c1, bias_s[c0 + j]);
                if (c0 + j > qi) v += MASK_NEG * LOG2E;
                p[j] = ex2(v - m);
                l += p[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                p[j] = ex2(fmaf(__uint_as_float(r[j]), c1, bias_s[c0 + j]) - m);
                l += p[j];
              }
            }
            split16(p, w);
            tmem_st16(tRow + c0, w);
          };
          const int sp = split(nc);
          tmem_ld16(tRow, ra);
          for (int c0 = 0; c0 < nc; c0 += 32) {
            if (c0 == sp) {                 // part 0 is complete: hand it to the tensor core
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(P_FULL + 2 * t));
            }
            tmem_ld_wait();
            if (c0 + 16 < nc) tmem_ld16(tRow + c0 + 16, rb);
            soft16(ra, c0);
            if (c0 + 16 < nc) {
              tmem_ld_wait();
              if (c0 + 32 < nc) tmem_ld16(tRow + c0 + 32, ra);
              soft16(rb, c0 + 16);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (sp >= nc) mbar_arrive(bar(P_FULL + 2 * t));
            mbar_arrive(bar(P_FULL + 2 * t + 1));
          }
        } else {
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(P_FULL + 2 * t));
            mbar_arrive(bar(P_FULL + 2 * t + 1));
          }
        }
        // epilogue: O = (P Vh + P Vl) / l
        if (row == 0) stamp(a.trace, 1 + t, li, 2);
        mbar_wait_sleep(bar(O_FULL + t), li & 1u);
        tc_fence_after();
        if (row == 0) stamp(a.trace, 1 + t, li, 3);
        uint32_t o0[32], o1[32];
        if (!idle) {
          tmem_ld32(tO + lane_off, o0);
          tmem_ld32(tO + lane_off + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(O_FREE));
        if (!idle && qi < a.Tq) {
          const float inv = 1.0f / l;
          float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + qi) * a.ld0 + h * DH);
#pragma unroll
          for (int j = 0; j < 8; j++)
            dst[j] = make_float4((__uint_as_float(o0[4 * j]) + __uint_as_float(o1[4 * j])) * inv,
                                 (__uint_as_float(o0[4 * j + 1]) + __uint_as_float(o1[4 * j + 1])) * inv,
                                 (__uint_as_float(o0[4 * j + 2]) + __uint_as_float(o1[4 * j + 2])) * inv,
                                 (__uint_as_float(o0[4 * j + 3]) + __uint_as_float(o1[4 * j + 3])) * inv);
          if (a.out1) a.out1[((int64_t)b * a.H + h) * a.Tq + qi] = m * 0.6931471805599453f + logf(l);
        }
        if (row == 0) stamp(a.trace, 1 + t, li, 4);
      }
    }
  }
  epilogue_free(c);
}

// ---------------------------------------------------------------------------------------------------
// backward.  One template for both kernels; "lanes" are the rows of the output (queries for dQ, keys for dK / dV),
// "columns" the other sequence, split in two halves of <= 112 between the warpgroups.
//
//   DKV = false (dq) : L = Q, Lg = dO (lane side), C = K, Cg = V (column side).
//       S = L C^T, dP = Lg Cg^T, dS = P (dP - delta_lane), acc[64] += dS [Ch | Cl]           -> dQ = scale * acc
//       operand tiles: C (= K) double-buffered (its last reader is the late dQ product), L, Lg, Cg single-buffered
//       and released as soon as the item's score MMAs are done.  Also computes delta = rowsum(dO * O) while
//       converting dO (O rides the ring as a fifth operand) and writes it to HBM for dkv.
//   DKV = true (dkv) : L = K, Lg = V (lane side), C = Q, Cg = dO (column side).
//       S^T = L C^T, dP^T = Lg Cg^T, P^T, dS^T = P^T (dP^T - delta_col), accV[32] += P^T Cg, accK[32] += dS^T C
//       operand tiles: C, Cg (= Q, dO) double-buffered, L, Lg single-buffered.
//
// tiles: slot 0,1 = A hi/lo buf 0, 2,3 = A hi/lo buf 1 (A = the double-buffered pair: K | Q), 4,5 = (dkv only: dO buf 0),
// ... laid out by tile index below.   tmem: S_w [224 w, +112)  dP_w [224 w + 112, +112)  acc [448, 512)
// ---------------------------------------------------------------------------------------------------
namespace bw {
enum { RING_FULL = 0, RING_EMPTY = NST, OP_FULL = 2 * NST /* [2 bufs][4 ops] */, DB_EMPTY = 2 * NST + 8 /* [2] */,
       SB_EMPTY = 2 * NST + 10, SP_FULL = 2 * NST + 11 /* [2] */, DS_FULL = 2 * NST + 13 /* [2] */,
       ACC_FULL = 2 * NST + 15 /* [2]: one per epilogue warpgroup */, ACC_FREE = 2 * NST + 17,
       VEC_FULL = 2 * NST + 18 /* [2] */, COUNT = 2 * NST + 20 };
}

template <bool DKV, bool CAUSAL>
__global__ void __launch_bounds__(THREADS, 1)
attn3_bwd_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
            const __grid_constant__ CUtensorMap mV, const __grid_constant__ CUtensorMap mdO,
            const __grid_constant__ CUtensorMap mO, const __grid_constant__ CUtensorMap mImg,
            const __grid_constant__ CUtensorMap mImgDo, const Args a) {
  using namespace bw;
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  prologue(c, smem_raw);
  const int QP = (a.Tq + 15) & ~15, KP = (a.Tk + 15) & ~15, TP = QP > KP ? QP : KP;
  constexpr int NTILES = 12;
  const Smem L = smem_plan(TP, NTILES, DKV ? 6 : 4, a.nst);
  const int tid = c.tid, warp = c.warp, lane = c.lane;
  const uint32_t base = c.base;
  uint8_t* gen = c.gen;
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int nq_ch = (a.Tq + 63) >> 6, nk_ch = (a.Tk + 63) >> 6;
  const int n_items = a.B * a.H;
  // lanes / columns
  const int TL = DKV ? a.Tk : a.Tq, TC = DKV ? a.Tq : a.Tk;
  const int ntiles = TL > 128 ? 2 : 1;
  const int nhalf = TC > HALF ? 2 : 1;
  // operand slots.  op index: 0 = K, 1 = Q, 2 = V, 3 = dO.  Double-buffered ops get two slot pairs.
  //   dq : K -> slots {0, 2} (+buf*2)   Q -> 4   V -> 6   dO -> {8, 10}
  //   dkv: Q -> slots {0, 2}            dO -> {4, 6}      K -> 8   V -> 10
  auto slot = [&](int op, int buf) -> int {
    if (!DKV) return op == 0 ? 2 * buf : (op == 1 ? 4 : (op == 2 ? 6 : 8 + 2 * buf));
    return op == 1 ? 2 * buf : (op == 3 ? 4 + 2 * buf : (op == 0 ? 8 : 10));
  };
  auto is_db = [&](int op) { return DKV ? (op == 1 || op == 3) : (op == 0 || op == 3); };
  auto nch_of = [&](int op) { return (op == 0 || op == 2) ? nk_ch : nq_ch; };
  // vectors: [0,256) / [256,512): per-item bias of the keys (dq: column side, dkv: lane side), double-buffered;
  // [512,768) / [768,1024): dq: delta of the item's queries; dkv: interleaved (lse2, delta) pairs need 2 x 256 x 2
  float* vec = reinterpret_cast<float*>(gen + L.vec);

  if (tid == 0) {
    for (int i = 0; i < NST; i++) {
      mbar_init(bar(RING_FULL + i), 1);
      mbar_init(bar(RING_EMPTY + i), NCONV);
    }
    // operands that arrive as images by one TMA transaction (direct mode: all but dO in dq) count a single arrival
    for (int b = 0; b < 2; b++) {
      for (int op = 0; op < 4; op++)
        mbar_init(bar(OP_FULL + 4 * b + op), (a.direct && (DKV || op != 3)) ? 1 : NCONV * nch_of(op));
      mbar_init(bar(VEC_FULL + b), NCONV);
    }
    mbar_init(bar(DB_EMPTY + 0), 1);
    mbar_init(bar(DB_EMPTY + 1), 1);
    mbar_init(bar(SB_EMPTY), 1);
    for (int w = 0; w < 2; w++) {
      mbar_init(bar(SP_FULL + w), 1);
      mbar_init(bar(DS_FULL + w), 4);
    }
    // dq: the warpgroups take turns at the epilogue, so each gets its OWN "accumulator ready" barrier and sees every
    // phase of it.  (With one barrier a warpgroup waited on every other phase; a warpgroup that ran ahead — short
    // sequences make a unit's math shorter than the MMA issue — found the parity of the phase before last already
    // complete and read the accumulators of a unit that had not finished.)
    mbar_init(bar(ACC_FULL), 1);
    mbar_init(bar(ACC_FULL + 1), 1);
    mbar_init(bar(ACC_FREE), 4);
    fence_barrier_init();
  }
  for (int i = tid; i < 1024; i += THREADS) reinterpret_cast<uint32_t*>(gen + L.vec - 4096)[i] = 0u;
  alloc_tmem(c, L.bars + 8u * COUNT, a.pdl_late);
  const uint32_t tmem = c.tmem;
  const uint32_t tAcc = tmem + 2 * MAXT;
  // columns unit (lane tile t, half w) covers: [HALF w, HALF w + ncols)
  volatile int* klen_s = reinterpret_cast<volatile int*>(gen + L.meta);      // [2 buffers]
  auto ncols = [&](int t, int w, int klen) {
    int hi = DKV ? TC : min(TC, klen);                   // dq: key columns past the effective length carry p = 0
    if (CAUSAL && !DKV) hi = min(hi, 128 * t + 128);     // queries of tile t see keys <= their own index
    int n = hi - HALF * w;
    n = n < 0 ? 0 : (n > HALF ? HALF : n);
    n = (n + 15) & ~15;
    return n < 16 ? 16 : n;
  };

  // dkv: key tile 1 (keys 128..) of an item whose effective length is <= 128 has zero gradients: not computed
  auto tiles_of = [&](int klen) { return (DKV && klen <= 128) ? 1 : ntiles; };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      prefetch_tmap(&mQ); prefetch_tmap(&mK); prefetch_tmap(&mV); prefetch_tmap(&mdO);
      uint32_t n = 0;
      auto load = [&](const CUtensorMap* m, int h, int row, int b) {
        const int st = n % a.nst;
        mbar_wait_sleep(bar(RING_EMPTY + st), ((n / a.nst) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(RING_FULL + st), STAGE_BYTES);
        tma_load_3d(base + L.ring + st * STAGE_BYTES, m, bar(RING_FULL + st), h * DH, row, b);
        n++;
      };
      uint32_t pli = 0;
      // direct mode: one [TP x 32] hi tile and one lo tile per operand, straight into the operand slot
      auto load_img = [&](const CUtensorMap* m, int op, int buf, int it, int tile_op) {
        const uint32_t dst = base + L.tiles + slot(op, buf) * L.tile_bytes;
        mbar_arrive_expect_tx(bar(OP_FULL + 4 * buf + op), 2 * L.tile_bytes);
        tma_load_2d(dst, m, bar(OP_FULL + 4 * buf + op), 0, ((2 * tile_op) * n_items + it) * TP);
        tma_load_2d(dst + L.tile_bytes, m, bar(OP_FULL + 4 * buf + op), 0, ((2 * tile_op + 1) * n_items + it) * TP);
      };
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, pli++) {
        const int b = it / a.H, h = it % a.H;
        if (a.direct) {
          const int buf = pli & 1;
          if (DKV) mbar_wait_sleep(bar(DB_EMPTY + buf), ((pli >> 1) & 1u) ^ 1u);
          if (!DKV) {
            // dO through the ring (fp32 -> tiles) first: the ring does not depend on any operand buffer, so these
            // rows are in flight a whole item ahead
            for (int ch = 0; ch < nq_ch; ch++) load(&mdO, h, 64 * ch, b);
            mbar_wait_sleep(bar(DB_EMPTY + buf), ((pli >> 1) & 1u) ^ 1u);
            load_img(&mImg, 0, buf, it, 0);                                  // K (double-buffered)
            mbar_wait_sleep(bar(SB_EMPTY), (pli & 1u) ^ 1u);
            load_img(&mImg, 1, buf, it, 1);                                  // Q
            load_img(&mImg, 2, buf, it, 2);                                  // V
          } else {
            load_img(&mImg, 1, buf, it, 1);                                  // Q  (double-buffered)
            load_img(&mImgDo, 3, buf, it, 0);                                // dO (double-buffered)
            mbar_wait_sleep(bar(SB_EMPTY), (pli & 1u) ^ 1u);
            load_img(&mImg, 0, buf, it, 0);                                  // K
            load_img(&mImg, 2, buf, it, 2);                                  // V
          }
          continue;
        }
        // order = order of first use: dq K, Q, V, (O, dO) pairs; dkv Q, dO, K, V
        if (!DKV) {
          for (int ch = 0; ch < nq_ch; ch++) load(&mdO, h, 64 * ch, b);
          for (int ch = 0; ch < nk_ch; ch++) load(&mK, h, 64 * ch, b);
          for (int ch = 0; ch < nq_ch; ch++) load(&mQ, h, 64 * ch, b);
          for (int ch = 0; ch < nk_ch; ch++) load(&mV, h, 64 * ch, b);
        } else {
          for (int ch = 0; ch < nq_ch; ch++) load(&mQ, h, 64 * ch, b);
          for (int ch = 0; ch < nq_ch; ch++) load(&mdO, h, 64 * ch, b);
          for (int ch = 0; ch < nk_ch; ch++) load(&mK, h, 64 * ch, b);
          for (int ch = 0; ch < nk_ch; ch++) load(&mV, h, 64 * ch, b);
        }
      }
      if (a.pdl_late) pdl_trigger();
    }
  } else if (warp >= 2 && warp < 2 + NCONV) {
    // ================= converters =================
    const int ct = tid - 64;
    uint32_t n = 0, li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, li++) {
      const int b = it / a.H, h = it % a.H, buf = li & 1;
      const int64_t bh = (int64_t)b * a.H + h;
      // double-buffered operands + vectors of buffer `buf`: free once the item two back has issued its last MMA
      // dq: this thread's share of O (for delta) does not depend on any buffer: in flight during the waits below
      float4 oval[4][CONV_ROUNDS][2];
      if (!DKV) {
        const float* Og = a.O + (int64_t)b * a.Tq * a.ldO + h * DH;
#pragma unroll
        for (int ch = 0; ch < 4; ch++)
          if (ch < nq_ch) prefetch_rows(oval[ch], Og, a.ldO, 64 * ch, a.Tq, ct);
      }
      mbar_wait_sleep(bar(DB_EMPTY + buf), ((li >> 1) & 1u) ^ 1u);
      if (ct == 0) stamp(a.trace, 3, li, 0);
      {
        const int kl = effective_klen(a.key_pad, a.mode, (int64_t)b * a.Tk, a.Tk, lane);
        if (ct == 0) klen_s[buf] = kl;
      }
      float* bias_s = vec + buf * 256;
      for (int j = ct; j < KP; j += NCONV * 32) {
        float bj = 0.f;
        if (j >= a.Tk) bj = -INFINITY;
        else if (a.mode >= 1 && a.key_pad && a.key_pad[(int64_t)b * a.Tk + j]) bj = MASK_NEG * LOG2E;
        bias_s[j] = bj;
      }
      if (DKV) {
        // per-query lse (log2 domain, +inf past the sequence: p = 0) and delta, column vectors of this item
        float* lse_s = vec + 512 + buf * 512;
        float* del_s = lse_s + 256;
        for (int j = ct; j < QP; j += NCONV * 32) {
          lse_s[j] = j < a.Tq ? a.lse[bh * a.Tq + j] * LOG2E : INFINITY;
          del_s[j] = j < a.Tq ? a.delta[bh * a.Tq + j] : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(VEC_FULL + buf));
      auto conv = [&](int op, bool first_sb) {
        const int nch = nch_of(op);
        uint8_t* hi = gen + L.tiles + slot(op, buf) * L.tile_bytes;
        uint8_t* lo = hi + L.tile_bytes;
        for (int ch = 0; ch < nch; ch++, n++) {
          // single-buffered operands: the previous item's score MMAs (their only readers) have completed
          if (first_sb && ch == 0) {
            if (ct == 0) stamp(a.trace, 3, li, 5);
            mbar_wait_sleep(bar(SB_EMPTY), (li & 1u) ^ 1u);
            if (ct == 0) stamp(a.trace, 3, li, 6);
          }
          const int st = n % a.nst;
          mbar_wait_sleep(bar(RING_FULL + st), (n / a.nst) & 1u);
          convert_stage<false>(gen + L.ring + st * STAGE_BYTES, hi, lo, 64 * ch, TP, ct);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(OP_FULL + 4 * buf + op));
            mbar_arrive(bar(RING_EMPTY + st));
          }
        }
        if (ct == 0) stamp(a.trace, 3, li, 1 + op);
      };
      if (DKV && a.direct) continue;          // every operand arrives as an image: only the vectors above
      if (!DKV) {
        if (a.img_do) {            // the image stores of the earlier items have read their dO tiles
          if (ct == 0) bulk_wait_read0();
          conv_bar_sync();
        }
        // dO chunks first (double-buffered tile): delta = rowsum(dO * O) from the fp32 values, dO -> tiles
        float* del_s = vec + 512 + buf * 256;
        uint8_t* hi = gen + L.tiles + slot(3, buf) * L.tile_bytes;
        uint8_t* lo = hi + L.tile_bytes;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          if (ch >= nq_ch) break;
          const int st = n % a.nst;
          mbar_wait_sleep(bar(RING_FULL + st), (n / a.nst) & 1u);
          convert_stage<true>(gen + L.ring + st * STAGE_BYTES, hi, lo, 64 * ch, TP, ct, oval[ch], del_s,
                              a.out1 ? a.out1 + bh * a.Tq : nullptr, a.Tq);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(OP_FULL + 4 * buf + 3));
            mbar_arrive(bar(RING_EMPTY + st));
          }
          n++;
        }
        if (ct == 0) stamp(a.trace, 3, li, 4);
        if (a.img_do && ct == 0) {             // dO images for dkv
          mbar_wait_sleep(bar(OP_FULL + 4 * buf + 3), (li >> 1) & 1u);
          tma_store_2d(&mImgDo, smem_u32(hi), 0, it * TP);
          tma_store_2d(&mImgDo, smem_u32(lo), 0, (n_items + it) * TP);
          bulk_commit();
        }
        if (!a.direct) {
          conv(0, false);          // K (double-buffered)
          conv(1, true);           // Q
          conv(2, false);          // V
        }
      } else {
        conv(1, false);          // Q  (double-buffered)
        conv(3, false);          // dO (double-buffered)
        conv(0, true);           // K
        conv(2, false);          // V
      }
    }
    if (!DKV && a.img_do && ct == 0) bulk_wait0();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc64 = idesc_bf16(128, 64, 1), idesc32 = idesc_bf16(128, 32, 1);
      auto tile_addr = [&](int op, int buf) { return base + L.tiles + slot(op, buf) * L.tile_bytes; };
      // score products of unit (item li, lane tile t, half w): S_w = L C^T, dP_w = Lg Cg^T
      auto issue_scores = [&](uint32_t item_li, int t, int w) {
        const int buf = item_li & 1;
        const uint32_t idesc = idesc_bf16(128, ncols(t, w, klen_s[buf]), 0);
        const int opL = DKV ? 0 : 1, opLg = DKV ? 2 : 3, opC = DKV ? 1 : 0, opCg = DKV ? 3 : 2;
        const uint32_t Lh = tile_addr(opL, buf), Lgh = tile_addr(opLg, buf), Ch = tile_addr(opC, buf), Cgh = tile_addr(opCg, buf);
        mma_scores(tmem + MAXT * w, Lh, Lh + L.tile_bytes, Ch, Ch + L.tile_bytes, 128 * t, HALF * w, idesc);
        mma_scores(tmem + MAXT * w + HALF, Lgh, Lgh + L.tile_bytes, Cgh, Cgh + L.tile_bytes, 128 * t, HALF * w, idesc);
        tc_commit(bar(SP_FULL + w));
      };
      auto wait_item = [&](uint32_t item_li) {
        const int buf = item_li & 1;
        const uint32_t par = (item_li >> 1) & 1u;
        for (int op = 0; op < 4; op++) mbar_wait_sleep(bar(OP_FULL + 4 * buf + op), par);
        mbar_wait_sleep(bar(VEC_FULL + buf), par);       // effective key length
        tc_fence_after();
      };
      const int first = blockIdx.x;
      uint32_t li = 0, u = 0;           // u: unit pairs processed so far by this CTA
      if (first < n_items) {
        wait_item(0);
        for (int w = 0; w < nhalf; w++) issue_scores(0, 0, w);
        if (tiles_of(klen_s[0]) == 1) tc_commit(bar(SB_EMPTY));
      }
      for (int it = first; it < n_items; it += gridDim.x, li++) {
        const int buf = li & 1;
        const bool more = it + (int)gridDim.x < n_items;
        const int klen = klen_s[buf];
        const int nt_item = tiles_of(klen);
        for (int t = 0; t < nt_item; t++, u++) {
          const bool last_tile = t + 1 == nt_item;
          const bool has_next = !last_tile || more;
          const uint32_t nli = last_tile ? li + 1 : li;
          const int nt = last_tile ? 0 : t + 1;
          for (int w = 0; w < nhalf; w++) {
            mbar_wait_sleep(bar(DS_FULL + w), u & 1u);
            stamp(a.trace, 0, li, 6 * t + 3 * w);
            if (w == 0 && u > 0) mbar_wait_sleep(bar(ACC_FREE), (u - 1) & 1u);      // previous epilogue drained the accumulators
            stamp(a.trace, 0, li, 6 * t + 3 * w + 1);
            tc_fence_after();
            const int ks = ncols(t, w, klen) >> 4;
            if (!DKV) {
              const uint32_t Kh = tile_addr(0, buf);
              mma_av_cat(tAcc, tmem + MAXT * w + HALF, Kh, L.tile_bytes, HALF * w, ks, idesc64, w > 0);
            } else {
              const uint32_t dOh = tile_addr(3, buf), Qh = tile_addr(1, buf);
              mma_av_32(tAcc, tmem + MAXT * w, dOh, dOh + L.tile_bytes, HALF * w, ks, idesc32, w > 0);            // dV += P^T dO
              mma_av_32(tAcc + 32, tmem + MAXT * w + HALF, Qh, Qh + L.tile_bytes, HALF * w, ks, idesc32, w > 0);  // dK += dS^T Q
            }
            if (w == nhalf - 1) {
              tc_commit(bar(ACC_FULL + ((DKV || nhalf == 1) ? 0 : (t & 1))));
              if (last_tile) tc_commit(bar(DB_EMPTY + buf));     // last reader of the double-buffered operands
            }
            if (has_next) {
              if (w == 0 && last_tile) wait_item(nli);
              issue_scores(nli, nt, w);
              // the single-buffered operands are free once the item's last score MMAs have completed
              if (w == nhalf - 1 && nt == tiles_of(klen_s[nli & 1]) - 1) tc_commit(bar(SB_EMPTY));
            }
            stamp(a.trace, 0, li, 6 * t + 3 * w + 2);
          }
        }
      }
    }
  } else if (warp >= MATH0) {
    // ================= math: warpgroup w owns column half w =================
    const int w = (warp - MATH0) >> 2;
    const int row = tid - 32 * MATH0 - 128 * w;     // TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float c1 = a.scale * LOG2E;
    const uint32_t tS = tmem + MAXT * w, tP = tS + HALF;
    // dkv: accumulator columns [0,32) = dV, [32,64) = dK of key row `krow` of item (eb, eh), unit parity eu
    auto dkv_epilogue = [&](int eb, int eh, int krow, bool eidle, uint32_t eu) {
      mbar_wait_sleep(bar(ACC_FULL), eu & 1u);
      tc_fence_after();
      uint32_t ov[32], ok[32];
      if (!eidle) {
        tmem_ld32(tAcc + lane_off, ov);
        tmem_ld32(tAcc + lane_off + 32, ok);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(ACC_FREE));
      if (!eidle && krow < a.Tk) {
        float4* dk = reinterpret_cast<float4*>(a.out0 + ((int64_t)eb * a.Tk + krow) * a.ld0 + eh * DH);
        float4* dv = reinterpret_cast<float4*>(a.out1 + ((int64_t)eb * a.Tk + krow) * a.ld1 + eh * DH);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          dv[j] = make_float4(__uint_as_float(ov[4 * j]), __uint_as_float(ov[4 * j + 1]), __uint_as_float(ov[4 * j + 2]),
                              __uint_as_float(ov[4 * j + 3]));
          dk[j] = make_float4(__uint_as_float(ok[4 * j]) * a.scale, __uint_as_float(ok[4 * j + 1]) * a.scale,
                              __uint_as_float(ok[4 * j + 2]) * a.scale, __uint_as_float(ok[4 * j + 3]) * a.scale);
        }
      }
    };
    bool have_prev = false, pidle = false;
    int pb = 0, ph = 0, prow = 0;
    uint32_t pu = 0;
    if (w < nhalf) {
      uint32_t li = 0, u = 0, nepi = 0;           // nepi: dq epilogues this warpgroup has done
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, li++) {
        const int b = it / a.H, h = it % a.H, buf = li & 1;
        const int64_t bh = (int64_t)b * a.H + h;
        const float* bias_s = vec + buf * 256;
        // vectors of this item are written before the converters arrive on these barriers
        mbar_wait_sleep(bar(VEC_FULL + buf), (li >> 1) & 1u);
        if (!DKV) mbar_wait_sleep(bar(OP_FULL + 4 * buf + 3), (li >> 1) & 1u);      // delta (with dO)
        const int klen = klen_s[buf];
        const int nt_item = tiles_of(klen);
        if (DKV && nt_item < ntiles) {
          // skipped key tile: its dK / dV rows are exactly zero (warpgroup 0 writes dK, warpgroup 1 dV; with one
          // column half, warpgroup 0 writes both)
          const int kr = 128 + row;
          if (kr < a.Tk) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w == 0) {
              float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tk + kr) * a.ld0 + h * DH);
#pragma unroll
              for (int j = 0; j < 8; j++) dst[j] = z;
            }
            if (w == 1 || nhalf == 1) {
              float4* dst = reinterpret_cast<float4*>(a.out1 + ((int64_t)b * a.Tk + kr) * a.ld1 + h * DH);
#pragma unroll
              for (int j = 0; j < 8; j++) dst[j] = z;
            }
          }
        }
        for (int t = 0; t < nt_item; t++, u++) {
          const int li_row = 128 * t + row;           // sequence index of this lane (query for dq, key for dkv)
          const int lw = 128 * t + (warp & 3) * 32;   // first lane index of this warp
          const int nc = ncols(t, w, klen);
          const int col0 = HALF * w;
          const bool idle = lw >= TL;
          float lane_bias = 0.f, lane_delta = 0.f, lane_lse2 = INFINITY;
          if (!DKV) {
            if (li_row < a.Tq) {
              lane_lse2 = a.lse[bh * a.Tq + li_row] * LOG2E;
              lane_delta = (vec + 512 + buf * 256)[li_row];
            }
          } else {
            lane_bias = li_row < KP ? bias_s[li_row] : -INFINITY;
          }
          const float* lse_s = vec + 512 + buf * 512;        // dkv only
          const float* del_s = lse_s + 256;
          mbar_wait_sleep(bar(SP_FULL + w), u & 1u);
          tc_fence_after();
          if (row == 0) stamp(a.trace, 1 + w, li, 4 * t);
          if (!idle) {
            uint32_t rs[16], rp[16];
            for (int c0 = 0; c0 < nc; c0 += 16) {
              tmem_ld16(tS + lane_off + c0, rs);
              tmem_ld16(tP + lane_off + c0, rp);
              tmem_ld_wait();
              const int cj0 = col0 + c0;              // sequence index of the chunk's first column
              float p[16], ds[16];
              uint32_t wv[16];
              if (!DKV) {
                const bool diag = CAUSAL && cj0 + 15 > lw;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  float v = fmaf(__uint_as_float(rs[j]), c1, bias_s[cj0 + j] - lane_lse2);
                  if (diag && cj0 + j > li_row) v += MASK_NEG * LOG2E;
                  p[j] = ex2(v);
                  ds[j] = p[j] * (__uint_as_float(rp[j]) - lane_delta);
                }
                split16(ds, wv);
                tmem_st16(tP + lane_off + c0, wv);
              } else {
                const bool diag = CAUSAL && lw + 31 > cj0;        // some (key lane, query column) pair has key > query
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  float v = fmaf(__uint_as_float(rs[j]), c1, lane_bias - lse_s[cj0 + j]);
                  if (diag && li_row > cj0 + j) v += MASK_NEG * LOG2E;
                  p[j] = ex2(v);
                  ds[j] = p[j] * (__uint_as_float(rp[j]) - del_s[cj0 + j]);
                }
                split16(p, wv);
                tmem_st16(tS + lane_off + c0, wv);
                split16(ds, wv);
                tmem_st16(tP + lane_off + c0, wv);
              }
            }
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(DS_FULL + w));
          if (row == 0) stamp(a.trace, 1 + w, li, 4 * t + 1);
          // epilogue.  dq: warpgroup (t & 1) stores dQ of tile t, right away.  dkv: warpgroup 0 stores dV and dK of a
          // key tile AFTER its math of the next unit: the accumulators are final only when the other warpgroup's half
          // has been consumed, and waiting for that here would keep this warpgroup from the next unit's scores, which
          // the tensor core finished long before (phase trace profiles/r2_a13_trace_dkv.log).
          if (!DKV) {
            if (nhalf == 1 || (t & 1) == w) {
              mbar_wait_sleep(bar(ACC_FULL + w), nepi & 1u);
              nepi++;
              tc_fence_after();
              if (row == 0) stamp(a.trace, 1 + w, li, 4 * t + 2);
              uint32_t o0[32], o1[32];
              if (!idle) {
                tmem_ld32(tAcc + lane_off, o0);
                tmem_ld32(tAcc + lane_off + 32, o1);
                tmem_ld_wait();
              }
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(ACC_FREE));
              if (!idle && li_row < a.Tq) {
                float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + li_row) * a.ld0 + h * DH);
#pragma unroll
                for (int j = 0; j < 8; j++)
                  dst[j] = make_float4((__uint_as_float(o0[4 * j]) + __uint_as_float(o1[4 * j])) * a.scale,
                                       (__uint_as_float(o0[4 * j + 1]) + __uint_as_float(o1[4 * j + 1])) * a.scale,
                                       (__uint_as_float(o0[4 * j + 2]) + __uint_as_float(o1[4 * j + 2])) * a.scale,
                                       (__uint_as_float(o0[4 * j + 3]) + __uint_as_float(o1[4 * j + 3])) * a.scale);
              }
            }
          } else if (w == 0) {
            if (have_prev) dkv_epilogue(pb, ph, prow, pidle, pu);
            have_prev = true; pb = b; ph = h; prow = li_row; pidle = idle; pu = u;
          }
        }
      }
      if (DKV && w == 0 && have_prev) dkv_epilogue(pb, ph, prow, pidle, pu);
    }
  }
  epilogue_free(c);
}

}  // namespace a3

// ---- host --------------------------------------------------------------------------------------
int attn3_supported(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, int Tq,
                    int Tk, int dh) {
  if (dh != a3::DH) return 0;
  if (Tq < 1 || Tk < 1 || Tq > a3::MAXT || Tk > a3::MAXT) return 0;
  if ((ldq | ldk | ldv) & 3) return 0;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V)) & 15) return 0;
  return 1;
}

int g_a3_trace = 0;
int attn3_trace_dump(long long* out, int n) {
  if (n > 4 * 8 * 16) n = 4 * 8 * 16;
  SKF_CUDA(cudaMemcpyFromSymbol(out, a3::g_trace, sizeof(long long) * n));
  return n;
}

static int a3_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
// as many ring stages (8 KB each, 4..NST) as fit beside the operand tiles in 227 KB of shared memory: the ring is the
// only data in flight from HBM, and at ~2 us per TMA round trip under load 32 KB per SM starve the kernels
static int a3_ring_stages(int TP, int ntiles, int nvec) {
  const int limit = 227 * 1024 - 1024;
  int nst = a3::NST;
  while (nst > 4 && (int)a3::smem_plan(TP, ntiles, nvec, nst).total > limit) nst--;
  return nst;
}
template <class Kern>
static int a3_attr(Kern k, int bytes) {
  SKF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return SKF_OK;
}
static int a3_map(CUtensorMap* m, const float* p, int64_t ld, int H, int T, int B) {
  return make_tmap_3d_f32(m, p, (uint64_t)H * a3::DH, (uint64_t)T, (uint64_t)B, (uint64_t)ld, (uint64_t)T * ld, 32, 64);
}

int make_tmap_2d_any(CUtensorMap* out, const void* ptr, int dtype, int esize, uint64_t inner, uint64_t outer,
                     uint64_t ld, uint32_t box_inner, uint32_t box_outer, int swizzle);
// bytes of the image buffer of one attention site (Q, K, V: 6 tiles per item) / of the dO images (2 tiles per item)
size_t attn3_image_bytes(int B, int H, int Tq, int Tk, int ntiles) {
  const int QP = (Tq + 15) & ~15, KP = (Tk + 15) & ~15, TP = QP > KP ? QP : KP;
  return (size_t)ntiles * B * H * TP * 64;
}
static int a3_img_map(CUtensorMap* m, const void* img, int ntiles, int n_items, int TP) {
  return make_tmap_2d_any(m, img, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 32, (uint64_t)ntiles * n_items * TP, 32, 32,
                          (uint32_t)TP, CU_TENSOR_MAP_SWIZZLE_64B);
}

int attn3_fwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
              int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh, int mode,
              void* img, cudaStream_t st) {
  SKF_REQUIRE(attn3_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn3_fwd: shape");
  SKF_REQUIRE((ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, SKF_ERR_BAD_ALIGN, "attn3_fwd: O");
  SKF_REQUIRE(mode == 0 || key_pad != nullptr || mode == 2, SKF_ERR_BAD_ARG, "attn3_fwd: mode 1 needs key_pad");
  SKF_REQUIRE(mode != 2 || Tq == Tk, SKF_ERR_BAD_ARG, "attn3_fwd: the causal mask needs Tq == Tk");
  CUtensorMap mQ, mK, mV;
  SKF_TRY(a3_map(&mQ, Q, ldq, H, Tq, B));
  SKF_TRY(a3_map(&mK, K, ldk, H, Tk, B));
  SKF_TRY(a3_map(&mV, V, ldv, H, Tk, B));
  a3::Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.scale = 1.0f / sqrtf((float)dh);
  a.trace = g_a3_trace == 1;
  a.pdl_late = (g_pdl & PDL_LATE) != 0;
  a.key_pad = key_pad; a.out0 = O; a.ld0 = ldo; a.out1 = lse;
  const int QP = (Tq + 15) & ~15, KP = (Tk + 15) & ~15, TP = QP > KP ? QP : KP;
  CUtensorMap mImg = mQ;
  a.img = img != nullptr;
  if (img) SKF_TRY(a3_img_map(&mImg, img, 6, B * H, TP));
  a.nst = a3_ring_stages(TP, 12, 2);
  const int smem = (int)a3::smem_plan(TP, 12, 2, a.nst).total + 1024;
  const int grid = B * H < a3_sm_count() ? B * H : a3_sm_count();
  if (mode == 2) {
    SKF_TRY(a3_attr(a3::attn3_fwd_k<true>, smem));
    launch_k(a3::attn3_fwd_k<true>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mImg, a);
  } else {
    SKF_TRY(a3_attr(a3::attn3_fwd_k<false>, smem));
    launch_k(a3::attn3_fwd_k<false>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mImg, a);
  }
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int attn3_bwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, const float* O,
              int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta, float* dQ, int64_t lddq,
              float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad, int B, int H, int Tq, int Tk,
              int dh, int mode, const void* img, void* img_do, cudaStream_t st) {
  SKF_REQUIRE(attn3_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn3_bwd: shape");
  SKF_REQUIRE(((lddo | lddq | lddk | lddv | ldo) & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(dQ) | reinterpret_cast<uintptr_t>(dK) |
                    reinterpret_cast<uintptr_t>(dV) | reinterpret_cast<uintptr_t>(O)) & 15) == 0,
              SKF_ERR_BAD_ALIGN, "attn3_bwd: alignment");
  SKF_REQUIRE(mode == 0 || key_pad != nullptr || mode == 2, SKF_ERR_BAD_ARG, "attn3_bwd: mode 1 needs key_pad");
  SKF_REQUIRE(mode != 2 || Tq == Tk, SKF_ERR_BAD_ARG, "attn3_bwd: the causal mask needs Tq == Tk");
  SKF_REQUIRE(lse != nullptr && delta != nullptr, SKF_ERR_BAD_ARG, "attn3_bwd: lse / delta buffers");
  CUtensorMap mQ, mK, mV, mdO, mO;
  SKF_TRY(a3_map(&mQ, Q, ldq, H, Tq, B));
  SKF_TRY(a3_map(&mK, K, ldk, H, Tk, B));
  SKF_TRY(a3_map(&mV, V, ldv, H, Tk, B));
  SKF_TRY(a3_map(&mdO, dO, lddo, H, Tq, B));
  SKF_TRY(a3_map(&mO, O, ldo, H, Tq, B));
  a3::Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.scale = 1.0f / sqrtf((float)dh);
  a.key_pad = key_pad; a.lse = lse; a.O = O; a.ldO = ldo;
  a.pdl_late = (g_pdl & PDL_LATE) != 0;
  const int QP = (Tq + 15) & ~15, KP = (Tk + 15) & ~15, TP = QP > KP ? QP : KP;
  const int grid = B * H < a3_sm_count() ? B * H : a3_sm_count();
  // with the forward kernel's operand images (and room for dO's) the backward kernels load bf16 tiles directly
  CUtensorMap mImg = mQ, mImgDo = mQ;
  a.direct = img != nullptr && img_do != nullptr;
  if (a.direct) {
    SKF_TRY(a3_img_map(&mImg, img, 6, B * H, TP));
    SKF_TRY(a3_img_map(&mImgDo, img_do, 2, B * H, TP));
  }
  {
    a.img_do = a.direct;
    a.out0 = dQ; a.ld0 = lddq; a.out1 = delta; a.delta = nullptr; a.trace = g_a3_trace == 2;
    a.nst = a3_ring_stages(TP, 12, 4);
    const int smem = (int)a3::smem_plan(TP, 12, 4, a.nst).total + 1024;
    if (mode == 2) {
      SKF_TRY(a3_attr(a3::attn3_bwd_k<false, true>, smem));
      launch_k(a3::attn3_bwd_k<false, true>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mdO, mO, mImg, mImgDo, a);
    } else {
      SKF_TRY(a3_attr(a3::attn3_bwd_k<false, false>, smem));
      launch_k(a3::attn3_bwd_k<false, false>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mdO, mO, mImg, mImgDo, a);
    }
    SKF_LAUNCH_CHECK();
  }
  {
    a.img_do = 0;
    a.out0 = dK; a.ld0 = lddk; a.out1 = dV; a.ld1 = lddv; a.delta = delta; a.trace = g_a3_trace == 3;
    a.nst = a3_ring_stages(TP, 12, 6);
    const int smem = (int)a3::smem_plan(TP, 12, 6, a.nst).total + 1024;
    if (mode == 2) {
      SKF_TRY(a3_attr(a3::attn3_bwd_k<true, true>, smem));
      launch_k(a3::attn3_bwd_k<true, true>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mdO, mO, mImg, mImgDo, a);
    } else {
      SKF_TRY(a3_attr(a3::attn3_bwd_k<true, false>, smem));
      launch_k(a3::attn3_bwd_k<true, false>, dim3(grid), dim3(a3::THREADS), smem, st, PDL_ATTN, mQ, mK, mV, mdO, mO, mImg, mImgDo, a);
    }
    SKF_LAUNCH_CHECK();
  }
  return SKF_OK;
}

}  // namespace skf
