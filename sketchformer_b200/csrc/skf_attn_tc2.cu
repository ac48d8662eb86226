// Scaled-dot-product attention (builders/utils.py:71-105, split_heads transformer.py:160-165) and its
// backward on the 5th-gen tensor cores, head dim 32 or 64 (template DH; a [rows x dh] tile is dh/32
// chunk tiles of 128-byte rows), sequences up to 1024, flash-style: the (Tq,Tk) score matrix never
// exists in HBM, masks come from per-key pad flags + indices, and the other sequence is streamed in
// 64-row blocks.  At head dim 32 a CTA needs 70-105 KB of shared memory and 128-256 TMEM columns and
// two or three CTAs share an SM (their TMA / MMA / softmax phases overlap each other); at head dim 64
// (cfg5: d 512, T 512) it is one CTA of 132-205 KB and up to 384 TMEM columns per SM.
//
//   fwd : per key block    S = Q K^T ; online softmax (running max / sum, O rescaled in TMEM) ; O += P V
//   dq  : per key block    S, dP = dO V^T ; dS = P (dP - delta) / sqrt(dh) ; dQ += dS K      (lanes = queries)
//   dkv : per query block  S^T = K Q^T, dP^T = V dO^T ; dV += P^T dO ; dK += dS^T Q         (lanes = keys)
//
// One thread per TMEM lane does the row math in registers and writes P / dS back to TMEM, where the
// second MMA reads it as its A operand (tcgen05.mma TS form).
//
// Precision (X3 = true, the default of the plan): products of two operands that come from memory
// (Q K^T, dO V^T) run as the error-compensated 3xTF32 sum  Ah*Bh + Ah*Bl + Al*Bh  with the hi/lo split
// done in shared memory after the TMA load (kind::tf32 truncates, so hi is what the tensor core sees
// of the raw tile and lo = x - trunc(x)).  Products whose A operand is produced in registers
// (P, dS) round it to TF32 (nearest) once and use two passes  A*Bh + A*Bl ; the softmax row sum is
// taken over the rounded P, so numerator and denominator see the same weights.
#include "skf_common.cuh"
#include "skf_ptx.cuh"

namespace skf {
using namespace ptx;

int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert);

constexpr int A2_DH = 32;              // columns of one tile chunk (one 128-byte swizzle row); head dim = NC chunks
constexpr int A2_BLK = 64;             // rows of the streamed operand per iteration
constexpr int A2_T128 = 128 * 128;     // 128 rows x 128 B
constexpr int A2_T64 = 64 * 128;       //  64 rows x 128 B
constexpr int A2_MAXT = 1024;          // longest sequence (bias / lse / delta staging); maximum_position_encoding = 1000
constexpr float A2_MASK_NEG = -1e9f;

struct Attn2Args {
  int H, Tq, Tk, mode;
  float scale;
  const uint8_t* key_pad;  // [B, Tk] or nullptr
  float* out0;             // fwd: O      dq: dQ     dkv: dK
  int64_t ld0;
  float* out1;             // fwd: lse    dq: delta  dkv: dV
  int64_t ld1;
  const float* lse;        // dq / dkv
  const float* delta;      // dkv
  const float* O;          // dq (for delta)
  int64_t ldO;
  const float* dO;         // dq (for delta)
  int64_t lddO;
};

__device__ __forceinline__ bool a2_masked(int mode, int qi, int kj, int padj) {
  bool m = false;
  if (mode >= 1) m = padj != 0;
  if (mode == 2) m = m || (kj > qi);
  return m;
}

// lo = x - trunc_tf32(x), elementwise over a tile (the raw and lo tiles share their swizzled layout)
__device__ __forceinline__ void a2_split(const uint8_t* raw, uint8_t* lo, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) {
    const float4 v = *reinterpret_cast<const float4*>(raw + i);
    float4 l;
    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(lo + i) = l;
  }
}

// A [rows x dh] tile is NC = dh / 32 chunks of [rows x 32] (128-byte rows), chunk c at + c * rows * 128 bytes;
// its lo tile follows the raw tile.  TMA: one box per chunk.
template <int NC>
__device__ __forceinline__ void a2_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, int col0, int row0, int rows) {
#pragma unroll
  for (int c = 0; c < NC; c++) tma_load_2d(dst + c * rows * 128, m, bar, col0 + A2_DH * c, row0);
}

// D[128 x N] = A B^T, both K-major smem tiles of NC chunks (4 k-steps each); X3: Ah*Bl + Al*Bh + Ah*Bh.
// ca / cb: chunk strides of the A / B tile in bytes (rows * 128).
template <bool X3, int NC>
__device__ __forceinline__ void a2_mma_ss(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl,
                                          uint32_t idesc, uint32_t ca, uint32_t cb) {
  uint32_t acc = 0;
  if (X3) {
#pragma unroll
    for (int ks = 0; ks < NC * 4; ks++) {
      const uint32_t ao = (ks >> 2) * ca + (ks & 3) * 32, bo = (ks >> 2) * cb + (ks & 3) * 32;
      mma_tf32_ss(d, make_desc_kmajor(ah + ao), make_desc_kmajor(bl + bo), idesc, acc);
      acc = 1;
    }
#pragma unroll
    for (int ks = 0; ks < NC * 4; ks++) {
      const uint32_t ao = (ks >> 2) * ca + (ks & 3) * 32, bo = (ks >> 2) * cb + (ks & 3) * 32;
      mma_tf32_ss(d, make_desc_kmajor(al + ao), make_desc_kmajor(bh + bo), idesc, 1u);
    }
  }
#pragma unroll
  for (int ks = 0; ks < NC * 4; ks++) {
    const uint32_t ao = (ks >> 2) * ca + (ks & 3) * 32, bo = (ks >> 2) * cb + (ks & 3) * 32;
    mma_tf32_ss(d, make_desc_kmajor(ah + ao), make_desc_kmajor(bh + bo), idesc, acc);
    acc = 1;
  }
}

// D[128 x dh] (+)= A[tmem, 64 columns] * B[64 rows x dh, MN-major smem: NC 32-wide N-chunks A2_T64 bytes apart];
// X3: A*Bl + A*Bh.  N comes from the instruction descriptor.
template <bool X3>
__device__ __forceinline__ void a2_mma_ts(uint32_t d, uint32_t a_tmem, uint32_t bh, uint32_t bl, uint32_t idesc,
                                          uint32_t accumulate) {
  uint32_t acc = accumulate;
  if (X3) {
#pragma unroll
    for (int k = 0; k < A2_BLK / 8; k++) {
      mma_tf32_ts(d, a_tmem + k * 8, make_desc_mnmajor(bl + k * 1024, A2_T64), idesc, acc);
      acc = 1;
    }
  }
#pragma unroll
  for (int k = 0; k < A2_BLK / 8; k++) {
    mma_tf32_ts(d, a_tmem + k * 8, make_desc_mnmajor(bh + k * 1024, A2_T64), idesc, acc);
    acc = 1;
  }
}

// A [rows x 32 fp32] tile has the same 128-byte rows in both operand roles; only the intra-row swizzle differs
// (pinned on B200 with tools/swz_probe.py): K-major tiles (CU_TENSOR_MAP_SWIZZLE_128B) keep logical 16-byte
// chunk q of row r at chunk q ^ (r & 7), MN-major tiles (SWIZZLE_128B_ATOM_32B) at chunk q ^ ((r & 3) << 1).
// Re-swizzling a tile in shared memory therefore replaces the second TMA load (and its ~2.5 us round trip)
// the backward kernels used to issue for the MN-major copy of the same rows.
__device__ __forceinline__ uint32_t a2_kchunk(int r, int q) { return (uint32_t)(q ^ (r & 7)) << 4; }
__device__ __forceinline__ uint32_t a2_mnchunk(int r, int q) { return (uint32_t)(q ^ ((r & 3) << 1)) << 4; }

// 64-row tile, 256 threads: thread t copies logical chunks 2g, 2g+1 (g = t >> 6) of row t & 63, raw and lo
__device__ __forceinline__ void a2_kmajor_to_mnmajor_64(const uint8_t* src, const uint8_t* src_lo, uint8_t* dst,
                                                        uint8_t* dst_lo, bool with_lo) {
  const int r = threadIdx.x & 63, g = threadIdx.x >> 6;
#pragma unroll
  for (int q = 2 * g; q < 2 * g + 2; q++) {
    const uint32_t so = r * 128 + a2_kchunk(r, q), dof = r * 128 + a2_mnchunk(r, q);
    *reinterpret_cast<float4*>(dst + dof) = *reinterpret_cast<const float4*>(src + so);
    if (with_lo) *reinterpret_cast<float4*>(dst_lo + dof) = *reinterpret_cast<const float4*>(src_lo + so);
  }
}
// in place, one thread per row (all 8 chunks through registers); rows [0, 64) of one tile
__device__ __forceinline__ void a2_row_kmajor_to_mnmajor_inplace(uint8_t* tile, int r) {
  float4 c[8];
#pragma unroll
  for (int q = 0; q < 8; q++) c[q] = *reinterpret_cast<const float4*>(tile + r * 128 + a2_kchunk(r, q));
#pragma unroll
  for (int q = 0; q < 8; q++) *reinterpret_cast<float4*>(tile + r * 128 + a2_mnchunk(r, q)) = c[q];
}

// D[128 x 2 dh] (+)= A[tmem, 64 columns] * [Bh | Bl]: the raw tile and its lo tile (contiguous 32-wide chunks,
// A2_T64 bytes apart in shared memory) are read by ONE MMA as the 2 NC N-chunks of an MN-major operand, so columns
// [0,dh) of D accumulate A*Bh and [dh,2dh) A*Bl (summed by the reader).  Dependent 128x32x8 MMAs cost ~80-100 clk each whatever
// their N (clock64 stamps: 16 of them 1300-1800 clk, 8 of N=64 850 clk), so half as many wider ones are cheaper.
__device__ __forceinline__ void a2_mma_ts_cat(uint32_t d, uint32_t a_tmem, uint32_t bh, uint32_t idesc64,
                                              uint32_t accumulate) {
  uint32_t acc = accumulate;
#pragma unroll
  for (int k = 0; k < A2_BLK / 8; k++) {
    mma_tf32_ts(d, a_tmem + k * 8, make_desc_mnmajor(bh + k * 1024, A2_T64), idesc64, acc);
    acc = 1;
  }
}

struct A2Common {
  uint32_t base;   // 1024-aligned smem window
  uint8_t* gen;
  uint32_t bar0;   // 8 mbarriers
  uint32_t tmem;
};

template <int TMEM_COLS>
__device__ __forceinline__ void a2_prologue(A2Common& c, uint8_t* smem_raw, int data_bytes) {
  const uint32_t raw = smem_u32(smem_raw);
  c.base = (raw + 1023u) & ~1023u;
  c.gen = smem_raw + (c.base - raw);
  c.bar0 = c.base + data_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; i++) mbar_init(c.bar0 + 8 * i, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(c.bar0 + 64, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *reinterpret_cast<volatile uint32_t*>(c.gen + data_bytes + 64);
}

template <int TMEM_COLS>
__device__ __forceinline__ void a2_free(const A2Common& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(c.tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// forward.  smem: Q | Ql | K | Kl | V | Vl | bias[256] | barriers     tmem: S/P [0,64)  O [64,96)
//
// Measured phase times of one CTA at cfg2 (clock64 stamps, 3 CTAs/SM, cycles per 64-key iteration of ~5400):
// K split + sync 460, S MMAs (12 x 128x64x8) 1000, softmax + P store + O rescale 1900, V split + sync 470,
// P V MMAs (16 x 128x32x8) 1300; prologue 2600, epilogue 1500.  Dependent small MMAs cost ~80 clk each whatever
// their N (feeding [Vh | Vl] as ONE N=64 operand cut the P V phase to 850 clk but not the kernel time), so the
// kernel is bound by the serial phase chain of each CTA, not by a pipe: the next step is a warp-specialised
// two-tile ping-pong (softmax of one query tile under the MMAs of the other), not more tuning of this structure.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int a2f_data(int nc) { return nc * (2 * A2_T128 + 4 * A2_T64) + A2_MAXT * 4; }
constexpr float A2_LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float a2_ex2(float x) {      // 2^x, MUFU.EX2 (2 ulp; the result is rounded to TF32 anyway)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool X3, int DH>
__global__ void __launch_bounds__(128, DH == 32 ? 3 : 1)
attn2_fwd_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
            const __grid_constant__ CUtensorMap mV, const Attn2Args a) {
  extern __shared__ uint8_t smem_raw[];
  A2Common c;
  constexpr int NC = DH / A2_DH, T128 = NC * A2_T128, T64 = NC * A2_T64;
  a2_prologue<128>(c, smem_raw, a2f_data(NC));
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sQ = c.base, sQl = sQ + T128, sK = sQl + T128, sKl = sK + T64, sV = sKl + T64,
                 sVl = sV + T64;
  uint8_t* gQ = c.gen;
  uint8_t* gK = gQ + 2 * T128;
  uint8_t* gV = gK + 2 * T64;
  float* bias_s = reinterpret_cast<float*>(gV + 2 * T64);
  const uint32_t bar_q = c.bar0, bar_k = c.bar0 + 8, bar_v = c.bar0 + 16, bar_s = c.bar0 + 24, bar_o = c.bar0 + 32;
  const int kmax = a.mode == 2 ? min(a.Tk, q0 + 128) : a.Tk;   // causal: keys above the tile's last row never count
  const int nblk = (kmax + A2_BLK - 1) / A2_BLK;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, T128);
    a2_load<NC>(sQ, &mQ, bar_q, h * DH, b * a.Tq + q0, 128);
    mbar_arrive_expect_tx(bar_k, T64);
    a2_load<NC>(sK, &mK, bar_k, h * DH, b * a.Tk, A2_BLK);
    mbar_arrive_expect_tx(bar_v, T64);
    a2_load<NC>(sV, &mV, bar_v, h * DH, b * a.Tk, A2_BLK);
  }
  // logits are handled in the log2 domain: t = S * (scale*log2e) + bias[key]; bias carries the reference's
  // additive -1e9 mask (builders/utils.py:93-97) for padded keys and -inf for keys outside the sequence
  for (int i = tid; i < nblk * A2_BLK; i += 128) {
    float bj = 0.f;
    if (i >= a.Tk) bj = -INFINITY;
    else if (a.mode >= 1 && a.key_pad && a.key_pad[(int64_t)b * a.Tk + i]) bj = A2_MASK_NEG * A2_LOG2E;
    bias_s[i] = bj;
  }
  mbar_wait(bar_q, 0);
  if (X3) a2_split(gQ, gQ + T128, T128);
  const float c1 = a.scale * A2_LOG2E;
  const int qw = q0 + (tid & ~31);            // first query row of this warp

  const uint32_t tS = c.tmem, tO = c.tmem + 64;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc_s = make_idesc_tf32(128, A2_BLK, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(128, DH, 0, 1);
  const int qi = q0 + tid;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nblk; j++) {
    const uint32_t ph = (uint32_t)j & 1u;
    const int k0 = j * A2_BLK;
    mbar_wait(bar_k, ph);
    if (X3) a2_split(gK, gK + T64, T64);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      a2_mma_ss<X3, NC>(tS, sQ, sQl, sK, sKl, idesc_s, A2_T128, A2_T64);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {     // K buffers are free again: prefetch the next key block
      mbar_arrive_expect_tx(bar_k, T64);
      a2_load<NC>(sK, &mK, bar_k, h * DH, b * a.Tk + k0 + A2_BLK, A2_BLK);
    }
    // ---- online softmax: this thread owns query row qi == TMEM lane tid ----------------------
    uint32_t r0[32], r1[32];
    __syncwarp();
    tmem_ld32(tS + lane_off, r0);
    tmem_ld32(tS + lane_off + 32, r1);
    tmem_ld_wait();
    float bmax = -INFINITY;
    const bool diag = a.mode == 2 && k0 + A2_BLK - 1 > qw;      // warp-uniform: some (q, k) pair has k > q
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + k0);
#pragma unroll
    for (int jj = 0; jj < 32; jj += 4) {
      const float4 bv = b4[jj >> 2];
      float t0 = fmaf(__uint_as_float(r0[jj]), c1, bv.x), t1 = fmaf(__uint_as_float(r0[jj + 1]), c1, bv.y);
      float t2 = fmaf(__uint_as_float(r0[jj + 2]), c1, bv.z), t3 = fmaf(__uint_as_float(r0[jj + 3]), c1, bv.w);
      if (diag) {
        const int kj = k0 + jj;
        if (kj > qi) t0 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 1 > qi) t1 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 2 > qi) t2 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 3 > qi) t3 += A2_MASK_NEG * A2_LOG2E;
      }
      r0[jj] = __float_as_uint(t0); r0[jj + 1] = __float_as_uint(t1);
      r0[jj + 2] = __float_as_uint(t2); r0[jj + 3] = __float_as_uint(t3);
      bmax = fmaxf(fmaxf(bmax, fmaxf(t0, t1)), fmaxf(t2, t3));
    }
#pragma unroll
    for (int jj = 0; jj < 32; jj += 4) {
      const float4 bv = b4[8 + (jj >> 2)];
      float t0 = fmaf(__uint_as_float(r1[jj]), c1, bv.x), t1 = fmaf(__uint_as_float(r1[jj + 1]), c1, bv.y);
      float t2 = fmaf(__uint_as_float(r1[jj + 2]), c1, bv.z), t3 = fmaf(__uint_as_float(r1[jj + 3]), c1, bv.w);
      if (diag) {
        const int kj = k0 + 32 + jj;
        if (kj > qi) t0 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 1 > qi) t1 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 2 > qi) t2 += A2_MASK_NEG * A2_LOG2E;
        if (kj + 3 > qi) t3 += A2_MASK_NEG * A2_LOG2E;
      }
      r1[jj] = __float_as_uint(t0); r1[jj + 1] = __float_as_uint(t1);
      r1[jj + 2] = __float_as_uint(t2); r1[jj + 3] = __float_as_uint(t3);
      bmax = fmaxf(fmaxf(bmax, fmaxf(t0, t1)), fmaxf(t2, t3));
    }
    const float m_new = fmaxf(m_run, bmax);            // finite: key 0 of every block row set exists (k0 < Tk)
    const float alpha = a2_ex2(m_run - m_new);         // 0 on the first block (m_run = -inf)
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
      const float p = round_tf32(a2_ex2(__uint_as_float(r0[jj]) - m_new));
      sum += p;
      r0[jj] = __float_as_uint(p);
    }
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
      const float p = round_tf32(a2_ex2(__uint_as_float(r1[jj]) - m_new));
      sum += p;
      r1[jj] = __float_as_uint(p);
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;
    __syncwarp();
    tmem_st32(tS + lane_off, r0);
    tmem_st32(tS + lane_off + 32, r1);
    if (j > 0) {                                       // rescale the running output (previous P V has completed)
#pragma unroll
      for (int cc = 0; cc < NC; cc++) {
        tmem_ld32(tO + lane_off + 32 * cc, r0);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; jj++) r0[jj] = __float_as_uint(__uint_as_float(r0[jj]) * alpha);
        tmem_st32(tO + lane_off + 32 * cc, r0);
      }
    }
    tmem_st_wait();
    mbar_wait(bar_v, ph);
    if (X3) a2_split(gV, gV + T64, T64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      a2_mma_ts<X3>(tO, tS, sV, sVl, idesc_o, j > 0 ? 1u : 0u);
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {     // V buffers are free again
      mbar_arrive_expect_tx(bar_v, T64);
      a2_load<NC>(sV, &mV, bar_v, h * DH, b * a.Tk + k0 + A2_BLK, A2_BLK);
    }
  }
  {
    uint32_t r[32];
    const float inv = 1.0f / l_run;
#pragma unroll
    for (int cc = 0; cc < NC; cc++) {
      __syncwarp();
      tmem_ld32(tO + lane_off + 32 * cc, r);
      tmem_ld_wait();
      if (qi < a.Tq) {
        float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + qi) * a.ld0 + h * DH + 32 * cc);
#pragma unroll
        for (int jj = 0; jj < 8; jj++)
          dst[jj] = make_float4(__uint_as_float(r[4 * jj]) * inv, __uint_as_float(r[4 * jj + 1]) * inv,
                                __uint_as_float(r[4 * jj + 2]) * inv, __uint_as_float(r[4 * jj + 3]) * inv);
      }
    }
    // natural-log logsumexp of the masked scaled logits (m_run is in the log2 domain)
    if (qi < a.Tq && a.out1) a.out1[((int64_t)b * a.H + h) * a.Tq + qi] = m_run * 0.6931471805599453f + logf(l_run);
  }
  a2_free<128>(c);
}

// ---------------------------------------------------------------------------------------------------
// backward kernels: 256 threads = two warpgroups over the same 128 TMEM lanes; warpgroup g owns columns
// [32g, 32g+32) of every 64-column block, so the row math of a block is spread over 8 warps.
// The logits are handled in the log2 domain: t = S * (scale*log2e) + bias, p = exp2(t - lse*log2e), where
// bias carries the reference's additive -1e9 mask (builders/utils.py:93-97) for padded keys and -inf for
// keys / queries outside the sequence.  dS is accumulated WITHOUT the trailing 1/sqrt(dh), which is
// applied once to the dQ / dK accumulators in the epilogue.
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// backward, dQ.  smem: Q | Ql | dO | dOl | K | Kl | X | Xl | bias[256] | barriers   (X = V block, then the
// MN-major copy of the K block).   tmem: S [0,64)  dP -> dS [64,128)  dQ [128,192) = dS*Kh | dS*Kl
// Measured phases of one CTA before the two changes below (clock64, cycles): prologue 8500 (of which the
// per-row global loads of O and dO for delta ~6000), per 64-key block S,dP MMAs 2200, re-swizzle 400-900,
// row math 1400-2100, dQ MMAs (16 x N=32) 1200-1800, splits 500; epilogue 2200.  Now the O tile comes in by
// TMA into the (not yet used) Q-lo buffer and delta is taken from shared memory, and dQ uses 8 N=64 MMAs.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int a2q_data(int nc) { return nc * (4 * A2_T128 + 4 * A2_T64) + A2_MAXT * 4; }
template <bool X3, int DH>
__global__ void __launch_bounds__(256, DH == 32 ? 2 : 1)
attn2_bwd_dq_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mdO,
               const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
               const __grid_constant__ CUtensorMap mO, const Attn2Args a) {
  extern __shared__ uint8_t smem_raw[];
  A2Common c;
  constexpr int NC = DH / A2_DH, T128 = NC * A2_T128, T64 = NC * A2_T64;
  constexpr int TCOLS = 256;                                      // S 64 + dP 64 + dQ (2) dh
  a2_prologue<TCOLS>(c, smem_raw, a2q_data(NC));
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, row = tid & 127;
  const uint32_t sQ = c.base, sQl = sQ + T128, sdO = sQl + T128, sdOl = sdO + T128, sK = sdOl + T128,
                 sKl = sK + T64, sX = sKl + T64, sXl = sX + T64;
  uint8_t* gQ = c.gen;
  uint8_t* gdO = gQ + 2 * T128;
  uint8_t* gK = gdO + 2 * T128;
  uint8_t* gX = gK + 2 * T64;
  float* bias_s = reinterpret_cast<float*>(gX + 2 * T64);
  const uint32_t bar_q = c.bar0, bar_k = c.bar0 + 8, bar_v = c.bar0 + 16, bar_x = c.bar0 + 24, bar_s = c.bar0 + 32,
                 bar_o = c.bar0 + 40;
  const int kmax = a.mode == 2 ? min(a.Tk, q0 + 128) : a.Tk;
  const int nblk = (kmax + A2_BLK - 1) / A2_BLK;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * T128);
    a2_load<NC>(sQ, &mQ, bar_q, h * DH, b * a.Tq + q0, 128);
    a2_load<NC>(sdO, &mdO, bar_q, h * DH, b * a.Tq + q0, 128);
    mbar_arrive_expect_tx(bar_x, T128);                      // O tile, parked in the Q-lo buffer
    a2_load<NC>(sQl, &mO, bar_x, h * DH, b * a.Tq + q0, 128);
    mbar_arrive_expect_tx(bar_k, T64);
    a2_load<NC>(sK, &mK, bar_k, h * DH, b * a.Tk, A2_BLK);
    mbar_arrive_expect_tx(bar_v, T64);
    a2_load<NC>(sX, &mV, bar_v, h * DH, b * a.Tk, A2_BLK);
  }
  for (int i = tid; i < nblk * A2_BLK; i += 256) {
    float bj = 0.f;
    if (i >= a.Tk) bj = -INFINITY;
    else if (a.mode >= 1 && a.key_pad && a.key_pad[(int64_t)b * a.Tk + i]) bj = A2_MASK_NEG * A2_LOG2E;
    bias_s[i] = bj;
  }
  const int qi = q0 + row;
  float delta = 0.f, lse2 = INFINITY;       // rows past the sequence: p = exp2(-inf) = 0
  if (qi < a.Tq) lse2 = a.lse[((int64_t)b * a.H + h) * a.Tq + qi] * A2_LOG2E;
  mbar_wait(bar_q, 0);
  mbar_wait(bar_x, 0);
  {
    // delta_i = sum_c dO[i,c] O[i,c] from the two K-major tiles in shared memory (row per thread; rows past
    // the sequence are TMA zero-fill)
#pragma unroll
    for (int cc = 0; cc < NC; cc++) {
      const uint8_t* orow = gQ + T128 + cc * A2_T128 + row * 128;
      const uint8_t* grow = gdO + cc * A2_T128 + row * 128;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float4 o = *reinterpret_cast<const float4*>(orow + a2_kchunk(row, q));
        const float4 g = *reinterpret_cast<const float4*>(grow + a2_kchunk(row, q));
        delta += (o.x * g.x + o.y * g.y) + (o.z * g.z + o.w * g.w);
      }
    }
    if (wg == 0 && qi < a.Tq) a.out1[((int64_t)b * a.H + h) * a.Tq + qi] = delta;
  }
  __syncthreads();                          // every thread has read the O tile: the Q-lo buffer may be overwritten
  if (X3) {
    a2_split(gQ, gQ + T128, T128);
    a2_split(gdO, gdO + T128, T128);
  }
  const uint32_t tS = c.tmem, tdP = c.tmem + 64, tdQ = c.tmem + 128;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t idesc_s = make_idesc_tf32(128, A2_BLK, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(128, X3 ? 2 * DH : DH, 0, 1);
  const float c1 = a.scale * A2_LOG2E;
  const int qw = q0 + (warp & 3) * 32;      // first query row of this warp

  for (int j = 0; j < nblk; j++) {
    const uint32_t ph = (uint32_t)j & 1u;
    const int k0 = j * A2_BLK;
    mbar_wait(bar_k, ph);
    mbar_wait(bar_v, ph);
    if (X3) {
      a2_split(gK, gK + T64, T64);
      a2_split(gX, gX + T64, T64);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      a2_mma_ss<X3, NC>(tS, sQ, sQl, sK, sKl, idesc_s, A2_T128, A2_T64);
      a2_mma_ss<X3, NC>(tdP, sdO, sdOl, sX, sXl, idesc_s, A2_T128, A2_T64);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    // the V block is dead: rebuild the key block (raw + lo) in its buffers with the MN-major swizzle for the
    // dQ MMA, then hand the K-major buffers to the TMA for the next key block
#pragma unroll
    for (int cc = 0; cc < NC; cc++)
      a2_kmajor_to_mnmajor_64(gK + cc * A2_T64, gK + T64 + cc * A2_T64, gX + cc * A2_T64, gX + T64 + cc * A2_T64, X3);
    __syncthreads();
    if (tid == 0 && j + 1 < nblk) {
      mbar_arrive_expect_tx(bar_k, T64);
      a2_load<NC>(sK, &mK, bar_k, h * DH, b * a.Tk + k0 + A2_BLK, A2_BLK);
    }
    {
      uint32_t rs[32], rp[32];
      const int kc = k0 + wg * 32;          // first key column of this thread's half block
      __syncwarp();
      tmem_ld32(tS + lane_off + wg * 32, rs);
      tmem_ld32(tdP + lane_off + wg * 32, rp);
      tmem_ld_wait();
      const bool diag = a.mode == 2 && kc + 31 > qw;      // warp-uniform: some (q, k) pair has k > q
      const float4* b4 = reinterpret_cast<const float4*>(bias_s + kc);
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const float4 bv = b4[j4];
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int jj = 4 * j4 + e;
          float t = fmaf(__uint_as_float(rs[jj]), c1, bb[e] - lse2);
          if (diag && kc + jj > qi) t += A2_MASK_NEG * A2_LOG2E;
          const float p = a2_ex2(t);
          rp[jj] = __float_as_uint(round_tf32(p * (__uint_as_float(rp[jj]) - delta)));
        }
      }
      __syncwarp();
      tmem_st32(tdP + lane_off + wg * 32, rp);
    }
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (X3) a2_mma_ts_cat(tdQ, tdP, sX, idesc_o, j > 0 ? 1u : 0u);
      else a2_mma_ts<false>(tdQ, tdP, sX, sXl, idesc_o, j > 0 ? 1u : 0u);
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {
      mbar_arrive_expect_tx(bar_v, T64);
      a2_load<NC>(sX, &mV, bar_v, h * DH, b * a.Tk + k0 + A2_BLK, A2_BLK);
    }
  }
  if (wg == 0) {
    uint32_t r[32];
#pragma unroll
    for (int cc = 0; cc < NC; cc++) {
      __syncwarp();
      tmem_ld32(tdQ + lane_off + 32 * cc, r);
      if (X3) {
        uint32_t rl[32];
        tmem_ld32(tdQ + lane_off + DH + 32 * cc, rl);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; jj++) r[jj] = __float_as_uint(__uint_as_float(r[jj]) + __uint_as_float(rl[jj]));
      }
      tmem_ld_wait();
      if (qi < a.Tq) {
        float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + qi) * a.ld0 + h * DH + 32 * cc);
#pragma unroll
        for (int jj = 0; jj < 8; jj++)
          dst[jj] = make_float4(__uint_as_float(r[4 * jj]) * a.scale, __uint_as_float(r[4 * jj + 1]) * a.scale,
                                __uint_as_float(r[4 * jj + 2]) * a.scale, __uint_as_float(r[4 * jj + 3]) * a.scale);
      }
    }
  }
  a2_free<TCOLS>(c);
}

// ---------------------------------------------------------------------------------------------------
// backward, dK / dV.  smem: K | Kl | V | Vl | A | Al | B | Bl | lse2[256] | delta[256] | barriers
//   (A, B = Q block, dO block K-major for the score MMAs, then their MN-major copies for the output MMAs)
// tmem: S^T -> P^T [0,64)  dP^T -> dS^T [64,128)  dV [128,192) = P^T dOh | P^T dOl   dK [192,256) likewise
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int a2k_data(int nc) { return nc * (4 * A2_T128 + 4 * A2_T64) + 2 * A2_MAXT * 4; }
template <bool X3, int DH>
__global__ void __launch_bounds__(256, DH == 32 ? 2 : 1)
attn2_bwd_dkv_k(const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mdO,
                const Attn2Args a) {
  extern __shared__ uint8_t smem_raw[];
  A2Common c;
  constexpr int NC = DH / A2_DH, T128 = NC * A2_T128, T64 = NC * A2_T64;
  constexpr int TCOLS = DH == 32 ? 256 : 512;                     // S 64 + dP 64 + dV 2 dh + dK 2 dh
  a2_prologue<TCOLS>(c, smem_raw, a2k_data(NC));
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, row = tid & 127;
  const uint32_t sK = c.base, sKl = sK + T128, sV = sKl + T128, sVl = sV + T128, sA = sVl + T128,
                 sAl = sA + T64, sB = sAl + T64, sBl = sB + T64;
  uint8_t* gK = c.gen;
  uint8_t* gV = gK + 2 * T128;
  uint8_t* gA = gV + 2 * T128;
  uint8_t* gB = gA + 2 * T64;
  float* lse_s = reinterpret_cast<float*>(gB + 2 * T64);     // lse * log2e, +inf for queries past the sequence
  float* del_s = lse_s + A2_MAXT;
  const uint32_t bar_q = c.bar0, bar_a = c.bar0 + 8, bar_b = c.bar0 + 16, bar_s = c.bar0 + 24, bar_o = c.bar0 + 32;
  const int nqb = (a.Tq + A2_BLK - 1) / A2_BLK;
  const int ib0 = a.mode == 2 ? j0 / A2_BLK : 0;     // causal: queries before the tile's first key see none of it

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * T128);
    a2_load<NC>(sK, &mK, bar_q, h * DH, b * a.Tk + j0, 128);
    a2_load<NC>(sV, &mV, bar_q, h * DH, b * a.Tk + j0, 128);
    if (ib0 < nqb) {
      mbar_arrive_expect_tx(bar_a, 2 * T64);
      a2_load<NC>(sA, &mQ, bar_a, h * DH, b * a.Tq + ib0 * A2_BLK, A2_BLK);
      a2_load<NC>(sB, &mdO, bar_a, h * DH, b * a.Tq + ib0 * A2_BLK, A2_BLK);
    }
  }
  for (int i = tid; i < nqb * A2_BLK; i += 256) {
    lse_s[i] = i < a.Tq ? a.lse[((int64_t)b * a.H + h) * a.Tq + i] * A2_LOG2E : INFINITY;
    del_s[i] = i < a.Tq ? a.delta[((int64_t)b * a.H + h) * a.Tq + i] : 0.f;
  }
  const int kj = j0 + row;
  float bias = 0.f;                                   // this thread's key: -inf outside the sequence, -1e9 if padded
  if (kj >= a.Tk) bias = -INFINITY;
  else if (a.mode >= 1 && a.key_pad && a.key_pad[(int64_t)b * a.Tk + kj]) bias = A2_MASK_NEG * A2_LOG2E;
  mbar_wait(bar_q, 0);
  if (X3) {
    a2_split(gK, gK + T128, T128);
    a2_split(gV, gV + T128, T128);
  }
  const uint32_t tS = c.tmem, tdP = c.tmem + 64, tdV = c.tmem + 128, tdK = c.tmem + 128 + 2 * DH;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t idesc_s = make_idesc_tf32(128, A2_BLK, 0, 0);
  const uint32_t idesc_o = make_idesc_tf32(128, X3 ? 2 * DH : DH, 0, 1);
  const float c1 = a.scale * A2_LOG2E;
  const int kw = j0 + (warp & 3) * 32;               // first key of this warp

  int it = 0;
  for (int ib = ib0; ib < nqb; ib++, it++) {
    const uint32_t ph = (uint32_t)it & 1u;
    const int i0 = ib * A2_BLK;
    mbar_wait(bar_a, ph);
    if (X3) {
      a2_split(gA, gA + T64, T64);
      a2_split(gB, gB + T64, T64);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      a2_mma_ss<X3, NC>(tS, sK, sKl, sA, sAl, idesc_s, A2_T128, A2_T64);      // S^T  = K Q^T
      a2_mma_ss<X3, NC>(tdP, sV, sVl, sB, sBl, idesc_s, A2_T128, A2_T64);     // dP^T = V dO^T
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    // the score MMAs are done with the K-major Q / dO blocks: re-swizzle them in place (raw and lo tiles, one
    // thread per 128-byte row) into the MN-major layout the output MMAs read
    {
#pragma unroll
      for (int cc = 0; cc < NC; cc++) {
        const int ti = (tid >> 6) + 4 * cc;          // gA, gA lo, gB, gB lo: 4 NC contiguous 64-row chunk tiles
        if (X3 || ((ti / NC) & 1) == 0) a2_row_kmajor_to_mnmajor_inplace(gA + ti * A2_T64, tid & 63);
      }
    }
    {
      uint32_t rs[32], rp[32];
      const int qc = i0 + wg * 32;          // first query column of this thread's half block
      __syncwarp();
      tmem_ld32(tS + lane_off + wg * 32, rs);
      tmem_ld32(tdP + lane_off + wg * 32, rp);
      tmem_ld_wait();
      const bool diag = a.mode == 2 && kw + 31 > qc;      // warp-uniform: some (q, k) pair has k > q
      const float4* l4 = reinterpret_cast<const float4*>(lse_s + qc);
      const float4* d4 = reinterpret_cast<const float4*>(del_s + qc);
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const float4 lv = l4[j4], dv = d4[j4];
        const float ll[4] = {lv.x, lv.y, lv.z, lv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int jj = 4 * j4 + e;
          float t = fmaf(__uint_as_float(rs[jj]), c1, bias - ll[e]);
          if (diag && kj > qc + jj) t += A2_MASK_NEG * A2_LOG2E;
          const float p = a2_ex2(t);
          rs[jj] = __float_as_uint(round_tf32(p));
          rp[jj] = __float_as_uint(round_tf32(p * (__uint_as_float(rp[jj]) - dd[e])));
        }
      }
      __syncwarp();
      tmem_st32(tS + lane_off + wg * 32, rs);
      tmem_st32(tdP + lane_off + wg * 32, rp);
    }
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (X3) {
        a2_mma_ts_cat(tdV, tS, sB, idesc_o, it > 0 ? 1u : 0u);          // dV += P^T [dOh | dOl]
        a2_mma_ts_cat(tdK, tdP, sA, idesc_o, it > 0 ? 1u : 0u);         // dK += dS^T [Qh | Ql]
      } else {
        a2_mma_ts<false>(tdV, tS, sB, sBl, idesc_o, it > 0 ? 1u : 0u);
        a2_mma_ts<false>(tdK, tdP, sA, sAl, idesc_o, it > 0 ? 1u : 0u);
      }
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && ib + 1 < nqb) {
      mbar_arrive_expect_tx(bar_a, 2 * T64);
      a2_load<NC>(sA, &mQ, bar_a, h * DH, b * a.Tq + i0 + A2_BLK, A2_BLK);
      a2_load<NC>(sB, &mdO, bar_a, h * DH, b * a.Tq + i0 + A2_BLK, A2_BLK);
    }
  }
  {
    // warpgroup 0 stores dK (scaled by 1/sqrt(dh)), warpgroup 1 stores dV
    uint32_t r[32];
#pragma unroll
    for (int cc = 0; cc < NC; cc++) {
      __syncwarp();
      if (it > 0) {
        tmem_ld32((wg == 0 ? tdK : tdV) + lane_off + 32 * cc, r);
        if (X3) {
          uint32_t rl[32];
          tmem_ld32((wg == 0 ? tdK : tdV) + lane_off + DH + 32 * cc, rl);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; jj++) r[jj] = __float_as_uint(__uint_as_float(r[jj]) + __uint_as_float(rl[jj]));
        }
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; jj++) r[jj] = 0u;
      }
      if (kj < a.Tk) {
        const float sc = wg == 0 ? a.scale : 1.0f;
        float* base = wg == 0 ? a.out0 + ((int64_t)b * a.Tk + kj) * a.ld0 : a.out1 + ((int64_t)b * a.Tk + kj) * a.ld1;
        float4* dst = reinterpret_cast<float4*>(base + h * DH + 32 * cc);
#pragma unroll
        for (int jj = 0; jj < 8; jj++)
          dst[jj] = make_float4(__uint_as_float(r[4 * jj]) * sc, __uint_as_float(r[4 * jj + 1]) * sc,
                                __uint_as_float(r[4 * jj + 2]) * sc, __uint_as_float(r[4 * jj + 3]) * sc);
      }
    }
  }
  a2_free<TCOLS>(c);
}

// ---- host --------------------------------------------------------------------------------------
int attn2_supported(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, int Tq,
                    int Tk, int dh) {
  if (dh != 32 && dh != 64) return 0;
  if (Tk > A2_MAXT || Tq > A2_MAXT || Tq < 1 || Tk < 1) return 0;
  if ((ldq | ldk | ldv) & 3) return 0;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V)) & 15) return 0;
  return 1;
}

// rows: box height (128 for the lane-side operand, 64 for the streamed one); conv: TMA rounds to TF32
template <class Kern>
static int a2_attr(Kern k, int bytes) {
  SKF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  SKF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return SKF_OK;
}

static int g_a2_dh = 32;   // head dim of the maps being built (set by attn2_fwd / attn2_bwd; host-side, single-threaded per call)
static int a2_map(CUtensorMap* m, const float* p, int64_t ld, int H, int64_t nrows, int rows, bool mn, bool conv) {
  return make_tmap_2d_typed(m, p, (uint64_t)H * g_a2_dh, (uint64_t)nrows, (uint64_t)ld, A2_DH, (uint32_t)rows, mn, conv);
}

template <class Kern, class... Args>
static int a2_launch(Kern k, bool& attr_done, dim3 grid, int threads, int smem, cudaStream_t st, const Args&... args) {
  if (!attr_done) { SKF_TRY(a2_attr(k, smem)); attr_done = true; }
  k<<<grid, threads, smem, st>>>(args...);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}
// one instantiation per (X3, DH); `done` flags are per instantiation
template <bool X3, int DH>
static int a2_fwd_launch(dim3 grid, cudaStream_t st, const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV,
                         const Attn2Args& a) {
  static bool done = false;
  return a2_launch(attn2_fwd_k<X3, DH>, done, grid, 128, a2f_data(DH / A2_DH) + 1024 + 128, st, mQ, mK, mV, a);
}
template <bool X3, int DH>
static int a2_dq_launch(dim3 grid, cudaStream_t st, const CUtensorMap& mQ, const CUtensorMap& mdO, const CUtensorMap& mK,
                        const CUtensorMap& mV, const CUtensorMap& mO, const Attn2Args& a) {
  static bool done = false;
  return a2_launch(attn2_bwd_dq_k<X3, DH>, done, grid, 256, a2q_data(DH / A2_DH) + 1024 + 128, st, mQ, mdO, mK, mV, mO, a);
}
template <bool X3, int DH>
static int a2_dkv_launch(dim3 grid, cudaStream_t st, const CUtensorMap& mK, const CUtensorMap& mV, const CUtensorMap& mQ,
                         const CUtensorMap& mdO, const Attn2Args& a) {
  static bool done = false;
  return a2_launch(attn2_bwd_dkv_k<X3, DH>, done, grid, 256, a2k_data(DH / A2_DH) + 1024 + 128, st, mK, mV, mQ, mdO, a);
}
#define A2_DISPATCH(fn, ...)                                                        \
  (dh == 32 ? (x3 ? fn<true, 32>(__VA_ARGS__) : fn<false, 32>(__VA_ARGS__))        \
            : (x3 ? fn<true, 64>(__VA_ARGS__) : fn<false, 64>(__VA_ARGS__)))

int attn2_fwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
              int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh, int mode, int x3,
              cudaStream_t st) {
  SKF_REQUIRE(attn2_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn2_fwd: shape");
  SKF_REQUIRE((ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, SKF_ERR_BAD_ALIGN, "attn2_fwd: O");
  SKF_REQUIRE(mode == 0 || key_pad != nullptr || mode == 2, SKF_ERR_BAD_ARG, "attn2_fwd: mode 1 needs key_pad");
  CUtensorMap mQ, mK, mV;
  g_a2_dh = dh;
  SKF_TRY(a2_map(&mQ, Q, ldq, H, (int64_t)B * Tq, 128, false, !x3));
  SKF_TRY(a2_map(&mK, K, ldk, H, (int64_t)B * Tk, A2_BLK, false, !x3));
  SKF_TRY(a2_map(&mV, V, ldv, H, (int64_t)B * Tk, A2_BLK, true, !x3));
  Attn2Args a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.scale = 1.0f / sqrtf((float)dh);
  a.key_pad = key_pad; a.out0 = O; a.ld0 = ldo; a.out1 = lse;
  dim3 grid(cdiv(Tq, 128), H, B);
  return A2_DISPATCH(a2_fwd_launch, grid, st, mQ, mK, mV, a);
}

int attn2_bwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, const float* O,
              int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta, float* dQ, int64_t lddq,
              float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad, int B, int H, int Tq, int Tk,
              int dh, int mode, int x3, cudaStream_t st) {
  SKF_REQUIRE(attn2_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn2_bwd: shape");
  SKF_REQUIRE(((lddo | lddq | lddk | lddv | ldo) & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(dQ) | reinterpret_cast<uintptr_t>(dK) |
                    reinterpret_cast<uintptr_t>(dV) | reinterpret_cast<uintptr_t>(O)) & 15) == 0,
              SKF_ERR_BAD_ALIGN, "attn2_bwd: alignment");
  const bool conv = !x3;
  CUtensorMap mQ128, mdO128, mO128, mK64, mV64, mK128, mV128, mQ64, mdO64;
  g_a2_dh = dh;
  SKF_TRY(a2_map(&mO128, O, ldo, H, (int64_t)B * Tq, 128, false, false));
  SKF_TRY(a2_map(&mQ128, Q, ldq, H, (int64_t)B * Tq, 128, false, conv));
  SKF_TRY(a2_map(&mdO128, dO, lddo, H, (int64_t)B * Tq, 128, false, conv));
  SKF_TRY(a2_map(&mK64, K, ldk, H, (int64_t)B * Tk, A2_BLK, false, conv));
  SKF_TRY(a2_map(&mV64, V, ldv, H, (int64_t)B * Tk, A2_BLK, false, conv));
  SKF_TRY(a2_map(&mK128, K, ldk, H, (int64_t)B * Tk, 128, false, conv));
  SKF_TRY(a2_map(&mV128, V, ldv, H, (int64_t)B * Tk, 128, false, conv));
  SKF_TRY(a2_map(&mQ64, Q, ldq, H, (int64_t)B * Tq, A2_BLK, false, conv));
  SKF_TRY(a2_map(&mdO64, dO, lddo, H, (int64_t)B * Tq, A2_BLK, false, conv));
  Attn2Args a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.scale = 1.0f / sqrtf((float)dh); a.key_pad = key_pad;
  a.lse = lse; a.O = O; a.ldO = ldo; a.dO = dO; a.lddO = lddo;
  {
    a.out0 = dQ; a.ld0 = lddq; a.out1 = delta;
    dim3 grid(cdiv(Tq, 128), H, B);
    SKF_TRY(A2_DISPATCH(a2_dq_launch, grid, st, mQ128, mdO128, mK64, mV64, mO128, a));
  }
  {
    a.out0 = dK; a.ld0 = lddk; a.out1 = dV; a.ld1 = lddv; a.delta = delta;
    dim3 grid(cdiv(Tk, 128), H, B);
    SKF_TRY(A2_DISPATCH(a2_dkv_launch, grid, st, mK128, mV128, mQ64, mdO64, a));
  }
  return SKF_OK;
}

}  // namespace skf

// ---- debug: raw shared-memory image of one TMA box (used once to pin the 32B-atom swizzle pattern) --------
namespace skf {
__global__ void tma_dump_k(const __grid_constant__ CUtensorMap m, float* out, int bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar = base + bytes;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(bar, bytes);
    ptx::tma_load_2d(base, &m, bar, 0, 0);
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(gen)[i];
}
}  // namespace skf
extern "C" int skf_debug_tma_dump(const float* src, int rows, int mn, float* out, void* stream) {
  CUtensorMap m;
  SKF_TRY(skf::make_tmap_2d_typed(&m, src, 32, (uint64_t)rows, 32, 32, (uint32_t)rows, mn != 0, false));
  const int bytes = rows * 128;
  skf::tma_dump_k<<<1, 128, bytes + 1024 + 64, skf::as_stream(stream)>>>(m, out, bytes);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}
