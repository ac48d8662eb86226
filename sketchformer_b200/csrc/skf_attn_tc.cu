// K3 / K13: scaled-dot-product attention (builders/utils.py:71-105) and its backward on the 5th-gen
// tensor cores, head dim 32.  Three kernels with one skeleton:
//
//   TMA (128-row boxes, 128B swizzle)  ->  score MMA(s) into TMEM (tcgen05.mma kind::tf32, SS form)
//   -> one thread per TMEM lane does the row-wise math in registers (scale, -1e9 mask, exp, ...)
//   -> result written back to TMEM (tcgen05.st) and used as the A operand of the second MMA (TS form)
//   -> epilogue from TMEM to global.
//
//   fwd : S = Q_t K^T ; P = softmax(S/sqrt(dh) + mask*-1e9) ; O = P V            (lanes = queries)
//   dq  : S, dP = dO_t V^T ; dS = P (dP - delta) / sqrt(dh) ; dQ = dS K          (lanes = queries)
//   dkv : S^T = K_t Q^T, dP^T = V_t dO^T ; dV = P^T dO ; dK = dS^T Q             (lanes = keys)
//
// The (Tq,Tk) score matrix never exists in HBM; masks come from per-key pad flags + indices.
// Forward can run the error-compensated 3xTF32 product (operands split hi/lo in shared memory after
// the TMA load) so that activations stay fp32-accurate (see DESIGN.md "precision").
#include "skf_common.cuh"
#include "skf_ptx.cuh"

namespace skf {
using namespace ptx;

int make_tmap_2d(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                 uint32_t box_inner, uint32_t box_outer, bool mn_major);
int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert);

constexpr int AT_DH = 32;             // head dim handled here (one 128-byte swizzle row)
constexpr int AT_TILE = 16384;        // 128 rows x 128 B
constexpr float AT_MASK_NEG = -1e9f;

struct AttnTcArgs {
  int H, Tq, Tk, mode;
  int No;                 // "other" extent rounded up to 16 (keys for fwd/dq, queries for dkv)
  float scale;
  const uint8_t* key_pad; // [B, Tk] or nullptr
  float* out0;            // fwd: O      dq: dQ     dkv: dK
  int64_t ld0;
  float* out1;            // fwd: lse    dq: delta  dkv: dV
  int64_t ld1;
  const float* lse;       // dq / dkv
  const float* delta;     // dkv
  const float* O;         // dq (for delta)
  int64_t ldO;
  const float* dO;        // dq (for delta)
  int64_t lddO;
};

__device__ __forceinline__ bool at_masked(int mode, int qi, int kj, int padj) {
  bool m = false;
  if (mode >= 1) m = padj != 0;
  if (mode == 2) m = m || (kj > qi);
  return m;
}

// one elected thread: load a [rows<=256, 32 cols] operand as 1 or 2 boxes of 128 rows
__device__ __forceinline__ void at_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, int col, int row, int nbox) {
  tma_load_2d(dst, m, bar, col, row);
  if (nbox > 1) tma_load_2d(dst + AT_TILE, m, bar, col, row + 128);
}

// hi = tf32(x) in place, lo = x - hi into `lo` (same swizzled position): elementwise, layout-agnostic
__device__ __forceinline__ void at_split(uint8_t* hi, uint8_t* lo, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) {
    float4 v = *reinterpret_cast<float4*>(hi + i), h, l;
    h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + i) = h;
    *reinterpret_cast<float4*>(lo + i) = l;
  }
}

struct AtCommon {
  uint32_t base;       // 1024-aligned smem window
  uint8_t* gen;
  uint32_t bar_load, bar_s, bar_o, tmem_slot;
  uint32_t tmem;
};

__device__ __forceinline__ void at_prologue(AtCommon& c, uint8_t* smem_raw, int data_bytes) {
  const uint32_t raw = smem_u32(smem_raw);
  c.base = (raw + 1023u) & ~1023u;
  c.gen = smem_raw + (c.base - raw);
  c.bar_load = c.base + data_bytes;
  c.bar_s = c.bar_load + 8;
  c.bar_o = c.bar_load + 16;
  c.tmem_slot = c.bar_load + 24;
  if (threadIdx.x == 0) {
    mbar_init(c.bar_load, 1);
    mbar_init(c.bar_s, 1);
    mbar_init(c.bar_o, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(c.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *reinterpret_cast<volatile uint32_t*>(c.gen + data_bytes + 24);
}

__device__ __forceinline__ void at_epilogue_free(const AtCommon& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(c.tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// forward
// smem: Qh | Ql | Kh(2 tiles) | Kl(2) | Vh(2) | Vl(2) | pad[256] | barriers
// tmem: S / P_hi [0,No) ; P_lo [240,240+No) ; O [480,512)
// ---------------------------------------------------------------------------------------------------
constexpr int ATF_DATA = 10 * AT_TILE + 256;
template <bool X3>
__global__ void __launch_bounds__(128, 1)
attn_fwd_tc_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
              const __grid_constant__ CUtensorMap mV, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  AtCommon c;
  at_prologue(c, smem_raw, ATF_DATA);
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nbox = a.No > 128 ? 2 : 1;
  const uint32_t sQh = c.base, sQl = c.base + AT_TILE, sKh = c.base + 2 * AT_TILE, sKl = c.base + 4 * AT_TILE,
                 sVh = c.base + 6 * AT_TILE, sVl = c.base + 8 * AT_TILE;
  uint8_t* pad_s = c.gen + 10 * AT_TILE;

  if (tid == 0) {
    mbar_arrive_expect_tx(c.bar_load, (1 + 2 * nbox) * AT_TILE);
    at_load(sQh, &mQ, c.bar_load, h * AT_DH, b * a.Tq + q0, 1);
    at_load(sKh, &mK, c.bar_load, h * AT_DH, b * a.Tk, nbox);
    at_load(sVh, &mV, c.bar_load, h * AT_DH, b * a.Tk, nbox);
  }
  for (int i = tid; i < 256; i += 128)
    pad_s[i] = (a.key_pad && i < a.Tk) ? a.key_pad[(int64_t)b * a.Tk + i] : 0;
  mbar_wait(c.bar_load, 0);
  if (X3) {
    at_split(c.gen, c.gen + AT_TILE, AT_TILE);
    at_split(c.gen + 2 * AT_TILE, c.gen + 4 * AT_TILE, nbox * AT_TILE);
    at_split(c.gen + 6 * AT_TILE, c.gen + 8 * AT_TILE, nbox * AT_TILE);
    fence_proxy_async();
  }
  __syncthreads();

  const uint32_t tS = c.tmem, tPl = c.tmem + 240, tO = c.tmem + 480;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, a.No, 0, 0);
    uint32_t acc = 0;
    for (int seg = 0; seg < (X3 ? 3 : 1); seg++) {
      const uint32_t qa = seg == 2 ? sQl : sQh;
      const uint32_t ka = seg == 1 ? sKl : sKh;
#pragma unroll
      for (int ks = 0; ks < AT_DH / 8; ks++) {
        mma_tf32_ss(tS, make_desc_kmajor(qa + ks * 32), make_desc_kmajor(ka + ks * 32), idesc, acc);
        acc = 1;
      }
    }
    tc_commit(c.bar_s);
  }
  mbar_wait(c.bar_s, 0);
  tc_fence_after();

  // ---- softmax: this thread owns query row (q0 + tid) == TMEM lane tid ------------------------
  const int qi = q0 + tid;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int nchunk = (a.No + 31) / 32;
  float mx = -INFINITY;
  for (int ch = 0; ch < nchunk; ch++) {
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(tS + lane_off + ch * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int kj = ch * 32 + j;
      float s = __uint_as_float(r[j]) * a.scale;
      if (at_masked(a.mode, qi, kj, pad_s[kj & 255])) s += AT_MASK_NEG;
      if (kj < a.Tk) mx = fmaxf(mx, s);
    }
  }
  float sum = 0.f;
  for (int ch = 0; ch < nchunk; ch++) {
    uint32_t r[32], ph[32], pl[32];
    __syncwarp();
    tmem_ld32(tS + lane_off + ch * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int kj = ch * 32 + j;
      float s = __uint_as_float(r[j]) * a.scale;
      if (at_masked(a.mode, qi, kj, pad_s[kj & 255])) s += AT_MASK_NEG;
      float p = kj < a.Tk ? expf(s - mx) : 0.f;
      sum += p;
      if (X3) {
        float hi = round_tf32(p);
        ph[j] = __float_as_uint(hi);
        pl[j] = __float_as_uint(p - hi);
      } else {
        ph[j] = __float_as_uint(p);
      }
    }
    __syncwarp();
    tmem_st32(tS + lane_off + ch * 32, ph);
    if (X3) tmem_st32(tPl + lane_off + ch * 32, pl);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, AT_DH, 0, 1);
    uint32_t acc = 0;
    const int nk = a.No / 8;
    for (int seg = 0; seg < (X3 ? 3 : 1); seg++) {
      const uint32_t pa = seg == 2 ? tPl : tS;
      const uint32_t va = seg == 1 ? sVl : sVh;
      for (int k = 0; k < nk; k++) {
        mma_tf32_ts(tO, pa + k * 8, make_desc_mnmajor(va + k * 1024, 4096), idesc, acc);
        acc = 1;
      }
    }
    tc_commit(c.bar_o);
  }
  mbar_wait(c.bar_o, 0);
  tc_fence_after();
  {
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(tO + lane_off, r);
    tmem_ld_wait();
    if (qi < a.Tq) {
      const float inv = 1.0f / sum;
      float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + qi) * a.ld0 + h * AT_DH);
#pragma unroll
      for (int j = 0; j < 8; j++)
        dst[j] = make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                             __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
      if (a.out1) a.out1[((int64_t)b * a.H + h) * a.Tq + qi] = mx + logf(sum);
    }
  }
  at_epilogue_free(c);
}

// ---------------------------------------------------------------------------------------------------
// backward, dQ.   smem: Q_t | dO_t | K(2) | V(2) | Kmn(2) | pad[256] | barriers
// tmem: S [0,No) ; dP -> dS [240,240+No) ; dQ [480,512)
// ---------------------------------------------------------------------------------------------------
constexpr int ATQ_DATA = 8 * AT_TILE + 256;
__global__ void __launch_bounds__(128, 1)
attn_bwd_dq_tc_k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mdO,
                 const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                 const __grid_constant__ CUtensorMap mKmn, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  AtCommon c;
  at_prologue(c, smem_raw, ATQ_DATA);
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nbox = a.No > 128 ? 2 : 1;
  const uint32_t sQ = c.base, sdO = c.base + AT_TILE, sK = c.base + 2 * AT_TILE, sV = c.base + 4 * AT_TILE,
                 sKmn = c.base + 6 * AT_TILE;
  uint8_t* pad_s = c.gen + 8 * AT_TILE;
  if (tid == 0) {
    mbar_arrive_expect_tx(c.bar_load, (2 + 3 * nbox) * AT_TILE);
    at_load(sQ, &mQ, c.bar_load, h * AT_DH, b * a.Tq + q0, 1);
    at_load(sdO, &mdO, c.bar_load, h * AT_DH, b * a.Tq + q0, 1);
    at_load(sK, &mK, c.bar_load, h * AT_DH, b * a.Tk, nbox);
    at_load(sV, &mV, c.bar_load, h * AT_DH, b * a.Tk, nbox);
    at_load(sKmn, &mKmn, c.bar_load, h * AT_DH, b * a.Tk, nbox);
  }
  for (int i = tid; i < 256; i += 128)
    pad_s[i] = (a.key_pad && i < a.Tk) ? a.key_pad[(int64_t)b * a.Tk + i] : 0;
  // delta_i = sum_c dO[i,c] O[i,c]  (row-per-thread, 128 B each from global)
  const int qi = q0 + tid;
  float delta = 0.f, lse = 0.f;
  if (qi < a.Tq) {
    const float4* o4 = reinterpret_cast<const float4*>(a.O + ((int64_t)b * a.Tq + qi) * a.ldO + h * AT_DH);
    const float4* g4 = reinterpret_cast<const float4*>(a.dO + ((int64_t)b * a.Tq + qi) * a.lddO + h * AT_DH);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float4 o = o4[j], g = g4[j];
      delta += (o.x * g.x + o.y * g.y) + (o.z * g.z + o.w * g.w);
    }
    lse = a.lse[((int64_t)b * a.H + h) * a.Tq + qi];
    a.out1[((int64_t)b * a.H + h) * a.Tq + qi] = delta;
  }
  mbar_wait(c.bar_load, 0);
  __syncthreads();
  const uint32_t tS = c.tmem, tdP = c.tmem + 240, tdQ = c.tmem + 480;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, a.No, 0, 0);
#pragma unroll
    for (int ks = 0; ks < AT_DH / 8; ks++)
      mma_tf32_ss(tS, make_desc_kmajor(sQ + ks * 32), make_desc_kmajor(sK + ks * 32), idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < AT_DH / 8; ks++)
      mma_tf32_ss(tdP, make_desc_kmajor(sdO + ks * 32), make_desc_kmajor(sV + ks * 32), idesc, ks != 0);
    tc_commit(c.bar_s);
  }
  mbar_wait(c.bar_s, 0);
  tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int nchunk = (a.No + 31) / 32;
  for (int ch = 0; ch < nchunk; ch++) {
    uint32_t rs[32], rp[32];
    __syncwarp();
    tmem_ld32(tS + lane_off + ch * 32, rs);
    tmem_ld32(tdP + lane_off + ch * 32, rp);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int kj = ch * 32 + j;
      float s = __uint_as_float(rs[j]) * a.scale;
      if (at_masked(a.mode, qi, kj, pad_s[kj & 255])) s += AT_MASK_NEG;
      const bool ok = kj < a.Tk && qi < a.Tq;
      float p = ok ? expf(s - lse) : 0.f;
      float ds = ok ? p * (__uint_as_float(rp[j]) - delta) * a.scale : 0.f;
      rp[j] = __float_as_uint(ds);
    }
    __syncwarp();
    tmem_st32(tdP + lane_off + ch * 32, rp);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, AT_DH, 0, 1);
    const int nk = a.No / 8;
    for (int k = 0; k < nk; k++)
      mma_tf32_ts(tdQ, tdP + k * 8, make_desc_mnmajor(sKmn + k * 1024, 4096), idesc, k != 0);
    tc_commit(c.bar_o);
  }
  mbar_wait(c.bar_o, 0);
  tc_fence_after();
  {
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(tdQ + lane_off, r);
    tmem_ld_wait();
    if (qi < a.Tq) {
      float4* dst = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tq + qi) * a.ld0 + h * AT_DH);
#pragma unroll
      for (int j = 0; j < 8; j++)
        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
    }
  }
  at_epilogue_free(c);
}

// ---------------------------------------------------------------------------------------------------
// backward, dK / dV.   smem: K_t | V_t | Q(2) | dO(2) | Qmn(2) | dOmn(2) | lse[256] | delta[256] | barriers
// tmem: S^T -> P^T [0,No) ; dP^T -> dS^T [224,224+No) ; dV [448,480) ; dK [480,512)     (No = queries)
// ---------------------------------------------------------------------------------------------------
constexpr int ATK_DATA = 10 * AT_TILE + 2048;
__global__ void __launch_bounds__(128, 1)
attn_bwd_dkv_tc_k(const __grid_constant__ CUtensorMap mK, const __grid_constant__ CUtensorMap mV,
                  const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mdO,
                  const __grid_constant__ CUtensorMap mQmn, const __grid_constant__ CUtensorMap mdOmn,
                  AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  AtCommon c;
  at_prologue(c, smem_raw, ATK_DATA);
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nbox = a.No > 128 ? 2 : 1;
  const uint32_t sK = c.base, sV = c.base + AT_TILE, sQ = c.base + 2 * AT_TILE, sdO = c.base + 4 * AT_TILE,
                 sQmn = c.base + 6 * AT_TILE, sdOmn = c.base + 8 * AT_TILE;
  float* lse_s = reinterpret_cast<float*>(c.gen + 10 * AT_TILE);
  float* del_s = lse_s + 256;
  if (tid == 0) {
    mbar_arrive_expect_tx(c.bar_load, (2 + 4 * nbox) * AT_TILE);
    at_load(sK, &mK, c.bar_load, h * AT_DH, b * a.Tk + j0, 1);
    at_load(sV, &mV, c.bar_load, h * AT_DH, b * a.Tk + j0, 1);
    at_load(sQ, &mQ, c.bar_load, h * AT_DH, b * a.Tq, nbox);
    at_load(sdO, &mdO, c.bar_load, h * AT_DH, b * a.Tq, nbox);
    at_load(sQmn, &mQmn, c.bar_load, h * AT_DH, b * a.Tq, nbox);
    at_load(sdOmn, &mdOmn, c.bar_load, h * AT_DH, b * a.Tq, nbox);
  }
  for (int i = tid; i < 256; i += 128) {
    lse_s[i] = i < a.Tq ? a.lse[((int64_t)b * a.H + h) * a.Tq + i] : 0.f;
    del_s[i] = i < a.Tq ? a.delta[((int64_t)b * a.H + h) * a.Tq + i] : 0.f;
  }
  const int kj = j0 + tid;
  const int padj = (a.key_pad && kj < a.Tk) ? a.key_pad[(int64_t)b * a.Tk + kj] : 0;
  mbar_wait(c.bar_load, 0);
  __syncthreads();
  const uint32_t tS = c.tmem, tdP = c.tmem + 224, tdV = c.tmem + 448, tdK = c.tmem + 480;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, a.No, 0, 0);
#pragma unroll
    for (int ks = 0; ks < AT_DH / 8; ks++)
      mma_tf32_ss(tS, make_desc_kmajor(sK + ks * 32), make_desc_kmajor(sQ + ks * 32), idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < AT_DH / 8; ks++)
      mma_tf32_ss(tdP, make_desc_kmajor(sV + ks * 32), make_desc_kmajor(sdO + ks * 32), idesc, ks != 0);
    tc_commit(c.bar_s);
  }
  mbar_wait(c.bar_s, 0);
  tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int nchunk = (a.No + 31) / 32;
  for (int ch = 0; ch < nchunk; ch++) {
    uint32_t rs[32], rp[32];
    __syncwarp();
    tmem_ld32(tS + lane_off + ch * 32, rs);
    tmem_ld32(tdP + lane_off + ch * 32, rp);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int qi = ch * 32 + j;
      float s = __uint_as_float(rs[j]) * a.scale;
      if (at_masked(a.mode, qi, kj, padj)) s += AT_MASK_NEG;
      const bool ok = kj < a.Tk && qi < a.Tq;
      float p = ok ? expf(s - lse_s[qi & 255]) : 0.f;
      float ds = ok ? p * (__uint_as_float(rp[j]) - del_s[qi & 255]) * a.scale : 0.f;
      rs[j] = __float_as_uint(p);
      rp[j] = __float_as_uint(ds);
    }
    __syncwarp();
    tmem_st32(tS + lane_off + ch * 32, rs);
    tmem_st32(tdP + lane_off + ch * 32, rp);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, AT_DH, 0, 1);
    const int nk = a.No / 8;
    for (int k = 0; k < nk; k++)
      mma_tf32_ts(tdV, tS + k * 8, make_desc_mnmajor(sdOmn + k * 1024, 4096), idesc, k != 0);
    for (int k = 0; k < nk; k++)
      mma_tf32_ts(tdK, tdP + k * 8, make_desc_mnmajor(sQmn + k * 1024, 4096), idesc, k != 0);
    tc_commit(c.bar_o);
  }
  mbar_wait(c.bar_o, 0);
  tc_fence_after();
  {
    uint32_t rv[32], rk[32];
    __syncwarp();
    tmem_ld32(tdV + lane_off, rv);
    tmem_ld32(tdK + lane_off, rk);
    tmem_ld_wait();
    if (kj < a.Tk) {
      float4* dk = reinterpret_cast<float4*>(a.out0 + ((int64_t)b * a.Tk + kj) * a.ld0 + h * AT_DH);
      float4* dv = reinterpret_cast<float4*>(a.out1 + ((int64_t)b * a.Tk + kj) * a.ld1 + h * AT_DH);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        dk[j] = make_float4(__uint_as_float(rk[4 * j]), __uint_as_float(rk[4 * j + 1]), __uint_as_float(rk[4 * j + 2]),
                            __uint_as_float(rk[4 * j + 3]));
        dv[j] = make_float4(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]), __uint_as_float(rv[4 * j + 2]),
                            __uint_as_float(rv[4 * j + 3]));
      }
    }
  }
  at_epilogue_free(c);
}

// ---- host --------------------------------------------------------------------------------------
int attn_tc_supported(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                      int Tq, int Tk, int dh) {
  if (dh != AT_DH) return 0;
  if (Tk > 224 || Tq > 224) return 0;   // TMEM budget (see kernel headers); longer sequences: SIMT path
  if ((ldq | ldk | ldv) & 3) return 0;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V)) & 15) return 0;
  return 1;
}

static int at_map(CUtensorMap* m, const float* p, int64_t ld, int H, int64_t rows, bool mn, bool conv) {
  return make_tmap_2d_typed(m, p, (uint64_t)H * AT_DH, (uint64_t)rows, (uint64_t)ld, AT_DH, 128, mn, conv);
}

template <class Kern>
static int at_smem_attr(Kern k, int bytes) {
  SKF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return SKF_OK;
}

int attn_fwd_tc(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
                int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh, int mode,
                int x3, cudaStream_t st) {
  SKF_REQUIRE(attn_tc_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn_fwd_tc: shape");
  SKF_REQUIRE((ldo & 3) == 0, SKF_ERR_BAD_ALIGN, "attn_fwd_tc: ldo");
  CUtensorMap mQ, mK, mV;
  // 3xTF32 splits the raw fp32 bits in shared memory; single-pass lets the TMA round to TF32
  SKF_TRY(at_map(&mQ, Q, ldq, H, (int64_t)B * Tq, false, !x3));
  SKF_TRY(at_map(&mK, K, ldk, H, (int64_t)B * Tk, false, !x3));
  SKF_TRY(at_map(&mV, V, ldv, H, (int64_t)B * Tk, true, !x3));
  AttnTcArgs a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.No = (Tk + 15) & ~15; a.scale = 1.0f / sqrtf((float)dh);
  a.key_pad = key_pad; a.out0 = O; a.ld0 = ldo; a.out1 = lse;
  const int smem = ATF_DATA + 1024 + 64;
  dim3 grid(cdiv(Tq, 128), H, B);
  if (x3) {
    static bool done = false;
    if (!done) { SKF_TRY(at_smem_attr(attn_fwd_tc_k<true>, smem)); done = true; }
    attn_fwd_tc_k<true><<<grid, 128, smem, st>>>(mQ, mK, mV, a);
  } else {
    static bool done = false;
    if (!done) { SKF_TRY(at_smem_attr(attn_fwd_tc_k<false>, smem)); done = true; }
    attn_fwd_tc_k<false><<<grid, 128, smem, st>>>(mQ, mK, mV, a);
  }
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int attn_bwd_tc(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                const float* O, int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta,
                float* dQ, int64_t lddq, float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad,
                int B, int H, int Tq, int Tk, int dh, int mode, cudaStream_t st) {
  SKF_REQUIRE(attn_tc_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, dh), SKF_ERR_UNSUPPORTED, "attn_bwd_tc: shape");
  SKF_REQUIRE(((lddo | lddq | lddk | lddv | ldo) & 3) == 0 && (reinterpret_cast<uintptr_t>(dO) & 15) == 0,
              SKF_ERR_BAD_ALIGN, "attn_bwd_tc: alignment");
  CUtensorMap mQ, mK, mV, mdO, mKmn, mQmn, mdOmn;
  SKF_TRY(at_map(&mQ, Q, ldq, H, (int64_t)B * Tq, false, true));
  SKF_TRY(at_map(&mK, K, ldk, H, (int64_t)B * Tk, false, true));
  SKF_TRY(at_map(&mV, V, ldv, H, (int64_t)B * Tk, false, true));
  SKF_TRY(at_map(&mdO, dO, lddo, H, (int64_t)B * Tq, false, true));
  SKF_TRY(at_map(&mKmn, K, ldk, H, (int64_t)B * Tk, true, true));
  SKF_TRY(at_map(&mQmn, Q, ldq, H, (int64_t)B * Tq, true, true));
  SKF_TRY(at_map(&mdOmn, dO, lddo, H, (int64_t)B * Tq, true, true));
  AttnTcArgs a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.Tq = Tq; a.Tk = Tk; a.mode = mode; a.scale = 1.0f / sqrtf((float)dh); a.key_pad = key_pad;
  a.lse = lse; a.O = O; a.ldO = ldo; a.dO = dO; a.lddO = lddo;
  {
    static bool done = false;
    const int smem = ATQ_DATA + 1024 + 64;
    if (!done) { SKF_TRY(at_smem_attr(attn_bwd_dq_tc_k, smem)); done = true; }
    a.No = (Tk + 15) & ~15; a.out0 = dQ; a.ld0 = lddq; a.out1 = delta;
    attn_bwd_dq_tc_k<<<dim3(cdiv(Tq, 128), H, B), 128, smem, st>>>(mQ, mdO, mK, mV, mKmn, a);
    SKF_LAUNCH_CHECK();
  }
  {
    static bool done = false;
    const int smem = ATK_DATA + 1024 + 64;
    if (!done) { SKF_TRY(at_smem_attr(attn_bwd_dkv_tc_k, smem)); done = true; }
    a.No = (Tq + 15) & ~15; a.out0 = dK; a.ld0 = lddk; a.out1 = dV; a.ld1 = lddv; a.delta = delta;
    attn_bwd_dkv_tc_k<<<dim3(cdiv(Tk, 128), H, B), 128, smem, st>>>(mK, mV, mQ, mdO, mQmn, mdOmn, a);
    SKF_LAUNCH_CHECK();
  }
  return SKF_OK;
}

}  // namespace skf
