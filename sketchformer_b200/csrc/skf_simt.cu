// Exact-fp32 SIMT implementations of the two contraction ops (GEMM and attention).  They are
// (a) the numerical cross-check for the tcgen05 kernels on the GPU itself (gemm_backend=1),
// (b) the path for shapes the TMA/tcgen05 kernels do not take (classifier head with C=345:
//     row stride not a multiple of 16 bytes; tiny M).
// They are CUDA kernels of this library, not a CPU fallback.
#include "skf_common.cuh"

namespace skf {

// ---------------------------------------------------------------------------------------------
// Generic strided GEMM, 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ---------------------------------------------------------------------------------------------
constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

__device__ __forceinline__ float epi_apply(float acc, int m, int n, const skf_gemm_epilogue& e) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.act == SKF_ACT_RELU) v = fmaxf(v, 0.f);
  else if (e.act == SKF_ACT_TANH) v = tanhf(v);
  if (e.mask) v = (e.mask[(int64_t)m * e.ld_mask + n] > 0.f) ? v : 0.f;
  if (e.resid) v += e.resid[(int64_t)m * e.ld_resid + n];
  if (e.round_tf32) v = round_tf32(v);
  return v;
}

__global__ void __launch_bounds__(256)
gemm_simt_k(const float* __restrict__ A, int64_t sa_m, int64_t sa_k, const float* __restrict__ B,
            int64_t sb_k, int64_t sb_n, float* __restrict__ Y, int64_t ldy, int M, int N, int K,
            skf_gemm_epilogue e, int k_chunk) {
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Bs[SB_K][SB_N + 4];
  int tid = threadIdx.x;
  int m0 = blockIdx.y * SB_M, n0 = blockIdx.x * SB_N;
  int k_begin = blockIdx.z * k_chunk;
  int k_end = min(K, k_begin + k_chunk);
  int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += SB_K) {
    // A tile: SB_M x SB_K = 1024 elements, 4 per thread; make the contiguous axis the fast one
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int idx = tid + r * 256;
      int mm, kk;
      if (sa_k == 1) { kk = idx % SB_K; mm = idx / SB_K; } else { mm = idx % SB_M; kk = idx / SB_M; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < k_end) ? A[(int64_t)gm * sa_m + (int64_t)gk * sa_k] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int idx = tid + r * 256;
      int nn, kk;
      if (sb_k == 1) { kk = idx % SB_K; nn = idx / SB_K; } else { nn = idx % SB_N; kk = idx / SB_N; }
      int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < k_end) ? B[(int64_t)gk * sb_k + (int64_t)gn * sb_n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SB_K; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* dst = Y + (int64_t)m * ldy + n;
      if (e.accumulate) {
        atomicAdd(dst, acc[i][j]);   // split-K partial sums (no epilogue terms)
      } else {
        *dst = epi_apply(acc[i][j], m, n, e);
      }
    }
  }
}

// Small problems (a handful of 64x64 tiles: the classifier head at batch 128, the [B, E] x [E, 2d] products of the
// folded cross attention): 32x32 tiles put 4x as many CTAs on the machine, and 64-deep k-steps with all of a step's
// global loads in flight cut the serial load -> sync -> FMA chain to K / 64 rounds.
constexpr int SS_M = 32, SS_N = 32, SS_K = 64;
__global__ void __launch_bounds__(256)
gemm_simt_small_k(const float* __restrict__ A, int64_t sa_m, int64_t sa_k, const float* __restrict__ B,
                  int64_t sb_k, int64_t sb_n, float* __restrict__ Y, int64_t ldy, int M, int N, int K,
                  skf_gemm_epilogue e, int k_chunk) {
  __shared__ float As[SS_K][SS_M + 1];
  __shared__ float Bs[SS_K][SS_N + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SS_M, n0 = blockIdx.x * SS_N;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int ty = tid / 16, tx = tid % 16;          // 2 x 2 outputs per thread: rows ty, ty + 16; columns tx, tx + 16
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = k_begin; k0 < k_end; k0 += SS_K) {
    float av[8], bv[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {                  // 32 x 64 elements of each operand, 8 per thread, all in flight
      const int idx = tid + r * 256;
      int mm, kk;
      if (sa_k == 1) { kk = idx % SS_K; mm = idx / SS_K; } else { mm = idx % SS_M; kk = idx / SS_M; }
      const int gm = m0 + mm, gk = k0 + kk;
      av[r] = (gm < M && gk < k_end) ? A[(int64_t)gm * sa_m + (int64_t)gk * sa_k] : 0.f;
      int nn, k2;
      if (sb_k == 1) { k2 = idx % SS_K; nn = idx / SS_K; } else { nn = idx % SS_N; k2 = idx / SS_N; }
      const int gn = n0 + nn, gk2 = k0 + k2;
      bv[r] = (gn < N && gk2 < k_end) ? B[(int64_t)gk2 * sb_k + (int64_t)gn * sb_n] : 0.f;
    }
    __syncthreads();                               // the previous step's reads of As / Bs are done
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int idx = tid + r * 256;
      int mm, kk;
      if (sa_k == 1) { kk = idx % SS_K; mm = idx / SS_K; } else { mm = idx % SS_M; kk = idx / SS_M; }
      As[kk][mm] = av[r];
      int nn, k2;
      if (sb_k == 1) { k2 = idx % SS_K; nn = idx / SS_K; } else { nn = idx % SS_N; k2 = idx / SS_N; }
      Bs[k2][nn] = bv[r];
    }
    __syncthreads();
#pragma unroll 16
    for (int kk = 0; kk < SS_K; kk++) {
      const float a0 = As[kk][ty], a1 = As[kk][ty + 16], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float* dst = Y + (int64_t)m * ldy + n;
      if (e.accumulate) atomicAdd(dst, acc[i][j]);
      else *dst = epi_apply(acc[i][j], m, n, e);
    }
  }
}

int gemm_simt(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
              int64_t sb_n, float* Y, int64_t ldy, int M, int N, int K,
              const skf_gemm_epilogue* epi, cudaStream_t st) {
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  if (epi) e = *epi;
  int split = e.split_k > 1 ? e.split_k : 1;
  SKF_REQUIRE(split == 1 || e.accumulate, SKF_ERR_BAD_ARG, "gemm: split_k needs accumulate");
  SKF_REQUIRE(!(e.accumulate && (e.bias || e.resid || e.mask || e.act)), SKF_ERR_BAD_ARG,
              "gemm: accumulate excludes other epilogue terms");
  SKF_REQUIRE(M > 0 && N > 0 && K > 0, SKF_ERR_BAD_SHAPE, "gemm: empty problem %d %d %d", M, N, K);
  if ((long)cdiv(M, SB_M) * cdiv(N, SB_N) * split <= 37) {     // fewer 64x64 tiles than a quarter of the SMs
    int kc = cdiv(cdiv(K, split), SS_K) * SS_K;
    split = cdiv(K, kc);
    dim3 grid_s(cdiv(N, SS_N), cdiv(M, SS_M), split);
    gemm_simt_small_k<<<grid_s, 256, 0, st>>>(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, e, kc);
    SKF_LAUNCH_CHECK();
    return SKF_OK;
  }
  int k_chunk = cdiv(cdiv(K, split), SB_K) * SB_K;
  split = cdiv(K, k_chunk);
  dim3 grid(cdiv(N, SB_N), cdiv(M, SB_M), split);
  gemm_simt_k<<<grid, 256, 0, st>>>(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, e, k_chunk);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

// ---------------------------------------------------------------------------------------------
// Attention (builders/utils.py:71-105) — flash-style, fp32 SIMT.
// CTA = (query tile of 32 rows, head, batch); 4 warps x 8 rows; keys streamed in chunks of 64
// through shared memory (rows padded to dh+1 floats: conflict-free for both access patterns).
// ---------------------------------------------------------------------------------------------
constexpr int AT_ROWS = 32, AT_KC = 64, AT_RPW = 8;  // rows/CTA, keys/chunk, rows/warp
constexpr float MASK_NEG = -1e9f;

template <int DH>
struct AttSmem {
  float q[AT_ROWS][DH];         // rows of the CTA (q, or k in the dK/dV kernel)
  float q2[AT_ROWS][DH];        // second row operand (dO in dQ kernel, v in dK/dV kernel)
  float k[AT_KC][DH + 1];       // streamed tile A
  float v[AT_KC][DH + 1];       // streamed tile B
  float p[4][AT_KC];            // per-warp probabilities
  float p2[4][AT_KC];
  float aux[AT_KC];             // per-key / per-query scalars of the streamed chunk
  float aux2[AT_KC];
  uint8_t pad[AT_KC];
};

__device__ __forceinline__ bool is_masked(int mode, int qi, int kj, uint8_t padj) {
  bool m = false;
  if (mode >= 1) m = padj != 0;
  if (mode == 2) m = m || (kj > qi);
  return m;
}

template <int DH>
__global__ void __launch_bounds__(128)
attn_fwd_simt_k(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ K, int64_t ldk,
                const float* __restrict__ V, int64_t ldv, float* __restrict__ O, int64_t ldo,
                float* __restrict__ lse, const uint8_t* __restrict__ key_pad, int H, int Tq, int Tk,
                int mode, float scale, int rnd) {
  extern __shared__ __align__(16) unsigned char smraw[];
  AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(smraw);
  constexpr int DPL = (DH + 31) / 32;  // dims per lane
  int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_ROWS;
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < AT_ROWS * DH; i += 128) {
    int r = i / DH, c = i % DH;
    int qi = q0 + r;
    sm.q[r][c] = qi < Tq ? Q[((int64_t)b * Tq + qi) * ldq + h * DH + c] : 0.f;
  }
  float mrun[AT_RPW], lrun[AT_RPW], oacc[AT_RPW][DPL];
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    mrun[r] = -INFINITY;
    lrun[r] = 0.f;
#pragma unroll
    for (int x = 0; x < DPL; x++) oacc[r][x] = 0.f;
  }
  int kmax = Tk;
  if (mode == 2) kmax = min(Tk, q0 + AT_ROWS);  // causal: keys beyond the tile's last row are masked
  // NOTE: skipping fully-masked chunks is exact: their probabilities are exp(-1e9 - m) == 0 in
  // fp32 whenever the row has at least one unmasked key (key 0 is never PAD: SOS / first point).
  for (int k0 = 0; k0 < kmax; k0 += AT_KC) {
    __syncthreads();
    for (int i = tid; i < AT_KC * DH; i += 128) {
      int r = i / DH, c = i % DH;
      int kj = k0 + r;
      bool ok = kj < Tk;
      sm.k[r][c] = ok ? K[((int64_t)b * Tk + kj) * ldk + h * DH + c] : 0.f;
      sm.v[r][c] = ok ? V[((int64_t)b * Tk + kj) * ldv + h * DH + c] : 0.f;
    }
    for (int i = tid; i < AT_KC; i += 128) {
      int kj = k0 + i;
      sm.pad[i] = (key_pad && kj < Tk) ? key_pad[(int64_t)b * Tk + kj] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AT_RPW; r++) {
      int rl = w * AT_RPW + r;
      int qi = q0 + rl;
      float s[2];
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        int j = lane + jj * 32;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < DH; c++) acc = fmaf(sm.q[rl][c], sm.k[j][c], acc);
        acc = acc * scale;
        int kj = k0 + j;
        if (is_masked(mode, qi, kj, sm.pad[j])) acc += MASK_NEG;
        s[jj] = (kj < Tk) ? acc : -INFINITY;
      }
      float mx = warp_max(fmaxf(s[0], s[1]));
      float mnew = fmaxf(mrun[r], mx);
      float corr = expf(mrun[r] - mnew);   // exp(-inf) = 0 on the first chunk
      float p0 = expf(s[0] - mnew), p1 = expf(s[1] - mnew);
      float ps = warp_sum(p0 + p1);
      lrun[r] = lrun[r] * corr + ps;
      mrun[r] = mnew;
      sm.p[w][lane] = p0;
      sm.p[w][lane + 32] = p1;
      __syncwarp();
#pragma unroll
      for (int x = 0; x < DPL; x++) {
        int c = lane + x * 32;
        float acc = oacc[r][x] * corr;
        if (c < DH) {
#pragma unroll 8
          for (int j = 0; j < AT_KC; j++) acc = fmaf(sm.p[w][j], sm.v[j][c], acc);
        }
        oacc[r][x] = acc;
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    int qi = q0 + w * AT_RPW + r;
    if (qi >= Tq) continue;
    float inv = 1.0f / lrun[r];
#pragma unroll
    for (int x = 0; x < DPL; x++) {
      int c = lane + x * 32;
      if (c < DH) {
        float o = oacc[r][x] * inv;
        O[((int64_t)b * Tq + qi) * ldo + h * DH + c] = rnd ? round_tf32(o) : o;
      }
    }
    if (lane == 0 && lse) lse[((int64_t)b * H + h) * Tq + qi] = mrun[r] + logf(lrun[r]);
  }
}

// dQ kernel: same tiling as forward.  Also writes delta[b,h,i] = sum_c dO*O.
template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dq_simt_k(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ K, int64_t ldk,
                   const float* __restrict__ V, int64_t ldv, const float* __restrict__ O, int64_t ldo,
                   const float* __restrict__ dO, int64_t lddo, const float* __restrict__ lse,
                   float* __restrict__ delta, float* __restrict__ dQ, int64_t lddq,
                   const uint8_t* __restrict__ key_pad, int H, int Tq, int Tk, int mode, float scale,
                   int rnd) {
  extern __shared__ __align__(16) unsigned char smraw[];
  AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(smraw);
  constexpr int DPL = (DH + 31) / 32;
  int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_ROWS;
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < AT_ROWS * DH; i += 128) {
    int r = i / DH, c = i % DH;
    int qi = q0 + r;
    bool ok = qi < Tq;
    sm.q[r][c] = ok ? Q[((int64_t)b * Tq + qi) * ldq + h * DH + c] : 0.f;
    sm.q2[r][c] = ok ? dO[((int64_t)b * Tq + qi) * lddo + h * DH + c] : 0.f;
  }
  __syncthreads();
  float dl[AT_RPW], ls[AT_RPW], dq[AT_RPW][DPL];
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    int rl = w * AT_RPW + r, qi = q0 + rl;
    float acc = 0.f;
    if (qi < Tq)
      for (int c = lane; c < DH; c += 32) acc += sm.q2[rl][c] * O[((int64_t)b * Tq + qi) * ldo + h * DH + c];
    acc = warp_sum(acc);
    dl[r] = acc;
    ls[r] = qi < Tq ? lse[((int64_t)b * H + h) * Tq + qi] : 0.f;
    if (lane == 0 && qi < Tq) delta[((int64_t)b * H + h) * Tq + qi] = acc;
#pragma unroll
    for (int x = 0; x < DPL; x++) dq[r][x] = 0.f;
  }
  int kmax = Tk;
  if (mode == 2) kmax = min(Tk, q0 + AT_ROWS);
  for (int k0 = 0; k0 < kmax; k0 += AT_KC) {
    __syncthreads();
    for (int i = tid; i < AT_KC * DH; i += 128) {
      int r = i / DH, c = i % DH;
      int kj = k0 + r;
      bool ok = kj < Tk;
      sm.k[r][c] = ok ? K[((int64_t)b * Tk + kj) * ldk + h * DH + c] : 0.f;
      sm.v[r][c] = ok ? V[((int64_t)b * Tk + kj) * ldv + h * DH + c] : 0.f;
    }
    for (int i = tid; i < AT_KC; i += 128) {
      int kj = k0 + i;
      sm.pad[i] = (key_pad && kj < Tk) ? key_pad[(int64_t)b * Tk + kj] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AT_RPW; r++) {
      int rl = w * AT_RPW + r, qi = q0 + rl;
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        int j = lane + jj * 32;
        float s = 0.f, dp = 0.f;
#pragma unroll 8
        for (int c = 0; c < DH; c++) {
          s = fmaf(sm.q[rl][c], sm.k[j][c], s);
          dp = fmaf(sm.q2[rl][c], sm.v[j][c], dp);
        }
        s *= scale;
        int kj = k0 + j;
        if (is_masked(mode, qi, kj, sm.pad[j])) s += MASK_NEG;
        float p = (kj < Tk) ? expf(s - ls[r]) : 0.f;
        sm.p[w][j] = p * (dp - dl[r]);   // dS
      }
      __syncwarp();
#pragma unroll
      for (int x = 0; x < DPL; x++) {
        int c = lane + x * 32;
        if (c < DH) {
          float acc = dq[r][x];
#pragma unroll 8
          for (int j = 0; j < AT_KC; j++) acc = fmaf(sm.p[w][j], sm.k[j][c], acc);
          dq[r][x] = acc;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    int qi = q0 + w * AT_RPW + r;
    if (qi >= Tq) continue;
#pragma unroll
    for (int x = 0; x < DPL; x++) {
      int c = lane + x * 32;
      if (c < DH) {
        float g = dq[r][x] * scale;
        dQ[((int64_t)b * Tq + qi) * lddq + h * DH + c] = rnd ? round_tf32(g) : g;
      }
    }
  }
}

// dK/dV kernel: CTA owns 32 key rows, streams query chunks of 64.
template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_simt_k(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ K, int64_t ldk,
                    const float* __restrict__ V, int64_t ldv, const float* __restrict__ dO,
                    int64_t lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                    float* __restrict__ dK, int64_t lddk, float* __restrict__ dV, int64_t lddv,
                    const uint8_t* __restrict__ key_pad, int H, int Tq, int Tk, int mode, float scale,
                    int rnd) {
  extern __shared__ __align__(16) unsigned char smraw[];
  AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(smraw);
  constexpr int DPL = (DH + 31) / 32;
  int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * AT_ROWS;
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < AT_ROWS * DH; i += 128) {
    int r = i / DH, c = i % DH;
    int kj = j0 + r;
    bool ok = kj < Tk;
    sm.q[r][c] = ok ? K[((int64_t)b * Tk + kj) * ldk + h * DH + c] : 0.f;
    sm.q2[r][c] = ok ? V[((int64_t)b * Tk + kj) * ldv + h * DH + c] : 0.f;
  }
  float dk[AT_RPW][DPL], dv[AT_RPW][DPL];
  uint8_t padj[AT_RPW];
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    int kj = j0 + w * AT_RPW + r;
    padj[r] = (key_pad && kj < Tk) ? key_pad[(int64_t)b * Tk + kj] : 0;
#pragma unroll
    for (int x = 0; x < DPL; x++) dk[r][x] = dv[r][x] = 0.f;
  }
  int qstart = 0;
  if (mode == 2) qstart = (j0 / AT_KC) * AT_KC;  // causal: queries before the tile's first key see none
  for (int i0 = qstart; i0 < Tq; i0 += AT_KC) {
    __syncthreads();
    for (int i = tid; i < AT_KC * DH; i += 128) {
      int r = i / DH, c = i % DH;
      int qi = i0 + r;
      bool ok = qi < Tq;
      sm.k[r][c] = ok ? Q[((int64_t)b * Tq + qi) * ldq + h * DH + c] : 0.f;
      sm.v[r][c] = ok ? dO[((int64_t)b * Tq + qi) * lddo + h * DH + c] : 0.f;
    }
    for (int i = tid; i < AT_KC; i += 128) {
      int qi = i0 + i;
      sm.aux[i] = qi < Tq ? lse[((int64_t)b * H + h) * Tq + qi] : 0.f;
      sm.aux2[i] = qi < Tq ? delta[((int64_t)b * H + h) * Tq + qi] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AT_RPW; r++) {
      int rl = w * AT_RPW + r, kj = j0 + rl;
#pragma unroll
      for (int ii = 0; ii < 2; ii++) {
        int i = lane + ii * 32;
        int qi = i0 + i;
        float s = 0.f, dp = 0.f;
#pragma unroll 8
        for (int c = 0; c < DH; c++) {
          s = fmaf(sm.k[i][c], sm.q[rl][c], s);     // q_i . k_j
          dp = fmaf(sm.v[i][c], sm.q2[rl][c], dp);  // dO_i . v_j
        }
        s *= scale;
        if (is_masked(mode, qi, kj, padj[r])) s += MASK_NEG;
        float p = (qi < Tq && kj < Tk) ? expf(s - sm.aux[i]) : 0.f;
        sm.p[w][i] = p;
        sm.p2[w][i] = p * (dp - sm.aux2[i]);
      }
      __syncwarp();
#pragma unroll
      for (int x = 0; x < DPL; x++) {
        int c = lane + x * 32;
        if (c < DH) {
          float a1 = dv[r][x], a2 = dk[r][x];
#pragma unroll 8
          for (int i = 0; i < AT_KC; i++) {
            a1 = fmaf(sm.p[w][i], sm.v[i][c], a1);
            a2 = fmaf(sm.p2[w][i], sm.k[i][c], a2);
          }
          dv[r][x] = a1;
          dk[r][x] = a2;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; r++) {
    int kj = j0 + w * AT_RPW + r;
    if (kj >= Tk) continue;
#pragma unroll
    for (int x = 0; x < DPL; x++) {
      int c = lane + x * 32;
      if (c < DH) {
        float gk = dk[r][x] * scale, gv = dv[r][x];
        dK[((int64_t)b * Tk + kj) * lddk + h * DH + c] = rnd ? round_tf32(gk) : gk;
        dV[((int64_t)b * Tk + kj) * lddv + h * DH + c] = rnd ? round_tf32(gv) : gv;
      }
    }
  }
}

template <int DH>
static int attn_fwd_launch(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V,
                           int64_t ldv, float* O, int64_t ldo, float* lse, const uint8_t* key_pad,
                           int B, int H, int Tq, int Tk, int mode, int rnd, cudaStream_t st) {
  size_t smem = sizeof(AttSmem<DH>);
  SKF_CUDA(cudaFuncSetAttribute(attn_fwd_simt_k<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(Tq, AT_ROWS), H, B);
  attn_fwd_simt_k<DH><<<grid, 128, smem, st>>>(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, H, Tq, Tk,
                                              mode, 1.0f / sqrtf((float)DH), rnd);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

template <int DH>
static int attn_bwd_launch(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V,
                           int64_t ldv, const float* O, int64_t ldo, const float* dO, int64_t lddo,
                           const float* lse, float* delta, float* dQ, int64_t lddq, float* dK,
                           int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad, int B, int H,
                           int Tq, int Tk, int mode, int rnd, cudaStream_t st) {
  size_t smem = sizeof(AttSmem<DH>);
  SKF_CUDA(cudaFuncSetAttribute(attn_bwd_dq_simt_k<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SKF_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_simt_k<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float scale = 1.0f / sqrtf((float)DH);
  attn_bwd_dq_simt_k<DH><<<dim3(cdiv(Tq, AT_ROWS), H, B), 128, smem, st>>>(
      Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, key_pad, H, Tq, Tk, mode, scale, rnd);
  SKF_LAUNCH_CHECK();
  attn_bwd_dkv_simt_k<DH><<<dim3(cdiv(Tk, AT_ROWS), H, B), 128, smem, st>>>(
      Q, ldq, K, ldk, V, ldv, dO, lddo, lse, delta, dK, lddk, dV, lddv, key_pad, H, Tq, Tk, mode, scale, rnd);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int attn_fwd_simt(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                  float* O, int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk,
                  int dh, int mode, int rnd, cudaStream_t st) {
  switch (dh) {
    case 16: return attn_fwd_launch<16>(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, mode, rnd, st);
    case 32: return attn_fwd_launch<32>(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, mode, rnd, st);
    case 64: return attn_fwd_launch<64>(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, mode, rnd, st);
    default: return fail(SKF_ERR_UNSUPPORTED, "attention: head dim %d not in {16,32,64}", dh);
  }
}

int attn_bwd_simt(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                  const float* O, int64_t ldo, const float* dO, int64_t lddo, const float* lse,
                  float* delta, float* dQ, int64_t lddq, float* dK, int64_t lddk, float* dV, int64_t lddv,
                  const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh, int mode, int rnd,
                  cudaStream_t st) {
#define SKF_ARGS Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, key_pad, B, H, Tq, Tk, mode, rnd, st
  switch (dh) {
    case 16: return attn_bwd_launch<16>(SKF_ARGS);
    case 32: return attn_bwd_launch<32>(SKF_ARGS);
    case 64: return attn_bwd_launch<64>(SKF_ARGS);
    default: return fail(SKF_ERR_UNSUPPORTED, "attention: head dim %d not in {16,32,64}", dh);
  }
#undef SKF_ARGS
}

}  // namespace skf
