// HBM-bound kernels of the Sketchformer hot path: embedding gather, LayerNorm fwd/bwd, column
// sums, pooling bottleneck, DenseExpander, softmax-CE heads, Adam.  Each kernel cites the
// reference op it replaces; all are coalesced along the feature axis and use warp shuffles for
// row reductions.  (The tensor-core GEMM / attention kernels live in skf_gemm_tc.cu etc.)
#include <math.h>
#include <stdarg.h>

#include "skf_common.cuh"

namespace skf {

thread_local char g_err[512] = {0};
thread_local int64_t g_launch_count = 0;
int g_pdl = 0;
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

__device__ __forceinline__ float maybe_round(float v, int r) { return r ? round_tf32(v) : v; }

// ---------------------------------------------------------------------------------------------
// K1: embedding gather * sqrt(d) + PE     (builders/layers/transformer.py:292-294, 330-332)
// one warp per (b,t) row, 128-bit loads/stores along d
// ---------------------------------------------------------------------------------------------
__global__ void embed_fwd_k(const int64_t* __restrict__ tok, int64_t tok_ld,
                            const float* __restrict__ table, const float* __restrict__ pe,
                            float* __restrict__ out, int B, int T, int d, int V, float scale,
                            int rnd) {
  pdl_enter();
  int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= B * T) return;
  int b = row / T, t = row - b * T;
  int64_t id = tok[(int64_t)b * tok_ld + t];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const float4* e4 = reinterpret_cast<const float4*>(table + id * d);
  const float4* p4 = reinterpret_cast<const float4*>(pe + (int64_t)t * d);
  float4* o4 = reinterpret_cast<float4*>(out + (int64_t)row * d);
  for (int c = threadIdx.x; c < d / 4; c += 32) {
    float4 e = __ldg(e4 + c), p = __ldg(p4 + c), o;
    // two roundings like the reference (x *= sqrt(d); x += pe), no FMA contraction
    o.x = maybe_round(__fadd_rn(__fmul_rn(e.x, scale), p.x), rnd);
    o.y = maybe_round(__fadd_rn(__fmul_rn(e.y, scale), p.y), rnd);
    o.z = maybe_round(__fadd_rn(__fmul_rn(e.z, scale), p.z), rnd);
    o.w = maybe_round(__fadd_rn(__fmul_rn(e.w, scale), p.w), rnd);
    o4[c] = o;
  }
}

// dtable[tok] += dout * sqrt(d)  (tf IndexedSlices -> dense-equivalent)
// One warp walks EMB_RUN consecutive positions of one sequence and merges runs of equal token ids in registers
// before touching memory: the PAD tail of a sequence (37 % of all positions at cfg2, all hitting table row 0)
// costs one atomic per run instead of one per position.
constexpr int EMB_RUN = 16;
template <int NV>        // d <= 128 * NV
__global__ void embed_bwd_k(const int64_t* __restrict__ tok, int64_t tok_ld,
                            const float* __restrict__ dout, float* __restrict__ dtable, int B,
                            int T, int d, int V, float scale) {
  pdl_enter();
  const int chunks = (T + EMB_RUN - 1) / EMB_RUN;
  const int w = blockIdx.x * blockDim.y + threadIdx.y;
  if (w >= B * chunks) return;
  const int b = w / chunks, t0 = (w - b * chunks) * EMB_RUN, t1 = min(T, t0 + EMB_RUN);
  const int lane = threadIdx.x, nv = d / 4;
  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t cur = -1;
  auto flush = [&]() {
    if (cur < 0) return;
    float* dst = dtable + cur * d;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = lane + 32 * k;
      if (c < nv) {
        // one 16-byte vector reduction (red.global.add.v4.f32, sm_90+) instead of four scalar atomics
        atomicAdd(reinterpret_cast<float4*>(dst) + c,
                  make_float4(acc[k].x * scale, acc[k].y * scale, acc[k].z * scale, acc[k].w * scale));
      }
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  for (int t = t0; t < t1; t++) {
    int64_t id = tok[(int64_t)b * tok_ld + t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    if (id != cur) { flush(); cur = id; }
    const float4* g = reinterpret_cast<const float4*>(dout + ((int64_t)b * T + t) * d);
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = lane + 32 * k;
      if (c < nv) {
        const float4 v = g[c];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
  }
  flush();
}

// builders/utils.py:17-32 in float64, rounded once to fp32
__global__ void posenc_k(float* __restrict__ pe, int rows, int d) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * d) return;
  int p = idx / d, i = idx - p * d;
  double rate = 1.0 / pow(10000.0, (double)(2 * (i / 2)) / (double)(float)d);
  double a = (double)p * rate;
  pe[idx] = (float)((i & 1) ? cos(a) : sin(a));
}

__global__ void pad_flags_k(const int64_t* __restrict__ tok, int64_t tok_ld,
                            uint8_t* __restrict__ pad, int B, int T) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * T) return;
  int b = idx / T, t = idx - b * T;
  pad[idx] = tok[(int64_t)b * tok_ld + t] == 0 ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------
// K1c / K10: continuous (stroke-5) input embedding and output loss, use_continuous_data=True
// (builders/layers/transformer.py:275-276, 292-294; builders/losses.py:43-66; builders/utils.py:38-39)
// ---------------------------------------------------------------------------------------------
// pad[b,t] = (x[b,t,4] == 1)
__global__ void pad_flags_cont_k(const float* __restrict__ x, int64_t x_ld, uint8_t* __restrict__ pad, int B, int T) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * T) return;
  int b = idx / T, t = idx - b * T;
  pad[idx] = x[(int64_t)b * x_ld + (int64_t)t * 5 + 4] == 1.0f ? 1 : 0;
}

// out[b,t,:] = (x[b,t,:5] . W[5,d] + bias) * sqrt(d) + PE[t,:]     (Dense(d_model) on the stroke-5 row)
__global__ void embed_cont_fwd_k(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ W,
                                 const float* __restrict__ bias, const float* __restrict__ pe,
                                 float* __restrict__ out, int B, int T, int d, float scale, int rnd) {
  int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= B * T) return;
  int b = row / T, t = row - b * T;
  const float* xr = x + (int64_t)b * x_ld + (int64_t)t * 5;
  float x0 = xr[0], x1 = xr[1], x2 = xr[2], x3 = xr[3], x4 = xr[4];
  for (int c = threadIdx.x; c < d; c += 32) {
    float acc = x0 * __ldg(W + c);
    acc = fmaf(x1, __ldg(W + d + c), acc);
    acc = fmaf(x2, __ldg(W + 2 * d + c), acc);
    acc = fmaf(x3, __ldg(W + 3 * d + c), acc);
    acc = fmaf(x4, __ldg(W + 4 * d + c), acc);
    acc += __ldg(bias + c);
    out[(int64_t)row * d + c] = maybe_round(__fadd_rn(__fmul_rn(acc, scale), __ldg(pe + (int64_t)t * d + c)), rnd);
  }
}

// dW[k,c] += sqrt(d) * sum_rows x[row,k] dout[row,c] ; db[c] += sqrt(d) * sum_rows dout[row,c]
// thread per column, blockIdx.y strides the rows
__global__ void embed_cont_bwd_k(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ dout,
                                 float* __restrict__ dW, float* __restrict__ db, int B, int T, int d, float scale) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, ab = 0.f;
  for (int row = blockIdx.y; row < B * T; row += gridDim.y) {
    int b = row / T, t = row - b * T;
    const float* xr = x + (int64_t)b * x_ld + (int64_t)t * 5;
    float g = dout[(int64_t)row * d + c];
    a0 = fmaf(xr[0], g, a0); a1 = fmaf(xr[1], g, a1); a2 = fmaf(xr[2], g, a2);
    a3 = fmaf(xr[3], g, a3); a4 = fmaf(xr[4], g, a4); ab += g;
  }
  atomicAdd(dW + c, a0 * scale); atomicAdd(dW + d + c, a1 * scale); atomicAdd(dW + 2 * d + c, a2 * scale);
  atomicAdd(dW + 3 * d + c, a3 * scale); atomicAdd(dW + 4 * d + c, a4 * scale); atomicAdd(db + c, ab * scale);
}

// builders/losses.py:43-66 sums.  stats[0] += sum loc*mask, stats[1] += sum CE3, stats[2] += sum mask
// loc = mean_2((real_xy - pred_xy)^2), CE3 = lse(pred[2:5]) - pred[2 + argmax(real[2:5])], mask = real[4] != 1
__device__ __forceinline__ void cont_row(const float* r, const float* p, float& loc, float& ce, float& mask,
                                         int& label, float* sm3) {
  float d0 = r[0] - p[0], d1 = r[1] - p[1];
  loc = 0.5f * (d0 * d0 + d1 * d1);
  label = 0;
  float best = r[2];
  if (r[3] > best) { best = r[3]; label = 1; }
  if (r[4] > best) { best = r[4]; label = 2; }
  float mx = fmaxf(p[2], fmaxf(p[3], p[4]));
  float e0 = expf(p[2] - mx), e1 = expf(p[3] - mx), e2 = expf(p[4] - mx);
  float se = e0 + e1 + e2;
  ce = mx + logf(se) - p[2 + label];
  mask = r[4] != 1.0f ? 1.f : 0.f;
  if (sm3) { float inv = 1.0f / se; sm3[0] = e0 * inv; sm3[1] = e1 * inv; sm3[2] = e2 * inv; }
}

__global__ void cont_loss_stats_k(const float* __restrict__ pred, int64_t ldp, const float* __restrict__ real,
                                  int64_t real_ld, int B, int Td, double* __restrict__ stats) {
  __shared__ float red[3][8];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float loc = 0.f, ce = 0.f, mask = 0.f;
  if (idx < B * Td) {
    int b = idx / Td, t = idx - b * Td, label;
    float l;
    cont_row(real + (int64_t)b * real_ld + (int64_t)t * 5, pred + (int64_t)idx * ldp, l, ce, mask, label, nullptr);
    loc = l * mask;
  }
  loc = warp_sum(loc); ce = warp_sum(ce); mask = warp_sum(mask);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = loc; red[1][w] = ce; red[2][w] = mask; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0, m = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { a += red[0][i]; c += red[1][i]; m += red[2][i]; }
    atomicAdd(stats + 0, a); atomicAdd(stats + 1, c); atomicAdd(stats + 2, m);
  }
}

// d loss / d pred with loss = weight * mean((loc + M) * mask), M = mean(CE3):
//   d/dxy = weight/N * mask * (pred - real) ; d/dpen = weight * (sum mask / N) / N * (softmax - onehot)
__global__ void cont_loss_grad_k(const float* pred, int64_t ldp, const float* __restrict__ real,
                                 int64_t real_ld, int B, int Td, const double* __restrict__ stats, float weight,
                                 float* dpred, int64_t lddp) {   // dpred may alias pred (row-local read, then write)
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Td) return;
  int b = idx / Td, t = idx - b * Td, label;
  const float* r = real + (int64_t)b * real_ld + (int64_t)t * 5;
  const float* p = pred + (int64_t)idx * ldp;
  float loc, ce, mask, sm3[3];
  cont_row(r, p, loc, ce, mask, label, sm3);
  const float invN = 1.0f / (float)((double)B * Td);
  const float wpen = weight * (float)(stats[2] * (double)invN) * invN;
  float* g = dpred + (int64_t)idx * lddp;
  const float p0 = p[0], p1 = p[1];
  g[0] = weight * invN * mask * (p0 - r[0]);
  g[1] = weight * invN * mask * (p1 - r[1]);
  g[2] = wpen * (sm3[0] - (label == 0 ? 1.f : 0.f));
  g[3] = wpen * (sm3[1] - (label == 1 ? 1.f : 0.f));
  g[4] = wpen * (sm3[2] - (label == 2 ? 1.f : 0.f));
  for (int j = 5; j < (int)lddp; j++) g[j] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// residual add + LayerNormalization(eps=1e-6)   (transformer.py:217-218, 222, 250-260)
// one warp per row; the row lives in registers (d <= 1024)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 8;  // float4 per lane -> d <= 1024 (kernels are instantiated for 1, 2, 4, 8)

template <int LN_NV>
__global__ void add_ln_fwd_k(const float* __restrict__ a, const float* __restrict__ b,
                             const float* __restrict__ bias, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ s_out,
                             float* __restrict__ y, float* __restrict__ mean_out,
                             float* __restrict__ rstd_out, int rows, int d, int rnd, DropSpec dr) {
  pdl_enter();
  int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= rows) return;
  int lane = threadIdx.x;
  int nv = d / 4;
  float4 v[LN_NV];
  float sum = 0.f;
  const uint32_t dkey = dr.thresh ? skf_drop_key(dr) : 0u;
#pragma unroll
  for (int k = 0; k < LN_NV; k++) {
    int c = lane + k * 32;
    if (c < nv) {
      float4 x = reinterpret_cast<const float4*>(a + (int64_t)row * d)[c];
      if (b) {
        float4 z = reinterpret_cast<const float4*>(b + (int64_t)row * d)[c];
        if (bias) {
          float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c);
          z.x += bb.x; z.y += bb.y; z.z += bb.z; z.w += bb.w;
        }
        if (dr.thresh) {          // Dropout on the sublayer output, before the residual add
          const uint32_t e0 = (uint32_t)row * (uint32_t)d + 4u * c;
          z.x = skf_drop_apply(z.x, dkey, dr.site, e0, dr.thresh, dr.scale);
          z.y = skf_drop_apply(z.y, dkey, dr.site, e0 + 1, dr.thresh, dr.scale);
          z.z = skf_drop_apply(z.z, dkey, dr.site, e0 + 2, dr.thresh, dr.scale);
          z.w = skf_drop_apply(z.w, dkey, dr.site, e0 + 3, dr.thresh, dr.scale);
        }
        x.x += z.x; x.y += z.y; x.z += z.z; x.w += z.w;
      }
      v[k] = x;
      sum += (x.x + x.y) + (x.z + x.w);
    }
  }
  sum = warp_sum(sum);
  float mean = sum / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < LN_NV; k++) {
    int c = lane + k * 32;
    if (c < nv) {
      float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  sq = warp_sum(sq);
  float rstd = 1.0f / sqrtf(sq / (float)d + 1e-6f);
#pragma unroll
  for (int k = 0; k < LN_NV; k++) {
    int c = lane + k * 32;
    if (c < nv) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = maybe_round((v[k].x - mean) * rstd * g.x + be.x, rnd);
      o.y = maybe_round((v[k].y - mean) * rstd * g.y + be.y, rnd);
      o.z = maybe_round((v[k].z - mean) * rstd * g.z + be.z, rnd);
      o.w = maybe_round((v[k].w - mean) * rstd * g.w + be.w, rnd);
      reinterpret_cast<float4*>(y + (int64_t)row * d)[c] = o;
      if (s_out) reinterpret_cast<float4*>(s_out + (int64_t)row * d)[c] = v[k];
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// LayerNorm backward.  Each block walks rows blockIdx.x, +gridDim.x, ...; one warp per row;
// dgamma/dbeta partials stay in registers across the walk, then smem -> one atomic per column.
template <int LN_NV, int R>
__global__ void __launch_bounds__(256) ln_bwd_k(const float* __restrict__ dy, const float* __restrict__ s,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, float* __restrict__ ds,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int d,
                         int rnd, float* __restrict__ ds_drop, DropSpec dr) {
  pdl_enter();
  extern __shared__ float sm[];  // [2][d]
  const uint32_t dkey = (ds_drop && dr.thresh) ? skf_drop_key(dr) : 0u;
  int lane = threadIdx.x, w = threadIdx.y, nw = blockDim.y;
  int nv = d / 4;
  float4 ag[LN_NV], ab[LN_NV];
#pragma unroll
  for (int k = 0; k < LN_NV; k++) ag[k] = ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = threadIdx.y * 32 + lane; c < 2 * d; c += nw * 32) sm[c] = 0.f;
  __syncthreads();
  // R rows per warp and iteration: their loads are issued together, so an SM keeps 16 warps x R x 2 rows of d
  // floats in flight (R = 2: 64 KB at d = 256; one row per warp left the kernel at half the HBM rate)
  const int rstride = gridDim.x * nw;
  for (int row0 = blockIdx.x * nw + w; row0 < rows; row0 += R * rstride) {
    float4 g[R][LN_NV], xh[R][LN_NV];
    float mu[R], rs[R], c1[R], c2[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int row = row0 + r * rstride;
      const bool live = row < rows;
      mu[r] = live ? mean[row] : 0.f;
      rs[r] = live ? rstd[row] : 0.f;
#pragma unroll
      for (int k = 0; k < LN_NV; k++) {
        const int c = lane + k * 32;
        g[r][k] = xh[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nv && live) {
          g[r][k] = reinterpret_cast<const float4*>(dy + (int64_t)row * d)[c];     // dy for now
          xh[r][k] = reinterpret_cast<const float4*>(s + (int64_t)row * d)[c];     // s for now
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      c1[r] = c2[r] = 0.f;
#pragma unroll
      for (int k = 0; k < LN_NV; k++) {
        const int c = lane + k * 32;
        if (c < nv) {
          const float4 dyv = g[r][k], sv = xh[r][k];
          const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c);
          xh[r][k] = make_float4((sv.x - mu[r]) * rs[r], (sv.y - mu[r]) * rs[r], (sv.z - mu[r]) * rs[r],
                                 (sv.w - mu[r]) * rs[r]);
          g[r][k] = make_float4(dyv.x * ga.x, dyv.y * ga.y, dyv.z * ga.z, dyv.w * ga.w);
          c1[r] += (g[r][k].x + g[r][k].y) + (g[r][k].z + g[r][k].w);
          c2[r] += (g[r][k].x * xh[r][k].x + g[r][k].y * xh[r][k].y) + (g[r][k].z * xh[r][k].z + g[r][k].w * xh[r][k].w);
          ag[k].x += dyv.x * xh[r][k].x; ag[k].y += dyv.y * xh[r][k].y;
          ag[k].z += dyv.z * xh[r][k].z; ag[k].w += dyv.w * xh[r][k].w;
          ab[k].x += dyv.x; ab[k].y += dyv.y; ab[k].z += dyv.z; ab[k].w += dyv.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      c1[r] = warp_sum(c1[r]) / (float)d;
      c2[r] = warp_sum(c2[r]) / (float)d;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int row = row0 + r * rstride;
      if (row >= rows) continue;
#pragma unroll
      for (int k = 0; k < LN_NV; k++) {
        int c = lane + k * 32;
        if (c < nv) {
          float4 o;
          o.x = maybe_round(rs[r] * (g[r][k].x - c1[r] - xh[r][k].x * c2[r]), rnd);
          o.y = maybe_round(rs[r] * (g[r][k].y - c1[r] - xh[r][k].y * c2[r]), rnd);
          o.z = maybe_round(rs[r] * (g[r][k].z - c1[r] - xh[r][k].z * c2[r]), rnd);
          o.w = maybe_round(rs[r] * (g[r][k].w - c1[r] - xh[r][k].w * c2[r]), rnd);
          reinterpret_cast<float4*>(ds + (int64_t)row * d)[c] = o;
          if (ds_drop) {
            if (dr.thresh) {
              const uint32_t e0 = (uint32_t)row * (uint32_t)d + 4u * c;
              o.x = skf_drop_apply(o.x, dkey, dr.site, e0, dr.thresh, dr.scale);
              o.y = skf_drop_apply(o.y, dkey, dr.site, e0 + 1, dr.thresh, dr.scale);
              o.z = skf_drop_apply(o.z, dkey, dr.site, e0 + 2, dr.thresh, dr.scale);
              o.w = skf_drop_apply(o.w, dkey, dr.site, e0 + 3, dr.thresh, dr.scale);
            }
            reinterpret_cast<float4*>(ds_drop + (int64_t)row * d)[c] = o;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LN_NV; k++) {
    int c = lane + k * 32;
    if (c < nv) {
      atomicAdd(&sm[c * 4 + 0], ag[k].x); atomicAdd(&sm[c * 4 + 1], ag[k].y);
      atomicAdd(&sm[c * 4 + 2], ag[k].z); atomicAdd(&sm[c * 4 + 3], ag[k].w);
      atomicAdd(&sm[d + c * 4 + 0], ab[k].x); atomicAdd(&sm[d + c * 4 + 1], ab[k].y);
      atomicAdd(&sm[d + c * 4 + 2], ab[k].z); atomicAdd(&sm[d + c * 4 + 3], ab[k].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.y * 32 + lane; c < d; c += nw * 32) {
    atomicAdd(dgamma + c, sm[c]);
    atomicAdd(dbeta + c, sm[d + c]);
  }
}

// y[i] = keep(i) ? x[i] / (1 - rate) : 0   (tf.keras.layers.Dropout, training=True)
__global__ void dropout_k(const float* x, float* y, int64_t n, DropSpec dr) {
  const uint32_t key = skf_drop_key(dr);
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    v.x = skf_drop_apply(v.x, key, dr.site, (uint32_t)i, dr.thresh, dr.scale);
    v.y = skf_drop_apply(v.y, key, dr.site, (uint32_t)i + 1, dr.thresh, dr.scale);
    v.z = skf_drop_apply(v.z, key, dr.site, (uint32_t)i + 2, dr.thresh, dr.scale);
    v.w = skf_drop_apply(v.w, key, dr.site, (uint32_t)i + 3, dr.thresh, dr.scale);
    *reinterpret_cast<float4*>(y + i) = v;
  } else {
    for (; i < n; i++) y[i] = skf_drop_apply(x[i], key, dr.site, (uint32_t)i, dr.thresh, dr.scale);
  }
}

// out[n] += sum_m x[m,n]   (bias gradients).  Each block owns 128 consecutive columns (one float4 per
// lane) and a strided set of rows; 8 warps x 4 rows in flight per warp keep ~16 KB of loads outstanding
// per block; partial sums go through shared memory to one atomic per column per block.
__global__ void __launch_bounds__(256) colsum_k(const float* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                int M, int N) {
  pdl_enter();
  __shared__ float4 sm[8][32];
  const int lane = threadIdx.x, w = threadIdx.y;
  const int col = blockIdx.x * 128 + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    const bool vec = (col + 3 < N) && ((ldx & 3) == 0);
    const int stride = gridDim.y * 8;
    int m = blockIdx.y * 8 + w;
    if (vec) {
      for (; m + 3 * stride < M; m += 4 * stride) {
        const float4 a0 = *reinterpret_cast<const float4*>(x + (int64_t)m * ldx + col);
        const float4 a1 = *reinterpret_cast<const float4*>(x + (int64_t)(m + stride) * ldx + col);
        const float4 a2 = *reinterpret_cast<const float4*>(x + (int64_t)(m + 2 * stride) * ldx + col);
        const float4 a3 = *reinterpret_cast<const float4*>(x + (int64_t)(m + 3 * stride) * ldx + col);
        acc.x += (a0.x + a1.x) + (a2.x + a3.x); acc.y += (a0.y + a1.y) + (a2.y + a3.y);
        acc.z += (a0.z + a1.z) + (a2.z + a3.z); acc.w += (a0.w + a1.w) + (a2.w + a3.w);
      }
      for (; m < M; m += stride) {
        const float4 a0 = *reinterpret_cast<const float4*>(x + (int64_t)m * ldx + col);
        acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
      }
    } else {
      for (; m < M; m += stride) {
        const float* r = x + (int64_t)m * ldx + col;
        acc.x += r[0];
        if (col + 1 < N) acc.y += r[1];
        if (col + 2 < N) acc.z += r[2];
        if (col + 3 < N) acc.w += r[3];
      }
    }
  }
  sm[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < N) {
    float4 t = sm[0][lane];
#pragma unroll
    for (int i = 1; i < 8; i++) { t.x += sm[i][lane].x; t.y += sm[i][lane].y; t.z += sm[i][lane].z; t.w += sm[i][lane].w; }
    atomicAdd(out + col, t.x);
    if (col + 1 < N) atomicAdd(out + col + 1, t.y);
    if (col + 2 < N) atomicAdd(out + col + 2, t.z);
    if (col + 3 < N) atomicAdd(out + col + 3, t.w);
  }
}

// ---------------------------------------------------------------------------------------------
// K6: pooling bottleneck SelfAttnV1 (transformer.py:61-74).  One CTA per sample.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; i++) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void pool_fwd_k(const float* __restrict__ x, const float* __restrict__ u,
                           const float* __restrict__ vattn, float* __restrict__ a_out,
                           float* __restrict__ emb, int T, int d, int U, int rnd) {
  extern __shared__ float sm[];  // sc[T] + red[32]
  float* sc = sm;
  float* red = sm + T;
  int b = blockIdx.x;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = w; t < T; t += nw) {
    const float* ur = u + ((int64_t)b * T + t) * U;
    float acc = 0.f;
    for (int j = lane; j < U; j += 32) acc += ur[j] * __ldg(vattn + j);
    acc = warp_sum(acc);
    if (lane == 0) sc[t] = acc;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) mx = fmaxf(mx, sc[t]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float e = expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  sum = block_reduce(sum, red, false);
  float inv = 1.0f / sum;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float p = sc[t] * inv;
    sc[t] = p;
    a_out[(int64_t)b * T + t] = p;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    const float* xb = x + (int64_t)b * T * d + c;
    for (int t = 0; t < T; t++) acc += sc[t] * xb[(int64_t)t * d];
    emb[(int64_t)b * d + c] = maybe_round(acc, rnd);
  }
}

__global__ void pool_bwd_k(const float* __restrict__ x, const float* __restrict__ u,
                           const float* __restrict__ vattn, const float* __restrict__ a,
                           const float* __restrict__ demb, float* __restrict__ dx, int dx_acc,
                           float* __restrict__ dupre, float* __restrict__ dvattn, int T, int d,
                           int U, int rnd) {
  extern __shared__ float sm[];  // da[T] + red[32]
  float* da = sm;
  float* red = sm + T;
  int b = blockIdx.x;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* de = demb + (int64_t)b * d;
  // da_t = demb . x_t
  for (int t = w; t < T; t += nw) {
    const float* xr = x + ((int64_t)b * T + t) * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += xr[c] * de[c];
    acc = warp_sum(acc);
    if (lane == 0) da[t] = acc;
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) dot += a[(int64_t)b * T + t] * da[t];
  dot = block_reduce(dot, red, false);
  __syncthreads();
  // dscore_t = a_t (da_t - dot)   (softmax over t)
  for (int t = threadIdx.x; t < T; t += blockDim.x) da[t] = a[(int64_t)b * T + t] * (da[t] - dot);
  __syncthreads();
  // dx[b,t,c] (+)= a_t * demb_c
  for (int t = w; t < T; t += nw) {
    float at = a[(int64_t)b * T + t];
    float* dr = dx + ((int64_t)b * T + t) * d;
    for (int c = lane; c < d; c += 32) {
      float v = at * de[c];
      dr[c] = dx_acc ? dr[c] + v : v;
    }
  }
  // dupre[b,t,j] = dscore_t * V_j * (1 - u^2);  dV_j += sum_t u[b,t,j] * dscore_t
  for (int j = threadIdx.x; j < U; j += blockDim.x) {
    float vj = __ldg(vattn + j), acc = 0.f;
    for (int t = 0; t < T; t++) {
      int64_t o = ((int64_t)b * T + t) * U + j;
      float uu = u[o];
      acc += uu * da[t];
      dupre[o] = maybe_round(da[t] * vj * (1.0f - uu * uu), rnd);
    }
    atomicAdd(dvattn + j, acc);
  }
}

// Cluster form of the two pooling kernels: one CTA per sample leaves 128 CTAs of serial T-loops on 148 SMs (68 us
// forward / 131 us backward at cfg2, for 26 MB / 131 MB of traffic).  A thread-block cluster of POOL_CL CTAs shares a
// sample: CTA r owns rows [r * ceil(T / POOL_CL), ...), the softmax statistics / the a.da dot product are exchanged
// through distributed shared memory (every CTA stores its partial into slot r of every peer, one cluster barrier),
// and the T-reductions (pooled vector, dV) are summed by the cluster's CTA 0 from partials stored into its smem.
constexpr int POOL_CL = 8;
__device__ __forceinline__ uint32_t pool_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pool_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the address of local shared variable `p` in CTA `rank` of the cluster
__device__ __forceinline__ void pool_st_remote(float* p, uint32_t rank, float v) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// block-wide reduction whose result every thread gets; `red` holds one float per warp
__device__ __forceinline__ float pool_block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; i++) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __cluster_dims__(POOL_CL, 1, 1) __launch_bounds__(256)
pool_fwd_cl_k(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ vattn,
              float* __restrict__ a_out, float* __restrict__ emb, int T, int d, int U, int rnd) {
  extern __shared__ float sm[];   // sc[chunk] | red[32] | xmax[POOL_CL] | xsum[POOL_CL] | gather[POOL_CL * d]
  const int chunk = (T + POOL_CL - 1) / POOL_CL;
  float* sc = sm;
  float* red = sc + chunk;
  float* xmax = red + 32;
  float* xsum = xmax + POOL_CL;
  float* gather = xsum + POOL_CL;
  const uint32_t r = pool_cta_rank();
  const int b = blockIdx.x / POOL_CL;
  const int t0 = min(T, (int)r * chunk), n = min(T, t0 + chunk) - t0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // every CTA of the cluster has started (its shared memory exists) before anyone stores into a peer's: the
  // programming model requires it, and compute-sanitizer flags the remote stores below without it
  pool_cluster_sync();
  // scores of this CTA's rows: warp per row
  for (int i = w; i < n; i += nw) {
    const float* ur = u + ((int64_t)b * T + t0 + i) * U;
    float acc = 0.f;
    for (int j = lane; j < U; j += 32) acc += ur[j] * __ldg(vattn + j);
    acc = warp_sum(acc);
    if (lane == 0) sc[i] = acc;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, sc[i]);
  mx = pool_block_reduce(mx, red, true);
  if (threadIdx.x < POOL_CL) pool_st_remote(xmax + r, threadIdx.x, mx);
  pool_cluster_sync();
  mx = xmax[0];
#pragma unroll
  for (int i = 1; i < POOL_CL; i++) mx = fmaxf(mx, xmax[i]);
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = expf(sc[i] - mx);
    sc[i] = e;
    sum += e;
  }
  sum = pool_block_reduce(sum, red, false);
  if (threadIdx.x < POOL_CL) pool_st_remote(xsum + r, threadIdx.x, sum);
  pool_cluster_sync();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < POOL_CL; i++) sum += xsum[i];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float pa = sc[i] * inv;
    sc[i] = pa;
    a_out[(int64_t)b * T + t0 + i] = pa;
  }
  __syncthreads();
  // partial pooled vector over this CTA's rows -> slot r of CTA 0
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float* xb = x + ((int64_t)b * T + t0) * d + c;
    float acc = 0.f;
#pragma unroll 5
    for (int i = 0; i < n; i++) acc += sc[i] * xb[(int64_t)i * d];
    pool_st_remote(gather + r * d + c, 0, acc);
  }
  pool_cluster_sync();
  if (r == 0) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < POOL_CL; i++) acc += gather[i * d + c];
      emb[(int64_t)b * d + c] = maybe_round(acc, rnd);
    }
  }
}

__global__ void __cluster_dims__(POOL_CL, 1, 1) __launch_bounds__(256)
pool_bwd_cl_k(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ vattn,
              const float* __restrict__ a, const float* __restrict__ demb, float* __restrict__ dx, int dx_acc,
              float* __restrict__ dupre, float* __restrict__ dvattn, int T, int d, int U, int rnd) {
  extern __shared__ float sm[];   // da[chunk] | at[chunk] | red[32] | xdot[POOL_CL] | gather[POOL_CL * U]
  const int chunk = (T + POOL_CL - 1) / POOL_CL;
  float* da = sm;
  float* at = da + chunk;
  float* red = at + chunk;
  float* xdot = red + 32;
  float* gather = xdot + POOL_CL;
  const uint32_t r = pool_cta_rank();
  const int b = blockIdx.x / POOL_CL;
  const int t0 = min(T, (int)r * chunk), n = min(T, t0 + chunk) - t0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // every CTA of the cluster has started (its shared memory exists) before anyone stores into a peer's: the
  // programming model requires it, and compute-sanitizer flags the remote stores below without it
  pool_cluster_sync();
  const float* de = demb + (int64_t)b * d;
  // da_t = demb . x_t for this CTA's rows
  for (int i = w; i < n; i += nw) {
    const float* xr = x + ((int64_t)b * T + t0 + i) * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += xr[c] * de[c];
    acc = warp_sum(acc);
    if (lane == 0) {
      da[i] = acc;
      at[i] = a[(int64_t)b * T + t0 + i];
    }
  }
  __syncthreads();
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dot += at[i] * da[i];
  dot = pool_block_reduce(dot, red, false);
  if (threadIdx.x < POOL_CL) pool_st_remote(xdot + r, threadIdx.x, dot);
  pool_cluster_sync();
  dot = 0.f;
#pragma unroll
  for (int i = 0; i < POOL_CL; i++) dot += xdot[i];
  // dscore_t = a_t (da_t - dot)   (softmax over t)
  for (int i = threadIdx.x; i < n; i += blockDim.x) da[i] = at[i] * (da[i] - dot);
  __syncthreads();
  // dx[b,t,c] (+)= a_t * demb_c
  for (int i = w; i < n; i += nw) {
    const float av = at[i];
    float* dr = dx + ((int64_t)b * T + t0 + i) * d;
    for (int c = lane; c < d; c += 32) {
      const float v = av * de[c];
      dr[c] = dx_acc ? dr[c] + v : v;
    }
  }
  // dupre[b,t,j] = dscore_t * V_j * (1 - u^2);  dV_j += sum_t u[b,t,j] * dscore_t (partials summed by CTA 0)
  for (int j = threadIdx.x; j < U; j += blockDim.x) {
    const float vj = __ldg(vattn + j);
    float acc = 0.f;
#pragma unroll 5
    for (int i = 0; i < n; i++) {
      const int64_t o = ((int64_t)b * T + t0 + i) * U + j;
      const float uu = u[o];
      acc += uu * da[i];
      dupre[o] = maybe_round(da[i] * vj * (1.0f - uu * uu), rnd);
    }
    pool_st_remote(gather + r * U + j, 0, acc);
  }
  pool_cluster_sync();
  if (r == 0) {
    for (int j = threadIdx.x; j < U; j += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < POOL_CL; i++) acc += gather[i * U + j];
      atomicAdd(dvattn + j, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K8: DenseExpander (transformer.py:370-376): pre[b,t,c] = emb[b,c]*w[t] + beta[t]
// ---------------------------------------------------------------------------------------------
__global__ void expand_fwd_k(const float* __restrict__ emb, const float* __restrict__ w,
                             const float* __restrict__ beta, float* __restrict__ pre, int B,
                             int T, int d, int rnd) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*T*d/4
  int64_t total = (int64_t)B * T * (d / 4);
  if (idx >= total) return;
  int c4 = (int)(idx % (d / 4));
  int64_t bt = idx / (d / 4);
  int t = (int)(bt % T);
  int b = (int)(bt / T);
  float4 e = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)b * d) + c4);
  float wt = __ldg(w + t), be = __ldg(beta + t);
  float4 o;
  o.x = maybe_round(__fadd_rn(__fmul_rn(e.x, wt), be), rnd);
  o.y = maybe_round(__fadd_rn(__fmul_rn(e.y, wt), be), rnd);
  o.z = maybe_round(__fadd_rn(__fmul_rn(e.z, wt), be), rnd);
  o.w = maybe_round(__fadd_rn(__fmul_rn(e.w, wt), be), rnd);
  reinterpret_cast<float4*>(pre)[idx] = o;
}

// demb[b,c] (+)= sum_t dpre[b,t,c] w[t]        (grid = B, thread per c)
__global__ void expand_bwd_emb_k(const float* __restrict__ dpre, const float* __restrict__ w,
                                 float* __restrict__ demb, int acc, int T, int d) {
  int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    const float* p = dpre + (int64_t)b * T * d + c;
    for (int t = 0; t < T; t++) s += p[(int64_t)t * d] * __ldg(w + t);
    float* o = demb + (int64_t)b * d + c;
    *o = acc ? *o + s : s;
  }
}
// dw[t] += sum_{b,c} dpre[b,t,c] emb[b,c];  dbeta[t] += sum_{b,c} dpre[b,t,c]   (grid = T)
__global__ void expand_bwd_w_k(const float* __restrict__ dpre, const float* __restrict__ emb,
                               float* __restrict__ dw, float* __restrict__ dbeta, int B, int T,
                               int d) {
  __shared__ float red[32];
  int t = blockIdx.x;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < B * d; i += blockDim.x) {
    int b = i / d, c = i - b * d;
    float g = dpre[((int64_t)b * T + t) * d + c];
    s1 += g * emb[i];
    s2 += g;
  }
  s1 = block_reduce(s1, red, false);
  s2 = block_reduce(s2, red, false);
  if (threadIdx.x == 0) {
    atomicAdd(dw + t, s1);
    atomicAdd(dbeta + t, s2);
  }
}

// ---------------------------------------------------------------------------------------------
// K7/K9: softmax cross-entropy heads (builders/losses.py:21-41; sketchformer.py:335-344).
// One warp per row.  argmax returns the lowest index among ties (tf.argmax).
// ---------------------------------------------------------------------------------------------
__global__ void softmax_ce_k(const float* __restrict__ logits, int64_t ldl,
                             const int64_t* __restrict__ tgt, int64_t tgt_ld, int rows, int Tc,
                             int ncols, int ignore_zero, float scale, float* dlogits,
                             int64_t lddl, float* __restrict__ probs, int64_t ldp,
                             int32_t* __restrict__ argmax_out, double* __restrict__ stats,
                             int rnd) {
  pdl_enter();
  __shared__ float s_loss[8];
  __shared__ int s_corr[8];
  int lane = threadIdx.x, w = threadIdx.y;
  int row = blockIdx.x * blockDim.y + w;
  float loss = 0.f;
  int corr = 0;
  if (row < rows) {
    const float* z = logits + (int64_t)row * ldl;
    int b = row / Tc, t = row - b * Tc;
    int64_t y = tgt ? tgt[(int64_t)b * tgt_ld + t] : -1;
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int j = lane; j < ncols; j += 32) {
      float v = z[j];
      if (v > mx) { mx = v; am = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    float se = 0.f;
    for (int j = lane; j < ncols; j += 32) se += expf(z[j] - mx);
    se = warp_sum(se);
    float lse = mx + logf(se);
    float inv = 1.0f / se;
    float mrow = (ignore_zero && y == 0) ? 0.f : 1.f;
    if (tgt) {
      float zt = (y >= 0 && y < ncols) ? z[y] : 0.f;
      loss = (lse - zt) * mrow;
      corr = (am == (int)y) ? 1 : 0;
    }
    if (argmax_out && lane == 0) argmax_out[row] = am;
    if (probs)
      for (int j = lane; j < ncols; j += 32) probs[(int64_t)row * ldp + j] = expf(z[j] - mx) * inv;
    if (dlogits) {
      float sc = scale * mrow;
      for (int j = lane; j < ncols; j += 32) {
        float p = expf(z[j] - mx) * inv;
        float g = (p - (j == (int)y ? 1.f : 0.f)) * sc;
        dlogits[(int64_t)row * lddl + j] = maybe_round(g, rnd);
      }
    }
  }
  if (stats) {
    if (lane == 0) { s_loss[w] = loss; s_corr[w] = corr; }
    __syncthreads();
    if (w == 0 && lane == 0) {
      double l = 0.0, c = 0.0;
      for (int i = 0; i < blockDim.y; i++) { l += (double)s_loss[i]; c += (double)s_corr[i]; }
      atomicAdd(stats + 0, l);
      atomicAdd(stats + 1, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K11: TF-2.1 Adam + Noam schedule, flat arena, 128-bit accesses
// ---------------------------------------------------------------------------------------------
__global__ void adam_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, const int64_t* __restrict__ step,
                       float rsqrt_d, float warm_pow, float b1, float b2, float eps,
                       float gscale) {
  int64_t it = *step;                       // optimizer.iterations BEFORE the increment
  float sf = (float)it;
  float arg1 = 1.0f / sqrtf(sf);            // inf at step 0
  float arg2 = sf * warm_pow;
  float lr = rsqrt_d * fminf(arg1, arg2);   // lr(0) == 0
  double t = (double)(it + 1);
  float alpha = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4);
    float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
#define SKF_ADAM1(c)                                         \
  {                                                          \
    float gr = gg.c * gscale;                                \
    mm.c += (gr - mm.c) * omb1;                              \
    vv.c += (gr * gr - vv.c) * omb2;                         \
    pp.c -= (mm.c * alpha) / (sqrtf(vv.c) + eps);            \
  }
    SKF_ADAM1(x) SKF_ADAM1(y) SKF_ADAM1(z) SKF_ADAM1(w)
#undef SKF_ADAM1
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n; i++) {
      float gr = g[i] * gscale;
      float mi = m[i] + (gr - m[i]) * omb1;
      float vi = v[i] + (gr * gr - v[i]) * omb2;
      m[i] = mi;
      v[i] = vi;
      p[i] -= (mi * alpha) / (sqrtf(vi) + eps);
    }
  }
}
__global__ void step_inc_k(int64_t* step) { *step += 1; }

// ---------------------------------------------------------------------------------------------
// Cross-attention keys / values of the expanded bottleneck (models/sketchformer.py:170-181,
// transformer.py:370-376 then MultiHeadAttention.wk / wv): pre[b,t,:] = emb[b,:] w[t] + beta[t] has rank
// two in t, so  pre W + bias = w[t] (emb W)[b,:] + beta[t] colsum(W) + bias  needs a [B,E]x[E,N] product
// instead of a [B T,E]x[E,N] one, and its backward is a reduction over the gradient rows.
// ---------------------------------------------------------------------------------------------
__global__ void kv_expand_fwd_k(const float* __restrict__ U, const float* __restrict__ s, const float* __restrict__ bias,
                                const float* __restrict__ w, const float* __restrict__ beta, float* __restrict__ KV,
                                int T, int N, int rows_per_cta) {
  const int b = blockIdx.x, t0 = blockIdx.y * rows_per_cta;
  const int t1 = min(T, t0 + rows_per_cta);
  for (int j = threadIdx.x * 4; j < N; j += blockDim.x * 4) {
    const float4 u = *reinterpret_cast<const float4*>(U + (int64_t)b * N + j);
    const float4 sv = *reinterpret_cast<const float4*>(s + j);
    const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = t0; t < t1; t++) {
      const float wt = __ldg(w + t), bt = __ldg(beta + t);
      float4 o;
      o.x = fmaf(wt, u.x, fmaf(bt, sv.x, bv.x)); o.y = fmaf(wt, u.y, fmaf(bt, sv.y, bv.y));
      o.z = fmaf(wt, u.z, fmaf(bt, sv.z, bv.z)); o.w = fmaf(wt, u.w, fmaf(bt, sv.w, bv.w));
      *reinterpret_cast<float4*>(KV + ((int64_t)b * T + t) * N + j) = o;
    }
  }
}

// One CTA = one sample x one slice of its T rows; a warp walks rows, a lane owns 4 columns of every 128-column
// group (NV groups per 128 NV-column chunk).  Everything is accumulated with atomics into zeroed / running outputs.
template <int NV>
__global__ void __launch_bounds__(256)
kv_expand_bwd_k(const float* __restrict__ G, const float* __restrict__ U, const float* __restrict__ s,
                const float* __restrict__ w, const float* __restrict__ beta, float* __restrict__ gU,
                float* __restrict__ gs, float* __restrict__ gbias, float* __restrict__ gw, float* __restrict__ gbeta,
                int T, int N, int rows_per_cta) {
  __shared__ float red[3][128 * NV];
  const int b = blockIdx.x, t0 = blockIdx.y * rows_per_cta;
  const int t1 = min(T, t0 + rows_per_cta);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c0 = 0; c0 < N; c0 += 128 * NV) {
    float4 u[NV], sv[NV], aU[NV], aS[NV], aB[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int j = c0 + 128 * v + 4 * lane;
      u[v] = *reinterpret_cast<const float4*>(U + (int64_t)b * N + j);
      sv[v] = *reinterpret_cast<const float4*>(s + j);
      aU[v] = aS[v] = aB[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = t0 + warp; t < t1; t += nw) {
      const float wt = __ldg(w + t), bt = __ldg(beta + t);
      const float* row = G + ((int64_t)b * T + t) * N + c0 + 4 * lane;
      float du = 0.f, ds = 0.f;
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const float4 g = *reinterpret_cast<const float4*>(row + 128 * v);
        du += (g.x * u[v].x + g.y * u[v].y) + (g.z * u[v].z + g.w * u[v].w);
        ds += (g.x * sv[v].x + g.y * sv[v].y) + (g.z * sv[v].z + g.w * sv[v].w);
        aU[v].x = fmaf(wt, g.x, aU[v].x); aU[v].y = fmaf(wt, g.y, aU[v].y);
        aU[v].z = fmaf(wt, g.z, aU[v].z); aU[v].w = fmaf(wt, g.w, aU[v].w);
        aS[v].x = fmaf(bt, g.x, aS[v].x); aS[v].y = fmaf(bt, g.y, aS[v].y);
        aS[v].z = fmaf(bt, g.z, aS[v].z); aS[v].w = fmaf(bt, g.w, aS[v].w);
        aB[v].x += g.x; aB[v].y += g.y; aB[v].z += g.z; aB[v].w += g.w;
      }
      du = warp_sum(du);
      ds = warp_sum(ds);
      if (lane == 0) {
        atomicAdd(gw + t, du);
        atomicAdd(gbeta + t, ds);
      }
    }
    // column sums: the 8 warps meet in shared memory, then one global atomic per column and CTA
    for (int i = threadIdx.x; i < 3 * 128 * NV; i += blockDim.x) (&red[0][0])[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int j = 128 * v + 4 * lane;
      atomicAdd(&red[0][j], aU[v].x); atomicAdd(&red[0][j + 1], aU[v].y);
      atomicAdd(&red[0][j + 2], aU[v].z); atomicAdd(&red[0][j + 3], aU[v].w);
      atomicAdd(&red[1][j], aS[v].x); atomicAdd(&red[1][j + 1], aS[v].y);
      atomicAdd(&red[1][j + 2], aS[v].z); atomicAdd(&red[1][j + 3], aS[v].w);
      atomicAdd(&red[2][j], aB[v].x); atomicAdd(&red[2][j + 1], aB[v].y);
      atomicAdd(&red[2][j + 2], aB[v].z); atomicAdd(&red[2][j + 3], aB[v].w);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 128 * NV; j += blockDim.x) {
      atomicAdd(gU + (int64_t)b * N + c0 + j, red[0][j]);
      atomicAdd(gs + c0 + j, red[1][j]);
      if (gbias) atomicAdd(gbias + c0 + j, red[2][j]);
    }
    __syncthreads();
  }
}

__global__ void add_row_broadcast_k(float* __restrict__ Y, const float* __restrict__ r, int64_t rows, int N) {
  pdl_enter();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= rows * N) return;
  const int j = (int)(i % N);
  float4 y = *reinterpret_cast<float4*>(Y + i);
  const float4 v = *reinterpret_cast<const float4*>(r + j);
  y.x += v.x; y.y += v.y; y.z += v.z; y.w += v.w;
  *reinterpret_cast<float4*>(Y + i) = y;
}


// ---------------------------------------------------------------------------------------------
// Cross-attention over the expanded bottleneck without materialising keys or values.  With
// K[b,t,:] = w[t] UK[b,:] + beta[t] sK + bK and V likewise (see kv_expand above), a query's logits are
// l[t] = a w[t] + c beta[t] + e  with the three per-(query, head) scalars a = Qh.UKh / sqrt(dh), c = Qh.sKh / sqrt(dh),
// e = Qh.bKh / sqrt(dh), and its output is  m1 UVh + m2 sVh + bVh  with m1 = sum_t P[t] w[t], m2 = sum_t P[t] beta[t]:
// scaled_dot_product_attention (builders/utils.py:71-105) of MultiHeadAttention.call (transformer.py:167-191)
// reduced to T exponentials per query and head.  One CTA per (head, sample), one warp per query, lane j owns
// elements j, j+32 of the head; additive -1e9 key-padding mask as in the reference.
// ---------------------------------------------------------------------------------------------
constexpr float XF_LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float xf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Shared by the forward and backward kernels of one (head, sample): the key table, centred so that second moments
// do not cancel, tab[t] = (w[t] - wbar, beta[t] - bbar, -1e9 log2e pad[t], 0), its extrema (for the softmax shift)
// and the six head vectors UKh, sKh, bKh, UVh, sVh, bVh.
struct XfShared {
  float4* tab;     // [T]
  float* vec;      // [6][DH]
  float* stat;     // wbar, bbar, wmin', wmax', bmin', bmax'  (primes: centred)
};

template <int DH>
__device__ __forceinline__ void xf_setup(const XfShared& S, const float* U, const float* s, const float* bias,
                                         const float* w, const float* beta, const uint8_t* key_pad, int b, int h, int T,
                                         int d) {
  __shared__ float red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float sw = 0.f, sb = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) { sw += w[t]; sb += beta[t]; }
  sw = warp_sum(sw); sb = warp_sum(sb);
  if (lane == 0) { red[0][warp] = sw; red[1][warp] = sb; }
  __syncthreads();
  float wbar = 0.f, bbar = 0.f;
  for (int k = 0; k < nw; k++) { wbar += red[0][k]; bbar += red[1][k]; }
  wbar /= (float)T; bbar /= (float)T;
  float wmn = INFINITY, wmx = -INFINITY, bmn = INFINITY, bmx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float a = w[t] - wbar, c = beta[t] - bbar;
    S.tab[t] = make_float4(a, c, (key_pad && key_pad[(int64_t)b * T + t]) ? -1e9f * XF_LOG2E : 0.f, 0.f);
    wmn = fminf(wmn, a); wmx = fmaxf(wmx, a); bmn = fminf(bmn, c); bmx = fmaxf(bmx, c);
  }
  wmn = -warp_max(-wmn); wmx = warp_max(wmx); bmn = -warp_max(-bmn); bmx = warp_max(bmx);
  if (lane == 0) { red[2][warp] = wmn; red[3][warp] = wmx; red[4][warp] = bmn; red[5][warp] = bmx; }
  for (int j = threadIdx.x; j < DH; j += blockDim.x) {
    const int ck = h * DH + j, cv = d + h * DH + j;
    S.vec[j] = U[(int64_t)b * 2 * d + ck];
    S.vec[DH + j] = s[ck];
    S.vec[2 * DH + j] = bias ? bias[ck] : 0.f;
    S.vec[3 * DH + j] = U[(int64_t)b * 2 * d + cv];
    S.vec[4 * DH + j] = s[cv];
    S.vec[5 * DH + j] = bias ? bias[cv] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = red[2][0], c = red[3][0], e = red[4][0], f = red[5][0];
    for (int k = 1; k < nw; k++) {
      a = fminf(a, red[2][k]); c = fmaxf(c, red[3][k]); e = fminf(e, red[4][k]); f = fmaxf(f, red[5][k]);
    }
    S.stat[0] = wbar; S.stat[1] = bbar; S.stat[2] = a; S.stat[3] = c; S.stat[4] = e; S.stat[5] = f;
  }
  __syncthreads();
}

// One query against the T keys: p[t] = 2^(a2 w'[t] + c2 b'[t] + e2 + mask[t] - shift); returns the sums of
// p, p w', p b' (and p w'^2, p w' b', p b'^2 when MOM).  `shift` is an upper bound of the unmasked logits.
template <bool MOM>
__device__ __forceinline__ void xf_row(const float4* __restrict__ tab, int T, float a2, float c2, float e2s, float& sum,
                                       float& s1, float& s2, float& s11, float& s12, float& s22) {
  sum = s1 = s2 = s11 = s12 = s22 = 0.f;
#pragma unroll 4
  for (int t = 0; t < T; t++) {
    const float4 k = tab[t];
    const float p = xf_ex2(fmaf(a2, k.x, fmaf(c2, k.y, e2s)) + k.z);
    const float pw = p * k.x, pb = p * k.y;
    sum += p; s1 += pw; s2 += pb;
    if (MOM) { s11 = fmaf(pw, k.x, s11); s12 = fmaf(pw, k.y, s12); s22 = fmaf(pb, k.y, s22); }
  }
}
// Softmax statistics of one query.  The shift is the sum of the per-term maxima (no pass over the keys); if that
// bound is so loose that everything underflows - or every key is masked, where the reference's additive -1e9 makes
// the row uniform - the exact maximum is taken instead.
template <bool MOM>
__device__ __forceinline__ float xf_softmax(const XfShared& S, int T, float a2, float c2, float e2, float& sum, float& s1,
                                            float& s2, float& s11, float& s12, float& s22) {
  float shift = fmaxf(a2 * S.stat[2], a2 * S.stat[3]) + fmaxf(c2 * S.stat[4], c2 * S.stat[5]) + e2;
  xf_row<MOM>(S.tab, T, a2, c2, e2 - shift, sum, s1, s2, s11, s12, s22);
  if (!(sum > 1e-30f)) {
    float mx = -INFINITY;
    for (int t = 0; t < T; t++) {
      const float4 k = S.tab[t];
      mx = fmaxf(mx, fmaf(a2, k.x, fmaf(c2, k.y, e2)) + k.z);
    }
    // (x + mask) - mx in the order the reference adds them: a fully masked row collapses to p = 1 everywhere
    shift = mx;
    sum = s1 = s2 = s11 = s12 = s22 = 0.f;
    for (int t = 0; t < T; t++) {
      const float4 k = S.tab[t];
      const float p = xf_ex2((fmaf(a2, k.x, fmaf(c2, k.y, e2)) + k.z) - mx);
      const float pw = p * k.x, pb = p * k.y;
      sum += p; s1 += pw; s2 += pb;
      if (MOM) { s11 = fmaf(pw, k.x, s11); s12 = fmaf(pw, k.y, s12); s22 = fmaf(pb, k.y, s22); }
    }
  }
  return shift;
}

// forward: one thread per query
template <int DH>
__global__ void __launch_bounds__(128)
xattn_fold_fwd_k(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ U, const float* __restrict__ s,
                 const float* __restrict__ bias, const float* __restrict__ w, const float* __restrict__ beta,
                 const uint8_t* __restrict__ key_pad, float* __restrict__ O, int64_t ldo, int Tq, int T, int d,
                 float scale) {
  pdl_enter();
  extern __shared__ float4 sm4[];
  XfShared S;
  S.tab = sm4;
  S.vec = reinterpret_cast<float*>(sm4 + T);
  S.stat = S.vec + 6 * DH;
  const int h = blockIdx.x, b = blockIdx.y;
  xf_setup<DH>(S, U, s, bias, w, beta, key_pad, b, h, T, d);
  const int q = blockIdx.z * blockDim.x + threadIdx.x;
  if (q >= Tq) return;
  const float4* qr = reinterpret_cast<const float4*>(Q + ((int64_t)b * Tq + q) * ldq + h * DH);
  const float4* uk = reinterpret_cast<const float4*>(S.vec);
  const float4* sk = reinterpret_cast<const float4*>(S.vec + DH);
  const float4* bk = reinterpret_cast<const float4*>(S.vec + 2 * DH);
  float a = 0.f, c = 0.f, e = 0.f;
#pragma unroll
  for (int j = 0; j < DH / 4; j++) {
    const float4 x = qr[j], u = uk[j], v = sk[j], z = bk[j];
    a += (x.x * u.x + x.y * u.y) + (x.z * u.z + x.w * u.w);
    c += (x.x * v.x + x.y * v.y) + (x.z * v.z + x.w * v.w);
    e += (x.x * z.x + x.y * z.y) + (x.z * z.z + x.w * z.w);
  }
  const float k2 = scale * XF_LOG2E;
  const float a2 = a * k2, c2 = c * k2;
  const float e2 = fmaf(a2, S.stat[0], fmaf(c2, S.stat[1], e * k2));     // centring constants folded into the offset
  float sum, s1, s2, s11, s12, s22;
  xf_softmax<false>(S, T, a2, c2, e2, sum, s1, s2, s11, s12, s22);
  const float inv = 1.0f / sum, m1 = fmaf(s1, inv, S.stat[0]), m2 = fmaf(s2, inv, S.stat[1]);
  float4* orow = reinterpret_cast<float4*>(O + ((int64_t)b * Tq + q) * ldo + h * DH);
  const float4* uv = reinterpret_cast<const float4*>(S.vec + 3 * DH);
  const float4* sv = reinterpret_cast<const float4*>(S.vec + 4 * DH);
  const float4* bv = reinterpret_cast<const float4*>(S.vec + 5 * DH);
#pragma unroll
  for (int j = 0; j < DH / 4; j++) {
    const float4 u = uv[j], v = sv[j], z = bv[j];
    orow[j] = make_float4(fmaf(m1, u.x, fmaf(m2, v.x, z.x)), fmaf(m1, u.y, fmaf(m2, v.y, z.y)),
                          fmaf(m1, u.z, fmaf(m2, v.z, z.z)), fmaf(m1, u.w, fmaf(m2, v.w, z.w)));
  }
}

// backward for one (head, sample), queries in chunks of blockDim:
//   phase 1 (thread = query)  scalars a, c, dm1 = dO.UVh, dm2 = dO.sVh, softmax moments -> dQ, and a record per query
//   phase 2 (thread = key t)  gw[t], gbeta[t] += sum_q p_q[t] (dm + coef (dm1 w'[t] + dm2 b'[t] - K_q))
//   phase 3 (thread = column) sums over q of  da Q, dc Q, m1 dO, m2 dO, dO  -> gU (stored), gs, gbias (atomics)
// The key bias has zero gradient (softmax is shift-invariant).
template <int DH>
__global__ void __launch_bounds__(256)
xattn_fold_bwd_k(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ U, const float* __restrict__ s,
                 const float* __restrict__ bias, const float* __restrict__ w, const float* __restrict__ beta,
                 const uint8_t* __restrict__ key_pad, const float* __restrict__ dO, int64_t lddo,
                 float* __restrict__ dQ, int64_t lddq, float* __restrict__ gU, float* __restrict__ gs,
                 float* __restrict__ gbias, float* __restrict__ gw, float* __restrict__ gbeta, int Tq, int T, int d,
                 float scale) {
  pdl_enter();
  extern __shared__ float4 sm4[];
  XfShared S;
  S.tab = sm4;
  float4* rec = sm4 + T;                                   // [3][blockDim]
  const int Tp = (T + 3) & ~3;                             // keeps the float4 vectors below 16-byte aligned
  float* gws = reinterpret_cast<float*>(rec + 3 * blockDim.x);   // [Tp]
  float* gbs = gws + Tp;                                   // [Tp]
  float* red = gbs + Tp;                                   // [5][DH]
  S.vec = red + 5 * DH;
  S.stat = S.vec + 6 * DH;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
  for (int t = tid; t < T; t += nthr) gws[t] = gbs[t] = 0.f;
  for (int i = tid; i < 5 * DH; i += nthr) red[i] = 0.f;
  xf_setup<DH>(S, U, s, bias, w, beta, key_pad, b, h, T, d);
  const float k2 = scale * XF_LOG2E;
  for (int q0 = 0; q0 < Tq; q0 += nthr) {
    const int q = q0 + tid, nq = min(nthr, Tq - q0);
    // ---- phase 1
    if (q < Tq) {
      const float4* qr = reinterpret_cast<const float4*>(Q + ((int64_t)b * Tq + q) * ldq + h * DH);
      const float4* gr = reinterpret_cast<const float4*>(dO + ((int64_t)b * Tq + q) * lddo + h * DH);
      const float4* v4 = reinterpret_cast<const float4*>(S.vec);
      float a = 0.f, c = 0.f, e = 0.f, dm1 = 0.f, dm2 = 0.f;
#pragma unroll
      for (int j = 0; j < DH / 4; j++) {
        const float4 x = qr[j], g = gr[j];
        const float4 u = v4[j], v = v4[DH / 4 + j], z = v4[2 * DH / 4 + j], uv = v4[3 * DH / 4 + j], sv = v4[4 * DH / 4 + j];
        a += (x.x * u.x + x.y * u.y) + (x.z * u.z + x.w * u.w);
        c += (x.x * v.x + x.y * v.y) + (x.z * v.z + x.w * v.w);
        e += (x.x * z.x + x.y * z.y) + (x.z * z.z + x.w * z.w);
        dm1 += (g.x * uv.x + g.y * uv.y) + (g.z * uv.z + g.w * uv.w);
        dm2 += (g.x * sv.x + g.y * sv.y) + (g.z * sv.z + g.w * sv.w);
      }
      const float a2 = a * k2, c2 = c * k2;
      const float e2 = fmaf(a2, S.stat[0], fmaf(c2, S.stat[1], e * k2));
      float sum, s1, s2, s11, s12, s22;
      const float shift = xf_softmax<true>(S, T, a2, c2, e2, sum, s1, s2, s11, s12, s22);
      const float inv = 1.0f / sum, m1 = s1 * inv, m2 = s2 * inv;          // centred means
      const float v11 = fmaf(s11, inv, -m1 * m1), v12 = fmaf(s12, inv, -m1 * m2), v22 = fmaf(s22, inv, -m2 * m2);
      const float da = fmaf(dm1, v11, dm2 * v12) * scale, dc = fmaf(dm1, v12, dm2 * v22) * scale;
      // logits carry 1/sqrt(dh): a_true = a scale, so d(logit)/dw[t] = a scale
      float4* dqr = reinterpret_cast<float4*>(dQ + ((int64_t)b * Tq + q) * lddq + h * DH);
#pragma unroll
      for (int j = 0; j < DH / 4; j++) {
        const float4 u = v4[j], v = v4[DH / 4 + j];
        dqr[j] = make_float4(fmaf(da, u.x, dc * v.x), fmaf(da, u.y, dc * v.y), fmaf(da, u.z, dc * v.z),
                             fmaf(da, u.w, dc * v.w));
      }
      rec[tid] = make_float4(a2, c2, e2 - shift + log2f(inv), dm1);
      rec[nthr + tid] = make_float4(dm2, fmaf(dm1, m1, dm2 * m2), a * scale, c * scale);
      rec[2 * nthr + tid] = make_float4(da, dc, m1 + S.stat[0], m2 + S.stat[1]);
    }
    __syncthreads();
    // ---- phase 2
    for (int t = tid; t < T; t += nthr) {
      const float4 k = S.tab[t];
      float accw = 0.f, accb = 0.f;
#pragma unroll 2
      for (int i = 0; i < nq; i++) {
        const float4 r0 = rec[i], r1 = rec[nthr + i];
        const float p = xf_ex2(fmaf(r0.x, k.x, fmaf(r0.y, k.y, r0.z)) + k.z);
        const float x = fmaf(r0.w, k.x, fmaf(r1.x, k.y, -r1.y));
        accw = fmaf(p, fmaf(r1.z, x, r0.w), accw);
        accb = fmaf(p, fmaf(r1.w, x, r1.x), accb);
      }
      gws[t] += accw;
      gbs[t] += accb;
    }
    // ---- phase 3: a warp takes every (nthr/32)-th query of the chunk, a lane the columns j, j + 32, ...
    {
      const int lane = tid & 31, wp = tid >> 5, nwp = nthr >> 5;
      const float* qb = Q + ((int64_t)b * Tq + q0) * ldq + h * DH;
      const float* gb = dO + ((int64_t)b * Tq + q0) * lddo + h * DH;
      for (int j = lane; j < DH; j += 32) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll 4
        for (int i = wp; i < nq; i += nwp) {
          const float4 r2 = rec[2 * nthr + i];
          const float x = qb[(int64_t)i * ldq + j], g = gb[(int64_t)i * lddo + j];
          a0 = fmaf(r2.x, x, a0); a1 = fmaf(r2.y, x, a1);
          a2 = fmaf(r2.z, g, a2); a3 = fmaf(r2.w, g, a3); a4 += g;
        }
        atomicAdd(red + j, a0); atomicAdd(red + DH + j, a1); atomicAdd(red + 2 * DH + j, a2);
        atomicAdd(red + 3 * DH + j, a3); atomicAdd(red + 4 * DH + j, a4);
      }
    }
    __syncthreads();
  }
  for (int j = tid; j < DH; j += nthr) {
    const int ck = h * DH + j, cv = d + h * DH + j;
    gU[(int64_t)b * 2 * d + ck] = red[j];
    gU[(int64_t)b * 2 * d + cv] = red[2 * DH + j];
    atomicAdd(gs + ck, red[DH + j]);
    atomicAdd(gs + cv, red[3 * DH + j]);
    if (gbias) atomicAdd(gbias + cv, red[4 * DH + j]);
  }
  for (int t = tid; t < T; t += nthr) {
    atomicAdd(gw + t, gws[t]);
    atomicAdd(gbeta + t, gbs[t]);
  }
}

}  // namespace skf

using namespace skf;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* skf_last_error(void) { return skf::err_buf(); }
int skf_version(void) { return 1; }
size_t skf_abi_sizeof(int which) {
  return which == 0 ? sizeof(skf_cfg) : which == 1 ? sizeof(skf_gemm_epilogue) : which == 2 ? sizeof(skf_param_entry) : 0;
}
int skf_device_ok(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

int skf_embed_fwd(const int64_t* tok, int64_t tok_ld, const float* table, const float* pe,
                  float* out, int B, int T, int d, int V, int round_tf32, void* stream) {
  SKF_REQUIRE(d % 4 == 0 && B > 0 && T > 0, SKF_ERR_BAD_SHAPE, "embed_fwd: d %% 4 != 0 or empty");
  dim3 blk(32, 8);
  launch_k(embed_fwd_k, dim3(cdiv((long)B * T, 8)), dim3(blk), 0, as_stream(stream), PDL_SIMT, tok, tok_ld, table, pe, out, B, T,
                                                                  d, V, sqrtf((float)d), round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_embed_bwd(const int64_t* tok, int64_t tok_ld, const float* dout, float* dtable, int B,
                  int T, int d, int V, void* stream) {
  SKF_REQUIRE(d % 4 == 0 && d <= 1024, SKF_ERR_BAD_SHAPE, "embed_bwd: d=%d unsupported", d);
  dim3 blk(32, 8);
  const int grid = cdiv((long)B * cdiv(T, EMB_RUN), 8);
  const float sc = sqrtf((float)d);
  cudaStream_t st = as_stream(stream);
  if (d <= 128) launch_k(embed_bwd_k<1>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, tok, tok_ld, dout, dtable, B, T, d, V, sc);
  else if (d <= 256) launch_k(embed_bwd_k<2>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, tok, tok_ld, dout, dtable, B, T, d, V, sc);
  else if (d <= 512) launch_k(embed_bwd_k<4>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, tok, tok_ld, dout, dtable, B, T, d, V, sc);
  else launch_k(embed_bwd_k<8>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, tok, tok_ld, dout, dtable, B, T, d, V, sc);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_positional_encoding(float* pe, int rows, int d, void* stream) {
  posenc_k<<<cdiv((long)rows * d, 256), 256, 0, as_stream(stream)>>>(pe, rows, d);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_pad_flags(const int64_t* tok, int64_t tok_ld, uint8_t* pad, int B, int T, void* stream) {
  pad_flags_k<<<cdiv((long)B * T, 256), 256, 0, as_stream(stream)>>>(tok, tok_ld, pad, B, T);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}


int skf_pad_flags_cont(const float* x, int64_t x_ld, uint8_t* pad, int B, int T, void* stream) {
  pad_flags_cont_k<<<cdiv((long)B * T, 256), 256, 0, as_stream(stream)>>>(x, x_ld, pad, B, T);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_embed_cont_fwd(const float* x, int64_t x_ld, const float* W, const float* bias, const float* pe,
                       float* out, int B, int T, int d, int round_tf32, void* stream) {
  SKF_REQUIRE(B > 0 && T > 0 && d > 0, SKF_ERR_BAD_SHAPE, "embed_cont_fwd: empty");
  dim3 blk(32, 8);
  embed_cont_fwd_k<<<cdiv((long)B * T, 8), blk, 0, as_stream(stream)>>>(x, x_ld, W, bias, pe, out, B, T, d,
                                                                       sqrtf((float)d), round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_embed_cont_bwd(const float* x, int64_t x_ld, const float* dout, float* dW, float* db, int B, int T,
                       int d, void* stream) {
  int gy = (int)((long)B * T / 64);
  gy = gy < 1 ? 1 : (gy > 296 ? 296 : gy);
  embed_cont_bwd_k<<<dim3(cdiv(d, 128), gy), 128, 0, as_stream(stream)>>>(x, x_ld, dout, dW, db, B, T, d,
                                                                         sqrtf((float)d));
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_cont_loss(const float* pred, int64_t ldp, const float* real, int64_t real_ld, int B, int Td,
                  float weight, float* dpred, int64_t lddp, double* stats, void* stream) {
  SKF_REQUIRE(pred && real && stats && B > 0 && Td > 0 && ldp >= 5, SKF_ERR_BAD_ARG, "cont_loss: bad argument");
  const int rows = B * Td;
  cont_loss_stats_k<<<cdiv(rows, 256), 256, 0, as_stream(stream)>>>(pred, ldp, real, real_ld, B, Td, stats);
  SKF_LAUNCH_CHECK();
  if (dpred) {
    SKF_REQUIRE(lddp >= 5, SKF_ERR_BAD_ARG, "cont_loss: lddp < 5");
    cont_loss_grad_k<<<cdiv(rows, 256), 256, 0, as_stream(stream)>>>(pred, ldp, real, real_ld, B, Td, stats, weight,
                                                                    dpred, lddp);
    SKF_LAUNCH_CHECK();
  }
  return SKF_OK;
}

static DropSpec make_dropout(float rate, uint32_t seed, uint32_t site, const int64_t* step) {
  DropSpec dr;
  dr.scale = 1.0f / (1.0f - rate);
  dr.thresh = (uint32_t)(rate * 16777216.0f);
  dr.seed = seed; dr.site = site; dr.step = step;
  return dr;
}
static int check_dropout(float rate, const int64_t* step) {
  SKF_REQUIRE(rate >= 0.f && rate < 1.f, SKF_ERR_BAD_ARG, "dropout rate %f outside [0, 1)", (double)rate);
  SKF_REQUIRE(rate == 0.f || step != nullptr, SKF_ERR_BAD_ARG, "dropout needs the device step counter");
  return SKF_OK;
}

int skf_dropout(const float* x, float* y, int64_t n, float rate, uint32_t seed, uint32_t site, const int64_t* step,
                void* stream) {
  SKF_REQUIRE(x && y && n > 0, SKF_ERR_BAD_ARG, "dropout: bad argument");
  SKF_TRY(check_dropout(rate, step));
  SKF_REQUIRE(rate > 0.f, SKF_ERR_BAD_ARG, "dropout: rate must be > 0 (skip the call otherwise)");
  SKF_REQUIRE(n < (int64_t)1 << 32, SKF_ERR_BAD_SHAPE, "dropout: more than 2^32 elements per site");
  dropout_k<<<cdiv(cdiv(n, 4), 256), 256, 0, as_stream(stream)>>>(x, y, n, make_dropout(rate, seed, site, step));
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_add_dropout_layernorm_fwd(const float* a, const float* b, const float* bias, const float* gamma,
                                  const float* beta, float* s, float* y, float* mean, float* rstd, int rows, int d,
                                  float rate, uint32_t seed, uint32_t site, const int64_t* step, void* stream) {
  SKF_REQUIRE(d % 4 == 0 && d <= 128 * LN_MAX_VEC, SKF_ERR_BAD_SHAPE, "layernorm: d=%d unsupported", d);
  SKF_TRY(check_dropout(rate, step));
  SKF_REQUIRE(rate == 0.f || b != nullptr, SKF_ERR_BAD_ARG, "layernorm: dropout needs the sublayer operand b");
  dim3 blk(32, 8);
  const int grid = cdiv(rows, 8);
  cudaStream_t st = as_stream(stream);
  DropSpec dr = make_dropout(rate, seed, site, step);
  if (d <= 128) launch_k(add_ln_fwd_k<1>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, 0, dr);
  else if (d <= 256) launch_k(add_ln_fwd_k<2>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, 0, dr);
  else if (d <= 512) launch_k(add_ln_fwd_k<4>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, 0, dr);
  else launch_k(add_ln_fwd_k<8>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, 0, dr);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_add_layernorm_fwd(const float* a, const float* b, const float* bias, const float* gamma,
                          const float* beta, float* s, float* y, float* mean, float* rstd,
                          int rows, int d, int round_tf32, void* stream) {
  SKF_REQUIRE(d % 4 == 0 && d <= 128 * LN_MAX_VEC, SKF_ERR_BAD_SHAPE, "layernorm: d=%d unsupported", d);
  dim3 blk(32, 8);
  const int grid = cdiv(rows, 8);
  cudaStream_t st = as_stream(stream);
  DropSpec dr = make_dropout(0.f, 0, 0, nullptr);
  if (d <= 128) launch_k(add_ln_fwd_k<1>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, round_tf32, dr);
  else if (d <= 256) launch_k(add_ln_fwd_k<2>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, round_tf32, dr);
  else if (d <= 512) launch_k(add_ln_fwd_k<4>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, round_tf32, dr);
  else launch_k(add_ln_fwd_k<8>, dim3(grid), dim3(blk), 0, st, PDL_SIMT, a, b, bias, gamma, beta, s, y, mean, rstd, rows, d, round_tf32, dr);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int g_ln_bwd_rows = 2;   // rows per warp and iteration of the LayerNorm backward for d <= 256 (skf_set_option "ln_bwd_rows": 2 | 4)
int g_ln_bwd_bpsm = 2;   // blocks per SM of the LayerNorm backward grid (skf_set_option "ln_bwd_bpsm")
static int ln_bwd_launch(const float* dy, const float* s, const float* mean, const float* rstd, const float* gamma,
                         float* ds, float* ds_drop, float* dgamma, float* dbeta, int rows, int d, int round_tf32,
                         DropSpec dr, void* stream);

int skf_layernorm_dropout_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                              const float* gamma, float* ds, float* ds_drop, float* dgamma, float* dbeta, int rows,
                              int d, float rate, uint32_t seed, uint32_t site, const int64_t* step, void* stream) {
  SKF_TRY(check_dropout(rate, step));
  SKF_REQUIRE(ds_drop != nullptr, SKF_ERR_BAD_ARG, "layernorm_dropout_bwd: ds_drop is NULL");
  return ln_bwd_launch(dy, s, mean, rstd, gamma, ds, ds_drop, dgamma, dbeta, rows, d, 0,
                       make_dropout(rate, seed, site, step), stream);
}

int skf_layernorm_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                      const float* gamma, float* ds, float* dgamma, float* dbeta, int rows,
                      int d, int round_tf32, void* stream) {
  return ln_bwd_launch(dy, s, mean, rstd, gamma, ds, nullptr, dgamma, dbeta, rows, d, round_tf32,
                       make_dropout(0.f, 0, 0, nullptr), stream);
}

static int ln_bwd_launch(const float* dy, const float* s, const float* mean, const float* rstd, const float* gamma,
                         float* ds, float* ds_drop, float* dgamma, float* dbeta, int rows, int d, int round_tf32,
                         DropSpec dr, void* stream) {
  SKF_REQUIRE(d % 4 == 0 && d <= 128 * LN_MAX_VEC, SKF_ERR_BAD_SHAPE, "layernorm: d=%d unsupported", d);
  dim3 blk(32, 8);
  // every block ends with 2*d global atomics onto the same dgamma / dbeta addresses: two blocks per SM keep that
  // at ~300 adds per address (1184 blocks cost ~8 us of serialised atomics out of 28 us at cfg2)
  int grid = cdiv(rows, 8 * 4);              // >= 4 rows per warp so the dgamma/dbeta partials amortise
  if (grid > 148 * g_ln_bwd_bpsm) grid = 148 * g_ln_bwd_bpsm;
  if (grid < 1) grid = 1;
  const size_t sm = 2 * d * sizeof(float);
  cudaStream_t st = as_stream(stream);
  if (d <= 128) {
    if (g_ln_bwd_rows >= 4) launch_k(ln_bwd_k<1, 4>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
    else launch_k(ln_bwd_k<1, 2>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
  }
  else if (d <= 256) {
    if (g_ln_bwd_rows >= 4) launch_k(ln_bwd_k<2, 4>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
    else launch_k(ln_bwd_k<2, 2>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
  }
  else if (d <= 512) launch_k(ln_bwd_k<4, 1>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
  else launch_k(ln_bwd_k<8, 1>, dim3(grid), dim3(blk), sm, st, PDL_SIMT, dy, s, mean, rstd, gamma, ds, dgamma, dbeta, rows, d, round_tf32, ds_drop, dr);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_colsum(const float* x, int64_t ldx, float* out, int M, int N, void* stream) {
  dim3 blk(32, 8);
  const int gx = cdiv(N, 128);
  int gy = (148 * 4 + gx - 1) / gx;           // ~4 blocks per SM in total
  const int maxy = cdiv(M, 8 * 4);
  if (gy > maxy) gy = maxy;
  if (gy < 1) gy = 1;
  launch_k(colsum_k, dim3(dim3(gx, gy)), dim3(blk), 0, as_stream(stream), PDL_SIMT, x, ldx, out, M, N);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_pool_fwd(const float* x, const float* u, const float* vattn, float* a, float* emb, int B,
                 int T, int d, int U, int round_tf32, void* stream) {
  const int chunk = (T + POOL_CL - 1) / POOL_CL;
  const size_t smc = (size_t)(chunk + 32 + 2 * POOL_CL + POOL_CL * d) * sizeof(float);
  if (smc <= 48 * 1024) {
    pool_fwd_cl_k<<<B * POOL_CL, 256, smc, as_stream(stream)>>>(x, u, vattn, a, emb, T, d, U, round_tf32);
    SKF_LAUNCH_CHECK();
    return SKF_OK;
  }
  size_t sm = (T + 32) * sizeof(float);
  pool_fwd_k<<<B, 256, sm, as_stream(stream)>>>(x, u, vattn, a, emb, T, d, U, round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_pool_bwd(const float* x, const float* u, const float* vattn, const float* a,
                 const float* demb, float* dx, int dx_accumulate, float* dupre, float* dvattn,
                 int B, int T, int d, int U, int round_tf32, void* stream) {
  const int chunk = (T + POOL_CL - 1) / POOL_CL;
  const size_t smc = (size_t)(2 * chunk + 32 + POOL_CL + POOL_CL * U) * sizeof(float);
  if (smc <= 48 * 1024) {
    pool_bwd_cl_k<<<B * POOL_CL, 256, smc, as_stream(stream)>>>(x, u, vattn, a, demb, dx, dx_accumulate, dupre,
                                                                dvattn, T, d, U, round_tf32);
    SKF_LAUNCH_CHECK();
    return SKF_OK;
  }
  size_t sm = (T + 32) * sizeof(float);
  pool_bwd_k<<<B, 256, sm, as_stream(stream)>>>(x, u, vattn, a, demb, dx, dx_accumulate, dupre, dvattn,
                                               T, d, U, round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_expand_fwd(const float* emb, const float* w, const float* beta, float* pre, int B, int T,
                   int d, int round_tf32, void* stream) {
  SKF_REQUIRE(d % 4 == 0, SKF_ERR_BAD_SHAPE, "expand: d %% 4");
  int64_t total = (int64_t)B * T * (d / 4);
  expand_fwd_k<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(emb, w, beta, pre, B, T, d, round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_expand_bwd(const float* dpre, const float* emb, const float* w, float* demb,
                   int demb_accumulate, float* dw, float* dbeta, int B, int T, int d,
                   void* stream) {
  expand_bwd_emb_k<<<B, 256, 0, as_stream(stream)>>>(dpre, w, demb, demb_accumulate, T, d);
  SKF_LAUNCH_CHECK();
  expand_bwd_w_k<<<T, 256, 0, as_stream(stream)>>>(dpre, emb, dw, dbeta, B, T, d);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_kv_expand_fwd(const float* U, const float* s, const float* bias, const float* w, const float* beta,
                      float* KV, int B, int T, int N, void* stream) {
  SKF_REQUIRE(U && s && w && beta && KV && B > 0 && T > 0 && N > 0 && N % 4 == 0, SKF_ERR_BAD_ARG,
              "kv_expand_fwd: bad argument");
  const int rows = 16;
  kv_expand_fwd_k<<<dim3(B, cdiv(T, rows)), 128, 0, as_stream(stream)>>>(U, s, bias, w, beta, KV, T, N, rows);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_kv_expand_bwd(const float* G, const float* U, const float* s, const float* w, const float* beta,
                      float* gU, float* gs, float* gbias, float* gw, float* gbeta, int B, int T, int N,
                      void* stream) {
  SKF_REQUIRE(G && U && s && w && beta && gU && gs && gw && gbeta && B > 0 && T > 0 && N > 0 && N % 128 == 0,
              SKF_ERR_BAD_ARG, "kv_expand_bwd: bad argument (N must be a multiple of 128)");
  const int slices = B >= 2 * 148 ? 1 : (2 * 148 + B - 1) / B;      // ~2 CTAs per SM
  const int rows = cdiv(T, slices < 1 ? 1 : (slices > T ? T : slices));
  dim3 grid(B, cdiv(T, rows));
  cudaStream_t st = as_stream(stream);
  if (N % 512 == 0) kv_expand_bwd_k<4><<<grid, 256, 0, st>>>(G, U, s, w, beta, gU, gs, gbias, gw, gbeta, T, N, rows);
  else if (N % 256 == 0) kv_expand_bwd_k<2><<<grid, 256, 0, st>>>(G, U, s, w, beta, gU, gs, gbias, gw, gbeta, T, N, rows);
  else kv_expand_bwd_k<1><<<grid, 256, 0, st>>>(G, U, s, w, beta, gU, gs, gbias, gw, gbeta, T, N, rows);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

#define SKF_XFOLD_DISPATCH(K, GRID, BLK, SMEM, ST, ...)                             \
  do {                                                                              \
    if (dh == 16) launch_k(K<16>, GRID, dim3(BLK), SMEM, ST, PDL_SIMT, __VA_ARGS__);    \
    else if (dh == 32) launch_k(K<32>, GRID, dim3(BLK), SMEM, ST, PDL_SIMT, __VA_ARGS__); \
    else launch_k(K<64>, GRID, dim3(BLK), SMEM, ST, PDL_SIMT, __VA_ARGS__);             \
  } while (0)

int skf_xattn_fold_fwd(const float* Q, int64_t ldq, const float* U, const float* s, const float* bias, const float* w,
                       const float* beta, const uint8_t* key_pad, float* O, int64_t ldo, int B, int H, int Tq, int T,
                       int dh, void* stream) {
  SKF_REQUIRE(Q && U && s && w && beta && O && B > 0 && H > 0 && Tq > 0 && T > 0, SKF_ERR_BAD_ARG,
              "xattn_fold_fwd: bad argument");
  SKF_REQUIRE(dh == 16 || dh == 32 || dh == 64, SKF_ERR_UNSUPPORTED, "xattn_fold_fwd: head dim %d", dh);
  SKF_REQUIRE(((ldq | ldo) & 3) == 0 && ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(O)) & 15) == 0,
              SKF_ERR_BAD_ALIGN, "xattn_fold_fwd: Q / O must be 16-byte aligned with row strides that are multiples of 4");
  const size_t smem = (size_t)T * 16 + (6 * dh + 8) * sizeof(float);
  SKF_REQUIRE(smem <= 48 * 1024, SKF_ERR_UNSUPPORTED, "xattn_fold_fwd: T = %d too long", T);
  const float scale = 1.0f / sqrtf((float)dh);
  cudaStream_t st = as_stream(stream);
  SKF_XFOLD_DISPATCH(xattn_fold_fwd_k, dim3(H, B, cdiv(Tq, 128)), 128, smem, st, Q, ldq, U, s, bias, w, beta, key_pad, O, ldo,
                     Tq, T, H * dh, scale);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_xattn_fold_bwd(const float* Q, int64_t ldq, const float* U, const float* s, const float* bias, const float* w,
                       const float* beta, const uint8_t* key_pad, const float* dO, int64_t lddo, float* dQ,
                       int64_t lddq, float* gU, float* gs, float* gbias, float* gw, float* gbeta, int B, int H, int Tq,
                       int T, int dh, void* stream) {
  SKF_REQUIRE(Q && U && s && w && beta && dO && dQ && gU && gs && gw && gbeta && B > 0 && H > 0 && Tq > 0 && T > 0,
              SKF_ERR_BAD_ARG, "xattn_fold_bwd: bad argument");
  SKF_REQUIRE(dh == 16 || dh == 32 || dh == 64, SKF_ERR_UNSUPPORTED, "xattn_fold_bwd: head dim %d", dh);
  SKF_REQUIRE(((ldq | lddo | lddq) & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(dQ)) & 15) == 0,
              SKF_ERR_BAD_ALIGN, "xattn_fold_bwd: Q / dO / dQ must be 16-byte aligned with row strides that are multiples of 4");
  const int nthr = 256;
  const size_t smem = (size_t)T * 16 + (size_t)3 * nthr * 16 + ((size_t)2 * ((T + 3) & ~3) + 11 * dh + 8) * sizeof(float);
  SKF_REQUIRE(smem <= 48 * 1024, SKF_ERR_UNSUPPORTED, "xattn_fold_bwd: T = %d too long", T);
  const float scale = 1.0f / sqrtf((float)dh);
  cudaStream_t st = as_stream(stream);
  SKF_XFOLD_DISPATCH(xattn_fold_bwd_k, dim3(H, B), nthr, smem, st, Q, ldq, U, s, bias, w, beta, key_pad, dO, lddo, dQ, lddq,
                     gU, gs, gbias, gw, gbeta, Tq, T, H * dh, scale);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}
#undef SKF_XFOLD_DISPATCH

int skf_add_row_broadcast(float* Y, const float* r, int rows, int N, void* stream) {
  SKF_REQUIRE(Y && r && rows > 0 && N > 0 && N % 4 == 0, SKF_ERR_BAD_ARG, "add_row_broadcast: bad argument");
  launch_k(add_row_broadcast_k, dim3(cdiv((int64_t)rows * N / 4, 256)), dim3(256), 0, as_stream(stream), PDL_SIMT, Y, r, rows, N);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_token_ce(const float* logits, int64_t ldl, const int64_t* tgt, int64_t tgt_ld, int B,
                 int Td, int V, float grad_scale, float* dlogits, int64_t lddl, double* stats,
                 int32_t* argmax_out, int round_tf32, void* stream) {
  dim3 blk(32, 8);
  int rows = B * Td;
  launch_k(softmax_ce_k, dim3(cdiv(rows, 8)), dim3(blk), 0, as_stream(stream), PDL_SIMT, logits, ldl, tgt, tgt_ld, rows, Td, V, 1,
                                                            grad_scale, dlogits, lddl, nullptr, 0,
                                                            argmax_out, stats, round_tf32);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_class_ce(const float* logits, int64_t ldl, const int64_t* labels, int B, int C,
                 float grad_scale, float* dlogits, int64_t lddl, double* stats,
                 int32_t* argmax_out, float* probs, void* stream) {
  dim3 blk(32, 8);
  launch_k(softmax_ce_k, dim3(cdiv(B, 8)), dim3(blk), 0, as_stream(stream), PDL_SIMT, logits, ldl, labels, 1, B, 1, C, 0, grad_scale,
                                                         dlogits, lddl, probs, C, argmax_out, stats, 0);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int skf_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, int64_t* step,
                  int d_model, int warmup_steps, float beta1, float beta2, float eps,
                  float grad_scale, void* stream) {
  float rsqrt_d = (float)(1.0 / sqrt((double)(float)d_model));
  float warm_pow = (float)pow((double)warmup_steps, -1.5);
  adam_k<<<cdiv(cdiv(n, 4), 256), 256, 0, as_stream(stream)>>>(params, grads, m, v, n, step, rsqrt_d,
                                                              warm_pow, beta1, beta2, eps, grad_scale);
  SKF_LAUNCH_CHECK();
  step_inc_k<<<1, 1, 0, as_stream(stream)>>>(step);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

}  // extern "C"
