// Greedy decoding kernels behind Transformer.predict / predict_from_embedding
// (models/sketchformer.py:201-221, 255-311).  The reference re-runs the whole decoder on the growing
// prefix at every step (O(T^2) layer passes, one host sync per token).  The decoder is causal and the
// key-padding flag of a generated position never changes once it is emitted, so the hidden states of
// earlier positions are step-invariant: here every step processes ONE new position per sample against
// per-layer key/value caches in HBM, and the EOS / pen-state bookkeeping stays on the device.
#include <math.h>

#include "skf_common.cuh"

namespace skf {

constexpr float DEC_MASK_NEG = -1e9f;   // builders/utils.py:93-97: added after the 1/sqrt(dh) scaling
constexpr int DEC_MAX_KEYS = 1024;      // maximum_position_encoding = 1000 (transformer.py:268, 307)

// One warp per (sample, head): single-query attention over `nkeys` cached keys.
//   pos >= 0 : first append this step's key / value rows (knew / vnew) to the cache at position pos
//   masked(j) = (pad && pad[b*pad_ld + j]) || j >= limit           -> logits += -1e9
template <int DH>
__global__ void __launch_bounds__(128)
decode_attn_k(const float* __restrict__ q, int64_t ldq, const float* __restrict__ knew,
              const float* __restrict__ vnew, int64_t ldn, float* Kc, float* Vc, int64_t ldc, int64_t cache_bs,
              int pos, int nkeys, const uint8_t* __restrict__ pad, int64_t pad_ld, int limit,
              float* __restrict__ out, int64_t ldo, int B, int H, float scale) {
  __shared__ float ps[4][DEC_MAX_KEYS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + w;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh - b * H;
  float* kb = Kc + (int64_t)b * cache_bs + h * DH;
  float* vb = Vc + (int64_t)b * cache_bs + h * DH;
  if (pos >= 0) {
    for (int c = lane; c < DH; c += 32) {
      kb[(int64_t)pos * ldc + c] = knew[(int64_t)b * ldn + h * DH + c];
      vb[(int64_t)pos * ldc + c] = vnew[(int64_t)b * ldn + h * DH + c];
    }
    __syncwarp();
  }
  float qr[DH];
  {
    const float4* q4 = reinterpret_cast<const float4*>(q + (int64_t)b * ldq + h * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; c++) {
      const float4 v = __ldg(q4 + c);
      qr[4 * c] = v.x; qr[4 * c + 1] = v.y; qr[4 * c + 2] = v.z; qr[4 * c + 3] = v.w;
    }
  }
  float mx = -INFINITY;
  for (int j = lane; j < nkeys; j += 32) {
    const float4* k4 = reinterpret_cast<const float4*>(kb + (int64_t)j * ldc);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; c++) {
      const float4 v = k4[c];
      acc = fmaf(qr[4 * c], v.x, acc); acc = fmaf(qr[4 * c + 1], v.y, acc);
      acc = fmaf(qr[4 * c + 2], v.z, acc); acc = fmaf(qr[4 * c + 3], v.w, acc);
    }
    float s = acc * scale;
    if ((pad && pad[(int64_t)b * pad_ld + j]) || j >= limit) s += DEC_MASK_NEG;
    ps[w][j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nkeys; j += 32) {
    const float e = expf(ps[w][j] - mx);
    ps[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  for (int c = lane; c < DH; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < nkeys; j++) acc = fmaf(ps[w][j], vb[(int64_t)j * ldc + c], acc);
    out[(int64_t)b * ldo + h * DH + c] = acc * inv;
  }
}

// state[0] = number of valid output columns once every row has finished (0 = still running)
__global__ void decode_init_k(int64_t* __restrict__ tokens, float* __restrict__ strokes, uint8_t* __restrict__ pad,
                              uint8_t* __restrict__ cross_pad, uint8_t* __restrict__ fin, int32_t* __restrict__ state,
                              const int32_t* __restrict__ expected_len, int B, int T, int64_t sos, int continuous) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j <= T; j += blockDim.x) {
    pad[(int64_t)b * (T + 1) + j] = 0;
    if (continuous) {
      float* r = strokes + ((int64_t)b * (T + 1) + j) * 5;
      r[0] = 0.f; r[1] = 0.f; r[2] = j == 0 ? 1.f : 0.f; r[3] = 0.f; r[4] = 0.f;   // start row [0,0,1,0,0]
    } else {
      tokens[(int64_t)b * (T + 1) + j] = j == 0 ? sos : 0;
    }
    if (j < T) cross_pad[(int64_t)b * T + j] = (expected_len && j >= expected_len[b]) ? 1 : 0;
  }
  if (threadIdx.x == 0) fin[b] = 0;
  if (b == 0 && threadIdx.x < 8) state[threadIdx.x] = 0;
}

// Token mode: argmax of the last position's logits (lowest index among ties, tf.argmax), appended as the
// next decoder input; a row is finished once it has emitted EOS (sketchformer.py:296-301).
__global__ void decode_pick_token_k(const float* __restrict__ logits, int64_t ldl, int V, int64_t* __restrict__ tokens,
                                    uint8_t* __restrict__ pad, uint8_t* __restrict__ fin,
                                    const int32_t* __restrict__ state, int B, int T, int step, int64_t eos) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B || state[0] != 0) return;
  const float* z = logits + (int64_t)b * ldl;
  float mx = -INFINITY;
  int am = 0x7fffffff;
  for (int j = lane; j < V; j += 32) {
    const float v = z[j];
    if (v > mx) { mx = v; am = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  if (lane == 0) {
    tokens[(int64_t)b * (T + 1) + step + 1] = am;
    pad[(int64_t)b * (T + 1) + step + 1] = am == 0 ? 1 : 0;
    if (am == eos) fin[b] = 1;
  }
}

// Continuous mode: next input = [dx, dy, softmax(pen logits)] (sketchformer.py:287-290); fin[b] is the
// flag of THIS step only (all rows must predict pen state 2 in the same step, 291-294).
__global__ void decode_pick_stroke_k(const float* __restrict__ pred, int64_t ldp, float* __restrict__ strokes,
                                     uint8_t* __restrict__ pad, uint8_t* __restrict__ fin,
                                     const int32_t* __restrict__ state, int B, int T, int step) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || state[0] != 0) return;
  const float* p = pred + (int64_t)b * ldp;
  const float mx = fmaxf(p[2], fmaxf(p[3], p[4]));
  const float e0 = expf(p[2] - mx), e1 = expf(p[3] - mx), e2 = expf(p[4] - mx);
  const float inv = 1.0f / (e0 + e1 + e2);
  const float s0 = e0 * inv, s1 = e1 * inv, s2 = e2 * inv;
  float* r = strokes + ((int64_t)b * (T + 1) + step + 1) * 5;
  r[0] = p[0]; r[1] = p[1]; r[2] = s0; r[3] = s1; r[4] = s2;
  pad[(int64_t)b * (T + 1) + step + 1] = s2 == 1.0f ? 1 : 0;       // builders/utils.py:38-39 on the next prefix
  int am = 0;
  float best = s0;
  if (s1 > best) { best = s1; am = 1; }
  if (s2 > best) { am = 2; }
  fin[b] = am == 2 ? 1 : 0;
}

__global__ void decode_check_k(const uint8_t* __restrict__ fin, int32_t* __restrict__ state, int B, int step) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) c += fin[b] ? 1 : 0;
  if (c) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    state[1] = step + 1;                                  // steps executed so far
    if (state[0] == 0 && cnt == B) state[0] = step + 2;   // output columns: start token + (step+1) predictions
  }
}

int decode_attn(const float* q, int64_t ldq, const float* knew, const float* vnew, int64_t ldn, float* Kc, float* Vc,
                int64_t ldc, int64_t cache_bs, int pos, int nkeys, const uint8_t* pad, int64_t pad_ld, int limit,
                float* out, int64_t ldo, int B, int H, int dh, cudaStream_t st) {
  SKF_REQUIRE(nkeys >= 1 && nkeys <= DEC_MAX_KEYS, SKF_ERR_BAD_SHAPE, "decode_attn: %d keys", nkeys);
  SKF_REQUIRE((ldq & 3) == 0 && (ldc & 3) == 0, SKF_ERR_BAD_ALIGN, "decode_attn: strides must be multiples of 4");
  const float scale = 1.0f / sqrtf((float)dh);
  const int grid = cdiv((long)B * H, 4);
#define SKF_DEC_ATTN(DH)                                                                                        \
  decode_attn_k<DH><<<grid, 128, 0, st>>>(q, ldq, knew, vnew, ldn, Kc, Vc, ldc, cache_bs, pos, nkeys, pad, pad_ld, \
                                          limit, out, ldo, B, H, scale)
  if (dh == 16) SKF_DEC_ATTN(16);
  else if (dh == 32) SKF_DEC_ATTN(32);
  else if (dh == 64) SKF_DEC_ATTN(64);
  else return fail(SKF_ERR_UNSUPPORTED, "decode_attn: head dim %d", dh);
#undef SKF_DEC_ATTN
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int decode_init(int64_t* tokens, float* strokes, uint8_t* pad, uint8_t* cross_pad, uint8_t* fin, int32_t* state,
                const int32_t* expected_len, int B, int T, int64_t sos, int continuous, cudaStream_t st) {
  decode_init_k<<<B, 128, 0, st>>>(tokens, strokes, pad, cross_pad, fin, state, expected_len, B, T, sos, continuous);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

// make_dummy_input with expected_len=None (models/sketchformer.py:230-253, 275): at step i the cross-attention keys
// [0, width = i + 1) are attended, the rest masked, for every sample
__global__ void decode_cross_width_k(uint8_t* __restrict__ cross_pad, int n, int T, int width) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) cross_pad[idx] = (idx % T) >= width ? 1 : 0;
}
int decode_cross_width(uint8_t* cross_pad, int B, int T, int width, cudaStream_t st) {
  decode_cross_width_k<<<cdiv((long)B * T, 256), 256, 0, st>>>(cross_pad, B * T, T, width);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

int decode_pick(const float* logits, int64_t ldl, int V, int64_t* tokens, float* strokes, uint8_t* pad, uint8_t* fin,
                int32_t* state, int B, int T, int step, int64_t eos, int continuous, cudaStream_t st) {
  if (continuous)
    decode_pick_stroke_k<<<cdiv(B, 128), 128, 0, st>>>(logits, ldl, strokes, pad, fin, state, B, T, step);
  else
    decode_pick_token_k<<<cdiv(B, 4), 128, 0, st>>>(logits, ldl, V, tokens, pad, fin, state, B, T, step, eos);
  SKF_LAUNCH_CHECK();
  decode_check_k<<<1, 256, 0, st>>>(fin, state, B, step);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

}  // namespace skf
