// Per-sketch preprocessing + dictionary tokenisation of the stroke3-distributed loader, fused on the device
// (dataloaders/distributed_stroke3.py:90-125, 138-153; utils/tokenizer.py:35-64; utils/sketch.py:31-50):
//   clip raw deltas to +-limit -> bounding box of the absolute path (origin included) -> divide (dx, dy) by
//   max(bbox_w, bbox_h, 1) -> nearest KMeans centre + 1 -> SEP after every pen-lift point -> [SOS] .. [EOS]
//   -> truncate to max_seq_len (EOS may be cut) -> PAD.
// One CTA per sketch.  The path integral is a sequential float64 sum in point order, exactly like the
// reference's Python loop, so the normalised coordinates are bit-identical to the host restatement; the
// 1000 x 2 centre table sits in shared memory and every thread scans it for its points.
#include "skf_common.cuh"

namespace skf {

constexpr int TK_MAX_T = 1024;

// The inputs are read with plain (coherent) loads, not through the read-only path: they are uploaded by the loader's
// prefetch thread right before the launch and the buffers are recycled from chunk to chunk.
__global__ void __launch_bounds__(256)
tokenize_strokes_k(const float* pts, const int64_t* offs, const float* cent, int K, float limit, int64_t* out, int T) {
  extern __shared__ float sm_c[];                 // K x 2 centres
  __shared__ int s_pos[TK_MAX_T];                 // token position of point i (i < T)
  __shared__ float s_inv[2];                      // max_dim (as float), unused slot
  __shared__ int s_total;                         // position of EOS
  const int b = blockIdx.x;
  const int64_t p0 = offs[b];
  const int n = (int)(offs[b + 1] - p0);
  const float* P = pts + p0 * 3;
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) sm_c[i] = __ldcg(cent + i);
  int64_t* row = out + (int64_t)b * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) row[t] = 0;   // PAD
  if (threadIdx.x == 0) {
    double ax = 0.0, ay = 0.0, mnx = 0.0, mxx = 0.0, mny = 0.0, mxy = 0.0;
    int lifts = 0;
    for (int i = 0; i < n; i++) {
      const float x = fminf(fmaxf(__ldcg(P + 3 * i), -limit), limit);
      const float y = fminf(fmaxf(__ldcg(P + 3 * i + 1), -limit), limit);
      const float pen = fminf(fmaxf(__ldcg(P + 3 * i + 2), -limit), limit);
      ax += (double)x; ay += (double)y;
      mnx = fmin(mnx, ax); mxx = fmax(mxx, ax); mny = fmin(mny, ay); mxy = fmax(mxy, ay);
      if (i < T) s_pos[i] = 1 + i + lifts;
      if (pen == 1.0f) lifts++;
    }
    const double md = fmax(fmax(mxx - mnx, mxy - mny), 1.0);
    s_inv[0] = (float)md;
    s_total = 1 + n + lifts;
  }
  __syncthreads();
  const float md = s_inv[0];
  if (threadIdx.x == 0) {
    row[0] = K + 2;                                        // SOS
    if (s_total < T) row[s_total] = K + 3;                 // EOS (cut when the sketch overflows max_seq_len)
  }
  const int np = n < T ? n : T;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const int pos = s_pos[i];
    if (pos >= T) continue;
    const float x = __fdiv_rn(fminf(fmaxf(__ldcg(P + 3 * i), -limit), limit), md);
    const float y = __fdiv_rn(fminf(fmaxf(__ldcg(P + 3 * i + 1), -limit), limit), md);
    const float pen = fminf(fmaxf(__ldcg(P + 3 * i + 2), -limit), limit);
    float best = INFINITY;
    int arg = 0;
    for (int k = 0; k < K; k++) {
      const float dx = __fsub_rn(x, sm_c[2 * k]), dy = __fsub_rn(y, sm_c[2 * k + 1]);
      const float dist = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));   // no FMA: same bits as numpy
      if (dist < best) { best = dist; arg = k; }
    }
    row[pos] = arg + 1;
    if (pen == 1.0f && pos + 1 < T) row[pos + 1] = K + 1;  // SEP
  }
}

}  // namespace skf

extern "C" int skf_tokenize_strokes(const float* points, const int64_t* offsets, int B, const float* centroids,
                                    int K, float limit, int64_t* out, int T, void* stream) {
  SKF_REQUIRE(points && offsets && centroids && out, SKF_ERR_BAD_ARG, "tokenize: NULL argument");
  SKF_REQUIRE(B > 0 && K > 0 && K <= 4096 && T > 1 && T <= skf::TK_MAX_T, SKF_ERR_BAD_SHAPE,
              "tokenize: B=%d K=%d T=%d unsupported", B, K, T);
  skf::tokenize_strokes_k<<<B, 256, (size_t)2 * K * sizeof(float), skf::as_stream(stream)>>>(points, offsets, centroids, K,
                                                                                       limit, out, T);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}
