// Whole-step orchestration: parameter arena layout, workspace layout, forward, losses, backward
// and optimizer of Transformer.model_trainer (models/sketchformer.py:325-349), plus the inference
// entry points (encode_from_seq / predict_class, sketchformer.py:162-168, 223-228).
// Host code only decides WHICH kernel runs on WHICH buffer; every byte of arithmetic happens in
// the CUDA kernels of this library.
#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_bf16.h>

#include "skf_common.cuh"

namespace skf {

int attn_fwd_simt(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float*,
                  const uint8_t*, int, int, int, int, int, int, int, cudaStream_t);
int attn_bwd_simt(const float*, int64_t, const float*, int64_t, const float*, int64_t, const float*, int64_t,
                  const float*, int64_t, const float*, float*, float*, int64_t, float*, int64_t, float*, int64_t,
                  const uint8_t*, int, int, int, int, int, int, int, cudaStream_t);
// 1 = tensor maps of the single-pass TF32 GEMM use CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 (the TMA unit rounds on load)
int g_tma_tf32_type = 1;
int gemm_tc2_supported(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const float*, int64_t, int, int,
                       int, const skf_gemm_epilogue*);
int gemm_tc2(const float*, int64_t, int64_t, const float*, int64_t, int64_t, float*, int64_t, int, int, int,
             const skf_gemm_epilogue*, int, cudaStream_t);
int gemm_tc3(const float*, int64_t, int64_t, const float*, int64_t, int64_t, float*, int64_t, int, int, int,
             const skf_gemm_epilogue*, cudaStream_t);
int gemm_tc3_wanted(int, int, int);
int g_gemm_pair = 1;
int g_wgrad_units = 1;     // side-stream wgrads: work units per cluster slot (finer units yield SMs sooner)
extern int g_t3_oversubscribe;
int g_kv_fold = 2;         // rank-2 form of the expanded bottleneck (see kv_fold()): 1 K / V only, 2 whole cross attention
int g_wgrad_overlap = 1;   // weight-gradient GEMMs of the backward pass on a second stream (see SideQueue)
extern int g_t3_bn128_max;
extern int g_t3_tail_split;  // pair GEMM: 128-column cluster tiles for N up to this
int g_gemm_tiny_simt = 0;   // 1: products with M <= 128 and few column tiles run on the small-tile SIMT kernel (no gain measured: 9.04 vs 9.01 ms)
extern int g_gemm_wsplit;   // 1: forward / dgrad products read pre-split bf16 images of the weights (skf_gemm_tc3.cu)
extern int g_gemm_mix;   // 1: 3xTF32 GEMMs with N > 128 run on CTA pairs (skf_gemm_tc3.cu)
int attn3_supported(const float*, int64_t, const float*, int64_t, const float*, int64_t, int, int, int);
int attn3_fwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float*,
              const uint8_t*, int, int, int, int, int, int, void*, cudaStream_t);
int attn3_bwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, const float*, int64_t,
              const float*, int64_t, const float*, float*, float*, int64_t, float*, int64_t, float*, int64_t,
              const uint8_t*, int, int, int, int, int, int, const void*, void*, cudaStream_t);
size_t attn3_image_bytes(int, int, int, int, int);
extern int g_a3_trace;
int attn3_trace_dump(long long*, int);
int g_attn3_img = 1;       // 1: the backward kernels read the bf16 operand images the forward kernel wrote (training only)
int g_attn3 = 3;           // bit 0 / bit 1: head dim 32 and sequences up to 224 run the warp-specialised whole-row forward /
                           // backward kernels (skf_attn_tc3.cu)
int attn2_supported(const float*, int64_t, const float*, int64_t, const float*, int64_t, int, int, int);
int attn2_fwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float*,
              const uint8_t*, int, int, int, int, int, int, int, cudaStream_t);
int attn2_bwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, const float*, int64_t,
              const float*, int64_t, const float*, float*, float*, int64_t, float*, int64_t, float*, int64_t,
              const uint8_t*, int, int, int, int, int, int, int, cudaStream_t);
int allreduce_sum_f32(float*, int64_t, void*, cudaStream_t);       // skf_dp.cu (NCCL, bound at run time)
int decode_attn(const float*, int64_t, const float*, const float*, int64_t, float*, float*, int64_t, int64_t, int, int,
                const uint8_t*, int64_t, int, float*, int64_t, int, int, int, cudaStream_t);
int decode_init(int64_t*, float*, uint8_t*, uint8_t*, uint8_t*, int32_t*, const int32_t*, int, int, int64_t, int,
                cudaStream_t);
int decode_pick(const float*, int64_t, int, int64_t*, float*, uint8_t*, uint8_t*, int32_t*, int, int, int, int64_t,
                int, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// parameter arena
// ---------------------------------------------------------------------------------------------
struct PEntry {
  std::string name;
  int64_t off;
  int rows, cols;
};
struct ParamLayout {
  std::vector<PEntry> e;
  std::map<std::string, int> idx;
  int64_t total = 0;
  void add(const std::string& n, int rows, int cols) {
    total = (total + 7) & ~(int64_t)7;  // 32-byte aligned tensors: 16 bytes (TMA / float4) in the bf16 images too
    idx[n] = (int)e.size();
    e.push_back({n, total, rows, cols});
    total += (int64_t)rows * cols;
  }
  int64_t off(const std::string& n) const { return e[idx.at(n)].off; }
  bool has(const std::string& n) const { return idx.count(n) != 0; }
};

static int check_cfg(const skf_cfg* c) {
  SKF_REQUIRE(c != nullptr, SKF_ERR_BAD_ARG, "cfg is NULL");
  SKF_REQUIRE(c->B > 0 && c->T > 1 && c->d > 0 && c->f > 0 && c->H > 0 && c->L > 0, SKF_ERR_BAD_SHAPE,
              "cfg: non-positive dimension");
  SKF_REQUIRE(c->d % c->H == 0, SKF_ERR_BAD_SHAPE, "cfg: d_model %% num_heads != 0");
  SKF_REQUIRE(c->d % 4 == 0 && c->f % 4 == 0, SKF_ERR_BAD_SHAPE, "cfg: d_model and dff must be multiples of 4");
  int dh = c->d / c->H;
  SKF_REQUIRE(dh == 16 || dh == 32 || dh == 64, SKF_ERR_UNSUPPORTED, "cfg: head dim %d not in {16,32,64}", dh);
  SKF_REQUIRE(c->u >= 0 && c->u % 4 == 0, SKF_ERR_BAD_SHAPE, "cfg: lowerdim must be a non-negative multiple of 4");
  // lowerdim=0: no bottleneck / classifier / expander, the decoder attends to the encoder output
  // (models/sketchformer.py:96-107, 157-158, 175-176); build_model() then creates no classify_layer
  SKF_REQUIRE(c->u > 0 || (!c->do_classification && c->do_reconstruction), SKF_ERR_BAD_ARG,
              "cfg: lowerdim=0 needs do_classification=0 and do_reconstruction=1");
  SKF_REQUIRE(c->continuous || (c->V > 0 && c->V % 4 == 0), SKF_ERR_BAD_SHAPE,
              "cfg: vocab size must be a positive multiple of 4");
  SKF_REQUIRE(!c->do_classification || c->C > 0, SKF_ERR_BAD_SHAPE, "cfg: n_classes");
  SKF_REQUIRE(c->T <= 1000, SKF_ERR_BAD_SHAPE, "cfg: max_seq_len > maximum_position_encoding (1000)");
  SKF_REQUIRE(c->dropout_rate >= 0.f && c->dropout_rate < 1.f, SKF_ERR_BAD_ARG, "cfg: dropout_rate outside [0, 1)");
  SKF_REQUIRE(c->class_dropout >= 0.f && c->class_dropout < 1.f, SKF_ERR_BAD_ARG, "cfg: class_dropout outside [0, 1)");
  SKF_REQUIRE(c->class_buffer_layers >= 0 && c->class_buffer_layers <= 8, SKF_ERR_BAD_SHAPE,
              "cfg: class_buffer_layers outside [0, 8]");
  return SKF_OK;
}

// SelfAttnV1: pooling hidden width = lowerdim, embedding width = d_model;
// SelfAttnV2: hidden width = d_model, Dense(lowerdim) after the pooling, embedding width = lowerdim.
static inline bool pool_v2(const skf_cfg& c) { return c.u > 0 && c.attn_version != 0 && c.attn_version != 1; }
static inline int emb_dim(const skf_cfg& c) { return pool_v2(c) ? c.u : c.d; }
static inline int pool_dim(const skf_cfg& c) { return pool_v2(c) ? c.d : c.u; }

static void add_mha_self(ParamLayout& L, const std::string& p, int d) {
  L.add(p + ".wqkv.k", d, 3 * d);
  L.add(p + ".wqkv.b", 1, 3 * d);
  L.add(p + ".wo.k", d, d);
  L.add(p + ".wo.b", 1, d);
}
static void add_ffn_ln(ParamLayout& L, const std::string& p, int d, int f, int nln) {
  L.add(p + ".ffn.w1.k", d, f);
  L.add(p + ".ffn.w1.b", 1, f);
  L.add(p + ".ffn.w2.k", f, d);
  L.add(p + ".ffn.w2.b", 1, d);
  for (int i = 1; i <= nln; i++) {
    L.add(p + ".ln" + std::to_string(i) + ".g", 1, d);
    L.add(p + ".ln" + std::to_string(i) + ".b", 1, d);
  }
}

static ParamLayout build_layout(const skf_cfg& c) {
  ParamLayout L;
  if (c.continuous) {            // Dense(d_model) on stroke-5 rows (transformer.py:275-276)
    L.add("enc.emb.k", 5, c.d);
    L.add("enc.emb.b", 1, c.d);
  } else {
    L.add("enc.emb", c.V, c.d);
  }
  for (int i = 0; i < c.L; i++) {
    std::string p = "enc.l" + std::to_string(i);
    add_mha_self(L, p + ".mha", c.d);
    add_ffn_ln(L, p, c.d, c.f, 2);
  }
  if (c.u > 0) {
    L.add("pool.W", c.d, pool_dim(c));
    L.add("pool.b", 1, pool_dim(c));
    L.add("pool.V", 1, pool_dim(c));
  }
  if (pool_v2(c)) {                       // SelfAttnV2.embeding_layer (transformer.py:99-100, 129-130)
    L.add("pool.emb.k", c.d, c.u);
    L.add("pool.emb.b", 1, c.u);
  }
  if (c.do_classification) {              // models/sketchformer.py:99-104
    int fin = emb_dim(c);
    for (int j = 0; j < c.class_buffer_layers; j++) {
      L.add("cbuf" + std::to_string(j) + ".k", fin, c.u);
      L.add("cbuf" + std::to_string(j) + ".b", 1, c.u);
      fin = c.u;
    }
    L.add("cls.k", fin, c.C);
    L.add("cls.b", 1, c.C);
  }
  if (c.do_reconstruction) {
    if (c.u > 0) {
      L.add("exp.k", 1, c.T);
      L.add("exp.b", 1, c.T);
    }
    if (c.continuous) {
      L.add("dec.emb.k", 5, c.d);
      L.add("dec.emb.b", 1, c.d);
    } else {
      L.add("dec.emb", c.V, c.d);
    }
    for (int i = 0; i < c.L; i++) {
      std::string p = "dec.l" + std::to_string(i);
      add_mha_self(L, p + ".mha1", c.d);
      L.add(p + ".mha2.wq.k", c.d, c.d);
      L.add(p + ".mha2.wq.b", 1, c.d);
      L.add(p + ".mha2.wkv.k", emb_dim(c), 2 * c.d);   // keys / values come from the expanded embedding
      L.add(p + ".mha2.wkv.b", 1, 2 * c.d);
      L.add(p + ".mha2.wo.k", c.d, c.d);
      L.add(p + ".mha2.wo.b", 1, c.d);
      add_ffn_ln(L, p, c.d, c.f, 3);
    }
    L.add("out.k", c.d, c.continuous ? 5 : c.V);   // sketchformer.py:86 / 91
    L.add("out.b", 1, c.continuous ? 5 : c.V);
  }
  L.total = (L.total + 7) & ~(int64_t)7;
  return L;
}

// Gradient buckets of the data-parallel exchange: contiguous arena ranges listed in the order in which
// the backward pass completes them.  The arena is laid out in forward order (encoder, bottleneck, classifier,
// expander, decoder, output layer), so the ranges are cut from its END: [second half of the decoder + output layer],
// [rest of the decoder + expander], [second half of the encoder + bottleneck + classifier], [rest of the encoder +
// its embedding].  stage[k] names the point of backward() after which bucket k is final:
//   0 = decoder layer L/2 done, 1 = decoder done, 2 = encoder layer L/2 done, 3 = everything done.
struct Buckets {
  int n = 0;
  int64_t lo[4], hi[4];
  int stage[4];
};
static Buckets build_buckets(const skf_cfg& c, const ParamLayout& PL, int want) {
  Buckets b;
  if (want < 1) want = 1;
  if (want > 4) want = 4;
  const int half = c.L / 2;
  std::vector<std::pair<int64_t, int>> cuts;        // (arena offset where a bucket starts, stage that completes it)
  if (c.do_reconstruction) {
    const int64_t dec_start = PL.off(c.u > 0 ? "exp.k" : (c.continuous ? "dec.emb.k" : "dec.emb"));
    if (want >= 4 && half > 0) cuts.push_back({PL.off("dec.l" + std::to_string(half) + ".mha1.wqkv.k"), 0});
    if (want >= 2) cuts.push_back({dec_start, 1});
  }
  if ((want >= 3 || (!c.do_reconstruction && want >= 2)) && half > 0)
    cuts.push_back({PL.off("enc.l" + std::to_string(half) + ".mha.wqkv.k"), 2});
  cuts.push_back({0, 3});
  int64_t end = PL.total;
  for (auto& ct : cuts) {
    if (ct.first >= end) continue;
    b.lo[b.n] = ct.first; b.hi[b.n] = end; b.stage[b.n] = ct.second;
    end = ct.first;
    b.n++;
  }
  return b;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct WEntry {
  size_t off, bytes;
};
struct Workspace {
  std::map<std::string, WEntry> t;
  size_t total = 0;
  void add(const std::string& n, size_t bytes) {
    total = (total + 255) & ~(size_t)255;
    t[n] = {total, bytes};
    total += bytes;
  }
};

// operand images of the whole-row attention kernels (skf_attn_tc3.cu): only for head dim 32 and sequences <= 224
static size_t attn_image_bytes(const skf_cfg& c, int T, int ntiles) {
  if (c.d / c.H != 32 || T > 224 || c.gemm_backend != 0 || c.attn_backend != 0) return 256;
  return attn3_image_bytes(c.B, c.H, T, T, ntiles);
}

static Workspace build_workspace(const skf_cfg& c, const ParamLayout& PL) {
  Workspace W;
  const size_t F = sizeof(float);
  const size_t Me = (size_t)c.B * c.T, Md = (size_t)c.B * (c.T - 1);
  const size_t d = c.d, f = c.f, H = c.H;
  W.add("step", 8);
  W.add("stats_step", 8 * sizeof(double));
  W.add("stats_run", 16 * sizeof(double));
  W.add("pe", (size_t)c.T * d * F);
  W.add("wc", g_tma_tf32_type ? 256 : (size_t)PL.total * F);  // TF32-rounded copy of the weights (tma_tf32=0 only)
  // bf16 images hi = bf16(w), lo = bf16(w - hi) of the parameter arena, same element offsets (pair GEMM, bf16x3)
  W.add("w16h", (size_t)PL.total * 2);
  W.add("w16l", (size_t)PL.total * 2);
  W.add("enc_pad", Me);
  W.add("dec_pad", Md);
  W.add("enc.x0", Me * d * F);
  for (int i = 0; i < c.L; i++) {
    std::string p = "enc.l" + std::to_string(i);
    W.add(p + ".qkv", Me * 3 * d * F);
    W.add(p + ".ctx", Me * d * F);
    W.add(p + ".lse", (size_t)c.B * H * c.T * F);
    W.add(p + ".img", attn_image_bytes(c, c.T, 6));        // bf16 hi / lo images of q, k, v (whole-row attention)
    W.add(p + ".s1", Me * d * F);
    W.add(p + ".mean1", Me * F);
    W.add(p + ".rstd1", Me * F);
    W.add(p + ".out1", Me * d * F);
    W.add(p + ".hff", Me * f * F);
    W.add(p + ".s2", Me * d * F);
    W.add(p + ".mean2", Me * F);
    W.add(p + ".rstd2", Me * F);
    W.add(p + ".out2", Me * d * F);
  }
  const size_t E = emb_dim(c);
  W.add("pool.u", Me * pool_dim(c) * F);
  W.add("pool.a", Me * F);
  W.add("pool.o", (size_t)c.B * d * F);            // pooled vector before SelfAttnV2's Dense
  W.add("embedding", (size_t)c.B * E * F);
  for (int j = 0; j < c.class_buffer_layers; j++) {
    W.add("cbuf" + std::to_string(j) + ".h", (size_t)c.B * c.u * F);    // relu output
    W.add("cbuf" + std::to_string(j) + ".hd", (size_t)c.B * c.u * F);   // after Dropout(class_dropout)
  }
  W.add("g.cb0", (size_t)c.B * c.u * F);
  W.add("g.cb1", (size_t)c.B * c.u * F);
  W.add("g.po", (size_t)c.B * d * F);
  size_t Cp = ((size_t)c.C + 3) & ~(size_t)3;
  W.add("cls.logits", (size_t)c.B * Cp * F);
  W.add("cls.dlogits", (size_t)c.B * Cp * F);
  W.add("cls.probs", (size_t)c.B * (c.C > 0 ? c.C : 1) * F);
  W.add("cls.argmax", (size_t)c.B * 4);
  if (c.do_reconstruction) {
    W.add("pre_decoder", Me * E * F);
    W.add("dec.x0", Md * d * F);
    for (int i = 0; i < c.L; i++) {
      std::string p = "dec.l" + std::to_string(i);
      W.add(p + ".qkv", Md * 3 * d * F);
      W.add(p + ".ctx1", Md * d * F);
      W.add(p + ".lse1", (size_t)c.B * H * (c.T - 1) * F);
      W.add(p + ".img1", attn_image_bytes(c, c.T - 1, 6));
      W.add(p + ".s1", Md * d * F);
      W.add(p + ".mean1", Md * F);
      W.add(p + ".rstd1", Md * F);
      W.add(p + ".out1", Md * d * F);
      W.add(p + ".q2", Md * d * F);
      W.add(p + ".kv2", Me * 2 * d * F);
      W.add(p + ".ctx2", Md * d * F);
      W.add(p + ".lse2", (size_t)c.B * H * (c.T - 1) * F);
      W.add(p + ".s2", Md * d * F);
      W.add(p + ".mean2", Md * F);
      W.add(p + ".rstd2", Md * F);
      W.add(p + ".out2", Md * d * F);
      W.add(p + ".hff", Md * f * F);
      W.add(p + ".s3", Md * d * F);
      W.add(p + ".mean3", Md * F);
      W.add(p + ".rstd3", Md * F);
      W.add(p + ".out3", Md * d * F);
    }
    W.add("recon.logits", Md * (size_t)(c.continuous ? 8 : c.V) * F);   // continuous: [Md, 5] with ld 8
    W.add("recon.argmax", Md * 4);
    // greedy decode (predict / predict_from_embedding): generated prefix, flags, per-layer self-attention
    // key/value caches [B, T, 2d] and one-position activations
    const size_t Bz = (size_t)c.B, Tp = (size_t)c.T + 1;
    W.add("gen.tokens", Bz * Tp * 8);
    W.add("gen.strokes", Bz * Tp * 5 * F);
    W.add("gen.pad", Bz * Tp);
    W.add("gen.cross_pad", Bz * c.T);
    W.add("gen.fin", Bz);
    W.add("gen.state", 8 * 4);
    for (int i = 0; i < c.L; i++) W.add("gen.kv" + std::to_string(i), Bz * c.T * 2 * d * F);
    W.add("gen.x", Bz * d * F);
    W.add("gen.qkv", Bz * 3 * d * F);
    W.add("gen.ctx", Bz * d * F);
    W.add("gen.s", Bz * d * F);
    W.add("gen.o1", Bz * d * F);
    W.add("gen.q2", Bz * d * F);
    W.add("gen.o2", Bz * d * F);
    W.add("gen.h", Bz * f * F);
    W.add("gen.logits", Bz * (size_t)(c.continuous ? 8 : c.V) * F);
  }
  // backward scratch
  W.add("g.a", Me * d * F);
  W.add("g.b", Me * d * F);
  W.add("g.c", Me * d * F);
  W.add("g.d", Me * d * F);
  for (const char* n : {"g.b1", "g.b2", "g.d1", "g.d2"}) W.add(n, Me * d * F);   // one (g.b, g.d) pair per LayerNorm site
  // rank-2 form of the cross-attention keys / values (kv_fold): U = emb Wkv per layer, s = colsum(Wkv), and the
  // reduced gradients gU, gs
  for (int i = 0; i < c.L; i++) {
    W.add("dec.l" + std::to_string(i) + ".kvU", (size_t)c.B * 2 * d * F);
    W.add("dec.l" + std::to_string(i) + ".kvs", (size_t)2 * d * F);
  }
  W.add("g.kvU", (size_t)c.B * 2 * d * F);
  W.add("g.kvs", (size_t)2 * d * F);
  W.add("g.h", Me * f * F);
  W.add("g.qkv", Me * 3 * d * F);
  W.add("g.q2", Me * d * F);
  W.add("g.kv2", Me * 2 * d * F);
  W.add("g.pre", Me * E * F);
  W.add("g.emb", (size_t)c.B * E * F);
  W.add("g.u", Me * pool_dim(c) * F);
  W.add("g.delta", (size_t)c.B * H * c.T * F);
  W.add("g.img_do", attn_image_bytes(c, c.T, 2));          // bf16 hi / lo images of dO (written by dq, read by dkv)
  W.total = (W.total + 255) & ~(size_t)255;
  return W;
}

}  // namespace skf

using namespace skf;

struct skf_plan {
  skf_cfg c;
  ParamLayout PL;
  Workspace W;
  float *params, *grads, *am, *av;
  uint8_t* ws;
  int64_t launches = 0;
  bool round_ops = false;  // round GEMM operands to TF32 when they are produced
  const float* wc = nullptr;
  bool training = false;   // inside skf_train_step: Dropout layers are active (training=True)
  int64_t gen_sos = 0, gen_eos = 0;   // greedy decode: start / stop token ids (utils/tokenizer.py:29-33)
  bool gen_ready = false;
  bool gen_uncached = false;          // blind_decoder_mask=False without expected_len: the prefix is recomputed every step
  // Backward pass, second stream: dW = X^T dY is off the critical path (nothing reads it before the optimizer),
  // so it is issued on `side` behind an event and fills the SMs the main chain leaves idle in the last wave of its
  // persistent GEMMs and attention kernels.  `pending` maps a gradient buffer to the event after which the side
  // stream no longer reads it; the main chain waits on it before it overwrites the buffer.
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> events;
  size_t ev_next = 0;
  std::unordered_map<const void*, cudaEvent_t> pending;
  bool side_active = false;
  bool w16_valid = false;   // the "w16h" / "w16l" images match `params` (refresh_compute_weights)
  // Data parallel (skf_plan_set_comm): gradient buckets are all-reduced on `comm_stream` as soon as the backward pass
  // has finished them, behind an event of the main chain and one of the side stream; the optimizer waits for the last
  void* comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  Buckets buckets;
  int buckets_sent = 0;
  cudaEvent_t next_event() { return events[ev_next++ % events.size()]; }
  ~skf_plan() {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (side) cudaStreamDestroy(side);
    if (comm_stream) cudaStreamDestroy(comm_stream);
  }

  float* wsf(const std::string& n) const { return reinterpret_cast<float*>(ws + W.t.at(n).off); }
  template <class T>
  T* wsp(const std::string& n) const { return reinterpret_cast<T*>(ws + W.t.at(n).off); }
  const float* w(const std::string& n) const { return wc + PL.off(n); }       // GEMM operand copy
  const float* pw(const std::string& n) const { return params + PL.off(n); }  // exact fp32 master
  float* g(const std::string& n) const { return grads + PL.off(n); }
};

namespace skf {

// lo = x - tf32(x).  rna: the TMA (TFLOAT32 maps) rounds to nearest; otherwise kind::tf32 truncates.
__global__ void split_lo_k(const float* __restrict__ x, float* __restrict__ lo, int64_t n, int rna) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(x + i), o;
    if (rna) {
      o.x = v.x - round_tf32(v.x); o.y = v.y - round_tf32(v.y);
      o.z = v.z - round_tf32(v.z); o.w = v.w - round_tf32(v.w);
    } else {
      o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
      o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
      o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
      o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    }
    *reinterpret_cast<float4*>(lo + i) = o;
  } else {
    for (; i < n; i++) {
      float v = x[i];
      lo[i] = rna ? v - round_tf32(v) : v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}
int split_lo(const float* x, float* lo, int64_t n, cudaStream_t st) {
  split_lo_k<<<cdiv(cdiv(n, 4), 256), 256, 0, st>>>(x, lo, n, g_tma_tf32_type);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

// hi = bf16(x), lo = bf16(x - hi): operand images of the bf16x3 product (x - hi is exact in fp32)
__global__ void split_bf16_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
    uint2 ph, pl;
    ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
    pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(hi + i) = ph;
    *reinterpret_cast<uint2*>(lo + i) = pl;
  } else {
    for (; i < n; i++) {
      const __nv_bfloat16 h = __float2bfloat16_rn(x[i]);
      hi[i] = h;
      lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h));
    }
  }
}
int split_bf16(const float* x, void* hi, void* lo, int64_t n, cudaStream_t st) {
  split_bf16_k<<<cdiv(cdiv(n, 4), 256), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(hi),
                                                      reinterpret_cast<__nv_bfloat16*>(lo), n);
  SKF_LAUNCH_CHECK();
  return SKF_OK;
}

__global__ void round_copy_k(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = round_tf32(src[i]);
}

// stats_step: [0] recon CE sum, [1] recon correct, [2] class CE sum, [3] class correct
// stats_run : [0] sum recon_loss [1] recon correct [2] recon count [3] sum class_loss [4] class correct
//             [5] class count [6] sum total_loss [7] steps | [8..12] last step: recon_loss, recon_acc,
//             class_loss, class_acc, total_loss     (builders/keras_metrics.py:19-42 running means)
__global__ void stats_finalize_k(const double* __restrict__ st, double* __restrict__ run, double n_recon,
                                 double n_class, double recon_w, double class_w, int do_recon, int do_class,
                                 int continuous) {
  double rl = do_recon ? recon_w * st[0] / n_recon : 0.0;
  // builders/losses.py:43-66: mean((loc + M) * mask) with M = mean(CE3); st[4..6] = sum loc*mask, sum CE3, sum mask
  if (do_recon && continuous) rl = recon_w * (st[4] + st[5] / n_recon * st[6]) / n_recon;
  double cl = do_class ? class_w * st[2] / n_class : 0.0;
  double tot = rl + cl;
  if (do_recon) { run[0] += rl; run[1] += continuous ? 0.0 : st[1]; run[2] += n_recon; }
  if (do_class) { run[3] += cl; run[4] += st[3]; run[5] += n_class; }
  run[6] += tot;
  run[7] += 1.0;
  run[8] = rl;
  run[9] = do_recon && !continuous ? st[1] / n_recon : 0.0;
  run[10] = cl;
  run[11] = do_class ? st[3] / n_class : 0.0;
  run[12] = tot;
}

// ---- GEMM front-end used by the plan ---------------------------------------------------------
struct Gemm {
  const skf_plan* P;
  cudaStream_t st;
  // a_elems / b_elems: number of floats of a dense activation operand (its low half is computed here);
  // 0 means the operand is a weight (its low half lives in the "wlo" arena at the same offset).
  int run(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* Y,
          int64_t ldy, int M, int N, int K, skf_gemm_epilogue e, bool x3 = false) const {
    // One 128-row block and a handful of column tiles (the [B, E] x [E, 2d] products of the folded cross attention,
    // the per-token products of the greedy decode at batch <= 128): a persistent tensor-core kernel would run its
    // whole k-loop on 2-4 SMs (18-20 us measured for 128 x 512 x 256); the small-tile fp32 kernel spreads it out.
    const bool tiny = g_gemm_tiny_simt && M <= 128 && !e.accumulate && (long)cdiv(M, 64) * cdiv(N, 64) <= 37;
    const bool tc = !tiny && P->c.gemm_backend == 0 &&
                    gemm_tc2_supported(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e);
    if (!P->round_ops) e.round_tf32 = 0;
    if (e.accumulate) {
      // split-K so that the persistent grid (one CTA per SM) has about one tile per SM
      const bool pair = tc && x3 && g_gemm_pair && gemm_tc3_wanted(M, N, K);
      const long tiles = pair ? (long)cdiv(M, 256) * cdiv(N, 256)
                         : tc ? (long)cdiv(M, 128) * cdiv(N, N <= 64 ? 64 : (N <= 128 ? 128 : 256))
                              : (long)cdiv(M, 64) * cdiv(N, 64);
      int want = (int)((pair ? 74 : tc ? 148 : 2 * 148) / tiles);
      if (pair && st == P->side && st != nullptr) want = (int)((long)74 * g_wgrad_units / tiles);
      int maxs = K / 256 > 0 ? K / 256 : 1;
      e.split_k = want < 1 ? 1 : (want > maxs ? maxs : want);
    }
    if (tc && x3 && g_gemm_pair && gemm_tc3_wanted(M, N, K))
      return gemm_tc3(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, st);
    return tc ? gemm_tc2(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, x3 ? 1 : 0, st)
              : gemm_simt(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, st);
  }
  // bf16 images of a weight inside the parameter arena (same element offset), when they are current
  void weight_images(const float* W, skf_gemm_epilogue* e) const {
    if (!P->w16_valid || W < P->params || W >= P->params + P->PL.total) return;
    const size_t off = (size_t)(W - P->params);
    e->b_hi16 = P->wsp<uint8_t>("w16h") + 2 * off;
    e->b_lo16 = P->wsp<uint8_t>("w16l") + 2 * off;
  }
  // Y[M,N] = act(X[M,K] * W[K,N] + bias) (+ resid)
  int fwd(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, int act,
          const float* resid, int rnd) const {
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    weight_images(W, &e);
    e.bias = bias; e.act = act; e.resid = resid; e.ld_resid = N; e.round_tf32 = rnd;
    return run(X, K, 1, W, N, 1, Y, N, M, N, K, e, P->c.gemm_precision >= 1);
  }
  // dX[M,K] = dY[M,N] * W[K,N]^T  (* (mask>0)) (+ resid)
  int dgrad(const float* dY, const float* W, float* dX, int M, int N, int K, const float* mask,
            const float* resid, int rnd) const {
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    weight_images(W, &e);
    e.mask = mask; e.ld_mask = K; e.resid = resid; e.ld_resid = K; e.round_tf32 = rnd;
    return run(dY, N, 1, W, 1, N, dX, K, M, K, N, e, P->c.gemm_precision >= 2);
  }
  // dW[K,N] += X[M,K]^T * dY[M,N] ; db[N] += colsum(dY)
  int wgrad(const float* X, const float* dY, float* dW, float* db, int M, int N, int K) const {
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.accumulate = 1;
    const bool x3 = P->c.gemm_precision == 2;
    // the pair kernel sums the columns of dY (bias gradient) from the tiles it stages anyway
    const bool fused = db && x3 && P->c.gemm_backend == 0 && g_gemm_pair && gemm_tc3_wanted(K, N, M) &&
                       gemm_tc2_supported(X, 1, K, dY, N, 1, dW, N, K, N, M, &e);
    if (fused) e.colsum = db;
    SKF_TRY(run(X, 1, K, dY, N, 1, dW, N, K, N, M, e, x3));
    if (db && !fused) SKF_TRY(skf_colsum(dY, N, db, M, N, st));
    return SKF_OK;
  }
  // fork: a front-end on the plan's side stream, ordered after everything issued so far on `st` (or *this when
  // the plan runs single-stream); merge: the side stream is done with `bufs` once it gets here
  int fork(Gemm* S) const {
    skf_plan* Q = const_cast<skf_plan*>(P);
    *S = *this;
    if (!Q->side_active || st == Q->side) return SKF_OK;
    cudaEvent_t ready = Q->next_event();
    SKF_CUDA(cudaEventRecord(ready, st));
    SKF_CUDA(cudaStreamWaitEvent(Q->side, ready, 0));
    S->st = Q->side;
    return SKF_OK;
  }
  int merge(const Gemm& S, std::initializer_list<const void*> bufs) const {
    skf_plan* Q = const_cast<skf_plan*>(P);
    if (S.st == st) return SKF_OK;
    cudaEvent_t done = Q->next_event();
    SKF_CUDA(cudaEventRecord(done, S.st));
    for (const void* b : bufs) Q->pending[b] = done;
    return SKF_OK;
  }
  // wgrad on the side stream
  int wgrad_async(const float* X, const float* dY, float* dW, float* db, int M, int N, int K) const {
    Gemm S{P, st};
    SKF_TRY(fork(&S));
    g_t3_oversubscribe = S.st != st && g_wgrad_units > 1;
    const int rc = S.wgrad(X, dY, dW, db, M, N, K);
    g_t3_oversubscribe = 0;
    SKF_TRY(rc);
    return merge(S, {dY});
  }
  // the main chain is about to overwrite (or read a result in) `buf`: wait until the side stream is done with it
  int before_write(const void* buf) const {
    skf_plan* Q = const_cast<skf_plan*>(P);
    auto it = Q->pending.find(buf);
    if (it == Q->pending.end()) return SKF_OK;
    SKF_CUDA(cudaStreamWaitEvent(st, it->second, 0));
    Q->pending.erase(it);
    return SKF_OK;
  }
  // rejoin: everything on the side stream completes before what follows on `st`
  int join() const {
    skf_plan* Q = const_cast<skf_plan*>(P);
    if (!Q->side_active) return SKF_OK;
    cudaEvent_t done = Q->next_event();
    SKF_CUDA(cudaEventRecord(done, Q->side));
    SKF_CUDA(cudaStreamWaitEvent(st, done, 0));
    Q->pending.clear();
    Q->side_active = false;
    return SKF_OK;
  }
};

// Dropout sites (one independent mask stream each): the reference's Dropout layers in call order
enum : uint32_t { SITE_ENC_X0 = 1, SITE_DEC_X0 = 2, SITE_ENC_L = 10, SITE_DEC_L = 100 };
static inline bool dropping(const skf_plan* P) { return P->training && P->c.dropout_rate > 0.f; }

// out = LN(resid + dropout(W-projection of `in`)) (transformer.py:217-218 etc.).  Without dropout the residual
// add is fused into the GEMM epilogue; with dropout the GEMM writes the bare sublayer output into `tmp` and the
// LayerNorm kernel applies mask and residual.  s receives the pre-LN sum (saved for backward).
static int proj_resid_ln(skf_plan* P, const Gemm& G, const float* in, const std::string& w, const std::string& b,
                         const std::string& ln, const float* resid, float* tmp, float* s, float* out, float* mean,
                         float* rstd, int M, int N, int K, uint32_t site, cudaStream_t st) {
  const int rnd = P->round_ops ? 1 : 0;
  if (!dropping(P)) {
    SKF_TRY(G.fwd(in, P->w(w), P->pw(b), s, M, N, K, SKF_ACT_NONE, resid, 0));
    return skf_add_layernorm_fwd(s, nullptr, nullptr, P->pw(ln + ".g"), P->pw(ln + ".b"), nullptr, out, mean, rstd, M,
                                 N, rnd, st);
  }
  SKF_TRY(G.fwd(in, P->w(w), P->pw(b), tmp, M, N, K, SKF_ACT_NONE, nullptr, 0));
  return skf_add_dropout_layernorm_fwd(resid, tmp, nullptr, P->pw(ln + ".g"), P->pw(ln + ".b"), s, out, mean, rstd, M,
                                       N, P->c.dropout_rate, P->c.dropout_seed, site, P->wsp<int64_t>("step"), st);
}
// backward of the LN of proj_resid_ln: ds -> gres (gradient of the residual input); returns in *gsub the
// gradient of the sublayer output (== gres without dropout, the masked copy in gdrop with dropout)
static int ln_bwd_any(skf_plan* P, const float* dy, const std::string& pfx, const std::string& idx, float* gres,
                      float* gdrop, const float** gsub, int M, int d, uint32_t site, cudaStream_t st) {
  const int rnd = P->round_ops ? 1 : 0;
  const std::string ln = pfx + ".ln" + idx;
  if (!dropping(P)) {
    *gsub = gres;
    return skf_layernorm_bwd(dy, P->wsf(pfx + ".s" + idx), P->wsf(pfx + ".mean" + idx), P->wsf(pfx + ".rstd" + idx),
                             P->pw(ln + ".g"), gres, P->g(ln + ".g"), P->g(ln + ".b"), M, d, rnd, st);
  }
  *gsub = gdrop;
  return skf_layernorm_dropout_bwd(dy, P->wsf(pfx + ".s" + idx), P->wsf(pfx + ".mean" + idx),
                                   P->wsf(pfx + ".rstd" + idx), P->pw(ln + ".g"), gres, gdrop, P->g(ln + ".g"),
                                   P->g(ln + ".b"), M, d, P->c.dropout_rate, P->c.dropout_seed, site,
                                   P->wsp<int64_t>("step"), st);
}
static int embed_dropout(skf_plan* P, float* x, int64_t n, uint32_t site, cudaStream_t st) {
  if (!dropping(P)) return SKF_OK;
  return skf_dropout(x, x, n, P->c.dropout_rate, P->c.dropout_seed, site, P->wsp<int64_t>("step"), st);
}

// img: workspace buffer for the operand images of this attention site (nullptr: none, e.g. cross attention)
static bool attn3_images(const skf_plan* P, const void* img, int Tq, int Tk) {
  return img != nullptr && P->training && (g_attn3 & 3) == 3 && (g_attn3_img != 0) && Tq == Tk &&
         P->c.gemm_precision >= 2;
}
static int attn_fwd_any(const skf_plan* P, const float* Q, int64_t ldq, const float* K, int64_t ldk,
                        const float* V, int64_t ldv, float* O, int64_t ldo, float* lse, const uint8_t* pad,
                        int Tq, int Tk, int mode, cudaStream_t st, void* img = nullptr) {
  const skf_cfg& c = P->c;
  if (c.attn_backend == 0 && (g_attn3 & 1) && c.gemm_precision >= 1 && attn3_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, c.d / c.H))
    return attn3_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, pad, c.B, c.H, Tq, Tk, c.d / c.H, mode,
                     attn3_images(P, img, Tq, Tk) ? img : nullptr, st);
  if (c.attn_backend == 0 && attn2_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, c.d / c.H))
    return attn2_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, pad, c.B, c.H, Tq, Tk, c.d / c.H, mode,
                     c.gemm_precision >= 1 ? 1 : 0, st);
  return attn_fwd_simt(Q, ldq, K, ldk, V, ldv, O, ldo, lse, pad, c.B, c.H, Tq, Tk, c.d / c.H, mode,
                       P->round_ops ? 1 : 0, st);
}
static int attn_bwd_any(const skf_plan* P, const float* Q, int64_t ldq, const float* K, int64_t ldk,
                        const float* V, int64_t ldv, const float* O, int64_t ldo, const float* dO, int64_t lddo,
                        const float* lse, float* delta, float* dQ, int64_t lddq, float* dK, int64_t lddk,
                        float* dV, int64_t lddv, const uint8_t* pad, int Tq, int Tk, int mode, cudaStream_t st,
                        const void* img = nullptr) {
  const skf_cfg& c = P->c;
  if (c.attn_backend == 0 && (g_attn3 & 2) && c.gemm_precision >= 2 && attn3_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, c.d / c.H)) {
    const bool direct = attn3_images(P, img, Tq, Tk);
    return attn3_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, pad, c.B, c.H,
                     Tq, Tk, c.d / c.H, mode, direct ? img : nullptr, direct ? P->wsp<void>("g.img_do") : nullptr, st);
  }
  if (c.attn_backend == 0 && attn2_supported(Q, ldq, K, ldk, V, ldv, Tq, Tk, c.d / c.H))
    return attn2_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, pad, c.B,
                     c.H, Tq, Tk, c.d / c.H, mode, c.gemm_precision >= 2 ? 1 : 0, st);
  return attn_bwd_simt(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, pad,
                       c.B, c.H, Tq, Tk, c.d / c.H, mode, P->round_ops ? 1 : 0, st);
}

// The decoder's cross-attention keys / values are Dense(2d) of pre[b,t,:] = emb[b,:] w[t] + beta[t]
// (DenseExpander, transformer.py:370-376), i.e. w[t] (emb Wkv)[b,:] + beta[t] colsum(Wkv) + bias: a [B,E]x[E,2d]
// product and an outer-product kernel replace the [B T,E]x[E,2d] GEMM, and a row reduction replaces its dgrad /
// wgrad and the expander's backward (skf_kv_expand_fwd / _bwd).  Same arithmetic up to fp32 summation order.
// Level 2 goes one step further: with such keys and values a query's logits are a w[t] + c beta[t] + const and its
// output is a combination of three vectors, so the second attention block needs no K / V / score tensors at all
// (skf_xattn_fold_fwd / _bwd); level 1 materialises K / V and runs the regular attention kernels.
static int kv_fold(const skf_plan* P) {
  if (!(g_kv_fold && P->c.u > 0 && !P->round_ops && (2 * P->c.d) % 128 == 0)) return 0;
  return g_kv_fold;
}

// ---- forward -----------------------------------------------------------------------------------
static int refresh_compute_weights(skf_plan* P, cudaStream_t st) {
  if (!P->round_ops) {
    P->wc = P->params;
    // bf16 images of the weights for the pair GEMM (Gemm::weight_images): one pass over the arena per forward
    P->w16_valid = false;
    if (g_gemm_wsplit && g_gemm_mix == 2 && g_gemm_pair && P->c.gemm_backend == 0 && P->c.gemm_precision >= 1) {
      SKF_TRY(split_bf16(P->params, P->wsp<uint8_t>("w16h"), P->wsp<uint8_t>("w16l"), P->PL.total, st));
      P->w16_valid = true;
    }
    return SKF_OK;
  }
  P->w16_valid = false;
  float* dst = P->wsf("wc");
  round_copy_k<<<cdiv(P->PL.total, 256), 256, 0, st>>>(P->params, dst, P->PL.total);
  SKF_LAUNCH_CHECK();
  P->wc = dst;
  return SKF_OK;
}

// pad flags + input embedding of `rows` positions per sample of the batch `in` ([B,T] int64 tokens, or
// [B,T,5] float stroke-5 rows when cfg.continuous); prefix = "enc" / "dec"
static int embed_any(skf_plan* P, const void* in, const std::string& prefix, uint8_t* pad, float* x, int rows,
                     cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int rnd = P->round_ops ? 1 : 0;
  if (c.continuous) {
    const float* xs = reinterpret_cast<const float*>(in);
    SKF_TRY(skf_pad_flags_cont(xs, (int64_t)c.T * 5, pad, c.B, rows, st));
    return skf_embed_cont_fwd(xs, (int64_t)c.T * 5, P->pw(prefix + ".emb.k"), P->pw(prefix + ".emb.b"), P->wsf("pe"),
                              x, c.B, rows, c.d, rnd, st);
  }
  const int64_t* tok = reinterpret_cast<const int64_t*>(in);
  SKF_TRY(skf_pad_flags(tok, c.T, pad, c.B, rows, st));
  return skf_embed_fwd(tok, c.T, P->pw(prefix + ".emb"), P->wsf("pe"), x, c.B, rows, c.d, c.V, rnd, st);
}
static int embed_any_bwd(skf_plan* P, const void* in, const std::string& prefix, const float* dx, int rows,
                         cudaStream_t st) {
  const skf_cfg& c = P->c;
  if (c.continuous)
    return skf_embed_cont_bwd(reinterpret_cast<const float*>(in), (int64_t)c.T * 5, dx, P->g(prefix + ".emb.k"),
                              P->g(prefix + ".emb.b"), c.B, rows, c.d, st);
  return skf_embed_bwd(reinterpret_cast<const int64_t*>(in), c.T, dx, P->g(prefix + ".emb"), c.B, rows, c.d, c.V, st);
}

// Transformer.classify_from_embedding (models/sketchformer.py:183-199): class_buffer_layers x
// (Dense(lowerdim, relu) + Dropout(class_dropout)), then the Dense(n_classes) logits.  All of it is B rows
// wide: exact-fp32 SIMT GEMMs (C = 345 columns are not 16-byte multiples).
enum : uint32_t { SITE_CBUF = 1000 };
static int classifier_fwd(skf_plan* P, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, Cp = (c.C + 3) & ~3;
  const float* fc = P->wsf("embedding");
  int fin = emb_dim(c);
  skf_gemm_epilogue e;
  for (int j = 0; j < c.class_buffer_layers; j++) {
    const std::string n = "cbuf" + std::to_string(j);
    memset(&e, 0, sizeof(e));
    e.bias = P->pw(n + ".b");
    e.act = SKF_ACT_RELU;
    float* h = P->wsf(n + ".h");
    SKF_TRY(gemm_simt(fc, fin, 1, P->pw(n + ".k"), c.u, 1, h, c.u, B, c.u, fin, &e, st));
    fc = h;
    if (P->training && c.class_dropout > 0.f) {
      float* hd = P->wsf(n + ".hd");
      SKF_TRY(skf_dropout(h, hd, (int64_t)B * c.u, c.class_dropout, c.dropout_seed, SITE_CBUF + j,
                          P->wsp<int64_t>("step"), st));
      fc = hd;
    }
    fin = c.u;
  }
  memset(&e, 0, sizeof(e));
  e.bias = P->pw("cls.b");
  return gemm_simt(fc, fin, 1, P->pw("cls.k"), c.C, 1, P->wsf("cls.logits"), Cp, B, c.C, fin, &e, st);
}
// gradient of the classification branch w.r.t. the embedding: gEmb (+)= ...
static int classifier_bwd(skf_plan* P, float* gEmb, bool accumulate, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, Cp = (c.C + 3) & ~3, nb = c.class_buffer_layers;
  const bool drop = P->training && c.class_dropout > 0.f;
  auto layer_in = [&](int j) -> const float* {       // input of buffer layer j / of the classifier (j == nb)
    if (j == 0) return P->wsf("embedding");
    return P->wsf("cbuf" + std::to_string(j - 1) + (drop ? ".hd" : ".h"));
  };
  const float* dl = P->wsf("cls.dlogits");
  int fin = nb ? c.u : emb_dim(c);
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.accumulate = 1;
  SKF_TRY(gemm_simt(layer_in(nb), 1, fin, dl, Cp, 1, P->g("cls.k"), c.C, fin, c.C, B, &e, st));
  SKF_TRY(skf_colsum(dl, Cp, P->g("cls.b"), B, c.C, st));
  const float* dy = dl;              // gradient of the current layer's output, [B, ncols] with row stride ldy
  int ncols = c.C, ldy = Cp;
  std::string wname = "cls.k";
  float* buf[2] = {P->wsf("g.cb0"), P->wsf("g.cb1")};
  for (int j = nb - 1; j >= -1; j--) {
    // d(input of layer j+1) = dy * W^T; for a buffer layer's output also through Dropout and ReLU
    const int kin = j >= 0 ? c.u : emb_dim(c);
    float* dst = j >= 0 ? buf[j & 1] : gEmb;
    memset(&e, 0, sizeof(e));
    if (j >= 0) { e.mask = P->wsf("cbuf" + std::to_string(j) + ".h"); e.ld_mask = c.u; }
    else if (accumulate) { e.resid = gEmb; e.ld_resid = kin; }
    SKF_TRY(gemm_simt(dy, ldy, 1, P->pw(wname), 1, ncols, dst, kin, B, kin, ncols, &e, st));
    if (j < 0) break;
    if (drop)
      SKF_TRY(skf_dropout(dst, dst, (int64_t)B * c.u, c.class_dropout, c.dropout_seed, SITE_CBUF + j,
                          P->wsp<int64_t>("step"), st));
    const std::string n = "cbuf" + std::to_string(j);
    const int jin = j == 0 ? emb_dim(c) : c.u;
    memset(&e, 0, sizeof(e));
    e.accumulate = 1;
    SKF_TRY(gemm_simt(layer_in(j), 1, jin, dst, c.u, 1, P->g(n + ".k"), c.u, jin, c.u, B, &e, st));
    SKF_TRY(skf_colsum(dst, c.u, P->g(n + ".b"), B, c.u, st));
    dy = dst; ncols = c.u; ldy = c.u; wname = n + ".k";
  }
  return SKF_OK;
}

static int forward_encoder(skf_plan* P, const void* tok, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, d = c.d, f = c.f, Me = B * T;
  const int rnd = P->round_ops ? 1 : 0;
  Gemm G{P, st};
  uint8_t* enc_pad = P->wsp<uint8_t>("enc_pad");
  float* x = P->wsf("enc.x0");
  SKF_TRY(embed_any(P, tok, "enc", enc_pad, x, T, st));
  SKF_TRY(embed_dropout(P, x, (int64_t)Me * d, SITE_ENC_X0, st));            // transformer.py:296
  float* tmp = P->wsf("g.c");                                                // forward scratch (free until backward)
  for (int i = 0; i < c.L; i++) {
    std::string p = "enc.l" + std::to_string(i);
    float* qkv = P->wsf(p + ".qkv");
    SKF_TRY(G.fwd(x, P->w(p + ".mha.wqkv.k"), P->pw(p + ".mha.wqkv.b"), qkv, Me, 3 * d, d, SKF_ACT_NONE, nullptr, 0));
    float* ctx = P->wsf(p + ".ctx");
    SKF_TRY(attn_fwd_any(P, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, ctx, d, P->wsf(p + ".lse"), enc_pad,
                         T, T, 1, st, P->wsp<void>(p + ".img")));
    float* out1 = P->wsf(p + ".out1");
    SKF_TRY(proj_resid_ln(P, G, ctx, p + ".mha.wo.k", p + ".mha.wo.b", p + ".ln1", x, tmp, P->wsf(p + ".s1"), out1,
                          P->wsf(p + ".mean1"), P->wsf(p + ".rstd1"), Me, d, d, SITE_ENC_L + 2 * i, st));
    float* hff = P->wsf(p + ".hff");
    SKF_TRY(G.fwd(out1, P->w(p + ".ffn.w1.k"), P->pw(p + ".ffn.w1.b"), hff, Me, f, d, SKF_ACT_RELU, nullptr, 1));
    float* out2 = P->wsf(p + ".out2");
    SKF_TRY(proj_resid_ln(P, G, hff, p + ".ffn.w2.k", p + ".ffn.w2.b", p + ".ln2", out1, tmp, P->wsf(p + ".s2"), out2,
                          P->wsf(p + ".mean2"), P->wsf(p + ".rstd2"), Me, d, f, SITE_ENC_L + 2 * i + 1, st));
    x = out2;
  }
  if (c.u == 0) return SKF_OK;      // lowerdim=0: embedding = enc_output (sketchformer.py:157-158)
  // bottleneck (SelfAttnV1.call transformer.py:61-74 / SelfAttnV2.call 116-131)
  const int UP = pool_dim(c);
  float* u = P->wsf("pool.u");
  SKF_TRY(G.fwd(x, P->w("pool.W"), P->pw("pool.b"), u, Me, UP, d, SKF_ACT_TANH, nullptr, 0));
  float* pooled = pool_v2(c) ? P->wsf("pool.o") : P->wsf("embedding");
  SKF_TRY(skf_pool_fwd(x, u, P->pw("pool.V"), P->wsf("pool.a"), pooled, B, T, d, UP, 0, st));
  if (pool_v2(c)) {
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = P->pw("pool.emb.b");
    SKF_TRY(gemm_simt(pooled, d, 1, P->pw("pool.emb.k"), c.u, 1, P->wsf("embedding"), c.u, B, c.u, d, &e, st));
  }
  if (c.do_classification) SKF_TRY(classifier_fwd(P, st));
  return SKF_OK;
}

static int forward_decoder(skf_plan* P, const void* tok, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, Td = c.T - 1, d = c.d, f = c.f, Me = B * T, Md = B * Td;
  const int rnd = P->round_ops ? 1 : 0;
  Gemm G{P, st};
  const int E = emb_dim(c);
  float* pre = P->wsf("pre_decoder");
  const int fold = kv_fold(P);
  if (fold) {
    for (int i = 0; i < c.L; i++) {
      std::string p = "dec.l" + std::to_string(i);
      float *U = P->wsf(p + ".kvU"), *sW = P->wsf(p + ".kvs");
      SKF_TRY(G.fwd(P->wsf("embedding"), P->w(p + ".mha2.wkv.k"), nullptr, U, B, 2 * d, E, SKF_ACT_NONE, nullptr, 0));
      SKF_CUDA(cudaMemsetAsync(sW, 0, (size_t)2 * d * sizeof(float), st));
      SKF_TRY(skf_colsum(P->pw(p + ".mha2.wkv.k"), 2 * d, sW, E, 2 * d, st));
      if (fold == 1)
        SKF_TRY(skf_kv_expand_fwd(U, sW, P->pw(p + ".mha2.wkv.b"), P->pw("exp.k"), P->pw("exp.b"), P->wsf(p + ".kv2"),
                                  B, T, 2 * d, st));
    }
  } else if (c.u > 0) {
    SKF_TRY(skf_expand_fwd(P->wsf("embedding"), P->pw("exp.k"), P->pw("exp.b"), pre, B, T, E, rnd, st));
  } else {
    pre = P->wsf("enc.l" + std::to_string(c.L - 1) + ".out2");      // sketchformer.py:175-176
  }
  uint8_t* dec_pad = P->wsp<uint8_t>("dec_pad");
  float* x = P->wsf("dec.x0");
  SKF_TRY(embed_any(P, tok, "dec", dec_pad, x, Td, st));   // tar_inp = tar[:, :-1]
  SKF_TRY(embed_dropout(P, x, (int64_t)Md * d, SITE_DEC_X0, st));            // transformer.py:334
  float* tmp = P->wsf("g.c");
  const uint8_t* cross_pad = c.blind_decoder_mask ? nullptr : P->wsp<uint8_t>("enc_pad");
  const int cross_mode = c.blind_decoder_mask ? 0 : 1;
  for (int i = 0; i < c.L; i++) {
    std::string p = "dec.l" + std::to_string(i);
    float* qkv = P->wsf(p + ".qkv");
    SKF_TRY(G.fwd(x, P->w(p + ".mha1.wqkv.k"), P->pw(p + ".mha1.wqkv.b"), qkv, Md, 3 * d, d, SKF_ACT_NONE, nullptr, 0));
    float* ctx1 = P->wsf(p + ".ctx1");
    SKF_TRY(attn_fwd_any(P, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, ctx1, d, P->wsf(p + ".lse1"), dec_pad,
                         Td, Td, 2, st, P->wsp<void>(p + ".img1")));
    float* out1 = P->wsf(p + ".out1");
    SKF_TRY(proj_resid_ln(P, G, ctx1, p + ".mha1.wo.k", p + ".mha1.wo.b", p + ".ln1", x, tmp, P->wsf(p + ".s1"), out1,
                          P->wsf(p + ".mean1"), P->wsf(p + ".rstd1"), Md, d, d, SITE_DEC_L + 3 * i, st));
    float* q2 = P->wsf(p + ".q2");
    SKF_TRY(G.fwd(out1, P->w(p + ".mha2.wq.k"), P->pw(p + ".mha2.wq.b"), q2, Md, d, d, SKF_ACT_NONE, nullptr, 0));
    float* kv2 = P->wsf(p + ".kv2");
    if (!fold)
      SKF_TRY(G.fwd(pre, P->w(p + ".mha2.wkv.k"), P->pw(p + ".mha2.wkv.b"), kv2, Me, 2 * d, E, SKF_ACT_NONE, nullptr, 0));
    float* ctx2 = P->wsf(p + ".ctx2");
    if (fold == 2)
      SKF_TRY(skf_xattn_fold_fwd(q2, d, P->wsf(p + ".kvU"), P->wsf(p + ".kvs"), P->pw(p + ".mha2.wkv.b"), P->pw("exp.k"),
                                 P->pw("exp.b"), cross_pad, ctx2, d, B, c.H, Td, T, d / c.H, st));
    else
      SKF_TRY(attn_fwd_any(P, q2, d, kv2, 2 * d, kv2 + d, 2 * d, ctx2, d, P->wsf(p + ".lse2"), cross_pad, Td, T,
                           cross_mode, st));
    float* out2 = P->wsf(p + ".out2");
    SKF_TRY(proj_resid_ln(P, G, ctx2, p + ".mha2.wo.k", p + ".mha2.wo.b", p + ".ln2", out1, tmp, P->wsf(p + ".s2"), out2,
                          P->wsf(p + ".mean2"), P->wsf(p + ".rstd2"), Md, d, d, SITE_DEC_L + 3 * i + 1, st));
    float* hff = P->wsf(p + ".hff");
    SKF_TRY(G.fwd(out2, P->w(p + ".ffn.w1.k"), P->pw(p + ".ffn.w1.b"), hff, Md, f, d, SKF_ACT_RELU, nullptr, 1));
    float* out3 = P->wsf(p + ".out3");
    SKF_TRY(proj_resid_ln(P, G, hff, p + ".ffn.w2.k", p + ".ffn.w2.b", p + ".ln3", out2, tmp, P->wsf(p + ".s3"), out3,
                          P->wsf(p + ".mean3"), P->wsf(p + ".rstd3"), Md, d, f, SITE_DEC_L + 3 * i + 2, st));
    x = out3;
  }
  if (c.continuous) {            // Dense(5), sketchformer.py:86 (5 columns: exact-fp32 SIMT GEMM, ld 8)
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = P->pw("out.b");
    return gemm_simt(x, d, 1, P->pw("out.k"), 5, 1, P->wsf("recon.logits"), 8, Md, 5, d, &e, st);
  }
  SKF_TRY(G.fwd(x, P->w("out.k"), P->pw("out.b"), P->wsf("recon.logits"), Md, c.V, d, SKF_ACT_NONE, nullptr, 0));
  return SKF_OK;
}

// ---- losses ------------------------------------------------------------------------------------
static int losses(skf_plan* P, const void* in, const int64_t* labels, bool want_grad, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, Td = T - 1;
  const int64_t* tok = reinterpret_cast<const int64_t*>(in);
  double* ss = P->wsp<double>("stats_step");
  SKF_CUDA(cudaMemsetAsync(ss, 0, 8 * sizeof(double), st));
  if (c.do_reconstruction && c.continuous) {
    // tar_real = data[:, 1:, :]; the gradient overwrites the predictions in place
    const float* x = reinterpret_cast<const float*>(in);
    float* pr = P->wsf("recon.logits");
    SKF_TRY(skf_cont_loss(pr, 8, x + 5, (int64_t)T * 5, B, Td, c.recon_weight, want_grad ? pr : nullptr, 8, ss + 4, st));
  } else if (c.do_reconstruction) {
    float* lg = P->wsf("recon.logits");
    float scale = c.recon_weight / (float)((double)B * Td);
    SKF_TRY(skf_token_ce(lg, c.V, tok + 1, T, B, Td, c.V, scale, want_grad ? lg : nullptr, c.V, ss + 0,
                         P->wsp<int32_t>("recon.argmax"), P->round_ops ? 1 : 0, st));
  }
  if (c.do_classification) {
    int Cp = (c.C + 3) & ~3;
    float scale = c.class_weight / (float)B;
    SKF_TRY(skf_class_ce(P->wsf("cls.logits"), Cp, labels, B, c.C, scale, want_grad ? P->wsf("cls.dlogits") : nullptr,
                         Cp, labels ? ss + 2 : nullptr, P->wsp<int32_t>("cls.argmax"), P->wsf("cls.probs"), st));
  }
  if (labels) {
    stats_finalize_k<<<1, 1, 0, st>>>(ss, P->wsp<double>("stats_run"), (double)B * Td, (double)B,
                                      (double)c.recon_weight, (double)c.class_weight, c.do_reconstruction,
                                      c.do_classification, c.continuous);
    SKF_LAUNCH_CHECK();
  }
  return SKF_OK;
}

// ---- data-parallel exchange ---------------------------------------------------------------------
// Stage `stage` of the backward pass is complete on the main chain (everything issued so far on `st`) and on the
// weight-gradient side stream: all-reduce every bucket that stage finishes, on the communication stream.
static int buckets_ready(skf_plan* P, int stage, cudaStream_t st) {
  if (!P->comm || P->c.world_size <= 1) return SKF_OK;
  while (P->buckets_sent < P->buckets.n && P->buckets.stage[P->buckets_sent] <= stage) {
    const int k = P->buckets_sent++;
    cudaEvent_t e_main = P->next_event();
    SKF_CUDA(cudaEventRecord(e_main, st));
    SKF_CUDA(cudaStreamWaitEvent(P->comm_stream, e_main, 0));
    if (P->side_active && P->side != st) {
      cudaEvent_t e_side = P->next_event();
      SKF_CUDA(cudaEventRecord(e_side, P->side));
      SKF_CUDA(cudaStreamWaitEvent(P->comm_stream, e_side, 0));
    }
    SKF_TRY(allreduce_sum_f32(P->grads + P->buckets.lo[k], P->buckets.hi[k] - P->buckets.lo[k], P->comm, P->comm_stream));
  }
  return SKF_OK;
}
// the main chain continues (optimizer) only after every bucket has been reduced
static int buckets_join(skf_plan* P, cudaStream_t st) {
  if (!P->comm || P->c.world_size <= 1) return SKF_OK;
  SKF_TRY(buckets_ready(P, 3, st));
  cudaEvent_t done = P->next_event();
  SKF_CUDA(cudaEventRecord(done, P->comm_stream));
  SKF_CUDA(cudaStreamWaitEvent(st, done, 0));
  return SKF_OK;
}

// ---- backward ----------------------------------------------------------------------------------
static int backward(skf_plan* P, const void* tok, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, Td = T - 1, d = c.d, f = c.f, Me = B * T, Md = B * Td;
  const int rnd = P->round_ops ? 1 : 0;
  Gemm G{P, st};
  SKF_CUDA(cudaMemsetAsync(P->grads, 0, (size_t)P->PL.total * sizeof(float), st));
  float *gA = P->wsf("g.a"), *gC = P->wsf("g.c"), *gH = P->wsf("g.h");
  // gradient of the pre-LN sum (gBs) and its Dropout-masked copy (gDs), one pair per LayerNorm site of a layer:
  // the weight-gradient GEMM that reads a pair runs on the side stream and may still be in flight when the next
  // site's LayerNorm backward runs, so a pair is only reused one layer later (Gemm::before_write orders that)
  float* gBs[3] = {P->wsf("g.b"), P->wsf("g.b1"), P->wsf("g.b2")};
  float* gDs[3] = {P->wsf("g.d"), P->wsf("g.d1"), P->wsf("g.d2")};
  P->pending.clear();
  P->side_active = g_wgrad_overlap && P->side != nullptr;
  P->buckets_sent = 0;
  float *gQKV = P->wsf("g.qkv"), *gQ2 = P->wsf("g.q2"), *gKV2 = P->wsf("g.kv2"), *gPre = P->wsf("g.pre");
  float *gEmb = P->wsf("g.emb"), *gU = P->wsf("g.u"), *delta = P->wsf("g.delta");
  const float* gS = nullptr;       // gradient of the current sublayer output (behind its Dropout)
  float *gB = gBs[0], *gD = gDs[0];
  const int fold = kv_fold(P);
  const uint8_t* enc_pad = P->wsp<uint8_t>("enc_pad");
  bool have_demb = false;

  if (c.do_reconstruction) {
    const uint8_t* dec_pad = P->wsp<uint8_t>("dec_pad");
    const uint8_t* cross_pad = c.blind_decoder_mask ? nullptr : enc_pad;
    const int cross_mode = c.blind_decoder_mask ? 0 : 1;
    float* pre = c.u > 0 ? P->wsf("pre_decoder") : P->wsf("enc.l" + std::to_string(c.L - 1) + ".out2");
    float* dlog = P->wsf("recon.logits");  // holds dlogits after losses()
    std::string last = "dec.l" + std::to_string(c.L - 1);
    if (c.continuous) {
      skf_gemm_epilogue e;
      memset(&e, 0, sizeof(e));
      e.accumulate = 1;
      e.split_k = 64;
      SKF_TRY(gemm_simt(P->wsf(last + ".out3"), 1, d, dlog, 8, 1, P->g("out.k"), 5, d, 5, Md, &e, st));
      SKF_TRY(skf_colsum(dlog, 8, P->g("out.b"), Md, 5, st));
      memset(&e, 0, sizeof(e));
      SKF_TRY(gemm_simt(dlog, 8, 1, P->pw("out.k"), 1, 5, gA, d, Md, d, 5, &e, st));
    } else {
      SKF_TRY(G.wgrad_async(P->wsf(last + ".out3"), dlog, P->g("out.k"), P->g("out.b"), Md, c.V, d));
      SKF_TRY(G.dgrad(dlog, P->w("out.k"), gA, Md, c.V, d, nullptr, nullptr, 0));
    }
    for (int i = c.L - 1; i >= 0; i--) {
      std::string p = "dec.l" + std::to_string(i);
      const float* xin = i == 0 ? P->wsf("dec.x0") : P->wsf("dec.l" + std::to_string(i - 1) + ".out3");
      // LN3 + FFN
      gB = gBs[2]; gD = gDs[2];
      SKF_TRY(G.before_write(gB)); SKF_TRY(G.before_write(gD));
      SKF_TRY(ln_bwd_any(P, gA, p, "3", gB, gD, &gS, Md, d, SITE_DEC_L + 3 * i + 2, st));
      SKF_TRY(G.wgrad_async(P->wsf(p + ".hff"), gS, P->g(p + ".ffn.w2.k"), P->g(p + ".ffn.w2.b"), Md, d, f));
      SKF_TRY(G.before_write(gH));
      SKF_TRY(G.dgrad(gS, P->w(p + ".ffn.w2.k"), gH, Md, d, f, P->wsf(p + ".hff"), nullptr, 1));
      SKF_TRY(G.wgrad_async(P->wsf(p + ".out2"), gH, P->g(p + ".ffn.w1.k"), P->g(p + ".ffn.w1.b"), Md, f, d));
      SKF_TRY(G.dgrad(gH, P->w(p + ".ffn.w1.k"), gA, Md, f, d, nullptr, gB, 0));
      // LN2 + cross attention
      gB = gBs[1]; gD = gDs[1];
      SKF_TRY(G.before_write(gB)); SKF_TRY(G.before_write(gD));
      SKF_TRY(ln_bwd_any(P, gA, p, "2", gB, gD, &gS, Md, d, SITE_DEC_L + 3 * i + 1, st));
      SKF_TRY(G.wgrad_async(P->wsf(p + ".ctx2"), gS, P->g(p + ".mha2.wo.k"), P->g(p + ".mha2.wo.b"), Md, d, d));
      SKF_TRY(G.dgrad(gS, P->w(p + ".mha2.wo.k"), gC, Md, d, d, nullptr, nullptr, 0));
      float* kv2 = P->wsf(p + ".kv2");
      float *gUkv = P->wsf("g.kvU"), *gskv = P->wsf("g.kvs");
      SKF_TRY(G.before_write(gQ2)); SKF_TRY(G.before_write(gKV2));
      if (fold == 2) {
        SKF_TRY(G.before_write(gUkv));           // the side stream is done with the previous layer's gU / gs
        SKF_CUDA(cudaMemsetAsync(gskv, 0, (size_t)2 * d * sizeof(float), st));
        SKF_TRY(skf_xattn_fold_bwd(P->wsf(p + ".q2"), d, P->wsf(p + ".kvU"), P->wsf(p + ".kvs"), P->pw(p + ".mha2.wkv.b"),
                                   P->pw("exp.k"), P->pw("exp.b"), cross_pad, gC, d, gQ2, d, gUkv, gskv,
                                   P->g(p + ".mha2.wkv.b"), P->g("exp.k"), P->g("exp.b"), B, c.H, Td, T, d / c.H, st));
      } else {
        SKF_TRY(attn_bwd_any(P, P->wsf(p + ".q2"), d, kv2, 2 * d, kv2 + d, 2 * d, P->wsf(p + ".ctx2"), d, gC, d,
                             P->wsf(p + ".lse2"), delta, gQ2, d, gKV2, 2 * d, gKV2 + d, 2 * d, cross_pad, Td, T,
                             cross_mode, st));
      }
      SKF_TRY(G.wgrad_async(P->wsf(p + ".out1"), gQ2, P->g(p + ".mha2.wq.k"), P->g(p + ".mha2.wq.b"), Md, d, d));
      if (fold) {
        // off the critical path (only the bottleneck gradient at the end of the decoder needs it): side stream
        Gemm S{P, st};
        SKF_TRY(G.fork(&S));
        const int E = emb_dim(c);
        float *gU = gUkv, *gs = gskv;
        if (fold == 1) {
          SKF_CUDA(cudaMemsetAsync(gU, 0, (size_t)B * 2 * d * sizeof(float), S.st));
          SKF_CUDA(cudaMemsetAsync(gs, 0, (size_t)2 * d * sizeof(float), S.st));
          SKF_TRY(skf_kv_expand_bwd(gKV2, P->wsf(p + ".kvU"), P->wsf(p + ".kvs"), P->pw("exp.k"), P->pw("exp.b"), gU, gs,
                                    P->g(p + ".mha2.wkv.b"), P->g("exp.k"), P->g("exp.b"), B, T, 2 * d, S.st));
        }
        SKF_TRY(S.wgrad(P->wsf("embedding"), gU, P->g(p + ".mha2.wkv.k"), nullptr, B, 2 * d, E));
        SKF_TRY(skf_add_row_broadcast(P->g(p + ".mha2.wkv.k"), gs, E, 2 * d, S.st));
        SKF_TRY(S.dgrad(gU, P->w(p + ".mha2.wkv.k"), gEmb, B, 2 * d, E, nullptr, i == c.L - 1 ? nullptr : gEmb, 0));
        SKF_TRY(G.merge(S, {gKV2, gEmb, gU}));
      } else {
        SKF_TRY(G.wgrad_async(pre, gKV2, P->g(p + ".mha2.wkv.k"), P->g(p + ".mha2.wkv.b"), Me, 2 * d, emb_dim(c)));
        SKF_TRY(G.dgrad(gKV2, P->w(p + ".mha2.wkv.k"), gPre, Me, 2 * d, emb_dim(c), nullptr,
                        i == c.L - 1 ? nullptr : gPre, 0));
      }
      SKF_TRY(G.dgrad(gQ2, P->w(p + ".mha2.wq.k"), gA, Md, d, d, nullptr, gB, 0));
      // LN1 + masked self attention
      gB = gBs[0]; gD = gDs[0];
      SKF_TRY(G.before_write(gB)); SKF_TRY(G.before_write(gD));
      SKF_TRY(ln_bwd_any(P, gA, p, "1", gB, gD, &gS, Md, d, SITE_DEC_L + 3 * i, st));
      SKF_TRY(G.wgrad_async(P->wsf(p + ".ctx1"), gS, P->g(p + ".mha1.wo.k"), P->g(p + ".mha1.wo.b"), Md, d, d));
      SKF_TRY(G.dgrad(gS, P->w(p + ".mha1.wo.k"), gC, Md, d, d, nullptr, nullptr, 0));
      float* qkv = P->wsf(p + ".qkv");
      SKF_TRY(G.before_write(gQKV));
      SKF_TRY(attn_bwd_any(P, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, P->wsf(p + ".ctx1"), d, gC, d,
                           P->wsf(p + ".lse1"), delta, gQKV, 3 * d, gQKV + d, 3 * d, gQKV + 2 * d, 3 * d, dec_pad,
                           Td, Td, 2, st, P->wsp<void>(p + ".img1")));
      SKF_TRY(G.wgrad_async(xin, gQKV, P->g(p + ".mha1.wqkv.k"), P->g(p + ".mha1.wqkv.b"), Md, 3 * d, d));
      SKF_TRY(G.dgrad(gQKV, P->w(p + ".mha1.wqkv.k"), gA, Md, 3 * d, d, nullptr, gB, 0));
      if (i == c.L / 2 && i > 0) SKF_TRY(buckets_ready(P, 0, st));    // decoder layers L/2.. and the output layer are final
    }
    SKF_TRY(embed_dropout(P, gA, (int64_t)Md * d, SITE_DEC_X0, st));
    SKF_TRY(embed_any_bwd(P, tok, "dec", gA, Td, st));
    if (fold) {
      SKF_TRY(G.before_write(gEmb));          // the side stream has finished accumulating the bottleneck gradient
      have_demb = true;
    } else if (c.u > 0) {
      SKF_TRY(skf_expand_bwd(gPre, P->wsf("embedding"), P->pw("exp.k"), gEmb, 0, P->g("exp.k"),
                             P->g("exp.b"), B, T, emb_dim(c), st));
      have_demb = true;
    }
  }
  if (c.do_reconstruction) SKF_TRY(buckets_ready(P, 1, st));          // decoder, its embedding and the expander are final
  if (c.do_classification) {
    SKF_TRY(classifier_bwd(P, gEmb, have_demb, st));
    have_demb = true;
  }
  if (c.u == 0) {
    // lowerdim=0: the encoder output reaches the loss only as the decoder's cross keys / values
    SKF_CUDA(cudaMemcpyAsync(gA, gPre, (size_t)Me * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if (!have_demb) {
    SKF_TRY(G.join());
    return buckets_join(P, st);
  } else {
    // bottleneck backward
    std::string last = "enc.l" + std::to_string(c.L - 1);
    const float* x = P->wsf(last + ".out2");
    const int UP = pool_dim(c);
    const float* gPool = gEmb;          // gradient of the pooled vector [B, d]
    if (pool_v2(c)) {                   // through SelfAttnV2's Dense(lowerdim)
      float* gO = P->wsf("g.po");
      skf_gemm_epilogue e;
      memset(&e, 0, sizeof(e));
      e.accumulate = 1;
      SKF_TRY(gemm_simt(P->wsf("pool.o"), 1, d, gEmb, c.u, 1, P->g("pool.emb.k"), c.u, d, c.u, B, &e, st));
      SKF_TRY(skf_colsum(gEmb, c.u, P->g("pool.emb.b"), B, c.u, st));
      memset(&e, 0, sizeof(e));
      SKF_TRY(gemm_simt(gEmb, c.u, 1, P->pw("pool.emb.k"), 1, c.u, gO, d, B, d, c.u, &e, st));
      gPool = gO;
    }
    SKF_TRY(skf_pool_bwd(x, P->wsf("pool.u"), P->pw("pool.V"), P->wsf("pool.a"), gPool, gA, 0, gU,
                         P->g("pool.V"), B, T, d, UP, rnd, st));
    SKF_TRY(G.wgrad_async(x, gU, P->g("pool.W"), P->g("pool.b"), Me, UP, d));
    SKF_TRY(G.dgrad(gU, P->w("pool.W"), gA, Me, UP, d, nullptr, gA, 0));
  }
  for (int i = c.L - 1; i >= 0; i--) {
    std::string p = "enc.l" + std::to_string(i);
    const float* xin = i == 0 ? P->wsf("enc.x0") : P->wsf("enc.l" + std::to_string(i - 1) + ".out2");
    gB = gBs[1]; gD = gDs[1];
    SKF_TRY(G.before_write(gB)); SKF_TRY(G.before_write(gD));
    SKF_TRY(ln_bwd_any(P, gA, p, "2", gB, gD, &gS, Me, d, SITE_ENC_L + 2 * i + 1, st));
    SKF_TRY(G.wgrad_async(P->wsf(p + ".hff"), gS, P->g(p + ".ffn.w2.k"), P->g(p + ".ffn.w2.b"), Me, d, f));
    SKF_TRY(G.before_write(gH));
    SKF_TRY(G.dgrad(gS, P->w(p + ".ffn.w2.k"), gH, Me, d, f, P->wsf(p + ".hff"), nullptr, 1));
    SKF_TRY(G.wgrad_async(P->wsf(p + ".out1"), gH, P->g(p + ".ffn.w1.k"), P->g(p + ".ffn.w1.b"), Me, f, d));
    SKF_TRY(G.dgrad(gH, P->w(p + ".ffn.w1.k"), gA, Me, f, d, nullptr, gB, 0));
    gB = gBs[0]; gD = gDs[0];
    SKF_TRY(G.before_write(gB)); SKF_TRY(G.before_write(gD));
    SKF_TRY(ln_bwd_any(P, gA, p, "1", gB, gD, &gS, Me, d, SITE_ENC_L + 2 * i, st));
    SKF_TRY(G.wgrad_async(P->wsf(p + ".ctx"), gS, P->g(p + ".mha.wo.k"), P->g(p + ".mha.wo.b"), Me, d, d));
    SKF_TRY(G.dgrad(gS, P->w(p + ".mha.wo.k"), gC, Me, d, d, nullptr, nullptr, 0));
    float* qkv = P->wsf(p + ".qkv");
    SKF_TRY(G.before_write(gQKV));
    SKF_TRY(attn_bwd_any(P, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, P->wsf(p + ".ctx"), d, gC, d,
                         P->wsf(p + ".lse"), delta, gQKV, 3 * d, gQKV + d, 3 * d, gQKV + 2 * d, 3 * d, enc_pad, T, T,
                         1, st, P->wsp<void>(p + ".img")));
    SKF_TRY(G.wgrad_async(xin, gQKV, P->g(p + ".mha.wqkv.k"), P->g(p + ".mha.wqkv.b"), Me, 3 * d, d));
    SKF_TRY(G.dgrad(gQKV, P->w(p + ".mha.wqkv.k"), gA, Me, 3 * d, d, nullptr, gB, 0));
    if (i == c.L / 2 && i > 0) SKF_TRY(buckets_ready(P, 2, st));      // encoder layers L/2.., bottleneck and classifier are final
  }
  SKF_TRY(embed_dropout(P, gA, (int64_t)Me * d, SITE_ENC_X0, st));
  SKF_TRY(embed_any_bwd(P, tok, "enc", gA, T, st));
  SKF_TRY(G.join());
  return buckets_join(P, st);
}

// ---- greedy decode (models/sketchformer.py:255-311) ----------------------------------------------
static int classify_embedding(skf_plan* P, cudaStream_t st) {
  const skf_cfg& c = P->c;
  int Cp = (c.C + 3) & ~3;
  SKF_TRY(classifier_fwd(P, st));
  return skf_class_ce(P->wsf("cls.logits"), Cp, nullptr, c.B, c.C, 0.f, nullptr, Cp, nullptr,
                      P->wsp<int32_t>("cls.argmax"), P->wsf("cls.probs"), st);
}

static int decode_begin(skf_plan* P, const float* embedding, const int32_t* expected_len, cudaStream_t st) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, d = c.d, Me = B * T;
  const int rnd = P->round_ops ? 1 : 0;
  Gemm G{P, st};
  SKF_TRY(refresh_compute_weights(P, st));
  float* pre = P->wsf("pre_decoder");
  if (c.u == 0) {
    // lowerdim=0: the "embedding" IS the encoder output [B, T, d] and feeds the decoder unexpanded
    const float* src = embedding ? embedding : P->wsf("enc.l" + std::to_string(c.L - 1) + ".out2");
    SKF_CUDA(cudaMemcpyAsync(pre, src, (size_t)Me * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    if (embedding && embedding != P->wsf("embedding"))
      SKF_CUDA(cudaMemcpyAsync(P->wsf("embedding"), embedding, (size_t)B * emb_dim(c) * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    if (c.do_classification) SKF_TRY(classify_embedding(P, st));     // sketchformer.py:307-310
    // the expanded embedding and every layer's cross-attention keys / values do not depend on the prefix
    SKF_TRY(skf_expand_fwd(P->wsf("embedding"), P->pw("exp.k"), P->pw("exp.b"), pre, B, T, emb_dim(c), rnd, st));
  }
  for (int i = 0; i < c.L; i++) {
    std::string p = "dec.l" + std::to_string(i);
    SKF_TRY(G.fwd(pre, P->w(p + ".mha2.wkv.k"), P->pw(p + ".mha2.wkv.b"), P->wsf(p + ".kv2"), Me, 2 * d, emb_dim(c),
                  SKF_ACT_NONE, nullptr, 0));
  }
  SKF_TRY(decode_init(P->wsp<int64_t>("gen.tokens"), P->wsf("gen.strokes"), P->wsp<uint8_t>("gen.pad"),
                      P->wsp<uint8_t>("gen.cross_pad"), P->wsp<uint8_t>("gen.fin"), P->wsp<int32_t>("gen.state"),
                      expected_len, B, T, P->gen_sos, c.continuous, st));
  return SKF_OK;
}

// position `i` of every sample: decoder input = column i of the generated prefix -> column i+1
// pick = false: only the position's self-attention keys / values are (re)written, no logits, no new token
static int decode_step(skf_plan* P, int i, bool cross_masked, cudaStream_t st, bool pick = true) {
  const skf_cfg& c = P->c;
  const int B = c.B, T = c.T, d = c.d, f = c.f, H = c.H, dh = c.d / c.H;
  const int rnd = P->round_ops ? 1 : 0;
  Gemm G{P, st};
  float *x = P->wsf("gen.x"), *qkv = P->wsf("gen.qkv"), *ctx = P->wsf("gen.ctx"), *sb = P->wsf("gen.s");
  float *o1 = P->wsf("gen.o1"), *q2 = P->wsf("gen.q2"), *o2 = P->wsf("gen.o2"), *hf = P->wsf("gen.h");
  uint8_t* pad = P->wsp<uint8_t>("gen.pad");
  const float* pe = P->wsf("pe") + (size_t)i * d;
  if (c.continuous)
    SKF_TRY(skf_embed_cont_fwd(P->wsf("gen.strokes") + (size_t)i * 5, (int64_t)(T + 1) * 5, P->pw("dec.emb.k"),
                               P->pw("dec.emb.b"), pe, x, B, 1, d, rnd, st));
  else
    SKF_TRY(skf_embed_fwd(P->wsp<int64_t>("gen.tokens") + i, T + 1, P->pw("dec.emb"), pe, x, B, 1, d, c.V, rnd, st));
  for (int l = 0; l < c.L; l++) {
    std::string p = "dec.l" + std::to_string(l);
    float* kvc = P->wsf("gen.kv" + std::to_string(l));
    SKF_TRY(G.fwd(x, P->w(p + ".mha1.wqkv.k"), P->pw(p + ".mha1.wqkv.b"), qkv, B, 3 * d, d, SKF_ACT_NONE, nullptr, 0));
    // combined mask of the prefix: keys 0..i, generated PAD positions masked (create_masks on `output`)
    SKF_TRY(decode_attn(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, kvc, kvc + d, 2 * d, (int64_t)T * 2 * d, i, i + 1, pad,
                        T + 1, 1 << 30, ctx, d, B, H, dh, st));
    SKF_TRY(G.fwd(ctx, P->w(p + ".mha1.wo.k"), P->pw(p + ".mha1.wo.b"), sb, B, d, d, SKF_ACT_NONE, x, 0));
    SKF_TRY(skf_add_layernorm_fwd(sb, nullptr, nullptr, P->pw(p + ".ln1.g"), P->pw(p + ".ln1.b"), nullptr, o1, nullptr,
                                  nullptr, B, d, rnd, st));
    SKF_TRY(G.fwd(o1, P->w(p + ".mha2.wq.k"), P->pw(p + ".mha2.wq.b"), q2, B, d, d, SKF_ACT_NONE, nullptr, 0));
    float* kv2 = P->wsf(p + ".kv2");
    SKF_TRY(decode_attn(q2, d, nullptr, nullptr, 0, kv2, kv2 + d, 2 * d, (int64_t)T * 2 * d, -1, T,
                        cross_masked ? P->wsp<uint8_t>("gen.cross_pad") : nullptr, T, 1 << 30, ctx, d, B, H, dh, st));
    SKF_TRY(G.fwd(ctx, P->w(p + ".mha2.wo.k"), P->pw(p + ".mha2.wo.b"), sb, B, d, d, SKF_ACT_NONE, o1, 0));
    SKF_TRY(skf_add_layernorm_fwd(sb, nullptr, nullptr, P->pw(p + ".ln2.g"), P->pw(p + ".ln2.b"), nullptr, o2, nullptr,
                                  nullptr, B, d, rnd, st));
    SKF_TRY(G.fwd(o2, P->w(p + ".ffn.w1.k"), P->pw(p + ".ffn.w1.b"), hf, B, f, d, SKF_ACT_RELU, nullptr, 1));
    SKF_TRY(G.fwd(hf, P->w(p + ".ffn.w2.k"), P->pw(p + ".ffn.w2.b"), sb, B, d, f, SKF_ACT_NONE, o2, 0));
    SKF_TRY(skf_add_layernorm_fwd(sb, nullptr, nullptr, P->pw(p + ".ln3.g"), P->pw(p + ".ln3.b"), nullptr, x, nullptr,
                                  nullptr, B, d, rnd, st));
  }
  if (!pick) return SKF_OK;
  float* lg = P->wsf("gen.logits");
  if (c.continuous) {
    skf_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = P->pw("out.b");
    SKF_TRY(gemm_simt(x, d, 1, P->pw("out.k"), 5, 1, lg, 8, B, 5, d, &e, st));
  } else {
    SKF_TRY(G.fwd(x, P->w("out.k"), P->pw("out.b"), lg, B, c.V, d, SKF_ACT_NONE, nullptr, 0));
  }
  return decode_pick(lg, c.continuous ? 8 : c.V, c.V, P->wsp<int64_t>("gen.tokens"), P->wsf("gen.strokes"),
                     P->wsp<uint8_t>("gen.pad"), P->wsp<uint8_t>("gen.fin"), P->wsp<int32_t>("gen.state"), B, T, i,
                     P->gen_eos, c.continuous, st);
}

// blind_decoder_mask=False and no expected_len (models/sketchformer.py:274-281): at step i the reference attends the
// first i + 1 cross-attention keys for EVERY decoder position, so the hidden states of the earlier positions — and
// with them the self-attention keys / values of layers >= 1 — change at every step and no cache survives a step.
// Like the reference, the prefix is recomputed: positions 0 .. i - 1 refill the caches under the step's mask, position
// i produces the token.  O(T^2) position steps per sequence (the reference: O(T^2) full decoder passes); the cached
// path above is what predict() and every blind model use.
int decode_cross_width(uint8_t* cross_pad, int B, int T, int width, cudaStream_t st);
static int decode_step_uncached(skf_plan* P, int i, cudaStream_t st) {
  SKF_TRY(decode_cross_width(P->wsp<uint8_t>("gen.cross_pad"), P->c.B, P->c.T, i + 1, st));
  for (int j = 0; j < i; j++) SKF_TRY(decode_step(P, j, true, st, false));
  return decode_step(P, i, true, st, true);
}

}  // namespace skf

// =============================================================================================
// C ABI
// =============================================================================================
extern int g_ln_bwd_bpsm;
extern int g_ln_bwd_rows;

extern "C" {

int skf_split_lo(const float* x, float* lo, int64_t n, void* stream) {
  SKF_REQUIRE(x && lo && n > 0, SKF_ERR_BAD_ARG, "split_lo: bad argument");
  return skf::split_lo(x, lo, n, as_stream(stream));
}

int skf_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  SKF_REQUIRE(x && hi && lo && n > 0, SKF_ERR_BAD_ARG, "split_bf16: bad argument");
  return skf::split_bf16(x, hi, lo, n, as_stream(stream));
}

int skf_set_option(const char* name, int value) {
  if (name && strcmp(name, "gemm_mix") == 0) {
    skf::g_gemm_mix = value < 0 ? 0 : (value > 2 ? 2 : value);
    return SKF_OK;
  }
  if (name && strcmp(name, "pdl") == 0) {
    skf::g_pdl = value & 15;        // programmatic dependent launch: bit mask of kernel classes (skf_common.cuh)
    return SKF_OK;
  }
  if (name && strcmp(name, "gemm_tail_split") == 0) {
    skf::g_t3_tail_split = value ? 1 : 0;     // pair GEMM: tiles of a partial last wave as two half-width units
    return SKF_OK;
  }
  if (name && strcmp(name, "gemm_bn128_max") == 0) {
    skf::g_t3_bn128_max = value;
    return SKF_OK;
  }
  if (name && strcmp(name, "ln_bwd_rows") == 0) {
    g_ln_bwd_rows = value >= 4 ? 4 : 2;
    return SKF_OK;
  }
  if (name && strcmp(name, "ln_bwd_bpsm") == 0) {
    g_ln_bwd_bpsm = value < 1 ? 1 : value;
    return SKF_OK;
  }
  if (name && strcmp(name, "gemm_tiny_simt") == 0) {
    skf::g_gemm_tiny_simt = value ? 1 : 0;
    return SKF_OK;
  }
  if (name && strcmp(name, "gemm_wsplit") == 0) {
    skf::g_gemm_wsplit = value ? 1 : 0;
    return SKF_OK;
  }
  if (name && strcmp(name, "attn3_img") == 0) {
    skf::g_attn3_img = value ? 1 : 0;
    return SKF_OK;
  }
  if (name && strcmp(name, "attn3_trace") == 0) {
    skf::g_a3_trace = value;
    return SKF_OK;
  }
  if (name && strcmp(name, "attn3") == 0) {
    skf::g_attn3 = value & 3;
    return SKF_OK;
  }
  if (name && strcmp(name, "kv_fold") == 0) {
    skf::g_kv_fold = value < 0 ? 0 : (value > 2 ? 2 : value);
    return SKF_OK;
  }
  if (name && strcmp(name, "wgrad_units") == 0) {
    skf::g_wgrad_units = value < 1 ? 1 : value;
    return SKF_OK;
  }
  if (name && strcmp(name, "wgrad_overlap") == 0) {
    skf::g_wgrad_overlap = value ? 1 : 0;
    return SKF_OK;
  }
  if (name && strcmp(name, "gemm_pair") == 0) {
    skf::g_gemm_pair = value ? 1 : 0;
    return SKF_OK;
  }
  if (name && strcmp(name, "tma_tf32") == 0) {
    skf::g_tma_tf32_type = value ? 1 : 0;
    return SKF_OK;
  }
  return skf::fail(SKF_ERR_BAD_ARG, "unknown option %s", name ? name : "(null)");
}

int64_t skf_param_count(const skf_cfg* cfg) {
  if (check_cfg(cfg) != SKF_OK) return -1;
  return build_layout(*cfg).total;
}

int skf_param_table(const skf_cfg* cfg, skf_param_entry* out, int cap) {
  SKF_TRY(check_cfg(cfg));
  ParamLayout L = build_layout(*cfg);
  int n = (int)L.e.size();
  for (int i = 0; i < n && i < cap; i++) {
    memset(&out[i], 0, sizeof(out[i]));
    strncpy(out[i].name, L.e[i].name.c_str(), sizeof(out[i].name) - 1);
    out[i].offset = L.e[i].off;
    out[i].rows = L.e[i].rows;
    out[i].cols = L.e[i].cols;
  }
  return n;
}

size_t skf_workspace_bytes(const skf_cfg* cfg) {
  if (check_cfg(cfg) != SKF_OK) return 0;
  ParamLayout L = build_layout(*cfg);
  return build_workspace(*cfg, L).total;
}

int skf_plan_create(const skf_cfg* cfg, float* params, float* grads, float* adam_m, float* adam_v, void* workspace,
                    size_t workspace_bytes, skf_plan** out) {
  SKF_TRY(check_cfg(cfg));
  SKF_REQUIRE(out != nullptr && params != nullptr && workspace != nullptr, SKF_ERR_BAD_ARG, "plan_create: NULL argument");
  SKF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SKF_ERR_BAD_ALIGN, "workspace must be 256-byte aligned");
  SKF_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0, SKF_ERR_BAD_ALIGN, "params must be 16-byte aligned");
  skf_plan* P = new skf_plan();
  P->c = *cfg;
  if (P->c.world_size < 1) P->c.world_size = 1;
  P->PL = build_layout(*cfg);
  P->W = build_workspace(*cfg, P->PL);
  if (workspace_bytes < P->W.total) {
    size_t need = P->W.total;
    delete P;
    return skf::fail(SKF_ERR_BAD_SHAPE, "workspace too small: %zu < %zu", workspace_bytes, need);
  }
  P->params = params; P->grads = grads; P->am = adam_m; P->av = adam_v;
  P->ws = reinterpret_cast<uint8_t*>(workspace);
  P->round_ops = (cfg->gemm_backend == 0) && !skf::g_tma_tf32_type;
  P->wc = params;
  {
    // second stream + event pool for the backward pass; without them (no device) the plan runs single-stream
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
        cudaStreamCreateWithPriority(&P->side, cudaStreamNonBlocking, lo) == cudaSuccess) {
      P->events.resize(512);
      for (cudaEvent_t& e : P->events)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
          e = nullptr;
          cudaStreamDestroy(P->side);
          P->side = nullptr;
          break;
        }
      if (!P->side) P->events.clear();
    } else {
      P->side = nullptr;
      (void)cudaGetLastError();
    }
  }
  *out = P;
  return SKF_OK;
}

void skf_plan_destroy(skf_plan* plan) { delete plan; }

int skf_plan_init(skf_plan* P, void* stream) {
  SKF_REQUIRE(P != nullptr, SKF_ERR_BAD_ARG, "plan is NULL");
  cudaStream_t st = as_stream(stream);
  int64_t l0 = skf::g_launch_count;
  SKF_CUDA(cudaMemsetAsync(P->ws + P->W.t.at("step").off, 0, 8, st));
  SKF_CUDA(cudaMemsetAsync(P->ws + P->W.t.at("stats_step").off, 0, 8 * sizeof(double), st));
  SKF_CUDA(cudaMemsetAsync(P->ws + P->W.t.at("stats_run").off, 0, 16 * sizeof(double), st));
  SKF_TRY(skf_positional_encoding(P->wsf("pe"), P->c.T, P->c.d, stream));
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int skf_plan_tensor(const skf_plan* P, const char* name, size_t* offset, size_t* bytes) {
  SKF_REQUIRE(P != nullptr && name != nullptr, SKF_ERR_BAD_ARG, "plan_tensor: NULL argument");
  auto it = P->W.t.find(name);
  SKF_REQUIRE(it != P->W.t.end(), SKF_ERR_BAD_ARG, "no workspace tensor named %s", name);
  if (offset) *offset = it->second.off;
  if (bytes) *bytes = it->second.bytes;
  return SKF_OK;
}

int skf_apply_adam(skf_plan* P, void* stream) {
  SKF_REQUIRE(P != nullptr && P->grads && P->am && P->av, SKF_ERR_BAD_ARG, "apply_adam: plan has no optimizer state");
  int64_t l0 = skf::g_launch_count;
  const skf_cfg& c = P->c;
  SKF_TRY(skf_adam_step(P->params, P->grads, P->am, P->av, P->PL.total, P->wsp<int64_t>("step"), c.d,
                        c.warmup_steps, c.adam_beta1, c.adam_beta2, c.adam_eps, 1.0f / (float)c.world_size, stream));
  P->w16_valid = false;          // the bf16 weight images are stale until the next refresh_compute_weights()
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int skf_train_step(skf_plan* P, const void* tokens, const int64_t* labels, int apply_adam, void* stream) {
  SKF_REQUIRE(P != nullptr && tokens != nullptr, SKF_ERR_BAD_ARG, "train_step: NULL argument");
  SKF_REQUIRE(P->grads != nullptr, SKF_ERR_BAD_ARG, "train_step: plan created without a gradient arena");
  SKF_REQUIRE(!P->c.do_classification || labels != nullptr, SKF_ERR_BAD_ARG, "train_step: labels required");
  cudaStream_t st = as_stream(stream);
  const void* tok = tokens;
  int64_t l0 = skf::g_launch_count;
  SKF_TRY(refresh_compute_weights(P, st));
  P->training = true;            // Dropout layers active (self.call(..., training=True), sketchformer.py:332)
  int rc = forward_encoder(P, tok, st);
  if (rc == SKF_OK && P->c.do_reconstruction) rc = forward_decoder(P, tok, st);
  if (rc == SKF_OK) rc = losses(P, tok, labels, true, st);
  if (rc == SKF_OK) rc = backward(P, tok, st);
  P->training = false;
  SKF_TRY(rc);
  P->launches += skf::g_launch_count - l0;
  if (apply_adam) SKF_TRY(skf_apply_adam(P, stream));
  return SKF_OK;
}

int skf_encode(skf_plan* P, const void* tokens, void* stream) {
  SKF_REQUIRE(P != nullptr && tokens != nullptr, SKF_ERR_BAD_ARG, "encode: NULL argument");
  cudaStream_t st = as_stream(stream);
  int64_t l0 = skf::g_launch_count;
  SKF_TRY(refresh_compute_weights(P, st));
  SKF_TRY(forward_encoder(P, tokens, st));
  if (P->c.do_classification) {
    int Cp = (P->c.C + 3) & ~3;
    SKF_TRY(skf_class_ce(P->wsf("cls.logits"), Cp, nullptr, P->c.B, P->c.C, 0.f, nullptr, Cp, nullptr,
                         P->wsp<int32_t>("cls.argmax"), P->wsf("cls.probs"), stream));
  }
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int skf_forward(skf_plan* P, const void* tokens, void* stream) {
  SKF_REQUIRE(P != nullptr && tokens != nullptr, SKF_ERR_BAD_ARG, "forward: NULL argument");
  cudaStream_t st = as_stream(stream);
  const void* tok = tokens;
  int64_t l0 = skf::g_launch_count;
  SKF_TRY(refresh_compute_weights(P, st));
  SKF_TRY(forward_encoder(P, tok, st));
  if (P->c.do_reconstruction) SKF_TRY(forward_decoder(P, tok, st));
  SKF_TRY(losses(P, tok, nullptr, false, st));
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int skf_decode_begin(skf_plan* P, const float* embedding, const int32_t* expected_len, int64_t sos_token,
                     int64_t eos_token, void* stream) {
  SKF_REQUIRE(P != nullptr, SKF_ERR_BAD_ARG, "decode_begin: plan is NULL");
  if (!P->c.do_reconstruction) {        // classification of a given embedding only (sketchformer.py:307-310)
    SKF_REQUIRE(P->c.do_classification, SKF_ERR_UNSUPPORTED, "decode: the model has neither decoder nor classifier");
    cudaStream_t st = as_stream(stream);
    int64_t l0 = skf::g_launch_count;
    SKF_TRY(refresh_compute_weights(P, st));    // like every other entry point: host-side weight-image flags are stale
                                                // after a CUDA-graph replay or a parameter load through torch
    if (embedding && embedding != P->wsf("embedding"))
      SKF_CUDA(cudaMemcpyAsync(P->wsf("embedding"), embedding, (size_t)P->c.B * emb_dim(P->c) * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    SKF_TRY(classify_embedding(P, st));
    P->launches += skf::g_launch_count - l0;
    return SKF_OK;
  }
  // blind_decoder_mask=False without expected_len: the reference's i+1 cross mask changes earlier positions at every
  // step and is not cacheable: decode_step_uncached recomputes the prefix per step
  P->gen_uncached = !P->c.blind_decoder_mask && expected_len == nullptr;
  int64_t l0 = skf::g_launch_count;
  P->gen_sos = sos_token;
  P->gen_eos = eos_token;
  SKF_TRY(decode_begin(P, embedding, P->c.blind_decoder_mask ? nullptr : expected_len, as_stream(stream)));
  P->gen_ready = true;
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int skf_decode_steps(skf_plan* P, int first_step, int n_steps, void* stream) {
  SKF_REQUIRE(P != nullptr && P->gen_ready, SKF_ERR_BAD_ARG, "decode_steps: call skf_decode_begin first");
  SKF_REQUIRE(first_step >= 0 && n_steps >= 0 && first_step + n_steps <= P->c.T, SKF_ERR_BAD_SHAPE,
              "decode_steps: steps [%d, %d) outside [0, max_seq_len=%d)", first_step, first_step + n_steps, P->c.T);
  int64_t l0 = skf::g_launch_count;
  for (int i = first_step; i < first_step + n_steps; i++) {
    if (P->gen_uncached) SKF_TRY(decode_step_uncached(P, i, as_stream(stream)));
    else SKF_TRY(decode_step(P, i, !P->c.blind_decoder_mask, as_stream(stream)));
  }
  P->launches += skf::g_launch_count - l0;
  return SKF_OK;
}

int64_t skf_plan_launches(const skf_plan* P) { return P ? P->launches : 0; }

size_t skf_attn_image_bytes(int B, int H, int Tq, int Tk, int which) {
  return skf::attn3_image_bytes(B, H, Tq, Tk, which == 0 ? 6 : 2);
}

int skf_attn_fwd_img(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
                     int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh,
                     int mask_mode, void* img, void* stream) {
  return attn3_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, dh, mask_mode, img, as_stream(stream));
}

int skf_attn_bwd_img(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                     const float* O, int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta,
                     float* dQ, int64_t lddq, float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad,
                     int B, int H, int Tq, int Tk, int dh, int mask_mode, const void* img, void* img_do, void* stream) {
  return attn3_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, key_pad, B, H,
                   Tq, Tk, dh, mask_mode, img, img_do, as_stream(stream));
}

int skf_debug_trace(int64_t* out, int n) {
  SKF_REQUIRE(out != nullptr && n > 0, SKF_ERR_BAD_ARG, "debug_trace: bad argument");
  return skf::attn3_trace_dump(reinterpret_cast<long long*>(out), n);
}

int skf_grad_buckets(const skf_cfg* cfg, int n_buckets, int64_t* lo, int64_t* hi) {
  SKF_TRY(check_cfg(cfg));
  SKF_REQUIRE(lo != nullptr && hi != nullptr && n_buckets >= 1, SKF_ERR_BAD_ARG, "grad_buckets: bad argument");
  ParamLayout L = build_layout(*cfg);
  Buckets b = build_buckets(*cfg, L, n_buckets);
  for (int i = 0; i < b.n; i++) { lo[i] = b.lo[i]; hi[i] = b.hi[i]; }
  return b.n;
}

int skf_plan_set_comm(skf_plan* P, void* comm, int n_buckets) {
  SKF_REQUIRE(P != nullptr, SKF_ERR_BAD_ARG, "plan_set_comm: plan is NULL");
  P->comm = comm;
  if (!comm) return SKF_OK;
  SKF_REQUIRE(P->grads != nullptr, SKF_ERR_BAD_ARG, "plan_set_comm: plan created without a gradient arena");
  SKF_REQUIRE(!P->events.empty(), SKF_ERR_CUDA, "plan_set_comm: the plan has no event pool (no CUDA device?)");
  if (!P->comm_stream) {
    int lo = 0, hi = 0;
    SKF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SKF_CUDA(cudaStreamCreateWithPriority(&P->comm_stream, cudaStreamNonBlocking, hi));
  }
  P->buckets = build_buckets(P->c, P->PL, n_buckets);
  return SKF_OK;
}

// ---- thin operator wrappers -------------------------------------------------------------------
int skf_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* Y,
             int64_t ldy, int M, int N, int K, const skf_gemm_epilogue* epi, int backend, void* stream) {
  SKF_REQUIRE(!(epi && epi->colsum) || backend == 5, SKF_ERR_UNSUPPORTED,
              "gemm: the fused column sum exists on the pair backend (5) only");
  if (backend == 0 || backend == 2)      // persistent tcgen05 kernel: single-pass TF32 / 3xTF32
    return gemm_tc2(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, epi, backend == 2, as_stream(stream));
  if (backend == 5) {                    // 3xTF32 on CTA pairs (cta_group::2), 256 x 256 cluster tiles
    SKF_REQUIRE(gemm_tc2_supported(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, epi), SKF_ERR_BAD_ALIGN,
                "gemm (pair): operand strides/alignment not TMA-compatible");
    return gemm_tc3(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, epi, as_stream(stream));
  }
  return gemm_simt(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, epi, as_stream(stream));
}

// ---- sublayer entry points (include/skf.h): the plan's launch sequences on caller-owned buffers --------------------
// the plan's choice of kernel for one product (Gemm::run without a plan): bf16x3 on CTA pairs where the shape is worth
// it, the single-CTA tcgen05 kernel (3xTF32) for the other TMA-compatible shapes, the fp32 SIMT kernel otherwise
static int gemm_auto(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* Y,
                     int64_t ldy, int M, int N, int K, skf_gemm_epilogue e, cudaStream_t st) {
  const bool tc = gemm_tc2_supported(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e) != 0;
  const bool pair = tc && gemm_tc3_wanted(M, N, K);
  if (e.accumulate) {            // split-K so that the persistent grid has about one unit per cluster / SM
    const long tiles = pair ? (long)cdiv(M, 256) * cdiv(N, 256)
                       : tc ? (long)cdiv(M, 128) * cdiv(N, N <= 64 ? 64 : (N <= 128 ? 128 : 256))
                            : (long)cdiv(M, 64) * cdiv(N, 64);
    const int want = (int)((pair ? 74 : tc ? 148 : 2 * 148) / tiles);
    const int maxs = K / 256 > 0 ? K / 256 : 1;
    e.split_k = want < 1 ? 1 : (want > maxs ? maxs : want);
  }
  if (!pair) e.colsum = nullptr, e.b_hi16 = e.b_lo16 = nullptr;
  if (pair) return gemm_tc3(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, st);
  return tc ? gemm_tc2(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, 1, st)
            : gemm_simt(A, sa_m, sa_k, B, sb_k, sb_n, Y, ldy, M, N, K, &e, st);
}
// dW[K,N] += X[M,K]^T dY[M,N]; db[N] += colsum(dY) (fused into the pair kernel when it runs, else its own launch)
static int wgrad_auto(const float* X, const float* dY, int64_t ldy, float* dW, float* db, int M, int N, int K,
                      cudaStream_t st) {
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.accumulate = 1;
  const bool fused = db && ldy == N && gemm_tc3_wanted(K, N, M) &&
                     gemm_tc2_supported(X, 1, K, dY, ldy, 1, dW, N, K, N, M, &e);
  if (fused) e.colsum = db;
  SKF_TRY(gemm_auto(X, 1, K, dY, ldy, 1, dW, N, K, N, M, e, st));
  if (db && !fused) SKF_TRY(skf_colsum(dY, ldy, db, M, N, st));
  return SKF_OK;
}

int skf_ffn_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                const float* resid, float* h, float* y, int M, int d, int f, const void* w1_hi16,
                const void* w1_lo16, const void* w2_hi16, const void* w2_lo16, void* stream) {
  SKF_REQUIRE(x && W1 && W2 && h && y && M > 0 && d > 0 && f > 0, SKF_ERR_BAD_ARG, "ffn_fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  const bool img = w1_hi16 && w1_lo16 && w2_hi16 && w2_lo16;
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.bias = b1; e.act = SKF_ACT_RELU;
  if (img) e.b_hi16 = w1_hi16, e.b_lo16 = w1_lo16;
  SKF_TRY(gemm_auto(x, d, 1, W1, f, 1, h, f, M, f, d, e, st));            // transformer.py:196
  memset(&e, 0, sizeof(e));
  e.bias = b2; e.resid = resid; e.ld_resid = d;
  if (img) e.b_hi16 = w2_hi16, e.b_lo16 = w2_lo16;
  return gemm_auto(h, f, 1, W2, d, 1, y, d, M, d, f, e, st);               // transformer.py:197 + the residual of 222
}

int skf_ffn_bwd(const float* x, const float* h, const float* dy, const float* W1, const float* W2, float* dh,
                float* dx, const float* dx_resid, float* dW1, float* db1, float* dW2, float* db2, int M, int d,
                int f, const void* w1_hi16, const void* w1_lo16, const void* w2_hi16, const void* w2_lo16,
                void* stream) {
  SKF_REQUIRE(x && h && dy && W1 && W2 && dh && dx && dW1 && dW2 && M > 0 && d > 0 && f > 0, SKF_ERR_BAD_ARG,
              "ffn_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  const bool img = w1_hi16 && w1_lo16 && w2_hi16 && w2_lo16;
  SKF_TRY(wgrad_auto(h, dy, d, dW2, db2, M, d, f, st));                    // dW2 += h^T dy, db2 += colsum(dy)
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.mask = h; e.ld_mask = f;                                               // ReLU backward: * [h > 0]
  if (img) e.b_hi16 = w2_hi16, e.b_lo16 = w2_lo16;
  SKF_TRY(gemm_auto(dy, d, 1, W2, 1, d, dh, f, M, f, d, e, st));          // dh = (dy W2^T) * [h > 0]
  SKF_TRY(wgrad_auto(x, dh, f, dW1, db1, M, f, d, st));                    // dW1 += x^T dh, db1 += colsum(dh)
  memset(&e, 0, sizeof(e));
  e.resid = dx_resid; e.ld_resid = d;
  if (img) e.b_hi16 = w1_hi16, e.b_lo16 = w1_lo16;
  return gemm_auto(dh, f, 1, W1, 1, f, dx, d, M, d, f, e, st);             // dx = dh W1^T (+ dx_resid)
}

int skf_logits_ce_fwd(const float* x, const float* W, const float* b, const int64_t* tgt, int64_t tgt_ld, int B,
                      int Td, int d, int V, float grad_scale, float* logits, int64_t ldl, double* stats,
                      int32_t* argmax_out, const void* w_hi16, const void* w_lo16, void* stream) {
  SKF_REQUIRE(x && W && tgt && logits && stats && B > 0 && Td > 0 && d > 0 && V > 0 && ldl >= V, SKF_ERR_BAD_ARG,
              "logits_ce_fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.bias = b;
  if (w_hi16 && w_lo16) e.b_hi16 = w_hi16, e.b_lo16 = w_lo16;
  SKF_TRY(gemm_auto(x, d, 1, W, V, 1, logits, ldl, B * Td, V, d, e, st));  // models/sketchformer.py:179
  return skf_token_ce(logits, ldl, tgt, tgt_ld, B, Td, V, grad_scale, logits, ldl, stats, argmax_out, 0, stream);
}

int skf_logits_ce_bwd(const float* x, const float* dlogits, int64_t ldl, const float* W, float* dx, float* dW,
                      float* db, int B, int Td, int d, int V, const void* w_hi16, const void* w_lo16, void* stream) {
  SKF_REQUIRE(x && dlogits && W && dx && dW && B > 0 && Td > 0 && d > 0 && V > 0 && ldl >= V, SKF_ERR_BAD_ARG,
              "logits_ce_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  const int M = B * Td;
  SKF_TRY(wgrad_auto(x, dlogits, ldl, dW, db, M, V, d, st));               // dW += x^T dlogits, db += colsum
  skf_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  if (w_hi16 && w_lo16) e.b_hi16 = w_hi16, e.b_lo16 = w_lo16;
  return gemm_auto(dlogits, ldl, 1, W, 1, V, dx, d, M, d, V, e, st);       // dx = dlogits W^T
}

int skf_attn_fwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
                 int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh,
                 int mask_mode, int round_tf32, int backend, void* stream) {
  // backend: 0 = tcgen05 flash kernels, single-pass TF32; 2 = the same with the 3xTF32 split (fp32-grade);
  //          1 = SIMT fp32
  if (backend == 6)                      // warp-specialised whole-row kernels (head dim 32, T <= 224), bf16x3
    return attn3_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, dh, mask_mode, nullptr, as_stream(stream));
  if (backend == 0 || backend == 2)
    return attn2_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, dh, mask_mode, backend == 2,
                     as_stream(stream));
  return attn_fwd_simt(Q, ldq, K, ldk, V, ldv, O, ldo, lse, key_pad, B, H, Tq, Tk, dh, mask_mode, round_tf32,
                       as_stream(stream));
}

int skf_attn_bwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                 const float* O, int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta,
                 float* dQ, int64_t lddq, float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad,
                 int B, int H, int Tq, int Tk, int dh, int mask_mode, int round_tf32, int backend, void* stream) {
  if (backend == 6)
    return attn3_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, key_pad, B, H,
                     Tq, Tk, dh, mask_mode, nullptr, nullptr, as_stream(stream));
  if (backend == 0 || backend == 2)
    return attn2_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, key_pad, B,
                     H, Tq, Tk, dh, mask_mode, backend == 2, as_stream(stream));
  return attn_bwd_simt(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, key_pad,
                       B, H, Tq, Tk, dh, mask_mode, round_tf32, as_stream(stream));
}

}  // extern "C"
