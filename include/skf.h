/*
 * skf.h — C ABI of libskf.so, the B200 (sm_100a) device path of the Sketchformer hot path.
 *
 * The reference (leosampaio/sketchformer, TF 2.1) has NO native/FFI boundary of its own: its
 * device path is a tf.function over stock TF ops.  The entry points below are therefore the
 * boundary SURVEY.md §8(b) dictates; each one cites the reference Python symbol it replaces
 * (paths relative to the reference repository).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every function returns 0 (SKF_OK) or a negative skf_status; skf_last_error() returns a
 *    thread-local message for the last failure on the calling thread;
 *  - no allocation, no ownership transfer: every buffer is a device pointer owned by the caller
 *    (torch tensors on the Python side); sizes come from the skf_*_bytes / skf_param_* queries;
 *  - every launch takes the cudaStream_t as `void* stream`, is asynchronous and never
 *    synchronises (CUDA-graph capturable); one caller per stream;
 *  - `round_tf32` arguments: when non-zero the stored result is rounded (to nearest) to TF32
 *    precision because a later tcgen05 kind::tf32 GEMM consumes it and the tensor core would
 *    otherwise truncate the low 13 mantissa bits (biased);
 *  - all floating-point tensors are fp32, row-major, leading dimension given in ELEMENTS;
 *    token / label ids are int64 (models/sketchformer.py:317-322);
 *  - weights use the Keras orientation kernel = (in, out).
 */
#ifndef SKF_H_
#define SKF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum skf_status {
  SKF_OK = 0,
  SKF_ERR_BAD_SHAPE = -1,
  SKF_ERR_BAD_ALIGN = -2,
  SKF_ERR_UNSUPPORTED = -3,
  SKF_ERR_CUDA = -4,
  SKF_ERR_NCCL = -5,
  SKF_ERR_BAD_ARG = -6
} skf_status;

/* Model / batch description.  Mirrors the hparam surface of models/sketchformer.py:24-53 and
 * dataloaders/distributed_stroke3.py:15-25 that changes device work. */
typedef struct skf_cfg {
  int32_t B;                  /* batch_size (per rank)                                  */
  int32_t T;                  /* max_seq_len: encoder length Te = T, decoder Td = T - 1 */
  int32_t d;                  /* d_model                                                */
  int32_t f;                  /* dff                                                    */
  int32_t H;                  /* num_heads                                              */
  int32_t L;                  /* num_layers                                             */
  int32_t V;                  /* vocab size (1004 for the dictionary tokenizer)         */
  int32_t C;                  /* n_classes                                              */
  int32_t u;                  /* lowerdim (hidden width of the pooling attention)       */
  int32_t continuous;         /* use_continuous_data (stroke-5 input, Dense(5) head)    */
  int32_t blind_decoder_mask; /* cross-attention mask all-zero when 1                   */
  int32_t do_classification;
  int32_t do_reconstruction;
  int32_t warmup_steps;       /* 5000: hard-coded at models/sketchformer.py:113-114      */
  float class_weight;
  float recon_weight;
  float adam_beta1, adam_beta2, adam_eps; /* 0.9, 0.98, 1e-9 (sketchformer.py:121-123)  */
  int32_t world_size;         /* data-parallel ranks: gradients are scaled by 1/world   */
  int32_t gemm_backend;       /* 0 = tcgen05 TF32 tensor-core GEMMs, 1 = SIMT fp32 (debug)*/
  int32_t attn_backend;       /* 0 = default, 1 = SIMT fp32 (debug)                      */
  int32_t gemm_precision;     /* tcgen05 GEMMs: 0 = TF32 everywhere, 1 = 3xTF32 (fp32-accurate) forward
                                 + TF32 backward, 2 = 3xTF32 everywhere, 3 = like 2 with single-pass
                                 TF32 weight gradients                                        */
  float dropout_rate;         /* models/sketchformer.py:31; applied by skf_train_step only (training=True) */
  uint32_t dropout_seed;      /* seed of the counter-based dropout masks (see skf_dropout)   */
  int32_t attn_version;       /* 1 = SelfAttnV1 (embedding width d_model), otherwise SelfAttnV2: W is (d,d) and a
                                 Dense(lowerdim) follows the pooling, embedding width = lowerdim (0 is read as 1) */
  int32_t class_buffer_layers;/* Dense(lowerdim, relu) + Dropout(class_dropout) blocks before the classifier      */
  float class_dropout;        /* models/sketchformer.py:38                                   */
} skf_cfg;

const char* skf_last_error(void);
/* Process-wide options.  "tma_tf32": 1 = tensor maps use CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 (the TMA
 * unit converts fp32->tf32 on load) instead of pre-rounded operands.  Set before creating plans. */
int skf_set_option(const char* name, int value);
int skf_version(void);
/* sizeof() of the ABI structs as this library was compiled: which = 0 skf_cfg, 1 skf_gemm_epilogue,
 * 2 skf_param_entry; 0 for anything else.  A binding checks its own struct layouts against these at load time. */
size_t skf_abi_sizeof(int which);
/* 1 if the current device can run the sm_100a kernels. */
int skf_device_ok(void);

/* ---------------- parameter arena ---------------------------------------------------------
 * All trainable variables live in one flat fp32 arena (same layout for params, grads and the
 * two Adam moments).  q/k/v kernels of a self-attention block are fused column-wise into
 * "<mha>.wqkv" (d, 3d); cross-attention keeps "<mha2>.wq" (d,d) and "<mha2>.wkv" (d, 2d).    */
typedef struct skf_param_entry {
  char name[56];
  int64_t offset; /* in floats */
  int32_t rows, cols;
} skf_param_entry;

int64_t skf_param_count(const skf_cfg* cfg);
/* Fills up to `cap` entries, returns the number of entries (or <0). */
int skf_param_table(const skf_cfg* cfg, skf_param_entry* out, int cap);

/* ---------------- operator-level entry points -------------------------------------------- */

/* Encoder.call / Decoder.call head, builders/layers/transformer.py:288-296, 325-334:
 * out[b,t,:] = table[tok[b,t],:] * sqrt(d) + PE[t,:].  tok has row stride tok_ld (int64).   */
int skf_embed_fwd(const int64_t* tok, int64_t tok_ld, const float* table, const float* pe,
                  float* out, int B, int T, int d, int V, int round_tf32, void* stream);
/* tf.keras.layers.Dropout (inverted dropout, training=True) with a counter-based mask: element i of dropout
 * site `site` is kept iff hash(seed, site, *step, i) >= rate * 2^24 (top 24 bits of two murmur3 finalisers)
 * and scaled by 1/(1-rate); `step` is a device pointer to optimizer.iterations, so every train step draws a
 * new mask, the backward pass regenerates it, and nothing is stored.  y may alias x.  Used after the input
 * embeddings (transformer.py:296, 334) and on the gradient flowing back through them.              */
int skf_dropout(const float* x, float* y, int64_t n, float rate, uint32_t seed, uint32_t site,
                const int64_t* step, void* stream);
/* dtable[tok[b,t],:] += dout[b,t,:] * sqrt(d)   (IndexedSlices scatter-add)     */
int skf_embed_bwd(const int64_t* tok, int64_t tok_ld, const float* dout, float* dtable, int B,
                  int T, int d, int V, void* stream);
/* builders/utils.py:17-32 — fills pe[rows, d] (float64 math, rounded to fp32).             */
int skf_positional_encoding(float* pe, int rows, int d, void* stream);
/* builders/utils.py:35-43 — pad[b,t] = (tok[b,t]==0) (tokens) as uint8.                    */
int skf_pad_flags(const int64_t* tok, int64_t tok_ld, uint8_t* pad, int B, int T, void* stream);

/* use_continuous_data=True (stroke-5 rows [dx,dy,p_down,p_up,p_pad], float32).  x has batch stride x_ld
 * (elements) and row stride 5.
 * builders/utils.py:38-39 — pad[b,t] = (x[b,t,4] == 1).                                      */
int skf_pad_flags_cont(const float* x, int64_t x_ld, uint8_t* pad, int B, int T, void* stream);
/* Encoder/Decoder.call with use_continuous_input (transformer.py:275-276, 292-294, 314-315):
 * out[b,t,:] = (x[b,t,:] . W[5,d] + bias) * sqrt(d) + PE[t,:].                               */
int skf_embed_cont_fwd(const float* x, int64_t x_ld, const float* W, const float* bias, const float* pe,
                       float* out, int B, int T, int d, int round_tf32, void* stream);
/* dW[5,d] += sqrt(d) x^T dout ; db[d] += sqrt(d) colsum(dout)                                 */
int skf_embed_cont_bwd(const float* x, int64_t x_ld, const float* dout, float* dW, float* db, int B, int T,
                       int d, void* stream);
/* Continuous reconstruction loss (builders/losses.py:43-66):
 * loss = weight * mean_{b,t}((loc + M) * mask), loc = mean_2((real_xy - pred_xy)^2),
 * M = mean_{b,t} CE3(argmax real[2:5], pred[2:5]) over ALL positions, mask = (real[...,4] != 1).
 * stats[0] += sum loc*mask, stats[1] += sum CE3, stats[2] += sum mask (doubles, zeroed by the caller);
 * dpred (optional, [B*Td, lddp]) = d loss / d pred.  real has batch stride real_ld, row stride 5. */
int skf_cont_loss(const float* pred, int64_t ldp, const float* real, int64_t real_ld, int B, int Td,
                  float weight, float* dpred, int64_t lddp, double* stats, void* stream);

/* tf.keras.layers.Dense (builders/layers/transformer.py:154-158,196-197) and its gradients.
 * Y[M,N] = epilogue(op(A)[M,K] * op(B)[K,N]).  A is addressed as A[m*sa_m + k*sa_k], B as
 * B[k*sb_k + n*sb_n] (element strides), so one entry point serves forward (X*W), dgrad (dY*W^T)
 * and wgrad (X^T*dY).  Epilogue: v = acc; v += bias[n]; act; v *= (mask[m,n] > 0);
 * v += resid[m,n]; Y = accumulate ? Y + v : v.  backend: 0 tcgen05, 1 SIMT.                */
typedef enum skf_act { SKF_ACT_NONE = 0, SKF_ACT_RELU = 1, SKF_ACT_TANH = 2 } skf_act;
typedef struct skf_gemm_epilogue {
  const float* bias;  /* [N] or NULL */
  const float* resid; /* [M, ld_resid] or NULL */
  int64_t ld_resid;
  const float* mask;  /* [M, ld_mask] or NULL: multiply by (mask > 0) (ReLU backward)        */
  int64_t ld_mask;
  int32_t act;        /* skf_act */
  int32_t accumulate; /* atomically add into Y (split-K weight gradients)                    */
  int32_t round_tf32; /* round the stored value to TF32 (operand of a later tensor-core GEMM)*/
  int32_t split_k;    /* >1: split the K range over this many CTAs (needs accumulate=1)       */
  /* Both non-NULL (tcgen05 backend only): error-compensated "3xTF32" product
   * A*B ~= Ah*Bh + Ah*Bl + Al*Bh with Xl = X - tf32(X) from skf_split_lo(); same strides as A / B. */
  const float* a_lo;
  const float* b_lo;
  /* Bias gradient fused into a weight-gradient product (pair backend 5, accumulate=1, B row-major [K, N]):
   * colsum[n] += sum_k B[k, n], taken from the B tiles the CTAs stage anyway (row-block 0 only).          */
  float* colsum;
  /* Both non-NULL (pair backend 5 in bf16x3 mode): B is a weight whose bf16 images hi = bf16(B) and
   * lo = bf16(B - hi) from skf_split_bf16() exist with the element layout and strides of B (base 16-byte aligned,
   * row pitch a multiple of 8 elements); the kernel then TMA-loads the bf16 tiles instead of splitting B's fp32
   * tiles in shared memory (Dense forward and dgrad: the same weights serve B*T rows).                       */
  const void* b_hi16;
  const void* b_lo16;
} skf_gemm_epilogue;
/* hi[i] = bf16(x[i]) (round to nearest even), lo[i] = bf16(x[i] - hi[i]): operand images of the bf16x3 product. */
int skf_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream);
/* lo[i] = x[i] - tf32(x[i])  (exact in fp32): the low half of the 3xTF32 operand split.     */
int skf_split_lo(const float* x, float* lo, int64_t n, void* stream);
int skf_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k,
             int64_t sb_n, float* Y, int64_t ldy, int M, int N, int K,
             const skf_gemm_epilogue* epi, int backend, void* stream);

/* ---- sublayer entry points: the operator groups SURVEY.md 8(b) names `skf_ffn_*` and `skf_logits_ce_*` ----------
 * Each is the sequence of launches the plan runs for that group, on caller-owned buffers.  They are NOT single
 * kernels: the hidden tensor / the logits go through HBM between the launches (DESIGN.md section 9 has the measured
 * reasons).  Products run as the plan's default: the bf16x3 pair GEMM where the shape allows it, with the weights'
 * bf16 images when all of them are given (skf_split_bf16; NULL = split in shared memory).
 *
 * Position-wise feed-forward sublayer with its residual (builders/layers/transformer.py:194-198, 220-222, 258-260):
 *   h = relu(x W1 + b1) [M, f] (an output: the backward pass needs it);  y = resid + h W2 + b2 [M, d]
 * x, resid (may be NULL), y: [M, d] row-major; W1 [d, f], W2 [f, d] Keras kernels.  Two launches.            */
int skf_ffn_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                const float* resid, float* h, float* y, int M, int d, int f, const void* w1_hi16,
                const void* w1_lo16, const void* w2_hi16, const void* w2_lo16, void* stream);
/* Its gradients from dy = dL/dy [M, d] (the residual branch passes dy through and stays with the caller):
 *   dW2 += h^T dy, db2 += colsum(dy);  dh = (dy W2^T) * [h > 0] [M, f] (scratch / output);
 *   dW1 += x^T dh, db1 += colsum(dh);  dx = dh W1^T (+ dx_resid when non-NULL) [M, d].
 * The weight gradients ACCUMULATE (zero them first); four GEMM launches, bias gradients fused where the shape
 * allows, else one column-sum launch each.                                                                   */
int skf_ffn_bwd(const float* x, const float* h, const float* dy, const float* W1, const float* W2, float* dh,
                float* dx, const float* dx_resid, float* dW1, float* db1, float* dW2, float* db2, int M, int d,
                int f, const void* w1_hi16, const void* w1_lo16, const void* w2_hi16, const void* w2_lo16,
                void* stream);
/* Output layer + token loss (models/sketchformer.py:179 `final_layer`; builders/losses.py:26-41; accuracy 335-339):
 *   logits = x W + b [B*Td, V] (row pitch ldl >= V, a multiple of 4), then skf_token_ce IN PLACE: stats[0] += CE sum,
 *   stats[1] += correct, argmax_out[row], and `logits` leaves as dlogits = (softmax - onehot) [tgt != 0] grad_scale.
 * Two launches.                                                                                               */
int skf_logits_ce_fwd(const float* x, const float* W, const float* b, const int64_t* tgt, int64_t tgt_ld, int B,
                      int Td, int d, int V, float grad_scale, float* logits, int64_t ldl, double* stats,
                      int32_t* argmax_out, const void* w_hi16, const void* w_lo16, void* stream);
/* dx = dlogits W^T [B*Td, d];  dW += x^T dlogits [d, V];  db += colsum(dlogits) [V].                           */
int skf_logits_ce_bwd(const float* x, const float* dlogits, int64_t ldl, const float* W, float* dx, float* dW,
                      float* db, int B, int Td, int d, int V, const void* w_hi16, const void* w_lo16, void* stream);

/* builders/utils.py:71-105 + split_heads (transformer.py:160-165), fused.
 * Q rows are b*Tq+i, K/V rows b*Tk+j; head h occupies columns [h*dh,(h+1)*dh) of each.
 * mask_mode: 0 none, 1 key padding (key_pad[b,j]!=0 masked), 2 causal | key padding.
 * Masked logits get -1e9 added after the 1/sqrt(dh) scaling exactly as the reference does.
 * lse[b,h,i] = logsumexp of the masked scaled logits (saved for backward).
 * backend: 0 = tcgen05 kind::tf32 (head dim 32, Tq,Tk <= 224), 2 = tcgen05 with the 3xTF32 split
 * (fp32-accurate, forward only), 1 = SIMT fp32 (any supported shape).                           */
int skf_attn_fwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V,
                 int64_t ldv, float* O, int64_t ldo, float* lse, const uint8_t* key_pad, int B,
                 int H, int Tq, int Tk, int dh, int mask_mode, int round_tf32, int backend,
                 void* stream);
/* delta is a scratch [B,H,Tq] buffer (rowsum(dO*O)).                                        */
int skf_attn_bwd(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V,
                 int64_t ldv, const float* O, int64_t ldo, const float* dO, int64_t lddo,
                 const float* lse, float* delta, float* dQ, int64_t lddq, float* dK,
                 int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad, int B, int H,
                 int Tq, int Tk, int dh, int mask_mode, int round_tf32, int backend, void* stream);

/* The whole-row kernels (backend 6: head dim 32, Tq, Tk <= 224) with operand images: the forward pass also writes the
 * bf16 hi / lo tiles of Q, K, V it built in shared memory to `img` (skf_attn_image_bytes(.., 0) bytes), and the backward
 * pass TMA-loads them straight into its operand slots instead of converting the fp32 rows again; `img_do`
 * (skf_attn_image_bytes(.., 1) bytes) carries the images of dO from the dQ kernel to the dK / dV kernel.  Same
 * arithmetic as skf_attn_fwd / skf_attn_bwd with backend 6; Tq == Tk.  This is what the plan runs in training.   */
size_t skf_attn_image_bytes(int B, int H, int Tq, int Tk, int which);
int skf_attn_fwd_img(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv, float* O,
                     int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Tq, int Tk, int dh,
                     int mask_mode, void* img, void* stream);
int skf_attn_bwd_img(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                     const float* O, int64_t ldo, const float* dO, int64_t lddo, const float* lse, float* delta,
                     float* dQ, int64_t lddq, float* dK, int64_t lddk, float* dV, int64_t lddv, const uint8_t* key_pad,
                     int B, int H, int Tq, int Tk, int dh, int mask_mode, const void* img, void* img_do, void* stream);

/* x + sublayer -> LayerNormalization(eps=1e-6) (transformer.py:209-224): s = a + b (+bias);
 * y = LN(s)*gamma+beta.  s, mean, rstd are saved for backward (any may be NULL).           */
int skf_add_layernorm_fwd(const float* a, const float* b, const float* bias, const float* gamma,
                          const float* beta, float* s, float* y, float* mean, float* rstd,
                          int rows, int d, int round_tf32, void* stream);
/* Same with dropout on the sublayer output (transformer.py:217-218, 221-222, 250-260):
 * s = a + dropout(b + bias), mask as in skf_dropout with element index row*d + col.          */
int skf_add_dropout_layernorm_fwd(const float* a, const float* b, const float* bias, const float* gamma,
                                  const float* beta, float* s, float* y, float* mean, float* rstd,
                                  int rows, int d, float rate, uint32_t seed, uint32_t site,
                                  const int64_t* step, void* stream);
/* ds = dLN(dy); dgamma += .., dbeta += .. (accumulated atomically into zeroed buffers)      */
int skf_layernorm_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                      const float* gamma, float* ds, float* dgamma, float* dbeta, int rows,
                      int d, int round_tf32, void* stream);
/* Same, additionally writing ds_drop = dropout_mask(ds) (the gradient of the sublayer output b of
 * skf_add_dropout_layernorm_fwd; ds itself is the gradient of the residual input a).          */
int skf_layernorm_dropout_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                              const float* gamma, float* ds, float* ds_drop, float* dgamma, float* dbeta,
                              int rows, int d, float rate, uint32_t seed, uint32_t site,
                              const int64_t* step, void* stream);
/* out[n] += sum_m x[m,n]  (bias gradients)                                                  */
int skf_colsum(const float* x, int64_t ldx, float* out, int M, int N, void* stream);

/* SelfAttnV1.call (transformer.py:61-74) after the tanh GEMM: score[b,t] = u[b,t,:].Vattn;
 * a = softmax over t (no padding mask); emb[b,:] = sum_t a[b,t] x[b,t,:].                  */
int skf_pool_fwd(const float* x, const float* u, const float* vattn, float* a, float* emb,
                 int B, int T, int d, int U, int round_tf32, void* stream);
/* Given demb: dx[b,t,:] (+)= a*demb ; dupre = (dscore*Vattn)*(1-u^2) ; dV += sum u*dscore    */
int skf_pool_bwd(const float* x, const float* u, const float* vattn, const float* a,
                 const float* demb, float* dx, int dx_accumulate, float* dupre, float* dvattn,
                 int B, int T, int d, int U, int round_tf32, void* stream);

/* DenseExpander.call (transformer.py:370-376): pre[b,t,c] = emb[b,c]*w[t] + beta[t].        */
int skf_expand_fwd(const float* emb, const float* w, const float* beta, float* pre, int B,
                   int T, int d, int round_tf32, void* stream);
int skf_expand_bwd(const float* dpre, const float* emb, const float* w, float* demb,
                   int demb_accumulate, float* dw, float* dbeta, int B, int T, int d,
                   void* stream);

/* Keys / values of the decoder's second attention block when they come from the expanded bottleneck
 * (models/sketchformer.py:170-181: expand_layer, then mha2.wk / wv of builders/layers/transformer.py:155-156):
 * pre[b,t,:] W + bias = w[t] U[b,:] + beta[t] s[:] + bias[:]  with U = emb W ([B,N]) and s = colsum(W) ([N]).
 * fwd writes KV [B,T,N]; bwd reduces G = dKV: gU[b,:] += sum_t w[t] G[b,t,:], gs += sum beta[t] G,
 * gbias += sum G, gw[t] += sum_{b,j} G U, gbeta[t] += sum_{b,j} G s   (N a multiple of 128 for bwd).          */
int skf_kv_expand_fwd(const float* U, const float* s, const float* bias, const float* w, const float* beta,
                      float* KV, int B, int T, int N, void* stream);
int skf_kv_expand_bwd(const float* G, const float* U, const float* s, const float* w, const float* beta,
                      float* gU, float* gs, float* gbias, float* gw, float* gbeta, int B, int T, int N,
                      void* stream);
/* The whole second attention block of a decoder layer on that rank-2 form, no K / V / score tensors: per query and
 * head a = Qh.UKh/sqrt(dh), c = Qh.sKh/sqrt(dh), e = Qh.bKh/sqrt(dh); P = softmax_t(a w[t] + c beta[t] + e - 1e9 pad[t]);
 * O = (sum P w) UVh + (sum P beta) sVh + bVh.  U [B, 2 H dh] (keys then values), s / bias [2 H dh], key_pad [B,T] or
 * NULL.  bwd: dQ, gU (overwritten), gs += , gbias += (value half; the key bias has zero gradient), gw += , gbeta += . */
int skf_xattn_fold_fwd(const float* Q, int64_t ldq, const float* U, const float* s, const float* bias, const float* w,
                       const float* beta, const uint8_t* key_pad, float* O, int64_t ldo, int B, int H, int Tq, int T,
                       int dh, void* stream);
int skf_xattn_fold_bwd(const float* Q, int64_t ldq, const float* U, const float* s, const float* bias, const float* w,
                       const float* beta, const uint8_t* key_pad, const float* dO, int64_t lddo, float* dQ,
                       int64_t lddq, float* gU, float* gs, float* gbias, float* gw, float* gbeta, int B, int H, int Tq,
                       int T, int dh, void* stream);
/* Y[r,:] += v[:] for every row (the colsum(W) term of the weight gradient above)                              */
int skf_add_row_broadcast(float* Y, const float* v, int rows, int N, void* stream);

/* Token reconstruction loss + accuracy (builders/losses.py:26-41, sketchformer.py:335-339):
 * loss = sum_rows CE(z,tgt)*[tgt!=0] / rows ; correct = #(argmax z == tgt) over ALL rows.
 * dlogits (may alias logits) = (softmax(z) - onehot)*[tgt!=0]*scale.  tgt row stride tgt_ld;
 * stats[0] += loss_sum (unnormalised), stats[1] += correct.  stats are doubles on device.   */
int skf_token_ce(const float* logits, int64_t ldl, const int64_t* tgt, int64_t tgt_ld, int B,
                 int Td, int V, float grad_scale, float* dlogits, int64_t lddl, double* stats,
                 int32_t* argmax_out, int round_tf32, void* stream);
/* Classification head loss (builders/losses.py:21-24): -log_softmax(z)[y] batch mean.       */
int skf_class_ce(const float* logits, int64_t ldl, const int64_t* labels, int B, int C,
                 float grad_scale, float* dlogits, int64_t lddl, double* stats,
                 int32_t* argmax_out, float* probs, void* stream);

/* TF-2.1 Adam with the Noam schedule evaluated on the device step counter
 * (builders/schedulers.py:25-29).  *step is read, used, and incremented.        */
int skf_adam_step(float* params, const float* grads, float* m, float* v, int64_t n,
                  int64_t* step, int d_model, int warmup_steps, float beta1, float beta2,
                  float eps, float grad_scale, void* stream);

/* ---------------- whole-step plan --------------------------------------------------------- */
typedef struct skf_plan skf_plan;

size_t skf_workspace_bytes(const skf_cfg* cfg);
/* params/grads/adam_m/adam_v: arenas of skf_param_count() floats; workspace of
 * skf_workspace_bytes() bytes, 256-byte aligned.  grads/adam may be NULL for inference.     */
int skf_plan_create(const skf_cfg* cfg, float* params, float* grads, float* adam_m,
                    float* adam_v, void* workspace, size_t workspace_bytes, skf_plan** out);
void skf_plan_destroy(skf_plan* plan);
/* One-time device initialisation of constants in the workspace (PE table, counters).       */
int skf_plan_init(skf_plan* plan, void* stream);
/* Offset (bytes) and size of a named workspace tensor, for tests / introspection.          */
int skf_plan_tensor(const skf_plan* plan, const char* name, size_t* offset, size_t* bytes);

/* Transformer.model_trainer (models/sketchformer.py:325-349): forward, losses, backward; when
 * apply_adam != 0 also the optimizer step.  tokens int64 [B,T] (or float [B,T,5] when
 * cfg.continuous), labels int64 [B,1]; both device pointers.  Metric sums accumulate in the
 * workspace tensor "stats" (doubles).                                                      */
int skf_train_step(skf_plan* plan, const void* tokens, const int64_t* labels, int apply_adam,
                   void* stream);
/* Only the optimizer (after an external gradient all-reduce).                               */
int skf_apply_adam(skf_plan* plan, void* stream);
/* Transformer.encode_from_seq / predict_class (sketchformer.py:162-168, 223-228):
 * writes embedding [B,d], class probabilities [B,C] and argmax [B] into the workspace.      */
int skf_encode(skf_plan* plan, const void* tokens, void* stream);
/* Transformer.call in inference mode incl. decoder logits for teacher-forced input.         */
int skf_forward(skf_plan* plan, const void* tokens, void* stream);
/* Transformer.predict_from_embedding (sketchformer.py:255-311): greedy decoding with per-layer key /
 * value caches.  skf_decode_begin: `embedding` device [B,d] (NULL = the workspace tensor "embedding" left
 * by skf_encode), `expected_len` device int32 [B] or NULL (cross-attention key mask j >= len when
 * blind_decoder_mask=0; ignored when blind); also classifies the embedding ("cls.argmax").  It expands the
 * embedding, projects every layer's cross keys / values once and seeds the prefix with the start token
 * (SOS, or the stroke row [0,0,1,0,0]).  skf_decode_steps runs decoder positions [first, first+n): each
 * appends argmax(logits) (or [dx,dy,softmax(pen)]) to the workspace tensor "gen.tokens" int64 [B,T+1]
 * ("gen.strokes" float [B,T+1,5]).  "gen.state" int32: [0] = number of valid output columns once every row
 * has emitted EOS (pen state 2), else 0; [1] = steps executed.  No host synchronisation inside.
 * blind_decoder_mask=0 with expected_len == NULL (sketchformer.py:275: the cross mask of step i attends the first
 * i + 1 keys for every position) cannot be cached: every step recomputes the prefix under its own mask — same
 * tokens as the reference, O(T^2) position steps per call instead of O(T).                                  */
int skf_decode_begin(skf_plan* plan, const float* embedding, const int32_t* expected_len,
                     int64_t sos_token, int64_t eos_token, void* stream);
int skf_decode_steps(skf_plan* plan, int first_step, int n_steps, void* stream);
/* DistributedStroke3DataLoader.preprocess + Tokenizer.encode for a batch of raw stroke-3 sketches
 * (dataloaders/distributed_stroke3.py:90-125, 138-153; utils/tokenizer.py:35-64; utils/sketch.py:31-50).
 * points: float [total, 3] rows (dx, dy, pen) of all sketches back to back; offsets: int64 [B+1] row ranges;
 * centroids: float [K, 2] (KMeans cluster_centers_); out: int64 [B, T] = [SOS] ids.. (SEP after pen lifts) [EOS]
 * PAD.., truncated to T.  Clip to +-limit, normalise by max(bbox_w, bbox_h, 1) of the absolute path.  */
int skf_tokenize_strokes(const float* points, const int64_t* offsets, int B, const float* centroids, int K,
                         float limit, int64_t* out, int T, void* stream);
/* Number of kernels launched by this plan since creation (bench `gpu_launches`).           */
int64_t skf_plan_launches(const skf_plan* plan);

/* Development aid: with skf_set_option("attn3_trace", 1) CTA 0 of the whole-row attention forward kernel stamps clock64()
 * at its phase boundaries ([role][item][16], see csrc/skf_attn_tc3.cu); copies up to n stamps to host memory `out`
 * (synchronises the device) and returns how many.                                                               */
int skf_debug_trace(int64_t* out, int n);

/* ---------------- data-parallel exchange --------------------------------------------------------
 * The reference is single-device (utils/gpu.py:4-15, train.py:41,69): these entry points are new work the north
 * star asks for.  One process per GPU; the only exchange of a step is a sum of the flat gradient arena over the
 * ranks (1/world is folded into the optimizer kernel, skf_cfg.world_size).  NCCL is bound at run time; every
 * failure of it is reported as SKF_ERR_NCCL with NCCL's own message.
 *  skf_nccl_version     : NCCL version code (e.g. 22809), or SKF_ERR_NCCL when no libnccl.so.2 can be loaded.
 *  skf_comm_unique_id   : rank 0 fills a 128-byte ncclUniqueId; the caller broadcasts it to the other ranks by any
 *                         means (torch.distributed, a file, the environment).
 *  skf_comm_create      : ncclCommInitRank on the current CUDA device; *comm is an opaque handle.
 *  skf_allreduce_bucket : in-place ncclAllReduce(sum, fp32) of `count` floats at `ptr` on `stream`; asynchronous and
 *                         CUDA-graph capturable.
 *  skf_grad_buckets     : the contiguous arena ranges [lo[i], hi[i]) (in floats) a plan with `n_buckets` (1..4) reduces,
 *                         in the order in which the backward pass completes them (output layer and decoder first,
 *                         encoder embedding last); returns how many ranges were written (<= n_buckets).
 *  skf_plan_set_comm    : from now on skf_train_step all-reduces those buckets on a high-priority communication stream
 *                         as soon as the backward pass (main chain and weight-gradient side stream) has produced
 *                         them, and the optimizer waits for the last one — all inside the step, so a CUDA graph of
 *                         the step contains the exchange.  comm = NULL turns it off.                              */
int skf_nccl_version(void);
int skf_comm_unique_id(void* id128);
int skf_comm_create(const void* id128, int world_size, int rank, void** comm);
int skf_comm_destroy(void* comm);
int skf_allreduce_bucket(float* ptr, int64_t count, void* comm, void* stream);
int skf_grad_buckets(const skf_cfg* cfg, int n_buckets, int64_t* lo, int64_t* hi);
int skf_plan_set_comm(skf_plan* plan, void* comm, int n_buckets);

#ifdef __cplusplus
}
#endif
#endif /* SKF_H_ */
