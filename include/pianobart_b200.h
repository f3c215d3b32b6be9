/* pianobart_b200 - C ABI of the sm_100a kernels behind the PianoBART hot path.
 *
 * The reference (RS2002/PianoBart) is pure Python and has no FFI layer: its "operator
 * interface" for this path is the Python module API (PianoBart.forward PianoBart.py:56-80,
 * PianoBartLM.forward model.py:20-66, Pretrainer.gen_mask / compute_loss pretrain.py:112-118,
 * 211-546).  The host-side mirror of that API lives in pianobart_b200/*.py; every device
 * operation it performs is one of the entry points below, bound with ctypes
 * (pianobart_b200/_lib.py).  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends
 *     in _host; no torch types; the caller owns every buffer (nothing is allocated here).
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and asynchronous.
 *   - return value 0 = success; non-zero = failure, message via pb_last_error().
 *   - dtype: 0 = fp32 buffers (parity mode), 1 = bf16 buffers (production mode).
 *   - there is NO CPU fallback: without a CUDA device these functions fail.
 */
#ifndef PIANOBART_B200_H
#define PIANOBART_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_DTYPE_F32 0
#define PB_DTYPE_BF16 1

/* ------------------------------------------------------------------ misc */
const char* pb_last_error(void);
int pb_version(void);
/* number of kernels launched by this library since the last reset (bench: gpu_launches) */
long long pb_launch_count(void);
void pb_reset_launch_count(void);
/* Programmatic dependent launch: when on (default; PIANOBART_B200_PDL=0 turns it off) consecutive kernels of a stream
 * overlap their prologue with the previous kernel's tail (every kernel waits for its predecessor with
 * griddepcontrol.wait before touching global memory).  Returns the previous setting.                              */
int pb_set_pdl(int on);

/* ------------------------------------------------------------------ GEMM
 * C[b,h][m,n] = epi(alpha * sum_k A[b,h][m,k] * B[b,h][n,k])
 * Replaces every nn.Linear / torch.bmm on the path: PianoBart.py:68,71 (in_linear),
 * HF modeling_bart.py BartAttention q/k/v/out_proj + eager_attention_forward matmuls,
 * BartEncoderLayer/BartDecoderLayer fc1/fc2, model.py:119-126 (8 heads as one N=1280 GEMM),
 * and autograd's backward products for them.
 *
 * a_mn_major = 0: A stored [M][K] (k contiguous), lda = row stride in elements
 * a_mn_major = 1: A stored [K][M] (m contiguous), lda = row stride in elements   (same for B / N)
 */
#define PB_GEMM_OUT_F32 1    /* C is fp32 (default: same dtype as the inputs)                */
#define PB_GEMM_GELU 2       /* exact erf GELU after bias (HF ACT2FN["gelu"])                */
#define PB_GEMM_ATOMIC_ACC 4 /* C += result with fp32 atomics (grad accumulation, split-K)   */
#define PB_GEMM_RES_F32 8    /* residual is fp32 (default: same dtype as C)                  */
#define PB_GEMM_AUX_PREACT 16 /* also store the pre-activation (after bias) to aux (dtype of C) */
#define PB_GEMM_MUL_DGELU 32  /* multiply the result by gelu'(aux[m,n]) (fc1 backward)          */
#define PB_GEMM_AUX_DGELU 64  /* with GELU: also store gelu'(pre-activation) to aux (dtype of C) - forward of fc1; the
                                 erf / exp evaluation is shared with the activation itself                        */
#define PB_GEMM_MUL_AUX 128   /* multiply the result by aux[m,n] (fc1 backward against a stored gelu')            */

typedef struct pb_gemm_desc {
  const void* a;
  const void* b;
  void* c;
  const float* bias;    /* [N] fp32 or NULL */
  const void* residual; /* added after activation, or NULL */
  int M, N, K;
  int a_mn_major, b_mn_major;
  long long lda, ldb, ldc, ldr;
  int batch_h, batch_b; /* two batch dimensions (0/1 = none) */
  long long a_stride_h, a_stride_b, b_stride_h, b_stride_b;
  long long c_stride_h, c_stride_b, r_stride_h, r_stride_b;
  float alpha;
  int flags;
  int split_k; /* >1 requires OUT_F32|ATOMIC_ACC */
  int causal;  /* 0 none; 1 skip output tiles with n > m (scores); 2 limit k to <= m (P.V)      */
  int block_n; /* 0 = auto, else 128 or 256 (tcgen05 path only)                                 */
  void* aux;   /* see PB_GEMM_AUX_* / PB_GEMM_MUL_*; row stride ldaux, not batched                 */
  long long ldaux;
  int r_row_mod; /* >0: residual row index = m % r_row_mod (learned position table broadcast over batch) */
  /* training-mode dropout on (acc + bias [+ activation]) BEFORE the residual is added (HF Bart*Layer.forward):
   * element (m, n) is kept iff hash(*drop_seed, drop_op, m*N + n) < drop_thresh and scaled by drop_scale.
   * drop_seed == NULL disables it. */
  const unsigned long long* drop_seed;
  unsigned int drop_op;
  unsigned int drop_thresh;
  float drop_scale;
  int cta_group; /* tcgen05 path: 0 = auto, 1 = single-CTA tiles, 2 = CTA pairs (cta_group::2, 256 x 256 tiles) */
} pb_gemm_desc;

/* bf16 operands, tcgen05/TMEM/TMA kernel */
int pb_gemm_bf16(const pb_gemm_desc* d, void* stream);
/* fp32 operands, SIMT kernel - the fp32 parity mode of the same graph */
int pb_gemm_f32(const pb_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------ fused attention (tcgen05, head_dim 128)
 * Replaces the core of HF BartAttention (eager_attention_forward: softmax(Q K^T * hd^-0.5 + mask) V) and its
 * autograd without materialising scores: encoder self-attention (key padding), decoder self-attention
 * (causal and key padding), decoder cross-attention (encoder key padding).  All tensors bf16, laid out
 * [B, S, H*hd] with row stride ld* (elements) so that slices of the fused QKV activation are used in place.
 * lse fp32 [B,H,Sq] (log2 domain) is written by forward and read by backward; dvec fp32 [B,H,Sq] is scratch.
 * Outputs (o, dq, dk, dv) are written with bulk tensor stores: pointers 16-byte aligned, row strides multiples of 8
 * elements.  Forward supports Sk <= 8192 (the key-padding bitmap of a sequence is kept in shared memory). */
typedef struct pb_attn_desc {
  const void* q; const void* k; const void* v;
  void* o;            /* forward output / backward input */
  const void* dout;   /* backward: gradient wrt o */
  void* dq; void* dk; void* dv;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float* lse; float* dvec;
  const uint8_t* key_keep; /* [B, Sk] non-zero = key visible, or NULL */
  int B, H, Sq, Sk, hd, causal;
  float scale;
} pb_attn_desc;
int pb_attn_fwd(const pb_attn_desc* d, void* stream);
int pb_attn_bwd(const pb_attn_desc* d, void* stream);
/* pb_attn_bwd in two parts: _prep fills dvec = rowsum(dout * o) (HBM-bound; may run on another stream next to an unrelated
 * GEMM), _main computes dq / dk / dv from a dvec that is complete */
int pb_attn_bwd_prep(const pb_attn_desc* d, void* stream);
int pb_attn_bwd_main(const pb_attn_desc* d, void* stream);
/* test hook (fault injection): the TMA producer threads of the attention kernels launched after this call stall a
 * pseudo-random number of cycles (< max_cycles) before each tile load, so tiles arrive late relative to the softmax warps
 * and the MMA thread; 0 turns it off.  Results must not depend on it (tests/test_gpu_parity.py).  Returns the old value. */
int pb_debug_set_attn_delay(int max_cycles);

/* ------------------------------------------------------------------ Octuple front end
 * out[m, 256*i + c] = table[row_off(i) + ids[m,i], c] for the 8 attributes (reference PianoBart.py:9-16,60-67:
 * eight nn.Embedding gathers * sqrt(256) + torch.cat).  `table` is the [1280,256] concatenation of the eight
 * tables in the activation dtype, already multiplied by 16.  ids: [M,8] int32 (ids_int64 = 0) or int64 (= 1).
 * n_tokens_host: 8 ints on the HOST.  err_flag (device int, may be NULL) is set to 1 on an out-of-range id. */
int pb_octuple_embed_fwd(const void* ids, int ids_int64, const void* table, void* out, long long M,
                         const int* n_tokens_host, int dtype, int* err_flag, void* stream);
/* dtable[row, c] += scale * dx[m, 256*i + c]  (fp32 atomics; autograd of the gathers above) */
int pb_octuple_embed_bwd(const void* ids, int ids_int64, const void* dx, float* dtable, long long M,
                         const int* n_tokens_host, float scale, int dtype, void* stream);

/* Fused front end (PianoBart.py:60-71 + modeling_bart.py:521-526 / 649-655) - the concatenated embedding [M,2048] never
 * exists: table_proj [1280, d] holds T[off_a + r] = 16 E_a[r] W_a^T (in_linear restricted to attribute a's 256 columns,
 * tabulated once per step by 8 small GEMMs), and per token
 *   y0[m] = sum_a T[off_a + ids[m,a]] + bias + pos_rows[m % S]          (pre-LayerNorm sum, kept for backward)
 *   h0[m] = dropout(LayerNorm(y0[m]; gamma, beta))                       (out_site as in pb_layernorm_fwd_drop)
 * d <= 2048 (bf16) / 1024 (fp32), multiple of the 16-byte pack.  pos_rows points at row 2 of the learned position table. */
typedef struct pb_drop_site pb_drop_site;
int pb_octuple_front_fwd(const void* ids, int ids_int64, const void* table_proj, const float* bias, const void* pos_rows,
                         int S, const float* gamma, const float* beta, void* y0, void* h0, float* mean, float* rstd,
                         long long M, int d, const int* n_tokens_host, float eps, const pb_drop_site* out_site, int dtype,
                         int* err_flag, void* stream);
/* one-hot rows of the Octuple ids: out[m, off_a + ids[m,a]] = 1 (a = 0..7), 0 elsewhere; out [M, sum(n_tokens)] in `dtype`.
 * Backward of the fused front end: G = onehot^T dy0 as one GEMM replaces the scatter-add of pb_octuple_embed_bwd. */
int pb_octuple_onehot(const void* ids, int ids_int64, void* out, long long M, const int* n_tokens_host, int dtype, void* stream);
/* Block-diagonal copy of the fused embedding table: out[off_a + r, emb_dim a + c] = emb[off_a + r, c]; out is
 * [sum(n_tokens), 8 emb_dim] in `dtype`, its off-diagonal blocks must have been zeroed once by the caller.  Turns the eight
 * per-attribute products of PianoBart.py:60-71 (tables x in_linear slices) and their weight gradients into single GEMMs. */
int pb_octuple_blockdiag(const void* emb, void* out, int emb_dim, const int* n_tokens_host, int dtype, void* stream);
/* g_emb[off_a + r, c] += alpha * dfull[off_a + r, emb_dim a + c]   (fp32; dfull = G W_in, [sum(n_tokens), 8 emb_dim]) */
int pb_octuple_blockdiag_grad(const float* dfull, float* g_emb, int emb_dim, const int* n_tokens_host, float alpha, void* stream);
/* stream-ordered memset(ptr, 0, bytes) */
int pb_fill_zero(void* ptr, long long bytes, void* stream);

/* ------------------------------------------------------------------ LayerNorm (HF BartEncoder/DecoderLayer
 * self_attn_layer_norm / encoder_attn_layer_norm / final_layer_norm / layernorm_embedding, eps 1e-5)
 * mean/rstd: [M] fp32 saved for backward.  d % 8 == 0 (bf16) / % 4 (fp32), d <= 2048 / 1024. */
int pb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     long long M, int d, float eps, int dtype, void* stream);
/* dropout sites (see pb_gemm_desc): `out` site masks the LayerNorm OUTPUT (layernorm_embedding -> dropout);
 * for backward, `in` site masks the incoming dy the same way, `out` site produces the second tensor dx_drop =
 * mask(dx) (gradient of the dropped-out Linear output feeding this LayerNorm) and dbias then sums dx_drop. */
struct pb_drop_site {
  const unsigned long long* seed; /* NULL = disabled */
  unsigned int op, thresh;
  float scale;
};
int pb_layernorm_fwd_drop(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                          long long M, int d, float eps, const pb_drop_site* out_site, int dtype, void* stream);
int pb_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dbias, long long M, int d,
                          const pb_drop_site* in_site, const pb_drop_site* out_site, int dtype, void* stream);
/* test hook: mask[i] = 1 iff element i of site (seed, op) is kept */
int pb_dropout_mask(const unsigned long long* seed, unsigned int op, unsigned int thresh, unsigned char* mask,
                    long long n, void* stream);
/* *ptr += inc  (per-step dropout seed bump, keeps the host out of the loop) */
int pb_add_u64(unsigned long long* ptr, unsigned long long inc, void* stream);
/* dx written; dgamma/dbeta accumulated (+=) with fp32 atomics; dbias (may be NULL) += column sums of dx, i.e. the
 * bias gradient of the Linear whose output (+ residual) fed this LayerNorm */
int pb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     void* dx, float* dgamma, float* dbeta, float* dbias, long long M, int d, int dtype, void* stream);

/* ------------------------------------------------------------------ masked softmax over attention scores
 * (HF eager_attention_forward: scores + mask -> softmax; masks of create_bidirectional_mask / create_causal_mask
 * built from pretrain.py:151-153).  scores fp32 [B,H,Sq,Sk] (already scaled); key_keep uint8 [B,Sk] or NULL
 * (non-zero = key visible); causal != 0: key j visible to query i only if j <= i.  Masked probabilities are
 * exactly 0; Sk <= 1024. */
int pb_softmax_fwd(const float* scores, void* probs, const uint8_t* key_keep, int B, int H, int Sq, int Sk,
                   int causal, int dtype, void* stream);
/* dscores = probs * (dprobs - sum_j probs_j dprobs_j); may run in place (dscores == probs) */
int pb_softmax_bwd(const void* probs, const float* dprobs, void* dscores, const uint8_t* key_keep, int B, int H,
                   int Sq, int Sk, int causal, int dtype, void* stream);

/* out[n] += sum_m x[m, n]   (bias gradients; also d(position table) with x viewed as [B, S*d]) */
int pb_colsum(const void* x, float* out, long long M, int N, long long ld, int dtype, void* stream);

/* ------------------------------------------------------------------ fused multi-head masked cross-entropy
 * Replaces pretrain.py:163-189 (8x permute + CrossEntropyLoss(reduction='none') * mask, host argmax accuracy)
 * and its autograd.  logits fp32 [M, sum(seg_sizes)], targets int32 [M,nseg], mask fp32 [M,nseg],
 * den[nseg] = (global) mask sums.  Accumulates loss_num[s] += CE*mask, correct[s] += (argmax==target)*mask and, if
 * dlogits != NULL, writes dlogits = (softmax - onehot) * mask * w[s] / (sum(w) * den[s]) * grad_scale in `dtype`.
 * argmax_out (int32 [M,nseg]) optional.  seg_sizes_host / weights_host live on the HOST. */
int pb_heads_ce(const float* logits, const int* targets, const float* mask, const float* den, float* loss_num,
                float* correct, void* dlogits, int* argmax_out, long long M, int nseg, const int* seg_sizes_host,
                const float* weights_host, float grad_scale, int dtype, void* stream);
/* The same computation FUSED with the heads' GEMM (north-star fusion 3; tcgen05, bf16): logits = h W^T + bias are
 * accumulated in TMEM per 128-row tile and reduced there - they never reach HBM.  h bf16 [M, K] (row stride ldh), w bf16
 * [sum(seg_sizes), K] (model.py:119-126, the eight Linear heads stacked), bias fp32; the other arguments as for pb_heads_ce.
 * dlogits bf16 [M, sum(seg_sizes)] (NULL in evaluation).  K % 64 == 0, every segment 32..496 classes, sum % 8 == 0. */
int pb_heads_ce_fused(const void* h, long long ldh, const void* w, const float* bias, const int* targets, const float* mask,
                      const float* den, float* loss_num, float* correct, void* dlogits, int* argmax_out, long long M, int K,
                      int nseg, const int* seg_sizes_host, const float* weights_host, float grad_scale, void* stream);
/* den[s] += sum_m mask[m, s]   (pretrain.py:117 denominators) */
int pb_mask_sums(const float* mask, float* den, long long M, int nseg, void* stream);

/* ------------------------------------------------------------------ optimizer (pretrain.py:76,195-196)
 * out[0] += sum(g^2) */
int pb_sumsq(const float* g, long long n, float* out, void* stream);
/* HF transformers 4.29 AdamW semantics + torch clip_grad_norm_ folded in (gnorm_sq = device scalar holding the
 * squared global gradient norm; max_norm <= 0 disables clipping).  p_bf16 (may be NULL) receives bf16(p*bf16_scale). */
int pb_adamw(float* p, float* m, float* v, const float* g, void* p_bf16, long long n, float lr, float beta1,
             float beta2, float eps, float wd, int step, const float* gnorm_sq, float max_norm, float grad_scale,
             float bf16_scale, void* stream);
/* y[m,:] = x[m,:] + table[m % S,:]  (decoder input embeddings supplied by the caller - reference
 * PianoBart.change_decoder_embedding, PianoBart.py:63-66,88-91 - plus BartLearnedPositionalEmbedding rows) */
int pb_add_rows_mod(const void* x, const void* table, void* y, long long M, int d, int S, int dtype, void* stream);
int pb_cast_from_f32(const float* src, void* dst, long long n, float scale, int dtype, void* stream);
int pb_cast_to_f32(const void* src, float* dst, long long n, int dtype, void* stream);

/* ------------------------------------------------------------------ BART noising (pretrain.py:128-153,211-546)
 * Applies a host-drawn plan (pianobart_b200/noising.py) to a batch: ori int16 [B,S,8]; src int32 [B,S] row codes
 * (>=0 source row, -1 PAD row, -2 MASK row, <=-3 row (-3-k) of rand_tok int32 [R,8]); loss_in uint8 [B,S];
 * loss_mode int32 [B] (0 host flags, 1 row-changed, 2 zero).  Outputs: enc_ids/dec_ids/targets int32 [B,S,8],
 * loss_mask fp32 [B,S,8], enc_keep/dec_keep uint8 [B,S] (Bar != <PAD>, computed after noising). */
int pb_noise_apply(const int16_t* ori, const int* src, const int* rand_tok, const uint8_t* loss_in,
                   const int* loss_mode, int* enc_ids, int* dec_ids, int* targets, float* loss_mask,
                   uint8_t* enc_keep, uint8_t* dec_keep, int B, int S, const int* pad_host, const int* mask_host,
                   const int* sos_host, void* stream);

/* ------------------------------------------------------------------ KV-cache decode (model.py:28-107)
 * The reference re-runs encoder + 1024-position decoder per generated token (model.py:42-45); these kernels
 * implement the same arithmetic incrementally.  bf16 only.  `t_dev` is the device-resident step index.
 *
 * pb_decode_finalize: acc fp32 [B,N] (split-K result of pb_gemm_bf16 with M = batch; zeroed again on exit) ->
 *   y = acc + bias; GELU if gelu; + residual (bf16 [B,N]); + pos_table[t+2] (bf16 rows of width N); LayerNorm if
 *   gamma != NULL; written as bf16 to `out` and/or fp32 to `out_f32`. */
int pb_decode_finalize(float* acc, const float* bias, const void* residual, const void* pos_table, const int* t_dev,
                       const float* gamma, const float* beta, void* out, float* out_f32, int B, int N, int gelu,
                       void* stream);
/* Small-batch (B <= 8) projection for the decode step: y[b,n] = epi(sum_k x[b,k] W[n,k]).  x_raw fp32 [B,K]; if gamma
 * != NULL the kernel first applies LayerNorm(gamma, beta) to x_raw (and stores the normalised vectors to x_norm_out,
 * the residual of the next sub-layer); W bf16 [N,K]; epilogue: + bias, GELU, + residual fp32 [B,N], + pos_table[t+2];
 * outputs fp32 and/or bf16 [B,N].  HBM-bound: each weight row is read once by one warp with 16-byte loads. */
int pb_decode_gemv(const float* x_raw, const float* gamma, const float* beta, float* x_norm_out, const void* W,
                   const float* bias, const float* residual, const void* pos_table, const int* t_dev, float* y_f32,
                   void* y_bf16, int B, int N, int K, int gelu, void* stream);
/* one query token per (batch, head).  q: bf16 [B, q_ld], head h at column h*hd.  Cache row j of batch b, head h at
 * k_cache + b*kv_batch_stride + j*kv_ld + h*hd (same for v).  append != 0: k_new/v_new (addressed like q) are
 * written to row t and keys 0..t are attended (HF BartAttention with past_key_values); else n_keys keys gated by
 * key_keep uint8 [B, n_keys] (cross attention on the encoder output).  Keys are split over ceil(max_keys/128) CTAs per
 * (batch, head); partial (max, sum, out) go through `workspace` and the last CTA (ticket) combines them. */
int pb_decode_attn(const void* q, int q_ld, const void* k_new, const void* v_new, void* k_cache, void* v_cache,
                   long long kv_batch_stride, int kv_ld, const uint8_t* key_keep, int n_keys, const int* t_dev,
                   int append, void* out, int out_ld, int B, int H, int hd, float scale, int max_keys,
                   float* workspace /* B*H*ceil(max_keys/128)*(hd+2) floats */, int* tickets /* B*H ints, zeroed once */,
                   float* out_f32 /* optional fp32 copy of the output (either output may be NULL) */, void* stream);
/* PianoBartLM.sample (model.py:68-78) + sampling/nucleus (model.py:84-107) for step t: logits fp32 [B, 1280];
 * uniforms double [B,S,8] drawn on the host from numpy's stream (one per attribute per step, as np.random.choice
 * consumes them); forced int32 [B,S,8] or NULL (teacher forcing).  Writes cur_tok int32 [B,8], sampled [B,S,8]. */
int pb_decode_sample(const float* logits, const double* uniforms, const int* forced, const int* t_dev, int* cur_tok,
                     int* sampled, int B, int S, const int* seg_sizes_host, const float* temp_host,
                     const float* top_p_host, void* stream);
/* model.py:59-65: stop when any attribute >= its <PAD> id (that step is not written), else result[b,t] = cur_tok;
 * then t += 1.  done int32 [B] sticky flags, n_written int32 [B]. */
int pb_decode_advance(const int* cur_tok, int* result, int* done, int* t_dev, int* n_written, int B, int S,
                      const int* pad_host, void* stream);

/* Generation post-processing on the device (reference demo.py:72-102, Octuple2Midi up to the MIDI writer): per sequence the
 * first row with any attribute >= its <PAD> id or a Pitch > 127 becomes the <EOS> row (PAD + 3) and later rows <PAD>; without
 * such a row the last row becomes <EOS>.  ids [B,S,8] int32 (ids_int64 = 0) or int64; out int64 [B,S,8]; len int64 [B] =
 * index of the <EOS> row = number of rows the reference hands to encoding_to_MIDI (0: "Generate Fail"). */
int pb_octuple_truncate(const void* ids, int ids_int64, long long* out, long long* len, int B, int S, const int* pad_host,
                        void* stream);

/* ------------------------------------------------------------------ finetuning classifier heads (all fp32)
 * SequenceClassification (model.py:128-143,195-218) and TokenClassification (model.py:236-272) around the backbone; the
 * wide Linears (1024 -> 128 / 256, 4096 -> 256, 64 -> 1024) are pb_gemm_* calls, these are the remaining pieces.
 * `act` folds the preceding activation into the operand: 0 none, 1 ReLU (nn.ReLU of both classifiers), 2 tanh
 * (SelfAttention, model.py:141).  Backward entry points ACCUMULATE (+=) into dw / dbias / dtable and overwrite dx / da / dp. */
/* nn.Dropout: y[i] = keep(site, i) ? x[i] * site->scale : 0 (same counter-based mask family as the backbone's sites);
 * applied to the incoming gradient it is its own backward */
int pb_dropout_apply(const float* x, float* y, long long n, const pb_drop_site* site, void* stream);
/* y[m, c] = sum_k act(x[m, k]) w[c, k] + bias[c],  N <= 16 outputs (Linear(256, class_num), SelfAttention.ws2) */
int pb_smalln_linear_fwd(const float* x, const float* w, const float* bias, float* y, long long M, int N, int K, int act,
                         void* stream);
/* dx[m, k] = act'(x[m, k]) sum_c dy[m, c] w[c, k] (dx may be NULL);  dw[c, k] += sum_m dy[m, c] act(x[m, k]);
 * dbias[c] += sum_m dy[m, c] (may be NULL) */
int pb_smalln_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias, long long M,
                         int N, int K, int act, void* stream);
/* softmax over the SEQUENCE axis of a [B, S, R] tensor (model.py:142 softmax(dim=1); padding rows are not masked), R <= 8 */
int pb_seq_softmax_fwd(const float* a, float* p, int B, int S, int R, void* stream);
int pb_seq_softmax_bwd(const float* p, const float* dp, float* da, int B, int S, int R, void* stream);
/* attention pooling (model.py:209 torch.bmm(attn_mat, x)): m[b, r, :] = sum_s p[b, s, r] x[b, s, :];
 * backward: dx[b, s, :] = sum_r p[b, s, r] dm[b, r, :], dp[b, s, r] = x[b, s, :] . dm[b, r, :] */
int pb_attn_pool_fwd(const float* p, const float* x, float* m, int B, int S, int R, int D, void* stream);
int pb_attn_pool_bwd(const float* p, const float* x, const float* dm, float* dx, float* dp, int B, int S, int R, int D,
                     void* stream);
/* Embeddings.forward (PianoBart.py:15-16) for a small table: out[m, :] = table[ids[m], :] * scale (ids int64; an id outside
 * [0, n_rows) sets *err_flag and yields a zero row);  backward: dtable[ids[m], :] += scale * dout[m, :], n_rows * d <= 12288 */
int pb_rows_gather(const long long* ids, const float* table, float* out, long long M, int n_rows, int d, float scale,
                   int* err_flag, void* stream);
int pb_rows_scatter_add(const long long* ids, const float* dout, float* dtable, long long M, int n_rows, int d, float scale,
                        void* stream);

/* ------------------------------------------------------------------ persistent whole-step decode (batch 1)
 * Replaces the per-token loop body of PianoBartLM.forward(generate=True) (model.py:42-65) for the default model geometry
 * (d 1024, 8 heads x 128, ffn 2048, 1280-entry Octuple vocabulary): ONE cooperative launch generates `n_steps` tokens.
 * Every SM streams its slice of the decoder weights and of the K/V caches through a shared-memory ring with bulk copies
 * that never wait for activations; activation vectors are exchanged between CTAs as tagged 8-byte words (payload + a
 * per-(token, hop) tag) instead of grid barriers; LayerNorm / bias / GELU / residual / sampling are fused into the
 * consumers (csrc/decode_persist.cu).
 * K/V caches use the split-major layout [head][split = key % 18][slot = key / 18][128] (bf16) per layer. */
#define PB_DECODE_MAX_LAYERS 8
#define PB_DECODE_NSPLIT 18
#define PB_DECODE_NSLOT 57
#define PB_DECODE_PART_WORDS 132 /* words per attention partial: max, sum, 2 pad, out[128] */
typedef struct pb_decode_layer {
  const void* wqkv; const float* bqkv;                 /* [3d,d] bf16 (q|k|v rows), [3d]            self_attn          */
  const void* wo; const float* bo;                     /* self_attn.out_proj                                           */
  const float* ln1_g; const float* ln1_b;              /* self_attn_layer_norm                                         */
  const void* wqc; const float* bqc;                   /* encoder_attn.q_proj                                          */
  const void* woc; const float* boc;                   /* encoder_attn.out_proj                                        */
  const float* ln2_g; const float* ln2_b;              /* encoder_attn_layer_norm                                      */
  const void* w1; const float* b1;                     /* fc1 [F,d]                                                    */
  const void* w2; const float* b2;                     /* fc2 [d,F]                                                    */
  const float* ln3_g; const float* ln3_b;              /* final_layer_norm                                             */
  void* self_k; void* self_v;                          /* [8][18][57][128] bf16, appended by the kernel                */
  const void* cross_k; const void* cross_v;            /* same layout, filled by pb_decode_kv_relayout                 */
} pb_decode_layer;
typedef struct pb_decode_persist_desc {
  pb_decode_layer layer[PB_DECODE_MAX_LAYERS];
  int n_layers, S_enc, S_max, stop_when_done;
  const void* emb_table;                               /* [1280,256] bf16, pre-scaled by 16                            */
  const void* w_in; const float* b_in;                 /* encoder_linear (= decoder_linear) [d,2048]                   */
  const void* pos_table;                               /* bart.decoder.embed_positions [max_pos+2, d] bf16             */
  const float* lne_g; const float* lne_b;              /* bart.decoder.layernorm_embedding                             */
  const void* w_heads; const float* b_heads;           /* 8 LM heads as one [1280,d]                                   */
  const uint8_t* enc_keep;                             /* [S_enc] encoder key-padding flags (non-zero = visible)       */
  /* generation state (device): same meaning as for pb_decode_sample / pb_decode_advance, batch 1 */
  int* t_dev; int* cur_tok; int* result; int* sampled; int* done; int* n_written;
  const double* uniforms; const int* forced;           /* [S_max,8]; forced may be NULL                                */
  float* logits_out;                                   /* [1280] fp32 logits of the last executed step (may be NULL)   */
  /* exchange buffers (device, zero-initialised once; 8-byte words) and the tag epoch */
  /* (the first seven are polled by every CTA and hold 4 words per 128-byte line: allocate 4x the word count) */
  unsigned long long* raw0; unsigned long long* raw1; unsigned long long* raw2;   /* 4 * d/2 words each                */
  unsigned long long* qkv;                             /* 4 * 3d/2 words                                               */
  unsigned long long* qc; unsigned long long* ob;      /* 4 * d/2 words each                                           */
  unsigned long long* f1;                              /* 4 * F/2 words                                                */
  unsigned long long* part;                            /* 8*18*PB_DECODE_PART_WORDS words                              */
  unsigned long long* logits_ll;                       /* 1280 words                                                   */
  unsigned long long* tok_ll;                          /* 8 words                                                      */
  unsigned int* epoch;                                 /* 1 word, zero-initialised once, advanced by the kernel        */
  int* error_flag;                                     /* set to a non-zero code before a bounded wait traps           */
  long long* trace;                                    /* developer hook (NULL in production): clock64 stamps of the hops
                                                          of the launch's LAST token, [gridDim][6 * 96] per CTA         */
  int dbg_flags;                                       /* developer experiment (0 in production): 1 = the producer does not
                                                          copy (no weight / KV traffic, results are garbage) - separates
                                                          the dependency-chain latency from the streaming time         */
} pb_decode_persist_desc;
/* generates up to n_steps tokens (stops early when stop_when_done and the stop rule fired); seg sizes / temperatures /
 * nucleus p / <PAD> ids live on the HOST */
int pb_decode_persist_run(const pb_decode_persist_desc* d, int n_steps, const int* seg_sizes_host, const float* temp_host,
                          const float* top_p_host, const int* pad_host, void* stream);
int pb_decode_persist_smem_bytes(void);
/* cross-attention K/V of one layer: projection layout [S_enc, 2d] (K | V, bf16) -> split-major cache layout */
int pb_decode_kv_relayout(const void* kv, void* k_out, void* v_out, int S_enc, void* stream);

/* ------------------------------------------------------------------ persistent whole-step decode (batch 64)
 * Same role as pb_decode_persist_run for a batch of 64 sequences (csrc/decode_batch.cu): one cooperative launch generates
 * n_steps tokens per sequence; projections on mma.sync tiles with the weights streamed ahead through a shared-memory ring,
 * LayerNorm / bias / GELU / residual fused, (sequence, head) attention units streaming their K / V from caches laid out
 * [sequence][head][key][128] (bf16), sampler and stop rule in the last phases; phases are separated by grid barriers.
 * All activation buffers are bf16 [64, width]; `layer[i].self_k/self_v` are [64][8][S_max][128], `cross_k/cross_v`
 * [64][8][S_enc][128] (pb_decode_kv_relayout_batch). */
typedef struct pb_decode_batch_desc {
  pb_decode_layer layer[PB_DECODE_MAX_LAYERS];
  int n_layers, S_enc, S_max, B;
  const void* emb_table; const void* w_in; const float* b_in; const void* pos_table;
  const float* lne_g; const float* lne_b; const void* w_heads; const float* b_heads;
  const uint8_t* enc_keep;                             /* [64, S_enc] or NULL                                           */
  int* t_dev; int* cur_tok; int* result; int* sampled; int* done; int* n_written;   /* as for pb_decode_sample / _advance */
  const double* uniforms; const int* forced;           /* [64, S_max, 8]; forced may be NULL                            */
  float* logits;                                       /* [64, 1280] fp32: logits of the last executed step             */
  void* xemb;                                          /* [64, 2048]                                                    */
  void* raw0; void* raw1; void* raw2; void* hn;        /* [64, d] each                                                  */
  void* qkv;                                           /* [64, 3d]                                                      */
  void* qc; void* ob;                                  /* [64, d] each                                                  */
  void* f1;                                            /* [64, F]                                                       */
  float* stats;                                        /* [3][64][2] fp32, zero-initialised once: LayerNorm row statistics  */
  unsigned int* barrier;                               /* grid-barrier counter (reset by every launch)                  */
  int* error_flag;
  long long* trace;                                    /* developer hook (NULL in production): clock64 before / after every
                                                          grid barrier of the launch's last token, [gridDim][2 * 80]     */
  float* part; int* part_cnt;                          /* cross-attention units of the last, partial round are cut into two
                                                          key halves: partial results [gridDim][132] fp32 and zero-initialised
                                                          pair counters [gridDim / 2]; both NULL = whole units only        */
} pb_decode_batch_desc;
int pb_decode_batch_run(const pb_decode_batch_desc* d, int n_steps, const int* seg_sizes_host, const float* temp_host,
                        const float* top_p_host, const int* pad_host, void* stream);
int pb_decode_kv_relayout_batch(const void* kv, void* k_out, void* v_out, int B, int S_enc, void* stream);

#ifdef __cplusplus
}
#endif
#endif
