/* dtts.h -- C ABI of the B200-native kernels behind detail_tts's synthesis hot path.
 *
 * Boundary: the reference (adelacvg/detail_tts) has no FFI; its hot path is plain PyTorch module
 * calls.  A maintainer swaps those module bodies for calls into this
 * library (ctypes stubs in INTEGRATION.md).  Every entry point takes DEVICE pointers owned by the
 * caller (torch tensors' data_ptr()), launches asynchronously on the given cudaStream_t and never
 * allocates device memory or synchronises.  Return 0 = ok, negative = error (dtts_last_error()).
 *
 * Activation layout ("rows"): channels-last [M, C] row-major with leading dimension ld (elements).
 * Utterance b owns rows [utt_off[b], utt_off[b]+utt_len[b]); rows between utterances are zero
 * separator rows (row_utt[m] == -1) that implement the reference's per-conv zero padding.
 *
 * Each function cites the reference code it replaces (file:line under /root/reference).
 */
#ifndef DTTS_H
#define DTTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DTTS_ABI_VERSION 2

/* activations usable in GEMM epilogues / elementwise kernels */
enum {
  DTTS_ACT_NONE = 0,
  DTTS_ACT_GELU_NEW = 1,   /* HF gelu_new, transformers modeling_gpt2.py MLP */
  DTTS_ACT_RELU = 2,       /* vqvae/modules/attentions.py:343 */
  DTTS_ACT_SILU = 3,       /* vqvae/diff_model.py:83,96,163 */
  DTTS_ACT_MISH = 4,       /* vqvae/modules/modules.py:496-501 */
  DTTS_ACT_LRELU = 5,      /* vqvae/model_24k.py:275,284; modules.py:317,320 (slope in act_param) */
  DTTS_ACT_TANH = 6,       /* vqvae/model_24k.py:286 */
  DTTS_ACT_LOG_CLAMP = 7,  /* log(max(x, act_param)): dynamic_range_compression_torch, vqvae/utils/data_utils.py:20-26 */
  /* pair activations: output column j is f(v[2j], v[2j+1]); N must be even, Nout = N/2.
     Weights are packed with the two halves of the reference's channel split interleaved. */
  DTTS_ACT_PAIR_TANH_SIGMOID = 16, /* fused_add_tanh_sigmoid_multiply, modules.py:15-22 */
  DTTS_ACT_PAIR_GLU = 17           /* Conv1dGLU x1*sigmoid(x2), modules.py:518-520 */
};

typedef struct {
  const void* A;        /* [M, K] fp16 (tc) or fp32 (simt), row-major, lda elements */
  const void* W;        /* [taps*N, K] same dtype as A, tap-major, ldw elements */
  int M, N, K, lda, ldw;
  int taps;             /* number of row taps (conv kernel size); 1 = plain GEMM */
  int tap_shift0;       /* row shift of tap 0 (e.g. -pad*dilation) */
  int tap_stride;       /* row shift increment per tap (dilation) */
  const float* bias;      /* [N] or NULL */
  const float* bias_utt;  /* [n_utt, N] per-utterance bias or NULL (needs row_utt) */
  const int* row_utt;     /* [M]: utterance id, or -1 => row is a separator: not stored by the register epilogues, stored as
                           * ZEROS by the TMA epilogues of dtts_gemm_f16_tc (separator rows are zero in every rows-layout buffer);
                           * NULL = all valid */
  const int* out_row_map; /* [M]: output/residual row index for A row m; NULL = identity */
  const float* res;       /* fp32 residual [*, Nout], ldr; added after the activation; NULL = none */
  float* out_f32;         /* fp32 output [*, Nout], ldo32; NULL = none */
  void* out_f16;          /* fp16 output [*, Nout], ldo16; NULL = none */
  int ldr, ldo32, ldo16;
  int act;                /* DTTS_ACT_* applied to acc+bias */
  int act16;              /* extra activation applied only to the fp16 copy (next GEMM's operand) */
  float act_param, act16_param;
  float alpha;            /* out = alpha*(act(acc+bias)+res) (+ out_old if accumulate) */
  int accumulate;
  /* dtts_gemm_tf32x3 only: */
  const void* A_lo;       /* fp32 low parts: A = A_hi + A_lo with A_hi tf32-exact (dtts_split_tf32) */
  const void* W_lo;       /* same for W */
  int split_k;            /* > 1: K is split over split_k CTAs per tile; raw partial sums are written to        */
  int64_t split_stride;   /*   out_f32 + s*split_stride (elements) and bias/act/res/out_f16 are ignored          */
  /* dtts_gemm_f16_tc only: GroupNorm statistics of the OUTPUT, accumulated by the epilogue (needs row_utt): */
  float* gn_stats;        /* [n_utt, N/gn_cpg, 2] (sum, sum of squares) += over this launch's rows; NULL = off */
  int gn_cpg;             /* channels per group (multiple of 8) */
} dtts_gemm_params;

/* D = epilogue(sum_taps A[m+shift_t,:] . W_t[n,:]) on tcgen05 tensor cores (fp16 operands, fp32
 * accumulation in TMEM, TMA-staged operands).  Replaces every dense contraction on the path:
 * nn.Conv1d / nn.Linear / ConvTranspose1d in vqvae/diff_model.py:81-99,160-209,
 * vqvae/utils/diff_util.py:199-203, vqvae/model_24k.py:236-267, vqvae/modules/modules.py:178-200,
 * 243-312, vqvae/modules/attentions.py:183-186,330-331, HF Conv1D in GPT2Block. */
int dtts_gemm_f16_tc(const dtts_gemm_params* p, void* stream);
/* Same contract in exact fp32 FMA arithmetic on CUDA cores (token-exact GPT path, small shapes). */
int dtts_gemm_f32(const dtts_gemm_params* p, void* stream);
/* Same contract at fp32-class accuracy on tcgen05 tensor cores: 3xTF32 (hi*hi + lo*hi + hi*lo, fp32
 * accumulation in TMEM) over fp32 operands pre-split into tf32-exact high and low parts.  This is the GPT
 * trunk's GEMM (HF Conv1D x@W+b in GPT2Block, gpt/model.py:149-173 via transformers modeling_gpt2.py):
 * token-exact sampling needs fp32-class logits.  With split_k > 1 the launch fills every SM for the
 * M = n_utterances decode step; dtts_splitk_reduce finishes the tile (fixed summation order). */
int dtts_gemm_tf32x3(const dtts_gemm_params* p, void* stream);

typedef struct {
  const void* x; int x_is_f16; int ldx;   /* [M, C] */
  int C, groups, n_utt;
  int max_len;                             /* longest utterance (sizes the smem cache); 0 = unknown */
  const int* utt_off; const int* utt_len;  /* [n_utt] */
  const float* gamma; const float* beta;   /* [C] */
  const float* film_scale; const float* film_shift; int ld_film;  /* [n_utt(or n_film), C] or NULL: y*(1+s)+b */
  const int* film_idx;                     /* [n_utt] row of film_* used by utterance b; NULL = b */
  int act;                                  /* DTTS_ACT_NONE or DTTS_ACT_SILU */
  float eps;
  float* out_f32; int ldo32; void* out_f16; int ldo16;
  int n_rows;                               /* rows of x (bounds of the TMA tensor map); 0 = unknown: no TMA staging */
} dtts_groupnorm_params;
/* GroupNorm32 (fp32 statistics over channels-in-group x frames of ONE utterance) + optional
 * timestep FiLM + SiLU: vqvae/utils/diff_util.py:113-133, vqvae/diff_model.py:107,113-115,242. */
int dtts_groupnorm(const dtts_groupnorm_params* p, void* stream);

typedef struct {
  const void* x; int x_is_f16; int ldx;   /* [M, C] */
  int M, C, cpg;
  const int* row_utt;                      /* [M]: utterance of each row, -1 = separator (not written) */
  const int* utt_len;                      /* [n_utt] frames per utterance (statistics are over utt_len * cpg values) */
  const float* stats;                      /* [n_utt, C/cpg, 2] (sum, sum of squares) from the producing GEMM (gn_stats) */
  const float* gamma; const float* beta;
  const float* film_scale; const float* film_shift; int ld_film; const int* film_idx;
  int act; float eps;
  float* out_f32; int ldo32; void* out_f16; int ldo16;
} dtts_gn_apply_params;
/* The apply half of GroupNorm32 when the statistics were accumulated by the epilogue of the GEMM that produced x
 * (dtts_gemm_params.gn_stats): one fully coalesced pass, y = (x - mean) * rstd * gamma + beta (+FiLM) (+SiLU). */
int dtts_groupnorm_apply(const dtts_gn_apply_params* p, void* stream);
typedef struct { float* ptr; int64_t n; } dtts_zero_params;
int dtts_zero_f32(const dtts_zero_params* p, void* stream);    /* ptr[0..n) = 0 (statistics buffers, inside a launch plan) */

typedef struct {
  const float* x; int ldx; int M, C;
  const float* gamma; const float* beta; float eps;
  const float* res; int ldr;      /* optional: normalises (x + res) */
  float* out_f32; int ldo32; void* out_f16; int ldo16;
  const int* row_utt;             /* optional [M]: rows with -1 are separators and are not written */
} dtts_layernorm_params;
/* LayerNorm over channels of each row: GPT2Block ln_1/ln_2/ln_f, gpt/model.py:322 final_norm,
 * vqvae/modules/modules.py:36-48 (channel LayerNorm in enc_p). */
int dtts_layernorm(const dtts_layernorm_params* p, void* stream);

enum { DTTS_ATTN_BIAS_NONE = 0, DTTS_ATTN_BIAS_RELPOS_TABLE = 1, DTTS_ATTN_BIAS_WINDOW_REL = 2 };
typedef struct {
  const void* q; const void* k; const void* v;  /* element (row r, head h, dim d) at ptr[r*ld + h*head_stride + d] */
  int is_f16; int ldq, ldk, ldv; int head_stride_q, head_stride_k, head_stride_v;
  int n_utt, n_heads, head_dim;
  const int* q_off; const int* q_len;    /* query rows of utterance b */
  const int* k_off; const int* k_len;    /* key rows of utterance b */
  int max_q_len, max_k_len;
  int causal;                 /* key j visible to query i iff j <= i + causal_offset[b] */
  const int* causal_offset;   /* [n_utt] or NULL (0) */
  float scale;                /* multiplies q.k */
  int bias_mode;
  const float* bias_table; int bias_half;   /* RELPOS_TABLE: [n_heads, 2*bias_half+1] indexed clamp(j-i)+bias_half */
  const float* rel_k; const float* rel_v; int window;  /* WINDOW_REL: [2*window+1, head_dim] shared over heads */
  float* out_f32; int ldo32; void* out_f16; int ldo16;  /* [rows, n_heads*head_dim] head-major */
  const int* o_off;           /* [n_utt] first output row of utterance b; NULL = q_off */
  float* out_lo; int ldo_lo;  /* decode fast path only: out_f32 receives the tf32-exact high part, out_lo the low part */
  int n_rows;                 /* dtts_attention_f16_tc only: rows of the q / k / v buffers (bounds of the TMA tensor maps) */
  /* decode fast path only (optional): the new token's q|k|v row is still the split-K partials of the QKV GEMM.
   * Each (utterance, head) warp first reduces its own 3*head_dim columns -- sum over qkv_splits in fixed order, + qkv_bias,
   * exactly dtts_splitk_reduce's arithmetic -- and stores them at row q_off[b] of q, k and v (column slices of one KV
   * arena whose newest key row is q_off[b]); saves the reduce launch of every layer of the latency-bound decode step.
   * Partials: qkv_ws[s*qkv_split_stride + b*qkv_ld_ws + part*n_heads*head_dim + h*head_dim + d], part = 0 q, 1 k, 2 v. */
  const float* qkv_ws; int qkv_splits; int64_t qkv_split_stride; int qkv_ld_ws; const float* qkv_bias;
} dtts_attention_params;
/* softmax(scale*q.k + bias) v in exact fp32 on CUDA cores, one query per CTA.  Covers the small
 * attentions: GPT-2 causal attention incl. KV-cache decode (modeling_gpt2.py:54-72),
 * MelStyleEncoder (modules.py:565-639), enc_p windowed rel-pos attention (attentions.py:198-239),
 * contextual_embedder / latent_conditioner (diff_util.py:145-169). */
int dtts_attention_f32(const dtts_attention_params* p, void* stream);
/* Flash-style tensor-core attention (fp16 operands, fp32 online softmax) for the diffusion
 * AttentionBlock hot loop: vqvae/utils/diff_util.py:145-169 + xtransformers.py:177-186.
 * Requires is_f16, head_dim 48, bias_mode NONE or RELPOS_TABLE, non-causal. */
int dtts_attention_f16_flash(const dtts_attention_params* p, void* stream);
/* Same contract on the 5th-generation tensor cores: persistent warp-specialised tcgen05 kernel, S and O accumulators
 * in TMEM, P fed back to the tensor core from TMEM, K/V chunks through a TMA ring, one softmax thread per query row.
 * Additionally requires n_rows and 16-byte aligned outputs. */
int dtts_attention_f16_tc(const dtts_attention_params* p, void* stream);

typedef struct {
  const float* logits; int ldl; int n_rows, vocab;
  const int64_t* ids; int ld_ids; int n_ids;     /* history incl. fake prefix (repetition penalty set) */
  const int* step_dev;                           /* optional device scalar added to n_ids (CUDA-graph replay) */
  float penalty, temperature, top_p; int top_k;
  int do_sample; int suppress_token;             /* -1 = none */
  float* probs; int ldp;                         /* do_sample: dense probabilities [n_rows, vocab] */
  int64_t* argmax;                               /* greedy: [n_rows] */
  float typical_mass;                            /* > 0: TypicalLogitsWarper(mass) between the penalty and the temperature */
} dtts_logits_params;
/* HF processor chain RepetitionPenalty -> [TypicalLogitsWarper] -> Temperature -> TopK -> TopP -> softmax (or argmax):
 * transformers generation/logits_process.py:298,407-410,522-532,582-585 as driven by
 * vqvae/model_24k.py:782-792 / gpt/model.py:540-544; the optional typical warper is the reference's own
 * gpt/modules/typical_sampling.py:5-33 (custom processors run after the penalty and before the sampling warpers). */
int dtts_process_logits(const dtts_logits_params* p, void* stream);

/* The whole KV-cached decode step of the GPT for a SMALL batch (B <= 32) as one persistent cooperative kernel: per layer
 * ln_1 -> c_attn -> attention over the KV arena -> c_proj + residual -> ln_2 -> c_fc -> gelu_new -> c_proj + residual, then
 * ln_f -> final_norm (latent) -> mel_head logits (gpt/model.py:107-185,392-417 + HF modeling_gpt2.py:229-310).  Exact fp32
 * FMA arithmetic; phases are separated by a grid barrier instead of ~84 launch boundaries.  Weights are plain fp32
 * [N, K] row-major (nn.Linear layout; HF Conv1D transposed). */
enum { DTTS_GPT_LAYER_PTRS = 13 };
typedef struct {
  int B, n_layers, d_model, n_heads, d_ff;
  /* device table [n_layers][DTTS_GPT_LAYER_PTRS] of device addresses, per layer in this order:
   * ln1_g ln1_b w_qkv[2304,768] b_qkv w_proj[768,768] b_proj ln2_g ln2_b w_fc[3072,768] b_fc w_out[768,3072] b_out arena */
  const uint64_t* layer_ptrs;
  const float* lnf_g; const float* lnf_b; const float* fn_g; const float* fn_b;     /* ln_f, final_norm */
  const float* w_head; const float* b_head; int vocab; int ld_logits;                /* mel_head [vocab, 768] */
  const float* x_in;            /* [B, 768] embedding of the current token (dtts_append_token output) */
  float* xa; float* xb;         /* [B, 768] scratch: residual stream ping-pong */
  float* att; float* u;         /* [B, 768], [B, 3072] scratch */
  float* part;                  /* [2, B, 768] scratch: split-K partials */
  float* hn;                    /* optional [B, 768]: final_norm output = the diffusion latent of this position */
  float* logits;                /* [B, ld_logits] */
  const int* kv_row;            /* [B] arena row of the new token */
  const int* k_off; const int* kv_len;   /* [B] first arena row / number of visible positions (incl. the new one) of utterance b */
  int arena_ld, max_k_len;      /* arena row pitch (>= 2304: q | k | v); bound on kv_len (shared-memory score buffer) */
  uint32_t* barrier;            /* [2], zero-initialised once by the caller: grid-barrier state */
  float ln_eps;
} dtts_gpt_step_params;
int dtts_gpt_decode_step(const dtts_gpt_step_params* p, void* stream);

typedef struct {
  int n_rows;
  const int64_t* next; int64_t* ids; int ld_ids; int n_ids;  /* ids[:, n_ids] = next (or stop if finished) */
  int* step_dev;                 /* optional device scalar: added to n_ids and pos, incremented by the kernel */
  int* unfinished; int* n_unfinished; int64_t stop_token;   /* n_unfinished: rows still running after this token */
  const float* tok_emb; const float* pos_emb; int pos; int dim;   /* next-step embedding row */
  float* x_out; int ldx;
  int* kv_row; int kv_stride;    /* optional [n_rows]: kv_row[b] = b*kv_stride + kv_pos0 + step (next KV/out row) */
  int kv_pos0; int* kv_len;      /* optional [n_rows]: kv_len[b] = kv_pos0 + step + 1 (keys visible next step) */
  const int* kv_pos_rows;        /* optional [n_rows]: per-row base position used instead of kv_pos0 (varlen prefixes) */
  float* x_stats;                /* optional [dim/128, n_rows, 2]: (sum, centred sum of squares) of every 128-column group of
                                    x_out -- the LayerNorm statistics format dtts_decode_gemm consumes (ln_stats) */
} dtts_append_params;
/* HF _sample bookkeeping (generation/utils.py:2797-2805) + next-token embedding
 * mel_embedding[id] + mel_pos_embedding[pos] (gpt/model.py:145-148). */
int dtts_append_token(const dtts_append_params* p, void* stream);

/* ---- the fused KV-cached decode step (round 2): 5 launches per GPT2Block instead of 8, no split-K partial buffers ------
 * dtts_decode_gemm: out[b, n] = act( sum_k LN(x)[b, k] * W[n, k] + bias[n] ) + res[b, n] for the B <= 128 utterance rows of
 * one decode step (HF GPT2Block's c_attn / c_proj / c_fc / mlp.c_proj and gpt/model.py:324 mel_head), "swap-AB" on tcgen05:
 * a 128-row slab of W is the M operand (TMA-staged, 128B swizzle), the activation rows are the N operand (N = B padded to
 * 16/32/64/128), 3xTF32 (hi*hi + lo*hi + hi*lo) with fp32 accumulators in TMEM.  The operand split x = hi + lo and the
 * optional LayerNorm in front of the GEMM (ln_1 / ln_2: transformers modeling_gpt2.py:262-310) happen while the activation
 * tile is written to shared memory, so there are no LayerNorm / split / reduce launches.  K is split over the CTAs of one
 * thread-block cluster (k_splits = 1, 2, 4 or 8); the partial accumulators are summed through distributed shared memory in
 * a fixed order (deterministic), each CTA finishing B/k_splits rows.  The weight TMA loads are issued before
 * griddepcontrol.wait, so under programmatic dependent launch (dtts_set_pdl) they overlap the predecessor's tail. */
typedef struct {
  const float* x; int ldx;               /* [B, K] fp32 activations */
  const float* W_hi; const float* W_lo;  /* [w_rows, K] fp32: tf32-exact high part / low part (pack.split_tf32_host) */
  int ldw, w_rows;                       /* w_rows >= N (rows beyond w_rows are TMA zero fill) */
  int B, N, K;                           /* B <= 128; K % (32 * k_splits) == 0 */
  const float* ln_stats; int ln_parts;   /* optional LayerNorm of x: [ln_parts, B, 2] (sum, centred sum of squares) over   */
  const float* ln_gamma; const float* ln_beta; float ln_eps;   /*   K/ln_parts columns each (out_stats of the producer)   */
  const float* bias; int act;            /* [N] or NULL; DTTS_ACT_NONE / DTTS_ACT_GELU_NEW */
  const float* res; int ldr;             /* optional residual [*, N] (indexed like out) */
  float* out; int ldo;                   /* [*, N] */
  const int* out_row_map;                /* optional [B]: output row of utterance b (KV arena row of the new token) */
  float* out_stats;                      /* optional [N/128, B, 2]: LayerNorm statistics of the OUTPUT rows per 128-column slab */
  int k_splits;                          /* cluster size along K */
} dtts_dgemm_params;
int dtts_decode_gemm(const dtts_dgemm_params* p, void* stream);

typedef struct {
  const float* x; int ldx; int B, C;     /* [B, C] residual stream after the last block */
  const float* g1; const float* b1;      /* ln_f   (HF GPT2Model.ln_f, gpt/model.py:322) */
  const float* g2; const float* b2;      /* final_norm (gpt/model.py:41,403); NULL = single LayerNorm */
  float eps;
  float* y; int ldy;                     /* [B, C] double-normed hidden = mel_head input */
  float* lat; int64_t lat_stride_b;      /* optional latent store: lat[b*lat_stride_b + (lat_pos0 + *step_dev)*C + c] = y */
  int lat_pos0; const int* step_dev;
  const int* row_step0; int lat_T;       /* continuous batching: row b's local step = *step_dev - row_step0[b]; positions >= lat_T (> 0) are not stored */
} dtts_final_ln_params;
/* ln_f -> final_norm of the B new positions (exact two-pass LayerNorm, one CTA per row) + the diffusion latent capture
 * (gpt/model.py:402-406: the latent of mel position j is this double-normed hidden). */
int dtts_final_ln(const dtts_final_ln_params* p, void* stream);

typedef struct {
  /* --- dtts_process_logits part --- */
  const float* logits; int ldl; int n_rows, vocab;
  int64_t* ids; int ld_ids; int n_ids;
  int* step_dev;                                   /* device step counter: read by every row, incremented once per launch */
  float penalty, temperature, top_p; int top_k;
  int do_sample; int suppress_token; float typical_mass;
  float* probs; int ldp;                           /* optional dense probabilities (debug / tests) */
  /* --- sampling: inverse CDF over the kept tokens in ascending id order, fp32 sequential sum, first id with cum > u --- */
  const float* uniforms; int ld_u;                 /* [n_steps, ld_u >= n_rows] pre-drawn U[0,1); row = *step_dev */
  /* --- dtts_append_token part --- */
  int* unfinished; int64_t stop_token;
  const float* tok_emb; const float* pos_emb; int pos; int dim;
  float* x_out; int ldx; float* x_stats;
  int* kv_row; int kv_stride; int* kv_len; const int* kv_pos_rows;
  uint32_t* done_counter;                          /* zero-initialised by the caller; self-resetting */
  /* --- continuous batching (slot reuse): row b decodes its own utterance since global step row_step0[b]; its history
   * column, mel position and KV row use the LOCAL step *step_dev - row_step0[b], the uniform row the global one.  A row is
   * finished by the stop token or by its max_new-th token; a finished row is FROZEN (nothing of it is written any more: its
   * codes / KV rows stay intact until the host harvests the slot and rebinds it to a waiting utterance). --- */
  const int* row_step0; int max_new;
} dtts_decode_tail_params;
/* dtts_process_logits + token choice (argmax, or inverse-CDF sampling from pre-drawn uniforms) + dtts_append_token in ONE
 * launch (one CTA per utterance row): the decode loop needs no host work per token (HF _sample: generation/utils.py:2743-2808). */
int dtts_decode_tail(const dtts_decode_tail_params* p, void* stream);

typedef struct {
  int M, C;                     /* rows, channels (128) */
  float* x; int ldx;            /* in/out fp32 state */
  const float* out_c; const float* out_u; int ldo;   /* [M, 2C] model outputs (eps | var) */
  const float* noise; int ldn;  /* [M, C] */
  float sqrt_recip, sqrt_recipm1, min_log, max_log, coef1, coef2, cfk, nonzero;
  void* x_f16; int ldx16;       /* optional fp16 copy of the new state */
} dtts_pstep_params;
/* One ancestral DDPM update with classifier-free guidance and learned-range variance:
 * vqvae/utils/diffusion.py:317-386,472-485. */
int dtts_p_sample_step(const dtts_pstep_params* p, void* stream);

typedef struct {
  const float* x; int ldx;       /* [M, Cp] fp32 stage input = ConvTranspose1d output (no activation); separators zero */
  int M, Cp;                     /* rows at this stage's rate; padded channel count (16 or 32) */
  const int* row_utt;            /* [M]: utterance of each row, -1 = separator (zero padding between utterances) */
  const void* w_frag;            /* fp16 weights of the 18 convs as mma.sync B fragments, conv order
                                    (k=3: c1_d1 c2 c1_d3 c2 c1_d5 c2), (k=7: ...), (k=11: ...); per conv [tap][Cp/8][Cp/16][32 lanes][4] */
  const float* bias;             /* [18, Cp] fp32, same conv order */
  float slope, slope_out;        /* leaky-ReLU slope inside the ResBlocks (0.1) / applied to the fp16 output */
  void* out_f16; int ldo16;      /* lrelu_slope_out(mean of the three ResBlocks): operand of the next ConvTranspose1d / conv_post */
  float* out_f32; int ldo32;     /* optional: the un-activated mean */
} dtts_voc_mrf_params;
/* Multi-receptive-field stage of the vocoder for the narrow stages, fused: xs = (RB_3(x) + RB_7(x) + RB_11(x)) / 3 with
 * RB_k = 3 x [lrelu -> conv(k, d in 1,3,5) -> lrelu -> conv(k, 1) -> + x]  (vqvae/model_24k.py:276-283,
 * vqvae/modules/modules.py:240-328).  One CTA keeps a 512-position tile (392 outputs + halo) in shared memory / registers
 * through all 18 convs; HBM traffic is x in, out16 out. */
int dtts_voc_mrf(const dtts_voc_mrf_params* p, void* stream);

typedef struct {
  const void* x; int ldx;        /* [M, 16] fp16 = lrelu_0.01(xs) of the last stage (12 channels padded to 16) */
  int M, C;                      /* rows, real channels (<= 16) */
  const int* row_utt;            /* [M] */
  const float* w;                /* [7, C] fp32 conv_post weight (tap-major), no bias */
  float* out; int ldo;           /* [M, ldo] waveform sample per row: tanh(conv) */
} dtts_conv_post_params;
/* conv_post + tanh: vqvae/model_24k.py:284-286 (Conv1d 12->1, k7, pad 3, no bias). */
int dtts_conv_post(const dtts_conv_post_params* p, void* stream);

typedef struct {
  const float* wav; int ldw; const int* wav_len;   /* [n_utt, ldw] fp32 waveforms (24 kHz), valid samples per utterance */
  int n_utt; const int* utt_off; const int* utt_len;   /* frame rows of utterance b: (wav_len + 2*pad - n_fft)/hop + 1 */
  int max_frames;
  int n_fft, hop, pad;          /* 1024, 256, (n_fft - hop)/2 = 384: reflect padding at both ends, center=False */
  const float* window;          /* [n_fft] periodic Hann (torch.hann_window); NULL = rectangular */
  float* out_hi; float* out_lo; int ld;   /* [M, n_fft] windowed frames split x = hi + lo (hi tf32-exact) */
  int zero_pad;                 /* 0: reflect padding (STFT); 1: zeros outside the waveform (polyphase resampler framing) */
} dtts_stft_frames_params;
/* Framing + windowing of mel_spectrogram_torch (vqvae/utils/data_utils.py:120-141: F.pad reflect + torch.stft framing).
 * The DFT itself is dtts_gemm_tf32x3 against a [cos | -sin] basis (fp32-class).  With zero_pad = 1 and no window the same
 * kernel frames the input of the sinc resampler api.py:37 applies to the prompt (torchaudio.transforms.Resample): frame
 * f = samples [f*orig - width, f*orig + width + orig), which dtts_gemm_tf32x3 multiplies with the [new, 2*width+orig] kernel. */
int dtts_stft_frames(const dtts_stft_frames_params* p, void* stream);

typedef struct {
  const float* spec; int lds;   /* [M, >= 2*n_bins] DFT output: real parts then imaginary parts */
  int M, n_bins; float eps;     /* 513 bins; sqrt(re^2 + im^2 + eps), eps = 1e-6 (data_utils.py:143) */
  float* out_hi; float* out_lo; int ld;   /* [M, ld >= n_bins] magnitudes split hi/lo, zero beyond n_bins */
} dtts_spec_mag_params;
int dtts_spec_mag(const dtts_spec_mag_params* p, void* stream);

/* elementwise / layout helpers (all on rows layout unless stated) */
typedef struct {
  const float* src; int B, C, T; const int* utt_off; const int* utt_len;  /* src [B, C, T] (reference layout) */
  float* dst_f32; int ld32; void* dst_f16; int ld16; float scale, shift;
} dtts_bct2rows_params;
int dtts_bct_to_rows(const dtts_bct2rows_params* p, void* stream);     /* [B,C,T] -> rows (x*scale+shift) */
typedef struct {
  const float* src; int ld; int B, C, T; const int* utt_off; const int* utt_len;
  float* dst; float scale, shift;                                       /* dst [B, C, T], zero beyond utt_len */
} dtts_rows2bct_params;
int dtts_rows_to_bct(const dtts_rows2bct_params* p, void* stream);
typedef struct {
  const float* x; int ldx; int M, C; int act; float act_param; float scale;
  float* out_f32; int ldo32; void* out_f16; int ldo16; const int* row_utt;
} dtts_eltwise_params;
int dtts_eltwise(const dtts_eltwise_params* p, void* stream);          /* out = act(x)*scale, masked rows -> 0 */
typedef struct {
  const int64_t* ids; int n; const float* table; int dim; const float* pos_table; const int* pos; /* pos[n] or NULL */
  float* out; int ldo;
  const int* dst_row;   /* optional [n]: output row of item i; NULL = i */
} dtts_embed_params;
int dtts_embed(const dtts_embed_params* p, void* stream);              /* gpt/model.py:134-136,517-519 */
typedef struct {
  const float* x; int ldx; int C; int n_utt; const int* utt_off; const int* utt_len; int repeat;
  float* out; int ldo; const int* out_off;
} dtts_repeat_rows_params;
int dtts_repeat_rows(const dtts_repeat_rows_params* p, void* stream);  /* F.interpolate nearest xN, diff_model.py:252 */
typedef struct {
  const float* x; int ldx; int C; int n_utt; const int* utt_off; const int* utt_len; float* out; int ldo;
} dtts_mean_rows_params;
int dtts_mean_rows(const dtts_mean_rows_params* p, void* stream);      /* masked temporal mean, modules.py:686-694; diff_model.py:228 */
typedef struct {
  const float* t; int n; int dim; float* out; int ldo; void* out_f16; int ldo16;
} dtts_tsemb_params;
int dtts_timestep_embedding(const dtts_tsemb_params* p, void* stream); /* vqvae/diff_model.py:20-38 */
typedef struct {
  float* x; int ldx; int M, half;            /* x [M, 2*half] */
  const float* m; int ldm; const int* row_utt;
  void* x0_f16; int ld16;                    /* optional: fp16 copy of x[:, :half] after the update (next pre-conv operand) */
  int flip_after;                            /* then flip channels in place (Flip, modules.py:393-400) */
} dtts_couple_params;
/* x1 = (x1 - m)*mask (m may be NULL = skip), modules.py:472-474, then optional Flip, then fp16 x0 copy */
int dtts_flow_couple(const dtts_couple_params* p, void* stream);
typedef struct {
  const float* m; const float* logs; int ld; const float* noise; int ldn; int M, C; float noise_scale;
  float* out; int ldo; const int* row_utt;
} dtts_zp_params;
int dtts_sample_zp(const dtts_zp_params* p, void* stream);             /* vqvae/model_24k.py:860 */
typedef struct {
  const void* src; int ld_src; int C; int n_utt;      /* fp16 rows */
  const int* src_off; const int* dst_off; const int* utt_len;   /* [n_utt] */
  void* dst; int ld_dst;
} dtts_copy_utts_params;
/* dst rows of utterance b <- src rows starting at src_off[b] (fp16): broadcasts the classifier-free-guidance
 * branch's timestep-integrator output, which depends only on (timestep, frame count), to every utterance of
 * that length (vqvae/diff_model.py:283-299, unconditioned_embedding.repeat + conditioning_timestep_integrator). */
int dtts_copy_utt_rows(const dtts_copy_utts_params* p, void* stream);
typedef struct { int* row_utt; int M; int n_utt; const int* utt_off; const int* utt_len; } dtts_rowutt_params;
int dtts_fill_row_utt(const dtts_rowutt_params* p, void* stream);

typedef struct {
  const float* x; int ldx; int M, C;
  float* hi; float* lo; int ld;
} dtts_split_params;
/* x = hi + lo with hi tf32-exact (low 13 mantissa bits zero): operand preparation for dtts_gemm_tf32x3. */
int dtts_split_tf32(const dtts_split_params* p, void* stream);

typedef struct {
  const float* ws; int64_t split_stride; int n_splits; int ld_ws;  /* partials [n_splits][M][ld_ws]; n_splits may be 0 */
  int M, N;
  const float* bias; int act; float act_param;
  const float* res; int ldr;           /* added after the activation */
  float* out_f32; int ldo32;           /* v = act(sum + bias) + res, row scattered through out_row_map */
  const int* out_row_map;
  const float* ln_gamma; const float* ln_beta; float ln_eps;   /* optional: y = LayerNorm(v) over the N columns (N <= 1024) */
  float* y_f32; int ldy;               /* optional fp32 copy of y (y = v when no LayerNorm) */
  float* y_hi; float* y_lo; int ld_hl; /* optional tf32 split of y: the next dtts_gemm_tf32x3 operand */
} dtts_reduce_params;
/* Deterministic split-K finish for the GPT decode step: fixed-order sum of the partial tiles + bias +
 * activation + residual, optionally fused with the LayerNorm that follows in GPT2Block (ln_1/ln_2/ln_f,
 * modeling_gpt2.py:262-310; gpt/model.py:322 final_norm) and the tf32 operand split of its output. */
int dtts_splitk_reduce(const dtts_reduce_params* p, void* stream);

/* library info */
int dtts_abi_version(void);
const char* dtts_last_error(void);
int dtts_sizeof(const char* struct_name);   /* sizeof(<struct>) for the binding's layout self-check */
int dtts_kernel_launches(void);             /* kernels launched by this library since load (bench gpu_launches) */
/* While on, the kernels of the GPT decode step (dtts_gemm_tf32x3, dtts_splitk_reduce, the KV-cache path of
 * dtts_attention_f32, dtts_process_logits) are launched with programmatic stream serialization (PDL): each is scheduled
 * while its predecessor drains and waits (griddepcontrol.wait) before touching memory.  Returns the previous setting. */
int dtts_set_pdl(int on);
int dtts_device_info(int* sm_count, int* cc_major, int* cc_minor);

#ifdef __cplusplus
}
#endif
#endif
