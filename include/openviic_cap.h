/*
 * openviic_cap.h -- C ABI of the B200-native caption hot path (libopenviic_cap.so).
 *
 * The reference (hieunghia-pat/OpenViIC) has NO native code and NO FFI: its hot path is Python
 * calling stock ATen ops behind a name->class Registry (builders/registry.py:8-90).  This header
 * is therefore the boundary *introduced* underneath the registered Python classes; each entry
 * point names the reference symbol (file:line, relative to the reference root) whose arithmetic
 * it replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions:
 *   - plain pointers and sizes only; no torch / C++ types in any signature;
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - activations are bf16 (uint16 storage), statistics / logits / log-probs fp32, ids int64 at the
 *     boundary (int32 inside), masks uint8 with 1 = masked (the reference's bool True = masked);
 *   - weights use nn.Linear layout [out, in] row-major;
 *   - every call takes an explicit cudaStream_t (as void*), never synchronises the device and is
 *     CUDA-graph capturable, except the *_host convenience calls which end with a stream sync;
 *   - return value: CAP_OK or an error code; cap_last_error() returns a thread-local message;
 *   - callers keep ownership of every buffer; the library allocates only inside handles
 *     (cap_beam_*, cap_engine_*) and frees in the matching destroy call.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef OPENVIIC_CAP_H
#define OPENVIIC_CAP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAP_ABI_VERSION 1

#define CAP_OK 0
#define CAP_ERR_INVALID 1 /* bad argument / unsupported shape */
#define CAP_ERR_CUDA 2    /* a CUDA runtime or driver call failed */
#define CAP_ERR_STATE 3   /* handle used in the wrong order */

typedef void* cap_stream_t; /* cudaStream_t */

enum cap_dtype { CAP_BF16 = 0, CAP_F32 = 1 };
enum cap_activation { CAP_ACT_NONE = 0, CAP_ACT_RELU = 1, CAP_ACT_SIGMOID = 2, CAP_ACT_LEAKY_RELU = 3 /* slope 0.01 */ };

int cap_abi_version(void);
const char* cap_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Per-operator entry points (used by the registered Python modules and by the parity tests)
 * ------------------------------------------------------------------------------------------ */

/* y[M,N] = act(x[M,K] . w[N,K]^T + bias[N]).  tcgen05/TMEM GEMM fed by TMA.
 * Replaces every nn.Linear on the path: models/modules/attentions.py:47-49,56,313-314;
 * models/modules/positionwise_feed_forward.py:24; models/modules/vision_embeddings.py:17;
 * models/modules/decoders.py:61,121.
 * x, w bf16; bias fp32 or NULL; y bf16 or fp32 (out_dtype).  ldx/ldy in elements; K % 8 == 0,
 * ldx % 8 == 0, x and w 16-byte aligned. */
int cap_linear(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy,
               int out_dtype, int act, int M, int N, int K, cap_stream_t stream);

/* Same contract on plain CUDA cores: the on-device cross-check for cap_linear in the tests. */
int cap_linear_simt(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy,
                    int out_dtype, int act, int M, int N, int K, cap_stream_t stream);

/* out = LayerNorm(residual + y) * gamma + beta [+ pos[row % pos_rows]] ; rows with
 * zero_rows[row] != 0 are written as 0.  y and residual are fp32 or bf16 (y_dtype / res_dtype; residual
 * may be NULL).  Two outputs, either may be NULL: `out` bf16 (the next GEMM's operand) and `out_f32`
 * (the next residual) -- the residual stream is fp32 end to end, as in the reference.
 * Replaces models/modules/attentions.py:308-309, positionwise_feed_forward.py:26,
 * encoders.py:20,36 (LN(x)+pos and the padded-row zeroing), decoders.py:26. */
int cap_add_layernorm(const void* y, int y_dtype, int ldy, const void* residual, int res_dtype,
                      int ldr, const float* gamma, const float* beta, float eps, const float* pos,
                      int pos_rows, const uint8_t* zero_rows, void* out, int ldo, float* out_f32,
                      int ldo32, int rows, int d, cap_stream_t stream);
/* Linear + residual + LayerNorm in ONE kernel: out = LayerNorm(residual + x.w^T + bias) * gamma + beta
 * [+ pos[row % pos_rows]], zero_rows as above; N = 128..1024 in steps of 128.  The N/128 CTAs of a row tile
 * form a thread-block cluster and exchange row statistics through distributed shared memory.
 * Replaces fc_o / fc2 followed by attentions.py:308-309 / positionwise_feed_forward.py:26. */
int cap_linear_layernorm(const void* x, int ldx, const void* w, const float* bias, const float* residual,
                         int ldr, const float* gamma, const float* beta, float eps, const float* pos,
                         int pos_rows, const uint8_t* zero_rows, void* out_bf16, int ldo, float* out_f32,
                         int ldo32, int M, int N, int K, cap_stream_t stream);

/* Visual-token padding mask + cast: mask[row] = (sum_k feats[row,k] == 0) (fp32 sum), and
 * out[row,:] = bf16(feats[row,:]).  feats fp32 or bf16.
 * Replaces models/utils.py:48-61 as called from models/modules/vision_embeddings.py:16. */
int cap_feature_mask_cast(const void* feats, int feat_dtype, void* out_bf16, uint8_t* mask,
                          int rows, int d_feature, cap_stream_t stream);

/* Box-relation bias g[B,H,n,n] = relu(W_g . embed(box_i, box_j) + b_g) (fp32).
 * w_g [H, d_g], b_g [H]; d_g = 4 (raw) or d_model/H (sin/cos) selected by `trig`.
 * Replaces models/utils.py:156-215 + models/modules/encoders.py:94-101. */
int cap_geometry_bias(const float* boxes, const float* w_g, const float* b_g, float* g, int B,
                      int n, int H, int d_g, int trig, cap_stream_t stream);

/* Locally-constrained mask of the dual-path encoder (models/utils.py:100-154 get_combine_masks): boxes fp32
 * [rows][4] in [0,1] image coordinates -> mask uint8 [rows][grid_size^2], 0 for the grid cells inside the box's
 * corner-to-corner cell rectangle, 1 (masked) elsewhere. */
int cap_region_grid_mask(const float* boxes, uint8_t* mask, int rows, int grid_size, cap_stream_t stream);

typedef struct cap_attention_args {
    const void* q;            /* bf16 (B, nq, H*64), row stride ldq, batch stride q_bs (elements) */
    const void* k;            /* bf16 (B, nk, H*64) */
    const void* v;            /* bf16 (B, nk, H*64) */
    void* out;                /* bf16 (B, nq, H*64) */
    int64_t q_bs, k_bs, v_bs, o_bs;
    int ldq, ldk, ldv, ldo;
    const uint8_t* mask;      /* (B, nq or 1, nk), 1 = masked; NULL = none */
    int64_t mask_bs;          /* batch stride of mask */
    int mask_qs;              /* query stride of mask (0 = broadcast over queries) */
    const float* geometry;    /* (B, H, nq, nk) fp32 >= 0, added as log(max(g,1e-6)); NULL = none */
    const void* mem_k;        /* bf16 (n_mem, H*64): sqrt(d_k)*m_k, never masked; NULL = none */
    const void* mem_v;        /* bf16 (n_mem, H*64): sqrt(n_mem)*m_v */
    int n_mem;
    int B, H, nq, nk;
    float scale;              /* 1/sqrt(d_k) */
    const void* sentinel;     /* bf16 (B, nq, H*64): query i's own extra key = value ("language signal" of
                                 AdaptiveScaledDotProductAttention, attentions.py:250-263), never masked; NULL = none */
    int64_t s_bs;             /* batch stride of sentinel (elements) */
    int lds;                  /* row stride of sentinel */
} cap_attention_args;

/* softmax(q.k^T*scale + mask + log g | memory slots | per-query sentinel).v per (batch, head); d_k = d_v = 64,
 * nk + n_mem <= 160.  Replaces the body of ScaledDotProductAttention.forward
 * (models/modules/attentions.py:51-55), AugmentedGeometry... (:104-111), AugmentedMemory... (:164-182) and
 * Adaptive... (:244-263) between the projections. */
int cap_attention(const cap_attention_args* args, cap_stream_t stream);

/* Decode-step self-attention over the beam-indirected KV cache (nq = 1 per row).
 *   qkv      bf16 [T][R][3*H*64]  (q|k|v of the token each row consumed at step t')
 *   ancestry int32 [T][R]  ancestry[t'][r] = cache row at step t' on row r's history (t' < t)
 *   padflag  uint8 [T][R]  1 if the token consumed at step t' by that cache row was <pad>
 * Replaces the stateful branch of MultiHeadAttention.forward (models/modules/attentions.py:
 * 297-304) + the running mask of Decoder.forward (decoders.py:101-103) + the per-step state
 * gather of BeamSearch._expand_state (beam_search.py:19-34) -- by indirection, no copy. */
int cap_decode_self_attention(const void* qkv, const int32_t* ancestry, const uint8_t* padflag,
                              void* out, int ldo, int t, int R, int H, float scale,
                              cap_stream_t stream);

/* Decode-step cross-attention: row r attends to image r / beam.  kv bf16 [B][n][2*H*64] (k|v,
 * projected once per image), key_mask uint8 [B][n].  out bf16 [R][H*64].
 * Replaces models/modules/decoders.py:23 (enc_attn) between its projections. */
int cap_decode_cross_attention(const void* q, int ldq, const void* kv, const uint8_t* key_mask,
                               void* out, int ldo, int B, int beam, int n, int H, float scale,
                               cap_stream_t stream);
/* The same queries over `levels` encoder levels in one launch (MeshedDecoderLayer, decoders.py:55-57): level i reads
 * kv + i * kv_level_stride elements and writes out + i * out_level_stride elements. */
int cap_decode_cross_attention_levels(const void* q, int ldq, const void* kv, size_t kv_level_stride,
                                      const uint8_t* key_mask, void* out, int ldo, size_t out_level_stride, int B,
                                      int beam, int n, int H, int levels, float scale, cap_stream_t stream);

/* x[r,:] = word_emb[token[r],:] + pos_table[position,:] (bf16 out); padflag_out[r] = token==pad.
 * Replaces models/modules/decoders.py:105-112 (stateful: position = t+1 for every row). */
int cap_embed_tokens(const int32_t* tokens, const void* word_emb_bf16, const float* pos_table,
                     int position, int pad_idx, void* out, float* out_f32, uint8_t* padflag_out, int R,
                     int d, cap_stream_t stream);

/* Meshed mix: out = sum_i sigmoid(a_i) * c_i / sqrt(levels)  (bf16), a_i fp32 [levels][R][d]
 * pre-activation gates, c_i bf16 or fp32 [levels][R][d]; bf16 and/or fp32 output.  Replaces models/modules/decoders.py:60-67. */
int cap_meshed_mix(const float* gates, const void* c, int c_dtype, void* out, float* out_f32,
                   int levels, int R, int d, cap_stream_t stream);

/* AoA gate: out = info * sigmoid(gate); ig fp32 [R][2*d] = (info | gate).  attentions.py:311-315 */
int cap_aoa_gate(const float* ig, void* out, float* out_f32, int R, int d, cap_stream_t stream);

/* out[r,:] = log_softmax(logits[r,:]) over V columns, fp32 (models/modules/decoders.py:123). */
int cap_log_softmax(const float* logits, int ld, float* out, int ldo, int rows, int V,
                    cap_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Beam-search state machine (models/modules/beam_search.py:36-118)
 * ------------------------------------------------------------------------------------------ */
typedef struct cap_beam cap_beam;

int cap_beam_create(int max_batch, int beam, int max_len, int vocab, int eos_idx, cap_beam** out);
int cap_beam_destroy(cap_beam* h);
/* seq_mask = 1, seq_logprob = 0, histories cleared, ancestry = identity; tokens = bos. */
int cap_beam_reset(cap_beam* h, int batch, int bos_idx, cap_stream_t stream);
/* One BeamSearch.iter(t): scores (R, ld) fp32 are either raw logits (is_logprob = 0: the
 * log-softmax of decoders.py:123 is fused in) or log-probs exactly as model.step returns them
 * (is_logprob = 1: selection is bit-exact with the reference's sort-based select()).
 * R = batch*beam rows at every t; at t = 0 only beam 0 of each image is a candidate (cur_beam = 1). */
int cap_beam_step(cap_beam* h, int t, const float* scores, int ld, int is_logprob,
                  cap_stream_t stream);
/* Vocabulary projection with log-softmax statistics: logits (M, ld) fp32 are stored once and, for every
 * row and every 32-column chunk, part_ms[row][chunk] = (max, sum exp(x - max)); *chunks_out = 8*ceil(N/256).
 * Replaces decoders.py:121-123 (fc + log_softmax) together with cap_beam_step_stats. */
int cap_vocab_logits_stats(const void* x, int ldx, const void* w, const float* bias, float* logits,
                           int ld, int M, int N, int K, float* part_ms, int* chunks_out,
                           cap_stream_t stream);
/* One BeamSearch.iter(t) from those statistics: log-sum-exp from the chunk pairs, then only the `beam`
 * chunks with the largest maxima are read back from `logits` (the row's best candidates lie there);
 * same state update as cap_beam_step.  Replaces the full sort of beam_search.py:37. */
int cap_beam_step_stats(cap_beam* h, int t, const float* logits, int ld, const float* part_ms,
                        int chunks, cap_stream_t stream);
/* Final descending sort by seq_logprob + gather; ids int64 (B,out_size,T), logp fp32 same shape. */
int cap_beam_finalize(cap_beam* h, int out_size, int64_t* ids, float* logp, cap_stream_t stream);
/* Device views of the running state (valid until destroy): */
const int32_t* cap_beam_tokens(cap_beam* h);    /* [R]    token each row consumes next       */
const int32_t* cap_beam_ancestry(cap_beam* h);  /* [T][R] see cap_decode_self_attention      */
const float* cap_beam_seq_logprob(cap_beam* h); /* [R]                                       */
const int32_t* cap_beam_parents(cap_beam* h);   /* [R]    selected_beam of the last step     */

/* ------------------------------------------------------------------------------------------
 * GEMM chains of a decode step (csrc/decode_fused.cu): BaseTransformer.step of the standard Decoder and of the
 * MeshedDecoder (decoders.py:95-123, 145-173 at nq = 1) as 1 + 2 * layers launches per step -- every CTA (pair)
 * carries a tile of 128 beam rows through a list of projection GEMMs with fused epilogues (bias, ReLU, residual +
 * LayerNorm, meshed gates + mix, vocabulary log-softmax statistics), the stand-alone attention kernels run between
 * the chains.  All pointers are DEVICE memory owned by the caller (the engine); the handle owns its scratch tiles and
 * (unless `stacked` is given) stacked copies of the weights.  Needs d_model 512, 8 heads, d_ff 2048, beam <= 5,
 * bias-free vocabulary projection; the meshed decoder with 3 encoder levels.
 * ------------------------------------------------------------------------------------------ */
typedef struct cap_fused_layer {
    const void *w_qkv, *w_o1, *w_q, *w_o2, *w_fc1, *w_fc2; /* bf16 [out,in]: self q|k|v, self fc_o, cross fc_q, cross fc_o, fc1, fc2 */
    const float *b_qkv, *b_o1, *ln1_g, *ln1_b;             /* self-attention biases + its LayerNorm */
    const float *b_q, *b_o2, *ln2_g, *ln2_b;               /* cross-attention */
    const float *b_fc1, *b_fc2, *ln3_g, *ln3_b;            /* feed-forward */
    int n_levels;                                          /* 0: DecoderLayer; 3: MeshedDecoderLayer (decoders.py:31-73) */
    const void* w_alpha[3];                                /* bf16 [d_model, 2 * d_model]: fc_alphas.i over [self_att ; enc_att_i] */
    const float* b_alpha[3];
} cap_fused_layer;

/* Stacked bf16 copies of the decoder's projection weights (one TMA tensor map then serves every GEMM of a chain).
 * A set built once can back any number of handles (cap_fused_desc::stacked): engines that pipeline independent
 * batches stream the same addresses, which stay L2-resident.  The caller keeps it alive while a handle uses it. */
typedef struct cap_fused_weights cap_fused_weights;
int cap_fused_weights_create(const cap_fused_layer* layers, int n_layers, cap_fused_weights** out);
int cap_fused_weights_destroy(cap_fused_weights* w);

typedef struct cap_fused_desc {
    int d_model, heads, d_ff, n_layers, vocab, max_len, beam, pad_idx;
    int max_rows;                 /* max_batch * beam */
    const cap_fused_layer* layers;
    const void* w_vocab;          /* bf16 [vocab, d_model] */
    const void* word_emb;         /* bf16 [vocab, d_model] */
    const float* word_pos;        /* fp32 [max_len + 1, d_model] */
    const int32_t* tokens;        /* [R]    cap_beam_tokens */
    uint8_t* padflag;             /* [T][R] written at step t, read at later steps */
    void* qkv_cache;              /* bf16 [layers][T][R][3*d_model] */
    float* logits;                /* fp32 [R][ld_logits], ld_logits % 32 == 0 */
    int ld_logits;
    float* part_ms;               /* fp32 [R][8*ceil(vocab/256)][2] */
    const void* att_in;           /* bf16 [max(levels,1)][max_rows][d_model]: outputs of cap_decode_{self,cross}_attention */
    void* q_out;                  /* bf16 [max_rows][d_model]: queries for cap_decode_cross_attention */
    const cap_fused_weights* stacked; /* NULL: the handle builds and owns its own stacked copies of `layers` */
} cap_fused_desc;

typedef struct cap_fused_decoder cap_fused_decoder;
int cap_fused_create(const cap_fused_desc* desc, cap_fused_decoder** out);
int cap_fused_destroy(cap_fused_decoder* f);
/* One chain of step t for B images (R = B*beam rows).  Per step:
 *   EMBED_QKV(0); for each layer L: cap_decode_self_attention -> SELF_OUT(L) -> cap_decode_cross_attention (one per
 *   encoder level, level i into att_in[i]) -> FFN(L); then cap_beam_step_stats.
 * EMBED_QKV: x = Emb[token] + pos, q|k|v of layer 0 into the cache slot of step t.  SELF_OUT(L): self fc_o + LN,
 * cross fc_q -> q_out (meshed: + the gates' self-attention part).  FFN(L): cross fc_o + LN (meshed: per level, then
 * gate and mix), fc1, fc2 + LN, and layer L+1's q|k|v projection or, after the last layer, the vocabulary projection
 * with its chunk statistics. */
enum cap_fused_chain_kind { CAP_CHAIN_EMBED_QKV = 0, CAP_CHAIN_SELF_OUT = 1, CAP_CHAIN_FFN = 2 };
int cap_fused_chain(cap_fused_decoder* f, int chain, int layer, int t, int B, cap_stream_t stream);
/* By default the vocabulary epilogue of the chains stores only the 32-column groups of logits that can contain one
 * of a row's `beam` best candidates (cap_beam_step_stats reads nothing else); on != 0 stores every logit (debug
 * views, cap_engine_decode_logits; OPENVIIC_FULL_LOGITS=1 sets it at creation). */
int cap_fused_set_full_logits(cap_fused_decoder* f, int on);
int cap_fused_get_full_logits(cap_fused_decoder* f);
/* Debug: when non-NULL (3 * 6 * 64 * 64 words of device memory), every later chain launch runs a tracing
 * instantiation that writes, for chain kind c and layer l, into region (c * 6 + l) * 4096 + tile * 64 + k:
 *   k = 0..4   issuer warp, SM clock cycles: total, waiting for weight stages, for free accumulators (epilogues),
 *              for the A tile, for streamed A blocks;
 *   k = 8..10  producer warp: total, waiting for free ring stages, for the hidden tile / free A slots;
 *   k = 40..47 %globaltimer stamps (ns) of the epilogue of fc1's fourth 256-column chunk: entry, bias staged,
 *              accumulator ready, the four 32-column groups stored, accumulator released. */
int cap_debug_fused_trace(unsigned long long* device_buffer);

/* ------------------------------------------------------------------------------------------
 * Encoder chains (csrc/decode_fused.cu): the vision embedding and the encoder layers (vision_embeddings.py:15-20,
 * encoders.py:17-40) on the chain kernel -- stage 0 = vision projection + LayerNorm + position table and layer 0's
 * q|k|v; stage 1 + l = layer l's fc_o + LN, fc1, fc2 + LN (padded rows zeroed, level output stored) and layer l + 1's
 * q|k|v or the decoder's cross-attention K|V projections of the level outputs it attends to; the encoder's
 * self-attention kernel (cap_attention on qkv_out -> att_in) runs between the stages.  d_model 512, d_ff 2048.
 * ------------------------------------------------------------------------------------------ */
#define CAP_ENC_MAX_KV 18
typedef struct cap_enc_layer {
    const void *w_qkv, *w_o, *w_fc1, *w_fc2;               /* bf16 [out,in] */
    const float *b_qkv, *b_o, *ln1_g, *ln1_b, *b_fc1, *b_fc2, *ln2_g, *ln2_b;
} cap_enc_layer;
typedef struct cap_enc_chain_weights cap_enc_chain_weights;
typedef struct cap_enc_chain_desc {
    int d_model, d_ff, d_feature, n_layers;
    int max_rows;                  /* max_batch * n_tokens */
    const cap_enc_layer* layers;
    const void* w_vis;             /* bf16 [d_model, d_feature]  vision_embedding.proj */
    const float *b_vis, *ln0_g, *ln0_b;   /* its bias; encoder.layer_norm */
    const float* pos;              /* fp32 [n_tokens, d_model] position table (row % n_tokens) */
    int n_kv;                      /* cross K|V projections computed while their source level is resident */
    const void* w_kv[CAP_ENC_MAX_KV];     /* bf16 [2*d_model, d_model] (fc_k | fc_v of a decoder layer's enc_attn) */
    const float* b_kv[CAP_ENC_MAX_KV];
    int kv_level[CAP_ENC_MAX_KV];         /* encoder layer whose output is projected */
    void* kv_dst[CAP_ENC_MAX_KV];         /* bf16 [rows, 2*d_model] */
    const void* feats;             /* bf16 [max_rows, d_feature] */
    void* qkv_out;                 /* bf16 [max_rows, 3*d_model] */
    const void* att_in;            /* bf16 [max_rows, d_model] */
    void* levels_out;              /* bf16 [n_layers][level_stride]: every layer's output, row-major */
    size_t level_stride;           /* elements */
    const uint8_t* row_mask;       /* uint8 [max_rows]: 1 = padded visual token */
    const cap_enc_chain_weights* stacked;   /* NULL: the handle builds and owns its stacked weight copies */
} cap_enc_chain_desc;
int cap_enc_chain_weights_create(const cap_enc_chain_desc* desc, cap_enc_chain_weights** out);
int cap_enc_chain_weights_destroy(cap_enc_chain_weights* w);
typedef struct cap_enc_chains cap_enc_chains;
int cap_enc_chains_create(const cap_enc_chain_desc* desc, cap_enc_chains** out);
int cap_enc_chains_destroy(cap_enc_chains* f);
int cap_enc_chain(cap_enc_chains* f, int stage, int rows, int n_tokens, cap_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole-path engine: encoder_forward once + max_len decode steps (models/base_transformer.py:
 * 30-53), one handle per GPU / rank.
 * ------------------------------------------------------------------------------------------ */
enum cap_encoder_kind { CAP_ENC_PLAIN = 0, CAP_ENC_MULTILEVEL = 1, CAP_ENC_GEOMETRIC = 2 };
enum cap_attention_kind { CAP_ATT_SDPA = 0, CAP_ATT_GEOMETRY = 1, CAP_ATT_MEMORY = 2 };
enum cap_decoder_kind { CAP_DEC_PLAIN = 0, CAP_DEC_MESHED = 1 };

typedef struct cap_model_desc {
    int d_model, heads, d_k, d_v, d_ff, d_feature;
    int enc_layers, dec_layers;
    int encoder_kind;      /* cap_encoder_kind */
    int enc_attention;     /* cap_attention_kind of the encoder self-attention */
    int n_memory;          /* memory slots when enc_attention == CAP_ATT_MEMORY */
    int trig_geometry;     /* TRIGNOMETRIC_EMBEDDING of GeometricEncoder */
    int decoder_kind;      /* cap_decoder_kind */
    int n_enc_levels;      /* encoder levels the meshed decoder attends to (1 for plain) */
    int aoa_enc, aoa_dec_self, aoa_dec_cross; /* USE_AOA flags */
    int vocab, max_len, pad_idx, bos_idx, eos_idx;
} cap_model_desc;

typedef struct cap_engine cap_engine;

int cap_engine_create(const cap_model_desc* desc, cap_engine** out);
int cap_engine_destroy(cap_engine* e);
/* A further engine over the SAME device weights as `parent` (which must be finalized): no upload, no copy -- only
 * its own workspaces, caches, beam state and CUDA graph after cap_engine_reserve.  Engines that caption independent
 * batches concurrently on several streams are created this way: one 48 MB weight set stays L2-resident instead of one
 * per engine.  The weights live until the last engine referencing them is destroyed. */
int cap_engine_create_shared(cap_engine* parent, cap_engine** out);
/* Upload one state_dict entry (fp32, HOST memory) under its reference name, e.g.
 * "encoder.layers.0.mhatt.attention.fc_q.weight".  Unknown names are an error. */
int cap_engine_load_weight(cap_engine* e, const char* name, const float* data_host,
                           const int64_t* shape, int ndim);
/* Check every required tensor arrived, build fused bf16 weights (q|k|v stacks, scaled memory
 * slots, position tables).  Must precede reserve/encode. */
int cap_engine_finalize(cap_engine* e);
/* Allocate workspaces, KV caches and the beam state for up to max_batch images of n_tokens
 * visual tokens decoded with `beam` beams. */
int cap_engine_reserve(cap_engine* e, int max_batch, int n_tokens, int beam);
/* encoder_forward + cross K/V projection; feats (B, n, d_feature) fp32 or bf16, boxes (B,n,4) fp32
 * (GeometricEncoder only, else NULL). */
int cap_engine_encode(cap_engine* e, const void* feats, int feat_dtype, const float* boxes, int B,
                      int n, cap_stream_t stream);
/* Decoder stack for step t on the beam state's current tokens -> logits (R, ld) fp32. */
int cap_engine_decode_logits(cap_engine* e, int t, cap_stream_t stream);
/* Production step: decoder stack + vocabulary GEMM with chunk statistics + beam update. */
int cap_engine_decode_step(cap_engine* e, int t, cap_stream_t stream);
/* cap_beam_step on the engine's own logits. */
int cap_engine_beam_advance(cap_engine* e, int t, cap_stream_t stream);
int cap_engine_begin_decode(cap_engine* e, cap_stream_t stream);
/* All max_len steps + finalize; ids int64 (B,out_size,T), logp fp32.  use_graph != 0 replays a
 * captured CUDA graph of one decode step. */
int cap_engine_beam_search(cap_engine* e, int out_size, int64_t* ids, float* logp, int use_graph,
                           cap_stream_t stream);
/* End-to-end with HOST buffers (pinned for full speed): H2D features, encode, beam search, D2H
 * ids/log-probs, stream sync.  The reference-facing call: trainers/vi_trainer.py:244 does
 * `items.to(device)` + `model.beam_search(...)` + `.tolist()`. */
int cap_engine_caption_host(cap_engine* e, const void* feats_host, int feat_dtype,
                            const float* boxes_host, int B, int n, int out_size, int64_t* ids_host,
                            float* logp_host, int use_graph, cap_stream_t stream);
/* Same, without the final stream synchronisation: several engines on several streams can then keep
 * H2D, compute and D2H of different batches in flight at once; the caller synchronises the stream
 * before reading ids_host / logp_host or reusing feats_host. */
int cap_engine_caption_host_async(cap_engine* e, const void* feats_host, int feat_dtype,
                                  const float* boxes_host, int B, int n, int out_size,
                                  int64_t* ids_host, float* logp_host, int use_graph,
                                  cap_stream_t stream);
/* Same path with DEVICE buffers on both sides (features already in HBM, ids int64 (B,out_size,T) and log-probs
 * fp32 left in HBM); asynchronous like the call above. */
int cap_engine_caption_device_async(cap_engine* e, const void* feats_dev, int feat_dtype,
                                    const float* boxes_dev, int B, int n, int out_size, int64_t* ids_dev,
                                    float* logp_dev, int use_graph, cap_stream_t stream);
/* Measurement hook: the 1 + 2*layers GEMM-chain launches of decode step t back to back, without the attention and
 * beam kernels between them (bench.py times them for the roofline).  CAP_ERR_STATE unless the engine runs chains. */
int cap_engine_debug_chains(cap_engine* e, int t, cap_stream_t stream);
/* Debug / parity views: */
const void* cap_engine_encoder_output(cap_engine* e);   /* bf16 [levels][B*n][d_model] */
const uint8_t* cap_engine_encoder_mask(cap_engine* e);  /* uint8 [B*n]                 */
const float* cap_engine_logits(cap_engine* e, int* ld); /* fp32 [R][ld]                */
cap_beam* cap_engine_beam(cap_engine* e);
/* Fault records: every bounded device-side wait (mbarrier waits of the tcgen05 / TMA kernels) that times out records
 * its source line in pinned host memory before it traps, so the cause of an "unspecified launch failure" can be read
 * AFTER the CUDA context has died.  Copies up to max_records source lines and returns the number of distinct waits
 * that timed out (0: none ever did). */
int cap_fault_records(unsigned int* out, int max_records);
/* Flight recorder (debug, OPENVIIC_FLIGHT=1 in the environment before the first launch): per kernel kind, how many
 * CTAs entered and how many left, counted in pinned host memory -- after a device hang the kinds with entered > left
 * are the ones that are stuck.  Copies [kind][entered, left] for up to max_kinds kinds, returns the number of kinds
 * (0: recorder off).  Kinds: csrc/cap_common.cuh FlightKind. */
int cap_flight_records(unsigned int* out, int max_kinds);
/* Debug: when non-NULL, every later cap_linear writes 8 %globaltimer stamps (ns) per CTA into
 * device_buffer[cta*8 + k]: 0 entry, 1 prologue done, 2 first TMA issued, 3 first stage landed,
 * 4 last MMA committed, 5 accumulator visible to the epilogue, 6 stores issued, 7 TMEM freed. */
int cap_debug_gemm_trace(unsigned long long* device_buffer);
/* ------------------------------------------------------------------------------------------------
 * Host glue either side of the path.  Plain host code: no GPU, no stream.
 * ------------------------------------------------------------------------------------------------ */
/* Collate: image i contributes n_rows[i] rows of D floats at rows[i]; the batch is (B, n_max, D), short images
 * zero-padded at the end -- InstanceList.__init__ / pad_values (reference utils/instance.py:32-55, 156-171)
 * writing directly into the buffer the H2D copy reads (pinned for full speed).  The bf16 variant rounds to
 * nearest even exactly as torch's .to(torch.bfloat16); `threads` host threads share the rows. */
int cap_host_collate_bf16(const float* const* rows, const int32_t* n_rows, int B, int n_max, int D,
                          uint16_t* out, int threads);
int cap_host_collate_f32(const float* const* rows, const int32_t* n_rows, int B, int n_max, int D,
                         float* out, int threads);
/* Vocabulary for ids -> text: word i is words[offsets[i] .. offsets[i+1]) (UTF-8, n_words + 1 offsets);
 * is_special[i] != 0 marks pad / bos / eos / unk (reference data_utils/vocab.py:41,66). */
typedef struct cap_vocab cap_vocab;
int cap_vocab_create(const char* words, const int64_t* offsets, int64_t n_words, const uint8_t* is_special,
                     int64_t eos_idx, cap_vocab** out);
int cap_vocab_destroy(cap_vocab* v);
/* Vocab.decode_caption(ids, join_words=True) (reference data_utils/vocab.py:104-122): per caption the
 * non-special words up to the first eos, joined by one space; captions are written back to back, each followed
 * by '\n'.  collapse_repeats != 0 also drops a word equal to the word before it -- the itertools.groupby pass
 * of the trainers (reference trainers/vi_trainer.py:251).  *out_bytes receives the bytes needed; if that exceeds
 * out_capacity nothing is written and CAP_ERR_INVALID is returned.  An id outside [0, n_words) is an error
 * (the reference raises IndexError). */
int cap_vocab_decode(const cap_vocab* v, const int64_t* ids, int64_t n_captions, int T, int collapse_repeats,
                     char* out, int64_t out_capacity, int64_t* out_bytes);
/* CIDEr-D: the self-critical reward and the evaluation loop's CIDEr -- reference
 * evaluation/cider/cider_scorer.py:9-167, evaluation/cider/cider.py:12-38, trainers/vi_trainer.py:137-145.  Captions are
 * int32 word-id sequences in ragged form: caption j is tokens[caption_offsets[j] .. caption_offsets[j+1]).
 *   cap_cider_set_corpus: image i owns captions [image_offsets[i], image_offsets[i+1]); counts in how many images each
 *     n-gram occurs (Cider(gts), cider.py:22-26).  ref_len = log(number of images) and log_table[c] = log(c) come
 *     from the caller so that they are the caller's (numpy's) logarithms; log_table may be NULL (libm's log).
 *   cap_cider_score: hypothesis i is scored against reference captions [ref_group_offsets[i], ref_group_offsets[i+1])
 *     (at least one); scores[i] is the reference's per-image CIDEr-D (x10).  Without a corpus the batch's reference
 *     groups are the documents and batch_ref_len / log_table apply (Cider(), cider.py:19-21).  Doubles throughout. */
typedef struct cap_cider cap_cider;
int cap_cider_create(int n, double sigma, cap_cider** out);
int cap_cider_destroy(cap_cider* c);
int cap_cider_set_corpus(cap_cider* c, const int32_t* tokens, const int64_t* caption_offsets,
                         const int64_t* image_offsets, int64_t n_images, double ref_len,
                         const double* log_table, int64_t log_table_len);
int64_t cap_cider_max_doc_freq(const cap_cider* c);
int cap_cider_score(const cap_cider* c, const int32_t* hyp_tokens, const int64_t* hyp_offsets, int64_t n_hyp,
                    const int32_t* ref_tokens, const int64_t* ref_caption_offsets,
                    const int64_t* ref_group_offsets, double batch_ref_len, const double* log_table,
                    int64_t log_table_len, double* scores, int threads);
/* Kernels launched by this library since load (all entry points); for bench.py's gpu_launches. */
int64_t cap_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * T1 -- the XE training step (trainers/vi_trainer.py:105-119; trainers/base_trainer.py:89-91, 114-117).
 * The forward pass uses cap_linear / cap_attention above plus the two *_fwd entry points here; the GEMMs of the backward
 * pass are cap_linear calls on transposed operands.  openviic_b200/training.py strings them together.
 * ------------------------------------------------------------------------------------------------------------------ */

/* pre = a (+ res) (kept for the backward pass); out = LayerNorm(pre) * gamma + beta (+ pos[row % pos_rows]), rows
 * flagged in zero_rows zeroed; fp32 and bf16 copies of out.  a, res, pre, out_f32: fp32 [rows][d] dense.
 * attentions.py:308-309, positionwise_feed_forward.py:26, encoders.py:20,36, decoders.py:26. */
int cap_train_layernorm_fwd(const float* a, const float* res, const float* gamma, const float* beta, float eps,
                            const float* pos, int pos_rows, const uint8_t* zero_rows, float* pre, float* out_f32,
                            void* out_bf16, int rows, int d, cap_stream_t stream);
/* Backward of the above: dout = dout_a (+ dout_b); dpre (fp32 and bf16) is the gradient of both summands of pre;
 * dgamma / dbeta [d] are ACCUMULATED into (atomicAdd). */
int cap_train_layernorm_bwd(const float* dout_a, const float* dout_b, const float* pre, const float* gamma, float eps,
                            const uint8_t* zero_rows, float* dpre_f32, void* dpre_bf16, float* dgamma, float* dbeta,
                            int rows, int d, cap_stream_t stream);
/* out[c][r] = in[r][c] (bf16), out rows padded with zeros up to ldo >= rows; colsum[c] += sum_r in[r][c] (fp32,
 * optional): the operands of dW = dY^T.X and the bias gradient of a Linear in one pass. */
int cap_transpose_bf16(const void* in, int ld, void* out, int ldo, float* colsum, int rows, int cols,
                       cap_stream_t stream);
/* Split-K product for few output tiles over a long contraction (the weight gradients dW = dY^T.X): `splits` CTAs per
 * output tile write fp32 partial products partials[z][M][ldy] over consecutive K ranges; cap_sum_partials adds them up
 * (count = M * ldy elements per partial).  x [M][K] and w [N][K] bf16 as in cap_linear; no bias, no activation. */
int cap_linear_splitk(const void* x, int ldx, const void* w, float* partials, int ldy, int M, int N, int K, int splits,
                      cap_stream_t stream);
int cap_sum_partials(const float* partials, int splits, int64_t count, float* out, cap_stream_t stream);
/* dh *= (h > 0), bf16, in place (positionwise_feed_forward.py:24). */
int cap_train_relu_bwd(void* dh, const void* h, int64_t count, cap_stream_t stream);
/* dst += src, fp32. */
int cap_axpy_f32(float* dst, const float* src, int64_t count, cap_stream_t stream);
/* Backward of cap_attention for the plain scaled dot-product attention (attentions.py:51-55): `args` as in the forward
 * call (geometry / memory / sentinel must be NULL; nq, nk <= 128); d_out has out's layout; dq / dk / dv (bf16) have
 * q's / k's / v's layout and strides. */
int cap_attention_backward(const cap_attention_args* args, const void* d_out, void* dq, void* dk, void* dv,
                           cap_stream_t stream);
/* Teacher-forcing token embedding: out[row] = emb[tokens[row]] + pos[tokens[row] == pad ? 0 : row % T + 1]
 * (decoders.py:105-112); emb, pos fp32; fp32 and bf16 outputs. */
int cap_train_embed_fwd(const int64_t* tokens, const float* emb, const float* pos, int T, int pad_idx, float* out_f32,
                        void* out_bf16, int rows, int d, cap_stream_t stream);
/* d_emb[tokens[row]] += g_a[row] (+ g_b[row]) for tokens != pad (nn.Embedding(padding_idx), text_embeddings.py:15). */
int cap_train_embed_bwd(const int64_t* tokens, const float* g_a, const float* g_b, int pad_idx, float* d_emb, int rows,
                        int d, cap_stream_t stream);
/* NLLLoss(ignore_index) over log_softmax(logits), mean over the counted targets (base_trainer.py:91, vi_trainer.py:110):
 * stats[0] = number of targets != ignore_index, stats[1] = sum of their negative log-likelihoods (loss = stats[1] /
 * stats[0]); dlogits bf16 [rows][ldd] = (softmax - onehot) / stats[0], zero for ignored rows and columns >= V.
 * row_weight (fp32 [rows], optional): stats[1] = sum w[row] * nll (the loss itself), dlogits = w[row] * (softmax -
 * onehot) -- the self-critical loss of vi_trainer.py:146-148 with w = advantage / (T * B * beam). */
int cap_train_xent(const float* logits, int ld, const int64_t* targets, int ignore_index, const float* row_weight,
                   float* stats, void* dlogits, int ldd, int rows, int V, cap_stream_t stream);
/* Dropout with a counter-based mask, in place: x[i] = hash(i, seed, site) >= threshold ? x[i] * scale : 0 (threshold =
 * floor(p * 2^32), scale = 1 / (1 - p)); the backward pass calls it on the gradient with the same (seed, site).  The
 * hash is restated in oracle/caption_oracle.py (dropout_keep).  nn.Dropout at vision_embeddings.py:18,
 * attentions.py:308, positionwise_feed_forward.py:24-25. */
int cap_train_dropout(void* x, int dtype, int64_t count, unsigned int threshold, float scale, unsigned int seed,
                      unsigned int site, cap_stream_t stream);
/* torch.optim.Adam (no weight decay) on flat fp32 buffers, step >= 1, and the bf16 copy of the new parameters. */
int cap_train_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                   int64_t count, float lr, float beta1, float beta2, float eps, int step, cap_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENVIIC_CAP_H */
