/* sonar_b200 -- C ABI of the B200-native SONAR text-embedding hot path.
 *
 * Plain C, no torch / CUDA types in the signatures: device buffers are `void*` /
 * typed raw pointers, the stream is an opaque `void*` (a `cudaStream_t`).
 * Every entry point returns 0 on success or a negative code and never throws;
 * `sb_last_error()` gives the message (thread-local).  No hidden device
 * allocations happen inside `sb_encoder_forward`; the caller owns all buffers.
 *
 * The reference (facebookresearch/SONAR) has NO native interface for this path --
 * it reaches fairseq2 Python modules.  Each entry point therefore cites the Python
 * interface it replaces:
 *
 *   sb_encoder_create / sb_encoder_forward
 *       <- SonarTextTransformerEncoderModel.forward(SequenceBatch) -> SonarEncoderOutput
 *          (sonar/models/sonar_text/model.py:130-143; built by
 *           SonarTextEncoderFactory.create_model, sonar/models/sonar_text/factory.py:72-120),
 *          called from TextToEmbeddingModelPipeline.predict via `.map(self.model)`
 *          (sonar/inference_pipelines/text.py:231-247).
 *   seq_lens / padded ids layout
 *       <- Collater(pad_idx) + extract_sequence_batch (text.py:241-242,
 *          sonar/inference_pipelines/utils.py:18-21): ids int64 [B,S] right-padded,
 *          PaddingMask(seq_lens).
 *   sb_pool
 *       <- SonarTextTransformerEncoderModel.static_pooling (model.py:86-128).
 *   weight layout
 *       <- fairseq2 state-dict names mapped in sonar/models/sonar_text/handler.py:71-92
 *          (nn.Linear [out,in] row-major).
 */
#ifndef SONAR_B200_H_
#define SONAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID (-1)  /* bad argument / unsupported shape */
#define SB_ERR_CUDA (-2)     /* CUDA runtime error */
#define SB_ERR_DRIVER (-3)   /* driver entry point / tensor-map failure */
#define SB_ERR_INPUT (-4)    /* device-side input check failed (e.g. token id out of range) */

/* Reference `Pooling` enum values (sonar/models/sonar_text/model.py:23-27). */
#define SB_POOL_MAX 1
#define SB_POOL_MEAN 2
#define SB_POOL_LAST 3

/* GEMM epilogues */
#define SB_EPI_BIAS 0
#define SB_EPI_BIAS_RELU 1
#define SB_EPI_BIAS_RESIDUAL 2
#define SB_EPI_BIAS_SILU 5

typedef struct SbEncoder SbEncoder;

/* Mirrors the fields of SonarTextEncoderConfig that reach the math
 * (sonar/models/sonar_text/config.py:14-84). */
typedef struct SbEncoderConfig {
  int32_t model_dim;     /* 1024 (multiple of 256, <= 1024; head_dim must be 64) */
  int32_t num_layers;    /* 24 */
  int32_t num_heads;     /* 16 */
  int32_t ffn_inner_dim; /* 8192 (multiple of 256) */
  int64_t vocab_size;    /* 256206 */
  int32_t pos_rows;      /* rows of the sinusoidal table = max_seq_len + pad_idx + 1 = 514 */
  int32_t pooling;       /* SB_POOL_* ; `basic` = SB_POOL_MEAN */
  float ln_eps;          /* 1e-5 */
  float embed_scale;     /* sqrt(model_dim) unless no_scale_embedding */
  int32_t cta_group;     /* 0/2 = paired-CTA tcgen05 tiles (default), 1 = single-CTA */
  int32_t num_sms;       /* 0 = query the device */
  int32_t ln_fold;       /* 1 = fold every encoder-layer LayerNorm into the GEMMs around it (no LayerNorm kernel runs:
                          *     the residual GEMMs emit per-row statistics + a bf16 copy of the stream, the QKV / FFN1 GEMMs
                          *     apply (mean, rstd) in their epilogue on weights pre-multiplied by gamma; sb_encoder_create
                          *     prepares those weights in device memory it owns -- the caller's weights are not modified);
                          * 2 = fold only the attention-block LayerNorm (FFN2 emits, QKV applies); the FFN-block LayerNorm
                          *     stays a kernel (the out-projection is HBM-bound, its epilogue has no slack for the extra work);
                          * 0 = separate LayerNorm kernels (the round-1 schedule) */
  int32_t epi_groups;    /* 0/2 = two epilogue warpgroups per GEMM CTA (default); 1 = one (round-1 kernel, kept for A/B runs;
                          *     needs ln_fold = 0) */
} SbEncoderConfig;

/* All pointers are DEVICE pointers and stay owned by the caller (must outlive the handle).
 * Matrices: bf16, row-major [out_features, in_features].  Vectors: fp32. */
typedef struct SbLayerWeights {
  const void* wqkv;   /* bf16 [3*D, D] = rows of q_proj | k_proj | v_proj */
  const float* bqkv;  /* [3*D] */
  const void* wo;     /* bf16 [D, D]   self_attn.output_proj */
  const float* bo;    /* [D] */
  const void* w1;     /* bf16 [F, D]   ffn.inner_proj */
  const float* b1;    /* [F] */
  const void* w2;     /* bf16 [D, F]   ffn.output_proj */
  const float* b2;    /* [D] */
  const float* ln1_g; /* self_attn_layer_norm */
  const float* ln1_b;
  const float* ln2_g; /* ffn_layer_norm */
  const float* ln2_b;
} SbLayerWeights;

typedef struct SbEncoderWeights {
  const void* embed;           /* bf16 [vocab, D]  encoder_frontend.embed.weight */
  const float* pos_table;      /* fp32 [pos_rows, D]; row t = sinusoid of position t + pad_idx + 1 */
  const float* final_ln_g;     /* layer_norm.weight */
  const float* final_ln_b;     /* layer_norm.bias */
  const SbLayerWeights* layers; /* HOST array of num_layers entries (copied at create) */
} SbEncoderWeights;

const char* sb_last_error(void);
int sb_version(void);

/* Allocates the handle's own device memory (a 256-byte input-check flag; with cfg->ln_fold the folded copies of the QKV and
 * FFN inner-projection weights, ~22 MB per layer) and synchronises the device once.  sb_encoder_forward never allocates. */
int sb_encoder_create(const SbEncoderConfig* cfg, const SbEncoderWeights* w, SbEncoder** out);
void sb_encoder_destroy(SbEncoder* enc);

/* Bytes of device workspace needed for a batch of <= max_batch sequences holding
 * <= max_tokens real (unpadded) tokens in total. */
int sb_encoder_workspace_bytes(const SbEncoder* enc, int32_t max_batch, int64_t max_tokens, size_t* bytes);

/* One pass of the hot path.
 *   ids            DEVICE int64 [batch, ids_row_stride >= seq_len], right-padded (any pad value)
 *   seq_lens_host  HOST int32 [batch] true lengths (1..seq_len); NULL = every row is full (no padding_mask)
 *   out            DEVICE fp32 [batch, model_dim] sentence embeddings
 *   encoded        DEVICE fp32 [batch, seq_len, model_dim] or NULL (final-LayerNormed states, padded rows zeroed)
 * Asynchronous on `stream` (does not synchronise). */
int sb_encoder_forward(SbEncoder* enc, const int64_t* ids, int64_t ids_row_stride, const int32_t* seq_lens_host,
                       int32_t batch, int32_t seq_len, float* out, float* encoded, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Same, but `ids_host` / `out_host` are HOST buffers (pinned for full speed); performs the
 * H2D copy of the ids, the forward and the D2H copy of the embeddings on `stream`, then
 * synchronises the stream.  `ids_staging` is DEVICE int64 [batch*seq_len], `out_staging`
 * DEVICE fp32 [batch*model_dim] (caller-owned). */
int sb_encoder_forward_host(SbEncoder* enc, const int64_t* ids_host, const int32_t* seq_lens_host, int32_t batch,
                            int32_t seq_len, float* out_host, int64_t* ids_staging, float* out_staging,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hook: every following forward records the two caller-owned CUDA events (cudaEvent_t, timing enabled)
 * around the FFN inner-projection GEMM of the middle layer -- the dominant kernel -- on the forward's stream, so its
 * duration can be read INSIDE a real step.  Pass NULL, NULL to switch it off. */
int sb_encoder_profile_ffn1(SbEncoder* enc, void* start_event, void* stop_event);

/* Checks the sticky device-side input flag (token id out of range) of the last forwards;
 * synchronises `stream`.  Returns SB_OK or SB_ERR_INPUT. */
int sb_encoder_check_inputs(SbEncoder* enc, void* workspace, void* stream);

/* ---- individual kernels (used by the parity tests and the micro-benchmarks) ---- */

/* LayerNorm folding (the schedule behind SbEncoderConfig.ln_fold; replaces F.layer_norm + F.linear pairs of the pre-LN
 * encoder layer, sonar/models/sonar_text/factory.py:122-153):
 *   LN(x; gamma, beta) . W^T + b  =  rstd * (x . Wf^T - mean * colsum) + bias_f
 * sb_fold_layernorm prepares Wf = bf16(W diag(gamma)) [N,K], colsum[n] = sum_k Wf[n,k], bias_f = bias + W beta;
 * sb_gemm_residual_stats computes x += A . W^T + bias (fp32, in place) and emits h_out = bf16(x) plus stats_out
 *   [M, N/128, 2] = (mean, M2) of N/128 disjoint 128-column subsets of the new rows (one per epilogue warpgroup and tile);
 * sb_gemm_ln_consumer computes C (bf16) = [relu](rstd * (A . Wf^T - mean * colsum) + bias_f) with A = the bf16 copy and
 *   (mean, rstd) merged from `stats` [M, K/128, 2].  K = a multiple of 128, <= 1024. */
int sb_fold_layernorm(const void* W, const float* bias, const float* gamma, const float* beta, int32_t N, int32_t K,
                      void* Wf, float* colsum, float* bias_f, void* stream);
int sb_gemm_ln_consumer(const void* A, int64_t lda, const void* Wf, int64_t ldw, void* C, int64_t ldc, const float* bias_f,
                        const float* colsum, const float* stats, float eps, int32_t M, int32_t N, int32_t K, int32_t relu,
                        void* stream);
int sb_gemm_residual_stats(const void* A, int64_t lda, const void* W, int64_t ldw, float* x, int64_t ldx, const float* bias,
                           void* h_out, int64_t ldh, float* stats_out, int32_t M, int32_t N, int32_t K, void* stream);

/* x += A . W^T + bias (fp32, in place) as the decoder step issues it (the residual additions of
 * fairseq2's StandardTransformerDecoderLayer, reached from sonar/models/sonar_text/factory.py:263-301): when the
 * [M/256, N/256] tile pairs leave SM pairs idle (2 560 hypothesis rows x 1 024 columns = 40 tiles on 74 pairs) the K
 * dimension is split into up to 4 slices run by different SM pairs, which add into x ONE AFTER THE OTHER (hand-over through
 * `counters`, >= 4 * tiles zero-initialised device ints that the call leaves zero): x + p0, + p1, + p2 -- the same bits on
 * every run.  With enough tiles it is the plain accumulate epilogue. */
int sb_gemm_residual_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* x, int64_t ldx, const float* bias,
                            int32_t M, int32_t N, int32_t K, int32_t* counters, int64_t n_counters, void* stream);

/* C[M,N] = epi(A[M,K] * W[N,K]^T + bias[N]) ; A, W bf16 row-major; C bf16 (out_fp32=0) or fp32;
 * residual (SB_EPI_BIAS_RESIDUAL) has C's dtype and may alias C.  N % 256 == 0, K % 64 == 0.
 * cta_group: 2 = paired-CTA tcgen05 tiles, 1 = single-CTA tiles, 0 = automatic (paired tiles, except that M <= 64 with
 * K % 256 == 0 takes the weight-streaming mma.sync path the decoder step uses; then only N % 8 == 0 is required). */
int sb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t out_fp32,
                 const float* bias, const void* residual, int64_t ldr, int32_t M, int32_t N, int32_t K, int32_t epi,
                 int32_t cta_group, void* stream);

/* y[T,D] (bf16) = LayerNorm(x[T,D] fp32) */
int sb_layernorm(const float* x, const float* gamma, const float* beta, float eps, void* y, int64_t T, int32_t D,
                 void* stream);

/* packed self-attention: qkv bf16 [total_tokens, 3*64*H], cu_seqlens DEVICE int32 [B+1], out bf16 [total_tokens, 64*H].
 * impl 0 = auto (= 2), 1 = mma.sync flash kernel (kept for tests / A-B timing), 2 = tcgen05 (any length; 128-key
 * tiles with online softmax beyond 128 tokens). */
int sb_attention(const void* qkv, const int32_t* cu_seqlens, int32_t B, int32_t max_len, int32_t H,
                 int64_t total_tokens, int32_t impl, void* out, void* stream);

/* x[cu[b]+t,:] = embed[ids[b,t],:]*scale + pos[t,:] ; err_flag DEVICE int32 (set to 1 on a bad id) */
int sb_embed(const int64_t* ids, int64_t ids_row_stride, const int32_t* cu_seqlens, int32_t B, int32_t S,
             const void* embed, int64_t vocab, const float* pos_table, int32_t pos_rows, int32_t D, float scale,
             float* x, int32_t* err_flag, void* stream);

/* (optional LayerNorm +) pooling of packed rows x fp32 [T,D] -> out fp32 [B,D] */
int sb_pool(const float* x, const int32_t* cu_seqlens, int32_t B, int32_t D, const float* gamma, const float* beta,
            float eps, int32_t apply_ln, int32_t pool_mode, float* out, float* encoded_padded, int32_t S_padded,
            void* stream);

/* ---- embedding -> text decoder, one incremental step at a time (BASELINE.json config 4) ----
 * Replaces ConditionalTransformerDecoderModel.decode + project (sonar/nn/conditional_decoder_model.py:60-94,
 * built by SonarTextDecoderFactory, sonar/models/sonar_text/factory.py:229-315) as driven by fairseq2's
 * BeamSearchSeq2SeqGenerator inside EmbeddingToTextModelPipeline.predict (sonar/inference_pipelines/text.py:305-346).
 * State-dict names: sonar/models/sonar_text/handler.py:136-158. */
typedef struct SbDecoder SbDecoder;

typedef struct SbDecoderConfig {
  int32_t model_dim;     /* 1024 */
  int32_t num_layers;    /* 24 */
  int32_t num_heads;     /* 16 */
  int32_t ffn_inner_dim; /* 8192 */
  int32_t input_dim;     /* dimensionality of the sentence embedding; must equal model_dim */
  int64_t vocab_size;    /* 256206 */
  int32_t pos_rows;      /* rows of the sinusoidal table (max_seq_len + pad_idx + 1) */
  int32_t eos_idx;       /* 3 */
  float ln_eps;          /* 1e-5 */
  float embed_scale;     /* sqrt(model_dim) */
} SbDecoderConfig;

/* DEVICE pointers, caller-owned; matrices bf16 [out,in], vectors fp32.  The encoder-decoder attention attends
 * over ONE key (the sentence embedding), so only its v_proj / output_proj reach the result. */
typedef struct SbDecoderLayerWeights {
  const void* wqkv;      /* bf16 [3D, D] self_attn q|k|v */
  const float* bqkv;
  const void* wo;        /* self_attn.output_proj */
  const float* bo;
  const void* cross_wv;  /* encoder_decoder_attn.v_proj [D, input_dim] */
  const float* cross_bv;
  const void* cross_wo;  /* encoder_decoder_attn.output_proj */
  const float* cross_bo;
  const void* w1;        /* ffn.inner_proj */
  const float* b1;
  const void* w2;        /* ffn.output_proj */
  const float* b2;
  const float* ln1_g;    /* self_attn_layer_norm */
  const float* ln1_b;
  const float* ln3_g;    /* ffn_layer_norm */
  const float* ln3_b;
} SbDecoderLayerWeights;

typedef struct SbDecoderWeights {
  const void* embed;       /* bf16 [vocab, D] decoder_frontend.embed.weight == final_proj.weight (tied) */
  const float* pos_table;  /* fp32 [pos_rows, D] */
  const float* final_ln_g; /* decoder.layer_norm */
  const float* final_ln_b;
  const SbDecoderLayerWeights* layers; /* HOST array */
} SbDecoderWeights;

int sb_decoder_create(const SbDecoderConfig* cfg, const SbDecoderWeights* w, SbDecoder** out);
void sb_decoder_destroy(SbDecoder* dec);
/* workspace for num_sentences x beam hypotheses of at most max_len positions (holds the KV caches) */
int sb_decoder_workspace_bytes(const SbDecoder* dec, int32_t num_sentences, int32_t beam, int32_t max_len, size_t* bytes);
/* start a batch: embeddings DEVICE fp32 [num_sentences, model_dim]; precomputes the per-layer cross-attention constants */
int sb_decoder_begin(SbDecoder* dec, const float* embeddings, int32_t num_sentences, int32_t beam, int32_t max_len,
                     void* workspace, size_t workspace_bytes, void* stream);
/* one decoding step at position t for all R = num_sentences*beam rows (row = sentence*beam + beam_slot):
 *   tokens  DEVICE int64 [R]           input token of every hypothesis at position t
 *   table   DEVICE int32 [R, max_len]  table[r, t'] = physical cache row holding position t' < t of hypothesis r
 *   out_lprob / out_tok DEVICE [R, 16] the 16 most probable next tokens (fp32 log-softmax over the whole vocabulary),
 *                                      ordered by (log-prob desc, token asc)
 *   out_eos_lprob DEVICE fp32 [R]      log P(eos)
 *   probe_tokens  DEVICE int64 [R] or NULL; with it out_probe_lprob DEVICE fp32 [R] = log P(probe_tokens[r]) -- the
 *                                      prompt-token scores fairseq2's generator accumulates while prefilling [fs2 _prefill] */
int sb_decoder_step(SbDecoder* dec, const int64_t* tokens, const int32_t* table, int32_t t, int32_t num_sentences,
                    int32_t beam, int32_t max_len, float* out_lprob, int32_t* out_tok, float* out_eos_lprob,
                    const int64_t* probe_tokens, float* out_probe_lprob, void* workspace, size_t workspace_bytes,
                    void* stream);
int sb_decoder_check_inputs(SbDecoder* dec, void* workspace, void* stream);

/* ---- speech feature frontend (BASELINE.json config 3, rows a9/a10) ----
 * Replaces fairseq2n WaveformToFbankConverter(num_mel_bins=80, waveform_scale=2**15, channel_last=True,
 * standardize=True) + Collater(pad_value=0, pad_to_multiple=2) (sonar/inference_pipelines/speech.py:120-127,139,
 * 283-290,444).  16 kHz mono input. */
size_t sb_fbank_tables_bytes(void);
/* fills a HOST buffer of sb_fbank_tables_bytes() (window, FFT twiddles, mel filters); upload it to the device once */
int sb_fbank_build_tables(void* host_buf);
/* waves DEVICE fp32 packed samples in [-1,1]; wave_offsets DEVICE int64 [B+1]; frame_offsets DEVICE int32 [B+1]
 * (cumulative frame counts, frames_b = 1 + (samples_b - 400) / 160); tables DEVICE (see above);
 * raw_out DEVICE fp32 [total_frames, 80] scratch; out DEVICE fp32 [B, padded_frames, 80] standardised, zero padded */
int sb_fbank(const float* waves, const int64_t* wave_offsets, const int32_t* frame_offsets, int32_t B,
             int32_t total_frames, const void* tables, float* raw_out, float* out, int32_t padded_frames, void* stream);

/* ---- speech encoder: w2v-BERT Conformer stack + attention pooler (BASELINE.json config 3, rows a11/a12) ----
 * Replaces SonarSpeechEncoderModel.forward (sonar/models/sonar_speech/model.py:59-77; factory.py:53-152;
 * sonar/nn/encoder_pooler.py:70-83); parameter names per sonar/models/sonar_speech/handler.py:63-100. */
typedef struct SbSpeechEncoder SbSpeechEncoder;

typedef struct SbSpeechConfig {
  int32_t model_dim;            /* 1024 */
  int32_t num_layers;           /* 24 Conformer blocks */
  int32_t num_heads;            /* 16 */
  int32_t ffn_inner_dim;        /* 4096 */
  int32_t conv_kernel;          /* 31 */
  int32_t pooler_layers;        /* 3 (english) / 6 (non_english) */
  int32_t pooler_ffn_inner_dim; /* 4096 */
  float ln_eps;                 /* 1e-5 */
  int32_t attn_impl;            /* relative-position attention: 0 = tcgen05 kernel, 1 = mma.sync kernel (the Python wrapper's
                                 * default: measured faster end to end) */
} SbSpeechConfig;

/* every field is a DEVICE pointer (matrices bf16 [out,in]; vectors fp32) */
typedef struct SbConformerLayerWeights {
  const float* ffn1_ln_g; const float* ffn1_ln_b;
  const void* ffn1_w1; const float* ffn1_b1; const void* ffn1_w2 /* x0.5 */; const float* ffn1_b2 /* x0.5 */;
  const float* attn_ln_g; const float* attn_ln_b;
  const void* wqkv; const float* bqkv; const void* wo; const float* bo;
  const void* wr;            /* self_attn.sdpa.r_proj.weight */
  const float* u_bias;       /* [H*64] */
  const float* v_bias;
  const float* conv_ln_g; const float* conv_ln_b;
  const void* pw1;           /* conv.pointwise_conv1 [2D, D] */
  const float* dw;           /* conv.depthwise_conv  fp32 [D, 31] */
  const float* bn_scale;     /* gamma / sqrt(running_var + eps) */
  const float* bn_shift;     /* beta - running_mean * bn_scale */
  const void* pw2;           /* conv.pointwise_conv2 [D, D] */
  const float* ffn2_ln_g; const float* ffn2_ln_b;
  const void* ffn2_w1; const float* ffn2_b1; const void* ffn2_w2 /* x0.5 */; const float* ffn2_b2 /* x0.5 */;
  const float* ln_g; const float* ln_b;   /* the block's final layer_norm */
} SbConformerLayerWeights;

typedef struct SbPoolerLayerWeights {
  const void* sa_wv; const float* sa_bv; const void* sa_wo; const float* sa_bo;
  const float* sa_ln_g; const float* sa_ln_b;
  const void* ca_wq; const float* ca_bq;
  const void* ca_wkv /* [2D, D] = k_proj | v_proj */; const float* ca_bkv;
  const void* ca_wo; const float* ca_bo;
  const float* ca_ln_g; const float* ca_ln_b;
  const void* w1; const float* b1; const void* w2; const float* b2;
  const float* ffn_ln_g; const float* ffn_ln_b;
} SbPoolerLayerWeights;

typedef struct SbSpeechWeights {
  const float* front_ln_g;  /* encoder_frontend.post_extract_layer_norm [160] */
  const float* front_ln_b;
  const void* front_w;      /* encoder_frontend.model_dim_proj, bf16 [D, 192] (columns 160..191 zero) */
  const float* front_b;
  const float* final_ln_g;  /* layer_norm (re-homed stack LayerNorm) */
  const float* final_ln_b;
  const float* pooler_q0;   /* fp32 [D] = embed[bos] * sqrt(D) + pos[0] */
  const void* proj_w;       /* encoder_pooler.projection_out.weight bf16 [D, D] */
  const float* zeros;       /* fp32 zeros, >= max(2D, relpos rows) entries (bias of the bias-free projections) */
  const SbConformerLayerWeights* layers; /* HOST arrays */
  const SbPoolerLayerWeights* pooler;
} SbSpeechWeights;

int sb_speech_encoder_create(const SbSpeechConfig* cfg, const SbSpeechWeights* w, SbSpeechEncoder** out);
void sb_speech_encoder_destroy(SbSpeechEncoder* enc);
int sb_speech_encoder_workspace_bytes(const SbSpeechEncoder* enc, int32_t batch, int64_t total_positions,
                                      int32_t max_positions, size_t* bytes);
/* fbank DEVICE fp32 [batch, padded_frames, 80] (sb_fbank output); cu_dev DEVICE int32 [batch+1] cumulative positions,
 * lens_host HOST int32 [batch] positions per utterance (= frames // 2); relpos_table DEVICE bf16 [relpos_rows, D] with
 * relpos_rows = roundup(2*max_len - 1, 256), row k = sinusoid of relative position (max_len - 1 - k), zero rows after
 * 2*max_len - 1; out DEVICE fp32 [batch, D]; encoded_packed DEVICE fp32 [total_positions, D] or NULL. */
int sb_speech_encoder_forward(SbSpeechEncoder* enc, const float* fbank, int32_t padded_frames, const int32_t* cu_dev,
                              const int32_t* lens_host, int32_t batch, const void* relpos_table, int32_t relpos_rows,
                              float* out, float* encoded_packed, void* workspace, size_t workspace_bytes, void* stream);

/* ---- xsim cosine k-NN / margin mining over sentence embeddings (BASELINE.json config 5) ----
 * Not a reference interface: the reference only ever does normalize + matmul
 * (tests/integration_tests/test_text_sonar.py:42,51); algorithm = public LASER xsim. */

int sb_xsim_workspace_bytes(int32_t n, int32_t m, int32_t d, size_t* bytes);

/* k nearest rows of y[m,d] (cosine) for every row of x[n,d]; x, y DEVICE fp32 row-major (raw, un-normalised).
 * out_val DEVICE fp64 [n,k] exact cosines, out_idx DEVICE int32 [n,k]; sorted by (cosine desc, index asc).
 * Candidates come from a bf16 tcgen05 GEMM with a fused running top-16, then are re-scored in fp64. */
int sb_xsim_knn(const float* x, const float* y, int32_t n, int32_t m, int32_t d, int32_t k, double* out_val,
                int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* Both k-NN directions from ONE pass over x^ . y^T ("row top-k and simultaneously per-column partial top-k"):
 * val_xy / idx_xy [n,k] as sb_xsim_knn; val_yx / idx_yx [m,k] = for every y row its k nearest x rows (exact fp64 cosines,
 * int32 row indices).  The reverse direction's candidates are the elements of the product above a per-column threshold:
 * the 16th best bf16 score of that y row against every 8th x row (a 1/8-size GEMM).  sb_xsim_knn keeps the 16 best bf16
 * scores of a row and re-scores them exactly, and a subset's 16th best never exceeds the 16th best over all rows, so the
 * result equals sb_xsim_knn(y, x) for any data; about 16 * 8 = 128 rows pass per column whatever the score distribution.
 * A y row that collects more than its 256 slots (heavy ties / duplicates) is marked idx_yx[j, :] = -2 and counted in
 * *overflow_flag (DEVICE int): the caller recomputes those rows with sb_xsim_knn(y[rows], x) (sonar_b200/xsim.py::knn_bidir
 * does). */
int sb_xsim_bidir_workspace_bytes(int32_t n, int32_t m, int32_t d, size_t* bytes);
int sb_xsim_knn_bidir(const float* x, const float* y, int32_t n, int32_t m, int32_t d, int32_t k, double* val_xy,
                      int32_t* idx_xy, double* val_yx, int32_t* idx_yx, int32_t* overflow_flag, void* workspace,
                      size_t workspace_bytes, void* stream);

/* pred[i] = forward candidate of row i with the best margin score.
 * margin_mode 0 = absolute, 1 = ratio, 2 = distance; val_yx = fp64 [m,k] k-NN cosines of y rows among x. */
int sb_xsim_margin_predict(const double* val_xy, const int32_t* idx_xy, const double* val_yx, int32_t n, int32_t m,
                           int32_t k, int32_t margin_mode, int32_t* pred, void* stream);

/* ---- beam search bookkeeping (one step, one launch) ----
 * The state transition of fairseq2's BeamSearchSeq2SeqGenerator [fs2] as the reference drives it
 * (sonar/inference_pipelines/text.py:315-333): from the decoder step's 16 best continuations per hypothesis
 * (lp / tok [N*beam,16], eos_lp [N*beam]) select the 2*beam best per sentence (score desc, then beam*vocab+token asc),
 * retire the EOS-terminated ones ranked inside the beam into fin_* until the sentence owns `beam` hypotheses (slot
 * CAP = 2*beam is scratch), and let the first `beam` others continue: seqs [N,beam,Tmax], the KV-cache ancestry table [N*beam,Tmax], tokens [N*beam], cum / alive
 * [N,beam] and done [N] are updated in place (alive / done: one byte per flag).  t = position of the step's input token,
 * g = number of tokens generated before this step; EOS is forbidden while g < eos_block (= min_gen_len - 1: fairseq2's
 * `step_nr < min_seq_len - 1`); score_div = (P+g)^len_penalty with P the prompt length (fairseq2 normalises by
 * seq_len - 1 counting prompt and EOS).  beam <= 7. */
int sb_beam_step(const float* lp, const int32_t* tok, const float* eos_lp, int64_t* seqs, int32_t* table,
                 int64_t* tokens, float* cum, uint8_t* alive, uint8_t* done, float* fin_score, int64_t* fin_seq,
                 int64_t* fin_len, int64_t* fin_count, int32_t N, int32_t beam, int32_t Tmax, int32_t t, int32_t g,
                 int32_t eos_block, int32_t max_gen, int64_t vocab, int32_t eos, int32_t unk, int32_t pad,
                 float unk_penalty, float score_div, int32_t normalize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SONAR_B200_H_ */
