// Internal (C++) interfaces shared by the sonar_b200 CUDA translation units.
// The public C ABI is include/sonar_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// EPI_BIAS_ACCUM (internal): C += A.W^T + bias with the add done by TMA reduce-add at L2; chosen
// automatically for EPI_BIAS_RESIDUAL when the residual aliases a fp32 C (the encoder's x += ... case).
// Bit-identical to EPI_BIAS_RESIDUAL: both compute fl32(x + fl32(acc + bias)).
// EPI_TOPK (internal): no C at all -- the epilogue keeps a running per-row top-k of the product (xsim mining).
// EPI_BIAS_SILU: x*sigmoid(x) (the Conformer's swish FFN activation)
// EPI_BIAS_RESIDUAL_STATS (internal): fp32 C = residual + A.W^T + bias, plus the bf16 copy and row statistics of LnFold
enum EpiMode { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_RESIDUAL = 2, EPI_BIAS_ACCUM = 3, EPI_TOPK = 4, EPI_BIAS_SILU = 5,
               EPI_BIAS_RESIDUAL_STATS = 6 };
constexpr int kTopkCandidates = 16;  // bf16-similarity candidates per row handed to the exact fp64 re-rank
enum PoolMode { POOL_MAX = 1, POOL_MEAN = 2, POOL_LAST = 3 };  // = reference `Pooling` enum values (model.py:23-27)

void set_last_error(const char* fmt, ...);

// cudaFuncSetAttribute is per device: returns true the first time `flags` (a per-call-site static array of 64 bools)
// is consulted for the current device.
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// LayerNorm FOLDED into the GEMMs on either side of it (the text encoder's default schedule; no LayerNorm kernel runs).
//   LN(x) . W^T + b  =  rstd * (x . W'^T  -  mean * c)  +  b'      with  W' = W diag(gamma),  c[n] = sum_k W'[n,k],
//                                                                        b' = b + W beta          (prepared once at create)
// so the GEMM that CONSUMES a LayerNorm runs on the un-normalised bf16 copy of the residual stream and applies the
// per-row (mean, rstd) in its epilogue (`stats_in`, `colsum`), and the GEMM that PRODUCES the residual stream
// (EPI_BIAS_RESIDUAL_STATS) emits, next to x, that bf16 copy (`h_out`) and per-row partial statistics (`stats_out`):
// one (mean, M2) pair per kLnPartCols columns of the row (each epilogue warpgroup's share of a 256-column tile), merged by
// the consumer with Chan's formula (no E[x^2]-mean^2 cancellation).  Saves the LayerNorm kernel's read of x and one of
// the two passes over h per LayerNorm.
constexpr int kLnPartCols = 128;
struct LnFold {
  const float* stats_in = nullptr;  // [M, chunks, 2] (mean, M2) of `chunks` disjoint kLnPartCols-column subsets of the input rows
  const float* colsum = nullptr;    // [N] c[n]
  int chunks = 0;                   // K / kLnPartCols of the LayerNorm the consumer folds (<= 8)
  float eps = 0.f;
  __nv_bfloat16* h_out = nullptr;   // producer: [M, N] bf16 copy of the new residual stream, leading dimension ldh
  long long ldh = 0;
  float* stats_out = nullptr;       // producer: [M, N / kLnPartCols, 2]
};

struct GemmArgs {
  const __nv_bfloat16* A;  // [M,K] row-major, ld = lda
  long long lda;
  const __nv_bfloat16* W;  // [N,K] row-major (nn.Linear layout), ld = ldw
  long long ldw;
  void* C;  // [M,N] bf16 or fp32
  long long ldc;
  int out_fp32;
  const float* bias;     // [N] fp32
  const void* residual;  // [M,N] same dtype as C (may alias C), ld = ldr
  long long ldr;
  int M, N, K;
  int epi;        // EpiMode
  int cta_group;  // 1 or 2
  int num_sms;    // 0 -> 148
  LnFold lf;      // LayerNorm folding (see above); default = off
  // Opt-in to the weight-streaming path for M <= 64 (gemm_skinny.cu).  It sums K in a different order than the tcgen05
  // tiles, so a caller that promises results independent of the batch size across the M = 64 boundary (the text
  // encoder: bitwise batch-composition invariance) leaves it off; the decoder step and the speech pooler turn it on.
  int allow_skinny = 0;
  // 2 (default) = two epilogue warpgroups / 5 mainloop stages; 1 = one epilogue warpgroup / 6 stages (A/B variant, only
  // for the text encoder's bias, bias+ReLU and accumulate epilogues with paired CTAs)
  int epi_groups = 2;
  // Ordered split-K of the accumulate epilogue (x += A.W^T + b with few tiles): zero-initialised device counters, one per
  // (tile, CTA of the pair, epilogue warpgroup); the kernel leaves them zero.  nullptr = never split.  Changes the
  // summation order (deterministically), so only callers that do not promise batch-size-independent bits pass it.
  int* splitk_flags = nullptr;
  long long splitk_flags_len = 0;
};

int gemm_bf16(const GemmArgs& g, cudaStream_t stream);

// M <= 64 rows: weight-streaming mma.sync path (gemm_skinny.cu); gemm_bf16 dispatches to it when eligible
bool gemm_skinny_eligible(const GemmArgs& g);
int gemm_skinny(const GemmArgs& g, cudaStream_t stream);

// Column filter of the top-k sweep (xsim: both k-NN directions from ONE pass over x . y^T): next to the per-row lists, every
// element above its column's threshold is appended to that column's candidate buffer -- entry = (bf16 product as fp32 bits,
// row index); `cnt[col]` counts the hits (beyond `cap` they are dropped and the column's threshold is raised to +inf so that
// degenerate inputs -- every product above its threshold -- cannot turn the sweep into a stream of atomics: the caller checks
// cnt > cap).
// `thr` must be readable up to the next multiple of 256 columns (pad with +inf); `thr8[g]` = min(thr[8g .. 8g+7]) lets the
// epilogue reject 8 columns of a row with one compare against the maximum it already has.
struct ColFilter {
  float* thr = nullptr;        // [N padded to 256]; nullptr = no column filter.  A column that fills up is closed (+inf)
  const float* thr8 = nullptr; // [N padded to 256, / 8]
  int* cnt = nullptr;          // [N]
  uint2* buf = nullptr;        // [N, cap]
  int cap = 0;
};

int gemm_topk_chunks(int M, int N, int cta_group, int num_sms);
// candidate lists per row that gemm_bf16_topk writes for a given n_chunks (two epilogue warpgroups per n-chunk)
constexpr int kTopkListsPerChunk = 2;
inline int gemm_topk_lists(int n_chunks) { return kTopkListsPerChunk * n_chunks; }
int gemm_bf16_topk(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, int M, int N, int K,
                   float* cand_val, int* cand_idx, float* lse_part, int n_chunks, int cta_group, int num_sms,
                   cudaStream_t stream, const ColFilter& cf = ColFilter());

int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                 int box_rows, int box_cols);

// x[cu[b]+t, :] = E[ids[b,t], :] * scale + pos[t, :]   (fp32 out)
int embed_tokens(const int64_t* ids, long long ids_stride, const int32_t* cu_seqlens, int B, int S,
                 const __nv_bfloat16* embed, long long vocab, const float* pos_table, int pos_rows, int D, float scale,
                 float* x, int* err_flag, cudaStream_t stream, int pos_offset = 0, __nv_bfloat16* h_out = nullptr,
                 float* stats_out = nullptr);  // h_out / stats_out: LnFold producer outputs (bf16 copy + row statistics)

// LnFold weight preparation: Wf = bf16(W diag(gamma)), colsum[n] = sum_k Wf[n,k], bias_f = bias + W beta
int fold_layernorm_weights(const __nv_bfloat16* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                           __nv_bfloat16* Wf, float* colsum, float* bias_f, cudaStream_t stream);

// y = LN(x) * gamma + beta, fp32 in, bf16 out, one warp per row
int layernorm_bf16(const float* x, const float* gamma, const float* beta, float eps, __nv_bfloat16* y, long long T,
                   int D, cudaStream_t stream);

// LN with an fp32 result (y32 may alias x) and/or a bf16 copy (either may be null)
int layernorm_dual(const float* x, const float* gamma, const float* beta, float eps, float* y32, __nv_bfloat16* y16,
                   long long T, int D, cudaStream_t stream);

// softmax(q k^T / sqrt(64)) v over packed sequences; qkv [T, 3*D] bf16 (q | k | v), out [T, D] bf16
// impl: 0 = auto (= 2), 1 = mma.sync flash kernel (tests / A-B only), 2 = tcgen05 (any length)
int attention_packed(const __nv_bfloat16* qkv, const int32_t* cu_seqlens, int B, int max_len, int H,
                     long long total_tokens, int impl, int num_sms, __nv_bfloat16* out, cudaStream_t stream);
int attention_packed_tc(const __nv_bfloat16* qkv, const int32_t* cu_seqlens, int B, int H, long long total_tokens,
                        __nv_bfloat16* out, int num_sms, cudaStream_t stream);

// Transformer-XL relative-position attention of the Conformer blocks on tcgen05 (attention_relpos_tc.cu):
// score(i,j) = ((q_i+u).k_j + (q_i+v).p[S_center-1-i+j]) / 8; qu / qv = [T, D] bf16 scratch for the biased queries
int attention_relpos_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* p, const float* u_bias, const float* v_bias,
                        const int32_t* cu_seqlens, int B, int H, long long total_tokens, int Npad, int S_center,
                        __nv_bfloat16* qu, __nv_bfloat16* qv, __nv_bfloat16* out, int num_sms, cudaStream_t stream);

// optional final LayerNorm + pooling over packed sequences -> out [B, D] fp32;
// optionally also scatters the (normalised) rows to a padded [B, S, D] fp32 tensor.
int ln_pool(const float* x, const int32_t* cu_seqlens, int B, int D, const float* gamma, const float* beta,
            float eps, int apply_ln, int pool_mode, float* out, float* encoded_padded, int S_padded,
            cudaStream_t stream);

}  // namespace sb
