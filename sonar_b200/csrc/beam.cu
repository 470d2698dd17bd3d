// One beam-search step of the embedding -> text generator as ONE kernel (one CTA per sentence).
//
// The reference drives fairseq2's BeamSearchSeq2SeqGenerator from Python (sonar/inference_pipelines/text.py:315-333):
// per generated token a few dozen small tensor ops select the 2*beam best continuations, retire the ones that end in
// EOS, and reorder the surviving hypotheses.  `sonar_b200/generation.py::_advance` restates that bookkeeping with
// vectorised torch ops (~60 kernel launches per step) and is the semantic definition -- it is what the CPU tests hold
// to the oracle (oracle/text_decoder.py::beam_search_step).  This kernel computes exactly the same state transition
// (bit for bit: tests/test_gpu_decoder.py) in one launch, which is what a launch-bound small batch needs.
//
// Ordering rule everywhere: score descending, then beam * vocab + token ascending, then candidate index ascending.
#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>

#include "../../include/sonar_b200.h"

namespace sb {
namespace {

constexpr int kBeamCand = kTopkCandidates;  // candidates per hypothesis handed over by the decoder step
constexpr int kMaxBeam = 7;                 // 2 * beam <= kBeamCand
constexpr int kMaxSel = 2 * kMaxBeam;
constexpr int kMaxCands = kMaxBeam * kBeamCand;

struct BeamStepParams {
  const float* lp;       // [R, 16] log-probs of the step's 16 best tokens per hypothesis
  const int32_t* tok;    // [R, 16]
  const float* eos_lp;   // [R]
  int64_t* seqs;         // [N, B, Tmax]
  int32_t* table;        // [R, Tmax] KV-cache ancestry
  int64_t* tokens;       // [R] next input tokens
  float* cum;            // [N, B]
  uint8_t* alive;        // [N, B] (torch.bool)
  uint8_t* done;         // [N]
  float* fin_score;      // [N, CAP + 1]
  int64_t* fin_seq;      // [N, CAP + 1, Tmax]
  int64_t* fin_len;      // [N, CAP + 1]
  int64_t* fin_count;    // [N]
  int N, B, Tmax, t, g, eos_block, max_gen;  // EOS is forbidden while g < eos_block (= min_gen_len - 1, [fs2] min_seq_len)
  long long vocab;
  int eos, unk, pad;
  float unk_penalty, score_div;
  int normalize;
};

__global__ void __launch_bounds__(128)
beam_step_kernel(const BeamStepParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int B = p.B, Tmax = p.Tmax, CAP = 2 * B, K2 = 2 * B, NC = B * kBeamCand;
  int64_t* old_seqs = reinterpret_cast<int64_t*>(smem);                       // [B][Tmax]
  int32_t* old_table = reinterpret_cast<int32_t*>(old_seqs + (size_t)B * Tmax);  // [B][Tmax]
  __shared__ float c_score[kMaxCands];
  __shared__ int c_tok[kMaxCands];
  __shared__ float s_score[kMaxSel];
  __shared__ int s_beam[kMaxSel], s_tok[kMaxSel];
  __shared__ int fin_dest[kMaxSel];                 // slot in fin_* or -1
  __shared__ int new_src[kMaxBeam], new_tok[kMaxBeam];
  __shared__ float new_cum[kMaxBeam];
  __shared__ int new_alive[kMaxBeam];

  // ---- stage the sentence's old hypotheses (the gathers below read them while the rows are overwritten) ----
  for (int i = tid; i < B * Tmax; i += blockDim.x) {
    old_seqs[i] = p.seqs[(size_t)n * B * Tmax + i];
    old_table[i] = p.table[(size_t)n * B * Tmax + i];
  }
  if (tid < K2) { s_score[tid] = -CUDART_INF_F; s_beam[tid] = 0; s_tok[tid] = 0; }
  const bool was_done = p.done[n] != 0;
  // ---- candidate scores ----
  for (int i = tid; i < NC; i += blockDim.x) {
    const int b = i / kBeamCand, c = i % kBeamCand;
    const int r = n * B + b;
    int tk = p.tok[(size_t)r * kBeamCand + c];
    float v = p.lp[(size_t)r * kBeamCand + c];
    if (tk < 0 || tk == p.pad) v = -CUDART_INF_F;
    if (p.unk_penalty != 0.f && tk == p.unk) v = v - p.unk_penalty;
    if (p.g < p.eos_block && tk == p.eos) v = -CUDART_INF_F;
    if (p.g >= p.max_gen - 1) {  // the last allowed token must be EOS
      v = -CUDART_INF_F;
      if (c == 0) { v = p.eos_lp[r]; tk = p.eos; }
    }
    float total = p.cum[n * B + b] + v;
    if (!p.alive[n * B + b]) total = -CUDART_INF_F;
    if (p.g == 0 && b > 0) total = -CUDART_INF_F;  // all beams are copies of the prompt
    c_score[i] = total;
    c_tok[i] = tk;
  }
  __syncthreads();
  // ---- the 2B best: rank by (score desc, beam * V + token asc, index asc) ----
  for (int i = tid; i < NC; i += blockDim.x) {
    const float si = c_score[i];
    const long long ki = (long long)(i / kBeamCand) * p.vocab + c_tok[i];
    int rank = 0;
    for (int j = 0; j < NC; ++j) {
      const float sj = c_score[j];
      const long long kj = (long long)(j / kBeamCand) * p.vocab + c_tok[j];
      const bool before = (sj > si) || (sj == si && (kj < ki || (kj == ki && j < i)));
      rank += before ? 1 : 0;
    }
    if (rank < K2) { s_score[rank] = si; s_beam[rank] = i / kBeamCand; s_tok[rank] = c_tok[i]; }
  }
  __syncthreads();
  // ---- sequential bookkeeping over the 2B selected candidates ----
  if (tid == 0) {
    long long fcount = p.fin_count[n];
    int nfin = 0, nkeep = 0;
    for (int b = 0; b < B; ++b) { new_src[b] = 0; new_tok[b] = p.pad; new_cum[b] = -CUDART_INF_F; new_alive[b] = 0; }
    for (int j = 0; j < K2; ++j) {
      const bool valid = s_score[j] > -CUDART_INF_F;
      const bool is_eos = valid && s_tok[j] == p.eos;
      fin_dest[j] = -1;
      if (is_eos && j < B && !was_done) {  // finalise EOS candidates ranked inside the beam
        const long long pos = fcount + nfin;
        ++nfin;
        if (pos < B) fin_dest[j] = (int)pos;  // the sentence closes the moment it owns `beam` hypotheses [fs2 _search_beam]
      }
      if (valid && !is_eos && !was_done) {  // next beam: the first B non-EOS candidates
        if (nkeep < B) {
          new_cum[nkeep] = s_score[j];
          new_alive[nkeep] = 1;
          new_src[nkeep] = s_beam[j];
          new_tok[nkeep] = s_tok[j];
        }
        ++nkeep;
      }
    }
    fcount += nfin;
    p.fin_count[n] = fcount;
    const bool now_done = was_done || fcount >= B;
    p.done[n] = now_done ? 1 : 0;
    p.fin_score[(size_t)n * (CAP + 1) + CAP] = -CUDART_INF_F;
    for (int b = 0; b < B; ++b) {
      p.cum[n * B + b] = new_cum[b];
      p.alive[n * B + b] = (new_alive[b] && !now_done) ? 1 : 0;
      p.tokens[n * B + b] = new_tok[b];
    }
  }
  __syncthreads();
  // ---- finished hypotheses ----
  for (int j = 0; j < K2; ++j) {
    const int dst = fin_dest[j];
    if (dst < 0) continue;
    int64_t* row = p.fin_seq + ((size_t)n * (CAP + 1) + dst) * Tmax;
    const int64_t* src = old_seqs + (size_t)s_beam[j] * Tmax;
    for (int i = tid; i < Tmax; i += blockDim.x) row[i] = (i == p.t + 1) ? (int64_t)s_tok[j] : src[i];
    if (tid == 0) {
      p.fin_score[(size_t)n * (CAP + 1) + dst] = p.normalize ? s_score[j] / p.score_div : s_score[j];
      p.fin_len[(size_t)n * (CAP + 1) + dst] = p.t + 2;
    }
  }
  // ---- the next beam: sequences and KV-cache ancestry follow their source hypothesis ----
  for (int b = 0; b < B; ++b) {
    const int sb_ = new_src[b];
    int64_t* srow = p.seqs + ((size_t)n * B + b) * Tmax;
    int32_t* trow = p.table + ((size_t)n * B + b) * Tmax;
    const int64_t* so = old_seqs + (size_t)sb_ * Tmax;
    const int32_t* to = old_table + (size_t)sb_ * Tmax;
    for (int i = tid; i < Tmax; i += blockDim.x) {
      srow[i] = (i == p.t + 1) ? (int64_t)new_tok[b] : so[i];
      trow[i] = (i == p.t) ? (n * B + sb_) : to[i];
    }
  }
}

}  // namespace
}  // namespace sb

extern "C" int sb_beam_step(const float* lp, const int32_t* tok, const float* eos_lp, int64_t* seqs, int32_t* table,
                            int64_t* tokens, float* cum, uint8_t* alive, uint8_t* done, float* fin_score,
                            int64_t* fin_seq, int64_t* fin_len, int64_t* fin_count, int32_t N, int32_t B, int32_t Tmax,
                            int32_t t, int32_t g, int32_t eos_block, int32_t max_gen, int64_t vocab, int32_t eos,
                            int32_t unk, int32_t pad, float unk_penalty, float score_div, int32_t normalize,
                            void* stream) {
  using namespace sb;
  if (!lp || !tok || !eos_lp || !seqs || !table || !tokens || !cum || !alive || !done || !fin_score || !fin_seq || !fin_len ||
      !fin_count) {
    set_last_error("sb_beam_step: null pointer");
    return SB_ERR_INVALID;
  }
  if (N <= 0 || B < 1 || B > kMaxBeam || Tmax < 2 || t < 0 || t + 1 >= Tmax || max_gen < 1) {
    set_last_error("sb_beam_step: bad shape (N=%d beam=%d Tmax=%d t=%d); beam must be 1..%d", N, B, Tmax, t, kMaxBeam);
    return SB_ERR_INVALID;
  }
  BeamStepParams p{lp, tok, eos_lp, seqs, table, tokens, cum, alive, done, fin_score, fin_seq, fin_len, fin_count,
                   N, B, Tmax, t, g, eos_block, max_gen, (long long)vocab, eos, unk, pad, unk_penalty, score_div, normalize};
  const size_t smem = (size_t)B * Tmax * (sizeof(int64_t) + sizeof(int32_t));
  if (smem > 48 * 1024) {
    set_last_error("sb_beam_step: beam * max_seq_len too large for the staging buffer (%zu bytes)", smem);
    return SB_ERR_INVALID;
  }
  beam_step_kernel<<<(unsigned)N, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("sb_beam_step launch failed: %s", cudaGetErrorString(e));
    return SB_ERR_CUDA;
  }
  return SB_OK;
}
