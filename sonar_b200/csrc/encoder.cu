// C ABI (include/sonar_b200.h) + the per-batch schedule of the SONAR text encoder:
//   embed -> 24 x [LN -> QKV GEMM -> attention -> out-proj GEMM(+residual)
//                  -> LN -> FFN1 GEMM(+ReLU) -> FFN2 GEMM(+residual)] -> final LN + pool
// following SonarTextTransformerEncoderModel.forward (sonar/models/sonar_text/model.py:130-143)
// with the `basic` wiring of sonar/models/sonar_text/factory.py:72-153.
//
// HBM layout of one batch (all buffers inside the caller's workspace):
//   tokens are PACKED: row cu_seqlens[b] + t holds token t of sentence b, so padded
//   positions never exist on the device (the reference computes them and masks them).
//   x   fp32 [T, D]   residual stream (fp32 so 48 residual adds do not accumulate bf16 rounding)
//   h   bf16 [T, D]   GEMM A operands: bf16 copy of x (LayerNorm folded into the GEMMs) or LayerNorm output; attention output
//   qkv bf16 [T, 3D]  fused q|k|v projections (its first [T, D] doubles as the bf16 copy of x behind the out-projection)
//   f   bf16 [T, F]   FFN inner activations
//
// Schedule with cfg.ln_fold = 1, 5 launches per layer, no LayerNorm kernel:
//   embed -> x, h = bf16(x), row stats
//   24 x [ QKV GEMM (folds LN1: stats + gamma/beta prepared into W', c, b') -> attention (tcgen05) ->
//          out-proj GEMM (+residual; emits x, bf16(x), stats) -> FFN1 GEMM (folds LN2, +ReLU) ->
//          FFN2 GEMM (+residual; emits x, bf16(x), stats) ] -> final LN + pool
// cfg.ln_fold = 0 (what the Python wrapper selects by default: measured faster, see bench.py `ab_layernorm_schedule`) keeps
// the classic schedule: separate LayerNorm kernels, residual adds by TMA reduce-add, 7 launches per layer.

#include "../../include/sonar_b200.h"
#include <stdlib.h>

#include "common.cuh"
#include "sonar_b200_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace sb {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  int32_t* cu;
  float* ln_stats;  // [T, D/128, 2] per-row LayerNorm partials (LnFold)
  float* x;
  __nv_bfloat16* h;
  __nv_bfloat16* qkv;
  __nv_bfloat16* f;
  size_t bytes;
};

}  // namespace sb

using namespace sb;

struct FoldedLayer {  // LnFold weights of one layer (device memory owned by the handle)
  __nv_bfloat16* wqkv = nullptr;
  __nv_bfloat16* w1 = nullptr;
  float* cqkv = nullptr;
  float* bqkv = nullptr;
  float* c1 = nullptr;
  float* b1 = nullptr;
};

struct SbEncoder {
  SbEncoderConfig cfg;
  std::vector<FoldedLayer> folded;
  void* fold_pool = nullptr;     // one allocation behind all FoldedLayer pointers
  int32_t* err_flag = nullptr;   // device: sticky "token id out of range" flag, cleared by sb_encoder_check_inputs
  const void* embed;
  const float* pos_table;
  const float* final_ln_g;
  const float* final_ln_b;
  std::vector<SbLayerWeights> layers;
  int num_sms;
  // pinned staging ring for cu_seqlens
  static constexpr int kSlots = 8;
  static constexpr int kSlotInts = 32768 + 8;
  int32_t* pinned = nullptr;
  cudaEvent_t ev[kSlots];
  bool ev_ok[kSlots];
  unsigned next_slot = 0;
  // optional in-step timing of the dominant kernel (FFN inner-projection GEMM of the middle layer)
  cudaEvent_t prof_start = nullptr, prof_stop = nullptr;
};

static Workspace carve(const SbEncoder* e, int32_t max_batch, int64_t max_tokens, void* base) {
  const size_t D = e->cfg.model_dim, F = e->cfg.ffn_inner_dim;
  const size_t T = (size_t)(max_tokens > 0 ? max_tokens : 1);
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  Workspace w;
  w.cu = reinterpret_cast<int32_t*>(p + off);
  off = align_up(off + sizeof(int32_t) * ((size_t)max_batch + 1), 1024);
  w.ln_stats = reinterpret_cast<float*>(p + off);
  off = align_up(off + T * (D / kLnPartCols) * 2 * sizeof(float), 1024);
  w.x = reinterpret_cast<float*>(p + off);
  off = align_up(off + T * D * 4, 1024);
  w.h = reinterpret_cast<__nv_bfloat16*>(p + off);
  off = align_up(off + T * D * 2, 1024);
  w.qkv = reinterpret_cast<__nv_bfloat16*>(p + off);
  off = align_up(off + T * 3 * D * 2, 1024);
  w.f = reinterpret_cast<__nv_bfloat16*>(p + off);
  off = align_up(off + T * F * 2, 1024);
  w.bytes = off;
  return w;
}

extern "C" {

const char* sb_last_error(void) { return g_err; }
int sb_version(void) { return 102; }

int sb_encoder_create(const SbEncoderConfig* cfg, const SbEncoderWeights* w, SbEncoder** out) {
  if (!cfg || !w || !out) { set_last_error("sb_encoder_create: null argument"); return SB_ERR_INVALID; }
  *out = nullptr;
  const int D = cfg->model_dim, H = cfg->num_heads, F = cfg->ffn_inner_dim;
  if (D <= 0 || D % 256 != 0 || D > 1024) {
    set_last_error("sb_encoder_create: model_dim must be a multiple of 256 and <= 1024 (got %d)", D);
    return SB_ERR_INVALID;
  }
  if (H <= 0 || D != H * 64) {
    set_last_error("sb_encoder_create: head_dim must be 64 (model_dim=%d, num_heads=%d)", D, H);
    return SB_ERR_INVALID;
  }
  if (F <= 0 || F % 256 != 0) { set_last_error("sb_encoder_create: ffn_inner_dim must be a multiple of 256"); return SB_ERR_INVALID; }
  if (cfg->ln_fold < 0 || cfg->ln_fold > 2) { set_last_error("sb_encoder_create: ln_fold must be 0, 1 or 2"); return SB_ERR_INVALID; }
  if (cfg->epi_groups < 0 || cfg->epi_groups > 2) { set_last_error("sb_encoder_create: epi_groups must be 0, 1 or 2"); return SB_ERR_INVALID; }
  if (cfg->epi_groups == 1 && (cfg->ln_fold != 0 || cfg->cta_group == 1)) {
    set_last_error("sb_encoder_create: epi_groups = 1 (the one-warpgroup epilogue kept for A/B runs) needs ln_fold = 0 and paired CTAs");
    return SB_ERR_INVALID;
  }
  if (cfg->num_layers < 0 || cfg->pos_rows <= 0 || cfg->vocab_size <= 0) {
    set_last_error("sb_encoder_create: bad num_layers / pos_rows / vocab_size");
    return SB_ERR_INVALID;
  }
  if (cfg->pooling != SB_POOL_MAX && cfg->pooling != SB_POOL_MEAN && cfg->pooling != SB_POOL_LAST) {
    set_last_error("sb_encoder_create: unsupported pooling %d", cfg->pooling);
    return SB_ERR_INVALID;
  }
  if (!w->embed || !w->pos_table || !w->final_ln_g || !w->final_ln_b || (cfg->num_layers > 0 && !w->layers)) {
    set_last_error("sb_encoder_create: missing weight pointer");
    return SB_ERR_INVALID;
  }
  int dev = 0, n_gpu = 0;
  if (cudaGetDeviceCount(&n_gpu) != cudaSuccess || n_gpu == 0) {
    set_last_error("sb_encoder_create: no CUDA device (this engine has no CPU path)");
    return SB_ERR_CUDA;
  }
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_last_error("sb_encoder_create: sm_100a kernels need a Blackwell B200-class GPU (found sm_%d%d)", prop.major,
                   prop.minor);
    return SB_ERR_CUDA;
  }
  SbEncoder* e = new (std::nothrow) SbEncoder();
  if (!e) { set_last_error("out of host memory"); return SB_ERR_INVALID; }
  e->cfg = *cfg;
  e->embed = w->embed;
  e->pos_table = w->pos_table;
  e->final_ln_g = w->final_ln_g;
  e->final_ln_b = w->final_ln_b;
  e->layers.assign(w->layers, w->layers + cfg->num_layers);
  for (int i = 0; i < cfg->num_layers; ++i) {
    const SbLayerWeights& l = e->layers[i];
    if (!l.wqkv || !l.bqkv || !l.wo || !l.bo || !l.w1 || !l.b1 || !l.w2 || !l.b2 || !l.ln1_g || !l.ln1_b ||
        !l.ln2_g || !l.ln2_b) {
      set_last_error("sb_encoder_create: layer %d has a null weight pointer", i);
      delete e;
      return SB_ERR_INVALID;
    }
  }
  e->num_sms = cfg->num_sms > 0 ? cfg->num_sms : prop.multiProcessorCount;
  for (int i = 0; i < SbEncoder::kSlots; ++i) e->ev_ok[i] = false;
  if (cudaMallocHost(reinterpret_cast<void**>(&e->pinned), sizeof(int32_t) * SbEncoder::kSlots * SbEncoder::kSlotInts) !=
      cudaSuccess) {
    set_last_error("sb_encoder_create: cudaMallocHost failed");
    delete e;
    return SB_ERR_CUDA;
  }
  if (cudaMalloc(reinterpret_cast<void**>(&e->err_flag), 256) != cudaSuccess ||
      cudaMemset(e->err_flag, 0, 256) != cudaSuccess) {
    set_last_error("sb_encoder_create: cudaMalloc of the input-check flag failed");
    sb_encoder_destroy(e);
    return SB_ERR_CUDA;
  }
  if (cfg->ln_fold && cfg->num_layers > 0) {
    // LayerNorm folding (LnFold): W' = W diag(gamma), c = row sums of W', b' = b + W beta for the two GEMMs that consume a
    // LayerNorm in every layer; prepared once here (the caller's weights are not modified)
    const size_t D_ = D, F_ = F;
    const size_t per_layer = align_up(3 * D_ * D_ * 2, 256) + align_up(F_ * D_ * 2, 256) + 2 * align_up(3 * D_ * 4, 256) +
                             2 * align_up(F_ * 4, 256);
    if (cudaMalloc(&e->fold_pool, per_layer * cfg->num_layers) != cudaSuccess) {
      set_last_error("sb_encoder_create: cudaMalloc of %zu bytes for the LayerNorm-folded weights failed",
                     per_layer * cfg->num_layers);
      sb_encoder_destroy(e);
      return SB_ERR_CUDA;
    }
    e->folded.resize(cfg->num_layers);
    uint8_t* p = reinterpret_cast<uint8_t*>(e->fold_pool);
    for (int i = 0; i < cfg->num_layers; ++i) {
      FoldedLayer& f = e->folded[i];
      const SbLayerWeights& l = e->layers[i];
      f.wqkv = reinterpret_cast<__nv_bfloat16*>(p); p += align_up(3 * D_ * D_ * 2, 256);
      f.w1 = reinterpret_cast<__nv_bfloat16*>(p); p += align_up(F_ * D_ * 2, 256);
      f.cqkv = reinterpret_cast<float*>(p); p += align_up(3 * D_ * 4, 256);
      f.bqkv = reinterpret_cast<float*>(p); p += align_up(3 * D_ * 4, 256);
      f.c1 = reinterpret_cast<float*>(p); p += align_up(F_ * 4, 256);
      f.b1 = reinterpret_cast<float*>(p); p += align_up(F_ * 4, 256);
      int rc = fold_layernorm_weights(reinterpret_cast<const __nv_bfloat16*>(l.wqkv), l.bqkv, l.ln1_g, l.ln1_b, 3 * D, D,
                                      f.wqkv, f.cqkv, f.bqkv, nullptr);
      if (!rc)
        rc = fold_layernorm_weights(reinterpret_cast<const __nv_bfloat16*>(l.w1), l.b1, l.ln2_g, l.ln2_b, F, D, f.w1, f.c1,
                                    f.b1, nullptr);
      if (rc) { sb_encoder_destroy(e); return rc; }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
      set_last_error("sb_encoder_create: folding the LayerNorm weights failed: %s", cudaGetErrorString(cudaGetLastError()));
      sb_encoder_destroy(e);
      return SB_ERR_CUDA;
    }
  }
  for (int i = 0; i < SbEncoder::kSlots; ++i) {
    if (cudaEventCreateWithFlags(&e->ev[i], cudaEventDisableTiming) != cudaSuccess) {
      set_last_error("sb_encoder_create: cudaEventCreate failed");
      sb_encoder_destroy(e);
      return SB_ERR_CUDA;
    }
    e->ev_ok[i] = true;
  }
  *out = e;
  return SB_OK;
}

void sb_encoder_destroy(SbEncoder* e) {
  if (!e) return;
  for (int i = 0; i < SbEncoder::kSlots; ++i)
    if (e->ev_ok[i]) cudaEventDestroy(e->ev[i]);
  if (e->pinned) cudaFreeHost(e->pinned);
  if (e->fold_pool) cudaFree(e->fold_pool);
  if (e->err_flag) cudaFree(e->err_flag);
  delete e;
}

int sb_encoder_workspace_bytes(const SbEncoder* enc, int32_t max_batch, int64_t max_tokens, size_t* bytes) {
  if (!enc || !bytes || max_batch <= 0 || max_tokens <= 0) {
    set_last_error("sb_encoder_workspace_bytes: bad argument");
    return SB_ERR_INVALID;
  }
  *bytes = carve(enc, max_batch, max_tokens, nullptr).bytes + 1024;  // + slack for base alignment
  return SB_OK;
}

int sb_encoder_forward(SbEncoder* e, const int64_t* ids, int64_t ids_row_stride, const int32_t* seq_lens_host,
                       int32_t B, int32_t S, float* out, float* encoded, void* workspace, size_t workspace_bytes,
                       void* stream_v) {
  if (!e || !ids || !out || !workspace) { set_last_error("sb_encoder_forward: null argument"); return SB_ERR_INVALID; }
  if (B <= 0 || S <= 0) { set_last_error("sb_encoder_forward: empty batch (B=%d, S=%d)", B, S); return SB_ERR_INVALID; }
  if (S > e->cfg.pos_rows) {
    set_last_error("sb_encoder_forward: seq_len %d exceeds the encoder's max_seq_len %d", S, e->cfg.pos_rows);
    return SB_ERR_INVALID;
  }
  if (B + 1 > SbEncoder::kSlotInts) { set_last_error("sb_encoder_forward: batch too large (%d)", B); return SB_ERR_INVALID; }
  if (ids_row_stride < S) { set_last_error("sb_encoder_forward: ids_row_stride < seq_len"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int D = e->cfg.model_dim, F = e->cfg.ffn_inner_dim, H = e->cfg.num_heads;

  // ---- cu_seqlens on the host, staged through a pinned ring ----
  const unsigned slot = e->next_slot++ % SbEncoder::kSlots;
  SB_CUDA_CHECK(cudaEventSynchronize(e->ev[slot]));  // only blocks if 8 forwards are still in flight
  int32_t* cu_h = e->pinned + (size_t)slot * SbEncoder::kSlotInts;
  long long T = 0;
  int max_len = 0;
  cu_h[0] = 0;
  for (int b = 0; b < B; ++b) {
    const int len = seq_lens_host ? seq_lens_host[b] : S;
    if (len < 0 || len > S) { set_last_error("sb_encoder_forward: seq_lens[%d]=%d outside [0,%d]", b, len, S); return SB_ERR_INVALID; }
    T += len;
    if (len > max_len) max_len = len;
    if (T > 0x7fffffffll) { set_last_error("sb_encoder_forward: too many tokens"); return SB_ERR_INVALID; }
    cu_h[b + 1] = (int32_t)T;
  }
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  Workspace w = carve(e, B, T, reinterpret_cast<void*>(base));
  if (base - reinterpret_cast<uintptr_t>(workspace) + w.bytes > workspace_bytes) {
    set_last_error("sb_encoder_forward: workspace too small (%zu bytes given, %zu needed for B=%d T=%lld)",
                   workspace_bytes, (size_t)(base - reinterpret_cast<uintptr_t>(workspace)) + w.bytes, B, T);
    return SB_ERR_INVALID;
  }
  SB_CUDA_CHECK(cudaMemcpyAsync(w.cu, cu_h, sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, stream));
  SB_CUDA_CHECK(cudaEventRecord(e->ev[slot], stream));
  if (T == 0) {
    SB_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * D, stream));
    return SB_OK;
  }

  const bool fold = e->cfg.ln_fold != 0 && e->cfg.num_layers > 0;  // LN1 (attention block) folded into FFN2 -> QKV
  const bool fold2 = fold && e->cfg.ln_fold == 1;                   // LN2 (FFN block) folded into out-proj -> FFN1 as well
  __nv_bfloat16* hn = w.qkv;  // [T, D] view of the (dead after attention) qkv buffer: bf16(x) behind the out-projection
  int rc;
  if ((rc = embed_tokens(ids, ids_row_stride, w.cu, B, S, reinterpret_cast<const __nv_bfloat16*>(e->embed),
                         e->cfg.vocab_size, e->pos_table, e->cfg.pos_rows, D, e->cfg.embed_scale, w.x, e->err_flag,
                         stream, 0, fold ? w.h : nullptr, fold ? w.ln_stats : nullptr)))
    return rc;

  GemmArgs g;
  g.cta_group = (e->cfg.cta_group == 1) ? 1 : 2;
  g.num_sms = e->num_sms;
  g.M = (int)T;
  g.epi_groups = (e->cfg.epi_groups == 1) ? 1 : 2;
  LnFold consume;  // what a LayerNorm-consuming GEMM needs
  consume.stats_in = w.ln_stats;
  consume.chunks = D / kLnPartCols;
  consume.eps = e->cfg.ln_eps;
  for (int li = 0; li < e->cfg.num_layers; ++li) {
    const SbLayerWeights& L = e->layers[li];
    // --- self-attention block: x += Wo . SDPA(LN1(x)) + bo ---
    if (!fold)
      if ((rc = layernorm_bf16(w.x, L.ln1_g, L.ln1_b, e->cfg.ln_eps, w.h, T, D, stream))) return rc;
    g.A = w.h; g.lda = D; g.ldw = D;
    g.C = w.qkv; g.ldc = 3 * D; g.out_fp32 = 0; g.residual = nullptr; g.ldr = 0;
    g.N = 3 * D; g.K = D; g.epi = EPI_BIAS;
    if (fold) {
      const FoldedLayer& f = e->folded[li];
      g.W = f.wqkv; g.bias = f.bqkv; g.lf = consume; g.lf.colsum = f.cqkv;
    } else {
      g.W = reinterpret_cast<const __nv_bfloat16*>(L.wqkv); g.bias = L.bqkv;
    }
    rc = gemm_bf16(g, stream);
    g.lf = LnFold();
    if (rc) return rc;
    if ((rc = attention_packed(w.qkv, w.cu, B, max_len, H, T, 0, e->num_sms, w.h, stream))) return rc;
    g.A = w.h; g.lda = D; g.W = reinterpret_cast<const __nv_bfloat16*>(L.wo); g.ldw = D;
    g.C = w.x; g.ldc = D; g.out_fp32 = 1; g.bias = L.bo; g.residual = w.x; g.ldr = D;
    g.N = D; g.K = D; g.epi = EPI_BIAS_RESIDUAL;
    if (fold2) {  // emits x, hn = bf16(x) and the statistics LN2 needs
      g.epi = EPI_BIAS_RESIDUAL_STATS;
      g.lf.h_out = hn; g.lf.ldh = D; g.lf.stats_out = w.ln_stats;
    }
    rc = gemm_bf16(g, stream);
    g.lf = LnFold();
    if (rc) return rc;
    // --- feed-forward block: x += W2 . relu(W1 . LN2(x) + b1) + b2 ---
    if (!fold2)
      if ((rc = layernorm_bf16(w.x, L.ln2_g, L.ln2_b, e->cfg.ln_eps, w.h, T, D, stream))) return rc;
    g.A = fold2 ? hn : w.h; g.lda = D; g.ldw = D;
    g.C = w.f; g.ldc = F; g.out_fp32 = 0; g.residual = nullptr; g.ldr = 0;
    g.N = F; g.K = D; g.epi = EPI_BIAS_RELU;
    if (fold2) {
      const FoldedLayer& f = e->folded[li];
      g.W = f.w1; g.bias = f.b1; g.lf = consume; g.lf.colsum = f.c1;
    } else {
      g.W = reinterpret_cast<const __nv_bfloat16*>(L.w1); g.bias = L.b1;
    }
    const bool prof = e->prof_start && li == e->cfg.num_layers / 2;
    if (prof) SB_CUDA_CHECK(cudaEventRecord(e->prof_start, stream));
    rc = gemm_bf16(g, stream);
    g.lf = LnFold();
    if (rc) return rc;
    if (prof) SB_CUDA_CHECK(cudaEventRecord(e->prof_stop, stream));
    g.A = w.f; g.lda = F; g.W = reinterpret_cast<const __nv_bfloat16*>(L.w2); g.ldw = F;
    g.C = w.x; g.ldc = D; g.out_fp32 = 1; g.bias = L.b2; g.residual = w.x; g.ldr = D;
    g.N = D; g.K = F; g.epi = EPI_BIAS_RESIDUAL;
    if (fold && li + 1 < e->cfg.num_layers) {  // emits x, h = bf16(x) and the statistics the next layer's LN1 needs
      g.epi = EPI_BIAS_RESIDUAL_STATS;
      g.lf.h_out = w.h; g.lf.ldh = D; g.lf.stats_out = w.ln_stats;
    }
    rc = gemm_bf16(g, stream);
    g.lf = LnFold();
    if (rc) return rc;
  }
  return ln_pool(w.x, w.cu, B, D, e->final_ln_g, e->final_ln_b, e->cfg.ln_eps, 1, e->cfg.pooling, out, encoded, S,
                 stream);
}

int sb_encoder_forward_host(SbEncoder* e, const int64_t* ids_host, const int32_t* seq_lens_host, int32_t B,
                            int32_t S, float* out_host, int64_t* ids_staging, float* out_staging, void* workspace,
                            size_t workspace_bytes, void* stream_v) {
  if (!e || !ids_host || !out_host || !ids_staging || !out_staging) {
    set_last_error("sb_encoder_forward_host: null argument");
    return SB_ERR_INVALID;
  }
  if (B <= 0 || S <= 0) { set_last_error("sb_encoder_forward_host: empty batch"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  SB_CUDA_CHECK(cudaMemcpyAsync(ids_staging, ids_host, sizeof(int64_t) * (size_t)B * S, cudaMemcpyHostToDevice, stream));
  int rc = sb_encoder_forward(e, ids_staging, S, seq_lens_host, B, S, out_staging, nullptr, workspace, workspace_bytes,
                              stream_v);
  if (rc) return rc;
  SB_CUDA_CHECK(cudaMemcpyAsync(out_host, out_staging, sizeof(float) * (size_t)B * e->cfg.model_dim,
                                cudaMemcpyDeviceToHost, stream));
  SB_CUDA_CHECK(cudaStreamSynchronize(stream));
  return SB_OK;
}

int sb_encoder_profile_ffn1(SbEncoder* e, void* start_event, void* stop_event) {
  if (!e || (!start_event) != (!stop_event)) { set_last_error("sb_encoder_profile_ffn1: bad argument"); return SB_ERR_INVALID; }
  e->prof_start = reinterpret_cast<cudaEvent_t>(start_event);
  e->prof_stop = reinterpret_cast<cudaEvent_t>(stop_event);
  return SB_OK;
}

int sb_encoder_check_inputs(SbEncoder* e, void* workspace, void* stream_v) {
  (void)workspace;  // (kept in the signature; the flag lives in the handle since v101)
  if (!e) { set_last_error("sb_encoder_check_inputs: null argument"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int32_t flag = 0;
  SB_CUDA_CHECK(cudaMemcpyAsync(&flag, e->err_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  SB_CUDA_CHECK(cudaMemsetAsync(e->err_flag, 0, sizeof(int32_t), stream));  // sticky until read: covers every forward since
  SB_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (flag != 0) {
    set_last_error("token id outside [0, vocab_size) in a batch passed to sb_encoder_forward since the last check");
    return SB_ERR_INPUT;
  }
  return SB_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel-level entry points
// ---------------------------------------------------------------------------------------------
int sb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int32_t out_fp32,
                 const float* bias, const void* residual, int64_t ldr, int32_t M, int32_t N, int32_t K, int32_t epi,
                 int32_t cta_group, void* stream) {
  if (!A || !W || !C) { set_last_error("sb_gemm_bf16: null pointer"); return SB_ERR_INVALID; }
  GemmArgs g;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.lda = lda;
  g.W = reinterpret_cast<const __nv_bfloat16*>(W); g.ldw = ldw;
  g.C = C; g.ldc = ldc; g.out_fp32 = out_fp32; g.bias = bias; g.residual = residual; g.ldr = ldr;
  g.M = M; g.N = N; g.K = K; g.epi = epi;
  g.cta_group = cta_group == 1 ? 1 : 2;
  g.allow_skinny = (cta_group == 0);  // 0 = automatic: M <= 64 may take the weight-streaming path
  int dev = 0, sms = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  SB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g.num_sms = sms;
  return gemm_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}

int sb_fold_layernorm(const void* W, const float* bias, const float* gamma, const float* beta, int32_t N, int32_t K,
                      void* Wf, float* colsum, float* bias_f, void* stream) {
  if (!W || !bias || !gamma || !beta || !Wf || !colsum || !bias_f || N <= 0 || K <= 0) {
    set_last_error("sb_fold_layernorm: bad argument");
    return SB_ERR_INVALID;
  }
  return fold_layernorm_weights(reinterpret_cast<const __nv_bfloat16*>(W), bias, gamma, beta, N, K,
                                reinterpret_cast<__nv_bfloat16*>(Wf), colsum, bias_f, reinterpret_cast<cudaStream_t>(stream));
}

int sb_gemm_ln_consumer(const void* A, int64_t lda, const void* Wf, int64_t ldw, void* C, int64_t ldc, const float* bias_f,
                        const float* colsum, const float* stats, float eps, int32_t M, int32_t N, int32_t K, int32_t relu,
                        void* stream) {
  if (!A || !Wf || !C || !bias_f || !colsum || !stats) { set_last_error("sb_gemm_ln_consumer: null pointer"); return SB_ERR_INVALID; }
  GemmArgs g;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.lda = lda;
  g.W = reinterpret_cast<const __nv_bfloat16*>(Wf); g.ldw = ldw;
  g.C = C; g.ldc = ldc; g.out_fp32 = 0; g.bias = bias_f; g.residual = nullptr; g.ldr = 0;
  g.M = M; g.N = N; g.K = K; g.epi = relu ? EPI_BIAS_RELU : EPI_BIAS;
  g.cta_group = 2;
  g.lf.stats_in = stats; g.lf.colsum = colsum; g.lf.chunks = K / kLnPartCols; g.lf.eps = eps;
  int dev = 0, sms = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  SB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g.num_sms = sms;
  return gemm_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}

int sb_gemm_residual_stats(const void* A, int64_t lda, const void* W, int64_t ldw, float* x, int64_t ldx, const float* bias,
                           void* h_out, int64_t ldh, float* stats_out, int32_t M, int32_t N, int32_t K, void* stream) {
  if (!A || !W || !x || !bias || !h_out || !stats_out) { set_last_error("sb_gemm_residual_stats: null pointer"); return SB_ERR_INVALID; }
  GemmArgs g;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.lda = lda;
  g.W = reinterpret_cast<const __nv_bfloat16*>(W); g.ldw = ldw;
  g.C = x; g.ldc = ldx; g.out_fp32 = 1; g.bias = bias; g.residual = x; g.ldr = ldx;
  g.M = M; g.N = N; g.K = K; g.epi = EPI_BIAS_RESIDUAL_STATS;
  g.cta_group = 2;
  g.lf.h_out = reinterpret_cast<__nv_bfloat16*>(h_out); g.lf.ldh = ldh; g.lf.stats_out = stats_out;
  int dev = 0, sms = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  SB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g.num_sms = sms;
  return gemm_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}

int sb_gemm_residual_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* x, int64_t ldx, const float* bias,
                            int32_t M, int32_t N, int32_t K, int32_t* counters, int64_t n_counters, void* stream) {
  if (!A || !W || !x || !bias || !counters) { set_last_error("sb_gemm_residual_splitk: null pointer"); return SB_ERR_INVALID; }
  GemmArgs g;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.lda = lda;
  g.W = reinterpret_cast<const __nv_bfloat16*>(W); g.ldw = ldw;
  g.C = x; g.ldc = ldx; g.out_fp32 = 1; g.bias = bias; g.residual = x; g.ldr = ldx;
  g.M = M; g.N = N; g.K = K; g.epi = EPI_BIAS_RESIDUAL;
  g.cta_group = 2;
  g.splitk_flags = counters;
  g.splitk_flags_len = n_counters;
  int dev = 0, sms = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  SB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g.num_sms = sms;
  return gemm_bf16(g, reinterpret_cast<cudaStream_t>(stream));
}

int sb_layernorm(const float* x, const float* gamma, const float* beta, float eps, void* y, int64_t T, int32_t D,
                 void* stream) {
  if (!x || !gamma || !beta || !y) { set_last_error("sb_layernorm: null pointer"); return SB_ERR_INVALID; }
  return layernorm_bf16(x, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y), T, D,
                        reinterpret_cast<cudaStream_t>(stream));
}

int sb_attention(const void* qkv, const int32_t* cu_seqlens, int32_t B, int32_t max_len, int32_t H,
                 int64_t total_tokens, int32_t impl, void* out, void* stream) {
  if (!qkv || !cu_seqlens || !out) { set_last_error("sb_attention: null pointer"); return SB_ERR_INVALID; }
  int dev = 0, sms = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  SB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return attention_packed(reinterpret_cast<const __nv_bfloat16*>(qkv), cu_seqlens, B, max_len, H, total_tokens, impl,
                          sms, reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<cudaStream_t>(stream));
}

int sb_embed(const int64_t* ids, int64_t ids_row_stride, const int32_t* cu_seqlens, int32_t B, int32_t S,
             const void* embed, int64_t vocab, const float* pos_table, int32_t pos_rows, int32_t D, float scale,
             float* x, int32_t* err_flag, void* stream) {
  if (!ids || !cu_seqlens || !embed || !pos_table || !x || !err_flag) {
    set_last_error("sb_embed: null pointer");
    return SB_ERR_INVALID;
  }
  return embed_tokens(ids, ids_row_stride, cu_seqlens, B, S, reinterpret_cast<const __nv_bfloat16*>(embed), vocab,
                      pos_table, pos_rows, D, scale, x, err_flag, reinterpret_cast<cudaStream_t>(stream));
}

int sb_pool(const float* x, const int32_t* cu_seqlens, int32_t B, int32_t D, const float* gamma, const float* beta,
            float eps, int32_t apply_ln, int32_t pool_mode, float* out, float* encoded_padded, int32_t S_padded,
            void* stream) {
  if (!x || !cu_seqlens || !out) { set_last_error("sb_pool: null pointer"); return SB_ERR_INVALID; }
  return ln_pool(x, cu_seqlens, B, D, gamma, beta, eps, apply_ln, pool_mode, out, encoded_padded, S_padded,
                 reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
