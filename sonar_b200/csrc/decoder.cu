// SONAR embedding -> text decoder, one incremental step at a time (BASELINE.json config 4).
//
// Reference: ConditionalTransformerDecoderModel.decode/project (sonar/nn/conditional_decoder_model.py:60-94),
// wiring sonar/models/sonar_text/factory.py:229-315 (pre-LN layers: causal self-attention with a KV cache,
// encoder-decoder attention, ReLU FFN; final LayerNorm; logits = h . E^T with the tied embedding matrix),
// driven one token at a time by fairseq2's BeamSearchSeq2SeqGenerator (sonar/inference_pipelines/text.py:305-346).
//
// The source is the sentence embedding as a SINGLE encoder position (sonar/models/sonar_translation/model.py:48-53),
// so every cross-attention softmax is over one key and equals 1: the layer's cross-attention output is the
// per-sentence constant  c_l = Wo_l (Wv_l e + bv_l) + bo_l  (q/k projections are dead compute).  sb_decoder_begin
// computes c_l once per sentence with two GEMMs per layer; each step then only adds it.
//
// Per step (R = sentences x beam rows, all at the same position t):
//   x = E[token] * sqrt(d) + pos[t]
//   24 x { h = LN(x); qkv = h Wqkv^T (tcgen05 GEMM); K/V appended to the cache; attention over positions 0..t
//          through a per-row ancestry table (beam reordering never moves the cache); x += o Wo^T + bo (TMA reduce-add);
//          x += c_l[sentence]; h = LN(x); x += W2 relu(W1 h + b1) + b2 }
//   h = LN_final(x);  logits = h E^T  -- never materialised: the tcgen05 GEMM's sweep epilogue keeps a running
//   top-16 and an online log-sum-exp per row over all 256 206 columns; a merge kernel turns the per-chunk partials
//   into the 16 best (log-prob, token) pairs per row plus log P(EOS).

#include "../../include/sonar_b200.h"
#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>
#include <new>
#include <vector>

namespace sb {

static inline size_t align_up_d(size_t v, size_t a) { return (v + a - 1) / a * a; }

// x[r,:] = E[token[r],:] * scale + pos[t,:]      (one warp per row)
__global__ void __launch_bounds__(256)
decode_embed_kernel(const int64_t* __restrict__ tokens, const __nv_bfloat16* __restrict__ embed, long long vocab,
                    const float* __restrict__ pos_row, int D, float scale, float* __restrict__ x, int R,
                    int* __restrict__ err_flag) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  long long id = tokens[r];
  if (id < 0 || id >= vocab) {
    if (lane == 0) atomicExch(err_flag, 1);
    id = 0;
  }
  const uint4* erow = reinterpret_cast<const uint4*>(embed + id * (long long)D);
  const float4* prow = reinterpret_cast<const float4*>(pos_row);
  float4* xrow = reinterpret_cast<float4*>(x + (long long)r * D);
  for (int c = lane; c < D / 8; c += 32) {
    const uint4 e = __ldg(erow + c);
    const float4 p0 = __ldg(prow + 2 * c), p1 = __ldg(prow + 2 * c + 1);
    const uint32_t w[4] = {e.x, e.y, e.z, e.w};
    float f[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
      f[2 * q] = __low2float(v);
      f[2 * q + 1] = __high2float(v);
    }
    xrow[2 * c] = make_float4(fmaf(f[0], scale, p0.x), fmaf(f[1], scale, p0.y), fmaf(f[2], scale, p0.z), fmaf(f[3], scale, p0.w));
    xrow[2 * c + 1] = make_float4(fmaf(f[4], scale, p1.x), fmaf(f[5], scale, p1.y), fmaf(f[6], scale, p1.z), fmaf(f[7], scale, p1.w));
  }
}

// x[r,:] += c[r / beam, :]   (fp32, one warp per row)
__global__ void __launch_bounds__(256)
add_sentence_const_kernel(float* __restrict__ x, const float* __restrict__ c, int R, int beam, int D) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float4* xr = reinterpret_cast<float4*>(x + (long long)r * D);
  const float4* cr = reinterpret_cast<const float4*>(c + (long long)(r / beam) * D);
  for (int q = lane; q < D / 4; q += 32) {
    float4 a = xr[q];
    const float4 b = __ldg(cr + q);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    xr[q] = a;
  }
}

// x[r,:] += c[r / beam, :], then h[r,:] = LayerNorm(x[r,:]) in bf16: the collapsed cross-attention residual and the FFN
// LayerNorm in one pass over the row (bit-identical to add_sentence_const_kernel followed by layernorm_bf16).
__global__ void __launch_bounds__(256)
add_const_layernorm_kernel(float* __restrict__ x, const float* __restrict__ c, int R, int beam, int D,
                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                           __nv_bfloat16* __restrict__ h) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int nvec = D / 128;
  float4 v[kMaxVec], cv[kMaxVec];
  load_row(x + (long long)r * D, nvec, lane, v);
  load_row(c + (long long)(r / beam) * D, nvec, lane, cv);
  float4* xr = reinterpret_cast<float4*>(x + (long long)r * D);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) {
      v[i].x += cv[i].x; v[i].y += cv[i].y; v[i].z += cv[i].z; v[i].w += cv[i].w;
      xr[i * 32 + lane] = v[i];
    }
  normalize_row(v, nvec, lane, D, gamma, beta, eps);
  uint2* hr = reinterpret_cast<uint2*>(h + (long long)r * D);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) hr[i * 32 + lane] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
}

// fp32 [n, D] -> bf16 (plain cast; A operand of the cross-attention constant GEMMs)
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

// Incremental causal self-attention: one warp per (row, head).  Appends this step's K/V to the cache at
// position t (physical row r) and attends over positions 0..t, where position t' < t of hypothesis r lives in
// physical cache row table[r, t'] (its ancestor at that step).
// The warp walks the keys FOUR at a time: lane group g = lane/8 owns keys g, g+4, ..., and inside a group each lane
// owns 8 of the 64 head dims, so every K and V access is one 16-byte load per lane and 128 contiguous bytes per group.
// Each group keeps its own online-softmax state (max, sum, 8 accumulators per lane); the four states are merged
// with two shuffle rounds at the end.  HBM-bound: 2 x 128 B per (hypothesis, head, cached position).
constexpr int kMaxDecodeLen = 512;  // positions <= the decoder's position table

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(b);
    f[2 * e + 1] = __high2float(b);
  }
}

__global__ void __launch_bounds__(128)
decode_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                        __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ table, int t, int Tmax, int H,
                        __nv_bfloat16* __restrict__ out) {
  const int r = blockIdx.y;
  const int h = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (h >= H) return;
  const int D = H * 64;
  const __nv_bfloat16* row = qkv + (long long)r * 3 * D + h * 64;
  {  // append this step's K / V (2 dims per lane) to the cache
    const long long own = ((long long)r * Tmax + t) * D + h * 64 + lane * 2;
    *reinterpret_cast<__nv_bfloat162*>(kcache + own) = *reinterpret_cast<const __nv_bfloat162*>(row + D + lane * 2);
    *reinterpret_cast<__nv_bfloat162*>(vcache + own) = *reinterpret_cast<const __nv_bfloat162*>(row + 2 * D + lane * 2);
  }
  const int grp = lane >> 3, sub = lane & 7;
  float q8[8];
  unpack8(*reinterpret_cast<const uint4*>(row + sub * 8), q8);
  __syncwarp();  // the row appended above is read back below by other lanes of this warp
  const int nk = t + 1;
  const int32_t* trow = table + (long long)r * Tmax;
  const float sl2 = 0.125f * 1.4426950408889634f;
  float m = -CUDART_INF_F, l = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // 16 keys per pass (4 per 8-lane group): the ancestry-table entries, then all eight 16-byte K / V loads of a lane are
  // in flight before the first dot product -- the loop is bound by loaded HBM latency, not by arithmetic.
  constexpr int U = 4;
  for (int base = 0; base < nk; base += 4 * U) {
    int tpv[U];
    bool val[U];
    int prow[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      tpv[u] = base + 4 * u + grp;
      val[u] = tpv[u] < nk;
      prow[u] = (val[u] && tpv[u] != t) ? __ldg(trow + tpv[u]) : r;
    }
    uint4 kq[U], vq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long off = ((long long)prow[u] * Tmax + (val[u] ? tpv[u] : t)) * D + h * 64 + sub * 8;
      kq[u] = vq[u] = make_uint4(0u, 0u, 0u, 0u);
      if (val[u]) {
        kq[u] = *reinterpret_cast<const uint4*>(kcache + off);
        vq[u] = *reinterpret_cast<const uint4*>(vcache + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float k8[8], v8[8];
      unpack8(kq[u], k8);
      unpack8(vq[u], v8);
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(q8[e], k8[e], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (val[u]) {  // uniform inside each 8-lane group
        const float mn = fmaxf(m, s);
        const float corr = exp2f((m - mn) * sl2);  // m = -inf on the group's first key -> 0
        const float pj = exp2f((s - mn) * sl2);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, v8[e], acc[e] * corr);
        m = mn;
      }
    }
  }
  // merge the four group states (a group that saw no key has m = -inf, l = 0)
  float M = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
  M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 16));
  const float sc = (m == -CUDART_INF_F) ? 0.f : exp2f((m - M) * sl2);
  l *= sc;
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float a = acc[e] * sc;
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 16);
    acc[e] = a;
  }
  if (grp == 0) {
    const float inv = 1.0f / l;
    *reinterpret_cast<uint4*>(out + (long long)r * D + h * 64 + sub * 8) =
        make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                   pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  }
}

// Merge the per-chunk partials of the vocabulary GEMM: one warp per row.
//   lse = log sum_j exp(logit_j) over all columns; out: 16 best (logit - lse, token), order (value desc, token asc);
//   eos_lprob = <h, E[eos]> - lse (needed when the generator must force EOS and EOS is not among the 16).
template <int KC>
__global__ void __launch_bounds__(256)
vocab_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, const float* __restrict__ lse_part,
                   int n_chunks, const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ embed, int D,
                   int eos_idx, int R, float* __restrict__ out_lprob, int* __restrict__ out_tok,
                   float* __restrict__ out_eos, const int64_t* __restrict__ probe_tokens, long long vocab,
                   float* __restrict__ out_probe) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  // ---- log-sum-exp ----
  float M = -CUDART_INF_F;
  for (int c = lane; c < n_chunks; c += 32) M = fmaxf(M, lse_part[((long long)r * n_chunks + c) * 2]);
  M = warp_max(M);
  float S = 0.f;
  for (int c = lane; c < n_chunks; c += 32) {
    const float mc = lse_part[((long long)r * n_chunks + c) * 2];
    const float sc = lse_part[((long long)r * n_chunks + c) * 2 + 1];
    if (mc > -CUDART_INF_F) S += sc * __expf(mc - M);
  }
  S = warp_sum(S);
  const float lse = M + logf(S);
  // ---- top-KC of the n_chunks*KC candidates: every lane keeps a strided slice, then KC rounds of warp arg-max ----
  const int total = n_chunks * KC;
  const float* cv = cand_val + (long long)r * total;
  const int* ci = cand_idx + (long long)r * total;
  unsigned long long taken0 = 0ull, taken1 = 0ull;  // lane-local bitmap over its slice (slice <= 128 entries: n_lists <= 256)
  for (int k = 0; k < KC; ++k) {
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff, bpos = -1;
    int s = 0;
    for (int p = lane; p < total; p += 32, ++s) {
      if (((s < 64) ? taken0 : taken1) & (1ull << (s & 63))) continue;
      const float v = cv[p];
      const int i = ci[p];
      if (i < 0) continue;
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bpos = s; }
    }
    // warp arg-max by (value desc, token asc)
    float wv = bv;
    int wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
    }
    if (bpos >= 0 && bv == wv && bi == wi) {  // token ids are unique -> exactly one lane
      if (bpos < 64) taken0 |= (1ull << bpos); else taken1 |= (1ull << (bpos - 64));
    }
    if (lane == 0) {
      const bool valid = wi != 0x7fffffff;
      out_lprob[(long long)r * KC + k] = valid ? (wv - lse) : -CUDART_INF_F;
      out_tok[(long long)r * KC + k] = valid ? wi : -1;
    }
  }
  // ---- log P(EOS) ----
  float dot = 0.f;
  const __nv_bfloat16* hr = h + (long long)r * D;
  const __nv_bfloat16* er = embed + (long long)eos_idx * D;
  for (int q = lane * 2; q < D; q += 64) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(hr + q);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(er + q);
    dot += __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
  }
  dot = warp_sum(dot);
  if (lane == 0) out_eos[r] = dot - lse;
  // ---- log P(probe token): the prompt scores the generator's prefill accumulates ----
  if (probe_tokens != nullptr) {
    long long pt = probe_tokens[r];
    if (pt < 0 || pt >= vocab) pt = 0;  // out-of-range ids are reported by decode_embed_kernel when they are fed back
    const __nv_bfloat16* pr = embed + pt * (long long)D;
    float pd = 0.f;
    for (int q = lane * 2; q < D; q += 64) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(hr + q);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(pr + q);
      pd += __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
    }
    pd = warp_sum(pd);
    if (lane == 0) out_probe[r] = pd - lse;
  }
}

}  // namespace sb

using namespace sb;

struct SbDecoder {
  SbDecoderConfig cfg;
  const void* embed;
  const float* pos_table;
  const float* final_ln_g;
  const float* final_ln_b;
  std::vector<SbDecoderLayerWeights> layers;
  int num_sms;
};

namespace {

constexpr long long kSplitkFlags = 4096;  // >= 4 * (tile pairs of an [R, D] residual GEMM) whenever splitting can pay

struct DecWs {
  int32_t* err_flag;
  int* splitk_flags;      // [kSplitkFlags] hand-over counters of the split-K residual GEMMs (zeroed by sb_decoder_begin)
  float* x;               // [R, D] fp32 residual stream of the current step
  __nv_bfloat16* h;       // [R, D]
  __nv_bfloat16* qkv;     // [R, 3D]
  __nv_bfloat16* f;       // [R, F]
  float* cross;           // [L, N, D] per-sentence cross-attention constants
  __nv_bfloat16* ebf;     // [N, D] bf16 sentence embeddings
  __nv_bfloat16* vtmp;    // [N, D]
  float* cand_val;        // [R, n_lists, 16]   n_lists = gemm_topk_lists(n_chunks)
  int* cand_idx;
  float* lse_part;        // [R, n_lists, 2]
  __nv_bfloat16* kcache;  // [L, R, Tmax, D]
  __nv_bfloat16* vcache;
  int n_chunks;
  size_t bytes;
};

DecWs carve_dec(const SbDecoder* d, int N, int beam, int Tmax, void* base) {
  const size_t D = d->cfg.model_dim, F = d->cfg.ffn_inner_dim, L = d->cfg.num_layers;
  const size_t R = (size_t)N * beam;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  DecWs w;
  auto take = [&](size_t bytes) { uint8_t* q = p + off; off = align_up_d(off + bytes, 1024); return q; };
  w.n_chunks = gemm_topk_chunks((int)R, (int)d->cfg.vocab_size, 2, d->num_sms);
  w.err_flag = reinterpret_cast<int32_t*>(take(256));
  w.splitk_flags = reinterpret_cast<int*>(take(kSplitkFlags * sizeof(int)));
  w.x = reinterpret_cast<float*>(take(R * D * 4));
  w.h = reinterpret_cast<__nv_bfloat16*>(take(R * D * 2));
  w.qkv = reinterpret_cast<__nv_bfloat16*>(take(R * 3 * D * 2));
  w.f = reinterpret_cast<__nv_bfloat16*>(take(R * F * 2));
  w.cross = reinterpret_cast<float*>(take(L * (size_t)N * D * 4));
  w.ebf = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * D * 2));
  w.vtmp = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * D * 2));
  const size_t n_lists = (size_t)gemm_topk_lists(w.n_chunks);
  w.cand_val = reinterpret_cast<float*>(take(R * n_lists * kTopkCandidates * 4));
  w.cand_idx = reinterpret_cast<int*>(take(R * n_lists * kTopkCandidates * 4));
  w.lse_part = reinterpret_cast<float*>(take(R * n_lists * 2 * 4));
  w.kcache = reinterpret_cast<__nv_bfloat16*>(take(L * R * (size_t)Tmax * D * 2));
  w.vcache = reinterpret_cast<__nv_bfloat16*>(take(L * R * (size_t)Tmax * D * 2));
  w.bytes = off;
  return w;
}

int check_ws(const SbDecoder* d, int N, int beam, int Tmax, void* workspace, size_t workspace_bytes, DecWs* out) {
  if (!d || !workspace) { set_last_error("sb_decoder: null argument"); return SB_ERR_INVALID; }
  if (N <= 0 || beam <= 0 || Tmax <= 0 || Tmax > d->cfg.pos_rows) {
    set_last_error("sb_decoder: bad N=%d beam=%d max_len=%d (position table has %d rows)", N, beam, Tmax, d->cfg.pos_rows);
    return SB_ERR_INVALID;
  }
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  *out = carve_dec(d, N, beam, Tmax, reinterpret_cast<void*>(base));
  if (base - reinterpret_cast<uintptr_t>(workspace) + out->bytes > workspace_bytes) {
    set_last_error("sb_decoder: workspace too small (%zu given, %zu needed)", workspace_bytes,
                   (size_t)(base - reinterpret_cast<uintptr_t>(workspace)) + out->bytes);
    return SB_ERR_INVALID;
  }
  if (gemm_topk_lists(out->n_chunks) > 256) { set_last_error("sb_decoder: vocabulary split into too many chunks"); return SB_ERR_INVALID; }
  return SB_OK;
}

}  // namespace

extern "C" {

int sb_decoder_create(const SbDecoderConfig* cfg, const SbDecoderWeights* w, SbDecoder** out) {
  if (!cfg || !w || !out) { set_last_error("sb_decoder_create: null argument"); return SB_ERR_INVALID; }
  *out = nullptr;
  const int D = cfg->model_dim, H = cfg->num_heads, F = cfg->ffn_inner_dim;
  if (D <= 0 || D % 256 != 0 || D > 1024 || H <= 0 || D != H * 64 || F <= 0 || F % 256 != 0) {
    set_last_error("sb_decoder_create: need model_dim %% 256 == 0 (<= 1024), head_dim 64, ffn %% 256 == 0");
    return SB_ERR_INVALID;
  }
  if (cfg->input_dim != D) {
    set_last_error("sb_decoder_create: input_dim (%d) must equal model_dim (%d)", cfg->input_dim, D);
    return SB_ERR_INVALID;
  }
  if (cfg->num_layers < 0 || cfg->pos_rows <= 0 || cfg->vocab_size <= 0 || cfg->eos_idx < 0 ||
      cfg->eos_idx >= cfg->vocab_size) {
    set_last_error("sb_decoder_create: bad num_layers / pos_rows / vocab_size / eos_idx");
    return SB_ERR_INVALID;
  }
  if (!w->embed || !w->pos_table || !w->final_ln_g || !w->final_ln_b || (cfg->num_layers > 0 && !w->layers)) {
    set_last_error("sb_decoder_create: missing weight pointer");
    return SB_ERR_INVALID;
  }
  int dev = 0, n_gpu = 0;
  if (cudaGetDeviceCount(&n_gpu) != cudaSuccess || n_gpu == 0) {
    set_last_error("sb_decoder_create: no CUDA device (this engine has no CPU path)");
    return SB_ERR_CUDA;
  }
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_last_error("sb_decoder_create: sm_100a kernels need a B200-class GPU (found sm_%d%d)", prop.major, prop.minor);
    return SB_ERR_CUDA;
  }
  SbDecoder* d = new (std::nothrow) SbDecoder();
  if (!d) { set_last_error("out of host memory"); return SB_ERR_INVALID; }
  d->cfg = *cfg;
  d->embed = w->embed;
  d->pos_table = w->pos_table;
  d->final_ln_g = w->final_ln_g;
  d->final_ln_b = w->final_ln_b;
  d->layers.assign(w->layers, w->layers + cfg->num_layers);
  for (int i = 0; i < cfg->num_layers; ++i) {
    const SbDecoderLayerWeights& l = d->layers[i];
    const void* ptrs[] = {l.wqkv, l.bqkv, l.wo, l.bo, l.cross_wv, l.cross_bv, l.cross_wo, l.cross_bo, l.w1, l.b1,
                          l.w2, l.b2, l.ln1_g, l.ln1_b, l.ln3_g, l.ln3_b};
    for (const void* q : ptrs)
      if (!q) {
        set_last_error("sb_decoder_create: layer %d has a null weight pointer", i);
        delete d;
        return SB_ERR_INVALID;
      }
  }
  d->num_sms = prop.multiProcessorCount;
  *out = d;
  return SB_OK;
}

void sb_decoder_destroy(SbDecoder* d) { delete d; }

int sb_decoder_workspace_bytes(const SbDecoder* d, int32_t num_sentences, int32_t beam, int32_t max_len, size_t* bytes) {
  if (!d || !bytes || num_sentences <= 0 || beam <= 0 || max_len <= 0) {
    set_last_error("sb_decoder_workspace_bytes: bad argument");
    return SB_ERR_INVALID;
  }
  *bytes = carve_dec(d, num_sentences, beam, max_len, nullptr).bytes + 1024;
  return SB_OK;
}

int sb_decoder_begin(SbDecoder* d, const float* embeddings, int32_t N, int32_t beam, int32_t max_len, void* workspace,
                     size_t workspace_bytes, void* stream_v) {
  DecWs w;
  int rc = check_ws(d, N, beam, max_len, workspace, workspace_bytes, &w);
  if (rc) return rc;
  if (!embeddings) { set_last_error("sb_decoder_begin: null embeddings"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int D = d->cfg.model_dim;
  SB_CUDA_CHECK(cudaMemsetAsync(w.err_flag, 0, sizeof(int32_t), stream));
  SB_CUDA_CHECK(cudaMemsetAsync(w.splitk_flags, 0, kSplitkFlags * sizeof(int), stream));
  const long long n = (long long)N * D;
  cast_bf16_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, stream>>>(embeddings, w.ebf, n);
  SB_CUDA_CHECK(cudaGetLastError());
  GemmArgs g;
  g.allow_skinny = 1;
  g.cta_group = 2;
  g.num_sms = d->num_sms;
  g.M = N;
  g.K = D;
  g.N = D;
  g.lda = D;
  g.ldw = D;
  g.ldc = D;
  g.residual = nullptr;
  g.ldr = 0;
  g.epi = EPI_BIAS;
  for (int li = 0; li < d->cfg.num_layers; ++li) {
    const SbDecoderLayerWeights& L = d->layers[li];
    // v = e Wv^T + bv  (bf16), c = v Wo^T + bo (fp32)
    g.A = w.ebf; g.W = reinterpret_cast<const __nv_bfloat16*>(L.cross_wv); g.C = w.vtmp; g.out_fp32 = 0; g.bias = L.cross_bv;
    if ((rc = gemm_bf16(g, stream))) return rc;
    g.A = w.vtmp; g.W = reinterpret_cast<const __nv_bfloat16*>(L.cross_wo); g.C = w.cross + (size_t)li * N * D; g.out_fp32 = 1;
    g.bias = L.cross_bo;
    if ((rc = gemm_bf16(g, stream))) return rc;
  }
  return SB_OK;
}

int sb_decoder_step(SbDecoder* d, const int64_t* tokens, const int32_t* table, int32_t t, int32_t N, int32_t beam,
                    int32_t max_len, float* out_lprob, int32_t* out_tok, float* out_eos_lprob,
                    const int64_t* probe_tokens, float* out_probe_lprob, void* workspace, size_t workspace_bytes,
                    void* stream_v) {
  DecWs w;
  int rc = check_ws(d, N, beam, max_len, workspace, workspace_bytes, &w);
  if (rc) return rc;
  if (!tokens || !table || !out_lprob || !out_tok || !out_eos_lprob) { set_last_error("sb_decoder_step: null pointer"); return SB_ERR_INVALID; }
  if ((probe_tokens != nullptr) != (out_probe_lprob != nullptr)) {
    set_last_error("sb_decoder_step: probe_tokens and out_probe_lprob go together");
    return SB_ERR_INVALID;
  }
  if (t < 0 || t >= max_len) { set_last_error("sb_decoder_step: position %d outside [0,%d)", t, max_len); return SB_ERR_INVALID; }
  if (max_len > kMaxDecodeLen) { set_last_error("sb_decoder_step: max_len %d > %d", max_len, kMaxDecodeLen); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int D = d->cfg.model_dim, F = d->cfg.ffn_inner_dim, H = d->cfg.num_heads;
  const int R = N * beam;
  if (R > 65535) { set_last_error("sb_decoder_step: too many rows (%d)", R); return SB_ERR_INVALID; }
  const unsigned row_blocks = (unsigned)((R + 7) / 8);
  decode_embed_kernel<<<row_blocks, 256, 0, stream>>>(tokens, reinterpret_cast<const __nv_bfloat16*>(d->embed),
                                                      d->cfg.vocab_size, d->pos_table + (size_t)t * D, D,
                                                      d->cfg.embed_scale, w.x, R, w.err_flag);
  SB_CUDA_CHECK(cudaGetLastError());
  GemmArgs g;
  g.allow_skinny = 1;
  g.cta_group = 2;
  g.num_sms = d->num_sms;
  g.M = R;
  g.splitk_flags = w.splitk_flags;  // the residual GEMMs may split K when their tiles leave SM pairs idle (2 560 rows)
  g.splitk_flags_len = kSplitkFlags;
  const size_t layer_stride = (size_t)R * max_len * D;
  for (int li = 0; li < d->cfg.num_layers; ++li) {
    const SbDecoderLayerWeights& L = d->layers[li];
    if ((rc = layernorm_bf16(w.x, L.ln1_g, L.ln1_b, d->cfg.ln_eps, w.h, R, D, stream))) return rc;
    g.A = w.h; g.lda = D; g.W = reinterpret_cast<const __nv_bfloat16*>(L.wqkv); g.ldw = D;
    g.C = w.qkv; g.ldc = 3 * D; g.out_fp32 = 0; g.bias = L.bqkv; g.residual = nullptr; g.ldr = 0;
    g.N = 3 * D; g.K = D; g.epi = EPI_BIAS;
    if ((rc = gemm_bf16(g, stream))) return rc;
    decode_attention_kernel<<<dim3((unsigned)((H + 3) / 4), (unsigned)R), 128, 0, stream>>>(
        w.qkv, w.kcache + li * layer_stride, w.vcache + li * layer_stride, table, t, max_len, H, w.h);
    SB_CUDA_CHECK(cudaGetLastError());
    g.A = w.h; g.lda = D; g.W = reinterpret_cast<const __nv_bfloat16*>(L.wo); g.ldw = D;
    g.C = w.x; g.ldc = D; g.out_fp32 = 1; g.bias = L.bo; g.residual = w.x; g.ldr = D;
    g.N = D; g.K = D; g.epi = EPI_BIAS_RESIDUAL;
    if ((rc = gemm_bf16(g, stream))) return rc;
    if (D % 128 == 0 && D <= 128 * kMaxVec) {
      add_const_layernorm_kernel<<<row_blocks, 256, 0, stream>>>(w.x, w.cross + (size_t)li * N * D, R, beam, D, L.ln3_g,
                                                                 L.ln3_b, d->cfg.ln_eps, w.h);
      SB_CUDA_CHECK(cudaGetLastError());
    } else {
      add_sentence_const_kernel<<<row_blocks, 256, 0, stream>>>(w.x, w.cross + (size_t)li * N * D, R, beam, D);
      SB_CUDA_CHECK(cudaGetLastError());
      if ((rc = layernorm_bf16(w.x, L.ln3_g, L.ln3_b, d->cfg.ln_eps, w.h, R, D, stream))) return rc;
    }
    g.A = w.h; g.lda = D; g.W = reinterpret_cast<const __nv_bfloat16*>(L.w1); g.ldw = D;
    g.C = w.f; g.ldc = F; g.out_fp32 = 0; g.bias = L.b1; g.residual = nullptr; g.ldr = 0;
    g.N = F; g.K = D; g.epi = EPI_BIAS_RELU;
    if ((rc = gemm_bf16(g, stream))) return rc;
    g.A = w.f; g.lda = F; g.W = reinterpret_cast<const __nv_bfloat16*>(L.w2); g.ldw = F;
    g.C = w.x; g.ldc = D; g.out_fp32 = 1; g.bias = L.b2; g.residual = w.x; g.ldr = D;
    g.N = D; g.K = F; g.epi = EPI_BIAS_RESIDUAL;
    if ((rc = gemm_bf16(g, stream))) return rc;
  }
  if ((rc = layernorm_bf16(w.x, d->final_ln_g, d->final_ln_b, d->cfg.ln_eps, w.h, R, D, stream))) return rc;
  if ((rc = gemm_bf16_topk(w.h, D, reinterpret_cast<const __nv_bfloat16*>(d->embed), D, R, (int)d->cfg.vocab_size, D,
                           w.cand_val, w.cand_idx, w.lse_part, w.n_chunks, 2, d->num_sms, stream)))
    return rc;
  vocab_merge_kernel<kTopkCandidates><<<row_blocks, 256, 0, stream>>>(
      w.cand_val, w.cand_idx, w.lse_part, gemm_topk_lists(w.n_chunks), w.h, reinterpret_cast<const __nv_bfloat16*>(d->embed), D,
      d->cfg.eos_idx, R, out_lprob, out_tok, out_eos_lprob, probe_tokens, (long long)d->cfg.vocab_size, out_probe_lprob);
  SB_CUDA_CHECK(cudaGetLastError());
  return SB_OK;
}

int sb_decoder_check_inputs(SbDecoder* d, void* workspace, void* stream_v) {
  if (!d || !workspace) { set_last_error("sb_decoder_check_inputs: null argument"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  int32_t flag = 0;
  SB_CUDA_CHECK(cudaMemcpyAsync(&flag, reinterpret_cast<void*>(base), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  SB_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (flag != 0) {
    set_last_error("token id outside [0, vocab_size) fed to the decoder");
    return SB_ERR_INPUT;
  }
  return SB_OK;
}

}  // extern "C"
