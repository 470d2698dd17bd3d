// 80-bin Kaldi log-mel filterbank + per-utterance standardisation for the SONAR speech path
// (BASELINE.json config 3).
//
// Replaces fairseq2n WaveformToFbankConverter(num_mel_bins=80, waveform_scale=2**15, channel_last=True,
// standardize=True) + Collater(pad_value=0, pad_to_multiple=2) as configured at
// sonar/inference_pipelines/speech.py:120-127,139,283-290,444 (= kaldi-native-fbank defaults: 25 ms / 10 ms frames,
// snip_edges, DC removal, pre-emphasis 0.97, Povey window, 512-point FFT, power spectrum, Kaldi mel 20 Hz..8 kHz,
// log floor FLT_EPSILON, no dither; standardise per bin over time with the unbiased std).
//
// fbank_kernel: ONE WARP PER FRAME, everything in shared memory: 400 samples in (coalesced), DC removal by a warp
// reduction, pre-emphasis + window, a 512-point real FFT done as a 256-point complex radix-2 FFT + split step,
// 80 triangular mel sums in a fixed order (deterministic), log.  HBM traffic = 1.6 KB in (overlapping frames hit
// L2) + 320 B out per frame; the kernel is latency/launch bound, not bandwidth bound.

#include "../../include/sonar_b200.h"
#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>

namespace sb {
namespace {

constexpr int kFrameLen = 400, kFrameShift = 160, kNfft = 512, kMel = 80, kBins = 257;
constexpr int kWarpsPerCta = 8;
constexpr float kWaveScale = 32768.0f, kPreemph = 0.97f, kLogFloor = 1.1920928955078125e-07f;

struct FbankTables {        // built once on the host (double precision -> fp32), lives in device memory
  float window[kFrameLen];  // Povey
  float tw256_cos[128], tw256_sin[128];  // exp(-2 pi i j / 256)
  float tw512_cos[256], tw512_sin[256];  // exp(-2 pi i k / 512)
  int mel_lo[kMel], mel_len[kMel];       // first FFT bin / number of bins with non-zero weight
  float mel_w[kMel][64];                 // weights, padded
};

__global__ void __launch_bounds__(kWarpsPerCta * 32)
fbank_kernel(const float* __restrict__ waves, const long long* __restrict__ wave_off, const int* __restrict__ frame_off,
             int B, const FbankTables* __restrict__ tab, float* __restrict__ out) {
  __shared__ float s_win[kFrameLen];
  __shared__ float s_c256[128], s_s256[128], s_c512[256], s_s512[256];
  __shared__ float s_re[kWarpsPerCta][kNfft];  // time samples, later power spectrum
  __shared__ float2 s_z[kWarpsPerCta][256];    // complex FFT buffer
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kFrameLen; i += blockDim.x) s_win[i] = tab->window[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { s_c256[i] = tab->tw256_cos[i]; s_s256[i] = tab->tw256_sin[i]; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_c512[i] = tab->tw512_cos[i]; s_s512[i] = tab->tw512_sin[i]; }
  __syncthreads();

  const int total_frames = frame_off[B];
  const int frame = blockIdx.x * kWarpsPerCta + warp;  // global (packed) frame index
  if (frame >= total_frames) return;
  // utterance of this frame: binary search in frame_off (B+1 ascending entries)
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (frame_off[mid] <= frame) lo = mid; else hi = mid;
  }
  const int b = lo;
  const float* src = waves + wave_off[b] + (long long)(frame - frame_off[b]) * kFrameShift;
  float* re = s_re[warp];
  float2* z = s_z[warp];

  // ---- load, scale, remove DC ----
  float v[13];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const int n = i * 32 + lane;
    v[i] = (n < kFrameLen) ? src[n] * kWaveScale : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / float(kFrameLen);
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const int n = i * 32 + lane;
    if (n < kFrameLen) re[n] = v[i] - mean;
  }
  __syncwarp();
  // ---- pre-emphasis + window; zero-pad to 512 ----
  float y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int n = i * 32 + lane;
    y[i] = (n < kFrameLen) ? (re[n] - kPreemph * re[n > 0 ? n - 1 : 0]) * s_win[n] : 0.f;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 16; ++i) re[i * 32 + lane] = y[i];
  __syncwarp();
  // ---- 256-point complex FFT of z[m] = y[2m] + i y[2m+1] (radix-2 DIT, bit-reversed load) ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = i * 32 + lane;
    z[__brev((unsigned)m) >> 24] = make_float2(re[2 * m], re[2 * m + 1]);
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int bf = j * 32 + lane;
      const int pos = bf & (half - 1);
      const int i0 = ((bf >> s) << (s + 1)) + pos, i1 = i0 + half;
      const int tw = pos << (7 - s);  // pos * 128 / half
      const float c = s_c256[tw], sn = s_s256[tw];  // w = c - i sn
      const float2 a = z[i0], bq = z[i1];
      const float tr = bq.x * c + bq.y * sn, ti = bq.y * c - bq.x * sn;
      z[i0] = make_float2(a.x + tr, a.y + ti);
      z[i1] = make_float2(a.x - tr, a.y - ti);
    }
    __syncwarp();
  }
  // ---- split step: X[k] = (Z[k] + conj Z[256-k])/2 - i e^{-2 pi i k/512} (Z[k] - conj Z[256-k])/2 ; power ----
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int k = i * 32 + lane;
    if (k <= 256) {
      const float2 a = z[k & 255], bq = z[(256 - k) & 255];
      const float er = 0.5f * (a.x + bq.x), ei = 0.5f * (a.y - bq.y);  // even part
      const float orr = 0.5f * (a.x - bq.x), oi = 0.5f * (a.y + bq.y);  // (Z[k] - conj Z[N-k]) / 2
      float c, sn;
      if (k < 256) { c = s_c512[k]; sn = s_s512[k]; } else { c = -1.f; sn = 0.f; }
      // -i * (c - i sn) * (orr + i oi) = -i * [(c orr + sn oi) + i (c oi - sn orr)] = (c oi - sn orr) - i (c orr + sn oi)
      const float xr = er + (c * oi - sn * orr), xi = ei - (c * orr + sn * oi);
      re[k] = xr * xr + xi * xi;
    }
  }
  __syncwarp();
  // ---- 80 mel bins (fixed summation order), log ----
  float* orow = out + (long long)frame * kMel;
  for (int m = lane; m < kMel; m += 32) {
    const int k0 = tab->mel_lo[m], len = tab->mel_len[m];
    float acc = 0.f;
    for (int q = 0; q < len; ++q) acc = fmaf(tab->mel_w[m][q], re[k0 + q], acc);
    orow[m] = logf(fmaxf(acc, kLogFloor));
  }
}

// one CTA per utterance: per-bin mean / unbiased std over its frames, write (f - mean)/std into the padded batch
__global__ void __launch_bounds__(320)
standardize_kernel(const float* __restrict__ raw, const int* __restrict__ frame_off, float* __restrict__ out, int Tpad) {
  __shared__ double s_part[4][kMel];
  __shared__ float s_mean[kMel], s_inv[kMel];
  const int b = blockIdx.x;
  const int g = threadIdx.x / kMel, bin = threadIdx.x % kMel;
  const int f0 = frame_off[b], T = frame_off[b + 1] - f0;
  const float* src = raw + (long long)f0 * kMel;
  double acc = 0.0;
  for (int t = g; t < T; t += 4) acc += (double)src[(long long)t * kMel + bin];
  s_part[g][bin] = acc;
  __syncthreads();
  if (g == 0) s_mean[bin] = (float)((s_part[0][bin] + s_part[1][bin] + s_part[2][bin] + s_part[3][bin]) / (double)T);
  __syncthreads();
  const float mean = s_mean[bin];
  acc = 0.0;
  for (int t = g; t < T; t += 4) {
    const double d = (double)src[(long long)t * kMel + bin] - (double)mean;
    acc += d * d;
  }
  s_part[g][bin] = acc;
  __syncthreads();
  if (g == 0) {
    const double var = (s_part[0][bin] + s_part[1][bin] + s_part[2][bin] + s_part[3][bin]) / (double)(T - 1);
    s_inv[bin] = (float)(1.0 / sqrt(var));  // T == 1 -> NaN, like torch.std_mean
  }
  __syncthreads();
  const float inv = s_inv[bin];
  float* dst = out + (long long)b * Tpad * kMel;
  for (int t = g; t < Tpad; t += 4)
    dst[(long long)t * kMel + bin] = (t < T) ? (src[(long long)t * kMel + bin] - mean) * inv : 0.f;
}

}  // namespace
}  // namespace sb

using namespace sb;

extern "C" {

size_t sb_fbank_tables_bytes(void) { return sizeof(FbankTables); }

// Fills a HOST buffer of sb_fbank_tables_bytes() with the window / twiddle / mel tables (the caller uploads it once).
int sb_fbank_build_tables(void* host_buf) {
  if (!host_buf) { set_last_error("sb_fbank_build_tables: null buffer"); return SB_ERR_INVALID; }
  FbankTables* t = reinterpret_cast<FbankTables*>(host_buf);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < kFrameLen; ++n)
    t->window[n] = (float)pow(0.5 - 0.5 * cos(2.0 * pi * n / (kFrameLen - 1)), 0.85);
  for (int j = 0; j < 128; ++j) { t->tw256_cos[j] = (float)cos(2.0 * pi * j / 256.0); t->tw256_sin[j] = (float)sin(2.0 * pi * j / 256.0); }
  for (int k = 0; k < 256; ++k) { t->tw512_cos[k] = (float)cos(2.0 * pi * k / 512.0); t->tw512_sin[k] = (float)sin(2.0 * pi * k / 512.0); }
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double lo = mel(20.0), hi = mel(8000.0), delta = (hi - lo) / (kMel + 1);
  for (int m = 0; m < kMel; ++m) {
    const double left = lo + m * delta, center = lo + (m + 1) * delta, right = lo + (m + 2) * delta;
    int first = -1, last = -1;
    double w[kBins];
    for (int k = 0; k < kBins; ++k) {
      double v = 0.0;
      if (k < kNfft / 2) {  // the Nyquist bin carries weight 0 (Kaldi)
        const double mk = mel(k * (16000.0 / kNfft));
        const double up = (mk - left) / (center - left), down = (right - mk) / (right - center);
        v = up < down ? up : down;
        if (v < 0.0) v = 0.0;
      }
      w[k] = v;
      if (v > 0.0) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    const int len = last - first + 1;
    if (len > 64) { set_last_error("sb_fbank_build_tables: mel filter wider than 64 bins"); return SB_ERR_INVALID; }
    t->mel_lo[m] = first;
    t->mel_len[m] = len;
    for (int q = 0; q < 64; ++q) t->mel_w[m][q] = (q < len) ? (float)w[first + q] : 0.f;
  }
  return SB_OK;
}

int sb_fbank(const float* waves, const int64_t* wave_offsets, const int32_t* frame_offsets, int32_t B,
             int32_t total_frames, const void* tables, float* raw_out, float* out, int32_t padded_frames, void* stream_v) {
  if (!waves || !wave_offsets || !frame_offsets || !tables || !raw_out || !out) {
    set_last_error("sb_fbank: null pointer");
    return SB_ERR_INVALID;
  }
  if (B <= 0 || total_frames <= 0 || padded_frames <= 0) { set_last_error("sb_fbank: empty input"); return SB_ERR_INVALID; }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  fbank_kernel<<<(unsigned)((total_frames + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, 0, stream>>>(
      waves, reinterpret_cast<const long long*>(wave_offsets), frame_offsets, B,
      reinterpret_cast<const FbankTables*>(tables), raw_out);
  SB_CUDA_CHECK(cudaGetLastError());
  standardize_kernel<<<(unsigned)B, 320, 0, stream>>>(raw_out, frame_offsets, out, padded_frames);
  SB_CUDA_CHECK(cudaGetLastError());
  return SB_OK;
}

}  // extern "C"
