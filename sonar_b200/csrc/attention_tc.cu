// tcgen05 self-attention for packed sequences of at most 128 tokens (the sentence regime of the SONAR
// text encoder: BASELINE config 2 is S = 128; longer inputs use attention.cu).
//
// One work item = (sentence b, head h): S = Q K^T and O = P V are each ONE accumulator tile:
//   S[128 x 128] = Q[128 x 64] . K[128 x 64]^T   4 x tcgen05.mma (M=128, N=128, K=16), both operands K-major
//   O[128 x 64]  = P[128 x 128] . V[128 x 64]    8 x tcgen05.mma (M=128, N=64,  K=16), V is the MN-major B operand
// Q/K/V tiles come straight out of the fused qkv activation [T, 3*D] with TMA (SWIZZLE_128B), accumulators
// live in TMEM (S: 128 columns, O: 64 columns), and the softmax runs with ONE THREAD PER QUERY ROW reading its
// row from TMEM (no shuffles, no shared-memory reductions); P is written back to shared memory as the bf16
// K-major A operand.  Keys >= len get probability exactly 0 (exp2(-inf)), so whatever the TMA box picked up
// beyond the sentence (the next sentence's rows, or zero fill past T) never contributes.
//
// CTA = 4 softmax/epilogue warps (TMEM lane quarter = warp index) + 1 control warp (TMA + MMA issue by one
// elected lane).  Persistent over items; 2 CTAs per SM (80 KB smem, 256 TMEM columns each) overlap one CTA's
// softmax with the other's loads and MMAs; within a CTA the next item's Q/K (and then V) are prefetched as
// soon as the MMAs that read them have retired.

#include "common.cuh"
#include "sonar_b200_internal.h"

#include <math_constants.h>

namespace sb {
namespace {

constexpr int kTile = 128 * 64 * 2;  // one [128 x 64] bf16 operand tile = 16 KB
constexpr int kSmemBytes = 3 * kTile + 2 * kTile + 256 + 1024;  // Q,K,V + P(2 tiles) + barriers + align slack
constexpr int kThreads = 160;

// MN-major B operand (V: rows = keys (K dim), 64 contiguous head dims = one 128 B swizzle row):
// 8-key groups are 1024 B apart (SBO); a single 64-wide atom along MN, LBO = 128 keys * 128 B.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t lo = ((smem_addr >> 4) & 0x3FFFu) | (uint64_t((128u * 128u) >> 4) << 16);
  uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int32_t* __restrict__ cu, int B, int H,
                    __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTile;
  uint8_t* sV = sK + kTile;
  uint8_t* sP = sV + kTile;  // 2 tiles: keys 0-63, keys 64-127
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTile);
  uint64_t* bar_qk = bars + 0;       // TMA: Q and K landed
  uint64_t* bar_v = bars + 1;        // TMA: V landed
  uint64_t* bar_s = bars + 2;        // MMA: S complete (Q, K smem free again)
  uint64_t* bar_p = bars + 3;        // softmax: P written (4 warp arrivals); also means S has been read
  uint64_t* bar_o = bars + 4;        // MMA: O complete (P, V smem free again)
  uint64_t* bar_drained = bars + 5;  // epilogue: O has been read out of TMEM (4 warp arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * 64;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      mbar_init(bar_drained, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_ptr_smem, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;        // columns [0,128)
  const uint32_t tmem_o = tmem_base + 128;  // columns [128,192)

  const int num_items = B * H;

  if (warp == 4) {
    // ============================ control warp: TMA + MMA issue ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(128, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32(128, 64) | (1u << 16);  // B operand MN-major
      auto load_qk = [&](int item) {
        const int b = item / H, h = item % H;
        const int tok0 = cu[b];
        mbar_arrive_expect_tx(bar_qk, 2 * kTile);
        tma_load_2d(sQ, &tm_qkv, bar_qk, h * 64, tok0);
        tma_load_2d(sK, &tm_qkv, bar_qk, D + h * 64, tok0);
      };
      auto load_v = [&](int item) {
        const int b = item / H, h = item % H;
        const int tok0 = cu[b];
        mbar_arrive_expect_tx(bar_v, kTile);
        tma_load_2d(sV, &tm_qkv, bar_v, 2 * D + h * 64, tok0);
      };
      int item = blockIdx.x;
      if (item < num_items) {
        load_qk(item);
        load_v(item);
      }
      uint32_t ph = 0;
      for (; item < num_items; item += gridDim.x, ph ^= 1) {
        const int next = item + gridDim.x;
        // ---- S = Q K^T ----
        mbar_wait(bar_qk, ph);
        tc_fence_after();
        {
          const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ));
          const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem_s, qd + uint64_t(2 * k), kd + uint64_t(2 * k), idesc_s, k != 0);
          umma_commit<1>(bar_s);
        }
        mbar_wait(bar_s, ph);  // Q, K consumed -> prefetch the next item's Q, K
        if (next < num_items) load_qk(next);
        // ---- O = P V ----
        mbar_wait(bar_p, ph);  // P in smem (and S fully read)
        mbar_wait(bar_v, ph);
        if (item != int(blockIdx.x)) mbar_wait(bar_drained, ph ^ 1);  // previous O read out of TMEM
        tc_fence_after();
        {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t pd = umma_desc_kmajor_sw128(smem_u32(sP + (k >> 2) * kTile)) + uint64_t(2 * (k & 3));
            const uint64_t vd = umma_desc_mnmajor_sw128(smem_u32(sV + k * 2048));  // 16 keys * 128 B per k-step
            umma_bf16<1>(tmem_o, pd, vd, idesc_o, k != 0);
          }
          umma_commit<1>(bar_o);
        }
        mbar_wait(bar_o, ph);  // P, V consumed -> prefetch the next item's V
        if (next < num_items) load_v(next);
      }
    }
    __syncwarp();
  } else {
    // ============================ softmax + epilogue: one thread per query row ============================
    const int row = warp * 32 + lane;
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ph ^= 1) {
      const int b = item / H, h = item % H;
      const int tok0 = cu[b];
      const int len = cu[b + 1] - tok0;  // 1..128 (host guarantees max_len <= 128)
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // the whole S row (128 fp32) lives in registers: one batch of TMEM loads, one wait
      const int nch = (len + 31) >> 5;  // 32-key chunks that hold valid keys (CTA-uniform)
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_s + lane_base + c * 32, v[c]);  // stale columns are masked below
      tmem_ld_wait();
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 + 32 > len) {  // chunk reaches past the sentence: keys >= len -> -inf -> probability exactly 0
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j >= len) v[c][j] = __float_as_uint(-CUDART_INF_F);
        }
        if (c < nch) {
          float m0 = __uint_as_float(v[c][0]), m1 = __uint_as_float(v[c][1]);
#pragma unroll
          for (int j = 2; j < 32; j += 2) {
            m0 = fmaxf(m0, __uint_as_float(v[c][j]));
            m1 = fmaxf(m1, __uint_as_float(v[c][j + 1]));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        }
      }
      const float mxs = mx * sl2;
      // p = exp2(s*c - max*c), row sum, bf16 P row -> swizzled K-major smem (zeros beyond the valid chunks)
      float sum0 = 0.f, sum1 = 0.f;
      uint8_t* prow = sP + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float p[32];
        if (c < nch) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            p[j] = fast_exp2(fmaf(__uint_as_float(v[c][j]), sl2, -mxs));
            p[j + 1] = fast_exp2(fmaf(__uint_as_float(v[c][j + 1]), sl2, -mxs));
            sum0 += p[j];
            sum1 += p[j + 1];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;  // 16-byte chunk inside the 64-key tile
          *reinterpret_cast<uint4*>(prow + (c >> 1) * kTile + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2(p[8 * q], p[8 * q + 1]), pack_bf16x2(p[8 * q + 2], p[8 * q + 3]),
                         pack_bf16x2(p[8 * q + 4], p[8 * q + 5]), pack_bf16x2(p[8 * q + 6], p[8 * q + 7]));
        }
      }
      const float sum = sum0 + sum1;
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // ---- epilogue: O / sum -> bf16 -> global ----
      const float inv = 1.0f / sum;
      mbar_wait(bar_o, ph);
      tc_fence_after();
      uint32_t o[2][32];
      tmem_ld_32x32(tmem_o + lane_base, o[0]);
      tmem_ld_32x32(tmem_o + lane_base + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_drained);
      if (row < len) {
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)(tok0 + row) * D + h * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t* s = &o[q >> 2][(q & 3) * 8];
          dst[q] = make_uint4(pack_bf16x2(__uint_as_float(s[0]) * inv, __uint_as_float(s[1]) * inv),
                              pack_bf16x2(__uint_as_float(s[2]) * inv, __uint_as_float(s[3]) * inv),
                              pack_bf16x2(__uint_as_float(s[4]) * inv, __uint_as_float(s[5]) * inv),
                              pack_bf16x2(__uint_as_float(s[6]) * inv, __uint_as_float(s[7]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(tmem_base, 256);
}

}  // namespace

int attention_packed_tc(const __nv_bfloat16* qkv, const int32_t* cu_seqlens, int B, int H, long long total_tokens,
                        __nv_bfloat16* out, int num_sms, cudaStream_t stream) {
  if (B <= 0 || total_tokens <= 0) return 0;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, 2, total_tokens, 3ll * H * 64, 3ll * H * 64, 128, 64);
  if (rc) return rc;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    SB_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  long long items = (long long)B * H;
  long long grid = 2ll * (num_sms > 0 ? num_sms : 148);
  if (grid > items) grid = items;
  attention_tc_kernel<<<(unsigned)grid, kThreads, kSmemBytes, stream>>>(tm, cu_seqlens, B, H, out);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sb
