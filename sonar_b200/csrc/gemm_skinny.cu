// Skinny GEMM for a handful of rows (M <= 64): C[M,N] = epi(A[M,K] . W[N,K]^T + bias).
//
// The decoder's beam step at the pipelines' default batch (5 sentences x beam 5 = 25 hypothesis rows) and the speech
// pooler's single-query layers multiply a few activation rows with full weight matrices.  The 128 x 256 tcgen05 tiles of
// gemm_tcgen05.cu spend a whole tile's MMA time on mostly-zero rows and put only N/256 CTA pairs on the machine, so a
// [25 x 8192] . [8192 x 1024] product takes ~50 us.  This path is the opposite design point: the work is streaming W once
// from HBM, so every CTA owns 8 rows of W (one n8 tile -> N/8 CTAs cover the SMs), its 8 warps split K, each lane pulls
// 16 contiguous bytes of "its" W row and of the activation rows straight from global memory, and the products run on
// mma.sync m16n8k16 with the activations as the A operand.  The K order inside a 32-element chunk is permuted identically
// for both operands (lane t owns elements 8t..8t+7), which a dot product does not see and which makes every load a full
// 16-byte vector.  The 8 per-warp partial tiles are summed through shared memory in warp order (deterministic) and the
// epilogue (bias, ReLU / SiLU, in-place fp32 residual) is applied once.
#include "common.cuh"
#include "sonar_b200_internal.h"

namespace sb {
namespace {

constexpr int kSkinnyWarps = 8;
constexpr int kSkinnyThreads = kSkinnyWarps * 32;

__device__ __forceinline__ void mma16816_f32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MT = number of 16-row activation tiles (M <= 16 * MT)
template <int MT, typename OutT>
__global__ void __launch_bounds__(kSkinnyThreads)
gemm_skinny_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ W, long long ldw,
                   OutT* C, long long ldc, const float* __restrict__ bias, int M, int N, int K, int epi) {
  __shared__ float part[kSkinnyWarps][MT * 16][8 + 1];
  const int n0 = blockIdx.x * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kslice = K / kSkinnyWarps;  // multiple of 32 (host checks K % 256 == 0)
  const int kbeg = warp * kslice;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  const __nv_bfloat16* wrow = W + (long long)(n0 + g) * ldw + kbeg + 8 * t;
  const __nv_bfloat16* arow[MT][2];
  bool aok[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = i * 16 + hh * 8 + g;
      aok[i][hh] = r < M;
      arow[i][hh] = A + (long long)(aok[i][hh] ? r : 0) * lda + kbeg + 8 * t;
    }
  float acc[MT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  constexpr int kUnroll = 4;  // 32-element chunks in flight per lane
  for (int k = 0; k < kslice; k += 32 * kUnroll) {
    uint4 w4[kUnroll], a4[kUnroll][MT][2];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const bool in = k + 32 * u < kslice;
      w4[u] = in ? __ldg(reinterpret_cast<const uint4*>(wrow + k + 32 * u)) : zero4;
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          a4[u][i][hh] = (in && aok[i][hh]) ? __ldg(reinterpret_cast<const uint4*>(arow[i][hh] + k + 32 * u)) : zero4;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        mma16816_f32(acc[i], a4[u][i][0].x, a4[u][i][1].x, a4[u][i][0].y, a4[u][i][1].y, w4[u].x, w4[u].y);
        mma16816_f32(acc[i], a4[u][i][0].z, a4[u][i][1].z, a4[u][i][0].w, a4[u][i][1].w, w4[u].z, w4[u].w);
      }
  }
  // accumulator fragment: c0,c1 -> (row g, cols 2t,2t+1); c2,c3 -> (row g+8, cols 2t,2t+1)
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    part[warp][i * 16 + g][2 * t] = acc[i][0];
    part[warp][i * 16 + g][2 * t + 1] = acc[i][1];
    part[warp][i * 16 + g + 8][2 * t] = acc[i][2];
    part[warp][i * 16 + g + 8][2 * t + 1] = acc[i][3];
  }
  __syncthreads();
  for (int o = tid; o < MT * 16 * 8; o += kSkinnyThreads) {
    const int r = o >> 3, c = o & 7;
    if (r >= M) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kSkinnyWarps; ++w) v += part[w][r][c];  // fixed order: bitwise reproducible
    if (bias) v += bias[n0 + c];
    if (epi == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
    if (epi == EPI_BIAS_SILU) v = silu_fast(v);
    OutT* dst = C + (long long)r * ldc + n0 + c;
    if constexpr (sizeof(OutT) == 4) {
      if (epi == EPI_BIAS_RESIDUAL || epi == EPI_BIAS_ACCUM) v = *dst + v;  // same rounding as the tcgen05 path: fl(x + fl(acc + bias))
      *dst = v;
    } else {
      *dst = __float2bfloat16_rn(v);
    }
  }
}

template <int MT>
int launch_skinny(const GemmArgs& g, cudaStream_t stream) {
  const dim3 grid((unsigned)(g.N / 8));
  if (g.out_fp32)
    gemm_skinny_kernel<MT, float><<<grid, kSkinnyThreads, 0, stream>>>(g.A, g.lda, g.W, g.ldw, reinterpret_cast<float*>(g.C),
                                                                         g.ldc, g.bias, g.M, g.N, g.K, g.epi);
  else
    gemm_skinny_kernel<MT, __nv_bfloat16><<<grid, kSkinnyThreads, 0, stream>>>(
        g.A, g.lda, g.W, g.ldw, reinterpret_cast<__nv_bfloat16*>(g.C), g.ldc, g.bias, g.M, g.N, g.K, g.epi);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("gemm_skinny launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace

bool gemm_skinny_eligible(const GemmArgs& g) {
  if (g.M <= 0 || g.M > 64 || g.N % 8 != 0 || g.K % 256 != 0) return false;
  if (g.lda % 8 != 0 || g.ldw % 8 != 0) return false;  // 16-byte vector loads
  if ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.W)) & 15) return false;
  switch (g.epi) {
    case EPI_BIAS:
    case EPI_BIAS_RELU:
    case EPI_BIAS_SILU:
      return true;
    case EPI_BIAS_RESIDUAL:
    case EPI_BIAS_ACCUM:  // only the in-place fp32 form x += A.W^T + b
      return g.out_fp32 && (g.epi == EPI_BIAS_ACCUM || (g.residual == g.C && g.ldr == g.ldc));
    default:
      return false;
  }
}

int gemm_skinny(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 16) return launch_skinny<1>(g, stream);
  if (g.M <= 32) return launch_skinny<2>(g, stream);
  return launch_skinny<4>(g, stream);
}

}  // namespace sb
