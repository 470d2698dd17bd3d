// Microbenchmark: peak rate of legacy warp-level mma.sync.m16n8k16 bf16 on sm_100a (what the flash kernels that still use
// it can hope for).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peak mma_sync_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
  if (s == 123.456f) out[0] = s;
}
int main() {
  float* d;
  cudaMalloc(&d, 4);
  for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
    const int iters = 20000, grid = 148 * ctas_per_sm;
    k<<<grid, 256>>>(d, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)grid * 8 /*warps*/ * iters * 8 * 4096.0;
    printf("mma.sync m16n8k16 bf16: %d CTAs/SM x 8 warps: %.1f TFLOP/s (%.3f ms)\n", ctas_per_sm, flop / ms / 1e9, ms);
  }
  return 0;
}
