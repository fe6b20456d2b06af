// Microbenchmark: issue ceiling of the two-FADD L1-distance term  acc = acc + |a - e|  (the inner loop of k_rank_tile),
// operands in registers, 16 independent accumulators per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float* out, float a0, float e0, int iters) {
  float acc[16], a[4], e[4];
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int i = 0; i < 4; ++i) { a[i] = a0 + threadIdx.x * 1e-3f + i; e[i] = e0 + i * 0.5f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[4 * i + j] = __fadd_rn(acc[4 * i + j], fabsf(__fsub_rn(a[i], e[j])));
      // perturb operands so nothing is hoisted (4 extra instructions per 32 FADD)
      a[u & 3] = __fadd_rn(a[u & 3], 1e-7f);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  const int iters = 4000;
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms * ctas, 256>>>(out, 1.f, 2.f, 10);
    cudaEventRecord(e0);
    k<<<sms * ctas, 256>>>(out, 1.f, 2.f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fadd_warp = (double)sms * ctas * 8 /*warps*/ * iters * 8.0 * 32.0;   // 2 FADD x 16 pairs per u
    const double per_clk_smsp = fadd_warp / (ms * 1e-3) / ((double)sms * 4) / (khz * 1e3);
    printf("ctas/SM %d: %.3f ms, %.3f G warp-FADD/s, %.3f warp-FADD per clk per SMSP (at the %d MHz max clock)\n", ctas, ms,
           fadd_warp / ms * 1e-6, per_clk_smsp, khz / 1000);
  }
  return 0;
}
