// Microbenchmark: does the packed FADD2 (sub.rn.f32x2, sm_100) raise the ceiling of the L1-distance term  acc += |a - e| ?
// Variant A: 16 scalar subs + 16 scalar |x| accumulates per k (k_rank_tile today, see fadd_peak.cu).
// Variant B: 8 packed subs (a pair of queries against one duplicated candidate value) + 16 scalar |x| accumulates.
// Variant C: packed sub, LOP3 abs on both halves, packed accumulate (8 + 16 + 8 issue slots, two pipes).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int V>
__global__ void __launch_bounds__(256) k(float* out, float a0, float e0, int iters) {
  float acc[16], a[4], e[4];
  u64 acc2[8];
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int i = 0; i < 8; ++i) acc2[i] = 0;
  for (int i = 0; i < 4; ++i) { a[i] = a0 + threadIdx.x * 1e-3f + i; e[i] = e0 + i * 0.5f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (V == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[4 * i + j] = __fadd_rn(acc[4 * i + j], fabsf(__fsub_rn(a[i], e[j])));
      } else {
        const u64 a01 = pack(a[0], a[1]), a23 = pack(a[2], a[3]);      // register pairs as an LDS.128 would deliver them
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const u64 ee = pack(e[j], e[j]);                             // duplicated candidate value (LDS.64 from a doubled tile)
          const u64 d0 = sub2(a01, ee), d1 = sub2(a23, ee);
          if (V == 1) {
            acc[j] = __fadd_rn(acc[j], fabsf(__uint_as_float((unsigned)d0)));
            acc[4 + j] = __fadd_rn(acc[4 + j], fabsf(__uint_as_float((unsigned)(d0 >> 32))));
            acc[8 + j] = __fadd_rn(acc[8 + j], fabsf(__uint_as_float((unsigned)d1)));
            acc[12 + j] = __fadd_rn(acc[12 + j], fabsf(__uint_as_float((unsigned)(d1 >> 32))));
          } else {
            acc2[j] = add2(acc2[j], d0 & 0x7fffffff7fffffffULL);
            acc2[4 + j] = add2(acc2[4 + j], d1 & 0x7fffffff7fffffffULL);
          }
        }
      }
      a[u & 3] = __fadd_rn(a[u & 3], 1e-7f);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, float* out, int sms, int khz) {
  const int iters = 4000;
  for (int ctas = 2; ctas <= 8; ctas *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<sms * ctas, 256>>>(out, 1.f, 2.f, 10);
    cudaEventRecord(e0);
    k<V><<<sms * ctas, 256>>>(out, 1.f, 2.f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double terms = (double)sms * ctas * 8 * iters * 8.0 * 16.0;       // warp-level |a-e| terms
    printf("%s ctas/SM %d: %.3f ms, %.3f warp-terms per clk per SMSP (at the %d MHz max clock)\n", name, ctas, ms,
           terms / (ms * 1e-3) / ((double)sms * 4) / (khz * 1e3), khz / 1000);
  }
}
int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  run<0>("A scalar sub + scalar |x| add  ", out, sms, khz);
  run<1>("B packed sub + scalar |x| add  ", out, sms, khz);
  run<2>("C packed sub + LOP3 + packed add", out, sms, khz);
  return 0;
}
