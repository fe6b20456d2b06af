// tc_numerics.cu -- what does tcgen05.mma (kind::f16, fp32 accumulate) do to the low bits?
//
// The exact mode of k_rank_tc (rank_tc_kernels.cu) needs a PROVEN bound of |accumulator - exact sum of the products|.  The PTX
// documentation does not state how the 16 products of one K=16 instruction and the incoming accumulator are aligned, summed and
// rounded, so this probe measures it: it runs ONE 128 x 256 tile with operands given bit for bit by a file and writes the
// fp32 accumulator tile back.  profiles/tc_numerics.py writes the experiments (known-answer families that separate the
// candidate hardware models) and reads the results.
//
//   tc_numerics <in.bin> <out.bin>
//   in.bin : int32 ktot (multiple of 16, <= 256), int32 fmt (0 = fp16, 1 = bf16), int32 nseg, int32 nrep,
//            uint16 A[128][ktot], uint16 B[256][ktot]   (row-major, K contiguous)
//   out.bin: float D[128][256]   with D[i][j] = sum_k A[i][k] * B[j][k], accumulated by ktot/16 MMAs; `nseg` > 1 splits the
//            K range into nseg runs and restarts nothing -- it only inserts a tcgen05.commit + wait between the runs (to see
//            whether a completed accumulator is re-read identically); `nrep` issues the whole K loop nrep times into the same
//            accumulator (D = nrep * sum: long accumulation runs without more shared memory).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/tc_numerics profiles/micro/tc_numerics.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}

// smem image: no-swizzle K-major core matrices, [k-chunk of 8][row][8 x 16 bit]; LBO = rows*16, SBO = 128
__global__ void __launch_bounds__(128, 1) k_probe(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, int ktot, int fmt, int nseg,
                                                 int nrep, int a_tmem, float* __restrict__ D) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  const int nkc = ktot / 8;
  unsigned char* smA = smem;
  unsigned char* smB = smem + (size_t)nkc * 128 * 16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smB + (size_t)nkc * 256 * 16);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int o = threadIdx.x; o < nkc * 128; o += blockDim.x) {
    const int row = o % 128, kc = o / 128;
    *reinterpret_cast<uint4*>(smA + (size_t)o * 16) = *reinterpret_cast<const uint4*>(A + (size_t)row * ktot + kc * 8);
  }
  for (int o = threadIdx.x; o < nkc * 256; o += blockDim.x) {
    const int row = o % 256, kc = o / 256;
    *reinterpret_cast<uint4*>(smB + (size_t)o * 16) = *reinterpret_cast<const uint4*>(B + (size_t)row * ktot + kc * 8);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores above -> visible to the MMA's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // a_tmem: the A operand lives in TMEM columns 256.. : thread t (= lane t = row t) stores its K elements, two 16-bit values per
  // 32-bit column, with tcgen05.st -- the layout the .ts form of tcgen05.mma is assumed to read (this probe checks the assumption)
  if (a_tmem) {
    const int warp_ = threadIdx.x >> 5;
    for (int c0 = 0; c0 < ktot / 2; c0 += 8) {
      uint32_t v[8];
      for (int e = 0; e < 8; ++e) {
        const int k = 2 * (c0 + e);
        v[e] = (uint32_t)A[(size_t)threadIdx.x * ktot + k] | ((uint32_t)A[(size_t)threadIdx.x * ktot + k + 1] << 16);
      }
      const uint32_t taddr = tmem + 256u + (uint32_t)c0 + ((uint32_t)(warp_ * 32) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                   "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // instruction descriptor: fp32 accumulate (bit 4), A/B format (bits 7.., 10..: 0 = fp16, 1 = bf16), N = 256, M = 128, K-major both
  const uint32_t idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const int nk = ktot / 16, nmma = nk * nrep, per = (nmma + nseg - 1) / nseg;
    uint32_t parity = 0;
    for (int jj = 0; jj < nmma; ++jj) {
      const int j = jj % nk;
      const uint64_t ad = desc(s32(smA) + (uint32_t)(2 * j) * 128 * 16, 128 * 16, 128);
      const uint64_t bd = desc(s32(smB) + (uint32_t)(2 * j) * 256 * 16, 256 * 16, 128);
      const uint32_t accumulate = jj > 0;
      if (a_tmem) {
        const uint32_t at = tmem + 256u + (uint32_t)(8 * j);      // 16 k = 8 columns per instruction
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(at),
                     "l"(bd), "r"(idesc), "r"(accumulate)
                     : "memory");
      } else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad),
                   "l"(bd), "r"(idesc), "r"(accumulate)
                   : "memory");
      if ((jj + 1) % per == 0 || jj + 1 == nmma) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
        mbar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 256; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 8; ++e) D[(size_t)(warp * 32 + lane) * 256 + c0 + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: tc_numerics in.bin out.bin\n"); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4) return 1;
  const int ktot = hdr[0], fmt = hdr[1], nseg = hdr[2] < 1 ? 1 : hdr[2], nrep = hdr[3] < 1 ? 1 : hdr[3];
  if (ktot <= 0 || ktot % 16 || ktot > 256) { fprintf(stderr, "ktot must be a multiple of 16, <= 256\n"); return 1; }
  std::vector<uint16_t> A((size_t)128 * ktot), B((size_t)256 * ktot);
  if (fread(A.data(), 2, A.size(), f) != A.size() || fread(B.data(), 2, B.size(), f) != B.size()) { fprintf(stderr, "short input\n"); return 1; }
  fclose(f);
  uint16_t *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const int smem = (ktot / 8) * (128 + 256) * 16 + 1024 + 64;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_probe<<<1, 128, smem>>>(dA, dB, ktot, fmt, nseg, nrep, getenv("TC_A_TMEM") ? 1 : 0, dD);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 256);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 1; }
  fwrite(D.data(), 4, D.size(), f);
  fclose(f);
  return 0;
}
