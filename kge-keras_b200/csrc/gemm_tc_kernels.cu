// gemm_tc_kernels.cu -- the Dense of the literal branch (EmbeddingModel.call, models.py:114-120,152-158) on the tensor cores.
//
// The three GEMMs of that step -- Y = act(X W + b), gW += X^T gZ, gX = gZ W^T, all with one dimension ~100 and fp32 parity at the
// 1e-5 bar -- ran as a CUDA-core FFMA kernel (literal_kernels.cu, 28 TFLOP/s).  Here they are tcgen05.mma kind::tf32 with a
// three-term split: every fp32 operand x is split on the way into shared memory into hi = x rounded to tf32 (11 significant bits)
// and lo = x - hi (exact in fp32; the MMA reads its top 11 bits), and the accumulator receives a_hi.b_hi + a_hi.b_lo + a_lo.b_hi.
// The dropped a_lo.b_lo and the rounding of the lo parts leave less than 1.3 * 2^-21 = 6e-7 of sum|a_k b_k| -- inside the 1e-5
// parity bar by a factor of 16, which plain tf32 (2^-11) is not.
//
// One CTA per 128 x BN output tile (or per K slice of one, when the output is a single tile and K is the batch: the weight
// gradient).  Warps 0-3 load: a thread owns one row of A and one of B per stage, reads 16 consecutive k (64 contiguous bytes, or 16
// coalesced scalars for a transposed operand: transposition is an addressing mode of the loader, no copies are made), splits, and
// stores 16-byte chunks straight into the no-swizzle K-major core-matrix layout the UMMA descriptors read -- [k-chunk][plane][row][4
// floats] -- then fences the generic-proxy stores towards the async proxy and arrives on the stage's mbarrier.  Warp 4 (one thread)
// issues 2 K=8 steps x 3 products per stage into a TMEM accumulator and commits the stage back.  After the last stage warps 0-3
// read the accumulator (tcgen05.ld, a thread per output row), add the bias, apply the activation and store, or write the raw
// partial of their K slice for the deterministic slice-order reduction (k_splitk_reduce, literal_kernels.cu).
#include "common.cuh"

namespace {

constexpr int GT_BM = 128;      // output rows per CTA (UMMA M: one TMEM lane per row)
constexpr int GT_BN = 128;      // output columns per CTA (UMMA N, padded to a multiple of 16)
constexpr int GT_BK = 16;       // k per stage: 4 chunks of 4 floats, 2 MMAs of K = 8 per product
constexpr int GT_STAGES = 3;
constexpr int GT_THREADS = 160; // warps 0-3: loaders, then epilogue; warp 4: TMEM allocation + MMA issue
constexpr uint32_t GT_A_STAGE = GT_BM * GT_BK * 4 * 2;   // 16 KB (two planes)
constexpr uint32_t GT_B_STAGE = GT_BN * GT_BK * 4 * 2;   // 16 KB
constexpr uint32_t GT_SMEM = GT_STAGES * (GT_A_STAGE + GT_B_STAGE) + 1024 + 128;

__device__ __forceinline__ uint32_t gt_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gt_s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gt_s32(bar)) : "memory");
}
__device__ __forceinline__ void gt_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tGT_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra GT_DONE;\n\tbra GT_WAIT;\n\tGT_DONE:\n\t}" ::"r"(
          gt_s32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t gt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void gt_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void gt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gt_s32(bar)) : "memory");
}

// x -> (hi, lo): hi = x rounded to the nearest tf32 (low 13 mantissa bits cleared after adding half), lo = (x - hi) rounded the same
// way (x - hi is exact in fp32; rounding it here, instead of letting the MMA truncate it, halves that term): |x - hi - lo| <= 2^-22 |x|
__device__ __forceinline__ float gt_tf32_rn(float x) {
  // round to nearest (ties away) by adding half of the dropped field; finite inputs within 2^-11 of FLT_MAX would carry into the
  // exponent field's top -- gradients and activations never are
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void gt_split(float x, float& hi, float& lo) {
  hi = gt_tf32_rn(x);
  lo = gt_tf32_rn(__fsub_rn(x, hi));
}

struct GtArgs {
  const float* A; const float* B; float* C; float* part; const float* bias;
  int64_t M; int N, K, lda, ldb, ldc, transA, transB, accumulate, act, kchunk, splits;
};

__device__ __forceinline__ float gt_act(float v, int act) {
  if (act == KGEB_ACT_TANH) return tanhf(v);
  if (act == KGEB_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == KGEB_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// One operand tile of one stage (128 rows x 16 k) -> hi / lo planes of the canonical layout; warp w owns rows 32w .. 32w+31.
//   transposed operand (element (r, k) at P[k * ld + r]): lane = row, 16 coalesced 128-byte loads, 16-byte vector stores;
//   row-major operand  (element (r, k) at P[r * ld + k]): lane = (k, row parity) with the rows rotated by 2 per 16-byte k-chunk, so
//   that a load instruction reads whole 16-byte pieces of eight rows and the 32 scalar stores of an instruction hit 32 different
//   banks (rows of one k-chunk are 16 bytes apart, the four k-chunks 4096 bytes: without the rotation they would collide 4 ways).
// phase 1: the 16 values a thread contributes to the stage, global -> registers (no dependent use: the loads of the NEXT stage are
// issued before the current stage's values are split and stored, so a stage's memory latency hides behind its predecessor)
__device__ __forceinline__ void gt_fetch(float (&x)[16], const float* __restrict__ P, int64_t row0, int64_t n_rows, int k0, int kend, int ld, int trans,
                                         int warp, int lane) {
  if (trans) {
    const int64_t r = row0 + 32 * warp + lane;
#pragma unroll
    for (int k = 0; k < GT_BK; ++k) x[k] = (r < n_rows && k0 + k < kend) ? __ldg(P + (int64_t)(k0 + k) * ld + r) : 0.f;
  } else {
    const int kk = lane & 15, c = kk >> 2, h = lane >> 4;
    const bool kok = k0 + kk < kend;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t r = row0 + 32 * warp + ((2 * i + h + 2 * c) & 31);
      x[i] = (kok && r < n_rows) ? __ldg(P + r * (int64_t)ld + k0 + kk) : 0.f;
    }
  }
}
// phase 2: split and store into the stage
template <int ROWS>
__device__ __forceinline__ void gt_store(const float (&x)[16], int trans, unsigned char* stage, int warp, int lane) {
  if (trans) {
    const int rt = 32 * warp + lane;
#pragma unroll
    for (int c = 0; c < GT_BK / 4; ++c) {
      float4 h, l;
      gt_split(x[4 * c], h.x, l.x); gt_split(x[4 * c + 1], h.y, l.y); gt_split(x[4 * c + 2], h.z, l.z); gt_split(x[4 * c + 3], h.w, l.w);
      *reinterpret_cast<float4*>(stage + ((size_t)(c * 2 + 0) * ROWS + rt) * 16) = h;
      *reinterpret_cast<float4*>(stage + ((size_t)(c * 2 + 1) * ROWS + rt) * 16) = l;
    }
  } else {
    const int kk = lane & 15, c = kk >> 2, j = kk & 3, h = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rt = 32 * warp + ((2 * i + h + 2 * c) & 31);
      float hi, lo;
      gt_split(x[i], hi, lo);
      *reinterpret_cast<float*>(stage + ((size_t)(c * 2 + 0) * ROWS + rt) * 16 + j * 4) = hi;
      *reinterpret_cast<float*>(stage + ((size_t)(c * 2 + 1) * ROWS + rt) * 16 + j * 4) = lo;
    }
  }
}

__global__ void __launch_bounds__(GT_THREADS, 2) k_gemm_tc(GtArgs a) {
  extern __shared__ unsigned char gt_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gt_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smA = smem;
  unsigned char* smB = smem + GT_STAGES * GT_A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_STAGES * (GT_A_STAGE + GT_B_STAGE));
  uint64_t* empty = full + GT_STAGES;
  uint64_t* acc_full = empty + GT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * GT_BM;
  const int n0 = blockIdx.x * GT_BN;
  const int kbeg = blockIdx.z * a.kchunk;
  const int kend = kbeg + a.kchunk < a.K ? kbeg + a.kchunk : a.K;
  const int nstage = (kend - kbeg + GT_BK - 1) / GT_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { gt_mbar_init(full + i, 128); gt_mbar_init(empty + i, 1); }
    gt_mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_s32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===== loaders: thread t owns row m0 + t of A and row n0 + t of B^T =====
    const int t = threadIdx.x;
    uint32_t stage = 0, phase = 0;
    float xa[16], xb[16], ya[16], yb[16];
    gt_fetch(xa, a.A, m0, a.M, kbeg, kend, a.lda, a.transA, warp, lane);
    gt_fetch(xb, a.B, (int64_t)n0, (int64_t)a.N, kbeg, kend, a.ldb, !a.transB, warp, lane);
    for (int s = 0; s < nstage; ++s) {
      if (s + 1 < nstage) {                       // the next stage's loads leave before this stage's values are consumed
        const int k1 = kbeg + (s + 1) * GT_BK;
        gt_fetch(ya, a.A, m0, a.M, k1, kend, a.lda, a.transA, warp, lane);
        gt_fetch(yb, a.B, (int64_t)n0, (int64_t)a.N, k1, kend, a.ldb, !a.transB, warp, lane);
      }
      gt_mbar_wait(empty + stage, phase ^ 1u);
      gt_store<GT_BM>(xa, a.transA, smA + stage * GT_A_STAGE, warp, lane);
      gt_store<GT_BN>(xb, !a.transB, smB + stage * GT_B_STAGE, warp, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA's async proxy
      gt_mbar_arrive(full + stage);
      if (++stage == GT_STAGES) { stage = 0; phase ^= 1u; }
#pragma unroll
      for (int i = 0; i < 16; ++i) { xa[i] = ya[i]; xb[i] = yb[i]; }
    }
    // ===== epilogue: thread t = output row m0 + t (TMEM lane t) =====
    gt_mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t row = m0 + t;
    const int ncols = a.N - n0 < GT_BN ? a.N - n0 : GT_BN;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
            "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < a.M) {
        float* dst = a.splits > 1 ? a.part + (size_t)blockIdx.z * (size_t)(a.M * a.N) + (size_t)row * a.N + n0 + c0
                                  : a.C + row * (int64_t)a.ldc + n0 + c0;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = __uint_as_float(r[j]);
          if (a.splits == 1 && a.bias && c0 + j < ncols) v[j] = gt_act(v[j] + __ldg(a.bias + n0 + c0 + j), a.act);
        }
        const bool vec = c0 + 16 <= ncols && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
        if (vec) {
          float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (a.splits == 1 && a.accumulate) { const float4 p = d4[q]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            d4[q] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < ncols) dst[j] = (a.splits == 1 && a.accumulate) ? dst[j] + v[j] : v[j];
        }
      }
    }
  } else if (lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor: fp32 accumulate (bit 4), A and B tf32 (format 2 at bits 7 and 10), both K-major, N at 17, M at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GT_BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
    const uint32_t lboA = 2 * GT_BM * 16, lboB = 2 * GT_BN * 16;   // distance between the two k-chunks of one K = 8 step (planes interleave)
    uint32_t stage = 0, phase = 0, accumulate = 0;
    for (int s = 0; s < nstage; ++s) {
      gt_mbar_wait(full + stage, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = gt_s32(smA + stage * GT_A_STAGE), sb = gt_s32(smB + stage * GT_B_STAGE);
#pragma unroll
      for (int j = 0; j < GT_BK / 8; ++j) {
        const uint32_t a_hi = sa + (uint32_t)(2 * j * 2) * GT_BM * 16, a_lo = a_hi + GT_BM * 16;
        const uint32_t b_hi = sb + (uint32_t)(2 * j * 2) * GT_BN * 16, b_lo = b_hi + GT_BN * 16;
        gt_mma_tf32(tmem_base, gt_desc(a_hi, lboA, 128), gt_desc(b_hi, lboB, 128), idesc, accumulate);
        accumulate = 1;
        gt_mma_tf32(tmem_base, gt_desc(a_hi, lboA, 128), gt_desc(b_lo, lboB, 128), idesc, 1);
        gt_mma_tf32(tmem_base, gt_desc(a_lo, lboA, 128), gt_desc(b_hi, lboB, 128), idesc, 1);
      }
      gt_commit(empty + stage);
      if (++stage == GT_STAGES) { stage = 0; phase ^= 1u; }
    }
    gt_commit(acc_full);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

}  // namespace

// C[M,N] (+)= op(A) op(B) (+ bias, activation) on the tensor cores; KGEB_ERR_UNSUPPORTED = shape not served (the caller falls back
// to the CUDA-core kernel).  splits_out / part: K is sliced when the output is less than one wave of tiles; the caller reduces the
// slices in slice order (k_splitk_reduce).
int kgeb_internal_gemm_tc(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
                          float* C, int ldc, int accumulate, const float* bias, int act, float* part, size_t part_bytes, int* splits_out,
                          cudaStream_t st) {
  *splits_out = 1;
  if (K < 8 || N < 8) return KGEB_ERR_UNSUPPORTED;          // nothing to gain on slivers
  const int64_t mt = (M + GT_BM - 1) / GT_BM;
  const int nt = (N + GT_BN - 1) / GT_BN;
  if (mt > 65535 || nt > 65535) return KGEB_ERR_UNSUPPORTED;
  const int64_t tiles = mt * nt;
  int splits = 1, kchunk = (K + GT_BK - 1) / GT_BK * GT_BK;
  if (!bias && tiles < ctx->num_sms && K >= 1024) {          // the weight gradient: one output tile, K = the batch
    int64_t want = (2 * (int64_t)ctx->num_sms + tiles - 1) / tiles;
    const int64_t fit = (int64_t)(part_bytes / ((size_t)M * N * sizeof(float)));
    if (want > fit) want = fit;
    if (want > (K + 255) / 256) want = (K + 255) / 256;
    if (want > 65535) want = 65535;
    if (want > 1) {
      kchunk = (int)(((K + want - 1) / want + GT_BK - 1) / GT_BK * GT_BK);
      splits = (K + kchunk - 1) / kchunk;
    }
  }
  static bool attr_done = false;
  if (!attr_done) {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    attr_done = true;
  }
  GtArgs a;
  a.A = A; a.B = B; a.C = C; a.part = part; a.bias = bias; a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldc = ldc;
  a.transA = transA; a.transB = transB; a.accumulate = accumulate; a.act = act; a.kchunk = kchunk; a.splits = splits;
  dim3 grid((unsigned)nt, (unsigned)mt, (unsigned)splits);
  k_gemm_tc<<<grid, GT_THREADS, GT_SMEM, st>>>(a);
  KGEB_LAUNCH_OK(ctx);
  *splits_out = splits;
  return KGEB_OK;
}
