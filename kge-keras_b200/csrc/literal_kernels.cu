// literal_kernels.cu -- the literal branch of EmbeddingModel.call (models.py:114-120,152-158): every gathered entity row is
// concatenated with its (non-trainable) literal row, both after Dropout, and sent through ONE shared Dense(e_dim, activation)
// before the score function:   s' = act([dp(E[s]) | dp(L[s])] W + b).
//
// The reference's own harness trains exactly this (tests.py:113,125-126: 5 random literals, tanh).  Per-reference dropout
// masks make the transformed row a function of the reference, not of the entity, so the step materialises the rows:
//   rows_concat      X[i]  = [dp(E[ids[i]]) | dp(L[ids[i]])]                       (also used for the relation rows, L = NULL)
//   sgemm            Z     = X W                                                   (small fp32 GEMMs: K = e_dim + literals)
//   dense_act_fwd    Y     = act(Z + b)   -> rows of a scratch "entity table" scored by the ordinary kernels with identity ids
//   dense_act_bwd    gZ    = gY * act'(Y),  gb += column sums
//   sgemm            gW    = X^T gZ,   gX = gZ W^T
//   rows_concat_bwd  gE[ids[i]] += dp-mask * gX[i, :e_dim]                         (literals are not trainable)
// Everything is fp32 on the CUDA cores: the Dense is tiny (K ~ 100) and the parity bar is 1e-5.
#include "frag.cuh"

namespace {

// keep * scale of element x of materialised row g (plain rows: the same counter layout as dp_keep_elem / the oracle)
__device__ __forceinline__ float lit_mult(const DropArgs& dp, int stream, uint64_t g, int x) {
  if (!dp.thr) return 1.f;
  return dp_keep_elem(dp, 0, stream, g, x) ? dp.scale : 0.f;
}

__global__ void __launch_bounds__(256) k_rows_concat(const float* __restrict__ E, int we, int ldE, const float* __restrict__ L, int wl, int ldL,
                                                     const int32_t* __restrict__ ids, int64_t n, DropArgs dp, int stream, float* __restrict__ X,
                                                     int ldX) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = ids[i];
    const float* e = E + r * ldE;
    float* x = X + i * (int64_t)ldX;
    for (int k = lane; k < we; k += 32) x[k] = __fmul_rn(__ldg(e + k), lit_mult(dp, stream, (uint64_t)i, k));
    if (L) {
      const float* l = L + r * ldL;
      for (int k = lane; k < wl; k += 32) x[we + k] = __fmul_rn(__ldg(l + k), lit_mult(dp, stream, (uint64_t)i, we + k));
    }
  }
}

__global__ void __launch_bounds__(256) k_rows_concat_bwd(const float* __restrict__ gX, int ldX, int we, const int32_t* __restrict__ ids, int64_t n,
                                                         DropArgs dp, int stream, float* gE, int ldE, uint8_t* touched) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = ids[i];
    const float* g = gX + i * (int64_t)ldX;
    float* dst = gE + r * ldE;
    if (lane == 0) touched[r] = 1;
    for (int k = lane; k < we; k += 32) {
      const float v = __fmul_rn(g[k], lit_mult(dp, stream, (uint64_t)i, k));
      if (v != 0.f) atomicAdd(dst + k, v);
    }
  }
}

// C[M,N] = op(A) op(B) (+ C if accumulate); op(A) is M x K, op(B) is K x N.  64 x 64 tile, 16-wide k-steps, 4 x 4 per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(int64_t M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                               float* __restrict__ C, int ldc, int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int x = tid; x < 64 * 16; x += 256) {
      // A tile: element (m, k)
      const int mm = TA ? x & 63 : x >> 4, kk = TA ? x >> 6 : x & 15;
      const int64_t m = m0 + mm;
      const int k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? (TA ? A[(int64_t)k * lda + m] : A[m * lda + k]) : 0.f;
      // B tile: element (k, n)
      const int nn = TB ? x >> 4 : x & 63, kb = TB ? x & 15 : x >> 6;
      const int n = n0 + nn, kq = k0 + kb;
      Bs[kb][nn] = (n < N && kq < K) ? (TB ? B[(int64_t)n * ldb + kq] : B[(int64_t)kq * ldb + n]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[m * ldc + n] = accumulate ? C[m * ldc + n] + acc[i][j] : acc[i][j];
    }
  }
}

__device__ __forceinline__ float act_fwd(int act, float z) {
  if (act == KGEB_ACT_TANH) return tanhf(z);
  if (act == KGEB_ACT_RELU) return fmaxf(z, 0.f);
  if (act == KGEB_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  return z;
}
__device__ __forceinline__ float act_bwd(int act, float y) {   // derivative expressed through the OUTPUT y
  if (act == KGEB_ACT_TANH) return 1.f - y * y;
  if (act == KGEB_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == KGEB_ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

__global__ void k_dense_act_fwd(float* Z, int64_t M, int N, int ld, const float* __restrict__ bias, int act) {
  const int64_t total = M * N;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = x / N;
    const int n = (int)(x - m * N);
    Z[m * ld + n] = act_fwd(act, Z[m * ld + n] + __ldg(bias + n));
  }
}

// gY <- gY * act'(Y) in place; gbias[n] += sum_m gY[m, n]: one block per 32-column slab walks a strided set of rows
__global__ void __launch_bounds__(256) k_dense_act_bwd(const float* __restrict__ Y, float* gY, int64_t M, int N, int ld, int act, float* gbias) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N) {
    for (int64_t m = (int64_t)blockIdx.y * 8 + w; m < M; m += (int64_t)gridDim.y * 8) {
      const float g = gY[m * ld + n] * act_bwd(act, Y[m * ld + n]);
      gY[m * ld + n] = g;
      s += g;
    }
  }
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    if (t != 0.f) atomicAdd(gbias + n, t);
  }
}

int lit_grid(const kgeb_ctx* ctx, int64_t warps) {
  int64_t b = (warps + 7) / 8;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

DropArgs lit_drop(const kgeb_ctx* ctx) {
  DropArgs d;
  d.thr = ctx->dp_thr; d.scale = ctx->dp_scale;
  d.k0 = (uint32_t)ctx->dp_seed ^ 0x44524F50u; d.k1 = (uint32_t)(ctx->dp_seed >> 32);
  d.ylo = (uint32_t)(ctx->dp_offset << 3); d.whi = (uint32_t)(ctx->dp_offset >> 29);
  return d;
}

}  // namespace

extern "C" int kgeb_rows_concat(kgeb_ctx* ctx, const float* E, int we, int ldE, const float* L, int wl, int ldL, const int32_t* ids, int64_t n,
                                float* X, int ldX, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, n >= 0 && we > 0 && ldE >= we && wl >= 0 && ldX >= we + wl, KGEB_ERR_INVALID, "rows_concat: bad shape");
  if (n == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, E && ids && X && (wl == 0 || (L && ldL >= wl)), KGEB_ERR_INVALID, "rows_concat: bad arguments");
  k_rows_concat<<<lit_grid(ctx, n), 256, 0, (cudaStream_t)stream>>>(E, we, ldE, wl ? L : nullptr, wl, ldL, ids, n, lit_drop(ctx), ctx->dp_stream, X, ldX);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_rows_concat_bwd(kgeb_ctx* ctx, const float* gX, int ldX, int we, const int32_t* ids, int64_t n, float* gE, int ldE,
                                    uint8_t* touched, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, n >= 0 && we > 0 && ldX >= we && ldE >= we, KGEB_ERR_INVALID, "rows_concat_bwd: bad shape");
  if (n == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, gX && ids && gE && touched, KGEB_ERR_INVALID, "rows_concat_bwd: bad arguments");
  k_rows_concat_bwd<<<lit_grid(ctx, n), 256, 0, (cudaStream_t)stream>>>(gX, ldX, we, ids, n, lit_drop(ctx), ctx->dp_stream, gE, ldE, touched);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_sgemm(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
                          float* C, int ldc, int accumulate, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, M >= 0 && N > 0 && K > 0 && ldc >= N, KGEB_ERR_INVALID, "sgemm: bad shape");
  if (M == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, A && B && C, KGEB_ERR_INVALID, "sgemm: null pointer");
  KGEB_REQUIRE(ctx, (M + 63) / 64 <= 65535, KGEB_ERR_UNSUPPORTED, "sgemm: too many row tiles");
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (transA && transB) k_sgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate);
  else if (transA) k_sgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate);
  else if (transB) k_sgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate);
  else k_sgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_dense_act_fwd(kgeb_ctx* ctx, float* Z, int64_t M, int N, int ld, const float* bias, int act, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, M >= 0 && N > 0 && ld >= N && act >= KGEB_ACT_LINEAR && act <= KGEB_ACT_SIGMOID, KGEB_ERR_INVALID, "dense_act_fwd: bad arguments");
  if (M == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, Z && bias, KGEB_ERR_INVALID, "dense_act_fwd: null pointer");
  int64_t blocks = (M * N + 255) / 256;
  if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
  k_dense_act_fwd<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(Z, M, N, ld, bias, act);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_dense_act_bwd(kgeb_ctx* ctx, const float* Y, float* gY, int64_t M, int N, int ld, int act, float* gbias, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, M >= 0 && N > 0 && ld >= N && act >= KGEB_ACT_LINEAR && act <= KGEB_ACT_SIGMOID, KGEB_ERR_INVALID, "dense_act_bwd: bad arguments");
  if (M == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, Y && gY && gbias, KGEB_ERR_INVALID, "dense_act_bwd: null pointer");
  int64_t rows = (M + 7) / 8;
  if (rows > 64) rows = 64;
  k_dense_act_bwd<<<dim3((unsigned)((N + 31) / 32), (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(Y, gY, M, N, ld, act, gbias);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
