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

// multipliers of the four elements of 128-bit chunk c of materialised row g: ONE Philox call per chunk (the counter layout of
// dp_keep_elem: lane = c & 31, call = c >> 6, half = (c >> 5) & 1)
__device__ __forceinline__ void lit_mult4(const DropArgs& dp, int stream, uint64_t g, int c, float (&mult)[4]) {
  if (!dp.thr) {
    mult[0] = mult[1] = mult[2] = mult[3] = 1.f;
    return;
  }
  const int t = c >> 5;
  const uint4 r = dp_philox(dp, 0, stream, g, c & 31, t >> 1);
#pragma unroll
  for (int e = 0; e < 4; ++e) mult[e] = (((dp_word(r, t & 1, e) >> (16 * (e & 1))) & 0xFFFFu) >= dp.thr) ? dp.scale : 0.f;
}

// a warp per row, a lane per 4-element chunk (one Philox call covers the chunk)
__global__ void __launch_bounds__(256) k_rows_concat(const float* __restrict__ E, int we, int ldE, const float* __restrict__ L, int wl, int ldL,
                                                     const int32_t* __restrict__ ids, int64_t n, DropArgs dp, int stream, float* __restrict__ X,
                                                     int ldX) {
  const int lane = threadIdx.x & 31;
  const int W = we + (L ? wl : 0), nchunk = (W + 3) >> 2;
  const bool vec = ((ldE | ldX | we) & 3) == 0 && !L && ((reinterpret_cast<uintptr_t>(E) | reinterpret_cast<uintptr_t>(X)) & 15) == 0;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = ids[i];
    const float* e = E + r * ldE;
    const float* l = L ? L + r * ldL : nullptr;
    float* x = X + i * (int64_t)ldX;
    for (int c = lane; c < nchunk; c += 32) {
      float mult[4];
      lit_mult4(dp, stream, (uint64_t)i, c, mult);
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(e) + c);
        *(reinterpret_cast<float4*>(x) + c) = make_float4(__fmul_rn(v.x, mult[0]), __fmul_rn(v.y, mult[1]), __fmul_rn(v.z, mult[2]), __fmul_rn(v.w, mult[3]));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = 4 * c + q;
          if (k < W) x[k] = __fmul_rn(k < we ? __ldg(e + k) : __ldg(l + (k - we)), mult[q]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_rows_concat_bwd(const float* __restrict__ gX, int ldX, int we, const int32_t* __restrict__ ids, int64_t n,
                                                         DropArgs dp, int stream, float* gE, int ldE, uint8_t* touched) {
  const int lane = threadIdx.x & 31;
  const int nchunk = (we + 3) >> 2;
  const bool vred = (ldE & 3) == 0 && (reinterpret_cast<uintptr_t>(gE) & 15) == 0;      // 16-byte aligned destination chunks
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = ids[i];
    const float* g = gX + i * (int64_t)ldX;
    float* dst = gE + r * ldE;
    if (lane == 0) touched[r] = 1;
    for (int c = lane; c < nchunk; c += 32) {
      float mult[4], v[4];
      lit_mult4(dp, stream, (uint64_t)i, c, mult);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = 4 * c + q;
        v[q] = k < we ? __fmul_rn(g[k], mult[q]) : 0.f;
      }
      if (v[0] == 0.f && v[1] == 0.f && v[2] == 0.f && v[3] == 0.f) continue;
      if (vred && 4 * c + 3 < we) {
        red_add4(dst + 4 * c, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (v[q] != 0.f) atomicAdd(dst + 4 * c + q, v[q]);
      }
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

// C[M,N] = op(A) op(B) (+ C if accumulate); op(A) is M x K, op(B) is K x N.  128 x 64 tile, 16-wide k-steps, 8 x 4 per thread
// (32 FFMA per 3 LDS.128); the next k-step's global loads are issued into registers before the current one is multiplied, so
// the L2 latency hides behind the 512 FFMAs a thread spends per step.
// Split-K (gridDim.z > 1, the weight gradient X^T gY whose K is the batch): slice z covers k in [z * kchunk, (z+1) * kchunk) and
// writes its partial tile to part[z][M][N]; k_splitk_reduce adds the slices in slice order, so the result does not depend on scheduling.
constexpr int GM = 128, GN = 64, GK = 16;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(int64_t M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                               float* __restrict__ C, int ldc, int accumulate, int kchunk, float* __restrict__ part, const float* __restrict__ bias,
                                               int act) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * GM;
  const int n0 = blockIdx.x * GN;
  const int kbeg = blockIdx.z * kchunk;
  if (gridDim.z > 1) {
    K = min(K, kbeg + kchunk);
    C = part + (int64_t)blockIdx.z * M * N;
    ldc = N;
    accumulate = 0;
  }
  // element x of a tile: the fastest-varying index follows the operand's contiguous axis so that a warp reads whole segments
  float pa[GM * GK / 256], pb[GN * GK / 256];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < GM * GK / 256; ++j) {
      const int x = tid + j * 256;
      const int mm = TA ? x & (GM - 1) : x >> 4, kk = TA ? x / GM : x & 15;
      const int64_t m = m0 + mm;
      const int k = k0 + kk;
      pa[j] = (m < M && k < K) ? (TA ? __ldg(A + (int64_t)k * lda + m) : __ldg(A + m * lda + k)) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < GN * GK / 256; ++j) {
      const int x = tid + j * 256;
      const int nn = TB ? x >> 4 : x & (GN - 1), kk = TB ? x & 15 : x / GN;
      const int n = n0 + nn, k = k0 + kk;
      pb[j] = (n < N && k < K) ? (TB ? __ldg(B + (int64_t)n * ldb + k) : __ldg(B + (int64_t)k * ldb + n)) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < GM * GK / 256; ++j) {
      const int x = tid + j * 256;
      As[TA ? x / GM : x & 15][TA ? x & (GM - 1) : x >> 4] = pa[j];
    }
#pragma unroll
    for (int j = 0; j < GN * GK / 256; ++j) {
      const int x = tid + j * 256;
      Bs[TB ? x & 15 : x / GN][TB ? x >> 4 : x & (GN - 1)] = pb[j];
    }
  };
  float acc[8][4] = {};
  if (kbeg < K) fetch(kbeg);
  for (int k0 = kbeg; k0 < K; k0 += GK) {
    stash();
    __syncthreads();
    if (k0 + GK < K) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = accumulate ? C[m * ldc + n] + acc[i][j] : acc[i][j];
      if (bias) v = act_fwd(act, v + __ldg(bias + n));      // fused Dense epilogue (never with split-K)
      C[m * ldc + n] = v;
    }
  }
}

__global__ void k_splitk_reduce(const float* __restrict__ part, int splits, int64_t MN, int N, float* __restrict__ C, int ldc, int accumulate) {
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < MN; x += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {          // eight loads in flight, added in slice order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(z + u) * MN + x];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += part[(int64_t)z * MN + x];
    const int64_t m = x / N;
    float* dst = C + m * ldc + (x - m * N);
    *dst = accumulate ? *dst + s : s;
  }
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
    const int64_t step = (int64_t)gridDim.y * 8;
    int64_t m = (int64_t)blockIdx.y * 8 + w;
    for (; m + 3 * step < M; m += 4 * step) {     // four rows in flight per warp: the loop is latency-bound otherwise
      float g[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { g[u] = gY[(m + u * step) * ld + n]; y[u] = Y[(m + u * step) * ld + n]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] *= act_bwd(act, y[u]);
        gY[(m + u * step) * ld + n] = g[u];
        s += g[u];
      }
    }
    for (; m < M; m += step) {
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

// tensor-core path (gemm_tc_kernels.cu): tcgen05 kind::tf32 with a three-term split; KGEB_GEMM=ffma keeps the CUDA-core kernel (A/B)
int kgeb_internal_gemm_tc(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
                          float* C, int ldc, int accumulate, const float* bias, int act, float* part, size_t part_bytes, int* splits_out,
                          cudaStream_t st);

static int launch_sgemm(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                        int ldc, int accumulate, const float* bias, int act, cudaStream_t st) {
  static const bool force_ffma = [] { const char* e = getenv("KGEB_GEMM"); return e && strcmp(e, "ffma") == 0; }();
  if (!force_ffma) {
    int tc_splits = 1;
    const int rc = kgeb_internal_gemm_tc(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias, act,
                                         reinterpret_cast<float*>(ctx->d_sort_scratch), KGEB_SORT_SCRATCH_BYTES, &tc_splits, st);
    if (rc == KGEB_OK) {
      if (tc_splits > 1) {
        const int64_t MN = M * N;
        k_splitk_reduce<<<(int)((MN + 63) / 64 < 2048 ? (MN + 63) / 64 : 2048), 64, 0, st>>>(reinterpret_cast<float*>(ctx->d_sort_scratch), tc_splits, MN, N, C,
                                                                                        ldc, accumulate);
        KGEB_LAUNCH_OK(ctx);
      }
      return KGEB_OK;
    }
    if (rc != KGEB_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid((unsigned)((N + GN - 1) / GN), (unsigned)((M + GM - 1) / GM));
  // few output tiles and a long K: slice K so that about two waves of CTAs are in flight (the slices fit the context scratch)
  const int64_t tiles = (int64_t)grid.x * grid.y;
  int splits = 1, kchunk = K;
  if (!bias && tiles < ctx->num_sms && K >= 1024) {
    int64_t want = (2 * (int64_t)ctx->num_sms + tiles - 1) / tiles;
    const int64_t fit = (int64_t)KGEB_SORT_SCRATCH_BYTES / (M * N * (int64_t)sizeof(float));
    if (want > fit) want = fit;
    if (want > (K + 255) / 256) want = (K + 255) / 256;
    if (want > 1) {
      kchunk = (int)(((K + want - 1) / want + 15) / 16 * 16);
      splits = (K + kchunk - 1) / kchunk;
    }
  }
  grid.z = (unsigned)splits;
  float* part = reinterpret_cast<float*>(ctx->d_sort_scratch);
  if (transA && transB) k_sgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, bias, act);
  else if (transA) k_sgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, bias, act);
  else if (transB) k_sgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, bias, act);
  else k_sgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, bias, act);
  if (splits > 1) {
    const int64_t MN = M * N;
    k_splitk_reduce<<<(int)((MN + 63) / 64 < 2048 ? (MN + 63) / 64 : 2048), 64, 0, st>>>(part, splits, MN, N, C, ldc, accumulate);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_sgemm(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
                          float* C, int ldc, int accumulate, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, M >= 0 && N > 0 && K > 0 && ldc >= N, KGEB_ERR_INVALID, "sgemm: bad shape");
  if (M == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, A && B && C, KGEB_ERR_INVALID, "sgemm: null pointer");
  KGEB_REQUIRE(ctx, (M + GM - 1) / GM <= 65535, KGEB_ERR_UNSUPPORTED, "sgemm: too many row tiles");
  return launch_sgemm(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int kgeb_dense_fwd(kgeb_ctx* ctx, const float* X, int ldx, int64_t rows, int kin, const float* W, int ldw, int width, const float* bias,
                              int act, float* Y, int ldy, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, rows >= 0 && kin > 0 && width > 0 && ldx >= kin && ldw >= width && ldy >= width && act >= KGEB_ACT_LINEAR && act <= KGEB_ACT_SIGMOID,
               KGEB_ERR_INVALID, "dense_fwd: bad arguments");
  if (rows == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, X && W && bias && Y, KGEB_ERR_INVALID, "dense_fwd: null pointer");
  KGEB_REQUIRE(ctx, (rows + GM - 1) / GM <= 65535, KGEB_ERR_UNSUPPORTED, "dense_fwd: too many row tiles");
  return launch_sgemm(ctx, 0, 0, rows, width, kin, X, ldx, W, ldw, Y, ldy, 0, bias, act, (cudaStream_t)stream);
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
  const int64_t slabs = (N + 31) / 32;
  int64_t rows = (M + 31) / 32, cap = ((int64_t)ctx->num_sms * 8 + slabs - 1) / slabs;
  if (rows > cap) rows = cap;
  if (rows > 65535) rows = 65535;
  k_dense_act_bwd<<<dim3((unsigned)((N + 31) / 32), (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(Y, gY, M, N, ld, act, gbias);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
