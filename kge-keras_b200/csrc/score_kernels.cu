// score_kernels.cu -- K1: gather + score (+ negatives + loss + gradient scatter) for DistMult / TransE / ComplEx,
// and the O(d^2) shared-memory HolE kernels.  One warp owns one triple (or one positive with its K negatives);
// each lane holds 128-bit chunks lane, lane+32, ... of every gathered row in registers, so a row is read from
// HBM/L2 exactly once per reference and gradients leave as 128-bit vector reductions (RED.ADD.F32x4).
//
// Reference arithmetic replaced: EmbeddingModel.call models.py:134-162, DistMult.func 171-172, TransE.func 186-190,
// ComplEx.func 212-222, HolE.func 231-237, losses utils.py:288-319, negatives utils.py:114-121.
#include "frag.cuh"

namespace {


// ---------------------------------------------------------------------------------------------------
// score_fwd: one warp per triple
// ---------------------------------------------------------------------------------------------------
template <int KIND, int NC, bool DP>
__global__ void __launch_bounds__(KGEB_THREADS) k_score_fwd(ScoreArgs a, const int32_t* __restrict__ idx, int64_t m,
                                                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5); t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    Frag<KIND, NC> S, P, O;
    load_frag(S, a.E, s, a.ldE, a.ge, lane);
    load_frag<KIND, NC, true>(P, a.R, p, a.ldR, a.gr, lane);
    load_frag(O, a.E, o, a.ldE, a.ge, lane);
    if (DP) {
      dp_apply(S, dp_bits<KIND, NC>(a.dp, 0, a.dp_stream, (uint64_t)t, lane), a.dp.scale);
      dp_apply(P, dp_bits<KIND, NC>(a.dp, 1, a.dp_stream, (uint64_t)t, lane), a.dp.scale);
      dp_apply(O, dp_bits<KIND, NC>(a.dp, 2, a.dp_stream, (uint64_t)t, lane), a.dp.scale);
    }
    const float sum = warp_sum(score_partial(S, P, O));
    if (lane == 0) out[t] = score_finish<KIND>(sum, a.gamma, a.param);
  }
}

// ---------------------------------------------------------------------------------------------------
// score_bwd: one warp per triple, gradient rows leave as vector reductions
// ---------------------------------------------------------------------------------------------------
template <int KIND, int NC, bool DP>
__global__ void __launch_bounds__(KGEB_THREADS) k_score_bwd(ScoreArgs a, const int32_t* __restrict__ idx, int64_t m,
                                                            const float* __restrict__ dscore, float* gE, float* gR,
                                                            uint8_t* tE, uint8_t* tR) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5); t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    const float coef = __ldg(dscore + t);
    if (lane == 0) { tE[s] = 1; tE[o] = 1; tR[p] = 1; }
    if (coef == 0.f) continue;
    Frag<KIND, NC> S, P, O, GS, GP, GO;
    load_frag(S, a.E, s, a.ldE, a.ge, lane);
    load_frag<KIND, NC, true>(P, a.R, p, a.ldR, a.gr, lane);
    load_frag(O, a.E, o, a.ldE, a.ge, lane);
    uint32_t mS = 0, mP = 0, mO = 0;
    if (DP) {
      mS = dp_bits<KIND, NC>(a.dp, 0, a.dp_stream, (uint64_t)t, lane);
      mP = dp_bits<KIND, NC>(a.dp, 1, a.dp_stream, (uint64_t)t, lane);
      mO = dp_bits<KIND, NC>(a.dp, 2, a.dp_stream, (uint64_t)t, lane);
      dp_apply(S, mS, a.dp.scale); dp_apply(P, mP, a.dp.scale); dp_apply(O, mO, a.dp.scale);
    }
    float sum = 0.f;
    if (KIND == TRANSE_L2 || KIND == MODE_L2) sum = warp_sum(score_partial(S, P, O));
    score_grads(S, P, O, coef, sum, a.gamma, GS, GP, GO, a.param);
    if (DP) { dp_apply(GS, mS, a.dp.scale); dp_apply(GP, mP, a.dp.scale); dp_apply(GO, mO, a.dp.scale); }
    red_frag(gE, s, a.ldE, a.ge, lane, GS);
    red_frag<KIND, NC, true>(gR, p, a.ldR, a.gr, lane, GP);
    red_frag(gE, o, a.ldE, a.ge, lane, GO);
  }
}

// ---------------------------------------------------------------------------------------------------
// HolE: f = sum_k p_k x_k,  x_k = sum_i s_i o_{(i+k) mod d}  (circular correlation, models.py:231-237), O(d^2) per triple.
// One warp per triple; lane l owns 8 CONSECUTIVE k (k0 = 8 l + 256 pass), so for a block of 8 consecutive i the 64 products
// s_{i+u} o_{i+u+k0+m} read a sliding window of 16 o-values: 2 + 4 LDS.128 per 64 FMA instead of two loads per FMA.  Rows
// are staged in shared memory, the candidate row twice in a row (o2[x] = o[x mod d]) so the window needs no modulo, the
// multiplier row zero-padded to a multiple of 8.  Each x_k is still accumulated over i in ascending order with fmaf.
// ---------------------------------------------------------------------------------------------------
// per-warp staging sizes in floats
__device__ __host__ __forceinline__ int hole_lin(int d) { return ((d + 7) & ~7); }               // zero-padded row
__device__ __host__ __forceinline__ int hole_dbl(int d) { return 2 * ((d + 7) & ~7) + 32; }      // doubled (circular) row

// dst[x] = row[(x - shift) mod d] for x < len (shift = 0: plain doubling), dropout applied per element
__device__ __forceinline__ void hole_stage_circ(float* dst, const float* __restrict__ src, int d, int len, int shift, int lane,
                                                const DropArgs& dp, int slot, int stream, uint64_t g) {
  for (int x = lane; x < len; x += 32) {
    int i = (x - shift) % d;
    if (i < 0) i += d;
    float v = __ldg(src + i);
    if (dp.thr) v = dp_keep_elem(dp, slot, stream, g, i) ? __fmul_rn(v, dp.scale) : 0.f;
    dst[x] = v;
  }
}
// dst[i] = row[i] for i < d, 0 up to the padded length
__device__ __forceinline__ void hole_stage_lin(float* dst, const float* __restrict__ src, int d, int lane, const DropArgs& dp, int slot,
                                               int stream, uint64_t g) {
  const int dp8 = hole_lin(d);
  for (int i = lane; i < dp8; i += 32) {
    float v = 0.f;
    if (i < d) {
      v = __ldg(src + i);
      if (dp.thr) v = dp_keep_elem(dp, slot, stream, g, i) ? __fmul_rn(v, dp.scale) : 0.f;
    }
    dst[i] = v;
  }
}
__device__ __forceinline__ void ld8(float (&v)[8], const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld16(float (&v)[16], const float* p) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(p + 4 * q);
    v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
  }
}

__global__ void __launch_bounds__(KGEB_THREADS) k_hole_fwd(ScoreArgs a, int d, const int32_t* __restrict__ idx, int64_t m,
                                                           float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int dl = hole_lin(d), dd = hole_dbl(d);
  float* ss = smem + (size_t)w * (dl + dd);
  float* so2 = ss + dl;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    __syncwarp();
    hole_stage_lin(ss, a.E + (int64_t)s * a.ldE, d, lane, a.dp, 0, a.dp_stream, (uint64_t)t);
    hole_stage_circ(so2, a.E + (int64_t)o * a.ldE, d, dd, 0, lane, a.dp, 2, a.dp_stream, (uint64_t)t);
    __syncwarp();
    const float* pr = a.R + (int64_t)p * a.ldR;
    float acc = 0.f;
    for (int k0 = 8 * lane; k0 < d; k0 += 256) {
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = 0.f;
      for (int i0 = 0; i0 < dl; i0 += 8) {
        float s8[8], w16[16];
        ld8(s8, ss + i0);
        ld16(w16, so2 + i0 + k0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = fmaf(s8[u], w16[u + q], x[q]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = k0 + q;
        if (k < d) {
          float pv = __ldg(pr + k);
          if (a.dp.thr) pv = dp_keep_elem(a.dp, 1, a.dp_stream, (uint64_t)t, k) ? __fmul_rn(pv, a.dp.scale) : 0.f;
          acc = fmaf(pv, x[q], acc);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[t] = acc;
  }
}

__global__ void __launch_bounds__(KGEB_THREADS) k_hole_bwd(ScoreArgs a, int d, const int32_t* __restrict__ idx, int64_t m,
                                                           const float* __restrict__ dscore, float* gE, float* gR,
                                                           uint8_t* tE, uint8_t* tR) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int dl = hole_lin(d), dd = hole_dbl(d);
  float* ss = smem + (size_t)w * 2 * (dl + dd);   // s, zero padded
  float* sp = ss + dl;                            // p, zero padded
  float* so2 = sp + dl;                           // o doubled: so2[x] = o[x mod d]
  float* sp2 = so2 + dd;                          // p doubled and shifted: sp2[x] = p[(x - dl) mod d]
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    const float coef = __ldg(dscore + t);
    if (lane == 0) { tE[s] = 1; tE[o] = 1; tR[p] = 1; }
    if (coef == 0.f) continue;
    __syncwarp();
    hole_stage_lin(ss, a.E + (int64_t)s * a.ldE, d, lane, a.dp, 0, a.dp_stream, (uint64_t)t);
    hole_stage_lin(sp, a.R + (int64_t)p * a.ldR, d, lane, a.dp, 1, a.dp_stream, (uint64_t)t);
    hole_stage_circ(so2, a.E + (int64_t)o * a.ldE, d, dd, 0, lane, a.dp, 2, a.dp_stream, (uint64_t)t);
    hole_stage_circ(sp2, a.R + (int64_t)p * a.ldR, d, dd, dl, lane, a.dp, 1, a.dp_stream, (uint64_t)t);
    __syncwarp();
    float* gs = gE + (int64_t)s * a.ldE;
    float* gp = gR + (int64_t)p * a.ldR;
    float* go = gE + (int64_t)o * a.ldE;
    for (int k0 = 8 * lane; k0 < d; k0 += 256) {
      // d/dp_k = sum_i s_i o_{i+k};  d/ds_k = sum_j p_j o_{k+j};  d/do_k = sum_j s_j p_{k-j}   (all times coef)
      float xp[8], xs[8], xo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { xp[q] = 0.f; xs[q] = 0.f; xo[q] = 0.f; }
      for (int i0 = 0; i0 < dl; i0 += 8) {
        float s8[8], p8[8], wo[16], wp[16];
        ld8(s8, ss + i0);
        ld8(p8, sp + i0);
        ld16(wo, so2 + i0 + k0);                  // o_{i0 + u + k0 + q} = wo[u + q]
        ld16(wp, sp2 + dl + k0 - i0 - 8);         // p_{k0 + q - i0 - u} = wp[8 + q - u]
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            xp[q] = fmaf(s8[u], wo[u + q], xp[q]);
            xs[q] = fmaf(p8[u], wo[u + q], xs[q]);
            xo[q] = fmaf(s8[u], wp[8 + q - u], xo[q]);
          }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = k0 + q;
        if (k < d) {
          float cp = coef, cs = coef, co = coef;   // gradient wrt the raw rows = gradient wrt the dropped rows * keep * scale
          if (a.dp.thr) {
            cp = dp_keep_elem(a.dp, 1, a.dp_stream, (uint64_t)t, k) ? coef * a.dp.scale : 0.f;
            cs = dp_keep_elem(a.dp, 0, a.dp_stream, (uint64_t)t, k) ? coef * a.dp.scale : 0.f;
            co = dp_keep_elem(a.dp, 2, a.dp_stream, (uint64_t)t, k) ? coef * a.dp.scale : 0.f;
          }
          atomicAdd(gp + k, cp * xp[q]);
          atomicAdd(gs + k, cs * xs[q]);
          atomicAdd(go + k, co * xo[q]);
        }
      }
    }
  }
}

// negatives only (parity of the sampler against its host restatement)
__global__ void k_negatives(const int32_t* __restrict__ pos, int64_t B, int K, uint64_t seed, uint64_t offset, uint32_t N,
                            int mode, int32_t* __restrict__ neg) {
  const int64_t total = B * (int64_t)K;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = g / K;
    const int s = __ldg(pos + 3 * i), p = __ldg(pos + 3 * i + 1), o = __ldg(pos + 3 * i + 2);
    int ns, no;
    philox_corrupt(s, o, (uint64_t)g, seed, offset, N, mode, ns, no);
    neg[3 * g] = ns; neg[3 * g + 1] = p; neg[3 * g + 2] = no;
  }
}

// oversample_data (utils.py:123-147) on the device: dst[i] = src[u_i], u_i uniform in [0, n) with replacement
__global__ void k_oversample(const int32_t* __restrict__ src, int64_t n, int64_t m, uint64_t seed, uint64_t offset, int stream_id,
                             int32_t* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                  make_uint2((uint32_t)seed ^ 0x4F565253u, (uint32_t)(seed >> 32) ^ (uint32_t)stream_id));
    const int64_t u = (int64_t)(((unsigned long long)r.x * (unsigned long long)n) >> 32);
    dst[3 * i] = __ldg(src + 3 * u); dst[3 * i + 1] = __ldg(src + 3 * u + 1); dst[3 * i + 2] = __ldg(src + 3 * u + 2);
  }
}

}  // namespace

extern "C" int kgeb_oversample(kgeb_ctx* ctx, const int32_t* src, int64_t n, int64_t m, uint64_t seed, uint64_t offset, int stream_id,
                               int32_t* dst, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, n >= 0 && m >= 0 && n < 0xffffffffLL, KGEB_ERR_INVALID, "oversample: bad sizes");
  if (m == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, src && dst && n > 0, KGEB_ERR_INVALID, "oversample: cannot draw from an empty list");
  int64_t blocks = (m + 255) / 256;
  if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
  k_oversample<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, n, m, seed, offset, stream_id, dst);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_score_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores,
                              kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, m >= 0, KGEB_ERR_INVALID, "score_fwd: negative batch");
  if (m == 0) return KGEB_OK;  /* empty batch: nothing to read, pointers may be null */
  KGEB_REQUIRE(ctx, idx && scores, KGEB_ERR_INVALID, "score_fwd: bad arguments");
  if (ctx->kind == KGEB_HAKE) return kgeb_internal_hake_fwd(ctx, E, R, idx, m, scores, (cudaStream_t)stream);
  if (int rc = check_tables(ctx, E, R)) return rc;
  if (m == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int nc;
  ScoreArgs a = make_score_args(ctx, E, R, &nc);
  const int grid = grid_for(ctx, m, 8);
  if (ctx->kind == KGEB_HOLE) {
    KGEB_REQUIRE(ctx, ctx->we == ctx->wr, KGEB_ERR_UNSUPPORTED, "HolE needs e_dim == r_dim");
    if (ctx->hole_fft) {
      const int rc = kgeb_internal_hole_fft_fwd(ctx, E, R, idx, m, scores, st);
      if (rc != KGEB_ERR_UNSUPPORTED) return rc;      // not a power-of-two width: the direct kernel below
    }
    const size_t sm = (size_t)KGEB_WARPS_PER_BLOCK * (hole_lin(ctx->we) + hole_dbl(ctx->we)) * sizeof(float);
    KGEB_REQUIRE(ctx, sm <= 200 * 1024, KGEB_ERR_UNSUPPORTED, "HolE dim too large for shared memory staging");
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fwd<<<grid, KGEB_THREADS, sm, st>>>(a, ctx->we, idx, m, scores);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  if (a.dp.thr) {
    KGEB_DISPATCH_KIND_ALL(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_fwd<KIND, NC, true><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, scores))));
  } else {
    KGEB_DISPATCH_KIND_ALL(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_fwd<KIND, NC, false><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, scores))));
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_score_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m,
                              const float* dscore, float* gE, float* gR, uint8_t* touchedE, uint8_t* touchedR,
                              kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, m >= 0, KGEB_ERR_INVALID, "score_bwd: negative batch");
  if (m == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, idx && dscore && gE && gR && touchedE && touchedR, KGEB_ERR_INVALID, "score_bwd: bad arguments");
  if (ctx->kind == KGEB_HAKE) return kgeb_internal_hake_bwd(ctx, E, R, idx, m, dscore, gE, gR, touchedE, touchedR, (cudaStream_t)stream);
  if (int rc = check_tables(ctx, E, R)) return rc;
  KGEB_REQUIRE(ctx, ctx->kind == KGEB_HOLE || (kgeb_aligned16(gE) && kgeb_aligned16(gR)), KGEB_ERR_ALIGNMENT,
               "gradient buffers must be 16-byte aligned");
  if (m == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int nc;
  ScoreArgs a = make_score_args(ctx, E, R, &nc);
  const int grid = grid_for(ctx, m, 8);
  if (ctx->kind == KGEB_HOLE) {
    if (ctx->hole_fft) {
      const int rc = kgeb_internal_hole_fft_bwd(ctx, E, R, idx, m, dscore, gE, gR, touchedE, touchedR, st);
      if (rc != KGEB_ERR_UNSUPPORTED) return rc;
    }
    const size_t sm = (size_t)KGEB_WARPS_PER_BLOCK * 2 * (hole_lin(ctx->we) + hole_dbl(ctx->we)) * sizeof(float);
    KGEB_REQUIRE(ctx, sm <= 200 * 1024, KGEB_ERR_UNSUPPORTED, "HolE dim too large for shared memory staging");
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_bwd<<<grid, KGEB_THREADS, sm, st>>>(a, ctx->we, idx, m, dscore, gE, gR, touchedE, touchedR);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  if (a.dp.thr) {
    KGEB_DISPATCH_KIND_ALL(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_bwd<KIND, NC, true><<<grid, KGEB_THREADS, 0, st>>>(
                                                           a, idx, m, dscore, gE, gR, touchedE, touchedR))));
  } else {
    KGEB_DISPATCH_KIND_ALL(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_bwd<KIND, NC, false><<<grid, KGEB_THREADS, 0, st>>>(
                                                           a, idx, m, dscore, gE, gR, touchedE, touchedR))));
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_negatives(kgeb_ctx* ctx, const int32_t* pos, int64_t B, int K, uint64_t seed, uint64_t offset,
                              int corrupt_mode, int32_t* neg, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, pos && neg && B >= 0 && K >= 0, KGEB_ERR_INVALID, "negatives: bad arguments");
  KGEB_REQUIRE(ctx, corrupt_mode >= 0 && corrupt_mode <= 3, KGEB_ERR_INVALID, "negatives: unknown corrupt mode");
  if (B * K == 0) return KGEB_OK;
  const int64_t total = B * (int64_t)K;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
  k_negatives<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pos, B, K, seed, offset, (uint32_t)ctx->N, corrupt_mode, neg);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

