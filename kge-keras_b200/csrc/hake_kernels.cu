// hake_kernels.cu -- HAKE.func (models.py:383-400) as the reference wrote it:
//   score_i = gamma - ( phase_weight * sum_k |sin((ps_k + pp_k - po_k) / 2)|  +  mod_weight * || U ||_F ),
//   U_ik = mod_s (mod_p + bias) - |mod_o| (1 - bias),  bias = where(min(bias_p, 1) < -|mod_p|, -|mod_p|, min(bias_p, 1)).
// The modulus norm has no axis (models.py:397): it is ONE scalar over every row of the call, so the scores of a call are
// coupled and so are their gradients.  Entity rows are [phase | modulus] (2d floats), relation rows [phase | modulus | bias] (3d).
//
//   forward   k_hake_rows (warp per triple: phase sum -> scores, sum_k U^2 -> rsq[i])
//             k_hake_finish (one block: T = sum_i rsq[i] in double, fixed order; scores -= mod_weight * sqrt(T))
//   backward  k_hake_rows + k_hake_finish again (T and C = sum_i dscore_i; the negative call ran in between)
//             k_hake_bwd  (warp per triple: d/d rows of  sum_i dscore_i * score_i, TensorFlow's conventions for |x|, clip, where)
// fp32 contract: phi = x / phase_div (phase_div = fp32(embedding_range / pi): the reference DIVIDES, models.py:390-392), a = ((phi_s + phi_p) - phi_o) * 0.5, sinf / cosf, sums over k lane-strided then butterfly.
#include "frag.cuh"

namespace {

struct HakeArgs {
  const float* E; const float* R; int ldE, ldR, d;
  float gamma, ps, pw, mw;
  DropArgs dp; int stream;
};

template <bool DP>
__device__ __forceinline__ float hake_ld(const HakeArgs& a, const float* row, int x, int slot, uint64_t g, float& mult) {
  if (DP) {
    mult = dp_keep_elem(a.dp, slot, a.stream, g, x) ? a.dp.scale : 0.f;
    return __fmul_rn(__ldg(row + x), mult);
  }
  mult = 1.f;
  return __ldg(row + x);
}

struct HakeElem {
  float a, u, ms, mo, mp, beta, b;
  bool cond;
  float m_sph, m_sm, m_oph, m_om, m_pph, m_pm, m_pb;   // dropout multipliers of the seven loaded elements
};

template <bool DP>
__device__ __forceinline__ HakeElem hake_elem(const HakeArgs& a, const float* es, const float* rp, const float* eo, int k, uint64_t g) {
  HakeElem e;
  const int d = a.d;
  const float phs = hake_ld<DP>(a, es, k, 0, g, e.m_sph), pho = hake_ld<DP>(a, eo, k, 2, g, e.m_oph), php = hake_ld<DP>(a, rp, k, 1, g, e.m_pph);
  e.ms = hake_ld<DP>(a, es, d + k, 0, g, e.m_sm);
  e.mo = hake_ld<DP>(a, eo, d + k, 2, g, e.m_om);
  e.mp = hake_ld<DP>(a, rp, d + k, 1, g, e.m_pm);
  e.beta = hake_ld<DP>(a, rp, 2 * d + k, 1, g, e.m_pb);
  e.a = __fmul_rn(__fsub_rn(__fadd_rn(__fdiv_rn(phs, a.ps), __fdiv_rn(php, a.ps)), __fdiv_rn(pho, a.ps)), 0.5f);
  const float b1 = fminf(e.beta, 1.f), nb = -fabsf(e.mp);
  e.cond = b1 < nb;
  e.b = e.cond ? nb : b1;
  e.u = __fsub_rn(__fmul_rn(e.ms, __fadd_rn(e.mp, e.b)), __fmul_rn(fabsf(e.mo), __fsub_rn(1.f, e.b)));
  return e;
}

template <bool DP>
__global__ void __launch_bounds__(KGEB_THREADS) k_hake_rows(HakeArgs a, const int32_t* __restrict__ idx, int64_t m, float* __restrict__ scores,
                                                            float* __restrict__ rsq) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < m; i += nwarps) {
    const float* es = a.E + (int64_t)idx[3 * i] * a.ldE;
    const float* rp = a.R + (int64_t)idx[3 * i + 1] * a.ldR;
    const float* eo = a.E + (int64_t)idx[3 * i + 2] * a.ldE;
    float psum = 0.f, usq = 0.f;
    for (int k = lane; k < a.d; k += 32) {
      const HakeElem e = hake_elem<DP>(a, es, rp, eo, k, (uint64_t)i);
      psum = __fadd_rn(psum, fabsf(sinf(e.a)));
      usq = fmaf(e.u, e.u, usq);
    }
    psum = warp_sum(psum);
    usq = warp_sum(usq);
    if (lane == 0) {
      if (scores) scores[i] = __fsub_rn(a.gamma, __fmul_rn(a.pw, psum));
      rsq[i] = usq;
    }
  }
}

// one block: out[0] = T = sum rsq, out[1] = C = sum dscore (0 without dscore); scores[i] -= mw * sqrt(T)
__global__ void __launch_bounds__(1024) k_hake_finish(const float* __restrict__ rsq, const float* __restrict__ dscore, int64_t m, float mw,
                                                      float* __restrict__ scores, double* __restrict__ out) {
  __shared__ double sh[2][32];
  __shared__ double tot[2];
  double t = 0.0, c = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    t += (double)rsq[i];
    if (dscore) c += (double)dscore[i];
  }
  t = warp_sum_d(t);
  c = warp_sum_d(c);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = t; sh[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double T = 0.0, Cs = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { T += sh[0][w]; Cs += sh[1][w]; }
    tot[0] = T; tot[1] = Cs;
    out[0] = T; out[1] = Cs;
  }
  __syncthreads();
  if (scores) {
    const float r = (float)((double)mw * sqrt(tot[0]));
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) scores[i] = __fsub_rn(scores[i], r);
  }
}

// inference through Keras' Model.predict scores the input in consecutive batches (batch_size, default 32), and every batch is
// its own call of func(): one norm per GROUP of `group` consecutive rows (the last group may be short).  One thread per group.
__global__ void k_hake_finish_groups(const float* __restrict__ rsq, int64_t m, int64_t group, float mw, float* __restrict__ scores) {
  const int64_t ng = (m + group - 1) / group;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = g * group, i1 = i0 + group < m ? i0 + group : m;
    double t = 0.0;
    for (int64_t i = i0; i < i1; ++i) t += (double)rsq[i];
    const float r = (float)((double)mw * sqrt(t));
    for (int64_t i = i0; i < i1; ++i) scores[i] = __fsub_rn(scores[i], r);
  }
}

__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.f) - (x < 0.f)); }
__device__ __forceinline__ void add_nz(float* p, float v) { if (v != 0.f) atomicAdd(p, v); }

template <bool DP>
__global__ void __launch_bounds__(KGEB_THREADS) k_hake_bwd(HakeArgs a, const int32_t* __restrict__ idx, int64_t m, const float* __restrict__ dscore,
                                                           const double* __restrict__ tc, float* gE, float* gR, uint8_t* tE, uint8_t* tR) {
  const int lane = threadIdx.x & 31, d = a.d;
  const double T = tc[0], Cs = tc[1];
  const float coefU = T > 0.0 ? (float)(-(Cs * (double)a.mw) / sqrt(T)) : 0.f;     // d(sum_i c_i score_i)/dU_ik = coefU * U_ik
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < m; i += nwarps) {
    const int64_t s = idx[3 * i], p = idx[3 * i + 1], o = idx[3 * i + 2];
    const float* es = a.E + s * a.ldE;
    const float* rp = a.R + p * a.ldR;
    const float* eo = a.E + o * a.ldE;
    float* gs = gE + s * a.ldE;
    float* gp = gR + p * a.ldR;
    float* go = gE + o * a.ldE;
    if (lane == 0) { tE[s] = 1; tE[o] = 1; tR[p] = 1; }
    const float cP = -a.pw * dscore[i];
    if (cP == 0.f && coefU == 0.f) continue;
    for (int k = lane; k < d; k += 32) {
      const HakeElem e = hake_elem<DP>(a, es, rp, eo, k, (uint64_t)i);
      const float dphi = __fdiv_rn(cP * sgn(sinf(e.a)) * cosf(e.a) * 0.5f, a.ps);
      add_nz(gs + k, dphi * e.m_sph);
      add_nz(gp + k, dphi * e.m_pph);
      add_nz(go + k, -dphi * e.m_oph);
      const float gu = coefU * e.u;
      const float db = gu * (e.ms + fabsf(e.mo));
      add_nz(gs + d + k, gu * (e.mp + e.b) * e.m_sm);
      add_nz(go + d + k, gu * (-sgn(e.mo) * (1.f - e.b)) * e.m_om);
      add_nz(gp + d + k, (gu * e.ms + (e.cond ? -sgn(e.mp) * db : 0.f)) * e.m_pm);
      add_nz(gp + 2 * d + k, ((!e.cond && e.beta <= 1.f) ? db : 0.f) * e.m_pb);
    }
  }
}

int hake_args(kgeb_ctx* ctx, const float* E, const float* R, int64_t m, HakeArgs* a) {
  KGEB_REQUIRE(ctx, E && R, KGEB_ERR_INVALID, "null table pointer");
  KGEB_REQUIRE(ctx, ctx->we % 2 == 0 && 2 * ctx->wr == 3 * ctx->we, KGEB_ERR_UNSUPPORTED,
               "HAKE needs entity rows [phase | modulus] (2d) and relation rows [phase | modulus | bias] (3d), models.py:371-372");
  KGEB_REQUIRE(ctx, ctx->hake_ps > 0.f, KGEB_ERR_INVALID, "HAKE: call kgeb_set_hake_params first");
  KGEB_REQUIRE(ctx, (size_t)m * sizeof(float) + 16 <= KGEB_SORT_SCRATCH_BYTES, KGEB_ERR_UNSUPPORTED, "HAKE: call too long for the context scratch (2 M triples)");
  a->E = E; a->R = R; a->ldE = ctx->ldE; a->ldR = ctx->ldR; a->d = ctx->we / 2;
  a->gamma = ctx->gamma; a->ps = ctx->hake_ps; a->pw = ctx->hake_pw; a->mw = ctx->hake_mw;
  a->dp.thr = ctx->dp_thr; a->dp.scale = ctx->dp_scale;
  a->dp.k0 = (uint32_t)ctx->dp_seed ^ 0x44524F50u; a->dp.k1 = (uint32_t)(ctx->dp_seed >> 32);
  a->dp.ylo = (uint32_t)(ctx->dp_offset << 3); a->dp.whi = (uint32_t)(ctx->dp_offset >> 29);
  a->stream = ctx->dp_stream;
  return KGEB_OK;
}

}  // namespace

int kgeb_internal_hake_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores, cudaStream_t st) {
  HakeArgs a;
  if (int rc = hake_args(ctx, E, R, m, &a)) return rc;
  double* tc = reinterpret_cast<double*>(ctx->d_sort_scratch);
  float* rsq = reinterpret_cast<float*>(ctx->d_sort_scratch + 16);
  const int grid = grid_for(ctx, m, 8);
  if (a.dp.thr) k_hake_rows<true><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, scores, rsq);
  else k_hake_rows<false><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, scores, rsq);
  KGEB_LAUNCH_OK(ctx);
  if (ctx->hake_group > 0 && ctx->hake_group < m) {
    const int64_t ng = (m + ctx->hake_group - 1) / ctx->hake_group;
    k_hake_finish_groups<<<(int)((ng + 127) / 128 < 1024 ? (ng + 127) / 128 : 1024), 128, 0, st>>>(rsq, m, ctx->hake_group, a.mw, scores);
  } else {
    k_hake_finish<<<1, 1024, 0, st>>>(rsq, nullptr, m, a.mw, scores, tc);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

int kgeb_internal_hake_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, const float* dscore, float* gE, float* gR,
                           uint8_t* tE, uint8_t* tR, cudaStream_t st) {
  HakeArgs a;
  if (int rc = hake_args(ctx, E, R, m, &a)) return rc;
  KGEB_REQUIRE(ctx, ctx->hake_group == 0, KGEB_ERR_UNSUPPORTED, "HAKE backward: kgeb_set_hake_call_rows() is an inference setting (Model.predict batches); reset it to 0");
  double* tc = reinterpret_cast<double*>(ctx->d_sort_scratch);
  float* rsq = reinterpret_cast<float*>(ctx->d_sort_scratch + 16);
  const int grid = grid_for(ctx, m, 8);
  if (a.dp.thr) k_hake_rows<true><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, nullptr, rsq);
  else k_hake_rows<false><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, nullptr, rsq);
  KGEB_LAUNCH_OK(ctx);
  k_hake_finish<<<1, 1024, 0, st>>>(rsq, dscore, m, a.mw, nullptr, tc);
  KGEB_LAUNCH_OK(ctx);
  if (a.dp.thr) k_hake_bwd<true><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, dscore, tc, gE, gR, tE, tR);
  else k_hake_bwd<false><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, dscore, tc, gE, gR, tE, tR);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
