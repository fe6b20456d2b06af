// score_kernels.cu -- K1: gather + score (+ negatives + loss + gradient scatter) for DistMult / TransE / ComplEx,
// and the O(d^2) shared-memory HolE kernels.  One warp owns one triple (or one positive with its K negatives);
// each lane holds 128-bit chunks lane, lane+32, ... of every gathered row in registers, so a row is read from
// HBM/L2 exactly once per reference and gradients leave as 128-bit vector reductions (RED.ADD.F32x4).
//
// Reference arithmetic replaced: EmbeddingModel.call models.py:134-162, DistMult.func 171-172, TransE.func 186-190,
// ComplEx.func 212-222, HolE.func 231-237, losses utils.py:288-319, negatives utils.py:114-121.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int DISTMULT = KGEB_DISTMULT, TRANSE_L1 = KGEB_TRANSE_L1, TRANSE_L2 = KGEB_TRANSE_L2, COMPLEX = KGEB_COMPLEX;

// Row fragment: NC chunks per lane (ComplEx: NC chunks of the real half followed by NC of the imaginary half).
template <int KIND, int NC>
struct Frag {
  static constexpr int NT = (KIND == COMPLEX) ? 2 * NC : NC;
  float4 v[NT];
};

struct RowGeom {
  int nch;   // 128-bit chunks per (half-)row handled by the lane loop: width/4, or (width/2)/4 for ComplEx
  int half4; // ComplEx: offset of the imaginary half in float4 units; otherwise 0
};

template <int KIND, int NC>
__device__ __forceinline__ void load_frag(Frag<KIND, NC>& f, const float* __restrict__ table, int64_t row, int ld,
                                          RowGeom g, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(table + row * (int64_t)ld);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    const bool ok = c < g.nch;
    f.v[i] = ok ? ldg4(r4 + c) : f4zero();
    if (KIND == COMPLEX) f.v[NC + i] = ok ? ldg4(r4 + g.half4 + c) : f4zero();
  }
}

template <int KIND, int NC>
__device__ __forceinline__ void zero_frag(Frag<KIND, NC>& f) {
#pragma unroll
  for (int i = 0; i < Frag<KIND, NC>::NT; ++i) f.v[i] = f4zero();
}
template <int KIND, int NC>
__device__ __forceinline__ void add_frag(Frag<KIND, NC>& a, const Frag<KIND, NC>& b) {
#pragma unroll
  for (int i = 0; i < Frag<KIND, NC>::NT; ++i) {
    a.v[i].x += b.v[i].x; a.v[i].y += b.v[i].y; a.v[i].z += b.v[i].z; a.v[i].w += b.v[i].w;
  }
}

template <int KIND, int NC>
__device__ __forceinline__ void red_frag(float* __restrict__ table, int64_t row, int ld, RowGeom g, int lane,
                                         const Frag<KIND, NC>& f) {
  float* r = table + row * (int64_t)ld;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    if (c < g.nch) {
      red_add4(r + 4 * c, f.v[i]);
      if (KIND == COMPLEX) red_add4(r + 4 * (g.half4 + c), f.v[NC + i]);
    }
  }
}

#define KGEB_FOR4(...)                    \
  {                                       \
    { constexpr int _e = 0; __VA_ARGS__ } \
    { constexpr int _e = 1; __VA_ARGS__ } \
    { constexpr int _e = 2; __VA_ARGS__ } \
    { constexpr int _e = 3; __VA_ARGS__ } \
  }
__device__ __forceinline__ float& comp(float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ float comp(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }

// lane-partial of the score sum; products are rounded like the reference's elementwise ops (no FMA contraction
// of s*p*o), the accumulation order is lane-strided + butterfly.
template <int KIND, int NC>
__device__ __forceinline__ float score_partial(const Frag<KIND, NC>& s, const Frag<KIND, NC>& p, const Frag<KIND, NC>& o) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    KGEB_FOR4({
      if (KIND == DISTMULT) {
        acc += __fmul_rn(__fmul_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e));
      } else if (KIND == TRANSE_L1) {
        acc += fabsf(__fsub_rn(__fadd_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e)));
      } else if (KIND == TRANSE_L2) {
        const float r = __fsub_rn(__fadd_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e));
        acc += __fmul_rn(r, r);
      } else {
        const float sr = comp(s.v[i], _e), si = comp(s.v[NC + i], _e);
        const float pr = comp(p.v[i], _e), pi = comp(p.v[NC + i], _e);
        const float orr = comp(o.v[i], _e), oi = comp(o.v[NC + i], _e);
        const float s1 = __fmul_rn(__fmul_rn(sr, pr), orr);
        const float s2 = __fmul_rn(__fmul_rn(pr, si), oi);
        const float s3 = __fmul_rn(__fmul_rn(pi, sr), oi);
        const float s4 = __fmul_rn(__fmul_rn(pi, si), orr);
        acc += __fsub_rn(__fadd_rn(__fadd_rn(s1, s2), s3), s4);
      }
    })
  }
  return acc;
}

template <int KIND>
__device__ __forceinline__ float score_finish(float sum, float gamma) {
  if (KIND == TRANSE_L1) return gamma > 0.f ? gamma - sum : sum;
  if (KIND == TRANSE_L2) { const float n = sqrtf(sum); return gamma > 0.f ? gamma - n : n; }
  return sum;
}

// gradient fragments of coef * score wrt the three rows.  `sum` is the un-finished score sum (needed by L2).
template <int KIND, int NC>
__device__ __forceinline__ void score_grads(const Frag<KIND, NC>& s, const Frag<KIND, NC>& p, const Frag<KIND, NC>& o,
                                            float coef, float sum, float gamma, Frag<KIND, NC>& gs, Frag<KIND, NC>& gp,
                                            Frag<KIND, NC>& go) {
  float k = coef;
  if (KIND == TRANSE_L1 || KIND == TRANSE_L2) {
    k = gamma > 0.f ? -coef : coef;
    if (KIND == TRANSE_L2) { const float n = sqrtf(sum); k = n > 0.f ? k / n : 0.f; }  // TF gives NaN at r == 0; we give 0
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    KGEB_FOR4({
      if (KIND == DISTMULT) {
        const float a = comp(s.v[i], _e), b = comp(p.v[i], _e), c = comp(o.v[i], _e);
        comp(gs.v[i], _e) = k * (b * c);
        comp(gp.v[i], _e) = k * (a * c);
        comp(go.v[i], _e) = k * (a * b);
      } else if (KIND == TRANSE_L1 || KIND == TRANSE_L2) {
        const float r = __fsub_rn(__fadd_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e));
        const float g = (KIND == TRANSE_L1) ? (r > 0.f ? k : (r < 0.f ? -k : 0.f)) : k * r;
        comp(gs.v[i], _e) = g;
        comp(gp.v[i], _e) = g;
        comp(go.v[i], _e) = -g;
      } else {
        const float sr = comp(s.v[i], _e), si = comp(s.v[NC + i], _e);
        const float pr = comp(p.v[i], _e), pi = comp(p.v[NC + i], _e);
        const float orr = comp(o.v[i], _e), oi = comp(o.v[NC + i], _e);
        comp(gs.v[i], _e) = k * (pr * orr + pi * oi);
        comp(gs.v[NC + i], _e) = k * (pr * oi - pi * orr);
        comp(gp.v[i], _e) = k * (sr * orr + si * oi);
        comp(gp.v[NC + i], _e) = k * (sr * oi - si * orr);
        comp(go.v[i], _e) = k * (sr * pr - pi * si);
        comp(go.v[NC + i], _e) = k * (pr * si + pi * sr);
      }
    })
  }
}

struct ScoreArgs {
  const float* E; const float* R; int ldE, ldR; RowGeom ge, gr; float gamma;
};

// ---------------------------------------------------------------------------------------------------
// score_fwd: one warp per triple
// ---------------------------------------------------------------------------------------------------
template <int KIND, int NC>
__global__ void __launch_bounds__(KGEB_THREADS) k_score_fwd(ScoreArgs a, const int32_t* __restrict__ idx, int64_t m,
                                                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5); t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    Frag<KIND, NC> S, P, O;
    load_frag(S, a.E, s, a.ldE, a.ge, lane);
    load_frag(P, a.R, p, a.ldR, a.gr, lane);
    load_frag(O, a.E, o, a.ldE, a.ge, lane);
    const float sum = warp_sum(score_partial(S, P, O));
    if (lane == 0) out[t] = score_finish<KIND>(sum, a.gamma);
  }
}

// ---------------------------------------------------------------------------------------------------
// score_bwd: one warp per triple, gradient rows leave as vector reductions
// ---------------------------------------------------------------------------------------------------
template <int KIND, int NC>
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
    load_frag(P, a.R, p, a.ldR, a.gr, lane);
    load_frag(O, a.E, o, a.ldE, a.ge, lane);
    float sum = 0.f;
    if (KIND == TRANSE_L2) sum = warp_sum(score_partial(S, P, O));
    score_grads(S, P, O, coef, sum, a.gamma, GS, GP, GO);
    red_frag(gE, s, a.ldE, a.ge, lane, GS);
    red_frag(gR, p, a.ldR, a.gr, lane, GP);
    red_frag(gE, o, a.ldE, a.ge, lane, GO);
  }
}

// ---------------------------------------------------------------------------------------------------
// fused train forward + gradient: one warp per positive and its K negatives
// ---------------------------------------------------------------------------------------------------
struct TrainArgs {
  ScoreArgs sc;
  const int32_t* pos; int64_t B; const int32_t* neg; int K;
  int loss_kind; float margin; float loss_weight; int corrupt_mode; uint64_t seed, offset; uint32_t N;
  float* loss; float* pos_scores; float* neg_scores; int32_t* neg_out;
  float* gE; float* gR; uint8_t* tE; uint8_t* tR;
  double* partials; unsigned int* ticket;
};

template <int KIND, int NC>
__global__ void __launch_bounds__(KGEB_THREADS) k_train_fwd_bwd(TrainArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const bool pairwise = a.loss_kind >= KGEB_PAIRWISE_HINGE;
  const double cntP = (double)a.B, cntN = (double)a.B * (double)a.K;
  // coefficient scales: pointwise positives /B, everything else /(B*K)
  const float scaleP = (float)((double)a.loss_weight / cntP), scaleN = (float)((double)a.loss_weight / cntN);
  double lp = 0.0, ln = 0.0;
  const ScoreArgs& sc = a.sc;
  for (int64_t g = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5); g < a.B; g += nw) {
    const int s = __ldg(a.pos + 3 * g), p = __ldg(a.pos + 3 * g + 1), o = __ldg(a.pos + 3 * g + 2);
    Frag<KIND, NC> S, P, O;
    load_frag(S, sc.E, s, sc.ldE, sc.ge, lane);
    load_frag(P, sc.R, p, sc.ldR, sc.gr, lane);
    load_frag(O, sc.E, o, sc.ldE, sc.ge, lane);
    const float tsum = warp_sum(score_partial(S, P, O));
    const float t = score_finish<KIND>(tsum, sc.gamma);
    if (lane == 0) {
      if (a.pos_scores) a.pos_scores[g] = t;
      a.tE[s] = 1; a.tE[o] = 1; a.tR[p] = 1;
    }
    float cpos = 0.f;  // dL/dt, already scaled
    if (!pairwise) {
      float val, der;
      pointwise_terms(a.loss_kind, a.margin, t, true, val, der);
      lp += (double)val;
      cpos = der * scaleP;
    }
    // gradient that accumulates on the positive's own rows (negatives that share a row add here, one RED at the end)
    Frag<KIND, NC> AS, AP, AO;
    zero_frag(AS); zero_frag(AP); zero_frag(AO);
    for (int j = 0; j < a.K; ++j) {
      const int64_t gi = g * (int64_t)a.K + j;
      int ns, np, no;
      if (a.neg) {
        ns = __ldg(a.neg + 3 * gi); np = __ldg(a.neg + 3 * gi + 1); no = __ldg(a.neg + 3 * gi + 2);
      } else {
        np = p;
        philox_corrupt(s, o, (uint64_t)gi, a.seed, a.offset, a.N, a.corrupt_mode, ns, no);
      }
      Frag<KIND, NC> Sn, Pn, On;
      if (ns != s) load_frag(Sn, sc.E, ns, sc.ldE, sc.ge, lane); else Sn = S;
      if (np != p) load_frag(Pn, sc.R, np, sc.ldR, sc.gr, lane); else Pn = P;
      if (no != o) load_frag(On, sc.E, no, sc.ldE, sc.ge, lane); else On = O;
      const float fsum = warp_sum(score_partial(Sn, Pn, On));
      const float f = score_finish<KIND>(fsum, sc.gamma);
      float cneg;
      if (pairwise) {
        float val, A;
        pairwise_terms(a.loss_kind, a.margin, f - t, val, A);
        ln += (double)val;
        cneg = A * scaleN;
        cpos -= cneg;
      } else {
        float val, der;
        pointwise_terms(a.loss_kind, a.margin, f, false, val, der);
        ln += (double)val;
        cneg = der * scaleN;
      }
      if (lane == 0) {
        if (a.neg_scores) a.neg_scores[gi] = f;
        if (a.neg_out) { a.neg_out[3 * gi] = ns; a.neg_out[3 * gi + 1] = np; a.neg_out[3 * gi + 2] = no; }
        a.tE[ns] = 1; a.tE[no] = 1; a.tR[np] = 1;
      }
      if (cneg != 0.f) {
        Frag<KIND, NC> GS, GP, GO;
        score_grads(Sn, Pn, On, cneg, fsum, sc.gamma, GS, GP, GO);
        if (ns != s) red_frag(a.gE, ns, sc.ldE, sc.ge, lane, GS); else add_frag(AS, GS);
        if (np != p) red_frag(a.gR, np, sc.ldR, sc.gr, lane, GP); else add_frag(AP, GP);
        if (no != o) red_frag(a.gE, no, sc.ldE, sc.ge, lane, GO); else add_frag(AO, GO);
      }
    }
    {
      Frag<KIND, NC> GS, GP, GO;
      if (cpos != 0.f) {
        score_grads(S, P, O, cpos, tsum, sc.gamma, GS, GP, GO);
        add_frag(AS, GS); add_frag(AP, GP); add_frag(AO, GO);
      }
      red_frag(a.gE, s, sc.ldE, sc.ge, lane, AS);
      red_frag(a.gR, p, sc.ldR, sc.gr, lane, AP);
      red_frag(a.gE, o, sc.ldE, sc.ge, lane, AO);
    }
  }
  // every lane carries the same lp/ln (warp-uniform values): count each warp once
  if (lane != 0) { lp = 0.0; ln = 0.0; }
  const double w = (double)a.loss_weight;
  block_finalize_loss(lp, ln, a.partials, a.ticket, pairwise ? 0.0 : w / cntP, w / cntN, a.loss);
}

// ---------------------------------------------------------------------------------------------------
// fused train forward + gradient, shared-memory ring version (the default path).
//
// The register version above is latency bound (profiles/r01_v1_summary.md: 8.9 warps stalled on global loads per
// issued instruction, L2 15 % busy).  Here every gathered row travels global -> shared memory as ONE bulk async copy
// (cp.async.bulk, SASS UBLKCP) that completes on an mbarrier, and each warp owns a private ring of D stages
// (one stage = the rows of one negative), so D negatives per warp are in flight without costing registers.  The warp is
// its own producer and consumer: no block-level synchronisation exists in the main loop.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "KGEB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KGEB_DONE;\n\t"
      "bra KGEB_WAIT;\n\t"
      "KGEB_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int KIND, int NC>
__device__ __forceinline__ void load_frag_smem(Frag<KIND, NC>& f, const float* row, RowGeom g, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    const bool ok = c < g.nch;
    f.v[i] = ok ? r4[c] : f4zero();
    if (KIND == COMPLEX) f.v[NC + i] = ok ? r4[g.half4 + c] : f4zero();
  }
}

template <int KIND, int NC>
__global__ void __launch_bounds__(KGEB_THREADS) k_train_ring(TrainArgs a, int D, int RB) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t region = (size_t)(3 + 3 * D) * RB;
  unsigned char* base = smem_raw + (size_t)w * region;
  float* posS = reinterpret_cast<float*>(base);
  float* posP = reinterpret_cast<float*>(base + RB);
  float* posO = reinterpret_cast<float*>(base + 2 * (size_t)RB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)KGEB_WARPS_PER_BLOCK * region) + w * (D + 1);
  if (lane == 0)
    for (int i = 0; i <= D; ++i) mbar_init(bars + i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const ScoreArgs& sc = a.sc;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const bool pairwise = a.loss_kind >= KGEB_PAIRWISE_HINGE;
  const double cntP = (double)a.B, cntN = (double)a.B * (double)a.K;
  const float scaleP = (float)((double)a.loss_weight / cntP), scaleN = (float)((double)a.loss_weight / cntN);
  const int K = a.K;
  double lp = 0.0, ln = 0.0;
  uint32_t phases = 0;  // bit 0: positive barrier, bit 1+s: stage s

  for (int64_t g = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; g < a.B; g += nw) {
    const int s = __ldg(a.pos + 3 * g), p = __ldg(a.pos + 3 * g + 1), o = __ldg(a.pos + 3 * g + 2);
    __syncwarp();  // every lane is done reading the previous group's buffers
    if (lane == 0) {
      mbar_expect_tx(bars, 3u * RB);
      bulk_g2s(posS, sc.E + (int64_t)s * sc.ldE, RB, bars);
      bulk_g2s(posP, sc.R + (int64_t)p * sc.ldR, RB, bars);
      bulk_g2s(posO, sc.E + (int64_t)o * sc.ldE, RB, bars);
    }
    // ids of this group's negatives: lane l owns negative (chunk*32 + l), drawn by Philox or read from the caller's
    // index tensor.  Two chunks are held (cur, ahead) because issue() runs up to D <= 32 negatives ahead of consumption.
    auto compute_chunk = [&](int chunk, int& ns, int& np, int& no) {
      const int j = chunk * 32 + lane;
      ns = s; np = p; no = o;
      if (j < K) {
        const int64_t gi = g * (int64_t)K + j;
        if (a.neg) {
          ns = __ldg(a.neg + 3 * gi); np = __ldg(a.neg + 3 * gi + 1); no = __ldg(a.neg + 3 * gi + 2);
        } else {
          philox_corrupt(s, o, (uint64_t)gi, a.seed, a.offset, a.N, a.corrupt_mode, ns, no);
        }
        if (a.neg_out) { a.neg_out[3 * gi] = ns; a.neg_out[3 * gi + 1] = np; a.neg_out[3 * gi + 2] = no; }
        a.tE[ns] = 1; a.tE[no] = 1; a.tR[np] = 1;
      }
    };
    int cur_chunk = 0, my_ns, my_np, my_no, ah_ns = s, ah_np = p, ah_no = o;
    compute_chunk(0, my_ns, my_np, my_no);
    if (K > 32) compute_chunk(1, ah_ns, ah_np, ah_no);
    auto ids_of = [&](int j, int& ns, int& np, int& no) {
      const bool in_cur = (j >> 5) == cur_chunk;  // warp-uniform; otherwise j lies in chunk cur_chunk + 1
      const int l = j & 31;
      ns = __shfl_sync(0xffffffffu, in_cur ? my_ns : ah_ns, l);
      np = __shfl_sync(0xffffffffu, in_cur ? my_np : ah_np, l);
      no = __shfl_sync(0xffffffffu, in_cur ? my_no : ah_no, l);
    };
    // issue the rows of negative j into stage j % D (rows shared with the positive are not copied again)
    auto issue = [&](int j) {
      int ns, np, no;
      ids_of(j, ns, np, no);
      if (lane == 0) {
        const int st = j % D;
        unsigned char* sb = base + (size_t)(3 + 3 * st) * RB;
        const uint32_t nrows = (ns != s) + (np != p) + (no != o);
        if (nrows) {
          mbar_expect_tx(bars + 1 + st, nrows * (uint32_t)RB);
          if (ns != s) bulk_g2s(sb, sc.E + (int64_t)ns * sc.ldE, RB, bars + 1 + st);
          if (np != p) bulk_g2s(sb + RB, sc.R + (int64_t)np * sc.ldR, RB, bars + 1 + st);
          if (no != o) bulk_g2s(sb + 2 * (size_t)RB, sc.E + (int64_t)no * sc.ldE, RB, bars + 1 + st);
        }
      }
    };
    const int prime = K < D ? K : D;
    for (int j = 0; j < prime; ++j) issue(j);

    // positive score
    mbar_wait(bars, phases & 1u);
    phases ^= 1u;
    float tsum, t;
    {
      Frag<KIND, NC> S, P, O;
      load_frag_smem(S, posS, sc.ge, lane);
      load_frag_smem(P, posP, sc.gr, lane);
      load_frag_smem(O, posO, sc.ge, lane);
      tsum = warp_sum(score_partial(S, P, O));
      t = score_finish<KIND>(tsum, sc.gamma);
    }
    if (lane == 0) {
      if (a.pos_scores) a.pos_scores[g] = t;
      a.tE[s] = 1; a.tE[o] = 1; a.tR[p] = 1;
    }
    float cpos = 0.f;
    if (!pairwise) {
      float val, der;
      pointwise_terms(a.loss_kind, a.margin, t, true, val, der);
      lp += (double)val;
      cpos = der * scaleP;
    }
    Frag<KIND, NC> AS, AP, AO;
    zero_frag(AS); zero_frag(AP); zero_frag(AO);
    bool usedS = false, usedP = false, usedO = false;

    for (int j = 0; j < K; ++j) {
      if ((j & 31) == 0 && j > 0) {  // crossed into the "ahead" chunk: rotate, then draw the chunk after it
        cur_chunk = j >> 5; my_ns = ah_ns; my_np = ah_np; my_no = ah_no;
        if ((cur_chunk + 1) * 32 < K) compute_chunk(cur_chunk + 1, ah_ns, ah_np, ah_no);
      }
      int ns, np, no;
      ids_of(j, ns, np, no);
      const int st = j % D;
      const unsigned char* sb = base + (size_t)(3 + 3 * st) * RB;
      if ((ns != s) | (np != p) | (no != o)) {
        mbar_wait(bars + 1 + st, (phases >> (1 + st)) & 1u);
        phases ^= 1u << (1 + st);
      }
      const float* rS = ns != s ? reinterpret_cast<const float*>(sb) : posS;
      const float* rP = np != p ? reinterpret_cast<const float*>(sb + RB) : posP;
      const float* rO = no != o ? reinterpret_cast<const float*>(sb + 2 * (size_t)RB) : posO;
      Frag<KIND, NC> Sn, Pn, On;
      load_frag_smem(Sn, rS, sc.ge, lane);
      load_frag_smem(Pn, rP, sc.gr, lane);
      load_frag_smem(On, rO, sc.ge, lane);
      const float fsum = warp_sum(score_partial(Sn, Pn, On));
      // the butterfly consumed every lane's loads of this stage, so the stage can be refilled for negative j + D
      if (j + D < K) issue(j + D);
      const float f = score_finish<KIND>(fsum, sc.gamma);
      float cneg;
      if (pairwise) {
        float val, A;
        pairwise_terms(a.loss_kind, a.margin, f - t, val, A);
        ln += (double)val;
        cneg = A * scaleN;
        cpos -= cneg;
      } else {
        float val, der;
        pointwise_terms(a.loss_kind, a.margin, f, false, val, der);
        ln += (double)val;
        cneg = der * scaleN;
      }
      if (lane == 0 && a.neg_scores) a.neg_scores[g * (int64_t)K + j] = f;
      if (cneg != 0.f) {
        Frag<KIND, NC> GS, GP, GO;
        score_grads(Sn, Pn, On, cneg, fsum, sc.gamma, GS, GP, GO);
        if (ns != s) red_frag(a.gE, ns, sc.ldE, sc.ge, lane, GS); else { add_frag(AS, GS); usedS = true; }
        if (np != p) red_frag(a.gR, np, sc.ldR, sc.gr, lane, GP); else { add_frag(AP, GP); usedP = true; }
        if (no != o) red_frag(a.gE, no, sc.ldE, sc.ge, lane, GO); else { add_frag(AO, GO); usedO = true; }
      }
    }
    if (cpos != 0.f) {
      Frag<KIND, NC> S, P, O, GS, GP, GO;
      load_frag_smem(S, posS, sc.ge, lane);
      load_frag_smem(P, posP, sc.gr, lane);
      load_frag_smem(O, posO, sc.ge, lane);
      score_grads(S, P, O, cpos, tsum, sc.gamma, GS, GP, GO);
      add_frag(AS, GS); add_frag(AP, GP); add_frag(AO, GO);
      usedS = usedP = usedO = true;
    }
    if (usedS) red_frag(a.gE, s, sc.ldE, sc.ge, lane, AS);
    if (usedP) red_frag(a.gR, p, sc.ldR, sc.gr, lane, AP);
    if (usedO) red_frag(a.gE, o, sc.ldE, sc.ge, lane, AO);
  }
  if (lane != 0) { lp = 0.0; ln = 0.0; }
  const double wgt = (double)a.loss_weight;
  block_finalize_loss(lp, ln, a.partials, a.ticket, pairwise ? 0.0 : wgt / cntP, wgt / cntN, a.loss);
}

// ---------------------------------------------------------------------------------------------------
// HolE: f = sum_k p_k sum_i s_i o_{(i+k) mod d}; rows staged in shared memory, O(d^2/32) per lane
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hole_stage(float* dst, const float* __restrict__ src, int d, int lane) {
  for (int i = lane; i < d; i += 32) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(KGEB_THREADS) k_hole_fwd(ScoreArgs a, int d, const int32_t* __restrict__ idx, int64_t m,
                                                           float* __restrict__ out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* ss = smem + (size_t)w * 2 * d;
  float* so = ss + d;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    __syncwarp();
    hole_stage(ss, a.E + (int64_t)s * a.ldE, d, lane);
    hole_stage(so, a.E + (int64_t)o * a.ldE, d, lane);
    __syncwarp();
    const float* pr = a.R + (int64_t)p * a.ldR;
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) {
      float x = 0.f;
      int j = k;
      for (int i = 0; i < d; ++i) {
        x = fmaf(ss[i], so[j], x);
        if (++j == d) j = 0;
      }
      acc = fmaf(__ldg(pr + k), x, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[t] = acc;
  }
}

__global__ void __launch_bounds__(KGEB_THREADS) k_hole_bwd(ScoreArgs a, int d, const int32_t* __restrict__ idx, int64_t m,
                                                           const float* __restrict__ dscore, float* gE, float* gR,
                                                           uint8_t* tE, uint8_t* tR) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* ss = smem + (size_t)w * 3 * d;
  float* sp = ss + d;
  float* so = sp + d;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t t = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; t < m; t += nw) {
    const int s = __ldg(idx + 3 * t), p = __ldg(idx + 3 * t + 1), o = __ldg(idx + 3 * t + 2);
    const float coef = __ldg(dscore + t);
    if (lane == 0) { tE[s] = 1; tE[o] = 1; tR[p] = 1; }
    if (coef == 0.f) continue;
    __syncwarp();
    hole_stage(ss, a.E + (int64_t)s * a.ldE, d, lane);
    hole_stage(sp, a.R + (int64_t)p * a.ldR, d, lane);
    hole_stage(so, a.E + (int64_t)o * a.ldE, d, lane);
    __syncwarp();
    float* gs = gE + (int64_t)s * a.ldE;
    float* gp = gR + (int64_t)p * a.ldR;
    float* go = gE + (int64_t)o * a.ldE;
    for (int k = lane; k < d; k += 32) {
      // dp_k = sum_i s_i o_{i+k};  ds_k = sum_j p_j o_{k+j};  do_k = sum_j p_j s_{k-j}
      float xp = 0.f, xs = 0.f, xo = 0.f;
      int jf = k, jb = k;
      for (int i = 0; i < d; ++i) {
        const float ov = so[jf];
        xp = fmaf(ss[i], ov, xp);
        xs = fmaf(sp[i], ov, xs);
        xo = fmaf(sp[i], ss[jb], xo);
        if (++jf == d) jf = 0;
        if (--jb < 0) jb = d - 1;
      }
      atomicAdd(gp + k, coef * xp);
      atomicAdd(gs + k, coef * xs);
      atomicAdd(go + k, coef * xo);
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

// ---------------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------------
int check_tables(kgeb_ctx* ctx, const void* E, const void* R) {
  KGEB_REQUIRE(ctx, E && R, KGEB_ERR_INVALID, "null table pointer");
  if (ctx->kind == KGEB_HOLE) return KGEB_OK;
  KGEB_REQUIRE(ctx, kgeb_aligned16(E) && kgeb_aligned16(R), KGEB_ERR_ALIGNMENT, "table pointers must be 16-byte aligned");
  KGEB_REQUIRE(ctx, ctx->ldE % 4 == 0 && ctx->ldR % 4 == 0, KGEB_ERR_ALIGNMENT, "row strides must be multiples of 4 floats");
  KGEB_REQUIRE(ctx, ctx->we == ctx->wr, KGEB_ERR_UNSUPPORTED, "entity and relation widths must match for this score kind");
  if (ctx->kind == KGEB_COMPLEX) {
    KGEB_REQUIRE(ctx, ctx->we % 8 == 0, KGEB_ERR_ALIGNMENT, "ComplEx needs dim % 4 == 0 (stored width % 8 == 0)");
  } else {
    KGEB_REQUIRE(ctx, ctx->we % 4 == 0, KGEB_ERR_ALIGNMENT, "row width must be a multiple of 4 floats");
  }
  return KGEB_OK;
}

ScoreArgs make_score_args(const kgeb_ctx* ctx, const float* E, const float* R, int* nc_out) {
  ScoreArgs a;
  a.E = E; a.R = R; a.ldE = ctx->ldE; a.ldR = ctx->ldR; a.gamma = ctx->gamma;
  if (ctx->kind == KGEB_COMPLEX) {
    a.ge.nch = ctx->we / 8; a.ge.half4 = ctx->we / 8;
  } else {
    a.ge.nch = ctx->we / 4; a.ge.half4 = 0;
  }
  a.gr = a.ge;
  *nc_out = (a.ge.nch + 31) / 32;
  return a;
}

int grid_for(const kgeb_ctx* ctx, int64_t warps_needed, int blocks_per_sm) {
  int64_t blocks = (warps_needed + KGEB_WARPS_PER_BLOCK - 1) / KGEB_WARPS_PER_BLOCK;
  const int64_t cap = (int64_t)ctx->num_sms * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// NC (chunks per lane) is a compile-time constant so rows live in registers: widths up to 512 floats
// (ComplEx: dim up to 256) are covered; wider rows are reported, never silently slow-pathed.
#define KGEB_DISPATCH_NC(NCV, ...)                                        \
  switch (NCV) {                                                          \
    case 1: { constexpr int NC = 1; __VA_ARGS__; } break;                 \
    case 2: { constexpr int NC = 2; __VA_ARGS__; } break;                 \
    case 3: { constexpr int NC = 3; __VA_ARGS__; } break;                 \
    case 4: { constexpr int NC = 4; __VA_ARGS__; } break;                 \
    default: return kgeb_fail(ctx, KGEB_ERR_UNSUPPORTED, "row too wide for the register-resident kernels (max 512 floats; ComplEx dim 256)"); \
  }
#define KGEB_DISPATCH_KIND(KINDV, ...)                                                   \
  switch (KINDV) {                                                                       \
    case KGEB_DISTMULT: { constexpr int KIND = KGEB_DISTMULT; __VA_ARGS__; } break;      \
    case KGEB_TRANSE_L1: { constexpr int KIND = KGEB_TRANSE_L1; __VA_ARGS__; } break;    \
    case KGEB_TRANSE_L2: { constexpr int KIND = KGEB_TRANSE_L2; __VA_ARGS__; } break;    \
    case KGEB_COMPLEX: { constexpr int KIND = KGEB_COMPLEX; __VA_ARGS__; } break;        \
    default: return kgeb_fail(ctx, KGEB_ERR_INVALID, "unknown score kind");          \
  }

}  // namespace

extern "C" int kgeb_score_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores,
                              kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, m >= 0, KGEB_ERR_INVALID, "score_fwd: negative batch");
  if (m == 0) return KGEB_OK;  /* empty batch: nothing to read, pointers may be null */
  KGEB_REQUIRE(ctx, idx && scores, KGEB_ERR_INVALID, "score_fwd: bad arguments");
  if (int rc = check_tables(ctx, E, R)) return rc;
  if (m == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int nc;
  ScoreArgs a = make_score_args(ctx, E, R, &nc);
  const int grid = grid_for(ctx, m, 8);
  if (ctx->kind == KGEB_HOLE) {
    KGEB_REQUIRE(ctx, ctx->we == ctx->wr, KGEB_ERR_UNSUPPORTED, "HolE needs e_dim == r_dim");
    const size_t sm = (size_t)KGEB_WARPS_PER_BLOCK * 2 * ctx->we * sizeof(float);
    KGEB_REQUIRE(ctx, sm <= 200 * 1024, KGEB_ERR_UNSUPPORTED, "HolE dim too large for shared memory staging");
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fwd<<<grid, KGEB_THREADS, sm, st>>>(a, ctx->we, idx, m, scores);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_fwd<KIND, NC><<<grid, KGEB_THREADS, 0, st>>>(a, idx, m, scores))));
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
  if (int rc = check_tables(ctx, E, R)) return rc;
  KGEB_REQUIRE(ctx, ctx->kind == KGEB_HOLE || (kgeb_aligned16(gE) && kgeb_aligned16(gR)), KGEB_ERR_ALIGNMENT,
               "gradient buffers must be 16-byte aligned");
  if (m == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int nc;
  ScoreArgs a = make_score_args(ctx, E, R, &nc);
  const int grid = grid_for(ctx, m, 8);
  if (ctx->kind == KGEB_HOLE) {
    const size_t sm = (size_t)KGEB_WARPS_PER_BLOCK * 3 * ctx->we * sizeof(float);
    KGEB_REQUIRE(ctx, sm <= 200 * 1024, KGEB_ERR_UNSUPPORTED, "HolE dim too large for shared memory staging");
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_bwd<<<grid, KGEB_THREADS, sm, st>>>(a, ctx->we, idx, m, dscore, gE, gR, touchedE, touchedR);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_score_bwd<KIND, NC><<<grid, KGEB_THREADS, 0, st>>>(
                                                         a, idx, m, dscore, gE, gR, touchedE, touchedR))));
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

extern "C" int kgeb_train_fwd_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* pos, int64_t B,
                                  const int32_t* neg, int K, int loss_kind, float margin, float loss_weight,
                                  int corrupt_mode, uint64_t seed, uint64_t offset, float* loss, float* pos_scores,
                                  float* neg_scores, int32_t* neg_out, float* gE, float* gR, uint8_t* touchedE,
                                  uint8_t* touchedR, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, pos && loss && gE && gR && touchedE && touchedR && B > 0 && K > 0, KGEB_ERR_INVALID,
               "train_fwd_bwd: bad arguments");
  KGEB_REQUIRE(ctx, loss_kind >= 0 && loss_kind <= KGEB_PAIRWISE_SQUARE, KGEB_ERR_INVALID, "unknown loss kind");
  KGEB_REQUIRE(ctx, corrupt_mode >= 0 && corrupt_mode <= 3, KGEB_ERR_INVALID, "unknown corrupt mode");
  KGEB_REQUIRE(ctx, ctx->kind != KGEB_HOLE, KGEB_ERR_UNSUPPORTED,
               "HolE uses the unfused sequence score_fwd -> loss_fwd_bwd -> score_bwd");
  if (int rc = check_tables(ctx, E, R)) return rc;
  KGEB_REQUIRE(ctx, kgeb_aligned16(gE) && kgeb_aligned16(gR), KGEB_ERR_ALIGNMENT, "gradient buffers must be 16-byte aligned");
  int nc;
  TrainArgs a;
  a.sc = make_score_args(ctx, E, R, &nc);
  a.pos = pos; a.B = B; a.neg = neg; a.K = K; a.loss_kind = loss_kind; a.margin = margin; a.loss_weight = loss_weight;
  a.corrupt_mode = corrupt_mode; a.seed = seed; a.offset = offset; a.N = (uint32_t)ctx->N;
  a.loss = loss; a.pos_scores = pos_scores; a.neg_scores = neg_scores; a.neg_out = neg_out;
  a.gE = gE; a.gR = gR; a.tE = touchedE; a.tR = touchedR; a.partials = ctx->d_partials; a.ticket = ctx->d_ticket;
  cudaStream_t st = (cudaStream_t)stream;
  // shared-memory ring version: D stages of 3 rows per warp + the positive's 3 rows; 2 CTAs/SM when it fits
  const int RB = ctx->we * (int)sizeof(float);
  auto stages_for = [&](size_t budget) { return (int)(((long long)(budget / KGEB_WARPS_PER_BLOCK) - 3LL * RB) / (3LL * RB)); };
  int ctas = 2, D = stages_for(112 * 1024);
  if (D < 2) { ctas = 1; D = stages_for(224 * 1024); }
  if (D > 8) D = 8;
  const char* force = getenv("KGEB_TRAIN_KERNEL");
  const bool use_ring = D >= 1 && !(force && strcmp(force, "reg") == 0);
  if (use_ring) {
    const size_t smem = (size_t)KGEB_WARPS_PER_BLOCK * (3 + 3 * D) * RB + (size_t)KGEB_WARPS_PER_BLOCK * (D + 1) * sizeof(uint64_t);
    int grid = grid_for(ctx, B, ctas);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, {
      KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_train_ring<KIND, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_train_ring<KIND, NC><<<grid, KGEB_THREADS, smem, st>>>(a, D, RB);
    }));
  } else {
    int grid = grid_for(ctx, B, 4);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_train_fwd_bwd<KIND, NC><<<grid, KGEB_THREADS, 0, st>>>(a))));
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
