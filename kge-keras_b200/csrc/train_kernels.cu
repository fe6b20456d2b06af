// train_kernels.cu -- K1: fused train step forward + gradient (gather, Philox negatives, score, loss, gradient scatter).
#include "frag.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// fused train forward + gradient: one warp per positive and its K negatives
// (register version: kept for A/B measurements, KGEB_TRAIN_KERNEL=reg; latency bound, see profiles/r01_v1_summary.md)
// ---------------------------------------------------------------------------------------------------

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
// Fragment <-> shared memory / global reductions with the chunk predicate hoisted: only the last chunk of a row can
// be partial (NC = ceil(nch/32)), so one precomputed bool replaces a compare per chunk per access.
template <int KIND, int NC>
__device__ __forceinline__ void lds_frag(Frag<KIND, NC>& f, const float* row, int half4, bool okl, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(row) + lane;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const bool k = (i < NC - 1) || okl;
    f.v[i] = k ? r4[32 * i] : f4zero();
    if (KIND == COMPLEX) f.v[NC + i] = k ? r4[half4 + 32 * i] : f4zero();
  }
}
// rowl = row base already advanced by 4*lane floats
template <int KIND, int NC>
__device__ __forceinline__ void red_frag_l(float* rowl, int half4, bool okl, const Frag<KIND, NC>& f) {
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    if ((i < NC - 1) || okl) {
      red_add4(rowl + 128 * i, f.v[i]);
      if (KIND == COMPLEX) red_add4(rowl + 4 * half4 + 128 * i, f.v[NC + i]);
    }
  }
}

// MODE 0: index-driven negatives (any ids, any relation; rows equal to the positive's are shared, not re-copied)
// MODE 1: in-kernel Philox, both sides corrupted (the reference's scheme): a stage is 2 rows, relation shared
// MODE 2: in-kernel Philox, one side corrupted: a stage is 1 row, relation and the other entity shared
template <int KIND, int NC, int MODE>
__global__ void __launch_bounds__(KGEB_THREADS) k_train_ring(TrainArgs a, int D, int RB) {
  constexpr int SR = MODE == 0 ? 3 : (MODE == 1 ? 2 : 1);  // rows per stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t region = (size_t)(3 + SR * D) * RB;
  unsigned char* base = smem_raw + (size_t)w * region;
  const float* posS = reinterpret_cast<const float*>(base);
  const float* posP = reinterpret_cast<const float*>(base + RB);
  const float* posO = reinterpret_cast<const float*>(base + 2 * (size_t)RB);
  unsigned char* ring = base + 3 * (size_t)RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)KGEB_WARPS_PER_BLOCK * region) + w * (D + 1);
  if (lane == 0)
    for (int i = 0; i <= D; ++i) mbar_init(bars + i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const ScoreArgs& sc = a.sc;
  const float* __restrict__ E = sc.E;
  const float* __restrict__ Rl = sc.R;
  const int ldE = sc.ldE, ldR = sc.ldR, half4 = sc.ge.half4;
  const float gamma = sc.gamma;
  const bool okl = lane + 32 * (NC - 1) < sc.ge.nch;
  float* const gEl = a.gE + 4 * lane;
  float* const gRl = a.gR + 4 * lane;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const int loss_kind = a.loss_kind;
  const float margin = a.margin;
  const bool pairwise = loss_kind >= KGEB_PAIRWISE_HINGE;
  const double cntP = (double)a.B, cntN = (double)a.B * (double)a.K;
  const float scaleP = (float)((double)a.loss_weight / cntP), scaleN = (float)((double)a.loss_weight / cntN);
  const int K = a.K;
  const uint32_t uRB = (uint32_t)RB;
  double lp = 0.0, ln = 0.0;
  uint32_t phases = 0;  // bit 0: positive barrier, bit 1+s: stage s

  for (int64_t g = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + w; g < a.B; g += nw) {
    const int s = __ldg(a.pos + 3 * g), p = __ldg(a.pos + 3 * g + 1), o = __ldg(a.pos + 3 * g + 2);
    __syncwarp();  // every lane is done reading the previous group's buffers
    if (lane == 0) {
      mbar_expect_tx(bars, 3u * uRB);
      bulk_g2s(base, E + (int64_t)s * ldE, uRB, bars);
      bulk_g2s(base + RB, Rl + (int64_t)p * ldR, uRB, bars);
      bulk_g2s(base + 2 * (size_t)RB, E + (int64_t)o * ldE, uRB, bars);
    }
    // ids of this group's negatives: lane l owns negative (chunk*32 + l).  Two chunks are held (cur, ahead) because
    // issue() runs up to D <= 32 negatives ahead of consumption.
    auto compute_chunk = [&](int chunk, int& ns, int& np, int& no, int& hs) {
      const int j = chunk * 32 + lane;
      ns = s; np = p; no = o; hs = 0;
      if (j < K) {
        const int64_t gi = g * (int64_t)K + j;
        if (MODE == 0) {
          ns = __ldg(a.neg + 3 * gi); np = __ldg(a.neg + 3 * gi + 1); no = __ldg(a.neg + 3 * gi + 2);
        } else {
          bool h;
          philox_corrupt(s, o, (uint64_t)gi, a.seed, a.offset, a.N, a.corrupt_mode, ns, no, h);
          hs = h ? 1 : 0;
        }
        if (a.neg_out) { a.neg_out[3 * gi] = ns; a.neg_out[3 * gi + 1] = np; a.neg_out[3 * gi + 2] = no; }
        a.tE[ns] = 1; a.tE[no] = 1;
        if (MODE == 0) a.tR[np] = 1;
      }
    };
    int cur_chunk = 0, my_ns, my_np, my_no, my_hs, ah_ns = s, ah_np = p, ah_no = o, ah_hs = 0;
    compute_chunk(0, my_ns, my_np, my_no, my_hs);
    if (K > 32) compute_chunk(1, ah_ns, ah_np, ah_no, ah_hs);
    auto ids_of = [&](int j, int& ns, int& np, int& no, int& hs) {
      const bool in_cur = (j >> 5) == cur_chunk;  // warp-uniform; otherwise j lies in chunk cur_chunk + 1
      const int l = j & 31;
      ns = __shfl_sync(0xffffffffu, in_cur ? my_ns : ah_ns, l);
      no = __shfl_sync(0xffffffffu, in_cur ? my_no : ah_no, l);
      np = MODE == 0 ? __shfl_sync(0xffffffffu, in_cur ? my_np : ah_np, l) : p;
      hs = MODE == 2 ? __shfl_sync(0xffffffffu, in_cur ? my_hs : ah_hs, l) : 0;
    };
    auto issue = [&](int j) {
      int ns, np, no, hs;
      ids_of(j, ns, np, no, hs);
      if (lane == 0) {
        const int st = j % D;
        unsigned char* sb = ring + (size_t)(SR * st) * RB;
        uint64_t* bar = bars + 1 + st;
        if (MODE == 1) {
          mbar_expect_tx(bar, 2u * uRB);
          bulk_g2s(sb, E + (int64_t)ns * ldE, uRB, bar);
          bulk_g2s(sb + RB, E + (int64_t)no * ldE, uRB, bar);
        } else if (MODE == 2) {
          mbar_expect_tx(bar, uRB);
          bulk_g2s(sb, E + (int64_t)(hs ? ns : no) * ldE, uRB, bar);
        } else {
          const uint32_t nrows = (ns != s) + (np != p) + (no != o);
          if (nrows) {
            mbar_expect_tx(bar, nrows * uRB);
            if (ns != s) bulk_g2s(sb, E + (int64_t)ns * ldE, uRB, bar);
            if (np != p) bulk_g2s(sb + RB, Rl + (int64_t)np * ldR, uRB, bar);
            if (no != o) bulk_g2s(sb + 2 * (size_t)RB, E + (int64_t)no * ldE, uRB, bar);
          }
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
      lds_frag(S, posS, half4, okl, lane);
      lds_frag(P, posP, half4, okl, lane);
      lds_frag(O, posO, half4, okl, lane);
      tsum = warp_sum(score_partial(S, P, O));
      t = score_finish<KIND>(tsum, gamma);
    }
    if (lane == 0) {
      if (a.pos_scores) a.pos_scores[g] = t;
      a.tE[s] = 1; a.tE[o] = 1; a.tR[p] = 1;
    }
    float cpos = 0.f;
    if (!pairwise) {
      float val, der;
      pointwise_terms(loss_kind, margin, t, true, val, der);
      lp += (double)val;
      cpos = der * scaleP;
    }
    Frag<KIND, NC> AS, AP, AO;  // gradient accumulating on the positive's own rows (AS/AO unused in MODE 1)
    zero_frag(AS); zero_frag(AP); zero_frag(AO);
    bool usedS = false, usedP = false, usedO = false;

    for (int j = 0; j < K; ++j) {
      if ((j & 31) == 0 && j > 0) {  // crossed into the "ahead" chunk: rotate, then draw the chunk after it
        cur_chunk = j >> 5; my_ns = ah_ns; my_np = ah_np; my_no = ah_no; my_hs = ah_hs;
        if ((cur_chunk + 1) * 32 < K) compute_chunk(cur_chunk + 1, ah_ns, ah_np, ah_no, ah_hs);
      }
      int ns, np, no, hs;
      ids_of(j, ns, np, no, hs);
      const int st = j % D;
      const unsigned char* sb = ring + (size_t)(SR * st) * RB;
      const float *rS, *rP, *rO;
      bool shS, shP, shO;  // row shared with the positive (warp-uniform)
      if (MODE == 1) { shS = false; shP = true; shO = false; }
      else if (MODE == 2) { shS = !hs; shP = true; shO = hs != 0; }
      else { shS = ns == s; shP = np == p; shO = no == o; }
      if (MODE != 0 || !(shS && shP && shO)) {
        mbar_wait(bars + 1 + st, (phases >> (1 + st)) & 1u);
        phases ^= 1u << (1 + st);
      }
      if (MODE == 1) {
        rS = reinterpret_cast<const float*>(sb); rP = posP; rO = reinterpret_cast<const float*>(sb + RB);
      } else if (MODE == 2) {
        rS = hs ? reinterpret_cast<const float*>(sb) : posS; rP = posP; rO = hs ? posO : reinterpret_cast<const float*>(sb);
      } else {
        rS = shS ? posS : reinterpret_cast<const float*>(sb);
        rP = shP ? posP : reinterpret_cast<const float*>(sb + RB);
        rO = shO ? posO : reinterpret_cast<const float*>(sb + 2 * (size_t)RB);
      }
      Frag<KIND, NC> Sn, Pn, On;
      lds_frag(Sn, rS, half4, okl, lane);
      lds_frag(Pn, rP, half4, okl, lane);
      lds_frag(On, rO, half4, okl, lane);
      const float fsum = warp_sum(score_partial(Sn, Pn, On));
      // the butterfly consumed every lane's loads of this stage, so the stage can be refilled for negative j + D
      if (j + D < K) issue(j + D);
      const float f = score_finish<KIND>(fsum, gamma);
      float cneg;
      if (pairwise) {
        float val, A;
        pairwise_terms(loss_kind, margin, f - t, val, A);
        ln += (double)val;
        cneg = A * scaleN;
        cpos -= cneg;
      } else {
        float val, der;
        pointwise_terms(loss_kind, margin, f, false, val, der);
        ln += (double)val;
        cneg = der * scaleN;
      }
      if (lane == 0 && a.neg_scores) a.neg_scores[g * (int64_t)K + j] = f;
      if (cneg != 0.f) {
        Frag<KIND, NC> GS, GP, GO;
        score_grads(Sn, Pn, On, cneg, fsum, gamma, GS, GP, GO);
        if (!shS) red_frag_l(gEl + (int64_t)ns * ldE, half4, okl, GS); else { add_frag(AS, GS); usedS = true; }
        if (!shP) red_frag_l(gRl + (int64_t)np * ldR, half4, okl, GP); else { add_frag(AP, GP); usedP = true; }
        if (!shO) red_frag_l(gEl + (int64_t)no * ldE, half4, okl, GO); else { add_frag(AO, GO); usedO = true; }
      }
    }
    if (cpos != 0.f) {
      Frag<KIND, NC> S, P, O, GS, GP, GO;
      lds_frag(S, posS, half4, okl, lane);
      lds_frag(P, posP, half4, okl, lane);
      lds_frag(O, posO, half4, okl, lane);
      score_grads(S, P, O, cpos, tsum, gamma, GS, GP, GO);
      add_frag(AP, GP);
      usedP = true;
      if (MODE == 1) {  // nothing else accumulates on the positive's entity rows: reduce them directly
        red_frag_l(gEl + (int64_t)s * ldE, half4, okl, GS);
        red_frag_l(gEl + (int64_t)o * ldE, half4, okl, GO);
      } else {
        add_frag(AS, GS); add_frag(AO, GO);
        usedS = usedO = true;
      }
    }
    if (MODE != 1 && usedS) red_frag_l(gEl + (int64_t)s * ldE, half4, okl, AS);
    if (usedP) red_frag_l(gRl + (int64_t)p * ldR, half4, okl, AP);
    if (MODE != 1 && usedO) red_frag_l(gEl + (int64_t)o * ldE, half4, okl, AO);
  }
  if (lane != 0) { lp = 0.0; ln = 0.0; }
  const double wgt = (double)a.loss_weight;
  block_finalize_loss(lp, ln, a.partials, a.ticket, pairwise ? 0.0 : wgt / cntP, wgt / cntN, a.loss);
}

}  // namespace

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
  // shared-memory ring version: per warp the positive's 3 rows + D stages of SR rows; as many CTAs/SM as fit (<= 3)
  const int RB = ctx->we * (int)sizeof(float);
  const int mode = neg ? 0 : (corrupt_mode == KGEB_CORRUPT_BOTH ? 1 : 2);
  const int SR = mode == 0 ? 3 : (mode == 1 ? 2 : 1);
  auto stages_for = [&](size_t budget) {
    return (int)(((long long)(budget / KGEB_WARPS_PER_BLOCK) - 3LL * RB - 8 * 16) / ((long long)SR * RB));
  };
  const size_t sm_budget = 226 * 1024;
  int ctas = 3, D = stages_for(sm_budget / 3 - 1024);
  if (D < 3) { ctas = 2; D = stages_for(sm_budget / 2 - 1024); }
  if (D < 2) { ctas = 1; D = stages_for(sm_budget - 1024); }
  if (D > 6) D = 6;
  if (const char* e = getenv("KGEB_RING_CTAS")) { ctas = atoi(e); D = stages_for(sm_budget / ctas - 1024); if (D > 8) D = 8; }
  if (const char* e = getenv("KGEB_RING_D")) D = atoi(e);
  const char* force = getenv("KGEB_TRAIN_KERNEL");
  const bool use_ring = D >= 1 && !(force && strcmp(force, "reg") == 0);
  if (use_ring) {
    const size_t smem = (size_t)KGEB_WARPS_PER_BLOCK * (3 + SR * D) * RB + (size_t)KGEB_WARPS_PER_BLOCK * (D + 1) * sizeof(uint64_t);
    KGEB_REQUIRE(ctx, smem <= 227 * 1024, KGEB_ERR_UNSUPPORTED, "ring stages do not fit shared memory");
    // persistent grid = SMs x CTAs that are really co-resident (registers may allow fewer than shared memory does)
#define KGEB_RING_LAUNCH(M)                                                                                              \
  {                                                                                                                      \
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_train_ring<KIND, NC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int occ = 0;                                                                                                         \
    KGEB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_train_ring<KIND, NC, M>, KGEB_THREADS, smem)); \
    int grid = grid_for(ctx, B, occ < 1 ? 1 : occ);                                                                      \
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;                                                  \
    KGEB_TIME_BEGIN(ctx, KGEB_TIMED_TRAIN, st);                                                                          \
    k_train_ring<KIND, NC, M><<<grid, KGEB_THREADS, smem, st>>>(a, D, RB);                                               \
    KGEB_TIME_END(ctx, KGEB_TIMED_TRAIN, st);                                                                            \
  }
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, {
      if (mode == 0) KGEB_RING_LAUNCH(0) else if (mode == 1) KGEB_RING_LAUNCH(1) else KGEB_RING_LAUNCH(2)
    }));
  } else {
    int grid = grid_for(ctx, B, 4);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_train_fwd_bwd<KIND, NC><<<grid, KGEB_THREADS, 0, st>>>(a))));
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
