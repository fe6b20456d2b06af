// train_kernels.cu -- K1: fused train step forward + gradient (gather, Philox negatives, score, loss, gradient scatter).
#include "train_ring.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// fused train forward + gradient: one warp per positive and its K negatives
// (register version: kept for A/B measurements, KGEB_TRAIN_KERNEL=reg; latency bound, see profiles/r01_v1_summary.md)
// ---------------------------------------------------------------------------------------------------

template <int KIND, int NC>
__global__ void __launch_bounds__(KGEB_THREADS) k_train_fwd_bwd(TrainArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const bool pairwise = KGEB_IS_PAIRWISE(a.loss_kind);
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


}  // namespace

extern "C" int kgeb_train_fwd_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* pos, int64_t B,
                                  const int32_t* neg, int K, int loss_kind, float margin, float loss_weight,
                                  int corrupt_mode, uint64_t seed, uint64_t offset, float* loss, float* pos_scores,
                                  float* neg_scores, int32_t* neg_out, float* gE, float* gR, uint8_t* touchedE,
                                  uint8_t* touchedR, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, pos && loss && gE && gR && touchedE && touchedR && B > 0 && K > 0, KGEB_ERR_INVALID,
               "train_fwd_bwd: bad arguments");
  KGEB_REQUIRE(ctx, loss_kind >= 0 && loss_kind <= KGEB_LOSS_KIND_MAX, KGEB_ERR_INVALID, "unknown loss kind");
  KGEB_REQUIRE(ctx, corrupt_mode >= 0 && corrupt_mode <= 3, KGEB_ERR_INVALID, "unknown corrupt mode");
  if (ctx->kind == KGEB_HOLE && ctx->hole_fft) {   // Fourier-domain mode: one fused kernel for the pointwise losses
    KGEB_REQUIRE(ctx, E && R, KGEB_ERR_INVALID, "null table pointer");
    const int rc = kgeb_internal_hole_fft_train(ctx, E, R, pos, B, neg, K, loss_kind, margin, loss_weight, corrupt_mode, seed, offset, loss, pos_scores,
                                                neg_scores, neg_out, gE, gR, touchedE, touchedR, (cudaStream_t)stream);
    if (rc != KGEB_ERR_UNSUPPORTED) return rc;
  }
  KGEB_REQUIRE(ctx, ctx->kind != KGEB_HOLE && ctx->kind != KGEB_HAKE, KGEB_ERR_UNSUPPORTED,
               "HolE (outside the Fourier-domain mode with a pointwise loss) and HAKE use the unfused sequence score_fwd -> loss_fwd_bwd -> score_bwd");
  if (int rc = check_tables(ctx, E, R)) return rc;
  KGEB_REQUIRE(ctx, kgeb_aligned16(gE) && kgeb_aligned16(gR), KGEB_ERR_ALIGNMENT, "gradient buffers must be 16-byte aligned");
  int nc;
  TrainArgs a;
  a.sc = make_score_args(ctx, E, R, &nc);
  a.pos = pos; a.B = B; a.neg = neg; a.K = K; a.loss_kind = loss_kind; a.margin = margin; a.loss_weight = loss_weight;
  a.corrupt_mode = corrupt_mode; a.seed = seed; a.offset = offset; a.N = (uint32_t)ctx->N;
  a.loss = loss; a.pos_scores = pos_scores; a.neg_scores = neg_scores; a.neg_out = neg_out;
  a.gE = gE; a.gR = gR; a.tE = touchedE; a.tR = touchedR; a.partials = ctx->d_partials; a.ticket = ctx->d_ticket;
  a.coef_out = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const bool sortseg = ctx->ss_ws != nullptr;   // kgeb_set_grad_sortseg: entity gradient rows by sort + segment instead of vector reductions
  if (sortseg) {
    KGEB_REQUIRE(ctx, a.sc.dp.thr == 0, KGEB_ERR_UNSUPPORTED, "the sort-segmented backward does not regenerate dropout masks: use the reduction path with dropout");
    int32_t* negbuf = nullptr;
    if (int rc = kgeb_internal_sortseg_carve(ctx, B, K, &a.coef_out, &negbuf)) return rc;
    if (!neg && !neg_out) a.neg_out = negbuf;   // the kernel's Philox ids are the segment kernel's input
  }
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
  if (ctx->env_ring_ctas > 0) { ctas = ctx->env_ring_ctas; D = stages_for(sm_budget / ctas - 1024); if (D > 8) D = 8; }
  if (ctx->env_ring_d > 0) D = ctx->env_ring_d;
  (void)ctas;
  const bool use_ring = D >= 1 && (a.sc.dp.thr != 0 || sortseg || !ctx->env_force_reg);  // dropout and sortseg live in the ring kernel only
  KGEB_REQUIRE(ctx, use_ring || !sortseg, KGEB_ERR_UNSUPPORTED, "the sort-segmented backward needs the ring kernel, whose stages do not fit for this width");
  KGEB_REQUIRE(ctx, use_ring || a.sc.dp.thr == 0, KGEB_ERR_UNSUPPORTED, "dropout needs the ring kernel, whose stages do not fit for this width");
  if (use_ring) {
    const size_t smem = (size_t)KGEB_WARPS_PER_BLOCK * (3 + SR * D) * RB + (size_t)KGEB_WARPS_PER_BLOCK * (D + 1) * sizeof(uint64_t);
    KGEB_REQUIRE(ctx, smem <= 227 * 1024, KGEB_ERR_UNSUPPORTED, "ring stages do not fit shared memory");
    const int rc = a.sc.dp.thr ? kgeb_internal_train_ring_dp(ctx, &a, nc, mode, D, RB, smem, B, st)
                               : launch_train_ring<false>(ctx, a, nc, mode, D, RB, smem, B, st);
    if (rc != KGEB_OK) return rc;
  } else {
    int grid = grid_for(ctx, B, 4);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    KGEB_TIME_BEGIN(ctx, KGEB_TIMED_TRAIN, st);
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_train_fwd_bwd<KIND, NC><<<grid, KGEB_THREADS, 0, st>>>(a))));
    KGEB_TIME_END(ctx, KGEB_TIMED_TRAIN, st);
  }
  KGEB_LAUNCH_OK(ctx);
  if (sortseg) return kgeb_internal_sortseg_entities(ctx, &a.sc, nc, pos, neg ? neg : a.neg_out, B, K, gE, st);
  return KGEB_OK;
}
