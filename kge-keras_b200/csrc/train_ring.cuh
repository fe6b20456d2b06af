// train_ring.cuh -- the fused train-step kernel (shared-memory ring version) as a template over the dropout flag.
#pragma once
#include "frag.cuh"

namespace {

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
template <int KIND, int NC, int MODE, bool DP>
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
  const bool pairwise = KGEB_IS_PAIRWISE(loss_kind);
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
    auto issue = [&](int j, int st) {   // st = ring stage of negative j, tracked by the callers (a runtime modulo costs ~30 instructions)
      int ns, np, no, hs;
      ids_of(j, ns, np, no, hs);
      if (lane == 0) {
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
    for (int j = 0; j < prime; ++j) issue(j, j);

    // positive score
    mbar_wait(bars, phases & 1u);
    phases ^= 1u;
    float tsum, t;
    uint32_t mposS = 0, mposP = 0, mposO = 0;  // dropout keep bits of the positive call's rows (reused by its backward)
    {
      Frag<KIND, NC> S, P, O;
      lds_frag(S, posS, half4, okl, lane);
      lds_frag(P, posP, half4, okl, lane);
      lds_frag(O, posO, half4, okl, lane);
      if (DP) {
        mposS = dp_bits<KIND, NC>(sc.dp, 0, 0, (uint64_t)g, lane);
        mposP = dp_bits<KIND, NC>(sc.dp, 1, 0, (uint64_t)g, lane);
        mposO = dp_bits<KIND, NC>(sc.dp, 2, 0, (uint64_t)g, lane);
        dp_apply(S, mposS, sc.dp.scale); dp_apply(P, mposP, sc.dp.scale); dp_apply(O, mposO, sc.dp.scale);
      }
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

    int st = -1;  // ring stage of negative j
    for (int j = 0; j < K; ++j) {
      if (++st == D) st = 0;
      if ((j & 31) == 0 && j > 0) {  // crossed into the "ahead" chunk: rotate, then draw the chunk after it
        cur_chunk = j >> 5; my_ns = ah_ns; my_np = ah_np; my_no = ah_no; my_hs = ah_hs;
        if ((cur_chunk + 1) * 32 < K) compute_chunk(cur_chunk + 1, ah_ns, ah_np, ah_no, ah_hs);
      }
      int ns, np, no, hs;
      ids_of(j, ns, np, no, hs);
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
      uint32_t mS = 0, mP = 0, mO = 0;
      if (DP) {  // the negative call draws its own masks, also for rows it shares with the positive
        const uint64_t gi = (uint64_t)(g * (int64_t)K + j);
        mS = dp_bits<KIND, NC>(sc.dp, 0, 1, gi, lane);
        mP = dp_bits<KIND, NC>(sc.dp, 1, 1, gi, lane);
        mO = dp_bits<KIND, NC>(sc.dp, 2, 1, gi, lane);
        dp_apply(Sn, mS, sc.dp.scale); dp_apply(Pn, mP, sc.dp.scale); dp_apply(On, mO, sc.dp.scale);
      }
      const float fsum = warp_sum(score_partial(Sn, Pn, On));
      // the butterfly consumed every lane's loads of this stage, so the stage can be refilled for negative j + D
      if (j + D < K) issue(j + D, st);
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
      if (a.coef_out) {   // sort-segmented mode: entity rows are reduced by segment later; only the relation gradient is formed here
        if (lane == 0) a.coef_out[a.B + g * (int64_t)K + j] = cneg;
        if (cneg != 0.f) {
          Frag<KIND, NC> GS, GP, GO;
          score_grads(Sn, Pn, On, cneg, fsum, gamma, GS, GP, GO);
          if (DP) dp_apply(GP, mP, sc.dp.scale);
          if (!shP) red_frag_l(gRl + (int64_t)np * ldR, half4, okl, GP); else { add_frag(AP, GP); usedP = true; }
        }
      } else if (cneg != 0.f) {
        Frag<KIND, NC> GS, GP, GO;
        score_grads(Sn, Pn, On, cneg, fsum, gamma, GS, GP, GO);
        if (DP) { dp_apply(GS, mS, sc.dp.scale); dp_apply(GP, mP, sc.dp.scale); dp_apply(GO, mO, sc.dp.scale); }
        if (!shS) red_frag_l(gEl + (int64_t)ns * ldE, half4, okl, GS); else { add_frag(AS, GS); usedS = true; }
        if (!shP) red_frag_l(gRl + (int64_t)np * ldR, half4, okl, GP); else { add_frag(AP, GP); usedP = true; }
        if (!shO) red_frag_l(gEl + (int64_t)no * ldE, half4, okl, GO); else { add_frag(AO, GO); usedO = true; }
      }
    }
    if (a.coef_out && lane == 0) a.coef_out[g] = cpos;
    if (cpos != 0.f) {
      Frag<KIND, NC> S, P, O, GS, GP, GO;
      lds_frag(S, posS, half4, okl, lane);
      lds_frag(P, posP, half4, okl, lane);
      lds_frag(O, posO, half4, okl, lane);
      if (DP) { dp_apply(S, mposS, sc.dp.scale); dp_apply(P, mposP, sc.dp.scale); dp_apply(O, mposO, sc.dp.scale); }
      score_grads(S, P, O, cpos, tsum, gamma, GS, GP, GO);
      if (DP) { dp_apply(GS, mposS, sc.dp.scale); dp_apply(GP, mposP, sc.dp.scale); dp_apply(GO, mposO, sc.dp.scale); }
      add_frag(AP, GP);
      usedP = true;
      if (a.coef_out) {
        // entity rows of the positive: left to the segment kernel
      } else if (MODE == 1) {  // nothing else accumulates on the positive's entity rows: reduce them directly
        red_frag_l(gEl + (int64_t)s * ldE, half4, okl, GS);
        red_frag_l(gEl + (int64_t)o * ldE, half4, okl, GO);
      } else {
        add_frag(AS, GS); add_frag(AO, GO);
        usedS = usedO = true;
      }
    }
    if (MODE != 1 && usedS && !a.coef_out) red_frag_l(gEl + (int64_t)s * ldE, half4, okl, AS);
    if (usedP) red_frag_l(gRl + (int64_t)p * ldR, half4, okl, AP);
    if (MODE != 1 && usedO && !a.coef_out) red_frag_l(gEl + (int64_t)o * ldE, half4, okl, AO);
  }
  if (lane != 0) { lp = 0.0; ln = 0.0; }
  const double wgt = (double)a.loss_weight;
  block_finalize_loss(lp, ln, a.partials, a.ticket, pairwise ? 0.0 : wgt / cntP, wgt / cntN, a.loss);
}

// host launcher of one DP flavour (train_kernels.cu instantiates DP = false, train_kernels_dp.cu DP = true, so the two
// sets of 48 kernels compile in parallel)
template <bool DP>
int launch_train_ring(kgeb_ctx* ctx, const TrainArgs& a, int nc, int mode, int D, int RB, size_t smem, int64_t B, cudaStream_t st) {
#define KGEB_RING_LAUNCH(M)                                                                                              \
  {                                                                                                                      \
    auto& cfg = ctx->ring_cfg[M][DP ? 1 : 0];   /* kind and NC are fixed per context */                                  \
    if (cfg.smem != (int)smem) {                                                                                         \
      KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_train_ring<KIND, NC, M, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      int occ = 0;                                                                                                       \
      KGEB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_train_ring<KIND, NC, M, DP>, KGEB_THREADS, smem)); \
      cfg.smem = (int)smem; cfg.occ = occ < 1 ? 1 : occ;                                                                 \
    }                                                                                                                    \
    int grid = grid_for(ctx, B, cfg.occ);                                                                                \
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;                                                  \
    KGEB_TIME_BEGIN(ctx, KGEB_TIMED_TRAIN, st);                                                                          \
    k_train_ring<KIND, NC, M, DP><<<grid, KGEB_THREADS, smem, st>>>(a, D, RB);                                           \
    KGEB_TIME_END(ctx, KGEB_TIMED_TRAIN, st);                                                                            \
  }
  KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, {
    if (mode == 0) KGEB_RING_LAUNCH(0) else if (mode == 1) KGEB_RING_LAUNCH(1) else KGEB_RING_LAUNCH(2)
  }));
#undef KGEB_RING_LAUNCH
  return KGEB_OK;
}

}  // namespace

// defined in train_kernels_dp.cu: the DP = true flavour (dropout armed by kgeb_set_dropout)
int kgeb_internal_train_ring_dp(kgeb_ctx* ctx, const void* args, int nc, int mode, int D, int RB, size_t smem, int64_t B, cudaStream_t st);
