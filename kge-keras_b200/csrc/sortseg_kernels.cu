// sortseg_kernels.cu -- K2 as the north star words it: the ENTITY gradient rows of a train step are sort-segmented by entity id
// and added segment by segment, instead of one vector reduction per row reference (train_ring.cuh).
//
//   records   (entity id, 2*t + slot) for every scored triple t (positives 0..B-1, then the B*K negatives) and slot (0 = s, 1 = o);
//   sort      cub::DeviceRadixSort::SortPairs on the id (only the bits an id can have) -- stable, so the order inside a
//             segment is the record order and the fp32 sums are REPRODUCIBLE run to run (the reduction path is not);
//   segments  a warp looks at 32 consecutive sorted records, finds the segment heads among them by ballot and owns every
//             segment that starts there: it re-forms each reference's gradient row from the triple's other two rows
//             (dL/dscore comes from the forward kernel, TrainArgs::coef_out), accumulates in registers in sorted order and
//             adds the finished row to the gradient buffer with ONE plain read-modify-write -- no atomics, one writer per row.
//
// The optimizer at segment end (kgeb_train_step_sortseg).  Re-forming a gradient row READS the parameter rows of the other entities
// of its triples, so a warp that updated its row in place would race with the warps still reading it.  The fused step therefore
// works on TWO copies of the entity table: every warp reads the current copy and writes the updated row -- parameters straight from
// the registers that hold the segment's own row, moments read and written in place (nobody else reads them) -- to the other copy;
// the host swaps the two after the step.  No gradient buffer exists on this path: per touched row the step moves
// p, m, v once in and once out instead of g twice + p, m, v twice + g twice more.  Rows touched by the previous step but not by this
// one are stale in the copy being written; a scan of the two steps' flag arrays copies them over (k_ss_sync_rows).  Without the
// fused entry the touched-row optimizer kernel (loss_opt_kernels.cu) runs after the plain segment kernel, unchanged.
#include <cub/cub.cuh>

#include "frag.cuh"

namespace {

struct SegArgs {
  ScoreArgs sc;
  const int32_t* pos; const int32_t* neg; int64_t B; int64_t m;   // m = B * (1 + K) scored triples
  const float* coef;                                              // [m]
  const uint32_t* keys; const uint32_t* vals; int64_t n;          // sorted records, n = 2 m
  float* gE;
};

__global__ void k_ss_records(const int32_t* __restrict__ pos, const int32_t* __restrict__ neg, int64_t B, int64_t m,
                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* tr = t < B ? pos + 3 * t : neg + 3 * (t - B);
    keys[2 * t] = (uint32_t)__ldg(tr);
    keys[2 * t + 1] = (uint32_t)__ldg(tr + 2);
    vals[2 * t] = (uint32_t)(2 * t);
    vals[2 * t + 1] = (uint32_t)(2 * t + 1);
  }
}

template <int KIND>
constexpr bool needs_sum() { return KIND == TRANSE_L2; }

// optimizer applied at segment end; E_next == nullptr: plain mode (the gradient row is added to SegArgs::gE)
struct SegOpt {
  float* E_next; float* m; float* v;
  int kind; float lr, b1, b2, eps, lr_t;
};
__device__ __forceinline__ void ss_opt_elem(const SegOpt& o, float g, float& p, float& m, float& v) {   // loss_opt_kernels.cu: opt_elem
  if (o.kind == KGEB_SGD) {
    p -= o.lr * g;
  } else if (o.kind == KGEB_ADAGRAD) {
    m += g * g;
    p -= o.lr * g / (sqrtf(m) + o.eps);
  } else {
    m = o.b1 * m + (1.f - o.b1) * g;
    v = o.b2 * v + (1.f - o.b2) * g * g;
    p -= o.lr_t * m / (sqrtf(v) + o.eps);
  }
}
__device__ __forceinline__ void ss_opt4(const SegOpt& o, float4 g, float4 p, int64_t off4) {
  float4 m = f4zero(), v = f4zero();
  float4* M4 = reinterpret_cast<float4*>(o.m);
  float4* V4 = reinterpret_cast<float4*>(o.v);
  if (o.kind != KGEB_SGD) m = M4[off4];
  if (o.kind >= KGEB_LAZY_ADAM) v = V4[off4];
  ss_opt_elem(o, g.x, p.x, m.x, v.x);
  ss_opt_elem(o, g.y, p.y, m.y, v.y);
  ss_opt_elem(o, g.z, p.z, m.z, v.z);
  ss_opt_elem(o, g.w, p.w, m.w, v.w);
  reinterpret_cast<float4*>(o.E_next)[off4] = p;
  if (o.kind != KGEB_SGD) M4[off4] = m;
  if (o.kind >= KGEB_LAZY_ADAM) V4[off4] = v;
}

// The walk of a segment is a chain of dependent loads (keys -> record -> triple ids -> partner rows -> moments); with ~2-3
// references per touched row a warp spends its time waiting, not moving bytes.  L2 prefetches -- no registers, no dependency --
// start the DRAM fetch of every row the warp is going to need as soon as its address is known: a lane prefetches the partner
// row of ITS record of the window while the rows are then consumed one reference at a time, and the segment's moment rows are
// requested before the first reference is looked at.
__device__ __forceinline__ void ss_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ss_prefetch_row_warp(const float* row, int row_bytes, int lane) {
  for (int b = lane * 128; b < row_bytes; b += 32 * 128) ss_prefetch_l2(reinterpret_cast<const char*>(row) + b);
}

template <int KIND, int NC, bool FUSED>
__global__ void __launch_bounds__(KGEB_THREADS) k_ss_segments(SegArgs a, SegOpt o) {
  const int lane = threadIdx.x & 31;
  const ScoreArgs& sc = a.sc;
  const int64_t nwin = (a.n + 31) >> 5;
  const int64_t nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t win = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5); win < nwin; win += nw) {
    const int64_t i0 = win << 5;
    const int64_t i = i0 + lane;
    const uint32_t key = i < a.n ? __ldg(a.keys + i) : 0xffffffffu;
    uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (lane == 0) prev = i0 > 0 ? __ldg(a.keys + i0 - 1) : 0xffffffffu;
    unsigned heads = __ballot_sync(0xffffffffu, i < a.n && (i == 0 || key != prev));
    while (heads) {
      const int b = __ffs(heads) - 1;
      heads &= heads - 1;
      const uint32_t e = __shfl_sync(0xffffffffu, key, b);
      Frag<KIND, NC> OWN, ACC;
      load_frag(OWN, sc.E, (int64_t)e, sc.ldE, sc.ge, lane);
      zero_frag(ACC);
      if (FUSED) {
        if (o.kind != KGEB_SGD) ss_prefetch_row_warp(o.m + (int64_t)e * sc.ldE, sc.ldE * 4, lane);
        if (o.kind >= KGEB_LAZY_ADAM) ss_prefetch_row_warp(o.v + (int64_t)e * sc.ldE, sc.ldE * 4, lane);
      }
      // walk the segment in windows of 32 records: every lane fetches the ids and the coefficient of ONE record, the loop
      // below then broadcasts them, so that the only dependent loads left per reference are the two partner rows
      for (int64_t w0 = i0 + b;; w0 += 32) {
        const int64_t r = w0 + lane;
        const bool mine = r < a.n && __ldg(a.keys + r) == e;
        int my_p = 0, my_x = 0, my_slot = 0;
        float my_c = 0.f;
        if (mine) {
          const uint32_t v = __ldg(a.vals + r);
          const int64_t t = v >> 1;
          my_slot = (int)(v & 1u);
          my_c = __ldg(a.coef + t);
          const int32_t* tr = t < a.B ? a.pos + 3 * t : a.neg + 3 * (t - a.B);
          my_p = __ldg(tr + 1);
          my_x = __ldg(tr + (my_slot ? 0 : 2));          // the OTHER entity of the triple
          if (my_c != 0.f) {                             // start the partner row's trip from DRAM now
            const char* xr = reinterpret_cast<const char*>(sc.E + (int64_t)my_x * sc.ldE);
            for (int b = 0; b < sc.ldE * 4; b += 128) ss_prefetch_l2(xr + b);
          }
        }
        const unsigned in_seg = __ballot_sync(0xffffffffu, mine);   // a prefix of the lanes: the records are sorted
        unsigned live = __ballot_sync(0xffffffffu, mine && my_c != 0.f);
        while (live) {
          const int l = __ffs(live) - 1;
          live &= live - 1;
          const int p = __shfl_sync(0xffffffffu, my_p, l), x = __shfl_sync(0xffffffffu, my_x, l);
          const int slot = __shfl_sync(0xffffffffu, my_slot, l);
          const float c = __shfl_sync(0xffffffffu, my_c, l);
          Frag<KIND, NC> P, X, GS, GP, GO;
          load_frag(P, sc.R, (int64_t)p, sc.ldR, sc.gr, lane);
          load_frag(X, sc.E, (int64_t)x, sc.ldE, sc.ge, lane);
          if (slot == 0) {      // this entity is the subject of the triple
            const float sum = needs_sum<KIND>() ? warp_sum(score_partial(OWN, P, X)) : 0.f;
            score_grads(OWN, P, X, c, sum, sc.gamma, GS, GP, GO);
            add_frag(ACC, GS);
          } else {
            const float sum = needs_sum<KIND>() ? warp_sum(score_partial(X, P, OWN)) : 0.f;
            score_grads(X, P, OWN, c, sum, sc.gamma, GS, GP, GO);
            add_frag(ACC, GO);
          }
        }
        if (in_seg != 0xffffffffu) break;   // the segment ended inside this window
      }
      if (FUSED) {   // one writer per row: the updated row goes to the OTHER table copy, the moments are updated in place
        const int64_t row4 = (int64_t)e * (sc.ldE >> 2);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const int c = lane + 32 * k;
          if (c < sc.ge.nch) {
            ss_opt4(o, ACC.v[k], OWN.v[k], row4 + c);
            if (split_row<KIND>()) ss_opt4(o, ACC.v[NC + k], OWN.v[NC + k], row4 + sc.ge.half4 + c);
          }
        }
        continue;
      }
      // one writer per row: plain read-modify-write of the gradient row
      float4* g4 = reinterpret_cast<float4*>(a.gE + (int64_t)e * sc.ldE);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        if (c < sc.ge.nch) {
          float4 g = g4[c];
          g.x += ACC.v[k].x; g.y += ACC.v[k].y; g.z += ACC.v[k].z; g.w += ACC.v[k].w;
          g4[c] = g;
          if (split_row<KIND>()) {
            float4 h = g4[sc.ge.half4 + c];
            h.x += ACC.v[NC + k].x; h.y += ACC.v[NC + k].y; h.z += ACC.v[NC + k].z; h.w += ACC.v[NC + k].w;
            g4[sc.ge.half4 + c] = h;
          }
        }
      }
    }
  }
}

// rows the previous step updated but this one did not: the copy being written still holds their pre-previous-step values.
// Also clears the previous step's flags: that array is the next step's "touched now".
__global__ void __launch_bounds__(KGEB_THREADS) k_ss_sync_rows(const float* __restrict__ E_cur, float* __restrict__ E_next, uint8_t* __restrict__ prev,
                                                               const uint8_t* __restrict__ now, int64_t n_rows, int ld4, int w4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const float4* S = reinterpret_cast<const float4*>(E_cur);
  float4* D = reinterpret_cast<float4*>(E_next);
  for (int64_t base = warp0 * 32; base < n_rows; base += nw * 32) {
    const int64_t r = base + lane;
    const bool was = r < n_rows && prev[r] != 0;
    const bool stale = was && now[r] == 0;
    if (was) prev[r] = 0;
    unsigned mask = __ballot_sync(0xffffffffu, stale);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const int64_t off = (base + b) * ld4;
      for (int c = lane; c < w4; c += 32) D[off + c] = ldg4(S + off + c);
    }
  }
}

size_t ss_align(size_t x) { return (x + 255) & ~(size_t)255; }

int ss_id_bits(int64_t N) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < N) ++b;
  return b;
}

size_t ss_cub_bytes(int64_t n, int bits) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (int)n, 0, bits);
  return tb;
}

}  // namespace

extern "C" int64_t kgeb_sortseg_workspace_bytes(const kgeb_ctx* ctx, int64_t B, int K) {
  if (!ctx || B <= 0 || K <= 0) return KGEB_ERR_INVALID;
  const int64_t m = B * (1 + (int64_t)K), n = 2 * m;
  if (n >= 0x7fffffffLL) return KGEB_ERR_UNSUPPORTED;
  return (int64_t)(ss_align((size_t)m * 4) + ss_align((size_t)B * K * 12) + 4 * ss_align((size_t)n * 4) + ss_align(ss_cub_bytes(n, ss_id_bits(ctx->N))));
}

extern "C" int kgeb_set_grad_sortseg(kgeb_ctx* ctx, void* workspace, int64_t workspace_bytes) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, workspace == nullptr || ((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && workspace_bytes > 0), KGEB_ERR_ALIGNMENT,
               "set_grad_sortseg: the workspace must be 256-byte aligned");
  ctx->ss_ws = static_cast<unsigned char*>(workspace);
  ctx->ss_bytes = workspace ? workspace_bytes : 0;
  return KGEB_OK;
}

// workspace carving shared with kgeb_train_fwd_bwd (train_kernels.cu)
int kgeb_internal_sortseg_carve(kgeb_ctx* ctx, int64_t B, int K, float** coef, int32_t** negbuf) {
  const int64_t need = kgeb_sortseg_workspace_bytes(ctx, B, K);
  if (need < 0) return kgeb_fail(ctx, (int)need, "%s", "sortseg: batch too large for 32-bit record ids");
  KGEB_REQUIRE(ctx, ctx->ss_ws && ctx->ss_bytes >= need, KGEB_ERR_WORKSPACE, "sortseg: workspace smaller than kgeb_sortseg_workspace_bytes(B, K)");
  const int64_t m = B * (1 + (int64_t)K);
  *coef = reinterpret_cast<float*>(ctx->ss_ws);
  *negbuf = reinterpret_cast<int32_t*>(ctx->ss_ws + ss_align((size_t)m * 4));
  return KGEB_OK;
}

int kgeb_internal_sortseg_entities(kgeb_ctx* ctx, const void* score_args, int nc, const int32_t* pos, const int32_t* neg, int64_t B, int K,
                                   float* gE, cudaStream_t st) {
  const int64_t m = B * (1 + (int64_t)K), n = 2 * m;
  unsigned char* w = ctx->ss_ws;
  const float* coef = reinterpret_cast<const float*>(w);
  w += ss_align((size_t)m * 4) + ss_align((size_t)B * K * 12);
  uint32_t* keys_a = reinterpret_cast<uint32_t*>(w); w += ss_align((size_t)n * 4);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(w); w += ss_align((size_t)n * 4);
  uint32_t* keys_b = reinterpret_cast<uint32_t*>(w); w += ss_align((size_t)n * 4);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(w); w += ss_align((size_t)n * 4);
  const int bits = ss_id_bits(ctx->N);
  size_t tb = ss_cub_bytes(n, bits);
  int64_t blocks = (m + 255) / 256;
  if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_AUX, st);
  k_ss_records<<<(int)blocks, 256, 0, st>>>(pos, neg, B, m, keys_a, vals_a);
  KGEB_LAUNCH_OK(ctx);
  KGEB_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(w, tb, keys_a, keys_b, vals_a, vals_b, (int)n, 0, bits, st));
  ctx->launches += 3;   // CUB's histogram + onesweep passes (not counted individually)
  SegArgs a;
  a.sc = *static_cast<const ScoreArgs*>(score_args);
  a.pos = pos; a.neg = neg; a.B = B; a.m = m; a.coef = coef; a.keys = keys_b; a.vals = vals_b; a.n = n; a.gE = gE;
  const int grid = grid_for(ctx, (n + 31) / 32, 8);
  SegOpt o = {};
  if (ctx->ss_fused) {
    o = *static_cast<const SegOpt*>(ctx->ss_fused);
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_ss_segments<KIND, NC, true><<<grid, KGEB_THREADS, 0, st>>>(a, o))));
  } else {
    KGEB_DISPATCH_KIND(ctx->kind, KGEB_DISPATCH_NC(nc, (k_ss_segments<KIND, NC, false><<<grid, KGEB_THREADS, 0, st>>>(a, o))));
  }
  KGEB_TIME_END(ctx, KGEB_TIMED_AUX, st);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

// ---- the whole entity half of a step with the optimizer fused at segment end (north-star item 2) ---------------------------------
extern "C" int kgeb_train_step_sortseg(kgeb_ctx* ctx, const float* E_cur, float* E_next, float* R, const int32_t* pos, int64_t B,
                                       const int32_t* neg, int K, int loss_kind, float margin, float loss_weight, int corrupt_mode,
                                       uint64_t seed, uint64_t offset, float* loss, float* gR, uint8_t* touchedR, uint8_t* touched_now,
                                       uint8_t* touched_prev, float* mE, float* vE, float* mR, float* vR, int opt_kind, float lr,
                                       float beta1, float beta2, float eps, int64_t step, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, E_cur && E_next && E_cur != E_next && R && gR && touchedR && touched_now && touched_prev && touched_now != touched_prev,
               KGEB_ERR_INVALID, "train_step_sortseg: two distinct table copies and two distinct flag arrays are needed");
  KGEB_REQUIRE(ctx, opt_kind == KGEB_SGD || opt_kind == KGEB_ADAGRAD || opt_kind == KGEB_LAZY_ADAM, KGEB_ERR_UNSUPPORTED,
               "train_step_sortseg updates touched rows only: SGD, Adagrad or LazyAdam (Keras' dense Adam decays every row's moments)");
  KGEB_REQUIRE(ctx, opt_kind == KGEB_SGD || mE, KGEB_ERR_INVALID, "optimizer state missing");
  KGEB_REQUIRE(ctx, opt_kind != KGEB_LAZY_ADAM || (vE && step >= 1), KGEB_ERR_INVALID, "Adam needs its second moment and step >= 1");
  KGEB_REQUIRE(ctx, ctx->ss_ws != nullptr, KGEB_ERR_WORKSPACE, "train_step_sortseg: arm the workspace first (kgeb_set_grad_sortseg)");
  KGEB_REQUIRE(ctx, kgeb_aligned16(E_next) && (!mE || kgeb_aligned16(mE)) && (!vE || kgeb_aligned16(vE)), KGEB_ERR_ALIGNMENT,
               "train_step_sortseg: 16-byte aligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  SegOpt o;
  o.E_next = E_next; o.m = mE; o.v = vE; o.kind = opt_kind; o.lr = lr; o.b1 = beta1; o.b2 = beta2; o.eps = eps; o.lr_t = lr;
  if (opt_kind == KGEB_LAZY_ADAM) {
    const double t = (double)step;
    o.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  }
  // forward (scores, loss, per-triple coefficients, relation gradient by vector reductions) + records + sort + fused segments
  ctx->ss_fused = &o;
  int rc = kgeb_train_fwd_bwd(ctx, E_cur, R, pos, B, neg, K, loss_kind, margin, loss_weight, corrupt_mode, seed, offset, loss, nullptr, nullptr,
                              nullptr, E_next /* never written as a gradient buffer on this path */, gR, touched_now, touchedR, stream);
  ctx->ss_fused = nullptr;
  if (rc != KGEB_OK) return rc;
  const int64_t blocks64 = (ctx->N + 32 * KGEB_WARPS_PER_BLOCK - 1) / (32 * KGEB_WARPS_PER_BLOCK);
  const int blocks = (int)(blocks64 > (int64_t)ctx->num_sms * 8 ? (int64_t)ctx->num_sms * 8 : blocks64);
  k_ss_sync_rows<<<blocks, KGEB_THREADS, 0, st>>>(E_cur, E_next, touched_prev, touched_now, ctx->N, ctx->ldE / 4, ctx->we / 4);
  KGEB_LAUNCH_OK(ctx);
  return kgeb_opt_update(ctx, R, gR, touchedR, ctx->R, ctx->wr, ctx->ldR, mR, vR, opt_kind, lr, beta1, beta2, eps, step, stream);
}
