// frag.cuh -- register row fragments and the score / gradient arithmetic shared by the scoring and training kernels.
#pragma once
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int DISTMULT = KGEB_DISTMULT, TRANSE_L1 = KGEB_TRANSE_L1, TRANSE_L2 = KGEB_TRANSE_L2, COMPLEX = KGEB_COMPLEX;
constexpr int ROTATE = KGEB_ROTATE, PROTATE = KGEB_PROTATE, MODE_L1 = KGEB_MODE_L1, MODE_L2 = KGEB_MODE_L2;

// Rows stored as two halves [re | im]: ComplEx (all three rows) and RotatE (entity rows; its relation row is d phases).
template <int KIND, bool REL = false>
constexpr bool split_row() { return KIND == COMPLEX || (KIND == ROTATE && !REL); }

// Row fragment: NC chunks per lane (split rows: NC chunks of the real half followed by NC of the imaginary half).
template <int KIND, int NC>
struct Frag {
  static constexpr int NT = split_row<KIND>() ? 2 * NC : NC;
  float4 v[NT];
};

struct RowGeom {
  int nch;   // 128-bit chunks per (half-)row handled by the lane loop: width/4, or (width/2)/4 for ComplEx
  int half4; // ComplEx: offset of the imaginary half in float4 units; otherwise 0
};

// REL = true marks the relation row (only RotatE treats it differently: a plain row of d phases in v[0..NC))
template <int KIND, int NC, bool REL = false>
__device__ __forceinline__ void load_frag(Frag<KIND, NC>& f, const float* __restrict__ table, int64_t row, int ld,
                                          RowGeom g, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(table + row * (int64_t)ld);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    const bool ok = c < g.nch;
    f.v[i] = ok ? ldg4(r4 + c) : f4zero();
    if (split_row<KIND, REL>()) f.v[NC + i] = ok ? ldg4(r4 + g.half4 + c) : f4zero();
    else if (split_row<KIND>()) f.v[NC + i] = f4zero();
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

template <int KIND, int NC, bool REL = false>
__device__ __forceinline__ void red_frag(float* __restrict__ table, int64_t row, int ld, RowGeom g, int lane,
                                         const Frag<KIND, NC>& f) {
  float* r = table + row * (int64_t)ld;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = lane + 32 * i;
    if (c < g.nch) {
      red_add4(r + 4 * c, f.v[i]);
      if (split_row<KIND, REL>()) red_add4(r + 4 * (g.half4 + c), f.v[NC + i]);
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
      } else if (KIND == MODE_L1 || KIND == MODE_L2) {   // models.py:418: s * p - o
        const float r = __fsub_rn(__fmul_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e));
        acc += (KIND == MODE_L1) ? fabsf(r) : __fmul_rn(r, r);
      } else if (KIND == PROTATE) {
        // models.py:481-485: |sin((ph(s) + ph(p) - ph(o)) / 2)| with ph(x) = atan2(sin x, cos x) = x - 2 pi k: the
        // half-angle turns the wrap into a sign flip, which the absolute value removes, so ph() is not evaluated
        acc += fabsf(sinf(0.5f * __fsub_rn(__fadd_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e))));
      } else if (KIND == ROTATE) {
        // models.py:441-453: cos/sin of atan2(sin p, cos p) are cos p / sin p; the SIGNED residuals are summed
        float sn, c;
        sincosf(comp(p.v[i], _e), &sn, &c);
        const float sr = comp(s.v[i], _e), si = comp(s.v[NC + i], _e);
        const float re = __fsub_rn(__fsub_rn(__fmul_rn(sr, c), __fmul_rn(si, sn)), comp(o.v[i], _e));
        const float im = __fsub_rn(__fadd_rn(__fmul_rn(sr, sn), __fmul_rn(si, c)), comp(o.v[NC + i], _e));
        acc += __fadd_rn(re, im);
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
__device__ __forceinline__ float score_finish(float sum, float gamma, float param = 1.f) {
  if (KIND == TRANSE_L1 || KIND == ROTATE) return gamma > 0.f ? gamma - sum : sum;
  if (KIND == TRANSE_L2) { const float n = sqrtf(sum); return gamma > 0.f ? gamma - n : n; }
  if (KIND == PROTATE) { const float a = __fmul_rn(sum, param); return gamma > 0.f ? gamma - a : a; }
  if (KIND == MODE_L1) return gamma - sum;            // models.py:418 subtracts unconditionally
  if (KIND == MODE_L2) return gamma - sqrtf(sum);
  return sum;
}

// gradient fragments of coef * score wrt the three rows.  `sum` is the un-finished score sum (needed by L2).
template <int KIND, int NC>
__device__ __forceinline__ void score_grads(const Frag<KIND, NC>& s, const Frag<KIND, NC>& p, const Frag<KIND, NC>& o,
                                            float coef, float sum, float gamma, Frag<KIND, NC>& gs, Frag<KIND, NC>& gp,
                                            Frag<KIND, NC>& go, float param = 1.f) {
  float k = coef;
  if (KIND == TRANSE_L1 || KIND == TRANSE_L2 || KIND == ROTATE || KIND == PROTATE) {
    k = gamma > 0.f ? -coef : coef;
    if (KIND == TRANSE_L2) { const float n = sqrtf(sum); k = n > 0.f ? k / n : 0.f; }  // TF gives NaN at r == 0; we give 0
    if (KIND == PROTATE) k *= 0.5f * param;
  }
  if (KIND == MODE_L1 || KIND == MODE_L2) {
    k = -coef;
    if (KIND == MODE_L2) { const float n = sqrtf(sum); k = n > 0.f ? k / n : 0.f; }
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
      } else if (KIND == MODE_L1 || KIND == MODE_L2) {
        const float a = comp(s.v[i], _e), b = comp(p.v[i], _e);
        const float r = __fsub_rn(__fmul_rn(a, b), comp(o.v[i], _e));
        const float g = (KIND == MODE_L1) ? (r > 0.f ? k : (r < 0.f ? -k : 0.f)) : k * r;
        comp(gs.v[i], _e) = g * b;
        comp(gp.v[i], _e) = g * a;
        comp(go.v[i], _e) = -g;
      } else if (KIND == PROTATE) {   // d|sin y|/dy = sign(sin y) cos y, y = (s + p - o)/2; k already holds param/2
        float sn, c;
        sincosf(0.5f * __fsub_rn(__fadd_rn(comp(s.v[i], _e), comp(p.v[i], _e)), comp(o.v[i], _e)), &sn, &c);
        const float g = sn > 0.f ? k * c : (sn < 0.f ? -k * c : 0.f);
        comp(gs.v[i], _e) = g;
        comp(gp.v[i], _e) = g;
        comp(go.v[i], _e) = -g;
      } else if (KIND == ROTATE) {
        float sn, c;
        sincosf(comp(p.v[i], _e), &sn, &c);
        const float sr = comp(s.v[i], _e), si = comp(s.v[NC + i], _e);
        comp(gs.v[i], _e) = k * (c + sn);
        comp(gs.v[NC + i], _e) = k * (c - sn);
        comp(gp.v[i], _e) = k * (sr * (c - sn) - si * (c + sn));
        comp(gp.v[NC + i], _e) = 0.f;
        comp(go.v[i], _e) = -k;
        comp(go.v[NC + i], _e) = -k;
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

// ---------------------------------------------------------------------------------------------------
// Dropout (models.py:122,144-150): an independent Bernoulli(1-dp) keep mask on every gathered s / p / o row of every
// model call, kept values scaled by 1/(1-dp).  Counter-based so forward, backward and the host restatement
// (oracle.kge_oracle.dropout_keep) agree without storing masks: 16 random bits per element, keep iff bits >= thr,
// thr = round(dp * 65536).  One Philox4x32-10 call covers two of a lane's 128-bit chunks:
//   counter = (32*call + lane, slot | stream<<2 | offset<<3, lo32(g), hi32(g) ^ (offset>>29)),  key = (lo32(seed)^'DROP', hi32(seed))
// slot 0/1/2 = s/p/o, stream 0 = the positive call, 1 = the negative call, g = index of the triple in its call.
// ---------------------------------------------------------------------------------------------------
struct DropArgs {
  uint32_t thr;   // 0 = dropout off
  float scale;    // 1 / (1 - dp)
  uint32_t k0, k1, ylo, whi;
};

__device__ __forceinline__ uint4 dp_philox(const DropArgs& d, int slot, int stream, uint64_t g, int lane, int call) {
  return philox4x32_10(make_uint4((uint32_t)(32 * call + lane), d.ylo | (uint32_t)slot | ((uint32_t)stream << 2), (uint32_t)g,
                                  (uint32_t)(g >> 32) ^ d.whi),
                       make_uint2(d.k0, d.k1));
}
__device__ __forceinline__ uint32_t dp_word(const uint4& r, int h, int e) { return h ? (e < 2 ? r.z : r.w) : (e < 2 ? r.x : r.y); }

// keep bits of this lane's fragment of one row instance: bit 4*t + e for chunk slot t, component e
template <int KIND, int NC>
__device__ __forceinline__ uint32_t dp_bits(const DropArgs& d, int slot, int stream, uint64_t g, int lane) {
  constexpr int NT = Frag<KIND, NC>::NT;
  uint32_t bits = 0;
#pragma unroll
  for (int c = 0; c < (NT + 1) / 2; ++c) {
    const uint4 r = dp_philox(d, slot, stream, g, lane, c);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = 2 * c + h;
      if (t < NT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t u = (dp_word(r, h, e) >> (16 * (e & 1))) & 0xFFFFu;
          bits |= (u >= d.thr ? 1u : 0u) << (4 * t + e);
        }
      }
    }
  }
  return bits;
}
// f <- f * keep * scale (forward on gathered rows, backward on gradient rows: the same linear map)
template <int KIND, int NC>
__device__ __forceinline__ void dp_apply(Frag<KIND, NC>& f, uint32_t bits, float scale) {
#pragma unroll
  for (int t = 0; t < Frag<KIND, NC>::NT; ++t) {
    KGEB_FOR4({ comp(f.v[t], _e) = ((bits >> (4 * t + _e)) & 1u) ? __fmul_rn(comp(f.v[t], _e), scale) : 0.f; })
  }
}
// element-wise form (HolE staging): x = element index in a plain (non-ComplEx) row
__device__ __forceinline__ bool dp_keep_elem(const DropArgs& d, int slot, int stream, uint64_t g, int x) {
  const int chunk = x >> 2, e = x & 3, t = chunk >> 5;
  const uint4 r = dp_philox(d, slot, stream, g, chunk & 31, t >> 1);
  return ((dp_word(r, t & 1, e) >> (16 * (e & 1))) & 0xFFFFu) >= d.thr;
}

struct ScoreArgs {
  const float* E; const float* R; int ldE, ldR; RowGeom ge, gr; float gamma, param;
  DropArgs dp; int dp_stream;
};

struct TrainArgs {
  ScoreArgs sc;
  const int32_t* pos; int64_t B; const int32_t* neg; int K;
  int loss_kind; float margin; float loss_weight; int corrupt_mode; uint64_t seed, offset; uint32_t N;
  float* loss; float* pos_scores; float* neg_scores; int32_t* neg_out;
  float* gE; float* gR; uint8_t* tE; uint8_t* tR;
  double* partials; unsigned int* ticket;
  // sort-segmented entity backward (sortseg_kernels.cu): when set, the kernel writes dL/dscore of every scored triple
  // (coef_out[g] for positive g, coef_out[B + g*K + j] for its negative j) and leaves the ENTITY gradient rows to the
  // segment kernel; relation gradients keep the in-register group sum + one vector reduction per group
  float* coef_out;
};

// ---------------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------------
int check_tables(kgeb_ctx* ctx, const void* E, const void* R) {
  KGEB_REQUIRE(ctx, E && R, KGEB_ERR_INVALID, "null table pointer");
  if (ctx->kind == KGEB_HOLE) return KGEB_OK;
  KGEB_REQUIRE(ctx, kgeb_aligned16(E) && kgeb_aligned16(R), KGEB_ERR_ALIGNMENT, "table pointers must be 16-byte aligned");
  KGEB_REQUIRE(ctx, ctx->ldE % 4 == 0 && ctx->ldR % 4 == 0, KGEB_ERR_ALIGNMENT, "row strides must be multiples of 4 floats");
  if (ctx->kind == KGEB_ROTATE) {
    KGEB_REQUIRE(ctx, ctx->we == 2 * ctx->wr, KGEB_ERR_UNSUPPORTED, "RotatE needs entity width == 2 * relation width (models.py:441-448)");
    KGEB_REQUIRE(ctx, ctx->we % 8 == 0, KGEB_ERR_ALIGNMENT, "RotatE needs dim % 4 == 0 (stored entity width % 8 == 0)");
    return KGEB_OK;
  }
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
  a.E = E; a.R = R; a.ldE = ctx->ldE; a.ldR = ctx->ldR; a.gamma = ctx->gamma; a.param = ctx->param;
  if (ctx->kind == KGEB_COMPLEX || ctx->kind == KGEB_ROTATE) {
    a.ge.nch = ctx->we / 8; a.ge.half4 = ctx->we / 8;
  } else {
    a.ge.nch = ctx->we / 4; a.ge.half4 = 0;
  }
  a.gr = a.ge;
  if (ctx->kind == KGEB_ROTATE) { a.gr.nch = ctx->wr / 4; a.gr.half4 = 0; }
  *nc_out = (a.ge.nch + 31) / 32;
  a.dp.thr = ctx->dp_thr; a.dp.scale = ctx->dp_scale;
  a.dp.k0 = (uint32_t)ctx->dp_seed ^ 0x44524F50u; a.dp.k1 = (uint32_t)(ctx->dp_seed >> 32);
  a.dp.ylo = (uint32_t)(ctx->dp_offset << 3); a.dp.whi = (uint32_t)(ctx->dp_offset >> 29);
  a.dp_stream = ctx->dp_stream;
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
    default: return kgeb_fail(ctx, KGEB_ERR_UNSUPPORTED, "the fused train kernel serves DistMult, TransE and ComplEx; other kinds train through score_fwd -> loss_fwd_bwd -> score_bwd"); \
  }
// every register-fragment kind (the scoring kernels): the four above + the sibling models
#define KGEB_DISPATCH_KIND_ALL(KINDV, ...)                                               \
  switch (KINDV) {                                                                       \
    case KGEB_DISTMULT: { constexpr int KIND = KGEB_DISTMULT; __VA_ARGS__; } break;      \
    case KGEB_TRANSE_L1: { constexpr int KIND = KGEB_TRANSE_L1; __VA_ARGS__; } break;    \
    case KGEB_TRANSE_L2: { constexpr int KIND = KGEB_TRANSE_L2; __VA_ARGS__; } break;    \
    case KGEB_COMPLEX: { constexpr int KIND = KGEB_COMPLEX; __VA_ARGS__; } break;        \
    case KGEB_ROTATE: { constexpr int KIND = KGEB_ROTATE; __VA_ARGS__; } break;          \
    case KGEB_PROTATE: { constexpr int KIND = KGEB_PROTATE; __VA_ARGS__; } break;        \
    case KGEB_MODE_L1: { constexpr int KIND = KGEB_MODE_L1; __VA_ARGS__; } break;        \
    case KGEB_MODE_L2: { constexpr int KIND = KGEB_MODE_L2; __VA_ARGS__; } break;        \
    default: return kgeb_fail(ctx, KGEB_ERR_INVALID, "unknown score kind");          \
  }


}  // namespace
