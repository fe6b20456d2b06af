// hole_fft_kernels.cu -- HolE through the Fourier domain, the formulation the reference itself uses (models.py:233-234:
// ifft(conj(fft(s)) * fft(o)) in complex64), for power-of-two widths.  The direct O(d^2) kernels in score_kernels.cu stay
// the default (their fp32 sequential contract is what the bit-exact tests pin); kgeb_set_hole_mode(ctx, KGEB_HOLE_FFT)
// selects these: ~3 length-d FFTs per triple instead of 3 d^2 multiply-adds.
//
//   Z = F(s + i o)            one complex FFT gives both spectra:  S_f = (Z_f + conj Z_-f)/2,  O_f = -i (Z_f - conj Z_-f)/2
//   score = (1/d) Re sum_f conj(P_f) conj(S_f) O_f                                   (Parseval; no inverse transform)
//   ds + i do = F^-1( c conj(P) O  +  i c P S )      one complex inverse FFT gives both real gradient rows
//   dp = F^-1( sum_run c conj(S) O )                 accumulated in the Fourier domain over a run of triples that share the
//                                                    relation (the K negatives of one positive), one inverse FFT per run
// A warp owns a contiguous range of triples, so runs stay together; P = F(p) is recomputed only when the relation changes.
// FFT: radix-4 (+ one radix-2 stage for odd log2 d) Stockham autosort in shared memory (two d-complex buffers per warp),
// twiddles from a per-block table.
#include "frag.cuh"

namespace {

constexpr int HF_WARPS = 4;

struct HoleFftArgs {
  const float* E; const float* R; int ldE, ldR, d, logd;
  DropArgs dp; int stream;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ping-pong between x and y: returns the buffer that holds the result.  INV uses conjugated twiddles (no 1/d scale).
// Radix-4 Stockham autosort stages (each pass reads and writes every element once, so half the shared-memory traffic of
// radix-2), plus one radix-2 stage when log2(d) is odd.  tw[k] = exp(-2 pi i k / d), k < d; behind it (tw + d) the compact tables of the
// radix-4 stages: stage p keeps w_m[k] = exp(-2 pi i m k / (4 p)), m = 1..3, k < p, at tw[d + (p - 1) + (m - 1) p + k] -- the same values
// the strided reads tw[(m k) << sh] returned, but consecutive in k (the strided reads cost 4-8 wavefronts each at p = 4 and 16).
// Bank conflicts of the stores: only the p = 4 stage has them (a warp's 32 outputs of one store fall on 8 banks: 8 wavefronts instead
// of 2), so that stage writes through the permutation hf_swz and the stage after it reads through it; every other stage is conflict free
// with plain indices (ncu before: 43 % of the shared-memory wavefronts of these kernels were conflicts).
__device__ __forceinline__ int hf_swz(int j) { return j ^ (((j >> 4) & 3) << 2); }

// one radix-4 Stockham stage.  P > 0: compile-time p (masks, twiddle offsets and the store shape fold away); P == 0: p at run time.
// RD / WR: read / write through hf_swz (the p = 16 stage reads what the p = 4 stage wrote).
template <bool INV, int P, bool RD, bool WR>
__device__ __forceinline__ void hf_stage4(const float2* __restrict__ x, float2* __restrict__ y, const float2* __restrict__ ctw, int d, int p_rt, int lane) {
  const int p = P > 0 ? P : p_rt, t = d >> 2;
  for (int i = lane; i < t; i += 32) {
    const int k = i & (p - 1), j = ((i - k) << 2) + k;
    float2 w1 = ctw[k], w2 = ctw[p + k], w3 = ctw[2 * p + k];
    if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
    const float2 u0 = x[RD ? hf_swz(i) : i], u1 = cmul(x[RD ? hf_swz(i + t) : i + t], w1), u2 = cmul(x[RD ? hf_swz(i + 2 * t) : i + 2 * t], w2),
                 u3 = cmul(x[RD ? hf_swz(i + 3 * t) : i + 3 * t], w3);
    const float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y), v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
    const float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
    const float2 df = make_float2(u1.x - u3.x, u1.y - u3.y);
    const float2 v3 = INV ? make_float2(-df.y, df.x) : make_float2(df.y, -df.x);      // (+/-) i * (u1 - u3)
    const float2 y0 = make_float2(v0.x + v2.x, v0.y + v2.y), y1 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 y2 = make_float2(v0.x - v2.x, v0.y - v2.y), y3 = make_float2(v1.x - v3.x, v1.y - v3.y);
    if (P == 1) {                                        // the four outputs are adjacent: two 128-bit stores
      *reinterpret_cast<float4*>(y + j) = make_float4(y0.x, y0.y, y1.x, y1.y);
      *reinterpret_cast<float4*>(y + j + 2) = make_float4(y2.x, y2.y, y3.x, y3.y);
    } else if (WR) {
      y[hf_swz(j)] = y0; y[hf_swz(j + p)] = y1; y[hf_swz(j + 2 * p)] = y2; y[hf_swz(j + 3 * p)] = y3;
    } else {
      y[j] = y0; y[j + p] = y1; y[j + 2 * p] = y2; y[j + 3 * p] = y3;
    }
  }
}

// widths 32 .. 1024 (the launchers check): the p = 1 and p = 4 stages always run and the p = 4 stage is never the last one
template <bool INV>
__device__ __forceinline__ float2* warp_fft(float2* x, float2* y, const float2* __restrict__ tw, int d, int logd, int lane) {
  const float2* ctw = tw + d;
  float2* tmp;
  __syncwarp();
  hf_stage4<INV, 1, false, false>(x, y, ctw, d, 1, lane);
  tmp = x; x = y; y = tmp;
  __syncwarp();
  hf_stage4<INV, 4, false, true>(x, y, ctw + 3, d, 4, lane);
  tmp = x; x = y; y = tmp;
  int p = 16, lp = 4;
  if (64 <= d) {
    __syncwarp();
    hf_stage4<INV, 16, true, false>(x, y, ctw + 15, d, 16, lane);
    tmp = x; x = y; y = tmp;
    for (p = 64, lp = 6; 4 * p <= d; p <<= 2, lp += 2) {
      __syncwarp();
      hf_stage4<INV, 0, false, false>(x, y, ctw + (p - 1), d, p, lane);
      tmp = x; x = y; y = tmp;
    }
  }
  if (p < d) {                                           // log2(d) odd: one radix-2 stage with p = d / 2
    const int t = d >> 1, sh = logd - 1 - lp;
    const bool rd_swz = p == 16;                         // d == 32: this stage reads what the p = 4 stage wrote
    __syncwarp();
    for (int i = lane; i < t; i += 32) {
      const int k = i & (p - 1), j = ((i - k) << 1) + k;
      float2 w = tw[k << sh];
      if (INV) w.y = -w.y;
      const float2 u0 = x[rd_swz ? hf_swz(i) : i], u1 = cmul(x[rd_swz ? hf_swz(i + t) : i + t], w);
      y[j] = make_float2(u0.x + u1.x, u0.y + u1.y);
      y[j + p] = make_float2(u0.x - u1.x, u0.y - u1.y);
    }
    tmp = x; x = y; y = tmp;
  }
  __syncwarp();
  return x;
}

// multipliers of the four elements of chunk c of row-slot `slot` of triple g (one Philox call per chunk, plain-row layout)
__device__ __forceinline__ void hf_mult4(const DropArgs& dp, int slot, int stream, uint64_t g, int c, float (&mult)[4]) {
  if (!dp.thr) { mult[0] = mult[1] = mult[2] = mult[3] = 1.f; return; }
  const int t = c >> 5;
  const uint4 r = dp_philox(dp, slot, stream, g, c & 31, t >> 1);
#pragma unroll
  for (int e = 0; e < 4; ++e) mult[e] = (((dp_word(r, t & 1, e) >> (16 * (e & 1))) & 0xFFFFu) >= dp.thr) ? dp.scale : 0.f;
}

struct HfWarp {
  float2 *a, *b, *P, *acc;     // two ping-pong buffers, the relation spectrum, the dp accumulator (backward only)
};

// z = dp(s) + i dp(o) into buf
__device__ __forceinline__ void hf_load_so(const HoleFftArgs& a, const float* es, const float* eo, uint64_t g, float2* buf, int lane) {
  for (int c = lane; c < (a.d >> 2); c += 32) {
    float ms[4], mo[4];
    hf_mult4(a.dp, 0, a.stream, g, c, ms);
    hf_mult4(a.dp, 2, a.stream, g, c, mo);
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(es) + c), o4 = __ldg(reinterpret_cast<const float4*>(eo) + c);
    buf[4 * c + 0] = make_float2(__fmul_rn(s4.x, ms[0]), __fmul_rn(o4.x, mo[0]));
    buf[4 * c + 1] = make_float2(__fmul_rn(s4.y, ms[1]), __fmul_rn(o4.y, mo[1]));
    buf[4 * c + 2] = make_float2(__fmul_rn(s4.z, ms[2]), __fmul_rn(o4.z, mo[2]));
    buf[4 * c + 3] = make_float2(__fmul_rn(s4.w, ms[3]), __fmul_rn(o4.w, mo[3]));
  }
}
__device__ __forceinline__ void hf_load_p(const HoleFftArgs& a, const float* rp, uint64_t g, float2* buf, int lane) {
  for (int c = lane; c < (a.d >> 2); c += 32) {
    float mp[4];
    hf_mult4(a.dp, 1, a.stream, g, c, mp);
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(rp) + c);
    buf[4 * c + 0] = make_float2(__fmul_rn(p4.x, mp[0]), 0.f);
    buf[4 * c + 1] = make_float2(__fmul_rn(p4.y, mp[1]), 0.f);
    buf[4 * c + 2] = make_float2(__fmul_rn(p4.z, mp[2]), 0.f);
    buf[4 * c + 3] = make_float2(__fmul_rn(p4.w, mp[3]), 0.f);
  }
}
__device__ __forceinline__ void hf_split(const float2* Z, int f, int d, float2& S, float2& O) {
  const float2 zf = Z[f], zm = cconj(Z[(d - f) & (d - 1)]);
  S = make_float2(0.5f * (zf.x + zm.x), 0.5f * (zf.y + zm.y));
  const float2 df = make_float2(zf.x - zm.x, zf.y - zm.y);     // O = -i * df / 2
  O = make_float2(0.5f * df.y, -0.5f * df.x);
}

__device__ __forceinline__ void hf_setup(const HoleFftArgs& a, float2* smem, float2*& tw, HfWarp& w) {
  tw = smem;                                   // d twiddles exp(-2 pi i k / d), shared by the block, then the compact radix-4 tables (< d entries)
  for (int k = threadIdx.x; k < a.d; k += blockDim.x) {
    float sn, cs;
    sincospif(-2.f * (float)k / (float)a.d, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  for (int p = 1; 4 * p <= a.d; p <<= 2)        // w_m[k] = exp(-2 pi i m k / (4 p)): the argument is an exact fp32 ratio, as in tw[]
    for (int e = threadIdx.x; e < 3 * p; e += blockDim.x) {
      const int m = e / p + 1, k = e - (m - 1) * p;
      float sn, cs;
      sincospif(-2.f * (float)(m * k) / (float)(4 * p), &sn, &cs);
      tw[a.d + (p - 1) + e] = make_float2(cs, sn);
    }
  float2* base = smem + 2 * a.d + (size_t)(threadIdx.x >> 5) * 4 * a.d;
  w.a = base; w.b = base + a.d; w.P = base + 2 * a.d; w.acc = base + 3 * a.d;
  __syncthreads();
}

template <bool DP>
__global__ void __launch_bounds__(HF_WARPS * 32) k_hole_fft_fwd(HoleFftArgs a, const int32_t* __restrict__ idx, int64_t m, float* __restrict__ scores) {
  extern __shared__ __align__(16) float2 hf_smem[];
  float2* tw; HfWarp w;
  hf_setup(a, hf_smem, tw, w);
  const int lane = threadIdx.x & 31, d = a.d;
  const int64_t nw = (int64_t)gridDim.x * HF_WARPS, wid = (int64_t)blockIdx.x * HF_WARPS + (threadIdx.x >> 5);
  const int64_t per = (m + nw - 1) / nw, i0 = wid * per, i1 = i0 + per < m ? i0 + per : m;
  int64_t prel = -1;
  const float inv_d = 1.f / (float)d;
  for (int64_t i = i0; i < i1; ++i) {
    const int64_t s = idx[3 * i], p = idx[3 * i + 1], o = idx[3 * i + 2];
    if (p != prel || DP) {                     // with dropout the relation row is masked per triple
      hf_load_p(a, a.R + p * a.ldR, (uint64_t)i, w.a, lane);
      const float2* Pf = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
      for (int f = lane; f < d; f += 32) w.P[f] = Pf[f];
      prel = p;
    }
    __syncwarp();
    hf_load_so(a, a.E + s * a.ldE, a.E + o * a.ldE, (uint64_t)i, w.a, lane);
    const float2* Z = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
    float part = 0.f;
    for (int f = lane; f < d; f += 32) {
      float2 S, O;
      hf_split(Z, f, d, S, O);
      const float2 q = cmul(cconj(w.P[f]), cconj(S));
      part += q.x * O.x - q.y * O.y;            // Re(q * O)
    }
    part = warp_sum(part);
    if (lane == 0) scores[i] = part * inv_d;
    __syncwarp();
  }
}

template <bool DP>
__global__ void __launch_bounds__(HF_WARPS * 32, 6) k_hole_fft_bwd(HoleFftArgs a, const int32_t* __restrict__ idx, int64_t m, const float* __restrict__ dscore,
                                                                float* gE, float* gR, uint8_t* tE, uint8_t* tR) {
  extern __shared__ __align__(16) float2 hf_smem[];
  float2* tw; HfWarp w;
  hf_setup(a, hf_smem, tw, w);
  const int lane = threadIdx.x & 31, d = a.d;
  const int64_t nw = (int64_t)gridDim.x * HF_WARPS, wid = (int64_t)blockIdx.x * HF_WARPS + (threadIdx.x >> 5);
  const int64_t per = (m + nw - 1) / nw, i0 = wid * per, i1 = i0 + per < m ? i0 + per : m;
  const float inv_d = 1.f / (float)d;
  int64_t prel = -1, pg = -1;
  bool acc_live = false;
  // dp of the finished run: gR[prel] += Re F^-1(acc) / d  (masked by the relation row's dropout when armed: runs have length 1 then)
  auto flush = [&]() {
    if (!acc_live) return;
    for (int f = lane; f < d; f += 32) w.a[f] = w.acc[f];
    const float2* g = warp_fft<true>(w.a, w.b, tw, d, a.logd, lane);
    float* dst = gR + prel * a.ldR;
    for (int c = lane; c < (d >> 2); c += 32) {
      float mp[4];
      hf_mult4(a.dp, 1, a.stream, (uint64_t)pg, c, mp);
      const float4 v = make_float4(g[4 * c].x * inv_d * mp[0], g[4 * c + 1].x * inv_d * mp[1], g[4 * c + 2].x * inv_d * mp[2], g[4 * c + 3].x * inv_d * mp[3]);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add4(dst + 4 * c, v);
    }
    __syncwarp();
    acc_live = false;
  };
  for (int64_t i = i0; i < i1; ++i) {
    const int64_t s = idx[3 * i], p = idx[3 * i + 1], o = idx[3 * i + 2];
    if (lane == 0) { tE[s] = 1; tE[o] = 1; tR[p] = 1; }
    const float c = dscore[i];
    if (c == 0.f) continue;
    if (p != prel || DP) {
      flush();
      hf_load_p(a, a.R + p * a.ldR, (uint64_t)i, w.a, lane);
      const float2* Pf = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
      for (int f = lane; f < d; f += 32) { w.P[f] = Pf[f]; w.acc[f] = make_float2(0.f, 0.f); }
      prel = p;
    }
    pg = i;
    __syncwarp();
    hf_load_so(a, a.E + s * a.ldE, a.E + o * a.ldE, (uint64_t)i, w.a, lane);
    float2* Z = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
    float2* Gb = (Z == w.a) ? w.b : w.a;       // the other buffer takes G = c conj(P) O + i c P S
    for (int f = lane; f < d; f += 32) {
      float2 S, O;
      hf_split(Z, f, d, S, O);
      const float2 P = w.P[f];
      const float2 gs = cmul(cconj(P), O), go = cmul(P, S);
      Gb[f] = make_float2(c * (gs.x - go.y), c * (gs.y + go.x));
      const float2 gp = cmul(cconj(S), O);
      w.acc[f].x += c * gp.x;
      w.acc[f].y += c * gp.y;
    }
    acc_live = true;
    __syncwarp();
    const float2* g = warp_fft<true>(Gb, Z, tw, d, a.logd, lane);
    float* ds = gE + s * a.ldE;
    float* dO = gE + o * a.ldE;
    for (int cch = lane; cch < (d >> 2); cch += 32) {
      float ms[4], mo[4];
      hf_mult4(a.dp, 0, a.stream, (uint64_t)i, cch, ms);
      hf_mult4(a.dp, 2, a.stream, (uint64_t)i, cch, mo);
      const float2 g0 = g[4 * cch], g1 = g[4 * cch + 1], g2 = g[4 * cch + 2], g3 = g[4 * cch + 3];
      red_add4(ds + 4 * cch, make_float4(g0.x * inv_d * ms[0], g1.x * inv_d * ms[1], g2.x * inv_d * ms[2], g3.x * inv_d * ms[3]));
      red_add4(dO + 4 * cch, make_float4(g0.y * inv_d * mo[0], g1.y * inv_d * mo[1], g2.y * inv_d * mo[2], g3.y * inv_d * mo[3]));
    }
    __syncwarp();
  }
  flush();
}

// The whole step for the POINTWISE losses (a triple's loss coefficient depends on its own score only): Z = F(s + i o) is
// transformed once, the score, the loss term and its derivative follow, and the same spectra give the gradients.  Triples are
// the concatenation [positives ; negatives]; negatives come from `neg` or from the in-kernel Philox draw (neg == NULL).
struct HoleTrainArgs {
  const int32_t* pos; int64_t B; const int32_t* neg; int K;
  int loss_kind; float margin, w; int corrupt_mode; uint64_t seed, offset; uint32_t N;
  float* loss; float* pos_scores; float* neg_scores; int32_t* neg_out;
  float* gE; float* gR; uint8_t* tE; uint8_t* tR;
  double* partials; unsigned int* ticket;
};

template <bool DP>
__global__ void __launch_bounds__(HF_WARPS * 32, 5) k_hole_fft_train(HoleFftArgs a, HoleTrainArgs t) {
  extern __shared__ __align__(16) float2 hf_smem[];
  float2* tw; HfWarp w;
  hf_setup(a, hf_smem, tw, w);
  const int lane = threadIdx.x & 31, d = a.d;
  const int64_t m = t.B * (1 + (int64_t)t.K);
  const int64_t nw = (int64_t)gridDim.x * HF_WARPS, wid = (int64_t)blockIdx.x * HF_WARPS + (threadIdx.x >> 5);
  const int64_t per = (m + nw - 1) / nw, i0 = wid * per, i1 = i0 + per < m ? i0 + per : m;
  const float inv_d = 1.f / (float)d;
  const float sP = (float)((double)t.w / (double)t.B), sN = (float)((double)t.w / ((double)t.B * (double)t.K));
  double lp = 0.0, ln = 0.0;
  int64_t prel = -1, pg = -1;
  int pstream = 0;
  bool acc_live = false;
  auto flush = [&]() {
    if (!acc_live) return;
    for (int f = lane; f < d; f += 32) w.a[f] = w.acc[f];
    const float2* g = warp_fft<true>(w.a, w.b, tw, d, a.logd, lane);
    float* dst = t.gR + prel * a.ldR;
    for (int c = lane; c < (d >> 2); c += 32) {
      float mp[4];
      hf_mult4(a.dp, 1, pstream, (uint64_t)pg, c, mp);
      const float4 v = make_float4(g[4 * c].x * inv_d * mp[0], g[4 * c + 1].x * inv_d * mp[1], g[4 * c + 2].x * inv_d * mp[2], g[4 * c + 3].x * inv_d * mp[3]);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add4(dst + 4 * c, v);
    }
    __syncwarp();
    acc_live = false;
  };
  for (int64_t x = i0; x < i1; ++x) {
    const bool is_pos = x < t.B;
    const int64_t g = is_pos ? x : x - t.B;            // index of the triple inside its call (dropout counter)
    int64_t s, p, o;
    if (is_pos) {
      s = t.pos[3 * x]; p = t.pos[3 * x + 1]; o = t.pos[3 * x + 2];
    } else if (t.neg) {
      s = t.neg[3 * g]; p = t.neg[3 * g + 1]; o = t.neg[3 * g + 2];
    } else {
      const int64_t i = g / t.K;
      int ns, no;
      philox_corrupt(t.pos[3 * i], t.pos[3 * i + 2], (uint64_t)g, t.seed, t.offset, t.N, t.corrupt_mode, ns, no);
      s = ns; p = t.pos[3 * i + 1]; o = no;
      if (t.neg_out && lane == 0) { t.neg_out[3 * g] = ns; t.neg_out[3 * g + 1] = (int32_t)p; t.neg_out[3 * g + 2] = no; }
    }
    if (lane == 0) { t.tE[s] = 1; t.tE[o] = 1; t.tR[p] = 1; }
    HoleFftArgs ac = a;
    ac.stream = is_pos ? 0 : 1;
    if (p != prel || DP) {
      flush();
      hf_load_p(ac, a.R + p * a.ldR, (uint64_t)g, w.a, lane);
      const float2* Pf = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
      for (int f = lane; f < d; f += 32) { w.P[f] = Pf[f]; w.acc[f] = make_float2(0.f, 0.f); }
      prel = p;
    }
    pg = g; pstream = ac.stream;
    __syncwarp();
    hf_load_so(ac, a.E + s * a.ldE, a.E + o * a.ldE, (uint64_t)g, w.a, lane);
    float2* Z = warp_fft<false>(w.a, w.b, tw, d, a.logd, lane);
    float part = 0.f;
    for (int f = lane; f < d; f += 32) {
      float2 S, O;
      hf_split(Z, f, d, S, O);
      const float2 q = cmul(cconj(w.P[f]), cconj(S));
      part += q.x * O.x - q.y * O.y;
    }
    const float score = warp_sum(part) * inv_d;
    float val, der;
    pointwise_terms(t.loss_kind, t.margin, score, is_pos, val, der);
    if (lane == 0) {
      if (is_pos) { lp += (double)val; if (t.pos_scores) t.pos_scores[x] = score; }
      else { ln += (double)val; if (t.neg_scores) t.neg_scores[g] = score; }
    }
    const float c = der * (is_pos ? sP : sN);
    if (c == 0.f) continue;                            // warp-uniform (the score is a full butterfly sum)
    float2* Gb = (Z == w.a) ? w.b : w.a;
    for (int f = lane; f < d; f += 32) {
      float2 S, O;
      hf_split(Z, f, d, S, O);
      const float2 P = w.P[f];
      const float2 gs = cmul(cconj(P), O), go = cmul(P, S);
      Gb[f] = make_float2(c * (gs.x - go.y), c * (gs.y + go.x));
      const float2 gp = cmul(cconj(S), O);
      w.acc[f].x += c * gp.x;
      w.acc[f].y += c * gp.y;
    }
    acc_live = true;
    __syncwarp();
    const float2* gr = warp_fft<true>(Gb, Z, tw, d, a.logd, lane);
    float* ds = t.gE + s * a.ldE;
    float* dO = t.gE + o * a.ldE;
    for (int cch = lane; cch < (d >> 2); cch += 32) {
      float ms[4], mo[4];
      hf_mult4(a.dp, 0, ac.stream, (uint64_t)g, cch, ms);
      hf_mult4(a.dp, 2, ac.stream, (uint64_t)g, cch, mo);
      const float2 g0 = gr[4 * cch], g1 = gr[4 * cch + 1], g2 = gr[4 * cch + 2], g3 = gr[4 * cch + 3];
      red_add4(ds + 4 * cch, make_float4(g0.x * inv_d * ms[0], g1.x * inv_d * ms[1], g2.x * inv_d * ms[2], g3.x * inv_d * ms[3]));
      red_add4(dO + 4 * cch, make_float4(g0.y * inv_d * mo[0], g1.y * inv_d * mo[1], g2.y * inv_d * mo[2], g3.y * inv_d * mo[3]));
    }
    __syncwarp();
  }
  flush();
  block_finalize_loss(lp, ln, t.partials, t.ticket, (double)t.w / (double)t.B, (double)t.w / ((double)t.B * (double)t.K), t.loss);
}

bool hf_args(const kgeb_ctx* ctx, const float* E, const float* R, HoleFftArgs* a) {
  const int d = ctx->we;
  if (d != ctx->wr || d < 32 || d > 1024 || (d & (d - 1)) != 0) return false;
  if ((ctx->ldE & 3) || (ctx->ldR & 3) || !kgeb_aligned16(E) || !kgeb_aligned16(R)) return false;
  a->E = E; a->R = R; a->ldE = ctx->ldE; a->ldR = ctx->ldR; a->d = d;
  a->logd = 0;
  while ((1 << a->logd) < d) ++a->logd;
  a->dp.thr = ctx->dp_thr; a->dp.scale = ctx->dp_scale;
  a->dp.k0 = (uint32_t)ctx->dp_seed ^ 0x44524F50u; a->dp.k1 = (uint32_t)(ctx->dp_seed >> 32);
  a->dp.ylo = (uint32_t)(ctx->dp_offset << 3); a->dp.whi = (uint32_t)(ctx->dp_offset >> 29);
  a->stream = ctx->dp_stream;
  return true;
}
size_t hf_smem_bytes(int d) { return ((size_t)2 * d + (size_t)HF_WARPS * 4 * d) * sizeof(float2); }
// grid = SMs x the blocks of THIS kernel that are really resident per SM (asked from the runtime: registers can bind before shared memory
// does -- with 125 registers the fused kernel fits 4 blocks where the shared-memory budget alone says 5, and a grid of 5 per SM would run
// a second wave a quarter full), each block a contiguous range of triples
template <typename Kern>
int hf_grid(const kgeb_ctx* ctx, Kern kern, int64_t m, size_t smem) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HF_WARPS * 32, smem) != cudaSuccess) per_sm = (int)((200u * 1024u) / smem);
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (int64_t)ctx->num_sms * per_sm;
  const int64_t need = (m + HF_WARPS * 4 - 1) / (HF_WARPS * 4);     // at least ~4 triples per warp
  if (blocks > need) blocks = need;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

// both return KGEB_ERR_UNSUPPORTED (without touching ctx->err) when the shape does not fit: the caller falls back to the direct kernels
int kgeb_internal_hole_fft_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores, cudaStream_t st) {
  HoleFftArgs a;
  if (!hf_args(ctx, E, R, &a)) return KGEB_ERR_UNSUPPORTED;
  const size_t sm = hf_smem_bytes(a.d);
  if (a.dp.thr) {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fft_fwd<true><<<hf_grid(ctx, k_hole_fft_fwd<true>, m, sm), HF_WARPS * 32, sm, st>>>(a, idx, m, scores);
  } else {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fft_fwd<false><<<hf_grid(ctx, k_hole_fft_fwd<false>, m, sm), HF_WARPS * 32, sm, st>>>(a, idx, m, scores);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

int kgeb_internal_hole_fft_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, const float* dscore, float* gE, float* gR,
                               uint8_t* tE, uint8_t* tR, cudaStream_t st) {
  HoleFftArgs a;
  if (!hf_args(ctx, E, R, &a) || !kgeb_aligned16(gE) || !kgeb_aligned16(gR)) return KGEB_ERR_UNSUPPORTED;
  const size_t sm = hf_smem_bytes(a.d);
  if (a.dp.thr) {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fft_bwd<true><<<hf_grid(ctx, k_hole_fft_bwd<true>, m, sm), HF_WARPS * 32, sm, st>>>(a, idx, m, dscore, gE, gR, tE, tR);
  } else {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hole_fft_bwd<false><<<hf_grid(ctx, k_hole_fft_bwd<false>, m, sm), HF_WARPS * 32, sm, st>>>(a, idx, m, dscore, gE, gR, tE, tR);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

// fused step; KGEB_ERR_UNSUPPORTED (ctx->err untouched) when the width or the loss is not served
int kgeb_internal_hole_fft_train(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* pos, int64_t B, const int32_t* neg, int K, int loss_kind,
                                 float margin, float loss_weight, int corrupt_mode, uint64_t seed, uint64_t offset, float* loss, float* pos_scores,
                                 float* neg_scores, int32_t* neg_out, float* gE, float* gR, uint8_t* tE, uint8_t* tR, cudaStream_t st) {
  HoleFftArgs a;
  if (KGEB_IS_PAIRWISE(loss_kind) || !hf_args(ctx, E, R, &a) || !kgeb_aligned16(gE) || !kgeb_aligned16(gR)) return KGEB_ERR_UNSUPPORTED;
  HoleTrainArgs t;
  t.pos = pos; t.B = B; t.neg = neg; t.K = K; t.loss_kind = loss_kind; t.margin = margin; t.w = loss_weight; t.corrupt_mode = corrupt_mode;
  t.seed = seed; t.offset = offset; t.N = (uint32_t)ctx->N; t.loss = loss; t.pos_scores = pos_scores; t.neg_scores = neg_scores; t.neg_out = neg_out;
  t.gE = gE; t.gR = gR; t.tE = tE; t.tR = tR; t.partials = ctx->d_partials; t.ticket = ctx->d_ticket;
  const size_t sm = hf_smem_bytes(a.d);
  if (a.dp.thr) {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_train<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int grid = hf_grid(ctx, k_hole_fft_train<true>, B * (1 + (int64_t)K), sm);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    k_hole_fft_train<true><<<grid, HF_WARPS * 32, sm, st>>>(a, t);
  } else {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_hole_fft_train<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int grid = hf_grid(ctx, k_hole_fft_train<false>, B * (1 + (int64_t)K), sm);
    if (grid > KGEB_MAX_PARTIAL_BLOCKS) grid = KGEB_MAX_PARTIAL_BLOCKS;
    k_hole_fft_train<false><<<grid, HF_WARPS * 32, sm, st>>>(a, t);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
