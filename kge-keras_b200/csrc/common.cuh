// common.cuh -- shared host/device helpers of libkgeb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/kgeb200.h"

#define KGEB_WARPS_PER_BLOCK 8
#define KGEB_THREADS (KGEB_WARPS_PER_BLOCK * 32)
#define KGEB_MAX_PARTIAL_BLOCKS 4096
#define KGEB_TIMED_KERNELS 6
#define KGEB_TIMED_POOL 64   // event pairs kept per timed kernel: kgeb_kernel_ms_sum adds up the launches since its last call
#define KGEB_SORT_SCRATCH_BYTES (8u << 20)

struct kgeb_ctx {
  int device;
  int64_t N, R;
  int we, wr, ldE, ldR;
  int kind;
  float gamma;
  float param;   // pRotatE modulus (kgeb_set_score_param), 1 otherwise
  int hole_fft;  // kgeb_set_hole_mode: HolE through the Fourier domain for power-of-two widths
  float hake_ps, hake_pw, hake_mw;   // HAKE: embedding_range / pi (the phase DIVISOR), phase_weight, mod_weight (kgeb_set_hake_params)
  int64_t hake_group;                // HAKE inference: rows per Model.predict batch (kgeb_set_hake_call_rows), 0 = the whole call is one batch
  int num_sms;
  int64_t launches;
  char err[512];
  // fixed scratch, allocated once in kgeb_create (nothing allocates afterwards)
  double* d_partials;        // [KGEB_MAX_PARTIAL_BLOCKS * 2] per-block loss partial sums
  unsigned int* d_ticket;    // last-block-done counter (self-resetting)
  unsigned char* d_sort_scratch;   // [KGEB_SORT_SCRATCH_BYTES] sorted chunks + prefix sums of the all-pairs hinge loss; split-K slices of kgeb_sgemm
  // dropout state (kgeb_set_dropout): thr = round(dp * 65536), 0 = off
  uint32_t dp_thr; float dp_scale; uint64_t dp_seed, dp_offset; int dp_stream;
  // tuning knobs read from the environment ONCE at kgeb_create (KGEB_RING_CTAS, KGEB_RING_D, KGEB_TRAIN_KERNEL=reg):
  // measurement aids, never consulted on the launch path
  int env_ring_ctas, env_ring_d, env_force_reg, env_allpairs_dense;   // KGEB_ALLPAIRS=dense forces the m x m loss kernel (A/B)
  int env_pairs_grouped;   // KGEB_PAIRS_GROUPED=0 / 1: never / always order the undecided pairs by entity before re-scoring; -1 (unset) = by size
  // launch configuration of the ring kernel per (negative mode, dropout flavour), filled on first use: the shared-memory
  // opt-in and the occupancy query are per-function driver calls that need not be repeated every step
  struct { int smem, occ; } ring_cfg[3][2];
  // optional kernel timing (kgeb_set_kernel_timing): events recorded on the caller's stream around selected launches
  int timing;
  cudaEvent_t ev[KGEB_TIMED_KERNELS][KGEB_TIMED_POOL][2];
  int ev_n[KGEB_TIMED_KERNELS];      // launches bracketed since the last kgeb_kernel_ms_sum (the pool is a ring)
  // sort-segmented entity backward (kgeb_set_grad_sortseg): caller-owned workspace, NULL = vector-reduction path
  unsigned char* ss_ws; int64_t ss_bytes;
  const void* ss_fused;   // non-NULL only inside kgeb_train_step_sortseg: the optimizer applied at segment end (sortseg_kernels.cu)
  // NCCL communicator of the sharded paths (kgeb_comm_init, comm.cu); NULL = single GPU
  void* comm; int comm_rank, comm_world;
};

// brackets one launch with events when timing is on (slot = kgeb_timed_kernel)
#define KGEB_TIME_BEGIN(ctx, slot, st) do { if ((ctx)->timing) cudaEventRecord((ctx)->ev[slot][(ctx)->ev_n[slot] % KGEB_TIMED_POOL][0], (st)); } while (0)
#define KGEB_TIME_END(ctx, slot, st) do { if ((ctx)->timing) { cudaEventRecord((ctx)->ev[slot][(ctx)->ev_n[slot] % KGEB_TIMED_POOL][1], (st)); (ctx)->ev_n[slot]++; } } while (0)

extern char g_kgeb_create_err[512];

static inline int kgeb_fail(kgeb_ctx* ctx, int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  char* dst = ctx ? ctx->err : g_kgeb_create_err;
  snprintf(dst, 512, fmt, a, b, c);
  return code;
}

#define KGEB_CUDA_OK(ctx, expr)                                                                          \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return kgeb_fail((ctx), KGEB_ERR_CUDA, "%s (line %lld)", cudaGetErrorString(_e), __LINE__); \
  } while (0)

#define KGEB_LAUNCH_OK(ctx)                                                                              \
  do {                                                                                                   \
    (ctx)->launches++;                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                                 \
    if (_e != cudaSuccess) return kgeb_fail((ctx), KGEB_ERR_CUDA, "launch failed: %s (line %lld)", cudaGetErrorString(_e), __LINE__); \
  } while (0)

#define KGEB_REQUIRE(ctx, cond, code, msg) \
  do { if (!(cond)) return kgeb_fail((ctx), (code), "%s", (msg)); } while (0)

static inline bool kgeb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// HolE through the Fourier domain (hole_fft_kernels.cu); KGEB_ERR_UNSUPPORTED = shape not served, use the direct kernels
int kgeb_internal_hole_fft_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores, cudaStream_t st);
int kgeb_internal_hole_fft_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, const float* dscore, float* gE, float* gR,
                               uint8_t* tE, uint8_t* tR, cudaStream_t st);
int kgeb_internal_hole_fft_train(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* pos, int64_t B, const int32_t* neg, int K, int loss_kind,
                                 float margin, float loss_weight, int corrupt_mode, uint64_t seed, uint64_t offset, float* loss, float* pos_scores,
                                 float* neg_scores, int32_t* neg_out, float* gE, float* gR, uint8_t* tE, uint8_t* tR, cudaStream_t st);
// sort-segmented entity backward (sortseg_kernels.cu), driven by kgeb_train_fwd_bwd when a workspace is armed
int kgeb_internal_sortseg_carve(kgeb_ctx* ctx, int64_t B, int K, float** coef, int32_t** negbuf);
int kgeb_internal_sortseg_entities(kgeb_ctx* ctx, const void* score_args, int nc, const int32_t* pos, const int32_t* neg, int64_t B, int K,
                                   float* gE, cudaStream_t st);
// comm.cu: releases the context's NCCL communicator (kgeb_destroy)
void kgeb_internal_comm_release(kgeb_ctx* ctx);
// HAKE's batch-coupled score (hake_kernels.cu), reached through kgeb_score_fwd / kgeb_score_bwd
int kgeb_internal_hake_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores, cudaStream_t st);
int kgeb_internal_hake_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, const float* dscore, float* gE, float* gR,
                           uint8_t* tE, uint8_t* tR, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Random123).  Counter-based: no state, any (counter,key) is independent.
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * (unsigned long long)c.x;
    const unsigned long long p1 = 0xCD9E8D57ull * (unsigned long long)c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    if (r < 9) { k.x += 0x9E3779B9u; k.y += 0xBB67AE85u; }
  }
  return c;
}

// One corrupted triple: g = i*K + j is the global negative index.  Mirrors oracle.kge_oracle.philox_negatives.
// head_side reports which slot was replaced in the one-sided modes (true = head).
__device__ __forceinline__ void philox_corrupt(int s, int o, uint64_t g, uint64_t seed, uint64_t offset, uint32_t N,
                                               int mode, int& ns, int& no, bool& head_side) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int nh = (int)(((unsigned long long)r.x * N) >> 32);
  const int nt = (int)(((unsigned long long)r.y * N) >> 32);
  head_side = mode == KGEB_CORRUPT_HEAD || (mode == KGEB_CORRUPT_EITHER && (r.z >> 31) != 0);
  ns = s; no = o;
  if (mode == KGEB_CORRUPT_BOTH) { ns = nh; no = nt; }
  else if (head_side) ns = nh;
  else no = nt;
}
__device__ __forceinline__ void philox_corrupt(int s, int o, uint64_t g, uint64_t seed, uint64_t offset, uint32_t N,
                                               int mode, int& ns, int& no) {
  bool hs;
  philox_corrupt(s, o, g, seed, offset, N, mode, ns, no, hs);
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// 128-bit vector reduction to global memory (sm_90+: RED.E.ADD.F32x4); result unused -> no return trip.
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Block-wide sum of two doubles, then "last block done" finalisation:  out = scale0*sum0 + scale1*sum1.
// partials must hold gridDim.x*2 doubles; ticket self-resets so the scratch is reusable by the next launch.
__device__ __forceinline__ void block_finalize_loss(double a, double b, double* partials, unsigned int* ticket,
                                                    double scale0, double scale1, float* out) {
  __shared__ double sh[2][KGEB_WARPS_PER_BLOCK];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if (lane == 0) { sh[0][w] = a; sh[1][w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0, sb = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { sa += sh[0][i]; sb += sh[1][i]; }
    partials[2 * blockIdx.x] = sa;
    partials[2 * blockIdx.x + 1] = sb;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    // warp 0 of the last block: lane-strided partial sums in a fixed order, then a fixed butterfly -> deterministic, and
    // the gridDim.x * 2 partials are fetched 32 at a time instead of one dependent load after the other
    __threadfence();
    double sa = 0, sb = 0;
    for (unsigned int i = lane; i < gridDim.x; i += 32) {
      sa += ((volatile double*)partials)[2 * i];
      sb += ((volatile double*)partials)[2 * i + 1];
    }
    sa = warp_sum_d(sa);
    sb = warp_sum_d(sb);
    if (lane == 0) {
      if (out) out[0] = (float)(scale0 * sa + scale1 * sb);
      *ticket = 0u;
    }
  }
}

// per-element loss pieces (utils.py:288-319).  C1 = fp32(1 + 1e-6) as in TF (EPSILON+1 cast to fp32).
#define KGEB_C1 1.00000095367431640625f

// pointwise: value and derivative of the positive term g(t) and negative term h(f)
__device__ __forceinline__ void pointwise_terms(int kind, float margin, float x, bool positive, float& val, float& der) {
  if (kind == KGEB_POINTWISE_HINGE) {
    const float z = positive ? margin - x : margin + x;
    val = fmaxf(z, 0.f);
    der = z > 0.f ? (positive ? -1.f : 1.f) : 0.f;
  } else if (kind == KGEB_POINTWISE_LOGISTIC) {
    const float e = expf(positive ? -x : x);
    val = logf(KGEB_C1 + e);
    const float sg = e / (KGEB_C1 + e);
    der = positive ? -sg : sg;
  } else if (kind == KGEB_POINTWISE_BCE) {  // Keras binary_crossentropy(1, t) / (0, f) on raw scores: clip, then -log(. + eps)
    const float eps = 1e-7f, hi = 1.f - 1e-7f;
    const float c = fminf(fmaxf(x, eps), hi);
    const bool inside = x >= eps && x <= hi;          // clip_by_value passes the gradient inside the closed interval
    const float arg = positive ? c + eps : (1.f - c) + eps;
    val = -logf(arg);
    der = inside ? (positive ? -1.f / arg : 1.f / arg) : 0.f;
  } else {  // square
    const float z = positive ? margin - x : margin + x;
    val = z * z;
    der = positive ? -2.f * z : 2.f * z;
  }
}
// pairwise: value l(x) and A = dl/dx with x = f - t  (dL/df = A/cnt, dL/dt = -sum A/cnt).
// FAST (the m x m all-pairs walk only, where the transcendental IS the work): ex2 / lg2 / rcp approximations, relative error
// ~1e-6 for |x| < 20 -- inside the stated 1e-5 bar, 3x fewer instructions per pair.
template <bool FAST = false>
__device__ __forceinline__ void pairwise_terms(int kind, float margin, float x, float& val, float& A) {
  if (kind == KGEB_PAIRWISE_HINGE) {
    const float z = margin + x;
    val = fmaxf(z, 0.f);
    A = z > 0.f ? 1.f : 0.f;
  } else if (kind == KGEB_PAIRWISE_LOGISTIC) {
    if (FAST && x < 80.f) {      // beyond ~88 the reference overflows to inf; leave that to the exact branch
      const float e = __expf(x);
      const float s = KGEB_C1 + e;
      val = __logf(s);
      A = __fdividef(e, s);
    } else {
      const float e = expf(x);
      val = logf(KGEB_C1 + e);
      A = e / (KGEB_C1 + e);
    }
  } else {
    val = -x * x;
    A = -2.f * x;
  }
}
