// rank_common.cuh -- the sequential-k fp32 score arithmetic shared by every 1-N ranking kernel (rank_kernels.cu,
// rank_tc_kernels.cu).  Arithmetic contract: every score is accumulated SEQUENTIALLY over k = 0..W-1 in fp32 with
// separately rounded multiply and add (no FMA), so any two code paths give bit-identical scores for the same (q,c) and
// oracle.kge_oracle.seqk_scores reproduces them on the CPU.
#pragma once
#include "common.cuh"

namespace kgeb_rank {

enum { OP_DOT = 0, OP_L1_TAIL = 1, OP_L1_HEAD = 2, OP_L2_TAIL = 3, OP_L2_HEAD = 4 };

// one term of the sequential accumulation; a = qvec[k] (a2 = qvec[W+k] for head TransE), e = E[c][k]
template <int OP>
__device__ __forceinline__ float seq_term(float acc, float a, float a2, float e) {
  if (OP == OP_DOT) return __fadd_rn(acc, __fmul_rn(a, e));
  float r;
  if (OP == OP_L1_TAIL || OP == OP_L2_TAIL) r = __fsub_rn(a, e);
  else r = __fsub_rn(__fadd_rn(e, a), a2);
  if (OP == OP_L1_TAIL || OP == OP_L1_HEAD) return __fadd_rn(acc, fabsf(r));
  return __fadd_rn(acc, __fmul_rn(r, r));
}
template <int OP>
__device__ __forceinline__ float seq_finish(float acc, float gamma) {
  if (OP == OP_DOT) return acc;
  if (OP == OP_L2_TAIL || OP == OP_L2_HEAD) acc = __fsqrt_rn(acc);
  return gamma > 0.f ? __fsub_rn(gamma, acc) : acc;
}
template <int OP>
__device__ __forceinline__ float seq_score(const float* __restrict__ q, const float* __restrict__ e, int W, float gamma) {
  float acc = 0.f;
  for (int k = 0; k < W; ++k) acc = seq_term<OP>(acc, q[k], (OP == OP_L1_HEAD || OP == OP_L2_HEAD) ? q[W + k] : 0.f, e[k]);
  return seq_finish<OP>(acc, gamma);
}
__device__ __forceinline__ float seq_score_dyn(int op, const float* q, const float* e, int W, float gamma) {
  switch (op) {
    case OP_DOT: return seq_score<OP_DOT>(q, e, W, gamma);
    case OP_L1_TAIL: return seq_score<OP_L1_TAIL>(q, e, W, gamma);
    case OP_L1_HEAD: return seq_score<OP_L1_HEAD>(q, e, W, gamma);
    case OP_L2_TAIL: return seq_score<OP_L2_TAIL>(q, e, W, gamma);
    default: return seq_score<OP_L2_HEAD>(q, e, W, gamma);
  }
}


}  // namespace kgeb_rank

// internal (hidden visibility) launchers shared between the translation units
int kgeb_internal_host_op_for(int kind, int side);
// counts[q] -= #{c in filter(q), c != target, c in shard : seq score(q,c) > tscore[q]}
int kgeb_internal_filter_correct(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                                 const float* tscore, const int32_t* triples, int64_t Q, int side,
                                 const int64_t* filt_offsets, const int32_t* filt_ids, int32_t* counts, cudaStream_t st);
// counts[q] += [seq score(q,c) > tscore[q]] for the first min(wl_hdr[0], wl_cap) pairs (q, shard-local c) of the worklist
int kgeb_internal_refine_worklist(kgeb_ctx* ctx, const float* E_shard, int64_t n_rows, const float* qvec, const float* tscore,
                                  int64_t Q, int side, const uint32_t* wl_hdr, uint32_t wl_cap, const uint2* wl, int32_t* counts,
                                  cudaStream_t st);
