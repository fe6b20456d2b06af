// rank_common.cuh -- the sequential-k fp32 score arithmetic shared by every 1-N ranking kernel (rank_kernels.cu,
// rank_tc_kernels.cu).  Arithmetic contract: every score is accumulated SEQUENTIALLY over k = 0..W-1 in fp32 with
// separately rounded multiply and add (no FMA), so any two code paths give bit-identical scores for the same (q,c) and
// oracle.kge_oracle.seqk_scores reproduces them on the CPU.
#pragma once
#include "common.cuh"

namespace kgeb_rank {

// TransE ops, then the sibling models: ModE head (tail queries reuse the TransE tail ops with q = s*p), pRotatE
enum { OP_DOT = 0, OP_L1_TAIL = 1, OP_L1_HEAD = 2, OP_L2_TAIL = 3, OP_L2_HEAD = 4,
       OP_ML1_HEAD = 5, OP_ML2_HEAD = 6, OP_PSIN_TAIL = 7, OP_PSIN_HEAD = 8, OP_COUNT = 9 };

template <int OP>
constexpr bool op_two_part() { return OP == OP_L1_HEAD || OP == OP_L2_HEAD || OP == OP_ML1_HEAD || OP == OP_ML2_HEAD || OP == OP_PSIN_HEAD; }

// one term of the sequential accumulation; a = qvec[k] (a2 = qvec[W+k] for the two-part head queries), e = E[c][k]
template <int OP>
__device__ __forceinline__ float seq_term(float acc, float a, float a2, float e) {
  if (OP == OP_DOT) return __fadd_rn(acc, __fmul_rn(a, e));
  float r;
  if (OP == OP_L1_TAIL || OP == OP_L2_TAIL || OP == OP_PSIN_TAIL) r = __fsub_rn(a, e);
  else if (OP == OP_ML1_HEAD || OP == OP_ML2_HEAD) r = __fsub_rn(__fmul_rn(e, a), a2);
  else r = __fsub_rn(__fadd_rn(e, a), a2);
  if (OP == OP_PSIN_TAIL || OP == OP_PSIN_HEAD) return __fadd_rn(acc, fabsf(sinf(__fmul_rn(r, 0.5f))));
  if (OP == OP_L1_TAIL || OP == OP_L1_HEAD || OP == OP_ML1_HEAD) return __fadd_rn(acc, fabsf(r));
  return __fadd_rn(acc, __fmul_rn(r, r));
}
// fin.x = gamma, fin.y = 1 if the distance is subtracted from gamma (TransE/pRotatE: gamma > 0; ModE: always), fin.z = param
struct SeqFin { float gamma; int sub; float param; };
template <int OP>
__device__ __forceinline__ float seq_finish(float acc, SeqFin f) {
  if (OP == OP_DOT) return acc;
  if (OP == OP_L2_TAIL || OP == OP_L2_HEAD || OP == OP_ML2_HEAD) acc = __fsqrt_rn(acc);
  if (OP == OP_PSIN_TAIL || OP == OP_PSIN_HEAD) acc = __fmul_rn(acc, f.param);
  return f.sub ? __fsub_rn(f.gamma, acc) : acc;
}

}  // namespace kgeb_rank

// internal (hidden visibility) launchers shared between the translation units
int kgeb_internal_host_op_for(int kind, int side);
kgeb_rank::SeqFin kgeb_internal_seq_fin(const kgeb_ctx* ctx);
// counts[q] -= #{c in filter(q), c != target, c in shard : seq score(q,c) > tscore[q]}
int kgeb_internal_filter_correct(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                                 const float* tscore, const int32_t* triples, int64_t Q, int side,
                                 const int64_t* filt_offsets, const int32_t* filt_ids, int32_t* counts, cudaStream_t st);
// counts[q] += [seq score(q,c) > tscore[q]] for the first min(wl_hdr[0], wl_cap) pairs (q, shard-local c) of the worklist
int kgeb_internal_refine_worklist(kgeb_ctx* ctx, const float* E_shard, int64_t n_rows, const float* qvec, const float* tscore,
                                  int64_t Q, int side, const uint32_t* wl_hdr, uint32_t wl_cap, const uint2* wl, int32_t* counts,
                                  cudaStream_t st, int grouped);
// wl_group_kernels.cu: the worklist ordered by entity on the device (counting sort; lengths never reach the host)
struct KgebWlGroupLayout { size_t bins, slot, out, cub, cub_bytes, total; };
KgebWlGroupLayout kgeb_internal_wl_group_layout(int64_t n_rows, int64_t cap);
int kgeb_internal_wl_group(kgeb_ctx* ctx, const uint32_t* hdr, uint32_t cap, const uint2* wl, int64_t n_rows, unsigned char* scratch,
                           const uint2** ordered, const uint32_t** length, cudaStream_t st);
