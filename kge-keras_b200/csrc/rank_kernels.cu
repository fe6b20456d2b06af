// rank_kernels.cu -- 1-N filtered ranking (validate / gen_tail_data / gen_head_data, utils.py:165-264) without ever
// materialising the Q x N score matrix: a register-tiled CUDA-core kernel scores a 64-query x 64-entity tile, compares
// against the query's target score in registers and keeps only integer counts.
//
//   rank(q) = 1 + #{c not in filter(q) u {target(q)} : score(q,c) > score(q,target)}          
//
// Arithmetic contract (what makes the integer counts reproducible and tie-consistent): every score -- candidate,
// target, filtered candidate -- is accumulated SEQUENTIALLY over k = 0..W-1 in fp32 with separately rounded multiply
// and add (__fmul_rn/__fadd_rn, no FMA), so the three code paths below give bit-identical scores for the same (q,c)
// and oracle.kge_oracle.seqk_scores reproduces them on the CPU.  The filter is applied as a correction pass:
// counts -= #{c in filter(q) : score(q,c) > target}.
#include <cuda.h>   // CUtensorMap (types only: the encoder is reached through cudaGetDriverEntryPoint, no libcuda at link time)
#include <stdlib.h>
#include "rank_common.cuh"

using namespace kgeb_rank;

namespace {

// ---------------------------------------------------------------------------------------------------
// query vectors
// ---------------------------------------------------------------------------------------------------
// one WARP per query: the triple and the row pointers are read once, the lanes stride over the output elements (coalesced)
__global__ void __launch_bounds__(256) k_query_prep(int kind, int side, int gamma_sub, const float* __restrict__ Es, int64_t row_offset,
                                                    int64_t n_rows, int ldE, const float* __restrict__ R, int ldR, int W,
                                                    const int32_t* __restrict__ triples, int64_t Q, int qw, float* __restrict__ qvec) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = warp0; q < Q; q += nwarps) {
    const int64_t src = side == KGEB_SIDE_TAIL ? triples[3 * q] : triples[3 * q + 2];
    if (src < row_offset || src >= row_offset + n_rows) continue;  // another shard owns this query's source row
    const float* e = Es + (src - row_offset) * (int64_t)ldE;   // s (tail queries) or o (head queries)
    const float* p = R + (int64_t)triples[3 * q + 1] * ldR;
    float* out = qvec + q * (int64_t)qw;
    if (kind == KGEB_DISTMULT) {
      for (int k = lane; k < W; k += 32) out[k] = side == KGEB_SIDE_TAIL ? __fmul_rn(e[k], p[k]) : __fmul_rn(p[k], e[k]);
    } else if (kind == KGEB_COMPLEX) {
      const int d = W / 2;
      for (int j = lane; j < d; j += 32) {
        const float er = e[j], ei = e[d + j], pr = p[j], pi = p[d + j];
        if (side == KGEB_SIDE_TAIL) {   // e = s:  [sr pr - si pi | si pr + sr pi]
          out[j] = __fsub_rn(__fmul_rn(er, pr), __fmul_rn(ei, pi));
          out[d + j] = __fadd_rn(__fmul_rn(ei, pr), __fmul_rn(er, pi));
        } else {                        // e = o:  [pr or + pi oi | pr oi - pi or]
          out[j] = __fadd_rn(__fmul_rn(pr, er), __fmul_rn(pi, ei));
          out[d + j] = __fsub_rn(__fmul_rn(pr, ei), __fmul_rn(pi, er));
        }
      }
    } else if (kind == KGEB_HOLE) {
      for (int k = lane; k < W; k += 32) {
        float acc = 0.f;
        if (side == KGEB_SIDE_TAIL) {  // q_j = sum_i s_i p_{(j-i) mod d}
          int jj = k;
          for (int i = 0; i < W; ++i) {
            acc = __fadd_rn(acc, __fmul_rn(e[i], p[jj]));
            if (--jj < 0) jj = W - 1;
          }
        } else {                        // q_i = sum_k p_k o_{(i+k) mod d}
          int jj = k;
          for (int i = 0; i < W; ++i) {
            acc = __fadd_rn(acc, __fmul_rn(p[i], e[jj]));
            if (++jj == W) jj = 0;
          }
        }
        out[k] = acc;
      }
    } else if (kind == KGEB_ROTATE) {
      // models.py:439-459: the score is linear in the candidate row.  tail: gamma - A = const + sum_j o_j -> q = 1;
      // head: gamma - A = const - e . [(c + s) | (c - s)], c/s = cos/sin of the relation phases.  The per-query constant
      // cancels inside one query's ranking (target and candidates share it); gamma <= 0 ranks by A itself: sign flips.
      const int d = W / 2;
      for (int k = lane; k < W; k += 32) {
        float v;
        if (side == KGEB_SIDE_TAIL) v = 1.f;
        else {
          float sn, c;
          sincosf(p[k < d ? k : k - d], &sn, &c);
          v = k < d ? -__fadd_rn(c, sn) : -__fsub_rn(c, sn);
        }
        out[k] = gamma_sub ? v : -v;
      }
    } else if (kind == KGEB_MODE_L1 || kind == KGEB_MODE_L2) {  // tail r = s*p - e; head r = e*p - o keeps [p | o]
      for (int k = lane; k < qw; k += 32) out[k] = side == KGEB_SIDE_TAIL ? __fmul_rn(e[k], p[k]) : (k < W ? p[k] : e[k - W]);
    } else {  // TransE, pRotatE: tail a = s + p; head keeps [p | o] so (e + p) - o follows the reference's order
      for (int k = lane; k < qw; k += 32) out[k] = side == KGEB_SIDE_TAIL ? __fadd_rn(e[k], p[k]) : (k < W ? p[k] : e[k - W]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// exact scores of explicit (query, candidate) pairs: the target scores, the CSR filter correction and the re-scoring of
// the pairs the tensor-core kernel could not decide all run through this one kernel.  A warp owns 32 pairs; the rows
// travel global -> shared memory as coalesced 128-byte segments (32 floats of every pair's query and entity row per
// pass), then lane p accumulates ITS pair sequentially over k from shared memory (stride 33: conflict free).  The
// arithmetic is exactly seq_score's (same order, same roundings), only the loads are cooperative.
// ---------------------------------------------------------------------------------------------------
enum { PAIRS_TARGET = 0, PAIRS_WORKLIST = 1, PAIRS_FILTER = 2, PAIRS_GROUPED = 3 };   // GROUPED: a worklist ordered by entity (wl_group_kernels.cu)
#ifndef KGEB_PAIR_WARPS
#define KGEB_PAIR_WARPS 2
#endif
constexpr int PAIR_WARPS = KGEB_PAIR_WARPS;

struct PairArgs {
  const float* Es; int64_t row_offset, n_rows; int ldE, W;
  const float* qvec; int qw;
  float* tscore;                 // written (TARGET) or read
  const int32_t* triples; int64_t Q; int side; SeqFin fin;
  const int64_t* offs; const int32_t* ids;          // FILTER
  const uint32_t* wl_hdr; uint32_t wl_cap; const uint2* wl;  // WORKLIST
  int32_t* counts;
};

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int PAIR_LDP = 36;    // row stride in floats: 16-byte aligned rows, conflict-free LDS.128 by row (4*lane groups)
#ifndef KGEB_PAIR_NBUF
#define KGEB_PAIR_NBUF 3
#endif
constexpr int PAIR_NBUF = KGEB_PAIR_NBUF;    // k-chunks of 32 floats in flight per warp (two ahead of the one being accumulated)

template <int OP, int SRC>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_seqk_pairs(PairArgs a) {
  constexpr bool HEAD = op_two_part<OP>();
  constexpr bool GROUPED = SRC == PAIRS_GROUPED;   // consecutive pairs share entity rows: a warp stages each DISTINCT row once
  constexpr int LDP = PAIR_LDP, NB = PAIR_NBUF;
  constexpr int ROWS = HEAD ? 3 : 2;            // q, e (and the second query part) per pair
  extern __shared__ __align__(16) float pair_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // per warp: [NB][ROWS][32 pairs][LDP]
  float* wbase = pair_smem + (size_t)w * NB * ROWS * 32 * LDP;
  auto buf = [&](int b, int r) { return wbase + ((size_t)(b * ROWS + r) * 32) * LDP; };
  const int sub = lane >> 3, l8 = lane & 7;  // 8 lanes x 16 bytes cover one 128-byte row segment; 4 pairs per copy instruction
  int64_t n;
  if (SRC == PAIRS_TARGET) n = a.Q;
  else if (SRC == PAIRS_WORKLIST || GROUPED) n = a.wl_hdr[0] < a.wl_cap ? a.wl_hdr[0] : a.wl_cap;
  else n = a.offs[a.Q];
  const int64_t warp0 = (int64_t)blockIdx.x * PAIR_WARPS + w, nwarps = (int64_t)gridDim.x * PAIR_WARPS;
  const int nkc = (a.W + 31) / 32;
  for (int64_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const int64_t x = base + lane;
    bool valid = x < n;
    int64_t q = 0, c = 0;
    if (valid) {
      if (SRC == PAIRS_TARGET) {
        q = x;
        c = (int64_t)(a.side == KGEB_SIDE_TAIL ? a.triples[3 * q + 2] : a.triples[3 * q]) - a.row_offset;
      } else if (SRC == PAIRS_WORKLIST || GROUPED) {
        const uint2 pr = a.wl[x];
        q = pr.x; c = pr.y;
      } else {
        int64_t lo = 0, hi = a.Q;  // largest q with offs[q] <= x
        while (hi - lo > 1) {
          const int64_t mid = (lo + hi) >> 1;
          if (a.offs[mid] <= x) lo = mid; else hi = mid;
        }
        q = lo;
        const int64_t id = a.ids[x];
        const int64_t tgt = a.side == KGEB_SIDE_TAIL ? a.triples[3 * q + 2] : a.triples[3 * q];
        c = id - a.row_offset;
        valid = id != tgt;
      }
      valid = valid && c >= 0 && c < a.n_rows;   // another shard owns this candidate
      if (!valid) { q = 0; c = 0; }
    }
    // element offsets of the two rows of MY pair; the copy lanes fetch them from the owning lane
    const int64_t qoff = q * a.qw, eoff = c * (int64_t)a.ldE;
    // GROUPED: lane p opens a new entity row when its candidate differs from lane p-1's; row p is staged into slot
    // (number of rows opened at or before p) - 1 and every lane reads its slot (same-slot lanes read one address: a broadcast)
    uint32_t firstmask = 0xffffffffu;
    if (GROUPED) {
      const int64_t cprev = __shfl_up_sync(0xffffffffu, c, 1);
      firstmask = __ballot_sync(0xffffffffu, lane == 0 || c != cprev);
    }
    const int eslot = GROUPED ? __popc(firstmask & ((2u << lane) - 1u)) - 1 : lane;
    // issue the asynchronous copies of k-chunk kc (32 floats of every pair's rows) into buffer kc % NB
    auto issue = [&](int kc) {
      const int k = kc * 32 + 4 * l8;
      const bool kin = k < a.W;                 // W % 4 == 0: a 16-byte piece is entirely inside or outside the row
      const int b = kc % NB;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = 4 * i + sub;
        const int64_t qo = __shfl_sync(0xffffffffu, qoff, p), eo = __shfl_sync(0xffffffffu, eoff, p);
        cp_async16(buf(b, 0) + p * LDP + 4 * l8, a.qvec + qo + (kin ? k : 0), kin);
        if (!GROUPED) cp_async16(buf(b, 1) + p * LDP + 4 * l8, a.Es + eo + (kin ? k : 0), kin);
        else if ((firstmask >> p) & 1u) cp_async16(buf(b, 1) + (__popc(firstmask & ((2u << p) - 1u)) - 1) * LDP + 4 * l8, a.Es + eo + (kin ? k : 0), kin);
        if (HEAD) cp_async16(buf(b, 2) + p * LDP + 4 * l8, a.qvec + qo + a.W + (kin ? k : 0), kin);
      }
      cp_async_commit();
    };
    __syncwarp();                                // every lane is done with the previous batch's buffers
#pragma unroll
    for (int pre = 0; pre < NB - 1; ++pre) {     // NB - 1 chunks ahead; empty groups keep the group count in step
      if (pre < nkc) issue(pre); else cp_async_commit();
    }
    float acc = 0.f;
    for (int kc = 0; kc < nkc; ++kc) {
      cp_async_wait<NB - 2>();                   // chunk kc has landed (only the NB - 2 newest groups may still be in flight)
      __syncwarp();
      const int b = kc % NB;
      const float* rq = buf(b, 0) + lane * LDP;
      const float* re = buf(b, 1) + eslot * LDP;
      const float* rq2 = HEAD ? buf(b, 2) + lane * LDP : nullptr;
      const int klim = a.W - kc * 32 < 32 ? a.W - kc * 32 : 32;   // multiple of 4
      for (int kk = 0; kk < klim; kk += 4) {
        const float4 x4 = *reinterpret_cast<const float4*>(rq + kk);
        const float4 e4 = *reinterpret_cast<const float4*>(re + kk);
        float4 y4 = f4zero();
        if (HEAD) y4 = *reinterpret_cast<const float4*>(rq2 + kk);
        acc = seq_term<OP>(acc, x4.x, y4.x, e4.x);
        acc = seq_term<OP>(acc, x4.y, y4.y, e4.y);
        acc = seq_term<OP>(acc, x4.z, y4.z, e4.z);
        acc = seq_term<OP>(acc, x4.w, y4.w, e4.w);
      }
      // buffer (kc + NB - 1) % NB == (kc - 1) % NB was read in the previous iteration; every lane has passed this iteration's
      // __syncwarp since, so it may be refilled
      if (kc + NB - 1 < nkc) issue(kc + NB - 1); else cp_async_commit();
    }
    cp_async_wait<0>();
    if (!valid) continue;
    const float sc = seq_finish<OP>(acc, a.fin);
    if (SRC == PAIRS_TARGET) a.tscore[q] = sc;
    else if (SRC == PAIRS_WORKLIST || GROUPED) { if (sc > a.tscore[q]) atomicAdd(a.counts + q, 1); }
    else { if (sc > a.tscore[q]) atomicSub(a.counts + q, 1); }
  }
}

constexpr size_t pair_smem_bytes(bool head) { return (size_t)PAIR_WARPS * PAIR_NBUF * (head ? 3 : 2) * 32 * PAIR_LDP * sizeof(float); }

// (more than 48 KB of dynamic shared memory needs the per-function opt-in; setting it again is harmless)
#define KGEB_PAIRS_LAUNCH(OPV)                                                                                             \
  {                                                                                                                        \
    const size_t psm = pair_smem_bytes(op_two_part<OPV>());                                                                \
    cudaFuncSetAttribute(k_seqk_pairs<OPV, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);                   \
    k_seqk_pairs<OPV, SRC><<<blocks, PAIR_WARPS * 32, psm, st>>>(a);                                                       \
  }
template <int SRC>
void launch_pairs(int op, int blocks, cudaStream_t st, const PairArgs& a) {
  switch (op) {
    case OP_ML1_HEAD: KGEB_PAIRS_LAUNCH(OP_ML1_HEAD); break;
    case OP_ML2_HEAD: KGEB_PAIRS_LAUNCH(OP_ML2_HEAD); break;
    case OP_PSIN_TAIL: KGEB_PAIRS_LAUNCH(OP_PSIN_TAIL); break;
    case OP_PSIN_HEAD: KGEB_PAIRS_LAUNCH(OP_PSIN_HEAD); break;
    case OP_DOT: KGEB_PAIRS_LAUNCH(OP_DOT); break;
    case OP_L1_TAIL: KGEB_PAIRS_LAUNCH(OP_L1_TAIL); break;
    case OP_L1_HEAD: KGEB_PAIRS_LAUNCH(OP_L1_HEAD); break;
    case OP_L2_TAIL: KGEB_PAIRS_LAUNCH(OP_L2_TAIL); break;
    default: KGEB_PAIRS_LAUNCH(OP_L2_HEAD); break;
  }
}

// ---------------------------------------------------------------------------------------------------
// the tile kernel: CTA = 64 queries x (a strided set of 64-entity tiles); thread = 4 x 4 micro-tile of INTERLEAVED rows
// (queries ty + 16 i, entities tx + 16 j).  Tiles sit in shared memory as they sit in global memory -- row major, k
// contiguous, rows padded to 36 floats -- so a tile arrives by plain 128-bit stores (no transposing scalar stores, no
// bank conflicts) and a thread fetches 4 consecutive k of a row with one LDS.128: the 8 threads of a quarter warp read
// 8 consecutive entity rows = 32 distinct banks, and one query row (broadcast).  Shared memory is double buffered: the
// next k-chunk is fetched into registers before the current one is consumed and stored into the other buffer behind it,
// one __syncthreads per chunk.  Per (query, entity) pair the terms are still accumulated in k order.
// ---------------------------------------------------------------------------------------------------
constexpr int BQ = 64, BN = 64, BK = 32, BKP = 36;

template <int OP>
__global__ void __launch_bounds__(256) k_rank_tile(const float* __restrict__ Es, int64_t n_rows, int ldE, int W,
                                                   const float* __restrict__ qvec, int qw, const float* __restrict__ tscore,
                                                   int64_t Q, SeqFin fin, int32_t* counts) {
  constexpr bool HEAD = op_two_part<OP>();
  extern __shared__ __align__(16) float tile_smem[];
  typedef float (*TileBuf)[BQ][BKP];   // [buffer][row][k]
  TileBuf As = reinterpret_cast<TileBuf>(tile_smem);
  TileBuf Bs = reinterpret_cast<TileBuf>(tile_smem + 2 * BQ * BKP);
  TileBuf As2 = reinterpret_cast<TileBuf>(tile_smem + 4 * BQ * BKP);   // two-part (head) queries only
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t q0 = (int64_t)blockIdx.x * BQ;
  const int64_t ntiles = (n_rows + BN - 1) / BN;
  // loader mapping: a 32-float k-chunk of a row is 8 float4; 32 rows per pass, 2 passes
  const int lr = tid >> 3, lk4 = tid & 7;
  const int nkc = (W + BK - 1) / BK;

  float tq[4];
  int cnt[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t q = q0 + ty + 16 * i;
    tq[i] = q < Q ? __ldg(tscore + q) : 0.f;
  }

  for (int64_t tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
    const int64_t n0 = tile * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra[2], ra2[2], rb[2];
    auto gload = [&](int kc) {
      const int k = kc * BK + lk4 * 4;
      const bool kin = k < W;   // W % 4 == 0: a float4 is entirely inside or outside the row
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t q = q0 + lr + 32 * h;
        const int64_t n = n0 + lr + 32 * h;
        ra[h] = (kin && q < Q) ? ldg4(reinterpret_cast<const float4*>(qvec + q * (int64_t)qw + k)) : f4zero();
        if (HEAD) ra2[h] = (kin && q < Q) ? ldg4(reinterpret_cast<const float4*>(qvec + q * (int64_t)qw + W + k)) : f4zero();
        rb[h] = (kin && n < n_rows) ? ldg4(reinterpret_cast<const float4*>(Es + n * (int64_t)ldE + k)) : f4zero();
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lr + 32 * h;
        *reinterpret_cast<float4*>(&As[buf][r][lk4 * 4]) = ra[h];
        if (HEAD) *reinterpret_cast<float4*>(&As2[buf][r][lk4 * 4]) = ra2[h];
        *reinterpret_cast<float4*>(&Bs[buf][r][lk4 * 4]) = rb[h];
      }
    };

    __syncthreads();   // the previous tile's last chunk is fully consumed before its buffer is overwritten
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kc = 0; kc < nkc; ++kc) {
      const int buf = kc & 1;
      if (kc + 1 < nkc) gload(kc + 1);  // in flight while this chunk is computed
      const int klim = (W - kc * BK) < BK ? (W - kc * BK) : BK;   // multiple of 4
#pragma unroll 2
      for (int k = 0; k < klim; k += 4) {
        float4 a[4], a2[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = *reinterpret_cast<const float4*>(&As[buf][ty + 16 * i][k]);
          a2[i] = HEAD ? *reinterpret_cast<const float4*>(&As2[buf][ty + 16 * i][k]) : f4zero();
          b[i] = *reinterpret_cast<const float4*>(&Bs[buf][tx + 16 * i][k]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float c = acc[i][j];
            c = seq_term<OP>(c, a[i].x, a2[i].x, b[j].x);
            c = seq_term<OP>(c, a[i].y, a2[i].y, b[j].y);
            c = seq_term<OP>(c, a[i].z, a2[i].z, b[j].z);
            c = seq_term<OP>(c, a[i].w, a2[i].w, b[j].w);
            acc[i][j] = c;
          }
      }
      if (kc + 1 < nkc) sstore(buf ^ 1);   // the other buffer was last read in iteration kc - 1, before the barrier below
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx + 16 * j;
        const float sc = seq_finish<OP>(acc[i][j], fin);
        cnt[i] += (n < n_rows && sc > tq[i]) ? 1 : 0;
      }
  }
  // the 16 tx lanes of one ty live in one half-warp: butterfly inside 16 lanes, then one integer atomic per query
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = cnt[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    const int64_t q = q0 + ty + 16 * i;
    if (tx == 0 && q < Q && c != 0) atomicAdd(counts + q, c);
  }
}

// ---------------------------------------------------------------------------------------------------
// the packed tile kernel (round 2): the same 64 x 64 tile and the same per-pair sequential-k fp32 arithmetic, issued as the
// sm_100 two-wide fp32 add (add.rn.f32x2 -> SASS FADD2).  The two halves of a packed accumulator are two
// DIFFERENT (query, entity) pairs -- one query against the entity pair (n, n + 32) at the same k -- so every pair still sees its
// terms in k order with separately rounded operations: bit-identical scores, half the issue slots.  ptxas folds what the scalar
// form got from operand modifiers into the packed instruction as well: the query value enters as a scalar broadcast operand
// (`R.F32`), the negation of e and the |x| of the L1 term as `-R.F32x2` / `|R|.F32x2` (checked in SASS; no MOV, no LOP3).
// Entity tiles sit in shared memory pair-interleaved -- row p holds (e_p[k], e_{p+32}[k]) for k = 0..31 -- so one LDS.128
// delivers two k of a pair already in aligned register pairs; the loader thread that fetches rows p and p + 32 interleaves
// them in registers and still stores 128 bits at a time (chunk order swapped in the upper half of a quarter warp: no bank
// conflicts).  Work is one flat list of (query tile, entity tile) items cut into equal contiguous ranges over a grid of
// exactly SMs x resident CTAs, so no wave is partially filled whatever Q and N are (the 2-D grid of the scalar kernel ran
// 637 CTAs on 296 slots on the WN18RR shape: a third wave 15 % full); counts are flushed when a range crosses into the next
// query tile.
// ---------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ u64 neg2(u64 v) { return pk2(-lo2(v), -hi2(v)); }
__device__ __forceinline__ u64 abs2(u64 v) { return pk2(fabsf(lo2(v)), fabsf(hi2(v))); }

// seq_term<OP> on two pairs at once: a, a2 belong to the query, e2 = (E[c][k], E[c + 32][k]).  x - y is issued as x + (-y): the same
// IEEE result.  Only the L1 ops are packed: ptxas contracts a packed multiply (mul.rn.f32x2, or fma.rn.f32x2 with a -0 addend) and
// the packed add behind it into ONE FFMA2 whatever the .rn qualifiers and --fmad say (seen in SASS), which rounds once where the
// contract rounds twice -- so the dot-product, L2 and ModE ops stay on the scalar kernel, whose __fmul_rn / __fadd_rn are honoured.
template <int OP>
__device__ __forceinline__ u64 seq_term2(u64 acc, float a, float a2, u64 e2) {
  static_assert(OP == OP_L1_TAIL || OP == OP_L1_HEAD, "packed ops: L1 only");
  const u64 r = OP == OP_L1_TAIL ? add2(bc2(a), neg2(e2)) : add2(add2(e2, bc2(a)), bc2(-a2));
  return add2(acc, abs2(r));
}
template <int OP>
constexpr bool op_packed() { return OP == OP_L1_TAIL || OP == OP_L1_HEAD; }
// the same term with the roles swapped (k_rank_tile3): e is the scalar, a / a2 the packed values of a query pair.  Tail: e - a, and
// |e - a| = |a - e| bit for bit; head: (e + a) - a2 as written.
template <int OP>
__device__ __forceinline__ u64 seq_term3(u64 acc, float e, u64 a, u64 a2) {
  static_assert(OP == OP_L1_TAIL || OP == OP_L1_HEAD, "packed ops: L1 only");
  const u64 r = OP == OP_L1_TAIL ? add2(bc2(e), neg2(a)) : add2(add2(bc2(e), a), neg2(a2));
  return add2(acc, abs2(r));
}

constexpr int BPP = 68;   // floats per pair row: 32 k x 2 + 4 (row stride = 4 banks mod 32: 8 rows of a quarter warp cover all banks)

template <int OP>
__global__ void __launch_bounds__(256, 2) k_rank_tile2(const float* __restrict__ Es, int64_t n_rows, int ldE, int W,
                                                       const float* __restrict__ qvec, int qw, const float* __restrict__ tscore,
                                                       int64_t Q, SeqFin fin, int32_t* counts) {
  constexpr bool HEAD = op_two_part<OP>();
  extern __shared__ __align__(16) float tile_smem[];
  typedef float (*TileBuf)[BQ][BKP];    // [buffer][query row][k]
  typedef float (*PairBuf)[BN / 2][BPP];   // [buffer][entity pair][k][2]
  TileBuf As = reinterpret_cast<TileBuf>(tile_smem);
  PairBuf Bp = reinterpret_cast<PairBuf>(tile_smem + 2 * BQ * BKP);
  TileBuf As2 = reinterpret_cast<TileBuf>(tile_smem + 2 * BQ * BKP + 2 * (BN / 2) * BPP);   // two-part (head) queries only
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t ntiles = (n_rows + BN - 1) / BN;
  const int64_t items = ((Q + BQ - 1) / BQ) * ntiles;
  const int64_t t0 = items * blockIdx.x / gridDim.x, t1 = items * (blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  const int lr = tid >> 3, lk4 = tid & 7;
  const int swp = (lk4 >> 2) & 1;
  const int nkc = (W + BK - 1) / BK;

  float tq[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt[4] = {0, 0, 0, 0};
  int64_t cur_qt = -1, q0 = 0;
  auto flush = [&]() {
    // the 16 tx lanes of one ty live in one half-warp: butterfly inside 16 lanes, then one integer atomic per query
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = cnt[i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      const int64_t q = q0 + ty + 16 * i;
      if (tx == 0 && q < Q && c != 0) atomicAdd(counts + q, c);
      cnt[i] = 0;
    }
  };

  // the k-chunks of all items of this CTA form ONE pipeline: while chunk (t, kc) is consumed, the chunk after it -- the first one
  // of item t + 1 after the last one of item t -- is in flight from global memory into registers, and is stored into the other
  // buffer behind the computation; one barrier per chunk, no load latency exposed between items
  float4 ra[2], ra2[2], rb[2];
  auto gload = [&](int64_t gq0, int64_t gn0, int kc) {   // first query / first entity row of the item the chunk belongs to
    const int k = kc * BK + lk4 * 4;
    const bool kin = k < W;   // W % 4 == 0: a float4 is entirely inside or outside the row
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t q = gq0 + lr + 32 * h;
      const int64_t n = gn0 + lr + 32 * h;
      ra[h] = (kin && q < Q) ? ldg4(reinterpret_cast<const float4*>(qvec + q * (int64_t)qw + k)) : f4zero();
      if (HEAD) ra2[h] = (kin && q < Q) ? ldg4(reinterpret_cast<const float4*>(qvec + q * (int64_t)qw + W + k)) : f4zero();
      rb[h] = (kin && n < n_rows) ? ldg4(reinterpret_cast<const float4*>(Es + n * (int64_t)ldE + k)) : f4zero();
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 32 * h;
      *reinterpret_cast<float4*>(&As[buf][r][lk4 * 4]) = ra[h];
      if (HEAD) *reinterpret_cast<float4*>(&As2[buf][r][lk4 * 4]) = ra2[h];
    }
    // rows lr and lr + 32 interleaved: chunk 2 lk4 = k, k+1; chunk 2 lk4 + 1 = k+2, k+3
    const float4 c0 = make_float4(rb[0].x, rb[1].x, rb[0].y, rb[1].y), c1 = make_float4(rb[0].z, rb[1].z, rb[0].w, rb[1].w);
    float* row = &Bp[buf][lr][lk4 * 8];
    *reinterpret_cast<float4*>(row + 4 * swp) = swp ? c1 : c0;
    *reinterpret_cast<float4*>(row + 4 * (1 - swp)) = swp ? c0 : c1;
  };

  int64_t qt = t0 / ntiles, nt = t0 - qt * ntiles;   // the only division: items advance entity tile first, then query tile
  gload(qt * BQ, nt * BN, 0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t n0 = nt * BN;
    int64_t qt_next = qt, nt_next = nt + 1;
    if (nt_next == ntiles) { nt_next = 0; ++qt_next; }
    if (qt != cur_qt) {   // CTA-uniform
      if (cur_qt >= 0) flush();
      cur_qt = qt; q0 = qt * BQ;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t q = q0 + ty + 16 * i;
        tq[i] = q < Q ? __ldg(tscore + q) : 0.f;
      }
    }
    u64 acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }

    for (int kc = 0; kc < nkc; ++kc) {
      const bool more = kc + 1 < nkc || t + 1 < t1;
      if (more) { if (kc + 1 < nkc) gload(q0, n0, kc + 1); else gload(qt_next * BQ, nt_next * BN, 0); }  // in flight while this chunk is computed
      const int klim = (W - kc * BK) < BK ? (W - kc * BK) : BK;   // multiple of 4
#pragma unroll 2
      for (int k = 0; k < klim; k += 4) {
        float4 a[4], a2[4];
        ulonglong2 b[2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = *reinterpret_cast<const float4*>(&As[buf][ty + 16 * i][k]);
          a2[i] = HEAD ? *reinterpret_cast<const float4*>(&As2[buf][ty + 16 * i][k]) : f4zero();
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          b[j][0] = *reinterpret_cast<const ulonglong2*>(&Bp[buf][tx + 16 * j][2 * k]);
          b[j][1] = *reinterpret_cast<const ulonglong2*>(&Bp[buf][tx + 16 * j][2 * k + 4]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            u64 c = acc[i][j];
            c = seq_term2<OP>(c, a[i].x, a2[i].x, b[j][0].x);
            c = seq_term2<OP>(c, a[i].y, a2[i].y, b[j][0].y);
            c = seq_term2<OP>(c, a[i].z, a2[i].z, b[j][1].x);
            c = seq_term2<OP>(c, a[i].w, a2[i].w, b[j][1].y);
            acc[i][j] = c;
          }
      }
      if (more) sstore(buf ^ 1);   // the other buffer was last read one chunk ago, before the previous barrier
      __syncthreads();
      buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t n = n0 + tx + 16 * j;   // the pair's rows: n and n + 32
        const float s_lo = seq_finish<OP>(lo2(acc[i][j]), fin), s_hi = seq_finish<OP>(hi2(acc[i][j]), fin);
        cnt[i] += (n < n_rows && s_lo > tq[i]) ? 1 : 0;
        cnt[i] += (n + 32 < n_rows && s_hi > tq[i]) ? 1 : 0;
      }
    qt = qt_next; nt = nt_next;
  }
  flush();
}

// ---------------------------------------------------------------------------------------------------
// the decoupled packed tile kernel: same arithmetic as k_rank_tile2, no load instruction and no CTA-wide barrier in the compute
// warps' steady state.  profiles/r02_rank_tile2_summary.md: the k-loop of k_rank_tile2 alone reaches ~90 % of the fp32 add pipe, staging
// the next chunk through registers costs 14 % and the per-chunk barrier 7 %.  Here
//   * ONE producer thread streams the candidate tiles by TMA: every k-chunk of a 64-row entity tile is one
//     cp.async.bulk.tensor.2d (SASS UTMALDG) of a 32-float x 64-row box of the table's tensor map into a ring of NST 8 KB stages that
//     complete on mbarriers (full: transaction bytes; empty: one arrival per consumer warp).  The box lands as dense 128-byte rows with
//     the 128-byte swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)), so the 8 rows a quarter warp reads with one LDS.128
//     fall in 8 different bank groups; rows past the shard and k past the row width arrive as zeros (never counted / never read).
//     (The first version issued one 128-byte cp.async.bulk per row: UBLKCP is a uniform instruction, a warp issues its lanes' copies one
//     at a time, and a single producer warp could not keep up -- 3.36 ms; four producer warps: 1.81 ms.)
//   * the 64-query tile is RESIDENT in shared memory for as long as the CTA's items stay in that query tile, stored pair-interleaved
//     (row p = queries p and p + 32, (q_p[k], q_{p+32}[k]) adjacent), so the packed operand of FADD2 is now the QUERY pair and the candidate
//     value is the scalar broadcast operand: d = e - q (|e - q| = |q - e| exactly), or (e + a) - a2 for the head side;
//   * two consumer groups of 8 warps share the resident query tile and take alternate entity tiles; their chunks interleave in the ring.
// One CTA of 17 warps per SM; used when the resident tile fits (TransE-L1 widths up to ~400 tail / ~200 head), else k_rank_tile2.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rk_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rk_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rk_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rk_smem(bar)) : "memory");
}
__device__ __forceinline__ void rk_tma_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(rk_smem(dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(rk_smem(bar))
               : "memory");
}
__device__ __forceinline__ void rk_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "RK_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra RK_DONE;\n\t"
      "bra RK_WAIT;\n\t"
      "RK_DONE:\n\t"
      "}" ::"r"(rk_smem(bar)),
      "r"(parity)
      : "memory");
}

constexpr int T3_NST = 8;                       // ring stages (a stage = one k-chunk of one 64-row entity tile: a 32 x 64 fp32 box)
constexpr int T3_STAGE_BYTES = BN * BK * 4;     // 8192: dense 128-byte rows, 128-byte swizzle (stage bases are 1024-byte aligned)
constexpr int T3_THREADS = 17 * 32;             // 2 consumer groups x 8 warps + the producer warp (one thread of it issues the TMA copies)
__host__ __device__ inline int t3_qstride(int W) { return 2 * W + ((2 * W) % 32 == 0 ? 8 : 0); }   // floats per pair row; adjacent rows in distinct banks
inline size_t t3_smem_bytes(int W, bool head) {
  return (size_t)T3_NST * T3_STAGE_BYTES + (size_t)(head ? 2 : 1) * (BQ / 2) * t3_qstride(W) * sizeof(float) + 2 * T3_NST * sizeof(uint64_t) + 1024;
}

template <int OP>
__global__ void __launch_bounds__(T3_THREADS, 1) k_rank_tile3(const __grid_constant__ CUtensorMap tmapE, int64_t n_rows, int W,
                                                              const float* __restrict__ qvec, int qw, const float* __restrict__ tscore,
                                                              int64_t Q, SeqFin fin, int32_t* counts) {
  constexpr bool HEAD = op_two_part<OP>();
  extern __shared__ __align__(1024) unsigned char tile3_smem[];
  unsigned char* ring = tile3_smem + ((1024u - (rk_smem(tile3_smem) & 1023u)) & 1023u);   // [NST][64 rows][128 B], 1024-byte aligned (swizzle atom)
  const int QS = t3_qstride(W);
  float* q2 = reinterpret_cast<float*>(ring + T3_NST * T3_STAGE_BYTES);          // [part][32 pairs][QS]
  uint64_t* full = reinterpret_cast<uint64_t*>(q2 + (HEAD ? 2 : 1) * (BQ / 2) * QS);
  uint64_t* empty = full + T3_NST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (n_rows + BN - 1) / BN;
  const int64_t items = ((Q + BQ - 1) / BQ) * ntiles;
  const int64_t t0 = items * blockIdx.x / gridDim.x, t1 = items * (blockIdx.x + 1) / gridDim.x;
  const int nkc = (W + BK - 1) / BK;
  if (threadIdx.x == 0) {
    for (int i = 0; i < T3_NST; ++i) { rk_mbar_init(full + i, 1); rk_mbar_init(empty + i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (t0 >= t1) return;
  // the CTA's items as runs inside one query tile: run r covers items [ta, tb) of query tile qt; inside a run the items pair up
  // (2 P, 2 P + 1) -> (group 0, group 1) and the chunks enter the ring in the order (P, kc, group)
  const int64_t qt_first = t0 / ntiles;

  if (warp == 16) {
    // ---------------- producer: one thread, one TMA box per stage ----------------
    if (lane != 0) return;
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapE) : "memory");
    uint32_t pos = 0;
    int64_t ta = t0;
    for (int64_t qt = qt_first; ta < t1; ++qt) {
      const int64_t tb = (qt + 1) * ntiles < t1 ? (qt + 1) * ntiles : t1;
      const int64_t nt0 = ta - qt * ntiles;
      const int n = (int)(tb - ta);
      for (int i = 0; i < n; i += 2) {
        const int both = i + 1 < n ? 2 : 1;
        for (int kc = 0; kc < nkc; ++kc) {
          for (int g = 0; g < both; ++g, ++pos) {
            const int st = pos % T3_NST;
            if (pos >= T3_NST) rk_mbar_wait(empty + st, ((pos / T3_NST) - 1) & 1u);
            rk_mbar_expect_tx(full + st, T3_STAGE_BYTES);     // the whole box is written: out-of-range elements arrive as zeros
            rk_tma_2d(ring + (size_t)st * T3_STAGE_BYTES, &tmapE, kc * BK, (int)((nt0 + i + g) * BN), full + st);
          }
        }
      }
      ta = tb;
    }
    return;
  }

  // ---------------- consumers ----------------
  const int g = warp >> 3;                       // consumer group
  const int tg = threadIdx.x & 255, tx = tg & 15, ty = tg >> 4;
  const int x7 = tx & 7;                         // swizzle key of this thread's candidate rows (tx + 16 j: the low three bits are tx's)
  uint32_t run_base = 0;
  int64_t ta = t0;
  for (int64_t qt = qt_first; ta < t1; ++qt) {
    const int64_t tb = (qt + 1) * ntiles < t1 ? (qt + 1) * ntiles : t1;
    const int64_t nt0 = ta - qt * ntiles, q0 = qt * BQ;
    const int n = (int)(tb - ta);
    // (re)load the resident query tile, pair-interleaved; every consumer thread is past the previous run when the first barrier opens
    asm volatile("bar.sync 1, 512;" ::: "memory");
    {
      const int w4 = W >> 2;
      for (int idx = threadIdx.x; idx < (HEAD ? 2 : 1) * (BQ / 2) * w4; idx += 512) {
        const int part = idx / ((BQ / 2) * w4), rem = idx - part * (BQ / 2) * w4;
        const int pr = rem / w4, k = (rem - pr * w4) * 4;
        const int64_t qa = q0 + pr, qb = q0 + pr + 32;
        const float4 a = qa < Q ? ldg4(reinterpret_cast<const float4*>(qvec + qa * (int64_t)qw + part * W + k)) : f4zero();
        const float4 b = qb < Q ? ldg4(reinterpret_cast<const float4*>(qvec + qb * (int64_t)qw + part * W + k)) : f4zero();
        float* row = q2 + ((size_t)part * (BQ / 2) + pr) * QS + 2 * k;
        *reinterpret_cast<float4*>(row) = make_float4(a.x, b.x, a.y, b.y);
        *reinterpret_cast<float4*>(row + 4) = make_float4(a.z, b.z, a.w, b.w);
      }
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    float tq[4];
    int cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t qa = q0 + ty + 16 * i, qb = qa + 32;
      tq[2 * i] = qa < Q ? __ldg(tscore + qa) : 0.f;
      tq[2 * i + 1] = qb < Q ? __ldg(tscore + qb) : 0.f;
    }
    const float* qrow0 = q2 + (size_t)ty * QS;
    const float* qrow1 = q2 + (size_t)(ty + 16) * QS;
    const size_t part2 = (size_t)(BQ / 2) * QS;   // offset of the second part (head queries)

    for (int i = g; i < n; i += 2) {
      const int both = (i | 1) < n ? 2 : 1;       // this pair of items has a partner for the other group
      const uint32_t base = run_base + (uint32_t)(i >> 1) * 2u * (uint32_t)nkc;
      const int64_t n0 = (nt0 + i) * BN;
      u64 acc[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = 0ull;
      for (int kc = 0; kc < nkc; ++kc) {
        const uint32_t pos = base + (both == 2 ? 2u * kc + g : (uint32_t)kc);
        const int st = pos % T3_NST;
        rk_mbar_wait(full + st, (pos / T3_NST) & 1u);
        const unsigned char* eb = ring + (size_t)st * T3_STAGE_BYTES + tx * 128;
        const int klim = (W - kc * BK) < BK ? (W - kc * BK) : BK;   // multiple of 4
        const float* qa0 = qrow0 + 2 * kc * BK;
        const float* qa1 = qrow1 + 2 * kc * BK;
#pragma unroll 2
        for (int k = 0; k < klim; k += 4) {
          ulonglong2 a[2][2], a2[2][2];
          float4 e[4];
          a[0][0] = *reinterpret_cast<const ulonglong2*>(qa0 + 2 * k);
          a[0][1] = *reinterpret_cast<const ulonglong2*>(qa0 + 2 * k + 4);
          a[1][0] = *reinterpret_cast<const ulonglong2*>(qa1 + 2 * k);
          a[1][1] = *reinterpret_cast<const ulonglong2*>(qa1 + 2 * k + 4);
          if (HEAD) {
            a2[0][0] = *reinterpret_cast<const ulonglong2*>(qa0 + part2 + 2 * k);
            a2[0][1] = *reinterpret_cast<const ulonglong2*>(qa0 + part2 + 2 * k + 4);
            a2[1][0] = *reinterpret_cast<const ulonglong2*>(qa1 + part2 + 2 * k);
            a2[1][1] = *reinterpret_cast<const ulonglong2*>(qa1 + part2 + 2 * k + 4);
          }
          const int sw = (((k >> 2) ^ x7) << 4);   // byte offset of k-chunk k/4 inside the swizzled 128-byte row
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = *reinterpret_cast<const float4*>(eb + 16 * j * 128 + sw);
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              u64 c = acc[p][j];
              c = seq_term3<OP>(c, e[j].x, a[p][0].x, a2[p][0].x);
              c = seq_term3<OP>(c, e[j].y, a[p][0].y, a2[p][0].y);
              c = seq_term3<OP>(c, e[j].z, a[p][1].x, a2[p][1].x);
              c = seq_term3<OP>(c, e[j].w, a[p][1].y, a2[p][1].y);
              acc[p][j] = c;
            }
        }
        __syncwarp();
        if (lane == 0) rk_mbar_arrive(empty + st);   // this warp is done with the stage
      }
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t nn = n0 + tx + 16 * j;
          const float s_lo = seq_finish<OP>(lo2(acc[p][j]), fin), s_hi = seq_finish<OP>(hi2(acc[p][j]), fin);
          cnt[2 * p] += (nn < n_rows && s_lo > tq[2 * p]) ? 1 : 0;
          cnt[2 * p + 1] += (nn < n_rows && s_hi > tq[2 * p + 1]) ? 1 : 0;
        }
    }
    // the 16 tx lanes of one ty live in one half-warp: butterfly inside 16 lanes, then one integer atomic per query and group
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = cnt[i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      const int64_t q = q0 + ty + 16 * (i >> 1) + 32 * (i & 1);
      if (tx == 0 && q < Q && c != 0) atomicAdd(counts + q, c);
    }
    run_base += (uint32_t)n * (uint32_t)nkc;
    ta = tb;
  }
}

// tensor map of a row shard for k_rank_tile3: fp32 [n_rows][W] with row stride ldE, 32 x 64 boxes, 128-byte swizzle, zeros out of range.
// cuTensorMapEncodeTiled is a host-side encoder (no GPU work); it is looked up once through the runtime.
typedef CUresult (*kgeb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool t3_tensor_map(const float* E_shard, int64_t n_rows, int W, int ldE, CUtensorMap* out) {
  static kgeb_encode_tiled_fn enc = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return (kgeb_encode_tiled_fn)fn;
  }();
  if (!enc || n_rows <= 0 || n_rows > 0x7fffffffLL) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ldE * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN}, estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(E_shard), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// launches the tile kernel of one op: the packed kernel on a grid of exactly SMs x resident CTAs, or (ops with a multiply or sinf; KGEB_RANK_TILE=scalar
// for A/B runs -- identical integer counts either way) the round-1 scalar kernel on its 2-D grid
template <int OP>
int launch_tile(kgeb_ctx* ctx, dim3 grid2d, int64_t items, cudaStream_t st, const float* E_shard, int64_t n_rows, const float* qvec,
                int qw, const float* tscore, int64_t Q, SeqFin fin, int32_t* counts) {
  static const bool scalar_tile = [] { const char* e = getenv("KGEB_RANK_TILE"); return e && !strcmp(e, "scalar"); }();
  if constexpr (op_packed<OP>()) {
    if (!scalar_tile) {
      static const bool no_t3 = [] { const char* e = getenv("KGEB_RANK_TILE"); return e && !strcmp(e, "packed2"); }();
      const size_t sm3 = t3_smem_bytes(ctx->we, op_two_part<OP>());
      CUtensorMap tmapE;
      if (!no_t3 && sm3 <= 227 * 1024 && t3_tensor_map(E_shard, n_rows, ctx->we, ctx->ldE, &tmapE)) {
        // the resident query tile fits: decoupled TMA producer / packed consumer kernel, one CTA per SM
        // (the attribute is per device and per function; setting it on every launch costs ~1 us and stays correct with several devices in one process)
        KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_rank_tile3<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        int64_t g3 = ctx->num_sms;
        if (g3 > items) g3 = items;
        k_rank_tile3<OP><<<(unsigned)g3, T3_THREADS, sm3, st>>>(tmapE, n_rows, ctx->we, qvec, qw, tscore, Q, fin, counts);
        return KGEB_OK;
      }
      auto kern = k_rank_tile2<OP>;
      const size_t tsm = ((size_t)(op_two_part<OP>() ? 4 : 2) * BQ * BKP + 2 * (BN / 2) * BPP) * sizeof(float);
      static int occ = 0;   // resident CTAs per SM of this instantiation, asked once (every device of a box is the same part)
      KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      if (occ == 0) {
        int o = 0;
        KGEB_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, tsm));
        occ = o < 1 ? 1 : o;
      }
      int64_t g = (int64_t)ctx->num_sms * occ;
      if (g > items) g = items;
      kern<<<(unsigned)g, 256, tsm, st>>>(E_shard, n_rows, ctx->ldE, ctx->we, qvec, qw, tscore, Q, fin, counts);
      return KGEB_OK;
    }
  }
  const size_t tsm = (size_t)(op_two_part<OP>() ? 6 : 4) * BQ * BKP * sizeof(float);
  KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_rank_tile<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
  k_rank_tile<OP><<<grid2d, 256, tsm, st>>>(E_shard, n_rows, ctx->ldE, ctx->we, qvec, qw, tscore, Q, fin, counts);
  return KGEB_OK;
}

int grid1d(const kgeb_ctx* ctx, int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int pairs_grid(const kgeb_ctx* ctx, int64_t pairs) {
  int64_t b = (pairs + PAIR_WARPS * 32 - 1) / (PAIR_WARPS * 32);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

int host_op_for(int kind, int side) {
  if (kind == KGEB_TRANSE_L1) return side == KGEB_SIDE_TAIL ? OP_L1_TAIL : OP_L1_HEAD;
  if (kind == KGEB_TRANSE_L2) return side == KGEB_SIDE_TAIL ? OP_L2_TAIL : OP_L2_HEAD;
  if (kind == KGEB_MODE_L1) return side == KGEB_SIDE_TAIL ? OP_L1_TAIL : OP_ML1_HEAD;
  if (kind == KGEB_MODE_L2) return side == KGEB_SIDE_TAIL ? OP_L2_TAIL : OP_ML2_HEAD;
  if (kind == KGEB_PROTATE) return side == KGEB_SIDE_TAIL ? OP_PSIN_TAIL : OP_PSIN_HEAD;
  return OP_DOT;  // DistMult, ComplEx, HolE, RotatE (linear in the candidate row)
}

SeqFin seq_fin(const kgeb_ctx* ctx) {
  SeqFin f;
  f.gamma = ctx->gamma; f.param = ctx->param;
  f.sub = (ctx->kind == KGEB_MODE_L1 || ctx->kind == KGEB_MODE_L2 || ctx->gamma > 0.f) ? 1 : 0;
  return f;
}

bool two_part_queries(int kind) {
  return kind == KGEB_TRANSE_L1 || kind == KGEB_TRANSE_L2 || kind == KGEB_MODE_L1 || kind == KGEB_MODE_L2 || kind == KGEB_PROTATE;
}

}  // namespace

extern "C" int kgeb_query_width(const kgeb_ctx* ctx, int side) {
  if (!ctx) return KGEB_ERR_INVALID;
  return (two_part_queries(ctx->kind) && side == KGEB_SIDE_HEAD) ? 2 * ctx->we : ctx->we;
}

extern "C" int kgeb_query_prep(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* R,
                               const int32_t* triples, int64_t Q, int side, float* qvec, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, E_shard && R && triples && qvec && Q >= 0 && n_rows >= 0 && row_offset >= 0, KGEB_ERR_INVALID,
               "query_prep: bad arguments");
  KGEB_REQUIRE(ctx, side == KGEB_SIDE_TAIL || side == KGEB_SIDE_HEAD, KGEB_ERR_INVALID, "unknown side");
  KGEB_REQUIRE(ctx, ctx->kind != KGEB_HAKE, KGEB_ERR_UNSUPPORTED,
               "HAKE has no 1-N ranking: its modulus term is one norm over the scoring call (models.py:397), so a candidate's score depends on its batch");
  KGEB_REQUIRE(ctx, ctx->kind == KGEB_ROTATE ? ctx->we == 2 * ctx->wr : ctx->we == ctx->wr, KGEB_ERR_UNSUPPORTED,
               "1-N ranking needs e_dim == r_dim (RotatE: e_dim == 2 * r_dim)");
  if (Q == 0) return KGEB_OK;
  const int qw = kgeb_query_width(ctx, side);
  k_query_prep<<<grid1d(ctx, Q * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      ctx->kind, side, ctx->gamma > 0.f ? 1 : 0, E_shard, row_offset, n_rows, ctx->ldE, R, ctx->ldR, ctx->we, triples, Q, qw, qvec);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_target_scores(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                                  const int32_t* triples, int64_t Q, int side, float* tscore, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, E_shard && qvec && triples && tscore && Q >= 0, KGEB_ERR_INVALID, "target_scores: bad arguments");
  if (Q == 0) return KGEB_OK;
  const int qw = kgeb_query_width(ctx, side);
  PairArgs a = {};
  a.Es = E_shard; a.row_offset = row_offset; a.n_rows = n_rows; a.ldE = ctx->ldE; a.W = ctx->we; a.qvec = qvec; a.qw = qw;
  a.tscore = tscore; a.triples = triples; a.Q = Q; a.side = side; a.fin = seq_fin(ctx);
  launch_pairs<PAIRS_TARGET>(host_op_for(ctx->kind, side), pairs_grid(ctx, Q), (cudaStream_t)stream, a);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_rank_1n(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                            const float* tscore, const int32_t* triples, int64_t Q, int side, const int64_t* filt_offsets,
                            const int32_t* filt_ids, int32_t* counts, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, E_shard && qvec && tscore && triples && counts && Q >= 0 && n_rows >= 0, KGEB_ERR_INVALID,
               "rank_1n: bad arguments");
  KGEB_REQUIRE(ctx, (filt_offsets == nullptr) == (filt_ids == nullptr), KGEB_ERR_INVALID, "filter CSR: give both arrays or neither");
  KGEB_REQUIRE(ctx, ctx->we % 4 == 0 && ctx->ldE % 4 == 0 && kgeb_aligned16(E_shard) && kgeb_aligned16(qvec),
               KGEB_ERR_ALIGNMENT, "rank_1n needs width, stride multiples of 4 floats and 16-byte aligned pointers");
  if (Q == 0 || n_rows == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int qw = kgeb_query_width(ctx, side);
  const int op = host_op_for(ctx->kind, side);
  const int64_t qtiles = (Q + BQ - 1) / BQ;
  const int64_t ntiles = (n_rows + BN - 1) / BN;
  // split the entity range so the grid covers ~4 CTAs per SM; each CTA keeps its counts in registers across its tiles
  int64_t split = ((int64_t)ctx->num_sms * 4 + qtiles - 1) / qtiles;
  if (split > ntiles) split = ntiles;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  dim3 grid((unsigned)qtiles, (unsigned)split);
  const SeqFin fin = seq_fin(ctx);
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_RANK_TILE, st);
  int rc = KGEB_OK;
#define KGEB_RANK_LAUNCH(OPV) rc = launch_tile<OPV>(ctx, grid, qtiles * ntiles, st, E_shard, n_rows, qvec, qw, tscore, Q, fin, counts)
  switch (op) {
    case OP_ML1_HEAD: KGEB_RANK_LAUNCH(OP_ML1_HEAD); break;
    case OP_ML2_HEAD: KGEB_RANK_LAUNCH(OP_ML2_HEAD); break;
    case OP_PSIN_TAIL: KGEB_RANK_LAUNCH(OP_PSIN_TAIL); break;
    case OP_PSIN_HEAD: KGEB_RANK_LAUNCH(OP_PSIN_HEAD); break;
    case OP_DOT: KGEB_RANK_LAUNCH(OP_DOT); break;
    case OP_L1_TAIL: KGEB_RANK_LAUNCH(OP_L1_TAIL); break;
    case OP_L1_HEAD: KGEB_RANK_LAUNCH(OP_L1_HEAD); break;
    case OP_L2_TAIL: KGEB_RANK_LAUNCH(OP_L2_TAIL); break;
    default: KGEB_RANK_LAUNCH(OP_L2_HEAD); break;
  }
  if (rc != KGEB_OK) return rc;
  KGEB_TIME_END(ctx, KGEB_TIMED_RANK_TILE, st);
  KGEB_LAUNCH_OK(ctx);
  if (filt_offsets)
    return kgeb_internal_filter_correct(ctx, E_shard, row_offset, n_rows, qvec, tscore, triples, Q, side, filt_offsets,
                                        filt_ids, counts, st);
  return KGEB_OK;
}

int kgeb_internal_host_op_for(int kind, int side) { return host_op_for(kind, side); }
kgeb_rank::SeqFin kgeb_internal_seq_fin(const kgeb_ctx* ctx) { return seq_fin(ctx); }

int kgeb_internal_filter_correct(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                                 const float* tscore, const int32_t* triples, int64_t Q, int side,
                                 const int64_t* filt_offsets, const int32_t* filt_ids, int32_t* counts, cudaStream_t st) {
  const int qw = kgeb_query_width(ctx, side);
  PairArgs a = {};
  a.Es = E_shard; a.row_offset = row_offset; a.n_rows = n_rows; a.ldE = ctx->ldE; a.W = ctx->we; a.qvec = qvec; a.qw = qw;
  a.tscore = const_cast<float*>(tscore); a.triples = triples; a.Q = Q; a.side = side; a.fin = seq_fin(ctx);
  a.offs = filt_offsets; a.ids = filt_ids; a.counts = counts;
  // nnz is on the device; size the grid from Q (lists are short) -- the kernel grid-strides over nnz
  launch_pairs<PAIRS_FILTER>(host_op_for(ctx->kind, side), pairs_grid(ctx, Q * 8), st, a);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

// grouped != 0: `wl` is ordered by entity (kgeb_internal_wl_group) and wl_hdr points at its length; the kinds the tensor-core pass
// serves have a grouped kernel, any other op takes the plain one (the order of a worklist never changes what is counted)
int kgeb_internal_refine_worklist(kgeb_ctx* ctx, const float* E_shard, int64_t n_rows, const float* qvec, const float* tscore,
                                  int64_t Q, int side, const uint32_t* wl_hdr, uint32_t wl_cap, const uint2* wl, int32_t* counts,
                                  cudaStream_t st, int grouped) {
  PairArgs a = {};
  a.Es = E_shard; a.row_offset = 0; a.n_rows = n_rows; a.ldE = ctx->ldE; a.W = ctx->we; a.qvec = qvec;
  a.qw = kgeb_query_width(ctx, side); a.tscore = const_cast<float*>(tscore); a.Q = Q; a.side = side; a.fin = seq_fin(ctx);
  a.wl_hdr = wl_hdr; a.wl_cap = wl_cap; a.wl = wl; a.counts = counts;
  const int op = host_op_for(ctx->kind, side), blocks = ctx->num_sms * 16;
  if (grouped && (op == OP_DOT || op == OP_L2_TAIL || op == OP_L2_HEAD)) {
    constexpr int SRC = PAIRS_GROUPED;
    if (op == OP_DOT) KGEB_PAIRS_LAUNCH(OP_DOT)
    else if (op == OP_L2_TAIL) KGEB_PAIRS_LAUNCH(OP_L2_TAIL)
    else KGEB_PAIRS_LAUNCH(OP_L2_HEAD)
  } else {
    launch_pairs<PAIRS_WORKLIST>(op, blocks, st, a);
  }
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

// ---- ranking of scores that came from somewhere else (validate() driven by a foreign model's predict) ---------------------------
namespace {
__global__ void __launch_bounds__(256) k_count_greater(const float* __restrict__ scores, int64_t n, int64_t target, int32_t* __restrict__ count) {
  const float t = __ldg(scores + target);
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += (i != target && __ldg(scores + i) > t) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c != 0) atomicAdd(count, c);
}
}  // namespace

extern "C" int kgeb_count_greater(kgeb_ctx* ctx, const float* scores, int64_t n, int64_t target, int32_t* count, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, scores && count && n > 0 && target >= 0 && target < n, KGEB_ERR_INVALID, "count_greater: bad arguments");
  k_count_greater<<<grid1d(ctx, n, 256), 256, 0, (cudaStream_t)stream>>>(scores, n, target, count);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
