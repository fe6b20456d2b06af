// rank_tc_kernels.cu -- 1-N filtered ranking for the dot-product score functions (DistMult models.py:171-172, ComplEx
// 212-222, HolE 231-237 as query x E^T) on the 5th-generation tensor cores.
//
//   counts[q] = #{c in shard, c != target(q) : score(q,c) > tscore[q]}  -  (filter correction, rank_kernels.cu)
//
// The Q x N score matrix never exists: tcgen05.mma accumulates a 128-query x 256-entity tile in TMEM, eight epilogue
// warps read it back with tcgen05.ld and keep nothing but integer counters.
//
// Exactness.  A 16-bit contraction cannot reproduce the fp32 sequential-k scores of rank_common.cuh bit for bit, so the
// tensor cores are used as a FILTER with a proven error bound and the few undecided pairs are re-scored exactly:
//   operands   Q = 2^b q and Ee = 2^a e (one power-of-two scale per query batch / per packed table, chosen on the device so
//              that the largest magnitude lands in [2^13, 2^14): exact, and it keeps the planes in fp16's normal range) are
//              split into fp16 planes  Q = qh + ql + rq,  Ee = eh + el + re.  fp16 keeps 11 significant bits, so
//              |ql| <= 2^-11 |Q| and |rq| <= 2^-22 |Q| + 2^-25 elementwise (the constant covers a subnormal low plane).
//              The kernel accumulates qh.eh + qh.el + ql.eh in fp32 (3 MMAs per k-step on the same staged tiles) and DROPS
//              ql.el + rq.Ee + (qh+ql).re  <=  3.001 * 2^-22 ||Q|| ||Ee||  +  2^-25 sqrt(K) (||Q|| + ||Ee||)   (Cauchy-Schwarz),
//   reference  the sequential fp32 sum itself: |seq - exact| <= u T + 16.01 u Z,  u = 2^-24, T = sum_k |q_k e_k| <= ||q|| ||e||,
//              Z = sum over the K/16 blocks j of T_j (T_j = the sum of |q_k e_k| up to the end of block j): every partial sum
//              of block j is at most T_j in magnitude and is rounded once per added term,
//   tensor     tcgen05.mma kind::f16 with fp32 accumulation, MEASURED (profiles/micro/tc_numerics.cu, r02_tc_numerics.md):
//              the 16 products and the incoming accumulator of one instruction are aligned to the largest exponent E among
//              them, each truncated toward zero to a multiple of 2^(E-25) (two bits below the fp32 ulp of the largest),
//              summed exactly, and the sum is truncated toward zero to fp32.  One instruction therefore loses less than
//              16 * 2^(E-25) + ulp(result) <= 10 u B, B = the sum of the magnitudes accumulated so far <= 1.002 T_j; the
//              bound budgets 12 u B per instruction, 3 instructions per block: 36.1 u Z,
//   spread     Z <= sum_j ||q_(blocks <= j)|| ||e_(blocks <= j)|| <= Aq * Ae  (Cauchy-Schwarz over j) with the WEIGHTED norms
//              A^2 = sum_k w_k x_k^2, w_k = number of blocks at or after k's.  For rows whose mass is spread over k this is
//              half of the worst case (K/16) ||q|| ||e|| that a plain norm bound has to assume,
//   bound      |S~ / 2^(a+b) - S_seqk| <= 1.02 * u * (13 ||q|| ||e|| + 52.11 Aq Ae) + 1.02 * 2^-25 sqrt(K) (||q|| 2^-a + ||e|| 2^-b),
//   epilogue   S~ > hi -> counted;   S~ < lo -> dropped;   otherwise the pair (q,c) goes to a worklist that k_seqk_pairs
//              (rank_kernels.cu) re-scores with the exact sequential-k arithmetic.  The bound used per 128-entity slice takes
//              the slice's largest ||e|| and largest Ae, so a thread compares against two constants.
// The integer result is therefore identical to k_rank_tile's.  KGEB_TC_BF16 (single bf16 plane, unscaled, bound forced to 0)
// is the approximate mode: one MMA per k-step, ranks can differ at near-ties (tolerance stated in the tests).
//
// TransE with the L2 norm (models.py:186-190, norm = 2) ranks by a distance, not by a dot product, but
//   D'^2(q', e) = ||q'||^2 - 2 S,   S = q'.e - ||e||^2 / 2,      q' = fl(s + p) (tail) or fl(o - p) (head),
// and S IS a contraction once three extra slots ride along in the LAST K=16 block(s): the entity row carries -W_e split into three
// fp16 values (W_e = 2^(2a) ||e||^2 / 2 from a double-precision sum; hi / mid / lo parts divided by 2^15, 2^4, 2^-7), the query
// row carries the matching powers of two.  The kernel, its pipeline and its filter epilogue are unchanged; what changes is the
// per-query centre of the band and a few more terms in it (k_tc_l2_query, the L2 branch of the epilogue):
//   decision   score_c > t  <=>  the reference's sequential accumulator acc_c (sum of fl(r_k^2), r_k = fl(q_k - e_k) or
//              fl(fl(e_k + p_k) - o_k)) lies below A(t); with eps = 1.01 u (2 D_t + |gamma|), D_t = gamma - t, both roundings of
//              gamma - sqrt(acc) are covered by   acc <= A_lo = (D_t - eps)^2  =>  counted,   acc >= A_hi = (D_t + eps)^2  =>  not
//              (gamma <= 0, where the score is +sqrt(acc): the roles swap and the query operand is negated),
//   reference  |acc - D'^2| <= eta D'^2 + c D',  eta = (W + 8) u (non-negative terms: the sequential sum's relative error) and, on the
//              head side only, c = 2.1 u (||e|| + ||p|| + ||q'||) for the reference's fl(e + p) and this kernel's fl(o - p);
//              both sides are monotone in D' beyond c/2, so checking the two boundary values X_lo, X_hi of D'^2 suffices,
//   tensor     the band of the dot-product case with T + V and Z + 2V in place of T and Z (V = ||e||^2 / 2 <= nmax^2 / 2: the
//              extra slots add their magnitude to at most the last two blocks) and 2^-31 V for the three-way split of W_e.
// The operands of this mode share the exponent window: entities are scaled to [2^9, 2^10) so that W_e / 2^15 fits fp16 for every
// width up to 4000, the query scale is the entities' or smaller (the extra slots need 2^(aQ - aE) in [2^-17, 1]; a query batch
// 2^17 times larger than the table is declared degenerate: empty planes, every pair re-scored).
//
// Data movement.  Both operands are pre-packed (k_tc_pack) into the exact shared-memory image the UMMA descriptors
// expect -- no-swizzle K-major core matrices, [tile][k-chunk][plane][row][8 x bf16] -- so one pipeline stage is two
// contiguous bulk async copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier; no tensor map is needed.
// Roles: warp 0 = copy producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator, warps 4-11 = epilogue.
#include <cuda.h>   // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "rank_common.cuh"

using namespace kgeb_rank;

namespace {

constexpr int TC_BM = 128;       // queries per tile  (UMMA M, one TMEM lane per query)
constexpr int TC_BN = 256;       // entities per tile (UMMA N, one TMEM column per entity)
constexpr int TC_STAGES = 4;
constexpr int TC2_STAGES = 6;    // CTA-pair kernel: a stage holds 16 KB of queries + 16 KB (half) of the entity tile
constexpr int TC_CH_TOTAL = 8;   // 16-byte k-chunks per stage summed over planes (8 / planes chunks of K each)
constexpr int TC_THREADS = 384;
constexpr uint32_t TC_A_STAGE = TC_BM * TC_CH_TOTAL * 16;  // 16 KB
constexpr uint32_t TC_B_STAGE = TC_BN * TC_CH_TOTAL * 16;  // 32 KB
constexpr int TC_WBUF = 128;     // undecided pairs buffered per epilogue warp before one aggregated append
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * (TC_A_STAGE + TC_B_STAGE) + 1024 + 256 + 8 * TC_WBUF * 8;
constexpr int TC_HDR_BYTES = 256;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t"
      "}" ::"r"(tc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The forms below are executed by ALL lanes of the (converged) issuing warp; elect.sync picks the one lane that performs the operation --
// always the same lane for the full mask, so the commits track the MMAs.  In a branch of the shape `if (lane == 0)` ptxas wraps every
// UTCHMMA in an ELECT / PLOP3 / BRA.U.ANY loop (one trip per distinct active thread), which cost more than the instruction itself.
__device__ __forceinline__ void tc_mma_bf16_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_w(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta_w(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(tc_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrives on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// ---- CTA-pair forms (cta_group::2: one MMA spans two SMs; the leader CTA of the pair issues it) ----
__device__ __forceinline__ uint32_t tc_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void tc_mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(tc_smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// 2-D TMA box into THIS CTA's shared memory whose transaction bytes complete on the barrier at the same offset in the LEADER CTA
// (cp.async.bulk.tensor ... .cta_group::2: the one copy form whose destination and barrier may sit in different CTAs of the pair)
__device__ __forceinline__ void tc_tma_2d_pair(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 rb;\n\t"
      "mapa.shared::cluster.u32 rb, %4, 0;\n\t"
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [rb];\n\t"
      "}" ::"r"(tc_smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(tc_smem_u32(bar))
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (thread = lane = query row).  The load is
// asynchronous: the registers are valid only after tc_ld_wait, which names them as in/out operands so that no use can
// be scheduled before it.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// one accumulator against the row's band [lo, hi]:  cmsk |= bit when s > hi (counted);  umsk |= bit when lo <= s <= hi
// (undecided).  Masks instead of a running count: the target and the CSR filter ids are then excluded by clearing bits, and
// the count is one popc per 32 columns -- the same four instructions per accumulator.
#define TC_CMP(S, HI, LO, CMSK, UMSK, BIT)                                                                                \
  asm("{\n\t.reg .pred ph, pu;\n\tsetp.gt.f32 ph, %2, %3;\n\t@ph or.b32 %0, %0, %5;\n\tsetp.ge.and.f32 pu, %2, %4, !ph;\n\t" \
      "@pu or.b32 %1, %1, %5;\n\t}"                                                                                       \
      : "+r"(CMSK), "+r"(UMSK)                                                                                            \
      : "r"(S), "f"(HI), "f"(LO), "n"(BIT))
template <int J>
__device__ __forceinline__ void tc_cmp32(const uint32_t (&r)[32], float hi, float lo, uint32_t& c0, uint32_t& c1, uint32_t& m0, uint32_t& m1) {
  if constexpr (J < 32) {  // two independent mask chains
    TC_CMP(r[J], hi, lo, c0, m0, 1u << J);
    TC_CMP(r[J + 1], hi, lo, c1, m1, 1u << (J + 1));
    tc_cmp32<J + 2>(r, hi, lo, c0, c1, m0, m1);
  }
}

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes stored contiguously (128 B);
// LBO = byte distance between the two 16-byte k-chunks of one K=16 step, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}
// the same descriptor as a variable low word (start address and LBO fields) and the constant high word (SBO = 128 bytes, version bit 46)
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint64_t tc_desc64(uint32_t lo) { return ((uint64_t)(((128u >> 4) & 0x3FFFu) | (1u << 14)) << 32) | lo; }
// instruction descriptor: fp32 accumulate, A and B both K-major, N = 256, M = 128; operand format fp16 (0) or bf16 (1)
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
constexpr uint32_t TC_IDESC_BF16 = TC_IDESC_F16 | (1u << 7) | (1u << 10);

// power-of-two scale that moves a largest magnitude `amax` (given as its fp32 bits) into [2^13, 2^14); 1 for an all-zero table
__device__ __forceinline__ int tc_scale_exp(uint32_t bits, int target) {
  const float amax = __uint_as_float(bits);
  if (!(amax > 0.f) || isinf(amax)) return 0;
  const int a = target - ilogbf(amax);
  return a < -60 ? -60 : (a > 60 ? 60 : a);
}
__device__ __forceinline__ float tc_scale_from_bits(uint32_t bits) { return ldexpf(1.f, tc_scale_exp(bits, 13)); }
// L2 mode: entities to [2^9, 2^10); the query scale never exceeds the entities' and stays within 2^-17 of it.  scQ = 0 flags the
// degenerate batch (queries more than 2^17 times larger than the table): empty query planes, every pair undecided.
__device__ __forceinline__ void tc_l2_scales(uint32_t ebits, uint32_t qbits, float& scE, float& scQ) {
  const int aE = tc_scale_exp(ebits, 9);
  int aQ = tc_scale_exp(qbits, 9);
  scE = ldexpf(1.f, aE);
  if (__uint_as_float(qbits) > 0.f && aQ < aE - 17) { scQ = 0.f; return; }
  if (aQ > aE || !(__uint_as_float(qbits) > 0.f)) aQ = aE;
  scQ = ldexpf(1.f, aQ);
}

struct TcArgs {
  const unsigned char* packedQ;  // [qtiles][nkc][planes][128][8] bf16
  const unsigned char* packedE;  // [ntiles][nkc][planes][256][8] bf16
  const float* qnorm;            // [qtiles*128]  upper bound of ||q||_2
  const float* qwnorm;           // [qtiles*128]  upper bound of the weighted norm Aq (exact mode)
  const float* emax;             // [ntiles*2]    max over each 128-entity slice of the upper bound of ||e||_2
  const float* ewmax;            // [ntiles*2]    max over each slice of the weighted norm Ae (exact mode)
  const uint32_t* escale_bits;   // largest |E| of the packed rows / largest |q| of the batch as fp32 bits (exact mode; else NULL)
  const uint32_t* qscale_bits;
  const float* tscore;           // [Q] exact target scores
  const int32_t* triples;        // [Q,3]
  const int64_t* filt_offs;      // [Q+1] CSR of known-true GLOBAL entity ids per query, sorted ascending (or NULL)
  const int32_t* filt_ids;
  int32_t* counts;               // [Q]
  uint32_t* wl_hdr;              // [0] = pairs appended (may exceed cap), [1] = cap
  uint2* wl;                     // (q, shard-local c)
  uint32_t wl_cap;
  int64_t Q, row_offset, n_rows;
  int nkc, planes, qtiles, ntiles, side, qgroup;
  int qtiles_real;               // CTA-pair kernel: `qtiles` counts PAIRS of query tiles, this is the number of 128-query tiles
  uint32_t idesc;
  float c_norm, c_wnorm, c_abs, t_slack;   // band = c_norm ||q|| ||e|| + c_wnorm Aq Ae + c_abs (||q|| / 2^a + ||e|| / 2^b)
  // TransE-L2 (see the header comment): per-query arrays [6][qpad] = ||q'||^2, A_lo, A_hi, sqrt(A_hi), D_t bound, c_q
  const float* l2q; int64_t l2_stride; int l2, l2_reverse;
  float l2_eta, l2_c_e, l2_c_v;
};

// item -> (entity tile, query tile).  Query tiles are walked in groups of `qgroup`: inside a group the query tile varies
// fastest, so the CTAs running at the same time share a handful of entity tiles (L2 hits) without all hammering one.
__device__ __forceinline__ void tc_item(const TcArgs& a, int64_t it, int64_t& tile, int64_t& qt) {
  const int64_t G = a.qgroup, per = G * a.ntiles, full = a.qtiles / G;
  if (it < full * per) {
    const int64_t g = it / per, rem = it - g * per;
    tile = rem / G;
    qt = g * G + (rem - tile * G);
  } else {
    const int64_t rem = it - full * per, Gl = a.qtiles - full * G;
    tile = rem / Gl;
    qt = full * G + (rem - tile * Gl);
  }
}

// the work unit (CTA, or CTA pair) and their number.  Read from blockIdx / gridDim at every use: hoisting them into 64-bit locals ahead of
// the role split made ptxas drop the uniform datapath for the single-thread producer / MMA-issue loops (UIADD3 / ULOP3 / USHF 303 -> 11 in
// SASS) and cost the approximate mode 8-18 % (measured A/B on one box)
template <int NCTA> __device__ __forceinline__ unsigned tc_unit() { return NCTA == 2 ? blockIdx.x >> 1 : blockIdx.x; }
template <int NCTA> __device__ __forceinline__ unsigned tc_units() { return NCTA == 2 ? gridDim.x >> 1 : gridDim.x; }
// NCTA = 1: one CTA per tile (128 queries x 256 entities).  NCTA = 2 (cta_group::2, clusters of two CTAs on the two SMs of a TPC): one MMA
// covers 256 queries x 256 entities -- each CTA stages its own 128-query tile and only HALF of the entity tile (128 rows), so the
// shared-memory fill and fetch per MMA flop drop by a third; the accumulator rows of a CTA's queries live in its own TMEM and its
// epilogue is unchanged.  Both CTAs fetch their stage with TMA boxes of the packed images (tensor maps over 64-bit words: one box = the 8
// chunk columns of a stage) whose bytes complete on the LEADER's barrier (.cta_group::2); the leader (cluster rank 0) issues the MMAs and
// its commits arrive on the stage / accumulator barriers of BOTH CTAs (multicast).  (First version: 1-D bulk copies completing on each
// CTA's own barrier, the peer forwarding "stage full" with a remote arrive -- 18.5 ms against 17.0 ms for one CTA per tile on C5 bf16.)
template <int NCTA>
__global__ void __launch_bounds__(TC_THREADS, 1) k_rank_tc_t(TcArgs a, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmE) {
  constexpr int STAGES = NCTA == 2 ? TC2_STAGES : TC_STAGES;
  constexpr uint32_t BROWS = TC_BN / NCTA;                  // entity rows of a stage held by this CTA
  constexpr uint32_t B_STAGE = TC_B_STAGE / NCTA;
  const uint32_t crank = NCTA == 2 ? tc_cluster_rank() : 0u;
  const bool leader = crank == 0;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smA = smem;
  unsigned char* smB = smem + STAGES * TC_A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (TC_A_STAGE + B_STAGE));
  uint64_t* full = bars;                     // [STAGES]  copies landed (pair: the leader's copy collects both CTAs' bytes)
  uint64_t* empty = bars + STAGES;           // [STAGES]  MMAs that read the stage completed
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]       accumulator ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;        // [2]       accumulator drained by all 8 epilogue warps (of both CTAs: the leader's copy counts)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint2* wbuf_all = reinterpret_cast<uint2*>(smem + STAGES * (TC_A_STAGE + B_STAGE) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes;
  const int ch_per_stage = TC_CH_TOTAL / planes;           // k-chunks (8 bf16 of K) per stage
  const int nstage = (a.nkc + ch_per_stage - 1) / ch_per_stage;
  const int64_t items = (int64_t)a.qtiles * a.ntiles;

  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) a.wl_hdr[1] = a.wl_cap;
    for (int i = 0; i < STAGES; ++i) { tc_mbar_init(full + i, 1); tc_mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { tc_mbar_init(acc_full + i, 1); tc_mbar_init(acc_empty + i, 8 * NCTA); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (NCTA == 2) {   // the same warp of both CTAs, the same slot address
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (NCTA == 2) tc_cluster_sync(); else __syncthreads();   // pair: the peer's barriers exist before anything arrives on them
  tc_fence_after();
  // The CTA owns all 512 columns of its SM's tensor memory, so the allocation starts at column 0, lane 0.  Checked once; the MMA issuer then
  // addresses the accumulators with constants -- a base loaded from shared memory is a VECTOR register to the compiler, and every MMA paid
  // an ELECT / R2UR.BROADCAST / BRA.U.ANY round trip to move it to the uniform datapath (SASS of the previous build).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===== producer: one thread streams the packed operand stages =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const size_t a_tile_bytes = (size_t)a.nkc * planes * TC_BM * 16, b_tile_bytes = (size_t)a.nkc * planes * TC_BN * 16;
      for (int64_t it = tc_unit<NCTA>(); it < items; it += tc_units<NCTA>()) {
        int64_t tile, qt;
        tc_item(a, it, tile, qt);
        if (NCTA == 2) {   // this CTA's query tile of the pair; an odd last pair re-reads the last tile (its epilogue rows are all q >= Q)
          qt = 2 * qt + crank;
          if (qt >= a.qtiles_real) qt = a.qtiles_real - 1;
        }
        const unsigned char* srcA = a.packedQ + (size_t)qt * a_tile_bytes;
        const unsigned char* srcB = a.packedE + (size_t)tile * b_tile_bytes;
        for (int s = 0; s < nstage; ++s) {
          const int c0 = s * ch_per_stage;
          const int nch = (a.nkc - c0 < ch_per_stage ? a.nkc - c0 : ch_per_stage) * planes;  // 16-byte chunk columns
          tc_mbar_wait(empty + stage, phase ^ 1u);
          if (NCTA == 1) {
            const uint32_t ab = (uint32_t)nch * TC_BM * 16, bb = (uint32_t)nch * BROWS * 16;
            tc_mbar_expect_tx(full + stage, ab + bb);
            tc_bulk_g2s(smA + stage * TC_A_STAGE, srcA + (size_t)c0 * planes * TC_BM * 16, ab, full + stage);
            tc_bulk_g2s(smB + stage * B_STAGE, srcB + (size_t)c0 * planes * TC_BN * 16, bb, full + stage);
          } else {
            // whole boxes always (a short last stage reads on into the next tile's columns or the zero fill past the image)
            if (leader) tc_mbar_expect_tx(full + stage, 2u * (TC_A_STAGE + B_STAGE));
            const int colq = (int)(qt * a.nkc + c0) * planes, cole = (int)(tile * a.nkc + c0) * planes;
            tc_tma_2d_pair(smA + stage * TC_A_STAGE, &tmQ, 0, colq, full + stage);
            tc_tma_2d_pair(smB + stage * B_STAGE, &tmE, (int)(crank * (BROWS * 2)), cole, full + stage);   // this CTA's 128 of the 256 rows (64-bit words)
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (NCTA == 2 && !leader) {
      // peer CTA: the leader issues the MMAs of the pair
    } else {
      // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (of the leader CTA in a pair) =====
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t lboA = (uint32_t)planes * TC_BM * 16, lboB = (uint32_t)planes * BROWS * 16;
      const uint32_t idesc = NCTA == 2 ? (a.idesc & ~(0x1Fu << 24)) | ((uint32_t)((TC_BM * 2) >> 4) << 24) : a.idesc;   // M = 256 across the pair
      for (int64_t it = tc_unit<NCTA>(); it < items; it += tc_units<NCTA>()) {
        tc_mbar_wait(acc_empty + acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        uint32_t accumulate = 0;
        // Stages are consumed in PAIRS: the MMAs of two stages are issued back to back and only then are both stages handed
        // back to the producer (two commits in a row).  Every commit between groups of MMAs costs the tensor pipe a hand-off
        // of ~0.3 us (measured: the kernel's time follows the number of MMA/commit alternations per tile -- 13, 7 or 4 -- not
        // the bytes moved or the depth of the ring), so with 4 stages two groups of two alternate.
        for (int s = 0; s < nstage; s += 2) {
          uint32_t st_idx[2];
          int nin = 0;
          for (int h = 0; h < 2 && s + h < nstage; ++h) {
            const int c0 = (s + h) * ch_per_stage;
            const int nk16 = (a.nkc - c0 < ch_per_stage ? a.nkc - c0 : ch_per_stage) >> 1;  // K=16 steps in this stage
            tc_mbar_wait(full + stage, phase);
            tc_fence_after();
            // descriptors of the stage's first K=16 step; the next step / the other plane is a constant added to the 14-bit start-address
            // field (shared memory ends below 2^18 bytes, so the field never carries): one uniform add per operand per MMA instead of
            // rebuilding the descriptor
            // (kept as 32-bit low words + one constant high word so that the arithmetic stays on the uniform datapath)
            uint32_t la = tc_desc_lo(tc_smem_u32(smA + stage * TC_A_STAGE), lboA), lb = tc_desc_lo(tc_smem_u32(smB + stage * B_STAGE), lboB);
            const uint32_t stepA = (uint32_t)(2 * planes) * TC_BM, stepB = (uint32_t)(2 * planes) * BROWS;   // (2 chunk columns x planes x rows x 16 B) >> 4
            for (int j = 0; j < nk16; ++j, la += stepA, lb += stepB) {
              if (NCTA == 2) {
                tc_mma_bf16_2cta_w(tmem_d, tc_desc64(la), tc_desc64(lb), idesc, accumulate);  // qh . eh
                if (planes == 2) {
                  tc_mma_bf16_2cta_w(tmem_d, tc_desc64(la), tc_desc64(lb + BROWS), idesc, 1);  // qh . el  (the other plane: rows x 16 B further)
                  tc_mma_bf16_2cta_w(tmem_d, tc_desc64(la + TC_BM), tc_desc64(lb), idesc, 1);  // ql . eh
                }
              } else {
                tc_mma_bf16_w(tmem_d, tc_desc64(la), tc_desc64(lb), idesc, accumulate);  // qh . eh
                if (planes == 2) {
                  tc_mma_bf16_w(tmem_d, tc_desc64(la), tc_desc64(lb + BROWS), idesc, 1);  // qh . el
                  tc_mma_bf16_w(tmem_d, tc_desc64(la + TC_BM), tc_desc64(lb), idesc, 1);  // ql . eh
                }
              }
              accumulate = 1;
            }
            st_idx[nin++] = stage;
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          for (int h = 0; h < nin; ++h) {   // the stages may be overwritten once these MMAs have read them
            if (NCTA == 2) tc_commit_2cta_w(empty + st_idx[h]); else tc_commit_w(empty + st_idx[h]);
          }
        }
        if (NCTA == 2) tc_commit_2cta_w(acc_full + acc); else tc_commit_w(acc_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).. and columns 128*((w-4)/4).. =====
    const int lq = warp & 3, chalf = (warp - 4) >> 2;
    uint2* wbuf = wbuf_all + (warp - 4) * TC_WBUF;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t wcnt = 0;  // pairs waiting in wbuf (warp-uniform)
    uint32_t acc = 0, acc_phase = 0;
    // exact mode: the accumulators carry the scores times 2^(a+b) (operand scales, powers of two: every product below is exact)
    float scE = a.escale_bits ? tc_scale_from_bits(__ldg(a.escale_bits)) : 1.f;
    float scQ = a.qscale_bits ? tc_scale_from_bits(__ldg(a.qscale_bits)) : 1.f;
    bool degenerate = false;
    if (a.l2) {
      tc_l2_scales(__ldg(a.escale_bits), __ldg(a.qscale_bits), scE, scQ);
      degenerate = scQ == 0.f;
      if (degenerate) scQ = 1.f;
    }
    const float sc = scE * scQ;

    auto flush = [&]() {
      __syncwarp();
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.wl_hdr, wcnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < wcnt; i += 32)
        if (base + i < a.wl_cap) a.wl[base + i] = wbuf[i];
      __syncwarp();
      wcnt = 0;
    };

    for (int64_t it = tc_unit<NCTA>(); it < items; it += tc_units<NCTA>()) {
      int64_t tile, qt;
      tc_item(a, it, tile, qt);
      if (NCTA == 2) qt = 2 * qt + crank;   // past the last tile for the second CTA of an odd last pair: every q >= Q below
      const int64_t q = qt * TC_BM + lq * 32 + lane;
      const bool qok = q < a.Q;
      const float t = qok ? __ldg(a.tscore + q) : 0.f;
      const float qn = qok ? __ldg(a.qnorm + q) : 0.f;
      const float epsq = __fmaf_rn(a.c_norm, qn, a.c_abs / scQ);                       // multiplies the slice's largest ||e||
      const float epsw = a.c_wnorm * ((qok && a.qwnorm) ? __ldg(a.qwnorm + q) : 0.f);  // multiplies the slice's largest Ae
      const float absq = a.c_abs * qn / scE;
      const int64_t tgt = qok ? (int64_t)__ldg(a.triples + 3 * q + (a.side == KGEB_SIDE_TAIL ? 2 : 0)) - a.row_offset : -1;
      const int64_t n0 = tile * TC_BN + chalf * 128;   // shard-local id of this warp's first column
      const float nmax = __ldg(a.emax + tile * 2 + chalf);
      const float wmax = a.ewmax ? __ldg(a.ewmax + tile * 2 + chalf) : 0.f;
      float band = __fmaf_rn(epsq, nmax, __fmaf_rn(epsw, wmax, absq));
      float hi, lo;
      if (!a.l2) {
        hi = __fmul_rn(__fadd_rn(__fmaf_rn(fabsf(t), a.t_slack, t), band), sc);
        lo = __fmul_rn(__fsub_rn(__fmaf_rn(-fabsf(t), a.t_slack, t), band), sc);
      } else {
        const int64_t qi = qok ? q : 0;
        const float qn2 = __ldg(a.l2q + qi), A_lo = __ldg(a.l2q + a.l2_stride + qi), A_hi = __ldg(a.l2q + 2 * a.l2_stride + qi);
        const float sA_hi = __ldg(a.l2q + 3 * a.l2_stride + qi), Dt_up = __ldg(a.l2q + 4 * a.l2_stride + qi);
        const float c = __fmaf_rn(a.l2_c_e, nmax, __ldg(a.l2q + 5 * a.l2_stride + qi));
        const float om = 1.f - a.l2_eta;
        const float X_lo = __fmaf_rn(A_lo, om, -c * Dt_up);                     // acc <= A_lo is certain for D'^2 <= X_lo
        const float y0 = (c + sA_hi) / om;
        const float X_hi = __fmaf_rn(c, y0, A_hi) / om;                         // acc >= A_hi is certain for D'^2 >= X_hi
        band = __fmaf_rn(a.l2_c_v * nmax, nmax, band);
        const float slack = 4.8e-7f * (qn2 + X_hi);                            // fp32 rounding of the handful of operations here
        float hiS = 0.5f * (qn2 - X_lo) + band + slack;                         // S >= hiS  =>  D'^2 <= X_lo
        const float loS = 0.5f * (qn2 - X_hi) - band - slack;                   // S <= loS  =>  D'^2 >= X_hi
        if (A_lo < 0.f || X_lo < 0.f) hiS = __int_as_float(0x7f800000);         // nothing can be that close
        if (!a.l2_reverse) { hi = hiS * sc; lo = loS * sc; }
        else { hi = -loS * sc; lo = -hiS * sc; }                                // the query operand was negated: S~ -> -S~
        if (degenerate) { hi = __int_as_float(0x7f800000); lo = __int_as_float(0xff800000); }
      }
      int cnt0 = 0;
      // cursor into this query's sorted filter list: first id at or beyond this warp's first column (global id space)
      int64_t fpos = 0, fend = 0;
      if (a.filt_offs && qok) {
        int64_t lo_i = __ldg(a.filt_offs + q);
        fend = __ldg(a.filt_offs + q + 1);
        const int64_t g0 = a.row_offset + n0;
        int64_t hi_i = fend;
        while (lo_i < hi_i) {
          const int64_t mid = (lo_i + hi_i) >> 1;
          if ((int64_t)__ldg(a.filt_ids + mid) < g0) lo_i = mid + 1; else hi_i = mid;
        }
        fpos = lo_i;
      }
      int64_t fnext = fpos < fend ? (int64_t)__ldg(a.filt_ids + fpos) - a.row_offset : INT64_MAX;   // shard-local column of the next filtered id

      // one 32-column chunk: counts, undecided mask, padding correction, append of the undecided pairs
      auto process = [&](uint32_t (&r)[32], int jc) {
        uint32_t k0 = 0, k1 = 0, m0 = 0, m1 = 0;
        tc_cmp32<0>(r, hi, lo, k0, k1, m0, m1);
        uint32_t cmask = k0 | k1, mask = m0 | m1;
        const int64_t c0 = n0 + jc;
        const int64_t left = a.n_rows - c0;                     // columns of this chunk that are real entities
        uint32_t keep = 0xffffffffu;
        if (left < 32) {                                        // warp-uniform: only in the last entity tile
          const int nvalid = left < 0 ? 0 : (int)left;          // padding rows are all-zero: never candidates
          keep = nvalid == 0 ? 0u : (0xffffffffu >> (32 - nvalid));
        }
        const int64_t trel = tgt - c0;                          // the target itself is never a candidate (utils.py:169)
        if (trel >= 0 && trel < 32) keep &= ~(1u << (int)trel);
        while (fnext < c0 + 32) {                               // known-true ids inside this chunk are not candidates either
          if (fnext >= c0) keep &= ~(1u << (int)(fnext - c0));
          ++fpos;
          fnext = fpos < fend ? (int64_t)__ldg(a.filt_ids + fpos) - a.row_offset : INT64_MAX;
        }
        cmask &= keep;
        mask &= keep;
        cnt0 += __popc(cmask);
        if (!qok) mask = 0;
        while (true) {
          const uint32_t have = __ballot_sync(0xffffffffu, mask != 0);
          if (have == 0) break;
          if (wcnt + 32 > TC_WBUF) flush();
          if (mask != 0) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            wbuf[wcnt + __popc(have & lt_mask)] = make_uint2((uint32_t)q, (uint32_t)(c0 + j));
          }
          wcnt += __popc(have);
        }
      };

      tc_mbar_wait(acc_full + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * TC_BN + chalf * 128 + ((uint32_t)(lq * 32) << 16);
      uint32_t rA[32], rB[32];
      tc_ld32_issue(taddr, rA);
      tc_ld_wait(rA);
      tc_ld32_issue(taddr + 32, rB);
      process(rA, 0);
      tc_ld_wait(rB);
      tc_ld32_issue(taddr + 64, rA);
      process(rB, 32);
      tc_ld_wait(rA);
      tc_ld32_issue(taddr + 96, rB);
      process(rA, 64);
      tc_ld_wait(rB);
      // every TMEM read of this accumulator has completed: hand it back to the MMA issuer before the last chunk
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (NCTA == 2 && !leader) tc_mbar_arrive_remote(acc_empty + acc, 0); else tc_mbar_arrive(acc_empty + acc); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      process(rB, 96);
      if (qok && cnt0 != 0) atomicAdd(a.counts + q, cnt0);
    }
    if (wcnt != 0) flush();
  }
  tc_fence_before();
  if (NCTA == 2) tc_cluster_sync(); else __syncthreads();   // pair: nothing of either CTA is still in flight towards the other
  if (warp == 2) {
    tc_fence_after();
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- the approximate mode with the QUERY tile resident in tensor memory -------------------------------------------------
// k_rank_tc is bound by shared-memory traffic in the single-plane mode: every K=16 step the MMA fetches 12 KB of operands from
// shared memory while the bulk copies refill it at the same rate (profiles/r01_tc_v2_summary.md).  A CTA keeps the same 128
// queries for thousands of entity tiles, so here their operand lives in TMEM (columns 256.., written once per query tile with
// tcgen05.st: lane = query, a 32-bit column = two consecutive k -- the layout the A-from-TMEM form of tcgen05.mma reads, checked
// bit for bit by profiles/micro/tc_numerics.cu with TC_A_TMEM=1) and only the entity operand streams: 128-row tiles, 16 KB
// stages, eight of them in flight, N = 128 accumulators double-buffered in columns 0..255.  Shared-memory bytes per flop
// drop by a third on both sides (fill and MMA fetch).  K <= 512 (256 columns of A); wider tables keep k_rank_tc.
// RESULT (profiles/r02_tc_ts.md): correct (every approximate-mode test passes, ranks identical to k_rank_tc's) and SLOWER --
// 22.0 ms against 16.0 ms per side on C5.  Tensor memory has one 64 B/clk read port: an M128 N128 K16 MMA fetches its 4 KB A
// operand from it in 64 clk (all of it), and the epilogue's tcgen05.ld of the 64 KB accumulator tile needs another 32 B/clk.
// The shared-memory operand path of k_rank_tc is the faster one on this chip; the kernel is kept behind KGEB_TC_TS=1.
constexpr int TS_BN = 128;
constexpr int TS_STAGES = 8;
constexpr uint32_t TS_B_STAGE = TS_BN * TC_CH_TOTAL * 16;   // 16 KB: 8 k-chunks = 64 k of 128 entity rows
constexpr uint32_t TS_SMEM_BYTES = TS_STAGES * TS_B_STAGE + 1024 + 256 + 8 * TC_WBUF * 8;
constexpr uint32_t TS_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TS_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void ts_mma_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// work split: the CTAs form a (G query lanes) x (T tile lanes) grid; inside a group of G query tiles CTA (lq, lt) keeps query tile
// g*G + lq and walks entity tiles lt, lt + T, ...: the query operand is reloaded only when the group changes
struct TsWalk {
  int G, T, lq, lt; bool idle;
  __device__ TsWalk(const TcArgs& a) {
    const int P = (int)gridDim.x;
    G = a.qgroup < P ? a.qgroup : P;
    if (G > a.qtiles) G = a.qtiles;
    T = P / G;
    lq = (int)blockIdx.x % G; lt = (int)blockIdx.x / G;
    idle = lt >= T;
  }
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_rank_tc_ts(TcArgs a) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smB = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TS_STAGES * TS_B_STAGE);
  uint64_t* full = bars;                       // [TS_STAGES]
  uint64_t* empty = bars + TS_STAGES;          // [TS_STAGES]
  uint64_t* acc_full = bars + 2 * TS_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* a_ready = acc_empty + 2;           // [1] the 8 epilogue warps have written the query operand of the current group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
  uint2* wbuf_all = reinterpret_cast<uint2*>(smem + TS_STAGES * TS_B_STAGE + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstage = (a.nkc + TC_CH_TOTAL - 1) / TC_CH_TOTAL;
  const TsWalk wk(a);
  const int ngroups = (a.qtiles + wk.G - 1) / wk.G;

  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) a.wl_hdr[1] = a.wl_cap;
    for (int i = 0; i < TS_STAGES; ++i) { tc_mbar_init(full + i, 1); tc_mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { tc_mbar_init(acc_full + i, 1); tc_mbar_init(acc_empty + i, 8); }
    tc_mbar_init(a_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + 256u;

  if (warp == 0) {
    // ===== producer: one thread streams the packed entity stages =====
    if (lane == 0 && !wk.idle) {
      uint32_t stage = 0, phase = 0;
      const size_t b_tile_bytes = (size_t)a.nkc * TS_BN * 16;
      for (int g = 0; g < ngroups; ++g) {
        if (g * wk.G + wk.lq >= a.qtiles) continue;
        for (int64_t tile = wk.lt; tile < a.ntiles; tile += wk.T) {
          const unsigned char* srcB = a.packedE + (size_t)tile * b_tile_bytes;
          for (int s = 0; s < nstage; ++s) {
            const int c0 = s * TC_CH_TOTAL;
            const int nch = a.nkc - c0 < TC_CH_TOTAL ? a.nkc - c0 : TC_CH_TOTAL;
            tc_mbar_wait(empty + stage, phase ^ 1u);
            const uint32_t bb = (uint32_t)nch * TS_BN * 16;
            tc_mbar_expect_tx(full + stage, bb);
            tc_bulk_g2s(smB + stage * TS_B_STAGE, srcB + (size_t)c0 * TS_BN * 16, bb, full + stage);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread =====
    if (lane == 0 && !wk.idle) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_phase = 0;
      for (int g = 0; g < ngroups; ++g) {
        if (g * wk.G + wk.lq >= a.qtiles) continue;
        tc_mbar_wait(a_ready, a_phase);          // the group's query operand is in TMEM (every MMA of the previous group has completed)
        a_phase ^= 1u;
        tc_fence_after();
        for (int64_t tile = wk.lt; tile < a.ntiles; tile += wk.T) {
          tc_mbar_wait(acc_empty + acc, acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * TS_BN;
          uint32_t accumulate = 0;
          // stages are consumed in groups of four: the MMAs of four stages are issued back to back and only then are the stages
          // handed back (every MMA/commit alternation costs the tensor pipe a hand-off; see k_rank_tc)
          for (int s = 0; s < nstage; s += 4) {
            uint32_t st_idx[4];
            int nin = 0;
            for (int h = 0; h < 4 && s + h < nstage; ++h) {
              const int c0 = (s + h) * TC_CH_TOTAL;
              const int nk16 = (a.nkc - c0 < TC_CH_TOTAL ? a.nkc - c0 : TC_CH_TOTAL) >> 1;
              tc_mbar_wait(full + stage, phase);
              tc_fence_after();
              const uint32_t sb = tc_smem_u32(smB + stage * TS_B_STAGE);
              for (int j = 0; j < nk16; ++j) {
                ts_mma_bf16(tmem_d, tmem_a + (uint32_t)(c0 / 2 + j) * 8u, tc_desc(sb + (uint32_t)(2 * j) * TS_BN * 16, TS_BN * 16, 128), TS_IDESC_BF16,
                            accumulate);
                accumulate = 1;
              }
              st_idx[nin++] = stage;
              if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
            }
            for (int h = 0; h < nin; ++h) tc_commit(empty + st_idx[h]);
          }
          tc_commit(acc_full + acc);
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4 && !wk.idle) {
    // ===== query-operand loaders + epilogue: 8 warps; warp w owns TMEM lanes 32*(w%4).. and columns 64*((w-4)/4).. of an accumulator =====
    const int lq = warp & 3, chalf = (warp - 4) >> 2;
    uint2* wbuf = wbuf_all + (warp - 4) * TC_WBUF;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t wcnt = 0;
    uint32_t acc = 0, acc_phase = 0;

    auto flush = [&]() {
      __syncwarp();
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.wl_hdr, wcnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < wcnt; i += 32)
        if (base + i < a.wl_cap) a.wl[base + i] = wbuf[i];
      __syncwarp();
      wcnt = 0;
    };

    for (int g = 0; g < ngroups; ++g) {
      const int64_t qt = (int64_t)g * wk.G + wk.lq;
      if (qt >= a.qtiles) continue;
      // ---- this group's query operand -> TMEM: the previous group's accumulators have all been read by this warp (program order),
      //      and they were complete, so no MMA still reads the old operand.  Warp (lq, chalf) writes its 32 lanes, half of the k range.
      {
        const uint4* src = reinterpret_cast<const uint4*>(a.packedQ + (size_t)qt * a.nkc * TC_BM * 16);
        const int kc0 = chalf * (a.nkc >> 1), kc1 = chalf ? a.nkc : (a.nkc >> 1);
        for (int kc = kc0; kc < kc1; ++kc) {
          const uint4 v = __ldg(src + (size_t)kc * TC_BM + lq * 32 + lane);
          const uint32_t taddr = tmem_a + (uint32_t)kc * 4u + ((uint32_t)(lq * 32) << 16);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(a_ready);
      }
      const int64_t q = qt * TC_BM + lq * 32 + lane;
      const bool qok = q < a.Q;
      const float t = qok ? __ldg(a.tscore + q) : 0.f;
      const int64_t tgt = qok ? (int64_t)__ldg(a.triples + 3 * q + (a.side == KGEB_SIDE_TAIL ? 2 : 0)) - a.row_offset : -1;
      int64_t f_lo = 0, f_end = 0;
      if (a.filt_offs && qok) { f_lo = __ldg(a.filt_offs + q); f_end = __ldg(a.filt_offs + q + 1); }
      int cnt0 = 0;

      for (int64_t tile = wk.lt; tile < a.ntiles; tile += wk.T) {
        const int64_t n0 = tile * TS_BN + chalf * 64;     // shard-local id of this warp's first column
        // cursor into the query's sorted filter list: first id at or beyond this warp's first column (global id space)
        int64_t fpos = f_end;
        if (f_lo < f_end) {
          int64_t lo_i = f_lo, hi_i = f_end;
          const int64_t g0 = a.row_offset + n0;
          while (lo_i < hi_i) {
            const int64_t mid = (lo_i + hi_i) >> 1;
            if ((int64_t)__ldg(a.filt_ids + mid) < g0) lo_i = mid + 1; else hi_i = mid;
          }
          fpos = lo_i;
        }
        int64_t fnext = fpos < f_end ? (int64_t)__ldg(a.filt_ids + fpos) - a.row_offset : INT64_MAX;

        auto process = [&](uint32_t (&r)[32], int jc) {
          uint32_t k0 = 0, k1 = 0, m0 = 0, m1 = 0;
          tc_cmp32<0>(r, t, t, k0, k1, m0, m1);              // approximate mode: no band -- counted above t, "undecided" exactly at t
          uint32_t cmask = k0 | k1, mask = m0 | m1;
          const int64_t c0 = n0 + jc;
          const int64_t left = a.n_rows - c0;
          uint32_t keep = 0xffffffffu;
          if (left < 32) {
            const int nvalid = left < 0 ? 0 : (int)left;
            keep = nvalid == 0 ? 0u : (0xffffffffu >> (32 - nvalid));
          }
          const int64_t trel = tgt - c0;
          if (trel >= 0 && trel < 32) keep &= ~(1u << (int)trel);
          while (fnext < c0 + 32) {
            if (fnext >= c0) keep &= ~(1u << (int)(fnext - c0));
            ++fpos;
            fnext = fpos < f_end ? (int64_t)__ldg(a.filt_ids + fpos) - a.row_offset : INT64_MAX;
          }
          cmask &= keep;
          mask &= keep;
          cnt0 += __popc(cmask);
          if (!qok) mask = 0;
          while (true) {
            const uint32_t have = __ballot_sync(0xffffffffu, mask != 0);
            if (have == 0) break;
            if (wcnt + 32 > TC_WBUF) flush();
            if (mask != 0) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              wbuf[wcnt + __popc(have & lt_mask)] = make_uint2((uint32_t)q, (uint32_t)(c0 + j));
            }
            wcnt += __popc(have);
          }
        };

        tc_mbar_wait(acc_full + acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * TS_BN + chalf * 64 + ((uint32_t)(lq * 32) << 16);
        uint32_t rA[32], rB[32];
        tc_ld32_issue(taddr, rA);
        tc_ld_wait(rA);
        tc_ld32_issue(taddr + 32, rB);
        process(rA, 0);
        tc_ld_wait(rB);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(acc_empty + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        process(rB, 32);
      }
      if (qok && cnt0 != 0) atomicAdd(a.counts + q, cnt0);
    }
    if (wcnt != 0) flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- operand packing ----------------------------------------------------------------------------------------------
// largest |x| of a table as fp32 bits (non-negative floats order like their bit patterns); *out must be zero before the launch
__global__ void k_tc_absmax(const float* __restrict__ src, int64_t rows, int W, int ld, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  const int64_t total = rows * (int64_t)W;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = o / W;
    const uint32_t b = __float_as_uint(__ldg(src + r * (int64_t)ld + (o - r * W))) & 0x7fffffffu;
    m = b > m ? b : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const uint32_t v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m; }
  if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(out, m);
}

// one thread per 16-byte output chunk: dst[tile][kchunk][plane][row][0..7].
//   scale_bits == NULL: bf16, plane 0 = bf16(x)                                  (approximate mode, one plane)
//   scale_bits != NULL: fp16 of y = x * 2^a, plane 0 = fp16(y), plane 1 = fp16(y - plane0)   (exact mode)
//   l2_role (TransE-L2): 1 = entity rows: scale from tc_l2_scales, positions W..W+2 of plane 0 take extras[row][0..2] (-W_e split);
//                        2 = query rows: scale and sign from tc_l2_scales / `negate`, positions W..W+2 take the matching powers of two
struct TcPackArgs {
  const float* src; int64_t rows; int W, ld, tile_rows, nkc, planes; int64_t ntile;
  const uint32_t* scale_bits; const uint32_t* other_bits; const __half* extras; int l2_role, negate;
  uint4* dst;
};
__global__ void k_tc_pack(TcPackArgs a) {
  const int64_t total = a.ntile * a.nkc * a.planes * a.tile_rows;
  float scale = a.scale_bits ? tc_scale_from_bits(__ldg(a.scale_bits)) : 1.f;
  float qx[3] = {0.f, 0.f, 0.f};
  if (a.l2_role == 1) {
    float scQ;
    tc_l2_scales(__ldg(a.scale_bits), 0u, scale, scQ);
  } else if (a.l2_role == 2) {
    float scE;
    tc_l2_scales(__ldg(a.other_bits), __ldg(a.scale_bits), scE, scale);      // scale = scQ (0 for a degenerate batch: empty planes)
    const float ratio = scale / scE;                                          // 2^(aQ - aE) in [2^-17, 1], or 0
    qx[0] = ratio * 32768.f; qx[1] = ratio * 16.f; qx[2] = ratio * 0.0078125f;
    if (a.negate) { scale = -scale; qx[0] = -qx[0]; qx[1] = -qx[1]; qx[2] = -qx[2]; }
  }
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(o % a.tile_rows);
    const int64_t o1 = o / a.tile_rows;
    const int plane = (int)(o1 % a.planes);
    const int64_t o2 = o1 / a.planes;
    const int kc = (int)(o2 % a.nkc);
    const int64_t tile = o2 / a.nkc;
    const int64_t r = tile * a.tile_rows + row;
    const int k0 = kc * 8;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (r < a.rows) {
      if (k0 + 8 <= a.W) {
        const float4 v0 = ldg4(reinterpret_cast<const float4*>(a.src + r * (int64_t)a.ld + k0));
        const float4 v1 = ldg4(reinterpret_cast<const float4*>(a.src + r * (int64_t)a.ld + k0 + 4));
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (k0 + e < a.W) x[e] = __ldg(a.src + r * (int64_t)a.ld + k0 + e);
      }
    }
    unsigned short h[8];
    if (a.scale_bits) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = __fmul_rn(x[e], scale);                  // exact: power of two, |y| < 2^14
        __half hi = __float2half_rn(y);
        if (plane == 1) hi = __float2half_rn(__fsub_rn(y, __half2float(hi)));   // the subtraction is exact in fp32
        h[e] = __half_as_ushort(hi);
      }
      if (a.l2_role && r < a.rows && k0 + 8 > a.W && k0 < a.W + 3) {   // the three extra slots (plane 0 only; plane 1 stays zero)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = k0 + e - a.W;
          if (j >= 0 && j < 3) {
            __half v = __float2half_rn(0.f);
            if (plane == 0) v = a.l2_role == 1 ? a.extras[r * 4 + j] : __float2half_rn(qx[j]);   // powers of two: exact
            h[e] = __half_as_ushort(v);
          }
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = __bfloat16_as_ushort(__float2bfloat16_rn(x[e]));
    }
    uint4 out;
    out.x = h[0] | ((uint32_t)h[1] << 16); out.y = h[2] | ((uint32_t)h[3] << 16);
    out.z = h[4] | ((uint32_t)h[5] << 16); out.w = h[6] | ((uint32_t)h[7] << 16);
    a.dst[o] = out;
  }
}

// TransE-L2 entity rows: extras[r][0..2] = -(hi, mid, lo) parts of W_e = 2^(2a) ||e_r||^2 / 2, divided by 2^15, 2^4, 2^-7 so that each
// fits fp16; the sum of squares runs in double, so W_e is exact to 2^-50 and the three-way split leaves less than 2^-33 W_e
__global__ void k_tc_l2_extras(const float* __restrict__ src, int64_t rows, int64_t rows_padded, int W, int ld,
                               const uint32_t* __restrict__ scale_bits, __half* __restrict__ extras) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float scE, scQ;
  tc_l2_scales(__ldg(scale_bits), 0u, scE, scQ);
  for (int64_t r = warp0; r < rows_padded; r += nwarps) {
    double acc = 0.0;
    if (r < rows)
      for (int k = lane; k < W; k += 32) { const double x = (double)__ldg(src + r * (int64_t)ld + k); acc += x * x; }
    acc = warp_sum_d(acc);
    if (lane == 0) {
      const double We = 0.5 * acc * (double)scE * (double)scE;
      const __half e1 = __float2half_rn((float)(We * (1.0 / 32768.0)));
      const double r1 = We - 32768.0 * (double)__half2float(e1);
      const __half e2 = __float2half_rn((float)(r1 * (1.0 / 16.0)));
      const double r2 = r1 - 16.0 * (double)__half2float(e2);
      const __half e3 = __float2half_rn((float)(r2 * 128.0));
      extras[r * 4 + 0] = __hneg(e1); extras[r * 4 + 1] = __hneg(e2); extras[r * 4 + 2] = __hneg(e3); extras[r * 4 + 3] = __float2half_rn(0.f);
    }
  }
}

// TransE-L2 queries (one warp each): q' = the vector whose distance to the candidate rows is ranked -- qvec itself (tail: fl(s + p))
// or fl(o - p) from the two-part head query [p | o] -- and the per-query constants of the epilogue's L2 branch
__global__ void k_tc_l2_query(const float* __restrict__ qvec, int qw, int W, int side, const float* __restrict__ tscore, int64_t Q,
                              int64_t qpad, float gamma, int sub, float* __restrict__ q2, float* __restrict__ l2q) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double u = 1.0 / 16777216.0;
  for (int64_t q = warp0; q < qpad; q += nwarps) {
    if (q >= Q) {
      if (lane < 6) l2q[lane * qpad + q] = 0.f;
      continue;
    }
    const float* src = qvec + q * (int64_t)qw;
    double n2 = 0.0, p2 = 0.0;
    for (int k = lane; k < W; k += 32) {
      float v;
      if (side == KGEB_SIDE_TAIL) v = src[k];
      else { v = __fsub_rn(src[W + k], src[k]); p2 += (double)src[k] * (double)src[k]; }
      q2[q * (int64_t)W + k] = v;
      n2 += (double)v * (double)v;
    }
    n2 = warp_sum_d(n2);
    p2 = warp_sum_d(p2);
    if (lane == 0) {
      const double t = (double)tscore[q];
      double A_lo, A_hi, Dt_up;
      if (sub) {                       // score = gamma - sqrt(acc)
        const double Dt = (double)gamma - t, Dp = Dt > 0.0 ? Dt : 0.0;
        const double eps = 1.01 * u * (2.0 * Dp + fabs((double)gamma)) + 1e-37;
        A_lo = Dt > eps ? (Dt - eps) * (Dt - eps) : -1.0;
        A_hi = (Dp + eps) * (Dp + eps);
        Dt_up = Dp + eps;
      } else {                         // score = +sqrt(acc): counted from (t (1 + 2u))^2 up, certainly not counted up to (t (1 - 2u))^2
        const double tp = t > 0.0 ? t : 0.0;
        A_lo = tp * (1.0 - 2.0 * u) * tp * (1.0 - 2.0 * u);
        A_hi = tp * (1.0 + 2.0 * u) * tp * (1.0 + 2.0 * u) + 1e-37;
        Dt_up = tp * (1.0 + 2.0 * u);
      }
      l2q[0 * qpad + q] = (float)n2;
      l2q[1 * qpad + q] = A_lo < 0.0 ? -1.f : (float)(A_lo * (1.0 - 1e-6));          // outward: a smaller A_lo, a larger A_hi
      l2q[2 * qpad + q] = (float)(A_hi * (1.0 + 1e-6));
      l2q[3 * qpad + q] = (float)(sqrt(A_hi) * (1.0 + 1e-6));
      l2q[4 * qpad + q] = (float)(Dt_up * (1.0 + 1e-6));
      l2q[5 * qpad + q] = side == KGEB_SIDE_TAIL ? 0.f : (float)(2.1 * u * (sqrt(p2) + sqrt(n2)) * 1.001);
    }
  }
}

// norms[r] = an upper bound of ||src[r,:]||_2 for r < rows; wnorms[r] (optional) = an upper bound of the weighted norm
// sqrt(sum_k w_k x_k^2), w_k = nblk - k/16 (the number of K=16 blocks at or after k's); rows..rows_padded-1 get `pad_value`
__global__ void k_tc_norms(const float* __restrict__ src, int64_t rows, int64_t rows_padded, int W, int ld, int nblk, float pad_value,
                           float* __restrict__ norms, float* __restrict__ wnorms) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows_padded; r += nwarps) {
    if (r >= rows) { if (lane == 0) { norms[r] = pad_value; if (wnorms) wnorms[r] = pad_value; } continue; }
    float acc = 0.f, wacc = 0.f;
    for (int k = lane; k < W; k += 32) {
      const float x = __ldg(src + r * (int64_t)ld + k);
      const float x2 = x * x;
      acc += x2;
      wacc = fmaf((float)(nblk - (k >> 4)), x2, wacc);
    }
    acc = warp_sum(acc);
    wacc = warp_sum(wacc);
    if (lane == 0) {
      norms[r] = sqrtf(acc) * 1.0001f;  // fp32 summation error of <= 2^12 non-negative terms is < 1e-4 relative
      if (wnorms) wnorms[r] = sqrtf(wacc) * 1.0001f;
    }
  }
}

// emax[slice] = max of the row-norm upper bounds over each 128-row slice (0 for an all-padding slice)
__global__ void k_tc_slice_max(const float* __restrict__ norms, int64_t nslices, float* __restrict__ emax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t sl = warp0; sl < nslices; sl += nwarps) {
    float m = 0.f;
    for (int i = lane; i < 128; i += 32) m = fmaxf(m, norms[sl * 128 + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) emax[sl] = m;
  }
}

int tc_planes(int mode) { return mode == KGEB_TC_BF16X3_EXACT ? 2 : 1; }
bool tc_l2_kind(const kgeb_ctx* ctx) { return ctx->kind == KGEB_TRANSE_L2; }
// K padded to 16, in 8-element chunks; the L2 mode appends three slots to every row
int tc_nkc(const kgeb_ctx* ctx) { return ((ctx->we + (tc_l2_kind(ctx) ? 3 : 0) + 15) / 16) * 2; }
int64_t tc_tiles(int64_t rows, int tile_rows) { return (rows + tile_rows - 1) / tile_rows; }
bool tc_kind_ok(const kgeb_ctx* ctx) {
  return ctx->kind == KGEB_DISTMULT || ctx->kind == KGEB_COMPLEX || ctx->kind == KGEB_HOLE || ctx->kind == KGEB_ROTATE || tc_l2_kind(ctx);
}
#define KGEB_TC_KINDS_MSG "tensor-core ranking serves DistMult, ComplEx, HolE, RotatE (dot-product queries) and TransE with the L2 norm"
size_t tc_align(size_t x) { return (x + 255) & ~(size_t)255; }

int tc_grid(const kgeb_ctx* ctx, int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// KGEB_TC_TS=1 (an experiment, OFF by default): approximate mode, K <= 512, with the query operand in TMEM (k_rank_tc_ts) and the
// entity image packed in 128-row tiles.  Measured on C5: 22.0 ms per side against 16.0 ms for k_rank_tc (profiles/r02_tc_ts.md).
bool tc_use_ts(const kgeb_ctx* ctx, int mode) {
  static const bool on = [] { const char* e = getenv("KGEB_TC_TS"); return e && e[0] == '1'; }();
  return on && mode == KGEB_TC_BF16 && !tc_l2_kind(ctx) && tc_nkc(ctx) <= 64;
}
int tc_bn(const kgeb_ctx* ctx, int mode) { return tc_use_ts(ctx, mode) ? TS_BN : TC_BN; }

// packed entity buffer: [operand image][row norms nt*bn][slice max][weighted norms nt*bn][their slice max][scale bits]  (bn = rows per tile)
struct TcPackedLayout { size_t image, norms, smax, wnorms, wmax, scale, extras, total; };
TcPackedLayout tc_packed_layout(const kgeb_ctx* ctx, int64_t n_rows, int mode) {
  const int bn = tc_bn(ctx, mode);
  const int64_t nt = tc_tiles(n_rows, bn), nsl = nt * (bn / 128);
  TcPackedLayout L;
  L.image = 0;
  L.norms = tc_align((size_t)nt * tc_nkc(ctx) * tc_planes(mode) * bn * 16);
  L.smax = L.norms + tc_align((size_t)nt * bn * 4);
  L.wnorms = L.smax + tc_align((size_t)nsl * 4);
  L.wmax = L.wnorms + tc_align((size_t)nt * bn * 4);
  L.scale = L.wmax + tc_align((size_t)nsl * 4);
  L.extras = L.scale + 256;                                                   // L2 mode: 4 halfs per (padded) row
  L.total = L.extras + (tc_l2_kind(ctx) ? tc_align((size_t)nt * bn * 8) : 0);
  return L;
}
// workspace: [header 256 B: appended, capacity, query scale bits at +16][worklist][packed queries][query norms][weighted query norms]
struct TcWsLayout { size_t wl, packedQ, qnorm, qwnorm, q2, l2q, grp, total; };
// The undecided pairs are ordered by entity before the exact re-scoring (wl_group_kernels.cu) when the pass is large enough for the
// four extra launches to pay: on the 1M-entity shape the re-scoring is bound by L2 -> SM bandwidth and an entity occurs in ~12
// pairs.  `rows` = the shard's rows at the call, the table's rows when the workspace is sized (never fewer).
inline bool tc_pairs_grouped(const kgeb_ctx* ctx, int64_t Q, int64_t rows, int mode) {
  if (mode != KGEB_TC_BF16X3_EXACT || ctx->env_pairs_grouped == 0) return false;
  return ctx->env_pairs_grouped == 1 || Q * rows >= ((int64_t)1 << 30);
}
TcWsLayout tc_ws_layout(const kgeb_ctx* ctx, int64_t Q, int mode, int64_t worklist_cap) {
  const int64_t qt = tc_tiles(Q, TC_BM);
  TcWsLayout L;
  L.wl = TC_HDR_BYTES;
  L.packedQ = L.wl + tc_align((size_t)worklist_cap * 8);
  L.qnorm = L.packedQ + tc_align((size_t)qt * tc_nkc(ctx) * tc_planes(mode) * TC_BM * 16);
  L.qwnorm = L.qnorm + tc_align((size_t)qt * TC_BM * 4);
  L.q2 = L.qwnorm + tc_align((size_t)qt * TC_BM * 4);                         // L2 mode: q' [qt*128, W] and six per-query arrays
  L.l2q = L.q2 + (tc_l2_kind(ctx) ? tc_align((size_t)qt * TC_BM * ctx->we * 4) : 0);
  L.grp = L.l2q + (tc_l2_kind(ctx) ? tc_align((size_t)6 * qt * TC_BM * 4) : 0);
  L.total = L.grp + (tc_pairs_grouped(ctx, Q, ctx->N, mode) ? tc_align(kgeb_internal_wl_group_layout(ctx->N, worklist_cap).total) : 0);
  return L;
}

// tensor map over a packed operand image seen as 64-bit words: one row per chunk column (tile_rows x 16 bytes), boxes of 128 rows'
// worth of words (256) x the 8 chunk columns of a pipeline stage, no swizzle: the box lands as the [chunk column][128 rows][16 B]
// image the UMMA descriptors read.  The encoder is host code, looked up once through the runtime.
typedef CUresult (*tc_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool tc_image_map(const void* image, int tile_rows, int64_t columns, CUtensorMap* out) {
  static tc_encode_tiled_fn enc = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return (tc_encode_tiled_fn)fn;
  }();
  if (!enc || columns <= 0 || columns > 0x7fffffffLL) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)tile_rows * 2, (cuuint64_t)columns};
  const cuuint64_t strides[1] = {(cuuint64_t)tile_rows * 16};
  const cuuint32_t box[2] = {256, (cuuint32_t)TC_CH_TOTAL}, estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(image), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int64_t kgeb_rank_tc_packed_bytes(const kgeb_ctx* ctx, int64_t n_rows, int mode) {
  if (!ctx || n_rows < 0 || (mode != KGEB_TC_BF16 && mode != KGEB_TC_BF16X3_EXACT)) return KGEB_ERR_INVALID;
  return (int64_t)tc_packed_layout(ctx, n_rows, mode).total;
}

extern "C" int kgeb_rank_tc_pack_entities(kgeb_ctx* ctx, const float* E_shard, int64_t n_rows, int mode, void* packed,
                                          int64_t packed_bytes, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, tc_kind_ok(ctx), KGEB_ERR_UNSUPPORTED, KGEB_TC_KINDS_MSG);
  KGEB_REQUIRE(ctx, E_shard && packed && n_rows > 0, KGEB_ERR_INVALID, "rank_tc_pack_entities: bad arguments");
  KGEB_REQUIRE(ctx, mode == KGEB_TC_BF16 || mode == KGEB_TC_BF16X3_EXACT, KGEB_ERR_INVALID, "unknown tensor-core mode");
  KGEB_REQUIRE(ctx, ctx->we % 4 == 0 && ctx->ldE % 4 == 0 && kgeb_aligned16(E_shard) && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0,
               KGEB_ERR_ALIGNMENT, "rank_tc needs width, stride multiples of 4 floats, a 16-byte aligned table and a 256-byte aligned packed buffer");
  KGEB_REQUIRE(ctx, packed_bytes >= kgeb_rank_tc_packed_bytes(ctx, n_rows, mode), KGEB_ERR_WORKSPACE, "packed entity buffer too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int bn = tc_bn(ctx, mode);
  const int64_t nt = tc_tiles(n_rows, bn), nsl = nt * (bn / 128);
  const int nkc = tc_nkc(ctx), planes = tc_planes(mode);
  const bool exact = mode == KGEB_TC_BF16X3_EXACT;
  const TcPackedLayout L = tc_packed_layout(ctx, n_rows, mode);
  unsigned char* base = static_cast<unsigned char*>(packed);
  uint32_t* scale_bits = reinterpret_cast<uint32_t*>(base + L.scale);
  const bool l2 = tc_l2_kind(ctx);
  KGEB_REQUIRE(ctx, !l2 || (exact && ctx->we <= 4000), KGEB_ERR_UNSUPPORTED, "TransE-L2 on the tensor cores: exact mode only, width <= 4000");
  __half* extras = reinterpret_cast<__half*>(base + L.extras);
  if (exact) {
    KGEB_CUDA_OK(ctx, cudaMemsetAsync(scale_bits, 0, 256, st));
    k_tc_absmax<<<tc_grid(ctx, n_rows * (int64_t)ctx->we, 256), 256, 0, st>>>(E_shard, n_rows, ctx->we, ctx->ldE, scale_bits);
    KGEB_LAUNCH_OK(ctx);
  }
  if (l2) {
    k_tc_l2_extras<<<tc_grid(ctx, nt * bn * 32, 256), 256, 0, st>>>(E_shard, n_rows, nt * bn, ctx->we, ctx->ldE, scale_bits, extras);
    KGEB_LAUNCH_OK(ctx);
  }
  {
    TcPackArgs pa = {};
    pa.src = E_shard; pa.rows = n_rows; pa.W = ctx->we; pa.ld = ctx->ldE; pa.tile_rows = bn; pa.nkc = nkc; pa.planes = planes; pa.ntile = nt;
    pa.scale_bits = exact ? scale_bits : nullptr; pa.other_bits = nullptr; pa.extras = l2 ? extras : nullptr; pa.l2_role = l2 ? 1 : 0; pa.negate = 0;
    pa.dst = reinterpret_cast<uint4*>(base);
    k_tc_pack<<<tc_grid(ctx, nt * nkc * planes * bn, 256), 256, 0, st>>>(pa);
    KGEB_LAUNCH_OK(ctx);
  }
  float* enorm = reinterpret_cast<float*>(base + L.norms);
  float* ewnorm = reinterpret_cast<float*>(base + L.wnorms);
  k_tc_norms<<<tc_grid(ctx, nt * bn * 32, 256), 256, 0, st>>>(E_shard, n_rows, nt * bn, ctx->we, ctx->ldE, nkc / 2, 0.f, enorm,
                                                                exact ? ewnorm : nullptr);
  KGEB_LAUNCH_OK(ctx);
  k_tc_slice_max<<<tc_grid(ctx, nsl * 32, 256), 256, 0, st>>>(enorm, nsl, reinterpret_cast<float*>(base + L.smax));
  KGEB_LAUNCH_OK(ctx);
  if (exact) {
    k_tc_slice_max<<<tc_grid(ctx, nsl * 32, 256), 256, 0, st>>>(ewnorm, nsl, reinterpret_cast<float*>(base + L.wmax));
    KGEB_LAUNCH_OK(ctx);
  }
  return KGEB_OK;
}

extern "C" int64_t kgeb_rank_tc_workspace_bytes(const kgeb_ctx* ctx, int64_t Q, int mode, int64_t worklist_cap) {
  if (!ctx || Q < 0 || worklist_cap < 0 || (mode != KGEB_TC_BF16 && mode != KGEB_TC_BF16X3_EXACT)) return KGEB_ERR_INVALID;
  return (int64_t)tc_ws_layout(ctx, Q, mode, worklist_cap).total;
}

extern "C" int kgeb_rank_1n_tc(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const void* packedE,
                               const float* qvec, const float* tscore, const int32_t* triples, int64_t Q, int side,
                               const int64_t* filt_offsets, const int32_t* filt_ids, int mode, void* workspace,
                               int64_t workspace_bytes, int64_t worklist_cap, int32_t* counts, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, tc_kind_ok(ctx), KGEB_ERR_UNSUPPORTED, KGEB_TC_KINDS_MSG);
  KGEB_REQUIRE(ctx, E_shard && packedE && qvec && tscore && triples && counts && workspace && Q >= 0 && n_rows >= 0 && row_offset >= 0,
               KGEB_ERR_INVALID, "rank_1n_tc: bad arguments");
  KGEB_REQUIRE(ctx, side == KGEB_SIDE_TAIL || side == KGEB_SIDE_HEAD, KGEB_ERR_INVALID, "unknown side");
  KGEB_REQUIRE(ctx, mode == KGEB_TC_BF16 || mode == KGEB_TC_BF16X3_EXACT, KGEB_ERR_INVALID, "unknown tensor-core mode");
  KGEB_REQUIRE(ctx, (filt_offsets == nullptr) == (filt_ids == nullptr), KGEB_ERR_INVALID, "filter CSR: give both arrays or neither");
  KGEB_REQUIRE(ctx, worklist_cap > 0 && worklist_cap < 0x7fffffffLL, KGEB_ERR_INVALID, "worklist capacity out of range");
  KGEB_REQUIRE(ctx, ctx->we % 4 == 0 && ctx->ldE % 4 == 0 && kgeb_aligned16(E_shard) && kgeb_aligned16(qvec) &&
                        (reinterpret_cast<uintptr_t>(packedE) & 255u) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               KGEB_ERR_ALIGNMENT, "rank_1n_tc needs 16-byte aligned tables and 256-byte aligned packed / workspace buffers");
  KGEB_REQUIRE(ctx, workspace_bytes >= kgeb_rank_tc_workspace_bytes(ctx, Q, mode, worklist_cap), KGEB_ERR_WORKSPACE,
               "rank_1n_tc: workspace smaller than kgeb_rank_tc_workspace_bytes()");
  if (Q == 0 || n_rows == 0) return KGEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nkc = tc_nkc(ctx), planes = tc_planes(mode), qw = kgeb_query_width(ctx, side);
  const bool exact = mode == KGEB_TC_BF16X3_EXACT;
  const bool ts = tc_use_ts(ctx, mode);
  const int64_t qt = tc_tiles(Q, TC_BM), nt = tc_tiles(n_rows, tc_bn(ctx, mode));
  KGEB_REQUIRE(ctx, qt < 0x7fffffffLL && nt < 0x7fffffffLL, KGEB_ERR_UNSUPPORTED, "too many tiles");
  const TcWsLayout W = tc_ws_layout(ctx, Q, mode, worklist_cap);
  const TcPackedLayout L = tc_packed_layout(ctx, n_rows, mode);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(ws);
  uint32_t* qscale_bits = hdr + 4;
  uint2* wl = reinterpret_cast<uint2*>(ws + W.wl);
  unsigned char* packedQ = ws + W.packedQ;
  float* qnorm = reinterpret_cast<float*>(ws + W.qnorm);
  float* qwnorm = reinterpret_cast<float*>(ws + W.qwnorm);

  const bool l2 = tc_l2_kind(ctx);
  KGEB_REQUIRE(ctx, !l2 || exact, KGEB_ERR_UNSUPPORTED, "TransE-L2 on the tensor cores: exact mode only");
  const int64_t qpad = qt * TC_BM;
  float* q2 = reinterpret_cast<float*>(ws + W.q2);
  float* l2q = reinterpret_cast<float*>(ws + W.l2q);
  const float* qsrc = qvec;          // what the operand image, the norms and the scale are taken from
  int qld = qw;
  const int reverse = l2 && !(ctx->gamma > 0.f);
  KGEB_CUDA_OK(ctx, cudaMemsetAsync(hdr, 0, TC_HDR_BYTES, st));
  if (l2) {
    k_tc_l2_query<<<tc_grid(ctx, qpad * 32, 256), 256, 0, st>>>(qvec, qw, ctx->we, side, tscore, Q, qpad, ctx->gamma, ctx->gamma > 0.f ? 1 : 0, q2, l2q);
    KGEB_LAUNCH_OK(ctx);
    qsrc = q2; qld = ctx->we;
  }
  if (exact) {
    k_tc_absmax<<<tc_grid(ctx, Q * (int64_t)ctx->we, 256), 256, 0, st>>>(qsrc, Q, ctx->we, qld, qscale_bits);
    KGEB_LAUNCH_OK(ctx);
  }
  {
    TcPackArgs pa = {};
    pa.src = qsrc; pa.rows = Q; pa.W = ctx->we; pa.ld = qld; pa.tile_rows = TC_BM; pa.nkc = nkc; pa.planes = planes; pa.ntile = qt;
    pa.scale_bits = exact ? qscale_bits : nullptr;
    pa.other_bits = l2 ? reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(packedE) + L.scale) : nullptr;
    pa.extras = nullptr; pa.l2_role = l2 ? 2 : 0; pa.negate = reverse;
    pa.dst = reinterpret_cast<uint4*>(packedQ);
    k_tc_pack<<<tc_grid(ctx, qt * nkc * planes * TC_BM, 256), 256, 0, st>>>(pa);
    KGEB_LAUNCH_OK(ctx);
  }
  k_tc_norms<<<tc_grid(ctx, qt * TC_BM * 32, 256), 256, 0, st>>>(qsrc, Q, qt * TC_BM, ctx->we, qld, nkc / 2, 0.f, qnorm, exact ? qwnorm : nullptr);
  KGEB_LAUNCH_OK(ctx);

  const unsigned char* pe = static_cast<const unsigned char*>(packedE);
  TcArgs a;
  a.packedQ = packedQ;
  a.packedE = pe;
  a.qnorm = qnorm;
  a.qwnorm = exact ? qwnorm : nullptr;
  a.emax = reinterpret_cast<const float*>(pe + L.smax);
  a.ewmax = exact ? reinterpret_cast<const float*>(pe + L.wmax) : nullptr;
  a.escale_bits = exact ? reinterpret_cast<const uint32_t*>(pe + L.scale) : nullptr;
  a.qscale_bits = exact ? qscale_bits : nullptr;
  a.tscore = tscore; a.triples = triples; a.counts = counts;
  a.filt_offs = filt_offsets; a.filt_ids = filt_ids;   // applied inside the epilogue as a mask: no correction pass
  a.wl_hdr = hdr; a.wl = wl; a.wl_cap = (uint32_t)worklist_cap;
  a.Q = Q; a.row_offset = row_offset; a.n_rows = n_rows;
  a.nkc = nkc; a.planes = planes; a.qtiles = (int)qt; a.ntiles = (int)nt; a.side = side;
  {
    // concurrent CTAs (one per SM) should cover a few entity tiles, not one and not all: groups of ~num_sms/4 query tiles
    const char* env = getenv("KGEB_TC_QGROUP");
    int64_t g = env ? atoll(env) : (ctx->num_sms + 3) / 4;
    if (g < 1) g = 1;
    if (g > qt) g = qt;
    a.qgroup = (int)g;
  }
  if (exact) {
    // the proven band of the header comment; 1.02 absorbs the 1.0001 of each norm bound and the fp32 rounding of the
    // half-dozen operations that form the thresholds in the epilogue
    const double u = 1.0 / 16777216.0;
    a.idesc = TC_IDESC_F16;
    a.c_norm = (float)(1.02 * 13.0 * u);
    a.c_wnorm = (float)(1.02 * 52.11 * u);
    a.c_abs = (float)(1.02 * sqrt((double)nkc * 8.0) / 33554432.0);
    a.t_slack = 2.4e-7f;
    if (l2) {   // no sequential-dot term: the reference's own rounding lives in eta and c (acc space); Z grows by 2 V, T by V
      a.c_norm = (float)(1.02 * 12.004 * u);
      a.c_wnorm = (float)(1.02 * 36.1 * u);
    }
  } else {
    a.idesc = TC_IDESC_BF16;
    a.c_norm = a.c_wnorm = a.c_abs = 0.f;
    a.t_slack = 0.f;
  }
  a.l2 = l2 ? 1 : 0; a.l2_reverse = reverse; a.l2q = l2q; a.l2_stride = qpad;
  {
    const double u = 1.0 / 16777216.0;
    a.l2_eta = (float)((ctx->we + 8) * u * 1.001);
    a.l2_c_e = side == KGEB_SIDE_HEAD ? (float)(2.1 * u * 1.001) : 0.f;
    // V = ||e||^2 / 2 <= nmax^2 / 2 enters Z twice (the slots may straddle the last two blocks) and the split of W_e once
    a.l2_c_v = (float)(1.02 * (2.0 * 36.1 * u + 1.0 / 2147483648.0) * 0.5);
  }
  CUtensorMap tmQ, tmE;
  memset(&tmQ, 0, sizeof(tmQ)); memset(&tmE, 0, sizeof(tmE));
  static bool attr_done_dev[64] = {};  // the attribute is per function AND per device (several devices in one process: each sets it once)
  bool& attr_done = attr_done_dev[ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0];
  if (!attr_done || ctx->device >= 64) {
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_rank_tc_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_rank_tc_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    KGEB_CUDA_OK(ctx, cudaFuncSetAttribute(k_rank_tc_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM_BYTES));
    attr_done = true;
  }
  const int64_t items = qt * nt;
  const int grid = (int)(items < ctx->num_sms ? items : ctx->num_sms);
  // CTA pairs (cta_group::2: one MMA over 256 queries x 256 entities, half of the entity tile staged per CTA) are the default: 4-5 % faster than
  // one CTA per tile in both modes on C5 (profiles/r02_tc_pair.md); KGEB_TC_2CTA=0 selects one CTA per tile
  static const bool pair = [] { const char* e = getenv("KGEB_TC_2CTA"); return !(e && e[0] == '0'); }();
  a.qtiles_real = (int)qt;
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_RANK_TC, st);
  if (ts) k_rank_tc_ts<<<grid, TC_THREADS, TS_SMEM_BYTES, st>>>(a);
  else if (pair && ctx->num_sms >= 2 && tc_image_map(packedQ, TC_BM, qt * nkc * planes, &tmQ) && tc_image_map(pe, TC_BN, nt * nkc * planes, &tmE)) {
    const int64_t qpairs = (qt + 1) / 2;
    a.qtiles = (int)qpairs;
    if (a.qgroup > qpairs) a.qgroup = (int)qpairs;
    const int64_t units = qpairs * nt, max_pairs = ctx->num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (units < max_pairs ? units : max_pairs)));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    KGEB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, k_rank_tc_t<2>, a, tmQ, tmE));
  } else k_rank_tc_t<1><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a, tmQ, tmE);
  KGEB_TIME_END(ctx, KGEB_TIMED_RANK_TC, st);
  KGEB_LAUNCH_OK(ctx);
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_RANK_REFINE, st);
  // hdr[0] = pairs appended, hdr[1] = capacity (written by the kernel): the host checks hdr[0] <= hdr[1] after the pass
  int rc;
  if (tc_pairs_grouped(ctx, Q, n_rows, mode)) {
    const uint2* ordered = nullptr;
    const uint32_t* length = nullptr;
    rc = kgeb_internal_wl_group(ctx, hdr, (uint32_t)worklist_cap, wl, n_rows, ws + W.grp, &ordered, &length, st);
    if (rc == KGEB_OK)
      rc = kgeb_internal_refine_worklist(ctx, E_shard, n_rows, qvec, tscore, Q, side, length, (uint32_t)worklist_cap, ordered, counts, st, 1);
  } else {
    rc = kgeb_internal_refine_worklist(ctx, E_shard, n_rows, qvec, tscore, Q, side, hdr, (uint32_t)worklist_cap, wl, counts, st, 0);
  }
  KGEB_TIME_END(ctx, KGEB_TIMED_RANK_REFINE, st);
  return rc;
}
