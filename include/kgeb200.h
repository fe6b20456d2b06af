/* kgeb200.h -- C ABI of libkgeb200.so: the B200 (sm_100a) triple-scoring hot path of KGE-Keras.
 *
 * The reference (NIVA-Knowledge-Graph/KGE-Keras, paths below relative to its root) has NO native/FFI seam:
 * its library layer calls stock Keras/TF ops.  This header is therefore the seam a maintainer would add
 * under the unchanged Python signatures; each entry point cites the reference code
 * whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C, no torch / TF types; every pointer named `const float* E`, `idx`, ... is a DEVICE pointer owned
 *     by the caller (the host framework owns the parameter tables, so get_weights()/init_entities keep working);
 *   - every call enqueues on the caller's `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; nothing allocates after kgeb_create;
 *   - return value 0 = ok, negative = kgeb_status; kgeb_last_error() gives the text; nothing throws;
 *   - a context is bound to one device and one model shape; contexts are independent, a single context is
 *     not thread-safe.  It owns a few MB of device scratch (loss partial sums, the sorted chunks of the all-pairs
 *     hinge loss, the split-K slices of kgeb_sgemm, HAKE's per-row sums): calls on ONE context must therefore be
 *     ordered on one stream (or by events); use one context per concurrent stream;
 *   - tables are row-major fp32, row stride `ld` floats (ld >= width).  The vectorised kernels need
 *     ld % 4 == 0 and 16-byte aligned base pointers; anything else returns KGEB_ERR_ALIGNMENT (never a slow
 *     silent path);
 *   - ids are int32 [m,3] row-major (s, p, o) as in EmbeddingModel.call (KGEkeras/models.py:140).
 */
#ifndef KGEB200_H
#define KGEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGEB_VERSION 110 /* 0.1.1: fp16 exact ranking planes, kgeb_comm_* */
#if defined(__GNUC__)
#define KGEB_API __attribute__((visibility("default")))
#else
#define KGEB_API
#endif

typedef struct kgeb_ctx kgeb_ctx;
typedef void* kgeb_stream; /* cudaStream_t */

typedef enum {
  KGEB_OK = 0,
  KGEB_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, unknown enum) */
  KGEB_ERR_CUDA = -2,      /* a CUDA runtime call failed; text in kgeb_last_error */
  KGEB_ERR_ALIGNMENT = -3, /* pointer / stride not usable by the vectorised kernels */
  KGEB_ERR_UNSUPPORTED = -4, /* valid but not implemented combination (e.g. row wider than the kernels handle) */
  KGEB_ERR_WORKSPACE = -5  /* caller workspace too small */
} kgeb_status;

/* score functions: DistMult models.py:171-172, TransE 186-190 (norm=1|2), ComplEx 212-222, HolE 231-237;
 * sibling models: RotatE 439-459 (entity rows [re | im] of width 2d, relation rows = d phases, so
 * we == 2*wr and the kernels need ldR >= we), pRotatE 479-489 (score scaled by kgeb_set_score_param's modulus),
 * ModE 417-418 (gamma - ||s*p - o||_norm, norm 1 | 2).  The siblings are served by score_fwd / score_bwd and the
 * CUDA-core ranking kernels (RotatE also by the tensor-core one: its score is linear in the candidate row); the fused
 * train kernel serves the first four kinds. */
typedef enum {
  KGEB_DISTMULT = 0, KGEB_TRANSE_L1 = 1, KGEB_TRANSE_L2 = 2, KGEB_COMPLEX = 3, KGEB_HOLE = 4,
  KGEB_ROTATE = 5, KGEB_PROTATE = 6, KGEB_MODE_L1 = 7, KGEB_MODE_L2 = 8,
  KGEB_HAKE = 9 /* models.py:383-400; score_fwd / score_bwd only, see kgeb_set_hake_params */
} kgeb_score_kind;

/* losses utils.py:288-319.  POINTWISE_BCE = pointwize_cross_entropy (utils.py:297-298): Keras binary_crossentropy on the raw
 * scores, i.e. -log(clip(t) + 1e-7) for positives and -log(1 - clip(f) + 1e-7) for negatives, clip to [1e-7, 1 - 1e-7]. */
typedef enum {
  KGEB_POINTWISE_HINGE = 0, KGEB_POINTWISE_LOGISTIC = 1, KGEB_POINTWISE_SQUARE = 2,
  KGEB_PAIRWISE_HINGE = 3, KGEB_PAIRWISE_LOGISTIC = 4, KGEB_PAIRWISE_SQUARE = 5,
  KGEB_POINTWISE_BCE = 6
} kgeb_loss_kind;
#define KGEB_LOSS_KIND_MAX KGEB_POINTWISE_BCE
#define KGEB_IS_PAIRWISE(k) ((k) >= KGEB_PAIRWISE_HINGE && (k) <= KGEB_PAIRWISE_SQUARE)

/* how pairwise losses pair true with false scores:
 * PAIRED   true (B,1) vs false (B,K)            -- the per-positive pairwise loss
 * ALLPAIRS true (m,)  vs false (m',1) -> (m',m) -- what tests.py:81 computes through broadcasting */
typedef enum { KGEB_PAIRED = 0, KGEB_ALLPAIRS = 1 } kgeb_pair_mode;

/* optimizers (tests.py:86 uses Keras 'adam'); ADAM_DENSE = Keras OptimizerV2 sparse-apply semantics
 * (m, v decayed on every row), LAZY_ADAM = touched rows only */
typedef enum { KGEB_SGD = 0, KGEB_ADAGRAD = 1, KGEB_LAZY_ADAM = 2, KGEB_ADAM_DENSE = 3 } kgeb_opt_kind;

/* which slot(s) the Philox sampler corrupts; BOTH is what utils.py:117-120 does */
typedef enum { KGEB_CORRUPT_BOTH = 0, KGEB_CORRUPT_HEAD = 1, KGEB_CORRUPT_TAIL = 2, KGEB_CORRUPT_EITHER = 3 } kgeb_corrupt_mode;

typedef enum { KGEB_SIDE_TAIL = 0, KGEB_SIDE_HEAD = 1 } kgeb_side;

/* ---- lifecycle -------------------------------------------------------------------------------------------- */
KGEB_API int kgeb_version(void);

/* One context per (device, model).  we/wr = STORED row widths in floats (ComplEx: 2*dim), ldE/ldR = row strides.
 * Replaces EmbeddingModel.__init__'s table allocation contract, models.py:84-112 (tables stay caller-owned). */
KGEB_API int kgeb_create(kgeb_ctx** out, int device, int64_t num_entities, int64_t num_relations, int we, int wr, int ldE,
                int ldR, int score_kind, float gamma);
KGEB_API int kgeb_destroy(kgeb_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL: last error of a failed kgeb_create). */
KGEB_API const char* kgeb_last_error(const kgeb_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
KGEB_API int64_t kgeb_launch_count(const kgeb_ctx* ctx);
KGEB_API int kgeb_num_sms(const kgeb_ctx* ctx);
/* Measurement aid for bench.py's roofline: with timing on, the library brackets the named kernels with CUDA events on
 * the caller's stream; kgeb_last_kernel_ms waits for the closing event and returns the device time of the last launch. */
typedef enum { KGEB_TIMED_RANK_TC = 0, KGEB_TIMED_RANK_REFINE = 1, KGEB_TIMED_RANK_TILE = 2, KGEB_TIMED_TRAIN = 3,
               KGEB_TIMED_OPT = 4, KGEB_TIMED_AUX = 5 } kgeb_timed_kernel;
KGEB_API int kgeb_set_kernel_timing(kgeb_ctx* ctx, int on);
KGEB_API int kgeb_last_kernel_ms(kgeb_ctx* ctx, int which, float* ms);
/* Sum of the device times of the (up to 64 most recent) launches of that kernel since the previous call, and how many
 * launches that was; resets the count.  Lets a timed region of several passes report the kernel's average duration. */
KGEB_API int kgeb_kernel_ms_sum(kgeb_ctx* ctx, int which, float* ms_sum, int* launches);

/* Scalar parameter of the score function besides gamma: pRotatE's modulus (models.py:476, modulus*embedding_range).
 * Default 1; ignored by the other kinds. */
KGEB_API int kgeb_set_score_param(kgeb_ctx* ctx, float param);

/* HolE arithmetic.  KGEB_HOLE_DIRECT (default): the O(d^2) circular correlation accumulated sequentially in fp32 (the contract the
 * bit-exact tests pin).  KGEB_HOLE_FFT: the reference's own formulation (models.py:233-234, ifft(conj(fft(s)) * fft(o))) with
 * shared-memory radix-4 FFTs, for power-of-two widths 32..1024 (other widths keep the direct kernels); same results to fp32
 * FFT rounding (1e-6 relative to the summed magnitudes), ~3 length-d FFTs per triple instead of 3 d^2 multiply-adds.
 * Applies to kgeb_score_fwd / kgeb_score_bwd.  Environment KGEB_HOLE=fft selects it at kgeb_create. */
typedef enum { KGEB_HOLE_DIRECT = 0, KGEB_HOLE_FFT = 1 } kgeb_hole_mode;
KGEB_API int kgeb_set_hole_mode(kgeb_ctx* ctx, int mode);

/* HAKE (models.py:362-400): entity rows [phase | modulus] (we = 2d), relation rows [phase | modulus | bias] (wr = 3d).
 * phase_div = embedding_range / pi -- the DIVISOR of models.py:390-392 (the phases are divided by it in fp32, op for op) --,
 * embedding_range = (gamma + epsilon) / we / 2 (models.py:381, we already doubled);
 * phase_weight, mod_weight: models.py:377-378.  The modulus term is ONE Frobenius norm over all m rows of a kgeb_score_fwd
 * call (tf.norm without axis, models.py:397): scores and gradients of a call are coupled, exactly as in the reference's
 * func(); kgeb_score_bwd therefore expects the same idx (and dropout state) as the forward call it differentiates.
 * No fused training kernel and no 1-N ranking (a candidate's score would depend on its batch).  m <= 2 M rows per call. */
KGEB_API int kgeb_set_hake_params(kgeb_ctx* ctx, float phase_div, float phase_weight, float mod_weight);
/* HAKE inference as Keras runs it: Model.predict scores its input in consecutive batches of `rows` triples (batch_size,
 * default 32), each batch its own call of func() with its own norm (models.py:397).  rows = 0 (default): the whole
 * kgeb_score_fwd call is one batch.  Forward only. */
KGEB_API int kgeb_set_hake_call_rows(kgeb_ctx* ctx, int64_t rows);

/* ---- dropout, models.py:122,144-150: Dropout(dp) on each gathered s / p / o row while training ----------------------
 * p > 0 arms it for the following score_fwd / score_bwd / train_fwd_bwd calls of this context (p = 0 disarms: predict,
 * validate).  Masks are counter-based (Philox4x32-10 keyed by seed, offset = step): forward and backward regenerate
 * the same mask; 16 bits per element, so the effective rate is round(p*65536)/65536.  stream_id separates the masks of
 * the model's positive call (0) from its negative call (1) for score_fwd/score_bwd; train_fwd_bwd uses both itself. */
KGEB_API int kgeb_set_dropout(kgeb_ctx* ctx, float p, uint64_t seed, uint64_t offset, int stream_id);

/* ---- scoring: EmbeddingModel.call + func, models.py:134-162 ------------------------------------------------------- */
/* scores[m] = func(E[idx[:,0]], R[idx[:,1]], E[idx[:,2]]) */
KGEB_API int kgeb_score_fwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, float* scores,
                   kgeb_stream stream);
/* Registered-gradient half of score_fwd: gE/gR += dscore[m] * d score / d row, touched flags set to 1.
 * Replaces TF autodiff through Embedding (IndexedSlices) for this path, SURVEY 8a row a19. */
KGEB_API int kgeb_score_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* idx, int64_t m, const float* dscore,
                   float* gE, float* gR, uint8_t* touchedE, uint8_t* touchedR, kgeb_stream stream);

/* ---- negatives: generate_negative utils.py:114-121 re-done as a counter-based device sampler -------------- */
/* neg[(i*K+j),:] = corrupt(pos[i]) with Philox4x32-10(counter=(i*K+j, offset), key=seed); relation kept. */
KGEB_API int kgeb_negatives(kgeb_ctx* ctx, const int32_t* pos, int64_t B, int K, uint64_t seed, uint64_t offset,
                   int corrupt_mode, int32_t* neg, kgeb_stream stream);

/* ---- oversampling: oversample_data utils.py:123-147 (random.choices with replacement) as a device sampler ----------
 * dst[i,:] = src[u_i,:], u_i = floor(Philox4x32-10(counter=(i, offset), key=seed^'OVRS', stream_id).x * n / 2^32), i < m.
 * The harness (tests.py:55-61) resamples the positive list (stream 0) and the negative list (stream 1) to a common length. */
KGEB_API int kgeb_oversample(kgeb_ctx* ctx, const int32_t* src, int64_t n, int64_t m, uint64_t seed, uint64_t offset, int stream_id,
                    int32_t* dst, kgeb_stream stream);

/* ---- losses on score vectors: utils.py:288-319 ------------------------------------------------------------ */
/* loss[0] = loss_weight * L(t, f); dt[B], df[Bf] = loss_weight * dL/dt, dL/df (either may be NULL).
 * Bf must be B*K for PAIRED pairwise kinds. */
KGEB_API int kgeb_loss_fwd_bwd(kgeb_ctx* ctx, const float* t, int64_t B, const float* f, int64_t Bf, int K, int loss_kind,
                      int pair_mode, float margin, float loss_weight, float* loss, float* dt, float* df,
                      kgeb_stream stream);

/* ---- sort-segmented entity backward (K2) --------------------------------------------------------------------
 * With a workspace armed, kgeb_train_fwd_bwd forms the ENTITY gradient rows by sorting (entity id, reference) records
 * (CUB radix sort on the id bits, stable) and adding each segment in registers -- one plain read-modify-write per
 * touched row, a reproducible summation order -- instead of one vector reduction per row reference.  Relation rows
 * (few rows, thousands of references each) keep the in-register group sums.  Replaces the IndexedSlices de-duplication
 * the reference's optimizer performs (tests.py:86, `adam` on embedding gradients).  workspace == NULL disarms it.
 * Not available together with dropout (the masks are not regenerated in the segment kernel). */
KGEB_API int64_t kgeb_sortseg_workspace_bytes(const kgeb_ctx* ctx, int64_t B, int K);
KGEB_API int kgeb_set_grad_sortseg(kgeb_ctx* ctx, void* workspace, int64_t workspace_bytes);

/* ---- fused train step forward + gradient (K1): gather, negatives, score, loss, gradient scatter ----------- */
/* One pass over B positives, each with K negatives: negatives come from `neg` ([B*K,3], index-driven parity
 * mode) or, when neg == NULL, from the Philox sampler above.  PAIRED pairing for pairwise losses.
 * Outputs: loss[0]; optional pos_scores[B], neg_scores[B*K], neg_out[B*K,3]; gradients are ADDED into gE/gR
 * (dense, same shape/stride as the tables) and touched flags set.  HolE and ALLPAIRS are served by the
 * unfused sequence score_fwd -> loss_fwd_bwd -> score_bwd. */
KGEB_API int kgeb_train_fwd_bwd(kgeb_ctx* ctx, const float* E, const float* R, const int32_t* pos, int64_t B,
                       const int32_t* neg, int K, int loss_kind, float margin, float loss_weight, int corrupt_mode,
                       uint64_t seed, uint64_t offset, float* loss, float* pos_scores, float* neg_scores,
                       int32_t* neg_out, float* gE, float* gR, uint8_t* touchedE, uint8_t* touchedR,
                       kgeb_stream stream);

/* ---- filter sets of validate(), utils.py:201-205 + the per-query list.remove of utils.py:169-171,185-187 -----------------
 * One sorted, de-duplicated CSR per side: ids[offsets[q] .. offsets[q+1]) = every known-true candidate of query q
 * (tail: all o with (s_q, p_q, o) in filt; head: all s with (s, p_q, o_q) in filt), ascending -- the layout kgeb_rank_1n /
 * kgeb_rank_1n_tc take.  Two steps because the caller owns the ids buffer:
 *   count: packs, radix-sorts and de-duplicates the F filtering triples into `workspace`, writes offsets[0..Q] (offsets[Q] = nnz);
 *   fill : (after reading nnz and allocating ids) copies the segments; *target_hit |= 1 if some query's own target is in its
 *          segment (the reference's list.remove raises there: the caller decides).
 * queries, filt: int32 [.,3] device; offsets int64 [Q+1], ids int32 [nnz], target_hit int32 [1] (zeroed by the caller) device.
 * workspace: device, kgeb_filter_csr_workspace bytes, untouched between count and fill. */
KGEB_API int kgeb_filter_csr_workspace(kgeb_ctx* ctx, int64_t Q, int64_t F, int64_t* bytes);
KGEB_API int kgeb_filter_csr_count(kgeb_ctx* ctx, const int32_t* queries, int64_t Q, const int32_t* filt, int64_t F, int side,
                          int64_t* offsets, void* workspace, int64_t workspace_bytes, kgeb_stream stream);
KGEB_API int kgeb_filter_csr_fill(kgeb_ctx* ctx, const int32_t* queries, int64_t Q, int64_t F, int side, const int64_t* offsets,
                         int32_t* ids, int32_t* target_hit, const void* workspace, kgeb_stream stream);

/* ---- l3_reg, models.py:10-11,79-82: loss[0] += w * sum |P|^3 ; G += 3 w P |P| on EVERY row (dense), all flags set ------ */
KGEB_API int kgeb_l3_reg(kgeb_ctx* ctx, const float* P, int64_t rows, int width, int ld, float w, float* G,
                         uint8_t* touched, float* loss, kgeb_stream stream);

/* ---- sparse optimizer (K2 second half): applies the update to touched rows, zeroes G and the flags -------- */
/* P, G: [rows, ld]; state0/state1: Adam m,v (or Adagrad accumulator in state0; unused -> NULL). step >= 1. */
KGEB_API int kgeb_opt_update(kgeb_ctx* ctx, float* P, float* G, uint8_t* touched, int64_t rows, int width, int ld,
                    float* state0, float* state1, int opt_kind, float lr, float beta1, float beta2, float eps,
                    int64_t step, kgeb_stream stream);

/* The same update on two tables (entity + relation) in one launch: the relation table of the bundled datasets has 11-237
 * rows, so its own launch is pure latency. */
KGEB_API int kgeb_opt_update2(kgeb_ctx* ctx, float* P0, float* G0, uint8_t* touched0, int64_t rows0, int width0, int ld0,
                     float* state00, float* state01, float* P1, float* G1, uint8_t* touched1, int64_t rows1, int width1,
                     int ld1, float* state10, float* state11, int opt_kind, float lr, float beta1, float beta2, float eps,
                     int64_t step, kgeb_stream stream);

/* ---- 1-N filtered ranking: validate / gen_tail_data / gen_head_data, utils.py:165-264 --------------------- */
/* Width in floats of one query vector for `side` (we, or 2*we for TransE head queries). */
KGEB_API int kgeb_query_width(const kgeb_ctx* ctx, int side);
/* qvec[q,:] for the queries whose source entity row lies in [row_offset, row_offset+n_rows) of the logical
 * table (E_shard points at row `row_offset`); other rows of qvec are left untouched (zero them first; with a
 * row-sharded table all-reduce(sum) qvec afterwards: exactly one rank contributes each row). */
KGEB_API int kgeb_query_prep(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* R,
                    const int32_t* triples, int64_t Q, int side, float* qvec, kgeb_stream stream);
/* tscore[q] = score of the query's own target entity, computed with the SAME arithmetic as the candidates
 * (so a candidate that ties the target is never counted).  Written only for targets inside the shard. */
KGEB_API int kgeb_target_scores(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                       const int32_t* triples, int64_t Q, int side, float* tscore, kgeb_stream stream);
/* count[0] += #{i != target : scores[i] > scores[target]} -- the rank of a score vector that came from a model the library
 * does not score itself (validate() driven through a foreign model's predict, utils.py:211-215: rankdata(res,'max') with the
 * target in row 0 gives rank = 1 + that count; ties are not counted). */
KGEB_API int kgeb_count_greater(kgeb_ctx* ctx, const float* scores, int64_t n, int64_t target, int32_t* count, kgeb_stream stream);

/* counts[q] += #{c in shard, c not in filter(q), c != target(q) : score(q,c) > tscore[q]}   (rank = 1 + sum over
 * shards).  filt_offsets int64[Q+1] / filt_ids int32[nnz] = CSR of known-true ids per query (may be
 * NULL/NULL).  Integer partial counts: summing them over any row partition gives bit-identical ranks. */
KGEB_API int kgeb_rank_1n(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const float* qvec,
                 const float* tscore, const int32_t* triples, int64_t Q, int side, const int64_t* filt_offsets,
                 const int32_t* filt_ids, int32_t* counts, kgeb_stream stream);

/* ---- 1-N ranking on the tensor cores (tcgen05) for the dot-product scores: DistMult, ComplEx, HolE ------------------
 * Same result contract as kgeb_rank_1n.  KGEB_TC_F16X3_EXACT: 16-bit hi/lo split contraction (three MMAs per k-step on
 * power-of-two scaled fp16 planes) used as a filter with a proven error bound + exact fp32 re-scoring of the undecided
 * pairs -> integer counts IDENTICAL to kgeb_rank_1n (rank_tc_kernels.cu states the bound; the tensor-core accumulation
 * model it rests on is measured by profiles/micro/tc_numerics.cu).  KGEB_TC_BF16X3_EXACT is the first build's name of the
 * same mode (its planes were bf16) and is kept as an alias.
 * KGEB_TC_BF16: one bf16 plane, no bound -> approximate (ranks may move at near-ties).  Replaces the per-query
 * model.predict + rankdata loop of validate, utils.py:210-245. */
typedef enum { KGEB_TC_BF16 = 0, KGEB_TC_BF16X3_EXACT = 1, KGEB_TC_F16X3_EXACT = 1 } kgeb_tc_mode;
/* (the filter CSR is applied inside the kernel's epilogue as a mask: ids must be sorted ascending per query)
 * bytes of the packed 16-bit image (+ row norms, slice maxima, scale) of an entity shard of n_rows rows; 256-byte aligned */
KGEB_API int64_t kgeb_rank_tc_packed_bytes(const kgeb_ctx* ctx, int64_t n_rows, int mode);
/* packs E_shard[n_rows, ldE] once per table version (the table is constant during validate) */
KGEB_API int kgeb_rank_tc_pack_entities(kgeb_ctx* ctx, const float* E_shard, int64_t n_rows, int mode, void* packed,
                               int64_t packed_bytes, kgeb_stream stream);
/* workspace for Q queries (256-byte aligned): 256-byte header {uint32 appended, uint32 capacity, .., uint32 query scale at
 * byte 16}, worklist of capacity (q,c) pairs, packed queries, query norms and -- in the exact mode, when Q x num_entities >= 2^30 or
 * KGEB_PAIRS_GROUPED=1 was set when the context was created -- the scratch of the on-device ordering of the worklist by entity
 * (bins, slots, the ordered copy: 12 bytes per worklist entry + 4 bytes per entity).  After the stream has run the call,
 * appended > capacity means the worklist overflowed: counts are then incomplete and the call must be repeated with a larger
 * worklist_cap. */
KGEB_API int64_t kgeb_rank_tc_workspace_bytes(const kgeb_ctx* ctx, int64_t Q, int mode, int64_t worklist_cap);
KGEB_API int kgeb_rank_1n_tc(kgeb_ctx* ctx, const float* E_shard, int64_t row_offset, int64_t n_rows, const void* packedE,
                    const float* qvec, const float* tscore, const int32_t* triples, int64_t Q, int side,
                    const int64_t* filt_offsets, const int32_t* filt_ids, int mode, void* workspace,
                    int64_t workspace_bytes, int64_t worklist_cap, int32_t* counts, kgeb_stream stream);

/* ---- row exchange of the row-sharded training path (the 1M-entity config): the all-to-all itself is NCCL's */
/* dst[i, :width] = P[idx[i], :width]  (rows another rank requested, packed contiguously for the send) */
KGEB_API int kgeb_rows_gather(kgeb_ctx* ctx, const float* P, int width, int ld, const int32_t* idx, int64_t n, float* dst,
                     kgeb_stream stream);
/* G[idx[i], :width] += src[i, :width]; touched[idx[i]] = 1 (gradient rows returned to their owner; duplicates add up) */
KGEB_API int kgeb_rows_scatter_add(kgeb_ctx* ctx, float* G, int width, int ld, uint8_t* touched, const int32_t* idx, int64_t n,
                          const float* src, kgeb_stream stream);

/* ---- one whole step, sort-segmented, with the optimizer applied at segment end (north-star item 2) -------------
 * Forward as kgeb_train_fwd_bwd (Philox negatives, scores, loss, relation gradient by vector reductions), then (entity id, triple,
 * slot) records are radix-sorted by id, one warp per segment re-forms and sums the row's gradient in registers in sorted order
 * and applies SGD / Adagrad / LazyAdam right there: the updated row is written to E_next (the kernel reads E_cur: rows of other
 * entities are still being read), the moments in place.  No entity gradient buffer is touched.  The caller swaps E_cur / E_next
 * and touched_now / touched_prev after every call; both table copies must start identical, both flag arrays zero.  Replaces
 * tests.py:86 (`adam` on IndexedSlices: dedup by id, then update) for tables far larger than the L2; on L2-resident tables the
 * vector-reduction path + kgeb_opt_update2 is faster (profiles/r02_sortseg.jsonl).  Needs kgeb_set_grad_sortseg armed. */
KGEB_API int kgeb_train_step_sortseg(kgeb_ctx* ctx, const float* E_cur, float* E_next, float* R, const int32_t* pos, int64_t B,
                            const int32_t* neg, int K, int loss_kind, float margin, float loss_weight, int corrupt_mode,
                            uint64_t seed, uint64_t offset, float* loss, float* gR, uint8_t* touchedR, uint8_t* touched_now,
                            uint8_t* touched_prev, float* mE, float* vE, float* mR, float* vR, int opt_kind, float lr,
                            float beta1, float beta2, float eps, int64_t step, kgeb_stream stream);

/* ---- the split form of the backward: (unique_ids, grad_rows) for a host that keeps its own optimizer ---------------
 * A tf.keras Embedding gives its optimizer IndexedSlices (models.py:89-112, tests.py:86); a registered gradient of the TF op shim
 * returns exactly this pair.  Compacts the touched rows of a dense gradient buffer filled by kgeb_train_fwd_bwd / kgeb_score_bwd:
 * ids_out[i] ascending, one per touched row, rows_out[i, :width] = G[ids_out[i], :width]; *count_out (device int64) = number of
 * touched rows (entries beyond `cap` are counted but not written: size the outputs with cap = n_rows, or repeat).  clear != 0
 * zeroes the consumed buffer rows and flags, leaving the buffer ready for the next step (what kgeb_opt_update does itself). */
KGEB_API int kgeb_grad_rows_export(kgeb_ctx* ctx, float* G, uint8_t* touched, int64_t n_rows, int width, int ld, int32_t* ids_out,
                          float* rows_out, int64_t cap, int64_t* count_out, int clear, kgeb_stream stream);

/* ---- collectives of the sharded paths: NCCL over NVLink behind the C ABI ------------------------------------
 * The reference is one Python process (tests.py:130-134) and has no counterpart; these are what a host without
 * torch.distributed binds to shard (a) the 1-N ranking -- every GPU ranks all queries against its entity rows and the int32
 * partial counts are summed -- and (b) the 1M-entity training step -- requested rows travel owners -> requesters, gradient rows
 * travel back.  NCCL is loaded at run time (libnccl.so.2; KGEB_NCCL_LIB overrides the name): KGEB_ERR_UNSUPPORTED if absent.
 * One communicator per context; every call is enqueued on the caller's stream and never synchronises. */
#define KGEB_COMM_ID_BYTES 128
/* rank 0 fills id_out[KGEB_COMM_ID_BYTES] (ncclGetUniqueId) and hands the bytes to the other ranks by any side channel */
KGEB_API int kgeb_comm_unique_id(void* id_out);
KGEB_API int kgeb_comm_version(void);                       /* NCCL version code of the loaded library, or KGEB_ERR_UNSUPPORTED */
/* collective: every rank of `world` calls it with the same id (ncclCommInitRank on the context's device) */
KGEB_API int kgeb_comm_init(kgeb_ctx* ctx, const void* id, int rank, int world);
/* orderly teardown (ncclCommDestroy): CUDA graphs that captured collectives of this communicator must be destroyed first --
 * NCCL waits for them.  kgeb_destroy on a context that still has a communicator aborts it instead (never blocks). */
KGEB_API int kgeb_comm_destroy(kgeb_ctx* ctx);
KGEB_API int kgeb_comm_info(const kgeb_ctx* ctx, int* rank, int* world);   /* (0, 1) without a communicator */
/* counts[n] <- sum over the ranks, in place: rank = 1 + the summed partial counts of kgeb_rank_1n / kgeb_rank_1n_tc */
KGEB_API int kgeb_allreduce_counts(kgeb_ctx* ctx, int32_t* counts, int64_t n, kgeb_stream stream);
/* buf[n] <- sum over the ranks (replicated relation gradient, loss); flags[n] <- max over the ranks (touched flags) */
KGEB_API int kgeb_allreduce_f32(kgeb_ctx* ctx, float* buf, int64_t n, kgeb_stream stream);
KGEB_API int kgeb_allreduce_flags(kgeb_ctx* ctx, uint8_t* flags, int64_t n, kgeb_stream stream);
/* rows of `width` floats: `send` holds send_rows[g] rows for rank g, concatenated in rank order; `recv` gets recv_rows[g] rows
 * from rank g in rank order.  send_rows / recv_rows are HOST arrays of length world (grouped ncclSend / ncclRecv). */
KGEB_API int kgeb_alltoall_rows(kgeb_ctx* ctx, const float* send, const int64_t* send_rows, float* recv, const int64_t* recv_rows,
                       int width, kgeb_stream stream);

/* ---- one table over the GPUs of an NVLink box (the 1M-entity config; peer.cu) -----------------------------------
 * `world` equal shards, shard g in the memory of GPU g, mapped back to back at ONE virtual range on every rank (CUDA virtual
 * memory management, POSIX file-descriptor handles): base + row * ld * 4 addresses a row wherever it lives, so kgeb_train_fwd_bwd
 * runs unchanged on GLOBAL ids -- its bulk-async row gathers become NVLink peer loads and its vector reductions NVLink peer
 * reductions into the owner's gradient shard: the row exchange happens inside the scoring kernel, row by row, not as two
 * all-to-alls around it.  Needs P2P access between the GPUs (KGEB_ERR_CUDA from kgeb_peer_import_fd otherwise).  The reference
 * (one process, one device, tests.py:130-134) has no counterpart. */
typedef struct kgeb_peer_table kgeb_peer_table;
KGEB_API int64_t kgeb_peer_granularity(int device);      /* shard sizes are rounded up to this many bytes (2 MiB today) */
/* allocates this rank's shard on `device`, reserves the range of world shards and maps the own shard at slot `rank` */
KGEB_API int kgeb_peer_create(kgeb_peer_table** out, int device, int rank, int world, int64_t shard_bytes);
KGEB_API int kgeb_peer_destroy(kgeb_peer_table* t);
KGEB_API const char* kgeb_peer_last_error(const kgeb_peer_table* t);
KGEB_API int64_t kgeb_peer_shard_bytes(const kgeb_peer_table* t);          /* after rounding */
KGEB_API void* kgeb_peer_base(const kgeb_peer_table* t);                   /* device pointer of slot 0 */
/* a file descriptor naming this rank's shard (owned by the table); hand it to the other ranks' processes (SCM_RIGHTS) */
KGEB_API int kgeb_peer_export_fd(kgeb_peer_table* t, int* fd);
/* maps rank src_rank's shard, received as a descriptor valid in this process, at slot src_rank */
KGEB_API int kgeb_peer_import_fd(kgeb_peer_table* t, int src_rank, int fd);
KGEB_API int kgeb_peer_complete(const kgeb_peer_table* t);                 /* 1 once every slot is mapped */
/* Owner-side reduce of per-rank gradient buffers mapped slot_stride floats apart (a peer table whose shard is a rank's WHOLE
 * private gradient buffer): for the owner's rows row0 .. row0+n_rows-1 whose flags[i] is set (flags may be NULL = all),
 *   G[rank][row] += sum over the other ranks p (ascending) of G[p][row];   G[p][row] = 0
 * -- a reduce-scatter over exactly the touched rows, pulled over NVLink by the kernel that feeds the touched-row optimizer. */
KGEB_API int kgeb_peer_pull_add(kgeb_ctx* ctx, float* G_base, int64_t slot_stride, int world, int rank, int64_t row0, int64_t n_rows,
                       int width, int ld, const uint8_t* flags, kgeb_stream stream);

/* ---- literal branch of EmbeddingModel.call, models.py:114-120,152-158 (the reference harness trains it: tests.py:113,125-126):
 *      s' = act([Dropout(E[s]) | Dropout(L[s])] W + b) before the score function, one shared Dense for s and o.
 * The step materialises the transformed rows (per-reference dropout masks make them a function of the reference, not of the
 * entity) and scores them with the ordinary kernels through identity ids; these five calls are the Dense part. -------------- */
typedef enum { KGEB_ACT_LINEAR = 0, KGEB_ACT_TANH = 1, KGEB_ACT_RELU = 2, KGEB_ACT_SIGMOID = 3 } kgeb_act;
/* X[i, :we] = mask * E[ids[i], :we],  X[i, we:we+wl] = mask * L[ids[i], :wl]  (L may be NULL with wl = 0: relation rows).
 * mask = the Dropout keep/scale armed by kgeb_set_dropout (row index i and column are the counter; stream_id separates calls);
 * disarmed -> plain gather. */
KGEB_API int kgeb_rows_concat(kgeb_ctx* ctx, const float* E, int we, int ldE, const float* L, int wl, int ldL, const int32_t* ids,
                     int64_t n, float* X, int ldX, kgeb_stream stream);
/* gE[ids[i], :we] += mask * gX[i, :we]; touched[ids[i]] = 1   (same dropout state as the forward call; literals get no gradient) */
KGEB_API int kgeb_rows_concat_bwd(kgeb_ctx* ctx, const float* gX, int ldX, int we, const int32_t* ids, int64_t n, float* gE, int ldE,
                         uint8_t* touched, kgeb_stream stream);
/* C[M,N] = op(A) op(B) (+ C when accumulate != 0), fp32, row-major; op(A) is M x K (transA: A is stored K x M), op(B) is K x N */
KGEB_API int kgeb_sgemm(kgeb_ctx* ctx, int transA, int transB, int64_t M, int N, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, int accumulate, kgeb_stream stream);
/* Y[rows, width] = act(X[rows, kin] W[kin, width] + bias): the Dense forward (models.py:118,156-157) with the bias and the
 * activation applied in the GEMM epilogue */
KGEB_API int kgeb_dense_fwd(kgeb_ctx* ctx, const float* X, int ldx, int64_t rows, int kin, const float* W, int ldw, int width,
                   const float* bias, int act, float* Y, int ldy, kgeb_stream stream);
/* Z <- act(Z + bias) in place (Keras Dense activation) */
KGEB_API int kgeb_dense_act_fwd(kgeb_ctx* ctx, float* Z, int64_t M, int N, int ld, const float* bias, int act, kgeb_stream stream);
/* gY <- gY * act'(Y) in place, gbias[n] += sum_m gY[m,n] */
KGEB_API int kgeb_dense_act_bwd(kgeb_ctx* ctx, const float* Y, float* gY, int64_t M, int N, int ld, int act, float* gbias, kgeb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* KGEB200_H */
