// filter_kernels.cu -- the filter sets of validate() (utils.py:201-205: dict (s,p) -> {o}, (p,o) -> {s}) as one sorted,
// de-duplicated CSR per side, built on the device: pack every filtering triple into one 64-bit key
// (key0 * span1 + key1) * spanV + value (span of a column = the number of ids it can hold: N for s and o, R for p), radix-sort, drop adjacent duplicates, then two binary searches per query give its
// segment.  The per-query list.remove of gen_tail_data / gen_head_data (utils.py:169-171,185-187) becomes "the target's id
// must not be in its own segment" (reported through a flag: the reference crashes there).  Sort / unique / scan are CUB
// device primitives (library code, like the NCCL calls); keys, searches and the gather are kernels of this file.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

struct FilterWs {
  unsigned long long *keys_a, *keys_b, *uniq;
  long long* nuniq;
  long long* lens;      // [Q + 1]
  void* cub_tmp;
  size_t cub_bytes, total;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

FilterWs carve(void* base, int64_t Q, int64_t F) {
  FilterWs w;
  size_t s_sort = 0, s_uniq = 0, s_scan = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, s_sort, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)F);
  cub::DeviceSelect::Unique(nullptr, s_uniq, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long*)nullptr, (int)F);
  cub::DeviceScan::ExclusiveSum(nullptr, s_scan, (const long long*)nullptr, (long long*)nullptr, (int)(Q + 1));
  w.cub_bytes = s_sort > s_uniq ? s_sort : s_uniq;
  if (s_scan > w.cub_bytes) w.cub_bytes = s_scan;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  w.keys_a = reinterpret_cast<unsigned long long*>(p + off); off += align256((size_t)F * 8);
  w.keys_b = reinterpret_cast<unsigned long long*>(p + off); off += align256((size_t)F * 8);
  w.uniq = reinterpret_cast<unsigned long long*>(p + off); off += align256((size_t)F * 8);
  w.nuniq = reinterpret_cast<long long*>(p + off); off += 256;
  w.lens = reinterpret_cast<long long*>(p + off); off += align256((size_t)(Q + 1) * 8);
  w.cub_tmp = p + off; off += align256(w.cub_bytes);
  w.total = off;
  return w;
}

struct FilterSpan { unsigned long long s1, sv; };   // spans of the second key column and of the value column

__global__ void k_filter_keys(const int32_t* __restrict__ filt, int64_t F, int c0, int c1, int cv, FilterSpan span, unsigned long long* keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((unsigned long long)filt[3 * i + c0] * span.s1 + (unsigned long long)filt[3 * i + c1]) * span.sv + (unsigned long long)filt[3 * i + cv];
}

__device__ __forceinline__ long long lower_bound(const unsigned long long* a, long long n, unsigned long long x) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// lens[q] = size of query q's segment; lens[Q] = 0 (so that the exclusive scan's last entry is nnz)
__global__ void k_filter_lens(const int32_t* __restrict__ queries, int64_t Q, int c0, int c1, FilterSpan span,
                              const unsigned long long* __restrict__ uniq, const long long* __restrict__ nuniq, long long* lens) {
  const long long n = *nuniq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (q == Q) { lens[q] = 0; continue; }
    const unsigned long long base = ((unsigned long long)queries[3 * q + c0] * span.s1 + (unsigned long long)queries[3 * q + c1]) * span.sv;
    lens[q] = lower_bound(uniq, n, base + span.sv) - lower_bound(uniq, n, base);
  }
}

// a warp per query copies its segment's values; target_hit |= 1 if the query's own target id is among them
__global__ void k_filter_fill(const int32_t* __restrict__ queries, int64_t Q, int c0, int c1, int cv, FilterSpan span,
                              const unsigned long long* __restrict__ uniq, const long long* __restrict__ nuniq,
                              const int64_t* __restrict__ offsets, int32_t* __restrict__ ids, int32_t* target_hit) {
  const long long n = *nuniq;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = warp0; q < Q; q += nwarps) {
    const unsigned long long base = ((unsigned long long)queries[3 * q + c0] * span.s1 + (unsigned long long)queries[3 * q + c1]) * span.sv;
    const long long lo = lower_bound(uniq, n, base);
    const int64_t o0 = offsets[q], len = offsets[q + 1] - o0;
    const int32_t tgt = queries[3 * q + cv];
    bool hit = false;
    for (int64_t j = lane; j < len; j += 32) {
      const int32_t v = (int32_t)(uniq[lo + j] - base);
      ids[o0 + j] = v;
      hit |= v == tgt;
    }
    if (hit) atomicOr(target_hit, 1);
  }
}

int side_cols(int side, int* c0, int* c1, int* cv) {
  if (side == KGEB_SIDE_TAIL) { *c0 = 0; *c1 = 1; *cv = 2; return 0; }
  if (side == KGEB_SIDE_HEAD) { *c0 = 1; *c1 = 2; *cv = 0; return 0; }
  return -1;
}
FilterSpan filter_span(const kgeb_ctx* ctx, int c1, int cv) {
  FilterSpan s;
  s.s1 = (unsigned long long)(c1 == 1 ? ctx->R : ctx->N);
  s.sv = (unsigned long long)(cv == 1 ? ctx->R : ctx->N);
  return s;
}

}  // namespace

extern "C" int kgeb_filter_csr_workspace(kgeb_ctx* ctx, int64_t Q, int64_t F, int64_t* bytes) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, bytes && Q >= 0 && F >= 0 && F < 0x7fffffffLL && Q < 0x7fffffffLL, KGEB_ERR_INVALID, "filter_csr_workspace: bad arguments");
  *bytes = (int64_t)carve(nullptr, Q, F > 0 ? F : 1).total;
  return KGEB_OK;
}

extern "C" int kgeb_filter_csr_count(kgeb_ctx* ctx, const int32_t* queries, int64_t Q, const int32_t* filt, int64_t F, int side, int64_t* offsets,
                                     void* workspace, int64_t workspace_bytes, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  int c0, c1, cv;
  KGEB_REQUIRE(ctx, side_cols(side, &c0, &c1, &cv) == 0, KGEB_ERR_INVALID, "unknown side");
  KGEB_REQUIRE(ctx, Q >= 0 && F > 0 && F < 0x7fffffffLL && Q < 0x7fffffffLL && offsets && workspace && filt && (Q == 0 || queries), KGEB_ERR_INVALID,
               "filter_csr_count: bad arguments");
  const FilterSpan span = filter_span(ctx, c1, cv);
  // every key is < N * R * N: with R = 1000 relations that is 96 M entities (the first build cubed max(N, R): 1.66 M)
  KGEB_REQUIRE(ctx, (long double)ctx->N * (long double)ctx->R * (long double)ctx->N < 9.2e18L, KGEB_ERR_UNSUPPORTED,
               "ids too large for the packed 64-bit filter key (N * R * N must stay below 2^63)");
  FilterWs w = carve(workspace, Q, F);
  KGEB_REQUIRE(ctx, (int64_t)w.total <= workspace_bytes, KGEB_ERR_INVALID, "filter_csr_count: workspace too small (kgeb_filter_csr_workspace)");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((F + 255) / 256 < ctx->num_sms * 8 ? (F + 255) / 256 : ctx->num_sms * 8);
  k_filter_keys<<<blocks, 256, 0, st>>>(filt, F, c0, c1, cv, span, w.keys_a);
  KGEB_LAUNCH_OK(ctx);
  size_t tb = w.cub_bytes;
  KGEB_CUDA_OK(ctx, cub::DeviceRadixSort::SortKeys(w.cub_tmp, tb, w.keys_a, w.keys_b, (int)F, 0, 64, st));
  tb = w.cub_bytes;
  KGEB_CUDA_OK(ctx, cub::DeviceSelect::Unique(w.cub_tmp, tb, w.keys_b, w.uniq, w.nuniq, (int)F, st));
  blocks = (int)((Q + 256) / 256 < ctx->num_sms * 8 ? (Q + 256) / 256 : ctx->num_sms * 8);
  k_filter_lens<<<blocks, 256, 0, st>>>(queries, Q, c0, c1, span, w.uniq, w.nuniq, w.lens);
  KGEB_LAUNCH_OK(ctx);
  tb = w.cub_bytes;
  KGEB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.lens, reinterpret_cast<long long*>(offsets), (int)(Q + 1), st));
  ctx->launches += 3;
  return KGEB_OK;
}

extern "C" int kgeb_filter_csr_fill(kgeb_ctx* ctx, const int32_t* queries, int64_t Q, int64_t F, int side, const int64_t* offsets, int32_t* ids,
                                    int32_t* target_hit, const void* workspace, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  int c0, c1, cv;
  KGEB_REQUIRE(ctx, side_cols(side, &c0, &c1, &cv) == 0, KGEB_ERR_INVALID, "unknown side");
  KGEB_REQUIRE(ctx, Q >= 0 && F > 0 && offsets && target_hit && workspace && (Q == 0 || queries), KGEB_ERR_INVALID, "filter_csr_fill: bad arguments");
  if (Q == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, ids, KGEB_ERR_INVALID, "filter_csr_fill: ids is null");
  FilterWs w = carve(const_cast<void*>(workspace), Q, F);
  int64_t blocks = (Q + 7) / 8;
  if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
  k_filter_fill<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(queries, Q, c0, c1, cv, filter_span(ctx, c1, cv), w.uniq, w.nuniq, offsets, ids, target_hit);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
