// wl_group_kernels.cu -- orders the worklist of undecided (query, entity) pairs of a tensor-core ranking pass by entity, on the
// device, without the host ever learning its length.
//
// Why (profiles/r02_seqk_pairs_c5_summary.md): the exact re-scoring kernel k_seqk_pairs stages both 2 KB rows of every pair through
// shared memory.  On the 1M-entity shape an entity occurs in ~12 undecided pairs, yet the worklist arrives in tile order where
// almost every pair of a warp names a different entity: 80 GB of L2 reads per side, lts__throughput 74.5 % of peak -- the kernel
// is bound by L2 -> SM bandwidth.  With the pairs of one entity adjacent, the 32 lanes of a warp share the staged copy of each
// distinct entity row (k_seqk_pairs<OP, PAIRS_GROUPED>) and the entity rows leave the L2 once per warp instead of once per pair.
//
// How: a counting sort keyed by the shard-local entity id, three launches plus one memset, every length read on the device:
//   k_wl_hist     slot[x] = atomicAdd(hist[c_x], 1) for the first n = min(appended, capacity) pairs (the order of the pairs of one
//                 entity is whatever the atomics decide: the re-scoring adds integers, any order counts the same),
//   ExclusiveSum  offsets over the n_rows + 1 bins (CUB; the bin count is known to the host), offsets[n_rows] = n,
//   k_wl_scatter  out[offsets[c_x] + slot[x]] = pair x.
#include <cub/cub.cuh>

#include "common.cuh"
#include "rank_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_wl_hist(const uint32_t* __restrict__ hdr, uint32_t cap, const uint2* __restrict__ wl,
                                                 uint32_t n_rows, uint32_t* __restrict__ hist, uint32_t* __restrict__ slot) {
  const uint32_t n = hdr[0] < cap ? hdr[0] : cap;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    const uint32_t c = wl[x].y;
    slot[x] = c < n_rows ? atomicAdd(hist + c, 1u) : 0xffffffffu;   // a pair outside the shard cannot occur; it would be dropped
  }
}

__global__ void __launch_bounds__(256) k_wl_scatter(const uint32_t* __restrict__ hdr, uint32_t cap, const uint2* __restrict__ wl,
                                                    const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ slot,
                                                    uint2* __restrict__ out) {
  const uint32_t n = hdr[0] < cap ? hdr[0] : cap;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    const uint32_t s = slot[x];
    if (s == 0xffffffffu) continue;
    const uint2 pr = wl[x];
    out[offsets[pr.y] + s] = pr;
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// byte offsets of the scratch regions inside a caller-provided block: [bins n_rows + 1][slot cap][ordered worklist cap][CUB temp]
KgebWlGroupLayout kgeb_internal_wl_group_layout(int64_t n_rows, int64_t cap) {
  KgebWlGroupLayout L;
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n_rows + 1));
  L.bins = 0;
  L.slot = L.bins + align256((size_t)(n_rows + 1) * 4);
  L.out = L.slot + align256((size_t)cap * 4);
  L.cub = L.out + align256((size_t)cap * 8);
  L.cub_bytes = tmp;
  L.total = L.cub + align256(tmp);
  return L;
}

// after the call (in stream order): ordered = the first min(hdr[0], cap) pairs of wl sorted by entity, *length = their number
int kgeb_internal_wl_group(kgeb_ctx* ctx, const uint32_t* hdr, uint32_t cap, const uint2* wl, int64_t n_rows, unsigned char* scratch,
                           const uint2** ordered, const uint32_t** length, cudaStream_t st) {
  const KgebWlGroupLayout L = kgeb_internal_wl_group_layout(ctx->N, cap);   // as the workspace was sized (n_rows <= N)
  uint32_t* bins = reinterpret_cast<uint32_t*>(scratch + L.bins);
  uint32_t* slot = reinterpret_cast<uint32_t*>(scratch + L.slot);
  uint2* out = reinterpret_cast<uint2*>(scratch + L.out);
  KGEB_CUDA_OK(ctx, cudaMemsetAsync(bins, 0, (size_t)(n_rows + 1) * 4, st));
  const int blocks = ctx->num_sms * 8;
  k_wl_hist<<<blocks, 256, 0, st>>>(hdr, cap, wl, (uint32_t)n_rows, bins, slot);
  KGEB_LAUNCH_OK(ctx);
  size_t tmp = L.cub_bytes;
  KGEB_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(scratch + L.cub, tmp, bins, bins, (int)(n_rows + 1), st));
  ctx->launches += 2;
  k_wl_scatter<<<blocks, 256, 0, st>>>(hdr, cap, wl, bins, slot, out);
  KGEB_LAUNCH_OK(ctx);
  *ordered = out;
  *length = bins + n_rows;
  return KGEB_OK;
}
