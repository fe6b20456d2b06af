// rows_kernels.cu -- row exchange helpers of the row-sharded training path (BASELINE configs[4]): the owner of
// a table shard gathers the rows other ranks asked for into a contiguous send buffer, and adds the gradient rows that
// come back into its dense gradient buffer.  One warp per row, 128-bit accesses; the NCCL all-to-all between the two
// calls is the host's (torch.distributed) business -- these kernels never touch another rank's memory.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(KGEB_THREADS) k_rows_gather(const float* __restrict__ P, int ld, int w4, const int32_t* __restrict__ idx,
                                                              int64_t n, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t i = warp0; i < n; i += nw) {
    const float4* src = reinterpret_cast<const float4*>(P + (int64_t)__ldg(idx + i) * ld);
    float4* d = reinterpret_cast<float4*>(dst + i * (int64_t)(4 * w4));
    for (int c = lane; c < w4; c += 32) d[c] = ldg4(src + c);
  }
}

__global__ void __launch_bounds__(KGEB_THREADS) k_rows_scatter_add(float* G, int ld, int w4, uint8_t* touched, const int32_t* __restrict__ idx,
                                                                   int64_t n, const float* __restrict__ src) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  for (int64_t i = warp0; i < n; i += nw) {
    const int64_t r = __ldg(idx + i);
    const float4* s = reinterpret_cast<const float4*>(src + i * (int64_t)(4 * w4));
    float* g = G + r * ld;
    for (int c = lane; c < w4; c += 32) red_add4(g + 4 * c, ldg4(s + c));   // several ranks may return the same row
    if (lane == 0 && touched) touched[r] = 1;
  }
}

int rows_grid(const kgeb_ctx* ctx, int64_t n) {
  int64_t b = (n + KGEB_WARPS_PER_BLOCK - 1) / KGEB_WARPS_PER_BLOCK;
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int kgeb_rows_gather(kgeb_ctx* ctx, const float* P, int width, int ld, const int32_t* idx, int64_t n, float* dst,
                                kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, n >= 0 && width > 0 && ld >= width, KGEB_ERR_INVALID, "rows_gather: bad arguments");
  if (n == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, P && idx && dst, KGEB_ERR_INVALID, "rows_gather: null pointer");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && kgeb_aligned16(P) && kgeb_aligned16(dst), KGEB_ERR_ALIGNMENT,
               "rows_gather needs width, ld multiples of 4 floats and 16-byte aligned buffers");
  k_rows_gather<<<rows_grid(ctx, n), KGEB_THREADS, 0, (cudaStream_t)stream>>>(P, ld, width / 4, idx, n, dst);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_rows_scatter_add(kgeb_ctx* ctx, float* G, int width, int ld, uint8_t* touched, const int32_t* idx, int64_t n,
                                     const float* src, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, n >= 0 && width > 0 && ld >= width, KGEB_ERR_INVALID, "rows_scatter_add: bad arguments");
  if (n == 0) return KGEB_OK;
  KGEB_REQUIRE(ctx, G && idx && src, KGEB_ERR_INVALID, "rows_scatter_add: null pointer");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && kgeb_aligned16(G) && kgeb_aligned16(src), KGEB_ERR_ALIGNMENT,
               "rows_scatter_add needs width, ld multiples of 4 floats and 16-byte aligned buffers");
  k_rows_scatter_add<<<rows_grid(ctx, n), KGEB_THREADS, 0, (cudaStream_t)stream>>>(G, ld, width / 4, touched, idx, n, src);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

// ---- owner-side reduction of the peer-mapped gradient buffers (peer.cu, KGEkeras/_peer.py) ----------------------------------------
// Every rank accumulates the gradient rows of ITS batch in a private full-size buffer (local vector reductions, as on one GPU); the
// buffers of all ranks are peer-mapped `slot_stride` floats apart.  After the step's barrier the OWNER of a row pulls the other
// ranks' partial rows over NVLink -- coalesced 128-bit loads, whole rows, ranks in a fixed order so the sum is reproducible -- adds
// them to its own, and clears the remote copies on the way (NVLink is full duplex: the zero stores ride the idle direction), so a
// reduce-scatter over exactly the touched rows happens inside one kernel, feeding the touched-row optimizer that follows.
namespace {

__global__ void __launch_bounds__(KGEB_THREADS) k_peer_pull_add(float* __restrict__ Gbase, int64_t slot_stride, int world, int rank, int64_t row0,
                                                                int64_t n_rows, int ld, int w4, const uint8_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * KGEB_WARPS_PER_BLOCK;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = warp0; i < n_rows; i += nw) {
    if (flags && !__ldg(flags + i)) continue;                       // nobody referenced the row this step: every copy is zero
    const int64_t off = (row0 + i) * ld;
    float4* mine = reinterpret_cast<float4*>(Gbase + (int64_t)rank * slot_stride + off);
    for (int c = lane; c < w4; c += 32) {
      float4 acc = mine[c];
      for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        float4* theirs = reinterpret_cast<float4*>(Gbase + (int64_t)p * slot_stride + off);
        const float4 v = theirs[c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        theirs[c] = zero;
      }
      mine[c] = acc;
    }
  }
}

}  // namespace

extern "C" int kgeb_peer_pull_add(kgeb_ctx* ctx, float* G_base, int64_t slot_stride, int world, int rank, int64_t row0, int64_t n_rows,
                                  int width, int ld, const uint8_t* flags, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world && row0 >= 0 && n_rows >= 0 && width > 0 && ld >= width && slot_stride >= 0,
               KGEB_ERR_INVALID, "peer_pull_add: bad arguments");
  if (n_rows == 0 || world == 1) return KGEB_OK;
  KGEB_REQUIRE(ctx, G_base, KGEB_ERR_INVALID, "peer_pull_add: null pointer");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && slot_stride % 4 == 0 && kgeb_aligned16(G_base), KGEB_ERR_ALIGNMENT,
               "peer_pull_add needs width, ld, slot stride multiples of 4 floats and a 16-byte aligned base");
  k_peer_pull_add<<<rows_grid(ctx, n_rows), KGEB_THREADS, 0, (cudaStream_t)stream>>>(G_base, slot_stride, world, rank, row0, n_rows, ld, width / 4, flags);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
