// export_kernels.cu -- the split form of the backward: (unique_ids, grad_rows) instead of an in-library update.
//
// A tf.keras Embedding hands its optimizer IndexedSlices (models.py:89-112; `tests.py:86` compiles with a stock optimizer), so a
// registered gradient for the TF op shim -- or any host that keeps its own optimizer -- needs the touched rows of the dense
// gradient buffer as a compact (ids, rows) pair: ids ascending, one entry per touched row, rows[i] = G[ids[i], :width].
// Three small kernels, no library calls: per-block counts of the row flags, one block scans the counts, then every block places
// its rows (a warp copies a row with 128-bit accesses) and, if asked, clears the buffer rows and flags it consumed so that the
// buffer is all-zero again for the next step.
#include "common.cuh"

namespace {

constexpr int EX_ROWS = 1024;   // rows per block

__global__ void __launch_bounds__(256) k_ex_count(const uint8_t* __restrict__ flags, int64_t n, uint32_t* __restrict__ block_count) {
  const int64_t r0 = (int64_t)blockIdx.x * EX_ROWS;
  int c = 0;
  for (int i = threadIdx.x; i < EX_ROWS; i += 256) c += (r0 + i < n && flags[r0 + i]) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int w[8];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += w[i];
    block_count[blockIdx.x] = (uint32_t)t;
  }
}

// exclusive scan of block_count in place (one block; the array has at most 2M entries), total -> *count_out
__global__ void __launch_bounds__(1024) k_ex_scan(uint32_t* __restrict__ block_count, int64_t nb, int64_t* __restrict__ count_out) {
  __shared__ uint32_t part[1024];
  const int64_t per = (nb + 1023) / 1024;
  const int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
  uint32_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += block_count[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
    *count_out = (int64_t)run;
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int64_t i = lo; i < hi; ++i) { const uint32_t v = block_count[i]; block_count[i] = run; run += v; }
}

__global__ void __launch_bounds__(256) k_ex_place(float* __restrict__ G, uint8_t* __restrict__ flags, int64_t n, int ld, int w4,
                                                 const uint32_t* __restrict__ block_base, int32_t* __restrict__ ids_out,
                                                 float* __restrict__ rows_out, int64_t cap, int clear) {
  __shared__ int list[EX_ROWS];
  __shared__ int wsum[8];
  __shared__ int total;
  const int64_t r0 = (int64_t)blockIdx.x * EX_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ordered compaction of this block's touched rows: thread t owns rows 4t .. 4t+3
  int f[4], c = 0;
  for (int j = 0; j < 4; ++j) { const int64_t r = r0 + 4 * threadIdx.x + j; f[j] = (r < n && flags[r]) ? 1 : 0; c += f[j]; }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) { int run = 0; for (int i = 0; i < 8; ++i) { const int v = wsum[i]; wsum[i] = run; run += v; } total = run; }
  __syncthreads();
  int pos = wsum[warp] + incl - c;
  for (int j = 0; j < 4; ++j) if (f[j]) list[pos++] = 4 * threadIdx.x + j;
  __syncthreads();
  const int64_t base = block_base[blockIdx.x];
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = warp; i < total; i += 8) {
    const int64_t r = r0 + list[i], o = base + i;
    if (o < cap) {
      float4* src = reinterpret_cast<float4*>(G + r * ld);
      float4* dst = reinterpret_cast<float4*>(rows_out + o * (int64_t)(4 * w4));
      for (int cc = lane; cc < w4; cc += 32) { dst[cc] = src[cc]; if (clear) src[cc] = zero; }
      if (lane == 0) { ids_out[o] = (int32_t)r; if (clear) flags[r] = 0; }
    }
  }
}

}  // namespace

extern "C" int kgeb_grad_rows_export(kgeb_ctx* ctx, float* G, uint8_t* touched, int64_t n_rows, int width, int ld, int32_t* ids_out,
                                     float* rows_out, int64_t cap, int64_t* count_out, int clear, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, G && touched && ids_out && rows_out && count_out && n_rows > 0 && cap >= 0 && width > 0 && ld >= width, KGEB_ERR_INVALID,
               "grad_rows_export: bad arguments");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && kgeb_aligned16(G) && kgeb_aligned16(rows_out), KGEB_ERR_ALIGNMENT,
               "grad_rows_export needs width, ld multiples of 4 floats and 16-byte aligned buffers");
  const int64_t nb = (n_rows + EX_ROWS - 1) / EX_ROWS;
  KGEB_REQUIRE(ctx, (size_t)nb * 4 <= KGEB_SORT_SCRATCH_BYTES, KGEB_ERR_UNSUPPORTED, "grad_rows_export: table too large for the context scratch");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* bc = reinterpret_cast<uint32_t*>(ctx->d_sort_scratch);
  k_ex_count<<<(int)nb, 256, 0, st>>>(touched, n_rows, bc);
  KGEB_LAUNCH_OK(ctx);
  k_ex_scan<<<1, 1024, 0, st>>>(bc, nb, count_out);
  KGEB_LAUNCH_OK(ctx);
  k_ex_place<<<(int)nb, 256, 0, st>>>(G, touched, n_rows, ld, width / 4, bc, ids_out, rows_out, cap, clear ? 1 : 0);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
