// loss_opt_kernels.cu -- losses on score vectors (utils.py:288-319) with their derivatives, and the touched-row
// optimizer (K2 second half; Keras Adam/Adagrad/SGD semantics).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

// -------- pointwise: L = mean_i g(t_i) + mean_j h(f_j) ---------------------------------------------------
__global__ void __launch_bounds__(KGEB_THREADS) k_loss_pointwise(const float* __restrict__ t, int64_t B,
                                                                 const float* __restrict__ f, int64_t Bf, int kind,
                                                                 float margin, float w, float* loss, float* dt, float* df,
                                                                 double* partials, unsigned int* ticket) {
  double lp = 0.0, ln = 0.0;
  const float sP = (float)((double)w / (double)B), sN = (float)((double)w / (double)Bf);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B + Bf; i += stride) {
    float val, der;
    if (i < B) {
      pointwise_terms(kind, margin, __ldg(t + i), true, val, der);
      lp += (double)val;
      if (dt) dt[i] = der * sP;
    } else {
      pointwise_terms(kind, margin, __ldg(f + (i - B)), false, val, der);
      ln += (double)val;
      if (df) df[i - B] = der * sN;
    }
  }
  block_finalize_loss(lp, ln, partials, ticket, (double)w / (double)B, (double)w / (double)Bf, loss);
}

// -------- pairwise PAIRED: X_ij = f_{iK+j} - t_i -----------------------------------------------------------
__global__ void __launch_bounds__(KGEB_THREADS) k_loss_paired(const float* __restrict__ t, int64_t B,
                                                              const float* __restrict__ f, int K, int kind, float margin,
                                                              float w, float* loss, float* dt, float* df, double* partials,
                                                              unsigned int* ticket) {
  double ln = 0.0;
  const float sN = (float)((double)w / ((double)B * (double)K));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
    const float ti = __ldg(t + i);
    float cpos = 0.f;
    for (int j = 0; j < K; ++j) {
      float val, A;
      pairwise_terms(kind, margin, __ldg(f + i * K + j) - ti, val, A);
      ln += (double)val;
      const float c = A * sN;
      if (df) df[i * K + j] = c;
      cpos -= c;
    }
    if (dt) dt[i] = cpos;
  }
  block_finalize_loss(0.0, ln, partials, ticket, 0.0, (double)w / ((double)B * (double)K), loss);
}

// -------- pairwise ALLPAIRS: X_ji = f_j - t_i over the full (Bf x B) grid (tests.py:81 broadcast) --------------
// blockIdx.z == 0: a thread owns one t_i and walks a slice of f (accumulates loss and dt_i);
// blockIdx.z == 1: a thread owns one f_j and walks a slice of t (accumulates df_j).  Partial sums meet in
// float atomics on dt/df (zeroed by the host wrapper first).
#define AP_TILE 256
__global__ void __launch_bounds__(AP_TILE) k_loss_allpairs(const float* __restrict__ t, int64_t B, const float* __restrict__ f,
                                                           int64_t Bf, int kind, float margin, float w, float* dt, float* df,
                                                           double* loss_acc) {
  __shared__ float tile[AP_TILE];
  const bool own_t = blockIdx.z == 0;
  const float* own = own_t ? t : f;
  const float* oth = own_t ? f : t;
  const int64_t n_own = own_t ? B : Bf, n_oth = own_t ? Bf : B;
  const int64_t nblk_own = (n_own + AP_TILE - 1) / AP_TILE;
  if ((int64_t)blockIdx.x >= nblk_own) return;
  const int64_t i = (int64_t)blockIdx.x * AP_TILE + threadIdx.x;
  const float mine = i < n_own ? __ldg(own + i) : 0.f;
  // slice of the other vector handled by this blockIdx.y
  const int64_t per = ((n_oth + gridDim.y - 1) / gridDim.y + AP_TILE - 1) / AP_TILE * AP_TILE;
  const int64_t j0 = (int64_t)blockIdx.y * per;
  const int64_t j1 = j0 + per < n_oth ? j0 + per : n_oth;
  float accA = 0.f;
  double accL = 0.0;
  for (int64_t base = j0; base < j1; base += AP_TILE) {
    __syncthreads();
    const int64_t jj = base + threadIdx.x;
    tile[threadIdx.x] = jj < j1 ? __ldg(oth + jj) : 0.f;
    __syncthreads();
    const int lim = (int)((j1 - base) < AP_TILE ? (j1 - base) : AP_TILE);
    float l32 = 0.f;
    for (int k = 0; k < lim; ++k) {
      float val, A;
      const float x = own_t ? tile[k] - mine : mine - tile[k];
      pairwise_terms<true>(kind, margin, x, val, A);
      accA += A;
      l32 += val;
    }
    accL += (double)l32;
  }
  const double cnt = (double)B * (double)Bf;
  const float sc = (float)((double)w / cnt);
  if (i < n_own) {
    if (own_t) { if (dt) atomicAdd(dt + i, -accA * sc); }
    else       { if (df) atomicAdd(df + i, accA * sc); }
  }
  if (own_t) {
    if (i >= n_own) accL = 0.0;
    accL = warp_sum_d(accL);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_acc, accL);
  }
}
__global__ void k_loss_allpairs_finish(const double* loss_acc, double scale, float* loss) { loss[0] = (float)(loss_acc[0] * scale); }

// -------- ALLPAIRS hinge without the m x m walk ----------------------------------------------------------------------------
// relu(margin + f_i - t_j) is non-zero exactly for the t_j below a threshold that depends on f_i, and the fp32 predicate
// fl(margin + fl(f_i - t_j)) > 0 is monotone in t_j and in f_i.  So: sort both vectors in 1024-element chunks (one CTA per
// chunk, bitonic network in shared memory), keep an exclusive prefix sum (double) of every sorted t-chunk, and let one thread
// per element binary-search every chunk of the OTHER vector with that same fp32 predicate:
//   df_i = #{j : pred} / cnt,  dt_j = -#{i : pred} / cnt      -- the integer counts of the m x m kernel, bit for bit
//   loss = sum_i [ #_i * (margin + f_i) - sum of the t_j counted ] / cnt       (accumulated in double)
// O(m * m/1024 * 10) probes instead of m^2 terms: the harness' m = 20 480 step (tests.py:81) drops from 251 us to a few us.
constexpr int SC = 1024;   // chunk length
__global__ void __launch_bounds__(SC / 2) k_chunk_sort(const float* __restrict__ t, int64_t B, const float* __restrict__ f, int64_t Bf,
                                                       int tchunks, float* __restrict__ ts, float* __restrict__ fs, double* __restrict__ tp) {
  __shared__ float s[SC];
  __shared__ double ps[SC / 2];
  const bool is_t = (int)blockIdx.x < tchunks;
  const int c = is_t ? blockIdx.x : blockIdx.x - tchunks;
  const float* src = is_t ? t : f;
  const int64_t n = is_t ? B : Bf, base = (int64_t)c * SC;
  const int tid = threadIdx.x;
  for (int i = tid; i < SC; i += SC / 2) s[i] = base + i < n ? __ldg(src + base + i) : __int_as_float(0x7f800000);  // +inf padding sorts last
  __syncthreads();
  for (int k = 2; k <= SC; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));   // the lower index of this thread's compare-exchange pair
      const int p = i | j;
      const bool asc = (i & k) == 0;
      const float a = s[i], b = s[p];
      if ((a > b) == asc) { s[i] = b; s[p] = a; }
      __syncthreads();
    }
  float* dst = (is_t ? ts : fs) + base;
  for (int i = tid; i < SC; i += SC / 2) dst[i] = s[i];
  if (!is_t) return;
  // exclusive prefix sums of the sorted t-chunk in double: tp[c][x] = sum of the x smallest (x = 0..SC)
  const double a0 = (double)s[2 * tid], a1 = (double)s[2 * tid + 1];
  const bool v0 = base + 2 * tid < n, v1 = base + 2 * tid + 1 < n;   // padding contributes nothing
  ps[tid] = (v0 ? a0 : 0.0) + (v1 ? a1 : 0.0);
  __syncthreads();
  for (int off = 1; off < SC / 2; off <<= 1) {
    const double add = tid >= off ? ps[tid - off] : 0.0;
    __syncthreads();
    ps[tid] += add;
    __syncthreads();
  }
  const double before = tid ? ps[tid - 1] : 0.0;
  double* d = tp + (int64_t)c * (SC + 1);
  d[2 * tid] = before;
  d[2 * tid + 1] = before + (v0 ? a0 : 0.0);
  if (tid == SC / 2 - 1) d[SC] = ps[tid];
}

// one thread per element of f (blockIdx.y == 0) or of t (blockIdx.y == 1)
__global__ void __launch_bounds__(256) k_allpairs_hinge_sorted(const float* __restrict__ t, int64_t B, const float* __restrict__ f, int64_t Bf,
                                                               const float* __restrict__ ts, const float* __restrict__ fs,
                                                               const double* __restrict__ tp, float margin, float w, float* dt, float* df,
                                                               double* loss_acc) {
  const bool f_side = blockIdx.y == 0;
  const int64_t n_own = f_side ? Bf : B, n_oth = f_side ? B : Bf;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = (int)((n_oth + SC - 1) / SC);
  double accL = 0.0;
  int64_t cnt = 0;
  if (i < n_own) {
    const float mine = __ldg((f_side ? f : t) + i);
    const float* oth = f_side ? ts : fs;
    // G chunks are searched in lockstep: their loads are independent, so a thread keeps G of them in flight instead of
    // walking 10 dependent L2 round trips per chunk (the grid is a fraction of a wave: this kernel is pure latency)
    constexpr int G = 8;
    for (int c0 = 0; c0 < nch; c0 += G) {
      int lo[G], hi[G], len[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int64_t left = n_oth - (int64_t)(c0 + g) * SC;
        len[g] = (c0 + g < nch) ? (int)(left < SC ? left : SC) : 0;
        lo[g] = 0;
        hi[g] = len[g];
      }
      // f side: pred(f_i, t) holds for the SMALLEST t's -> count = first index where it fails
      // t side: pred(f, t_j) holds for the LARGEST f's  -> count = len - first index where it holds
#pragma unroll 1
      for (int it = 0; it < 11; ++it) {           // ceil(log2(SC + 1)) halvings finish every chunk
        float o[G];
#pragma unroll
        for (int g = 0; g < G; ++g) o[g] = lo[g] < hi[g] ? __ldg(oth + (int64_t)(c0 + g) * SC + ((lo[g] + hi[g]) >> 1)) : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (lo[g] < hi[g]) {
            const int mid = (lo[g] + hi[g]) >> 1;
            const bool pr = f_side ? (__fadd_rn(margin, __fsub_rn(mine, o[g])) > 0.f) : (__fadd_rn(margin, __fsub_rn(o[g], mine)) > 0.f);
            if (f_side ? pr : !pr) lo[g] = mid + 1; else hi[g] = mid;
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (c0 + g >= nch) continue;
        if (f_side) {
          cnt += lo[g];
          accL += (double)lo[g] * ((double)margin + (double)mine) - tp[(int64_t)(c0 + g) * (SC + 1) + lo[g]];
        } else {
          cnt += len[g] - lo[g];
        }
      }
    }
    const float sc = (float)((double)w / ((double)B * (double)Bf));
    if (f_side) { if (df) df[i] = (float)cnt * sc; }
    else        { if (dt) dt[i] = -(float)cnt * sc; }
  }
  if (f_side) {
    accL = warp_sum_d(accL);
    if ((threadIdx.x & 31) == 0 && accL != 0.0) atomicAdd(loss_acc, accL);
  }
}

// -------- optimizer: a warp scans 32 row flags at a time, then updates the touched rows cooperatively -----------
struct OptArgs {
  float* P; float* G; uint8_t* touched; int64_t rows; int w4; int ld4; float* s0; float* s1;
  int kind; float lr, b1, b2, eps, lr_t;
  int win;  // rows scanned per warp trip (1..32): small when most rows are touched so the work spreads over all SMs
};

__device__ __forceinline__ void opt_elem(int kind, float g, float& p, float& m, float& v, const OptArgs& a) {
  if (kind == KGEB_SGD) {
    p -= a.lr * g;
  } else if (kind == KGEB_ADAGRAD) {
    m += g * g;
    p -= a.lr * g / (sqrtf(m) + a.eps);
  } else {
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    p -= a.lr_t * m / (sqrtf(v) + a.eps);
  }
}

__device__ __forceinline__ void opt_update_body(const OptArgs& a, int nblocks, int block) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)block * KGEB_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)nblocks * KGEB_WARPS_PER_BLOCK;
  const bool dense = a.kind == KGEB_ADAM_DENSE;
  float4* P4 = reinterpret_cast<float4*>(a.P);
  float4* G4 = reinterpret_cast<float4*>(a.G);
  float4* M4 = reinterpret_cast<float4*>(a.s0);
  float4* V4 = reinterpret_cast<float4*>(a.s1);
  for (int64_t base = warp * a.win; base < a.rows; base += nw * a.win) {
    const int64_t r = base + lane;
    const bool inwin = lane < a.win && r < a.rows;
    const bool flag = inwin && (a.touched[r] != 0);
    const bool work = inwin && (dense || flag);
    unsigned mask = __ballot_sync(0xffffffffu, work);
    const unsigned tmask = __ballot_sync(0xffffffffu, flag);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const int64_t row = base + b;
      const bool was_touched = (tmask >> b) & 1u;
      for (int c = lane; c < a.w4; c += 32) {
        const int64_t off = row * a.ld4 + c;
        float4 g = was_touched ? G4[off] : f4zero();
        float4 p = P4[off];
        float4 m = f4zero(), v = f4zero();
        if (a.kind != KGEB_SGD) m = M4[off];
        if (a.kind >= KGEB_LAZY_ADAM) v = V4[off];
        opt_elem(a.kind, g.x, p.x, m.x, v.x, a);
        opt_elem(a.kind, g.y, p.y, m.y, v.y, a);
        opt_elem(a.kind, g.z, p.z, m.z, v.z, a);
        opt_elem(a.kind, g.w, p.w, m.w, v.w, a);
        P4[off] = p;
        if (a.kind != KGEB_SGD) M4[off] = m;
        if (a.kind >= KGEB_LAZY_ADAM) V4[off] = v;
        if (was_touched) G4[off] = f4zero();
      }
      if (was_touched && lane == 0) a.touched[row] = 0;
    }
  }
}

__global__ void __launch_bounds__(KGEB_THREADS) k_opt_update(OptArgs a) { opt_update_body(a, gridDim.x, blockIdx.x); }

// entity and relation table in ONE launch: the relation table of the bundled datasets has 11-237 rows, a launch of its own
// costs more in latency (8 us on C2) than in work.  The second table's blocks come FIRST in the grid so its short
// dependent chain starts at once and hides under the first table's streaming.
struct OptArgs2 { OptArgs t[2]; int nblocks[2]; };
__global__ void __launch_bounds__(KGEB_THREADS) k_opt_update2(OptArgs2 a) {
  const int b = blockIdx.x;
  const bool second = b < a.nblocks[1];
  const OptArgs my = second ? a.t[1] : a.t[0];   // one copy into registers: the body then reads no indexed constants
  opt_update_body(my, second ? a.nblocks[1] : a.nblocks[0], second ? b : b - a.nblocks[1]);
}

int fill_opt_args(kgeb_ctx* ctx, OptArgs& a, int* nblocks, float* P, float* G, uint8_t* touched, int64_t rows, int width, int ld,
                  float* state0, float* state1, int opt_kind, float lr, float beta1, float beta2, float eps, int64_t step) {
  KGEB_REQUIRE(ctx, P && G && touched && rows >= 0 && width > 0 && ld >= width, KGEB_ERR_INVALID, "opt_update: bad arguments");
  KGEB_REQUIRE(ctx, opt_kind >= KGEB_SGD && opt_kind <= KGEB_ADAM_DENSE, KGEB_ERR_INVALID, "unknown optimizer");
  KGEB_REQUIRE(ctx, opt_kind == KGEB_SGD || state0, KGEB_ERR_INVALID, "optimizer state0 missing");
  KGEB_REQUIRE(ctx, opt_kind < KGEB_LAZY_ADAM || (state1 && step >= 1), KGEB_ERR_INVALID, "Adam needs state1 and step >= 1");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && kgeb_aligned16(P) && kgeb_aligned16(G) &&
                        (!state0 || kgeb_aligned16(state0)) && (!state1 || kgeb_aligned16(state1)),
               KGEB_ERR_ALIGNMENT, "opt_update needs width, ld multiples of 4 floats and 16-byte aligned buffers");
  a.P = P; a.G = G; a.touched = touched; a.rows = rows; a.w4 = width / 4; a.ld4 = ld / 4; a.s0 = state0; a.s1 = state1;
  a.kind = opt_kind; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
  a.lr_t = lr;
  if (opt_kind >= KGEB_LAZY_ADAM) {
    const double t = (double)step;
    a.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  }
  // window: ~32 warps per SM worth of windows when the table is small, 32 rows per trip when it is huge
  int64_t win = rows / ((int64_t)ctx->num_sms * 32);
  a.win = win >= 32 ? 32 : (win >= 16 ? 16 : (win >= 8 ? 8 : (win >= 4 ? 4 : (win >= 2 ? 2 : 1))));
  int64_t blocks = (rows + (int64_t)a.win * KGEB_WARPS_PER_BLOCK - 1) / ((int64_t)a.win * KGEB_WARPS_PER_BLOCK);
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  *nblocks = (int)blocks;
  return KGEB_OK;
}

}  // namespace

extern "C" int kgeb_loss_fwd_bwd(kgeb_ctx* ctx, const float* t, int64_t B, const float* f, int64_t Bf, int K, int loss_kind,
                                 int pair_mode, float margin, float loss_weight, float* loss, float* dt, float* df,
                                 kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, t && f && loss && B > 0 && Bf > 0, KGEB_ERR_INVALID, "loss_fwd_bwd: bad arguments");
  KGEB_REQUIRE(ctx, loss_kind >= 0 && loss_kind <= KGEB_LOSS_KIND_MAX, KGEB_ERR_INVALID, "unknown loss kind");
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = ctx->num_sms * 4 < KGEB_MAX_PARTIAL_BLOCKS ? ctx->num_sms * 4 : KGEB_MAX_PARTIAL_BLOCKS;
  if (!KGEB_IS_PAIRWISE(loss_kind)) {
    int64_t blocks = (B + Bf + KGEB_THREADS - 1) / KGEB_THREADS;
    if (blocks > cap) blocks = cap;
    k_loss_pointwise<<<(int)blocks, KGEB_THREADS, 0, st>>>(t, B, f, Bf, loss_kind, margin, loss_weight, loss, dt, df,
                                                           ctx->d_partials, ctx->d_ticket);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  if (pair_mode == KGEB_PAIRED) {
    KGEB_REQUIRE(ctx, K > 0 && Bf == B * (int64_t)K, KGEB_ERR_INVALID, "PAIRED needs len(false) == len(true) * K");
    int64_t blocks = (B + KGEB_THREADS - 1) / KGEB_THREADS;
    if (blocks > cap) blocks = cap;
    k_loss_paired<<<(int)blocks, KGEB_THREADS, 0, st>>>(t, B, f, K, loss_kind, margin, loss_weight, loss, dt, df,
                                                        ctx->d_partials, ctx->d_ticket);
    KGEB_LAUNCH_OK(ctx);
    return KGEB_OK;
  }
  KGEB_REQUIRE(ctx, pair_mode == KGEB_ALLPAIRS, KGEB_ERR_INVALID, "unknown pair mode");
  {
    const int64_t tch = (B + SC - 1) / SC, fch = (Bf + SC - 1) / SC;
    const size_t need = (size_t)(tch + fch) * SC * sizeof(float) + (size_t)tch * (SC + 1) * sizeof(double);
    if (loss_kind == KGEB_PAIRWISE_HINGE && need <= KGEB_SORT_SCRATCH_BYTES && !ctx->env_allpairs_dense) {
      // sorted path (exact same counts as the m x m walk, see k_allpairs_hinge_sorted)
      double* tp = reinterpret_cast<double*>(ctx->d_sort_scratch);
      float* ts = reinterpret_cast<float*>(tp + tch * (SC + 1));
      float* fs = ts + tch * SC;
      double* loss_acc = ctx->d_partials;
      KGEB_CUDA_OK(ctx, cudaMemsetAsync(loss_acc, 0, sizeof(double), st));
      k_chunk_sort<<<(unsigned)(tch + fch), SC / 2, 0, st>>>(t, B, f, Bf, (int)tch, ts, fs, tp);
      KGEB_LAUNCH_OK(ctx);
      const int64_t nmax = B > Bf ? B : Bf;
      k_allpairs_hinge_sorted<<<dim3((unsigned)((nmax + 63) / 64), 2), 64, 0, st>>>(t, B, f, Bf, ts, fs, tp, margin, loss_weight, dt, df, loss_acc);
      KGEB_LAUNCH_OK(ctx);
      k_loss_allpairs_finish<<<1, 1, 0, st>>>(loss_acc, (double)loss_weight / ((double)B * (double)Bf), loss);
      KGEB_LAUNCH_OK(ctx);
      return KGEB_OK;
    }
  }
  if (dt) KGEB_CUDA_OK(ctx, cudaMemsetAsync(dt, 0, (size_t)B * sizeof(float), st));
  if (df) KGEB_CUDA_OK(ctx, cudaMemsetAsync(df, 0, (size_t)Bf * sizeof(float), st));
  double* loss_acc = ctx->d_partials;  // slot 0 doubles as the ALLPAIRS accumulator
  KGEB_CUDA_OK(ctx, cudaMemsetAsync(loss_acc, 0, sizeof(double), st));
  const int64_t nb = ((B > Bf ? B : Bf) + AP_TILE - 1) / AP_TILE;
  int slices = (int)((int64_t)ctx->num_sms * 8 / (2 * nb));
  if (slices < 1) slices = 1;
  if (slices > 256) slices = 256;
  dim3 grid((unsigned)nb, (unsigned)slices, 2);
  k_loss_allpairs<<<grid, AP_TILE, 0, st>>>(t, B, f, Bf, loss_kind, margin, loss_weight, dt, df, loss_acc);
  KGEB_LAUNCH_OK(ctx);
  k_loss_allpairs_finish<<<1, 1, 0, st>>>(loss_acc, (double)loss_weight / ((double)B * (double)Bf), loss);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

namespace {
// l3_reg (models.py:10-11): w * ||P||_3^3 over the whole table; its gradient 3 w p |p| touches every row.
__global__ void __launch_bounds__(KGEB_THREADS) k_l3_reg(const float* __restrict__ P, int64_t rows, int w4, int ld4, float w,
                                                         float* G, uint8_t* touched, float* loss, double* partials,
                                                         unsigned int* ticket) {
  const float4* P4 = reinterpret_cast<const float4*>(P);
  float4* G4 = reinterpret_cast<float4*>(G);
  double acc = 0.0;
  const int64_t total = rows * (int64_t)w4;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = x / w4;
    const int c = (int)(x - r * w4);
    const int64_t off = r * ld4 + c;
    const float4 p = P4[off];
    float4 g = G4[off];
    const float a0 = fabsf(p.x), a1 = fabsf(p.y), a2 = fabsf(p.z), a3 = fabsf(p.w);
    acc += (double)(a0 * a0 * a0) + (double)(a1 * a1 * a1) + (double)(a2 * a2 * a2) + (double)(a3 * a3 * a3);
    g.x += 3.f * w * p.x * a0; g.y += 3.f * w * p.y * a1; g.z += 3.f * w * p.z * a2; g.w += 3.f * w * p.w * a3;
    G4[off] = g;
    if (c == 0) touched[r] = 1;
  }
  // the last block writes w * sum into `loss` (a scratch slot here; the host wrapper adds it to the running loss)
  block_finalize_loss(acc, 0.0, partials, ticket, (double)w, 0.0, loss);
}
__global__ void k_add_scalar(float* dst, const float* src) { dst[0] += src[0]; }
}  // namespace

extern "C" int kgeb_l3_reg(kgeb_ctx* ctx, const float* P, int64_t rows, int width, int ld, float w, float* G, uint8_t* touched,
                           float* loss, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, P && G && touched && loss && rows > 0 && width > 0 && ld >= width, KGEB_ERR_INVALID, "l3_reg: bad arguments");
  KGEB_REQUIRE(ctx, width % 4 == 0 && ld % 4 == 0 && kgeb_aligned16(P) && kgeb_aligned16(G), KGEB_ERR_ALIGNMENT,
               "l3_reg needs width, ld multiples of 4 floats and 16-byte aligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (rows * (width / 4) + KGEB_THREADS - 1) / KGEB_THREADS;
  const int64_t cap = ctx->num_sms * 8 < KGEB_MAX_PARTIAL_BLOCKS ? ctx->num_sms * 8 : KGEB_MAX_PARTIAL_BLOCKS;
  if (blocks > cap) blocks = cap;
  float* tmp = reinterpret_cast<float*>(ctx->d_partials + 2 * KGEB_MAX_PARTIAL_BLOCKS - 2);  // last scratch slot
  k_l3_reg<<<(int)blocks, KGEB_THREADS, 0, st>>>(P, rows, width / 4, ld / 4, w, G, touched, tmp, ctx->d_partials, ctx->d_ticket);
  KGEB_LAUNCH_OK(ctx);
  k_add_scalar<<<1, 1, 0, st>>>(loss, tmp);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_opt_update(kgeb_ctx* ctx, float* P, float* G, uint8_t* touched, int64_t rows, int width, int ld,
                               float* state0, float* state1, int opt_kind, float lr, float beta1, float beta2, float eps,
                               int64_t step, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  OptArgs a;
  int blocks = 0;
  if (int rc = fill_opt_args(ctx, a, &blocks, P, G, touched, rows, width, ld, state0, state1, opt_kind, lr, beta1, beta2, eps, step)) return rc;
  if (rows == 0) return KGEB_OK;
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_OPT, (cudaStream_t)stream);
  k_opt_update<<<blocks, KGEB_THREADS, 0, (cudaStream_t)stream>>>(a);
  KGEB_TIME_END(ctx, KGEB_TIMED_OPT, (cudaStream_t)stream);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}

extern "C" int kgeb_opt_update2(kgeb_ctx* ctx, float* P0, float* G0, uint8_t* touched0, int64_t rows0, int width0, int ld0,
                                float* state00, float* state01, float* P1, float* G1, uint8_t* touched1, int64_t rows1,
                                int width1, int ld1, float* state10, float* state11, int opt_kind, float lr, float beta1,
                                float beta2, float eps, int64_t step, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  OptArgs2 a;
  if (int rc = fill_opt_args(ctx, a.t[0], &a.nblocks[0], P0, G0, touched0, rows0, width0, ld0, state00, state01, opt_kind, lr, beta1, beta2, eps, step)) return rc;
  if (int rc = fill_opt_args(ctx, a.t[1], &a.nblocks[1], P1, G1, touched1, rows1, width1, ld1, state10, state11, opt_kind, lr, beta1, beta2, eps, step)) return rc;
  if (rows0 == 0) a.nblocks[0] = 0;
  if (rows1 == 0) a.nblocks[1] = 0;
  const int gx = a.nblocks[0] + a.nblocks[1];
  if (gx == 0) return KGEB_OK;
  KGEB_TIME_BEGIN(ctx, KGEB_TIMED_OPT, (cudaStream_t)stream);
  k_opt_update2<<<gx, KGEB_THREADS, 0, (cudaStream_t)stream>>>(a);
  KGEB_TIME_END(ctx, KGEB_TIMED_OPT, (cudaStream_t)stream);
  KGEB_LAUNCH_OK(ctx);
  return KGEB_OK;
}
