// api.cu -- context lifecycle of libkgeb200 (see include/kgeb200.h).
#include "common.cuh"
#include <new>
#include <stdlib.h>

char g_kgeb_create_err[512] = "";

extern "C" int kgeb_version(void) { return KGEB_VERSION; }

extern "C" int kgeb_create(kgeb_ctx** out, int device, int64_t num_entities, int64_t num_relations, int we, int wr, int ldE,
                           int ldR, int score_kind, float gamma) {
  if (!out) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: out is null");
  *out = nullptr;
  if (num_entities <= 0 || num_relations <= 0 || we <= 0 || wr <= 0 || ldE < we || ldR < wr)
    return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: bad table shape");
  if (num_entities > 0x7fffffffLL || num_relations > 0x7fffffffLL)
    return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: ids are int32");
  if (score_kind < KGEB_DISTMULT || score_kind > KGEB_HAKE)
    return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: unknown score kind");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return kgeb_fail(nullptr, KGEB_ERR_CUDA, "kgeb_create: no CUDA device (%s)", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: bad device index");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return kgeb_fail(nullptr, KGEB_ERR_CUDA, "kgeb_create: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return kgeb_fail(nullptr, KGEB_ERR_UNSUPPORTED, "%s: built for sm_100a only, device is sm_%lld%lld", "kgeb_create", prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return kgeb_fail(nullptr, KGEB_ERR_CUDA, "kgeb_create: %s", cudaGetErrorString(e));
  kgeb_ctx* c = new (std::nothrow) kgeb_ctx();
  if (!c) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "kgeb_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->device = device; c->N = num_entities; c->R = num_relations; c->we = we; c->wr = wr; c->ldE = ldE; c->ldR = ldR;
  c->kind = score_kind; c->gamma = gamma; c->param = 1.f; c->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("KGEB_RING_CTAS")) c->env_ring_ctas = atoi(e);
  if (const char* e = getenv("KGEB_RING_D")) c->env_ring_d = atoi(e);
  if (const char* e = getenv("KGEB_TRAIN_KERNEL")) c->env_force_reg = strcmp(e, "reg") == 0;
  if (const char* e = getenv("KGEB_HOLE")) c->hole_fft = strcmp(e, "fft") == 0;
  if (const char* e = getenv("KGEB_ALLPAIRS")) c->env_allpairs_dense = strcmp(e, "dense") == 0;
  c->env_pairs_grouped = -1;
  if (const char* e = getenv("KGEB_PAIRS_GROUPED")) c->env_pairs_grouped = atoi(e) != 0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) c->ring_cfg[i][j].smem = -1;
  if ((e = cudaMalloc(&c->d_partials, sizeof(double) * 2 * KGEB_MAX_PARTIAL_BLOCKS)) != cudaSuccess ||
      (e = cudaMalloc(&c->d_ticket, sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaMalloc(&c->d_sort_scratch, KGEB_SORT_SCRATCH_BYTES)) != cudaSuccess ||
      (e = cudaMemset(c->d_ticket, 0, sizeof(unsigned int))) != cudaSuccess) {
    kgeb_fail(nullptr, KGEB_ERR_CUDA, "kgeb_create: %s", cudaGetErrorString(e));
    if (c->d_partials) cudaFree(c->d_partials);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->d_sort_scratch) cudaFree(c->d_sort_scratch);
    delete c;
    return KGEB_ERR_CUDA;
  }
  *out = c;
  return KGEB_OK;
}

extern "C" int kgeb_destroy(kgeb_ctx* ctx) {
  if (!ctx) return KGEB_ERR_INVALID;
  cudaSetDevice(ctx->device);
  kgeb_internal_comm_release(ctx);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  if (ctx->d_sort_scratch) cudaFree(ctx->d_sort_scratch);
  for (int i = 0; i < KGEB_TIMED_KERNELS; ++i)
    for (int p = 0; p < KGEB_TIMED_POOL; ++p)
      for (int j = 0; j < 2; ++j)
        if (ctx->ev[i][p][j]) cudaEventDestroy(ctx->ev[i][p][j]);
  delete ctx;
  return KGEB_OK;
}

extern "C" const char* kgeb_last_error(const kgeb_ctx* ctx) { return ctx ? ctx->err : g_kgeb_create_err; }
extern "C" int64_t kgeb_launch_count(const kgeb_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int kgeb_num_sms(const kgeb_ctx* ctx) { return ctx ? ctx->num_sms : KGEB_ERR_INVALID; }

extern "C" int kgeb_set_kernel_timing(kgeb_ctx* ctx, int on) {
  if (!ctx) return KGEB_ERR_INVALID;
  if (on) {
    for (int i = 0; i < KGEB_TIMED_KERNELS; ++i)
      for (int p = 0; p < KGEB_TIMED_POOL; ++p)
        for (int j = 0; j < 2; ++j)
          if (!ctx->ev[i][p][j]) KGEB_CUDA_OK(ctx, cudaEventCreate(&ctx->ev[i][p][j]));
  }
  for (int i = 0; i < KGEB_TIMED_KERNELS; ++i) ctx->ev_n[i] = 0;
  ctx->timing = on ? 1 : 0;
  return KGEB_OK;
}

extern "C" int kgeb_last_kernel_ms(kgeb_ctx* ctx, int which, float* ms) {
  if (!ctx || !ms || which < 0 || which >= KGEB_TIMED_KERNELS) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, ctx->ev_n[which] > 0, KGEB_ERR_INVALID, "kernel timing: that kernel has not been launched with timing on");
  const int p = (ctx->ev_n[which] - 1) % KGEB_TIMED_POOL;
  KGEB_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev[which][p][1]));
  KGEB_CUDA_OK(ctx, cudaEventElapsedTime(ms, ctx->ev[which][p][0], ctx->ev[which][p][1]));
  return KGEB_OK;
}

extern "C" int kgeb_kernel_ms_sum(kgeb_ctx* ctx, int which, float* ms_sum, int* launches) {
  if (!ctx || !ms_sum || !launches || which < 0 || which >= KGEB_TIMED_KERNELS) return KGEB_ERR_INVALID;
  const int n = ctx->ev_n[which] < KGEB_TIMED_POOL ? ctx->ev_n[which] : KGEB_TIMED_POOL;   // older launches fell out of the ring
  float total = 0.f;
  for (int i = 0; i < n; ++i) {
    const int p = (ctx->ev_n[which] - 1 - i) % KGEB_TIMED_POOL;
    float ms = 0.f;
    KGEB_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev[which][p][1]));
    KGEB_CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev[which][p][0], ctx->ev[which][p][1]));
    total += ms;
  }
  *ms_sum = total; *launches = n;
  ctx->ev_n[which] = 0;
  return KGEB_OK;
}

extern "C" int kgeb_set_hole_mode(kgeb_ctx* ctx, int mode) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, mode == KGEB_HOLE_DIRECT || mode == KGEB_HOLE_FFT, KGEB_ERR_INVALID, "set_hole_mode: unknown mode");
  ctx->hole_fft = mode == KGEB_HOLE_FFT;
  return KGEB_OK;
}

extern "C" int kgeb_set_hake_params(kgeb_ctx* ctx, float phase_div, float phase_weight, float mod_weight) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, ctx->kind == KGEB_HAKE, KGEB_ERR_INVALID, "set_hake_params: the context is not a HAKE context");
  KGEB_REQUIRE(ctx, phase_div > 0.f, KGEB_ERR_INVALID, "set_hake_params: phase_div = embedding_range / pi must be positive");
  ctx->hake_ps = phase_div; ctx->hake_pw = phase_weight; ctx->hake_mw = mod_weight;
  return KGEB_OK;
}

extern "C" int kgeb_set_hake_call_rows(kgeb_ctx* ctx, int64_t rows) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, ctx->kind == KGEB_HAKE && rows >= 0, KGEB_ERR_INVALID, "set_hake_call_rows: HAKE context and rows >= 0 expected");
  ctx->hake_group = rows;
  return KGEB_OK;
}

extern "C" int kgeb_set_score_param(kgeb_ctx* ctx, float param) {
  if (!ctx) return KGEB_ERR_INVALID;
  ctx->param = param;
  return KGEB_OK;
}

extern "C" int kgeb_set_dropout(kgeb_ctx* ctx, float p, uint64_t seed, uint64_t offset, int stream_id) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, p >= 0.f && p < 1.f, KGEB_ERR_INVALID, "dropout rate must be in [0, 1)");
  KGEB_REQUIRE(ctx, stream_id == 0 || stream_id == 1, KGEB_ERR_INVALID, "dropout stream id is 0 (positive call) or 1 (negative call)");
  const double thr = (double)p * 65536.0 + 0.5;
  ctx->dp_thr = (uint32_t)thr;
  ctx->dp_scale = 1.0f / (1.0f - p);
  ctx->dp_seed = seed; ctx->dp_offset = offset; ctx->dp_stream = stream_id;
  return KGEB_OK;
}
