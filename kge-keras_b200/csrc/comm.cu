// comm.cu -- the collectives of the two sharded paths behind the C ABI (kgeb_comm_init / kgeb_allreduce_counts /
// kgeb_alltoall_rows), so that a host WITHOUT torch.distributed (the TF op shim, a C++ server) can shard:
//   * 1-N filtered ranking, entity rows split over the GPUs: one in-place sum of the int32 partial rank counts;
//   * row-sharded training (the 1M-entity config): requested rows travel owners -> requesters and gradient rows back, as grouped
//     ncclSend / ncclRecv (NCCL has no all-to-all primitive), the replicated relation gradient is summed, touched flags are max-ed.
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library carries no link-time dependency, a process that already loaded an
// NCCL (PyTorch's) shares that copy, and a box without NCCL can still use every single-GPU entry point.  Everything is enqueued on
// the caller's stream; nothing synchronises.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*CommAbort)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
};
NcclApi g_nccl = {};
char g_nccl_err[256] = "";

bool nccl_load() {
  if (g_nccl.handle) return true;
  const char* names[] = {getenv("KGEB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { snprintf(g_nccl_err, sizeof g_nccl_err, "cannot load NCCL (libnccl.so.2): %s", dlerror()); return false; }
#define KGEB_SYM(field, name)                                                                             \
  do {                                                                                                    \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                                            \
    if (!g_nccl.field) { snprintf(g_nccl_err, sizeof g_nccl_err, "NCCL symbol %s is missing", name); dlclose(h); return false; } \
  } while (0)
  KGEB_SYM(GetUniqueId, "ncclGetUniqueId");
  KGEB_SYM(CommInitRank, "ncclCommInitRank");
  KGEB_SYM(CommDestroy, "ncclCommDestroy");
  KGEB_SYM(CommAbort, "ncclCommAbort");
  KGEB_SYM(AllReduce, "ncclAllReduce");
  KGEB_SYM(Send, "ncclSend");
  KGEB_SYM(Recv, "ncclRecv");
  KGEB_SYM(GroupStart, "ncclGroupStart");
  KGEB_SYM(GroupEnd, "ncclGroupEnd");
  KGEB_SYM(GetErrorString, "ncclGetErrorString");
  KGEB_SYM(GetVersion, "ncclGetVersion");
#undef KGEB_SYM
  g_nccl.handle = h;
  return true;
}

#define KGEB_NCCL_OK(ctx, expr)                                                                           \
  do {                                                                                                    \
    ncclResult_t r_ = (expr);                                                                             \
    if (r_ != ncclSuccess) return kgeb_fail((ctx), KGEB_ERR_CUDA, "NCCL: %s (line %lld)", g_nccl.GetErrorString(r_), __LINE__); \
  } while (0)

inline ncclComm_t comm_of(const kgeb_ctx* ctx) { return static_cast<ncclComm_t>(ctx->comm); }

}  // namespace

static_assert(KGEB_COMM_ID_BYTES == sizeof(ncclUniqueId), "kgeb200.h states the size of an NCCL unique id");

extern "C" int kgeb_comm_unique_id(void* id_out) {
  if (!id_out) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "comm_unique_id: null output");
  if (!nccl_load()) return kgeb_fail(nullptr, KGEB_ERR_UNSUPPORTED, "%s", g_nccl_err);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return kgeb_fail(nullptr, KGEB_ERR_CUDA, "NCCL: %s", g_nccl.GetErrorString(r));
  memcpy(id_out, &id, sizeof id);
  return KGEB_OK;
}

extern "C" int kgeb_comm_version(void) {
  int v = 0;
  if (!nccl_load() || g_nccl.GetVersion(&v) != ncclSuccess) return KGEB_ERR_UNSUPPORTED;
  return v;
}

extern "C" int kgeb_comm_init(kgeb_ctx* ctx, const void* id, int rank, int world) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_REQUIRE(ctx, id && world >= 1 && rank >= 0 && rank < world, KGEB_ERR_INVALID, "comm_init: bad rank / world / id");
  KGEB_REQUIRE(ctx, ctx->comm == nullptr, KGEB_ERR_INVALID, "comm_init: the context already has a communicator (kgeb_comm_destroy first)");
  if (!nccl_load()) return kgeb_fail(ctx, KGEB_ERR_UNSUPPORTED, "%s", g_nccl_err);
  KGEB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  KGEB_NCCL_OK(ctx, g_nccl.CommInitRank(&c, world, uid, rank));
  ctx->comm = c; ctx->comm_rank = rank; ctx->comm_world = world;
  return KGEB_OK;
}

extern "C" int kgeb_comm_destroy(kgeb_ctx* ctx) {
  if (!ctx) return KGEB_ERR_INVALID;
  if (ctx->comm) {
    KGEB_NCCL_OK(ctx, g_nccl.CommDestroy(comm_of(ctx)));
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 0;
  }
  return KGEB_OK;
}
// kgeb_destroy's hook (api.cu): never fails and never blocks.  ncclCommDestroy waits for every CUDA graph that captured a
// collective of the communicator to be destroyed first; a context torn down at interpreter exit cannot order that, so the
// implicit release aborts the communicator instead (kgeb_comm_destroy is the orderly path).
void kgeb_internal_comm_release(kgeb_ctx* ctx) {
  if (ctx->comm && g_nccl.handle) g_nccl.CommAbort(comm_of(ctx));
  ctx->comm = nullptr;
}

extern "C" int kgeb_comm_info(const kgeb_ctx* ctx, int* rank, int* world) {
  if (!ctx || !rank || !world) return KGEB_ERR_INVALID;
  *rank = ctx->comm ? ctx->comm_rank : 0;
  *world = ctx->comm ? ctx->comm_world : 1;
  return KGEB_OK;
}

#define KGEB_NEED_COMM(ctx, what) KGEB_REQUIRE(ctx, (ctx)->comm != nullptr, KGEB_ERR_INVALID, what ": call kgeb_comm_init first")

// counts[n] <- sum over ranks, in place (integers: the summed ranks are identical for every partition of the rows)
extern "C" int kgeb_allreduce_counts(kgeb_ctx* ctx, int32_t* counts, int64_t n, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_NEED_COMM(ctx, "allreduce_counts");
  KGEB_REQUIRE(ctx, counts && n >= 0, KGEB_ERR_INVALID, "allreduce_counts: bad arguments");
  if (n == 0) return KGEB_OK;
  KGEB_NCCL_OK(ctx, g_nccl.AllReduce(counts, counts, (size_t)n, ncclInt32, ncclSum, comm_of(ctx), (cudaStream_t)stream));
  ctx->launches++;
  return KGEB_OK;
}

// buf[n] <- sum over ranks (the replicated relation gradient of the row-sharded step, the loss)
extern "C" int kgeb_allreduce_f32(kgeb_ctx* ctx, float* buf, int64_t n, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_NEED_COMM(ctx, "allreduce_f32");
  KGEB_REQUIRE(ctx, buf && n >= 0, KGEB_ERR_INVALID, "allreduce_f32: bad arguments");
  if (n == 0) return KGEB_OK;
  KGEB_NCCL_OK(ctx, g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, comm_of(ctx), (cudaStream_t)stream));
  ctx->launches++;
  return KGEB_OK;
}

// flags[n] <- max over ranks (touched flags of the replicated relation table)
extern "C" int kgeb_allreduce_flags(kgeb_ctx* ctx, uint8_t* flags, int64_t n, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_NEED_COMM(ctx, "allreduce_flags");
  KGEB_REQUIRE(ctx, flags && n >= 0, KGEB_ERR_INVALID, "allreduce_flags: bad arguments");
  if (n == 0) return KGEB_OK;
  KGEB_NCCL_OK(ctx, g_nccl.AllReduce(flags, flags, (size_t)n, ncclUint8, ncclMax, comm_of(ctx), (cudaStream_t)stream));
  ctx->launches++;
  return KGEB_OK;
}

// rows of `width` floats: send holds send_rows[g] rows for every rank g, concatenated in rank order; recv receives
// recv_rows[g] rows from every rank g in rank order.  send_rows / recv_rows are HOST arrays of length world.
extern "C" int kgeb_alltoall_rows(kgeb_ctx* ctx, const float* send, const int64_t* send_rows, float* recv, const int64_t* recv_rows,
                                  int width, kgeb_stream stream) {
  if (!ctx) return KGEB_ERR_INVALID;
  KGEB_NEED_COMM(ctx, "alltoall_rows");
  KGEB_REQUIRE(ctx, send_rows && recv_rows && width > 0, KGEB_ERR_INVALID, "alltoall_rows: bad arguments");
  const int world = ctx->comm_world, me = ctx->comm_rank;
  int64_t so = 0, ro = 0, self_so = -1, self_ro = -1;
  for (int g = 0; g < world; ++g) {
    KGEB_REQUIRE(ctx, send_rows[g] >= 0 && recv_rows[g] >= 0, KGEB_ERR_INVALID, "alltoall_rows: negative row count");
    if (g == me) { self_so = so; self_ro = ro; }
    so += send_rows[g]; ro += recv_rows[g];
  }
  KGEB_REQUIRE(ctx, (so == 0 || send) && (ro == 0 || recv), KGEB_ERR_INVALID, "alltoall_rows: null buffer");
  KGEB_REQUIRE(ctx, send_rows[me] == recv_rows[me], KGEB_ERR_INVALID, "alltoall_rows: a rank sends itself as many rows as it receives from itself");
  cudaStream_t st = (cudaStream_t)stream;
  // the rank's own rows never touch NCCL
  if (send_rows[me] > 0)
    KGEB_CUDA_OK(ctx, cudaMemcpyAsync(recv + self_ro * width, send + self_so * width, (size_t)send_rows[me] * width * 4, cudaMemcpyDeviceToDevice, st));
  if (world > 1) {
    KGEB_NCCL_OK(ctx, g_nccl.GroupStart());
    so = 0; ro = 0;
    ncclResult_t bad = ncclSuccess;
    for (int g = 0; g < world; ++g) {
      if (g != me) {
        ncclResult_t r = ncclSuccess;
        if (send_rows[g] > 0) r = g_nccl.Send(send + so * width, (size_t)send_rows[g] * width, ncclFloat32, g, comm_of(ctx), st);
        if (r == ncclSuccess && recv_rows[g] > 0) r = g_nccl.Recv(recv + ro * width, (size_t)recv_rows[g] * width, ncclFloat32, g, comm_of(ctx), st);
        if (r != ncclSuccess && bad == ncclSuccess) bad = r;
      }
      so += send_rows[g]; ro += recv_rows[g];
    }
    ncclResult_t end = g_nccl.GroupEnd();
    KGEB_NCCL_OK(ctx, bad);
    KGEB_NCCL_OK(ctx, end);
  }
  ctx->launches++;
  return KGEB_OK;
}
