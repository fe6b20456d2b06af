// peer.cu -- one table, many GPUs: a row-sharded embedding table (and its gradient buffer) mapped at ONE contiguous virtual
// address range on every GPU of an NVLink / NVSwitch box (the 1M-entity config).
//
// Shard g -- rows [g * rows_per_shard, (g + 1) * rows_per_shard) -- is physical memory of GPU g (cuMemCreate); every rank imports
// the other shards through POSIX file-descriptor handles and maps all of them back to back (cuMemAddressReserve + cuMemMap), so
// `base + row * ld` is the address of a row wherever it lives.  The training kernels then run UNCHANGED on global ids: the
// bulk-async row gathers of k_train_ring become NVLink peer loads for remote rows and its `red.global.add.v4.f32` gradient rows
// become NVLink peer reductions resolved in the owner's L2 -- the row exchange and the gradient return of the sharded step
// happen inside the scoring kernel, row by row, overlapped with the arithmetic, instead of as two NCCL all-to-alls around it.
//
// The driver API is reached through cudaGetDriverEntryPoint (no link-time dependency on libcuda: the library still loads on a
// box without a driver).  File descriptors travel between the ranks' processes by whatever the host has (the Python host sends
// them over UNIX sockets, KGEkeras/_peer.py).
#include <cuda.h>
#include <unistd.h>

#include <new>

#include "common.cuh"

struct kgeb_peer_table {
  int device, rank, world;
  size_t shard_bytes;            // rounded up to the allocation granularity
  CUdeviceptr base;              // reserved range of world * shard_bytes
  CUmemGenericAllocationHandle* handles;   // [world]; 0 = not there yet
  unsigned char* mapped;         // [world]
  int own_fd;
  char err[256];
};

namespace {

struct DriverApi {
  bool ok;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*GetErrorString)(CUresult, const char**);
};
DriverApi g_drv = {};
char g_drv_err[256] = "";

bool drv_load() {
  if (g_drv.ok) return true;
#define KGEB_DRV(field, name)                                                                                                   \
  do {                                                                                                                          \
    cudaDriverEntryPointQueryResult q;                                                                                          \
    void* fn = nullptr;                                                                                                         \
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);                                                  \
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {                                                          \
      snprintf(g_drv_err, sizeof g_drv_err, "driver entry point %s is not available (%s)", name, cudaGetErrorString(e));        \
      return false;                                                                                                             \
    }                                                                                                                           \
    *reinterpret_cast<void**>(&g_drv.field) = fn;                                                                               \
  } while (0)
  KGEB_DRV(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
  KGEB_DRV(MemCreate, "cuMemCreate");
  KGEB_DRV(MemRelease, "cuMemRelease");
  KGEB_DRV(MemAddressReserve, "cuMemAddressReserve");
  KGEB_DRV(MemAddressFree, "cuMemAddressFree");
  KGEB_DRV(MemMap, "cuMemMap");
  KGEB_DRV(MemUnmap, "cuMemUnmap");
  KGEB_DRV(MemSetAccess, "cuMemSetAccess");
  KGEB_DRV(MemExportToShareableHandle, "cuMemExportToShareableHandle");
  KGEB_DRV(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
  KGEB_DRV(GetErrorString, "cuGetErrorString");
#undef KGEB_DRV
  g_drv.ok = true;
  return true;
}

int peer_fail(kgeb_peer_table* t, int code, const char* what, CUresult r) {
  const char* s = nullptr;
  if (g_drv.ok) g_drv.GetErrorString(r, &s);
  char* dst = t ? t->err : g_kgeb_create_err;
  snprintf(dst, t ? sizeof t->err : 512, "%s: %s", what, s ? s : "driver error");
  return code;
}
#define KGEB_DRV_OK(t, what, expr)                                          \
  do {                                                                      \
    CUresult r_ = (expr);                                                   \
    if (r_ != CUDA_SUCCESS) return peer_fail((t), KGEB_ERR_CUDA, (what), r_); \
  } while (0)

CUmemAllocationProp shard_prop(int device) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof p);
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

}  // namespace

extern "C" int64_t kgeb_peer_granularity(int device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return kgeb_fail(nullptr, KGEB_ERR_CUDA, "%s", "peer_granularity: no such CUDA device");
  if (!drv_load()) return kgeb_fail(nullptr, KGEB_ERR_UNSUPPORTED, "%s", g_drv_err);
  const CUmemAllocationProp p = shard_prop(device);
  size_t g = 0;
  if (g_drv.MemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || g == 0)
    return kgeb_fail(nullptr, KGEB_ERR_CUDA, "%s", "cuMemGetAllocationGranularity failed");
  return (int64_t)g;
}

extern "C" int kgeb_peer_create(kgeb_peer_table** out, int device, int rank, int world, int64_t shard_bytes) {
  if (!out) return KGEB_ERR_INVALID;
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || shard_bytes <= 0) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "peer_create: bad rank / world / size");
  const int64_t gran = kgeb_peer_granularity(device);
  if (gran < 0) return (int)gran;
  kgeb_peer_table* t = new (std::nothrow) kgeb_peer_table();
  if (!t) return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "peer_create: out of host memory");
  memset(t, 0, sizeof *t);
  t->device = device; t->rank = rank; t->world = world; t->own_fd = -1;
  t->shard_bytes = ((size_t)shard_bytes + (size_t)gran - 1) / (size_t)gran * (size_t)gran;
  t->handles = new (std::nothrow) CUmemGenericAllocationHandle[world]();
  t->mapped = new (std::nothrow) unsigned char[world]();
  if (!t->handles || !t->mapped) { delete[] t->handles; delete[] t->mapped; delete t; return kgeb_fail(nullptr, KGEB_ERR_INVALID, "%s", "peer_create: out of host memory"); }
  const CUmemAllocationProp p = shard_prop(device);
  CUresult r = g_drv.MemCreate(&t->handles[rank], t->shard_bytes, &p, 0);
  if (r == CUDA_SUCCESS) r = g_drv.MemAddressReserve(&t->base, t->shard_bytes * (size_t)world, (size_t)gran, 0, 0);
  if (r == CUDA_SUCCESS) r = g_drv.MemMap(t->base + t->shard_bytes * (size_t)rank, t->shard_bytes, 0, t->handles[rank], 0);
  if (r == CUDA_SUCCESS) {
    t->mapped[rank] = 1;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = g_drv.MemSetAccess(t->base + t->shard_bytes * (size_t)rank, t->shard_bytes, &acc, 1);
  }
  if (r != CUDA_SUCCESS) {
    peer_fail(nullptr, KGEB_ERR_CUDA, "peer_create (cuMemCreate / reserve / map)", r);
    kgeb_peer_destroy(t);
    return KGEB_ERR_CUDA;
  }
  *out = t;
  return KGEB_OK;
}

extern "C" const char* kgeb_peer_last_error(const kgeb_peer_table* t) { return t ? t->err : g_kgeb_create_err; }
extern "C" int64_t kgeb_peer_shard_bytes(const kgeb_peer_table* t) { return t ? (int64_t)t->shard_bytes : KGEB_ERR_INVALID; }
extern "C" void* kgeb_peer_base(const kgeb_peer_table* t) { return t ? reinterpret_cast<void*>(t->base) : nullptr; }

// a file descriptor of this rank's shard (created on first use, owned by the table); send it to the other ranks' processes
extern "C" int kgeb_peer_export_fd(kgeb_peer_table* t, int* fd) {
  if (!t || !fd) return KGEB_ERR_INVALID;
  if (t->own_fd < 0) {
    int f = -1;
    KGEB_DRV_OK(t, "cuMemExportToShareableHandle", g_drv.MemExportToShareableHandle(&f, t->handles[t->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    t->own_fd = f;
  }
  *fd = t->own_fd;
  return KGEB_OK;
}

// maps the shard of `src_rank`, received as a file descriptor valid in THIS process, at its slot; the descriptor stays the caller's
extern "C" int kgeb_peer_import_fd(kgeb_peer_table* t, int src_rank, int fd) {
  if (!t) return KGEB_ERR_INVALID;
  if (src_rank < 0 || src_rank >= t->world || src_rank == t->rank || fd < 0 || t->mapped[src_rank]) {
    snprintf(t->err, sizeof t->err, "peer_import_fd: bad source rank %d / descriptor, or that shard is already mapped", src_rank);
    return KGEB_ERR_INVALID;
  }
  KGEB_DRV_OK(t, "cuMemImportFromShareableHandle",
              g_drv.MemImportFromShareableHandle(&t->handles[src_rank], (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  const CUdeviceptr at = t->base + t->shard_bytes * (size_t)src_rank;
  KGEB_DRV_OK(t, "cuMemMap (peer shard)", g_drv.MemMap(at, t->shard_bytes, 0, t->handles[src_rank], 0));
  t->mapped[src_rank] = 1;
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = t->device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  KGEB_DRV_OK(t, "cuMemSetAccess (peer shard: are the GPUs NVLink / P2P connected?)", g_drv.MemSetAccess(at, t->shard_bytes, &acc, 1));
  return KGEB_OK;
}

// 1 when every shard is mapped (the base pointer may then be used as one table)
extern "C" int kgeb_peer_complete(const kgeb_peer_table* t) {
  if (!t) return KGEB_ERR_INVALID;
  for (int g = 0; g < t->world; ++g)
    if (!t->mapped[g]) return 0;
  return 1;
}

extern "C" int kgeb_peer_destroy(kgeb_peer_table* t) {
  if (!t) return KGEB_ERR_INVALID;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  if (g_drv.ok) {
    for (int g = 0; g < t->world; ++g) {
      if (t->mapped && t->mapped[g]) g_drv.MemUnmap(t->base + t->shard_bytes * (size_t)g, t->shard_bytes);
      if (t->handles && t->handles[g]) g_drv.MemRelease(t->handles[g]);
    }
    if (t->base) g_drv.MemAddressFree(t->base, t->shard_bytes * (size_t)t->world);
  }
  if (t->own_fd >= 0) close(t->own_fd);
  delete[] t->handles;
  delete[] t->mapped;
  delete t;
  return KGEB_OK;
}
