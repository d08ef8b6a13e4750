// liblkb200 - the one exchange step of the path: an NCCL all-gather of the per-rank power blocks.
//
// One process per GPU.  NCCL is bound at RUN TIME (dlopen of libnccl.so.2 - inside a PyTorch process that is the
// copy torch already loaded, otherwise the system one), so the library has no link-time dependency on it and every
// single-GPU entry point works on a box without NCCL.  The communicator is bootstrapped the standard way: rank 0
// calls lkb_nccl_unique_id, the 128-byte id travels to the other ranks over whatever side channel the host
// program has (torch.distributed's store, MPI, a pipe), then every rank calls lkb_nccl_init.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace lkb {
namespace {

// the slice of nccl.h this file needs (stable NCCL 2.x ABI)
struct NcclUniqueId { char internal[LKB_NCCL_ID_BYTES]; };
typedef void* NcclComm;
constexpr int kNcclFloat32 = 7;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

std::mutex g_comm_mu;
NcclApi g_nccl;
NcclComm g_comm = nullptr;
int g_rank = -1, g_world = 0;

int nccl_load() {
  if (g_nccl.handle) return LKB_OK;
  const char* names[] = {getenv("LKB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (n == nullptr || *n == 0) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("NCCL not found (tried $LKB_NCCL_LIB, libnccl.so.2, libnccl.so): %s", dlerror());
    return LKB_E_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
    set_error("the NCCL library lacks a required symbol");
    dlclose(h);
    return LKB_E_UNSUPPORTED;
  }
  g_nccl = a;
  return LKB_OK;
}

#define LKB_NCCL_CHECK(expr)                                                                     \
  do {                                                                                           \
    int _r = (expr);                                                                             \
    if (_r != 0) {                                                                               \
      set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
      return LKB_E_NCCL;                                                                         \
    }                                                                                            \
  } while (0)

}  // namespace
}  // namespace lkb

using namespace lkb;

extern "C" {

int lkb_nccl_version(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  if (nccl_load() != LKB_OK || !g_nccl.GetVersion) return 0;
  int v = 0;
  return g_nccl.GetVersion(&v) == 0 ? v : 0;
}

int lkb_nccl_unique_id(void* id_out) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  LKB_REQUIRE(id_out != nullptr, "lkb_nccl_unique_id: id_out is NULL");
  LKB_TRY(nccl_load());
  NcclUniqueId id;
  memset(&id, 0, sizeof(id));
  LKB_NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return LKB_OK;
}

int lkb_nccl_init(int rank, int world_size, const void* id) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  LKB_REQUIRE(id != nullptr, "lkb_nccl_init: id is NULL");
  LKB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "lkb_nccl_init: rank outside [0, world_size)");
  LKB_REQUIRE(g_comm == nullptr, "lkb_nccl_init: a communicator already exists (call lkb_nccl_shutdown first)");
  LKB_TRY(ensure_device());                       // the communicator binds to the device lkb_init selected
  LKB_TRY(nccl_load());
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NcclComm comm = nullptr;
  LKB_NCCL_CHECK(g_nccl.CommInitRank(&comm, world_size, uid, rank));
  g_comm = comm;
  g_rank = rank;
  g_world = world_size;
  return LKB_OK;
}

int lkb_nccl_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  if (g_comm == nullptr) return LKB_OK;
  cudaDeviceSynchronize();
  NcclComm c = g_comm;
  g_comm = nullptr;
  g_rank = -1;
  g_world = 0;
  LKB_NCCL_CHECK(g_nccl.CommDestroy(c));
  return LKB_OK;
}

int lkb_nccl_rank(void) { return g_rank; }
int lkb_nccl_world_size(void) { return g_world; }

int lkb_allgather_f32(const float* local, int64_t n_local, float* global, void* stream) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  LKB_REQUIRE(g_comm != nullptr, "lkb_allgather_f32: no communicator (call lkb_nccl_init on every rank first)");
  LKB_REQUIRE(n_local >= 0, "lkb_allgather_f32: n_local < 0");
  if (n_local == 0) return LKB_OK;
  LKB_REQUIRE(local != nullptr && global != nullptr, "lkb_allgather_f32: NULL buffer");
  LKB_NCCL_CHECK(g_nccl.AllGather(local, global, (size_t)n_local, kNcclFloat32, g_comm, (cudaStream_t)stream));
  g_launches++;
  return LKB_OK;
}

}  // extern "C"
