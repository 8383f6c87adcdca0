// nccl_dl.cpp — run-time binding to NCCL (dlopen), see emvs_internal.h.
#include "emvs_internal.h"

#include <dlfcn.h>
#include <mutex>

namespace emvs {

namespace {
NcclApi g_api;
bool g_ok = false;
std::once_flag g_once;

void* open_nccl()
{
  // Prefer an NCCL that is already mapped in this process (e.g. the one torch bundles), so
  // that only one NCCL runtime is alive; otherwise load the system library.
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL)) return h;
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) return h;
  return nullptr;
}

void load()
{
  void* h = open_nccl();
  if (!h) return;
#define BIND(field, sym)                                   \
  *(void**)(&g_api.field) = dlsym(h, sym);                 \
  if (!g_api.field) return;
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllReduce, "ncclAllReduce")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_ok = true;
}
}  // namespace

const NcclApi* nccl_api()
{
  std::call_once(g_once, load);
  if (!g_ok) {
    set_error("NCCL not available: %s", dlerror() ? dlerror() : "libnccl.so.2 could not be loaded");
    return nullptr;
  }
  return &g_api;
}

}  // namespace emvs
