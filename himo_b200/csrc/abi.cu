// himo_b200/csrc/abi.cu -- ABI version + status strings.
#include "common.cuh"
#include "himo_b200.h"

extern "C" int himo_abi_version(void) { return HIMO_B200_ABI_VERSION; }

extern "C" const char* himo_status_string(int status) {
  switch (status) {
    case HIMO_OK: return "ok";
    case HIMO_ERR_ARG: return "invalid argument";
    case HIMO_ERR_WORKSPACE: return "workspace too small";
    case HIMO_ERR_UNSUPPORTED: return "unsupported configuration";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown himo status";
}

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
extern "C" void himo_count_launch_(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long himo_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
