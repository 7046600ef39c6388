// capi.cu -- library-wide bookkeeping behind the C ABI (include/golf_b200.h).
#include <atomic>

#include "common.cuh"

namespace golf {
static std::atomic<uint64_t> g_launches{0};
static thread_local int g_last_cuda = 0;
std::atomic<int> g_pdl{1};

void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int note_cuda(cudaError_t e) {
  if (e == cudaSuccess) return GOLF_OK;
  g_last_cuda = (int)e;
  return GOLF_ERR_CUDA;
}
}  // namespace golf

GOLF_API int golf_abi_version(void) { return GOLF_B200_ABI_VERSION; }

GOLF_API const char* golf_strerror(int code) {
  switch (code) {
    case GOLF_OK: return "ok";
    case GOLF_ERR_INVALID: return "invalid argument (shape, null pointer or alignment)";
    case GOLF_ERR_UNSUPPORTED: return "configuration not supported by the compiled kernels";
    case GOLF_ERR_WORKSPACE: return "workspace missing or too small";
    case GOLF_ERR_CUDA: return "CUDA runtime error (see golf_last_cuda_error)";
  }
  return "unknown golf_b200 error code";
}

GOLF_API int golf_last_cuda_error(void) { return golf::g_last_cuda; }
GOLF_API uint64_t golf_launch_count(void) { return golf::g_launches.load(std::memory_order_relaxed); }

GOLF_API void golf_set_pdl(int on) { golf::g_pdl = on ? 1 : 0; }
