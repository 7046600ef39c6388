// common.cuh -- shared device helpers for libgolf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/golf_b200.h"

#define GOLF_API extern "C" __attribute__((visibility("default")))

namespace golf {

// ---- host-side bookkeeping (capi.cu) -------------------------------------------
void note_launch(int n = 1);
int launch_frame_reduction(const float* g, const float* y, int64_t y_stride, float* d_a, int B, int L, int F, int M, int hop,
                           cudaStream_t st);  // lpc_ss.cu
int launch_gain_reduction(const float* u, const float* ex, int64_t ex_stride, float* d_gain, int B, int L, int F, int hop,
                          cudaStream_t st);  // lpc_ss.cu
int note_cuda(cudaError_t e);  // records e, returns GOLF_ERR_CUDA if e != cudaSuccess else 0

#define GOLF_CHECK_LAUNCH()                              \
  do {                                                   \
    golf::note_launch();                                 \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) return golf::note_cuda(_e);   \
  } while (0)

#define GOLF_CUDA(call)                                  \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return golf::note_cuda(_e);   \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may have its CTAs scheduled as
// soon as every CTA of the previous kernel in the stream has called pdl_trigger() (or exited); it runs
// whatever does not depend on that kernel and blocks in pdl_wait() until it has completed and its writes
// are visible.  Both are no-ops in a kernel launched the ordinary way.  Used for ONE edge: knot prefix ->
// flow kernel, whose long prologue (polyphase taps, three interpolated table rows) then runs beside the
// scan.  Measured on the other edges of the decoder chain (wait at the top of the dependent kernel) it
// gained nothing: early-launched CTAs only hold registers while they wait (profiles/README.md).
// golf_set_pdl(0) turns the attribute off.
extern std::atomic<int> g_pdl;  // process-wide switches are relaxed atomics: host threads may flip them while others launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: launch sites keep one bit per device ordinal in a
// static mask (one process per GPU is the supported mode, but a process that drives several must not trip here)
// Read-only test: the caller sets the attributes, then calls mark_used_on_device() once ALL of them succeeded, so a
// failed cudaFuncSetAttribute is retried by the next call instead of leaving the bit set (atomic: host threads may race).
static inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  return ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull) == 0;
}
static inline void mark_used_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return;
  __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- ATen's linear-upsample arithmetic (align_corners=True) ---------------------
// Reference: models/audiotensor/audiotensor.py:11-17 -> F.interpolate(mode="linear",
// align_corners=True).  ATen CPU computes, in float:
//   scale = (n_in-1)/(n_out-1);  src = scale*dst;  i0 = min(floor(src), n_in-1);
//   l1 = clamp(src - i0, 0, 1);  l0 = 1 - l1;  i1 = i0 + (i0 < n_in-1);
//   out = fma(l0, x[i0], l1*x[i1])
// (pinned bit-for-bit against ATen on CPU by the test-suite).  Everything here uses
// explicit _rn intrinsics so nvcc cannot re-contract it differently.
struct Lerp {
  int i0, i1;
  float l0, l1;
};

__host__ __device__ __forceinline__ float lerp_scale(int n_in, int hop) {
  // (n_in-1) / ((n_in-1)*hop) in float, correctly rounded == 1/hop correctly rounded
  return n_in > 1 ? (float)(n_in - 1) / (float)((int64_t)(n_in - 1) * hop) : 0.f;
}

__device__ __forceinline__ Lerp lerp_at(int t, float scale, int n_in) {
  Lerp r;
  float src = __fmul_rn(scale, (float)t);
  int i0 = (int)floorf(src);
  i0 = min(i0, n_in - 1);
  float l1 = __fsub_rn(src, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  r.i0 = i0;
  r.i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  r.l1 = l1;
  r.l0 = __fsub_rn(1.f, l1);
  return r;
}

__device__ __forceinline__ float lerp_apply(const Lerp& w, float x0, float x1) {
  return __fmaf_rn(w.l0, x0, __fmul_rn(w.l1, x1));
}

// ---- small PTX wrappers: mbarrier + 1-D bulk copy (TMA unit) --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy through the TMA unit; dst/src 16-B aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store through the TMA unit
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// 4-byte asynchronous global -> shared copy (LDGSTS); !valid: the destination is zero-filled
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace golf
