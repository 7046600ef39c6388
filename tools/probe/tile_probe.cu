// Ceiling of the register-tiled FIR inner loops (fir_tile.cuh) with operands already in shared memory:
// scalar tile (8 outputs/lane) vs packed tile (16 outputs/lane, fma.rn.f32x2), n one-warp CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I golf_b200/csrc -I include -o tile_probe tile_probe.cu
#include <cstdio>
#include "fir_tile.cuh"
using namespace golf;
constexpr int K = 520, REP = 64;  // taps per pass, passes per CTA
__global__ void __launch_bounds__(32) k_scalar(float* out) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                 // 256 + 528 + 32 floats, fir_sw layout
  float* ks = smem + 1024;
  for (int i = threadIdx.x; i < 1024; i += 32) xs[i] = 1e-3f * i;
  for (int i = threadIdx.x; i < 528; i += 32) ks[i] = 1e-3f;
  __syncwarp();
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  for (int r = 0; r < REP; ++r) fir_tile8_sw(xs, threadIdx.x * kR, ks, 528, acc);
  float s = 0;
#pragma unroll
  for (int i = 0; i < kR; ++i) s += acc[i];
  out[blockIdx.x * 32 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(32) k_packed(float* out) {
  extern __shared__ __align__(16) float smem[];
  float* xs0 = smem;
  float* xs1 = smem + 1088;
  float* kd = smem + 2176;
  for (int i = threadIdx.x; i < 1088; i += 32) xs0[i] = 1e-3f * i, xs1[i] = 1e-3f * (i + 1);
  for (int i = threadIdx.x; i < 2 * K; i += 32) kd[i] = 1e-3f;
  __syncwarp();
  f32x2 acc[kR2 / 2];
#pragma unroll
  for (int i = 0; i < kR2 / 2; ++i) acc[i] = 0ull;
  for (int r = 0; r < REP; ++r) fir_tile16_x2(xs0, xs1, threadIdx.x * kR2, kd, K, acc);
  float s = 0;
#pragma unroll
  for (int i = 0; i < kR2 / 2; ++i) {
    float lo, hi;
    unpack2(acc[i], lo, hi);
    s += lo + hi;
  }
  out[blockIdx.x * 32 + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 32 * 32 * sizeof(float));
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {1, 2, 4, 8, 12, 16, 24, 32}) {
    for (int which = 0; which < 2; ++which) {
      // dynamic smem chosen so that exactly n CTAs fit an SM (227 KB usable)
      size_t need = which ? (2176 + 2 * K) * 4 : (1024 + 528) * 4;
      size_t sm = (size_t)(227 * 1024 / n) - 1024;
      if (sm > 64 * 1024) sm = 64 * 1024;
      if (sm < need) sm = need;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_scalar<<<148 * n, 32, sm>>>(out);
        else k_packed<<<148 * n, 32, sm>>>(out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)REP * (which ? K * 16.0 : 528 * 8.0) * 32 * n;  // lane-FMA per SM
      printf("%-6s CTAs/SM %2d: %.3f ms  %.1f lane-FMA/clk/SM\n", which ? "packed" : "scalar", n, ms, fma / (ms * 1e-3 * clk_khz * 1e3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
