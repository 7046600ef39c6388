// FP32 issue-rate probe for sm_100a: scalar FFMA (3 distinct registers) vs packed fma.rn.f32x2 (FFMA2).
// Prints lane-FMA per clock per SM for several warps/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NCH = 8, ITERS = 4096;
__global__ void k_ffma(float* out, float a, float b) {
  float x[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x * 1e-3f + i;
  float c0 = a, c1 = b;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) x[i] = __fmaf_rn(x[i], c0, c1);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[NCH], c0, c1;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
  }
  asm("mov.b64 %0, {%1, %1};" : "=l"(c0) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(b));
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(c0), "l"(c1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// register-tile pattern of the FIR kernels: acc[j] += tap[k] * win[j + k], all operands distinct registers
__global__ void k_tile(float* out, float a, float b) {
  float acc[8], tap[4], win[12];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) tap[i] = a + i * b;
#pragma unroll
  for (int i = 0; i < 12; ++i) win[i] = threadIdx.x * 1e-3f + i * b;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(tap[k], win[j + k], acc[j]);
    win[it & 7] += 1e-6f;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same tile with packed FMAs: acc pairs (j, j+1) need win pairs at even AND odd offsets -> keep two packings
__global__ void k_tile2(float* out, float a, float b) {
  unsigned long long acc[4], tapp[4], we[6], wo[6];
  for (int i = 0; i < 4; ++i) acc[i] = 0ull;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float t = a + i * b; asm("mov.b64 %0, {%1, %1};" : "=l"(tapp[i]) : "f"(t)); }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float w0 = threadIdx.x * 1e-3f + 2 * i * b, w1 = w0 + b, w2 = w1 + b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(we[i]) : "f"(w0), "f"(w1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(wo[i]) : "f"(w1), "f"(w2));
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned long long w = (k & 1) ? wo[j + k / 2] : we[j + k / 2];
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(tapp[k]), "l"(w));
      }
    we[it & 3] ^= 1ull;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float) * 4);
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s SMs %d clock %d kHz\n", pr.name, pr.multiProcessorCount, clk_khz);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int wps : {4, 8, 16, 32}) {
    for (int which = 0; which < 4; ++which) {
      dim3 grid(pr.multiProcessorCount), block(32 * wps);
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_ffma<<<grid, block>>>(out, 0.999f, 0.001f);
        else if (which == 1) k_ffma2<<<grid, block>>>(out, 0.999f, 0.001f);
        else if (which == 2) k_tile<<<grid, block>>>(out, 0.999f, 0.001f);
        else k_tile2<<<grid, block>>>(out, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)ITERS * 4 * NCH * ((which & 1) ? 2 : 1) * 32 * wps;  if (which == 3) fma /= 2;  // lane-FMA per SM
      double clocks = ms * 1e-3 * clk_khz * 1e3;
      printf("%-6s warps/SM %2d: %.3f ms, %.1f lane-FMA/clk/SM (at nominal clock)\n", (which == 0 ? "FFMA" : which == 1 ? "FFMA2" : which == 2 ? "TILE" : "TILE2"), wps, ms, fma / clocks);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
