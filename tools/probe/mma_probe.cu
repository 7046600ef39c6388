// Legacy tensor-core path probe for sm_100a: mma.sync m16n8k8 TF32 and m16n8k16 BF16 issue rates (independent accumulator
// chains) and the latency of a dependent chain (D of one MMA is C -- or, re-used as A -- of the next).
// Prints MMA per clock per SM and TFLOP/s at the measured clock for several warps/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int KIND, int NCH>
__global__ void k_rate(float* out, long long* clk) {
  float d[NCH][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * threadIdx.x + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(2e-3f * threadIdx.x + i);
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (KIND == 0) mma_tf32(d[c], a, b); else mma_bf16(d[c], a, b);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
// dependent chain: the D fragment of step i is the A fragment (registers re-used as they are) of step i+1
__global__ void k_chain_da(float* out, long long* clk) {
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * threadIdx.x + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(1e-3f * threadIdx.x + i);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]),
                   "f"(z[0]), "f"(z[1]), "f"(z[2]), "f"(z[3]));
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(d[i]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d[0] + d[1] + d[2] + d[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <class F>
static void run(const char* name, F launch, int nch, double flop_per_mma) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float)); cudaMalloc(&clk, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wps : {1, 2, 4, 8, 16}) {
    launch(148, 32 * wps, out, clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(148, 32 * wps, out, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double mmas = (double)ITERS * nch * wps;  // per SM
    printf("%-28s warps/SM %2d: %6.3f MMA/clk/SM  (%7.1f clk per MMA per warp)  %7.1f TFLOP/s by events\n", name, wps, mmas / c,
           (double)c / (ITERS * nch), mmas * 148 * flop_per_mma / (ms * 1e-3) / 1e12);
  }
  cudaFree(out); cudaFree(clk);
}
int main() {
  run("tf32 m16n8k8, 8 chains", [](int g, int b, float* o, long long* c) { k_rate<0, 8><<<g, b>>>(o, c); }, 8, 2.0 * 16 * 8 * 8);
  run("tf32 m16n8k8, 2 chains", [](int g, int b, float* o, long long* c) { k_rate<0, 2><<<g, b>>>(o, c); }, 2, 2.0 * 16 * 8 * 8);
  run("bf16 m16n8k16, 8 chains", [](int g, int b, float* o, long long* c) { k_rate<1, 8><<<g, b>>>(o, c); }, 8, 2.0 * 16 * 8 * 16);
  run("tf32 chain D->A (latency)", [](int g, int b, float* o, long long* c) { k_chain_da<<<g, b>>>(o, c); }, 1, 2.0 * 16 * 8 * 8);
  return 0;
}
