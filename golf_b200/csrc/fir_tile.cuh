// fir_tile.cuh -- register-tiled FIR correlation shared by the noise FIR, the room FIR and
// the oscillator's polyphase decimator.
#pragma once
#include "common.cuh"

namespace golf {

constexpr int kR = 8;  // outputs per thread

// acc[i] += sum_{j<ntaps12} k[j] * x[i + j],  i < 8;  ntaps12 % 12 == 0, x and k 16-B aligned,
// x readable up to index ntaps12 + 19.
__device__ __forceinline__ void fir_tile8(const float* __restrict__ x, const float* __restrict__ k, int ntaps12,
                                          float (&acc)[kR]) {
  float xw[12];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const float4 t = *reinterpret_cast<const float4*>(x + 4 * v);
    xw[4 * v] = t.x, xw[4 * v + 1] = t.y, xw[4 * v + 2] = t.z, xw[4 * v + 3] = t.w;
  }
#pragma unroll 1
  for (int j = 0; j < ntaps12; j += 12) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 kv = *reinterpret_cast<const float4*>(k + j + 4 * g);
      const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < kR; ++i) acc[i] = __fmaf_rn(kk[jj], xw[(4 * g + jj + i) % 12], acc[i]);
      const float4 t = *reinterpret_cast<const float4*>(x + j + 4 * g + 12);
      xw[(4 * g) % 12] = t.x, xw[(4 * g + 1) % 12] = t.y, xw[(4 * g + 2) % 12] = t.z, xw[(4 * g + 3) % 12] = t.w;
    }
  }
}

// Swizzled variant.  A thread's 8 outputs start at logical index 8*tid, so the float4 loads of a
// quarter-warp are 32 bytes apart: lanes i and i+4 hit the same four banks (2-way conflict, and the
// x loads are 4 of every 5 shared-memory wavefronts of this loop).  Storing logical element q at
// physical index sw(q) = q + 4*(q/32) shifts every 32-float row by one float4: the eight lanes of a
// quarter-warp then cover eight different bank quads.  q % 4 == 0 -> sw(q) % 4 == 0, and a float4
// never straddles a row, so aligned vector loads still work.
__host__ __device__ __forceinline__ int fir_sw(int q) { return q + ((q >> 5) << 2); }

// acc[i] += sum_{j<ntaps12} k[j] * x[q0 + i + j],  i < 8;  xs holds x in the sw() layout, q0 % 4 == 0,
// logical indices up to q0 + ntaps12 + 19 must be readable.
__device__ __forceinline__ void fir_tile8_sw(const float* __restrict__ xs, int q0, const float* __restrict__ k, int ntaps12,
                                             float (&acc)[kR]) {
  float xw[12];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const float4 t = *reinterpret_cast<const float4*>(xs + fir_sw(q0 + 4 * v));
    xw[4 * v] = t.x, xw[4 * v + 1] = t.y, xw[4 * v + 2] = t.z, xw[4 * v + 3] = t.w;
  }
#pragma unroll 1
  for (int j = 0; j < ntaps12; j += 12) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 kv = *reinterpret_cast<const float4*>(k + j + 4 * g);
      const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < kR; ++i) acc[i] = __fmaf_rn(kk[jj], xw[(4 * g + jj + i) % 12], acc[i]);
      const float4 t = *reinterpret_cast<const float4*>(xs + fir_sw(q0 + j + 4 * g + 12));
      xw[(4 * g) % 12] = t.x, xw[(4 * g + 1) % 12] = t.y, xw[(4 * g + 2) % 12] = t.z, xw[(4 * g + 3) % 12] = t.w;
    }
  }
}

}  // namespace golf
