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

}  // namespace golf
