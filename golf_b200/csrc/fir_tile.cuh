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

// ---- packed-FP32 tile (FFMA2) ---------------------------------------------------------------
// An sm_100a SM issues one warp instruction per scheduler per cycle and its FP32 pipe retires 32 lane-FMAs
// per scheduler per cycle, so a scalar-FFMA inner loop must spend EVERY issue slot on an FFMA to reach the
// pipe's peak; the tile above spends 34 slots per 32 FMAs and the loop around it more.  fma.rn.f32x2 (SASS
// FFMA2) does two FMAs per issued instruction (measured on B200: 126 lane-FMA/clk/SM, the pipe's peak, at
// half the issue rate -- tools/probe/fma_probe.cu), which leaves every other slot for the loads.
//
// Packing that keeps each output's taps in sequential order (same bits as the scalar tile): a 64-bit
// accumulator holds outputs (2i, 2i+1); tap j needs the input pair (x[2i+j], x[2i+j+1]) -- an aligned
// register pair of the strip for even j, and of the strip SHIFTED BY ONE SAMPLE for odd j.  So the strip is
// staged twice (xs0[n] = x[n], xs1[n] = x[n+1]) and the taps duplicated (kd[2j] = kd[2j+1] = k[j]).
// A thread owns 16 consecutive outputs; per 4 taps it issues 32 FFMA2 + 4 LDS.128.
constexpr int kR2 = 16;       // outputs per thread
constexpr int kTapStep = 20;  // taps per unrolled iteration (a 20-float register ring per strip)

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// XOR swizzle for strips read as float4 by threads 16 floats apart: float4 index c -> c ^ ((c >> 3) & 7);
// the eight lanes of a quarter-warp then cover the eight bank quads (no padding needed).
__host__ __device__ __forceinline__ int fir_sw16(int q) { return ((((q >> 2) ^ ((q >> 5) & 7))) << 2) | (q & 3); }

// acc[i] (outputs q0 + 2i, q0 + 2i + 1) += sum_{j<ntaps20} k[j] * x[q0 + 2i (+1) + j];  q0 % 4 == 0,
// strips in the fir_sw16() layout, logical indices up to q0 + ntaps20 + 23 readable, kd 16-B aligned.
__device__ __forceinline__ void fir_tile16_x2(const float* __restrict__ xs0, const float* __restrict__ xs1, int q0,
                                              const float* __restrict__ kd, int ntaps20, f32x2 (&acc)[kR2 / 2]) {
  f32x2 we[10], wo[10];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const float4 a = *reinterpret_cast<const float4*>(xs0 + fir_sw16(q0 + 4 * v));
    const float4 b = *reinterpret_cast<const float4*>(xs1 + fir_sw16(q0 + 4 * v));
    we[2 * v] = pack2(a.x, a.y), we[2 * v + 1] = pack2(a.z, a.w);
    wo[2 * v] = pack2(b.x, b.y), wo[2 * v + 1] = pack2(b.z, b.w);
  }
#pragma unroll 1
  for (int j = 0; j < ntaps20; j += kTapStep) {
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      const float4 ka = *reinterpret_cast<const float4*>(kd + 2 * (j + 4 * g));
      const float4 kb = *reinterpret_cast<const float4*>(kd + 2 * (j + 4 * g) + 4);
      const f32x2 t0 = pack2(ka.x, ka.y), t1 = pack2(ka.z, ka.w), t2 = pack2(kb.x, kb.y), t3 = pack2(kb.z, kb.w);
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) ffma2(acc[i], t0, we[(2 * g + i) % 10]);
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) ffma2(acc[i], t1, wo[(2 * g + i) % 10]);
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) ffma2(acc[i], t2, we[(2 * g + 1 + i) % 10]);
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) ffma2(acc[i], t3, wo[(2 * g + 1 + i) % 10]);
      const float4 a = *reinterpret_cast<const float4*>(xs0 + fir_sw16(q0 + j + 4 * g + 20));
      const float4 b = *reinterpret_cast<const float4*>(xs1 + fir_sw16(q0 + j + 4 * g + 20));
      we[(2 * g) % 10] = pack2(a.x, a.y), we[(2 * g + 1) % 10] = pack2(a.z, a.w);
      wo[(2 * g) % 10] = pack2(b.x, b.y), wo[(2 * g + 1) % 10] = pack2(b.z, b.w);
    }
  }
}

// Same strips and rings, taps NOT duplicated: a 64-bit accumulator holds the even-tap and the odd-tap partial sum of ONE
// output, so the natural tap pair (k[j], k[j+1]), j even, meets the input pair (x[o+j], x[o+j+1]) -- the even-aligned ring for
// even outputs, the shifted ring for odd ones.  The same 32 FFMA2 per 4 taps with 3 LDS.128 instead of 4, half the tap
// staging and half the shared memory for taps (more CTAs per SM); the result is lo + hi of each accumulator.
// acc[o] (output q0 + o) += sum_{j<ntaps20} k[j] * x[q0 + o + j];  k: ntaps20 floats, 16-B aligned.
__device__ __forceinline__ void fir_tile16_eo(const float* __restrict__ xs0, const float* __restrict__ xs1, int q0,
                                              const float* __restrict__ k, int ntaps20, f32x2 (&acc)[kR2]) {
  f32x2 we[10], wo[10];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const float4 a = *reinterpret_cast<const float4*>(xs0 + fir_sw16(q0 + 4 * v));
    const float4 b = *reinterpret_cast<const float4*>(xs1 + fir_sw16(q0 + 4 * v));
    we[2 * v] = pack2(a.x, a.y), we[2 * v + 1] = pack2(a.z, a.w);
    wo[2 * v] = pack2(b.x, b.y), wo[2 * v + 1] = pack2(b.z, b.w);
  }
#pragma unroll 1
  for (int j = 0; j < ntaps20; j += kTapStep) {
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      const float4 kq = *reinterpret_cast<const float4*>(k + j + 4 * g);
      const f32x2 t0 = pack2(kq.x, kq.y), t1 = pack2(kq.z, kq.w);
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) {
        ffma2(acc[2 * i], t0, we[(2 * g + i) % 10]);
        ffma2(acc[2 * i + 1], t0, wo[(2 * g + i) % 10]);
      }
#pragma unroll
      for (int i = 0; i < kR2 / 2; ++i) {
        ffma2(acc[2 * i], t1, we[(2 * g + 1 + i) % 10]);
        ffma2(acc[2 * i + 1], t1, wo[(2 * g + 1 + i) % 10]);
      }
      const float4 a = *reinterpret_cast<const float4*>(xs0 + fir_sw16(q0 + j + 4 * g + 20));
      const float4 b = *reinterpret_cast<const float4*>(xs1 + fir_sw16(q0 + j + 4 * g + 20));
      we[(2 * g) % 10] = pack2(a.x, a.y), we[(2 * g + 1) % 10] = pack2(a.z, a.w);
      wo[(2 * g) % 10] = pack2(b.x, b.y), wo[(2 * g + 1) % 10] = pack2(b.z, b.w);
    }
  }
}

}  // namespace golf
