// lpc_ss_tc.cuh -- GOLF-ss pass 1 (chunk responses) on the tensor cores: padded order 24, forward form.
//
// Same contract as ss_response_kernel<24, MT, 0> (lpc_ss.cuh): per chunk, W[col][k] = end-state component k of the
// recurrence started from unit state `col` (col < M: the transition matrix Phi) and, in column M, from rest driven by the
// chunk's excitation (z).  The FP32 kernel runs the M + 1 recurrences column by column, M FMAs per column and sample --
// M (M + 1) = 506 FMA per sample at M = 22.  Here the columns are the M dimension of a matrix product:
//
//   block of 8 steps t0 .. t0+7:   y_new = L y_new + H y_old + e      c[s][i] = -a_i(t0 + s), interpolated (ATen arithmetic)
//                                   L[s][r] = c[s][s-1-r] (r < s),  H[s][j] = c[s][s+23-j] (y_old[j] = y[t0-24+j])
//                                   y_new = G y_old + g_e,   G = (I - L)^-1 H  (8 x 24),  g_e = (I - L)^-1 e
//   G and g_e depend on the coefficients (and the excitation) only: one forward substitution per block in FP32, shared by all
//   columns -- lane j owns column j of G (28 FMAs), lane 24 owns g_e.  672 + 28 FMAs per 8 samples instead of 4048.
//   Y_new^T (columns x 8 steps) = Y_old^T (columns x 24 old steps) G^T: mma.sync.m16n8k8 TF32, 2 row tiles x 3 k-steps.
//   The D fragment of a block (rows g, g+8; steps 2t, 2t+1) IS the A fragment of the next blocks under the k-permutation
//   k = t <-> step 2t, k = t+4 <-> step 2t+1, so the state never leaves registers and no shuffle is needed; the B fragment
//   rows are read with the same permutation (one LDS.64 per k-step: G[g][8q+2t], G[g][8q+2t+1]).
//
// PREC 3: error-compensated products (x = hi + lo, hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly;
//   hi*hi + lo*hi + hi*lo in separate accumulators: ~2^-20 per product) -- float32-grade Phi / z, three MMAs per product.
// PREC 1: one TF32 product (round-to-nearest operands, 2^-11 per product): measured on B200 -- Phi / z are only good to ~1e-3
//   and one refinement round does NOT recover the parity bar on resonant filters (rows at 1e-2 .. 1e+1), so it is not
//   instantiated; kept in the template for the record.
//
// Measured (B200, B = 32 x 2 s, M = 22): 83 us against 74 us for the FP32 kernel -- the per-block overhead (staging,
// substitution, fragment splits: ~300 issued instructions per 8-step block and warp, 4 warps per scheduler) outweighs the
// 5.8x fewer FP32 FMAs, and the tensor cores' truncating accumulation leaves Phi ~10x less accurate than the FP32 kernel's,
// so the adaptive refinement round fires for most sequences.  Opt-in (golf_lpc_ss_set_response(1)); DESIGN.md 3.1.
//
// One warp per chunk, 4 warps per CTA, no block-level barrier.
#pragma once
#include <type_traits>

namespace golf {

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// hi = x with the 13 low mantissa bits cleared (what the tensor core reads of x anyway), lo = x - hi (exact)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(__fsub_rn(x, __uint_as_float(hi)));
}

constexpr int kTcWarps = 4;
// per warp: c rows [8][32] (taps 24..31 stay zero: H[s][j] = c[s][s+23-j] reads them for j < s), e at stride 33 behind them,
// and two G buffers ([8][24] + g_e at stride 24) so that the substitution of block b+1 overlaps the MMAs of block b
constexpr int kTcC = 8 * 32 + 8 * 33 + 8;
constexpr int kTcG = 8 * 24 + 8 * 24;
constexpr int kTcWarpFloats = kTcC + 2 * kTcG;

template <int PREC>
struct TcResp {
  static constexpr int MP = 24;
  uint32_t ah[3][2][4], al[3][2][4];  // A fragments of the three previous blocks (slot = block index mod 3)
  float qa0[6], qa1[6];               // this lane's taps of the current frame pair, negated
  int qf0, qf1;

  // (1) coefficient rows c[s][i] = -(l0 a[i0][i] + l1 a[i1][i]) (ATen arithmetic) and inputs e[s] of block blk
  __device__ __forceinline__ void stage(const SsParams& p, float* cb, const float* ab, const float* gb, const float* inb, int pi,
                                        int blk, int g, int t) {
    const int tt = pi * p.Lc + 8 * blk + g;
    const bool ok = tt < p.L;
    const Lerp w = lerp_at(ok ? tt : 0, p.scale, p.F);
    if (w.i0 != qf0 || w.i1 != qf1) {
      qf0 = w.i0, qf1 = w.i1;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const bool in = 6 * t + i < p.M;
        qa0[i] = in ? -__ldg(ab + (size_t)w.i0 * p.M + 6 * t + i) : 0.f;
        qa1[i] = in ? -__ldg(ab + (size_t)w.i1 * p.M + 6 * t + i) : 0.f;
      }
    }
    const float l0 = ok ? w.l0 : 0.f, l1 = ok ? w.l1 : 0.f;
    float* dst = cb + g * 32 + 6 * t;
#pragma unroll
    for (int i = 0; i < 6; i += 2) {
      float2 v;
      v.x = __fmaf_rn(l0, qa0[i], __fmul_rn(l1, qa1[i]));
      v.y = __fmaf_rn(l0, qa0[i + 1], __fmul_rn(l1, qa1[i + 1]));
      *reinterpret_cast<float2*>(dst + i) = v;
    }
    if (t == 0) {
      float e = 0.f;
      if (inb && ok) {
        e = __ldg(inb + tt);
        if (gb) e = __fmul_rn(e, __fmaf_rn(w.l0, __ldg(gb + w.i0), __fmul_rn(w.l1, __ldg(gb + w.i1))));
      }
      cb[8 * 32 + 33 * g] = e;
    }
  }

  // (2) forward substitution G = (I - L)^-1 H: lane j < 24 owns column j, lane 24 the forced column g_e.
  // hp: this lane's H diagonal (stride 33: c[s][s+23-j], or e[s] for lane 24); gp: where its column goes (stride 24)
  __device__ __forceinline__ void substitute(const float* cb, const float* hp, float* gp, bool store) {
    float h[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) h[s] = hp[33 * s];
    // L[s][r] = c[s][s-1-r]: taps 0..6 of rows 1..7, broadcast
    float l[8][8];
#pragma unroll
    for (int s = 1; s < 8; ++s) {
      const float4 v0 = *reinterpret_cast<const float4*>(cb + 32 * s);
      l[s][0] = v0.x, l[s][1] = v0.y, l[s][2] = v0.z, l[s][3] = v0.w;
      if (s > 4) {
        const float4 v1 = *reinterpret_cast<const float4*>(cb + 32 * s + 4);
        l[s][4] = v1.x, l[s][5] = v1.y, l[s][6] = v1.z, l[s][7] = v1.w;
      }
    }
    float G[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      float v = h[s];
#pragma unroll
      for (int r = 0; r < s; ++r) v = __fmaf_rn(l[s][s - 1 - r], G[r], v);
      G[s] = v;
    }
    if (store) {
#pragma unroll
      for (int s = 0; s < 8; ++s) gp[24 * s] = G[s];
    }
  }

  // (3a) issue the products of block blk: Y_new^T = Y_old^T G^T; the q-th oldest block sits in slot (R + q) % 3
  template <int R>
  __device__ __forceinline__ void products(const float* gbuf, int g, int t, float (&d)[2][4], float (&d1)[2][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int r = 0; r < 4; ++r) d[mt][r] = d1[mt][r] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      constexpr int dummy = 0;
      (void)dummy;
      const int slot = (R + q) % 3;
      const float2 bv = *reinterpret_cast<const float2*>(gbuf + g * 24 + 8 * q + 2 * t);
      uint32_t bh[2], bl[2];
      if (PREC == 3) {
        tf32_split(bv.x, bh[0], bl[0]);
        tf32_split(bv.y, bh[1], bl[1]);
      } else {
        bh[0] = tf32_rna(bv.x), bh[1] = tf32_rna(bv.y);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_tf32_16x8x8(d[mt], ah[slot][mt], bh);
        if (PREC == 3) {
          mma_tf32_16x8x8(d1[mt], al[slot][mt], bh);
          mma_tf32_16x8x8(d1[mt], ah[slot][mt], bl);
        }
      }
    }
  }

  // (3b) take the results: + g_e on the forced row, new A fragments into slot R, end state to W from the last 3 blocks
  template <int R>
  __device__ __forceinline__ void retire(const SsParams& p, float (&d)[2][4], float (&d1)[2][4], const float* geb, bool z_here, int zmt,
                                         int z_off, float* wb, int blk, int NB, int g, int t, bool has_in) {
    if (PREC == 3) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) d[mt][r] += d1[mt][r];
    }
    if (z_here) {
      const float gx = geb[24 * (2 * t)], gy = geb[24 * (2 * t + 1)];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int o = 0; o < 4; o += 2)
          if (mt == zmt && o == z_off) d[mt][o] += gx, d[mt][o + 1] += gy;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      // D (row g: c0,c1; row g+8: c2,c3) -> A (row g: a0,a2; row g+8: a1,a3)
      const float v[4] = {d[mt][0], d[mt][2], d[mt][1], d[mt][3]};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (PREC == 3) tf32_split(v[r], ah[R][mt][r], al[R][mt][r]);
        else ah[R][mt][r] = tf32_rna(v[r]);
      }
    }
    if (blk >= NB - 3) {
      // end-state component k <-> chunk step Lc-1-k: this block holds k = 8 (NB-1-blk) + 7 - s
      const int k0 = 8 * (NB - 1 - blk) + 6 - 2 * t;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int col = 16 * mt + g + 8 * hr;
          if (col < p.M || (col == p.M && has_in))
            *reinterpret_cast<float2*>(wb + col * MP + k0) = make_float2(d[mt][2 * hr + 1], d[mt][2 * hr]);
        }
    }
  }
};

template <int PREC>
__global__ void __launch_bounds__(32 * kTcWarps, 4) ss_response_tc_kernel(SsParams p) {
  constexpr int MP = 24;
  __shared__ __align__(16) float sm[kTcWarps][kTcWarpFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nresp = p.C - 1;
  const int task = blockIdx.x * kTcWarps + warp;
  if (task >= p.B * nresp) return;
  const int b = task / nresp, pi = task - b * nresp;
  float* __restrict__ cb = sm[warp];
  float* __restrict__ gbuf0 = cb + kTcC;
  const int g = lane >> 2, t = lane & 3;
  const float* __restrict__ ab = p.a + (size_t)b * p.F * p.M;
  const float* __restrict__ gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  const float* __restrict__ inb = p.in ? p.in + (size_t)b * p.in_stride : nullptr;
  const int M = p.M;

  // zero padding of the c rows (taps 24..31), written once
  for (int i = lane; i < 64; i += 32) cb[(i >> 3) * 32 + 24 + (i & 7)] = 0.f;

  TcResp<PREC> st;
  st.qf0 = st.qf1 = -1;
  // state starts as the identity (column m from unit state m: y[-1-m] = 1, i.e. old step j = 23 - m), column M (z) from
  // rest; before block 0 (R = 0) the q-th oldest block is slot q
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = 16 * mt + g + ((r & 1) ? 8 : 0);    // a0,a2: row g; a1,a3: row g+8
        const int j = 8 * q + 2 * t + ((r & 2) ? 1 : 0);  // a0,a1: k = t <-> step 2t; a2,a3: k = t+4 <-> step 2t+1
        st.ah[q][mt][r] = (m < M && 23 - m == j) ? __float_as_uint(1.f) : 0u;
        st.al[q][mt][r] = 0u;
      }
  // the forced row (column M of [Phi | z]) within this lane's fragment, if it is here
  const int zmt = M >> 4, zr = M & 15;
  const bool z_here = inb != nullptr && g == (zr & 7);
  const int z_off = (zr >> 3) * 2;
  // substitution role: H diagonal and output column of this lane
  const int j = lane;
  const float* hp = j < 24 ? cb + 23 - j : cb + 8 * 32;   // lanes >= 24 read e (only lane 24 stores)
  const int gcol = j < 24 ? j : 8 * 24;                    // g_e lives behind G, same stride
  const bool gstore = j <= 24;
  const int NB = p.Lc >> 3;
  float* wb = p.W + ((size_t)b * nresp + pi) * ((MP + 1) * MP);

  __syncwarp();
  st.stage(p, cb, ab, gb, inb, pi, 0, g, t);
  __syncwarp();
  st.substitute(cb, hp, gbuf0 + gcol, gstore);
  __syncwarp();

  auto body = [&](auto Rtag, int blk) {
    constexpr int R = decltype(Rtag)::value;
    const float* gcur = gbuf0 + (blk & 1) * kTcG;
    float* gnext = gbuf0 + ((blk + 1) & 1) * kTcG;
    float d[2][4], d1[2][4];
    st.template products<R>(gcur, g, t, d, d1);
    if (blk + 1 < NB) {  // the next block's G while the tensor cores work
      st.stage(p, cb, ab, gb, inb, pi, blk + 1, g, t);
      __syncwarp();
      st.substitute(cb, hp, gnext + gcol, gstore);
    }
    st.template retire<R>(p, d, d1, gcur + 8 * 24, z_here, zmt, z_off, wb, blk, NB, g, t, inb != nullptr);
    __syncwarp();
  };
#pragma unroll 1
  for (int blk = 0; blk < NB; blk += 3) {
    body(std::integral_constant<int, 0>{}, blk);
    if (blk + 1 < NB) body(std::integral_constant<int, 1>{}, blk + 1);
    if (blk + 2 < NB) body(std::integral_constant<int, 2>{}, blk + 2);
  }
}

extern std::atomic<int> g_ss_response_mode;  // 0: FP32 kernel (default), 1: tensor cores, 3 x TF32

inline bool response_tc_applies(const SsParams& p, int MP, int FORM) {
  return g_ss_response_mode != 0 && FORM == 0 && MP == 24 && p.Lc % 8 == 0 && p.Lc >= 24 && p.M <= 24;
}

inline int launch_response_tc(const SsParams& p, cudaStream_t st) {
  const int tasks = p.B * (p.C - 1);
  if (tasks <= 0) return GOLF_OK;
  const int grid = ceil_div(tasks, kTcWarps);
  ss_response_tc_kernel<3><<<grid, 32 * kTcWarps, 0, st>>>(p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

}  // namespace golf
