// osc.cu -- glottal-flow wavetable oscillator (models/synth.py:213-263) on sm_100a.
//
// Reference pipeline (IndexedGlottalFlowTable.forward): interpolate two table rows per
// control frame -> tables[B,Fw,P]; phase/os upsampled linearly to the os-times
// oversampled rate; float cumsum; mod 1; bilinear (frame-time x phase) read through
// F.grid_sample (GlottalFlowTable.generate, synth.py:124-177); * rsqrt(phase);
// kazane.Decimate(os).  Five [B, 4T] streams round-trip through memory there.
//
// Here:
//   osc_tables_kernel       tables[b,f,:] = table[lo]*(1-p) + table[lo+1]*p   (L2 resident)
//   osc_knot_prefix_*       the phase increments are piecewise linear between the Np
//                           control knots, so the running sum has a closed form inside a
//                           knot interval; only the per-knot prefix needs a scan.  Default:
//                           exact 64-bit fixed point (Q0.64 cycles, wraps at one period,
//                           integer adds -> order-independent, no FP64 unit needed).
//                           accumulate=1 ("aten_cpu") keeps the unwrapped sum and rounds it
//                           to float32 before `% 1`, which is what ATen's CPU cumsum does
//                           (double accumulator, float32 output) -- used for parity checks.
//   osc_flow_decimate_kernel one CTA per 1024 output samples: evaluates the wrapped phase
//                           (closed form or stored), does the bilinear table read with
//                           grid_sample's align_corners=True arithmetic, scales by
//                           rsqrt(increment), lays the oversampled flow out polyphase in
//                           shared memory and applies the decimation FIR as os register-
//                           tiled correlations.  The oversampled stream never leaves the SM.
#include "fir_tile.cuh"

namespace golf {

__global__ void osc_tables_kernel(const float* __restrict__ w, const float* __restrict__ table,
                                  float* __restrict__ tables, int BF, int n_tab, int P) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BF * P) return;
  const int bf = idx / P, c = idx % P;
  const float raw = __fmul_rn(w[bf], (float)(n_tab - 1));
  int lo = (int)raw;  // .long() truncates toward zero
  lo = min(max(lo, 0), n_tab - 2);
  const float p = __fsub_rn(raw, (float)lo);
  const float v0 = __fmul_rn(table[(size_t)lo * P + c], __fsub_rn(1.f, p));
  const float v1 = __fmul_rn(table[(size_t)(lo + 1) * P + c], p);
  tables[idx] = __fadd_rn(v0, v1);
}

// Exclusive per-knot prefix of the oversampled phase increments, float64, frac part.
// x_k = phase[k]/os.  Interval k (hp = phase_hop*os samples, t = k*hp + r):
//   inc(t) = x_k + (x_{k+1}-x_k) r/hp ;  sum_{r<hp} inc = hp x_k + (x_{k+1}-x_k)(hp-1)/2
__global__ void __launch_bounds__(256) osc_knot_prefix_kernel(const float* __restrict__ phase, double* __restrict__ pref,
                                                              int Np, int hp, int os, int wrap) {
  __shared__ double part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* __restrict__ ph = phase + (size_t)b * Np;
  double* pb = pref + (size_t)b * Np;
  const int per = (Np + 255) / 256;
  const int k0 = tid * per, k1 = min(Np, k0 + per);
  const double inv_os = 1.0 / (double)os, dh = (double)hp, half = 0.5 * (double)(hp - 1);
  // interval k contributes hp*x_k + (x_{k+1}-x_k)(hp-1)/2; the loads are independent of the
  // running sum, so unrolling lets them overlap
  double s = 0.0;
#pragma unroll 8
  for (int k = k0; k < k1; ++k) {
    const double xk = (double)__ldg(ph + k) * inv_os;
    const double xn = (double)__ldg(ph + min(k + 1, Np - 1)) * inv_os;
    s += dh * xk + (xn - xk) * half;
  }
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int i = 0; i < 256; ++i) {
      const double v = part[i];
      part[i] = run;
      run += v;
      if (wrap) run -= floor(run);
    }
  }
  __syncthreads();
  double run = part[tid];
#pragma unroll 8
  for (int k = k0; k < k1; ++k) {
    pb[k] = run;
    const double xk = (double)__ldg(ph + k) * inv_os;
    const double xn = (double)__ldg(ph + min(k + 1, Np - 1)) * inv_os;
    run += dh * xk + (xn - xk) * half;
    if (wrap) run -= floor(run);
  }
}

// phase / os exactly as the reference divides it; a power-of-two os makes the reciprocal exact
__device__ __forceinline__ float div_os(float x, float os_f, float inv_os_f, bool pow2) {
  return pow2 ? __fmul_rn(x, inv_os_f) : __fdiv_rn(x, os_f);
}

// ---- exact phase in 64-bit fixed point ------------------------------------------------
// Default accumulation.  A phase in cycles only matters mod 1, so it is kept as an unsigned
// Q0.64 fraction: adding increments is exact integer arithmetic that wraps exactly at one
// cycle, is associative (any scan order gives the same bits) and needs no FP64 unit.
__device__ __forceinline__ uint64_t q64_from_float(float x) {  // x in [0, 1): x * 2^64, exact for x >= 2^-41
  const uint32_t u = __float_as_uint(x);
  const int sh = (int)((u >> 23) & 0xff) - 86;  // value = m * 2^(e-150); times 2^64
  const uint64_t m = (uint64_t)((u & 0x7fffffu) | 0x800000u);
  return (sh >= 0 && sh < 64) ? m << sh : 0ull;  // increments below 2^-41 cycles per sample (and zero) flush to zero
}
// increment sum of knot interval k: hp*X_k + (X_{k+1}-X_k)(hp-1)/2   (mod 2^64)
__device__ __forceinline__ uint64_t q64_interval(uint64_t xk, uint64_t xn, int hp) {
  const int64_t half = ((int64_t)(xn - xk)) >> 1;
  return xk * (uint64_t)hp + (uint64_t)(half * (int64_t)(hp - 1));
}

// One WARP per span of the knots (kPrefSplit spans per utterance, kPrefWarps of them per CTA: the grid is
// B * kPrefSplit / kPrefWarps CTAs, 256 at B = 32, so the scan runs on every SM -- its first version used one
// CTA per utterance, 32 of 148 SMs, 38 us).  The warp walks its span 32*kPrefR knots at a time: a lane owns
// kPrefR CONSECUTIVE knots (serial adds in registers), only the 32 lane totals go through a shuffle scan.
// Output: the span-local exclusive prefix pref[b][k] and the span totals tot[b][span]; readers turn the
// 32 totals into exclusive span offsets with one warp scan in their prologue (span_offsets) and add
// pref[k] + soff[k / span] (integer adds wrap at one cycle, any order gives the same bits).
constexpr int kPrefSplit = 32;
constexpr int kPrefWarps = 4;
constexpr int kPrefR = 8;
__global__ void __launch_bounds__(32 * kPrefWarps) osc_knot_prefix_q64_kernel(const float* __restrict__ phase,
                                                                              unsigned long long* __restrict__ pref,
                                                                              unsigned long long* __restrict__ tot, int Np,
                                                                              int hp, float os_f, int span) {
  pdl_trigger();  // the flow kernel's prologue (taps, table rows) runs beside this scan
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = blockIdx.x * kPrefWarps + (threadIdx.x >> 5);
  const float* __restrict__ ph = phase + (size_t)b * Np;
  unsigned long long* pb = pref + (size_t)b * Np;
  const bool pow2 = ((int)os_f & ((int)os_f - 1)) == 0;
  const float inv_os_f = 1.f / os_f;
  const int s0 = min(warp * span, Np), s1 = min(Np, s0 + span);
  unsigned long long carry = 0;
  float f[kPrefR + 1], fnext[kPrefR + 1];
  auto fetch = [&](int k0, float (&dst)[kPrefR + 1]) {
#pragma unroll
    for (int i = 0; i <= kPrefR; ++i) dst[i] = __ldg(ph + min(k0 + i, Np - 1));
  };
  fetch(s0 + lane * kPrefR, f);
  for (int base = s0; base < s1; base += 32 * kPrefR) {
    const int k0 = base + lane * kPrefR;
    if (base + 32 * kPrefR < s1) fetch(k0 + 32 * kPrefR, fnext);  // next round's loads fly during this round's scan
    uint64_t x[kPrefR + 1], v[kPrefR];
#pragma unroll
    for (int i = 0; i <= kPrefR; ++i) x[i] = q64_from_float(div_os(f[i], os_f, inv_os_f, pow2));
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < kPrefR; ++i) {
      v[i] = k0 + i < s1 ? q64_interval(x[i], x[i + 1], hp) : 0ull;
      mine += v[i];
    }
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += u;
    }
    unsigned long long acc = carry + inc - mine;
#pragma unroll
    for (int i = 0; i < kPrefR; ++i) {
      if (k0 + i < s1) pb[k0 + i] = acc;
      acc += v[i];
    }
    carry += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
    for (int i = 0; i <= kPrefR; ++i) f[i] = fnext[i];
  }
  if (lane == 0) tot[(size_t)b * kPrefSplit + warp] = carry;
}

// reader prologue (first warp of the CTA): exclusive scan of the utterance's span totals into shared memory
__device__ __forceinline__ void span_offsets(const unsigned long long* __restrict__ tot_b, unsigned long long* soff_s,
                                             const double* __restrict__ phase0 = nullptr, int b = 0) {
  const int lane = threadIdx.x & 31;
  // phase0[b]: running phase (cycles) the utterance starts from -- a stream continuing a previous call.  float64
  // so that the hand-over is good to 2^-53 of a cycle (a float32 offset would move the table read by 1e-4 of a
  // column); converted to Q0.64 in two 32-bit halves
  unsigned long long start = 0ull;
  if (phase0) {
    const double fr = (phase0[b] - floor(phase0[b])) * 4294967296.0;
    const double hi = floor(fr);
    start = ((unsigned long long)(unsigned int)hi << 32) | (unsigned long long)(unsigned int)((fr - hi) * 4294967296.0);
  }
  const unsigned long long w = tot_b[lane];
  unsigned long long winc = w;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, d);
    if (lane >= d) winc += u;
  }
  soff_s[lane] = winc - w + start;
}

// upsampled increment at oversampled time t, ATen arithmetic on phase/os
__device__ __forceinline__ float osc_inc(const float* ph, int t, float scale, int Np, float os_f, float inv_os_f, bool pow2) {
  const Lerp w = lerp_at(t, scale, Np);
  return lerp_apply(w, div_os(__ldg(ph + w.i0), os_f, inv_os_f, pow2), div_os(__ldg(ph + w.i1), os_f, inv_os_f, pow2));
}

// bilinear table read, F.grid_sample(align_corners=True, zeros padding) arithmetic.  The column
// (phase) coordinate follows ATen operation by operation -- the table is steep around glottal
// closure; the row (time) coordinate uses t * (1/ydenom): rows differ slowly, the 1-ulp
// difference against ATen's division moves the result by < 1e-7.
struct OscTaps {
  int row0, row1, c0, c1;      // table rows (already clamped to the R supplied ones) and columns; c1 < 0: outside
  bool has_row1;
  float w00, w01, w10, w11;    // bilinear weights of (row0,c0) (row0,c1) (row1,c0) (row1,c1)
};
__device__ __forceinline__ OscTaps osc_taps(int R, int P, float wrapped, int t, float inv_ydenom, int blocks) {
  const float gx = __fsub_rn(__fmul_rn(wrapped, 2.f), 1.f);
  const float gy = __fsub_rn(__fmul_rn(__fmul_rn((float)t, inv_ydenom), 2.f), 1.f);
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)P);  // (x+1)/2: *0.5 is the same float
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)blocks);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = __fsub_rn(ix, x0f), fy = __fsub_rn(iy, y0f);
  const int x0 = min(max((int)x0f, 0), P), y0 = min(max((int)y0f, 0), blocks);
  OscTaps o;
  // column P wraps to column 0; column P+1 / row blocks+1 are outside the image (weight 0 anyway);
  // rows beyond the R supplied ones replicate the last (F.pad replicate, synth.py:138-148)
  o.c0 = x0 == P ? 0 : x0;
  o.c1 = x0 + 1 >= P ? (x0 + 1 == P ? 0 : -1) : x0 + 1;
  o.row0 = min(y0, R - 1), o.row1 = min(y0 + 1, R - 1);
  o.has_row1 = y0 + 1 <= blocks;
  const float gx1 = __fsub_rn(1.f, fx), gy1 = __fsub_rn(1.f, fy);
  o.w00 = __fmul_rn(gx1, gy1), o.w01 = __fmul_rn(fx, gy1), o.w10 = __fmul_rn(gx1, fy), o.w11 = __fmul_rn(fx, fy);
  return o;
}
__device__ __forceinline__ float osc_read(const float* __restrict__ tb, int R, int P, float wrapped, int t,
                                          float inv_ydenom, int blocks) {
  const OscTaps o = osc_taps(R, P, wrapped, t, inv_ydenom, blocks);
  const float* r0p = tb + (size_t)o.row0 * P;
  const float* r1p = tb + (size_t)o.row1 * P;
  const float t00 = __ldg(r0p + o.c0), t01 = o.c1 >= 0 ? __ldg(r0p + o.c1) : 0.f;
  const float t10 = o.has_row1 ? __ldg(r1p + o.c0) : 0.f, t11 = (o.has_row1 && o.c1 >= 0) ? __ldg(r1p + o.c1) : 0.f;
  float v = __fmul_rn(t00, o.w00);
  v = __fmaf_rn(t01, o.w01, v);
  v = __fmaf_rn(t10, o.w10, v);
  v = __fmaf_rn(t11, o.w11, v);
  return v;
}

__global__ void wavetable_read_kernel(const float* __restrict__ wrapped, const float* __restrict__ tables,
                                      float* __restrict__ out, int N, int R, int P, int hop_tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= N) return;
  const int blocks = (N + hop_tab - 1) / hop_tab;
  out[(size_t)b * N + t] =
      osc_read(tables + (size_t)b * R * P, R, P, wrapped[(size_t)b * N + t], t, 1.f / (float)((int64_t)hop_tab * blocks), blocks);
}

// ---- fused flow + decimation --------------------------------------------------------
constexpr int kOscTile = 1024;  // outputs per CTA (128 threads x 8)
constexpr int kMaxOs = 8;

struct OscParams {
  const float* phase;     // [B,Np]
  const float* tables;    // [B,Fw,P]
  const double* pref;     // [B,Np] exclusive knot prefix (frac part, or unwrapped when aten_cpu)
  int aten_cpu;           // 1: round the running sum to float32 before mod 1 (ATen CPU cumsum semantics)
  const unsigned long long* totals;  // [B][kPrefSplit] span totals of the Q0.64 prefix (readers: span_offsets)
  int span;               // knots per span
  const double* phase0;   // [B] initial running phase in cycles, or null (exact-phase mode only)
  const float* dec;       // [2*zeros*os+1]
  float* out;             // [B,n_out]
  int B, Np, hp, N, n_out, Fw, P, hop_tab, blocks, os, zeros, equal_energy;
  int plen;               // per-phase strip length (floats, multiple of 4)
  int kp12;               // per-phase taps padded to a multiple of 12
  float scale, ydenom, inv_2hp;
};

// running phase of the `os` oversampled samples of output-rate index mj (they share one knot interval)
struct KnotPhase {
  int r0;
  double xk, dk, pk;   // aten_cpu: float64 running sum
  uint64_t qx, qp;     // default: Q0.64 fixed point
  int64_t qq;
  __device__ __forceinline__ void init(const OscParams& p, const float* __restrict__ ph, int b, int mj, float os_f,
                                       float inv_os_f, bool pow2, const unsigned long long* soff_s) {
    const int phase_hop = p.hp / p.os;
    int k = phase_hop == 1 ? mj : mj / phase_hop;
    k = min(k, p.Np - 1);
    r0 = (mj - k * phase_hop) * p.os;
    const float fk = __ldg(ph + k), fn = __ldg(ph + min(k + 1, p.Np - 1));
    if (p.aten_cpu) {
      const double inv_os = 1.0 / (double)p.os;
      xk = (double)fk * inv_os;
      dk = (double)fn * inv_os - xk;
      pk = __ldg(p.pref + (size_t)b * p.Np + k);
    } else {
      qx = q64_from_float(div_os(fk, os_f, inv_os_f, pow2));
      const uint64_t qn = q64_from_float(div_os(fn, os_f, inv_os_f, pow2));
      // slope term per r(r+1): (x_{k+1}-x_k)/(2hp).  A float reciprocal is plenty: its 6e-8 relative
      // error moves the phase by < 1e-9 cycles (the term itself is a second-order correction)
      qq = __float2ll_rn(__ll2float_rn((int64_t)(qn - qx)) * p.inv_2hp);
      qp = reinterpret_cast<const unsigned long long*>(p.pref)[(size_t)b * p.Np + k];
      qp += soff_s[k / p.span];  // exclusive span offset
    }
  }
  __device__ __forceinline__ float wrapped(const OscParams& p, int phs) const {
    const int r = r0 + phs;
    if (p.aten_cpu) {  // cumsum output is float32, then `% 1` in float32
      const double phi = pk + (double)(r + 1) * xk + dk * ((double)r * (double)(r + 1)) / (2.0 * (double)p.hp);
      const float f = (float)phi;
      return __fsub_rn(f, floorf(f));
    }
    const uint64_t phi = qp + (uint64_t)(r + 1) * qx + (uint64_t)(qq * (int64_t)(r * (r + 1)));
    const float wr = __fmul_rn(__ull2float_rn(phi), 5.42101086242752217e-20f);  // * 2^-64
    return wr >= 1.f ? 0.f : wr;
  }
};

template <int OS>  // oversampling factor known at compile time (0: runtime)
__global__ void __launch_bounds__(128) osc_flow_decimate_kernel(OscParams p) {
  extern __shared__ __align__(16) float smem[];
  float* vp = smem;                      // [os][plen] polyphase oversampled flow
  float* hp_ = smem + p.os * p.plen;     // [os][kp12] polyphase taps
  __shared__ unsigned long long soff_s[kPrefSplit];
  if (threadIdx.x < 32 && !p.aten_cpu) span_offsets(p.totals + (size_t)blockIdx.y * kPrefSplit, soff_s, p.phase0, blockIdx.y);
  const int b = blockIdx.y, tid = threadIdx.x;
  const int m0 = blockIdx.x * kOscTile;
  const float* ph = p.phase + (size_t)b * p.Np;
  const float* tb = p.tables + (size_t)b * p.Fw * p.P;
  const float os_f = (float)p.os;
  const int Z = p.zeros;
  // polyphase taps: out[m] = sum_n h[n] v[(m-Z)*os + n]; n = q*os + ph -> v[(m-Z+q)*os + ph]
  for (int i = tid; i < p.os * p.kp12; i += blockDim.x) {
    const int phs = i / p.kp12, q = i % p.kp12;
    const int n = q * p.os + phs;
    hp_[i] = (n <= 2 * Z * p.os) ? (p.dec ? p.dec[n] : 1.f) : 0.f;
  }
  // strip index j <-> output-rate index mj = m0 - Z + j; its `os` oversampled samples
  // t = mj*os + phs share one knot interval (knot spacing hp = phase_hop*os), so the float64
  // prefix / slope are fetched once per j:  vp[phs][j] = v[mj*os + phs]
  const bool pow2 = (p.os & (p.os - 1)) == 0;
  const float inv_os_f = 1.f / os_f, inv_yd = 1.f / p.ydenom;
  __syncthreads();  // soff_s
  for (int j = tid; j < p.plen; j += blockDim.x) {
    const int mj = m0 - Z + j;
    const bool in = mj >= 0 && (int64_t)mj * p.os < p.N;
    KnotPhase kp;
    if (in) kp.init(p, ph, b, mj, os_f, inv_os_f, pow2, soff_s);
    const int nos = OS ? OS : p.os;
#pragma unroll
    for (int phs = 0; phs < nos; ++phs) {
      const int t = mj * nos + phs;
      float v = 0.f;
      if (in && t < p.N) {
        const float wr = kp.wrapped(p, phs);
        v = osc_read(tb, p.Fw, p.P, wr, t, inv_yd, p.blocks);
        // torch.rsqrt: MUFU.RSQ-based rsqrtf is within 2 ulp of ATen's CPU 1/sqrt
        if (p.equal_energy) v = __fmul_rn(v, rsqrtf(osc_inc(ph, t, p.scale, p.Np, os_f, inv_os_f, pow2)));
      }
      vp[phs * p.plen + j] = v;
    }
  }
  __syncthreads();
  const int r0 = tid * kR;  // outputs m0 + r0 .. m0 + r0 + 7
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  const int nos2 = OS ? OS : p.os;
#pragma unroll
  for (int phs = 0; phs < nos2; ++phs) fir_tile8(vp + phs * p.plen + r0, hp_ + phs * p.kp12, p.kp12, acc);
  float* ob = p.out + (size_t)b * p.n_out;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (m0 + r0 + i < p.n_out) ob[m0 + r0 + i] = acc[i];
}


// ---- fused flow + decimation, v2 (exact-phase mode, compile-time oversampling) --------------
// Same function as osc_flow_decimate_kernel in its exact-phase mode, restructured so a sample costs a
// fraction of the instructions and none of its loads leave the SM:
//  * the (at most three) interpolated table rows a tile can touch are built straight from the base
//    table into shared memory (osc_tables_kernel's arithmetic; that launch and its [B,Fw,P] tensor
//    disappear), so the four bilinear taps are LDS instead of dependent L2 gathers;
//  * the Q0.64 phase of the `os` samples of one output index advances by exact integer increments
//    (phi += d; d += 2q) instead of re-evaluating the closed form;
//  * table coordinates come straight from the integers: with P a power of two the column is the top
//    log2(P) bits of the Q0.64 phase and the column fraction the next 32 bits; the row is t / hop_tab
//    and the row fraction (t % hop_tab) / hop_tab, found by one compare because a tile is shorter than a
//    row interval.  grid_sample's float detour (normalise to [-1,1], unnormalise, floor) rounds these
//    same quantities to 24 bits on the way: the two agree to ~1e-7 of a column / row, far inside what the
//    exact phase already differs from the reference's float32 cumsum by;
//  * the upsampled increment (equal-energy scaling) reuses the knot pair already in registers;
//  * the polyphase strips use the conflict-free fir_sw() layout.
constexpr int kOscRows = 3;
template <int OS>
__global__ void __launch_bounds__(128) osc_flow_v2_kernel(OscParams p, const float* __restrict__ w,
                                                          const float* __restrict__ table, int n_tab) {
  extern __shared__ __align__(16) float smem[];
  const int plen_sw = fir_sw(p.plen) + 4;
  float* vp = smem;                         // [OS][plen_sw] polyphase oversampled flow, fir_sw() layout
  float* hp_ = vp + OS * plen_sw;           // [OS][kp12] polyphase taps
  float* rows = hp_ + OS * p.kp12;          // [kOscRows][P] interpolated table rows ybase .. ybase+2
  __shared__ unsigned long long soff_s[kPrefSplit];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int m0 = blockIdx.x * kOscTile;
  const float* __restrict__ ph = p.phase + (size_t)b * p.Np;
  const int Z = p.zeros, P = p.P;
  const float os_f = (float)OS, inv_os_f = 1.f / os_f, inv_yd = 1.f / p.ydenom;
  constexpr bool pow2 = (OS & (OS - 1)) == 0;
  // first table row this tile can touch.  The tile (with its FIR halo) is shorter than one row
  // interval (checked on the host), so its samples see y0 = t / hop_tab in {ybase, ybase+1} and read
  // rows ybase .. ybase+2; staged row i is control frame min(ybase+i, Fw-1) (replicate padding).
  const int t_first = max((m0 - Z) * OS, 0);
  const int ybase = t_first / p.hop_tab;
  const int trow0 = ybase * p.hop_tab;
  // ---- polyphase taps
  for (int i = tid; i < OS * p.kp12; i += blockDim.x) {
    const int phs = i / p.kp12, q = i % p.kp12;
    const int n = q * OS + phs;
    hp_[i] = (n <= 2 * Z * OS) ? (p.dec ? p.dec[n] : 1.f) : 0.f;
  }
  // ---- table rows: rows[i][c] = table[lo]*(1-frac) + table[lo+1]*frac of control frame min(ybase+i, Fw-1)
  for (int i = 0; i < kOscRows; ++i) {
    const int f = min(ybase + i, p.Fw - 1);
    const float raw = __fmul_rn(__ldg(w + (size_t)b * p.Fw + f), (float)(n_tab - 1));
    int lo = (int)raw;
    lo = min(max(lo, 0), n_tab - 2);
    const float fr = __fsub_rn(raw, (float)lo), fr1 = __fsub_rn(1.f, fr);
    const float4* t0 = reinterpret_cast<const float4*>(table + (size_t)lo * P);
    const float4* t1 = reinterpret_cast<const float4*>(table + (size_t)(lo + 1) * P);
    float4* dst = reinterpret_cast<float4*>(rows + i * P);
    for (int c = tid; c < P / 4; c += blockDim.x) {
      const float4 a = __ldg(t0 + c), d = __ldg(t1 + c);
      float4 o;
      o.x = __fadd_rn(__fmul_rn(a.x, fr1), __fmul_rn(d.x, fr));
      o.y = __fadd_rn(__fmul_rn(a.y, fr1), __fmul_rn(d.y, fr));
      o.z = __fadd_rn(__fmul_rn(a.z, fr1), __fmul_rn(d.z, fr));
      o.w = __fadd_rn(__fmul_rn(a.w, fr1), __fmul_rn(d.w, fr));
      dst[c] = o;
    }
  }
  pdl_wait();  // everything above read launch inputs only; the knot prefix comes from the previous kernel
  if (threadIdx.x < 32) span_offsets(p.totals + (size_t)blockIdx.y * kPrefSplit, soff_s, p.phase0, blockIdx.y);
  __syncthreads();
  // ---- flow: strip index j <-> output-rate index mj = m0 - Z + j, samples t = mj*OS + phs
  const int phase_hop = p.hp / OS;
  const bool ph1 = phase_hop == 1;                   // sample-rate f0 (training mode): r0 == 0 everywhere
  const int lgP = 31 - __clz(P);                     // P is a power of two (host dispatch)
  const int lg2hp = 31 - __clz(2 * p.hp);
  const bool hp_pow2 = (p.hp & (p.hp - 1)) == 0;
  const float inv_hop_tab = 1.f / (float)p.hop_tab, inv_hp = 1.f / (float)p.hp, inv_span = 1.f / (float)p.span;
  const unsigned long long* __restrict__ prefb = reinterpret_cast<const unsigned long long*>(p.pref) + (size_t)b * p.Np;
  // the knot pair and its prefix are fetched one iteration ahead: the loop is otherwise bound by the
  // latency of these three global loads
  struct Knot {
    int k;
    float fk, fn;
    unsigned long long pk;
  };
  auto fetch = [&](int j) {
    Knot q;
    const int mj = min(max(m0 - Z + j, 0), p.n_out - 1);
    q.k = min(ph1 ? mj : mj / phase_hop, p.Np - 1);
    q.fk = __ldg(ph + q.k), q.fn = __ldg(ph + min(q.k + 1, p.Np - 1));
    q.pk = __ldg(prefb + q.k);
    return q;
  };
  Knot nx = fetch(min(tid, p.plen - 1));
  for (int j = tid; j < p.plen; j += blockDim.x) {
    const int mj = m0 - Z + j;
    const bool in = mj >= 0 && (int64_t)mj * OS < p.N;
    const Knot kn = nx;
    nx = fetch(min(j + (int)blockDim.x, p.plen - 1));
    float v[OS];
#pragma unroll
    for (int phs = 0; phs < OS; ++phs) v[phs] = 0.f;
    if (in) {
      const int k = kn.k;
      const int r0 = ph1 ? 0 : (mj - k * phase_hop) * OS;
      const float xk = div_os(kn.fk, os_f, inv_os_f, pow2), xn = div_os(kn.fn, os_f, inv_os_f, pow2);
      const uint64_t qx = q64_from_float(xk), qn = q64_from_float(xn);
      // slope term per r(r+1): (x_{k+1}-x_k)/(2hp)
      const int64_t dq = (int64_t)(qn - qx);
      const int64_t qq = hp_pow2 ? (dq >> lg2hp) : __float2ll_rn(__ll2float_rn(dq) * p.inv_2hp);
      int sp = (int)((float)k * inv_span);  // k / span without the integer division
      sp -= (sp * p.span > k) ? 1 : 0;
      sp += ((sp + 1) * p.span <= k) ? 1 : 0;
      const uint64_t qp = kn.pk + soff_s[sp];
      // phi(r) = qp + (r+1) qx + qq r (r+1);  phi(r+1) - phi(r) = qx + 2 qq (r+1)
      const uint64_t dd = (uint64_t)(2 * qq);
      uint64_t phi = qp + qx, d = qx + dd;
      if (!ph1) {
        phi = qp + (uint64_t)(r0 + 1) * qx + (uint64_t)(qq * (int64_t)(r0 * (r0 + 1)));
        d = qx + (uint64_t)(2 * qq * (int64_t)(r0 + 1));
      }
      const int t0 = mj * OS;
      const int trel = t0 - trow0;
      auto sample = [&](int phs) -> float {
        // column: top lgP bits of the phase; fraction: the next 32 bits (rounded to float)
        const uint32_t hi = (uint32_t)(phi >> 32), lo = (uint32_t)phi;
        const int c0 = (int)(hi >> (32 - lgP));
        const int c1 = (c0 + 1) & (P - 1);  // column P wraps to column 0
        const float fx = __fmul_rn(__uint2float_rn(__funnelshift_l(lo, hi, lgP)), 2.3283064365386963e-10f);  // * 2^-32
        // row: t / hop_tab relative to the first staged row (0 or 1), fraction (t % hop_tab) / hop_tab
        const int tr = trel + phs;
        const bool up = tr >= p.hop_tab;
        const float fy = __fmul_rn((float)(up ? tr - p.hop_tab : tr), inv_hop_tab);
        const float* r0p = rows + (up ? P : 0);
        const float* r1p = r0p + P;
        const float t00 = r0p[c0], t01 = r0p[c1], t10 = r1p[c0], t11 = r1p[c1];
        const float gx1 = __fsub_rn(1.f, fx), gy1 = __fsub_rn(1.f, fy);
        float val = __fmul_rn(t00, __fmul_rn(gx1, gy1));
        val = __fmaf_rn(t01, __fmul_rn(fx, gy1), val);
        val = __fmaf_rn(t10, __fmul_rn(gx1, fy), val);
        val = __fmaf_rn(t11, __fmul_rn(fx, fy), val);
        if (p.equal_energy) {  // upsampled increment at t: lerp of the knot pair at r / hp
          const float l1 = __fmul_rn((float)(r0 + phs), inv_hp);
          const float inc = __fmaf_rn(__fsub_rn(1.f, l1), xk, __fmul_rn(l1, xn));
          float rs;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(inc));  // one MUFU.RSQ, within 2 ulp of 1/sqrt
          val = __fmul_rn(val, rs);
        }
        phi += d;
        d += dd;
        return val;
      };
      if (t0 + OS <= p.N) {  // all `os` samples exist (every index but possibly the last): no per-sample branch
#pragma unroll
        for (int phs = 0; phs < OS; ++phs) v[phs] = sample(phs);
      } else {
#pragma unroll
        for (int phs = 0; phs < OS; ++phs)
          if (t0 + phs < p.N) v[phs] = sample(phs);
      }
    }
    const int js = fir_sw(j);
#pragma unroll
    for (int phs = 0; phs < OS; ++phs) vp[phs * plen_sw + js] = v[phs];
  }
  __syncthreads();
  const int r0 = tid * kR;  // outputs m0 + r0 .. m0 + r0 + 7
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
#pragma unroll
  for (int phs = 0; phs < OS; ++phs) fir_tile8_sw(vp + phs * plen_sw, r0, hp_ + phs * p.kp12, p.kp12, acc);
  float* ob = p.out + (size_t)b * p.n_out;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (m0 + r0 + i < p.n_out) ob[m0 + r0 + i] = acc[i];
}

// ---- adjoint w.r.t. the table-selection weight -----------------------------------------
// d_w[b,f] = (n_tab-1) * sum_t gv[t] rs(t) * sum_{taps of t in row f} weight * (T[lo_f+1][col] - T[lo_f][col])
// with gv = transposed decimation of the output gradient.  Same tiling as the forward: a CTA owns
// 1024 output-rate indices; gv comes from the polyphase correlations with reversed taps, the table
// geometry from the same phase code; per-row partial sums go through shared memory, one float
// atomicAdd per (CTA, row) to global (order-dependent in the last bit).
struct OscBwdParams {
  OscParams f;          // forward geometry (out unused)
  const float* gout;    // [B, n_out]
  const float* w;       // [B, Fw]
  const float* table;   // [n_tab, P]
  float* d_w;           // [B, Fw], zero-initialised
  float* d_table;       // [n_tab, P], zero-initialised (table adjoint only)
  int n_tab;
};

__global__ void __launch_bounds__(128) osc_dw_kernel(OscBwdParams q) {
  const OscParams& p = q.f;
  extern __shared__ __align__(16) float smem[];
  float* gs = smem;                       // gout[m0 - Z + i]
  float* hr = smem + p.plen;              // [os][kp12] reversed polyphase taps
  float* dws = hr + p.os * p.kp12;        // [Fw] partial sums
  __shared__ unsigned long long soff_s[kPrefSplit];
  if (threadIdx.x < 32 && !p.aten_cpu) span_offsets(p.totals + (size_t)blockIdx.y * kPrefSplit, soff_s, p.phase0, blockIdx.y);
  const int b = blockIdx.y, tid = threadIdx.x;
  const int m0 = blockIdx.x * kOscTile;
  const float* ph = p.phase + (size_t)b * p.Np;
  const float os_f = (float)p.os;
  const int Z = p.zeros;
  const bool pow2 = (p.os & (p.os - 1)) == 0;
  const float inv_os_f = 1.f / os_f, inv_yd = 1.f / p.ydenom;
  const float* gb = q.gout + (size_t)b * p.n_out;
  for (int i = tid; i < p.plen; i += blockDim.x) {
    const int m = m0 - Z + i;
    gs[i] = (m >= 0 && m < p.n_out) ? __ldg(gb + m) : 0.f;
  }
  for (int i = tid; i < p.os * p.kp12; i += blockDim.x) {
    const int phs = i / p.kp12, qr = i % p.kp12;
    const int n = (2 * Z - qr) * p.os + phs;  // reversed: hr[phs][q'] = h[(2Z - q')*os + phs]
    hr[i] = (qr <= 2 * Z && n >= 0 && n <= 2 * Z * p.os) ? (p.dec ? p.dec[n] : 1.f) : 0.f;
  }
  for (int i = tid; i < p.Fw; i += blockDim.x) dws[i] = 0.f;
  __syncthreads();
  const int r0 = tid * kR;
  int cur = -1;          // row whose partial sums live in a0 (row cur) / a1 (row cur+1, clamped)
  float a0 = 0.f, a1 = 0.f;
  auto flush = [&]() {
    if (cur >= 0) {
      atomicAdd(dws + cur, a0);
      atomicAdd(dws + min(cur + 1, p.Fw - 1), a1);
    }
    a0 = a1 = 0.f;
  };
  for (int phs = 0; phs < p.os; ++phs) {
    float gv[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) gv[i] = 0.f;
    fir_tile8(gs + r0, hr + phs * p.kp12, p.kp12, gv);  // gv[(m0+r0+i)*os + phs]
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const int mj = m0 + r0 + i;
      const int t = mj * p.os + phs;
      if (mj >= p.n_out || t >= p.N) continue;
      KnotPhase kp;
      kp.init(p, ph, b, mj, os_f, inv_os_f, pow2, soff_s);
      const float wr = kp.wrapped(p, phs);
      const OscTaps o = osc_taps(p.Fw, p.P, wr, t, inv_yd, p.blocks);
      float g = gv[i];
      if (p.equal_energy) g = __fmul_rn(g, rsqrtf(osc_inc(ph, t, p.scale, p.Np, os_f, inv_os_f, pow2)));
      if (o.row0 != cur) {
        flush();
        cur = o.row0;
      }
      // slope of the row-interpolated table w.r.t. its selection weight: T[lo+1] - T[lo]
      auto slope = [&](int row, int col) -> float {
        const float raw = __fmul_rn(__ldg(q.w + (size_t)b * p.Fw + row), (float)(q.n_tab - 1));
        const int lo = min(max((int)raw, 0), q.n_tab - 2);
        return __ldg(q.table + (size_t)(lo + 1) * p.P + col) - __ldg(q.table + (size_t)lo * p.P + col);
      };
      float c0 = o.w00 * slope(o.row0, o.c0);
      if (o.c1 >= 0) c0 = __fmaf_rn(o.w01, slope(o.row0, o.c1), c0);
      a0 = __fmaf_rn(g, c0, a0);
      if (o.has_row1) {
        float c1 = o.w10 * slope(o.row1, o.c0);
        if (o.c1 >= 0) c1 = __fmaf_rn(o.w11, slope(o.row1, o.c1), c1);
        if (o.row1 == o.row0) a0 = __fmaf_rn(g, c1, a0); else a1 = __fmaf_rn(g, c1, a1);
      }
    }
  }
  flush();
  __syncthreads();
  for (int i = tid; i < p.Fw; i += blockDim.x)
    if (dws[i] != 0.f) atomicAdd(q.d_w + (size_t)b * p.Fw + i, dws[i] * (float)(q.n_tab - 1));
}

// v2 of the weight adjoint (exact-phase mode, compile-time oversampling, power-of-two tables): the same
// restructuring as osc_flow_v2_kernel.  The slope rows T[lo+1] - T[lo] of the (at most three) control frames a
// tile can touch are staged in shared memory, so the four bilinear taps are LDS instead of eight dependent
// L2 gathers per sample; table coordinates come from the integers (column = top bits of the Q0.64 phase,
// row = t / hop_tab by one compare); the knot setup is done once per output index and the phase advances by
// integer increments over its `os` samples; the three per-row sums are reduced by shuffles, one shared-memory
// atomic per warp and row, one global atomic per CTA and row.
//
// TABLE = true is the adjoint w.r.t. the wavetable itself (IndexedGlottalFlowTable(trainable=True), models/synth.py:
// the table is an nn.Parameter there).  The output is bilinear in the table, so every oversampled sample scatters
// g * {1-fy, fy} * {1-p, p} * {1-fx, fx} onto 8 table entries.  The tile touches at most kOscRows control frames and
// two table rows per frame, so the scatter goes into [kOscRows][2][P] shared-memory accumulators (48 KB at P = 2048)
// with shared atomics and is flushed once per CTA with global atomics on the non-zero entries.
template <int OS, bool TABLE>
__global__ void __launch_bounds__(128) osc_dw_v2_kernel(OscBwdParams q) {
  const OscParams& p = q.f;
  extern __shared__ __align__(16) float smem[];
  float* gs = smem;                       // gout[m0 - Z + i]
  float* hr = smem + p.plen;              // [OS][kp12] reversed polyphase taps
  float* slope = hr + OS * p.kp12;        // [kOscRows][P] T[lo+1] - T[lo] of control frames ybase .. ybase+2
                                          // TABLE: [kOscRows][2][P] accumulators for rows lo, lo+1 of those frames
  __shared__ unsigned long long soff_s[kPrefSplit];
  __shared__ float dws[kOscRows];
  if (threadIdx.x < 32) span_offsets(p.totals + (size_t)blockIdx.y * kPrefSplit, soff_s, p.phase0, blockIdx.y);
  const int b = blockIdx.y, tid = threadIdx.x, P = p.P;
  const int m0 = blockIdx.x * kOscTile;
  const float* __restrict__ ph = p.phase + (size_t)b * p.Np;
  const int Z = p.zeros;
  const float os_f = (float)OS, inv_os_f = 1.f / os_f;
  constexpr bool pow2 = (OS & (OS - 1)) == 0;
  const int ybase = (m0 * OS) / p.hop_tab;
  const int trow0 = ybase * p.hop_tab;
  const float* __restrict__ gb = q.gout + (size_t)b * p.n_out;
  for (int i = tid; i < p.plen; i += blockDim.x) {
    const int m = m0 - Z + i;
    gs[i] = (m >= 0 && m < p.n_out) ? __ldg(gb + m) : 0.f;
  }
  for (int i = tid; i < OS * p.kp12; i += blockDim.x) {
    const int phs = i / p.kp12, qr = i % p.kp12;
    const int n = (2 * Z - qr) * OS + phs;  // reversed: hr[phs][q'] = h[(2Z - q')*os + phs]
    hr[i] = (qr <= 2 * Z && n >= 0 && n <= 2 * Z * OS) ? (p.dec ? p.dec[n] : 1.f) : 0.f;
  }
  float pr[kOscRows];  // TABLE: row blend p = w (n_tab-1) - lo of the staged control frames
  int lor[kOscRows];
#pragma unroll
  for (int i = 0; i < kOscRows; ++i) {
    const int f = min(ybase + i, p.Fw - 1);
    const float raw = __fmul_rn(__ldg(q.w + (size_t)b * p.Fw + f), (float)(q.n_tab - 1));
    const int lo = min(max((int)raw, 0), q.n_tab - 2);
    pr[i] = __fsub_rn(raw, (float)lo);
    lor[i] = lo;
    if constexpr (TABLE) {
      float4* dst = reinterpret_cast<float4*>(slope + (size_t)i * 2 * P);
      for (int c = tid; c < P / 2; c += blockDim.x) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* t0 = reinterpret_cast<const float4*>(q.table + (size_t)lo * P);
    const float4* t1 = reinterpret_cast<const float4*>(q.table + (size_t)(lo + 1) * P);
    float4* dst = reinterpret_cast<float4*>(slope + i * P);
    for (int c = tid; c < P / 4; c += blockDim.x) {
      const float4 a = __ldg(t0 + c), d = __ldg(t1 + c);
      dst[c] = make_float4(d.x - a.x, d.y - a.y, d.z - a.z, d.w - a.w);
    }
  }
  if (tid < kOscRows) dws[tid] = 0.f;
  __syncthreads();
  const int r0 = tid * kR;
  // gv[phs][i] = gradient w.r.t. the oversampled sample (m0+r0+i)*OS + phs: transposed decimation
  float gv[OS][kR];
#pragma unroll
  for (int phs = 0; phs < OS; ++phs) {
#pragma unroll
    for (int i = 0; i < kR; ++i) gv[phs][i] = 0.f;
    fir_tile8(gs + r0, hr + phs * p.kp12, p.kp12, gv[phs]);
  }
  const int phase_hop = p.hp / OS;
  const bool ph1 = phase_hop == 1;
  const int lgP = 31 - __clz(P), lg2hp = 31 - __clz(2 * p.hp);
  const bool hp_pow2 = (p.hp & (p.hp - 1)) == 0;
  const float inv_hop_tab = 1.f / (float)p.hop_tab, inv_hp = 1.f / (float)p.hp, inv_span = 1.f / (float)p.span;
  const unsigned long long* __restrict__ prefb = reinterpret_cast<const unsigned long long*>(p.pref) + (size_t)b * p.Np;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;  // sums for staged rows 0, 1, 2
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const int mj = m0 + r0 + i;
    if (mj >= p.n_out || (int64_t)mj * OS >= p.N) continue;
    const int k = min(ph1 ? mj : mj / phase_hop, p.Np - 1);
    const int rr0 = ph1 ? 0 : (mj - k * phase_hop) * OS;
    const float xk = div_os(__ldg(ph + k), os_f, inv_os_f, pow2), xn = div_os(__ldg(ph + min(k + 1, p.Np - 1)), os_f, inv_os_f, pow2);
    const uint64_t qx = q64_from_float(xk), qn = q64_from_float(xn);
    const int64_t dq = (int64_t)(qn - qx);
    const int64_t qq = hp_pow2 ? (dq >> lg2hp) : __float2ll_rn(__ll2float_rn(dq) * p.inv_2hp);
    int sp = (int)((float)k * inv_span);
    sp -= (sp * p.span > k) ? 1 : 0;
    sp += ((sp + 1) * p.span <= k) ? 1 : 0;
    const uint64_t qp = __ldg(prefb + k) + soff_s[sp];
    const uint64_t dd = (uint64_t)(2 * qq);
    uint64_t phi = qp + (uint64_t)(rr0 + 1) * qx + (uint64_t)(qq * (int64_t)(rr0 * (rr0 + 1)));
    uint64_t d = qx + (uint64_t)(2 * qq * (int64_t)(rr0 + 1));
    const int t0 = mj * OS;
#pragma unroll
    for (int phs = 0; phs < OS; ++phs) {
      if (t0 + phs < p.N) {
        const uint32_t hi = (uint32_t)(phi >> 32), lo = (uint32_t)phi;
        const int c0 = (int)(hi >> (32 - lgP));
        const int c1 = (c0 + 1) & (P - 1);
        const float fx = __fmul_rn(__uint2float_rn(__funnelshift_l(lo, hi, lgP)), 2.3283064365386963e-10f);
        const int tr = t0 + phs - trow0;
        const bool up = tr >= p.hop_tab;
        const float fy = __fmul_rn((float)(up ? tr - p.hop_tab : tr), inv_hop_tab);
        float g = gv[phs][i];
        if (p.equal_energy) {
          const float l1 = __fmul_rn((float)(rr0 + phs), inv_hp);
          const float inc = __fmaf_rn(__fsub_rn(1.f, l1), xk, __fmul_rn(l1, xn));
          float rs;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(inc));
          g = __fmul_rn(g, rs);
        }
        const float gx1 = 1.f - fx;
        if constexpr (TABLE) {
          float* A0 = slope + (up ? 2 * P : 0);  // frame `up`: planes lo, lo+1
          float* A1 = A0 + 2 * P;                // next frame
          const float p0 = up ? pr[1] : pr[0], p1 = up ? pr[2] : pr[1];
          const float gl = g * (1.f - fy), gh = g * fy;
          const float gl0 = gl * (1.f - p0), gl1 = gl * p0, gh0 = gh * (1.f - p1), gh1 = gh * p1;
          atomicAdd(A0 + c0, gl0 * gx1);
          atomicAdd(A0 + c1, gl0 * fx);
          atomicAdd(A0 + P + c0, gl1 * gx1);
          atomicAdd(A0 + P + c1, gl1 * fx);
          atomicAdd(A1 + c0, gh0 * gx1);
          atomicAdd(A1 + c1, gh0 * fx);
          atomicAdd(A1 + P + c0, gh1 * gx1);
          atomicAdd(A1 + P + c1, gh1 * fx);
          phi += d;
          d += dd;
          continue;
        }
        const float* s0p = slope + (up ? P : 0);
        const float* s1p = s0p + P;
        const float lo_part = g * (1.f - fy) * __fmaf_rn(fx, s0p[c1], gx1 * s0p[c0]);
        const float hi_part = g * fy * __fmaf_rn(fx, s1p[c1], gx1 * s1p[c0]);
        a0 += up ? 0.f : lo_part;
        a1 += up ? lo_part : hi_part;
        a2 += up ? hi_part : 0.f;
      }
      phi += d;
      d += dd;
    }
  }
  if constexpr (TABLE) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kOscRows; ++i)
      for (int c = tid; c < 2 * P; c += blockDim.x) {
        const float v = slope[(size_t)i * 2 * P + c];
        if (v != 0.f) atomicAdd(q.d_table + (size_t)lor[i] * P + c, v);  // plane 1 is row lo+1: contiguous after row lo
      }
    return;
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, sh);
    a1 += __shfl_xor_sync(0xffffffffu, a1, sh);
    a2 += __shfl_xor_sync(0xffffffffu, a2, sh);
  }
  if ((tid & 31) == 0) {
    atomicAdd(dws + 0, a0);
    atomicAdd(dws + 1, a1);
    atomicAdd(dws + 2, a2);
  }
  __syncthreads();
  if (tid < kOscRows && dws[tid] != 0.f)
    atomicAdd(q.d_w + (size_t)b * p.Fw + min(ybase + tid, p.Fw - 1), dws[tid] * (float)(q.n_tab - 1));
}

struct OscLayout {
  int hp, N, n_out, hop_tab, blocks;
  size_t off_tables, off_pref, bytes;
};

static bool osc_layout(int B, int Np, int phase_hop, int Fw, int P, int os, OscLayout* L) {
  if (B <= 0 || Np <= 0 || phase_hop <= 0 || Fw <= 0 || P <= 0 || os <= 0 || os > kMaxOs) return false;
  const int64_t hp = (int64_t)phase_hop * os, N = (int64_t)(Np - 1) * hp + 1;
  if (N > INT32_MAX / 2) return false;
  L->hp = (int)hp, L->N = (int)N, L->n_out = (int)((N - 1) / os + 1);
  L->off_tables = 0;
  L->off_pref = align_up((size_t)B * Fw * P * 4, 256);
  L->bytes = L->off_pref + align_up((size_t)B * Np * 8, 256) + align_up((size_t)B * kPrefSplit * 8, 256);
  return true;
}

static OscParams osc_params(const float* phase, const float* tables, const double* pref, const unsigned long long* totals,
                            int span, const float* dec_kernel, float* out, int B, int Np, int Fw, int w_hop, int P, int os,
                            int zeros, int accumulate, int flags, const OscLayout& L) {
  OscParams p{};
  p.phase = phase, p.tables = tables, p.pref = pref, p.aten_cpu = accumulate;
  p.totals = totals, p.span = span;
  p.dec = os > 1 ? dec_kernel : nullptr, p.out = out;
  p.B = B, p.Np = Np, p.hp = L.hp, p.N = L.N, p.n_out = L.n_out, p.Fw = Fw, p.P = P;
  p.hop_tab = w_hop * os;
  p.blocks = (L.N + p.hop_tab - 1) / p.hop_tab;
  p.os = os, p.zeros = os > 1 ? zeros : 0, p.equal_energy = flags & 1;
  p.kp12 = ceil_div(2 * p.zeros + 1, 12) * 12;
  p.plen = (int)align_up((size_t)kOscTile + p.kp12 + 24, 4);
  p.scale = lerp_scale(Np, L.hp);
  p.ydenom = (float)((int64_t)p.hop_tab * p.blocks);
  p.inv_2hp = 1.f / (2.f * (float)L.hp);
  return p;
}

static std::atomic<int> g_osc_v2{1};

}  // namespace golf

using namespace golf;

GOLF_API void golf_glottal_osc_set_variant(int v2) { g_osc_v2 = v2 ? 1 : 0; }

GOLF_API size_t golf_glottal_osc_workspace_bytes(int B, int Np, int phase_hop, int Fw, int P, int os) {
  OscLayout L;
  return osc_layout(B, Np, phase_hop, Fw, P, os, &L) ? L.bytes : 0;
}

static int glottal_osc_fwd(const float* phase, const float* w, const float* table, const float* dec_kernel, float* out, int B,
                           int Np, int phase_hop, int Fw, int w_hop, int n_tab, int P, int os, int zeros, int accumulate,
                           int flags, void* workspace, size_t workspace_bytes, void* stream, const double* phase0) {
  if (phase0 && accumulate != 0) return GOLF_ERR_UNSUPPORTED;  // the initial phase rides on the fixed-point prefix
  if (!phase || !w || !table || !out || n_tab < 2 || w_hop <= 0 || zeros < 0) return GOLF_ERR_INVALID;
  if (os > 1 && !dec_kernel) return GOLF_ERR_INVALID;
  OscLayout L;
  if (!osc_layout(B, Np, phase_hop, Fw, P, os, &L)) return GOLF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < L.bytes) return GOLF_ERR_WORKSPACE;
  if (B > 65535) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  float* tables = reinterpret_cast<float*>(ws + L.off_tables);
  double* pref = reinterpret_cast<double*>(ws + L.off_pref);

  if (accumulate != 0 && accumulate != 1) return GOLF_ERR_INVALID;
  unsigned long long* totals = reinterpret_cast<unsigned long long*>(ws + L.off_pref + align_up((size_t)B * Np * 8, 256));
  const int span = ceil_div(Np, kPrefSplit);
  OscParams p = osc_params(phase, tables, pref, totals, span, dec_kernel, out, B, Np, Fw, w_hop, P, os, zeros, accumulate, flags, L);
  p.phase0 = phase0;
  dim3 grid(ceil_div(L.n_out, kOscTile), B);
  // v2: exact-phase mode, compile-time oversampling, table rows built in shared memory.  Needs a tile
  // (plus FIR halo) shorter than one table-row interval so three staged rows always suffice.
  const size_t sm2 = ((size_t)os * (fir_sw(p.plen) + 4 + p.kp12) + (size_t)kOscRows * P) * sizeof(float);
  const bool v2 = g_osc_v2 && accumulate == 0 && (os == 1 || os == 2 || os == 4) && P >= 4 && (P & (P - 1)) == 0 && sm2 <= 200 * 1024 &&
                  (int64_t)p.plen * os < (int64_t)p.hop_tab;
  if (v2) {
    osc_knot_prefix_q64_kernel<<<dim3(kPrefSplit / kPrefWarps, B), 32 * kPrefWarps, 0, st>>>(phase, reinterpret_cast<unsigned long long*>(pref), totals, Np,
                                                                                          L.hp, (float)os, span);
    GOLF_CHECK_LAUNCH();
    static unsigned long long attr = 0;
    if (first_use_on_device(attr)) {
      GOLF_CUDA(cudaFuncSetAttribute(osc_flow_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_flow_v2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_flow_v2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      mark_used_on_device(attr);
    }
    switch (os) {
      case 1: GOLF_CUDA(launch_pdl(osc_flow_v2_kernel<1>, grid, dim3(128), sm2, st, p, w, table, n_tab)); break;
      case 2: GOLF_CUDA(launch_pdl(osc_flow_v2_kernel<2>, grid, dim3(128), sm2, st, p, w, table, n_tab)); break;
      default: GOLF_CUDA(launch_pdl(osc_flow_v2_kernel<4>, grid, dim3(128), sm2, st, p, w, table, n_tab)); break;
    }
    GOLF_CHECK_LAUNCH();
    return GOLF_OK;
  }
  osc_tables_kernel<<<ceil_div(B * Fw * P, 256), 256, 0, st>>>(w, table, tables, B * Fw, n_tab, P);
  GOLF_CHECK_LAUNCH();
  if (accumulate == 0)
    osc_knot_prefix_q64_kernel<<<dim3(kPrefSplit / kPrefWarps, B), 32 * kPrefWarps, 0, st>>>(phase, reinterpret_cast<unsigned long long*>(pref), totals, Np, L.hp,
                                                            (float)os, span);
  else
    osc_knot_prefix_kernel<<<B, 256, 0, st>>>(phase, pref, Np, L.hp, os, 0);
  GOLF_CHECK_LAUNCH();
  const size_t sm = (size_t)os * (p.plen + p.kp12) * sizeof(float);
  if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
  switch (os) {
    case 1: osc_flow_decimate_kernel<1><<<grid, 128, sm, st>>>(p); break;
    case 2: osc_flow_decimate_kernel<2><<<grid, 128, sm, st>>>(p); break;
    case 4: osc_flow_decimate_kernel<4><<<grid, 128, sm, st>>>(p); break;
    default: osc_flow_decimate_kernel<0><<<grid, 128, sm, st>>>(p); break;
  }
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}


GOLF_API int golf_glottal_osc_fwd(const float* phase, const float* w, const float* table, const float* dec_kernel,
                                  float* out, int B, int Np, int phase_hop, int Fw, int w_hop, int n_tab, int P, int os,
                                  int zeros, int accumulate, int flags, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  return glottal_osc_fwd(phase, w, table, dec_kernel, out, B, Np, phase_hop, Fw, w_hop, n_tab, P, os, zeros, accumulate, flags,
                         workspace, workspace_bytes, stream, nullptr);
}

GOLF_API int golf_glottal_osc_fwd_from(const float* phase, const float* w, const float* table, const float* dec_kernel,
                                       float* out, const double* phase0, int B, int Np, int phase_hop, int Fw, int w_hop,
                                       int n_tab, int P, int os, int zeros, int accumulate, int flags, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  return glottal_osc_fwd(phase, w, table, dec_kernel, out, B, Np, phase_hop, Fw, w_hop, n_tab, P, os, zeros, accumulate, flags,
                         workspace, workspace_bytes, stream, phase0);
}

// d_w and d_table are each optional (at least one): the two adjoints are separate launches of one kernel template
static int glottal_osc_bwd(const float* gout, const float* phase, const float* w, const float* table,
                           const float* dec_kernel, float* d_w, float* d_table, int B, int Np, int phase_hop, int Fw, int w_hop,
                           int n_tab, int P, int os, int zeros, int accumulate, int flags, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!gout || !phase || !w || !table || (!d_w && !d_table) || n_tab < 2 || w_hop <= 0 || zeros < 0) return GOLF_ERR_INVALID;
  if (os > 1 && !dec_kernel) return GOLF_ERR_INVALID;
  if (accumulate != 0 && accumulate != 1) return GOLF_ERR_INVALID;
  OscLayout L;
  if (!osc_layout(B, Np, phase_hop, Fw, P, os, &L)) return GOLF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < L.bytes) return GOLF_ERR_WORKSPACE;
  if (B > 65535 || Fw > 1024) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  double* pref = reinterpret_cast<double*>(ws + L.off_pref);
  unsigned long long* totals = reinterpret_cast<unsigned long long*>(ws + L.off_pref + align_up((size_t)B * Np * 8, 256));
  const int span = ceil_div(Np, kPrefSplit);
  // the phase prefix is recomputed (cheap) rather than saved between forward and backward
  if (accumulate == 0)
    osc_knot_prefix_q64_kernel<<<dim3(kPrefSplit / kPrefWarps, B), 32 * kPrefWarps, 0, st>>>(phase, reinterpret_cast<unsigned long long*>(pref), totals, Np, L.hp,
                                                            (float)os, span);
  else
    osc_knot_prefix_kernel<<<B, 256, 0, st>>>(phase, pref, Np, L.hp, os, 0);
  GOLF_CHECK_LAUNCH();
  if (d_w) GOLF_CUDA(cudaMemsetAsync(d_w, 0, (size_t)B * Fw * sizeof(float), st));
  if (d_table) GOLF_CUDA(cudaMemsetAsync(d_table, 0, (size_t)n_tab * P * sizeof(float), st));
  OscBwdParams q{};
  q.f = osc_params(phase, nullptr, pref, totals, span, dec_kernel, nullptr, B, Np, Fw, w_hop, P, os, zeros, accumulate, flags, L);
  q.gout = gout, q.w = w, q.table = table, q.d_w = d_w, q.d_table = d_table, q.n_tab = n_tab;
  dim3 grid(ceil_div(L.n_out, kOscTile), B);
  const size_t smv2 = ((size_t)q.f.plen + (size_t)os * q.f.kp12 + (size_t)kOscRows * P * (d_table ? 2 : 1)) * sizeof(float);
  if (g_osc_v2 && accumulate == 0 && (os == 1 || os == 2 || os == 4) && P >= 4 && (P & (P - 1)) == 0 && smv2 <= 200 * 1024 &&
      (int64_t)kOscTile * os < (int64_t)q.f.hop_tab) {
    static unsigned long long attr = 0;
    if (first_use_on_device(attr)) {
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GOLF_CUDA(cudaFuncSetAttribute(osc_dw_v2_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      mark_used_on_device(attr);
    }
    if (d_w) {
      const size_t smw = smv2 - (d_table ? (size_t)kOscRows * P * sizeof(float) : 0);
      switch (os) {
        case 1: osc_dw_v2_kernel<1, false><<<grid, 128, smw, st>>>(q); break;
        case 2: osc_dw_v2_kernel<2, false><<<grid, 128, smw, st>>>(q); break;
        default: osc_dw_v2_kernel<4, false><<<grid, 128, smw, st>>>(q); break;
      }
      GOLF_CHECK_LAUNCH();
    }
    if (d_table) {
      switch (os) {
        case 1: osc_dw_v2_kernel<1, true><<<grid, 128, smv2, st>>>(q); break;
        case 2: osc_dw_v2_kernel<2, true><<<grid, 128, smv2, st>>>(q); break;
        default: osc_dw_v2_kernel<4, true><<<grid, 128, smv2, st>>>(q); break;
      }
      GOLF_CHECK_LAUNCH();
    }
    return GOLF_OK;
  }
  if (d_table) return GOLF_ERR_UNSUPPORTED;  // table adjoint: integer-phase path only (power-of-two P, os 1/2/4)
  const size_t sm = (size_t)(q.f.plen + os * q.f.kp12 + Fw) * sizeof(float);
  if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
  osc_dw_kernel<<<grid, 128, sm, st>>>(q);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_glottal_osc_bwd_w(const float* gout, const float* phase, const float* w, const float* table,
                                    const float* dec_kernel, float* d_w, int B, int Np, int phase_hop, int Fw, int w_hop,
                                    int n_tab, int P, int os, int zeros, int accumulate, int flags, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!d_w) return GOLF_ERR_INVALID;
  return glottal_osc_bwd(gout, phase, w, table, dec_kernel, d_w, nullptr, B, Np, phase_hop, Fw, w_hop, n_tab, P, os, zeros,
                         accumulate, flags, workspace, workspace_bytes, stream);
}

GOLF_API int golf_glottal_osc_bwd(const float* gout, const float* phase, const float* w, const float* table,
                                  const float* dec_kernel, float* d_w, float* d_table, int B, int Np, int phase_hop, int Fw,
                                  int w_hop, int n_tab, int P, int os, int zeros, int accumulate, int flags, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return glottal_osc_bwd(gout, phase, w, table, dec_kernel, d_w, d_table, B, Np, phase_hop, Fw, w_hop, n_tab, P, os, zeros,
                         accumulate, flags, workspace, workspace_bytes, stream);
}

GOLF_API int golf_wavetable_read_fwd(const float* wrapped, const float* tables, float* out, int B, int N, int R, int P,
                                     int hop_tab, void* stream) {
  if (!wrapped || !tables || !out || B <= 0 || N <= 0 || R <= 0 || P <= 0 || hop_tab <= 0) return GOLF_ERR_INVALID;
  if (B > 65535) return GOLF_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(N, 256), B);
  wavetable_read_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(wrapped, tables, out, N, R, P, hop_tab);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
