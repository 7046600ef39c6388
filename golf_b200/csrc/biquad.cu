// biquad.cu -- the second-order-section row of the hot path:
//
//   * golf_biquad_cascade_bwd : adjoint of BatchSecondOrderLPCSynth.forward (models/lpc.py:94-131: per frame a cascade
//                               of K second-order all-pole sections, lfilter(x, [a0,a1,a2], [1,0,0]) K times, Hann OLA);
//   * golf_lfilter_allpole_*  : the `lfilter`-shaped twin of models/lpc.py:11-16 (lpc_synthesis: x[C,N], gains[C], a[C,M]
//                               -> per-channel LTI all-pole from zero state) and its adjoint (torchaudio's
//                               DifferentiableIIR.backward for b = [gain, 0, ...]);
//   * golf_biquad_params_*    : logits -> stable sections [1,a1,a2] (coef | conj | real, models/utils.py:487-525) ->
//                               order-2K polynomial (biquads2lpc / coeff_product, models/utils.py:444-484), one thread
//                               per frame, forward and adjoint -- what LTVMinimumPhaseFilter*.ctrl runs at frame rate
//                               for the ISMIR-23 checkpoints (models/filters.py:73-78).
#include "lpc_ff.cuh"

namespace golf {

// =====================================================================================================================
// Cascade adjoint.  Geometry of ff_backward_kernel (lpc_ff.cu): a CTA owns 32 consecutive frames of one utterance
// (lane = frame, warp 0 runs the recurrences) and the hop-sized segments those frames fully determine.
//   phase 1: forward cascade; every section's output goes to the CTA's scratch S[j][n][lane] (lane fastest: each
//            warp store is one 128-byte line);
//   phase 2: reversed time, sections K-1 .. 0:  u_j[n] = g_j[n] + na1_j u_j[n+1] + na2_j u_j[n+2],
//            g_{j-1}[n] = b0_j u_j[n];  d_b0_j += u_j[n] x_j[n], d_na1_j += u_j[n] y_j[n-1], d_na2_j += u_j[n] y_j[n-2]
//            (x_j = y_{j-1}, x_0 = gain_k * ex);  g_{-1} is the gradient of the frame's scaled input:
//            d_gain_k += g_{-1}[n] ex[n], d_ex += gain_k g_{-1}[n] (segment accumulators, no atomics).
// lfilter normalises by a0: b0 = 1/a0, na1 = -a1/a0, na2 = -a2/a0, so
//   d_a0 = (-d_b0 + a1 d_na1 ... ) -> see the epilogue.
template <int KP>
__global__ void __launch_bounds__(kFfThreads) biquad_backward_kernel(FfParams p, float* __restrict__ scratch, float* __restrict__ d_bq) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x % p.ctas_per_seq, b = blockIdx.x / p.ctas_per_seq;
  const FfGeom g = ff_geom(p, cta);
  float* __restrict__ strip = smem;                            // [NSTRIP][hop+1] gy / norm, padded coordinates
  float* __restrict__ xstrip = strip + g.NSTRIP * g.seg_stride;   // [NSTRIP][hop+1] ex
  float* __restrict__ acc = xstrip + g.NSTRIP * g.seg_stride;     // [NS][hop+1]     d_ex accumulators
  float* __restrict__ wsm = acc + g.NS * g.seg_stride;            // [win], then the norm row [hop]
  const int K = p.M;
  const int k = g.k0 + lane;
  const bool frame_ok = (k >= 0) && (k < p.n_frames);
  const bool own = frame_ok && (lane >= p.NQ - 1 || cta == 0);
  float* __restrict__ S = scratch + (size_t)blockIdx.x * K * p.win * 32;

  for (int i = tid; i < p.win; i += kFfThreads) wsm[i] = __ldg(p.window + i);
  for (int i = tid; i < g.NS * g.seg_stride; i += kFfThreads) acc[i] = 0.f;
  __syncthreads();
  ff_stage_adjoint_strips<true>(p, g, p.ex + (size_t)b * p.ex_stride, p.vws_ex + (size_t)b * p.ex_stride2, strip, xstrip,
                                wsm + p.win, wsm);

  if (warp == 0) {
    float b0[KP], na1[KP], na2[KP], s1[KP], s2[KP];
    const float* q = p.coef + ((size_t)b * p.F + (frame_ok ? k : 0)) * K * 3;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const bool on = frame_ok && j < K;
      const float a0 = on ? q[3 * j] : 1.f;
      b0[j] = 1.f / a0;
      na1[j] = on ? -(q[3 * j + 1] / a0) : 0.f;
      na2[j] = on ? -(q[3 * j + 2] / a0) : 0.f;
      s1[j] = s2[j] = 0.f;
    }
    const float gframe = frame_ok ? __ldg(p.gain + (size_t)b * p.F + k) : 0.f;
    // ---- phase 1: forward, all section outputs to the scratch
    {
      int q0 = 0, r0 = 0;
#pragma unroll 1
      for (int n = 0; n < p.win; ++n) {
        float x = __fmul_rn(xstrip[(lane + q0) * g.seg_stride + r0], gframe);
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          if (j < K) {
            float v = __fmul_rn(x, b0[j]);
            v = __fmaf_rn(na2[j], s2[j], v);
            v = __fmaf_rn(na1[j], s1[j], v);
            s2[j] = s1[j], s1[j] = v;
            S[((size_t)j * p.win + n) * 32 + lane] = v;
            x = v;
          }
        }
        if (++r0 == p.hop) r0 = 0, ++q0;
      }
    }
    __syncwarp();
    // ---- phase 2: reversed time
    float db0[KP], dn1[KP], dn2[KP], ym1[KP];  // s1/s2 become u[n+1], u[n+2]; ym1 = y_j[n-1] carried to the next step as y_j[n]
    float dg = 0.f;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      db0[j] = dn1[j] = dn2[j] = 0.f, s1[j] = s2[j] = 0.f;
      ym1[j] = (j < K && p.win >= 2) ? S[((size_t)j * p.win + p.win - 2) * 32 + lane] : 0.f;
    }
    int q0 = p.NQ - 1, r0 = p.hop - 1;
#pragma unroll 1
    for (int n = p.win - 1; n >= 0; --n) {
      float gcur = frame_ok ? __fmul_rn(strip[(lane + q0) * g.seg_stride + r0], wsm[n]) : 0.f;
      const float x0 = xstrip[(lane + q0) * g.seg_stride + r0];
#pragma unroll
      for (int j = KP - 1; j >= 0; --j) {
        if (j < K) {
          float u = __fmaf_rn(na2[j], s2[j], gcur);
          u = __fmaf_rn(na1[j], s1[j], u);
          s2[j] = s1[j], s1[j] = u;
          const float xin = j > 0 ? S[((size_t)(j - 1) * p.win + n) * 32 + lane] : __fmul_rn(x0, gframe);
          const float y1 = ym1[j];
          const float y2 = n >= 2 ? S[((size_t)j * p.win + n - 2) * 32 + lane] : 0.f;
          db0[j] = __fmaf_rn(u, xin, db0[j]);
          dn1[j] = __fmaf_rn(u, y1, dn1[j]);
          dn2[j] = __fmaf_rn(u, y2, dn2[j]);
          ym1[j] = y2;  // y_j[(n-1)-1]
          gcur = __fmul_rn(b0[j], u);
        }
      }
      dg = __fmaf_rn(gcur, x0, dg);
      const int sj = lane + q0 - (p.NQ - 1);
      if (frame_ok && sj >= 0 && sj < g.NS) acc[sj * g.seg_stride + r0] += gcur * gframe;
      if (--r0 < 0) r0 = p.hop - 1, --q0;
    }
    if (own) {
      if (p.d_gain) p.d_gain[(size_t)b * p.F + k] = dg;
      if (d_bq) {
        float* dst = d_bq + ((size_t)b * p.F + k) * K * 3;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          if (j < K) {
            // b0 = 1/a0, na1 = -a1/a0, na2 = -a2/a0
            const float inv = b0[j];
            dst[3 * j + 0] = -inv * (db0[j] * inv + dn1[j] * na1[j] + dn2[j] * na2[j]);
            dst[3 * j + 1] = -dn1[j] * inv;
            dst[3 * j + 2] = -dn2[j] * inv;
          }
        }
      }
    }
  }
  __syncthreads();
  ff_write_adjoint_rows(p, g, acc, p.y + (size_t)b * p.Le);
}

template <int KP>
static int launch_biquad_bwd(const FfParams& pb, float* scratch, float* d_bq, cudaStream_t st) {
  const int NS = 33 - pb.NQ, NSTRIP = 32 + pb.NQ - 1;
  const size_t sm = ((size_t)(NS + 2 * NSTRIP) * (pb.hop + 1) + pb.win + pb.hop) * sizeof(float);
  if (sm > 220 * 1024) return GOLF_ERR_UNSUPPORTED;
  static size_t sm_allowed_dev[64];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t& sm_allowed = sm_allowed_dev[dev_ & 63];
  if (sm > 48 * 1024 && sm > sm_allowed) {
    GOLF_CUDA(cudaFuncSetAttribute(biquad_backward_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sm_allowed = sm;
  }
  biquad_backward_kernel<KP><<<pb.B * pb.ctas_per_seq, kFfThreads, sm, st>>>(pb, scratch, d_bq);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

static int biquad_geom(FfParams* p, int B, int T_ex, int F, int K, int hop, int win) {
  if (B <= 0 || T_ex <= 0 || F <= 0 || K <= 0 || hop <= 0 || win < hop) return GOLF_ERR_INVALID;
  if (K > 16 || hop < 16 || (win - hop) % 2 != 0) return GOLF_ERR_UNSUPPORTED;
  p->B = B, p->F = F, p->M = K, p->hop = hop, p->win = win, p->interp_gain = 0;
  int rc = fill_geometry(p, T_ex, (win - hop) / 2);
  if (rc) return rc;
  // the adjoint covers every input position [0, T_ex)
  p->nseg = (p->pad + p->Le - 1) / hop - p->nseg0 + 1;
  p->ctas_per_seq = ceil_div(p->nseg, 33 - p->NQ);
  return GOLF_OK;
}

// =====================================================================================================================
// lfilter-shaped all-pole filter: lane = channel, 32 channels per warp, MP steps per tile with static register
// rotation; tiles are transposed through shared memory so that global accesses run along time.
//   REVERSE: time runs backwards (position N-1-n): the same kernel is the adjoint recurrence.
template <int MP>
__global__ void __launch_bounds__(32) lfilter_allpole_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ gain,
                                                             const float* __restrict__ a, float* __restrict__ y, int C, int N, int M,
                                                             int reverse) {
  __shared__ float tile[32 * (MP + 1)];
  const int lane = threadIdx.x;
  const int c0 = blockIdx.x * 32, c = c0 + lane;
  const bool ok = c < C;
  AllPole<MP> f;
  {
    const float* ar = a + (size_t)(ok ? c : 0) * M;
#pragma unroll
    for (int j = 0; j < MP; ++j) {
      const int i = MP - 1 - j;
      f.na[j] = (ok && i < M) ? -__ldg(ar + i) : 0.f;
      f.h[j] = 0.f;
    }
  }
  const float gch = (gain && ok) ? __ldg(gain + c) : 1.f;
#pragma unroll 1
  for (int n0 = 0; n0 < N; n0 += MP) {
    for (int i = lane; i < 32 * MP; i += 32) {
      const int rr = i / MP, s = i - rr * MP;
      const int n = n0 + s;
      const int pos = reverse ? N - 1 - n : n;
      tile[rr * (MP + 1) + s] = (c0 + rr < C && n < N) ? __ldg(x + (size_t)(c0 + rr) * x_stride + pos) : 0.f;
    }
    __syncwarp();
    float xs[MP], ys[MP];
#pragma unroll
    for (int s = 0; s < MP; ++s) xs[s] = __fmul_rn(tile[lane * (MP + 1) + s], gch);
    TileSteps<AllPole<MP>, 0, MP>::run(f, xs, ys);
#pragma unroll
    for (int s = 0; s < MP; ++s) tile[lane * (MP + 1) + s] = ys[s];
    __syncwarp();
    for (int i = lane; i < 32 * MP; i += 32) {
      const int rr = i / MP, s = i - rr * MP;
      const int n = n0 + s;
      const int pos = reverse ? N - 1 - n : n;
      if (c0 + rr < C && n < N) y[(size_t)(c0 + rr) * N + pos] = tile[rr * (MP + 1) + s];
    }
    __syncwarp();
  }
}

// one warp per channel: d_x = gain * u, d_gain = sum_n u[n] x[n], d_a[i] = -sum_n u[n] y[n-1-i]
__global__ void __launch_bounds__(32) lfilter_grad_kernel(const float* __restrict__ u, const float* __restrict__ x, int64_t x_stride,
                                                          const float* __restrict__ y, const float* __restrict__ gain,
                                                          float* __restrict__ d_x, float* __restrict__ d_gain, float* __restrict__ d_a,
                                                          int C, int N, int M) {
  const int c = blockIdx.x, lane = threadIdx.x;
  const float* uc = u + (size_t)c * N;
  const float* xc = x + (size_t)c * x_stride;
  const float* yc = y + (size_t)c * N;
  const float gch = gain ? __ldg(gain + c) : 1.f;
  float dg = 0.f;
  float da[40];
#pragma unroll
  for (int i = 0; i < 40; ++i) da[i] = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float un = uc[n];
    dg = __fmaf_rn(un, xc[n], dg);
    if (d_x) d_x[(size_t)c * N + n] = __fmul_rn(un, gch);
    if (d_a) {
#pragma unroll
      for (int i = 0; i < 40; ++i)
        if (i < M && n - 1 - i >= 0) da[i] = __fmaf_rn(un, yc[n - 1 - i], da[i]);
    }
  }
  for (int o = 16; o; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o);
  if (d_gain && lane == 0) d_gain[c] = dg;
  if (d_a) {
#pragma unroll
    for (int i = 0; i < 40; ++i) {
      if (i < M) {
        float v = da[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) d_a[(size_t)c * M + i] = -v;
      }
    }
  }
}

template <int MP>
static int launch_lfilter(const float* x, int64_t xs, const float* gain, const float* a, float* y, int C, int N, int M, int reverse,
                          cudaStream_t st) {
  lfilter_allpole_kernel<MP><<<ceil_div(C, 32), 32, 0, st>>>(x, xs, gain, a, y, C, N, M, reverse);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
static int dispatch_lfilter(const float* x, int64_t xs, const float* gain, const float* a, float* y, int C, int N, int M, int reverse,
                            cudaStream_t st) {
  if (M <= 4) return launch_lfilter<4>(x, xs, gain, a, y, C, N, M, reverse, st);
  if (M <= 8) return launch_lfilter<8>(x, xs, gain, a, y, C, N, M, reverse, st);
  if (M <= 12) return launch_lfilter<12>(x, xs, gain, a, y, C, N, M, reverse, st);
  if (M <= 16) return launch_lfilter<16>(x, xs, gain, a, y, C, N, M, reverse, st);
  if (M <= 24) return launch_lfilter<24>(x, xs, gain, a, y, C, N, M, reverse, st);
  if (M <= 32) return launch_lfilter<32>(x, xs, gain, a, y, C, N, M, reverse, st);
  return launch_lfilter<40>(x, xs, gain, a, y, C, N, M, reverse, st);
}

// =====================================================================================================================
// Biquad parameterisations + polynomial product, one thread per frame.
//   rep 0 "coef": a1 = 2 rho tanh(l0),            a2 = ((2 - |a1|) rho tanh(l1) + |a1|) / 2
//   rep 1 "conj": m = rho sigmoid(l0), a1 = -2 m tanh(l1), a2 = m^2
//   rep 2 "real": z1 = rho tanh(l0), z2 = rho tanh(l1),    a1 = -z1 - z2, a2 = z1 z2
constexpr int kMaxSections = 16;

__device__ __forceinline__ void section_fwd(int rep, float rho, float l0, float l1, float* a1, float* a2) {
  if (rep == 0) {
    const float v = tanhf(l0) * rho * 2.f;
    const float m = fabsf(v);
    *a1 = v;
    *a2 = 0.5f * ((2.f - m) * tanhf(l1) * rho + m);
  } else if (rep == 1) {
    const float m = (1.f / (1.f + expf(-l0))) * rho;
    *a1 = -2.f * m * tanhf(l1);
    *a2 = m * m;
  } else {
    const float z1 = tanhf(l0) * rho, z2 = tanhf(l1) * rho;
    *a1 = -z1 - z2;
    *a2 = z1 * z2;
  }
}

// (d_a1, d_a2) -> (d_l0, d_l1)
__device__ __forceinline__ void section_bwd(int rep, float rho, float l0, float l1, float d1, float d2, float* g0, float* g1) {
  if (rep == 0) {
    const float t0 = tanhf(l0), t1 = tanhf(l1);
    const float v = t0 * rho * 2.f;
    const float sgn = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
    const float m = fabsf(v);
    // a2 = 0.5 ((2 - m) t1 rho + m):  d/dm = 0.5 (1 - t1 rho);  d/dt1 = 0.5 (2 - m) rho
    const float dv = d1 + d2 * 0.5f * (1.f - t1 * rho) * sgn;
    *g0 = dv * 2.f * rho * (1.f - t0 * t0);
    *g1 = d2 * 0.5f * (2.f - m) * rho * (1.f - t1 * t1);
  } else if (rep == 1) {
    const float s = 1.f / (1.f + expf(-l0)), t1 = tanhf(l1);
    const float m = s * rho;
    const float dm = d1 * (-2.f * t1) + d2 * 2.f * m;
    *g0 = dm * rho * s * (1.f - s);
    *g1 = d1 * (-2.f * m) * (1.f - t1 * t1);
  } else {
    const float t0 = tanhf(l0), t1 = tanhf(l1);
    const float z1 = t0 * rho, z2 = t1 * rho;
    const float dz1 = -d1 + d2 * z2, dz2 = -d1 + d2 * z1;
    *g0 = dz1 * rho * (1.f - t0 * t0);
    *g1 = dz2 * rho * (1.f - t1 * t1);
  }
}

__global__ void biquad_params_kernel(const float* __restrict__ logits, float* __restrict__ biquads, float* __restrict__ a, int N, int K,
                                     int rep, float rho) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  float poly[2 * kMaxSections + 1];
  poly[0] = 1.f;
  for (int i = 1; i <= 2 * K; ++i) poly[i] = 0.f;
  const float* lg = logits + (size_t)f * K * 2;
  for (int j = 0; j < K; ++j) {
    float a1, a2;
    section_fwd(rep, rho, lg[2 * j], lg[2 * j + 1], &a1, &a2);
    if (biquads) {
      float* q = biquads + ((size_t)f * K + j) * 3;
      q[0] = 1.f, q[1] = a1, q[2] = a2;
    }
    // poly *= (1 + a1 z^-1 + a2 z^-2): degree 2j -> 2j+2, in place from the top
    for (int n = 2 * j + 2; n >= 1; --n) {
      float v = poly[n];
      v = __fmaf_rn(a1, poly[n - 1], v);
      if (n >= 2) v = __fmaf_rn(a2, poly[n - 2], v);
      poly[n] = v;
    }
  }
  if (a)
    for (int i = 0; i < 2 * K; ++i) a[(size_t)f * 2 * K + i] = poly[i + 1];
}

// d_logits from d_a (through the product) and / or d_biquads (sections used directly)
__global__ void biquad_params_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ d_biquads, const float* __restrict__ d_a,
                                         float* __restrict__ d_logits, int N, int K, int rep, float rho) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  constexpr int D = 2 * kMaxSections + 1;
  float pref[(kMaxSections + 1) * D];  // pref[j] = product of sections 0..j-1 (local memory; frame rate)
  float a1s[kMaxSections], a2s[kMaxSections];
  const float* lg = logits + (size_t)f * K * 2;
  for (int i = 0; i < D; ++i) pref[i] = 0.f;
  pref[0] = 1.f;
  for (int j = 0; j < K; ++j) {
    section_fwd(rep, rho, lg[2 * j], lg[2 * j + 1], &a1s[j], &a2s[j]);
    const float* src = pref + j * D;
    float* dst = pref + (j + 1) * D;
    for (int n = 0; n < D; ++n) {
      float v = src[n];
      if (n >= 1) v = __fmaf_rn(a1s[j], src[n - 1], v);
      if (n >= 2) v = __fmaf_rn(a2s[j], src[n - 2], v);
      dst[n] = v;
    }
  }
  float dP[D], dQ[D];
  dP[0] = 0.f;
  for (int i = 1; i < D; ++i) dP[i] = (d_a && i <= 2 * K) ? d_a[(size_t)f * 2 * K + i - 1] : 0.f;
  for (int j = K - 1; j >= 0; --j) {
    const float* src = pref + j * D;  // P_j (degree 2j), P_{j+1} = P_j * q_j
    float d1 = 0.f, d2 = 0.f;
    for (int n = 1; n <= 2 * j + 2; ++n) {
      d1 = __fmaf_rn(dP[n], src[n - 1], d1);
      if (n >= 2) d2 = __fmaf_rn(dP[n], src[n - 2], d2);
    }
    for (int n = 0; n <= 2 * j; ++n) dQ[n] = dP[n] + a1s[j] * dP[n + 1] + a2s[j] * dP[n + 2];
    for (int n = 0; n <= 2 * j; ++n) dP[n] = dQ[n];
    for (int n = 2 * j + 1; n < D; ++n) dP[n] = 0.f;
    if (d_biquads) {
      const float* q = d_biquads + ((size_t)f * K + j) * 3;
      d1 += q[1], d2 += q[2];
    }
    float g0, g1;
    section_bwd(rep, rho, lg[2 * j], lg[2 * j + 1], d1, d2, &g0, &g1);
    d_logits[((size_t)f * K + j) * 2] = g0;
    d_logits[((size_t)f * K + j) * 2 + 1] = g1;
  }
}

}  // namespace golf

using namespace golf;

GOLF_API int golf_biquad_cascade_fwd(const float* ex, int64_t ex_stride, const float* gain, const float* biquads, const float* window,
                                     float* y, int B, int T_ex, int F, int K, int hop, int win, void* stream) {
  return golf_biquad_ff_fwd(ex, ex_stride, gain, biquads, window, y, B, T_ex, F, K, hop, win, stream);
}

GOLF_API size_t golf_biquad_cascade_bwd_workspace_bytes(int B, int T_ex, int F, int K, int hop, int win) {
  FfParams p{};
  if (biquad_geom(&p, B, T_ex, F, K, hop, win)) return 0;
  return align_up((size_t)B * p.ctas_per_seq * K * win * 32 * sizeof(float), 256);
}

GOLF_API int golf_biquad_cascade_bwd(const float* gy, const float* ex, int64_t ex_stride, const float* gain, const float* biquads,
                                     const float* window, float* d_ex, float* d_gain, float* d_biquads, int B, int T_ex, int F, int K,
                                     int hop, int win, void* workspace, size_t workspace_bytes, void* stream) {
  if (!gy || !ex || !gain || !biquads || !window || !d_ex) return GOLF_ERR_INVALID;
  FfParams p{};
  int rc = biquad_geom(&p, B, T_ex, F, K, hop, win);
  if (rc) return rc;
  const size_t need = golf_biquad_cascade_bwd_workspace_bytes(B, T_ex, F, K, hop, win);
  if (!workspace || workspace_bytes < need) return GOLF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_gain) GOLF_CUDA(cudaMemsetAsync(d_gain, 0, (size_t)B * F * sizeof(float), st));
  if (d_biquads) GOLF_CUDA(cudaMemsetAsync(d_biquads, 0, (size_t)B * F * K * 3 * sizeof(float), st));
  p.ex = gy, p.ex_stride = p.out_len, p.vws_ex = ex, p.ex_stride2 = ex_stride, p.gain = gain, p.coef = biquads, p.window = window;
  p.y = d_ex, p.d_gain = d_gain;
  float* scratch = reinterpret_cast<float*>(workspace);
  if (K <= 4) return launch_biquad_bwd<4>(p, scratch, d_biquads, st);
  if (K <= 8) return launch_biquad_bwd<8>(p, scratch, d_biquads, st);
  if (K <= 12) return launch_biquad_bwd<12>(p, scratch, d_biquads, st);
  return launch_biquad_bwd<16>(p, scratch, d_biquads, st);
}

GOLF_API int golf_lfilter_allpole_fwd(const float* x, int64_t x_stride, const float* gain, const float* a, float* y, int C, int N, int M,
                                      void* stream) {
  if (!x || !a || !y || C <= 0 || N <= 0 || M <= 0 || x_stride < N) return GOLF_ERR_INVALID;
  if (M > 40) return GOLF_ERR_UNSUPPORTED;
  return dispatch_lfilter(x, x_stride, gain, a, y, C, N, M, 0, (cudaStream_t)stream);
}

GOLF_API size_t golf_lfilter_allpole_bwd_workspace_bytes(int C, int N) {
  if (C <= 0 || N <= 0) return 0;
  return align_up((size_t)C * N * sizeof(float), 256);
}

GOLF_API int golf_lfilter_allpole_bwd(const float* gy, const float* x, int64_t x_stride, const float* y, const float* gain, const float* a,
                                      float* d_x, float* d_gain, float* d_a, int C, int N, int M, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  if (!gy || !x || !y || !a || C <= 0 || N <= 0 || M <= 0 || x_stride < N) return GOLF_ERR_INVALID;
  if (M > 40) return GOLF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < golf_lfilter_allpole_bwd_workspace_bytes(C, N)) return GOLF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* u = reinterpret_cast<float*>(workspace);
  // u[n] = gy[n] - sum_i a_i u[n+1+i]: the same recurrence on reversed time
  int rc = dispatch_lfilter(gy, N, nullptr, a, u, C, N, M, 1, st);
  if (rc) return rc;
  lfilter_grad_kernel<<<C, 32, 0, st>>>(u, x, x_stride, y, gain, d_x, d_gain, d_a, C, N, M);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_biquad_params_fwd(const float* logits, float* biquads, float* a, int N, int K, int rep, float max_abs_pole,
                                    void* stream) {
  if (!logits || (!biquads && !a) || N <= 0 || K <= 0 || rep < 0 || rep > 2) return GOLF_ERR_INVALID;
  if (K > kMaxSections) return GOLF_ERR_UNSUPPORTED;
  biquad_params_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(logits, biquads, a, N, K, rep, max_abs_pole);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_biquad_params_bwd(const float* logits, const float* d_biquads, const float* d_a, float* d_logits, int N, int K, int rep,
                                    float max_abs_pole, void* stream) {
  if (!logits || (!d_biquads && !d_a) || !d_logits || N <= 0 || K <= 0 || rep < 0 || rep > 2) return GOLF_ERR_INVALID;
  if (K > kMaxSections) return GOLF_ERR_UNSUPPORTED;
  biquad_params_bwd_kernel<<<ceil_div(N, 64), 64, 0, (cudaStream_t)stream>>>(logits, d_biquads, d_a, d_logits, N, K, rep, max_abs_pole);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
