// lpc_ff.cuh -- geometry and per-frame filter state shared by the frame-wise (GOLF-ff family) kernels:
// lpc_ff.cu (LTVMinimumPhaseFilter, BatchLPCSynth) and biquad.cu (BatchSecondOrderLPCSynth adjoint).
#pragma once
#include "common.cuh"

namespace golf {

#ifndef GOLF_FF_THREADS
#define GOLF_FF_THREADS 128
#endif
constexpr int kFfThreads = GOLF_FF_THREADS;  // warp 0 runs the recurrences; all warps stage and write out

struct FfParams {
  const float* ex;       // fwd: excitation [B, ex_stride]; bwd: gy [B, out_len]
  int64_t ex_stride;
  const float* gain;     // [B,F]
  const float* coef;     // all-pole: a [B,F,M]; biquad: [B,F,K,3]
  const float* window;   // [win]
  float* y;              // fwd: [B, out_len]; bwd: d_e [B, Le]
  float* vws;            // [B*n_frames, win] frame outputs (fwd STORE_V writes, bwd reads)
  float* d_a;            // bwd: [B,F,M]
  float* d_gain;         // bwd with per-frame gain (interp_gain == 0): [B,F]
  const float* vws_ex;   // bwd with per-frame gain: the excitation [B, ex_stride2]
  int64_t ex_stride2;
  int B, Le, F, M, hop, win, NQ, pad, n_frames, out_len, nseg0, nseg, ctas_per_seq;
  int interp_gain;       // 1: strip holds ex*up(gain) (ff); 0: gain applied per frame (biquad synth)
  float scale;
};

// ---- per-frame filters ------------------------------------------------------------
// EXACT: libtorchaudio's CPU arithmetic bit for bit -- one accumulator, taps oldest first, a rounded multiply and a rounded
// subtract per tap (the arithmetic the test-suite pins bitwise against torchaudio.functional.lfilter on the CPU); the
// default sums three interleaved FMA chains (a third of the dependent latency, float32 rounding in a different order).
template <int MP, bool EXACT = false>
struct AllPole {
  static constexpr int TILE = MP;
  float na[MP];  // na[j] = -a[MP-1-j]: index j pairs with the output MP-j steps back (oldest first)
  float h[MP];   // h[s] = output of tile position s (static rotation)
  __device__ __forceinline__ void load(const FfParams& p, int b, int k, bool ok) {
    const float* a = p.coef + ((size_t)b * p.F + (ok ? k : 0)) * p.M;
#pragma unroll
    for (int j = 0; j < MP; ++j) {
      const int i = MP - 1 - j;  // tap index (a[i] multiplies y[n-1-i])
      na[j] = (ok && i < p.M) ? -__ldg(a + i) : 0.f;
      h[j] = 0.f;
    }
  }
  template <int S>
  __device__ __forceinline__ float step(float x) {
    if (EXACT) {
      float acc = x;
#pragma unroll
      for (int j = 0; j < MP; ++j) acc = __fadd_rn(acc, __fmul_rn(na[j], h[(S + j) % MP]));  // acc - a*y: the negation is exact
      h[S] = acc;
      return acc;
    }
    float acc0 = x, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
    for (int j = 0; j < MP - 1; ++j) {  // y[n-MP+j] sits in slot (S+j)%MP
      if (j % 3 == 0) acc0 = __fmaf_rn(na[j], h[(S + j) % MP], acc0);
      if (j % 3 == 1) acc1 = __fmaf_rn(na[j], h[(S + j) % MP], acc1);
      if (j % 3 == 2) acc2 = __fmaf_rn(na[j], h[(S + j) % MP], acc2);
    }
    const float y = __fmaf_rn(na[MP - 1], h[(S + MP - 1) % MP], (acc0 + acc1) + acc2);
    h[S] = y;
    return y;
  }
};

template <int KP>
struct BiquadCascade {
  static constexpr int TILE = 16;
  float b0[KP], na1[KP], na2[KP], y1[KP], y2[KP];
  int K;
  __device__ __forceinline__ void load(const FfParams& p, int b, int k, bool ok) {
    K = p.M;
    const float* q = p.coef + ((size_t)b * p.F + (ok ? k : 0)) * p.M * 3;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const bool on = ok && j < p.M;
      const float a0 = on ? q[3 * j] : 1.f;
      b0[j] = 1.f / a0;
      na1[j] = on ? -(q[3 * j + 1] / a0) : 0.f;
      na2[j] = on ? -(q[3 * j + 2] / a0) : 0.f;
      y1[j] = y2[j] = 0.f;
    }
  }
  template <int S>
  __device__ __forceinline__ float step(float x) {
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      if (j < K) {
        float acc = __fmul_rn(x, b0[j]);
        acc = __fmaf_rn(na2[j], y2[j], acc);
        acc = __fmaf_rn(na1[j], y1[j], acc);
        y2[j] = y1[j];
        y1[j] = acc;
        x = acc;
      }
    }
    return x;
  }
};

template <class Filt, int S, int N>
struct TileSteps {
  __device__ __forceinline__ static void run(Filt& f, const float* xs, float* ys) {
    ys[S] = f.template step<S>(xs[S]);
    if constexpr (S + 1 < N) TileSteps<Filt, S + 1, N>::run(f, xs, ys);
  }
};

struct FfGeom {
  int NS, NSTRIP, seg_stride, P0, k0;
};
__device__ __forceinline__ FfGeom ff_geom(const FfParams& p, int w) {
  FfGeom g;
  g.NS = 33 - p.NQ;          // complete segments per CTA
  g.NSTRIP = 32 + p.NQ - 1;  // segments the CTA's 32 frames touch
  g.seg_stride = p.hop + 1;
  g.P0 = p.nseg0 + w * g.NS;  // first padded segment owned by this CTA
  g.k0 = g.P0 - (p.NQ - 1);   // frame handled by lane 0
  return g;
}
// overlap-added window at padded segment P, offset r
__device__ __forceinline__ float ff_norm(const FfParams& p, const float* wsm, int P, int r) {
  float norm = 0.f;
  for (int q = p.NQ - 1; q >= 0; --q) {  // frame P-q contributes its q-th hop of the window
    const int kk = P - q;
    if (kk >= 0 && kk < p.n_frames) norm += wsm[q * p.hop + r];
  }
  return norm;
}

// ---- staging / write-out shared by the adjoint kernels (lpc_ff.cu, biquad.cu) ------------------------------------------
// strip = gy / (overlap-added window), xstrip (optional) = excitation, both in padded coordinates.  The copies go out
// through cp.async (LDGSTS), all in flight at once; the division follows in a pass over shared memory, interior segments
// sharing the norm row `nrow` [hop].  (A load -> divide -> store loop paid one L2 round trip per element, 66 times per
// thread at hop 240.)  wsm must be complete (barrier) on entry; ends with a barrier.
template <bool WITH_EX>
__device__ __forceinline__ void ff_stage_adjoint_strips(const FfParams& p, const FfGeom& g, const float* __restrict__ gyb,
                                                        const float* __restrict__ exb, float* __restrict__ strip,
                                                        float* __restrict__ xstrip, float* __restrict__ nrow,
                                                        const float* __restrict__ wsm) {
  constexpr int kWarps = kFfThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int sg = warp; sg < g.NSTRIP; sg += kWarps) {
    const int o0 = (g.k0 + sg) * p.hop - p.pad;
    float* __restrict__ row = strip + sg * g.seg_stride;
    for (int r = lane; r < p.hop; r += 32) {
      const int o = o0 + r;
      cp_async4(row + r, gyb + min(max(o, 0), p.out_len - 1), o >= 0 && o < p.out_len);
      if (WITH_EX) cp_async4(xstrip + sg * g.seg_stride + r, exb + min(max(o, 0), p.Le - 1), o >= 0 && o < p.Le);
    }
  }
  for (int r = tid; r < p.hop; r += kFfThreads) {
    float norm = 0.f;
    for (int q = p.NQ - 1; q >= 0; --q) norm += wsm[q * p.hop + r];  // ff_norm's order
    nrow[r] = norm;
  }
  cp_async_wait_all();
  __syncthreads();
  for (int sg = warp; sg < g.NSTRIP; sg += kWarps) {
    const int P = g.k0 + sg;
    const int o0 = P * p.hop - p.pad;
    float* __restrict__ row = strip + sg * g.seg_stride;
    if (P - (p.NQ - 1) >= 0 && P < p.n_frames) {
#pragma unroll 4
      for (int r = lane; r < p.hop; r += 32) row[r] = __fdiv_rn(row[r], nrow[r]);  // 0 / norm stays 0 outside the signal
    } else {
#pragma unroll 1
      for (int r = lane; r < p.hop; r += 32) {
        const int o = o0 + r;
        if (o >= 0 && o < p.out_len) row[r] = __fdiv_rn(row[r], ff_norm(p, wsm, P, r));
      }
    }
  }
  __syncthreads();
}

// the CTA's NS accumulator rows -> d_e [Le] (positions outside the signal dropped)
__device__ __forceinline__ void ff_write_adjoint_rows(const FfParams& p, const FfGeom& g, const float* __restrict__ acc,
                                                      float* __restrict__ deb) {
  constexpr int kWarps = kFfThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sj = warp; sj < g.NS; sj += kWarps) {
    const int pos0 = (g.P0 + sj) * p.hop - p.pad;
    const float* __restrict__ arow = acc + sj * g.seg_stride;
#pragma unroll 4
    for (int r = lane; r < p.hop; r += 32) {
      const int pos = pos0 + r;
      if (pos >= 0 && pos < p.Le) deb[pos] = arow[r];
    }
  }
}

static inline size_t ff_smem_bytes(const FfParams& p, int vt_floats) {
  const int NS = 33 - p.NQ, NSTRIP = 32 + p.NQ - 1;
  return ((size_t)(NS + NSTRIP) * (p.hop + 1) + p.win + vt_floats) * sizeof(float);
}

static inline int fill_geometry(FfParams* p, int T_ex, int pad) {
  if (p->win % p->hop != 0) return GOLF_ERR_UNSUPPORTED;
  p->NQ = p->win / p->hop;
  if (p->NQ < 2 || p->NQ > 8) return GOLF_ERR_UNSUPPORTED;
  p->pad = pad;
  const int64_t up = (int64_t)(p->F - 1) * p->hop + 1;
  p->Le = p->interp_gain ? (int)(T_ex < up ? T_ex : up) : T_ex;
  if (p->Le + 2 * pad < p->win) return GOLF_ERR_INVALID;
  p->n_frames = (p->Le + 2 * pad - p->win) / p->hop + 1;
  if (p->n_frames > p->F) return GOLF_ERR_INVALID;  // the reference asserts the same
  p->out_len = (p->n_frames - 1) * p->hop + p->win - 2 * pad;
  if (p->out_len <= 0) return GOLF_ERR_INVALID;
  p->nseg0 = pad / p->hop;
  const int last = (pad + p->out_len - 1) / p->hop;
  p->nseg = last - p->nseg0 + 1;
  p->ctas_per_seq = ceil_div(p->nseg, 33 - p->NQ);
  p->scale = lerp_scale(p->F, p->hop);
  return GOLF_OK;
}

}  // namespace golf
