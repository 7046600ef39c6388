// lpc_ff.cu -- GOLF-ff: frame-wise LTI all-pole filtering with Hann overlap-add, the
// cascaded-biquad variant, and the inverse (analysis) filter.
//
// Replaces models/filters.py:131-184 (LTVMinimumPhaseFilter.forward: ex*gain, pad,
// unfold(win, hop), models/lpc.py:11-16 lpc_synthesis -> torchaudio lfilter per frame,
// conv_transpose1d OLA with a dense diag(window) kernel, divide by the OLA'd window),
// models/lpc.py:94-131 (BatchSecondOrderLPCSynth) and models/filters.py:186-195.
//
// Mapping: frames are independent, each a serial recurrence of `win` steps from zero
// state -- one LANE per frame, coefficients and filter state in registers (static
// rotation, loop unrolled by the padded order MP), taps visited oldest-first exactly
// like libtorchaudio's CPU loop so consecutive steps overlap in the FMA pipe.
// A warp owns 32 consecutive frames of one utterance and the 33-NQ padded-coordinate
// segments (hop samples each, NQ = win/hop) those frames fully determine, so the
// overlap-add is a shared-memory accumulate with no atomics and no 4x unfold in HBM:
//   * the warp's excitation strip (ex*up(gain), computed once per sample) is staged in
//     shared memory, segment stride hop+1 so the per-lane strided reads are bank-
//     conflict free;
//   * lane l at step n = q*hop + r adds window[n]*y to segment l+q-(NQ-1), offset r;
//   * write-out divides by the overlap-added window and stores coalesced.
// Adjacent warps recompute NQ-1 frames of overlap (10% at NQ=4) instead of exchanging
// partial sums: deterministic and launch-local.
//
// Algorithmic HBM bytes per output sample: 4 (ex) + 4 (y) + 4(M+1)/hop = 8.383 B.
#include "common.cuh"

namespace golf {

struct FfParams {
  const float* ex;
  int64_t ex_stride;
  const float* gain;    // [B,F]
  const float* coef;    // all-pole: a [B,F,M]; biquad: [B,F,K,3]
  const float* window;  // [win]
  float* y;             // [B, out_len]
  int B, Le, F, M, hop, win, NQ, pad, n_frames, out_len, nseg0, nseg, warps_per_seq;
  int interp_gain;      // 1: strip holds ex*up(gain) (ff); 0: gain applied per frame (biquad synth)
  float scale;
};

// ---- per-frame filters ------------------------------------------------------------
template <int MP>
struct AllPole {
  static constexpr int TILE = MP;
  float na[MP];  // na[j] = -a[M-1-j'] arranged oldest-first: index j pairs with y[n-MP+j]
  float h[MP];   // h[s] = output of tile position s (static rotation)
  __device__ __forceinline__ void load(const FfParams& p, int b, int k, bool ok) {
    const float* a = p.coef + ((size_t)b * p.F + (ok ? k : 0)) * p.M;
#pragma unroll
    for (int j = 0; j < MP; ++j) {
      const int i = MP - 1 - j;  // tap index (a[i] multiplies y[n-1-i])
      na[j] = (ok && i < p.M) ? -a[i] : 0.f;
      h[j] = 0.f;
    }
  }
  template <int S>
  __device__ __forceinline__ float step(float x) {
    float acc = x;
#pragma unroll
    for (int j = 0; j < MP; ++j) acc = __fmaf_rn(na[j], h[(S + j) % MP], acc);  // y[n-MP+j] sits in slot (S+j)%MP
    h[S] = acc;
    return acc;
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

template <class Filt, int S>
struct StepRunner {
  // (q0, r0) = (n0 / hop, n0 % hop); hop >= TILE so a tile crosses at most one hop boundary
  __device__ __forceinline__ static void run(Filt& f, const FfParams& p, const float* strip, float* acc, const float* wsm,
                                             int lane, int n0, int q0, int r0, bool frame_ok, float gframe) {
    const int n = n0 + S;
    if (n < p.win) {
      const bool wrap = r0 + S >= p.hop;
      const int q = q0 + (wrap ? 1 : 0), r = r0 + S - (wrap ? p.hop : 0);
      float x = strip[(lane + q) * (p.hop + 1) + r];
      if (!p.interp_gain) x = __fmul_rn(x, gframe);
      const float yv = f.template step<S>(x);
      const int sj = lane + q - (p.NQ - 1);
      if (frame_ok && sj >= 0 && sj < 33 - p.NQ) acc[sj * (p.hop + 1) + r] += wsm[n] * yv;
    }
    if constexpr (S + 1 < Filt::TILE) StepRunner<Filt, S + 1>::run(f, p, strip, acc, wsm, lane, n0, q0, r0, frame_ok, gframe);
  }
};

template <class Filt>
__global__ void __launch_bounds__(32) ff_frames_kernel(FfParams p) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x;
  const int b = blockIdx.x / p.warps_per_seq, w = blockIdx.x % p.warps_per_seq;
  const int NS = 33 - p.NQ;              // complete segments per warp
  const int NSTRIP = 32 + p.NQ - 1;      // segments of excitation the warp's frames touch
  const int seg_stride = p.hop + 1;
  float* strip = smem;                        // [NSTRIP][hop+1]  padded-coordinate excitation
  float* acc = strip + NSTRIP * seg_stride;   // [NS][hop+1]      overlap-add accumulators
  float* wsm = acc + NS * seg_stride;         // [win]            window
  const int P0 = p.nseg0 + w * NS;            // first padded segment owned by this warp
  const int k0 = P0 - (p.NQ - 1);             // frame handled by lane 0
  const int k = k0 + lane;
  const bool frame_ok = (k >= 0) && (k < p.n_frames);

  // ---- stage window, zero accumulators, build the excitation strip
  for (int i = lane; i < p.win; i += 32) wsm[i] = p.window[i];
  for (int i = lane; i < NS * seg_stride; i += 32) acc[i] = 0.f;
  const float* exb = p.ex + (size_t)b * p.ex_stride;
  const float* gb = p.gain + (size_t)b * p.F;
  for (int sg = 0; sg < NSTRIP; ++sg) {
    const int xbase = (k0 + sg) * p.hop - p.pad;  // signal position of the segment start
    for (int r = lane; r < p.hop; r += 32) {
      const int pos = xbase + r;
      float v = 0.f;
      if (pos >= 0 && pos < p.Le) {
        v = exb[pos];
        if (p.interp_gain) {
          const Lerp lw = lerp_at(pos, p.scale, p.F);
          v = __fmul_rn(v, lerp_apply(lw, gb[lw.i0], gb[lw.i1]));
        }
      }
      strip[sg * seg_stride + r] = v;
    }
  }
  Filt f;
  f.load(p, b, k, frame_ok);
  const float gframe = (!p.interp_gain && frame_ok) ? gb[k] : 1.f;
  __syncwarp();

  // ---- the serial part: `win` recurrence steps per lane
  int q0 = 0, r0 = 0;
#pragma unroll 1
  for (int n0 = 0; n0 < p.win; n0 += Filt::TILE) {
    StepRunner<Filt, 0>::run(f, p, strip, acc, wsm, lane, n0, q0, r0, frame_ok, gframe);
    r0 += Filt::TILE;
    if (r0 >= p.hop) r0 -= p.hop, ++q0;
    __syncwarp();
  }

  // ---- normalise by the overlap-added window and store
  float* yb = p.y + (size_t)b * p.out_len;
  for (int sj = 0; sj < NS; ++sj) {
    const int P = P0 + sj;  // padded segment
    for (int r = lane; r < p.hop; r += 32) {
      const int o = P * p.hop + r - p.pad;
      if (o < 0 || o >= p.out_len) continue;
      float norm = 0.f;
      for (int q = p.NQ - 1; q >= 0; --q) {  // frames P-q contribute their q-th hop of the window
        const int kk = P - q;
        if (kk >= 0 && kk < p.n_frames) norm += wsm[q * p.hop + r];
      }
      yb[o] = acc[sj * seg_stride + r] / norm;
    }
  }
}

// ---- inverse / analysis filter ------------------------------------------------------
__global__ void lpc_inverse_kernel(const float* __restrict__ y, int64_t y_stride, const float* __restrict__ a,
                                   float* __restrict__ r, int B, int L, int F, int M, float scale) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= L) return;
  const Lerp w = lerp_at(t, scale, F);
  const float* a0 = a + ((size_t)b * F + w.i0) * M;
  const float* a1 = a + ((size_t)b * F + w.i1) * M;
  const float* yb = y + (size_t)b * y_stride;
  // fir_filt (models/utils.py:433-441) sums y[t-M..t] * [a_M .. a_1, 1] oldest-first
  float acc = 0.f;
  for (int i = M - 1; i >= 0; --i) {
    const int ty = t - 1 - i;
    if (ty >= 0) acc = __fmaf_rn(lerp_apply(w, a0[i], a1[i]), yb[ty], acc);
  }
  r[(size_t)b * L + t] = acc + yb[t];
}

template <class Filt>
static int launch_ff(const FfParams& p, cudaStream_t st) {
  const int NS = 33 - p.NQ, NSTRIP = 32 + p.NQ - 1;
  const size_t sm = ((size_t)(NS + NSTRIP) * (p.hop + 1) + p.win) * sizeof(float);
  if (sm > 220 * 1024) return GOLF_ERR_UNSUPPORTED;
  static size_t sm_allowed = 48 * 1024;  // per instantiation
  if (sm > sm_allowed) {
    GOLF_CUDA(cudaFuncSetAttribute(ff_frames_kernel<Filt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sm_allowed = sm;
  }
  ff_frames_kernel<Filt><<<p.B * p.warps_per_seq, 32, sm, st>>>(p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

static int fill_geometry(FfParams* p, int T_ex, int pad) {
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
  p->warps_per_seq = ceil_div(p->nseg, 33 - p->NQ);
  p->scale = lerp_scale(p->F, p->hop);
  return GOLF_OK;
}

}  // namespace golf

using namespace golf;

GOLF_API int golf_lpc_ff_fwd(const float* ex, int64_t ex_stride, const float* gain, const float* a, const float* window,
                             float* y, int B, int T_ex, int F, int M, int hop, int win, void* stream) {
  if (!ex || !gain || !a || !window || !y || B <= 0 || T_ex <= 0 || F <= 0 || M <= 0 || hop <= 0 || win < 2 * hop)
    return GOLF_ERR_INVALID;
  if (M > 40 || hop < 40) return GOLF_ERR_UNSUPPORTED;
  FfParams p{};
  p.ex = ex, p.ex_stride = ex_stride, p.gain = gain, p.coef = a, p.window = window, p.y = y;
  p.B = B, p.F = F, p.M = M, p.hop = hop, p.win = win, p.interp_gain = 1;
  int rc = fill_geometry(&p, T_ex, win / 2);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 4) return launch_ff<AllPole<4>>(p, st);
  if (M <= 8) return launch_ff<AllPole<8>>(p, st);
  if (M <= 12) return launch_ff<AllPole<12>>(p, st);
  if (M <= 16) return launch_ff<AllPole<16>>(p, st);
  if (M <= 20) return launch_ff<AllPole<20>>(p, st);
  if (M <= 24) return launch_ff<AllPole<24>>(p, st);
  if (M <= 32) return launch_ff<AllPole<32>>(p, st);
  return launch_ff<AllPole<40>>(p, st);
}

GOLF_API int golf_biquad_ff_fwd(const float* ex, int64_t ex_stride, const float* gain, const float* biquads,
                                const float* window, float* y, int B, int T_ex, int F, int K, int hop, int win,
                                void* stream) {
  if (!ex || !gain || !biquads || !window || !y || B <= 0 || T_ex <= 0 || F <= 0 || K <= 0 || hop <= 0 || win < hop)
    return GOLF_ERR_INVALID;
  if (K > 16 || hop < 16 || (win - hop) % 2 != 0) return GOLF_ERR_UNSUPPORTED;
  FfParams p{};
  p.ex = ex, p.ex_stride = ex_stride, p.gain = gain, p.coef = biquads, p.window = window, p.y = y;
  p.B = B, p.F = F, p.M = K, p.hop = hop, p.win = win, p.interp_gain = 0;
  int rc = fill_geometry(&p, T_ex, (win - hop) / 2);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 4) return launch_ff<BiquadCascade<4>>(p, st);
  if (K <= 8) return launch_ff<BiquadCascade<8>>(p, st);
  if (K <= 12) return launch_ff<BiquadCascade<12>>(p, st);
  return launch_ff<BiquadCascade<16>>(p, st);
}

GOLF_API int golf_lpc_inverse_fwd(const float* y, int64_t y_stride, const float* a, float* r, int B, int L, int F, int M,
                                  int hop, void* stream) {
  if (!y || !a || !r || B <= 0 || L <= 0 || F <= 0 || M <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  if ((int64_t)L > (int64_t)(F - 1) * hop + 1 || y_stride < L) return GOLF_ERR_INVALID;
  dim3 grid(ceil_div(L, 256), B);
  lpc_inverse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, y_stride, a, r, B, L, F, M, lerp_scale(F, hop));
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
