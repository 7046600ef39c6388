// lpc_ff.cu -- GOLF-ff: frame-wise LTI all-pole filtering with Hann overlap-add (forward and
// adjoint), the cascaded-biquad variant, and the inverse (analysis) filter.
//
// Replaces models/filters.py:131-184 (LTVMinimumPhaseFilter.forward: ex*gain, pad,
// unfold(win, hop), models/lpc.py:11-16 lpc_synthesis -> torchaudio lfilter per frame,
// conv_transpose1d OLA with a dense diag(window) kernel, divide by the OLA'd window) and
// torchaudio's DifferentiableIIR.backward; models/lpc.py:94-131 (BatchSecondOrderLPCSynth);
// models/filters.py:186-195 (reverse) + models/utils.py:433-441 (fir_filt).
//
// Mapping: frames are independent, each a serial recurrence of `win` steps from zero state --
// one LANE per frame, coefficients and filter state in registers (static rotation, loop
// unrolled by the padded order MP).  Taps are visited oldest-first like libtorchaudio's CPU
// loop, in three interleaved chains with the newest tap last, so only one FMA per step sits on
// the serial dependency.  A CTA owns 32 consecutive frames of one utterance (warp 0 runs them)
// and the 33-NQ padded-coordinate segments (hop samples each, NQ = win/hop) those frames fully
// determine; all 4 warps of the CTA stage and write out:
//   * the excitation strip (ex*up(gain), computed once per sample) is staged in shared memory,
//     segment stride hop+1 so the per-lane strided reads are bank-conflict free;
//   * lane l at step n = q*hop + r adds window[n]*y to segment l+q-(NQ-1), offset r;
//   * write-out divides by the overlap-added window and stores coalesced.
// Adjacent CTAs recompute NQ-1 frames of overlap (10% at NQ=4) instead of exchanging partial
// sums: deterministic, no atomics, no 4x unfold in HBM.
//
// The adjoint has the same geometry (a frame's input and output positions coincide): the strip
// holds gy/norm, the lane multiplies by window[n], runs the same recurrence on reversed time
// (u), scatters u into the segment accumulators (-> d_e) and accumulates
// d_a[i] -= u[n] * v[n-1-i] against the frame's forward output v, which a forward pass with
// STORE_V left in a workspace ([B*n_frames, win], L2 resident).
//
// Algorithmic HBM bytes per output sample: 4 (ex) + 4 (y) + 4(M+1)/hop = 8.383 B.
#include "lpc_ff.cuh"

namespace golf {

#ifdef GOLF_FF_TIMING  // phase clocks of CTA 0 (variant builds only: tools/gpu/time_ff.py)
__device__ long long g_ff_clk[8];
#define FF_CLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_ff_clk[i] = clock64(); } while (0)
#else
#define FF_CLK(i) do { } while (0)
#endif

// ---- forward ------------------------------------------------------------------------
// ALIGNED: hop % TILE == 0 and win % TILE == 0 -> a tile never crosses a hop boundary, every
// shared-memory access of a tile is base + constant offset (no per-step index arithmetic).
template <class Filt, bool STORE_V, bool ALIGNED>
__global__ void __launch_bounds__(kFfThreads) ff_forward_kernel(FfParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.ctas_per_seq;
  const FfGeom g = ff_geom(p, blockIdx.x % p.ctas_per_seq);
  float* __restrict__ strip = smem;                       // [NSTRIP][hop+1] padded-coordinate excitation
  float* __restrict__ acc = strip + g.NSTRIP * g.seg_stride;  // [NS][hop+1]     overlap-add accumulators
  float* __restrict__ wsm = acc + g.NS * g.seg_stride;        // [win]           window
  float* __restrict__ vt = wsm + p.win;                       // [32][TILE+1]    v tile (STORE_V)
  const int k = g.k0 + lane;
  const bool frame_ok = (k >= 0) && (k < p.n_frames);
  FF_CLK(0);

  // ---- all warps: window, zeroed accumulators, excitation strip
  const float* __restrict__ exb = p.ex + (size_t)b * p.ex_stride;
  const float* __restrict__ gb = p.gain + (size_t)b * p.F;
  // The strip is latency bound (66 elements per thread at hop 240): with a load -> multiply -> store loop every element
  // paid its own L2 round trip (~20 us of an 80 us kernel).  All copies go out at once through cp.async (LDGSTS,
  // zero-filled outside the signal); the gain envelope is applied in a second pass over shared memory.
  constexpr int kWarps = kFfThreads / 32;
  for (int sg = warp; sg < g.NSTRIP; sg += kWarps) {
    const int pos0 = (g.k0 + sg) * p.hop - p.pad;  // signal position of the segment's first sample
    float* __restrict__ row = strip + sg * g.seg_stride;
    for (int r = lane; r < p.hop; r += 32) {
      const int pos = pos0 + r;
      const bool ok = pos >= 0 && pos < p.Le;
      cp_async4(row + r, exb + min(max(pos, 0), p.Le - 1), ok);
    }
  }
  for (int i = tid; i < p.win; i += kFfThreads) wsm[i] = __ldg(p.window + i);
  for (int i = tid; i < g.NS * g.seg_stride; i += kFfThreads) acc[i] = 0.f;
  cp_async_wait_all();
  FF_CLK(1);
  if (p.interp_gain) {  // each thread revisits the elements it copied itself: no barrier needed in between
    for (int sg = warp; sg < g.NSTRIP; sg += kWarps) {
      const int pos0 = (g.k0 + sg) * p.hop - p.pad;
      float* __restrict__ row = strip + sg * g.seg_stride;
      for (int r4 = lane; r4 < p.hop; r4 += 4 * 32) {  // four elements per round: their gain loads are issued together
        Lerp lw[4];
        float g0[4], g1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pos = min(max(pos0 + r4 + 32 * u, 0), p.Le - 1);
          lw[u] = lerp_at(pos, p.scale, p.F);
          g0[u] = __ldg(gb + lw[u].i0), g1[u] = __ldg(gb + lw[u].i1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r4 + 32 * u, pos = pos0 + r;
          if (r < p.hop && pos >= 0 && pos < p.Le) row[r] = __fmul_rn(row[r], lerp_apply(lw[u], g0[u], g1[u]));
        }
      }
    }
  }
  __syncthreads();
  FF_CLK(2);

  // ---- warp 0: the serial part, `win` recurrence steps per lane
  if (warp == 0) {
    Filt f;
    f.load(p, b, k, frame_ok);
    FF_CLK(3);
    const float gframe = (!p.interp_gain && frame_ok) ? __ldg(gb + k) : 1.f;
    int q0 = 0, r0 = 0;  // (n0 / hop, n0 % hop); hop >= TILE so a tile crosses at most one hop boundary
#pragma unroll 1
    for (int n0 = 0; n0 < p.win; n0 += Filt::TILE) {
      float xs[Filt::TILE], ys[Filt::TILE];
      if (ALIGNED) {
        const float* __restrict__ xrow = strip + (lane + q0) * g.seg_stride + r0;
#pragma unroll
        for (int s = 0; s < Filt::TILE; ++s) xs[s] = xrow[s];
      } else {
#pragma unroll
        for (int s = 0; s < Filt::TILE; ++s) {
          const bool wrap = r0 + s >= p.hop;
          const int q = q0 + (wrap ? 1 : 0), r = r0 + s - (wrap ? p.hop : 0);
          xs[s] = (n0 + s < p.win) ? strip[(lane + q) * g.seg_stride + r] : 0.f;
        }
      }
      if (!p.interp_gain) {
#pragma unroll
        for (int s = 0; s < Filt::TILE; ++s) xs[s] = __fmul_rn(xs[s], gframe);
      }
      TileSteps<Filt, 0, Filt::TILE>::run(f, xs, ys);  // static tile positions
      if (ALIGNED) {
        const int sj = lane + q0 - (p.NQ - 1);
        if (frame_ok && sj >= 0 && sj < g.NS) {
          float* __restrict__ arow = acc + sj * g.seg_stride + r0;
          const float* __restrict__ wrow = wsm + n0;
#pragma unroll
          for (int s = 0; s < Filt::TILE; ++s) arow[s] = __fmaf_rn(wrow[s], ys[s], arow[s]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < Filt::TILE; ++s) {
          const bool wrap = r0 + s >= p.hop;
          const int q = q0 + (wrap ? 1 : 0), r = r0 + s - (wrap ? p.hop : 0);
          const int sj = lane + q - (p.NQ - 1);
          if (n0 + s < p.win && frame_ok && sj >= 0 && sj < g.NS) acc[sj * g.seg_stride + r] += wsm[n0 + s] * ys[s];
        }
      }
      if (STORE_V) {
#pragma unroll
        for (int s = 0; s < Filt::TILE; ++s) vt[lane * (Filt::TILE + 1) + s] = ys[s];
      }
      if (STORE_V) {  // coalesced rows of the frame-output workspace
        __syncwarp();
        for (int i = lane; i < 32 * Filt::TILE; i += 32) {
          const int rr = i / Filt::TILE, s = i - rr * Filt::TILE;
          const int kk = g.k0 + rr;
          // frames owned by this CTA only (overlap frames are written by the neighbour that owns them)
          const bool own = kk >= 0 && kk < p.n_frames && rr >= p.NQ - 1 - (blockIdx.x % p.ctas_per_seq == 0 ? p.NQ - 1 : 0);
          if (own && n0 + s < p.win) p.vws[((size_t)b * p.n_frames + kk) * p.win + n0 + s] = vt[rr * (Filt::TILE + 1) + s];
        }
        __syncwarp();
      }
      r0 += Filt::TILE;
      if (r0 >= p.hop) r0 -= p.hop, ++q0;
    }
  }
  FF_CLK(4);
  __syncthreads();

  // ---- all warps: normalise by the overlap-added window and store
  // Interior segments (all NQ frames exist) share one normalisation row: it is summed once into the strip (free now)
  // in ff_norm's order; edge segments take the generic path.  (Per-element index division + the norm loop made this
  // phase 16 us of an 80 us kernel.)
  float* __restrict__ yb = p.y + (size_t)b * p.out_len;
  float* __restrict__ nrow = strip;
  for (int r = tid; r < p.hop; r += kFfThreads) {
    float norm = 0.f;
    for (int q = p.NQ - 1; q >= 0; --q) norm += wsm[q * p.hop + r];
    nrow[r] = norm;
  }
  __syncthreads();
  for (int sj = warp; sj < g.NS; sj += kWarps) {
    const int P = g.P0 + sj;
    const int o0 = P * p.hop - p.pad;
    const bool interior = P - (p.NQ - 1) >= 0 && P < p.n_frames;
    const float* __restrict__ arow = acc + sj * g.seg_stride;
    if (o0 >= p.out_len || o0 + p.hop <= 0) continue;
    if (interior && o0 >= 0 && o0 + p.hop <= p.out_len) {  // the common case: no per-element checks
#pragma unroll 4
      for (int r = lane; r < p.hop; r += 32) yb[o0 + r] = __fdiv_rn(arow[r], nrow[r]);
    } else {
#pragma unroll 1
      for (int r = lane; r < p.hop; r += 32) {
        const int o = o0 + r;
        if (o >= 0 && o < p.out_len) yb[o] = __fdiv_rn(arow[r], interior ? nrow[r] : ff_norm(p, wsm, P, r));
      }
    }
  }
  FF_CLK(5);
}

// ---- adjoint (all-pole only) ----------------------------------------------------------
// ex := gy [B,out_len]; y := d_e [B,Le]; vws = forward frame outputs; d_a [B,F,M].
// FRAME_GAIN (BatchLPCSynth, models/lpc.py:62-91): the gain is one value per frame instead of an interpolated
// envelope -- the excitation strip is staged too, d_gain[k] = sum_n u[n] ex[n] is accumulated by the frame's lane
// and the segment accumulators receive gain_k * u (= d_ex directly).
template <int MP, bool FRAME_GAIN>
__global__ void __launch_bounds__(kFfThreads) ff_backward_kernel(FfParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x % p.ctas_per_seq, b = blockIdx.x / p.ctas_per_seq;
  const FfGeom g = ff_geom(p, cta);
  float* __restrict__ strip = smem;                           // [NSTRIP][hop+1] gy / norm in padded coordinates
  float* __restrict__ acc = strip + g.NSTRIP * g.seg_stride;  // [NS][hop+1]     d_e accumulators
  float* __restrict__ wsm = acc + g.NS * g.seg_stride;        // [win]
  float* __restrict__ vt = wsm + p.win;                       // [2][32][2*MP+1] v tiles (double buffered): column c <-> v[nhi - 2*MP + c]
  float* __restrict__ xstrip = vt + 2 * 32 * (2 * MP + 1);    // [NSTRIP][hop+1] excitation in padded coordinates (FRAME_GAIN)
  const int k = g.k0 + lane;
  const bool frame_ok = (k >= 0) && (k < p.n_frames);
  // a frame's d_a is produced by the CTA that owns it (not by the neighbour that recomputes it)
  const bool own = frame_ok && (lane >= p.NQ - 1 || cta == 0);

  for (int i = tid; i < p.win; i += kFfThreads) wsm[i] = __ldg(p.window + i);
  for (int i = tid; i < g.NS * g.seg_stride; i += kFfThreads) acc[i] = 0.f;
  __syncthreads();
  ff_stage_adjoint_strips<FRAME_GAIN>(p, g, p.ex + (size_t)b * p.ex_stride, FRAME_GAIN ? p.vws_ex + (size_t)b * p.ex_stride2 : nullptr,
                                      strip, xstrip, xstrip + (FRAME_GAIN ? g.NSTRIP * g.seg_stride : 0), wsm);

  if (warp == 0) {
    AllPole<MP> f;
    f.load(p, b, k, frame_ok);
    const float gframe = (FRAME_GAIN && frame_ok) ? __ldg(p.gain + (size_t)b * p.F + k) : 0.f;
    float dg = 0.f;
    float da[MP];  // da[i] accumulates -sum_n u[n] v[n-1-i]
    float vh[MP];  // vh[m mod MP] = v[m] for the M most recent m < n (win % MP == 0 keeps slots static)
#pragma unroll
    for (int i = 0; i < MP; ++i) da[i] = 0.f, vh[i] = 0.f;
    constexpr int VT = 2 * MP + 1;
    // v tiles stream through two shared-memory buffers filled by cp.async (LDGSTS, zero-fill where a frame
    // or a sample does not exist): the copy of the NEXT tile runs while this one is consumed.  (Staging each
    // tile with load -> store pairs cost one L2 latency per 32 elements, 49 times per tile: 383 us.)
    auto stage = [&](float* dst, int nhi_) {
      for (int i = lane; i < 32 * VT; i += 32) {
        const int rr = i / VT, c = i - rr * VT;
        const int kk = g.k0 + rr, m = nhi_ - 2 * MP + c;
        const bool ok = kk >= 0 && kk < p.n_frames && m >= 0;
        cp_async4(dst + i, p.vws + ((size_t)b * p.n_frames + (ok ? kk : 0)) * p.win + (ok ? m : 0), ok);
      }
    };
    stage(vt, p.win - 1);
    // reversed time: tau = 0..win-1 <-> n = win-1-tau.  Tile tau0..tau0+MP-1 covers n in [nhi-MP+1, nhi].
    int q0 = p.NQ - 1, r0 = p.hop - 1;  // (n / hop, n % hop) at the tile's first step
    int buf = 0;
#pragma unroll 1
    for (int tau0 = 0; tau0 < p.win; tau0 += MP) {
      const int nhi = p.win - 1 - tau0;
      // ---- v[nhi-2*MP .. nhi] of all 32 frames: wait for this tile, start the next one
      cp_async_wait_all();
      __syncwarp();
      float* __restrict__ vcur = vt + buf * (32 * VT);
      if (tau0 + MP < p.win) stage(vt + (buf ^ 1) * (32 * VT), nhi - MP);
      buf ^= 1;
      if (tau0 == 0) {  // history for the first step: v[win-2 .. win-1-MP]
#pragma unroll
        for (int i = 0; i < MP; ++i) {
          const int c = 2 * MP - 1 - i;  // m = nhi - 1 - i
          vh[((MP - 2 - i) % MP + MP) % MP] = vcur[lane * VT + c];
        }
      }
      // hop % MP == 0 (checked on the host): the tile stays inside hop-segment q0, offsets r0-s
      float xs[MP], us[MP];
      {
        const float* __restrict__ xrow = strip + (lane + q0) * g.seg_stride + r0;
        const float* __restrict__ wrow = wsm + nhi;
#pragma unroll
        for (int s = 0; s < MP; ++s) xs[s] = __fmul_rn(xrow[-s], wrow[-s]);
      }
      TileSteps<AllPole<MP>, 0, MP>::run(f, xs, us);
      if (FRAME_GAIN) {
        const float* __restrict__ erow = xstrip + (lane + q0) * g.seg_stride + r0;
#pragma unroll
        for (int s = 0; s < MP; ++s) dg = __fmaf_rn(us[s], erow[-s], dg);
      }
      const float* __restrict__ vrow_t = vcur + lane * VT;
#pragma unroll
      for (int s = 0; s < MP; ++s) {
        // n = nhi - s; n mod MP = (MP-1-s) since win % MP == 0 and tau0 % MP == 0
        // v[n-1-i] sits in slot (n-1-i) mod MP = (2*MP-2-s-i) % MP
#pragma unroll
        for (int i = 0; i < MP; ++i) da[i] = __fmaf_rn(us[s], vh[(2 * MP - 2 - s - i) % MP], da[i]);
        // slide: v[n-1] leaves, v[n-1-MP] enters (same slot)
        vh[(2 * MP - 2 - s) % MP] = vrow_t[MP - 1 - s];  // column of m = n-1-MP = nhi-s-1-MP
      }
      {
        const int sj = lane + q0 - (p.NQ - 1);
        if (frame_ok && sj >= 0 && sj < g.NS) {
          float* __restrict__ arow = acc + sj * g.seg_stride + r0;
#pragma unroll
          for (int s = 0; s < MP; ++s) arow[-s] += FRAME_GAIN ? us[s] * gframe : us[s];
        }
      }
      r0 -= MP;
      if (r0 < 0) r0 += p.hop, --q0;
    }
    if (own && p.d_a) {
      float* dst = p.d_a + ((size_t)b * p.F + k) * p.M;
      for (int i = 0; i < p.M; ++i) dst[i] = -da[i];
    }
    if (FRAME_GAIN && own && p.d_gain) p.d_gain[(size_t)b * p.F + k] = dg;
  }
  __syncthreads();
  ff_write_adjoint_rows(p, g, acc, p.y + (size_t)b * p.Le);
}

// d_ex = d_e * up(gain); d_gain[b,k] = sum_t w_k(t) d_e[t] ex[t]; frames beyond n_frames get d_a = 0
__global__ void ff_finish_kernel(const float* __restrict__ d_e, const float* __restrict__ ex, int64_t ex_stride,
                                 const float* __restrict__ gain, float* __restrict__ d_ex, int64_t dex_stride,
                                 float* __restrict__ d_gain, float* __restrict__ d_a, int B, int T_ex, int Le, int F, int M,
                                 int hop, int n_frames, int skip, float scale) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (d_ex && t < T_ex + skip) {
    float v = 0.f;
    const int te = t - skip;
    if (te >= 0 && te < Le) {
      const Lerp w = lerp_at(te, scale, F);
      v = __fmul_rn(d_e[(size_t)b * Le + te], lerp_apply(w, gain[(size_t)b * F + w.i0], gain[(size_t)b * F + w.i1]));
    }
    d_ex[(size_t)b * dex_stride + t] = v;
  }
  if (t < F) {
    if (d_gain) {
      float acc = 0.f;
      const int t_lo = max(0, (t - 1) * hop), t_hi = min(Le - 1, (t + 1) * hop);
      for (int tt = t_lo; tt <= t_hi; ++tt) {
        const Lerp w = lerp_at(tt, scale, F);
        float wk = 0.f;
        if (w.i0 == t) wk += w.l0;
        if (w.i1 == t) wk += w.l1;
        if (wk != 0.f) acc = __fmaf_rn(wk, d_e[(size_t)b * Le + tt] * ex[(size_t)b * ex_stride + tt], acc);
      }
      d_gain[(size_t)b * F + t] = acc;
    }
    if (d_a && t >= n_frames)
      for (int i = 0; i < M; ++i) d_a[((size_t)b * F + t) * M + i] = 0.f;
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

// adjoint w.r.t. the filtered signal: d_y[s] = g[s] + sum_i c(s+1+i, i) g[s+1+i], c(t, .) the coefficient row
// the forward interpolated at time t (same arithmetic)
__global__ void lpc_inverse_dy_kernel(const float* __restrict__ g, const float* __restrict__ a, float* __restrict__ d_y, int B,
                                      int L, int F, int M, float scale) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (s >= L) return;
  const float* gb = g + (size_t)b * L;
  const float* ab = a + (size_t)b * F * M;
  float acc = gb[s];
  for (int i = 0; i < M; ++i) {
    const int t = s + 1 + i;
    if (t >= L) break;
    const Lerp w = lerp_at(t, scale, F);
    acc = __fmaf_rn(lerp_apply(w, ab[(size_t)w.i0 * M + i], ab[(size_t)w.i1 * M + i]), gb[t], acc);
  }
  d_y[(size_t)b * L + s] = acc;
}

// ---- host side ------------------------------------------------------------------------

template <class Filt, bool STORE_V, bool ALIGNED>
static int launch_ff_fwd_a(const FfParams& p, cudaStream_t st) {
  const size_t sm = ff_smem_bytes(p, STORE_V ? 32 * (Filt::TILE + 1) : 0);
  if (sm > 220 * 1024) return GOLF_ERR_UNSUPPORTED;
  static size_t sm_allowed_dev[64];  // per instantiation and device (0: the 48 KB default)
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t& sm_allowed = sm_allowed_dev[dev_ & 63];
  if (sm > 48 * 1024 && sm > sm_allowed) {
    GOLF_CUDA(cudaFuncSetAttribute(ff_forward_kernel<Filt, STORE_V, ALIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sm_allowed = sm;
  }
  ff_forward_kernel<Filt, STORE_V, ALIGNED><<<p.B * p.ctas_per_seq, kFfThreads, sm, st>>>(p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
template <class Filt, bool STORE_V>
static int launch_ff_fwd(const FfParams& p, cudaStream_t st) {
  if (p.hop % Filt::TILE == 0 && p.win % Filt::TILE == 0) return launch_ff_fwd_a<Filt, STORE_V, true>(p, st);
  return launch_ff_fwd_a<Filt, STORE_V, false>(p, st);
}

template <int MP, bool FRAME_GAIN>
static int launch_ff_bwd(const FfParams& pf, const FfParams& pb, cudaStream_t st) {
  int rc = launch_ff_fwd<AllPole<MP>, true>(pf, st);
  if (rc) return rc;
  const int NSTRIP = 32 + pb.NQ - 1;
  const size_t sm = ff_smem_bytes(pb, 2 * 32 * (2 * MP + 1) + (FRAME_GAIN ? NSTRIP * (pb.hop + 1) : 0) + pb.hop);
  if (sm > 220 * 1024) return GOLF_ERR_UNSUPPORTED;
  static size_t sm_allowed_dev[64];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t& sm_allowed = sm_allowed_dev[dev_ & 63];
  if (sm > 48 * 1024 && sm > sm_allowed) {
    GOLF_CUDA(cudaFuncSetAttribute(ff_backward_kernel<MP, FRAME_GAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sm_allowed = sm;
  }
  ff_backward_kernel<MP, FRAME_GAIN><<<pb.B * pb.ctas_per_seq, kFfThreads, sm, st>>>(pb);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

// The adjoint keeps static ring slots / tile offsets, so its padded order must divide the window and the hop: the smallest
// compiled order >= M that does (the extra taps are zero coefficients), e.g. window 1024 / hop 256 / order 22 -> 32.
static inline int ff_adjoint_order(int M, int win, int hop) {
  for (int mp : {4, 8, 12, 16, 20, 24, 32, 40})
    if (mp >= M && win % mp == 0 && hop % mp == 0) return mp;
  return 0;
}

template <bool FRAME_GAIN>
static int dispatch_ff_bwd(int mp, const FfParams& pf, const FfParams& pb, cudaStream_t st) {
  switch (mp) {
    case 4: return launch_ff_bwd<4, FRAME_GAIN>(pf, pb, st);
    case 8: return launch_ff_bwd<8, FRAME_GAIN>(pf, pb, st);
    case 12: return launch_ff_bwd<12, FRAME_GAIN>(pf, pb, st);
    case 16: return launch_ff_bwd<16, FRAME_GAIN>(pf, pb, st);
    case 20: return launch_ff_bwd<20, FRAME_GAIN>(pf, pb, st);
    case 24: return launch_ff_bwd<24, FRAME_GAIN>(pf, pb, st);
    case 32: return launch_ff_bwd<32, FRAME_GAIN>(pf, pb, st);
    default: return launch_ff_bwd<40, FRAME_GAIN>(pf, pb, st);
  }
}

std::atomic<int> g_ff_exact{0};  // 1: forward filters reproduce libtorchaudio's CPU summation order (golf_lpc_ff_set_exact_order)

template <bool STORE_V>
static int dispatch_ff_fwd(int M, const FfParams& p, cudaStream_t st) {
  if constexpr (!STORE_V) {
    if (g_ff_exact.load()) {
      if (M <= 4) return launch_ff_fwd<AllPole<4, true>, false>(p, st);
      if (M <= 8) return launch_ff_fwd<AllPole<8, true>, false>(p, st);
      if (M <= 12) return launch_ff_fwd<AllPole<12, true>, false>(p, st);
      if (M <= 16) return launch_ff_fwd<AllPole<16, true>, false>(p, st);
      if (M <= 20) return launch_ff_fwd<AllPole<20, true>, false>(p, st);
      if (M <= 24) return launch_ff_fwd<AllPole<24, true>, false>(p, st);
      if (M <= 32) return launch_ff_fwd<AllPole<32, true>, false>(p, st);
      return launch_ff_fwd<AllPole<40, true>, false>(p, st);
    }
  }
  if (M <= 4) return launch_ff_fwd<AllPole<4>, STORE_V>(p, st);
  if (M <= 8) return launch_ff_fwd<AllPole<8>, STORE_V>(p, st);
  if (M <= 12) return launch_ff_fwd<AllPole<12>, STORE_V>(p, st);
  if (M <= 16) return launch_ff_fwd<AllPole<16>, STORE_V>(p, st);
  if (M <= 20) return launch_ff_fwd<AllPole<20>, STORE_V>(p, st);
  if (M <= 24) return launch_ff_fwd<AllPole<24>, STORE_V>(p, st);
  if (M <= 32) return launch_ff_fwd<AllPole<32>, STORE_V>(p, st);
  return launch_ff_fwd<AllPole<40>, STORE_V>(p, st);
}

}  // namespace golf

using namespace golf;

#ifdef GOLF_FF_TIMING
GOLF_API int golf_debug_ff_clocks(long long* out8) {
  GOLF_CUDA(cudaDeviceSynchronize());
  GOLF_CUDA(cudaMemcpyFromSymbol(out8, g_ff_clk, sizeof(long long) * 8));
  return GOLF_OK;
}
#endif

GOLF_API void golf_lpc_ff_set_exact_order(int on) { g_ff_exact = on ? 1 : 0; }

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
  return dispatch_ff_fwd<false>(M, p, st);
}

GOLF_API size_t golf_lpc_ff_bwd_workspace_bytes(int B, int T_ex, int F, int hop, int win) {
  if (B <= 0 || T_ex <= 0 || F <= 0 || hop <= 0 || win < 2 * hop) return 0;
  FfParams p{};
  p.B = B, p.F = F, p.hop = hop, p.win = win, p.interp_gain = 1;
  if (fill_geometry(&p, T_ex, win / 2)) return 0;
  return align_up((size_t)B * p.n_frames * win * 4, 256) + align_up((size_t)B * p.Le * 4, 256) + align_up((size_t)B * p.out_len * 4, 256);
}

GOLF_API int golf_lpc_ff_bwd(const float* gy, const float* ex, int64_t ex_stride, const float* gain, const float* a,
                             const float* window, float* d_ex, int64_t dex_stride, float* d_gain, float* d_a, int B, int T_ex,
                             int F, int M, int hop, int win, void* workspace, size_t workspace_bytes, void* stream) {
  if (!gy || !ex || !gain || !a || !window || B <= 0 || T_ex <= 0 || F <= 0 || M <= 0 || hop <= 0 || win < 2 * hop)
    return GOLF_ERR_INVALID;
  if (M > 40 || hop < 40) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  FfParams pf{};
  pf.ex = ex, pf.ex_stride = ex_stride, pf.gain = gain, pf.coef = a, pf.window = window;
  pf.B = B, pf.F = F, pf.M = M, pf.hop = hop, pf.win = win, pf.interp_gain = 1;
  int rc = fill_geometry(&pf, T_ex, win / 2);
  if (rc) return rc;
  const size_t need = golf_lpc_ff_bwd_workspace_bytes(B, T_ex, F, hop, win);
  if (!workspace || workspace_bytes < need) return GOLF_ERR_WORKSPACE;
  char* ws = reinterpret_cast<char*>(workspace);
  float* vws = reinterpret_cast<float*>(ws);
  float* d_e = reinterpret_cast<float*>(ws + align_up((size_t)B * pf.n_frames * win * 4, 256));
  float* yscratch = reinterpret_cast<float*>(ws + align_up((size_t)B * pf.n_frames * win * 4, 256) + align_up((size_t)B * pf.Le * 4, 256));
  pf.vws = vws, pf.y = yscratch;
  FfParams pb = pf;
  pb.ex = gy, pb.ex_stride = pf.out_len, pb.y = d_e, pb.d_a = d_a;
  // the adjoint writes d_e over every input position [0, Le), which can reach past the last output
  pb.nseg = (pf.pad + pf.Le - 1) / hop - pf.nseg0 + 1;
  pb.ctas_per_seq = ceil_div(pb.nseg, 33 - pf.NQ);
  const int mp = ff_adjoint_order(M, win, hop);
  if (mp == 0) return GOLF_ERR_UNSUPPORTED;
  rc = dispatch_ff_bwd<false>(mp, pf, pb, st);
  if (rc) return rc;
  // d_ex covers the caller's full excitation row (zeros beyond the filtered span)
  const int span = T_ex > F ? T_ex : F;
  // the gain gradient is a per-frame reduction over 2*hop samples: one warp per (utterance, frame) instead of
  // one THREAD per frame walking its 481 samples (131 us for the whole launch)
  const bool fast_gain = d_gain && launch_gain_reduction(d_e, ex, ex_stride, d_gain, B, pf.Le, F, hop, st) == GOLF_OK;
  ff_finish_kernel<<<dim3(ceil_div(span, 256), B), 256, 0, st>>>(d_e, ex, ex_stride, gain, d_ex, dex_stride, fast_gain ? nullptr : d_gain,
                                                                d_a, B, T_ex, pf.Le, F, M, hop, pf.n_frames, 0, pf.scale);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

// ---- BatchLPCSynth / LPCSynth (models/lpc.py:19-91): one gain per frame, frames padded by (win - hop) / 2 ----
static int frames_params(FfParams* p, const float* ex, int64_t ex_stride, const float* gain, const float* a, const float* window,
                         int B, int T_ex, int F, int M, int hop, int win) {
  if (!ex || !gain || !a || !window || B <= 0 || T_ex <= 0 || F <= 0 || M <= 0 || hop <= 0 || win < 2 * hop) return GOLF_ERR_INVALID;
  if (M > 40 || hop < 40 || (win - hop) % 2 != 0) return GOLF_ERR_UNSUPPORTED;
  p->ex = ex, p->ex_stride = ex_stride, p->gain = gain, p->coef = a, p->window = window;
  p->B = B, p->F = F, p->M = M, p->hop = hop, p->win = win, p->interp_gain = 0;
  return fill_geometry(p, T_ex, (win - hop) / 2);
}

GOLF_API int golf_lpc_frames_out_length(int T_ex, int F, int hop, int win) {
  FfParams p{};
  p.B = 1, p.F = F, p.hop = hop, p.win = win, p.interp_gain = 0;
  if (T_ex <= 0 || F <= 0 || hop <= 0 || win < hop || (win - hop) % 2 != 0) return 0;
  if (fill_geometry(&p, T_ex, (win - hop) / 2)) return 0;
  return p.out_len;
}

GOLF_API int golf_lpc_frames_fwd(const float* ex, int64_t ex_stride, const float* gain, const float* a, const float* window,
                                 float* y, int B, int T_ex, int F, int M, int hop, int win, void* stream) {
  if (!y) return GOLF_ERR_INVALID;
  FfParams p{};
  int rc = frames_params(&p, ex, ex_stride, gain, a, window, B, T_ex, F, M, hop, win);
  if (rc) return rc;
  p.y = y;
  return dispatch_ff_fwd<false>(M, p, (cudaStream_t)stream);
}

GOLF_API size_t golf_lpc_frames_bwd_workspace_bytes(int B, int T_ex, int F, int hop, int win) {
  FfParams p{};
  p.B = B, p.F = F, p.hop = hop, p.win = win, p.interp_gain = 0;
  if (B <= 0 || T_ex <= 0 || F <= 0 || hop <= 0 || win < 2 * hop || (win - hop) % 2 != 0) return 0;
  if (fill_geometry(&p, T_ex, (win - hop) / 2)) return 0;
  return align_up((size_t)B * p.n_frames * win * 4, 256) + align_up((size_t)B * p.out_len * 4, 256);
}

GOLF_API int golf_lpc_frames_bwd(const float* gy, const float* ex, int64_t ex_stride, const float* gain, const float* a,
                                 const float* window, float* d_ex, float* d_gain, float* d_a, int B, int T_ex, int F, int M,
                                 int hop, int win, void* workspace, size_t workspace_bytes, void* stream) {
  if (!gy || !d_ex) return GOLF_ERR_INVALID;
  FfParams pf{};
  int rc = frames_params(&pf, ex, ex_stride, gain, a, window, B, T_ex, F, M, hop, win);
  if (rc) return rc;
  const size_t need = golf_lpc_frames_bwd_workspace_bytes(B, T_ex, F, hop, win);
  if (!workspace || workspace_bytes < need) return GOLF_ERR_WORKSPACE;
  const int mp = ff_adjoint_order(M, win, hop);
  if (mp == 0) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  pf.vws = reinterpret_cast<float*>(ws);
  pf.y = reinterpret_cast<float*>(ws + align_up((size_t)B * pf.n_frames * win * 4, 256));
  // frames beyond n_frames receive no gradient
  if (d_gain) GOLF_CUDA(cudaMemsetAsync(d_gain, 0, (size_t)B * F * sizeof(float), st));
  if (d_a) GOLF_CUDA(cudaMemsetAsync(d_a, 0, (size_t)B * F * M * sizeof(float), st));
  FfParams pb = pf;
  pb.ex = gy, pb.ex_stride = pf.out_len, pb.y = d_ex, pb.d_a = d_a, pb.d_gain = d_gain;
  pb.vws_ex = ex, pb.ex_stride2 = ex_stride;
  pb.nseg = (pf.pad + pf.Le - 1) / hop - pf.nseg0 + 1;  // d_ex over every input position [0, T_ex)
  pb.ctas_per_seq = ceil_div(pb.nseg, 33 - pf.NQ);
  return dispatch_ff_bwd<true>(mp, pf, pb, st);
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
  if (K <= 4) return launch_ff_fwd<BiquadCascade<4>, false>(p, st);
  if (K <= 8) return launch_ff_fwd<BiquadCascade<8>, false>(p, st);
  if (K <= 12) return launch_ff_fwd<BiquadCascade<12>, false>(p, st);
  return launch_ff_fwd<BiquadCascade<16>, false>(p, st);
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

GOLF_API int golf_lpc_inverse_bwd(const float* g, const float* y, int64_t y_stride, const float* a, float* d_y, float* d_a,
                                  int B, int L, int F, int M, int hop, void* stream) {
  if (!g || !y || !a || B <= 0 || L <= 0 || F <= 0 || M <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  if ((int64_t)L > (int64_t)(F - 1) * hop + 1 || y_stride < L || B > 65535) return GOLF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_y) {
    lpc_inverse_dy_kernel<<<dim3(ceil_div(L, 256), B), 256, 0, st>>>(g, a, d_y, B, L, F, M, lerp_scale(F, hop));
    GOLF_CHECK_LAUNCH();
  }
  if (d_a) return launch_frame_reduction(g, y, y_stride, d_a, B, L, F, M, hop, st);
  return GOLF_OK;
}
