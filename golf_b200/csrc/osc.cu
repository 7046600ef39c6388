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
//   osc_knot_prefix_kernel  the phase increments are piecewise linear between the Np
//                           control knots, so the running sum has a closed form inside a
//                           knot interval; only the per-knot prefix needs a scan.  Done in
//                           float64, one CTA per utterance (frac part kept).
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
  const float* ph = phase + (size_t)b * Np;
  double* pb = pref + (size_t)b * Np;
  const int per = (Np + 255) / 256;
  const int k0 = tid * per, k1 = min(Np, k0 + per);
  const double inv_os = 1.0 / (double)os;
  double s = 0.0;
  for (int k = k0; k < k1; ++k) {
    const double xk = (double)ph[k] * inv_os;
    const double xn = (double)ph[min(k + 1, Np - 1)] * inv_os;
    s += (double)hp * xk + (xn - xk) * 0.5 * (double)(hp - 1);
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
  for (int k = k0; k < k1; ++k) {
    pb[k] = run;
    const double xk = (double)ph[k] * inv_os;
    const double xn = (double)ph[min(k + 1, Np - 1)] * inv_os;
    run += (double)hp * xk + (xn - xk) * 0.5 * (double)(hp - 1);
    if (wrap) run -= floor(run);
  }
}

// upsampled increment at oversampled time t, ATen arithmetic on phase/os
__device__ __forceinline__ float osc_inc(const float* ph, int t, float scale, int Np, float os_f) {
  const Lerp w = lerp_at(t, scale, Np);
  return lerp_apply(w, __fdiv_rn(ph[w.i0], os_f), __fdiv_rn(ph[w.i1], os_f));
}

// bilinear table read, F.grid_sample(align_corners=True, zeros padding) arithmetic
__device__ __forceinline__ float osc_read(const float* __restrict__ tb, int R, int P, float wrapped, int t,
                                          float ydenom, int blocks) {
  const float gx = __fsub_rn(__fmul_rn(wrapped, 2.f), 1.f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn((float)t, ydenom), 2.f), 1.f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)P);
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)blocks);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = __fsub_rn(ix, x0f), fy = __fsub_rn(iy, y0f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  // rows beyond the R supplied ones replicate the last (F.pad replicate, synth.py:138-148)
  auto tap = [&](int yy, int xx) -> float {
    if (xx < 0 || xx > P || yy < 0 || yy > blocks) return 0.f;
    const int row = min(yy, R - 1);
    const int col = xx == P ? 0 : xx;  // column P wraps to column 0
    return tb[(size_t)row * P + col];
  };
  const float gx1 = __fsub_rn(1.f, fx), gy1 = __fsub_rn(1.f, fy);
  float v = __fmul_rn(tap(y0, x0), __fmul_rn(gx1, gy1));
  v = __fmaf_rn(tap(y0, x0 + 1), __fmul_rn(fx, gy1), v);
  v = __fmaf_rn(tap(y0 + 1, x0), __fmul_rn(gx1, fy), v);
  v = __fmaf_rn(tap(y0 + 1, x0 + 1), __fmul_rn(fx, fy), v);
  return v;
}

__global__ void wavetable_read_kernel(const float* __restrict__ wrapped, const float* __restrict__ tables,
                                      float* __restrict__ out, int N, int R, int P, int hop_tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= N) return;
  const int blocks = (N + hop_tab - 1) / hop_tab;
  out[(size_t)b * N + t] =
      osc_read(tables + (size_t)b * R * P, R, P, wrapped[(size_t)b * N + t], t, (float)((int64_t)hop_tab * blocks), blocks);
}

// ---- fused flow + decimation --------------------------------------------------------
constexpr int kOscTile = 1024;  // outputs per CTA (128 threads x 8)
constexpr int kMaxOs = 8;

struct OscParams {
  const float* phase;     // [B,Np]
  const float* tables;    // [B,Fw,P]
  const double* pref;     // [B,Np] exclusive knot prefix (frac part, or unwrapped when aten_cpu)
  int aten_cpu;           // 1: round the running sum to float32 before mod 1 (ATen CPU cumsum semantics)
  const float* dec;       // [2*zeros*os+1]
  float* out;             // [B,n_out]
  int B, Np, hp, N, n_out, Fw, P, hop_tab, blocks, os, zeros, equal_energy;
  int plen;               // per-phase strip length (floats, multiple of 4)
  int kp12;               // per-phase taps padded to a multiple of 12
  float scale, ydenom;
};

__global__ void __launch_bounds__(128) osc_flow_decimate_kernel(OscParams p) {
  extern __shared__ __align__(16) float smem[];
  float* vp = smem;                      // [os][plen] polyphase oversampled flow
  float* hp_ = smem + p.os * p.plen;     // [os][kp12] polyphase taps
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
  // oversampled samples u in [(m0-Z)*os, (m0-Z)*os + os*plen): vp[phs][j] = v[(m0-Z+j)*os + phs]
  const int total = p.os * p.plen;
  for (int i = tid; i < total; i += blockDim.x) {
    const int j = i / p.os, phs = i % p.os;  // consecutive threads -> consecutive oversampled times
    const int t = (m0 - Z + j) * p.os + phs;
    float v = 0.f;
    if (t >= 0 && t < p.N) {
      const float inc = osc_inc(ph, t, p.scale, p.Np, os_f);
      float wr;
      {
        const int k = min(t / p.hp, p.Np - 1), r = t - k * p.hp;
        const double xk = (double)ph[k] / (double)p.os;
        const double xn = (double)ph[min(k + 1, p.Np - 1)] / (double)p.os;
        double phi = p.pref[(size_t)b * p.Np + k] + (double)(r + 1) * xk +
                     (xn - xk) * ((double)r * (double)(r + 1)) / (2.0 * (double)p.hp);
        if (p.aten_cpu) {  // cumsum output is float32, then `% 1` in float32
          const float f = (float)phi;
          wr = __fsub_rn(f, floorf(f));
        } else {
          phi -= floor(phi);
          wr = (float)phi;
          if (wr >= 1.f) wr = 0.f;
        }
      }
      v = osc_read(tb, p.Fw, p.P, wr, t, p.ydenom, p.blocks);
      if (p.equal_energy) v = __fmul_rn(v, __fdiv_rn(1.f, __fsqrt_rn(inc)));
    }
    vp[phs * p.plen + j] = v;
  }
  __syncthreads();
  const int r0 = tid * kR;  // outputs m0 + r0 .. m0 + r0 + 7
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  for (int phs = 0; phs < p.os; ++phs) fir_tile8(vp + phs * p.plen + r0, hp_ + phs * p.kp12, p.kp12, acc);
  float* ob = p.out + (size_t)b * p.n_out;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (m0 + r0 + i < p.n_out) ob[m0 + r0 + i] = acc[i];
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
  L->bytes = L->off_pref + align_up((size_t)B * Np * 8, 256);
  return true;
}

}  // namespace golf

using namespace golf;

GOLF_API size_t golf_glottal_osc_workspace_bytes(int B, int Np, int phase_hop, int Fw, int P, int os) {
  OscLayout L;
  return osc_layout(B, Np, phase_hop, Fw, P, os, &L) ? L.bytes : 0;
}

GOLF_API int golf_glottal_osc_fwd(const float* phase, const float* w, const float* table, const float* dec_kernel,
                                  float* out, int B, int Np, int phase_hop, int Fw, int w_hop, int n_tab, int P, int os,
                                  int zeros, int accumulate, int flags, void* workspace, size_t workspace_bytes,
                                  void* stream) {
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

  osc_tables_kernel<<<ceil_div(B * Fw * P, 256), 256, 0, st>>>(w, table, tables, B * Fw, n_tab, P);
  GOLF_CHECK_LAUNCH();
  if (accumulate != 0 && accumulate != 1) return GOLF_ERR_INVALID;
  osc_knot_prefix_kernel<<<B, 256, 0, st>>>(phase, pref, Np, L.hp, os, accumulate == 0 ? 1 : 0);
  GOLF_CHECK_LAUNCH();
  OscParams p{};
  p.phase = phase, p.tables = tables, p.pref = pref, p.aten_cpu = accumulate;
  p.dec = os > 1 ? dec_kernel : nullptr, p.out = out;
  p.B = B, p.Np = Np, p.hp = L.hp, p.N = L.N, p.n_out = L.n_out, p.Fw = Fw, p.P = P;
  p.hop_tab = w_hop * os;
  p.blocks = (L.N + p.hop_tab - 1) / p.hop_tab;
  p.os = os, p.zeros = os > 1 ? zeros : 0, p.equal_energy = flags & 1;
  p.kp12 = ceil_div(2 * p.zeros + 1, 12) * 12;
  p.plen = (int)align_up((size_t)kOscTile + p.kp12 + 24, 4);
  p.scale = lerp_scale(Np, L.hp);
  p.ydenom = (float)((int64_t)p.hop_tab * p.blocks);
  const size_t sm = (size_t)os * (p.plen + p.kp12) * sizeof(float);
  if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(L.n_out, kOscTile), B);
  osc_flow_decimate_kernel<<<grid, 128, sm, st>>>(p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
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
