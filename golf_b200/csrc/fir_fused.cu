// fir_fused.cu -- noise branch of the GOLF decoder in ONE kernel: (optional) white-noise generation, zero-phase FIR
// design from log-magnitudes, block FIR, + harmonic source.
//
// Replaces, for the shipped configuration (n_mag = 256 -> 510 taps, cfg/ae/decoder/golf.yaml:20-25):
//   models/noise.py:34-35      torch.randn_like                      (opt-in: counter-based generator in the kernel)
//   models/filters.py:294-306  exp -> irfft -> fftshift -> window    (cuFFT + two helper kernels + a [B,F,510] tensor)
//   models/filters.py:360-384  block-wise FIR
//   models/sf.py:53-56         harm + filtered noise
//
// FIR design.  The spectrum X_k = exp(log_mag_k) is real, so its inverse real FFT is the cosine series
//   raw[n] = (1/K) [ X_0 + (-1)^n X_{N-1} + 2 sum_{0<k<N-1} X_k cos(2 pi k n / K) ],   K = 2(N-1) = 510,
// symmetric (raw[n] = raw[K-n]); fftshift and the (symmetric) window then give tap[i] = raw[|i - K/2|] * w[i].  Only
// raw[0..255] is needed, and cos(2 pi k (K/2 - d) / K) = (-1)^k cos(2 pi k d / K) pairs d with 255 - d:
//   raw[d] = E[d] + O[d],  raw[255-d] = E[d] - O[d],   E / O = the even-k / odd-k halves of the series, d < 128.
// That is two [frames x 128] x [128 x 128] matrix products per group of frames: a CTA designs the 8 frames it is
// about to apply, thread d keeps 8 + 8 accumulators, the spectra sit in shared memory ([k][frame], broadcast reads) and
// the cosine matrix entries are formed in registers by angle addition from two small tables (see g_design_aa below; the
// first version streamed the whole 128 KB matrix through L1 / L2 once per CTA).  32.8 k FMA per frame, a quarter of the
// 122 k FMA the FIR itself costs -- against 13 MB written and read back plus three launches for the cuFFT route.
//
// Noise.  With ex == NULL the strip is filled by Philox4x32-10 (counter = sample index / 4, utterance, call offset;
// key = seed) + Box-Muller: same distribution as torch.randn, not the same stream -- opt-in (the exact-stream mode
// passes the torch.randn tensor as `ex`).  Neighbouring CTAs regenerate the same halo samples from the same counters.
//
// FIR.  The packed-FP32 register tile of fir_tile.cuh (16 outputs per lane, FFMA2; fir_tile16_eo: even- and odd-tap partial sums
// per output, taps stored once).
#include <atomic>

#include "fir_tile.cuh"

namespace golf {

constexpr int kDN = 256;               // n_mag
constexpr int kDK = 2 * (kDN - 1);     // 510 taps
constexpr int kDH = kDK / 2;           // 255
constexpr int kDD = kDN / 2;           // 128 paired outputs d
constexpr int kDFB = 8;                // frames (blocks) designed and applied per CTA
constexpr int kDK20 = (kDK + kTapStep - 1) / kTapStep * kTapStep;  // 520

// Cosine matrix of the design, never stored whole (its 128 KB would stream through L1 / L2 once per CTA: 102 MB of L2 reads
// per launch, the kernel's largest stall).  With k' = 16 kh + kl the entry for even k = 2k' is cos(A_h + A_l), A_h = 2 pi 16 kh d
// / 255, A_l = 2 pi kl d / 255, and for odd k = 2k' + 1 it is cos(A_h + A_l + pi d / 255): a thread (fixed d) keeps cos / sin of
// the 16 A_l, of the 16 (A_l + pi d / 255) -- scaled by 2/K -- in 64 registers and fetches cos / sin of A_h once per 16 terms.
// [kind][index][d], kinds: 0 cos A_h, 1 sin A_h (index kh < 8), 2 / 3 cos / sin A_l, 4 / 5 cos / sin (A_l + pi d / 255); 48 KB,
// built once per device in double precision.  (c_0 = c_{N-1} = 1/K instead of 2/K: the spectrum values of those two bins are
// halved when they are staged.)
__device__ float g_design_aa[6][16][kDD];
static std::atomic<unsigned long long> g_design_ready{0};  // bit per device ordinal

__global__ void design_table_kernel() {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 6 * 16 * kDD) return;
  const int kind = idx / (16 * kDD), i = (idx / kDD) % 16, d = idx % kDD;
  double v;
  if (kind < 2) {
    const int m = (16 * i * d) % kDH;  // A_h = 2 pi m / 255
    v = kind == 0 ? cospi(2.0 * m / (double)kDH) : sinpi(2.0 * m / (double)kDH);
    if (i >= 8) v = 0.0;
  } else if (kind < 4) {
    const int m = (i * d) % kDH;
    v = (2.0 / kDK) * (kind == 2 ? cospi(2.0 * m / (double)kDH) : sinpi(2.0 * m / (double)kDH));
  } else {
    const int m = ((2 * i + 1) * d) % kDK;  // A_l + pi d / 255 = pi m / 255
    v = (2.0 / kDK) * (kind == 4 ? cospi(m / (double)kDH) : sinpi(m / (double)kDH));
  }
  g_design_aa[kind][i][d] = (float)v;
}

static int ensure_design_table(cudaStream_t st) {
  int dev = 0;
  GOLF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev > 63) return GOLF_ERR_UNSUPPORTED;
  if ((g_design_ready.load(std::memory_order_acquire) >> dev) & 1ull) return GOLF_OK;
  // first use on this device: built on the caller's stream, ahead of the kernel that reads it (if that first use is
  // being captured into a CUDA graph the build is captured too and simply repeats on every replay)
  design_table_kernel<<<ceil_div(6 * 16 * kDD, 256), 256, 0, st>>>();
  GOLF_CHECK_LAUNCH();
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
    // other streams of this process must not race ahead of the build
    GOLF_CUDA(cudaStreamSynchronize(st));
    g_design_ready.fetch_or(1ull << dev, std::memory_order_release);
  }
  return GOLF_OK;
}

// ---- Philox4x32-10 + Box-Muller ------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// four standard normals for samples 4g .. 4g+3 of utterance b
__device__ __forceinline__ float4 normal4(unsigned int g, unsigned int b, unsigned long long seed, unsigned long long offset) {
  const uint4 r = philox4x32_10(make_uint4(g, b, (unsigned int)offset, (unsigned int)(offset >> 32)),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = fmaf((float)r.x, k, 0.5f * k), u1 = (float)r.y * k;
  const float u2 = fmaf((float)r.z, k, 0.5f * k), u3 = (float)r.w * k;
  const float a = sqrtf(-2.f * logf(fminf(u0, 1.f - 0.5f * k))), c = sqrtf(-2.f * logf(fminf(u2, 1.f - 0.5f * k)));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  return make_float4(a * c0, a * s0, c * c1, c * s1);
}

// plain generator (tests, and the exact twin of what the fused kernel draws): out[b, t]
__global__ void philox_normal_kernel(float* __restrict__ out, int T, const unsigned long long* __restrict__ state) {
  const int b = blockIdx.y;
  const unsigned long long seed = state[0], offset = state[1];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; 4 * g < T; g += gridDim.x * blockDim.x) {
    const float4 v = normal4((unsigned int)g, (unsigned int)b, seed, offset);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * g + i < T) out[(size_t)b * T + 4 * g + i] = vv[i];
  }
}

__global__ void rng_advance_kernel(unsigned long long* state) { state[1] += 1ull; }

// ---- the fused kernel ------------------------------------------------------------------------------------
// grid (ceil(n_blocks / 8), B), 128 threads (4 warps x 2 blocks of `hop` outputs; 128 < hop <= 256, hop % 4 == 0).
// smem: xs0[XS] | xs1[XS] | kd[8][K20] | sx[256][8] (spectra, later raw[8][256])
// Resident CTAs per SM the register budget is set for.  The taps-stored-once tile leaves 44.5 KB of shared memory per CTA, so
// five fit; that needs <= 102 registers (96 allocated, 8 bytes of spill) against 127 unconstrained.  Interleaved A/B on one
// B200 (tools/gpu/ab_libs.sh, 3 rounds): value with 8 passes in flight 7.42-7.46e9 -> 7.73-7.80e9 samples/s, GOLF-ff decoder
// 9.65e9 -> 1.00e10; one pass alone 0.3138 -> 0.3184 ms.  (The tile change by itself, 4 CTAs per SM at 127 registers, was
// neutral against the duplicated-tap tile at 3 CTAs per SM.)
#ifndef GOLF_FIRD_MINB
#define GOLF_FIRD_MINB 5
#endif
template <bool PHILOX>
__global__ void __launch_bounds__(128, GOLF_FIRD_MINB) noise_fir_design_kernel(const float* __restrict__ ex, int64_t ex_stride,
                                                               const unsigned long long* __restrict__ rng_state,
                                                               const float* __restrict__ log_mag, const float* __restrict__ window,
                                                               const float* __restrict__ add, int64_t add_stride,
                                                               float* __restrict__ y, int T, int F, int hop, int n_blocks, int xs_len,
                                                               int XS, int TPB, int vec_ok) {
  extern __shared__ __align__(16) float smem[];
  float* xs0 = smem;
  float* xs1 = smem + XS;
  float* kd = smem + 2 * XS;
  float* sx = kd + kDFB * kDK20;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  const int k0 = blockIdx.x * kDFB;
  const int nb = min(kDFB, n_blocks - k0);
  constexpr int p = (kDK - 1) / 2;
  const int start = k0 * hop - p;  // signal position of logical strip index 0

  // ---- the noise strip (two copies, the second shifted by one sample: fir_tile.cuh)
  if (PHILOX) {
    const unsigned long long seed = rng_state[0], offset = rng_state[1];
    const int g_lo = start >= 0 ? start / 4 : -((-start + 3) / 4);  // floor(start / 4)
    const int ng = (xs_len + 1 + 3 + 3) / 4 + 1;
    for (int gi = tid; gi < ng; gi += 128) {
      const int g = g_lo + gi;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g >= 0 && 4 * g < T) v = normal4((unsigned int)g, (unsigned int)b, seed, offset);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pos = 4 * g + q, i = pos - start;
        const float val = (pos >= 0 && pos < T) ? vv[q] : 0.f;
        if (i >= 0 && i < xs_len) xs0[fir_sw16(i)] = val;
        if (i >= 1 && i <= xs_len) xs1[fir_sw16(i - 1)] = val;
      }
    }
  } else {
    const float* __restrict__ exb = ex + (size_t)b * ex_stride;
    for (int i = tid; i <= xs_len; i += 128) {
      const int pos = start + i;
      const bool ok = pos >= 0 && pos < T;
      const float* src = exb + min(max(pos, 0), T - 1);
      if (i < xs_len) cp_async4(xs0 + fir_sw16(i), src, ok);
      if (i > 0) cp_async4(xs1 + fir_sw16(i - 1), src, ok);
    }
  }
  // ---- spectra of the CTA's frames: sx[k][f] = exp(log_mag[b, k0 + f, k])
  {
    const float* __restrict__ lm = log_mag + ((size_t)b * F + k0) * kDN;
    for (int i = tid; i < kDFB * kDN; i += 128) {
      const int f = i / kDN, k = i - f * kDN;
      sx[k * kDFB + f] = f < nb ? expf(__ldg(lm + (size_t)f * kDN + k)) * ((k == 0 || k == kDN - 1) ? 0.5f : 1.f) : 0.f;
    }
  }
  __syncthreads();
  // ---- design: thread d accumulates E[f][d] (even k) and O[f][d] (odd k) for the 8 frames
  float aE[kDFB], aO[kDFB];
#pragma unroll
  for (int f = 0; f < kDFB; ++f) aE[f] = aO[f] = 0.f;
  {
    const int d = tid;
    float cl[16], sl[16], clo[16], slo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      cl[i] = __ldg(&g_design_aa[2][i][d]), sl[i] = __ldg(&g_design_aa[3][i][d]);
      clo[i] = __ldg(&g_design_aa[4][i][d]), slo[i] = __ldg(&g_design_aa[5][i][d]);
    }
#pragma unroll 1
    for (int kh = 0; kh < kDD / 16; ++kh) {
      const float ch = __ldg(&g_design_aa[0][kh][d]), sh = __ldg(&g_design_aa[1][kh][d]);
#pragma unroll
      for (int kl = 0; kl < 16; ++kl) {
        const int kp = 16 * kh + kl;
        const float ce = fmaf(ch, cl[kl], -sh * sl[kl]), co = fmaf(ch, clo[kl], -sh * slo[kl]);
        const float4* xe = reinterpret_cast<const float4*>(sx + (2 * kp) * kDFB);
        const float4* xo = reinterpret_cast<const float4*>(sx + (2 * kp + 1) * kDFB);
        const float4 e0 = xe[0], e1 = xe[1], o0 = xo[0], o1 = xo[1];
        aE[0] = fmaf(ce, e0.x, aE[0]), aE[1] = fmaf(ce, e0.y, aE[1]), aE[2] = fmaf(ce, e0.z, aE[2]), aE[3] = fmaf(ce, e0.w, aE[3]);
        aE[4] = fmaf(ce, e1.x, aE[4]), aE[5] = fmaf(ce, e1.y, aE[5]), aE[6] = fmaf(ce, e1.z, aE[6]), aE[7] = fmaf(ce, e1.w, aE[7]);
        aO[0] = fmaf(co, o0.x, aO[0]), aO[1] = fmaf(co, o0.y, aO[1]), aO[2] = fmaf(co, o0.z, aO[2]), aO[3] = fmaf(co, o0.w, aO[3]);
        aO[4] = fmaf(co, o1.x, aO[4]), aO[5] = fmaf(co, o1.y, aO[5]), aO[6] = fmaf(co, o1.z, aO[6]), aO[7] = fmaf(co, o1.w, aO[7]);
      }
    }
  }
  __syncthreads();  // every thread has read the spectra: the buffer becomes raw[f][0..255]
  {
    const int d = tid;
#pragma unroll
    for (int f = 0; f < kDFB; ++f) {
      sx[f * kDN + d] = aE[f] + aO[f];
      sx[f * kDN + (kDH - d)] = aE[f] - aO[f];
    }
  }
  __syncthreads();
  // ---- taps, duplicated for the packed tile: tap[i] = raw[|i - 255|] * window[i]
  for (int i = tid; i < kDK20; i += 128) {
    const float w = i < kDK ? __ldg(window + i) : 0.f;
    const int dd = i < kDK ? abs(i - kDH) : 0;
#pragma unroll
    for (int f = 0; f < kDFB; ++f) {
      kd[f * kDK20 + i] = i < kDK ? sx[f * kDN + dd] * w : 0.f;
    }
  }
  if (!PHILOX) cp_async_wait_all();
  __syncthreads();
  // ---- FIR: warp w applies blocks 2w and 2w+1
  const int bi = lane / TPB, c = lane - bi * TPB;
  const int blk = 2 * warp + bi;
  if (bi >= 2 || blk >= nb) return;
  const int r0 = c * kR2;
  f32x2 acc[kR2];  // (even-tap sum, odd-tap sum) per output
#pragma unroll
  for (int i = 0; i < kR2; ++i) acc[i] = 0ull;
  fir_tile16_eo(xs0, xs1, blk * hop + r0, kd + blk * kDK20, kDK20, acc);
  float o[kR2];
#pragma unroll
  for (int i = 0; i < kR2; ++i) {
    float lo, hi;
    unpack2(acc[i], lo, hi);
    o[i] = lo + hi;
  }
  float* yb = y + (size_t)b * n_blocks * hop + (size_t)(k0 + blk) * hop + r0;
  const float* ab = add ? add + (size_t)b * add_stride + (size_t)(k0 + blk) * hop + r0 : nullptr;
  if (vec_ok && r0 + kR2 <= hop) {
#pragma unroll
    for (int v = 0; v < kR2 / 4; ++v) {
      float4 r = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
      if (ab) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(ab) + v);
        r.x = a4.x + r.x, r.y = a4.y + r.y, r.z = a4.z + r.z, r.w = a4.w + r.w;
      }
      reinterpret_cast<float4*>(yb)[v] = r;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kR2; ++i)
      if (r0 + i < hop) yb[i] = ab ? ab[i] + o[i] : o[i];
  }
}

}  // namespace golf

using namespace golf;

GOLF_API int golf_noise_fir_design_supported(int n_mag, int hop) {
  return (n_mag == kDN && hop > 128 && hop <= 256 && hop % 4 == 0) ? 1 : 0;
}

GOLF_API int golf_noise_fir_design_fwd(const float* ex, int64_t ex_stride, const uint64_t* rng_state, const float* log_mag,
                                       const float* window, const float* add, int64_t add_stride, float* y, int B, int T, int F,
                                       int n_mag, int hop, void* stream) {
  if ((!ex && !rng_state) || !log_mag || !window || !y || B <= 0 || T <= 0 || F <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  if (!golf_noise_fir_design_supported(n_mag, hop) || B > 65535) return GOLF_ERR_UNSUPPORTED;
  constexpr int p = (kDK - 1) / 2;
  if (T + 2 * p < kDK + hop - 1) return GOLF_ERR_INVALID;
  int n_blocks = (T + 2 * p - (kDK + hop - 1)) / hop + 1;
  if (n_blocks > F) n_blocks = F;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_design_table(st);
  if (rc) return rc;
  const int TPB = ceil_div(hop, kR2);
  const int xs_len = (kDFB - 1) * hop + (TPB - 1) * kR2 + kDK20 + 24;
  const int XS = (int)align_up((size_t)xs_len + 1, 32);
  const size_t sm = ((size_t)2 * XS + (size_t)kDFB * kDK20 + (size_t)kDFB * kDN) * sizeof(float);
  const bool aligned = ((uintptr_t)y % 16 == 0) && (!add || ((uintptr_t)add % 16 == 0 && add_stride % 4 == 0));
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {  // sized for the largest supported hop (256), not for this call's
    const int xs_max = (kDFB - 1) * 256 + 15 * kR2 + kDK20 + 24;
    const size_t sm_max = ((size_t)2 * align_up((size_t)xs_max + 1, 32) + (size_t)kDFB * kDK20 + (size_t)kDFB * kDN) * sizeof(float);
    GOLF_CUDA(cudaFuncSetAttribute(noise_fir_design_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_max));
    GOLF_CUDA(cudaFuncSetAttribute(noise_fir_design_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_max));
    mark_used_on_device(attr);
  }
  const dim3 grid(ceil_div(n_blocks, kDFB), B);
  if (ex)
    noise_fir_design_kernel<false><<<grid, 128, sm, st>>>(ex, ex_stride, nullptr, log_mag, window, add, add_stride, y, T, F, hop,
                                                          n_blocks, xs_len, XS, TPB, aligned ? 1 : 0);
  else
    noise_fir_design_kernel<true><<<grid, 128, sm, st>>>(nullptr, 0, reinterpret_cast<const unsigned long long*>(rng_state), log_mag,
                                                         window, add, add_stride, y, T, F, hop, n_blocks, xs_len, XS, TPB,
                                                         aligned ? 1 : 0);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_philox_normal(float* out, int B, int T, const uint64_t* rng_state, void* stream) {
  if (!out || !rng_state || B <= 0 || T <= 0 || B > 65535) return GOLF_ERR_INVALID;
  philox_normal_kernel<<<dim3(ceil_div(ceil_div(T, 4), 256), B), 256, 0, (cudaStream_t)stream>>>(
      out, T, reinterpret_cast<const unsigned long long*>(rng_state));
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_rng_advance(uint64_t* rng_state, void* stream) {
  if (!rng_state) return GOLF_ERR_INVALID;
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(rng_state));
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
