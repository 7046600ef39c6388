// fir.cu -- the FIR stages of the GOLF decoder and two frame-rate helpers.
//
//   golf_noise_fir_fwd   models/filters.py:350-384  LTVZeroPhaseFIRFilter.forward: block k
//                        (hop samples) = valid cross-correlation of the zero-padded input
//                        with that frame's 510-tap kernel.  The reference unfolds to
//                        [B*F, 749] and runs a grouped conv1d with B*F groups.
//   golf_room_fir_fwd    models/filters.py:443-450  LTIAcousticFilter.forward.
//   golf_linear_upsample models/audiotensor/audiotensor.py:11-17.
//   golf_rc2lpc_fwd      models/utils.py:581-593 (+ tanh*max_abs of models/filters.py:80).
//
// Both FIRs are register-tiled correlations: a thread owns R=8 consecutive outputs, keeps
// a 12-deep sliding input window in registers, and per 4 taps issues one LDS.128 of taps
// (warp broadcast) and one LDS.128 of new input for 32 FMAs.  Input segment and taps are
// staged once per CTA in shared memory (coalesced), so HBM sees each sample once:
// algorithmic bytes per output sample = 4 (in) + 4 (out) + 4*K/hop (kernels) (+4 `add`).
#include <algorithm>

#include "fir_tile.cuh"

namespace golf {

std::atomic<int> g_fir_x2{1};  // 1 (default): packed-FP32 (FFMA2) kernels where they apply; 0: scalar register tile

// ---- time-varying block FIR ---------------------------------------------------------
// grid (n_blocks, B), 32*ceil(hop/256) threads.  smem: xs[hop + K12 + 32] | ks[K12]
__global__ void __launch_bounds__(256) noise_fir_kernel(const float* __restrict__ ex, int64_t ex_stride,
                                                        const float* __restrict__ kernel, const float* __restrict__ window,
                                                        const float* __restrict__ add, int64_t add_stride,
                                                        float* __restrict__ y, int T, int F, int K, int hop, int n_blocks,
                                                        int K12, int xs_len) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                      // input strip, fir_sw() layout
  float* ks = smem + fir_sw(xs_len) + 4;
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int p = (K - 1) / 2;
  const float* exb = ex + (size_t)b * ex_stride;
  const float* kb = kernel + ((size_t)b * F + k) * K;
  const int start = k * hop - p;  // signal position of logical xs[0]
  for (int i = tid; i < xs_len; i += blockDim.x) {
    const int pos = start + i;
    xs[fir_sw(i)] = (pos >= 0 && pos < T && i < hop + K - 1) ? exb[pos] : 0.f;
  }
  // taps: either final, or (window given) the raw irfft output: fftshift + windowing fused here
  for (int i = tid; i < K12; i += blockDim.x) {
    float v = 0.f;
    if (i < K) v = window ? __fmul_rn(kb[(i + K / 2) % K], window[i]) : kb[i];
    ks[i] = v;
  }
  __syncthreads();
  const int r0 = tid * kR;
  if (r0 >= hop) return;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  fir_tile8_sw(xs, r0, ks, K12, acc);
  float* yb = y + (size_t)b * n_blocks * hop + (size_t)k * hop;
  const float* ab = add ? add + (size_t)b * add_stride + (size_t)k * hop : nullptr;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (r0 + i < hop) yb[r0 + i] = ab ? ab[r0 + i] + acc[i] : acc[i];
}

// ---- time-varying block FIR, packed-FP32 (FFMA2) version ------------------------------------
// One WARP per CTA, NBW consecutive blocks of one utterance per warp (TPB = ceil(hop/16) lanes per block,
// NBW = 32 / TPB: hop 240 -> 2 blocks on 30 lanes), so the grid is thousands of small CTAs that the block
// scheduler keeps balanced over the 148 SMs.  smem: xs0[XS] | xs1[XS] | kd[NBW][2*K20]  (fir_tile.cuh).
__global__ void __launch_bounds__(32) noise_fir_x2_kernel(const float* __restrict__ ex, int64_t ex_stride,
                                                          const float* __restrict__ kernel, const float* __restrict__ window,
                                                          const float* __restrict__ add, int64_t add_stride,
                                                          float* __restrict__ y, int T, int F, int K, int hop, int n_blocks,
                                                          int K20, int xs_len, int XS, int TPB, int NBW, int vec_ok) {
  extern __shared__ __align__(16) float smem[];
  float* xs0 = smem;
  float* xs1 = smem + XS;
  float* kd = smem + 2 * XS;
  const int lane = threadIdx.x, b = blockIdx.y;
  const int k0 = blockIdx.x * NBW;
  const int p = (K - 1) / 2;
  const float* __restrict__ exb = ex + (size_t)b * ex_stride;
  const int start = k0 * hop - p;  // signal position of logical strip index 0
  // staging is latency bound (a warp alone, every element one global load).  The strip goes through
  // cp.async (LDGSTS: global -> shared without a register in between, zero-filled outside the signal), all
  // copies in flight at once; the taps (which need a multiply on the way) follow in register batches
  // while the strip lands.
  for (int i = lane; i <= xs_len; i += 32) {
    const int pos = start + i;
    const bool ok = pos >= 0 && pos < T;
    const float* src = exb + min(max(pos, 0), T - 1);
    if (i < xs_len) cp_async4(xs0 + fir_sw16(i), src, ok);
    if (i > 0) cp_async4(xs1 + fir_sw16(i - 1), src, ok);
  }
  constexpr int kStageU = 17;  // 510 taps: one batch per block
  // taps, duplicated: either final, or (window given) the raw irfft output: fftshift + windowing fused here
  const int nb = min(NBW, n_blocks - k0);
  for (int bi = 0; bi < nb; ++bi) {
    const float* __restrict__ kb = kernel + ((size_t)b * F + k0 + bi) * K;
    float2* dst = reinterpret_cast<float2*>(kd + (size_t)bi * 2 * K20);
    const int half = window ? K / 2 : 0;
    for (int i0 = lane; i0 < K20; i0 += 32 * kStageU) {
      float v[kStageU], wv[kStageU];
#pragma unroll
      for (int u = 0; u < kStageU; ++u) {
        const int i = min(i0 + 32 * u, K - 1);
        const int src = i + half >= K ? i + half - K : i + half;
        v[u] = __ldg(kb + src);
        wv[u] = window ? __ldg(window + i) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < kStageU; ++u) {
        const int i = i0 + 32 * u;
        const float t = i < K ? (window ? __fmul_rn(v[u], wv[u]) : v[u]) : 0.f;
        if (i < K20) dst[i] = make_float2(t, t);
      }
    }
  }
  cp_async_wait_all();
  __syncwarp();
  const int bi = lane / TPB, c = lane - bi * TPB;
  if (bi >= nb) return;
  const int r0 = c * kR2;
  f32x2 acc[kR2 / 2];
#pragma unroll
  for (int i = 0; i < kR2 / 2; ++i) acc[i] = 0ull;
  fir_tile16_x2(xs0, xs1, bi * hop + r0, kd + (size_t)bi * 2 * K20, K20, acc);
  float o[kR2];
#pragma unroll
  for (int i = 0; i < kR2 / 2; ++i) unpack2(acc[i], o[2 * i], o[2 * i + 1]);
  float* yb = y + (size_t)b * n_blocks * hop + (size_t)(k0 + bi) * hop + r0;
  const float* ab = add ? add + (size_t)b * add_stride + (size_t)(k0 + bi) * hop + r0 : nullptr;
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

// ---- room FIR: out[t] = x[t] + sum_{j<n} k[j] x[t-n+j] ---------------------------------
// grid (ceil(T/TILE), B), 128 threads, TILE = 1024 outputs.  Taps are staged as
// [k_0 .. k_{n-1}, 1, 0...] so the direct path is tap n.
constexpr int kRoomTile = 1024;
__global__ void __launch_bounds__(128) room_fir_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                       float* __restrict__ out, int T, int n, int K12, int xs_len) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                      // input strip, fir_sw() layout
  float* ks = smem + fir_sw(xs_len) + 4;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int t0 = blockIdx.x * kRoomTile;
  const float* xb = x + (size_t)b * T;
  for (int i = tid; i < xs_len; i += blockDim.x) {
    const int pos = t0 - n + i;
    // the reference pads x[:-1]: the last sample never feeds the taps (it is never needed:
    // tap j < n reaches at most t-1 <= T-2), only the direct path
    xs[fir_sw(i)] = (pos >= 0 && pos < T) ? xb[pos] : 0.f;
  }
  for (int i = tid; i < K12; i += blockDim.x) ks[i] = i < n ? k[i] : (i == n ? 1.f : 0.f);
  __syncthreads();
  const int r0 = tid * kR;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  fir_tile8_sw(xs, r0, ks, K12, acc);
  float* ob = out + (size_t)b * T;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (t0 + r0 + i < T) ob[t0 + r0 + i] = acc[i];
}


// (A packed-FP32 version of this kernel -- fir_tile16_x2, 2048 outputs per CTA, two strip copies -- measured 16.9 us against
// 15.5 us for this one at 32 x 47 760, 128 taps: with only 140 padded taps the second strip copy and the duplicated taps cost
// more than FFMA2 saves, unlike the 510-tap noise FIR.)

// ---- adjoints of the block FIR --------------------------------------------------------
// d_kernel[b,k,j] = sum_r gy[k*hop + r] * xpad[k*hop + r + j]: the same correlation with gy as
// the taps.  One CTA per (block, utterance), threads own 8 consecutive j.
__global__ void __launch_bounds__(128) noise_fir_dkernel_kernel(const float* __restrict__ gy, const float* __restrict__ ex,
                                                                int64_t ex_stride, float* __restrict__ d_kernel, int T, int F,
                                                                int K, int hop, int n_blocks, int H12, int xs_len) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* gs = smem + xs_len;
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int p = (K - 1) / 2;
  const float* exb = ex + (size_t)b * ex_stride;
  const float* gb = gy + (size_t)b * n_blocks * hop + (size_t)k * hop;
  const int start = k * hop - p;
  for (int i = tid; i < xs_len; i += blockDim.x) {
    const int pos = start + i;
    xs[i] = (pos >= 0 && pos < T && i < hop + K - 1) ? exb[pos] : 0.f;
  }
  for (int i = tid; i < H12; i += blockDim.x) gs[i] = i < hop ? gb[i] : 0.f;
  __syncthreads();
  for (int j0 = tid * kR; j0 < K; j0 += blockDim.x * kR) {
    float acc[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) acc[i] = 0.f;
    fir_tile8(xs + j0, gs, H12, acc);
    float* dk = d_kernel + ((size_t)b * F + k) * K;
#pragma unroll
    for (int i = 0; i < kR; ++i)
      if (j0 + i < K) dk[j0 + i] = acc[i];
  }
}

// d_ex[u - p] = sum_v kernel[block(v)][u - v] gy[v]: every input-gradient sample gathers from the
// (up to ceil((K+hop-1)/hop)) blocks whose windows cover it.  One CTA per hop-sized tile of u (padded
// coordinates); per contributing block one register-tiled correlation of the block's gy (zero
// outside) with the reversed taps.  Deterministic (no atomics).
__global__ void __launch_bounds__(256) noise_fir_dex_kernel(const float* __restrict__ gy, const float* __restrict__ kernel,
                                                            float* __restrict__ d_ex, int T, int F, int K, int hop, int n_blocks,
                                                            int K12, int gs_len) {
  extern __shared__ __align__(16) float smem[];
  float* gs = smem;           // gy of one block, zero-padded by K12 in front and 24 behind
  float* ks = smem + gs_len;  // reversed taps
  const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int p = (K - 1) / 2;
  const int u0 = tile * hop;  // padded coordinate of the first output of this tile
  const int r0 = tid * kR;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  const int k_lo = max(0, (u0 - (K - 1)) / hop - 1), k_hi = min(n_blocks - 1, (u0 + hop - 1) / hop);
  for (int k = k_lo; k <= k_hi; ++k) {
    // out[u] += sum_j kern_k[j] * gy_k[u - k*hop - j],  0 <= u - k*hop - j < hop
    // as a correlation: x[m] = gy_k[m - K12 + (u0 - k*hop)] ... taps t[jj] = kern_k[K12 - 1 - jj]
    __syncthreads();
    const float* gk = gy + (size_t)b * n_blocks * hop + (size_t)k * hop;
    const int off = u0 - k * hop - (K12 - 1);  // gy index seen by x[0]
    for (int i = tid; i < gs_len; i += blockDim.x) {
      const int r = off + i;
      gs[i] = (r >= 0 && r < hop) ? gk[r] : 0.f;
    }
    const float* kb = kernel + ((size_t)b * F + k) * K;
    for (int i = tid; i < K12; i += blockDim.x) {
      const int j = K12 - 1 - i;
      ks[i] = j < K ? kb[j] : 0.f;
    }
    __syncthreads();
    if (r0 < hop) fir_tile8(gs + r0, ks, K12, acc);
  }
  if (r0 < hop) {
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const int t = u0 + r0 + i - p;
      if (r0 + i < hop && t >= 0 && t < T) d_ex[(size_t)b * T + t] = acc[i];
    }
  }
}

// ---- adjoints of the room FIR -----------------------------------------------------------
// d_x[t] = gy[t] + sum_j k[j] gy[t + n - j]   (anti-causal); taps staged reversed with the direct
// path first: ks = [1, k_{n-1}, ..., k_0].
__global__ void __launch_bounds__(128) room_fir_dx_kernel(const float* __restrict__ gy, const float* __restrict__ k,
                                                          float* __restrict__ d_x, int T, int n, int K12, int xs_len) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* ks = smem + xs_len;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int t0 = blockIdx.x * kRoomTile;
  const float* gb = gy + (size_t)b * T;
  for (int i = tid; i < xs_len; i += blockDim.x) {
    const int pos = t0 + i;
    xs[i] = pos < T ? gb[pos] : 0.f;
  }
  for (int i = tid; i < K12; i += blockDim.x) ks[i] = i == 0 ? 1.f : (i <= n ? k[n - i] : 0.f);
  __syncthreads();
  const int r0 = tid * kR;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  fir_tile8(xs + r0, ks, K12, acc);
  float* ob = d_x + (size_t)b * T;
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (t0 + r0 + i < T) ob[t0 + r0 + i] = acc[i];
}

// d_k[j] = sum_{b,t} gy[b,t] x[b,t-n+j]: per CTA a tile of t for one utterance, 8 taps per thread
// via the same correlation (gy as taps), partial sums added atomically (float adds: the order,
// hence the last bit, is not reproducible run to run).
__global__ void __launch_bounds__(32) room_fir_dk_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                         float* __restrict__ d_k, int T, int n, int tile_len, int G12,
                                                         int xs_len) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* gs = smem + xs_len;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int t0 = blockIdx.x * tile_len;
  const float* __restrict__ xb = x + (size_t)b * T;
  const float* __restrict__ gb = gy + (size_t)b * T;
  // a lone warp per CTA: stage with batched, branch-free loads (the loop was bound by one global-load
  // latency per 32 elements)
  constexpr int U = 8;
  for (int i0 = tid; i0 < xs_len; i0 += 32 * U) {
    float v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int pos = t0 - n + i0 + 32 * q;
      const float raw = __ldg(xb + min(max(pos, 0), T - 1));
      v[q] = (pos >= 0 && pos < T) ? raw : 0.f;
    }
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (i0 + 32 * q < xs_len) xs[i0 + 32 * q] = v[q];
  }
  for (int i0 = tid; i0 < G12; i0 += 32 * U) {
    float v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int i = i0 + 32 * q;
      const float raw = __ldg(gb + min(t0 + i, T - 1));
      v[q] = (i < tile_len && t0 + i < T) ? raw : 0.f;
    }
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (i0 + 32 * q < G12) gs[i0 + 32 * q] = v[q];
  }
  __syncthreads();
  // d_k[j] += sum_r gs[r] * xs[r + j].  With n <= 128 taps only 16 lanes own a group of 8 taps: the two
  // half-warps then split the tile's samples (G12 / 2 each, a multiple of 12)
  const int groups = (n + kR - 1) / kR;
  const bool split = groups <= 16 && (G12 / 2) % 12 == 0;
  const int h = split ? tid / 16 : 0, q = split ? tid % 16 : tid;
  const int span = split ? G12 / 2 : G12;
  for (int j0 = q * kR; j0 < n; j0 += (split ? 16 : 32) * kR) {
    float acc[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) acc[i] = 0.f;
    fir_tile8(xs + j0 + h * span, gs + h * span, span, acc);
#pragma unroll
    for (int i = 0; i < kR; ++i)
      if (j0 + i < n) atomicAdd(d_k + j0 + i, acc[i]);
  }
}

// Persistent version for n <= 128 taps.  The kernel above issues one atomicAdd per (CTA, tap): at B = 32 x 2 s
// that is ~2000 float atomics on each of 127 addresses, and same-address atomics serialise in L2 (~40 ns
// each: 97 us, IPC 0.3).  Here a grid of one CTA per SM walks the (tile, utterance) tasks, four warps per CTA
// each with its own strip, partial sums stay in registers across tasks, the half-warps and the warps are
// combined through shuffles / shared memory, and each CTA issues ONE atomic per tap (148 per address).
constexpr int kDkWarps = 4;
__global__ void __launch_bounds__(32 * kDkWarps) room_fir_dk2_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                                     float* __restrict__ d_k, int B, int T, int n, int tile_len,
                                                                     int G12, int xs_len, int n_tiles) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = smem + (size_t)warp * (xs_len + G12);
  float* gs = xs + xs_len;
  float* red = smem + (size_t)kDkWarps * (xs_len + G12);  // [kDkWarps][128]
  const int h = lane / 16, q = lane % 16, span = G12 / 2, j0 = q * kR;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  const int n_tasks = n_tiles * B;
  constexpr int U = 8;
  for (int task = blockIdx.x * kDkWarps + warp; task < n_tasks; task += gridDim.x * kDkWarps) {
    const int b = task / n_tiles, t0 = (task % n_tiles) * tile_len;
    const float* __restrict__ xb = x + (size_t)b * T;
    const float* __restrict__ gb = gy + (size_t)b * T;
    __syncwarp();
    for (int i0 = lane; i0 < xs_len; i0 += 32 * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = t0 - n + i0 + 32 * u;
        const float raw = __ldg(xb + min(max(pos, 0), T - 1));
        v[u] = (pos >= 0 && pos < T) ? raw : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + 32 * u < xs_len) xs[i0 + 32 * u] = v[u];
    }
    for (int i0 = lane; i0 < G12; i0 += 32 * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 32 * u;
        const float raw = __ldg(gb + min(t0 + i, T - 1));
        v[u] = (i < tile_len && t0 + i < T) ? raw : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + 32 * u < G12) gs[i0 + 32 * u] = v[u];
    }
    __syncwarp();
    if (j0 < n) fir_tile8(xs + j0 + h * span, gs + h * span, span, acc);  // acc[i] += sum_r gs[r] xs[r + j0 + i]
  }
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  if (h == 0) {
#pragma unroll
    for (int i = 0; i < kR; ++i) red[warp * 128 + j0 + i] = acc[i];
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < n) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kDkWarps; ++w) sum += red[w * 128 + j];
    atomicAdd(d_k + j, sum);
  }
}

// ---- FIR design, first step: spectrum = exp(log_mag) + 0j (models/filters.py:295-296) -----------------
// One pass instead of torch's exp kernel + real->complex copy; expf is the same libdevice routine ATen calls.
__global__ void exp_to_complex_kernel(const float* __restrict__ x, float2* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float2(expf(x[i]), 0.f);
}

// ---- linear upsample -----------------------------------------------------------------
__global__ void linear_upsample_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int hop, int L,
                                       float scale) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (d >= L) return;
  const Lerp w = lerp_at(d, scale, n);
  const float* xr = x + (size_t)r * n;
  out[(size_t)r * L + d] = lerp_apply(w, xr[w.i0], xr[w.i1]);
}

// ---- reflection coefficients -> LPC (step-up), one thread per frame ----------------------
constexpr int kMaxOrder = 64;
__global__ void rc2lpc_kernel(const float* __restrict__ logits, float* __restrict__ a, int N, int M, float max_abs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  float cur[kMaxOrder + 1], nxt[kMaxOrder + 1];
  const float* lg = logits + (size_t)idx * M;
  cur[0] = 1.f;
  for (int n = 0; n < M; ++n) {
    const float kn = __fmul_rn(tanhf(lg[n]), max_abs);
    // poly (degree n) extended by a zero, plus kn * reversed
    for (int i = 0; i <= n + 1; ++i) {
      const float pi_ = i <= n ? cur[i] : 0.f;
      const float pr = (n + 1 - i) <= n ? cur[n + 1 - i] : 0.f;
      nxt[i] = __fadd_rn(pi_, __fmul_rn(kn, pr));
    }
    for (int i = 0; i <= n + 1; ++i) cur[i] = nxt[i];
  }
  for (int i = 0; i < M; ++i) a[(size_t)idx * M + i] = cur[i + 1];
}

// adjoint of the step-up recursion: one thread per frame recomputes the polynomial of every level (kept in
// local memory, triangular: level n has n+2 coefficients) and walks the levels backwards,
//   d_k_n = sum_j dpoly[j] ext[n+1-j],   d_ext[j] = dpoly[j] + k_n dpoly[n+1-j],
// then d_logit = d_k * max_abs * (1 - tanh^2).
constexpr int kMaxOrderBwd = 40;
__global__ void rc2lpc_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ d_a, float* __restrict__ d_logits,
                                  int N, int M, float max_abs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  float lev[(kMaxOrderBwd + 1) * (kMaxOrderBwd + 2) / 2 + kMaxOrderBwd + 2];  // level n starts at n*(n+3)/2... offsets below
  float kk[kMaxOrderBwd], th[kMaxOrderBwd], dp[kMaxOrderBwd + 2], dn[kMaxOrderBwd + 2];
  const float* lg = logits + (size_t)idx * M;
  auto off = [](int n) { return n * (n + 3) / 2; };  // level n holds n+2 coefficients: 2, 3, 4, ... -> offsets 0, 2, 5, 9
  // level 0: [1, k0]
  th[0] = tanhf(lg[0]);
  kk[0] = __fmul_rn(th[0], max_abs);
  lev[0] = 1.f, lev[1] = kk[0];
  for (int n = 1; n < M; ++n) {
    th[n] = tanhf(lg[n]);
    kk[n] = __fmul_rn(th[n], max_abs);
    const float* cur = lev + off(n - 1);  // n+1 coefficients
    float* nxt = lev + off(n);            // n+2 coefficients
    for (int i = 0; i <= n + 1; ++i) {
      const float pi_ = i <= n ? cur[i] : 0.f;
      const float pr = (n + 1 - i) <= n ? cur[n + 1 - i] : 0.f;
      nxt[i] = __fadd_rn(pi_, __fmul_rn(kk[n], pr));
    }
  }
  const float* g = d_a + (size_t)idx * M;
  dp[0] = 0.f;
  for (int i = 0; i < M; ++i) dp[i + 1] = g[i];
  float* dl = d_logits + (size_t)idx * M;
  for (int n = M - 1; n >= 1; --n) {
    const float* ext = lev + off(n - 1);  // ext[j] = poly^{(n-1)}[j] for j <= n, 0 at j = n+1
    float dk = 0.f;
    for (int j = 0; j <= n + 1; ++j) {
      const int r = n + 1 - j;
      if (r <= n) dk = __fmaf_rn(dp[j], ext[r], dk);
    }
    for (int j = 0; j <= n; ++j) dn[j] = __fmaf_rn(kk[n], dp[n + 1 - j], dp[j]);
    for (int j = 0; j <= n; ++j) dp[j] = dn[j];
    dl[n] = dk * max_abs * (1.f - th[n] * th[n]);
  }
  dl[0] = dp[1] * max_abs * (1.f - th[0] * th[0]);
}

}  // namespace golf

using namespace golf;

GOLF_API void golf_fir_set_variant(int x2) { g_fir_x2 = x2 ? 1 : 0; }

GOLF_API int golf_noise_fir_fwd(const float* ex, int64_t ex_stride, const float* kernel, const float* window,
                                const float* add, int64_t add_stride, float* y, int B, int T, int F, int K, int hop,
                                void* stream) {
  if (!ex || !kernel || !y || B <= 0 || T <= 0 || F <= 0 || K <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  const int p = (K - 1) / 2;
  if (T + 2 * p < K + hop - 1) return GOLF_ERR_INVALID;
  int n_blocks = (T + 2 * p - (K + hop - 1)) / hop + 1;
  if (n_blocks > F) n_blocks = F;
  if (hop > 2048) return GOLF_ERR_UNSUPPORTED;
  if (window && (K & 1)) return GOLF_ERR_UNSUPPORTED;  // fused fftshift assumes an even tap count (2*(n_mag-1))
  if (B > 65535) return GOLF_ERR_UNSUPPORTED;
  // packed-FP32 kernel: hop a multiple of 4 (float4 strip reads), at most 32 lanes per block
  if (g_fir_x2 && hop % 4 == 0 && hop <= 32 * kR2) {
    const int TPB = ceil_div(hop, kR2), NBW = 32 / TPB;
    const int K20 = ceil_div(K, kTapStep) * kTapStep;
    const int xs_len = (NBW - 1) * hop + (TPB - 1) * kR2 + K20 + 24;  // last strip index a lane reads, + 1
    const int XS = (int)align_up((size_t)xs_len + 1, 32);
    const size_t sm = ((size_t)2 * XS + (size_t)NBW * 2 * K20) * sizeof(float);
    if (sm <= 48 * 1024) {
      const bool aligned = ((uintptr_t)y % 16 == 0) && (!add || ((uintptr_t)add % 16 == 0 && add_stride % 4 == 0));
      noise_fir_x2_kernel<<<dim3(ceil_div(n_blocks, NBW), B), 32, sm, (cudaStream_t)stream>>>(
          ex, ex_stride, kernel, window, add, add_stride, y, T, F, K, hop, n_blocks, K20, xs_len, XS, TPB, NBW, aligned ? 1 : 0);
      GOLF_CHECK_LAUNCH();
      return GOLF_OK;
    }
  }
  const int K12 = ceil_div(K, 12) * 12;
  const int threads = 32 * ceil_div(hop, 32 * kR);
  const int xs_len = (int)align_up((size_t)threads * kR + K12 + 24, 4);
  const size_t sm = (size_t)(fir_sw(xs_len) + 4 + K12) * sizeof(float);
  if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
  dim3 grid(n_blocks, B);
  noise_fir_kernel<<<grid, threads, sm, (cudaStream_t)stream>>>(ex, ex_stride, kernel, window, add, add_stride, y, T, F, K,
                                                              hop, n_blocks, K12, xs_len);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_room_fir_fwd(const float* x, const float* k, float* out, int B, int T, int n, void* stream) {
  if (!x || !k || !out || B <= 0 || T <= 0 || n <= 0) return GOLF_ERR_INVALID;
  if (B > 65535) return GOLF_ERR_UNSUPPORTED;
  const int K12 = ceil_div(n + 1, 12) * 12;
  const int xs_len = (int)align_up((size_t)kRoomTile + K12 + 24, 4);
  const size_t sm = (size_t)(fir_sw(xs_len) + 4 + K12) * sizeof(float);
  if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(T, kRoomTile), B);
  room_fir_kernel<<<grid, 128, sm, (cudaStream_t)stream>>>(x, k, out, T, n, K12, xs_len);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}


GOLF_API int golf_noise_fir_bwd(const float* gy, const float* ex, int64_t ex_stride, const float* kernel, float* d_ex,
                                float* d_kernel, int B, int T, int F, int K, int hop, void* stream) {
  if (!gy || !ex || !kernel || B <= 0 || T <= 0 || F <= 0 || K <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  const int p = (K - 1) / 2;
  if (T + 2 * p < K + hop - 1) return GOLF_ERR_INVALID;
  int n_blocks = (T + 2 * p - (K + hop - 1)) / hop + 1;
  if (n_blocks > F) n_blocks = F;
  if (hop > 2048) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int K12 = ceil_div(K, 12) * 12;
  if (d_kernel) {
    // frames beyond n_blocks receive no gradient
    if (n_blocks < F) {
      for (int b = 0; b < B; ++b)
        GOLF_CUDA(cudaMemsetAsync(d_kernel + ((size_t)b * F + n_blocks) * K, 0, (size_t)(F - n_blocks) * K * sizeof(float), st));
    }
    const int H12 = ceil_div(hop, 12) * 12;
    const int threads = 128;
    const int xs_len = (int)align_up((size_t)ceil_div(K, threads * kR) * threads * kR + H12 + 24, 4);
    const size_t sm = (size_t)(xs_len + H12) * sizeof(float);
    if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
    noise_fir_dkernel_kernel<<<dim3(n_blocks, B), threads, sm, st>>>(gy, ex, ex_stride, d_kernel, T, F, K, hop, n_blocks, H12,
                                                                    xs_len);
    GOLF_CHECK_LAUNCH();
  }
  if (d_ex) {
    const int threads = 32 * ceil_div(hop, 32 * kR);
    const int gs_len = (int)align_up((size_t)threads * kR + K12 + 24, 4);
    const size_t sm = (size_t)(gs_len + K12) * sizeof(float);
    if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
    const int tiles = ceil_div(T + p, hop);  // padded coordinates u = t + p, t in [0, T)
    noise_fir_dex_kernel<<<dim3(tiles, B), threads, sm, st>>>(gy, kernel, d_ex, T, F, K, hop, n_blocks, K12, gs_len);
    GOLF_CHECK_LAUNCH();
  }
  return GOLF_OK;
}

GOLF_API int golf_room_fir_bwd(const float* gy, const float* x, const float* k, float* d_x, float* d_k, int B, int T, int n,
                               void* stream) {
  if (!gy || !x || !k || B <= 0 || T <= 0 || n <= 0) return GOLF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int K12 = ceil_div(n + 1, 12) * 12;
  if (d_x) {
    const int xs_len = (int)align_up((size_t)kRoomTile + K12 + 24, 4);
    const size_t sm = (size_t)(xs_len + K12) * sizeof(float);
    if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
    room_fir_dx_kernel<<<dim3(ceil_div(T, kRoomTile), B), 128, sm, st>>>(gy, k, d_x, T, n, K12, xs_len);
    GOLF_CHECK_LAUNCH();
  }
  if (d_k) {
    GOLF_CUDA(cudaMemsetAsync(d_k, 0, (size_t)n * sizeof(float), st));
    const int tile_len = 1536;
    const int G12 = ceil_div(tile_len, 12) * 12;
    const int xs_len = (int)align_up((size_t)ceil_div(n, 32 * kR) * 32 * kR + G12 + 24, 4);
    const size_t sm = (size_t)(xs_len + G12) * sizeof(float);
    if (sm > 48 * 1024) return GOLF_ERR_UNSUPPORTED;
    if (n <= 128) {
      const int tl = 768, g12 = 768;  // two half-warp spans of 384 = 32 * 12 samples
      const int xl = (int)align_up((size_t)128 + g12 + 24, 4);
      const size_t sm2 = ((size_t)kDkWarps * (xl + g12) + kDkWarps * 128) * sizeof(float);
      const int n_tiles = ceil_div(T, tl);
      int sms = 148;
      {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
      }
      const int grid = std::min(sms, ceil_div(n_tiles * B, kDkWarps));
      room_fir_dk2_kernel<<<grid, 32 * kDkWarps, sm2, st>>>(gy, x, d_k, B, T, n, tl, g12, xl, n_tiles);
      GOLF_CHECK_LAUNCH();
      return GOLF_OK;
    }
    room_fir_dk_kernel<<<dim3(ceil_div(T, tile_len), B), 32, sm, st>>>(gy, x, d_k, T, n, tile_len, G12, xs_len);
    GOLF_CHECK_LAUNCH();
  }
  return GOLF_OK;
}

GOLF_API int golf_exp_to_complex(const float* x, float* out_interleaved, int64_t n, void* stream) {
  if (!x || !out_interleaved || n <= 0) return GOLF_ERR_INVALID;
  exp_to_complex_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<float2*>(out_interleaved), n);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_linear_upsample(const float* x, float* out, int R, int n, int hop, void* stream) {
  if (!x || !out || R <= 0 || n <= 0 || hop <= 0) return GOLF_ERR_INVALID;
  const int64_t L = (int64_t)(n - 1) * hop + 1;
  if (L > INT32_MAX || R > 65535) return GOLF_ERR_UNSUPPORTED;
  dim3 grid(ceil_div((int)L, 256), R);
  linear_upsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, n, hop, (int)L, lerp_scale(n, hop));
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_rc2lpc_fwd(const float* logits, float* a, int N, int M, float max_abs, void* stream) {
  if (!logits || !a || N <= 0 || M <= 0) return GOLF_ERR_INVALID;
  if (M > kMaxOrder) return GOLF_ERR_UNSUPPORTED;
  rc2lpc_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(logits, a, N, M, max_abs);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API int golf_rc2lpc_bwd(const float* logits, const float* d_a, float* d_logits, int N, int M, float max_abs, void* stream) {
  if (!logits || !d_a || !d_logits || N <= 0 || M <= 0) return GOLF_ERR_INVALID;
  if (M > kMaxOrderBwd) return GOLF_ERR_UNSUPPORTED;
  rc2lpc_bwd_kernel<<<ceil_div(N, 64), 64, 0, (cudaStream_t)stream>>>(logits, d_a, d_logits, N, M, max_abs);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}
