// lpc_ss.cu -- GOLF-ss: time-varying all-pole (LPC) filter and its adjoint on sm_100a.
//
// Replaces models/filters.py:99-113 (LTVMinimumPhaseFilterPrecise.forward: ex*gain,
// a.reduce_hop_length(), torchlpc.sample_wise_lpc) and torchlpc's autograd, without
// ever materialising the [B,T,M] sample-rate coefficient tensor.
//
// The recurrence  y[t] = e[t] - sum_i a[t,i] y[t-1-i]  is serial in t; at B=32 a
// one-sequence-per-warp mapping would leave 99% of a B200 idle.  It is linear though,
// so time is cut into chunks of Lc samples and solved in three launches:
//
//   1. ss_response_kernel   one WARP per (sequence, chunk).  Lanes 0..M-1 run the
//      homogeneous responses to the M unit initial states, lane M the zero-state
//      response to the excitation; all lanes share the chunk's coefficients, which the
//      warp interpolates tile by tile (ATen arithmetic) into shared memory and reads
//      back as LDS.128 broadcasts.  Output: the chunk's M x M transition matrix Phi and
//      its zero-state end state z   (workspace W, L2 resident).
//   2. ss_stitch_kernel     one warp per sequence walks the chunks: s <- z + Phi s.
//      Phi/z blocks are prefetched by the TMA unit (cp.async.bulk + mbarrier ring).
//   3. ss_solve_kernel      one LANE per (sequence, chunk): re-runs the recurrence from
//      the now-known initial state with the reference's tap order and writes y.
//      Frame coefficient pairs live in registers; inputs/outputs are staged through
//      shared memory so global accesses stay coalesced.
//
// The adjoint  u[t] = g[t] - sum_i a[t+1+i,i] u[t+1+i]  runs through the same three
// kernels on reversed time in transposed form (FORM 1): every product pairs u[s] with
// the coefficient row of its own time s, so coefficient staging is identical.
//
// Algorithmic HBM bytes per sample: 4 (ex) + 4 (y) + 4(M+1)/hop (controls) = 8.383 B at
// M=22, hop=240; the work is ~M(M+1) FMA/sample in pass 1 -- FP32-issue bound, see
// DESIGN.md.
#pragma once
#include "common.cuh"

namespace golf {

struct SsParams {
  const float* in;     // FORM0: ex [B, in_stride]; FORM1: gy [B, L]
  int64_t in_stride;
  const float* gain;   // [B,F] or null (== 1)
  const float* a;      // [B,F,M]
  float* out;          // FORM0: y [B,L]; FORM1: u [B,L]
  float* out2;         // FORM1: d_ex = u * up(gain) [B,L] or null
  float* W;            // [B][C-1][(MP+1)*MP]   chunk responses
  float* S;            // [B][C][MP]            state entering each chunk (processing order)
  const float* zi;     // [B,M] or null (FORM0 only)
  int B, L, F, M, hop, Lc, C, HB;
  float scale;
};

// processing index p, step n within chunk -> absolute time
template <int FORM>
__device__ __forceinline__ int time_of(const SsParams& p, int pi, int n) {
  return FORM == 0 ? pi * p.Lc + n : (p.C - pi) * p.Lc - 1 - n;
}

// ------------------------------------------------------------------ pass 1 --------
template <int MP, int FORM>
__global__ void __launch_bounds__(128) ss_response_kernel(SsParams p) {
  constexpr int SLOTS = (MP + 1 + 31) / 32;  // columns per lane (M+1 columns in all)
  constexpr int TAPS = (MP + 31) / 32;       // taps per lane while staging coefficients
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nresp = p.C - 1;
  const int wg = blockIdx.x * (blockDim.x >> 5) + warp;
  if (wg >= p.B * nresp) return;
  const int b = wg / nresp, pi = wg % nresp;
  float* ctile = smem + warp * (MP * MP + MP);  // [MP rows][MP taps], negated coefficients
  float* etile = ctile + MP * MP;               // [MP] chunk input

  // state registers: FORM0 h[k] holds the output of tile position k (newest = MP-1);
  // FORM1 r[k] holds the (negated) pending sum that is consumed at tile position k.
  float st[SLOTS][MP];
  bool zsr[SLOTS];
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    const int col = lane + 32 * sl;
    zsr[sl] = (col == p.M);
#pragma unroll
    for (int k = 0; k < MP; ++k) {
      const int comp = FORM == 0 ? MP - 1 - k : k;  // state component stored in slot k
      st[sl][k] = (col < p.M && comp == col) ? 1.f : 0.f;
    }
  }

  const float* ab = p.a + (size_t)b * p.F * p.M;
  const float* gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  const float* inb = p.in + (size_t)b * p.in_stride;

#pragma unroll 1
  for (int tile = 0; tile < p.Lc / MP; ++tile) {
    __syncwarp();
    // ---- stage MP coefficient rows + inputs (lanes = rows for the weights, = taps for the values)
    Lerp wrow[TAPS];
    bool vrow[TAPS];
#pragma unroll
    for (int q = 0; q < TAPS; ++q) {
      const int s = lane + 32 * q;
      const int t = time_of<FORM>(p, pi, tile * MP + s);
      vrow[q] = (s < MP) && (t < p.L) && (t >= 0);
      wrow[q] = lerp_at(vrow[q] ? t : 0, p.scale, p.F);
      if (s < MP) {
        float e = 0.f;
        if (vrow[q]) {
          e = inb[t];
          if (FORM == 0 && gb) e = __fmul_rn(e, lerp_apply(wrow[q], gb[wrow[q].i0], gb[wrow[q].i1]));
        }
        etile[s] = e;
      }
    }
    int cur0 = -1;
    float a0v[TAPS], a1v[TAPS];
#pragma unroll
    for (int s = 0; s < MP; ++s) {
      const int q = s / 32, src = s % 32;
      const int i0 = __shfl_sync(0xffffffffu, wrow[q].i0, src);
      const int i1 = __shfl_sync(0xffffffffu, wrow[q].i1, src);
      const float l0 = __shfl_sync(0xffffffffu, wrow[q].l0, src);
      const float l1 = __shfl_sync(0xffffffffu, wrow[q].l1, src);
      const bool ok = __shfl_sync(0xffffffffu, (int)vrow[q], src) != 0;
      if (i0 != cur0) {  // warp-uniform: new frame pair
        cur0 = i0;
#pragma unroll
        for (int q2 = 0; q2 < TAPS; ++q2) {
          const int i = lane + 32 * q2;
          a0v[q2] = i < p.M ? ab[(size_t)i0 * p.M + i] : 0.f;
          a1v[q2] = i < p.M ? ab[(size_t)i1 * p.M + i] : 0.f;
        }
      }
#pragma unroll
      for (int q2 = 0; q2 < TAPS; ++q2) {
        const int i = lane + 32 * q2;
        if (i < MP) {
          const float v = __fmaf_rn(l0, a0v[q2], __fmul_rn(l1, a1v[q2]));
          ctile[s * MP + i] = ok ? -v : 0.f;
        }
      }
    }
    __syncwarp();
    // ---- MP recurrence steps, fully unrolled so the state rotates through registers
#pragma unroll
    for (int s = 0; s < MP; ++s) {
      const float e = etile[s];
      float c[MP];
#pragma unroll
      for (int i4 = 0; i4 < MP / 4; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(ctile + s * MP + 4 * i4);
        c[4 * i4] = v.x, c[4 * i4 + 1] = v.y, c[4 * i4 + 2] = v.z, c[4 * i4 + 3] = v.w;
      }
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        if (FORM == 0) {
          float acc = zsr[sl] ? e : 0.f;
#pragma unroll
          for (int i = 0; i < MP; ++i) acc = __fmaf_rn(c[i], st[sl][(s - 1 - i + 2 * MP) % MP], acc);
          st[sl][s] = acc;
        } else {
          const float u = (zsr[sl] ? e : 0.f) + st[sl][s];
#pragma unroll
          for (int k = 0; k < MP - 1; ++k) st[sl][(s + 1 + k) % MP] = __fmaf_rn(c[k], u, st[sl][(s + 1 + k) % MP]);
          st[sl][s] = __fmul_rn(c[MP - 1], u);
        }
      }
    }
  }
  // ---- emit column `col` of [Phi | z]: W[col][k] = end-state component k
  float* wb = p.W + ((size_t)b * nresp + pi) * ((MP + 1) * MP);
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    const int col = lane + 32 * sl;
    if (col <= p.M) {
#pragma unroll
      for (int k4 = 0; k4 < MP / 4; ++k4) {
        float4 v;
        if (FORM == 0) {
          v = make_float4(st[sl][MP - 1 - 4 * k4], st[sl][MP - 2 - 4 * k4], st[sl][MP - 3 - 4 * k4], st[sl][MP - 4 - 4 * k4]);
        } else {
          v = make_float4(st[sl][4 * k4], st[sl][4 * k4 + 1], st[sl][4 * k4 + 2], st[sl][4 * k4 + 3]);
        }
        *reinterpret_cast<float4*>(wb + col * MP + 4 * k4) = v;
      }
    }
  }
}

constexpr int kStitchStages = 8;
__global__ void ss_stitch_kernel(SsParams p, int MP);

// ------------------------------------------------------------------ pass 3 --------
// One lane per chunk.  GENERIC: coefficients fetched from global every step (any hop /
// chunk relation).  !GENERIC: requires HB % MP == 0 with HB = min(hop, Lc) dividing
// max(hop, Lc): the frame pair sits in registers and is reloaded at tile starts only.
template <int MP, int FORM, bool GENERIC>
__global__ void __launch_bounds__(128) ss_solve_kernel(SsParams p) {
  constexpr int TST = MP + 1;  // tile row stride (odd -> conflict-free per-lane rows)
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (p.C + 31) / 32;
  const int wg = blockIdx.x * (blockDim.x >> 5) + warp;
  if (wg >= p.B * G) return;
  const int b = wg / G, g = wg % G;
  const int pi = g * 32 + lane;
  const bool active = pi < p.C;
  const int pic = active ? pi : p.C - 1;
  float* tile = smem + warp * (FORM == 0 ? 1 : 2) * 32 * TST;
  float* tile2 = tile + 32 * TST;  // FORM1 only: d_ex

  const float* ab = p.a + (size_t)b * p.F * p.M;
  const float* gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  const float* inb = p.in + (size_t)b * p.in_stride;
  float* outb = p.out + (size_t)b * p.L;
  float* out2b = (FORM == 1 && p.out2) ? p.out2 + (size_t)b * p.L : nullptr;

  float st[MP];
  {
    const float* s0 = p.S + ((size_t)b * p.C + pic) * MP;
#pragma unroll
    for (int k = 0; k < MP; ++k) {
      const int comp = FORM == 0 ? MP - 1 - k : k;
      st[k] = active ? s0[comp] : 0.f;
    }
  }
  float na0[MP], na1[MP];  // negated frame pair (registers; !GENERIC)
  float g0 = 1.f, g1 = 1.f, kregf = 0.f;
  int kreg = -1;
#pragma unroll
  for (int i = 0; i < MP; ++i) na0[i] = na1[i] = 0.f;

#pragma unroll 1
  for (int tl = 0; tl < p.Lc / MP; ++tl) {
    const int n0 = tl * MP;
    __syncwarp();
    // ---- stage this tile's input rows (row r = chunk 32g+r), coalesced along time
    for (int r = 0; r < 32; ++r) {
      const int pr = g * 32 + r;
      if (pr >= p.C) break;
      for (int s = lane; s < MP; s += 32) {
        const int t = time_of<FORM>(p, pr, n0 + s);
        tile[r * TST + s] = (t >= 0 && t < p.L) ? inb[t] : 0.f;
      }
    }
    __syncwarp();
    const bool reload = !GENERIC && (n0 % p.HB == 0);
#pragma unroll
    for (int s = 0; s < MP; ++s) {
      const int t = time_of<FORM>(p, pic, n0 + s);
      const bool valid = active && t >= 0 && t < p.L;
      const int tc = valid ? t : 0;
      float nc[MP];
      float gv = 1.f;
      bool slow = GENERIC;
      Lerp w;
      if (!GENERIC) {
        if (s == 0 && reload) {  // warp-uniform: load the frame pair the coming steps live in
          kreg = min(tc / p.hop, p.F - 1);
          kregf = (float)kreg;
          const int k1 = min(kreg + 1, p.F - 1);
#pragma unroll
          for (int i = 0; i < MP; ++i) {
            na0[i] = i < p.M ? -ab[(size_t)kreg * p.M + i] : 0.f;
            na1[i] = i < p.M ? -ab[(size_t)k1 * p.M + i] : 0.f;
          }
          if (FORM == 0 && gb) g0 = gb[kreg], g1 = gb[k1];
          if (FORM == 1 && gb) g0 = gb[kreg], g1 = gb[k1];
        }
        const float src = __fmul_rn(p.scale, (float)tc);
        if (s == (FORM == 0 ? 0 : MP - 1)) {
          // only here can ATen's floor(src) fall outside the register pair (t % hop == 0)
          w = lerp_at(tc, p.scale, p.F);
          slow = __any_sync(0xffffffffu, w.i0 != kreg);
        }
        if (!slow) {
          float l1 = __fsub_rn(src, kregf);
          l1 = fminf(fmaxf(l1, 0.f), 1.f);
          const float l0 = __fsub_rn(1.f, l1);
#pragma unroll
          for (int i = 0; i < MP; ++i) nc[i] = __fmaf_rn(l0, na0[i], __fmul_rn(l1, na1[i]));
          if (gb) gv = __fmaf_rn(l0, g0, __fmul_rn(l1, g1));
        }
      }
      if (slow) {
        w = lerp_at(tc, p.scale, p.F);
        const float* r0 = ab + (size_t)w.i0 * p.M;
        const float* r1 = ab + (size_t)w.i1 * p.M;
#pragma unroll
        for (int i = 0; i < MP; ++i) nc[i] = i < p.M ? -lerp_apply(w, r0[i], r1[i]) : 0.f;
        if (gb) gv = lerp_apply(w, gb[w.i0], gb[w.i1]);
      }
      if (!valid) {
#pragma unroll
        for (int i = 0; i < MP; ++i) nc[i] = 0.f;
      }
      const float x = tile[lane * TST + s];
      if (FORM == 0) {
        float acc = valid ? __fmul_rn(x, gv) : 0.f;
        if (!gb) acc = valid ? x : 0.f;
#pragma unroll
        for (int i = 0; i < MP; ++i) acc = __fmaf_rn(nc[i], st[(s - 1 - i + 2 * MP) % MP], acc);
        st[s] = acc;
        tile[lane * TST + s] = acc;
      } else {
        const float u = (valid ? x : 0.f) + st[s];
#pragma unroll
        for (int k = 0; k < MP - 1; ++k) st[(s + 1 + k) % MP] = __fmaf_rn(nc[k], u, st[(s + 1 + k) % MP]);
        st[s] = __fmul_rn(nc[MP - 1], u);
        tile[lane * TST + s] = u;
        tile2[lane * TST + s] = __fmul_rn(u, gv);
      }
    }
    __syncwarp();
    // ---- write the tile back, coalesced
    for (int r = 0; r < 32; ++r) {
      const int pr = g * 32 + r;
      if (pr >= p.C) break;
      for (int s = lane; s < MP; s += 32) {
        const int t = time_of<FORM>(p, pr, n0 + s);
        if (t >= 0 && t < p.L) {
          outb[t] = tile[r * TST + s];
          if (FORM == 1 && out2b) out2b[t] = tile2[r * TST + s];
        }
      }
    }
  }
  // FORM1 with zi: nothing more here; d_zi is produced by ss_grad_kernel.
}

template <int MP, int FORM>
int launch_mp(const SsParams& p, bool generic, int passes, cudaStream_t st) {
  const int nresp = p.C - 1;
  if (nresp > 0 && (passes & 1)) {
    const int warps = 4;
    const size_t sm = warps * (MP * MP + MP) * sizeof(float);
    ss_response_kernel<MP, FORM><<<ceil_div(p.B * nresp, warps), warps * 32, sm, st>>>(p);
    GOLF_CHECK_LAUNCH();
  }
  if (passes & 2) {
    const size_t sm = (size_t)kStitchStages * (MP + 1) * MP * 4 + (((MP + 3) / 4) * 4 + 4) * 4 + kStitchStages * 8 + 128;
    static bool attr_done = false;
    if (!attr_done) {
      GOLF_CUDA(cudaFuncSetAttribute(ss_stitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_done = true;
    }
    ss_stitch_kernel<<<p.B, 32, sm, st>>>(p, MP);
    GOLF_CHECK_LAUNCH();
  }
  if (passes & 4) {
    const int warps = 4;
    const int G = ceil_div(p.C, 32);
    const size_t sm = warps * (FORM == 0 ? 1 : 2) * 32 * (MP + 1) * sizeof(float);
    if (generic)
      ss_solve_kernel<MP, FORM, true><<<ceil_div(p.B * G, warps), warps * 32, sm, st>>>(p);
    else
      ss_solve_kernel<MP, FORM, false><<<ceil_div(p.B * G, warps), warps * 32, sm, st>>>(p);
    GOLF_CHECK_LAUNCH();
  }
  return GOLF_OK;
}


}  // namespace golf
