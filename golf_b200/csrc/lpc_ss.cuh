// lpc_ss.cuh -- GOLF-ss: time-varying all-pole (LPC) filter and its adjoint on sm_100a.
//
// Replaces models/filters.py:99-113 (LTVMinimumPhaseFilterPrecise.forward: ex*gain,
// a.reduce_hop_length(), torchlpc.sample_wise_lpc) and torchlpc's autograd, without
// ever materialising the [B,T,M] sample-rate coefficient tensor.
//
// The recurrence  y[t] = e[t] - sum_i a[t,i] y[t-1-i]  is serial in t; at B=32 a
// one-sequence-per-warp mapping would leave 99% of a B200 idle.  It is linear though,
// so time is cut into chunks of Lc samples and solved in three kinds of launch:
//
//   1. ss_response_kernel   per chunk, M+1 recurrences share the chunk's coefficients: the
//      homogeneous responses to the M unit initial states and the zero-state response to
//      the excitation.  A lane runs 4 of those columns (so every coefficient it reads from
//      shared memory feeds 4 FMAs -- the smem-operand : FP32-pipe ratio of the SM), a
//      warp hosts 32/ceil((M+1)/4) chunks; the warp interpolates the coefficient rows tile
//      by tile (ATen arithmetic) into shared memory.  Output: each chunk's M x M
//      transition matrix Phi and zero-state end state z (workspace W, L2 resident).
//   2. ss_stitch_kernel     one warp per sequence walks the chunks: s <- z + Phi s.
//      Phi/z blocks are prefetched by the TMA unit (cp.async.bulk + mbarrier ring).
//   3. ss_solve_kernel      one LANE per (sequence, chunk): re-runs the recurrence from
//      the now-known initial state and writes y; also records the state it ends in (E).
//      Frame coefficient pairs live in registers; inputs/outputs are staged through
//      shared memory so global accesses stay coalesced.
//   4. (refinement, optional) the end state E_p a chunk really reached differs from the
//      stitched S_{p+1} by the float32 error of Phi/z.  ss_stitch_kernel(refine) propagates
//      that mismatch, delta_{p+1} = Phi_p delta_p + (E_p - S_{p+1}), corrects S, and the
//      solve runs once more: the result is then a piecewise *sequential* float32
//      recurrence with consistent states, i.e. as accurate as the reference's own loop
//      even for high-gain / near-unstable filters where Phi is badly conditioned.
//
// The adjoint  u[t] = g[t] - sum_i a[t+1+i,i] u[t+1+i]  runs through the same kernels on
// reversed time in transposed form (FORM 1): every product pairs u[s] with the
// coefficient row of its own time s, so coefficient staging is identical.
//
// Algorithmic HBM bytes per sample: 4 (ex) + 4 (y) + 4(M+1)/hop (controls) = 8.383 B at
// M=22, hop=240; the work is ~M(M+1) FMA/sample in pass 1 -- FP32-pipe bound, see
// DESIGN.md.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "fir_tile.cuh"  // f32x2 / pack2 / unpack2 / ffma2

namespace golf {

extern std::atomic<int> g_solve_systolic;  // 1 (default): 4-lanes-per-chunk solve where it applies; 0: lane-per-chunk
extern std::atomic<int> g_ss_tail;         // stitch + solve + refinement (+ room) in one cluster launch: 0 (default) never, 1 where it applies, 2 small batches
constexpr int kTailAutoMaxBatch = 8;

struct SsParams {
  const float* in;     // FORM0: ex [B, in_stride]; FORM1: gy [B, L]
  int64_t in_stride;
  const float* gain;   // [B,F] or null (== 1)
  const float* a;      // [B,F,M]
  float* out;          // FORM0: y [B,L]; FORM1: u [B,L]
  float* out2;         // FORM1: d_ex = u * up(gain) [B,L] or null
  float* W;            // [B][C-1][(MP+1)*MP]   chunk responses
  float* S;            // [B][C][MP]            state entering each chunk (processing order)
  float* E;            // [B][C][MP]            state each chunk ended in (written by the solve)
  const float* zi;     // [B,M] or null (FORM0 only)
  unsigned int* flags; // [B][2]: float bits of max |E_p - S_{p+1}| and max |S_p| (adaptive refinement)
  int B, L, F, M, hop, Lc, C, HB;
  float scale;
  float refine_tol;    // refine sequence b only if mismatch > refine_tol * max|S| (0: always)
  // ---- one-launch tail (lpc_ss_tail.cuh): two-level stitch + solve + refinement (+ room FIR) per sequence
  float* Gw;           // [B][NG][(MP+1)*MP]  composed transition [Phi_grp | z_grp] of each group of G chunks
  float* Sg;           // [B][NG][MP]         state entering each group
  float* Dg;           // [B][NG][MP]         refinement: mismatch accumulated over a group, then the correction entering it
  const float* room_k; // [room_n] learned taps of the room FIR fused behind the filter, or null
  float* room_out;     // [B][L] final output when room_k is given (p.out then holds the filter output y)
  int NG, G, room_n;
};

// processing index p, step n within chunk -> absolute time
template <int FORM>
__device__ __forceinline__ int time_of(const SsParams& p, int pi, int n) {
  return FORM == 0 ? pi * p.Lc + n : (p.C - pi) * p.Lc - 1 - n;
}

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }

// ------------------------------------------------------------------ pass 1 --------
// Chunk responses.  Two resources decide the mapping:
//  * operand bandwidth: shared memory hands a warp 32 lane-words per cycle, the FP32 pipe wants 128
//    lane-FMAs per cycle, so every coefficient a lane reads must feed >= 4 of its FMAs: a lane owns
//    NC columns of one chunk's [Phi | z] (LPC = ceil((MT+1)/NC) lanes per chunk, CPW = 32/LPC chunks
//    side by side in a warp);
//  * the register file: 4 schedulers x 16384 registers.  NC*MP state registers + a coefficient row per lane:
//    NC = 5 (M = 22: 5 lanes x 5 columns, 6 chunks per warp) needs every register of the SM for 8 warps and is the
//    fastest mapping for ONE decoder pass at a time by ~2 %; NC = 3 (8 lanes per chunk, 4 chunks per warp, all
//    32 lanes busy, 168 registers, 12 resident warps) is as fast alone (76 vs 78 us) and leaves a third of the
//    register file to the other kernels when several passes are in flight (bench.py `value`: 7.19e9 vs 6.77e9
//    samples/s), so it is the default.  CTAs are 4 warps that meet at a barrier every tile, and each lane's NC FMA
//    chains are independent so one warp alone can keep the pipe busy.
// MT = taps actually summed (M <= MT <= MP): the ring keeps MP outputs so that Lc % MP == 0 tiles
// work, but only MT products per column and step are issued (M = 22: 22 of 24, 8 % fewer FMAs).
// Per step a lane issues ceil(MT/4) LDS.128 for its chunk's coefficient row and NC*MT FFMA.
// Coefficient rows are interpolated (ATen arithmetic, negated) tile by tile by lanes mapped to
// (chunk, group of TPL taps) whose frame pair sits in registers.
#ifndef GOLF_RESP_NC
#define GOLF_RESP_NC 3  // columns per lane at padded order 24; -DGOLF_RESP_NC=4|5|6 for experiments (tools/resp_nc_sweep.sh)
#endif
#ifndef GOLF_RESP_F2
// 1: direct-form recurrence on packed FP32 (fma.rn.f32x2; SASS FFMA2), half the FMA issue slots.  Measured on B200
// (B = 32 x 47 760, M = 22): the kernel takes 73.8 us either way -- it is bound by the per-tile latency chain, not by
// issue -- and the two partial sums per column round differently from the sequential solve, so the chunk-boundary
// mismatch rises above refine_tol and the refinement launches run (filter 137 -> 197 us alone; 7.42e9 vs 7.44e9
// samples/s with 8 passes in flight).  Kept as a compile-time experiment, off by default.
#define GOLF_RESP_F2 0
#endif
#ifndef GOLF_RESP_WPB
#define GOLF_RESP_WPB 4  // warps per CTA: they meet at a barrier every tile, which keeps them in the same part of the 48 KB loop body (81 -> 77 us; 8 per CTA: the same; 2 per CTA: value with 8 passes in flight 7.76e9 -> 7.67e9 in an interleaved A/B)
#endif
#ifndef GOLF_RESP_RES_BIG
#define GOLF_RESP_RES_BIG 8  // resident warps per SM the padded orders 32 / 40 are compiled for (4: 255 registers, one CTA)
#endif
template <int MP>
struct RespCfg {
  static constexpr int NC = (MP == 24) ? GOLF_RESP_NC : 4;  // columns per lane
  // resident one-warp CTAs per SM the register budget is set for: NC*MP state registers + ~80
#ifdef GOLF_RESP_RES
  static constexpr int RES = MP > 24 ? 4 : GOLF_RESP_RES;
#else
  static constexpr int RES = MP > 24 ? GOLF_RESP_RES_BIG : (NC <= 4 ? 12 : 8);
#endif
};

template <int MP, int MT, int FORM>
__global__ void __launch_bounds__(32 * GOLF_RESP_WPB, (RespCfg<MP>::RES / GOLF_RESP_WPB > 0 ? RespCfg<MP>::RES / GOLF_RESP_WPB : 1)) ss_response_kernel(SsParams p) {
  constexpr int WPB = GOLF_RESP_WPB;
  constexpr int NC = RespCfg<MP>::NC;
  constexpr int LPC = (MT + 1 + NC - 1) / NC;  // lanes per chunk
  constexpr int CPW = 32 / LPC;                // chunks per warp
  constexpr int ROWS = CPW * MP;               // coefficient rows staged per tile
  constexpr int NA = (ROWS + 31) / 32;
  constexpr int TSTR = MP * MP + 4;  // per-chunk tile stride: +4 floats puts the chunks in different bank quads
  constexpr int SQ0 = 32 / CPW;                          // staging lanes available per chunk
  constexpr int TPL = ((MP + SQ0 - 1) / SQ0 + 1) / 2 * 2;  // taps per staging lane (even: 8-byte stores)
  constexpr int SQ = (MP + TPL - 1) / TPL;                // staging lanes used per chunk
  constexpr int NQ = (MT + 3) / 4;                        // coefficient quads read per step
  constexpr bool F2 = GOLF_RESP_F2 && FORM == 0 && MT <= MP - 1 && MP % 4 == 0;
  static_assert(LPC * CPW <= 32 && SQ * CPW <= 32 && MT <= MP && MT >= 1, "response kernel geometry");
  extern __shared__ __align__(128) float smem_all[];
  constexpr int kWarpFloats = (CPW * TSTR + ROWS * 5 + 31) / 32 * 32;  // shared memory of one warp
  float* smem = smem_all + (WPB > 1 ? (threadIdx.x >> 5) * kWarpFloats : 0);
  const int lane = threadIdx.x & 31;
  const int nresp = p.C - 1;
  const int wps = (nresp + CPW - 1) / CPW;  // warps per sequence
  const int wg_raw = blockIdx.x * WPB + (threadIdx.x >> 5);
  const bool warp_on = wg_raw < p.B * wps;
  if (WPB == 1 && !warp_on) return;
  const int wg = warp_on ? wg_raw : p.B * wps - 1;  // WPB > 1: idle warps shadow the last task (they must reach the barriers)
  const int b = wg / wps, w0 = (wg % wps) * CPW;
  const int gi = min(lane / LPC, CPW - 1), li = lane % LPC;
  const bool lane_on = lane < LPC * CPW;
  const int pi = w0 + gi;
  const bool chunk_on = warp_on && lane_on && pi < nresp;

  float* ctile = smem;                                       // [CPW][MP rows][MP taps] negated coefficients
  float* etile = ctile + CPW * TSTR;                         // [CPW][MP] chunk inputs
  float4* lw = reinterpret_cast<float4*>(etile + ROWS);      // [CPW][MP] (l0, l1, i0, i1) per row
  const float* __restrict__ ab = p.a + (size_t)b * p.F * p.M;
  const float* __restrict__ gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  // in == nullptr: transition matrices only (they depend on the coefficients alone, so the host
  // can run this launch on a side stream while the excitation is still being produced); the
  // zero-state responses then come from ss_solve_kernel(round -1)
  const float* __restrict__ inb = p.in ? p.in + (size_t)b * p.in_stride : nullptr;

  // state: FORM0 st[c][k] = output of tile position k; FORM1 st[c][k] = (negated) pending sum
  // consumed at tile position k.  Column `col` starts from the unit state `col`; column M
  // (zero-state response) starts from rest and is driven by the input.
  float st[NC][MP];
  bool zsr[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = NC * li + c;
    zsr[c] = (col == p.M);
#pragma unroll
    for (int k = 0; k < MP; ++k) {
      const int comp = FORM == 0 ? MP - 1 - k : k;  // state component stored in slot k
      st[c][k] = (col < p.M && comp == col) ? 1.f : 0.f;
    }
  }
  f32x2 sp[F2 ? NC : 1][F2 ? MP / 2 : 1];  // packed ring (F2): slots 2k, 2k+1
  if constexpr (F2) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < MP / 2; ++k) sp[c][k] = pack2(st[c][2 * k], st[c][2 * k + 1]);
  }
  const float* myc = ctile + gi * TSTR;
  const float* mye = etile + gi * MP;
  // staging role of this lane: chunk sk, taps [TPL*sg, TPL*sg + TPL); its frame pair in registers
  const int sk = lane / SQ, sg = lane - sk * SQ;
  const bool stager = lane < CPW * SQ;
  float qa0[TPL], qa1[TPL];
  int qf0 = -1, qf1 = -1;
  // inputs (+ gain pair) of the (chunk, row) pairs this lane publishes, fetched one tile ahead so
  // the global-load latency hides behind the recurrence
  float pre_x[NA], pre_g0[NA], pre_g1[NA];
  auto row_time = [&](int r, int tile, bool& ok) {
    const int k = r / MP, sr = r - k * MP;
    const int ppi = w0 + k;
    const int t = time_of<FORM>(p, ppi, tile * MP + sr);
    ok = r < ROWS && ppi < nresp && t >= 0 && t < p.L;
    return t;
  };
  auto prefetch = [&](int tile) {  // raw loads only: nothing here waits on them
    if (!inb) return;
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      bool ok;
      const int t = row_time(lane + 32 * j, tile, ok);
      const Lerp w = lerp_at(ok ? t : 0, p.scale, p.F);
      pre_x[j] = ok ? __ldg(inb + t) : 0.f;
      pre_g0[j] = (FORM == 0 && gb) ? __ldg(gb + w.i0) : 1.f;
      pre_g1[j] = (FORM == 0 && gb) ? __ldg(gb + w.i1) : 1.f;
    }
  };
#pragma unroll
  for (int j = 0; j < NA; ++j) pre_x[j] = 0.f, pre_g0[j] = pre_g1[j] = 1.f;
  prefetch(0);

#pragma unroll 1
  for (int tile = 0; tile < p.Lc / MP; ++tile) {
    if (WPB > 1) __syncthreads();  // keeps the CTA's warps in the same tile of the (large) loop body
    __syncwarp();
    // ---- stage A: interpolation weights of every row + the inputs prefetched for this tile
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const int r = lane + 32 * j;
      if (r < ROWS) {
        bool ok;
        const int t = row_time(r, tile, ok);
        const Lerp w = lerp_at(ok ? t : 0, p.scale, p.F);
        float e = pre_x[j];
        if (FORM == 0 && gb) e = __fmul_rn(e, __fmaf_rn(w.l0, pre_g0[j], __fmul_rn(w.l1, pre_g1[j])));
        etile[r] = e;
        lw[r] = make_float4(ok ? w.l0 : 0.f, ok ? w.l1 : 0.f, __int_as_float(w.i0), __int_as_float(w.i1));
      }
    }
    __syncwarp();
    // ---- stage B: coefficient rows (ATen arithmetic, negated).  The frame pair is re-fetched only
    // when a row's frames differ (frame change, or ATen's floor() landing one frame low at t % hop == 0).
    if (stager) {
#pragma unroll 4
      for (int sr = 0; sr < MP; ++sr) {
        const float4 wv = lw[sk * MP + sr];
        const int i0 = __float_as_int(wv.z), i1 = __float_as_int(wv.w);
        if (i0 != qf0 || i1 != qf1) {
          qf0 = i0, qf1 = i1;
#pragma unroll
          for (int i = 0; i < TPL; ++i) {
            const bool in = TPL * sg + i < p.M;
            qa0[i] = in ? -__ldg(ab + (size_t)i0 * p.M + TPL * sg + i) : 0.f;
            qa1[i] = in ? -__ldg(ab + (size_t)i1 * p.M + TPL * sg + i) : 0.f;
          }
        }
        float* dst = ctile + sk * TSTR + sr * MP + TPL * sg;
#pragma unroll
        for (int i = 0; i < TPL; i += 2) {
          if (TPL * sg + i < MP) {  // -(l0*a0 + l1*a1) == fma(l0, -a0, l1*(-a1)): negation is exact
            float2 v;
            v.x = __fmaf_rn(wv.x, qa0[i], __fmul_rn(wv.y, qa1[i]));
            v.y = __fmaf_rn(wv.x, qa0[i + 1], __fmul_rn(wv.y, qa1[i + 1]));
            if constexpr (F2) {
              // packed layout: the row of step sr holds lag L at word (MP - (sr & 1) - L) mod MP -- descending lags, so
              // that an aligned pair of words meets an aligned pair of ring slots (see the recurrence below)
              float* row = ctile + sk * TSTR + sr * MP;
              const int t0 = TPL * sg + i;  // taps t0, t0 + 1 = lags t0 + 1, t0 + 2
              if ((sr & 1) == 0) {
                *reinterpret_cast<float2*>(row + (MP - 2 - t0)) = make_float2(v.y, v.x);
              } else {
                row[MP - 2 - t0] = v.x;
                row[(MP - 3 - t0 + MP) % MP] = v.y;  // lag MP (tap MP-1, zero: M < MP) lands on the lag-0 word
              }
            } else {
              *reinterpret_cast<float2*>(dst + i) = v;
            }
          }
        }
      }
    }
    __syncwarp();
    if (tile + 1 < p.Lc / MP) prefetch(tile + 1);
    // ---- MP recurrence steps, fully unrolled so the state rotates through registers
#pragma unroll
    for (int s = 0; s < MP; ++s) {
      const float e = mye[s];
      if constexpr (F2) {
        // Ring slots 2k, 2k+1 share a 64-bit register (sp[cc][k]).  At step s (parity par) the lags of slots 2k and
        // 2k+1 are L and L-1 with L = (s - 2k) mod MP of parity par, and the coefficient row was staged in descending
        // lag order starting at lag MP - par: word pair m = lags (MP-par-2m, MP-1-par-2m) meets ring pair
        // ((s+par)/2 + m) mod MP/2.  Lags 0 and > M carry zero coefficients.  11 or 12 packed FMAs per column and
        // step for M = 22 instead of 22 scalar ones; two partial sums (even / odd lags), added at the end.
        constexpr int HP = MP / 2;
        const int par = s & 1;
        f32x2 acc[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) acc[cc] = pack2(zsr[cc] ? e : 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < MP / 4; ++q) {
          // pairs 2q, 2q+1: lags MP-par-4q .. MP-3-par-4q; needed if any lag is in [1, MT]
          const int lag_hi = MP - par - 4 * q, lag_lo = MP - 3 - par - 4 * q;
          if (lag_lo > MT || lag_hi < 1) continue;
          const float4 v = *reinterpret_cast<const float4*>(myc + s * MP + 4 * q);
          const f32x2 c0 = pack2(v.x, v.y), c1 = pack2(v.z, v.w);
          const bool use0 = lag_hi - 1 <= MT && lag_hi >= 1, use1 = lag_lo <= MT && lag_lo + 1 >= 1;
          const int k0 = ((s + par) / 2 + 2 * q) % HP, k1 = ((s + par) / 2 + 2 * q + 1) % HP;
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) {
            if (use0) ffma2(acc[cc], c0, sp[cc][k0]);
            if (use1) ffma2(acc[cc], c1, sp[cc][k1]);
          }
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          float a0, a1, lo, hi;
          unpack2(acc[cc], a0, a1);
          unpack2(sp[cc][s / 2], lo, hi);
          const float y = a0 + a1;
          sp[cc][s / 2] = par ? pack2(lo, y) : pack2(y, hi);
        }
        continue;
      }
      float c[4 * NQ];
#pragma unroll
      for (int i4 = 0; i4 < NQ; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(myc + s * MP + 4 * i4);
        c[4 * i4] = v.x, c[4 * i4 + 1] = v.y, c[4 * i4 + 2] = v.z, c[4 * i4 + 3] = v.w;
      }
      if (FORM == 0) {
        // oldest tap first (lag L pairs coefficient c[L-1] with the output L steps back); the NC
        // columns' chains are interleaved tap by tap so the FMA pipe always has NC independent ops
        float acc[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) acc[cc] = zsr[cc] ? e : 0.f;
#pragma unroll
        for (int L = MT; L >= 1; --L)
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) acc[cc] = __fmaf_rn(c[L - 1], st[cc][(s - L + 2 * MP) % MP], acc[cc]);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) st[cc][s] = acc[cc];
      } else {
        float u[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) u[cc] = (zsr[cc] ? e : 0.f) + st[cc][s];
#pragma unroll
        for (int k = 0; k < MT - 1; ++k)
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) st[cc][(s + 1 + k) % MP] = __fmaf_rn(c[k], u[cc], st[cc][(s + 1 + k) % MP]);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) st[cc][(s + MT) % MP] = __fmul_rn(c[MT - 1], u[cc]);  // slot consumed MP-MT steps ago (or just now)
      }
    }
  }
  if constexpr (F2) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < MP / 2; ++k) unpack2(sp[c][k], st[c][2 * k], st[c][2 * k + 1]);
  }
  // ---- emit this lane's columns of [Phi | z]: W[col][k] = end-state component k
  if (chunk_on) {
    float* wb = p.W + ((size_t)b * nresp + pi) * ((MP + 1) * MP);
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) {
      const int col = NC * li + cc;
      if (col < p.M || (col == p.M && inb)) {
#pragma unroll
        for (int k4 = 0; k4 < MP / 4; ++k4) {
          float4 v;
          if (FORM == 0) {
            v = make_float4(st[cc][MP - 1 - 4 * k4], st[cc][MP - 2 - 4 * k4], st[cc][MP - 3 - 4 * k4], st[cc][MP - 4 - 4 * k4]);
          } else {  // slots >= MT hold sums that were already consumed: those components are zero
            v = make_float4(4 * k4 < MT ? st[cc][4 * k4] : 0.f, 4 * k4 + 1 < MT ? st[cc][4 * k4 + 1] : 0.f,
                            4 * k4 + 2 < MT ? st[cc][4 * k4 + 2] : 0.f, 4 * k4 + 3 < MT ? st[cc][4 * k4 + 3] : 0.f);
          }
          *reinterpret_cast<float4*>(wb + col * MP + 4 * k4) = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ pass 2 --------
// One warp per sequence.  s_{p+1} = z_p + Phi_p s_p for p = 0..C-2; S[b][p] = s_p.
// refine: delta_{p+1} = Phi_p delta_p + (E_p - S_{p+1}); S_{p+1} += delta_{p+1}.
// Each chunk's block ([Phi | z], and in refine mode the E_p and S_{p+1} rows) streams
// through a ring of shared-memory stages filled by the TMA unit (1-D cp.async.bulk,
// completion counted on an mbarrier per stage).
constexpr int kStitchStages = 4;  // ring depth (the kernel is bound by its dependent chain, not by the copies)
constexpr int kStitchGroup = 4;   // chunk blocks per stage (one mbarrier wait per group)

// MC: number of Phi columns when known at compile time (M == MC), 0: runtime p.M.
// The loop over a group is software-pipelined by hand: the Phi row of the NEXT chunk is loaded in the
// latency shadow of the state broadcast of the current one, so a step costs the dependent chain
// (state LDS -> 6-deep FMA chains -> STS -> __syncwarp) plus ~60 issue slots.
template <int MP, int MC>
__global__ void __launch_bounds__(32) ss_stitch_kernel(SsParams p, int refine) {
  constexpr int Q = (MP + 31) / 32;
  constexpr int SLOT = (MP + 1) * MP;           // floats per chunk block
  constexpr int GW = kStitchGroup * SLOT;        // W floats per stage
  constexpr int STAGE = GW + 2 * kStitchGroup * MP;  // + E rows + S rows
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x, b = blockIdx.x;
  float* ring = smem;                               // [stages][STAGE]
  float* svec = ring + kStitchStages * STAGE;       // [2][MP] state, double buffered
  uint64_t* bars = reinterpret_cast<uint64_t*>(svec + 2 * MP);
  const int nresp = p.C - 1;
  const int ngroups = (nresp + kStitchGroup - 1) / kStitchGroup;
  const int ncol = MC ? MC : p.M;
  const float* wb = p.W + (size_t)b * nresp * SLOT;
  float* sb = p.S + (size_t)b * p.C * MP;
  const float* eb = p.E + (size_t)b * p.C * MP;

  // one elected lane programs the TMA unit: group g -> stage g % kStitchStages
  auto issue = [&](int g) {
    const int stg = g % kStitchStages;
    const int first = g * kStitchGroup;
    const int n = min(kStitchGroup, nresp - first);
    float* dst = ring + stg * STAGE;
    const uint32_t wbytes = (uint32_t)(((n - 1) * SLOT + (p.M + 1) * MP) * sizeof(float));
    const uint32_t rbytes = (uint32_t)(n * MP * sizeof(float));
    mbar_expect_tx(&bars[stg], wbytes + (refine ? 2 * rbytes : 0));
    bulk_g2s(dst, wb + (size_t)first * SLOT, wbytes, &bars[stg]);
    if (refine) {
      bulk_g2s(dst + GW, eb + (size_t)first * MP, rbytes, &bars[stg]);
      bulk_g2s(dst + GW + kStitchGroup * MP, sb + (size_t)(first + 1) * MP, rbytes, &bars[stg]);
    }
  };

  if (refine) {  // adaptive: skip sequences whose chunk-boundary mismatch is already negligible
    const float mism = __uint_as_float(p.flags[2 * b]), smax = __uint_as_float(p.flags[2 * b + 1]);
    if (!(mism > p.refine_tol * smax)) return;
  } else if (lane == 0) {
    p.flags[2 * b] = 0u, p.flags[2 * b + 1] = 0u;
  }
  if (lane == 0) {
    for (int s = 0; s < kStitchStages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int k = lane + 32 * q;
    if (k < MP) {
      const float v = (!refine && p.zi && k < p.M) ? p.zi[(size_t)b * p.M + k] : 0.f;
      svec[k] = v;
      if (!refine) sb[k] = v;
    }
  }
  __syncwarp();
  if (lane == 0)
    for (int g = 0; g < kStitchStages && g < ngroups; ++g) issue(g);

  // row `k` of chunk block `blk`: Phi[k][0..ncol) and the additive term (z, or E - S when refining)
  float phi[Q][MP], base[Q], old[Q];
  auto load_row = [&](const float* grp, int c, float (&ph)[Q][MP], float (&bs)[Q], float (&od)[Q]) {
    const float* blk = grp + c * SLOT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int k = min(lane + 32 * q, MP - 1);
#pragma unroll
      for (int j = 0; j < MP; ++j) ph[q][j] = j < ncol ? blk[j * MP + k] : 0.f;
      if (refine) {
        od[q] = grp[GW + kStitchGroup * MP + c * MP + k];
        bs[q] = grp[GW + c * MP + k] - od[q];
      } else {
        od[q] = 0.f;
        bs[q] = blk[p.M * MP + k];
      }
    }
  };

  int buf = 0;
#pragma unroll 1
  for (int g = 0; g < ngroups; ++g) {
    const int stg = g % kStitchStages;
    mbar_wait(&bars[stg], (uint32_t)((g / kStitchStages) & 1));
    const float* grp = ring + stg * STAGE;
    const int n = min(kStitchGroup, nresp - g * kStitchGroup);
    load_row(grp, 0, phi, base, old);
#pragma unroll
    for (int c = 0; c < kStitchGroup; ++c) {
      if (c < n) {
        const int pi = g * kStitchGroup + c;
        // dependent part first in program order: broadcast-read the running state ...
        float sv[MP];
#pragma unroll
        for (int j4 = 0; j4 < MP / 4; ++j4) {
          const float4 v = *reinterpret_cast<const float4*>(svec + buf * MP + 4 * j4);
          sv[4 * j4] = v.x, sv[4 * j4 + 1] = v.y, sv[4 * j4 + 2] = v.z, sv[4 * j4 + 3] = v.w;
        }
        // ... and, while those loads are in flight, fetch the next chunk's row
        float nphi[Q][MP], nbase[Q], nold[Q];
        if (c + 1 < kStitchGroup && c + 1 < n) load_row(grp, c + 1, nphi, nbase, nold);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int k = lane + 32 * q;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < MP; ++j)
            if (MC == 0 || j < MC) acc[j & 3] = __fmaf_rn(phi[q][j], sv[j], acc[j & 3]);
          const float nxt = base[q] + ((acc[0] + acc[1]) + (acc[2] + acc[3]));
          if (k < MP) {
            svec[(buf ^ 1) * MP + k] = nxt;
            sb[(size_t)(pi + 1) * MP + k] = old[q] + nxt;
          }
        }
        buf ^= 1;
        __syncwarp();
        if (c + 1 < kStitchGroup && c + 1 < n) {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
#pragma unroll
            for (int j = 0; j < MP; ++j) phi[q][j] = nphi[q][j];
            base[q] = nbase[q], old[q] = nold[q];
          }
        }
      }
    }
    // the whole warp has consumed this stage: refill it with the group kStitchStages ahead
    if (lane == 0 && g + kStitchStages < ngroups) issue(g + kStitchStages);
  }
}

// ------------------------------------------------------------------ pass 3 --------
// One lane per chunk, one warp per CTA (the grid is small; spread it over every SM).
// GENERIC: coefficients fetched from global every step (any hop / chunk relation).
// !GENERIC: requires HB % MP == 0 with HB = min(hop, Lc) dividing max(hop, Lc): the frame
// pair sits in registers and is reloaded at tile starts only.
// Taps are summed oldest-first in three interleaved chains with the newest tap last, so
// consecutive steps overlap in the FMA pipe (the serial dependency is one FMA per step).
// round 0: solve from the stitched states S, write the output and the end states E;
// round 1: the same after the refinement stitch, only for the sequences that need it;
// round -1: from REST, no output -- the end state is the chunk's zero-state response z, written
//           into the z slot of W (used when pass 1 ran without the excitation).
template <int MP, int FORM, bool GENERIC>
__global__ void __launch_bounds__(32) ss_solve_kernel(SsParams p, int round) {
  constexpr int TST = MP + 1;  // tile row stride (odd -> conflict-free per-lane rows)
  constexpr int NLD = MP;      // staged elements per lane per tile (32 rows x MP / 32 lanes)
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x;
  const int G = (p.C + 31) / 32;
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  if (round == 1) {  // refinement round: only for the sequences that need it
    const float mism = __uint_as_float(p.flags[2 * b]), smax = __uint_as_float(p.flags[2 * b + 1]);
    if (!(mism > p.refine_tol * smax)) return;
  }
  const int pi = g * 32 + lane;
  const bool active = pi < p.C;
  const int pic = active ? pi : p.C - 1;
  float* tile = smem;
  float* tile2 = tile + 32 * TST;  // FORM1 only: d_ex

  const float* __restrict__ ab = p.a + (size_t)b * p.F * p.M;
  const float* __restrict__ gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  const float* __restrict__ inb = p.in + (size_t)b * p.in_stride;
  float* __restrict__ outb = p.out + (size_t)b * p.L;
  float* __restrict__ out2b = (FORM == 1 && p.out2) ? p.out2 + (size_t)b * p.L : nullptr;

  float st[MP];
  {
    const float* s0 = p.S + ((size_t)b * p.C + pic) * MP;
#pragma unroll
    for (int k = 0; k < MP; ++k) {
      const int comp = FORM == 0 ? MP - 1 - k : k;
      st[k] = (active && round >= 0) ? s0[comp] : 0.f;
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
    // ---- stage this tile's input rows (row r = chunk 32g+r): all loads first, then stores
    {
      float v[NLD];
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        const int idx = lane + 32 * j, r = idx / MP, s = idx - r * MP;
        const int prr = g * 32 + r;
        const int t = time_of<FORM>(p, prr, n0 + s);
        v[j] = (prr < p.C && t >= 0 && t < p.L) ? __ldg(inb + t) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        const int idx = lane + 32 * j, r = idx / MP, s = idx - r * MP;
        tile[r * TST + s] = v[j];
      }
    }
    __syncwarp();
    const bool reload = !GENERIC && (n0 % p.HB == 0);
    // per-tile scalars: step s of this lane is at time t0 +- s; it is a real sample iff lo_s <= s < hi_s
    const int t0 = time_of<FORM>(p, pic, n0);
    const int lo_s = FORM == 0 ? 0 : max(0, t0 - p.L + 1);
    const int hi_s = active ? (FORM == 0 ? min(MP, p.L - t0) : min(MP, t0 + 1)) : 0;
    const float tf0 = (float)max(t0, 0);
#pragma unroll
    for (int s = 0; s < MP; ++s) {
      const bool valid = s >= lo_s && s < hi_s;
      float nc[MP];
      float gv = 1.f;
      bool slow = GENERIC;
      Lerp w;
      int tc = 0;
      if (GENERIC || s == 0 || s == MP - 1) tc = valid ? (FORM == 0 ? t0 + s : t0 - s) : 0;
      if (!GENERIC) {
        if (s == 0 && reload) {  // warp-uniform: load the frame pair the coming steps live in
          kreg = min(tc / p.hop, p.F - 1);
          kregf = (float)kreg;
          const int k1 = min(kreg + 1, p.F - 1);
#pragma unroll
          for (int i = 0; i < MP; ++i) {
            na0[i] = i < p.M ? -__ldg(ab + (size_t)kreg * p.M + i) : 0.f;
            na1[i] = i < p.M ? -__ldg(ab + (size_t)k1 * p.M + i) : 0.f;
          }
          if (gb) g0 = __ldg(gb + kreg), g1 = __ldg(gb + k1);
        }
        if (s == (FORM == 0 ? 0 : MP - 1)) {
          // only here can ATen's floor(src) fall outside the register pair (t % hop == 0)
          w = lerp_at(tc, p.scale, p.F);
          slow = __any_sync(0xffffffffu, valid && w.i0 != kreg);
        }
        if (!slow) {
          // float(t) = tf0 +- s exactly (t < 2^24); ATen: src = scale * float(t)
          const float src = __fmul_rn(p.scale, FORM == 0 ? tf0 + (float)s : tf0 - (float)s);
          float l1 = __fsub_rn(src, kregf);
          l1 = fminf(fmaxf(l1, 0.f), 1.f);
          const float l0 = __fsub_rn(1.f, l1);
          const float2 l0p = f2(l0, l0), l1p = f2(l1, l1);
#pragma unroll
          for (int i = 0; i < MP; i += 2) {  // two taps per packed FMUL2/FFMA2
            const float2 c2 = __ffma2_rn(l0p, f2(na0[i], na0[i + 1]), __fmul2_rn(l1p, f2(na1[i], na1[i + 1])));
            nc[i] = c2.x, nc[i + 1] = c2.y;
          }
          if (gb) gv = __fmaf_rn(l0, g0, __fmul_rn(l1, g1));
        }
      }
      if (slow) {
        w = lerp_at(tc, p.scale, p.F);
        const float* r0 = ab + (size_t)w.i0 * p.M;
        const float* r1 = ab + (size_t)w.i1 * p.M;
#pragma unroll
        for (int i = 0; i < MP; ++i) nc[i] = i < p.M ? -lerp_apply(w, __ldg(r0 + i), __ldg(r1 + i)) : 0.f;
        if (gb) gv = lerp_apply(w, __ldg(gb + w.i0), __ldg(gb + w.i1));
      }
      // (steps outside [0, L) keep their coefficients: their input is zero, their output is
      //  never stored and no later chunk reads their end state)
      const float x = tile[lane * TST + s];
      if (FORM == 0) {
        // lag L pairs nc[L-1] with st[(s - L) mod MP]; chains over lags MP..2, newest (L=1) last
        float acc0 = valid ? (gb ? __fmul_rn(x, gv) : x) : 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
        for (int L = MP; L >= 2; --L) {
          const float h = st[(s - L + 2 * MP) % MP];
          if ((MP - L) % 3 == 0) acc0 = __fmaf_rn(nc[L - 1], h, acc0);
          if ((MP - L) % 3 == 1) acc1 = __fmaf_rn(nc[L - 1], h, acc1);
          if ((MP - L) % 3 == 2) acc2 = __fmaf_rn(nc[L - 1], h, acc2);
        }
        const float y = __fmaf_rn(nc[0], st[(s - 1 + MP) % MP], (acc0 + acc1) + acc2);
        st[s] = y;
        tile[lane * TST + s] = y;
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
    if (round < 0) continue;
    // ---- write the tile back, coalesced
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
      const int idx = lane + 32 * j, r = idx / MP, s = idx - r * MP;
      const int prr = g * 32 + r;
      const int t = time_of<FORM>(p, prr, n0 + s);
      if (prr < p.C && t >= 0 && t < p.L) {
        outb[t] = tile[r * TST + s];
        if (FORM == 1 && out2b) out2b[t] = tile2[r * TST + s];
      }
    }
  }
  if (round < 0) {  // zero-state response of this chunk -> W[b][pi][col M][:]
    if (active && pi < p.C - 1) {
      float* z = p.W + ((size_t)b * (p.C - 1) + pi) * ((MP + 1) * MP) + p.M * MP;
#pragma unroll
      for (int k4 = 0; k4 < MP / 4; ++k4) {
        float4 v;
        if (FORM == 0) {
          v = make_float4(st[MP - 1 - 4 * k4], st[MP - 2 - 4 * k4], st[MP - 3 - 4 * k4], st[MP - 4 - 4 * k4]);
        } else {
          v = make_float4(st[4 * k4], st[4 * k4 + 1], st[4 * k4 + 2], st[4 * k4 + 3]);
        }
        *reinterpret_cast<float4*>(z + 4 * k4) = v;
      }
    }
    return;
  }
  // ---- the state this chunk really ended in (input of the refinement stitch), and how far
  // it is from the stitched state of the next chunk (decides whether refinement is needed)
  if (round == 0 && p.E) {
    float mism = 0.f, smax = 0.f;
    if (active) {
      float* e0 = p.E + ((size_t)b * p.C + pi) * MP;
      const float* s1 = p.S + ((size_t)b * p.C + min(pi + 1, p.C - 1)) * MP;
#pragma unroll
      for (int k = 0; k < MP; ++k) {
        const float ev = st[FORM == 0 ? MP - 1 - k : k];
        e0[k] = ev;
        if (pi + 1 < p.C && k < p.M) {
          const float sv = s1[k];
          // NaN/inf states must force the comparison to "needs refinement"-agnostic: fmaxf drops NaN
          mism = fmaxf(mism, fabsf(ev - sv));
          smax = fmaxf(smax, fabsf(sv));
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mism = fmaxf(mism, __shfl_xor_sync(0xffffffffu, mism, d));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, d));
    }
    if (lane == 0) {  // non-negative floats order like their bit patterns
      atomicMax(p.flags + 2 * b, __float_as_uint(mism));
      atomicMax(p.flags + 2 * b + 1, __float_as_uint(smax));
    }
  }
}

// ------------------------------------------------------- pass 3, systolic form ----
// The lane-per-chunk solve above is bound by ONE warp's in-order issue: ~50 instructions per
// sample, 240 dependent samples.  Here a chunk is solved by LB = 4 neighbouring lanes instead:
// lane j owns the taps of lags TB*j+1 .. TB*j+TB (TB = MP/4) and their interpolation, so a sample
// costs every lane ~TB FMAs + TB coefficient lerps.  Only lane 0's block touches the newest
// outputs; the older blocks can be summed ahead of time, so lane j runs D*j samples AHEAD of lane 0
// and the partial sums ripple down (lane 3 -> 2 -> 1 -> 0, one __shfl_down per sample, consumed D
// iterations later) while the outputs ripple up through a delay line (lane j+1's newest history
// element is lane j's lag-(TB-D) element, one __shfl_up per sample).  Both shuffles have D
// iterations of slack, so the only serial dependency per sample is lane 0's last FMA.
//   iteration n (n = -PRE .. Lc-1, PRE = D*(LB-1)), lane j, local time tl = n + D*j:
//     ring_j[k] = y[n - (TB-D)*j - (k+1)],  k < TB          (y[m < 0] = initial state)
//     out = q_in + (j == 0 ? e[n] : 0) + sum_k c[tl][TB*j + k] * ring_j[k]       oldest k first
//     lane 0: out = y[n];  lane j > 0: out = partial sum over lags > TB*j for time tl
// Coefficient frames are (re)loaded when a lane's local time enters a new frame; the first sample
// of a frame (where ATen's floor() may land one frame low) uses coefficients computed with the
// exact reference arithmetic at load time.  FORM 0, frame-aligned chunks only.
// One warp solves the GPW = 8 chunks 8g .. 8g+7 of sequence b.  xin_all: [2][GPW*MP] floats, yout: [GPW*MP] floats of
// shared memory private to the calling warp (16-byte aligned).  Warp-level synchronisation only, so several warps of a
// CTA may run it side by side (ss_tail_kernel) or one warp per CTA (ss_solve_sys_kernel).  READ_CG: S was written by
// another CTA of this kernel (read through L2).
template <int MP, bool READ_CG>
__device__ __forceinline__ void solve_sys_body(const SsParams& p, const int b, const int g, const int round, float* xin_all,
                                               float* yout, const int lane) {
  constexpr int LB = 4, TB = MP / LB, D = 2, HOPD = TB - D, GPW = 32 / LB, PRE = D * (LB - 1);
  constexpr int NLD = (GPW * MP + 31) / 32;  // staged elements per lane per tile
  static_assert(MP % LB == 0 && HOPD >= 1 && TB % 2 == 0, "systolic solve geometry");
  const int grp = lane / LB, j = lane % LB;
  const int pi = g * GPW + grp;
  const bool active = pi < p.C;
  const int pic = active ? pi : p.C - 1;
  const float* __restrict__ ab = p.a + (size_t)b * p.F * p.M;
  const float* __restrict__ gb = p.gain ? p.gain + (size_t)b * p.F : nullptr;
  const float* __restrict__ inb = p.in + (size_t)b * p.in_stride;
  float* __restrict__ outb = p.out ? p.out + (size_t)b * p.L : nullptr;
  const int tap0 = TB * j;  // this lane's taps: tap0 .. tap0+TB-1 (tap i multiplies y[t-1-i])

  // history ring and the initial-state values lane 0 pushes during the PRE warm-up iterations
  float ring[TB], pre[PRE];
  {
    const float* s0 = p.S + ((size_t)b * p.C + pic) * MP;
    const bool ld = active && round >= 0;
#pragma unroll
    for (int k = 0; k < TB; ++k) ring[k] = ld ? (READ_CG ? __ldcg(s0 + PRE + HOPD * j + k) : s0[PRE + HOPD * j + k]) : 0.f;
#pragma unroll
    for (int i = 0; i < PRE; ++i) pre[i] = ld ? (READ_CG ? __ldcg(s0 + PRE - 1 - i) : s0[PRE - 1 - i]) : 0.f;
  }
  // coefficient frames (negated) of this lane's taps, gain pair, and the exact set for the first
  // sample of the frame
  float na0[TB], na1[TB], cg[TB];
  float g0 = 1.f, g1 = 1.f, gg = 1.f, kregf = 0.f;
  int tnext = 0;           // local time at which this lane (re)loads its frames
  int tl = -PRE + D * j;   // local time of this lane's output in the coming iteration
  const int base_t = pic * p.Lc;
#pragma unroll
  for (int i = 0; i < TB; ++i) na0[i] = na1[i] = cg[i] = 0.f;
  bool first = false;      // the coming iteration is the first sample of a freshly loaded frame
  auto reload = [&]() {
    const int t = base_t + tl;
    if (t < p.L) {
      const int kreg = min(t / p.hop, p.F - 1), k1 = min(kreg + 1, p.F - 1);
      const Lerp w = lerp_at(t, p.scale, p.F);
      const float* r0 = ab + (size_t)kreg * p.M + tap0;
      const float* r1 = ab + (size_t)k1 * p.M + tap0;
      const float* e0 = ab + (size_t)w.i0 * p.M + tap0;
      const float* e1 = ab + (size_t)w.i1 * p.M + tap0;
#pragma unroll
      for (int i = 0; i < TB; ++i) {
        const bool in = tap0 + i < p.M;
        na0[i] = in ? -__ldg(r0 + i) : 0.f;
        na1[i] = in ? -__ldg(r1 + i) : 0.f;
        cg[i] = in ? -lerp_apply(w, __ldg(e0 + i), __ldg(e1 + i)) : 0.f;
      }
      if (gb) {
        g0 = __ldg(gb + kreg), g1 = __ldg(gb + k1);
        gg = lerp_apply(w, __ldg(gb + w.i0), __ldg(gb + w.i1));
      }
      kregf = (float)kreg;
      first = true;
    }
    tnext = p.hop < p.Lc ? tnext + p.hop : 0x7fffffff;
  };
  float f0 = 0.f, f1 = 0.f;  // partial sums received from lane j+1, consumed D iterations later
  float tf = (float)(base_t + tl);  // float(t), advanced by exact +1 (t < 2^24)

  // one iteration; X = excitation sample of lane 0's time (unused by the other lanes)
  auto step = [&](float x, bool push_out, float pre_v) -> float {
    float c[TB], gv;
    if (first) {  // (divergent at most once per frame and lane)
#pragma unroll
      for (int i = 0; i < TB; ++i) c[i] = cg[i];
      gv = gg;
      first = false;
    } else {
      const float src = __fmul_rn(p.scale, tf);
      float l1 = __fsub_rn(src, kregf);
      l1 = fminf(fmaxf(l1, 0.f), 1.f);
      const float l0 = __fsub_rn(1.f, l1);
      const float2 l0p = f2(l0, l0), l1p = f2(l1, l1);
#pragma unroll
      for (int i = 0; i < TB; i += 2) {
        const float2 c2 = __ffma2_rn(l0p, f2(na0[i], na0[i + 1]), __fmul2_rn(l1p, f2(na1[i], na1[i + 1])));
        c[i] = c2.x, c[i + 1] = c2.y;
      }
      gv = __fmaf_rn(l0, g0, __fmul_rn(l1, g1));
    }
    const float q_in = (j == LB - 1) ? 0.f : f0;
    f0 = f1;
    float acc = q_in + (j == 0 ? (gb ? __fmul_rn(x, gv) : x) : 0.f);
#pragma unroll
    for (int k = TB - 1; k >= 0; --k) acc = __fmaf_rn(c[k], ring[k], acc);
    f1 = __shfl_down_sync(0xffffffffu, acc, 1);
    const float up = __shfl_up_sync(0xffffffffu, ring[HOPD - 1], 1);
    const float nv = (j == 0) ? (push_out ? acc : pre_v) : up;
#pragma unroll
    for (int k = TB - 1; k > 0; --k) ring[k] = ring[k - 1];
    ring[0] = nv;
    tf = __fadd_rn(tf, 1.f);
    ++tl;
    return acc;
  };

  // ---- stage tile 0 of the inputs
  const int ntiles = p.Lc / MP;
  float v[NLD];
  auto fetch = [&](int tile) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = lane + 32 * i, r = idx / MP, sx = idx - r * MP;
      const int prr = g * GPW + r;
      const int t = prr * p.Lc + tile * MP + sx;
      v[i] = (idx < GPW * MP && prr < p.C && t < p.L) ? __ldg(inb + t) : 0.f;
    }
  };
  auto publish = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = lane + 32 * i;
      if (idx < GPW * MP) xin_all[buf * (GPW * MP) + idx] = v[i];
    }
  };
  fetch(0);
  // ---- warm-up: the lanes ahead of lane 0 start their partial sums; lane 0 replays the initial state
#pragma unroll
  for (int i = 0; i < PRE; ++i) {
    if (tl == tnext) reload();
    step(0.f, false, pre[i]);
  }
  publish(0);
  __syncwarp();
#pragma unroll 1
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) fetch(tile + 1);
    const float* xg = xin_all + buf * (GPW * MP) + grp * MP;
#pragma unroll
    for (int sx = 0; sx < MP; ++sx) {
      // frame changes of lane j happen D*j iterations before lane 0's (which are tile aligned)
      if ((sx == 0 || sx >= MP - PRE) && tl == tnext) reload();
      const float y = step(xg[sx], true, 0.f);
      if (j == 0) yout[grp * MP + sx] = y;
    }
    __syncwarp();
    if (round >= 0) {  // write the tile back: GPW segments of MP contiguous samples
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int idx = lane + 32 * i, r = idx / MP, sx = idx - r * MP;
        const int prr = g * GPW + r;
        const int t = prr * p.Lc + tile * MP + sx;
        if (idx < GPW * MP && prr < p.C && t < p.L) outb[t] = yout[idx];
      }
    }
    if (tile + 1 < ntiles) publish(buf ^ 1);
    if (tile + 1 < ntiles) __syncwarp();
  }
  // ---- end state of each chunk = its last MP outputs (still in yout): component k = y[Lc-1-k]
  if (round < 0) {  // zero-state response -> W[b][pi][col M][:]
    if (active && pi < p.C - 1) {
      float* z = p.W + ((size_t)b * (p.C - 1) + pi) * ((MP + 1) * MP) + p.M * MP;
#pragma unroll
      for (int k = 0; k < TB; ++k) z[TB * j + k] = yout[grp * MP + MP - 1 - (TB * j + k)];
    }
    return;
  }
  if (round == 0 && p.E) {
    float mism = 0.f, smax = 0.f;
    if (active) {
      float* e0 = p.E + ((size_t)b * p.C + pi) * MP;
      const float* s1 = p.S + ((size_t)b * p.C + min(pi + 1, p.C - 1)) * MP;
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        const int comp = TB * j + k;
        const float ev = yout[grp * MP + MP - 1 - comp];
        e0[comp] = ev;
        if (pi + 1 < p.C && comp < p.M) {
          const float sv = READ_CG ? __ldcg(s1 + comp) : s1[comp];
          mism = fmaxf(mism, fabsf(ev - sv));
          smax = fmaxf(smax, fabsf(sv));
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mism = fmaxf(mism, __shfl_xor_sync(0xffffffffu, mism, d));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, d));
    }
    if (lane == 0) {  // non-negative floats order like their bit patterns
      atomicMax(p.flags + 2 * b, __float_as_uint(mism));
      atomicMax(p.flags + 2 * b + 1, __float_as_uint(smax));
    }
  }
}

template <int MP>
__global__ void __launch_bounds__(32) ss_solve_sys_kernel(SsParams p, int round) {
  constexpr int GPW = 8;
  __shared__ __align__(16) float xin[2 * GPW * MP];
  __shared__ __align__(16) float yout[GPW * MP];
  const int G = (p.C + GPW - 1) / GPW;
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  if (round == 1) {  // refinement round: only for the sequences that need it
    const float mism = __uint_as_float(p.flags[2 * b]), smax = __uint_as_float(p.flags[2 * b + 1]);
    if (!(mism > p.refine_tol * smax)) return;
  }
  solve_sys_body<MP, false>(p, b, g, round, xin, yout, (int)threadIdx.x);
}

}  // namespace golf
#include "lpc_ss_tail.cuh"
#include "lpc_ss_tc.cuh"
namespace golf {

template <int MP, int MT, int FORM>
int launch_response(const SsParams& p, cudaStream_t st) {
  if constexpr (MP == 24 && FORM == 0) {
    if (response_tc_applies(p, MP, FORM)) return launch_response_tc(p, st);
  }
  constexpr int NC = RespCfg<MP>::NC;
  constexpr int LPC = (MT + 1 + NC - 1) / NC, CPW = 32 / LPC;
  const int nresp = p.C - 1;
  const int wps = ceil_div(nresp, CPW);
  constexpr int WPB = GOLF_RESP_WPB;
  const size_t sm = (size_t)((CPW * (MP * MP + 4) + CPW * MP * 5 + 31) / 32 * 32) * sizeof(float) * WPB;
  if (sm > 220 * 1024) return GOLF_ERR_UNSUPPORTED;
  static unsigned long long attr = 0;
  if (sm > 48 * 1024 && first_use_on_device(attr)) {
    GOLF_CUDA(cudaFuncSetAttribute(ss_response_kernel<MP, MT, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    mark_used_on_device(attr);
  }
  ss_response_kernel<MP, MT, FORM><<<ceil_div(p.B * wps, WPB), 32 * WPB, sm, st>>>(p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

// passes: bit0 responses (without z when p.in is null), bit1 stitch, bit2 solve, bit3 refinement
// (stitch + solve again), bit4 zero-state responses by a solve from rest (before the stitch)
template <int MP, int FORM>
int launch_mp(const SsParams& p, bool generic, int passes, cudaStream_t st) {
  const int nresp = p.C - 1;
  // one-launch tail: two-level stitch + solve + refinement (+ room FIR) by a cluster per sequence (lpc_ss_tail.cuh)
  bool use_tail = false;
  if constexpr (FORM == 0 && MP >= 16 && MP % 8 == 0 && MP <= 32) {
    // g_ss_tail: 0 (default) never, 1 wherever it applies, 2 for batches of at most kTailAutoMaxBatch sequences.  The
    // cluster kernel shortens the serial part of the FILTER (B = 1: 150 -> 142 us, B = 32: 178 -> 163 us) but holds whole
    // SMs for it (166 registers x 256 threads, 100 KB of shared memory per CTA, 4 CTAs per sequence) and its fused room
    // FIR runs on those 4 SMs only: with several passes in flight at B = 32 it costs throughput (bench.py `value`
    // 6.05e9 vs 7.10e9 samples/s with the light stitch / solve launches) and the whole decoder is no faster even at
    // B = 1 (250 vs 246 us).  It stays an opt-in schedule (profiles/README.md, round 2).
    const bool tail_wanted = g_ss_tail == 1 || (g_ss_tail == 2 && p.B <= kTailAutoMaxBatch);
    use_tail = !generic && g_solve_systolic && tail_wanted && (passes & 6) == 6 && p.Gw != nullptr;
  }
  if (p.room_k && !use_tail) return GOLF_ERR_UNSUPPORTED;  // the fused room FIR exists only in the tail kernel
  if (nresp > 0 && (passes & 1)) {
    constexpr int MTS = MP >= 8 ? MP - 2 : MP;  // short-tap variant for M <= MP - 2
    int rc;
    if constexpr (MP == 40) {  // order 25..32 at hops that 32 does not divide (240, 120) runs at padded order 40: 32 taps, not 38
      if (p.M <= 32)
        rc = launch_response<MP, 32, FORM>(p, st);
      else if (p.M <= MTS)
        rc = launch_response<MP, MTS, FORM>(p, st);
      else
        rc = launch_response<MP, MP, FORM>(p, st);
    } else if (p.M <= MTS) {
      rc = launch_response<MP, MTS, FORM>(p, st);
    } else {
      rc = launch_response<MP, MP, FORM>(p, st);
    }
    if (rc) return rc;
  }
  if constexpr (FORM == 0 && MP >= 16 && MP % 8 == 0 && MP <= 32) {
    if (use_tail) return launch_tail<MP>(p, passes, st);
  }
  const size_t sm_stitch =
      ((size_t)kStitchStages * kStitchGroup * ((MP + 1) * MP + 2 * MP) + 2 * MP) * sizeof(float) + kStitchStages * 8 + 128;
  constexpr int MCS = MP >= 8 ? MP - 2 : 0;  // compile-time column counts: M == MP - 2 and M == MP
  constexpr int MC8 = MP == 40 ? 32 : 0;     // ... and order 32 at padded order 40 (the run-time-M walk is 6x slower: 174 vs 27 us)
  const int mc = p.M == MP ? MP : (MCS && p.M == MCS ? MCS : (MC8 && p.M == MC8 ? MC8 : 0));
  static unsigned long long attr2 = 0;
  if (first_use_on_device(attr2)) {
    GOLF_CUDA(cudaFuncSetAttribute(ss_stitch_kernel<MP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_stitch));
    GOLF_CUDA(cudaFuncSetAttribute(ss_stitch_kernel<MP, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_stitch));
    if (MCS) GOLF_CUDA(cudaFuncSetAttribute(ss_stitch_kernel<MP, (MCS ? MCS : MP)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_stitch));
    if (MC8) GOLF_CUDA(cudaFuncSetAttribute(ss_stitch_kernel<MP, (MC8 ? MC8 : MP)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_stitch));
    mark_used_on_device(attr2);
  }
  const int G = ceil_div(p.C, 32);
  const size_t sm_solve = (FORM == 0 ? 1 : 2) * 32 * (MP + 1) * sizeof(float);
  auto solve = [&](int round) -> int {
    if constexpr (FORM == 0 && MP >= 16 && MP % 8 == 0) {
      if (!generic && g_solve_systolic) {
        ss_solve_sys_kernel<MP><<<p.B * ceil_div(p.C, 8), 32, 0, st>>>(p, round);
        GOLF_CHECK_LAUNCH();
        return GOLF_OK;
      }
    }
    if (generic)
      ss_solve_kernel<MP, FORM, true><<<p.B * G, 32, sm_solve, st>>>(p, round);
    else
      ss_solve_kernel<MP, FORM, false><<<p.B * G, 32, sm_solve, st>>>(p, round);
    GOLF_CHECK_LAUNCH();
    return GOLF_OK;
  };
  if (nresp > 0 && (passes & 16)) {
    const int rc = solve(-1);
    if (rc) return rc;
  }
  for (int round = 0; round < 2; ++round) {
    const bool refine = round == 1;
    if (refine && !((passes & 8) && nresp > 0)) break;
    if (refine || (passes & 2)) {
      if (mc == MP)
        ss_stitch_kernel<MP, MP><<<p.B, 32, sm_stitch, st>>>(p, refine ? 1 : 0);
      else if (MC8 && mc == MC8)
        ss_stitch_kernel<MP, (MC8 ? MC8 : MP)><<<p.B, 32, sm_stitch, st>>>(p, refine ? 1 : 0);
      else if (mc != 0)
        ss_stitch_kernel<MP, (MCS ? MCS : MP)><<<p.B, 32, sm_stitch, st>>>(p, refine ? 1 : 0);
      else
        ss_stitch_kernel<MP, 0><<<p.B, 32, sm_stitch, st>>>(p, refine ? 1 : 0);
      GOLF_CHECK_LAUNCH();
    }
    if (refine || (passes & 4)) {
      const int rc = solve(round);
      if (rc) return rc;
    }
  }
  return GOLF_OK;
}

}  // namespace golf
