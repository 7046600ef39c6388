// lpc_ss_solve_tr.cuh -- GOLF-ss pass 3 in TRANSPOSED systolic form.
//
// The direct-form solve (solve_sys_body) spends ~230 cycles per sample on ONE dependent chain per step: interpolate
// the coefficient row of time t, then a serial 6-FMA accumulate, then a shuffle (ncu: 68 instructions per step at
// 0.3 IPC, stall reason `wait`).  Here every product is instead pushed FORWARD to the sample it contributes to:
//
//     y[m] known  ->  q[t] += c_k[t] * y[m]   for t = m+1+k, k < M        (c_k[t] = -a_up[t, k], the row of time t)
//     y[t] = e[t] + q[t]                       once all M contributions have arrived
//
// so the only dependency from one sample to the next is  y[m] -> one FMA -> one ADD -> y[m+1];  the M products of a
// step are independent of each other and the coefficient rows (which depend on t alone) are computed off that path.
// Four lanes share a chunk (lane j owns taps TB*j .. TB*j+TB-1, TB = MP/4) as a pipeline: lane j runs D = 2 steps
// behind lane j-1, receives y[m] from it by shuffle, keeps TB pending sums (for t = m+1+TB*j .. m+TB+TB*j) and hands
// each completed partial sum down to lane j-1, which merges it into its own pending sum of the same t; both shuffles
// are consumed two iterations after they were issued.  Interpolation weights of the destination times slide through a
// TB-deep register window (one new weight pair per step, ATen arithmetic as everywhere else).
//
// A chunk must lie inside one control frame (Lc divides hop); its first MP iterations replay the entry state
// (S[i] = y[t0-1-i]) through the pipeline so that the pending sums of t >= t0 hold the contributions of the past.
// The first sample of a frame (where ATen's floor() may land one frame low) uses coefficients computed with the exact
// reference arithmetic.  Same interface and epilogue as solve_sys_body.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace golf {

// xin_all: [2][8*MP], yout: [8*MP], ssm: [8*MP] floats of shared memory private to the calling warp
template <int MP, bool READ_CG>
__device__ __forceinline__ void solve_tr_body(const SsParams& p, const int b, const int g, const int round, float* xin_all,
                                              float* yout, float* ssm, const int lane) {
  constexpr int LB = 4, TB = MP / LB, D = 2, GPW = 32 / LB, KS = TB - D - 2;
  constexpr int NLD = (GPW * MP + 31) / 32;  // staged elements per lane per tile
  static_assert(MP % LB == 0 && KS >= 0, "transposed solve geometry");
  // (the caller may be a non-inlined function holding `p` in local memory: read each field once)
  const int C = p.C, L = p.L, Lc = p.Lc, F = p.F, M = p.M, hop = p.hop;
  const float scale = p.scale;
  const int grp = lane / LB, j = lane % LB;
  const int pi = g * GPW + grp;
  const bool active = pi < C;
  const int pic = active ? pi : C - 1;
  const float* __restrict__ ab = p.a + (size_t)b * F * M;
  const float* __restrict__ gb = p.gain ? p.gain + (size_t)b * F : nullptr;
  const float* __restrict__ inb = p.in + (size_t)b * p.in_stride;
  float* __restrict__ outb = p.out ? p.out + (size_t)b * L : nullptr;
  const int tap0 = TB * j;
  const int t0 = pic * Lc;

  // ---- entry states of the 8 chunks -> shared memory (zero: from rest)
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    const int idx = lane + 32 * i, r = idx / MP, c = idx - r * MP;
    const int prr = g * GPW + r;
    if (idx < GPW * MP) {
      float v = 0.f;
      if (round >= 0 && prr < C) {
        const float* sp = p.S + ((size_t)b * C + prr) * MP + c;
        v = READ_CG ? __ldcg(sp) : *sp;
      }
      ssm[idx] = v;
    }
  }
  // ---- this lane's taps of the chunk's frame pair (negated), the exact set for a frame's first sample, gain pair
  const int kreg = min(t0 / hop, F - 1), k1 = min(kreg + 1, F - 1);
  const float kregf = (float)kreg;
  const bool first_exact = (t0 % hop) == 0;  // dest time t0 is the first sample of frame kreg
  float na0[TB], na1[TB], cg[TB];
  {
    const Lerp w = lerp_at(min(t0, L - 1), scale, F);
    const float* r0 = ab + (size_t)kreg * M + tap0;
    const float* r1 = ab + (size_t)k1 * M + tap0;
    const float* e0 = ab + (size_t)w.i0 * M + tap0;
    const float* e1 = ab + (size_t)w.i1 * M + tap0;
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      const bool in = tap0 + i < M;
      na0[i] = in ? -__ldg(r0 + i) : 0.f;
      na1[i] = in ? -__ldg(r1 + i) : 0.f;
      cg[i] = in ? -lerp_apply(w, __ldg(e0 + i), __ldg(e1 + i)) : 0.f;
    }
  }
  // interpolation weights of time tf (ATen: src = scale * t; l1 = clamp(src - i0); l0 = 1 - l1), frame kreg
  auto weights = [&](float tf, float& l0, float& l1) {
    const float src = __fmul_rn(scale, tf);
    float v = __fsub_rn(src, kregf);
    v = fminf(fmaxf(v, 0.f), 1.f);
    l1 = v;
    l0 = __fsub_rn(1.f, v);
  };
  // pending sums and the weights of their destination times.  Iteration i: this lane processes source m = i - D*j;
  // logical slot k' (physical register (k' + s) % TB at unrolled step s) is destination m + 1 + TB*j + k'.
  float q[TB], wl0[TB], wl1[TB];
  float tf_top;  // float(t0 + destination of the slot that becomes the top one at the NEXT iteration)
  {
    const int m0 = -MP - D * j;  // source of the first iteration (i = -MP)
#pragma unroll
    for (int k = 0; k < TB; ++k) {
      q[k] = 0.f;
      // logical k at step s = 0 is physical k; the top slot (k = TB-1) gets its weights inside the first iteration
      weights((float)(t0 + m0 + 1 + tap0 + k), wl0[k], wl1[k]);
    }
    tf_top = (float)(t0 + m0 + tap0 + TB);
  }
  float done_last = 0.f, pin_cur = 0.f, y_last = 0.f, yin_cur = 0.f, done_prev = 0.f;

  const float* smine = ssm + grp * MP;

  // TB iterations starting at iteration i0 (the slot rotation has period TB, so this is the smallest body with static
  // register indices; a 24-iteration body is 19 KB of straight-line code and starves on instruction fetch).
  // EARLY: sources with m < 0 come from the entry state and the destination t0 may need the exact first-sample
  // coefficients; eg: this lane-group's excitation samples e[i0 ..] (null while i < 0); yo: where lane 0 stores y.
  auto subtile = [&](auto early_tag, const int i0, const float* __restrict__ eg, float* __restrict__ yo) {
    constexpr bool EARLY = decltype(early_tag)::value;
#pragma unroll
    for (int s = 0; s < TB; ++s) {
      // ---- source sample of this lane.  lane 0: y[i] = e[i] + (sum completed last iteration)
      const float ycomp = (eg ? eg[s] : 0.f) + done_prev;
      float ysrc = (j == 0) ? ycomp : yin_cur;
      if (EARLY) {
        const int m = i0 + s - D * j;
        if (m < 0) {
          const int idx = -1 - m;
          ysrc = idx < MP ? smine[min(idx, MP - 1)] : 0.f;
        }
      }
      if (j == 0 && yo) yo[s] = ycomp;
      // ---- partial sum handed down by lane j+1 joins the pending sum of the same destination
      q[(KS + s) % TB] = __fadd_rn(q[(KS + s) % TB], pin_cur);
      // ---- the slot that enters at the top: weights of its destination time
      weights(tf_top, wl0[(TB - 1 + s) % TB], wl1[(TB - 1 + s) % TB]);
      tf_top = __fadd_rn(tf_top, 1.f);
      // ---- push y[m] to the TB destinations of this lane's taps
      float done = 0.f;
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        const int ph = (k + s) % TB;
        float c = __fmaf_rn(wl0[ph], na0[k], __fmul_rn(wl1[ph], na1[k]));
        if (EARLY && first_exact) {
          const int dest = i0 + s - D * j + 1 + tap0 + k;
          if (dest == 0) c = cg[k];
        }
        if (k == 0)
          done = __fmaf_rn(c, ysrc, q[ph]);
        else if (k == TB - 1)
          q[ph] = __fmul_rn(c, ysrc);
        else
          q[ph] = __fmaf_rn(c, ysrc, q[ph]);
      }
      // ---- hand-offs (both consumed two iterations after the value was produced)
      const float pin_n = __shfl_down_sync(0xffffffffu, done_last, 1);
      const float yin_n = __shfl_up_sync(0xffffffffu, y_last, 1);
      pin_cur = (j == LB - 1) ? 0.f : pin_n;
      yin_cur = yin_n;
      done_last = done;
      y_last = ysrc;
      done_prev = done;
    }
  };
  using True = std::true_type;
  using False = std::false_type;

  // ---- input staging: tile tl covers iterations tl*MP .. tl*MP+MP-1 of the 8 chunks; the excitation e = x * up(gain)
  // is formed here, in parallel over the lanes, with the reference's interpolation arithmetic (lerp_at)
  const int ntiles = Lc / MP;
  // (raw loads only in fetch(): nothing waits on them until publish(), one tile later)
  float vx[NLD], vg0[NLD], vg1[NLD];
  auto fetch = [&](int tl) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = lane + 32 * i, r = idx / MP, sx = idx - r * MP;
      const int prr = g * GPW + r;
      const int t = prr * Lc + tl * MP + sx;
      const bool ok = idx < GPW * MP && prr < C && t < L;
      vx[i] = ok ? __ldg(inb + t) : 0.f;
      if (gb) {
        const Lerp w = lerp_at(ok ? t : 0, scale, F);
        vg0[i] = __ldg(gb + w.i0), vg1[i] = __ldg(gb + w.i1);
      }
    }
  };
  auto publish = [&](int buf, int tl) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = lane + 32 * i, r = idx / MP, sx = idx - r * MP;
      float e = vx[i];
      if (gb) {
        const int t = (g * GPW + r) * Lc + tl * MP + sx;
        const Lerp w = lerp_at(t < L ? t : 0, scale, F);
        e = __fmul_rn(e, lerp_apply(w, vg0[i], vg1[i]));
      }
      if (idx < GPW * MP) xin_all[buf * (GPW * MP) + idx] = e;
    }
  };
  fetch(0);
  __syncwarp();  // ssm visible
#pragma unroll 1
  for (int u = 0; u < MP / TB; ++u) subtile(True{}, -MP + u * TB, nullptr, nullptr);  // replay the entry state
  publish(0, 0);
  __syncwarp();
#pragma unroll 1
  for (int tl = 0; tl < ntiles; ++tl) {
    const int buf = tl & 1;
    if (tl + 1 < ntiles) fetch(tl + 1);
    const float* eg = xin_all + buf * (GPW * MP) + grp * MP;
    float* yo = yout + grp * MP;
    if (tl == 0) {
#pragma unroll 1
      for (int u = 0; u < MP / TB; ++u) subtile(True{}, u * TB, eg + u * TB, yo + u * TB);
    } else {
#pragma unroll 1
      for (int u = 0; u < MP / TB; ++u) subtile(False{}, tl * MP + u * TB, eg + u * TB, yo + u * TB);
    }
    __syncwarp();
    if (round >= 0) {  // write the tile back: GPW segments of MP contiguous samples
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int idx = lane + 32 * i, r = idx / MP, sx = idx - r * MP;
        const int prr = g * GPW + r;
        const int t = prr * Lc + tl * MP + sx;
        if (idx < GPW * MP && prr < C && t < L) outb[t] = yout[idx];
      }
    }
    if (tl + 1 < ntiles) publish(buf ^ 1, tl + 1);
    if (tl + 1 < ntiles) __syncwarp();
  }
  // ---- end state of each chunk = its last MP outputs (still in yout): component k = y[Lc-1-k]
  if (round < 0) {  // zero-state response -> W[b][pi][col M][:]
    if (active && pi < C - 1) {
      float* z = p.W + ((size_t)b * (C - 1) + pi) * ((MP + 1) * MP) + M * MP;
#pragma unroll
      for (int k = 0; k < TB; ++k) z[TB * j + k] = yout[grp * MP + MP - 1 - (TB * j + k)];
    }
    return;
  }
  if (round == 0 && p.E) {
    float mism = 0.f, smax = 0.f;
    if (active) {
      float* e0 = p.E + ((size_t)b * C + pi) * MP;
      const float* s1 = p.S + ((size_t)b * C + min(pi + 1, C - 1)) * MP;
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        const int comp = TB * j + k;
        const float ev = yout[grp * MP + MP - 1 - comp];
        e0[comp] = ev;
        if (pi + 1 < C && comp < M) {
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

}  // namespace golf
