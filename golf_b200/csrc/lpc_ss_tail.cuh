// lpc_ss_tail.cuh -- GOLF-ss passes 2..4 (+ the room FIR) as ONE launch: a thread-block cluster per sequence.
//
// After the chunk-response pass the filter still has to (2) stitch the chunk states, s_{p+1} = z_p + Phi_p s_p,
// (3) re-run every chunk from its state, (4) if the states the chunks really ended in disagree with the stitched
// ones, propagate the mismatch and solve again, and the decoder then applies the learned room FIR
// (models/filters.py:443-450).  As separate kernels that is a serial tail: a 198-step dependent walk on ONE warp
// per sequence (38 us), the solve (28 us), two launches that only find out that no refinement is needed (9 us), the
// room FIR (17 us) and five launch gaps -- ~110 us during which a 148-SM part is nearly idle (ncu: IPC 0.03 / 0.5).
//
// Here the 8 CTAs x 4 warps of a cluster own one sequence and move through the phases together, separated by
// cluster barriers (barrier.cluster, release/acquire at cluster scope); data crossing CTAs goes through L2
// (__ldcg loads of S / E / group blocks / y -- never a stale L1 line):
//
//   0  (optional) zero-state responses by a solve from rest          [when pass 1 ran without the excitation]
//   1  compose: the C-1 chunk blocks are cut into NG <= 32 groups of G; a warp per group multiplies the group's
//      affine maps, [Phi_grp | z_grp] = prod_k [Phi_k | z_k]  (G-1 products of MP x MP by
//      MP x (M+1), operands streamed from L2, four columns = 16 FMA chains in flight)
//   2  one warp walks the NG groups:  s_{g+1} = z_grp + Phi_grp s_g                        (depth NG instead of C)
//   3  expand: a warp per group walks its G chunks from s_g and writes every chunk's entry state S
//   4  solve: a warp re-runs 8 chunks (4 lanes per chunk; transposed pipeline, lpc_ss_solve_tr.cuh), writes y and the
//      state each chunk ended in (E), and reduces max|E_p - S_{p+1}| / max|S| for the sequence
//   5  if the mismatch exceeds the tolerance (a cluster-uniform decision): the same two-level walk on
//      delta_{p+1} = Phi_p delta_p + (E_p - S_{p+1}), S += delta, and the solve once more
//   6  room FIR over the sequence's y (register-tiled correlation, fir_tile.cuh), 256 outputs per warp and tile
//
// Depth of the serial part: (G-1) matrix products + NG + G matrix-vector steps (6 + 29 + 7 at C = 199) against 198.
#pragma once
#include <cooperative_groups.h>

#include "fir_tile.cuh"
#include "lpc_ss_solve_tr.cuh"

namespace golf {
namespace cg = cooperative_groups;

constexpr int kTailCtas = 8;                      // CTAs per cluster (one cluster per sequence; 8 is the portable maximum)
constexpr int kTailWarps = 4;                     // warps per CTA (160 registers: three CTAs per SM)
constexpr int kTailNW = kTailCtas * kTailWarps;   // warps per sequence
constexpr int kRoomTile = 256;                    // room-FIR outputs per warp and tile (8 per lane)
constexpr int kRoomMaxTaps = 252;                 // learned taps supported by the fused room FIR (K12 <= 264)

template <int MP>
struct TailCfg {
  static constexpr int SLOT = (MP + 1) * MP;
  static constexpr int CH = MP + 1;                                             // columns of [Phi | z] per compose warp (all of them: one warp per group)
  static constexpr int kCompose = 2 * CH * MP;                                  // the running product, double-buffered
  static constexpr int kSolve = 4 * 8 * MP;                                     // xin[2][8*MP] + yout[8*MP] + entry states[8*MP]
  static constexpr int kRoomStrip = kRoomTile + 264 + 20;                       // logical strip length (max taps)
  static constexpr int kRoomStripSw = kRoomStrip + 4 * (kRoomStrip / 32) + 8;   // fir_sw() layout
  static constexpr int kMax2 = kCompose > kSolve ? kCompose : kSolve;
  static constexpr int kNeed = kMax2 > kRoomStripSw ? kMax2 : kRoomStripSw;
  static constexpr int kWarpFloats = (kNeed + 31) / 32 * 32;
};

// row r of a chunk (or group) block [Phi | z] stored column-major (column j at j*MP): MP coefficients + the additive term
template <int MP>
__device__ __forceinline__ void tail_load_row(const float* __restrict__ blk, int r, int M, float (&row)[MP], float& z) {
#pragma unroll
  for (int j = 0; j < MP; ++j) row[j] = j < M ? __ldcg(blk + j * MP + r) : 0.f;
  z = __ldcg(blk + M * MP + r);
}

// sum_j row[j] * sv[j]  (sv: shared memory, broadcast reads), four interleaved chains
template <int MP>
__device__ __forceinline__ float tail_dot(const float (&row)[MP], const float* __restrict__ sv) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j4 = 0; j4 < MP / 4; ++j4) {
    const float4 v = *reinterpret_cast<const float4*>(sv + 4 * j4);
    acc[0] = __fmaf_rn(row[4 * j4], v.x, acc[0]);
    acc[1] = __fmaf_rn(row[4 * j4 + 1], v.y, acc[1]);
    acc[2] = __fmaf_rn(row[4 * j4 + 2], v.z, acc[2]);
    acc[3] = __fmaf_rn(row[4 * j4 + 3], v.w, acc[3]);
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ---- phase 1: [Phi_grp | z_grp] of group g; this warp owns columns [h*CH, h*CH + CH) of it (the columns of a matrix
// product are independent, so a group could be split over several warps; with CH = MP+1 one warp takes them all).
// pb: [2][CH*MP] floats of this warp's shared memory.  Four columns are in flight at a time: 16 independent FMA chains
// hide the FMA latency a lone warp would otherwise wait out.
template <int MP>
__device__ __forceinline__ void tail_compose(const SsParams& p, int b, int g, int h, float* pb, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT, CH = TailCfg<MP>::CH, HALF = CH * MP;
  const int nresp = p.C - 1;
  const int first = g * p.G, last = min(first + p.G, nresp);
  if (first >= last) return;
  const int M = p.M;
  const int c0 = h * CH, ncols = min(CH, M + 1 - c0);  // columns c0 .. c0+ncols-1 (column M is the affine term z)
  if (ncols <= 0) return;
  const float* __restrict__ wb = p.W + (size_t)b * nresp * SLOT;
  const int r = min(lane, MP - 1);
  {  // running product starts as the first block (columns are contiguous, z follows column M-1)
    const float* __restrict__ src = wb + (size_t)first * SLOT + c0 * MP;
    for (int i = lane; i < ncols * MP; i += 32) pb[i] = __ldcg(src + i);
  }
  int cur = 0;
  float row[MP], zk = 0.f;
  if (first + 1 < last) tail_load_row<MP>(wb + (size_t)(first + 1) * SLOT, r, M, row, zk);
  __syncwarp();
#pragma unroll 1
  for (int k = first + 1; k < last; ++k) {
    float nrow[MP], nz = 0.f;
    if (k + 1 < last) tail_load_row<MP>(wb + (size_t)(k + 1) * SLOT, r, M, nrow, nz);  // next block's row rides under the product
    const float* __restrict__ src = pb + cur * HALF;
    float* __restrict__ dst = pb + (cur ^ 1) * HALF;
#pragma unroll 1
    for (int c = 0; c < ncols; c += 4) {  // P'[:, c] = Phi_k P[:, c]   (+ z_k for the affine column)
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = tail_dot<MP>(row, src + min(c + u, ncols - 1) * MP);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c0 + c + u == M) v[u] += zk;
        if (lane < MP && c + u < ncols) dst[(c + u) * MP + lane] = v[u];
      }
    }
    __syncwarp();
    cur ^= 1;
    if (k + 1 < last) {
#pragma unroll
      for (int j = 0; j < MP; ++j) row[j] = nrow[j];
      zk = nz;
    }
  }
  float* __restrict__ gb = p.Gw + ((size_t)b * p.NG + g) * SLOT + c0 * MP;
  const float* __restrict__ res = pb + cur * HALF;
  for (int i = lane; i < ncols * MP; i += 32) gb[i] = res[i];
}

// ---- one matrix-vector step of a walk: rows stream from L2 through two register sets (the row of step k+2 is
// requested as soon as step k has consumed its set)
template <int MP>
struct TailRow {
  float r[MP];
  float z, e, s;
};

// ---- phases 2 / 5B: walk the groups.  refine == 0: states (from zi or rest); 1: corrections (from zero).
// sv: [2][MP] floats of shared memory.
template <int MP>
__device__ __forceinline__ void tail_walk_groups(const SsParams& p, int b, int refine, float* sv, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  const int r = min(lane, MP - 1);
  const bool on = lane < MP;
  float* __restrict__ sg = (refine ? p.Dg : p.Sg) + (size_t)b * p.NG * MP;
  const float* __restrict__ gw = p.Gw + (size_t)b * p.NG * SLOT;
  float s = 0.f;
  if (!refine && p.zi && r < p.M) s = p.zi[(size_t)b * p.M + r];
  if (on) sv[r] = s;
  int cur = 0;
  const int nstep = p.NG - 1;  // step g: s_{g+1} = add_g + Phi_grp_g s_g
  auto fetch = [&](int g, TailRow<MP>& R) {
    tail_load_row<MP>(gw + (size_t)g * SLOT, r, p.M, R.r, R.z);
    // refine: the additive term is the mismatch accumulated over group g (phase 5A); it is replaced in place by the
    // correction ENTERING group g when the step runs (same lane, same address: read here, written there)
    if (refine) R.z = __ldcg(sg + (size_t)g * MP + r);
  };
  auto step = [&](int g, const TailRow<MP>& R) {
    if (on) sg[(size_t)g * MP + r] = s;
    const float nxt = R.z + tail_dot<MP>(R.r, sv + cur * MP);
    if (on) sv[(cur ^ 1) * MP + r] = nxt;
    s = nxt;
    cur ^= 1;
    __syncwarp();
  };
  TailRow<MP> R0, R1;
  if (0 < nstep) fetch(0, R0);
  if (1 < nstep) fetch(1, R1);
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < nstep; g += 2) {
    step(g, R0);
    if (g + 2 < nstep) fetch(g + 2, R0);
    if (g + 1 < nstep) {
      step(g + 1, R1);
      if (g + 3 < nstep) fetch(g + 3, R1);
    }
  }
  if (on) sg[(size_t)(p.NG - 1) * MP + r] = s;
}

// ---- phases 3 / 5A / 5C: walk the chunks of group g.
//   mode 0 (expand):      s from Sg[g];  S[first] = s;  s <- Phi_k s + z_k;  S[k+1] = s
//   mode 1 (accumulate):  d = 0;  d <- Phi_k d + (E_k - S_{k+1});  Dg[g] = d
//   mode 2 (correct):     d from Dg[g];  S[first] += d;  d <- Phi_k d + (E_k - S_{k+1});  S[k+1] += d
// A group never writes the entry state of the next group's first chunk (that group does).
template <int MP>
__device__ __forceinline__ void tail_walk_chunks(const SsParams& p, int b, int g, int mode, float* sv, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  const int nresp = p.C - 1;
  const int first = g * p.G, last = min(first + p.G, nresp);
  if (first > nresp || (first == nresp && g > 0)) return;
  const int r = min(lane, MP - 1);
  const bool on = lane < MP;
  const float* __restrict__ wb = p.W + (size_t)b * nresp * SLOT;
  float* __restrict__ Sb = p.S + (size_t)b * p.C * MP;
  const float* __restrict__ Eb = p.E + (size_t)b * p.C * MP;
  const bool owns_end = last == nresp;  // the last group also owns the entry state of the final chunk
  float s = 0.f;
  if (mode == 0) s = __ldcg(p.Sg + ((size_t)b * p.NG + g) * MP + r);
  if (mode == 2) s = __ldcg(p.Dg + ((size_t)b * p.NG + g) * MP + r);
  if (on) {
    sv[r] = s;
    if (mode == 0) Sb[(size_t)first * MP + r] = s;
    if (mode == 2 && g > 0) Sb[(size_t)first * MP + r] = __ldcg(Sb + (size_t)first * MP + r) + s;
  }
  // mode 2: the step into the next group's first chunk is that group's business (and it may already have
  // corrected the S it would read)
  const int stop = (mode == 2 && !owns_end) ? last - 1 : last;
  int cur = 0;
  auto fetch = [&](int k, TailRow<MP>& R) {
    tail_load_row<MP>(wb + (size_t)k * SLOT, r, p.M, R.r, R.z);
    if (mode != 0) {
      R.e = __ldcg(Eb + (size_t)k * MP + r);
      R.s = __ldcg(Sb + (size_t)(k + 1) * MP + r);
    }
  };
  auto step = [&](int k, const TailRow<MP>& R) {
    const float add = mode == 0 ? R.z : R.e - R.s;
    const float nxt = add + tail_dot<MP>(R.r, sv + cur * MP);
    if (on) {
      sv[(cur ^ 1) * MP + r] = nxt;
      const bool mine = (k + 1 < last) || owns_end;
      if (mode == 0 && mine) Sb[(size_t)(k + 1) * MP + r] = nxt;
      if (mode == 2 && mine) Sb[(size_t)(k + 1) * MP + r] = R.s + nxt;
    }
    s = nxt;
    cur ^= 1;
    __syncwarp();
  };
  TailRow<MP> R0, R1;
  if (first < stop) fetch(first, R0);
  if (first + 1 < stop) fetch(first + 1, R1);
  __syncwarp();
#pragma unroll 1
  for (int k = first; k < stop; k += 2) {
    step(k, R0);
    if (k + 2 < stop) fetch(k + 2, R0);
    if (k + 1 < stop) {
      step(k + 1, R1);
      if (k + 3 < stop) fetch(k + 3, R1);
    }
  }
  if (mode == 1 && on) p.Dg[((size_t)b * p.NG + g) * MP + r] = s;
}

// ---- phase 6: out[t] = y[t] + sum_{j<n} k[j] y[t-n+j] over one tile of kRoomTile outputs.
// ks: [K12] taps staged as [k_0 .. k_{n-1}, 1, 0 ...] (shared by the CTA); xs: this warp's strip (fir_sw layout).
__device__ __forceinline__ void tail_room_tile(const float* __restrict__ yb, float* __restrict__ ob, int L, int n, int K12, int t0,
                                               const float* __restrict__ ks, float* __restrict__ xs, int lane) {
  const int xs_len = kRoomTile + K12 + 20;
  constexpr int U = 7;  // 256 + 132 + 20 = 408 elements: two batches of 7 loads per lane
#pragma unroll 1
  for (int i0 = lane; i0 < xs_len; i0 += 32 * U) {
    float v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int pos = t0 - n + i0 + 32 * q;
      const float raw = __ldcg(yb + min(max(pos, 0), L - 1));
      v[q] = (pos >= 0 && pos < L) ? raw : 0.f;
    }
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (i0 + 32 * q < xs_len) xs[fir_sw(i0 + 32 * q)] = v[q];
  }
  __syncwarp();
  const int r0 = lane * kR;
  float acc[kR];
#pragma unroll
  for (int i = 0; i < kR; ++i) acc[i] = 0.f;
  fir_tile8_sw(xs, r0, ks, K12, acc);
#pragma unroll
  for (int i = 0; i < kR; ++i)
    if (t0 + r0 + i < L) ob[t0 + r0 + i] = acc[i];
  __syncwarp();
}

// the solve of 8 chunks by one warp: transposed pipeline when every chunk lies inside one control frame, else the
// direct-form one.  Not inlined: the tail kernel calls it from three phases.
template <int MP>
__device__ __noinline__ void tail_solve(const SsParams& p, int b, int g, int round, float* wsm, int lane) {
  if (p.hop % p.Lc == 0)
    solve_tr_body<MP, true>(p, b, g, round, wsm, wsm + 2 * 8 * MP, wsm + 3 * 8 * MP, lane);
  else
    solve_sys_body<MP, true>(p, b, g, round, wsm, wsm + 2 * 8 * MP, lane);
}

// passes: bit1|bit2 (stitch + solve, always), bit3 refinement allowed, bit4 zero-state responses by a solve from rest
template <int MP>
__global__ void __cluster_dims__(kTailCtas, 1, 1) __launch_bounds__(32 * kTailWarps, 3) ss_tail_kernel(SsParams p, int passes) {
  constexpr int WF = TailCfg<MP>::kWarpFloats;
  __shared__ __align__(16) float smem[kTailWarps * WF];
  __shared__ __align__(16) float room_taps[264];
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / kTailCtas;
  const int crank = blockIdx.x % kTailCtas;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = crank * kTailWarps + warp;  // warp index within the sequence's cluster
  float* wsm = smem + warp * WF;
  const int nsolve = (p.C + 7) / 8;  // groups of 8 chunks for the solve
  const int K12 = p.room_k ? (p.room_n + 1 + 11) / 12 * 12 : 0;
  if (p.room_k) {
    for (int i = threadIdx.x; i < K12; i += blockDim.x) room_taps[i] = i < p.room_n ? p.room_k[i] : (i == p.room_n ? 1.f : 0.f);
  }
  // the solve warps are spread over the cluster's CTAs (warp w of the sequence -> CTA w % 8) so that no SM hosts more
  // of these latency-bound warps than it must
  const int sw = warp * kTailCtas + crank;

  if (passes & 16) {  // ---- phase 0: zero-state responses into the z column of W
    for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, -1, wsm, lane);
    __threadfence();
    cluster.sync();
  }
  // ---- phase 1: compose the groups
  if (sw < p.NG) tail_compose<MP>(p, b, sw, 0, wsm, lane);
  __threadfence();
  cluster.sync();
  // ---- phase 2: walk the groups
  if (cw == 0) {
    if (lane == 0) p.flags[2 * b] = 0u, p.flags[2 * b + 1] = 0u;
    tail_walk_groups<MP>(p, b, 0, wsm, lane);
  }
  __threadfence();
  cluster.sync();
  // ---- phase 3: entry state of every chunk
  if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 0, wsm, lane);
  __threadfence();
  cluster.sync();
  // ---- phases 4 / 5: solve; if the states the chunks ended in disagree with the stitched ones (a decision every
  // thread of the cluster takes from the same two words), propagate the mismatch and solve again
  for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, 0, wsm, lane);
  __threadfence();
  cluster.sync();
  const float mism = __uint_as_float(__ldcg(p.flags + 2 * b)), smax = __uint_as_float(__ldcg(p.flags + 2 * b + 1));
  const bool refine = (passes & 8) && p.C > 1 && (mism > p.refine_tol * smax);
  if (refine) {
    if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 1, wsm, lane);
    __threadfence();
    cluster.sync();
    if (cw == 0) tail_walk_groups<MP>(p, b, 1, wsm, lane);
    __threadfence();
    cluster.sync();
    if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 2, wsm, lane);
    __threadfence();
    cluster.sync();
    for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, 1, wsm, lane);
    __threadfence();
    cluster.sync();
  }
  // ---- phase 6: room FIR
  if (p.room_k) {
    __syncthreads();  // room_taps
    const float* __restrict__ yb = p.out + (size_t)b * p.L;
    float* __restrict__ ob = p.room_out + (size_t)b * p.L;
    for (int t0 = cw * kRoomTile; t0 < p.L; t0 += kTailNW * kRoomTile) tail_room_tile(yb, ob, p.L, p.room_n, K12, t0, room_taps, wsm, lane);
  }
}

// host: group geometry for C chunks (at most one group per warp of the cluster)
constexpr int kTailMaxGroups = kTailNW;
static inline void tail_groups(int C, int* NG, int* G) {
  const int nresp = C - 1;
  const int g = nresp > 0 ? (nresp + kTailMaxGroups - 1) / kTailMaxGroups : 1;
  *G = g;
  *NG = nresp > 0 ? (nresp + g - 1) / g : 1;
}

template <int MP>
int launch_tail(const SsParams& p, int passes, cudaStream_t st) {
  if (p.room_k && (p.room_n < 1 || p.room_n > kRoomMaxTaps)) return GOLF_ERR_UNSUPPORTED;
  ss_tail_kernel<MP><<<p.B * kTailCtas, 32 * kTailWarps, 0, st>>>(p, passes);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

}  // namespace golf
