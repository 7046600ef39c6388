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
//   1  compose: the C-1 chunk blocks are cut into NG <= 32 groups of G; warp g multiplies its group's affine maps,
//      [Phi_grp | z_grp] = prod_k [Phi_k | z_k]  (G-1 products of MP x MP by MP x (M+1), operands streamed from L2)
//   2  one warp walks the NG groups:  s_{g+1} = z_grp + Phi_grp s_g                        (depth NG instead of C)
//   3  expand: warp g walks the G chunks of its group from s_g and writes every chunk's entry state S
//   4  solve: warp w re-runs chunks 8w .. 8w+7 (systolic, 4 lanes per chunk: solve_sys_body), writes y and the state
//      each chunk ended in (E), and reduces max|E_p - S_{p+1}| / max|S| for the sequence
//   5  if the mismatch exceeds the tolerance (a cluster-uniform decision): the same two-level walk on
//      delta_{p+1} = Phi_p delta_p + (E_p - S_{p+1}), S += delta, and the solve once more
//   6  room FIR over the sequence's y (register-tiled correlation, fir_tile.cuh), 256 outputs per warp and tile
//
// Depth of the serial part: (G-1) matrix products + NG + G matrix-vector steps (6 + 29 + 7 at C = 199) against 198.
#pragma once
#include <cooperative_groups.h>

#include "fir_tile.cuh"

namespace golf {
namespace cg = cooperative_groups;

constexpr int kTailCtas = 8;                      // CTAs per cluster (one cluster per sequence; 8 is the portable maximum)
constexpr int kTailWarps = 4;                     // warps per CTA
constexpr int kTailNW = kTailCtas * kTailWarps;   // warps per sequence
constexpr int kRoomTile = 256;                    // room-FIR outputs per warp and tile (8 per lane)
constexpr int kRoomMaxTaps = 252;                 // learned taps supported by the fused room FIR (K12 <= 264)

template <int MP>
struct TailCfg {
  static constexpr int SLOT = (MP + 1) * MP;
  static constexpr int kRoomStrip = kRoomTile + 264 + 20;                       // logical strip length (max taps)
  static constexpr int kRoomStripSw = kRoomStrip + 4 * (kRoomStrip / 32) + 8;   // fir_sw() layout
  static constexpr int kNeed = 2 * SLOT > kRoomStripSw ? 2 * SLOT : kRoomStripSw;  // compose needs 2 blocks; solve 3*8*MP < 2*SLOT
  static constexpr int kWarpFloats = (kNeed + 31) / 32 * 32;
  static_assert(3 * 8 * MP <= 2 * SLOT, "solve staging must fit the per-warp shared memory");
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

// ---- phase 1: [Phi_grp | z_grp] of group g.  pb: [2][SLOT] floats of this warp's shared memory.
template <int MP>
__device__ __forceinline__ void tail_compose(const SsParams& p, int b, int g, float* pb, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  const int nresp = p.C - 1;
  const int first = g * p.G, last = min(first + p.G, nresp);
  if (first >= last) return;
  const int M = p.M, ncopy = (M + 1) * MP;
  const float* __restrict__ wb = p.W + (size_t)b * nresp * SLOT;
  const int r = min(lane, MP - 1);
  {  // running product starts as the first block (columns 0..M-1 and z are contiguous)
    const float* __restrict__ src = wb + (size_t)first * SLOT;
    for (int i = lane; i < ncopy; i += 32) pb[i] = __ldcg(src + i);
  }
  int cur = 0;
  float row[MP], zk = 0.f;
  if (first + 1 < last) tail_load_row<MP>(wb + (size_t)(first + 1) * SLOT, r, M, row, zk);
  __syncwarp();
#pragma unroll 1
  for (int k = first + 1; k < last; ++k) {
    float nrow[MP], nz = 0.f;
    if (k + 1 < last) tail_load_row<MP>(wb + (size_t)(k + 1) * SLOT, r, M, nrow, nz);  // next block's row rides under the product
    const float* __restrict__ src = pb + cur * SLOT;
    float* __restrict__ dst = pb + (cur ^ 1) * SLOT;
#pragma unroll 1
    for (int c = 0; c <= M; ++c) {  // P'[:, c] = Phi_k P[:, c]   (+ z_k for the affine column c == M)
      float v = tail_dot<MP>(row, src + c * MP);
      if (c == M) v += zk;
      if (lane < MP) dst[c * MP + lane] = v;
    }
    __syncwarp();
    cur ^= 1;
    if (k + 1 < last) {
#pragma unroll
      for (int j = 0; j < MP; ++j) row[j] = nrow[j];
      zk = nz;
    }
  }
  float* __restrict__ gb = p.Gw + ((size_t)b * p.NG + g) * SLOT;
  const float* __restrict__ res = pb + cur * SLOT;
  for (int i = lane; i < ncopy; i += 32) gb[i] = res[i];
}

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
  float row[MP], z = 0.f;
  if (p.NG > 1) tail_load_row<MP>(gw, r, p.M, row, z);
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g + 1 < p.NG; ++g) {
    float nrow[MP], nz = 0.f;
    if (g + 2 < p.NG) tail_load_row<MP>(gw + (size_t)(g + 1) * SLOT, r, p.M, nrow, nz);
    // refine: the additive term is the mismatch accumulated over group g (phase 5A), replaced in place by the
    // correction ENTERING group g
    float add = z;
    if (refine) {
      add = __ldcg(sg + (size_t)g * MP + r);
    }
    if (on) sg[(size_t)g * MP + r] = s;
    const float nxt = add + tail_dot<MP>(row, sv + cur * MP);
    if (on) sv[(cur ^ 1) * MP + r] = nxt;
    s = nxt;
    cur ^= 1;
    __syncwarp();
    if (g + 2 < p.NG) {
#pragma unroll
      for (int j = 0; j < MP; ++j) row[j] = nrow[j];
      z = nz;
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
  float row[MP], z = 0.f, e1 = 0.f, s1 = 0.f;
  auto fetch = [&](int k, float (&rw)[MP], float& zz, float& ee, float& ss) {
    tail_load_row<MP>(wb + (size_t)k * SLOT, r, p.M, rw, zz);
    if (mode != 0) {
      ee = __ldcg(Eb + (size_t)k * MP + r);
      ss = __ldcg(Sb + (size_t)(k + 1) * MP + r);
    }
  };
  if (first < stop) fetch(first, row, z, e1, s1);
  __syncwarp();
#pragma unroll 1
  for (int k = first; k < stop; ++k) {
    float nrow[MP], nz = 0.f, ne = 0.f, ns = 0.f;
    if (k + 1 < stop) fetch(k + 1, nrow, nz, ne, ns);
    const float add = mode == 0 ? z : e1 - s1;
    const float nxt = add + tail_dot<MP>(row, sv + cur * MP);
    if (on) {
      sv[(cur ^ 1) * MP + r] = nxt;
      const bool mine = (k + 1 < last) || owns_end;
      if (mode == 0 && mine) Sb[(size_t)(k + 1) * MP + r] = nxt;
      if (mode == 2 && mine) Sb[(size_t)(k + 1) * MP + r] = s1 + nxt;
    }
    s = nxt;
    cur ^= 1;
    __syncwarp();
    if (k + 1 < stop) {
#pragma unroll
      for (int j = 0; j < MP; ++j) row[j] = nrow[j];
      z = nz, e1 = ne, s1 = ns;
    }
  }
  if (mode == 1 && on) p.Dg[((size_t)b * p.NG + g) * MP + r] = s;
}

// ---- phase 6: out[t] = y[t] + sum_{j<n} k[j] y[t-n+j] over one tile of kRoomTile outputs.
// ks: [K12] taps staged as [k_0 .. k_{n-1}, 1, 0 ...] (shared by the CTA); xs: this warp's strip (fir_sw layout).
__device__ __forceinline__ void tail_room_tile(const float* __restrict__ yb, float* __restrict__ ob, int L, int n, int K12, int t0,
                                               const float* __restrict__ ks, float* __restrict__ xs, int lane) {
  const int xs_len = kRoomTile + K12 + 20;
  constexpr int U = 4;
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

// passes: bit1|bit2 (stitch + solve, always), bit3 refinement allowed, bit4 zero-state responses by a solve from rest
template <int MP>
__global__ void __cluster_dims__(kTailCtas, 1, 1) __launch_bounds__(32 * kTailWarps) ss_tail_kernel(SsParams p, int passes) {
  constexpr int WF = TailCfg<MP>::kWarpFloats;
  __shared__ __align__(16) float smem[kTailWarps * WF];
  __shared__ __align__(16) float room_taps[264];
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / kTailCtas;
  const int crank = blockIdx.x % kTailCtas;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = crank * kTailWarps + warp;  // warp index within the sequence's cluster
  float* wsm = smem + warp * WF;
  const int nsolve = (p.C + 7) / 8;  // groups of 8 chunks for the systolic solve
  const int K12 = p.room_k ? (p.room_n + 1 + 11) / 12 * 12 : 0;
  if (p.room_k) {
    for (int i = threadIdx.x; i < K12; i += blockDim.x) room_taps[i] = i < p.room_n ? p.room_k[i] : (i == p.room_n ? 1.f : 0.f);
  }

  if (passes & 16) {  // ---- phase 0: zero-state responses into the z column of W
    for (int g = cw; g < nsolve; g += kTailNW) solve_sys_body<MP, true>(p, b, g, -1, wsm, wsm + 2 * 8 * MP, lane);
    __threadfence();
    cluster.sync();
  }
  // ---- phase 1: compose the groups
  if (cw < p.NG) tail_compose<MP>(p, b, cw, wsm, lane);
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
  if (cw < p.NG) tail_walk_chunks<MP>(p, b, cw, 0, wsm, lane);
  __threadfence();
  cluster.sync();
  // ---- phase 4: solve
  for (int g = cw; g < nsolve; g += kTailNW) solve_sys_body<MP, true>(p, b, g, 0, wsm, wsm + 2 * 8 * MP, lane);
  __threadfence();
  cluster.sync();
  // ---- phase 5: refinement, decided once for the sequence (every thread of the cluster reads the same two words)
  const float mism = __uint_as_float(__ldcg(p.flags + 2 * b)), smax = __uint_as_float(__ldcg(p.flags + 2 * b + 1));
  const bool refine = (passes & 8) && p.C > 1 && (mism > p.refine_tol * smax);
  if (refine) {
    if (cw < p.NG) tail_walk_chunks<MP>(p, b, cw, 1, wsm, lane);
    __threadfence();
    cluster.sync();
    if (cw == 0) tail_walk_groups<MP>(p, b, 1, wsm, lane);
    __threadfence();
    cluster.sync();
    if (cw < p.NG) tail_walk_chunks<MP>(p, b, cw, 2, wsm, lane);
    __threadfence();
    cluster.sync();
    for (int g = cw; g < nsolve; g += kTailNW) solve_sys_body<MP, true>(p, b, g, 1, wsm, wsm + 2 * 8 * MP, lane);
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

// host: group geometry for C chunks
static inline void tail_groups(int C, int* NG, int* G) {
  const int nresp = C - 1;
  const int g = nresp > 0 ? (nresp + kTailNW - 1) / kTailNW : 1;
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
