// lpc_ss_tail.cuh -- GOLF-ss passes 2..4 (+ the room FIR) as ONE launch: a thread-block cluster per sequence.
//
// After the chunk-response pass the filter still has to (2) stitch the chunk states, s_{p+1} = z_p + Phi_p s_p,
// (3) re-run every chunk from its state, (4) if the states the chunks really ended in disagree with the stitched
// ones, propagate the mismatch and solve again, and the decoder then applies the learned room FIR
// (models/filters.py:443-450).  As separate kernels that is a serial tail: a 198-step dependent walk on ONE warp
// per sequence (38 us), the solve (28 us), two launches that only find out that no refinement is needed (9 us), the
// room FIR (17 us) and five launch gaps -- ~110 us during which a 148-SM part is nearly idle (ncu: IPC 0.03 / 0.5).
//
// Here a cluster of 4 CTAs x 8 warps owns one sequence (one CTA per SM: the warps of the latency-bound phases must not
// share a scheduler) and moves through the phases together, separated by cluster barriers (barrier.cluster,
// release/acquire at cluster scope).  Data crossing CTAs goes through L2; the chunk / group blocks a walk consumes
// stream into a per-warp shared-memory ring through the TMA unit (cp.async.bulk + mbarrier), three blocks ahead:
//
//   0  (optional) zero-state responses by a solve from rest          [when pass 1 ran without the excitation]
//   1  compose: the C-1 chunk blocks are cut into NG <= 32 groups of G; a warp per group multiplies the group's
//      affine maps, [Phi_grp | z_grp] = prod_k [Phi_k | z_k]  (G-1 products of MP x MP by MP x (M+1), four columns =
//      16 FMA chains in flight)
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
#include <cstdio>

#include "fir_tile.cuh"
#include "lpc_ss_solve_tr.cuh"

namespace golf {
namespace cg = cooperative_groups;

constexpr int kTailCtas = 4;                      // CTAs per cluster (one cluster per sequence)
constexpr int kTailWarps = 8;                     // warps per CTA
constexpr int kTailNW = kTailCtas * kTailWarps;   // warps per sequence
constexpr int kRoomTile = 256;                    // room-FIR outputs per warp and tile (8 per lane)
constexpr int kRoomMaxTaps = 252;                 // learned taps supported by the fused room FIR (K12 <= 264)

// per-warp shared memory (floats):  ring[NS][STG] | pbuf[2][SLOT] | sv[2][MP] | mbarriers
// (the solve's staging and the room FIR's strip alias the ring)
// TMA ring depth of a walking / composing warp.  A bulk copy takes ~2 000 cycles from issue to completion (measured: three
// blocks in flight gave 640 cycles per matrix-vector step), so the ring is as deep as the shared memory of a CTA that
// owns its SM allows: 7 stages up to MP = 24 (186 KB per CTA), 3 at MP = 32.
template <int MP>
struct TailCfg {
  static constexpr int NS = MP <= 24 ? 7 : 3;
  static constexpr int SLOT = (MP + 1) * MP;
  static constexpr int STG = SLOT + 2 * MP;                                     // a block + an E row + an S row
  static constexpr int kRing = NS * STG;
  static constexpr int kSolve = 4 * 8 * MP;                                     // xin[2][8*MP] + yout[8*MP] + entry states[8*MP]
  static constexpr int kRoomStrip = kRoomTile + 264 + 20;                       // logical strip length (max taps)
  static constexpr int kRoomStripSw = kRoomStrip + 4 * (kRoomStrip / 32) + 8;   // fir_sw() layout
  static constexpr int kPbuf = kRing, kSv = kPbuf + 2 * SLOT, kBars = kSv + 2 * MP;
  static constexpr int kWarpFloats = (kBars + 2 * NS + 31) / 32 * 32;
  static_assert(kSolve <= kRing && kRoomStripSw <= kRing && (kBars % 2) == 0 && (STG % 4) == 0 && (SLOT % 4) == 0, "tail smem layout");
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// One warp's stream of blocks through the TMA unit.  Items are numbered n = 0, 1, 2, ... over the whole kernel;
// item n completes on mbarrier n % NS with parity (n / NS) & 1.  issue() is called by lane 0 only.
template <int MP>
struct TailStream {
  float* ring;
  uint64_t* bars;
  int n_issue, n_wait;
  __device__ __forceinline__ void init(float* wsm, int lane) {
    ring = wsm;
    bars = reinterpret_cast<uint64_t*>(wsm + TailCfg<MP>::kBars);
    n_issue = n_wait = 0;
    if (lane == 0) {
      for (int i = 0; i < TailCfg<MP>::NS; ++i) mbar_init(&bars[i], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ float* stage(int n) const { return ring + (n % TailCfg<MP>::NS) * TailCfg<MP>::STG; }
  // block (bytes) -> dst; optional rows e_src / s_src (MP floats each) -> the stage's E / S slots
  __device__ __forceinline__ void issue(float* dst, const float* blk, uint32_t bytes, const float* e_src, const float* s_src) {
    uint64_t* bar = &bars[n_issue % TailCfg<MP>::NS];
    float* stg = stage(n_issue);
    constexpr uint32_t rb = MP * sizeof(float);
    mbar_expect_tx(bar, bytes + (e_src ? rb : 0u) + (s_src ? rb : 0u));
    bulk_g2s(dst, blk, bytes, bar);
    if (e_src) bulk_g2s(stg + TailCfg<MP>::SLOT, e_src, rb, bar);
    if (s_src) bulk_g2s(stg + TailCfg<MP>::SLOT + MP, s_src, rb, bar);
    ++n_issue;
  }
  __device__ __forceinline__ void wait() {  // all lanes
    mbar_wait(&bars[n_wait % TailCfg<MP>::NS], (uint32_t)((n_wait / TailCfg<MP>::NS) & 1));
    ++n_wait;
  }
};

// row r of a block [Phi | z] in shared memory, column-major (column j at j*MP): conflict-free across lanes
template <int MP>
__device__ __forceinline__ void tail_row_smem(const float* __restrict__ blk, int r, int M, float (&row)[MP], float& z) {
#pragma unroll
  for (int j = 0; j < MP; ++j) row[j] = j < M ? blk[j * MP + r] : 0.f;
  z = blk[M * MP + r];
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

// ---- phase 1: [Phi_grp | z_grp] of group g.  The running product lives in pbuf (double-buffered); the blocks it is
// multiplied by stream through the ring.  Four columns are in flight at a time: 16 independent FMA chains hide the FMA
// latency a lone warp would otherwise wait out.
template <int MP>
__device__ __forceinline__ void tail_compose(const SsParams& p, int b, int g, TailStream<MP>& ts, float* wsm, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  const int nresp = p.C - 1;
  const int first = g * p.G, last = min(first + p.G, nresp);
  if (first >= last) return;
  const int M = p.M, ncols = M + 1;  // column M is the affine term z
  const uint32_t bytes = (uint32_t)(ncols * MP * sizeof(float));
  const float* __restrict__ wb = p.W + (size_t)b * nresp * SLOT;
  float* pbuf = wsm + TailCfg<MP>::kPbuf;
  const int r = min(lane, MP - 1);
  const int K = last - first;
  if (lane == 0) {
    fence_proxy_async_all();
    ts.issue(pbuf, wb + (size_t)first * SLOT, bytes, nullptr, nullptr);  // the product starts as the first block
    for (int i = 1; i < TailCfg<MP>::NS && i < K; ++i) ts.issue(ts.stage(ts.n_issue), wb + (size_t)(first + i) * SLOT, bytes, nullptr, nullptr);
  } else {
    ts.n_issue += min(TailCfg<MP>::NS, K);
  }
  ts.wait();
  int cur = 0;
#pragma unroll 1
  for (int i = 1; i < K; ++i) {
    const float* stg = ts.stage(ts.n_wait);
    ts.wait();
    float row[MP], zk;
    tail_row_smem<MP>(stg, r, M, row, zk);
    const float* __restrict__ src = pbuf + cur * SLOT;
    float* __restrict__ dst = pbuf + (cur ^ 1) * SLOT;
#pragma unroll 1
    for (int c = 0; c < ncols; c += 4) {  // P'[:, c] = Phi_k P[:, c]   (+ z_k for the affine column)
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = tail_dot<MP>(row, src + min(c + u, ncols - 1) * MP);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c + u == M) v[u] += zk;
        if (lane < MP && c + u < ncols) dst[(c + u) * MP + lane] = v[u];
      }
    }
    __syncwarp();  // the stage has been read by every lane, P' is complete
    cur ^= 1;
    if (i + TailCfg<MP>::NS - 1 < K) {
      if (lane == 0) ts.issue(ts.stage(ts.n_issue), wb + (size_t)(first + i + TailCfg<MP>::NS - 1) * SLOT, bytes, nullptr, nullptr);
      else ++ts.n_issue;
    }
  }
  float* __restrict__ gb = p.Gw + ((size_t)b * p.NG + g) * SLOT;
  const float* __restrict__ res = pbuf + cur * SLOT;
  for (int i = lane; i < ncols * MP; i += 32) gb[i] = res[i];
}

// ---- phases 2 / 5B: walk the groups.  refine == 0: states (from zi or rest); 1: corrections (from zero).
template <int MP>
__device__ __forceinline__ void tail_walk_groups(const SsParams& p, int b, int refine, TailStream<MP>& ts, float* wsm, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  float* sv = wsm + TailCfg<MP>::kSv;
  const int r = min(lane, MP - 1);
  const bool on = lane < MP;
  float* __restrict__ sg = (refine ? p.Dg : p.Sg) + (size_t)b * p.NG * MP;
  const float* __restrict__ gw = p.Gw + (size_t)b * p.NG * SLOT;
  const uint32_t bytes = (uint32_t)((p.M + 1) * MP * sizeof(float));
  float s = 0.f;
  if (!refine && p.zi && r < p.M) s = p.zi[(size_t)b * p.M + r];
  if (on) sv[r] = s;
  int cur = 0;
  const int K = p.NG - 1;  // step g: s_{g+1} = add_g + Phi_grp_g s_g
  // refine: the additive term is the mismatch accumulated over group g (phase 5A); it rides in the stage's E slot and is
  // replaced in global memory by the correction ENTERING group g when the step runs
  auto issue = [&](int g) { ts.issue(ts.stage(ts.n_issue), gw + (size_t)g * SLOT, bytes, refine ? sg + (size_t)g * MP : nullptr, nullptr); };
  if (lane == 0) {
    fence_proxy_async_all();
    for (int i = 0; i < TailCfg<MP>::NS && i < K; ++i) issue(i);
  } else {
    ts.n_issue += min(TailCfg<MP>::NS, K);
  }
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < K; ++g) {
    const float* stg = ts.stage(ts.n_wait);
    ts.wait();
    float row[MP], z;
    tail_row_smem<MP>(stg, r, p.M, row, z);
    const float add = refine ? stg[SLOT + r] : z;
    if (on) sg[(size_t)g * MP + r] = s;
    const float nxt = add + tail_dot<MP>(row, sv + cur * MP);
    if (on) sv[(cur ^ 1) * MP + r] = nxt;
    s = nxt;
    cur ^= 1;
    __syncwarp();
    if (g + TailCfg<MP>::NS < K) {
      if (lane == 0) issue(g + TailCfg<MP>::NS);
      else ++ts.n_issue;
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
__device__ __forceinline__ void tail_walk_chunks(const SsParams& p, int b, int g, int mode, TailStream<MP>& ts, float* wsm, int lane) {
  constexpr int SLOT = TailCfg<MP>::SLOT;
  float* sv = wsm + TailCfg<MP>::kSv;
  const int nresp = p.C - 1;
  const int first = g * p.G, last = min(first + p.G, nresp);
  if (first > nresp || (first == nresp && g > 0)) return;
  const int r = min(lane, MP - 1);
  const bool on = lane < MP;
  const float* __restrict__ wb = p.W + (size_t)b * nresp * SLOT;
  float* __restrict__ Sb = p.S + (size_t)b * p.C * MP;
  const float* __restrict__ Eb = p.E + (size_t)b * p.C * MP;
  const uint32_t bytes = (uint32_t)((p.M + 1) * MP * sizeof(float));
  const bool owns_end = last == nresp;  // the last group also owns the entry state of the final chunk
  float s = 0.f;
  if (mode == 0) s = __ldcg(p.Sg + ((size_t)b * p.NG + g) * MP + r);
  if (mode == 2) s = __ldcg(p.Dg + ((size_t)b * p.NG + g) * MP + r);
  // mode 2: the step into the next group's first chunk is that group's business (and it may already have
  // corrected the S it would read)
  const int stop = (mode == 2 && !owns_end) ? last - 1 : last;
  const int K = max(stop - first, 0);
  auto issue = [&](int k) {
    ts.issue(ts.stage(ts.n_issue), wb + (size_t)k * SLOT, bytes, mode ? Eb + (size_t)k * MP : nullptr, mode ? Sb + (size_t)(k + 1) * MP : nullptr);
  };
  if (lane == 0) {
    fence_proxy_async_all();
    for (int i = 0; i < TailCfg<MP>::NS && i < K; ++i) issue(first + i);
  } else {
    ts.n_issue += min(TailCfg<MP>::NS, K);
  }
  if (on) {
    sv[r] = s;
    if (mode == 0) Sb[(size_t)first * MP + r] = s;
    if (mode == 2 && g > 0) Sb[(size_t)first * MP + r] = __ldcg(Sb + (size_t)first * MP + r) + s;
  }
  int cur = 0;
  __syncwarp();
#pragma unroll 1
  for (int k = first; k < stop; ++k) {
    const float* stg = ts.stage(ts.n_wait);
    ts.wait();
    float row[MP], z;
    tail_row_smem<MP>(stg, r, p.M, row, z);
    const float s1 = mode ? stg[SLOT + MP + r] : 0.f;
    const float add = mode == 0 ? z : stg[SLOT + r] - s1;
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
    if (k + TailCfg<MP>::NS < stop) {
      if (lane == 0) issue(k + TailCfg<MP>::NS);
      else ++ts.n_issue;
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
      const float raw = yb[min(max(pos, 0), L - 1)];  // (plain load: y was written before the last cluster barrier)
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

// -DGOLF_TAIL_TIMING (instrumented builds only, tools/): cluster 0 prints the duration of every phase
#ifdef GOLF_TAIL_TIMING
#define TAIL_MARK(i)                                                            \
  do {                                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                  \
      unsigned long long t_;                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
      tail_marks[i] = t_;                                                       \
      tail_clk[i] = clock64();                                                  \
    }                                                                           \
  } while (0)
#else
#define TAIL_MARK(i)
#endif

// every thread: make this thread's global writes visible to the cluster (generic and async proxies), then barrier
#define TAIL_SYNC()          \
  do {                       \
    __threadfence();         \
    fence_proxy_async_all(); \
    cluster.sync();          \
  } while (0)

// passes: bit1|bit2 (stitch + solve, always), bit3 refinement allowed, bit4 zero-state responses by a solve from rest
template <int MP>
__global__ void __cluster_dims__(kTailCtas, 1, 1) __launch_bounds__(32 * kTailWarps, 1) ss_tail_kernel(SsParams p, int passes) {
  constexpr int WF = TailCfg<MP>::kWarpFloats;
  extern __shared__ __align__(128) float smem[];  // [kTailWarps][WF] | room taps [264]
  float* room_taps = smem + kTailWarps * WF;
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
  // tasks are dealt to the cluster's warps CTA by CTA (task w -> CTA w % 4), so that no SM hosts more of these
  // latency-bound warps than it must
  const int sw = warp * kTailCtas + crank;
  TailStream<MP> ts;
  ts.init(wsm, lane);
#ifdef GOLF_TAIL_TIMING
  __shared__ unsigned long long tail_marks[8];
  __shared__ long long tail_clk[8];
#endif
  TAIL_MARK(0);

  if (passes & 16) {  // ---- phase 0: zero-state responses into the z column of W
    for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, -1, wsm, lane);
    TAIL_SYNC();
  }
  // ---- phase 1: compose the groups
  if (sw < p.NG) tail_compose<MP>(p, b, sw, ts, wsm, lane);
  TAIL_SYNC();
  TAIL_MARK(1);
  // ---- phase 2: walk the groups
  if (cw == 0) {
    if (lane == 0) p.flags[2 * b] = 0u, p.flags[2 * b + 1] = 0u;
    tail_walk_groups<MP>(p, b, 0, ts, wsm, lane);
  }
  TAIL_SYNC();
  TAIL_MARK(2);
  // ---- phase 3: entry state of every chunk
  if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 0, ts, wsm, lane);
  TAIL_SYNC();
  TAIL_MARK(3);
  // ---- phases 4 / 5: solve; if the states the chunks ended in disagree with the stitched ones (a decision every
  // thread of the cluster takes from the same two words), propagate the mismatch and solve again
  for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, 0, wsm, lane);
  TAIL_SYNC();
  TAIL_MARK(4);
  const float mism = __uint_as_float(__ldcg(p.flags + 2 * b)), smax = __uint_as_float(__ldcg(p.flags + 2 * b + 1));
  const bool refine = (passes & 8) && p.C > 1 && (mism > p.refine_tol * smax);
  if (refine) {
    if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 1, ts, wsm, lane);
    TAIL_SYNC();
    if (cw == 0) tail_walk_groups<MP>(p, b, 1, ts, wsm, lane);
    TAIL_SYNC();
    if (sw < p.NG) tail_walk_chunks<MP>(p, b, sw, 2, ts, wsm, lane);
    TAIL_SYNC();
    for (int g = sw; g < nsolve; g += kTailNW) tail_solve<MP>(p, b, g, 1, wsm, lane);
    TAIL_SYNC();
  }
  TAIL_MARK(5);
  // ---- phase 6: room FIR
  if (p.room_k) {
    __syncthreads();  // room_taps
    const float* __restrict__ yb = p.out + (size_t)b * p.L;
    float* __restrict__ ob = p.room_out + (size_t)b * p.L;
    for (int t0 = cw * kRoomTile; t0 < p.L; t0 += kTailNW * kRoomTile) tail_room_tile(yb, ob, p.L, p.room_n, K12, t0, room_taps, wsm, lane);
  }
#ifdef GOLF_TAIL_TIMING
  cluster.sync();
  TAIL_MARK(6);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("tail phases (ns): compose %llu  groups %llu  expand %llu  solve %llu  refine %llu  room %llu  total %llu\n",
           tail_marks[1] - tail_marks[0], tail_marks[2] - tail_marks[1], tail_marks[3] - tail_marks[2], tail_marks[4] - tail_marks[3],
           tail_marks[5] - tail_marks[4], tail_marks[6] - tail_marks[5], tail_marks[6] - tail_marks[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("tail phases (clk): compose %lld  groups %lld  expand %lld  solve %lld  refine %lld  room %lld  total %lld\n",
           tail_clk[1] - tail_clk[0], tail_clk[2] - tail_clk[1], tail_clk[3] - tail_clk[2], tail_clk[4] - tail_clk[3],
           tail_clk[5] - tail_clk[4], tail_clk[6] - tail_clk[5], tail_clk[6] - tail_clk[0]);
#endif
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
  const size_t sm = ((size_t)kTailWarps * TailCfg<MP>::kWarpFloats + 264) * sizeof(float);
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    GOLF_CUDA(cudaFuncSetAttribute(ss_tail_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    mark_used_on_device(attr);
  }
  ss_tail_kernel<MP><<<p.B * kTailCtas, 32 * kTailWarps, sm, st>>>(p, passes);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

}  // namespace golf
