// synth_fused.cu -- one entry point for the whole GOLF-ss decoder pass (models/sf.py:47-64 with the modules of
// cfg/ae/decoder/golf-precise.yaml):
//
//   phase, w  --> glottal wavetable oscillator (knot-prefix scan + flow/decimation kernel)        -> harm   [workspace]
//   noise | rng, log_mag --> FIR design + block FIR + harm   (fir_fused.cu)                       -> src    [workspace]
//   src, gain, a --> GOLF-ss chunk responses, then stitch + solve + refinement + room FIR (cluster kernel) -> out
//
// Five launches per pass, no library (cuFFT / ATen) kernel, no allocation; the intermediate streams (harm, src, y) are
// 6 MB each at B = 32 x 2 s and stay in the 126 MB L2 between producer and consumer.  Everything is enqueued on the
// caller's stream; capturable into a CUDA graph.
#include "common.cuh"

using namespace golf;

namespace {
struct FusedGeom {
  int T_osc, n_blocks, Ls, L;
  size_t osc_ws, harm_b, src_b, lpc_ws;
};

bool fused_geom(int B, int Np, int phase_hop, int Fw, int P, int os, int F, int M, int hop, int n_mag, FusedGeom* g) {
  if (B <= 0 || Np <= 0 || phase_hop <= 0 || Fw <= 0 || P <= 0 || os <= 0 || F <= 0 || M <= 0 || hop <= 0 || n_mag < 2) return false;
  const int64_t n_os = (int64_t)(Np - 1) * phase_hop * os + 1;
  const int64_t t_osc = (n_os - 1) / os + 1;
  if (t_osc > INT32_MAX) return false;
  g->T_osc = (int)t_osc;
  const int K = 2 * (n_mag - 1), p = (K - 1) / 2;
  if (g->T_osc + 2 * p < K + hop - 1) return false;
  g->n_blocks = (g->T_osc + 2 * p - (K + hop - 1)) / hop + 1;
  if (g->n_blocks > F) g->n_blocks = F;
  g->Ls = g->n_blocks * hop;
  const int64_t lmax = (int64_t)(F - 1) * hop + 1;
  g->L = (int)(g->Ls < lmax ? g->Ls : lmax);
  g->osc_ws = align_up(golf_glottal_osc_workspace_bytes(B, Np, phase_hop, Fw, P, os), 256);
  g->harm_b = align_up((size_t)B * g->T_osc * sizeof(float), 256);
  g->src_b = align_up((size_t)B * g->Ls * sizeof(float), 256);
  g->lpc_ws = golf_lpc_ss_room_workspace_bytes(B, g->L, M, hop, 0);
  return g->lpc_ws != 0;
}
}  // namespace

GOLF_API size_t golf_synth_fused_workspace_bytes(int B, int Np, int phase_hop, int Fw, int P, int os, int F, int M, int hop,
                                                 int n_mag) {
  FusedGeom g;
  if (!fused_geom(B, Np, phase_hop, Fw, P, os, F, M, hop, n_mag, &g)) return 0;
  return g.osc_ws + g.harm_b + g.src_b + g.lpc_ws;
}

GOLF_API int golf_synth_fused_out_length(int Np, int phase_hop, int os, int F, int hop, int n_mag) {
  FusedGeom g;
  if (!fused_geom(1, Np, phase_hop, 1, 4, os, F, 1, hop, n_mag, &g)) return 0;
  return g.L;
}

GOLF_API int golf_synth_fused_fwd(const float* phase, const float* w, const float* table, const float* dec_kernel,
                                  const float* noise, int64_t noise_stride, uint64_t* rng_state, const float* log_mag,
                                  const float* fir_window, const float* gain, const float* a, const float* room_k, int room_n,
                                  float* out, int B, int Np, int phase_hop, int Fw, int w_hop, int n_tab, int P, int os, int zeros,
                                  int osc_accumulate, int osc_flags, int F, int M, int hop, int n_mag, int refine, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!phase || !w || !table || (!noise && !rng_state) || !log_mag || !fir_window || !a || !out) return GOLF_ERR_INVALID;
  FusedGeom g;
  if (!fused_geom(B, Np, phase_hop, Fw, P, os, F, M, hop, n_mag, &g)) return GOLF_ERR_UNSUPPORTED;
  if (!golf_noise_fir_design_supported(n_mag, hop)) return GOLF_ERR_UNSUPPORTED;
  if (noise && noise_stride < g.T_osc) return GOLF_ERR_INVALID;
  if (!workspace || workspace_bytes < g.osc_ws + g.harm_b + g.src_b + g.lpc_ws) return GOLF_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 255) != 0) return GOLF_ERR_INVALID;
  char* ws = reinterpret_cast<char*>(workspace);
  void* osc_ws = ws;
  float* harm = reinterpret_cast<float*>(ws + g.osc_ws);
  float* src = reinterpret_cast<float*>(ws + g.osc_ws + g.harm_b);
  void* lpc_ws = ws + g.osc_ws + g.harm_b + g.src_b;
  // 1-2: oscillator (prefix scan + flow / decimation)
  int rc = golf_glottal_osc_fwd(phase, w, table, dec_kernel, harm, B, Np, phase_hop, Fw, w_hop, n_tab, P, os, zeros, osc_accumulate,
                                osc_flags, osc_ws, g.osc_ws, stream);
  if (rc) return rc;
  // 3: noise branch + harm -> src
  rc = golf_noise_fir_design_fwd(noise, noise_stride, noise ? nullptr : rng_state, log_mag, fir_window, harm, g.T_osc, src, B, g.T_osc, F,
                                 n_mag, hop, stream);
  if (rc) return rc;
  // 4-5: GOLF-ss filter (+ room FIR); the tail kernel advances the generator state once the pass has consumed it
  if (room_k) {
    rc = golf_lpc_ss_room_fwd(src, g.Ls, gain, a, nullptr, room_k, room_n, nullptr, out, B, g.L, F, M, hop, 0, refine, lpc_ws, g.lpc_ws,
                              stream);
  } else {
    rc = golf_lpc_ss_fwd_passes(src, g.Ls, gain, a, nullptr, out, B, g.L, F, M, hop, 0, lpc_ws, g.lpc_ws, refine ? 15 : 7, stream);
  }
  if (rc) return rc;
  if (!noise) rc = golf_rng_advance(rng_state, stream);
  return rc;
}
