/*
 * golf_b200.h -- C ABI of libgolf_b200.so: the B200 (sm_100a) implementation of
 * GOLF's sample-recurrent synthesis hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream (passed as void* == cudaStream_t); none allocates,
 * none synchronises the device, all enqueue on the given stream and are
 * CUDA-graph capturable.  Return value: 0 on success, a negative GOLF_ERR_* code
 * otherwise (golf_strerror() names it).  All tensors are contiguous float32,
 * row-major, in the layouts the reference's Python passes around.
 *
 * Reference interfaces replaced (paths under the reference repo iamycy/golf):
 *   golf_lpc_ss_*            models/filters.py:99-113  LTVMinimumPhaseFilterPrecise.forward
 *                            (ex*gain, upsample a, torchlpc.sample_wise_lpc) and its autograd
 *                            with hop == 1, gain == NULL: torchlpc.sample_wise_lpc(x, a, zi)
 *                            as called at models/filters.py:112,789 and models/lru/lru.py:15
 *   golf_lpc_ss_room_fwd     models/sf.py:64 room_filter(end_filter(...)): filters.py:99-113 followed by :443-450
 *   golf_lpc_ff_*            models/filters.py:131-184 LTVMinimumPhaseFilter.forward
 *                            (unfold, models/lpc.py:11-16 lpc_synthesis -> torchaudio lfilter, Hann OLA)
 *   golf_biquad_ff_fwd       models/lpc.py:94-131     BatchSecondOrderLPCSynth.forward
 *   golf_biquad_cascade_fwd/bwd  the same (SURVEY 8b's names) and its autograd (torchaudio DifferentiableIIR K times + OLA)
 *   golf_lpc_frames_fwd/bwd  models/lpc.py:19-91      LPCSynth / BatchLPCSynth.forward (per-frame gain, pad (win-hop)/2) + autograd
 *   golf_lfilter_allpole_fwd/bwd  models/lpc.py:11-16 lpc_synthesis -> torchaudio.functional.lfilter(x, [1,a], [g,0..], clamp=False)
 *                            and DifferentiableIIR.backward for that b
 *   golf_biquad_params_fwd/bwd  models/utils.py:487-525 get_logits2biquads (coef|conj|real) + :444-484 biquads2lpc /
 *                            coeff_product, as composed at models/filters.py:73-78, and their autograd
 *   golf_lpc_inverse_fwd/bwd models/filters.py:186-195 reverse() + models/utils.py:433-441 fir_filt, and its
 *                            autograd (inverse-target training, ltng/vocoder.py:192-198)
 *   golf_noise_fir_*         models/filters.py:350-384 LTVZeroPhaseFIRFilter.forward (block FIR)
 *   golf_noise_fir_design_fwd models/filters.py:294-306 + 360-384 (+ models/noise.py:34-35 when it draws the noise itself)
 *   golf_room_fir_*          models/filters.py:443-450 LTIAcousticFilter.forward
 *   golf_synth_fused_fwd     models/sf.py:47-64 SourceFilterSynth.forward, GOLF-ss modules, inference
 *   golf_glottal_osc_fwd     models/synth.py:213-263   IndexedGlottalFlowTable.forward
 *   golf_glottal_osc_fwd_from  the same with phase_offset (models/synth.py:195-218,251-252) as a per-utterance constant
 *   golf_glottal_osc_bwd_w   autograd of the above w.r.t. table_select_weight
 *   golf_glottal_osc_bwd     ... and w.r.t. the wavetable (trainable tables)
 *   golf_wavetable_read_fwd  models/synth.py:124-177   GlottalFlowTable.generate
 *   golf_linear_upsample     models/audiotensor/audiotensor.py:11-17 linear_upsample
 *   golf_rc2lpc_fwd/bwd      models/utils.py:581-593   rc2lpc (with the tanh*max_abs of filters.py:80) and its autograd
 *   golf_mss_*               loss/spec.py:11-67 SSSLoss / MSSLoss (torchaudio Spectrogram power=1 -> L1 + alpha log2-L1) and its
 *                            autograd w.r.t. the prediction: the training-step consumer of the decoder output (ltng/ae.py:119-121)
 *   golf_exp_to_complex      models/filters.py:295-296 `torch.exp(log_mag) + 0j` of get_zero_phase_fir
 *   (golf_set_pdl, golf_fir_set_variant, golf_lpc_ss_set_*, golf_glottal_osc_set_variant: process-wide kernel
 *    selection switches for A/B timing and tests; no reference counterpart)
 */
#ifndef GOLF_B200_H_
#define GOLF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOLF_B200_ABI_VERSION 8

enum {
  GOLF_OK = 0,
  GOLF_ERR_INVALID = -1,      /* bad shape / null pointer / misaligned buffer            */
  GOLF_ERR_UNSUPPORTED = -2,  /* valid request outside what the kernels are built for    */
  GOLF_ERR_WORKSPACE = -3,    /* workspace too small (ask golf_*_workspace_bytes)        */
  GOLF_ERR_CUDA = -4          /* a CUDA runtime call or launch failed (see golf_last_cuda_error) */
};

int golf_abi_version(void);
const char *golf_strerror(int code);
/* last cudaError_t seen by this library on the calling thread (0 = none) */
int golf_last_cuda_error(void);
/* 1 (default): the oscillator's flow kernel is launched with programmatic stream serialization
 * (its prologue -- taps, table rows -- runs beside the knot-prefix scan and it waits on-device for the
 * scan); 0: ordinary launches.  Results are identical. */
void golf_set_pdl(int on);
/* number of kernel launches issued by this library since load (all threads) */
uint64_t golf_launch_count(void);

/* ------------------------------------------------------------------ GOLF-ss ---- */
/* Time-varying all-pole filter on FRAME-RATE controls, coefficients interpolated
 * to sample rate in-kernel with ATen's align_corners=True arithmetic:
 *   e[t] = ex[t] * up(gain)[t];  y[t] = e[t] - sum_{i<M} up(a)[t,i] * y[t-1-i]
 * ex [B,T_ex] (row stride ex_stride), gain [B,F] (may be NULL == 1), a [B,F,M],
 * zi [B,M] or NULL (zi[:,j] = y[-1-j]), y [B,L] with L = min(T_ex,(F-1)*hop+1)
 * (pass L; it is checked).  hop == 1 gives torchlpc.sample_wise_lpc exactly.
 * chunk = 0 lets the library pick the time-chunk length. */
size_t golf_lpc_ss_workspace_bytes(int B, int L, int M, int hop, int chunk);
/* The refinement round runs per sequence only when the states the chunks really
 * ended in differ from the stitched ones by more than tol * max|state| (default 1e-4; 0 = always).
 * Process-wide setting. */
void golf_lpc_ss_set_refine_tolerance(float tol);
float golf_lpc_ss_get_refine_tolerance(void);
/* Solve-pass variant of the forward filter: 1 (default) = four lanes per chunk (tap blocks pipelined
 * across lanes), 0 = one lane per chunk.  Same recurrence, different summation order
 * (results agree to float32 rounding).  Process-wide; for A/B timing and tests. */
void golf_lpc_ss_set_solver(int systolic);
int golf_lpc_ss_fwd(const float *ex, int64_t ex_stride, const float *gain, const float *a,
                    const float *zi, float *y, int B, int L, int F, int M, int hop,
                    int chunk, void *workspace, size_t workspace_bytes, void *stream);
/* Same, running only the selected passes (bit 0: chunk responses, bit 1: stitch, bit 2:
 * solve, bit 3: one refinement round = mismatch stitch + solve again, bit 4: zero-state chunk
 * responses by a solve from rest, run before the stitch).  golf_lpc_ss_fwd is passes == 15; 7
 * skips the refinement (faster, less accurate on high-gain filters).
 * Two-call form for overlapping with the producer of `ex`: the chunk transition matrices depend
 * on `a` alone, so  passes == 1 with ex == NULL, y == NULL  may be enqueued on a side stream as
 * soon as `a` exists; once ex is ready (and after a stream join)  passes == 30  on the SAME
 * workspace finishes the job.  Results are bit-identical to passes == 15. */
int golf_lpc_ss_fwd_passes(const float *ex, int64_t ex_stride, const float *gain,
                           const float *a, const float *zi, float *y, int B, int L, int F,
                           int M, int hop, int chunk, void *workspace, size_t workspace_bytes,
                           int passes, void *stream);
/* Passes 2..4 (stitch, solve, refinement -- and the room FIR of golf_lpc_ss_room_fwd) as ONE launch, a thread-block
 * cluster per sequence with a two-level stitch: mode 0 (default) never, 1 wherever the kernel applies,
 * 2 for batches of at most 8 sequences.  An opt-in schedule: it shortens the filter's serial part and brings a decoder
 * pass down to five launches, but it occupies whole SMs, which costs throughput once several passes are in flight
 * (measured, profiles/README.md).  Same recurrences, different grouping of the state propagation (results agree to
 * float32 rounding).  Process-wide. */
void golf_lpc_ss_set_tail(int mode);
int golf_lpc_ss_get_tail(void);
/* Pass 1 (chunk responses) at padded order 24 (orders 21..24, chunk a multiple of 8 samples, forward form): mode 0
 * (default) the FP32 kernel, 1 the mma.sync TF32 tensor-core kernel with error-compensated (3 x TF32) products.  An opt-in
 * experiment: on B200 it is slower (83 vs 74 us at B = 32 x 2 s) and its transition matrices are ~10x less accurate, so the
 * refinement round runs for most sequences.  Other shapes and the adjoint always use the FP32 kernel. */
void golf_lpc_ss_set_response(int mode);
int golf_lpc_ss_get_response(void);
/* GOLF-ss end filter + the room filter behind it (models/sf.py:64: room_filter(end_filter(src, gain, a)) with
 * models/filters.py:99-113 and :443-450) in the same launches: out[t] = y[t] + sum_{j<room_n} room_k[j] y[t-room_n+j],
 * y = golf_lpc_ss_fwd(ex, gain, a, zi).  y may be NULL (the filter output then lives in the workspace only; pass a
 * buffer to keep it for the adjoint); refine != 0 allows the refinement round.  room_n <= 252 taps. */
size_t golf_lpc_ss_room_workspace_bytes(int B, int L, int M, int hop, int chunk);
int golf_lpc_ss_room_fwd(const float *ex, int64_t ex_stride, const float *gain, const float *a, const float *zi,
                         const float *room_k, int room_n, float *y, float *out, int B, int L, int F, int M,
                         int hop, int chunk, int refine, void *workspace, size_t workspace_bytes, void *stream);
/* Adjoint.  Inputs: gy = dL/dy [B,L], saved y, ex, gain, a, zi.  Outputs (any may
 * be NULL): d_ex [B,L], d_gain [B,F], d_a [B,F,M], d_zi [B,M]. */
size_t golf_lpc_ss_bwd_workspace_bytes(int B, int L, int M, int hop, int chunk);
int golf_lpc_ss_bwd(const float *gy, const float *y, const float *ex, int64_t ex_stride,
                    const float *gain, const float *a, const float *zi, float *d_ex,
                    float *d_gain, float *d_a, float *d_zi, int B, int L, int F, int M,
                    int hop, int chunk, int refine, void *workspace, size_t workspace_bytes,
                    void *stream);

/* ------------------------------------------------------------------ GOLF-ff ---- */
/* Frame-wise filter: e = ex*up(gain); frames of `win` samples every `hop` from
 * e zero-padded by win/2 each side; per-frame LTI all-pole from zero state with
 * a[b,k,:]; Hann(periodic) overlap-add, divided by the overlap-added window.
 * n_frames = (L_e + 2*(win/2) - win)/hop + 1 <= F where L_e = min(T_ex,(F-1)*hop+1);
 * y [B, (n_frames-1)*hop].  Requires win == 4*hop (the reference's shipped ratio).
 * window: [win] device array (the module's window buffer). */
/* 1: the per-frame recurrences of golf_lpc_ff_fwd / golf_lpc_frames_fwd sum their taps exactly like libtorchaudio's CPU loop
 * (one accumulator, oldest tap first, rounded multiply then rounded subtract): per frame bit-identical to
 * torchaudio.functional.lfilter, ~2.5x the dependent latency.  0 (default): three interleaved FMA chains.  Process-wide. */
void golf_lpc_ff_set_exact_order(int on);
int golf_lpc_ff_fwd(const float *ex, int64_t ex_stride, const float *gain, const float *a,
                    const float *window, float *y, int B, int T_ex, int F, int M, int hop,
                    int win, void *stream);
/* Adjoint of golf_lpc_ff_fwd: gy [B,(n_frames-1)*hop] -> d_ex [B,T_ex] (row stride dex_stride),
 * d_gain [B,F], d_a [B,F,M] (any may be NULL).  Recomputes the per-frame recurrences into the
 * workspace.  Requires win % MP == 0 for the padded order MP (true for every shipped config). */
size_t golf_lpc_ff_bwd_workspace_bytes(int B, int T_ex, int F, int hop, int win);
int golf_lpc_ff_bwd(const float *gy, const float *ex, int64_t ex_stride, const float *gain,
                    const float *a, const float *window, float *d_ex, int64_t dex_stride,
                    float *d_gain, float *d_a, int B, int T_ex, int F, int M, int hop, int win,
                    void *workspace, size_t workspace_bytes, void *stream);
/* Cascade of K second-order all-pole sections per frame (biquads [B,F,K,3]),
 * gain applied per frame, zero-pad (win-hop)/2, Hann OLA + normalise. */
int golf_biquad_ff_fwd(const float *ex, int64_t ex_stride, const float *gain,
                       const float *biquads, const float *window, float *y, int B, int T_ex,
                       int F, int K, int hop, int win, void *stream);

/* BatchSecondOrderLPCSynth under the name SURVEY 8(b) gives it; identical to golf_biquad_ff_fwd. */
int golf_biquad_cascade_fwd(const float *ex, int64_t ex_stride, const float *gain,
                            const float *biquads, const float *window, float *y, int B, int T_ex,
                            int F, int K, int hop, int win, void *stream);
/* Adjoint of the cascade for an upstream gradient gy [B,out_len]: d_ex [B,T_ex] (contiguous, required),
 * d_gain [B,F], d_biquads [B,F,K,3] (either may be NULL; frames beyond n_frames get zeros).  The kernel
 * recomputes every section's output into the workspace (32*K*win floats per CTA). */
size_t golf_biquad_cascade_bwd_workspace_bytes(int B, int T_ex, int F, int K, int hop, int win);
int golf_biquad_cascade_bwd(const float *gy, const float *ex, int64_t ex_stride, const float *gain,
                            const float *biquads, const float *window, float *d_ex, float *d_gain,
                            float *d_biquads, int B, int T_ex, int F, int K, int hop, int win,
                            void *workspace, size_t workspace_bytes, void *stream);

/* LPCSynth / BatchLPCSynth (models/lpc.py:19-91): frames of `win` samples every `hop`, zero-padded by
 * (win-hop)/2, per-frame LTI all-pole a[b,k,:] on gain[b,k]*frame from zero state, window OLA + normalise.
 * y [B, out_len], out_len = golf_lpc_frames_out_length(...) (0: invalid geometry). */
int golf_lpc_frames_out_length(int T_ex, int F, int hop, int win);
int golf_lpc_frames_fwd(const float *ex, int64_t ex_stride, const float *gain, const float *a,
                        const float *window, float *y, int B, int T_ex, int F, int M, int hop,
                        int win, void *stream);
size_t golf_lpc_frames_bwd_workspace_bytes(int B, int T_ex, int F, int hop, int win);
int golf_lpc_frames_bwd(const float *gy, const float *ex, int64_t ex_stride, const float *gain,
                        const float *a, const float *window, float *d_ex, float *d_gain, float *d_a,
                        int B, int T_ex, int F, int M, int hop, int win, void *workspace,
                        size_t workspace_bytes, void *stream);

/* lfilter-shaped twin of models/lpc.py:11-16: y[c,n] = gain[c]*x[c,n] - sum_i a[c,i] y[c,n-1-i], zero
 * initial state; x [C,N] (row stride x_stride), gain [C] or NULL (= 1), a [C,M], y [C,N]. */
int golf_lfilter_allpole_fwd(const float *x, int64_t x_stride, const float *gain, const float *a,
                             float *y, int C, int N, int M, void *stream);
/* Adjoint: gy, y [C,N] -> d_x [C,N], d_gain [C], d_a [C,M] (any may be NULL). */
size_t golf_lfilter_allpole_bwd_workspace_bytes(int C, int N);
int golf_lfilter_allpole_bwd(const float *gy, const float *x, int64_t x_stride, const float *y,
                             const float *gain, const float *a, float *d_x, float *d_gain, float *d_a,
                             int C, int N, int M, void *workspace, size_t workspace_bytes, void *stream);

/* Biquad parameterisations (rep 0 "coef", 1 "conj", 2 "real"; models/utils.py:487-525) and the polynomial
 * product of the K sections (models/utils.py:444-484): logits [N,K,2] -> biquads [N,K,3] and / or
 * a [N,2K] (either may be NULL).  K <= 16. */
int golf_biquad_params_fwd(const float *logits, float *biquads, float *a, int N, int K, int rep,
                           float max_abs_pole, void *stream);
/* d_logits [N,K,2] from d_biquads [N,K,3] and / or d_a [N,2K] (either may be NULL). */
int golf_biquad_params_bwd(const float *logits, const float *d_biquads, const float *d_a,
                           float *d_logits, int N, int K, int rep, float max_abs_pole, void *stream);

/* Inverse (analysis) filter: r[t] = y[t] + sum_i up(a)[t,i] y[t-1-i], [B,L]. */
int golf_lpc_inverse_fwd(const float *y, int64_t y_stride, const float *a, float *r, int B,
                         int L, int F, int M, int hop, void *stream);
/* Adjoint of the inverse filter for an upstream gradient g [B,L]: d_y [B,L] (w.r.t. the first L samples of
 * y; later samples do not reach r) and d_a [B,F,M]; either may be NULL.  Used when the reference trains with
 * an inverse-filtered target (ltng/vocoder.py:192-198). */
int golf_lpc_inverse_bwd(const float *g, const float *y, int64_t y_stride, const float *a, float *d_y,
                         float *d_a, int B, int L, int F, int M, int hop, void *stream);

/* ------------------------------------------------- multi-scale spectral loss ---- */
/* loss/spec.py:11-67 with torchaudio's Spectrogram defaults (centre / reflect padding, periodic Hann of n_fft, hop =
 * n_fft - int(0.75 n_fft), onesided, power 1):  loss = ratio * sum_scales ( mean|Sp - St| + alpha mean|log2(St + eps) -
 * log2(Sp + eps)| ).  The STFTs are DFT-as-GEMM on the tcgen05 tensor cores (the shipped sizes 509 / 1021 / 2053 are primes),
 * prec3: bit 0 error-compensated 3 x TF32 products (float32-grade) in the forward GEMMs, bit 1 in the adjoint GEMM (3 = both).  tables[i]: the DFT bases of n_ffts[i], built once by
 * golf_mss_build_tables into golf_mss_tables_bytes(n_fft) bytes.  loss: one float on the device.  d_pred (optional, [B, L]
 * with row stride dpred_stride): d loss / d pred.  pred, target [B, L].  hops: one hop per scale, or NULL for the 75 % overlap
 * of the shipped config.  At most 8 scales, n_fft <= 4096, L > n_fft / 2. */
size_t golf_mss_tables_bytes(int n_fft);
int golf_mss_build_tables(int n_fft, float *tables, void *stream);
size_t golf_mss_workspace_bytes(int B, int L, const int *n_ffts, const int *hops, int n_scales);
int golf_mss_loss(const float *pred, int64_t pred_stride, const float *target, int64_t target_stride, int B,
                  int L, const int *n_ffts, const int *hops, int n_scales, const float *const *tables, float alpha,
                  float ratio,
                  float eps, float *loss, float *d_pred, int64_t dpred_stride, int prec3, void *workspace,
                  size_t workspace_bytes, void *stream);
/* The GEMM underneath (D [M,N] = A [M,K] . Bt [N,K]^T, float32 in and out, row pitches in floats and multiples of 4, bn =
 * N tile: a multiple of 16 up to 256): TMA-staged operands, tcgen05.mma kind::tf32, accumulator in TMEM. */
int golf_mss_gemm(const float *A, int64_t pitchA, const float *Bt, int64_t pitchB, float *D, int64_t pitchD, int M,
                  int N, int K, int bn, int prec3, void *stream);

/* -------------------------------------------------------------- FIR stages ---- */
/* Block-wise time-varying FIR: output block k (hop samples) is the valid
 * cross-correlation of ex zero-padded by (K-1)/2 with kernel[b,k,:];
 * n_blocks = min((T + 2*((K-1)/2) - (K+hop-1))/hop + 1, F); y [B, n_blocks*hop].
 * If add != NULL ([B, >= n_blocks*hop], row stride add_stride) it is added to the
 * result (fuses `harm + noise_filter(noise)`, models/sf.py:53-56).
 * If window != NULL ([K], K even), `kernel` is the RAW irfft(exp(log_mag)) output and the
 * fftshift + windowing of models/filters.py:294-306 happen while the taps are staged. */
int golf_noise_fir_fwd(const float *ex, int64_t ex_stride, const float *kernel,
                       const float *window, const float *add, int64_t add_stride, float *y,
                       int B, int T, int F, int K, int hop, void *stream);
/* 1 (default): packed-FP32 (fma.rn.f32x2) FIR kernels where they apply; 0: the scalar register
 * tile.  Both sum each output's taps in the same order: results are bit-identical. */
void golf_fir_set_variant(int x2);
/* Adjoint w.r.t. the input (d_ex [B,T]) and the final taps (d_kernel [B,F,K]); either may be
 * NULL.  gy [B, n_blocks*hop]. */
int golf_noise_fir_bwd(const float *gy, const float *ex, int64_t ex_stride,
                       const float *kernel, float *d_ex, float *d_kernel, int B, int T, int F,
                       int K, int hop, void *stream);
/* Noise branch in one kernel (n_mag == 256, 128 < hop <= 256, hop % 4 == 0: golf_noise_fir_design_supported):
 * taps designed in-kernel from log_mag [B,F,n_mag] (models/filters.py:294-306: exp -> irfft -> fftshift -> window, as
 * a cosine series -- no FFT library, no [B,F,K] kernel tensor), block FIR as golf_noise_fir_fwd, optional `add`.
 * window: [2*(n_mag-1)] the module's window (NOT divided by K).  Noise: `ex` [B,T] (e.g. the torch.randn draw of
 * models/noise.py:34-35), or ex == NULL and rng_state -> device {seed, offset} (two uint64): white N(0,1) noise from
 * Philox4x32-10 + Box-Muller inside the kernel (same distribution as randn_like, not the same stream; the caller
 * advances the offset between calls: golf_rng_advance).  golf_philox_normal writes that very draw to memory. */
int golf_noise_fir_design_supported(int n_mag, int hop);
int golf_noise_fir_design_fwd(const float *ex, int64_t ex_stride, const uint64_t *rng_state, const float *log_mag,
                              const float *window, const float *add, int64_t add_stride, float *y, int B, int T,
                              int F, int n_mag, int hop, void *stream);
int golf_philox_normal(float *out, int B, int T, const uint64_t *rng_state, void *stream);
int golf_rng_advance(uint64_t *rng_state, void *stream);
/* out[t] = x[t] + sum_{j<n} k[j] x[t-n+j]   (n = length-1 learned taps) */
int golf_room_fir_fwd(const float *x, const float *k, float *out, int B, int T, int n,
                      void *stream);
/* Adjoint: d_x [B,T] and d_k [n] (either may be NULL).  d_k is accumulated with float
 * atomics: its last bit is not reproducible run to run. */
int golf_room_fir_bwd(const float *gy, const float *x, const float *k, float *d_x, float *d_k,
                      int B, int T, int n, void *stream);

/* -------------------------------------------------------------- oscillator ---- */
/* Glottal-flow wavetable oscillator at `os`x oversampling.
 * phase [B,Np] cycles/sample at hop phase_hop; w [B,Fw] in [0,1] at hop w_hop;
 * table [n_tab, P]; dec_kernel [2*zeros*os+1] (os > 1).  N_os = (Np-1)*phase_hop*os+1
 * oversampled samples, out [B, (N_os-1)/os+1].
 * accumulate: 0 = exact running phase in 64-bit fixed point (Q0.64 cycles; default,
 * closer to exact than the reference), 1 = "aten_cpu": float64 running sum rounded to float32
 * before `% 1`, the arithmetic of ATen's CPU cumsum (parity checks).
 * flags bit0: equal_energy (multiply by rsqrt(upsampled phase)). */
size_t golf_glottal_osc_workspace_bytes(int B, int Np, int phase_hop, int Fw, int P, int os);
/* Forward kernel variant in exact-phase mode: 1 (default) = table rows staged in shared memory,
 * incremental fixed-point phase; 0 = the first-generation kernel.  Same arithmetic.
 * Process-wide; for A/B timing and tests. */
void golf_glottal_osc_set_variant(int v2);
int golf_glottal_osc_fwd(const float *phase, const float *w, const float *table,
                         const float *dec_kernel, float *out, int B, int Np, int phase_hop,
                         int Fw, int w_hop, int n_tab, int P, int os, int zeros,
                         int accumulate, int flags, void *workspace, size_t workspace_bytes,
                         void *stream);
/* Same with an initial running phase phase0 [B] (cycles, float64; `phase_offset` of models/synth.py:195-218,251-252
 * for a per-utterance constant): the stream continues where a previous call stopped.  Exact-phase mode only. */
int golf_glottal_osc_fwd_from(const float *phase, const float *w, const float *table,
                              const float *dec_kernel, float *out, const double *phase0, int B, int Np,
                              int phase_hop, int Fw, int w_hop, int n_tab, int P, int os, int zeros,
                              int accumulate, int flags, void *workspace, size_t workspace_bytes,
                              void *stream);
/* Adjoint w.r.t. the selection weight: gout [B, n_out] -> d_w [B,Fw] (same geometry arguments
 * and workspace size as the forward; the gradient w.r.t. phase is not provided -- the shipped
 * configs detach f0).  Accumulated with float atomics (last bit not reproducible). */
int golf_glottal_osc_bwd_w(const float *gout, const float *phase, const float *w,
                           const float *table, const float *dec_kernel, float *d_w, int B, int Np,
                           int phase_hop, int Fw, int w_hop, int n_tab, int P, int os, int zeros,
                           int accumulate, int flags, void *workspace, size_t workspace_bytes,
                           void *stream);
/* Both adjoints of the oscillator that the reference's autograd provides for a detached f0: d_w [B,Fw] as above and
 * d_table [n_tab,P], the gradient w.r.t. the wavetable itself (GlottalFlowTable(trainable=True) registers the table as
 * an nn.Parameter, models/synth.py:157-166).  Either pointer may be NULL (not both).  d_table needs the exact-phase
 * mode, oversampling 1/2/4 and a power-of-two table length (GOLF_ERR_UNSUPPORTED otherwise). */
int golf_glottal_osc_bwd(const float *gout, const float *phase, const float *w, const float *table,
                         const float *dec_kernel, float *d_w, float *d_table, int B, int Np,
                         int phase_hop, int Fw, int w_hop, int n_tab, int P, int os, int zeros,
                         int accumulate, int flags, void *workspace, size_t workspace_bytes,
                         void *stream);
/* GlottalFlowTable.generate: wrapped [B,N] in [0,1), tables [B,R,P] at hop hop_tab. */
int golf_wavetable_read_fwd(const float *wrapped, const float *tables, float *out, int B,
                            int N, int R, int P, int hop_tab, void *stream);

/* ------------------------------------------------------------ whole decoder ---- */
/* SourceFilterSynth.forward (models/sf.py:47-64) for the GOLF-ss decoder of cfg/ae/decoder/golf-precise.yaml in one
 * call: oscillator -> (+ FIR-filtered noise, golf_noise_fir_design_fwd) -> sample-wise LPC filter -> room FIR.  Five
 * launches (six with the in-kernel generator's offset bump), no library kernel, no allocation; harm / src / y live in
 * the workspace.  Arguments as in golf_glottal_osc_fwd, golf_noise_fir_design_fwd and golf_lpc_ss_room_fwd;
 * noise == NULL selects the in-kernel generator (rng_state is advanced); room_k == NULL skips the room filter.
 * out [B, golf_synth_fused_out_length(...)].  The workspace must be 256-byte aligned. */
size_t golf_synth_fused_workspace_bytes(int B, int Np, int phase_hop, int Fw, int P, int os, int F, int M, int hop,
                                        int n_mag);
int golf_synth_fused_out_length(int Np, int phase_hop, int os, int F, int hop, int n_mag);
int golf_synth_fused_fwd(const float *phase, const float *w, const float *table, const float *dec_kernel,
                         const float *noise, int64_t noise_stride, uint64_t *rng_state, const float *log_mag,
                         const float *fir_window, const float *gain, const float *a, const float *room_k, int room_n,
                         float *out, int B, int Np, int phase_hop, int Fw, int w_hop, int n_tab, int P, int os,
                         int zeros, int osc_accumulate, int osc_flags, int F, int M, int hop, int n_mag, int refine,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------- frame-rate helpers ---- */
/* x [R,n] -> out [R,(n-1)*hop+1], F.interpolate(linear, align_corners=True) */
int golf_linear_upsample(const float *x, float *out, int R, int n, int hop, void *stream);
/* a = step_up(tanh(logits) * max_abs); logits, a: [N, M] */
int golf_rc2lpc_fwd(const float *logits, float *a, int N, int M, float max_abs, void *stream);
/* Adjoint: d_logits [N,M] from d_a [N,M] (M <= 40); the recursion is recomputed per frame. */
int golf_rc2lpc_bwd(const float *logits, const float *d_a, float *d_logits, int N, int M,
                    float max_abs, void *stream);
/* out[i] = (exp(x[i]), 0) as interleaved complex64: the spectrum handed to the inverse real FFT of the
 * zero-phase FIR design (models/filters.py:295-297), in one pass. */
int golf_exp_to_complex(const float *x, float *out_interleaved, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GOLF_B200_H_ */
