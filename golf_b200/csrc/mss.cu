// mss.cu -- multi-scale spectral loss (loss/spec.py:11-67: MSSLoss over SSSLoss = L1 + alpha * log2-L1 of STFT magnitudes,
// n_fft 509 / 1021 / 2053 at 75 % overlap in cfg/ae/vctk.yaml:58-67) on the 5th-generation tensor cores.
//
// The three STFT sizes are primes: cuFFT runs them as Bluestein transforms (~2 ms forward + backward per training step at
// B = 32 x 2 s, two thirds of the decoder-side step).  A prime-size DFT of many frames is a GEMM,
//     [frames x n_fft] (windowed frames)  x  [n_fft x 2 nbins] (cos | -sin, interleaved)  ->  (re, im) per frame and bin,
// and so is its adjoint.  This file runs both on tcgen05:
//
//   mss_frames_kernel   reflect padding (torch.stft center=True), framing (hop = n_fft - int(0.75 n_fft) is odd, so the frames
//                       cannot be a strided TMA view of the signal: 16-byte stride rule) and the periodic Hann window
//   mss_gemm_kernel     D[128 x N_tile] tiles: operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 2-stage
//                       shared-memory ring, tcgen05.mma.kind::tf32 issued by one thread, accumulator in TMEM, read back with
//                       tcgen05.ld by four epilogue warps.  Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM allocation),
//                       2..5 splitter during the main loop, epilogue afterwards.
//                       Float32-grade products (the log-magnitude term reads bins 60 dB below a frame's peak): the tensor
//                       core reads the top 19 bits of an operand (hi); the splitter warps write lo = rna_tf32(x - hi) into a
//                       second pair of tiles in the SAME swizzled layout; every k-step issues hi*hi + lo*hi + hi*lo.
//                       Epilogues: magnitudes of the target (mode 0); loss terms + d loss / d(re, im) for the prediction
//                       (mode 1: the two signals' spectra never exist in HBM as complex tensors); plain store (mode 2, adjoint).
//   mss_ola_kernel      adjoint of framing + window + reflect padding: gather, no atomics
//
// C ABI: golf_mss_*.  Tables (the DFT bases, built once per n_fft in double precision) and the workspace are caller-provided.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace golf {

constexpr int kMssMaxScales = 8;
constexpr int kGemmThreads = 512;  // 4 warpgroups: producer + MMA | splitter | epilogue (column half 0) | epilogue (column half 1)
constexpr int kBM = 128;          // rows per tile (TMEM lanes)
constexpr int kBK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kAccCols = 112;     // float32 REGISTER accumulators per epilogue thread (one row, one column half of the tile)
constexpr int kMaxBN = 208;       // columns per tile (two epilogue warpgroups share a row: 7 + 6 groups of 16); rows of a B tile
constexpr int kTmemCols = 256;    // TMEM columns per accumulator (allocation granularity)
constexpr int kStages = 2;
constexpr int kAccBufs = 2;       // TMEM accumulators: the tensor cores run one chunk ahead of the register accumulation
constexpr int kTileA = kBM * kBK * 4;        // 16 KB
constexpr int kTileB = kMaxBN * kBK * 4;     // 26 KB
constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;  // A | A_lo | B | B_lo
constexpr int kSst = kMaxBN / 2 + 1;         // row stride of the staged target magnitudes (conflict-free column reads)
constexpr int kSstBytes = kBM * kSst * 4;
constexpr int kGemmSmem = kStages * kStageBytes + kSstBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct MssGemmParams {
  int M, N, K;          // D [M x N] = A [M x K] . Bt[N x K]^T
  int bn;               // N tile (multiple of 16, <= 256)
  int mode;             // 0: magnitudes -> out [M, N/2]; 1: loss + gradient -> out [M, N] (needs s_true); 2: plain store
  int prec3;            // error-compensated products
  int chunk_kb;         // k-blocks summed inside the tensor core before the sum moves to float32 registers
  float* out;
  int64_t out_pitch;
  const float* s_true;  // [M, N/2]
  int64_t st_pitch;
  double* loss_acc;     // [gridDim.x][2] per-CTA partial sums: |Sp - St|, |log2(St + eps) - log2(Sp + eps)| (no atomics: same-address
                        // atomics serialise in L2 -- 1 184 of them cost 46 us per launch -- and the sum stays deterministic)
  float alpha, eps;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B (what the TMA box {32 floats, rows} produces at a 1024-byte aligned
// address): start address >> 4, LBO = 1 (ignored for swizzled K-major), SBO = 8 rows * 128 B = 1024 B, version 1, layout 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, both operands K-major: c_format 1 @4, a/b_format 2 (TF32) @7/@10, N >> 3 @17, M >> 4 @24
__host__ __device__ inline uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- the GEMM -------------------------------------------------------------------------------------------------------
// Persistent: CTA c works on tiles c, c + gridDim.x, ... (n fastest, so the CTAs of a wave share A row blocks in L2).
//
// Accumulation.  The tensor core adds products into its fp32 accumulator with TRUNCATION; over the 770 MMA steps of a
// K = 2053 product the bias reaches 2e-5 of the largest partial sum, which is fatal for this loss: the partial sums of a
// spectral valley next to a strong harmonic are as large as the harmonic until the window closes, and the log-magnitude term
// (and its 1/S gradient) reads exactly those valleys.  So TMEM only ever holds the sum of a few k-blocks (p.chunk_kb = 4: K = 128, 48
// MMAs): four 128-column accumulators rotate, and the epilogue warps add each finished chunk into float32 REGISTERS
// (round-to-nearest, one row of <= 128 columns per thread) while the tensor cores run the next chunk.  Measured: gradient
// error in -60 dB valleys from 1e-1 to the level of torch's own float32 cuFFT path.
//   warp 0      TMA producer            full[s]  <- TMA bytes          (waits empty[s])
//   warp 1      MMA issuer, TMEM alloc  empty[s] <- tcgen05.commit     (waits ready[s] | full[s], tempty[b]);  tfull[b] <- commit
//   warps 2..5  splitter (prec3)        ready[s] <- 128 arrivals       (waits full[s])
//   warps 6..9  accumulate + epilogue   tempty[b] <- 128 arrivals      (waits tfull[b])
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
    mss_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, MssGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  float* sst = reinterpret_cast<float*>(tiles + kStages * kStageBytes);  // [128][kSst] target magnitudes of the tile (mode 1)
  constexpr int kRegProducer = 40, kRegSplit = 96, kRegEpilogue = 184;  // 128 x (40 + 96 + 2 x 184) = 64 512 <= 65 536
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + kStages * kStageBytes + kSstBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* ready = bars + kStages;      // [kStages]
  uint64_t* empty = bars + 2 * kStages;  // [kStages]
  uint64_t* tfull = bars + 3 * kStages;  // [kAccBufs]
  uint64_t* tempty = bars + 3 * kStages + kAccBufs;  // [kAccBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 2 * kAccBufs);
  double* red = reinterpret_cast<double*>(bars + 3 * kStages + 2 * kAccBufs + 1);  // [16] per-warp loss partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + kBK - 1) / kBK;
  const int nch = (nkb + p.chunk_kb - 1) / p.chunk_kb;
  const int tiles_n = (p.N + p.bn - 1) / p.bn, tiles_m = (p.M + kBM - 1) / kBM;
  const int n_tiles = tiles_n * tiles_m;
  const uint32_t stage_tx = (uint32_t)(kBM * kBK * 4 + p.bn * kBK * 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(ready + s, 128);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 256);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kAccBufs * kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegProducer));
  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * p.bn;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty + s, ((it / kStages) - 1) & 1);
          uint8_t* st = tiles + s * kStageBytes;
          mbar_expect_tx(full + s, stage_tx);
          tma_load_2d(st, &tmA, kb * kBK, m0, full + s);
          tma_load_2d(st + 2 * kTileA, &tmB, kb * kBK, n0, full + s);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kBM, p.bn);
      int it = 0, gc = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < nch; ++ch, ++gc) {
          const int buf = gc % kAccBufs;
          if (gc >= kAccBufs) {
            mbar_wait(tempty + buf, ((gc / kAccBufs) - 1) & 1);  // the accumulate warps have drained this accumulator
            tc_fence_after();
          }
          const uint32_t acc = tmem + (uint32_t)(buf * kTmemCols);
          const int kb_end = min(nkb, (ch + 1) * p.chunk_kb);
          for (int kb = ch * p.chunk_kb; kb < kb_end; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(p.prec3 ? ready + s : full + s, (it / kStages) & 1);
            tc_fence_after();
            const uint32_t sa = base + s * kStageBytes;
            const uint64_t da = umma_desc(sa), dal = umma_desc(sa + kTileA), db = umma_desc(sa + 2 * kTileA),
                           dbl = umma_desc(sa + 2 * kTileA + kTileB);
            const bool first = kb == ch * p.chunk_kb;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes: +2 in the (>> 4) start-address field
              const uint64_t o = (uint64_t)(2 * k);
              tc_mma_tf32(acc, da + o, db + o, idesc, (first && k == 0) ? 0u : 1u);
              if (p.prec3) {
                tc_mma_tf32(acc, dal + o, db + o, idesc, 1u);
                tc_mma_tf32(acc, da + o, dbl + o, idesc, 1u);
              }
            }
            tc_commit(empty + s);  // frees the stage once these MMAs have read it
          }
          tc_commit(tfull + buf);
        }
      }
    }
  }
  // (warps 2 and 3 of the producer warpgroup idle)
  } else if (warp < 8) {
    // ===== splitter: lo = x - hi for both operand tiles, same swizzled positions
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegSplit));
    if (p.prec3) {
      const int st_tid = threadIdx.x - 128;  // 0..127
      // The tensor core reads the top 19 bits of an fp32 operand, i.e. hi = x truncated to TF32, for free; the remainder
      // x - hi is exact in fp32 (13 significant bits) and is stored ROUNDED to TF32 (cvt.rna), so what the tensor core then
      // truncates away of it is nothing: x = hi + lo to 2^-22 |x|, unbiased.  (Writing a rounded hi back in place as well was
      // 2x closer still but cost a third more shared-memory traffic -- the resource this kernel is bound by: TMA fill +
      // splitter + three operand reads per MMA.)
      auto lo_of = [](float x) {
        const float r = __fsub_rn(x, __uint_as_float(__float_as_uint(x) & 0xffffe000u));
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(r));
        return __uint_as_float(h);
      };
      const int nb16 = p.bn >> 4;  // B: 8 float4 per row, 128 threads -> bn / 16 float4 per thread
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full + s, (it / kStages) & 1);
          const float4* a = reinterpret_cast<const float4*>(tiles + s * kStageBytes) + st_tid;
          float4* al = reinterpret_cast<float4*>(tiles + s * kStageBytes + kTileA) + st_tid;
          const float4* b = reinterpret_cast<const float4*>(tiles + s * kStageBytes + 2 * kTileA) + st_tid;
          float4* bl = reinterpret_cast<float4*>(tiles + s * kStageBytes + 2 * kTileA + kTileB) + st_tid;
#pragma unroll
          for (int i = 0; i < kBM * kBK / 4 / 128; ++i) {
            const float4 v = a[128 * i];
            al[128 * i] = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
          }
#pragma unroll 4
          for (int i = 0; i < nb16; ++i) {
            const float4 v = b[128 * i];
            bl[128 * i] = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          mbar_arrive(ready + s);
        }
      }
    }
  } else {
    // ===== accumulate + epilogue: thread = one row of the tile, one column half
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegEpilogue));
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int nbins = p.N >> 1;
    const int trow = 32 * q + lane;  // row within the tile
    // the tile's 16-column groups are divided between the two epilogue warpgroups: [0, cs) and [cs, bn)
    const int ngrp = p.bn >> 4, g_half = (ngrp + 1) >> 1;
    const int cs = warp >= 12 ? 16 * g_half : 0;                  // first column of this thread's part
    const int cw = warp >= 12 ? p.bn - 16 * g_half : 16 * g_half;  // its width (<= kAccCols)
    double lin_d = 0.0, lg_d = 0.0;
    int gc = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * p.bn;
      const int row = m0 + trow;
      const bool row_ok = row < p.M;
      if (MODE == 1) {
        // this warp's 32 rows of the target magnitudes, read along the bins (coalesced, independent loads) long before use
        const int hb = cw >> 1, b0 = (n0 + cs) >> 1, s0 = cs >> 1;  // this warp: its 32 rows, its column half
        __syncwarp();  // the previous tile's reads of sst are done
        // asynchronous copies (LDGSTS, zero fill outside the matrix): nothing waits for them until the tile's last chunk
        for (int rr = 0; rr < 32; ++rr) {
          const int grow = m0 + 32 * q + rr;
          for (int j = lane; j < hb; j += 32) {
            const bool ok = grow < p.M && b0 + j < nbins;
            cp_async4(sst + (32 * q + rr) * kSst + s0 + j, p.s_true + (ok ? (size_t)grow * p.st_pitch + b0 + j : 0), ok);
          }
        }
      }
      float acc[kAccCols];
#pragma unroll
      for (int i = 0; i < kAccCols; ++i) acc[i] = 0.f;
      for (int ch = 0; ch < nch; ++ch, ++gc) {
        const int buf = gc % kAccBufs;
        mbar_wait(tfull + buf, (gc / kAccBufs) & 1);
        tc_fence_after();
        const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * kTmemCols + cs);
#pragma unroll
        for (int c = 0; c < kAccCols / 16; ++c) {
          if (16 * c < cw) {
            uint32_t r[16];
            tc_ld16(tq + (uint32_t)(16 * c), r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty + buf);
      }
      // ---- the tile's epilogue, from registers
      if (MODE == 1) {
        cp_async_wait_all();
        __syncwarp();
      }
      float lin = 0.f, lg = 0.f;
#pragma unroll
      for (int c = 0; c < kAccCols / 16; ++c) {
        if (16 * c < cw) {
          const int col0 = n0 + cs + 16 * c;
          if (MODE == 2) {
            if (row_ok) {
              float* dst = p.out + (size_t)row * p.out_pitch + col0;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (col0 + j + 3 < p.N) {
                  *reinterpret_cast<float4*>(dst + j) = make_float4(acc[16 * c + j], acc[16 * c + j + 1], acc[16 * c + j + 2], acc[16 * c + j + 3]);
                } else {
                  for (int jj = j; jj < j + 4; ++jj)
                    if (col0 + jj < p.N) dst[jj] = acc[16 * c + jj];
                }
            }
          } else {
            const int bin0 = col0 >> 1;
            float mag[8], gq[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float re = acc[16 * c + 2 * j], im = acc[16 * c + 2 * j + 1];
              const float s2 = fmaf(re, re, im * im);
              const float rs = s2 > 0.f ? rsqrtf(s2) : 0.f;  // 1 / |S|; 0 at the origin, where torch's abs() backward is 0 too
              const float sp = s2 * rs;
              mag[j] = sp;
              gq[2 * j] = gq[2 * j + 1] = 0.f;
              if (MODE == 1 && row_ok && bin0 + j < nbins) {
                const float st = sst[trow * kSst + (cs >> 1) + 8 * c + j];
                const float dl = sp - st;
                // log2 through the special-function unit (absolute error ~2^-22): the loss averages millions of these
                const float dg = __log2f(st + p.eps) - __log2f(sp + p.eps);
                lin += fabsf(dl);
                lg += fabsf(dg);
                // d/dSp [ |Sp - St| + alpha |log2(St+eps) - log2(Sp+eps)| ]
                const float sgn_l = dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f);
                const float sgn_g = dg > 0.f ? -1.f : (dg < 0.f ? 1.f : 0.f);
                const float gs = fmaf(p.alpha * sgn_g * 1.4426950408889634f, __frcp_rn(sp + p.eps), sgn_l);
                const float inv = gs * rs;
                gq[2 * j] = inv * re, gq[2 * j + 1] = inv * im;
              }
            }
            if (MODE == 1 && row_ok) {  // the gradient buffer's pitch is a multiple of 32 columns: whole 16-byte groups fit
              float* dst = p.out + (size_t)row * p.out_pitch + col0;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (col0 + j < p.out_pitch) *reinterpret_cast<float4*>(dst + j) = make_float4(gq[j], gq[j + 1], gq[j + 2], gq[j + 3]);
            }
            if (MODE == 0 && row_ok) {  // out pitch is a multiple of 4: whole 16-byte groups inside the pitch are written
              float* dst = p.out + (size_t)row * p.out_pitch + bin0;
              if (bin0 + 3 < p.out_pitch) *reinterpret_cast<float4*>(dst) = make_float4(mag[0], mag[1], mag[2], mag[3]);
              if (bin0 + 7 < p.out_pitch) *reinterpret_cast<float4*>(dst + 4) = make_float4(mag[4], mag[5], mag[6], mag[7]);
            }
          }
        }
      }
      lin_d += (double)lin, lg_d += (double)lg;  // float partial sums only within one tile row
    }
    if (MODE == 1) {
      for (int o = 16; o; o >>= 1) {
        lin_d += __shfl_xor_sync(0xffffffffu, lin_d, o);
        lg_d += __shfl_xor_sync(0xffffffffu, lg_d, o);
      }
      if (lane == 0) red[2 * (warp - 8)] = lin_d, red[2 * (warp - 8) + 1] = lg_d;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 1 && threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) a += red[2 * w], b += red[2 * w + 1];
    p.loss_acc[2 * blockIdx.x] = a, p.loss_acc[2 * blockIdx.x + 1] = b;
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kAccBufs * kTmemCols) : "memory");
}

// ---- framing / overlap-add -----------------------------------------------------------------------------------------------
// frames[b * nfr + f][c] = xpad[f * hop + c] * w[c], xpad = reflect padding by n_fft / 2 (torch.stft center=True);
// blockIdx.z selects the signal (0: x0 -> fr0, 1: x1 -> fr1); a thread writes 4 consecutive columns of kFrRows rows
constexpr int kFrRows = 8;
__global__ void __launch_bounds__(128) mss_frames_kernel(const float* __restrict__ x0, int64_t x0_stride, float* __restrict__ fr0,
                                                         const float* __restrict__ x1, int64_t x1_stride, float* __restrict__ fr1,
                                                         int64_t pitch, int L, int n_fft, int hop, int nfr, int rows) {
  const int c0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c0 >= n_fft) return;
  const float* __restrict__ x = blockIdx.z ? x1 : x0;
  const int64_t xs = blockIdx.z ? x1_stride : x0_stride;
  float* __restrict__ fr = blockIdx.z ? fr1 : fr0;
  // periodic Hann (torch.hann_window(n_fft), scipy get_window("hann", n_fft)): 0.5 - 0.5 cos(2 pi c / n_fft)
  const float inv_n = 2.f / (float)n_fft;
  float w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = c0 + i < n_fft ? 0.5f - 0.5f * cospif((float)(c0 + i) * inv_n) : 0.f;
  const int r_end = min(rows, (int)(blockIdx.y + 1) * kFrRows);
  for (int r = blockIdx.y * kFrRows; r < r_end; ++r) {
    const int b = r / nfr, f = r - b * nfr;
    const float* xb = x + (size_t)b * xs;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = f * hop + c0 + i - n_fft / 2;
      if (idx < 0) idx = -idx;
      if (idx >= L) idx = 2 * (L - 1) - idx;
      v[i] = __ldg(xb + min(max(idx, 0), L - 1)) * w[i];
    }
    *reinterpret_cast<float4*>(fr + (size_t)r * pitch + c0) = make_float4(v[0], v[1], v[2], v[3]);  // pitch % 32 == 0: in bounds
  }
}

// d_x[b][t] (+)= scale * sum over padded positions P that read x[t] and frames f covering P of dfr[b*nfr+f][P - f*hop] * w
__global__ void mss_ola_kernel(const float* __restrict__ dfr, int64_t pitch, float* __restrict__ dx, int64_t dx_stride, int B, int L,
                               int n_fft, int hop, int nfr, float scale, int accumulate) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= L) return;
  const int pad = n_fft / 2;
  int P[3];
  int np = 0;
  P[np++] = t + pad;
  if (t >= 1 && t <= pad) P[np++] = pad - t;
  if (t <= L - 2 && t >= L - 1 - pad) P[np++] = pad + 2 * (L - 1) - t;
  float acc = 0.f;
  for (int i = 0; i < np; ++i) {
    const int pp = P[i];
    int f_hi = pp / hop;
    if (f_hi > nfr - 1) f_hi = nfr - 1;
    int f_lo = (pp - n_fft + 1 + hop - 1) / hop;
    if (pp - n_fft + 1 <= 0) f_lo = 0;
    for (int f = f_lo; f <= f_hi; ++f) {
      const int c = pp - f * hop;
      if (c < 0 || c >= n_fft) continue;
      const float w = 0.5f - 0.5f * cospif(2.f * (float)c / (float)n_fft);
      acc = fmaf(__ldg(dfr + ((size_t)b * nfr + f) * pitch + c), w, acc);
    }
  }
  float* dst = dx + (size_t)b * dx_stride + t;
  *dst = accumulate ? *dst + acc * scale : acc * scale;
}

// DFT bases: fwd[N rows][Kp]: row 2k = cos(2 pi k n / n_fft), row 2k+1 = -sin(.), n along the row (K-major B operand of the
// forward GEMM); bwd[n_fft rows][Kb]: row n, column 2k = cos, 2k+1 = -sin (K-major B operand of the adjoint GEMM)
__global__ void mss_tables_kernel(float* __restrict__ fwd, int64_t kp, float* __restrict__ bwd, int64_t kb, int n_fft) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (n >= n_fft) return;
  const long long m = ((long long)k * n) % n_fft;
  double s, c;
  sincospi(2.0 * (double)m / (double)n_fft, &s, &c);
  fwd[(size_t)(2 * k) * kp + n] = (float)c;
  fwd[(size_t)(2 * k + 1) * kp + n] = (float)(-s);
  bwd[(size_t)n * kb + 2 * k] = (float)c;
  bwd[(size_t)n * kb + 2 * k + 1] = (float)(-s);
}

struct MssCounts {
  double inv[kMssMaxScales];
  int grid[kMssMaxScales];
};
constexpr int kMaxGrid = 256;  // per-CTA partial sums kept per scale (the persistent grid is at most one CTA per SM)
__global__ void __launch_bounds__(32) mss_finish_kernel(const double* __restrict__ acc, MssCounts cnt, int n_scales, float alpha,
                                                        float ratio, float* __restrict__ loss) {
  double tot = 0.0;
  for (int s = 0; s < n_scales; ++s) {
    double lin = 0.0, lg = 0.0;
    for (int c = threadIdx.x; c < cnt.grid[s]; c += 32) lin += acc[(size_t)(s * kMaxGrid + c) * 2], lg += acc[(size_t)(s * kMaxGrid + c) * 2 + 1];
    tot += (lin + (double)alpha * lg) * cnt.inv[s];
  }
  for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (threadIdx.x == 0) loss[0] = (float)(tot * ratio);
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [rows x cols] fp32, row pitch in floats (multiple of 4), box {32 floats, box_rows}, 128-byte swizzle, zero fill outside
static int make_map(CUtensorMap* tm, const float* ptr, int rows, int cols, int64_t pitch, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return GOLF_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GOLF_OK : GOLF_ERR_CUDA;
}

static int gemm_grid(const MssGemmParams& p) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms > kMaxGrid) sms = kMaxGrid;
  const int n_tiles = ceil_div(p.N, p.bn) * ceil_div(p.M, kBM);
  return n_tiles < sms ? n_tiles : sms;
}

static int launch_gemm(const float* A, int64_t pitchA, const float* Bt, int64_t pitchB, MssGemmParams p, cudaStream_t st) {
  if (p.bn % 16 != 0 || p.bn < 16 || p.bn > kMaxBN || (pitchA & 3) || (pitchB & 3)) return GOLF_ERR_INVALID;
  if (((uintptr_t)A & 15) || ((uintptr_t)Bt & 15)) return GOLF_ERR_INVALID;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, A, p.M, p.K, pitchA, kBM);
  if (rc) return rc;
  rc = make_map(&tmB, Bt, p.N, p.K, pitchB, p.bn);
  if (rc) return rc;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    GOLF_CUDA(cudaFuncSetAttribute(mss_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    GOLF_CUDA(cudaFuncSetAttribute(mss_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    GOLF_CUDA(cudaFuncSetAttribute(mss_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    mark_used_on_device(attr);
  }
  const int grid = gemm_grid(p);
  if (p.mode == 0) mss_gemm_kernel<0><<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, p);
  else if (p.mode == 1) mss_gemm_kernel<1><<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, p);
  else mss_gemm_kernel<2><<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, p);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

struct MssScale {
  int n_fft, hop, nbins, nbp, N, Kp, Kb, nfr, rows, bn_f, bn_b;
};
static bool mss_scale(int n_fft, int hop, int B, int L, MssScale* s) {
  if (n_fft < 16 || n_fft > 4096 || L <= n_fft / 2) return false;  // reflect padding needs pad < L
  s->n_fft = n_fft;
  s->hop = hop > 0 ? hop : (int)(n_fft - n_fft * 0.75);  // loss/spec.py:57: hop_length = int(n_fft - n_fft * overlap)
  if (s->hop > n_fft) return false;
  s->nbins = n_fft / 2 + 1;
  s->N = 2 * s->nbins;
  s->nbp = (s->nbins + 3) / 4 * 4;  // row pitch of the magnitude buffer
  s->Kp = (n_fft + kBK - 1) / kBK * kBK;
  s->Kb = (s->N + kBK - 1) / kBK * kBK;
  s->nfr = 1 + L / s->hop;
  s->rows = B * s->nfr;
  const int m_tiles = ceil_div(s->rows, kBM);
  auto pick = [m_tiles](int n) {  // N tile (multiple of 16): useful fraction of the work of the last wave x of the padded columns
    int best = kMaxBN;
    double best_eff = 0.0;
    for (int bn = kMaxBN; bn >= 128; bn -= 16) {
      const int nt = ceil_div(n, bn), tiles = nt * m_tiles;
      const double eff = ((double)n / (nt * bn)) * ((double)tiles / (ceil_div(tiles, 148) * 148));
      if (eff > best_eff + 1e-9) best_eff = eff, best = bn;
    }
    return best;
  };
  s->bn_f = pick(s->N);
  s->bn_b = pick(n_fft);
  return true;
}
static size_t tables_floats(const MssScale& s) { return (size_t)s.N * s.Kp + (size_t)s.n_fft * s.Kb; }

}  // namespace golf

using namespace golf;

GOLF_API size_t golf_mss_tables_bytes(int n_fft) {
  MssScale s;
  if (!mss_scale(n_fft, 0, 1, 1 << 20, &s)) return 0;
  return align_up(tables_floats(s) * sizeof(float), 256);
}

GOLF_API int golf_mss_build_tables(int n_fft, float* tables, void* stream) {
  MssScale s;
  if (!tables || !mss_scale(n_fft, 0, 1, 1 << 20, &s)) return GOLF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  GOLF_CUDA(cudaMemsetAsync(tables, 0, tables_floats(s) * sizeof(float), st));
  float* fwd = tables;
  float* bwd = tables + (size_t)s.N * s.Kp;
  mss_tables_kernel<<<dim3(ceil_div(n_fft, 128), s.nbins), 128, 0, st>>>(fwd, s.Kp, bwd, s.Kb, n_fft);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

GOLF_API size_t golf_mss_workspace_bytes(int B, int L, const int* n_ffts, const int* hops, int n_scales) {
  if (!n_ffts || n_scales <= 0 || n_scales > kMssMaxScales || B <= 0) return 0;
  size_t frames = 0, strue = 0, g = 0;
  for (int i = 0; i < n_scales; ++i) {
    MssScale s;
    if (!mss_scale(n_ffts[i], hops ? hops[i] : 0, B, L, &s)) return 0;
    frames = std::max(frames, (size_t)s.rows * s.Kp);
    strue = std::max(strue, (size_t)s.rows * s.nbp);
    g = std::max(g, (size_t)s.rows * s.Kb);
  }
  // frames(pred) | frames(true) / d_frames | S_true | G | accumulators
  return align_up(frames * 4, 256) * 2 + align_up(strue * 4, 256) + align_up(g * 4, 256) + (size_t)kMssMaxScales * kMaxGrid * 2 * sizeof(double) + 1024;
}

// loss[0] = ratio * sum_s ( mean|Sp - St| + alpha * mean|log2(St+eps) - log2(Sp+eps)| );  d_pred (optional) = d loss / d pred.
// pred, target: [B, L] (row strides given); tables[i]: golf_mss_build_tables(n_ffts[i]).
GOLF_API int golf_mss_loss(const float* pred, int64_t pred_stride, const float* target, int64_t target_stride, int B, int L,
                           const int* n_ffts, const int* hops, int n_scales, const float* const* tables, float alpha, float ratio, float eps,
                           float* loss, float* d_pred, int64_t dpred_stride, int prec3, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (!pred || !target || !n_ffts || !tables || !loss || B <= 0 || L <= 0 || n_scales <= 0 || n_scales > kMssMaxScales)
    return GOLF_ERR_INVALID;
  const size_t need = golf_mss_workspace_bytes(B, L, n_ffts, hops, n_scales);
  if (need == 0) return GOLF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) return GOLF_ERR_WORKSPACE;
  if (B > 65535) return GOLF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  size_t frames = 0, strue = 0, g = 0;
  MssScale sc[kMssMaxScales];
  for (int i = 0; i < n_scales; ++i) {
    mss_scale(n_ffts[i], hops ? hops[i] : 0, B, L, &sc[i]);
    frames = std::max(frames, (size_t)sc[i].rows * sc[i].Kp);
    strue = std::max(strue, (size_t)sc[i].rows * sc[i].nbp);
    g = std::max(g, (size_t)sc[i].rows * sc[i].Kb);
  }
  char* ws = reinterpret_cast<char*>(workspace);
  float* fr_p = reinterpret_cast<float*>(ws);
  float* fr_t = reinterpret_cast<float*>(ws + align_up(frames * 4, 256));  // target frames, later the frame gradients
  float* s_true = reinterpret_cast<float*>(ws + 2 * align_up(frames * 4, 256));
  float* G = reinterpret_cast<float*>(ws + 2 * align_up(frames * 4, 256) + align_up(strue * 4, 256));
  double* acc = reinterpret_cast<double*>(ws + 2 * align_up(frames * 4, 256) + align_up(strue * 4, 256) + align_up(g * 4, 256));
  MssCounts cnt{};
  for (int i = 0; i < n_scales; ++i) cnt.inv[i] = 1.0 / ((double)sc[i].rows * sc[i].nbins);
  for (int i = 0; i < n_scales; ++i) {
    const MssScale& s = sc[i];
    const float* fwd = tables[i];
    const float* bwd = tables[i] + (size_t)s.N * s.Kp;
    const dim3 fgrid(ceil_div(s.n_fft, 512), ceil_div(s.rows, kFrRows), 2);
    mss_frames_kernel<<<fgrid, 128, 0, st>>>(target, target_stride, fr_t, pred, pred_stride, fr_p, s.Kp, L, s.n_fft, s.hop, s.nfr, s.rows);
    GOLF_CHECK_LAUNCH();
    MssGemmParams p{};
    p.M = s.rows, p.N = s.N, p.K = s.n_fft, p.bn = s.bn_f, p.prec3 = (prec3 & 1) ? 1 : 0, p.chunk_kb = 4, p.alpha = alpha, p.eps = eps;
    p.mode = 0, p.out = s_true, p.out_pitch = s.nbp;
    int rc = launch_gemm(fr_t, s.Kp, fwd, s.Kp, p, st);
    if (rc) return rc;
    p.mode = 1, p.out = G, p.out_pitch = s.Kb, p.s_true = s_true, p.st_pitch = s.nbp, p.loss_acc = acc + (size_t)i * kMaxGrid * 2;
    cnt.grid[i] = gemm_grid(p);
    rc = launch_gemm(fr_p, s.Kp, fwd, s.Kp, p, st);
    if (rc) return rc;
    if (d_pred) {
      MssGemmParams b{};
      b.M = s.rows, b.N = s.n_fft, b.K = s.N, b.bn = s.bn_b, b.prec3 = (prec3 & 2) ? 1 : 0, b.chunk_kb = (prec3 & 2) ? 4 : 16, b.mode = 2, b.out = fr_t, b.out_pitch = s.Kp;
      rc = launch_gemm(G, s.Kb, bwd, s.Kb, b, st);
      if (rc) return rc;
      mss_ola_kernel<<<dim3(ceil_div(L, 256), B), 256, 0, st>>>(fr_t, s.Kp, d_pred, dpred_stride, B, L, s.n_fft, s.hop, s.nfr,
                                                               (float)(ratio * cnt.inv[i]), i > 0 ? 1 : 0);
      GOLF_CHECK_LAUNCH();
    }
  }
  mss_finish_kernel<<<1, 32, 0, st>>>(acc, cnt, n_scales, alpha, ratio, loss);
  GOLF_CHECK_LAUNCH();
  return GOLF_OK;
}

// D [M x N] = A [M x K] . Bt [N x K]^T on the same kernel (plain-store epilogue): bring-up / test entry point
GOLF_API int golf_mss_gemm(const float* A, int64_t pitchA, const float* Bt, int64_t pitchB, float* D, int64_t pitchD, int M, int N,
                           int K, int bn, int prec3, void* stream) {
  if (!A || !Bt || !D || M <= 0 || N <= 0 || K <= 0 || (pitchD & 3) || ((uintptr_t)D & 15)) return GOLF_ERR_INVALID;
  MssGemmParams p{};
  p.M = M, p.N = N, p.K = K, p.bn = bn, p.prec3 = prec3 ? 1 : 0, p.chunk_kb = prec3 ? 4 : 16, p.mode = 2, p.out = D, p.out_pitch = pitchD;
  return launch_gemm(A, pitchA, Bt, pitchB, p, (cudaStream_t)stream);
}
