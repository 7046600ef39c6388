"""Block-wise (streaming) synthesis with the GOLF-ss decoder: arbitrarily long utterances in bounded memory.

The reference synthesises long files by cross-fading independently decoded windows
(ltng/vocoder.py:350-383).  Every stage of the decoder except two is a finite-memory map of its inputs, and the
two recursive ones have a small explicit state, so a stream can instead be continued *exactly*:

  oscillator   running phase: carried as a float64 sum of the knot-interval closed forms and handed to the
               kernel as its initial phase (golf_glottal_osc_fwd_from; `phase_offset` of models/synth.py:251);
               the 4x decimator's +-16-sample support and the table-row interpolation are covered by running
               the oscillator over one extra table frame (w_hop samples) on each side and keeping the middle
  noise FIR    +-255 samples of noise around the block: the noise of a block is drawn once and kept until
               both neighbours have used it; the block FIR runs over two extra hops on each side
  LPC filter   the last M output samples (zi of torchlpc.sample_wise_lpc, models/filters.py:112)
  room FIR     the last 127 filtered samples

One block of look-ahead is needed (next knot / next control frame / noise to the right), so push() returns the
waveform of the PREVIOUS block and flush() the last one.  Concatenated, the blocks equal the one-shot decoder
on the same controls and the same noise to float32 rounding (tests/test_gpu_streaming.py); the one-shot
decoder drops its final hop (the reference's block FIR has no look-ahead there), the stream does not.

Inference only (no autograd), GOLF-ss configuration (SourceFilterSynth with IndexedGlottalFlowTable,
LTVZeroPhaseFIRFilter, LTVMinimumPhaseFilterPrecise, LTIAcousticFilter, subtract_harmonics=False).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import functional as G
from ._lib import GolfError
from .filters import LTIAcousticFilter, LTVMinimumPhaseFilterPrecise, LTVZeroPhaseFIRFilter
from .synth import IndexedGlottalFlowTable


class _Block:
    __slots__ = ("phase", "w", "log_mag", "gain", "a", "noise", "n")

    def __init__(self, phase, w, log_mag, gain, a, noise):
        self.phase, self.w, self.log_mag, self.gain, self.a, self.noise = phase, w, log_mag, gain, a, noise
        self.n = noise.shape[1]


class StreamingSynth:
    def __init__(self, decoder, hop: int = 240, w_hop: int = 2400, phase_hop: int = 1,
                 noise_fn: Optional[Callable[[int, int, torch.device], torch.Tensor]] = None):
        osc, nf, ef, rf = decoder.harm_oscillator, decoder.noise_filter, decoder.end_filter, decoder.room_filter
        if not (isinstance(osc, IndexedGlottalFlowTable) and type(nf) is LTVZeroPhaseFIRFilter
                and type(ef) is LTVMinimumPhaseFilterPrecise and isinstance(rf, LTIAcousticFilter)) or decoder.subtract_harmonics:
            raise GolfError("StreamingSynth: GOLF-ss decoder configuration expected")
        if osc.phase_accumulation != "exact":
            raise GolfError("StreamingSynth: the carried phase needs phase_accumulation = 'exact'")
        if w_hop % hop or w_hop % phase_hop or hop % phase_hop and phase_hop % hop:
            raise GolfError("StreamingSynth: hop and phase_hop must divide w_hop")
        self.dec, self.hop, self.w_hop, self.phase_hop = decoder, hop, w_hop, phase_hop
        self.noise_fn = noise_fn or (lambda b, n, dev: torch.randn(b, n, dtype=torch.float32, device=dev))
        self.prev: Optional[_Block] = None   # block p-1 (already emitted; context only)
        self.cur: Optional[_Block] = None    # block p (waiting for its look-ahead)
        self.cum = None        # [B] float64: running phase (cycles) before the first oversampled sample of block p
        self.zi = None         # [B, M] last LPC outputs, most recent first
        self.room_hist = None  # [B, 127] last LPC outputs in time order (room FIR history)
        self.emitted = 0

    # -------------------------------------------------------------------------------- helpers
    def _phase_sum(self, knots: torch.Tensor) -> torch.Tensor:
        """sum of the oversampled phase increments over the k intervals of k+1 knots, float64, from the closed form
        of every knot interval: hp*x_k + (x_{k+1} - x_k)(hp - 1)/2 with x = phase/os, hp = phase_hop*os"""
        os_ = self.dec.harm_oscillator.oversampling
        hp = self.phase_hop * os_
        x = knots.to(torch.float64) / os_
        return (hp * x[:, :-1] + (x[:, 1:] - x[:, :-1]) * ((hp - 1) / 2.0)).sum(1)

    def _emit(self, nxt: Optional[_Block], w_end: Optional[torch.Tensor] = None) -> torch.Tensor:
        osc, nf, ef, rf = self.dec.harm_oscillator, self.dec.noise_filter, self.dec.end_filter, self.dec.room_filter
        prev, cur = self.prev, self.cur
        B, n, dev = cur.noise.shape[0], cur.n, cur.noise.device
        hop, w_hop, ph = self.hop, self.w_hop, self.phase_hop
        kh = w_hop // ph  # knots per table frame
        K = 2 * (cur.log_mag.shape[-1] - 1)
        hf = -(-(K // 2 + 1) // hop)  # control frames of FIR context on each side (>= K/2 + 1 samples)
        halo = hf * hop
        if halo > w_hop:
            raise GolfError("StreamingSynth: the noise FIR reaches beyond one table frame")

        # ---- oscillator over [s0 - w_hop, s1 + w_hop) (what exists of it), initial phase carried
        parts_p = ([prev.phase[:, -kh:]] if prev else []) + [cur.phase] + ([nxt.phase[:, : kh + 1]] if nxt else [])
        w_tail = nxt.w[:, :2] if nxt is not None else (w_end if w_end is not None else cur.w[:, -1:])
        parts_w = ([prev.w[:, -1:]] if prev else []) + [cur.w, w_tail]
        if nxt is not None and nxt.w.shape[1] < 2:  # the row after the right halo: replicate
            parts_w.append(nxt.w[:, -1:])
        phase_ext, w_ext = torch.cat(parts_p, 1).contiguous(), torch.cat(parts_w, 1).contiguous()
        if self.cum is None:
            self.cum = torch.zeros(B, dtype=torch.float64, device=dev)
        # self.cum: running phase before block p; the oscillator starts one table frame earlier when there is one
        phase0 = self.cum - self._phase_sum(torch.cat([prev.phase[:, -kh:], cur.phase[:, :1]], 1)) if prev else self.cum
        dk = osc.decimater.kernel if osc.oversampling > 1 else None
        harm_ext = G.glottal_osc(phase_ext, ph, w_ext, w_hop, osc.table, dk, osc.oversampling, osc.equal_energy, "exact", phase0)
        lead = w_hop if prev else 0  # samples of left halo in harm_ext
        # harm over [s0 - halo, s1 + halo) for the fused add (zeros where the stream has not started / has ended)
        harm_fir = torch.zeros(B, n + 2 * halo, dtype=torch.float32, device=dev)
        lo = lead - halo
        src_lo, dst_lo = max(lo, 0), max(-lo, 0)
        take = min(harm_ext.shape[1] - src_lo, n + 2 * halo - dst_lo)
        harm_fir[:, dst_lo : dst_lo + take] = harm_ext[:, src_lo : src_lo + take]
        if nxt is None:  # end of stream: nothing sounds beyond s1 (the one-shot decoder's zero padding)
            harm_fir[:, halo + n :] = 0.0

        # ---- noise FIR (+ harm) over the same range, kernels of control frames f0-hf .. f1+hf
        zeros = torch.zeros(B, halo, dtype=torch.float32, device=dev)
        noise_ext = torch.cat([prev.noise[:, -halo:] if prev else zeros, cur.noise, nxt.noise[:, :halo] if nxt else zeros], 1)
        lm = cur.log_mag
        lm_ext = torch.cat([prev.log_mag[:, -hf:] if prev else lm[:, :1].expand(-1, hf, -1), lm,
                            nxt.log_mag[:, : hf + 1] if nxt else lm[:, -1:].expand(-1, hf + 1, -1)], 1).contiguous()
        raw = nf.raw_kernels(lm_ext)
        src_ext = G.ltv_fir_blocks(noise_ext.contiguous(), raw, hop, add=harm_fir, window=nf._window(raw.shape[-1], raw))
        src = src_ext[:, halo : halo + n].contiguous()

        # ---- LPC from the carried state; frame f1 comes from the look-ahead (replicated at the end)
        gain = torch.cat([cur.gain, nxt.gain[:, :1] if nxt else cur.gain[:, -1:]], 1).contiguous()
        a = torch.cat([cur.a, nxt.a[:, :1] if nxt else cur.a[:, -1:]], 1).contiguous()
        y = G.lpc_ss(src, gain, a, hop, zi=self.zi)
        M = a.shape[-1]
        nk = rf.kernel.shape[0]
        keep = max(nk, M)
        hist = y if self.room_hist is None else torch.cat([self.room_hist, y], 1)
        if hist.shape[1] < keep:  # a first block shorter than the histories: the stream started from rest
            hist = torch.nn.functional.pad(hist, (keep - hist.shape[1], 0))
        self.zi = hist[:, -M:].flip(1).contiguous()

        # ---- room FIR with its history prepended
        if self.room_hist is None:
            out = G.room_fir(y.contiguous(), rf.kernel)
        else:
            out = G.room_fir(torch.cat([self.room_hist, y], 1).contiguous(), rf.kernel)[:, self.room_hist.shape[1] :]
        self.room_hist = hist[:, -keep:].contiguous()

        # ---- the carried phase moves past block p
        if nxt is not None:
            self.cum = self.cum + self._phase_sum(torch.cat([cur.phase, nxt.phase[:, :1]], 1))
            self.cum = self.cum - torch.floor(self.cum)
        self.emitted += n
        return out

    # -------------------------------------------------------------------------------- API
    @torch.no_grad()
    def push(self, phase, w, log_mag, gain, a) -> Optional[torch.Tensor]:
        """One block of n table frames (n*w_hop samples): phase [B, n*w_hop/phase_hop] (cycles per sample), w [B, n],
        log_mag [B, n*w_hop/hop, n_mag], gain [B, n*w_hop/hop], a [B, n*w_hop/hop, M].  Returns the waveform of
        the previous block [B, its samples], or None on the first call."""
        n = w.shape[1] * self.w_hop
        if phase.shape[1] * self.phase_hop != n or gain.shape[1] * self.hop != n or a.shape[1] != gain.shape[1] or log_mag.shape[1] != gain.shape[1]:
            raise GolfError("StreamingSynth.push: a block is a whole number of table frames in every control")
        blk = _Block(phase, w, log_mag, gain, a, self.noise_fn(w.shape[0], n, w.device))
        out = None
        if self.cur is not None:
            out = self._emit(blk)
            self.prev = self.cur
        self.cur = blk
        return out

    @torch.no_grad()
    def flush(self, w_end: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """the last block (zero look-ahead, like the end of a one-shot call); the stream is reset.  w_end [B,1]:
        the table weight that closes the last table frame (the reference's encoder emits T/w_hop + 1 of them);
        default: the last one is held."""
        out = self._emit(None, w_end) if self.cur is not None else None
        self.prev = self.cur = self.cum = self.zi = self.room_hist = None
        return out


def synthesize_long(decoder, phase, w, log_mag, gain, a, block_frames: int = 10, **kw) -> torch.Tensor:
    """whole-utterance convenience wrapper: cut the controls into blocks of `block_frames` table frames, stream
    them, concatenate.  phase [B, T/phase_hop], w [B, T/w_hop (+1 closing frame)], log_mag/gain/a at the control
    hop (T/hop frames)."""
    st = StreamingSynth(decoder, **kw)
    kh, fh = st.w_hop // st.phase_hop, st.w_hop // st.hop
    n_w = gain.shape[1] // fh  # whole table frames
    w_end = w[:, n_w : n_w + 1] if w.shape[1] > n_w else None
    outs: List[torch.Tensor] = []
    for f in range(0, n_w, block_frames):
        g = min(block_frames, n_w - f)
        o = st.push(phase[:, f * kh : (f + g) * kh], w[:, f : f + g], log_mag[:, f * fh : (f + g) * fh], gain[:, f * fh : (f + g) * fh],
                    a[:, f * fh : (f + g) * fh])
        if o is not None:
            outs.append(o)
    o = st.flush(w_end)
    if o is not None:
        outs.append(o)
    return torch.cat(outs, 1)
