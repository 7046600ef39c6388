"""Config 4 of BASELINE.json: the `autoencode.py fit` training step on synthetic 24 kHz batches, one process per GPU.

Mirrors VoiceAutoEncoder.training_step (ltng/ae.py:86-143) with the shipped GOLF settings (cfg/ae/vctk.yaml:
train_with_true_f0, MSS loss n_fft 509/1021/2053 at 75 % overlap, Adam lr 1e-4, gradient clip 0.5) and Lightning's DDP
strategy (autoencode.py:9-16):

  x [B,48000], f0 [B,48000] (sample rate; unvoiced -> one U(50,500) draw per item, ae.py:96-101)
  encoder(x) -> [B,200,343] logits -> the decoder's own .ctrl transforms (golf_b200 kernels: rc2lpc, downsampler MLP)
  decoder(phase, **params)  -- golf_b200 SourceFilterSynth, autograd path through the CUDA forward kernels
  MSS loss -> backward through the CUDA adjoints and the encoder -> gradient all-reduce (NCCL) -> clip -> Adam

The encoder is a STAND-IN: the reference's U-Net (models/unet.py, 6.08 M parameters) is torch code outside the hot
path and is not available on the GPU box, so a plain MLP over 960-sample frames with the same parameter count (6.07 M)
and the same zero-initialised output layer (models/enc.py:24-27) produces the logits.  What the step measures is
therefore the decoder side exactly, plus a DDP gradient exchange of the right size overlapped with a backward of
comparable shape (torch DistributedDataParallel buckets, as Lightning uses).

    python tools/fit_step.py [steps] [ss|ff] [tcgen05|torch]                                  # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/fit_step.py [steps] [ss|ff] [tcgen05|torch]

Prints one JSON line: samples/s over all ranks, ms per step (max over ranks), the per-phase split, and the time of a
bare 24.3 MB all-reduce on the same communicator (what DDP has to hide).
"""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from golf_b200 import synth as gsynth  # noqa: E402
from golf_b200.audiotensor import AudioTensor  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variant = sys.argv[2] if len(sys.argv) > 2 else "ss"  # "ff": GOLF-ff (frame-wise filter), cfg/ae/decoder/golf.yaml
loss_impl = sys.argv[3] if len(sys.argv) > 3 else "tcgen05"  # "tcgen05-fastbwd": single-TF32 adjoint GEMM; "torch": the torch.stft / cuFFT restatement of loss/spec.py
B, T, SR, HOP = bench.BATCH, bench.T, bench.SR, bench.HOP
FRAMES = T // HOP  # 200: the U-Net's frame count for sample-rate f0 (models/unet.py:160-162)


class EncoderStandIn(nn.Module):
    """[B,T] -> [B,200,out] at hop 240; 6.07 M parameters; output layer zero-initialised like BackboneModelInterface"""

    def __init__(self, out_channels: int, hidden: int = 1080, layers: int = 4, win: int = 960):
        super().__init__()
        self.win = win
        mods, d = [], win
        for _ in range(layers + 1):
            mods += [nn.Linear(d, hidden), nn.GELU()]
            d = hidden
        self.body = nn.Sequential(*mods)
        self.out_linear = nn.Linear(hidden, out_channels)
        nn.init.zeros_(self.out_linear.weight)
        nn.init.zeros_(self.out_linear.bias)

    def forward(self, x):
        pad = (self.win - HOP) // 2
        fr = torch.nn.functional.pad(x, (pad, pad)).unfold(1, self.win, HOP)  # [B,200,960]
        return self.out_linear(self.body(fr))


class AutoEncoder(nn.Module):
    """encoder logits -> split -> .ctrl transforms -> decoder, as VocoderParameterEncoderInterface.forward
    (models/enc.py:73-98) and VoiceAutoEncoder.forward do"""

    def __init__(self, decoder):
        super().__init__()
        self.decoder = decoder
        self.split_sizes, self.trsfms, self.args_keys = decoder.split_sizes_and_trsfms
        self.encoder = EncoderStandIn(sum(sum(self.split_sizes, ())))

    def forward(self, x, phase):
        h = self.encoder(x)
        flat = torch.split(h, [n for grp in self.split_sizes for n in grp], dim=2)
        params, i = {}, 0
        for key, grp, fn in zip(self.args_keys, self.split_sizes, self.trsfms):
            args = [AudioTensor(t.squeeze(2) if t.shape[2] == 1 else t, hop_length=HOP) for t in flat[i : i + len(grp)]]
            params[key] = fn(*args)
            i += len(grp)
        return self.decoder(phase=AudioTensor(phase, hop_length=1), **params).as_tensor()


torch.manual_seed(2434 + rank)
dec = bench.build_decoder(dev, variant).train()
gsynth.CHECK_INPUTS = "off"
model = AutoEncoder(dec).to(dev)
n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
windows = {n: torch.hann_window(n, device=dev) for n in (509, 1021, 2053)}

# synthetic batch: a target waveform and its f0 track (20 % unvoiced), fixed across steps (ltng/data.py shapes)
s = bench.make_inputs(1, B, seed=2434 + rank)[0]
f0 = (s["phase"] * SR).to(dev)
unv = (torch.rand(B, T // 2400, device=dev) < 0.2).repeat_interleave(2400, 1)
f0 = torch.where(unv, torch.zeros_like(f0), f0)
x = (torch.randn(B, T, device=dev) * 0.05).contiguous()


from golf_b200.loss import MSSLoss  # noqa: E402

criterion = MSSLoss([509, 1021, 2053], alpha=1.0, overlap=0.75, window="hann", precision=1 if loss_impl == "tcgen05-fastbwd" else 3)  # cfg/ae/vctk.yaml:58-67 on tcgen05


def mss(pred, true):
    """loss/spec.py:11-67 (Spectrogram power=1, periodic Hann, centre / reflect padding): golf_b200.loss.MSSLoss (DFT-as-GEMM
    on the tensor cores, csrc/mss.cu) or, with `torch` as the third argument, the torch.stft / cuFFT restatement"""
    if loss_impl != "torch":
        return criterion(pred.contiguous(), true.contiguous())
    loss = 0.0
    for n, win in windows.items():
        sp, st = (torch.stft(v, n, hop_length=int(n - n * 0.75), window=win, return_complex=True).abs() for v in (pred, true))
        loss = loss + (sp - st).abs().mean() + ((st + 1e-8).log2() - (sp + 1e-8).log2()).abs().mean()
    return loss


ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]


def step():
    opt.zero_grad(set_to_none=True)
    random_f0 = torch.empty(B, 1, device=dev).uniform_(50, 500)
    phase = torch.where(f0 == 0, random_f0, f0) / SR  # ae.py:96-101
    ev[0].record()
    y = ddp(x, phase)
    ev[1].record()
    loss = mss(y[:, :T], x[:, : y.shape[1]])
    ev[2].record()
    loss.backward()  # DDP: bucketed all-reduce overlapped with the rest of backward
    ev[3].record()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    opt.step()
    ev[4].record()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
tot = [0.0] * 4
for _ in range(steps):
    loss = step()
    torch.cuda.synchronize()
    for i in range(4):
        tot[i] += ev[i].elapsed_time(ev[i + 1])
ms = sum(tot) / steps
t = torch.tensor([ms], device=dev)
allreduce_ms = 0.0
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bucket = torch.zeros(n_params, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        dist.all_reduce(bucket)
    e0.record()
    for _ in range(10):
        dist.all_reduce(bucket)
    e1.record()
    torch.cuda.synchronize()
    allreduce_ms = e0.elapsed_time(e1) / 10
if rank == 0:
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in model.parameters())
    print(json.dumps({
        "workload": f"GOLF-{variant} fit step (stand-in encoder {n_params / 1e6:.2f} M params -> .ctrl -> decoder fwd -> MSS loss -> bwd "
                    f"(CUDA adjoints){' + DDP all-reduce' if world > 1 else ''} -> clip 0.5 -> Adam), {B} x 2 s per GPU, eager",
        "n_gpus": world, "mss_loss": f"golf_b200.loss ({loss_impl} DFT-as-GEMM)" if loss_impl != "torch" else "torch.stft (cuFFT Bluestein)", "ms_per_step": float(t), "samples_per_s": world * B * T / (float(t) * 1e-3),
        "split_ms": {"encoder_decoder_fwd": tot[0] / steps, "mss_loss_fwd": tot[1] / steps, "backward_incl_allreduce": tot[2] / steps,
                     "clip_adam": tot[3] / steps},
        "bare_allreduce_ms": allreduce_ms, "grad_bytes": 4 * n_params, "loss": float(loss)}))
if world > 1:
    dist.destroy_process_group()
