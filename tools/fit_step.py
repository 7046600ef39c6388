"""Config 4 of BASELINE.json, decoder side: one training step through the GOLF-ss decoder on synthetic
24 kHz batches -- forward (autograd path, CUDA forward kernels), multi-scale spectral loss
(loss/spec.py:11-67 restated with torch.stft: n_fft 509/1021/2053, 75 % overlap, L1 + log2-L1), backward
through the CUDA adjoints to the controls the encoder would produce (gain, a, log_mag, table weight) and to
the decoder's own parameters (room kernel), then -- with more than one rank -- the gradient all-reduce
DDP would do (NCCL), for the decoder's parameters plus a 6.08 M-parameter stand-in for the encoder's
(autoencode.py:9-16; the encoder itself is the reference's torch U-Net and stays out of scope).

    python tools/fit_step.py [steps] [ss|ff]               # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/fit_step.py

Prints one JSON line: samples/s over all ranks, ms per step (max over ranks), and the split.
"""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import synth as gsynth
from golf_b200.audiotensor import AudioTensor

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variant = sys.argv[2] if len(sys.argv) > 2 else "ss"  # "ff": GOLF-ff (frame-wise filter), cfg/ae/decoder/golf.yaml
dec = bench.build_decoder(dev, variant).train()
gsynth.CHECK_INPUTS = "off"
s = {k: v.to(dev) for k, v in bench.make_inputs(1, bench.BATCH, seed=2434 + rank)[0].items()}
leaves = {k: s[k].clone().requires_grad_() for k in ("w", "log_mag", "gain", "a")}
target = torch.randn(bench.BATCH, bench.T, device=dev) * 0.05
encoder_grads = torch.zeros(6_080_000, device=dev)  # stand-in bucket for the encoder's gradients
windows = {n: torch.hann_window(n, device=dev) for n in (509, 1021, 2053)}


def mss(pred, true):
    loss = 0.0
    for n, win in windows.items():
        sp, st = (torch.stft(x, n, hop_length=int(n - n * 0.75), window=win, return_complex=True).abs() for x in (pred, true))
        loss = loss + (sp - st).abs().mean() + ((st + 1e-8).log2() - (sp + 1e-8).log2()).abs().mean()
    return loss


def step():
    for p in list(leaves.values()) + list(dec.parameters()):
        p.grad = None
    y = dec(phase=AudioTensor(s["phase"], hop_length=1), harm_oscillator_params=(AudioTensor(leaves["w"], hop_length=2400),),
            noise_generator_params=(), noise_filter_params=(AudioTensor(leaves["log_mag"], hop_length=bench.HOP),),
            end_filter_params=(AudioTensor(leaves["gain"], hop_length=bench.HOP), AudioTensor(leaves["a"], hop_length=bench.HOP))).as_tensor()
    ev[1].record()
    loss = mss(y, target[:, : y.shape[1]])
    ev[2].record()
    loss.backward()
    ev[3].record()
    if world > 1:
        flat = torch.cat([p.grad.flatten() for p in dec.parameters() if p.grad is not None] + [encoder_grads])
        dist.all_reduce(flat)
    return loss


ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
tot = [0.0] * 4
for _ in range(steps):
    ev[0].record()
    loss = step()
    ev[4].record()
    torch.cuda.synchronize()
    for i in range(4):
        tot[i] += ev[i].elapsed_time(ev[i + 1])
ms = sum(tot) / steps
t = torch.tensor([ms], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    assert all(torch.isfinite(p.grad).all() for p in leaves.values())
    print(json.dumps({"workload": f"GOLF-{variant} decoder fit step (fwd + MSS loss + bwd" + (" + NCCL grad all-reduce" if world > 1 else "") + "), 32 x 2 s per GPU, eager",
                      "n_gpus": world, "ms_per_step": float(t), "samples_per_s": world * bench.BATCH * bench.T / (float(t) * 1e-3),
                      "split_ms": {"decoder_fwd": tot[0] / steps, "mss_loss_fwd": tot[1] / steps, "backward": tot[2] / steps, "allreduce": tot[3] / steps},
                      "loss": float(loss)}))
if world > 1:
    dist.destroy_process_group()
