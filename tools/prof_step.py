"""Driver for ncu: a few full GOLF-ss decoder steps at the bench shape (device-resident inputs)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import synth as gsynth
from golf_b200.audiotensor import AudioTensor
dev = torch.device("cuda:0")
dec = bench.build_decoder(dev)
gsynth.CHECK_INPUTS = "off"
sets = [{k: v.to(dev) for k, v in s.items()} for s in bench.make_inputs(2, bench.BATCH)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for i in range(n):
        s = sets[i % 2]
        dec(phase=AudioTensor(s["phase"], hop_length=1), harm_oscillator_params=(AudioTensor(s["w"], hop_length=2400),),
            noise_generator_params=(), noise_filter_params=(AudioTensor(s["log_mag"], hop_length=bench.HOP),),
            end_filter_params=(AudioTensor(s["gain"], hop_length=bench.HOP), AudioTensor(s["a"], hop_length=bench.HOP)))
torch.cuda.synchronize()
print("done")
