"""Driver for the DRAM-traffic capture: full GOLF-ss decoder passes at the bench shape over 8 rotating input sets (208 MB, more
than the 126 MB L2, as in bench.py), one pass at a time.  Run under
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none
so that every kernel runs ONCE in the cache state the previous kernel left (no flush, no replay): the per-kernel DRAM bytes
then add up to what a decoder pass really moves to and from HBM -- intermediates (harm, src, chunk blocks, y) that the next
kernel finds in L2 are not charged twice.  tools/dram_per_step.py turns the CSV into the per-pass table."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import synth as gsynth
from golf_b200.audiotensor import AudioTensor
dev = torch.device("cuda:0")
fused = os.environ.get("GOLF_BENCH_NOISE", "fused") != "torch"
dec = bench.build_decoder(dev, fused_noise=fused)
gsynth.CHECK_INPUTS = "off"
sets = [{k: v.to(dev) for k, v in s.items()} for s in bench.make_inputs(8, bench.BATCH)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with torch.no_grad():
    for i in range(n):
        s = sets[i % 8]
        dec(phase=AudioTensor(s["phase"], hop_length=1), harm_oscillator_params=(AudioTensor(s["w"], hop_length=2400),),
            noise_generator_params=(), noise_filter_params=(AudioTensor(s["log_mag"], hop_length=bench.HOP),),
            end_filter_params=(AudioTensor(s["gain"], hop_length=bench.HOP), AudioTensor(s["a"], hop_length=bench.HOP)))
torch.cuda.synchronize()
print("done")
