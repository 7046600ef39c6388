"""H2D / D2H copy rate per rank with ordinary pinned host memory vs write-combined pinned memory (cudaHostAllocWriteCombined),
all ranks copying at once: does the host side of the e2e pipeline have headroom?  torchrun or single process."""
import ctypes, os, sys, torch
import torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return torch.frombuffer((ctypes.c_float * (nbytes // 4)).from_address(p.value), dtype=torch.float32)
N_IN, N_OUT = 13289216, 6113280
d_in = torch.empty(N_IN // 4, device=dev); d_out = torch.ones(N_OUT // 4, device=dev)
bufs = {"pinned": torch.zeros(N_IN // 4).pin_memory(), "write-combined": host_alloc(N_IN, 0x04), "portable|wc": host_alloc(N_IN, 0x05)}
h_out = torch.empty(N_OUT // 4).pin_memory()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
for name, h in bufs.items():
    h.fill_(1.0)
    for both in (False, True):
        torch.cuda.synchronize()
        if world > 1: dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_stream(torch.cuda.current_stream()); s_out.wait_stream(torch.cuda.current_stream())
        for _ in range(50):
            with torch.cuda.stream(s_in): d_in.copy_(h, non_blocking=True)
            if both:
                with torch.cuda.stream(s_out): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        t = torch.tensor([ms], device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"N={world} {name:16s} {'H2D + D2H' if both else 'H2D only '}: {float(t):.4f} ms/step  H2D {N_IN / float(t) / 1e6:.1f} GB/s per rank, {world * N_IN / float(t) / 1e6:.1f} GB/s total", flush=True)
if world > 1: dist.destroy_process_group()
