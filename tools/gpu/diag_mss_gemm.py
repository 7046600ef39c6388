"""Bring-up of the tcgen05 GEMM under the MSS loss: D = A . Bt^T against float64, single and compensated products, timing."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from golf_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
def gemm(A, Bt, bn, prec3):
    M, K = A.shape; N = Bt.shape[0]
    pd = (N + 3) // 4 * 4
    D = torch.full((M, pd), float("nan"), device=dev)
    rc = L.golf_mss_gemm(A.data_ptr(), A.stride(0), Bt.data_ptr(), Bt.stride(0), D.data_ptr(), pd, M, N, K, bn, prec3, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, (rc, L.golf_last_cuda_error())
    torch.cuda.synchronize()
    return D[:, :N]
def padded(rows, cols, gen):
    p = (cols + 31) // 32 * 32
    t = torch.zeros(rows, p, device=dev)
    t[:, :cols] = torch.randn(rows, cols, generator=gen, device=dev)
    return t[:, :cols]  # a view with pitch p
g = torch.Generator(device=dev).manual_seed(0)
for (M, N, K, bn) in ((128, 16, 32, 16), (128, 112, 32, 112), (128, 64, 64, 64), (300, 80, 100, 80), (1000, 510, 509, 176), (3008, 2054, 2053, 208), (3008, 2053, 2054, 208)):
    A, Bt = padded(M, K, g), padded(N, K, g)
    ref = A.double() @ Bt.double().T
    for prec3 in (0, 1):
        D = gemm(A, Bt, bn, prec3)
        err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
        nan = int(torch.isnan(D).sum())
        print(f"M={M} N={N} K={K} bn={bn} prec3={prec3}: max err / max |ref| = {err:.2e}  nan={nan}", flush=True)
A, Bt = padded(3008, 2053, g), padded(2054, 2053, g)
pd = 2056
D = torch.empty(3008, pd, device=dev)
st = torch.cuda.current_stream().cuda_stream
for prec3 in (0, 1):
    for _ in range(3): L.golf_mss_gemm(A.data_ptr(), A.stride(0), Bt.data_ptr(), Bt.stride(0), D.data_ptr(), pd, 3008, 2054, 2053, 208, prec3, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.golf_mss_gemm(A.data_ptr(), A.stride(0), Bt.data_ptr(), Bt.stride(0), D.data_ptr(), pd, 3008, 2054, 2053, 208, prec3, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2 * 3008 * 2054 * 2053 * (3 if prec3 else 1)
    print(f"3008 x 2054 x 2053 prec3={prec3}: {ms*1e3:.1f} us, {fl/ms/1e9:.0f} TFLOP/s of tensor work ({2*3008*2054*2053/ms/1e9:.0f} useful)")
