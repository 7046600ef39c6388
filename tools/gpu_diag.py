"""Diagnostic sweep on a GPU box: every kernel vs the CPU oracle, errors printed, nothing asserted.
    gpurun -- 'python tools/gpu_diag.py > gpurun_out/diag.log 2>&1'
"""
import os, sys, time, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from golf_b200 import functional as G
from oracle import golf_oracle as O

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
GD = os.path.join(ROOT, "tests", "golden")

def rel(x, y):
    x = x.detach().double().cpu(); y = y.detach().double().cpu()
    if x.shape != y.shape: return f"SHAPE {tuple(x.shape)} vs {tuple(y.shape)}"
    e = torch.sqrt(((x - y) ** 2).mean(-1) / (y ** 2).mean(-1).clamp_min(1e-300))
    return f"{float(e.max()):.3e}"

def smooth(x, n=16):
    xt = x.transpose(1, -1) if x.ndim > 2 else x
    flat = xt.reshape(-1, 1, xt.shape[-1])
    y = torch.nn.functional.conv1d(torch.nn.functional.pad(flat, (n - 1, 0), mode="replicate"), torch.ones(1, 1, n) / n)
    y = (y * n ** 0.5).view(xt.shape)
    return y.transpose(1, -1) if x.ndim > 2 else y

def controls(B, Fr, M, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = O.rc2lpc(torch.tanh(0.15 * smooth(torch.randn(B, Fr, M, generator=g))))
    gain = torch.exp(smooth(torch.randn(B, Fr, generator=g)) - 6)
    return gain, a

def run(name, fn):
    try:
        t0 = time.time(); fn(); torch.cuda.synchronize()
        print(f"[ok  ] {name} ({time.time()-t0:.2f}s)")
    except Exception:
        print(f"[FAIL] {name}\n{traceback.format_exc()}")
    sys.stdout.flush()

def t_ss():
    for (B, T, H, M) in [(2, 4800, 240, 22), (3, 12000, 240, 22), (2, 12000, 120, 22), (2, 9600, 240, 12), (2, 9600, 240, 20),
                         (2, 9600, 240, 32), (2, 9600, 240, 8), (2, 7000, 240, 22), (1, 500, 240, 22), (2, 9600, 256, 22), (2, 9600, 100, 22)]:
        Fr = T // H + 1
        gain, a = controls(B, Fr, M, seed=M)
        ex = torch.randn(B, T, generator=torch.Generator().manual_seed(1))
        ref = O.lpc_ss_fused(ex, gain, a, H); ref64 = O.lpc_ss_fused(ex, gain, a, H, double=True)
        y = G.lpc_ss(ex.to(dev), gain.to(dev), a.to(dev), H)
        print(f"  ss B{B} T{T} H{H} M{M}: vs oracle32 {rel(y, ref)}  vs f64 {rel(y, ref64)}  (oracle32 vs f64 {rel(ref, ref64)})")

def t_ss_gt():
    g = np.load(os.path.join(GD, "controls_gt.npz"))
    gain, a, H = torch.tensor(g["gain"]), torch.tensor(g["a"]), int(g["hop"])
    ex = torch.randn(gain.shape[0], (gain.shape[1] - 1) * H, generator=torch.Generator().manual_seed(0))
    ref = O.lpc_ss_fused(ex, gain, a, H); ref64 = O.lpc_ss_fused(ex, gain, a, H, double=True)
    y = G.lpc_ss(ex.to(dev), gain.to(dev), a.to(dev), H)
    print(f"  ss encoder-derived: vs oracle32 {rel(y, ref)} vs f64 {rel(y, ref64)} (oracle32 vs f64 {rel(ref, ref64)})")
    for chunk in (480, 120, 960):
        y = G.lpc_ss(ex.to(dev), gain.to(dev), a.to(dev), H, chunk=chunk)
        print(f"     chunk {chunk}: vs oracle32 {rel(y, ref)} vs f64 {rel(y, ref64)}")

def t_swl():
    B, T, M = 2, 3000, 6
    g = torch.Generator().manual_seed(3)
    A = O.rc2lpc(torch.tanh(0.3 * smooth(torch.randn(B, T, M, generator=g), 64)))
    x = torch.randn(B, T, generator=g); zi = torch.randn(B, M, generator=g)
    ref = O.sample_wise_lpc(x, A, zi)
    y = G.sample_wise_lpc(x.to(dev), A.to(dev), zi.to(dev))
    print(f"  sample_wise_lpc dense A + zi: {rel(y, ref)}")
    lam = torch.rand(B, T, 1, generator=g) * 0.9
    ref = O.sample_wise_lpc(x, -lam, zi[:, :1]); y = G.sample_wise_lpc(x.to(dev), -lam.to(dev), zi[:, :1].to(dev))
    print(f"  order-1: {rel(y, ref)}")

def t_ss_bwd():
    g = np.load(os.path.join(GD, "grads_ss.npz"))
    H = int(g["hop"])
    ex = torch.tensor(g["ex"]).to(dev).requires_grad_(); gain = torch.tensor(g["gain"]).to(dev).requires_grad_(); a = torch.tensor(g["a"]).to(dev).requires_grad_()
    y = G.lpc_ss(ex, gain, a, H)
    print(f"  fwd vs ref y {rel(y, torch.tensor(g['ss_y']))}")
    dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), torch.tensor(g["ss_up"]).to(dev))
    print(f"  d_ex {rel(dex[:, :y.shape[1]], torch.tensor(g['ss_dex'])[:, :y.shape[1]])} d_gain {rel(dgain, torch.tensor(g['ss_dgain']))} d_a(flat) {rel(da.flatten(1), torch.tensor(g['ss_da']).flatten(1))}")

def t_ff():
    gd = np.load(os.path.join(GD, "filters_rand.npz")); H = int(gd["hop"])
    for M in (8, 20, 22):
        ex, gain, a = (torch.tensor(gd[f"{k}_{M}"]) for k in ("ex", "gain", "a"))
        win = torch.hann_window(4 * H)
        y = G.lpc_ff(ex.to(dev), gain.to(dev), a.to(dev), win.to(dev), H)
        print(f"  ff M{M}: vs reference golden {rel(y, torch.tensor(gd[f'ff_{M}']))}  vs oracle {rel(y, O.lpc_ff(ex, gain, a, H, 4*H))}")
        ss = G.lpc_ss(ex.to(dev), gain.to(dev), a.to(dev), H)
        print(f"  ss M{M}: vs reference golden {rel(ss, torch.tensor(gd[f'ss_{M}']))}")
        r = G.lpc_inverse(torch.tensor(gd[f"target_{M}"]).to(dev), a.to(dev), H)
        print(f"  inverse M{M}: {rel(r, torch.tensor(gd[f'inverse_{M}']))}")
    bq = torch.tensor(gd["biquads_8"]); ex, gain = torch.tensor(gd["ex_8"]), torch.tensor(gd["gain_8"])
    y = G.biquad_ff(ex.to(dev), gain.to(dev), bq.to(dev), torch.hann_window(4 * H).to(dev), H)
    print(f"  biquad cascade: vs golden {rel(y, torch.tensor(gd['bq_cascade_8']))}")
    for (B, T, Hh, M) in [(2, 48000, 240, 22), (2, 24000, 120, 22), (1, 9600, 240, 32), (1, 9600, 240, 12)]:
        Fr = T // Hh + 1
        gain, a = controls(B, Fr, M, seed=5)
        ex = torch.randn(B, T, generator=torch.Generator().manual_seed(2))
        y = G.lpc_ff(ex.to(dev), gain.to(dev), a.to(dev), torch.hann_window(4 * Hh).to(dev), Hh)
        print(f"  ff B{B} T{T} H{Hh} M{M}: {rel(y, O.lpc_ff(ex, gain, a, Hh, 4*Hh))}")

def t_fir():
    for v in ("ss", "ff"):
        g = np.load(os.path.join(GD, f"stages_{v}.npz")); H = int(g["hop"])
        kern = O.zero_phase_fir(torch.tensor(g["log_mag"]))
        noise = torch.tensor(g["noise"])[:, : g["harm"].shape[1]]
        y = G.ltv_fir_blocks(noise.to(dev), kern.to(dev), H)
        print(f"  noise FIR ({v}) vs reference golden {rel(y, torch.tensor(g['noise_filtered']))}")
        y2 = G.ltv_fir_blocks(noise.to(dev), kern.to(dev), H, add=torch.tensor(g["harm"]).to(dev))
        src = torch.tensor(g["harm"])[:, : y.shape[1]] + torch.tensor(g["noise_filtered"])
        print(f"  noise FIR + add: {rel(y2, src)}")
        r = G.room_fir(torch.tensor(g["lpc"]).to(dev), torch.tensor(g["room_kernel"]).to(dev))
        print(f"  room FIR ({v}) vs golden {rel(r, torch.tensor(g['out']))}")

def t_osc():
    table, _ = O.glottal_table()
    dk = O.decimate_kernel(4)
    for v in ("ss",):
        g = np.load(os.path.join(GD, f"stages_{v}.npz"))
        ph, w = torch.tensor(g["phase"]), torch.tensor(g["w"])
        for acc in ("aten_cpu", "fp64"):
            y = G.glottal_osc(ph.to(dev), int(g["phase_hop"]), w.to(dev), int(g["w_hop"]), table.to(dev), dk.to(dev), 4, True, acc)
            r64 = O.glottal_osc(ph, int(g["phase_hop"]), w, int(g["w_hop"]), table, 4, True, "fp64")
            print(f"  osc rtf-mode acc={acc}: vs reference golden {rel(y, torch.tensor(g['harm']))}  vs oracle-fp64-phase {rel(y, r64)}  (ref vs fp64 {rel(torch.tensor(g['harm']), r64)})")
    # training mode: sample-rate f0 random walk
    B, T = 2, 24000
    gen = torch.Generator().manual_seed(4)
    f0 = (200 + 80 * torch.cumsum(torch.randn(B, T, generator=gen), 1) / 100).clamp(80, 400)
    ph = f0 / 24000; w = torch.rand(B, T // 2400 + 1, generator=gen)
    for acc in ("aten_cpu", "fp64"):
        y = G.glottal_osc(ph.to(dev), 1, w.to(dev), 2400, table.to(dev), dk.to(dev), 4, True, acc)
        r32 = O.glottal_osc(ph, 1, w, 2400, table, 4, True, "fp32"); r64 = O.glottal_osc(ph, 1, w, 2400, table, 4, True, "fp64")
        print(f"  osc train-mode acc={acc}: vs oracle fp32 {rel(y, r32)} vs oracle fp64 {rel(y, r64)} (oracle32 vs 64 {rel(r32, r64)})")
    # generate()
    tabs = O.select_tables(table, w)
    up = O.upsample_time(ph / 4, 4); wr = O.phase_accumulate(up, "fp32")
    y = G.wavetable_read(wr.to(dev), tabs.to(dev), 9600)
    print(f"  wavetable_read vs oracle {rel(y, O.wavetable_read(wr, tabs, 9600))}")

def t_misc():
    x = torch.randn(5, 201)
    y = G.linear_upsample(x.to(dev), 240)
    ref = O.upsample_time(x, 240)
    print(f"  upsample bit-equal frac {float((y.cpu() == ref).float().mean()):.6f} max {float((y.cpu()-ref).abs().max()):.2e}")
    lg = torch.randn(300, 22) * 0.5
    a = G.rc2lpc(lg.to(dev))
    print(f"  rc2lpc max abs err {float((a.cpu() - O.rc2lpc(torch.tanh(lg))).abs().max()):.2e}")

def t_perf():
    B, T, H, M = 32, 48000, 240, 22
    Fr = T // H
    gain, a = controls(B, Fr, M, seed=9)
    ex = torch.randn(B, T).to(dev); gain = gain.to(dev); a = a.to(dev)
    win = torch.hann_window(960).to(dev)
    kern = torch.randn(B, Fr, 510, device=dev) * 0.01
    rk = torch.randn(127, device=dev) * 0.01
    table, _ = O.glottal_table(); table = table.to(dev); dk = O.decimate_kernel(4).to(dev)
    ph = torch.full((B, T), 150.0 / 24000, device=dev); w = torch.rand(B, 21, device=dev)
    def timeit(name, fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        print(f"  {name}: {s.elapsed_time(e)/n*1000:.1f} us")
    timeit("lpc_ss B32", lambda: G.lpc_ss(ex, gain, a, H))
    for ch in (480, 960):
        timeit(f"lpc_ss B32 chunk{ch}", lambda: G.lpc_ss(ex, gain, a, H, chunk=ch))
    timeit("lpc_ff B32", lambda: G.lpc_ff(ex, gain, a, win, H))
    exf = ex.clone().requires_grad_(); gf = gain.clone().requires_grad_(); af = a.clone().requires_grad_()
    yf = G.lpc_ff(exf, gf, af, win, H); upf = torch.randn_like(yf)
    timeit("lpc_ff bwd B32", lambda: torch.autograd.grad(yf, (exf, gf, af), upf, retain_graph=True))
    timeit("noise_fir B32", lambda: G.ltv_fir_blocks(ex, kern, H))
    timeit("room_fir B32", lambda: G.room_fir(ex, rk))
    timeit("osc B32 fp64", lambda: G.glottal_osc(ph, 1, w, 2400, table, dk, 4, True, "fp64"))
    exg = ex.clone().requires_grad_(); gg = gain.clone().requires_grad_(); ag = a.clone().requires_grad_()
    y = G.lpc_ss(exg, gg, ag, H); up = torch.randn_like(y)
    timeit("lpc_ss bwd B32", lambda: torch.autograd.grad(y, (exg, gg, ag), up, retain_graph=True))

for name, fn in [("ss", t_ss), ("ss encoder-derived", t_ss_gt), ("sample_wise_lpc", t_swl), ("ss backward", t_ss_bwd), ("ff/biquad/inverse", t_ff),
                 ("fir", t_fir), ("osc", t_osc), ("misc", t_misc), ("perf", t_perf)]:
    print(f"== {name}"); run(name, fn)
