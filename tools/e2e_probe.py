"""Where the host-entry time goes (B=256): full call, last-row-only call, raw copies through torch."""
import os, sys, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
B = 256
hx = [(torch.from_numpy(synth(7000 + i, B)[0]).pin_memory(), torch.from_numpy(synth(7000 + i, B)[1]).pin_memory()) for i in range(2)]
hy = torch.empty((B, 40, 131), dtype=torch.float32).pin_memory()
hl = torch.empty((B, 131), dtype=torch.float32).pin_memory()
def timeit(f, n=50):
    for _ in range(4): f(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("forward_host full        %.1f us" % timeit(lambda i: m.forward_host(hx[i % 2][0], hx[i % 2][1], out=hy)))
print("forward_host last row    %.1f us" % timeit(lambda i: m.forward_host(hx[i % 2][0], hx[i % 2][1], last_row_only=True, out=hl)))
dxi = torch.empty((B, 40, 90), device='cuda'); dxs = torch.empty((B, 40, 131), device='cuda')
def torch_path(i):
    dxi.copy_(hx[i % 2][0], non_blocking=True); dxs.copy_(hx[i % 2][1], non_blocking=True)
    y = m(dxi, dxs); hy.copy_(y, non_blocking=True); torch.cuda.synchronize()
print("torch copies + model()   %.1f us" % timeit(torch_path))
def copies_only(i):
    dxi.copy_(hx[i % 2][0], non_blocking=True); dxs.copy_(hx[i % 2][1], non_blocking=True); torch.cuda.synchronize()
print("H2D only (2 copies)      %.1f us" % timeit(copies_only))
y = m(dxi, dxs)
def d2h_only(i):
    hy.copy_(y, non_blocking=True); torch.cuda.synchronize()
print("D2H only                 %.1f us" % timeit(d2h_only))
def fwd_only(i):
    m(dxi, dxs); torch.cuda.synchronize()
print("model() + sync           %.1f us" % timeit(fwd_only))
