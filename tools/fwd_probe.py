"""Device time per forward (CUDA events, L2 flushed) and e2e host-entry time at B=256."""
import os, sys, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
B = int(os.environ.get('B', '256'))
sets = []
for i in range(4):
    xi, xs = synth(1 + 1000 * i, B)
    sets.append((torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for i in range(12): m(*sets[i % 4])
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
for i in range(60):
    flush.zero_(); ev[i][0].record(); m(*sets[i % 4]); ev[i][1].record()
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev)
print("B", B, "device us/fwd median %.1f min %.1f" % (t[30] * 1e3, t[0] * 1e3), "launches", m.last_launch_count())
hx = [(torch.from_numpy(synth(7000 + i, B)[0]).pin_memory(), torch.from_numpy(synth(7000 + i, B)[1]).pin_memory()) for i in range(2)]
hy = torch.empty((B, 40, 131), dtype=torch.float32).pin_memory()
for i in range(4): m.forward_host(hx[i % 2][0], hx[i % 2][1], out=hy)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50): y = m.forward_host(hx[i % 2][0], hx[i % 2][1], out=hy)
torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 50
print("e2e us/step %.1f -> %.0f frames/s" % (el * 1e6, B / el))
ref = m(hx[1][0].cuda(), hx[1][1].cuda()).cpu()
print("host-entry vs device-call max diff", float((ref - y).abs().max()))
