"""Host-entry time per pinned input buffer set (NUMA / placement effects) at B=256."""
import sys, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch
from bench import build_model, load_weights, synth
sd, _ = load_weights(); m = build_model(sd, torch.device('cuda:0'))
B = 256
sets = [(torch.from_numpy(synth(7000 + i, B)[0]).pin_memory(), torch.from_numpy(synth(7000 + i, B)[1]).pin_memory()) for i in range(4)]
out = torch.empty((B, 40, 131)).pin_memory()
def t(f, n=30):
    for _ in range(4): f(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
for k in range(4):
    print("set", k, "forward_host %.1f us" % t(lambda i: m.forward_host(sets[k][0], sets[k][1], out=out)))
print("alternating 0/1 %.1f us" % t(lambda i: m.forward_host(sets[i % 2][0], sets[i % 2][1], out=out)))
print("alternating 0..3 %.1f us" % t(lambda i: m.forward_host(sets[i % 4][0], sets[i % 4][1], out=out)))
d = torch.empty((B, 40, 131), device='cuda')
for k in range(4):
    print("set", k, "H2D of x_s alone %.1f us" % t(lambda i: (d.copy_(sets[k][1], non_blocking=True), torch.cuda.synchronize())))
print("alternating H2D of x_s %.1f us" % t(lambda i: (d.copy_(sets[i % 2][1], non_blocking=True), torch.cuda.synchronize())))
