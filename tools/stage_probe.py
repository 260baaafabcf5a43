import sys, os, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np, contextlib, io
from bench import build_model, load_weights, synth
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
m.set_profile(True)
for B in (1, 8, 16, 64, 128, 256, 512):
    xi, xs = synth(1, B)
    xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
    for _ in range(3): m(xi, xs)
    torch.cuda.synchronize()
    acc = {}
    for _ in range(5):
        m(xi, xs); 
        for n, l, ms in m.profile(): acc.setdefault(n, []).append(ms)
    print(B, {k: round(float(np.mean(v))*1e3,1) for k, v in acc.items()})
