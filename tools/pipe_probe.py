"""End-to-end job pipeline (HostPipeline, pinned host buffers, one pose row per window back) over lanes x depth."""
import os, sys, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
from tip_b200.pipeline import HostPipeline
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
B, K = 256, int(os.environ.get("K", "60"))
NBMAX = 25
hx = [(torch.from_numpy(synth(7000 + i, B)[0]).pin_memory(), torch.from_numpy(synth(7000 + i, B)[1]).pin_memory()) for i in range(NBMAX)]
hl = [torch.empty((B, 131), dtype=torch.float32).pin_memory() for _ in range(NBMAX)]
for spec in os.environ.get("SPECS", "3x2,4x2,5x2,5x3,6x2,5x4").split(","):
    nl, per = (int(x) for x in spec.split("x"))
    depth = nl * per
    nb = depth + 1
    pipe = HostPipeline(m, depth=depth, lanes=nl, last_row_only=True)
    for i in range(3 * nb):
        pipe.submit(hx[i % nb][0], hx[i % nb][1], hl[i % nb])
    for _ in pipe.drain(): pass
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        for i in range(K):
            pipe.submit(hx[i % nb][0], hx[i % nb][1], hl[i % nb])
        for _ in pipe.drain(): pass
        best = min(best, time.perf_counter() - t0)
    print(f"lanes {nl} x {per} jobs in flight: {best / K * 1e6:7.1f} us per step -> {B * K / best:9.0f} frames/s", flush=True)
    del pipe
