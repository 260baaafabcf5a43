"""Aggregate device time per whole-batch forward over LANES lanes (default 3) -- run under different TIP_* switches."""
import os, sys
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch
from bench import build_model, load_weights, synth
from tip_b200.pipeline import ForwardLanes
sd, _ = load_weights()
dev = torch.device('cuda:0')
B = int(os.environ.get('B', '256')); N = int(os.environ.get('N', '192')); NS = 16
model = build_model(sd, dev)
if os.environ.get('MODE') == 'as_shipped':      # the consumers' default: train mode, fresh Dropout(0.8) per call
    model.train(); model.past_state_dropout = 0.8
sets = []
for i in range(NS):
    xi, xs = synth(1 + 1000 * i, B)
    sets.append((torch.from_numpy(xi).to(dev), torch.from_numpy(xs).to(dev)))
for nl in [int(x) for x in os.environ.get('LANES', '3').split(',')]:
    lanes = ForwardLanes(model, nl)
    outs = [torch.empty((B, 40, 131), device=dev) for _ in range(nl)]
    def run(n):
        lanes.fork()
        for i in range(n):
            lanes.forward(i, *sets[i % NS], out=outs[i % nl])
        lanes.join()
    run(3 * NS * nl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(); run(N); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / N)
    print(os.environ.get('TAG', ''), "lanes %d: %.1f us per forward -> %.0f frames/s" % (nl, best * 1e3, B / best * 1e3), flush=True)
