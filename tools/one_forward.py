"""Two forwards at B=256 (profiling target for ncu)."""
import os, sys
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch
from bench import build_model, load_weights, synth
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
if os.environ.get('MODE') == 'as_shipped':      # the consumers' default: train mode, fresh Dropout(0.8) per call
    m.train(); m.past_state_dropout = 0.8
B = int(os.environ.get('B', '256'))
xi, xs = synth(1, B)
xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
for _ in range(int(os.environ.get('N', '2'))):
    m(xi, xs)
torch.cuda.synchronize()
