import sys, os, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np, contextlib, io
from bench import build_model, load_weights, synth
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
import os
if os.environ.get('ENGINE'): m.set_gemm_engine(int(os.environ['ENGINE']))
if os.environ.get('MODE') == 'as_shipped':
    m.train(); m.past_state_dropout = 0.8
m.set_profile(True)
for B in [int(b) for b in os.environ.get('BS','1,8,16,64,128,256,512').split(',')]:
    xi, xs = synth(1, B)
    xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
    for _ in range(3): m(xi, xs)
    torch.cuda.synchronize()
    acc = {}
    for _ in range(5):
        m(xi, xs); 
        for n, l, ms in m.profile(): acc.setdefault(n, []).append(ms)
    d={k: round(float(np.mean(v))*1e3,1) for k, v in acc.items()}
    tot=sum(v*(4 if k in ('qkv','attention','qkv_attn','out_proj_ln','ff1','ff2_ln','ffn_ln') else 1) for k,v in d.items())
    print(B, round(tot,1), d)
