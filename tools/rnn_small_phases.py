import ctypes as C, os, sys
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
from tip_b200 import capi
lib = capi.load_library()
buf = (C.c_ulonglong * 8)()
lib.tip_debug_rnn_timestamps.argtypes = [C.c_void_p]
lib.tip_debug_rnn_timestamps(buf)      # allocate
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
m.set_use_graphs(False)
for B in (1, 8):
    xi, xs = synth(1, B)
    xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
    for _ in range(3): m(xi, xs)
    torch.cuda.synchronize()
    lib.tip_debug_rnn_timestamps(buf)
    t = list(buf)
    print(B, "wait_done->reduced %.2f us, reduced->sent %.2f us, sent->next wait_done %.2f us, step %.2f" % (
        (t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[4]-t[2])/1e3, (t[4]-t[0])/1e3))
