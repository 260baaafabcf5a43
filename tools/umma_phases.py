"""Phase timestamps of CTA 0 of the tcgen05 GEMMs (debug build hook TIP_DBG=4)."""
import ctypes as C, os, sys
os.environ["TIP_TS"] = "1"
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
from tip_b200 import capi
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
xi, xs = synth(1, B)
xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
for _ in range(3): m(xi, xs)
torch.cuda.synchronize()
lib = capi.load_library()
lib.tip_debug_timestamps.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * (64 * 8))()
assert lib.tip_debug_timestamps(buf, 64 * 8) == 0
names = ["in", "qkv", "out_ln", "ff1", "ff2_ln", "ih", "head_r", "head_e"]
lab = ["start", "first_full|chunk0_done", "mma_issued", "tfull_seen", "ln_pass1|ldtm_done", "ln_pass2|sts_done", "epi_done", "exit"]
for w in range(8):
    for layer in (0, 16):
        t = [buf[8 * (w + layer) + i] for i in range(8)]
        if t[0] == 0: continue
        print(names[w], "layer>0" if layer else "layer0", {lab[i]: (t[i] - t[0]) / 1e3 for i in range(1, 8) if t[i]})
