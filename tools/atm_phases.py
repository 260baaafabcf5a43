"""Phase timestamps of CTA 0 of the A-in-TMEM GEMMs (TIP_TS=1; run with TIP_ATM=15)."""
import ctypes as C, os, sys
os.environ["TIP_TS"] = "1"
os.environ.setdefault("TIP_ATM", "15")
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, numpy as np
from bench import build_model, load_weights, synth
from tip_b200 import capi
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
xi, xs = synth(1, B)
xi, xs = torch.from_numpy(xi).cuda(), torch.from_numpy(xs).cuda()
m.set_profile(True)
for _ in range(3): m(xi, xs)
torch.cuda.synchronize()
lib = capi.load_library()
lib.tip_debug_timestamps.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 2048)()
assert lib.tip_debug_timestamps(buf, 2048) == 0
for k, name in enumerate(["in_linear", "qkv", "ff1", "rnn_ih"]):
    t = [buf[1024 + 64 * k + i] for i in range(64)]
    if t[0] == 0: continue
    r = lambda i: round((t[i] - t[0]) / 1e3, 2) if t[i] else None
    print(name, "A_full", r(1), "A_cp_issued", r(2), "W_first_full", r(3), "end", r(36))
    print("   tempty_passed", [r(40 + i) for i in range(8)])
    print("   mma_issued   ", [r(4 + i) for i in range(8)])
    print("   tfull_seen   ", [r(12 + i) for i in range(8)])
    print("   acc_in_regs  ", [r(20 + i) for i in range(8)])
    print("   epi_done     ", [r(28 + i) for i in range(8)])
t = [buf[1500 + i] for i in range(6)]
if t[0]:
    print("attention CTA 0 (us from start): loads issued + Q fragments requested", (t[1] - t[0]) / 1e3, "K/V landed", (t[2] - t[0]) / 1e3,
          "K fragments in registers", (t[3] - t[0]) / 1e3, "S / softmax / PV done", (t[4] - t[0]) / 1e3, "stored", (t[5] - t[0]) / 1e3)
