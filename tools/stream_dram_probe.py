"""BASELINE configs[4]: DRAM bytes per closed-loop streaming frame (one stream per GPU; replicas are identical, so the per-GPU
figure of N = 1 is the per-GPU figure of any N).  Run under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv
and sum over the kernels of the LAST frame (the session is in steady state, L = 40, frame = one graph of kernels)."""
import os, sys
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import numpy as np, torch
from bench import build_model, load_weights
from tip_b200.streaming import StreamSession
sd, _ = load_weights()
m = build_model(sd, torch.device('cuda:0'))
m.set_use_graphs(False)                     # kernels stay visible to ncu one by one
S = int(os.environ.get("S", "1"))
sess = StreamSession(m, n_streams=S)
s0 = np.zeros((S, 114)); s0[:, 2] = 0.95
sess.set_state(s0 if S > 1 else s0[0])
rs = np.random.RandomState(3)
for t in range(int(os.environ.get("FRAMES", "45"))):
    raw = np.concatenate((np.tile(np.eye(3).reshape(9), 6), 3.0 * rs.standard_normal(18))).astype(np.float32)
    raw = np.tile(raw[None], (S, 1))
    if t == int(os.environ.get("FRAMES", "45")) - 1:
        torch.cuda.synchronize(); torch.cuda.nvtx.range_push("last_frame")
    st = sess.step_closed(raw)
torch.cuda.synchronize()
