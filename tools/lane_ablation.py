"""Marginal cost of each stage of the forward under execution lanes: the laned B = 256 forward is timed with one stage
NOT launched at a time (TIP_SKIP, a diagnostic switch of the library; outputs are garbage in those runs).  A stage whose
removal saves its whole single-lane duration is not being overlapped by the other lanes; one that saves only its share
of SM-time is.  usage: python tools/lane_ablation.py  (LANES=3 N=96)"""
import os, subprocess, sys
stages = [("none", 0), ("condition", 1), ("in_linear", 2), ("qkv", 4), ("attention", 8), ("out_proj_ln", 16), ("ff1", 32),
          ("ff2_ln", 64), ("rnn_ih", 128), ("rnn", 256), ("head", 512)]
base = None
for name, bit in stages:
    env = dict(os.environ, TIP_SKIP=str(bit), LANES=os.environ.get("LANES", "3"), N=os.environ.get("N", "96"), TAG="skip_" + name)
    out = subprocess.run([sys.executable, "tools/lane_knobs_probe.py"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    us = float(out[-1].split(":")[1].split("us")[0])
    if base is None: base = us
    print(f"{name:12s} {us:7.1f} us per forward   saved {base - us:6.1f} us", flush=True)
