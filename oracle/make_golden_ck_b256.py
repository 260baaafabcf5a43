"""Mint the batch=256 golden for the released checkpoint (BASELINE configs[1] with the weights the bench uses) from the
UNMODIFIED reference module.  Build container only; same conventions as make_golden.py (test infrastructure).

    python oracle/make_golden_ck_b256.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tip_oracle as O  # noqa: E402
from oracle.make_golden import GOLD, REF, build_ref, load_reference_class, ref_forward  # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    TF = load_reference_class()
    ck = "model-with-dip9and10"
    sd = {k: v.numpy() for k, v in torch.load(os.path.join(REF, "output", ck + ".pt"), map_location="cpu").items()}
    m = build_ref(TF, sd)
    xseed = 1
    x_imu, x_s = O.synth_inputs(xseed, 256, 40)
    y = ref_forward(m, x_imu, x_s)
    idx = np.array([0, 31, 128, 255])
    np.savez_compressed(os.path.join(GOLD, "ck_b256_l40_sub.npz"), checkpoint=ck, xseed=xseed, idx=idx,
                        y_sub=y[idx], y_last=y[:, -1, :],
                        y_sum=np.float64(y.astype(np.float64).sum()),
                        y_abs_sum=np.float64(np.abs(y.astype(np.float64)).sum()))
    print("ck_b256_l40_sub", y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
