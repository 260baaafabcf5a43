"""Mint golden vectors from the UNMODIFIED reference module (test infrastructure).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_golden.py

Imports ``/root/reference/simple_transformer_with_state.py`` read-only, puts it into the
deterministic mode SURVEY.md section 8c defines (``eval()``, ``past_state_dropout = 0``,
``no_grad``, fp32, CPU) and writes small ``.npz`` fixtures to ``tests/golden/``.  Random-weight
cases regenerate their weights from a seed at test time (``oracle.tip_oracle.random_state_dict``);
checkpoint cases need ``baseline/_ref/model-*.pt`` (copied there by ``__graft_entry__.build()``).
"""
import contextlib
import io
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tip_oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference_class():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_simple_transformer_with_state", os.path.join(REF, "simple_transformer_with_state.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TF_RNN_Past_State


def build_ref(TF, sd, size_s=131, with_rnn=True, with_acc_sum=True):
    with contextlib.redirect_stdout(io.StringIO()):
        m = TF(72, size_s, rnn_hid_size=512, tf_hid_size=1024, tf_in_dim=256, n_heads=16,
               tf_layers=4, dropout=0.0, in_dropout=0.0, past_state_dropout=0.8,
               with_rnn=with_rnn, with_acc_sum=with_acc_sum)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m.eval()
    m.past_state_dropout = 0.0
    return m


def ref_forward(m, x_imu, x_s):
    with torch.no_grad():
        return m(torch.from_numpy(x_imu), torch.from_numpy(x_s)).numpy()


def feedback_row(y_last):
    """Closed-loop load generator of SURVEY.md section 8d (a simplification of
    real_time_runner_minimal.py:87-112,152-167,196): normalise each joint's two 3-vectors,
    keep root-vel, SBP logit>0 -> flag, offsets / 5."""
    s = y_last.astype(np.float64).copy()
    r = s[:108].reshape(18, 3, 2)
    r = r / (np.linalg.norm(r, axis=1, keepdims=True) + 1e-6)
    s[:108] = r.reshape(-1)
    c = s[111:]
    c[0::4] = (c[0::4] > 0) * 1.0
    c[1::4] /= 5.0
    c[2::4] /= 5.0
    c[3::4] /= 5.0
    return s.astype(np.float32)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    TF = load_reference_class()

    # ---- random-weight cases (weights regenerate from the seed) ------------------------------
    cases = [
        # name, wseed, xseed, B, L, size_s, with_rnn, with_acc_sum, keep_mask?
        ("rw_b1_l40", 11, 0, 1, 40, 131, True, True, False),
        ("rw_b1_l1", 11, 1, 1, 1, 131, True, True, False),
        ("rw_b1_l7", 11, 2, 1, 7, 131, True, True, False),
        ("rw_b3_l39", 11, 3, 3, 39, 131, True, True, False),
        ("rw_b5_l40_s119", 12, 4, 5, 40, 119, True, True, False),
        ("rw_b2_l40_noacc", 13, 5, 2, 40, 131, True, False, False),
        ("rw_b2_l33_nornn", 14, 6, 2, 33, 131, False, True, False),
        ("rw_b4_l40_mask", 11, 7, 4, 40, 131, True, True, True),
    ]
    for name, wseed, xseed, B, L, size_s, with_rnn, with_acc_sum, use_mask in cases:
        sd = O.random_state_dict(wseed, size_s=size_s, with_rnn=with_rnn, with_acc_sum=with_acc_sum)
        m = build_ref(TF, sd, size_s, with_rnn, with_acc_sum)
        x_imu, x_s = O.synth_inputs(xseed, B, L, size_s=size_s, with_acc_sum=with_acc_sum)
        out = dict(x_imu=x_imu, x_s=x_s, wseed=wseed, size_s=size_s, with_rnn=with_rnn,
                   with_acc_sum=with_acc_sum)
        if use_mask:
            # explicit keep-mask x5, applied OUTSIDE the reference (algebraically what :77 does)
            rs = np.random.RandomState(1000 + xseed)
            keep = (rs.uniform(size=x_s.shape) < 0.2).astype(np.float32)
            out["keep_mask"] = keep
            out["past_scale"] = np.float32(5.0)
            xs_in = np.where(np.isnan(x_s), 0, x_s) * keep * np.float32(5.0)
            y = ref_forward(m, x_imu, xs_in.astype(np.float32))
        else:
            y = ref_forward(m, x_imu, x_s)
        out["y"] = y
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, y.shape, float(np.abs(y).max()))

    # ---- B=256 (BASELINE config 2), random weights: store a subsample + checksums ------------
    sd = O.random_state_dict(11)
    m = build_ref(TF, sd)
    x_imu, x_s = O.synth_inputs(1, 256, 40)
    y = ref_forward(m, x_imu, x_s)
    idx = np.array([0, 17, 100, 255])
    np.savez_compressed(os.path.join(GOLD, "rw_b256_l40_sub.npz"), wseed=11, xseed=1, idx=idx,
                        y_sub=y[idx], y_last=y[:, -1, :],
                        y_sum=np.float64(y.astype(np.float64).sum()),
                        y_abs_sum=np.float64(np.abs(y.astype(np.float64)).sum()))
    print("rw_b256_l40_sub", y.shape)

    # ---- checkpoint cases ---------------------------------------------------------------------
    os.makedirs(os.path.join(ROOT, "baseline", "_ref"), exist_ok=True)
    for ck in ("model-with-dip9and10", "model-without-dip9and10"):
        src = os.path.join(REF, "output", ck + ".pt")
        dst = os.path.join(ROOT, "baseline", "_ref", ck + ".pt")
        if not os.path.exists(dst):
            shutil.copyfile(src, dst)
        sd = {k: v.numpy() for k, v in torch.load(src, map_location="cpu").items()}
        m = build_ref(TF, sd)
        for (B, L, xseed) in ((1, 40, 0), (3, 39, 21), (1, 7, 22)):
            x_imu, x_s = O.synth_inputs(xseed, B, L)
            y = ref_forward(m, x_imu, x_s)
            np.savez_compressed(os.path.join(GOLD, f"ck_{ck}_b{B}_l{L}.npz"),
                                x_imu=x_imu, x_s=x_s, y=y, checkpoint=ck)
            print(ck, B, L, float(np.abs(y).max()))

    # ---- 200-frame streaming trace (ring-buffer path), checkpoint weights ---------------------
    ck = "model-with-dip9and10"
    sd = {k: v.numpy() for k, v in torch.load(os.path.join(REF, "output", ck + ".pt"),
                                              map_location="cpu").items()}
    m = build_ref(TF, sd)
    T = 200
    imu_rows, _ = O.synth_inputs(2, 1, T)
    imu_rows = imu_rows[0]
    _, s0 = O.synth_inputs(3, 1, 1, nan_frac=0.0)
    s_rows = [feedback_row(s0[0, 0])]
    s_rows[0][111:] = 0.0                      # s_init has all-zero SBPs (runner :45)
    y_last = np.zeros((T, 131), np.float32)
    for t in range(T):
        lo = max(0, t + 1 - 40)
        xi = imu_rows[lo:t + 1][None]
        xs = np.array(s_rows[lo:t + 1])[None]
        y = ref_forward(m, xi, xs)
        y_last[t] = y[0, -1]
        s_rows.append(feedback_row(y[0, -1]))
    np.savez_compressed(os.path.join(GOLD, "ck_stream200.npz"), imu_rows=imu_rows,
                        s_rows=np.array(s_rows[:T]), y_last=y_last, checkpoint=ck)
    print("stream200 done", float(np.abs(y_last).max()))


if __name__ == "__main__":
    main()
