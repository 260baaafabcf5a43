"""Synthetic workload generators of the product side (bench.py, examples): seeded random weights with the reference's
parameter shapes (simple_transformer_with_state.py:20-42) and synthetic 6-IMU windows with the value distributions of
SURVEY.md section 8d / Appendix A (there is no DIP data in the container).  Pure numpy, no dependency on ``oracle/``
(the oracle keeps its own copy for the tests; ``tests/test_oracle.py`` checks that both produce identical arrays)."""
from __future__ import annotations

import numpy as np


# state-dict key order of the reference module (56 tensors for tf_layers=4, with_rnn=True);
# this is also the order the C-ABI ``tip_pack_weights`` takes its pointers in.
def state_dict_keys(tf_layers: int = 4, with_rnn: bool = True):
    keys = ["in_linear.weight", "in_linear.bias"]
    for i in range(tf_layers):
        p = f"tf_encode.layers.{i}."
        keys += [p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias",
                 p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                 p + "linear1.weight", p + "linear1.bias",
                 p + "linear2.weight", p + "linear2.bias",
                 p + "norm1.weight", p + "norm1.bias",
                 p + "norm2.weight", p + "norm2.bias"]
    if with_rnn:
        keys += ["rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0"]
    keys += ["linear.weight", "linear.bias"]
    return keys


def random_state_dict(seed: int, input_size_imu=72, size_s=131, rnn_hid_size=512,
                      tf_hid_size=1024, tf_in_dim=256, n_heads=16, tf_layers=4,
                      with_rnn=True, with_acc_sum=True, dtype=np.float32):
    """Seeded random weights with the reference's parameter shapes
    (simple_transformer_with_state.py:20-42).  Uses the frozen legacy
    ``numpy.random.RandomState`` stream so fixtures regenerate bit-identically.
    Scales are chosen so activations stay O(1) like the shipped checkpoints."""
    rs = np.random.RandomState(seed)
    d_in = input_size_imu + size_s + (18 if with_acc_sum else 0)
    E, F, R = tf_in_dim, tf_hid_size, rnn_hid_size

    def lin(n_out, n_in):
        bound = 1.0 / np.sqrt(n_in)
        return (rs.uniform(-bound, bound, size=(n_out, n_in)).astype(dtype),
                rs.uniform(-bound, bound, size=(n_out,)).astype(dtype))

    sd = {}
    sd["in_linear.weight"], sd["in_linear.bias"] = lin(E, d_in)
    for i in range(tf_layers):
        p = f"tf_encode.layers.{i}."
        sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"] = lin(3 * E, E)
        sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"] = lin(E, E)
        sd[p + "linear1.weight"], sd[p + "linear1.bias"] = lin(F, E)
        sd[p + "linear2.weight"], sd[p + "linear2.bias"] = lin(E, F)
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = (1.0 + 0.1 * rs.standard_normal(E)).astype(dtype)
            sd[p + n + ".bias"] = (0.1 * rs.standard_normal(E)).astype(dtype)
    if with_rnn:
        sd["rnn.weight_ih_l0"], sd["rnn.bias_ih_l0"] = lin(R, E)
        sd["rnn.weight_hh_l0"], sd["rnn.bias_hh_l0"] = lin(R, R)
        sd["linear.weight"], sd["linear.bias"] = lin(size_s, R)
    else:
        sd["linear.weight"], sd["linear.bias"] = lin(size_s, E)
    return {k: sd[k] for k in state_dict_keys(tf_layers, with_rnn)}


def synth_inputs(seed: int, B: int, L: int, input_size_imu=72, size_s=131,
                 with_acc_sum=True, nan_frac=0.05, dtype=np.float32):
    """Seeded synthetic IMU windows with the value distributions of SURVEY.md section 8d /
    Appendix A (no DIP data in the container).  Returns (x_imu (B,L,72|90), x_s (B,L,size_s))."""
    rs = np.random.RandomState(seed)
    n_imu = input_size_imu + (18 if with_acc_sum else 0)
    x_imu = np.empty((B, L, n_imu), dtype=np.float64)
    x_imu[..., :54] = rs.uniform(-1, 1, size=(B, L, 54))          # 6 rotation matrices
    x_imu[..., 54:72] = 3.0 * rs.standard_normal((B, L, 18))      # smoothed accelerations
    if with_acc_sum:
        x_imu[..., 72:90] = 2.0 * rs.standard_normal((B, L, 18))  # acc-sum / 15
    x_s = np.empty((B, L, size_s), dtype=np.float64)
    x_s[..., :108] = rs.uniform(-1, 1, size=(B, L, 108))          # 18 joints x 2 columns of R
    x_s[..., 108:111] = rs.standard_normal((B, L, 3))             # root velocity (zeroed in model)
    n_c = size_s - 111
    c = rs.uniform(-0.15, 0.15, size=(B, L, n_c))
    c[..., 0::4] = (rs.uniform(size=(B, L, n_c // 4)) < 0.5) * 1.0
    x_s[..., 111:] = c
    if nan_frac > 0:
        rows = rs.uniform(size=(B, L)) < nan_frac
        x_s[rows, 108:111] = np.nan                                # DIP rows carry NaN root velocity
    return x_imu.astype(dtype), x_s.astype(dtype)
