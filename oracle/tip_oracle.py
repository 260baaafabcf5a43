"""CPU oracle for the TIP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path
(``transformer-inertial-poser_b200/``) never imports it and has no CPU fallback.

What it is: an independent numpy restatement of the reference's
``TF_RNN_Past_State.forward`` (``/root/reference/simple_transformer_with_state.py:60-102``)
in its *deterministic* mode: module in ``eval()``,
``past_state_dropout = 0`` (or an explicit keep-mask supplied by the caller),
``torch.no_grad()``, fp32.  The arithmetic the reference delegates to PyTorch
(``nn.Linear``, ``nn.TransformerEncoderLayer`` post-norm / relu / eps 1e-5,
``nn.MultiheadAttention``, ``nn.RNN`` tanh) is restated from PyTorch's published
definitions; torch is a third-party dependency of the reference, not pinned by it
(README says "only tested" with 1.7.1; this image has 2.11.0).

Pinning: the reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference module itself,
run in this container by ``oracle/make_golden.py`` (imports
``/root/reference/simple_transformer_with_state.py`` unmodified) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle.py`` checks this file against every one
of those fixtures.

Each function cites the reference line(s) it follows.
"""
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


def _layer_norm(x, w, b, eps=1e-5):
    # torch.nn.LayerNorm: biased variance over the last dim, eps inside the sqrt.
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * w + b


def forward(sd, x_imu, x_s, n_heads=16, with_rnn=True, keep_mask=None, past_scale=1.0,
            dtype=np.float32, return_intermediates=False):
    """Deterministic restatement of TF_RNN_Past_State.forward.

    sd         : dict key -> ndarray in the reference's state-dict layout.
    x_imu, x_s : (B, L, 72|90), (B, L, size_s).  Not modified (reference clones, :63-64).
    keep_mask  : optional (B, L, size_s) 0/1 array; x_s is multiplied by keep_mask*past_scale,
                 which is what ``nn.Dropout(past_state_dropout)`` at :77 does for a given mask
                 (past_scale = 1/(1-p)).  None reproduces past_state_dropout = 0.
    dtype      : np.float32 (parity oracle) or np.float64 (error budgeting).
    """
    W = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    x_imu = np.array(x_imu, dtype=dtype, copy=True)                 # :63
    x_s = np.array(x_s, dtype=dtype, copy=True)                     # :64
    x_s[np.isnan(x_s)] = 0                                          # :65
    B, L = x_imu.shape[0], x_imu.shape[1]
    x_s[:, :, 18 * 6: 18 * 6 + 3] *= 0                              # :75 root velocity removed
    if keep_mask is not None:                                       # :77 with an explicit mask
        x_s = x_s * (np.asarray(keep_mask, dtype=dtype) * dtype(past_scale))
    x = np.concatenate((x_imu, x_s), axis=2)                        # :78
    x = x @ W["in_linear.weight"].T + W["in_linear.bias"]           # :79   (B, L, E)
    E = x.shape[-1]
    d = E // n_heads
    # :88-89 feature permutation h*d+j -> j*n_heads+h
    x = x.reshape(B, L, n_heads, d).transpose(0, 1, 3, 2).reshape(B, L, E)
    inter = {"embed": x.copy()} if return_intermediates else None

    # :56-58, :85 causal mask: key index <= query index
    causal = np.tril(np.ones((L, L), dtype=bool))
    scale = dtype(1.0 / np.sqrt(d))
    n_layers = 1 + max(int(k.split(".")[2]) for k in W if k.startswith("tf_encode.layers."))
    for i in range(n_layers):                                       # :91 nn.TransformerEncoder
        p = f"tf_encode.layers.{i}."
        qkv = x @ W[p + "self_attn.in_proj_weight"].T + W[p + "self_attn.in_proj_bias"]
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        q = q.reshape(B, L, n_heads, d).transpose(0, 2, 1, 3) * scale
        k = k.reshape(B, L, n_heads, d).transpose(0, 2, 1, 3)
        v = v.reshape(B, L, n_heads, d).transpose(0, 2, 1, 3)
        s = q @ k.transpose(0, 1, 3, 2)                             # (B, H, L, L)
        s = np.where(causal, s, dtype(-np.inf))
        s = s - s.max(axis=-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(axis=-1, keepdims=True)
        o = (pr @ v).transpose(0, 2, 1, 3).reshape(B, L, E)
        a = o @ W[p + "self_attn.out_proj.weight"].T + W[p + "self_attn.out_proj.bias"]
        x = _layer_norm(x + a, W[p + "norm1.weight"], W[p + "norm1.bias"])
        f = np.maximum(x @ W[p + "linear1.weight"].T + W[p + "linear1.bias"], 0)
        f = f @ W[p + "linear2.weight"].T + W[p + "linear2.bias"]
        x = _layer_norm(x + f, W[p + "norm2.weight"], W[p + "norm2.bias"])
        if return_intermediates:
            inter[f"layer{i}"] = x.copy()

    if with_rnn:                                                    # :95-99 tanh RNN, h0 = 0
        R = W["rnn.weight_hh_l0"].shape[0]
        gi = x @ W["rnn.weight_ih_l0"].T + W["rnn.bias_ih_l0"]
        h = np.zeros((B, R), dtype=dtype)
        hs = np.empty((B, L, R), dtype=dtype)
        Whh_t = np.ascontiguousarray(W["rnn.weight_hh_l0"].T)
        for t in range(L):
            h = np.tanh(gi[:, t] + h @ Whh_t + W["rnn.bias_hh_l0"])
            hs[:, t] = h
        x = hs
        if return_intermediates:
            inter["rnn"] = x.copy()
    y = x @ W["linear.weight"].T + W["linear.bias"]                 # :102
    return (y, inter) if return_intermediates else y


# ---------------------------------------------------------------------------------------------
# Caller-side window assembly (rows a10 / N1): restatement of the pre-model slice of
# RTRunnerMin.step (real_time_runner_minimal.py:59-76,131-147) and imu_rotate_to_local
# (data_utils.py:190-219).  Pure numpy, float64 like the reference buffers.

IMU_N_SMOOTH = 5            # constants.py:15
ACC_MOVING_AVE_LEN = 11     # constants.py:16
ACC_SUM_WIN_LEN = 40        # constants.py:17
ACC_SUM_DOWN_SCALE = 15.0   # constants.py:18


def imu_rotate_to_local(batch_imu):
    """data_utils.py:190-219: rotations / accelerations of the 5 non-root IMUs into the root frame."""
    batch_imu = np.asarray(batch_imu, dtype=np.float64)
    root_r = batch_imu[:, :9].reshape(-1, 3, 3)
    inv = np.linalg.inv(root_r)
    other_r = batch_imu[:, 9:54].reshape(-1, 5, 3, 3)
    other_r_local = np.einsum("tij,tnjk->tnik", inv, other_r)
    root_acc = batch_imu[:, 54:57]
    other_acc = batch_imu[:, 57:72].reshape(-1, 5, 3)
    other_acc_local = np.einsum("tij,tnj->tni", inv, other_acc)
    return np.concatenate((root_r.reshape(-1, 9), other_r_local.reshape(-1, 45),
                           root_acc, other_acc_local.reshape(-1, 15)), axis=1)


class WindowAssembler:
    """Append-only buffers exactly as RTRunnerMin keeps them (real_time_runner_minimal.py:33-37);
    ``push`` returns the (L, 72|90) model window or None during the 5-call warm-up."""

    def __init__(self, max_input_l=40, with_acc_sum=True):
        self.max_input_l = max_input_l
        self.with_acc_sum = with_acc_sum
        self.raw, self.smoothed, self.acc_sum = [], [], []

    def push(self, cur_imu):
        cur_imu = np.asarray(cur_imu, dtype=np.float64)
        if len(self.raw) == 0:                                       # :60-63
            for _ in range(IMU_N_SMOOTH):
                self.raw.append(cur_imu.copy())
        self.raw.append(cur_imu.copy())                              # :66
        if len(self.raw) >= ACC_MOVING_AVE_LEN:                      # :68-74
            win = np.array(self.raw[-ACC_MOVING_AVE_LEN:])
            self.smoothed.append(np.concatenate((self.raw[-IMU_N_SMOOTH - 1][:54],
                                                 np.mean(win[:, 54:72], axis=0))))
        if len(self.smoothed) < 1:                                   # :125
            return None
        in_imu = imu_rotate_to_local(np.array(self.smoothed[-self.max_input_l:]))   # :131-132
        if self.with_acc_sum:                                        # :134-141
            self.acc_sum.append(np.sum(in_imu[-ACC_SUM_WIN_LEN:, 54:72], axis=0))
            win = np.array(self.acc_sum[-self.max_input_l:]) / ACC_SUM_DOWN_SCALE
            in_imu = np.concatenate((in_imu, win), axis=1)
        return in_imu
