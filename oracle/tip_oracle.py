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


# ---------------------------------------------------------------------------------------------
# Dropout masks of the product's as-shipped (stochastic) mode.  The reference draws them from torch's Philox
# stream (fresh nn.Dropout at simple_transformer_with_state.py:73,77; the encoder layers' p = 0.1 dropouts
# while the module is in train mode, offline_testing_simple.py:98); the product draws them from a counter
# hash documented in include/tip_b200.h (tip_dropout).  This is the numpy restatement of THAT generator, so a
# stochastic device forward can be checked mask for mask (the statistics -- keep rate 1-p, scale 1/(1-p), sites
# -- are checked on these masks in tests/test_oracle.py).
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_IN, SEED_PAST = 0x1111, 0x2222
SEED_CHUNK = 0x632BE59BD9B4E019


def seed_attn(l): return 101 * (l + 1)
def seed_out(l): return 211 * (l + 1)
def seed_ff1(l): return 307 * (l + 1)
def seed_ff2(l): return 401 * (l + 1)


def hash_u64(seed, idx):
    """splitmix64 finaliser of seed + 0x9E3779B97F4A7C15 * (idx + 1), modulo 2^64 (tip_common.cuh hash_u64)."""
    with np.errstate(over="ignore"):
        idx = np.asarray(idx, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(0x9E3779B97F4A7C15) * (idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def drop_threshold(p):
    p = np.float32(p)
    return 0 if p <= 0 else (65536 if p >= 1 else int(np.float32(p * np.float32(65536.0) + np.float32(0.5))))


def dropout_factors(seed, idx, p, dtype=np.float32):
    """nn.Dropout(p) factors (0 or 1/(1-p)) of the elements ``idx`` (any integer array) of the site ``seed``."""
    idx = np.asarray(idx, dtype=np.uint64)
    thr = drop_threshold(p)
    if thr == 0:
        return np.ones(idx.shape, dtype=dtype)
    h = hash_u64(seed, idx >> np.uint64(2))
    u = (h >> (np.uint64(16) * (idx & np.uint64(3)))) & np.uint64(0xFFFF)
    inv = np.float32(0.0) if p >= 1 else np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u < np.uint64(thr), np.float32(0.0), inv).astype(dtype)


def attn_drop_index(B, n_heads, L):
    """Element indices (B, H, L, L) of the attention-probability dropout (tip_common.cuh attn_drop_index): queries r
    and r + 8 of a 16-row tile x two adjacent keys share one hash group."""
    u = np.uint64
    bh = (np.arange(B, dtype=u)[:, None] * u(n_heads) + np.arange(n_heads, dtype=u))[:, :, None, None]
    q = np.arange(L, dtype=u)[None, None, :, None]
    k = np.arange(L, dtype=u)[None, None, None, :]
    grp = ((bh * u(3) + (q >> u(4))) * u(8) + (q & u(7))) * u(20) + (k >> u(1))
    return u(4) * grp + (u(2) * ((q >> u(3)) & u(1)) + (k & u(1)))


def forward(sd, x_imu, x_s, n_heads=16, with_rnn=True, keep_mask=None, past_scale=1.0,
            dtype=np.float32, return_intermediates=False, dropout=None):
    """Restatement of TF_RNN_Past_State.forward: deterministic (``dropout=None``), or as shipped with the
    product's mask generator: ``dropout = dict(seed=, in_dropout=, past_state_dropout=, encoder_dropout=)``
    (reference :73, :77 and the nn.TransformerEncoderLayer dropouts: on the attention probabilities, on the
    attention block's output before the residual (dropout1), after the FFN's ReLU, on the FFN's output
    before the residual (dropout2)).  B <= 1024 (one chunk of the product).

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
    dp = dict(seed=0, in_dropout=0.0, past_state_dropout=0.0, encoder_dropout=0.0)
    dp.update(dropout or {})
    seed, p_enc = int(dp["seed"]), float(dp["encoder_dropout"])
    n_imu, size_s = x_imu.shape[2], x_s.shape[2]
    kin_pad = -(-(n_imu + size_s) // 64) * 64
    rows = np.arange(B * L, dtype=np.uint64).reshape(B, L, 1)
    if dp["in_dropout"] > 0:                                        # :73
        idx = rows * np.uint64(kin_pad) + np.arange(n_imu, dtype=np.uint64)
        x_imu = x_imu * dropout_factors(seed + SEED_IN, idx, dp["in_dropout"], dtype)
    if keep_mask is not None:                                       # :77 with an explicit mask
        x_s = x_s * (np.asarray(keep_mask, dtype=dtype) * dtype(past_scale))
    elif dp["past_state_dropout"] > 0:                              # :77 with the product's generator
        idx = rows * np.uint64(kin_pad) + np.uint64(n_imu) + np.arange(size_s, dtype=np.uint64)
        x_s = x_s * dropout_factors(seed + SEED_PAST, idx, dp["past_state_dropout"], dtype)
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
        if p_enc > 0:                                               # attention-probability dropout
            idx = attn_drop_index(B, n_heads, L)
            pr = pr * dropout_factors(seed + seed_attn(i), idx, p_enc, dtype)
        o = (pr @ v).transpose(0, 2, 1, 3).reshape(B, L, E)
        a = o @ W[p + "self_attn.out_proj.weight"].T + W[p + "self_attn.out_proj.bias"]
        if p_enc > 0:                                               # dropout1
            a = a * dropout_factors(seed + seed_out(i), rows * np.uint64(E) + np.arange(E, dtype=np.uint64), p_enc, dtype)
        x = _layer_norm(x + a, W[p + "norm1.weight"], W[p + "norm1.bias"])
        f = np.maximum(x @ W[p + "linear1.weight"].T + W[p + "linear1.bias"], 0)
        if p_enc > 0:                                               # dropout inside the FFN (after the activation)
            Fh = f.shape[-1]
            f = f * dropout_factors(seed + seed_ff1(i), rows * np.uint64(Fh) + np.arange(Fh, dtype=np.uint64), p_enc, dtype)
        f = f @ W[p + "linear2.weight"].T + W[p + "linear2.bias"]
        if p_enc > 0:                                               # dropout2
            f = f * dropout_factors(seed + seed_ff2(i), rows * np.uint64(E) + np.arange(E, dtype=np.uint64), p_enc, dtype)
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


# ---------------------------------------------------------------------------------------------
# Post-model step (row N3): restatement of the model-visible part of RTRunnerMin.step after the model
# call (real_time_runner_minimal.py:87-112 smooth_and_split_s_c, :150-167 state assembly, :78-85/:196
# record_state_aa_and_c) and of the rotation-representation helpers it uses (data_utils.py:164-187,
# fairmotion conversions A2R / R2A = scipy Rotation).  PyBullet FK and the SBP root correction
# (:169-194) only produce the root translation s_t[0:3], which is never fed back to the model, and
# stay on the CPU (out of scope).  Closed-form numpy, float64 -- except where the reference itself
# computes in float32 (see PostProcessor.step).

N_DOFS = 57                 # constants.py:24
DT = 1.0 / 60               # constants.py:7


def nearest_rotation(M):
    """Special-orthogonal estimate of (n,3,3) matrices (orthogonal Procrustes, U V^T) -- what
    scipy >= 1.11 ``Rotation.from_matrix`` applies to non-orthogonal input (the reference reaches it through
    fairmotion ``conversions.R2A``, data_utils.py:177; the model's 2-axis output is never exactly
    orthonormal).  Older scipy skipped this step (version-dependent reference behaviour; goldens were
    minted with scipy 1.18)."""
    M = np.asarray(M, dtype=np.float64)
    U, _, Vt = np.linalg.svd(M)
    R = U @ Vt
    neg = np.linalg.det(R) < 0
    if np.any(neg):
        U = U.copy()
        U[neg, :, -1] *= -1
        R = U @ Vt
    return R


def rotmat_to_quat(R):
    """(n,3,3) rotation matrices -> (n,4) unit quaternions xyzw (Markley / Shepperd decision method)."""
    R = np.asarray(R, dtype=np.float64)
    n = R.shape[0]
    dec = np.empty((n, 4))
    dec[:, 0], dec[:, 1], dec[:, 2] = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    dec[:, 3] = dec[:, :3].sum(axis=1)
    ch = dec.argmax(axis=1)
    q = np.empty((n, 4))
    for idx in range(n):
        c, Rm = ch[idx], R[idx]
        if c != 3:
            i, j, k = c, (c + 1) % 3, (c + 2) % 3
            q[idx, i] = 1 - dec[idx, 3] + 2 * Rm[i, i]
            q[idx, j] = Rm[j, i] + Rm[i, j]
            q[idx, k] = Rm[k, i] + Rm[i, k]
            q[idx, 3] = Rm[k, j] - Rm[j, k]
        else:
            q[idx, 0] = Rm[2, 1] - Rm[1, 2]
            q[idx, 1] = Rm[0, 2] - Rm[2, 0]
            q[idx, 2] = Rm[1, 0] - Rm[0, 1]
            q[idx, 3] = 1 + dec[idx, 3]
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def quat_to_aa(q):
    """(n,4) xyzw -> (n,3) rotation vectors with angle in [0, pi] (scipy as_rotvec)."""
    q = np.array(q, dtype=np.float64)
    q[q[:, 3] < 0] *= -1
    nrm = np.linalg.norm(q[:, :3], axis=1)
    angle = 2 * np.arctan2(nrm, q[:, 3])
    small = angle <= 1e-3
    scale = np.empty_like(angle)
    a2 = angle[small] ** 2
    scale[small] = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    scale[~small] = angle[~small] / np.sin(angle[~small] / 2)
    return q[:, :3] * scale[:, None]


def rotmat_to_aa(R):
    """fairmotion conversions.R2A == scipy Rotation.from_matrix(R).as_rotvec()."""
    return quat_to_aa(rotmat_to_quat(nearest_rotation(R)))


def aa_to_rotmat(A):
    """fairmotion conversions.A2R == scipy Rotation.from_rotvec(A).as_matrix() (Rodrigues via quaternion)."""
    A = np.asarray(A, dtype=np.float64)
    angle = np.linalg.norm(A, axis=1)
    small = angle <= 1e-3
    scale = np.empty_like(angle)
    a2 = angle[small] ** 2
    scale[small] = 0.5 - a2 / 48 + a2 * a2 / 3840
    scale[~small] = np.sin(angle[~small] / 2) / angle[~small]
    x, y, z = (A * scale[:, None]).T
    w = np.cos(angle / 2)
    R = np.empty((A.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rot6_to_aa(rm):
    """data_utils.py:164-179 batch_rot_mat_2axis_to_aa for one frame: (Nj*6,) -> (Nj*3,).  Runs in the dtype
    of ``rm`` up to the R2A call, like the reference (float32 while the runner passes the raw model row)."""
    rm = np.asarray(rm).reshape(-1, 3, 2)
    a1 = rm[:, :, 0] / (np.linalg.norm(rm[:, :, 0], axis=1, keepdims=True) + 1e-6)
    a2 = rm[:, :, 1] / (np.linalg.norm(rm[:, :, 1], axis=1, keepdims=True) + 1e-6)
    a3 = np.cross(a1, a2)
    full = np.stack((a1, a2, a3), axis=2)
    return rotmat_to_aa(full).reshape(-1)


def state_to_row(cur_s, cur_c):
    """real_time_runner_minimal.py:78-85 + data_utils.py:182-187: (114,) qdq and (20,) constraints ->
    the (131,) row appended to s_and_c_in_buffer."""
    aa = np.asarray(cur_s, dtype=np.float64)[3:N_DOFS].reshape(-1, 3)
    r = aa_to_rotmat(aa)[:, :, :2].reshape(-1)
    return np.concatenate((r, cur_s[N_DOFS:N_DOFS + 3], cur_c))


class PostProcessor:
    """One stream's post-model state machine.  ``step(y_last, root_R_row)`` consumes the model's last
    output row (float32, as ``y.squeeze(0)[-1].numpy()`` :150) and the 9 root-rotation entries of the
    newest window row (:163) and returns (s_t[3:60], c_t, next_row) where next_row is what
    record_state_aa_and_c appends for the next call."""

    def __init__(self, n_sbps=5):
        self.n_c = n_sbps * 4
        self.coeff = 0.6 ** np.arange(6)[::-1]                        # :57
        self.buf = []
        self.last_tail = None                                           # last_s[6:60]

    def step(self, y_last, root_R_row, with_root_v=False):
        y = np.array(y_last, dtype=np.float32)                          # the buffer keeps THIS array (:91)
        self.buf.append(y)
        if len(self.buf) >= 6:                                          # :93-96
            s = np.array(self.buf[-6:]) * self.coeff[:, None]
            s_smooth = np.sum(s, axis=0) / np.sum(self.coeff)
        else:                                                           # :98: a float32 VIEW of the buffer
            s_smooth = y                                                # entry -- :107-110 edit it in place
        st = s_smooth[:-self.n_c]
        c_t = s_smooth[-self.n_c:]
        c_t[0::4] = (c_t[0::4] > 0.0) * 1.0                            # :107
        c_t[1::4] /= 5.0
        c_t[2::4] /= 5.0
        c_t[3::4] /= 5.0
        root_v = st[-3:]                                                # :154 (integrated into the root position at :159)
        root_v_out = np.array(root_v, dtype=np.float64)
        st_aa = rot6_to_aa(st[:-3])                                     # :155
        tail = np.zeros(N_DOFS - 6 + 3)                                 # s_t[6:60]
        tail[:N_DOFS - 6] = st_aa[3:]                                   # :160
        tail[N_DOFS - 6:] = root_v                                      # :158
        root_aa = rotmat_to_aa(np.asarray(root_R_row, dtype=np.float64).reshape(1, 3, 3))[0]   # :161-162
        if self.last_tail is not None:                                  # :165-166 (s_t[60:] stays 0)
            tail = (tail + self.last_tail) / 2.0
        self.last_tail = tail.copy()
        s = np.concatenate((root_aa, tail))                             # s_t[3:60]
        row = np.concatenate((aa_to_rotmat(s[:54].reshape(-1, 3))[:, :, :2].reshape(-1), s[54:57], c_t))
        if with_root_v:
            return s, np.array(c_t, dtype=np.float64), row, root_v_out
        return s, np.array(c_t, dtype=np.float64), row
