"""CPU PyTorch restatement of the reference forward -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
file.  It restates ``TF_RNN_Past_State.forward``
(/root/reference/simple_transformer_with_state.py:60-102, deterministic mode of SURVEY.md 8c) with
the same multi-threaded ATen CPU operators the reference module dispatches to (``addmm`` via
``F.linear``, ``layer_norm``, softmax, ``tanh``), so that timing it on the GPU box's host cores
is a fair stand-in for "the reference's own CPU PyTorch path": /root/reference does not exist
on the GPU box and reference sources may not be copied into this repo.  It is the same
algorithm as ``oracle/tip_oracle.py`` (numpy) and is pinned by the same golden vectors
(tests/test_oracle.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def to_torch_state(sd):
    return {k: torch.as_tensor(v, dtype=torch.float32) for k, v in sd.items()}


@torch.no_grad()
def forward(W, x_imu, x_s, n_heads=16, with_rnn=True):
    x_imu = torch.as_tensor(x_imu, dtype=torch.float32).clone()             # :63
    x_s = torch.as_tensor(x_s, dtype=torch.float32).clone()                 # :64
    x_s = torch.nan_to_num(x_s, nan=0.0, posinf=float("inf"), neginf=float("-inf"))   # :65
    B, L = x_imu.shape[0], x_imu.shape[1]
    x_s[:, :, 108:111] = 0.0                                                # :75
    x = F.linear(torch.cat((x_imu, x_s), dim=2), W["in_linear.weight"], W["in_linear.bias"])  # :78-79
    E = x.shape[-1]
    d = E // n_heads
    x = x.view(B, L, n_heads, d).transpose(2, 3).reshape(B, L, E)           # :88-89
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    n_layers = 1 + max(int(k.split(".")[2]) for k in W if k.startswith("tf_encode.layers."))
    for i in range(n_layers):                                               # :91
        p = f"tf_encode.layers.{i}."
        qkv = F.linear(x, W[p + "self_attn.in_proj_weight"], W[p + "self_attn.in_proj_bias"])
        q, k, v = (t.view(B, L, n_heads, d).transpose(1, 2) for t in qkv.split(E, dim=-1))
        s = (q * (d ** -0.5)) @ k.transpose(-1, -2)
        s = s.masked_fill(~causal, float("-inf"))
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, E)
        a = F.linear(o, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + a, (E,), W[p + "norm1.weight"], W[p + "norm1.bias"], 1e-5)
        f = F.linear(torch.relu(F.linear(x, W[p + "linear1.weight"], W[p + "linear1.bias"])),
                     W[p + "linear2.weight"], W[p + "linear2.bias"])
        x = F.layer_norm(x + f, (E,), W[p + "norm2.weight"], W[p + "norm2.bias"], 1e-5)
    if with_rnn:                                                            # :95-99
        gi = F.linear(x, W["rnn.weight_ih_l0"], W["rnn.bias_ih_l0"])
        Whh, bhh = W["rnn.weight_hh_l0"], W["rnn.bias_hh_l0"]
        h = torch.zeros(B, Whh.shape[0])
        hs = []
        for t in range(L):
            h = torch.tanh(gi[:, t] + F.linear(h, Whh, bhh))
            hs.append(h)
        x = torch.stack(hs, dim=1)
    return F.linear(x, W["linear.weight"], W["linear.bias"])                # :102
