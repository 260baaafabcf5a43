"""Python mirror of the reference's model-call surface for the TIP hot path.

``TF_RNN_Past_State`` here keeps the constructor, the 56-key ``state_dict`` and the
``forward(x_imu, x_s)`` contract of the reference class
(/root/reference/simple_transformer_with_state.py:8-102) so the reference's consumers
(offline_testing_simple.py:79-99, live_demo_new.py:202-212, real_time_runner_minimal.py:149,
real_time_runner.py:431) can import it by the same module name and run unmodified.  All
arithmetic happens in ``libtip_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/tip_b200.h``); torch is used for parameter storage, device memory and streams only.
There is no CPU path: calling the module with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
import operator
import math

import torch
from torch import nn

from . import capi


def state_dict_keys(tf_layers: int = 4, with_rnn: bool = True):
    """Key order of the reference module's state_dict (56 tensors for the shipped checkpoints);
    tip_pack_weights takes its pointers in this order."""
    keys = ["in_linear.weight", "in_linear.bias"]
    for i in range(tf_layers):
        p = f"tf_encode.layers.{i}."
        keys += [p + s for s in ("self_attn.in_proj_weight", "self_attn.in_proj_bias",
                                 "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                                 "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                                 "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")]
    if with_rnn:
        keys += ["rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0"]
    return keys + ["linear.weight", "linear.bias"]


class _Bag(nn.Module):
    """Parameter container; exists only so state_dict keys match the reference's."""


_VERSION_OF = operator.attrgetter("_version")


def _linear_bag(n_out: int, n_in: int, gen: torch.Generator) -> _Bag:
    b = _Bag()
    bound = 1.0 / math.sqrt(n_in)
    b.weight = nn.Parameter((torch.rand(n_out, n_in, generator=gen) * 2 - 1) * bound)
    b.bias = nn.Parameter((torch.rand(n_out, generator=gen) * 2 - 1) * bound)
    return b


def _norm_bag(n: int) -> _Bag:
    b = _Bag()
    b.weight = nn.Parameter(torch.ones(n))
    b.bias = nn.Parameter(torch.zeros(n))
    return b


class TF_RNN_Past_State(nn.Module):
    """Drop-in for the reference class of the same name (reference :8).

    Stochastic behaviour follows the reference as shipped: ``past_state_dropout`` and
    ``in_dropout`` are applied on EVERY call (the reference builds fresh ``nn.Dropout`` modules,
    :73/:77, which ignore ``eval()``), the encoder's p=0.1 dropouts only in ``train()`` mode.  The
    masks come from a counter-based generator seeded from torch's CPU generator (so
    ``torch.manual_seed`` makes runs repeatable) -- statistically equivalent to, not bit-equal
    with, torch's Philox stream.  Deterministic parity mode: ``m.eval()`` and
    ``m.past_state_dropout = 0.0``.
    """

    ENCODER_DROPOUT = 0.1   # nn.TransformerEncoderLayer default; the reference never overrides it (:26-28)

    def __init__(self, input_size_imu, size_s, rnn_hid_size, tf_hid_size, tf_in_dim, n_heads,
                 tf_layers, dropout, in_dropout, past_state_dropout, with_rnn=True,
                 with_acc_sum=False):
        super().__init__()
        gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        d_in = input_size_imu + size_s + (18 if with_acc_sum else 0)
        if with_acc_sum:
            print("model with acc sum")                      # reference :21
        self.in_linear = _linear_bag(tf_in_dim, d_in, gen)
        self.tf_encode = _Bag()
        self.tf_encode.layers = nn.ModuleList()
        for _ in range(tf_layers):
            layer = _Bag()
            layer.self_attn = _Bag()
            qkv = _linear_bag(3 * tf_in_dim, tf_in_dim, gen)
            layer.self_attn.in_proj_weight = qkv.weight
            layer.self_attn.in_proj_bias = qkv.bias
            layer.self_attn.out_proj = _linear_bag(tf_in_dim, tf_in_dim, gen)
            layer.linear1 = _linear_bag(tf_hid_size, tf_in_dim, gen)
            layer.linear2 = _linear_bag(tf_in_dim, tf_hid_size, gen)
            layer.norm1 = _norm_bag(tf_in_dim)
            layer.norm2 = _norm_bag(tf_in_dim)
            self.tf_encode.layers.append(layer)
        self.with_rnn = with_rnn
        if with_rnn:
            self.rnn = _Bag()
            ih = _linear_bag(rnn_hid_size, tf_in_dim, gen)
            hh = _linear_bag(rnn_hid_size, rnn_hid_size, gen)
            self.rnn.weight_ih_l0, self.rnn.weight_hh_l0 = ih.weight, hh.weight
            self.rnn.bias_ih_l0, self.rnn.bias_hh_l0 = ih.bias, hh.bias
            self.linear = _linear_bag(size_s, rnn_hid_size, gen)
        else:
            print("no RNN layer")                            # reference :44
            self.rnn = None
            self.linear = _linear_bag(size_s, tf_in_dim, gen)

        self.rnn_hid_size = rnn_hid_size
        self.in_dropout = in_dropout
        self.n_heads = n_heads
        self.past_state_dropout = past_state_dropout
        print("number of parameters: %e", sum(p.numel() for p in self.parameters()))   # :54

        self._dims = capi.TipDims(input_size_imu, size_s, rnn_hid_size, tf_hid_size, tf_in_dim,
                                  n_heads, tf_layers, int(bool(with_rnn)), int(bool(with_acc_sum)))
        self._size_s = size_s
        self._n_imu = input_size_imu + (18 if with_acc_sum else 0)
        self._lib = None
        self._handle = None
        self._device = None
        self._packed_sig = None
        self._packed_versions = None
        self._plist = None
        self._owner = None          # lanes: the module whose packed weights this handle shares
        self._owner_handle = None

    # -------------------------------------------------------------------------------------------
    def _ordered_params(self):
        """The tensors in the reference's state_dict() order (= tip_pack_weights order)."""
        if self._plist is None:     # Parameter objects are stable across load_state_dict / .cuda()
            sd = self.state_dict(keep_vars=True)
            self._plist = [sd[k] for k in state_dict_keys(self._dims.tf_layers,
                                                          bool(self._dims.with_rnn))]
        return self._plist

    def _release(self):
        if self._handle is not None and self._lib is not None:
            self._lib.tip_destroy(self._handle)
        self._handle = None
        self._packed_sig = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, device: torch.device, fast: bool = False):
        """Create the C handle on ``device`` and (re)pack when any parameter changed
        (load_state_dict, .cuda(), in-place optimiser updates bump ``_version``).  ``fast`` (the streaming
        session's per-frame calls): compare the version counters only; storage swaps through ``_apply`` reset
        the signature, a bare ``param.data = tensor`` is picked up by the next full check (any ``forward``)."""
        if self._lib is None:
            self._lib = capi.load_library()
        if self._owner is not None:
            # a lane: the owner packs (once, for all lanes); this handle shares its packed weights (tip_create_lane)
            oh = self._owner._ensure(device, fast=fast)
            if self._handle is None or self._device != device or self._owner_handle != oh.value:
                self._release()
                with torch.cuda.device(device):
                    h = C.c_void_p()
                    rc = self._lib.tip_create_lane(oh, C.byref(h))
                    capi.check(self._lib, None, rc, "tip_create_lane")
                self._handle, self._device, self._owner_handle = h, device, oh.value
            return self._handle
        if self._handle is None or self._device != device:
            self._release()
            with torch.cuda.device(device):
                h = C.c_void_p()
                rc = self._lib.tip_create(C.byref(self._dims), C.byref(h))
                capi.check(self._lib, None, rc, "tip_create")
            self._handle, self._device = h, device
        params = self._ordered_params()
        if fast and self._packed_versions is not None and self._packed_sig is not None and \
                tuple(map(_VERSION_OF, params)) == self._packed_versions:
            return self._handle          # per-frame callers: version counters only (6 us instead of 15 us of Python)
        sig = tuple((p.data_ptr(), p._version) for p in params)
        if sig != self._packed_sig:
            for p in params:
                if p.device != device or p.dtype != torch.float32:
                    raise RuntimeError(
                        "TF_RNN_Past_State: parameters must be fp32 on the input's CUDA device "
                        f"(got {p.dtype} on {p.device}; call .cuda() as the reference does)")
            keep = [p.detach().contiguous() for p in params]
            n = len(keep)
            ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
            numels = (C.c_int64 * n)(*[t.numel() for t in keep])
            stream = torch.cuda.current_stream(device).cuda_stream
            with torch.cuda.device(device):
                rc = self._lib.tip_pack_weights(self._handle, ptrs, numels, n, C.c_void_p(stream))
            capi.check(self._lib, self._handle, rc, "tip_pack_weights")
            self._packed_sig = sig
        self._packed_versions = tuple(map(_VERSION_OF, params))
        return self._handle

    def _apply(self, fn, *args, **kwargs):
        # .cuda() / .to() / .float(): parameter storage may move without a version bump
        self._packed_sig = None
        self._packed_versions = None
        return super()._apply(fn, *args, **kwargs)

    def _dropout_struct(self):
        p_enc = self.ENCODER_DROPOUT if self.training else 0.0
        p_in, p_past = float(self.in_dropout), float(self.past_state_dropout)
        if p_enc == 0.0 and p_in == 0.0 and p_past == 0.0:
            return None
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        return capi.TipDropout(p_in, p_past, p_enc, seed & 0xFFFFFFFFFFFFFFFF)

    def _check_inputs(self, x_imu, x_s):
        if not (x_imu.is_cuda and x_s.is_cuda):
            raise RuntimeError("TF_RNN_Past_State (tip_b200): inputs must be CUDA tensors; the "
                               "B200 hot path has no CPU fallback")
        if x_imu.dim() != 3 or x_s.dim() != 3 or x_imu.shape[:2] != x_s.shape[:2]:
            raise RuntimeError(f"expected x_imu (B,L,{self._n_imu}) and x_s (B,L,{self._size_s}), "
                               f"got {tuple(x_imu.shape)} and {tuple(x_s.shape)}")
        if x_imu.shape[2] != self._n_imu or x_s.shape[2] != self._size_s:
            raise RuntimeError(f"expected feature widths {self._n_imu} and {self._size_s}, got "
                               f"{x_imu.shape[2]} and {x_s.shape[2]}")
        return (x_imu.detach().to(torch.float32).contiguous(),
                x_s.detach().to(torch.float32).contiguous())

    def forward(self, x_imu, x_s, keep_mask=None, past_scale=1.0, out=None):
        """(B, L, 72|90), (B, L, size_s) -> (B, L, size_s); reference :60-102.  Inputs are not
        modified.  ``keep_mask`` (test hook): explicit 0/1 mask used instead of drawing the
        past-state dropout mask; x_s is multiplied by keep_mask * past_scale.  ``out`` (beyond the
        reference surface): pre-allocated contiguous fp32 CUDA tensor of the result's shape; with stable
        input / output addresses a repeated forward is one CUDA-graph replay."""
        x_imu, x_s = self._check_inputs(x_imu, x_s)
        dev = x_imu.device
        h = self._ensure(dev)
        B, L = x_imu.shape[0], x_imu.shape[1]
        if out is None:
            y = torch.empty((B, L, self._size_s), dtype=torch.float32, device=dev)
        else:
            y = out
            if y.device != dev or y.dtype != torch.float32 or tuple(y.shape) != (B, L, self._size_s) \
                    or not y.is_contiguous():
                raise RuntimeError(f"out must be a contiguous fp32 tensor of shape {(B, L, self._size_s)} on {dev}")
        if B == 0:
            return y
        drop = self._dropout_struct()
        km = None
        if keep_mask is not None:
            km = keep_mask.detach().to(device=dev, dtype=torch.float32).contiguous()
            if drop is not None:
                drop.past_state_dropout = 0.0
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = self._lib.tip_forward(h, x_imu.data_ptr(), x_s.data_ptr(), y.data_ptr(), B, L,
                                       km.data_ptr() if km is not None else None,
                                       float(past_scale), C.byref(drop) if drop else None,
                                       C.c_void_p(stream))
        capi.check(self._lib, h, rc, "tip_forward")
        return y

    # ---- extras beyond the reference surface --------------------------------------------------
    def make_lane(self):
        """A second execution lane: a module that SHARES this module's Parameter objects (no copy; a
        ``load_state_dict`` / optimiser step / ``.cuda()`` on either is seen by both) and its PACKED weights
        (``tip_create_lane``: one packed copy per GPU, packed once) but owns its own C handle, i.e. its own
        workspace, captured graphs and job slots.  Forwards of
        different lanes may run concurrently on different CUDA streams -- the LayerNorm GEMMs (80 row
        tiles at B = 256) and the recurrence (104 SMs) leave SMs idle that another lane's kernels fill:
        two lanes give 598 k instead of 498 k frames/s at B = 256.  Dropout settings and train/eval mode
        are copied at creation (``ForwardLanes`` / ``HostPipeline`` refresh them on every call)."""
        lane = object.__new__(type(self))
        nn.Module.__init__(lane)
        for name, child in self._modules.items():
            lane._modules[name] = child
        for k in ("with_rnn", "rnn_hid_size", "in_dropout", "n_heads", "past_state_dropout", "_dims",
                  "_size_s", "_n_imu"):
            setattr(lane, k, getattr(self, k))
        if "rnn" not in self._modules:
            lane.rnn = None
        lane.training = self.training
        lane._lib = self._lib
        lane._handle = lane._device = lane._packed_sig = lane._packed_versions = lane._plist = None
        # (plain attribute, not a registered sub-module: the lane's state_dict stays the reference's 56 keys)
        object.__setattr__(lane, "_owner", self if self._owner is None else self._owner)
        lane._owner_handle = None
        return lane

    def _sync_lane_settings(self, src):
        self.in_dropout, self.past_state_dropout, self.training = src.in_dropout, src.past_state_dropout, src.training

    def forward_host(self, x_imu, x_s, last_row_only=False, out=None):
        """``model(x_imu.cuda(), x_s.cuda()).cpu()`` in one C call with host (numpy / CPU tensor)
        buffers: H2D, forward, D2H, stream sync (real_time_runner_minimal.py:149).  Runs on the
        device the parameters live on.  Pinned buffers (``tensor.pin_memory()``) are used in place and,
        the forward is a CUDA-graph replay from the second call on; pageable
        buffers go through the handle's pinned staging.  ``out``: optional pre-allocated CPU tensor."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("forward_host: move the module to a CUDA device first (.cuda())")
        h = self._ensure(dev, fast=True)
        xi = torch.as_tensor(x_imu, dtype=torch.float32).contiguous()
        xs = torch.as_tensor(x_s, dtype=torch.float32).contiguous()
        if xi.is_cuda or xs.is_cuda:
            raise RuntimeError("forward_host takes host buffers; call the module itself with CUDA tensors")
        if xi.dim() != 3 or xs.dim() != 3 or xi.shape[:2] != xs.shape[:2] or xi.shape[2] != self._n_imu \
                or xs.shape[2] != self._size_s:
            raise RuntimeError(f"expected x_imu (B,L,{self._n_imu}) and x_s (B,L,{self._size_s}), "
                               f"got {tuple(xi.shape)} and {tuple(xs.shape)}")
        B, L = xi.shape[0], xi.shape[1]
        shape = (B, self._size_s) if last_row_only else (B, L, self._size_s)
        if out is None:
            y = torch.empty(shape, dtype=torch.float32)
        else:
            y = out
            if y.is_cuda or y.dtype != torch.float32 or tuple(y.shape) != shape or not y.is_contiguous():
                raise RuntimeError(f"forward_host: out must be a contiguous fp32 CPU tensor of shape {shape}")
        drop = self._dropout_struct()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = self._lib.tip_forward_host(h, xi.data_ptr(), xs.data_ptr(), y.data_ptr(), B, L,
                                            int(last_row_only), C.byref(drop) if drop else None,
                                            C.c_void_p(stream))
        capi.check(self._lib, h, rc, "tip_forward_host")
        return y

    def forward_host_submit(self, slot, x_imu, x_s, out, last_row_only=False):
        """Queue one ``forward_host`` job in job slot ``slot`` (0..3) and return at once
        (tip_forward_host_submit): upload, forward and download run on the handle's own streams and
        overlap with the other slots' jobs.  All three tensors must be pinned, contiguous fp32 CPU
        tensors and must not be touched until ``forward_host_wait(slot)`` returns."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("forward_host_submit: move the module to a CUDA device first (.cuda())")
        h = self._ensure(dev)          # full signature check: a lane does not see the owner's _apply
        for name, t in (("x_imu", x_imu), ("x_s", x_s), ("out", out)):
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() \
                    or not t.is_pinned():
                raise RuntimeError(f"forward_host_submit: {name} must be a pinned, contiguous fp32 CPU tensor "
                                   "(tensor.pin_memory())")
        if x_imu.dim() != 3 or x_s.dim() != 3 or x_imu.shape[:2] != x_s.shape[:2] or x_imu.shape[2] != self._n_imu \
                or x_s.shape[2] != self._size_s:
            raise RuntimeError(f"expected x_imu (B,L,{self._n_imu}) and x_s (B,L,{self._size_s}), "
                               f"got {tuple(x_imu.shape)} and {tuple(x_s.shape)}")
        B, L = x_imu.shape[0], x_imu.shape[1]
        shape = (B, self._size_s) if last_row_only else (B, L, self._size_s)
        if tuple(out.shape) != shape:
            raise RuntimeError(f"forward_host_submit: out must have shape {shape}")
        drop = self._dropout_struct()
        with torch.cuda.device(dev):
            rc = self._lib.tip_forward_host_submit(h, int(slot), x_imu.data_ptr(), x_s.data_ptr(), out.data_ptr(),
                                                   B, L, int(last_row_only), C.byref(drop) if drop else None)
        capi.check(self._lib, h, rc, "tip_forward_host_submit")

    def forward_host_wait(self, slot):
        """Block until the job in ``slot`` is complete (its ``out`` is filled); no-op for an idle slot."""
        if self._handle is None:
            return
        with torch.cuda.device(self._device):
            rc = self._lib.tip_forward_host_wait(self._handle, int(slot))
        capi.check(self._lib, self._handle, rc, "tip_forward_host_wait")

    def set_gemm_engine(self, engine: int):
        """0 auto (= 2), 1 FFMA fp32 cross-check kernels, 2 tcgen05 3xFP16-split kernels (tip_set_gemm_engine)."""
        dev = next(self.parameters()).device
        h = self._ensure(dev)
        capi.check(self._lib, h, self._lib.tip_set_gemm_engine(h, int(engine)), "tip_set_gemm_engine")

    def set_tuning(self, key: str, value: int):
        """Kernel-selection knob of the tcgen05 engine (tip_set_tuning): "atm", "atm_grid", "atm_min_tiles", "dyn_sched",
        "ln_pair", "ln_grid", "ln_share", "rnn_clusters", "atm_pair", "attn_grid"; see include/tip_b200.h.  Per handle: lanes made by make_lane() have their own."""
        dev = next(self.parameters()).device
        h = self._ensure(dev)
        capi.check(self._lib, h, self._lib.tip_set_tuning(h, key.encode(), int(value)), "tip_set_tuning")

    def set_use_graphs(self, enable: bool):
        """CUDA-graph replay of repeated forwards / streaming frames (tip_set_use_graphs); on by default."""
        dev = next(self.parameters()).device
        h = self._ensure(dev)
        capi.check(self._lib, h, self._lib.tip_set_use_graphs(h, int(enable)), "tip_set_use_graphs")

    def algorithmic_cost(self, B: int, L: int):
        """(bytes, flops) of one forward as SURVEY.md 8d defines them."""
        dev = next(self.parameters()).device
        h = self._ensure(dev)
        b, f = C.c_double(), C.c_double()
        capi.check(self._lib, h, self._lib.tip_algorithmic_cost(h, B, L, C.byref(b), C.byref(f)),
                   "tip_algorithmic_cost")
        return b.value, f.value

    def set_profile(self, enable: bool):
        dev = next(self.parameters()).device
        h = self._ensure(dev)
        capi.check(self._lib, h, self._lib.tip_set_profile(h, int(enable)), "tip_set_profile")

    def profile(self):
        """[(stage name, layer, milliseconds)] of the last forward (needs set_profile(True))."""
        out = []
        buf = C.create_string_buffer(64)
        for i in range(self._lib.tip_profile_stages(self._handle)):
            layer, ms = C.c_int(), C.c_float()
            rc = self._lib.tip_profile_get(self._handle, i, buf, 64, C.byref(layer), C.byref(ms))
            capi.check(self._lib, self._handle, rc, "tip_profile_get")
            out.append((buf.value.decode(), layer.value, ms.value))
        return out

    def last_launch_count(self) -> int:
        return int(self._lib.tip_last_launch_count(self._handle)) if self._handle else 0

    def debug_tensor(self, name: str, cols: int):
        """Copy of an internal activation buffer of the last forward (test hook)."""
        n = C.c_int64()
        rc = self._lib.tip_debug_tensor(self._handle, name.encode(), None, 0, C.byref(n), None)
        capi.check(self._lib, self._handle, rc, "tip_debug_tensor")
        out = torch.empty(n.value, dtype=torch.float32, device=self._device)
        stream = torch.cuda.current_stream(self._device).cuda_stream
        rc = self._lib.tip_debug_tensor(self._handle, name.encode(), out.data_ptr(), n.value,
                                        C.byref(n), C.c_void_p(stream))
        capi.check(self._lib, self._handle, rc, "tip_debug_tensor")
        return out.view(-1, cols)
