"""ctypes binding of ``include/tip_b200.h`` (the C-ABI drop-in boundary of the TIP hot path).

There is no CPU fallback: if ``libtip_b200.so`` is missing or a call fails, a ``RuntimeError``
is raised.  Build the library with ``python __graft_entry__.py build`` (or
``tip_b200.build.build_library()``); it is kept in-tree at
``transformer-inertial-poser_b200/lib/libtip_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libtip_b200.so")

TIP_OK = 0
TIP_HOST_SLOTS = 4          # job slots of tip_forward_host_submit / _wait
STATUS_NAMES = {0: "TIP_OK", 1: "TIP_ERR_INVALID_ARG", 2: "TIP_ERR_NOT_PACKED", 3: "TIP_ERR_CUDA",
                4: "TIP_ERR_NO_DEVICE", 5: "TIP_ERR_OOM"}


class TipDims(C.Structure):
    """tip_dims: constructor arguments of TF_RNN_Past_State (reference :9-17)."""
    _fields_ = [("input_size_imu", C.c_int32), ("size_s", C.c_int32), ("rnn_hid_size", C.c_int32),
                ("tf_hid_size", C.c_int32), ("tf_in_dim", C.c_int32), ("n_heads", C.c_int32),
                ("tf_layers", C.c_int32), ("with_rnn", C.c_int32), ("with_acc_sum", C.c_int32)]


class TipDropout(C.Structure):
    """tip_dropout: per-call stochastic behaviour (reference :73, :77 + encoder dropouts)."""
    _fields_ = [("in_dropout", C.c_float), ("past_state_dropout", C.c_float),
                ("encoder_dropout", C.c_float), ("seed", C.c_uint64)]


_VP = C.c_void_p
# every symbol include/tip_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tip_abi_version": (C.c_int, []),
    "tip_create": (C.c_int, [C.POINTER(TipDims), C.POINTER(_VP)]),
    "tip_destroy": (None, [_VP]),
    "tip_create_lane": (C.c_int, [_VP, C.POINTER(_VP)]),
    "tip_last_error": (C.c_char_p, [_VP]),
    "tip_num_weight_tensors": (C.c_int, [_VP]),
    "tip_pack_weights": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(C.c_int64), C.c_int, _VP]),
    "tip_forward": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, _VP, C.c_float,
                              C.POINTER(TipDropout), _VP]),
    "tip_forward_host": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(TipDropout), _VP]),
    "tip_forward_host_submit": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(TipDropout)]),
    "tip_forward_host_wait": (C.c_int, [_VP, C.c_int]),
    "tip_stream_reset": (C.c_int, [_VP, C.c_int]),
    "tip_stream_step": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.POINTER(TipDropout), _VP]),
    "tip_stream_length": (C.c_int, [_VP]),
    "tip_stream_step_raw": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.POINTER(TipDropout), _VP,
                                      C.POINTER(C.c_int)]),
    "tip_stream_set_state": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "tip_stream_state_width": (C.c_int, [_VP]),
    "tip_stream_step_closed": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.POINTER(TipDropout), _VP,
                                         C.POINTER(C.c_int)]),
    "tip_algorithmic_cost": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]),
    "tip_last_launch_count": (C.c_int, [_VP]),
    "tip_set_gemm_engine": (C.c_int, [_VP, C.c_int]),
    "tip_set_tuning": (C.c_int, [_VP, C.c_char_p, C.c_int]),
    "tip_set_use_graphs": (C.c_int, [_VP, C.c_int]),
    "tip_set_profile": (C.c_int, [_VP, C.c_int]),
    "tip_profile_stages": (C.c_int, [_VP]),
    "tip_profile_get": (C.c_int, [_VP, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_float)]),
    "tip_debug_tensor": (C.c_int, [_VP, C.c_char_p, _VP, C.c_int64, C.POINTER(C.c_int64), _VP]),
}

_lib = None


def load_library(path: str | None = None):
    """Load libtip_b200.so and attach prototypes.  Raises RuntimeError when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"tip_b200: CUDA extension {p} is not built (run `python __graft_entry__.py build`); "
            "there is no CPU fallback for the TIP hot path")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.tip_abi_version() != 2:
        raise RuntimeError("tip_b200: ABI version mismatch between capi.py and libtip_b200.so")
    if path is None:
        _lib = lib
    return lib


def check(lib, handle, rc: int, what: str):
    """Non-zero status -> RuntimeError carrying tip_last_error (the C side never throws)."""
    if rc != TIP_OK:
        msg = lib.tip_last_error(handle)
        msg = msg.decode() if msg else ""
        raise RuntimeError(f"{what}: {STATUS_NAMES.get(rc, rc)}: {msg}")
