"""CPU: the C-ABI library loads, exports every symbol include/tip_b200.h declares, and the host
logic of the Python mirror (state-dict surface, error behaviour) holds.  No compute calls."""
import contextlib
import ctypes as C
import io
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_checkpoint
from oracle import tip_oracle as O
from tip_b200 import TF_RNN_Past_State, capi, state_dict_keys


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "tip_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(tip_[a-z_0-9]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)
    assert lib.tip_abi_version() == 1


def test_create_rejects_unsupported_dims_without_touching_a_device():
    lib = capi.load_library()
    h = C.c_void_p()
    bad = capi.TipDims(72, 131, 512, 1024, 128, 16, 4, 1, 1)      # tf_in_dim != 256
    assert lib.tip_create(C.byref(bad), C.byref(h)) == 1          # TIP_ERR_INVALID_ARG
    assert b"unsupported" in lib.tip_last_error(None)
    assert not h.value


def _make(size_s=131, with_rnn=True, with_acc_sum=True):
    with contextlib.redirect_stdout(io.StringIO()):
        return TF_RNN_Past_State(72, size_s, rnn_hid_size=512, tf_hid_size=1024, tf_in_dim=256,
                                 n_heads=16, tf_layers=4, dropout=0.0, in_dropout=0.0,
                                 past_state_dropout=0.8, with_rnn=with_rnn,
                                 with_acc_sum=with_acc_sum)


@pytest.mark.parametrize("size_s,with_rnn,with_acc_sum",
                         [(131, True, True), (119, True, True), (131, False, True), (131, True, False)])
def test_state_dict_surface_matches_reference(size_s, with_rnn, with_acc_sum):
    m = _make(size_s, with_rnn, with_acc_sum)
    sd_ref = O.random_state_dict(3, size_s=size_s, with_rnn=with_rnn, with_acc_sum=with_acc_sum)
    sd = m.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys()) == state_dict_keys(4, with_rnn)
    for k in sd:
        assert tuple(sd[k].shape) == sd_ref[k].shape, k
        assert sd[k].dtype == torch.float32
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd_ref.items()})      # strict
    np.testing.assert_array_equal(m.in_linear.weight.detach().numpy(), sd_ref["in_linear.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: torch.from_numpy(v) for k, v in list(sd_ref.items())[:-1]})


def test_released_checkpoint_loads_strictly():
    sd = load_checkpoint("model-with-dip9and10")
    if sd is None:
        pytest.skip("baseline/_ref checkpoint not staged")
    m = _make()
    assert len(sd) == 56 and sum(v.size for v in sd.values()) == 3677315
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    assert sum(p.numel() for p in m.parameters()) == 3677315


def test_attributes_and_no_cpu_fallback():
    m = _make()
    assert m.past_state_dropout == 0.8 and m.in_dropout == 0.0
    assert m.n_heads == 16 and m.rnn_hid_size == 512
    m.past_state_dropout = 0.0
    assert m.eval() is m and not m.training
    x_imu, x_s = O.synth_inputs(0, 1, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.from_numpy(x_imu), torch.from_numpy(x_s))
