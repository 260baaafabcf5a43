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
    assert lib.tip_abi_version() == 2


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


def test_make_lane_shares_parameters_and_owns_nothing_else():
    m = _make()
    m.past_state_dropout = 0.0
    lane = m.make_lane()
    assert type(lane) is type(m) and lane._handle is None
    assert list(lane.state_dict().keys()) == list(m.state_dict().keys())
    for a, b in zip(lane.parameters(), m.parameters()):
        assert a is b                                            # the same Parameter objects, not copies
    assert lane.past_state_dropout == 0.0 and lane.n_heads == 16 and lane.training == m.training
    m.train()
    m.in_dropout = 0.3
    lane._sync_lane_settings(m)
    assert lane.training and lane.in_dropout == 0.3
    norn = _make(with_rnn=False).make_lane()
    assert norn.rnn is None and "rnn.weight_hh_l0" not in norn.state_dict()


def test_host_pipeline_slot_rotation_and_order():
    """Host logic of HostPipeline with a recording stand-in for the module: jobs come back in submission order,
    never more than `depth` in flight, a (lane, slot) pair is never re-used while its job is in flight."""
    from tip_b200.pipeline import HostPipeline

    class Rec:
        def __init__(self, log, name="L0"):
            self.log, self.name, self.n = log, name, 0
            self.in_dropout = self.past_state_dropout = 0.0
            self.training = False

        def make_lane(self):
            self.n += 1
            return Rec(self.log, f"L{self.n}")

        def _sync_lane_settings(self, src):
            pass

        def forward_host_submit(self, slot, xi, xs, out, last_row_only=False):
            self.log.append(("submit", self.name, slot, out))

        def forward_host_wait(self, slot):
            self.log.append(("wait", self.name, slot))

    for depth, lanes in ((1, 1), (2, 1), (4, 1), (4, 2), (6, 3), (5, 2)):
        log = []
        pipe = HostPipeline(Rec(log), depth=depth, lanes=lanes)
        got = []
        for j in range(23):
            r = pipe.submit("xi", "xs", j)
            assert len(pipe) <= depth
            if r is not None:
                got.append(r[2])
        got += [r[2] for r in pipe.drain()]
        assert got == list(range(23))
        busy = {}
        for e in log:
            if e[0] == "submit":
                assert (e[1], e[2]) not in busy, (depth, lanes, e)
                busy[(e[1], e[2])] = e[3]
            else:
                busy.pop((e[1], e[2]))
        assert not busy
        assert {e[1] for e in log} == {f"L{i}" for i in range(lanes)}
    with pytest.raises(ValueError):
        HostPipeline(Rec([]), depth=5, lanes=1)
    with pytest.raises(ValueError):
        HostPipeline(Rec([]), depth=0)
