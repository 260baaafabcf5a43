"""GPU: rows N1 + N3 + N4 against the trace of the UNMODIFIED reference runner (tests/golden/
runner_min_trace.npz, minted by oracle/make_runner_golden.py from RTRunnerMin + the released checkpoint)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_checkpoint
from oracle import tip_oracle as O
from test_gpu_parity import make_model
from tip_b200.evaluate import run_motions
from tip_b200.streaming import StreamSession, state_to_row

pytestmark = pytest.mark.gpu
TRACE = os.path.join(GOLD, "runner_min_trace.npz")


@pytest.fixture(scope="module")
def trace():
    return np.load(TRACE)


@pytest.fixture(scope="module")
def model(trace):
    sd = load_checkpoint(str(trace["ckpt"]).replace(".pt", ""))
    if sd is None:
        pytest.skip("baseline/_ref checkpoint not staged")
    return make_model(sd)


def test_state_to_row_matches_reference(trace):
    np.testing.assert_allclose(state_to_row(trace["s_init"], np.zeros(20)), trace["s_and_c_in"][0], atol=1e-12)


def test_post_step_teacher_forced_matches_reference_runner(trace, model):
    """N3 in isolation: with the model's output row replaced by the one the reference model produced, the
    device post step must reproduce the reference runner's state to float64 round-off (1e-6 during the
    first five calls, which the reference computes in float32 -- see PostProcessor.step)."""
    sess = StreamSession(model, n_streams=1)
    sess.set_state(trace["s_init"])
    assert sess.state_width == 80
    pp = O.PostProcessor()
    k = 0
    for t in range(trace["imu"].shape[0]):
        yo = trace["y_last"][k][None] if t >= 5 else None
        st = sess.step_closed(trace["imu"][t][None], y_override=yo)
        if t < 5:
            assert st is None
            continue
        tol = 2e-6 if k < 5 else 1e-9
        np.testing.assert_allclose(st[0, :57], trace["qdq"][t][3:60], atol=tol, err_msg=f"call {t}")
        np.testing.assert_allclose(st[0, 57:77], trace["ct"][t], atol=tol, err_msg=f"call {t}")
        _, _, _, rv = pp.step(trace["y_last"][k], trace["x_imu_last"][k][:9], with_root_v=True)
        np.testing.assert_allclose(st[0, 77:80], rv, atol=tol, err_msg=f"root_v, call {t}")
        k += 1
    # the rows fed back on the device are what the reference runner appended to s_and_c_in_buffer
    win_s = sess.window("win_s").cpu().numpy()[0]
    np.testing.assert_allclose(win_s, trace["s_and_c_in"][k - 40:k].astype(np.float32), atol=1e-6)
    win_i = sess.window("win_imu").cpu().numpy()[0]
    np.testing.assert_allclose(win_i[-1], trace["x_imu_last"][k - 1], atol=1e-6)


def test_closed_loop_tracks_reference_runner(trace, model):
    """N1 + forward + N3 free-running for 145 model calls: every frame's pose stays within the model's own
    fp32 tolerance band of the reference runner (feedback does not amplify the 1e-5 forward error)."""
    sess = StreamSession(model, n_streams=1)
    sess.set_state(trace["s_init"])
    worst = 0.0
    flips = 0
    for t in range(trace["imu"].shape[0]):
        st = sess.step_closed(trace["imu"][t][None])
        if t < 5:
            assert st is None
            continue
        worst = max(worst, np.abs(st[0, :57] - trace["qdq"][t][3:60]).max())
        flips += int((st[0, 57:77:4] != trace["ct"][t][0::4]).sum())
        np.testing.assert_allclose(st[0, 58:77:4], trace["ct"][t][1::4], atol=1e-3)
    assert worst < 1e-3, worst
    assert flips == 0, flips          # contact flags (logit > 0) agree on every frame of this trace


def test_batched_evaluator_equals_single_streams(trace, model):
    """N4: motions of different lengths run as parallel streams give, per motion, the single-stream result."""
    imu = trace["imu"]
    rs = np.random.RandomState(0)
    s2 = trace["s_init"].copy()
    s2[3:57] = rs.uniform(-0.3, 0.3, 54)
    motions = [imu, imu[:90], imu[20:80]]
    inits = [trace["s_init"], s2, trace["s_init"]]
    res = run_motions(model, motions, inits)
    assert [r["state"].shape[0] for r in res] == [150, 90, 60]
    for i, (mo, s0) in enumerate(zip(motions, inits)):
        single = run_motions(model, [mo], [s0])[0]
        np.testing.assert_allclose(res[i]["state"], single["state"], atol=2e-5)
        np.testing.assert_allclose(res[i]["ct"][:, 1::4], single["ct"][:, 1::4], atol=2e-5)
        assert not res[i]["valid"][:5].any() and res[i]["valid"][5:].all()
        np.testing.assert_array_equal(res[i]["state"][0], np.asarray(s0)[3:60])
    assert np.abs(res[0]["state"][5:] - trace["qdq"][5:, 3:60]).max() < 1e-3


def test_closed_loop_device_tensors(trace, model):
    """CUDA tensors in / out (no host round trip) give the same states as the host-buffer call."""
    a = StreamSession(model, n_streams=2)
    a.set_state(np.stack([trace["s_init"], trace["s_init"]]))
    outs = []
    for t in range(60):
        raw = torch.from_numpy(np.stack([trace["imu"][t], trace["imu"][t + 1]])).cuda()
        st = a.step_closed(raw)
        if st is not None:
            assert st.is_cuda and st.dtype == torch.float64
            outs.append(st.cpu().numpy())
    b = StreamSession(model, n_streams=2)
    b.set_state(np.stack([trace["s_init"], trace["s_init"]]))
    k = 0
    for t in range(60):
        st = b.step_closed(np.stack([trace["imu"][t], trace["imu"][t + 1]]))
        if st is not None:
            np.testing.assert_allclose(st, outs[k], atol=1e-12)
            k += 1
    assert k == len(outs) == 55


def test_closed_loop_requires_state(model):
    sess = StreamSession(model, n_streams=1)
    with pytest.raises(RuntimeError, match="tip_stream_set_state"):
        sess.step_closed(np.zeros((1, 72), dtype=np.float32))
