"""CPU: pin the runner-side restatements (rows N1 window assembly, N3 post-model step) against a trace
minted from the UNMODIFIED reference runner ``RTRunnerMin`` (oracle/make_runner_golden.py), and check
the import shims (row N2) that made that run possible."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLD, ROOT
from oracle import tip_oracle as O

TRACE = os.path.join(GOLD, "runner_min_trace.npz")
REF = "/root/reference"


@pytest.fixture(scope="module")
def trace():
    return np.load(TRACE)


def test_window_assembler_matches_reference_runner(trace):
    """N1: record_raw_imu + imu_rotate_to_local + acc-sum rows equal the windows the reference runner fed
    to the model, bit for bit after the runner's ``.float()`` cast (real_time_runner_minimal.py:146)."""
    wa = O.WindowAssembler()
    k = 0
    for t in range(trace["imu"].shape[0]):
        win = wa.push(trace["imu"][t].astype(np.float64))
        if t < 5:
            assert win is None                       # :125-128 the first 5 calls return s_init
            continue
        assert win.shape == (trace["L"][k], 90)
        np.testing.assert_array_equal(win[-1].astype(np.float32), trace["x_imu_last"][k])
        k += 1
    assert k == trace["L"].shape[0] and trace["L"].max() == 40 and trace["L"][0] == 1


def test_post_processor_matches_reference_runner(trace):
    """N3: smoothing filter, SBP split, 2-axis -> axis-angle, averaging with the previous state and the row
    fed back to the model equal the reference runner's to 1e-12 (teacher-forced on the recorded y)."""
    pp = O.PostProcessor()
    row0 = O.state_to_row(trace["s_init"], np.zeros(20))
    np.testing.assert_allclose(row0, trace["s_and_c_in"][0], atol=1e-12)
    for i in range(trace["L"].shape[0]):
        s, c, row = pp.step(trace["y_last"][i], trace["x_imu_last"][i][:9])
        t = i + 5                                    # runner call index of model call i
        np.testing.assert_allclose(s, trace["qdq"][t][3:60], atol=1e-12)
        np.testing.assert_allclose(c, trace["ct"][t], atol=0)
        np.testing.assert_allclose(row, trace["s_and_c_in"][i + 1], atol=1e-12)
        if i + 1 < trace["L"].shape[0]:              # what the next model call saw as its newest x_s row
            np.testing.assert_array_equal(row.astype(np.float32), trace["x_s_last"][i + 1])


def test_first_calls_edit_the_smoothing_buffer_in_place(trace):
    """The reference thresholds / rescales the SBP block of the buffered raw rows during the first five
    model calls (:98 returns a view, :107-110 write through it); the restatement must keep that."""
    pp = O.PostProcessor()
    for i in range(5):
        pp.step(trace["y_last"][i], trace["x_imu_last"][i][:9])
        assert set(np.unique(pp.buf[i][111::4])) <= {0.0, 1.0}
    pp.step(trace["y_last"][5], trace["x_imu_last"][5][:9])
    np.testing.assert_array_equal(pp.buf[5], trace["y_last"][5])     # from the 6th call on the rows stay raw


def test_rotation_conversions_match_scipy():
    from scipy.spatial.transform import Rotation
    rs = np.random.RandomState(0)
    A = rs.uniform(-3, 3, (300, 3))
    A[:10] *= 1e-5
    R = Rotation.from_rotvec(A).as_matrix()
    np.testing.assert_allclose(O.aa_to_rotmat(A), R, atol=1e-14)
    np.testing.assert_allclose(O.rotmat_to_aa(R), Rotation.from_matrix(R).as_rotvec(), atol=1e-12)
    M = R + 0.05 * rs.standard_normal(R.shape)      # the model's 2-axis output is not orthonormal
    np.testing.assert_allclose(O.rotmat_to_aa(M), Rotation.from_matrix(M).as_rotvec(), atol=1e-12)


def test_pybullet_shim_forward_kinematics():
    """N2: the kinematic PyBullet stand-in -- link frames compose parent * origin * joint rotation."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree (URDF) not present on this box")
    sys.path.insert(0, os.path.join(ROOT, "tools", "ref_env", "shims"))
    try:
        import importlib
        pb = importlib.import_module("pybullet")
        from scipy.spatial.transform import Rotation
        bid = pb.loadURDF(os.path.join(REF, "data", "amass.urdf"), flags=pb.URDF_MAINTAIN_LINK_ORDER)
        assert pb.getNumJoints(bid) == 19
        types = [pb.getJointInfo(bid, j)[2] for j in range(19)]
        assert types.count(pb.JOINT_FIXED) == 2 and types[14] == pb.JOINT_FIXED and types[18] == pb.JOINT_FIXED
        names = [pb.getJointInfo(bid, j)[12].decode() for j in range(19)]
        assert names[:6] == ["lhip", "lknee", "lankle", "rhip", "rknee", "rankle"]
        # rest pose: knee joint frame = hip frame + the knee joint origin (pure translations at rest)
        pb.resetBasePositionAndOrientation(bid, [0, 0, 1.0], [0, 0, 0, 1])
        ls = pb.getLinkStates(bid, list(range(19)))
        hip, knee = np.array(ls[0][4]), np.array(ls[1][4])
        q = Rotation.from_euler("z", 90, degrees=True).as_quat()
        pb.resetJointStatesMultiDof(bid, [0], [q], [np.zeros(3)])
        ls2 = pb.getLinkStates(bid, list(range(19)))
        np.testing.assert_allclose(np.array(ls2[0][4]), hip, atol=1e-12)          # the hip does not move
        off = knee - hip
        np.testing.assert_allclose(np.array(ls2[1][4]) - hip, Rotation.from_quat(q).apply(off), atol=1e-12)
        np.testing.assert_allclose(np.linalg.norm(np.array(ls2[1][4]) - hip), np.linalg.norm(off), atol=1e-12)
    finally:
        sys.path.remove(os.path.join(ROOT, "tools", "ref_env", "shims"))
        sys.modules.pop("pybullet", None)


def test_product_state_to_row_matches_reference_runner(trace):
    """Host logic of the product (tip_b200.streaming.state_to_row, used by StreamSession.set_state) against the row
    the reference runner's constructor appended (record_state_aa_and_c(s_init, zeros), :47) and against later rows."""
    from tip_b200.streaming import state_to_row
    np.testing.assert_allclose(state_to_row(trace["s_init"], np.zeros(20)), trace["s_and_c_in"][0], atol=1e-12)
    for t in (6, 40, 149):                         # rows the runner appended after frames t (closed loop)
        i = t - 5                                  # model call index of runner call t
        np.testing.assert_allclose(state_to_row(trace["qdq"][t], trace["ct"][t]), trace["s_and_c_in"][i + 1], atol=1e-12)


def test_evaluator_bookkeeping_without_a_gpu(monkeypatch, trace):
    """run_motions (row N4): stream/time bookkeeping with a fake session -- motions of different lengths, warm-up rows
    hold s_init, finished motions are dropped, per-motion outputs come from their own stream."""
    import tip_b200.evaluate as ev

    class FakeSession:
        def __init__(self, model, n_streams):
            self.S, self.t, self.state_width = n_streams, 0, 80
        def set_state(self, s):
            assert s.shape == (self.S, 114)
        def step_closed(self, frame, y_override=None):
            self.t += 1
            if self.t <= 5:
                return None
            out = np.zeros((self.S, 80))
            out[:, 0] = frame[:, 0] * 10 + np.arange(self.S)          # depends on the stream's own frame only
            out[:, 57] = 1.0
            out[:, 77] = self.t
            return out

    monkeypatch.setattr(ev, "StreamSession", FakeSession)
    imu = trace["imu"]
    s0 = trace["s_init"]
    res = ev.run_motions(None, [imu[:30], imu[:12], imu[:8]], [s0, s0, s0])
    assert [r["state"].shape for r in res] == [(30, 57), (12, 57), (8, 57)]
    for i, r in enumerate(res):
        assert not r["valid"][:5].any() and r["valid"][5:].all()
        np.testing.assert_array_equal(r["state"][:5], np.tile(s0[3:60], (5, 1)))
        np.testing.assert_allclose(r["state"][5:, 0], imu[5:len(r["state"]), 0].astype(np.float32) * 10 + i, rtol=1e-6)
        assert (r["ct"][5:, 0] == 1.0).all() and (r["ct"][:5] == 0).all()
        np.testing.assert_array_equal(r["root_v"][5:, 0], np.arange(6, len(r["state"]) + 1))
