"""GPU: the reference's UNMODIFIED consumers driving the drop-in on the B200.

``from simple_transformer_with_state import TF_RNN_Past_State`` (offline_testing_simple.py:80, live_demo_new.py:16)
resolves to the B200-native class; ``RTRunnerMin`` (real_time_runner_minimal.py), ``RTRunner`` (real_time_runner.py,
the live demo's runner) and the evaluation script ``offline_testing_simple.py`` run from the staged reference
install (baseline/_ref/reference) without edits.  The comparison model is the reference module itself, run by
torch on the same GPU (or the committed trace minted from it on the CPU)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "ref_env"))
import consumers as CE  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(CE.reference_dir() is None or not os.path.exists(os.path.join(CE.CKPT_DIR, "model-with-dip9and10.pt")),
                                 reason="reference install not staged")]


def _run_min(dropin, g, T):
    with CE.consumer_env(dropin=dropin, deterministic=True) as stws:
        from real_time_runner_minimal import RTRunnerMin
        m = CE.build_model(stws)
        r = RTRunnerMin(CE.make_char(), m, 40, g["s_init"], with_acc_sum=True)
        prev = g["s_init"][:3].copy()
        qdq, ct = [], []
        for t in range(T):
            with torch.no_grad():
                res = r.step(g["imu"][t].astype(np.float64), prev)
            prev = res["qdq"][:3].copy()
            qdq.append(np.array(res["qdq"]))
            ct.append(np.array(res["ct"]))
        return np.array(qdq), np.array(ct), type(m).__module__


def test_unmodified_rtrunnermin_through_the_dropin_matches_the_reference_trace():
    """150 frames of RTRunnerMin.step (window growth 1..40, then sliding) with the drop-in as ``self.model``
    (real_time_runner_minimal.py:149: ``self.model(x_imu.cuda(), x_s.cuda()).cpu()``), closed loop, against the
    trace the same runner produced with the reference model (tests/golden/runner_min_trace.npz)."""
    g = np.load(os.path.join(GOLD, "runner_min_trace.npz"))
    T = g["imu"].shape[0]
    qdq, ct, mod = _run_min(True, g, T)
    assert mod.startswith("tip_b200")
    assert np.abs(qdq[:, 3:60] - g["qdq"][:, 3:60]).max() < 1e-3          # pose + root velocity (model-visible loop)
    np.testing.assert_array_equal(ct[:, 0::4], g["ct"][:, 0::4])            # contact flags identical
    assert np.abs(ct - g["ct"]).max() < 1e-3
    assert np.abs(qdq[:, :3] - g["qdq"][:, :3]).max() < 5e-3                # root translation (integrated, FK-corrected)
    # and the reference model on THIS GPU (torch eager) gives the same trace: the comparison is apples to apples
    qdq_r, ct_r, mod_r = _run_min(False, g, 60)
    assert mod_r == "simple_transformer_with_state"
    assert np.abs(qdq_r[:, 3:60] - g["qdq"][:60, 3:60]).max() < 1e-3


def test_unmodified_live_demo_runner_through_the_dropin():
    """``RTRunner`` (real_time_runner.py; what live_demo_new.py:256-285 drives) with the drop-in vs the reference
    model on the same GPU: terrain / IK history corrections stay on the CPU, the model call is ours."""
    g = np.load(os.path.join(GOLD, "runner_min_trace.npz"))

    def run(dropin):
        with CE.consumer_env(dropin=dropin, deterministic=True) as stws:
            from real_time_runner import RTRunner
            import constants as cst
            m = CE.build_model(stws)
            r = RTRunner(CE.make_char(), m, 40, g["s_init"], map_bound=cst.MAP_BOUND, grid_size=cst.GRID_SIZE,
                         play_back_gt=False, five_sbp=True, with_acc_sum=True, multi_sbp_terrain_and_correction=False)
            prev = g["s_init"][:3].copy()
            out = []
            for t in range(100):
                with torch.no_grad():
                    res = r.step(g["imu"][t].astype(np.float64), prev, t=t)
                prev = res["qdq"][:3].copy()
                out.append(np.concatenate((res["qdq"], res["ct"])))
            return np.array(out)
    a, b = run(True), run(False)
    assert np.isfinite(a).all()
    assert np.abs(a[:, 3:60] - b[:, 3:60]).max() < 2e-3
    assert np.abs(a[:, 114:] - b[:, 114:]).max() < 2e-3


def test_unmodified_offline_testing_simple_through_the_dropin(tmp_path):
    """BASELINE configs[3]: ``offline_testing_simple.py`` on DIP-format motions (synthetic: the recordings are
    absent) with model-with-dip9and10.pt -- once with the drop-in, once with the reference model, same motions.
    Deterministic mode: the two predicted trajectories agree (pose error between them, the reference's own
    metrics); as shipped (train mode, p = 0.8: the script's default): both land on the same accuracy."""
    wd = str(tmp_path)
    with CE.consumer_env(dropin=False, workdir=wd):
        CE.write_synthetic_dip(wd, n_motions=2, T=260)
    ours = CE.run_offline_testing_simple(dropin=True, workdir=wd, deterministic=True)
    ref = CE.run_offline_testing_simple(dropin=False, workdir=wd, deterministic=True)
    with CE.consumer_env(dropin=False, workdir=wd):
        char = CE.make_char()
        for a, b in zip(ours["ours_list"], ref["ours_list"]):
            e = CE.pose_error_between(char, a, b)
            assert e["mpjpe_cm"] < 0.05 and e["joint_angle_deg"] < 0.05, e
    for k, v in ref["metrics"].items():
        assert abs(ours["metrics"][k] - v) <= 2e-2 * max(1.0, abs(v)), (k, ours["metrics"][k], v)
    torch.manual_seed(0)
    ours_s = CE.run_offline_testing_simple(dropin=True, workdir=wd, deterministic=False)
    ref_s = CE.run_offline_testing_simple(dropin=False, workdir=wd, deterministic=False)
    for r in (ours_s, ref_s):
        assert all(np.isfinite(v) for v in r["metrics"].values())
    # stochastic on both sides (different generators): same accuracy band
    assert abs(ours_s["metrics"]["joint_angle_err_deg"] - ref_s["metrics"]["joint_angle_err_deg"]) < 0.25 * ref_s["metrics"]["joint_angle_err_deg"]
    assert abs(ours_s["metrics"]["joint_pos_err_cm"] - ref_s["metrics"]["joint_pos_err_cm"]) < 0.25 * ref_s["metrics"]["joint_pos_err_cm"]


def test_long_closed_loop_600_frames_no_drift(tmp_path):
    """Closed loop far beyond one window (600 frames = 10 s at 60 Hz, the script's --test_len): the drop-in and the
    reference model (torch eager, same GPU) drive the unmodified RTRunnerMin on the same synthetic motion; the fed-back
    state must not drift apart (the loop is contractive: the root rotation comes from the IMU every frame)."""
    import pickle
    wd = str(tmp_path)
    with CE.consumer_env(dropin=False, workdir=wd):
        path = CE.write_synthetic_dip(wd, n_motions=1, T=600, seed=21)[0]
    motion = pickle.load(open(path, "rb"))

    def run(dropin):
        with CE.consumer_env(dropin=dropin, deterministic=True, workdir=wd) as stws:
            from real_time_runner_minimal import RTRunnerMin
            m = CE.build_model(stws)
            r = RTRunnerMin(CE.make_char(), m, 40, motion["nimble_qdq"][0], with_acc_sum=True)
            prev = motion["nimble_qdq"][0][:3].copy()
            out = []
            for t in range(600):
                with torch.no_grad():
                    res = r.step(motion["imu"][t], prev)
                prev = res["qdq"][:3].copy()
                out.append(np.concatenate((res["qdq"], res["ct"])))
            return np.array(out)
    a, b = run(True), run(False)
    err = np.abs(a[:, 3:60] - b[:, 3:60]).max(axis=1)
    assert err.max() < 2e-3, (err.max(), int(err.argmax()))
    assert err[400:].max() < 2 * max(err[:200].max(), 2e-4)                 # no growth over the run
    flips = int((a[:, 114::4] != b[:, 114::4]).sum())
    assert flips <= 2, flips                                                # contact flags (thresholded logits)
