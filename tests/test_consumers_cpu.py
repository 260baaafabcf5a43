"""CPU: the environment that runs the reference's UNMODIFIED consumers (tools/ref_env/consumers.py) -- here
with the reference model on the CPU; the GPU tests (test_gpu_consumers.py) run the same consumers against the
drop-in.  Skipped when the reference install is not staged (baseline/_ref/reference, made by build())."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "ref_env"))
import consumers as CE  # noqa: E402

pytestmark = pytest.mark.skipif(CE.reference_dir() is None or not os.path.exists(os.path.join(CE.CKPT_DIR, "model-with-dip9and10.pt")),
                                reason="reference install not staged")


def test_env_resolves_the_model_module_by_name():
    with CE.consumer_env(dropin=False) as stws:
        assert os.path.realpath(stws.__file__).startswith(os.path.realpath(CE.reference_dir()))
    with CE.consumer_env(dropin=True) as stws:
        assert os.path.realpath(stws.__file__).startswith(os.path.realpath(CE.PKG))
        from tip_b200 import TF_RNN_Past_State
        assert stws.TF_RNN_Past_State is TF_RNN_Past_State
    assert "simple_transformer_with_state" not in sys.modules and os.getcwd() != CE.reference_dir()


def test_unmodified_runner_with_reference_model_reproduces_the_golden_trace():
    """The staged install + shims + deterministic wrapper give back the committed trace bit for bit."""
    g = np.load(os.path.join(GOLD, "runner_min_trace.npz"))
    torch.set_num_threads(1)
    with CE.consumer_env(dropin=False, deterministic=True) as stws:
        from real_time_runner_minimal import RTRunnerMin
        m = CE.build_model(stws)
        assert not m.training and m.past_state_dropout == 0.0
        r = RTRunnerMin(CE.make_char(), m, 40, g["s_init"], with_acc_sum=True)
        prev = g["s_init"][:3].copy()
        for t in range(30):
            with torch.no_grad():
                res = r.step(g["imu"][t].astype(np.float64), prev)
            prev = res["qdq"][:3].copy()
            np.testing.assert_allclose(res["qdq"], g["qdq"][t], atol=1e-12)
            np.testing.assert_allclose(res["ct"], g["ct"][t], atol=1e-12)


def test_synthetic_dip_files_and_unmodified_offline_script(tmp_path):
    """Synthetic motions in the DIP pkl layout (preprocess_DIP_TC_new.py:211) and the unmodified evaluation
    script end to end with the reference model (as shipped: train mode, p = 0.8)."""
    wd = str(tmp_path)
    with CE.consumer_env(dropin=False, workdir=wd):
        paths = CE.write_synthetic_dip(wd, n_motions=2, T=170)
    for p in paths:
        d = pickle.load(open(p, "rb"))
        assert set(d) == {"imu", "nimble_qdq"} and d["imu"].shape == (170, 72) and d["nimble_qdq"].shape == (170, 114)
        R = d["imu"][:, :54].reshape(-1, 3, 3)
        np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-9)
        assert np.abs(d["imu"][:, 54:]).max() < 60 and np.isfinite(d["imu"]).all()
    torch.set_num_threads(1)
    res = CE.run_offline_testing_simple(dropin=False, workdir=wd, deterministic=False)
    assert len(res["ours_list"]) == 2 and res["ours_list"][0].shape == (170, 114)
    assert set(res["metrics"]) == {"joint_angle_err_deg", "joint_pos_err_cm", "root_drift_2s_m", "root_drift_5s_m",
                                   "root_drift_10s_m", "jerk_all", "jerk_root"}
    assert all(np.isfinite(v) for v in res["metrics"].values())
    assert 0 < res["metrics"]["joint_angle_err_deg"] < 60
