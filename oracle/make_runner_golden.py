"""Mint a golden streaming trace from the UNMODIFIED reference runner (test infrastructure, rows N1-N3).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_runner_golden.py

What runs unmodified from ``/root/reference``: ``real_time_runner_minimal.RTRunnerMin`` (the per-frame
state machine, :20-200), ``data_utils`` (rotation representations, root-local IMU rotation, SBP root
correction), ``bullet_agent.SimAgent`` / ``bullet_utils`` / ``bullet_client``, ``amass_char_info``,
``constants`` and the model ``simple_transformer_with_state.TF_RNN_Past_State`` with the released
checkpoint.  What is substituted (``tools/ref_env/shims``): ``fairmotion`` (numpy/scipy restatement of the few
conversions used) and ``pybullet`` (kinematic FK over ``data/amass.urdf``) -- neither is installed here.
Two process-level adaptations, both outside the reference sources:
  * ``torch.Tensor.cuda`` is made the identity (the runner hard-codes ``.cuda()`` at :149; this container
    has no GPU), so the reference model runs on the CPU;
  * the model is put into the deterministic mode of SURVEY.md 8c (``eval()``, ``past_state_dropout = 0``).

The trace pins, per runner call: the raw IMU frame fed in, the last row of both model windows, the model's
last output row, and the runner's outputs ``qdq`` / ``ct``.  The model-visible closed loop
(``qdq[3:60]`` and ``ct`` -> next ``x_s`` row) does not depend on the FK stand-in; ``qdq[0:3]`` does.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SHIMS = os.path.join(ROOT, "tools", "ref_env", "shims")
GOLD = os.path.join(ROOT, "tests", "golden")


def random_rotation_walk(rs, T, step=0.06):
    """(T, 3, 3) smooth random-walk rotations."""
    from scipy.spatial.transform import Rotation
    R = Rotation.from_rotvec(rs.uniform(-0.5, 0.5, 3))
    w = rs.standard_normal(3) * step
    out = []
    for _ in range(T):
        w = 0.9 * w + 0.1 * rs.standard_normal(3) * step
        R = Rotation.from_rotvec(w) * R
        out.append(R.as_matrix())
    return np.array(out)


def synth_raw_imu(seed, T):
    """(T, 72) raw runner input: 6 global rotations (54, row-major) + 6 global accelerations (18)
    (real_time_runner_minimal.py:118), rounded to float32-representable values."""
    rs = np.random.RandomState(seed)
    rots = np.stack([random_rotation_walk(rs, T) for _ in range(6)], axis=1)        # (T, 6, 3, 3)
    acc = np.zeros((T, 18))
    a = np.zeros(18)
    for t in range(T):
        a = 0.8 * a + 0.2 * rs.standard_normal(18) * 6.0
        acc[t] = a
    imu = np.concatenate((rots.reshape(T, 54), acc), axis=1)
    return imu.astype(np.float32).astype(np.float64)


def main(T=150, seed=3, ckpt="model-with-dip9and10.pt"):
    sys.path.insert(0, REF)
    sys.path.insert(0, SHIMS)
    os.chdir(REF)                                   # SimAgent loads "data/amass.urdf" relative to the repo
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU in this container (see module docstring)
    torch.manual_seed(0)
    torch.set_num_threads(1)

    import importlib.util
    import bullet_client
    import pybullet as pb
    from bullet_agent import SimAgent
    from real_time_runner_minimal import RTRunnerMin
    from simple_transformer_with_state import TF_RNN_Past_State
    import constants as cst

    spec = importlib.util.spec_from_file_location("char_info", os.path.join(REF, "amass_char_info.py"))
    char_info = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(char_info)

    pb_c = bullet_client.BulletClient(connection_mode=pb.DIRECT)
    with contextlib.redirect_stdout(io.StringIO()):
        char = SimAgent(name="sim_agent_0", pybullet_client=pb_c, model_file="data/amass.urdf",
                        char_info=char_info, ref_scale=1.0, self_collision=False, kinematic_only=True,
                        verbose=True)
        m = TF_RNN_Past_State(72, 131, rnn_hid_size=512, tf_hid_size=1024, tf_in_dim=256, n_heads=16,
                              tf_layers=4, dropout=0.0, in_dropout=0.0, past_state_dropout=0.8,
                              with_acc_sum=True)
    m.load_state_dict(torch.load(os.path.join(REF, "output", ckpt), map_location="cpu"))
    m.eval()
    m.past_state_dropout = 0.0

    calls = {"x_imu": [], "x_s": [], "y": [], "L": []}

    class Recorder(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x_imu, x_s):
            with torch.no_grad():
                y = self.inner(x_imu, x_s)
            calls["x_imu"].append(x_imu[0, -1].numpy().copy())
            calls["x_s"].append(x_s[0, -1].numpy().copy())
            calls["y"].append(y[0, -1].numpy().copy())
            calls["L"].append(x_imu.shape[1])
            return y

    rs = np.random.RandomState(seed + 1000)
    s_init = np.zeros(cst.n_dofs * 2)
    s_init[:3] = [0.0, 0.0, 0.95]
    s_init[3:cst.n_dofs] = rs.uniform(-0.4, 0.4, cst.n_dofs - 3)
    imu = synth_raw_imu(seed, T)

    runner = RTRunnerMin(char, Recorder(m), 40, s_init, with_acc_sum=True)
    qdq, ct, viz = [], [], []
    prev_xyz = s_init[:3].copy()
    for t in range(T):
        res = runner.step(imu[t], prev_xyz)
        qdq.append(np.array(res["qdq"], dtype=np.float64))
        ct.append(np.array(res["ct"], dtype=np.float64))
        viz.append(np.array(res["viz_locs"], dtype=np.float64))
        prev_xyz = qdq[-1][:3].copy()

    out = os.path.join(GOLD, "runner_min_trace.npz")
    np.savez_compressed(
        out, imu=imu.astype(np.float32), s_init=s_init, qdq=np.array(qdq), ct=np.array(ct),
        viz_locs=np.array(viz), x_imu_last=np.array(calls["x_imu"]), x_s_last=np.array(calls["x_s"]),
        y_last=np.array(calls["y"]), L=np.array(calls["L"], dtype=np.int32),
        s_and_c_in=np.array(runner.s_and_c_in_buffer), ckpt=np.array(ckpt))
    print("wrote", out, "calls:", len(calls["L"]), "L max", max(calls["L"]),
          "|y| max %.3f" % np.abs(np.array(calls["y"])).max(), "root xyz end", qdq[-1][:3])


if __name__ == "__main__":
    main()
