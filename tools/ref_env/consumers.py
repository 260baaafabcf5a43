"""Run the reference's UNMODIFIED consumers of the hot path -- ``RTRunnerMin`` / ``RTRunner``
(real_time_runner_minimal.py, real_time_runner.py) and the evaluation script
``offline_testing_simple.py`` -- against either model:

* the drop-in (``transformer-inertial-poser_b200/simple_transformer_with_state.py`` first on ``sys.path``, so
  ``from simple_transformer_with_state import TF_RNN_Past_State`` -- offline_testing_simple.py:80 -- resolves
  to the B200-native class), or
* the reference module itself (from the staged reference install).

The reference sources are read from the staged install ``baseline/_ref/reference`` (git-ignored; copied there by
``__graft_entry__.build()`` in the build container, shipped to the GPU box by gpurun), else from
``/root/reference``.  Nothing is patched inside those files.  What the environment supplies around them:

* ``tools/ref_env/shims``: stand-ins for the third-party packages this image lacks (``fairmotion`` subset,
  kinematic ``pybullet``, ``imageio``);
* a scratch working directory laid out like the reference checkout (``data/amass.urdf``, ``amass_char_info.py``,
  the ``data/<dataset>`` folders the script scans), with SYNTHETIC motions in the DIP pkl format
  (``preprocess_DIP_TC_new.py:211``: ``{"imu": (T, 72), "nimble_qdq": (T, 114)}``) -- the real DIP-IMU
  recordings are not redistributable and absent;
* optionally the deterministic parity mode of SURVEY.md 8c (``eval()``, ``past_state_dropout = 0``), applied by
  wrapping the constructor of whichever class the consumer imports -- the consumers themselves leave the module
  in train mode with p = 0.8 (offline_testing_simple.py:93,98), which is the "as shipped" mode;
* on a box without a GPU (the build container) ``Tensor.cuda`` / ``Module.cuda`` become the identity and
  ``torch.load`` maps to the CPU, so that the REFERENCE model runs on the CPU; the drop-in has no CPU path.

Test / bench infrastructure: the product never imports this file.
"""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import runpy
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "transformer-inertial-poser_b200")
SHIMS = os.path.join(ROOT, "tools", "ref_env", "shims")
STAGED = os.path.join(ROOT, "baseline", "_ref", "reference")
CKPT_DIR = os.path.join(ROOT, "baseline", "_ref")

# reference files the consumers import (copied verbatim by build(); never edited)
REFERENCE_FILES = [
    "simple_transformer_with_state.py", "real_time_runner_minimal.py", "real_time_runner.py", "data_utils.py",
    "constants.py", "amass_char_info.py", "bullet_agent.py", "bullet_utils.py", "bullet_client.py",
    "render_funcs.py", "learning_utils.py", "offline_testing_simple.py", "data/amass.urdf", "LICENSE",
]
# modules that must be re-imported when the environment switches between the drop-in and the reference model
_VOLATILE = ("simple_transformer_with_state", "real_time_runner_minimal", "real_time_runner", "data_utils", "constants",
             "bullet_agent", "bullet_utils", "bullet_client", "render_funcs", "learning_utils", "pybullet", "imageio",
             "char_info")
# the dataset folders offline_testing_simple.py:310-318 scans under data/
DATASET_DIRS = ["syn_AMASS_CMU_v0", "syn_Eyes_Japan_Dataset_v0", "syn_KIT_v0", "syn_HUMAN4D_v0", "syn_ACCAD_v0",
                "syn_DFaust_67_v0", "syn_HumanEva_v0", "syn_MPI_Limits_v0", "syn_MPI_mosh_v0", "syn_SFU_v0",
                "syn_Transitions_mocap_v0", "preprocessed_DIP_IMU_v0", "preprocessed_TotalCapture_v0",
                "syn_TotalCapture_v0", "syn_DanceDB_v0"]


def reference_dir():
    """Directory holding the unmodified reference sources, or None."""
    for d in (STAGED, "/root/reference"):
        if os.path.exists(os.path.join(d, "real_time_runner_minimal.py")) and os.path.exists(os.path.join(d, "data", "amass.urdf")):
            return d
    return None


def stage_reference(src="/root/reference", dst=STAGED):
    """Copy the reference files the consumers need into the git-ignored install (build container only)."""
    if not os.path.isdir(src):
        return None
    for f in REFERENCE_FILES:
        s, d = os.path.join(src, f), os.path.join(dst, f)
        if os.path.exists(s) and not (os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s)):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    return dst


def _purge():
    for k in list(sys.modules):
        if k in _VOLATILE or k == "fairmotion" or k.startswith("fairmotion."):
            del sys.modules[k]


@contextlib.contextmanager
def consumer_env(dropin: bool, deterministic: bool = False, workdir: str | None = None, cpu_reference: bool | None = None):
    """Import environment + working directory in which the reference consumers run unmodified.

    dropin        : ``simple_transformer_with_state`` resolves to the B200 drop-in (True) or the reference module.
    deterministic : force ``eval()`` + ``past_state_dropout = 0`` on every model the consumer constructs.
    workdir       : directory to run in (default: the reference directory itself, for ``data/amass.urdf``).
    cpu_reference : make ``.cuda()`` the identity (default: only when no GPU is present and dropin is False).
    """
    import torch
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference sources are not staged (baseline/_ref/reference); run __graft_entry__.build() in the build container")
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    saved_cuda = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.load)
    _purge()
    model_dir = PKG if dropin else ref
    sys.path[:] = [model_dir, SHIMS, ref] + [p for p in saved_path if p not in (model_dir, SHIMS, ref, PKG)] + ([PKG] if not dropin else [])
    if not dropin:
        # (tip_b200 stays importable as a package, but the bare module name must resolve to the reference file)
        sys.path[:] = [ref, SHIMS] + [p for p in saved_path if p not in (ref, SHIMS)]
    if cpu_reference is None:
        cpu_reference = (not dropin) and not torch.cuda.is_available()
    try:
        if cpu_reference:
            torch.Tensor.cuda = lambda self, *a, **k: self
            torch.nn.Module.cuda = lambda self, *a, **k: self
            _load = saved_cuda[2]
            # (the released checkpoints hold CUDA tensors; offline_testing_simple.py:96 loads them without map_location)
            torch.load = lambda f, *a, **k: _load(f, *a, **{**k, "map_location": k.get("map_location", "cpu")})
        os.chdir(workdir or ref)
        import simple_transformer_with_state as stws
        expect = os.path.realpath(os.path.join(model_dir, "simple_transformer_with_state.py"))
        assert os.path.realpath(stws.__file__) == expect, (stws.__file__, expect)
        if deterministic:
            cls = stws.TF_RNN_Past_State
            orig_init = cls.__init__

            def init(self, *a, **k):
                orig_init(self, *a, **k)
                self.eval()
                self.past_state_dropout = 0.0
            cls.__init__ = init
        try:
            yield stws
        finally:
            if deterministic:
                cls.__init__ = orig_init
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.load = saved_cuda
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        _purge()


def load_char_info(ref):
    import importlib.util
    spec = importlib.util.spec_from_file_location("char_info", os.path.join(ref, "amass_char_info.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_char():
    """The kinematic character the runners use for FK (render_funcs.py:170-178 without the GUI)."""
    import bullet_client
    import pybullet as pb
    from bullet_agent import SimAgent
    pb_c = bullet_client.BulletClient(connection_mode=pb.DIRECT)
    with contextlib.redirect_stdout(io.StringIO()):
        char = SimAgent(name="sim_agent_0", pybullet_client=pb_c, model_file="data/amass.urdf",
                        char_info=load_char_info(os.getcwd()), ref_scale=1.0, self_collision=False,
                        kinematic_only=True, verbose=True)
    return char


def build_model(stws, ckpt="model-with-dip9and10.pt", device=None):
    """The model exactly as offline_testing_simple.py:81-99 builds it (NOT put into eval mode)."""
    import torch
    with contextlib.redirect_stdout(io.StringIO()):
        m = stws.TF_RNN_Past_State(72, 18 * 6 + 3 + 20, rnn_hid_size=512, tf_hid_size=1024, tf_in_dim=256,
                                   n_heads=16, tf_layers=4, dropout=0.0, in_dropout=0.0,
                                   past_state_dropout=0.8, with_acc_sum=True)
    m.load_state_dict(torch.load(os.path.join(CKPT_DIR, ckpt), map_location="cpu"))
    return m.cuda() if device is None else m.to(device)


# ---- synthetic motions in the DIP pkl format ------------------------------------------------------------------
def synth_qdq(seed: int, T: int):
    """(T, 114) smooth random motion in the reference's state layout (``nimble_qdq``): q[0:3] root position,
    q[3:6] root axis-angle, q[6:57] 17 joints' axis-angles (Nimble order, amass_char_info.py:89-109), dq = 0 except
    the root velocity dq[0:3] (what ``get_raw_motion_info_nimble_q_dummy_dq`` stores, preprocess_DIP_TC_new.py:206)."""
    rs = np.random.RandomState(seed)
    t = np.arange(T) / 60.0
    q = np.zeros((T, 57))
    for j in range(3, 57):                                   # a few low-frequency sinusoids per DoF
        amp = 0.35 if j >= 6 else 0.25
        for _ in range(3):
            f = rs.uniform(0.1, 0.9)
            q[:, j] += amp / 3 * rs.uniform(0.3, 1.0) * np.sin(2 * np.pi * f * t + rs.uniform(0, 2 * np.pi))
    q[:, 3:6] += np.array([1.2, 1.2, 1.2])                   # upright (z-up world: constants.rot_up_Q as an axis-angle)
    walk = np.cumsum(0.6 / 60.0 * np.stack([np.cos(0.2 * t), np.sin(0.2 * t)], axis=1), axis=0)
    q[:, 0:2] = walk
    q[:, 2] = 0.95 + 0.02 * np.sin(2 * np.pi * 1.7 * t)
    dq = np.zeros((T, 57))
    dq[1:, 0:3] = (q[1:, 0:3] - q[:-1, 0:3]) * 60.0
    dq[0, 0:3] = dq[1, 0:3]
    return np.concatenate((q, dq), axis=1)


def synth_imu_from_motion(char, qdq):
    """(T, 72) IMU readings of a motion, as the reference synthesises them (data-gen-and-viz-bullet-new.py:148-213):
    global orientations of [root, lwrist, rwrist, lknee, rknee, upperneck] (row-major 3x3 each) and second finite
    differences of their positions over +-acc_fd_N frames; runs the reference's own FK wrapper over the character."""
    import constants as cst
    from data_utils import our_pose_2_bullet_format, viz_current_frame_and_store_fk_info_include_fixed
    from fairmotion.ops import conversions
    info = char.get_char_info()
    bodies = [0] + [1 + j for j in (info.lwrist, info.rwrist, info.lknee, info.rknee, info.upperneck)]
    T = qdq.shape[0]
    pq = np.array([viz_current_frame_and_store_fk_info_include_fixed(char, our_pose_2_bullet_format(char, qdq[t]))
                   for t in range(T)])                       # (T, 20, 7): position + xyzw quaternion per body
    H = np.zeros((T, 72))
    H[:, :54] = conversions.Q2R(pq[:, bodies, 3:]).reshape(T, 54)
    n = cst.acc_fd_N
    p = pq[:, bodies, :3]
    acc = (-2 * p[n:-n] + p[2 * n:] + p[:-2 * n]) / (cst.DT_FIN_ACC ** 2)
    H[n:-n, 54:] = acc.reshape(-1, 18)
    H[:n, 54:] = H[n, 54:]
    H[-n:, 54:] = H[-n - 1, 54:]
    return H


def write_synthetic_dip(workdir, n_motions=2, T=400, seed=9):
    """Lay out ``workdir`` like the reference checkout and write ``n_motions`` synthetic DIP-format pkl files
    (named like the DIP-IMU s_09 / s_10 test split, preprocess_DIP_TC_new.py:317-338).  Must run inside
    ``consumer_env`` (uses the reference's FK).  Returns the pkl paths."""
    ref = reference_dir()
    os.makedirs(os.path.join(workdir, "data"), exist_ok=True)
    for f in ("data/amass.urdf", "amass_char_info.py"):
        if not os.path.exists(os.path.join(workdir, f)):
            shutil.copyfile(os.path.join(ref, f), os.path.join(workdir, f))
    for d in DATASET_DIRS:
        os.makedirs(os.path.join(workdir, "data", d), exist_ok=True)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        char = make_char()
        paths = []
        for i in range(n_motions):
            qdq = synth_qdq(seed + i, T)
            imu = synth_imu_from_motion(char, qdq)
            path = os.path.join(workdir, "data", "preprocessed_DIP_IMU_v0", f"dipimu_s_{9 + i % 2:02d}_{i:02d}.pkl")
            with open(path, "wb") as fh:
                pickle.dump({"imu": imu, "nimble_qdq": qdq}, fh, protocol=pickle.HIGHEST_PROTOCOL)
            paths.append(path)
    finally:
        os.chdir(cwd)
    return paths


def run_offline_testing_simple(dropin: bool, workdir: str, deterministic: bool, ckpt="model-with-dip9and10.pt",
                               test_len=600, seed=42):
    """Execute the unmodified ``offline_testing_simple.py`` (as ``python offline_testing_simple.py --name_contains
    "dipimu_s_09 dipimu_s_10" --ours_path_name_kin <ckpt> --with_acc_sum --five_sbp --compare_gt``, README.md:113) in
    ``workdir``.  Returns {"metrics": the 7 mean metrics it prints (:447-453), "ours_list", "gt_list", "stdout"}."""
    ref = reference_dir()
    argv = ["offline_testing_simple.py", "--name_contains", "dipimu_s_09 dipimu_s_10", "--ours_path_name_kin",
            os.path.join(CKPT_DIR, ckpt), "--with_acc_sum", "--five_sbp", "--compare_gt", "--test_len", str(test_len),
            "--seed", str(seed)]
    out = io.StringIO()
    with consumer_env(dropin, deterministic=deterministic, workdir=workdir):
        saved_argv = sys.argv
        sys.argv = argv
        try:
            with contextlib.redirect_stdout(out):
                runpy.run_path(os.path.join(ref, "offline_testing_simple.py"), run_name="__main__")
        finally:
            sys.argv = saved_argv
    text = out.getvalue()
    with open(os.path.join(workdir, "test-output-tmp.pkl"), "rb") as fh:
        dump = pickle.load(fh)
    nums = []
    for line in text.splitlines():
        try:
            nums.append(float(line.strip()))
        except ValueError:
            pass
    names = ["joint_angle_err_deg", "joint_pos_err_cm", "root_drift_2s_m", "root_drift_5s_m", "root_drift_10s_m",
             "jerk_all", "jerk_root"]
    # (the script prints the file count first, then the 7 means, then 7 "max <file>" lines)
    assert len(nums) >= 7, text[-2000:]
    return {"metrics": dict(zip(names, nums[-7:])), "ours_list": dump["ours_list"], "gt_list": dump["gt_list"],
            "stdout": text}


def pose_error_between(char, traj_a, traj_b):
    """Pose error between two predicted trajectories (T, 114) with the reference's own metrics
    (data_utils.py:314-338 ``loss_angle`` / ``loss_j_pos`` over its FK, as offline_testing_simple.py:423-441 uses
    them): mean joint-angle difference in degrees and mean joint-position difference (root-relative) in cm --
    the MPJPE of BASELINE configs[3].  Must run inside ``consumer_env``."""
    from data_utils import (loss_angle, loss_j_pos, our_pose_2_bullet_format,
                            viz_current_frame_and_store_fk_info_include_fixed)
    ta = np.array([our_pose_2_bullet_format(char, s) for s in traj_a])
    tb = np.array([our_pose_2_bullet_format(char, s) for s in traj_b])
    pa = np.array([viz_current_frame_and_store_fk_info_include_fixed(char, s) for s in ta])
    pb = np.array([viz_current_frame_and_store_fk_info_include_fixed(char, s) for s in tb])
    return {"joint_angle_deg": float(loss_angle(ta, tb, pa, pb)), "mpjpe_cm": float(loss_j_pos(ta, tb, pa, pb)),
            "max_abs_qdq": float(np.abs(np.asarray(traj_a)[:, 3:60] - np.asarray(traj_b)[:, 3:60]).max())}


def scratch_dir(prefix="tip_consumer_"):
    return tempfile.mkdtemp(prefix=prefix)
