"""Device-resident sliding windows for the per-frame call of the runners.

The reference re-assembles the last <=40 rows of two append-only Python lists, converts them to
tensors and copies both to the GPU on EVERY frame (real_time_runner_minimal.py:131-149,
real_time_runner.py:413-431).  Rows are immutable once written, so a
device-side window that is shifted by one row per frame reproduces the same model inputs;
``StreamSession.step`` pushes one (imu_row, s_row) per stream and returns ``y[:, L-1, :]``
(what :150 consumes).  In steady state (L == 40) the whole frame is one CUDA-graph launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi


class StreamSession:
    def __init__(self, model, n_streams: int = 1):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("StreamSession: move the model to a CUDA device first")
        self.model, self.device, self.n_streams = model, dev, int(n_streams)
        self._h = model._ensure(dev)
        self._lib = model._lib
        with torch.cuda.device(dev):
            capi.check(self._lib, self._h, self._lib.tip_stream_reset(self._h, self.n_streams),
                       "tip_stream_reset")

    @property
    def length(self) -> int:
        return int(self._lib.tip_stream_length(self._h))

    def reset(self):
        with torch.cuda.device(self.device):
            capi.check(self._lib, self._h, self._lib.tip_stream_reset(self._h, self.n_streams),
                       "tip_stream_reset")

    def step(self, imu_row, s_row):
        """imu_row (S, 72|90), s_row (S, size_s): numpy / CPU tensors (copied in, result returned
        as numpy after a stream sync) or CUDA tensors (asynchronous, CUDA tensor returned)."""
        m = self.model
        h = m._ensure(self.device, fast=True)
        S = self.n_streams
        on_host = not (isinstance(imu_row, torch.Tensor) and imu_row.is_cuda)
        if on_host:
            xi = np.ascontiguousarray(np.asarray(imu_row, dtype=np.float32).reshape(S, m._n_imu))
            xs = np.ascontiguousarray(np.asarray(s_row, dtype=np.float32).reshape(S, m._size_s))
            y = np.empty((S, m._size_s), dtype=np.float32)
            pi, ps, py = xi.ctypes.data, xs.ctypes.data, y.ctypes.data
        else:
            xi = imu_row.detach().to(torch.float32).contiguous().view(S, m._n_imu)
            xs = s_row.detach().to(device=self.device, dtype=torch.float32).contiguous().view(S, m._size_s)
            y = torch.empty((S, m._size_s), dtype=torch.float32, device=self.device)
            pi, ps, py = xi.data_ptr(), xs.data_ptr(), y.data_ptr()
        drop = m._dropout_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.tip_stream_step(h, pi, ps, py, int(on_host),
                                           C.byref(drop) if drop else None, C.c_void_p(stream))
        capi.check(self._lib, h, rc, "tip_stream_step")
        return y

    def step_raw(self, raw_imu, s_row):
        """Row N1: push RAW IMU frames (S, 72) = 6 global rotations + 6 global accelerations exactly as
        ``RTRunnerMin.step`` receives ``cur_imu`` (real_time_runner_minimal.py:118); smoothing, root-local
        rotation and the acc-sum feature run on the device.  Returns None for the first 5 calls (the
        runner returns ``s_init`` then, :125-128), afterwards ``y[:, L-1, :]`` like ``step``."""
        m = self.model
        h = m._ensure(self.device, fast=True)
        S = self.n_streams
        on_host = not (isinstance(raw_imu, torch.Tensor) and raw_imu.is_cuda)
        if on_host:
            xi = np.ascontiguousarray(np.asarray(raw_imu, dtype=np.float32).reshape(S, 72))
            xs = np.ascontiguousarray(np.asarray(s_row, dtype=np.float32).reshape(S, m._size_s))
            y = np.empty((S, m._size_s), dtype=np.float32)
            pi, ps, py = xi.ctypes.data, xs.ctypes.data, y.ctypes.data
        else:
            xi = raw_imu.detach().to(torch.float32).contiguous().view(S, 72)
            xs = s_row.detach().to(device=self.device, dtype=torch.float32).contiguous().view(S, m._size_s)
            y = torch.empty((S, m._size_s), dtype=torch.float32, device=self.device)
            pi, ps, py = xi.data_ptr(), xs.data_ptr(), y.data_ptr()
        drop = m._dropout_struct()
        produced = C.c_int(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.tip_stream_step_raw(h, pi, ps, py, int(on_host), C.byref(drop) if drop else None,
                                               C.c_void_p(stream), C.byref(produced))
        capi.check(self._lib, h, rc, "tip_stream_step_raw")
        return y if produced.value else None

    # ---- row N3: closed loop (post-model step on the device) -----------------------------------------
    @property
    def state_width(self) -> int:
        return int(self._lib.tip_stream_state_width(self._h))

    def set_state(self, s_init):
        """Initial state of every stream, like ``RTRunnerMin.__init__`` does with ``s_init``
        (real_time_runner_minimal.py:45-47): ``s_init`` is (S, 114) qdq rows (or (114,) for one stream);
        the x_s row ``record_state_aa_and_c(s_init, zeros)`` is computed here on the host and parked on
        the device, where the post step replaces it after every frame."""
        m = self.model
        S = self.n_streams
        s_init = np.asarray(s_init, dtype=np.float64).reshape(S, -1)
        rows = np.stack([state_to_row(s, np.zeros(m._size_s - 111)) for s in s_init]).astype(np.float32)
        rows = np.ascontiguousarray(rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.tip_stream_set_state(self._h, rows.ctypes.data, 1, C.c_void_p(stream))
        capi.check(self._lib, self._h, rc, "tip_stream_set_state")

    def step_closed(self, raw_imu, y_override=None):
        """Push one RAW IMU frame per stream (S, 72) and get the runner's per-frame pose back:
        (S, W) float64 = [s_t[3:60] | c_t | root_v] (see ``tip_stream_step_closed``), or None during the runner's
        5 warm-up calls.  The state row is fed back to the model on the device (no per-frame x_s upload).
        ``y_override`` (S, size_s) teacher-forces the post step (parity-test hook)."""
        m = self.model
        h = m._ensure(self.device, fast=True)
        S, W = self.n_streams, self.state_width
        on_host = not (isinstance(raw_imu, torch.Tensor) and raw_imu.is_cuda)
        keep = None
        if on_host:
            xi = np.ascontiguousarray(np.asarray(raw_imu, dtype=np.float32).reshape(S, 72))
            out = np.empty((S, W), dtype=np.float64)
            pi, po, py = xi.ctypes.data, out.ctypes.data, None
            if y_override is not None:
                keep = np.ascontiguousarray(np.asarray(y_override, dtype=np.float32).reshape(S, m._size_s))
                py = keep.ctypes.data
        else:
            xi = raw_imu.detach().to(torch.float32).contiguous().view(S, 72)
            out = torch.empty((S, W), dtype=torch.float64, device=self.device)
            pi, po, py = xi.data_ptr(), out.data_ptr(), None
            if y_override is not None:
                keep = y_override.detach().to(device=self.device, dtype=torch.float32).contiguous().view(S, m._size_s)
                py = keep.data_ptr()
        drop = m._dropout_struct()
        produced = C.c_int(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.tip_stream_step_closed(h, pi, py, po, int(on_host), C.byref(drop) if drop else None,
                                                  C.c_void_p(stream), C.byref(produced))
        capi.check(self._lib, h, rc, "tip_stream_step_closed")
        return out if produced.value else None

    def window(self, which="win_imu"):
        """Copy of the device-resident windows (test hook): (S, 40, 72|90) or (S, 40, size_s)."""
        m = self.model
        n = C.c_int64()
        capi.check(self._lib, self._h, self._lib.tip_debug_tensor(self._h, which.encode(), None, 0, C.byref(n), None),
                   "tip_debug_tensor")
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self._lib, self._h, self._lib.tip_debug_tensor(self._h, which.encode(), out.data_ptr(), n.value,
                                                                  C.byref(n), C.c_void_p(stream)), "tip_debug_tensor")
        return out.view(self.n_streams, 40, -1)


def _aa_to_rotmat(A):
    """Rotation vectors (n, 3) -> matrices (n, 3, 3): fairmotion ``conversions.A2R`` (= scipy
    ``Rotation.from_rotvec(A).as_matrix()``), closed form."""
    A = np.asarray(A, dtype=np.float64)
    angle = np.linalg.norm(A, axis=1)
    small = angle <= 1e-3
    a2 = angle * angle
    scale = np.where(small, 0.5 - a2 / 48 + a2 * a2 / 3840, np.sin(angle / 2) / np.where(small, 1.0, angle))
    x, y, z = (A * scale[:, None]).T
    w = np.cos(angle / 2)
    R = np.empty((A.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def state_to_row(qdq, c):
    """``record_state_aa_and_c`` (real_time_runner_minimal.py:78-85; ``batch_to_rot_mat_2axis``,
    data_utils.py:182-187): (114,) qdq + (n_c,) constraints -> the (111 + n_c,) x_s row: for the root and
    the 17 joints the first two columns of R, row-major, then the root velocity, then the constraints."""
    qdq = np.asarray(qdq, dtype=np.float64)
    r = _aa_to_rotmat(qdq[3:57].reshape(-1, 3))[:, :, :2].reshape(-1)
    return np.concatenate((r, qdq[57:60], np.asarray(c, dtype=np.float64)))
