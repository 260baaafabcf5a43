"""fairmotion.ops.quaternion subset (test shim); quaternions are xyzw unless stated."""
import numpy as np


def Q_op(Q, op, xyzw_in=True):
    Q = np.array(Q, dtype=float)
    if "normalize" in op:
        Q = Q / np.linalg.norm(Q)
    if "halfspace" in op:
        w_idx = 3 if xyzw_in else 0
        if Q[w_idx] < 0.0:
            Q = -Q
    if "change_order" in op:
        Q = Q[[3, 0, 1, 2]] if xyzw_in else Q[[1, 2, 3, 0]]
    return Q


def Q_mult(Q1, Q2):
    """Hamilton product Q1 * Q2 (xyzw), un-normalised, any leading batch dimensions (data_utils.py:395 passes a
    quaternion DIFFERENCE, so nothing may be re-normalised here)."""
    Q1, Q2 = np.asarray(Q1, dtype=float), np.asarray(Q2, dtype=float)
    ax, ay, az, aw = Q1[..., 0], Q1[..., 1], Q1[..., 2], Q1[..., 3]
    bx, by, bz, bw = Q2[..., 0], Q2[..., 1], Q2[..., 2], Q2[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], axis=-1)


def Q_diff(Q1, Q2):
    """Q1^-1 * Q2 (batched; data_utils.py:318 loss_angle)."""
    return Q_mult(np.asarray(Q1, dtype=float) * np.array([-1.0, -1.0, -1.0, 1.0]), Q2)
