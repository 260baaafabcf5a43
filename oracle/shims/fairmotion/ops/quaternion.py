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
    """Hamilton product Q1 * Q2 (xyzw)."""
    ax, ay, az, aw = Q1
    bx, by, bz, bw = Q2
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def Q_diff(Q1, Q2):
    """Q1^-1 * Q2."""
    return Q_mult(np.asarray(Q1, dtype=float) * np.array([-1.0, -1.0, -1.0, 1.0]), Q2)
