"""fairmotion.ops.math subset (test shim)."""
import numpy as np

from fairmotion.ops import conversions


def projectionOnVector(v1, v2):
    v2 = np.asarray(v2, dtype=float)
    return np.dot(v1, v2) / np.dot(v2, v2) * v2


def random_unit_vector(dim=3):
    while True:
        v = np.random.uniform(-1, 1, dim)
        n = np.linalg.norm(v)
        if 0 < n <= 1:
            return v / n


def project_rotation_1D(R, axis):
    Q = conversions.R2Q(R)
    axis = np.asarray(axis, dtype=float) / np.linalg.norm(axis)
    p = np.dot(Q[:3], axis) * axis
    Qp = np.concatenate((p, [Q[3]]))
    Qp = Qp / np.linalg.norm(Qp)
    a = conversions.Q2A(Qp)
    return float(np.dot(a, axis))


def project_rotation_2D(R, axis1, axis2, order="zyx"):
    raise NotImplementedError("not used on the TIP hot path")


def project_rotation_3D(R):
    return conversions.R2A(R)


def project_angular_vel_1D(w, axis):
    return np.linalg.norm(projectionOnVector(w, axis))


def project_angular_vel_2D(w, axis1, axis2):
    return np.array([np.linalg.norm(projectionOnVector(w, axis1)), np.linalg.norm(projectionOnVector(w, axis2))])


def project_angular_vel_3D(w):
    return w
