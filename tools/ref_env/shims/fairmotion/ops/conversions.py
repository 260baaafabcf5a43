"""fairmotion.ops.conversions subset (test shim): scipy Rotation based, quaternions xyzw, any leading
batch dimensions -- the same definitions fairmotion publishes."""
import numpy as np
from scipy.spatial.transform import Rotation


def _batch(x, fn, in_tail, out_tail):
    x = np.asarray(x, dtype=float)
    lead = x.shape[: x.ndim - in_tail]
    flat = x.reshape((-1,) + x.shape[x.ndim - in_tail:])
    out = np.asarray(fn(flat))
    return out.reshape(lead + out.shape[1:]) if lead else out.reshape(out.shape[1:])


def A2R(A):
    return _batch(A, lambda a: Rotation.from_rotvec(a).as_matrix(), 1, 2)


def R2A(R):
    return _batch(R, lambda r: Rotation.from_matrix(r).as_rotvec(), 2, 1)


def A2Q(A):
    return _batch(A, lambda a: Rotation.from_rotvec(a).as_quat(), 1, 1)


def Q2A(Q):
    return _batch(Q, lambda q: Rotation.from_quat(q).as_rotvec(), 1, 1)


def Q2R(Q):
    return _batch(Q, lambda q: Rotation.from_quat(q).as_matrix(), 1, 2)


def R2Q(R):
    return _batch(R, lambda r: Rotation.from_matrix(r).as_quat(), 2, 1)


def Ax2R(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])


def Ay2R(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0, s], [0, 1.0, 0], [-s, 0, c]])


def Az2R(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def Rp2T(R, p):
    R, p = np.asarray(R, dtype=float), np.asarray(p, dtype=float)
    T = np.zeros(R.shape[:-2] + (4, 4))
    T[..., :3, :3] = R
    T[..., :3, 3] = p
    T[..., 3, 3] = 1.0
    return T


def T2Rp(T):
    T = np.asarray(T, dtype=float)
    return T[..., :3, :3], T[..., :3, 3]


def T2R(T):
    return np.asarray(T, dtype=float)[..., :3, :3]


def T2p(T):
    return np.asarray(T, dtype=float)[..., :3, 3]


def R2T(R):
    return Rp2T(R, np.zeros(np.asarray(R).shape[:-2] + (3,)))


def p2T(p):
    p = np.asarray(p, dtype=float)
    return Rp2T(np.broadcast_to(np.eye(3), p.shape[:-1] + (3, 3)), p)


def Qp2T(Q, p):
    return Rp2T(Q2R(Q), p)


def Q2T(Q):
    return R2T(Q2R(Q))


def T2Qp(T):
    R, p = T2Rp(T)
    return R2Q(R), p
