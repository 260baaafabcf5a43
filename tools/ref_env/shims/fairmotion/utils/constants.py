"""fairmotion.utils.constants subset (test shim)."""
import numpy as np

EPSILON = np.finfo(float).eps
EYE_R = np.eye(3, dtype=float)
EYE_T = np.eye(4, dtype=float)
ZERO_P = np.zeros(3, dtype=float)
ZERO_R = np.zeros((3, 3), dtype=float)


def eye_T():
    return EYE_T.copy()


def eye_R():
    return EYE_R.copy()


def zero_p():
    return ZERO_P.copy()
