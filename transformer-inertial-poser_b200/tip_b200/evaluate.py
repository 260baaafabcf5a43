"""Row N4: batched offline evaluation -- many recorded motions as parallel streams.

The reference evaluates recorded motions one file after another, one frame after another, one model call
per frame at B = 1 (offline_testing_simple.py:360-399 -> :109-155).  The motions are independent, so here
motion i becomes stream i of ONE closed-loop ``StreamSession`` and every frame is a single batched call at
B = number of motions: the raw IMU rows of all motions go in, the per-frame poses of all motions come
back (``StreamSession.step_closed``: IMU pre-processing, forward, post-model step and state feedback all on
the device).  Per motion the result is exactly what a single-stream run produces (streams never mix).

What stays with the caller, unchanged and on the CPU: PyBullet FK, the SBP root-translation correction
and the metrics (offline_testing_simple.py:414-461, data_utils.py:314-391) -- they consume the arrays
returned here.
"""
from __future__ import annotations

import numpy as np

from .streaming import StreamSession

IMU_N_SMOOTH = 5            # constants.py:15: the first 5 runner calls return s_init (real_time_runner_minimal.py:125-128)


def run_motions(model, imus, s_inits, y_overrides=None):
    """Stream ``len(imus)`` recorded motions through the model in lockstep.

    imus     : list of (T_i, 72) raw IMU arrays (``data['imu']`` of the reference's pkl files,
               offline_testing_simple.py:366-368); lengths may differ.
    s_inits  : list of (114,) initial qdq states (``s_gt[0]``, :118).
    Returns a list of dicts, one per motion, with
      ``state`` (T_i, 57)  s_t[3:60] per runner call (rows of the 5 warm-up calls hold s_init[3:60]),
      ``ct``    (T_i, n_c) constraints per call (zeros during warm-up, like the runner),
      ``root_v``(T_i, 3)   filtered root velocity per call (what :159 integrates into the root position),
      ``valid`` (T_i,)     False for the warm-up calls.
    """
    S = len(imus)
    assert S >= 1 and len(s_inits) == S
    imus = [np.asarray(x, dtype=np.float32).reshape(-1, 72) for x in imus]
    lens = [x.shape[0] for x in imus]
    T = max(lens)
    sess = StreamSession(model, n_streams=S)
    sess.set_state(np.stack([np.asarray(s, dtype=np.float64) for s in s_inits]))
    W = sess.state_width
    n_c = W - 60
    out = [dict(state=np.zeros((n, 57)), ct=np.zeros((n, n_c)), root_v=np.zeros((n, 3)), valid=np.zeros(n, dtype=bool))
           for n in lens]
    frame = np.zeros((S, 72), dtype=np.float32)
    for t in range(T):
        for i in range(S):                      # a finished motion keeps replaying its last frame; its output is dropped
            frame[i] = imus[i][min(t, lens[i] - 1)]
        yo = None
        if y_overrides is not None:
            yo = np.stack([y_overrides[i][min(max(t - IMU_N_SMOOTH, 0), len(y_overrides[i]) - 1)] for i in range(S)])
        st = sess.step_closed(frame, y_override=yo)
        for i in range(S):
            if t >= lens[i]:
                continue
            if st is None:
                out[i]["state"][t] = np.asarray(s_inits[i], dtype=np.float64)[3:60]
            else:
                out[i]["state"][t] = st[i, :57]
                out[i]["ct"][t] = st[i, 57:57 + n_c]
                out[i]["root_v"][t] = st[i, 57 + n_c:]
                out[i]["valid"][t] = True
    return out
