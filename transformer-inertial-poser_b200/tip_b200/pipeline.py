"""Job pipeline over host buffers: many batches through ``tip_forward_host_submit`` / ``_wait``.

The reference evaluates recorded motions with one blocking ``model(x.cuda(), s.cuda()).cpu()`` after the
other (offline_testing_simple.py:360-399, real_time_runner_minimal.py:149-150): upload, forward and
download never overlap.  When the next batch is known before the previous result is needed, the three
legs of consecutive jobs can run concurrently; this class keeps ``depth`` jobs in flight and hands the
results back in submission order.
"""
from __future__ import annotations

from collections import deque

from . import capi


class HostPipeline:
    """``pipe = HostPipeline(model, depth=2)``; ``done = pipe.submit(x_imu, x_s, out)`` returns the oldest
    finished job's ``(x_imu, x_s, out)`` once ``depth`` jobs are in flight (else ``None``); ``pipe.drain()``
    yields the rest.  All tensors are pinned fp32 CPU tensors owned by the caller; a job's tensors may be
    re-used as soon as the job has been handed back."""

    def __init__(self, model, depth: int = 2, last_row_only: bool = False):
        if not 1 <= depth <= capi.TIP_HOST_SLOTS:
            raise ValueError(f"depth must be in 1..{capi.TIP_HOST_SLOTS}")
        self.model, self.depth, self.last_row_only = model, depth, last_row_only
        self._jobs = deque()          # (slot, x_imu, x_s, out), oldest first
        self._next = 0

    def __len__(self):
        return len(self._jobs)

    def _pop(self):
        slot, xi, xs, out = self._jobs.popleft()
        self.model.forward_host_wait(slot)
        return xi, xs, out

    def submit(self, x_imu, x_s, out):
        done = self._pop() if len(self._jobs) >= self.depth else None
        slot = self._next
        self._next = (self._next + 1) % self.depth
        self.model.forward_host_submit(slot, x_imu, x_s, out, last_row_only=self.last_row_only)
        self._jobs.append((slot, x_imu, x_s, out))
        return done

    def drain(self):
        while self._jobs:
            yield self._pop()
