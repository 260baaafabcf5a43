"""Keeping the GPU busy across calls: execution lanes and the host-buffer job pipeline.

The reference evaluates recorded motions with one blocking ``model(x.cuda(), s.cuda()).cpu()`` after the
other (offline_testing_simple.py:360-399, real_time_runner_minimal.py:149-150): upload, forward and
download never overlap, and one forward at a time leaves SMs idle in its narrow phases (the LayerNorm
GEMMs have 80 row tiles at B = 256, the recurrence occupies 104 SMs).  When the next batch is known
before the previous result is needed -- offline evaluation, a serving loop over many streams' windows --

* ``ForwardLanes`` runs forwards of device-resident batches on several lanes (``model.make_lane()``: shared
  parameters, own C handle / workspace) on their own CUDA streams, and
* ``HostPipeline`` keeps several host-buffer jobs in flight through ``tip_forward_host_submit`` / ``_wait``
  (upload, forward and download of consecutive jobs overlap), optionally spread over lanes,

both handing results back in submission order.  Windows stay independent; no arithmetic changes: outputs are
bit-identical to a blocking call's.  Handles that own or are lanes run in THROUGHPUT MODE (tip_set_tuning "auto"):
their wide GEMMs and LayerNorm GEMMs are launched narrow (about a quarter of the GPU each), so that several
lanes' kernels run side by side; 4-5 lanes is the measured sweet spot at B = 256 on a B200.
"""
from __future__ import annotations

from collections import deque

import torch

from . import capi


def _lanes_of(model, n_lanes: int):
    if n_lanes < 1:
        raise ValueError("lanes must be >= 1")
    return [model] + [model.make_lane() for _ in range(n_lanes - 1)]


class ForwardLanes:
    """``lanes = ForwardLanes(model, n_lanes=2)``; ``lanes.fork()``; ``y = lanes.forward(k, x_imu, x_s, out=...)``
    for k = 0, 1, 2, ... (job k runs on lane ``k % n_lanes``'s stream); ``lanes.join()`` makes the current
    stream wait for everything submitted.  Tensors passed in / returned are only safe to touch from another
    stream after ``join()``."""

    def __init__(self, model, n_lanes: int = 2):
        self.model = model
        self.models = _lanes_of(model, n_lanes)
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ForwardLanes: move the module to a CUDA device first (.cuda())")
        self.device = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
        self._ev = torch.cuda.Event()

    def __len__(self):
        return len(self.models)

    def fork(self):
        """Lane streams wait for the work queued on the current stream so far."""
        self._ev.record(torch.cuda.current_stream(self.device))
        for s in self.streams:
            s.wait_event(self._ev)

    def forward(self, k: int, x_imu, x_s, out=None):
        i = k % len(self.models)
        m = self.models[i]
        if m is not self.model:
            m._sync_lane_settings(self.model)
        caller = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.streams[i]):
            y = m(x_imu, x_s, out=out)
        if out is None:
            y.record_stream(caller)      # allocated on the lane's stream, consumed (after join()) on the caller's
        return y

    def join(self):
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def last_launch_count(self, k: int) -> int:
        return self.models[k % len(self.models)].last_launch_count()


class HostPipeline:
    """``pipe = HostPipeline(model, depth=2, lanes=1)``; ``done = pipe.submit(x_imu, x_s, out)`` returns the
    oldest finished job's ``(x_imu, x_s, out)`` once ``depth`` jobs are in flight (else ``None``);
    ``pipe.drain()`` yields the rest.  All tensors are pinned fp32 CPU tensors owned by the caller; a job's
    tensors may be re-used as soon as the job has been handed back.  ``lanes`` > 1 alternates the jobs over
    that many execution lanes, whose forwards overlap on the GPU (``depth`` >= 2 * lanes keeps every lane's
    copies hidden under its forwards)."""

    def __init__(self, model, depth: int = 2, last_row_only: bool = False, lanes: int = 1, lane_models=None):
        self.model = model
        # lane_models: re-use existing lanes (e.g. ``ForwardLanes.models``) instead of creating new handles
        self.models = list(lane_models) if lane_models is not None else _lanes_of(model, lanes)
        lanes = len(self.models)
        per_lane = -(-depth // lanes)
        if depth < 1 or per_lane > capi.TIP_HOST_SLOTS:
            raise ValueError(f"depth must be in 1..{capi.TIP_HOST_SLOTS * lanes} for {lanes} lane(s)")
        self.depth, self.last_row_only = depth, last_row_only
        self._per_lane = per_lane
        self._jobs = deque()          # (lane, slot, x_imu, x_s, out), oldest first
        self._n = 0

    def __len__(self):
        return len(self._jobs)

    def _pop(self):
        lane, slot, xi, xs, out = self._jobs.popleft()
        self.models[lane].forward_host_wait(slot)
        return xi, xs, out

    def submit(self, x_imu, x_s, out):
        done = self._pop() if len(self._jobs) >= self.depth else None
        lane = self._n % len(self.models)
        slot = (self._n // len(self.models)) % self._per_lane
        self._n += 1
        m = self.models[lane]
        if m is not self.model:
            m._sync_lane_settings(self.model)
        m.forward_host_submit(slot, x_imu, x_s, out, last_row_only=self.last_row_only)
        self._jobs.append((lane, slot, x_imu, x_s, out))
        return done

    def drain(self):
        while self._jobs:
            yield self._pop()
