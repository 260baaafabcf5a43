"""CPU, world_size=2 over gloo: the replica launcher's one collective (weight broadcast at init)
and the stream -> rank partitioning."""
import contextlib
import io
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tip_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "transformer-inertial-poser_b200"))
    from tip_b200 import TF_RNN_Past_State
    from tip_b200.replicas import broadcast_weights, flatten_params, local_streams
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                    # ranks start from DIFFERENT weights
    with contextlib.redirect_stdout(io.StringIO()):
        m = TF_RNN_Past_State(72, 131, 512, 1024, 256, 16, 4, 0.0, 0.0, 0.8, with_acc_sum=True)
    if rank == 0:
        sd = O.random_state_dict(5)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    before = flatten_params(m).clone()
    v0 = m.in_linear.weight._version
    moved = broadcast_weights(m, src=0)
    after = flatten_params(m)
    q.put((rank, moved, float(before.double().sum()), float(after.double().sum()),
           m.in_linear.weight._version > v0, local_streams(8, rank, world)))
    dist.destroy_process_group()


def test_weight_broadcast_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sd = O.random_state_dict(5)
    want = float(sum(np.asarray(v, dtype=np.float64).sum() for v in sd.values()))
    (r0, moved0, b0, a0, bumped0, s0), (r1, moved1, b1, a1, bumped1, s1) = res
    assert moved0 == moved1 == 14709260                 # one 14.7 MB blob
    assert abs(a0 - want) < 1e-6 and abs(a1 - want) < 1e-6 and abs(b1 - want) > 1e-3
    assert bumped1 and not bumped0                      # receivers re-pack; the source does not
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5, 7]
