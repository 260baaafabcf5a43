"""Per-rank PCIe copy bandwidth with all ranks copying AT ONCE (the e2e leg's situation), with and without binding
each rank to its GPU's CPUs / NUMA node before the pinned buffers are allocated.

    BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_probe_ranks.py

Prints one line per rank (GPU NUMA node from sysfs, CPUs allowed, duplex H2D / D2H GB/s of 9.05 MB + 5.37 MB copies --
the bench's per-step bytes) and the aggregate on rank 0."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
if world == 0:
    world = 1
bind = os.environ.get("BIND", "1") == "1"
n_aff = None
if bind:
    from bench import bind_to_gpu_cpus
    n_aff = bind_to_gpu_cpus(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
numa = "?"
try:
    import pynvml
    pynvml.nvmlInit()
    bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
    bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    numa = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
except Exception as e:
    numa = f"? ({type(e).__name__})"
H2D, D2H = 9_052_160, 5_365_760
hs = [torch.empty(H2D, dtype=torch.uint8).pin_memory() for _ in range(3)]
ho = [torch.empty(D2H, dtype=torch.uint8).pin_memory() for _ in range(3)]
for t in hs + ho:
    t.fill_(1)                                   # first touch on the bound CPUs
ds = [torch.empty(H2D, dtype=torch.uint8, device="cuda") for _ in range(3)]
do = [torch.empty(D2H, dtype=torch.uint8, device="cuda") for _ in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(n):
    for i in range(n):
        with torch.cuda.stream(s1):
            ds[i % 3].copy_(hs[i % 3], non_blocking=True)
        with torch.cuda.stream(s2):
            ho[i % 3].copy_(do[i % 3], non_blocking=True)
    torch.cuda.synchronize()


run(10)
if world > 1:
    dist.barrier()
N = 200
t0 = time.perf_counter()
run(N)
el = time.perf_counter() - t0
if world > 1:
    dist.barrier()
h2d, d2h = H2D * N / el / 1e9, D2H * N / el / 1e9
print(f"rank {rank} gpu {local} numa {numa} bind {int(bind)} cpus_allowed {len(os.sched_getaffinity(0))} of {os.cpu_count()} "
      f"(nvml mask -> {n_aff}): duplex H2D {h2d:.1f} GB/s + D2H {d2h:.1f} GB/s, {el / N * 1e6:.0f} us per step-equivalent "
      f"-> copy-bound ceiling {256 * N / el / 1e3:.0f} k frames/s", flush=True)
if world > 1:
    t = torch.tensor([h2d, d2h])
    dist.all_reduce(t)
    if rank == 0:
        print(f"aggregate over {world} ranks: H2D {t[0]:.1f} GB/s, D2H {t[1]:.1f} GB/s", flush=True)
    dist.destroy_process_group()
