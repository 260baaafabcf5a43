import torch, time
for mb in (2.3, 9.05, 5.4, 64):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
    for direction in ('h2d', 'd2h'):
        for _ in range(3): (d.copy_(h, non_blocking=True) if direction == 'h2d' else h.copy_(d, non_blocking=True)); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): (d.copy_(h, non_blocking=True) if direction == 'h2d' else h.copy_(d, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{mb} MB {direction}: {ms*1e3:.0f} us, {n/ms/1e6:.1f} GB/s")
# full duplex
n = int(9e6); h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 10
print(f"duplex 9 MB each way: {el*1e6:.0f} us per pair")
