"""Execution lanes: aggregate device time per whole-batch forward with 1..4 lanes (ForwardLanes) and
end-to-end host-buffer throughput with the job pipeline spread over lanes (HostPipeline)."""
import os, sys, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch
from bench import build_model, load_weights, synth
from tip_b200.pipeline import ForwardLanes, HostPipeline
sd, _ = load_weights()
dev = torch.device('cuda:0')
B = int(os.environ.get('B', '256'))
N = int(os.environ.get('N', '192'))
model = build_model(sd, dev)
NS = 16
sets = []
for i in range(NS):
    xi, xs = synth(1 + 1000 * i, B)
    sets.append((torch.from_numpy(xi).to(dev), torch.from_numpy(xs).to(dev)))
ref = model(*sets[5]).clone()
for nl in (1, 2, 3, 4):
    lanes = ForwardLanes(model, nl)
    outs = [torch.empty((B, 40, 131), device=dev) for _ in range(nl)]
    def run(n):
        lanes.fork()
        for i in range(n):
            lanes.forward(i, *sets[i % NS], out=outs[i % nl])
        lanes.join()
    run(3 * NS * nl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(); run(N); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / N)
    lanes.fork(); y = lanes.forward(nl - 1, *sets[5], out=outs[nl - 1]); lanes.join(); torch.cuda.synchronize()
    print("device lanes %d: %.1f us per forward -> %.0f frames/s; same result: %s" % (nl, best * 1e3, B / best * 1e3, bool(torch.equal(y, ref))))
    del lanes
hx = [(torch.from_numpy(synth(7000 + i, B)[0]).pin_memory(), torch.from_numpy(synth(7000 + i, B)[1]).pin_memory()) for i in range(9)]
hy = [torch.empty((B, 40, 131)).pin_memory() for _ in range(9)]
for nl, depth in ((1, 2), (2, 2), (2, 4), (2, 6), (3, 6)):
    pipe = HostPipeline(model, depth=depth, lanes=nl)
    nb = depth + 1
    for i in range(4 * nb):
        pipe.submit(hx[i % nb][0], hx[i % nb][1], hy[i % nb])
    for _ in pipe.drain(): pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    chk = 0.0
    for i in range(N):
        d = pipe.submit(hx[i % nb][0], hx[i % nb][1], hy[i % nb])
        if d is not None: chk += float(d[2][0, -1, 0])
    for d in pipe.drain(): chk += float(d[2][0, -1, 0])
    el = (time.perf_counter() - t0) / N
    print("host pipeline lanes %d depth %d: %.1f us per job -> %.0f frames/s" % (nl, depth, el * 1e6, B / el))
