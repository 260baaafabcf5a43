"""SM clock / power draw while the laned forward runs for a few seconds (is throughput mode power-limited?)."""
import os, sys, threading, time
sys.path.insert(0, 'transformer-inertial-poser_b200'); sys.path.insert(0, '.')
import torch, pynvml
from bench import build_model, load_weights, synth
from tip_b200.pipeline import ForwardLanes
sd, _ = load_weights()
dev = torch.device('cuda:0')
B = 256; NS = 16
model = build_model(sd, dev)
sets = []
for i in range(NS):
    xi, xs = synth(1 + 1000 * i, B)
    sets.append((torch.from_numpy(xi).to(dev), torch.from_numpy(xs).to(dev)))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("power limit W:", pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1e3, "max sm MHz:", pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
for nl in [int(x) for x in os.environ.get("LANES", "1,3,5").split(",")]:
    lanes = ForwardLanes(model, nl)
    outs = [torch.empty((B, 40, 131), device=dev) for _ in range(nl)]
    def run(n):
        lanes.fork()
        for i in range(n):
            lanes.forward(i, *sets[i % NS], out=outs[i % nl])
        lanes.join()
    run(3 * NS * nl); torch.cuda.synchronize()
    samples = []; stop = [False]
    def poll():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.01)
    th = threading.Thread(target=poll); th.start()
    N = int(os.environ.get("N", "4000"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(N); e1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    ms = e0.elapsed_time(e1)
    body = samples[len(samples) // 4:]
    clk = sorted(s[0] for s in body); pw = sorted(s[1] for s in body)
    reasons = 0
    for s in body: reasons |= s[2]
    print(f"lanes {nl}: {ms / N * 1e3:.1f} us per forward over {ms / 1e3:.2f} s; SM MHz min/median/max {clk[0]}/{clk[len(clk)//2]}/{clk[-1]}; "
          f"power W median/max {pw[len(pw)//2]:.0f}/{pw[-1]:.0f}; throttle reasons mask 0x{reasons:x}", flush=True)
    del lanes
