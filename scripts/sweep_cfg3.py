"""configs[2] (16x480x864, bf16 decoder) step time against pipeline depth / plan switches.
    python scripts/sweep_cfg3.py [timeline]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stemseg_b200.pipeline import build_davis_pipeline  # noqa: E402

device = torch.device("cuda:0")
t16 = 16
feats = {s: f.to(device) for s, f in bench.make_features_cpu(seed=0, t=t16).items()}
mask = torch.ones((t16, bench.H4, bench.W4), dtype=torch.uint8, device=device)
flops = 2 * 2 * 564.87e9
peak = bench.load_peaks()["bf16_tflops_sustained"]


def run(pipe, steps):
    queue = []
    for _ in range(steps):
        queue.append(pipe.submit(feats, fg_mask=mask))
        if len(queue) > pipe.steps_in_flight:
            queue.pop(0).result()
    for q in queue:
        q.result()


def measure(depth, precision="bf16", steps=20, **switches):
    pipe = build_davis_pipeline(device, num_frames=t16, precision=precision)
    pipe.steps_in_flight = depth
    group = pipe._head_group()
    for k, v in switches.items():
        setattr(group, k, v)
    run(pipe, 6)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    run(pipe, steps)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print("depth %d %s %s: %.3f ms/step  %.1f TFLOP/s  frac %.3f" % (
        depth, precision, switches, ms, flops / ms / 1e9, flops / ms / 1e9 / peak), flush=True)
    return pipe


if len(sys.argv) > 1 and sys.argv[1] == "timeline":
    from torch.profiler import ProfilerActivity, profile
    pipe = measure(2)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(pipe, 4)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    for e in evs:
        name = e.name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").split("(")[0][:60]
        print("%9.1f %8.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, name))
elif len(sys.argv) > 1 and sys.argv[1] == "fp32":
    # configs[1]: 8 frames, fp32-parity arithmetic (1.13 TFLOP algorithmic per clip)
    t16 = 8
    feats = {s: f.to(device) for s, f in bench.make_features_cpu(seed=0, t=t16).items()}
    mask = torch.ones((t16, bench.H4, bench.W4), dtype=torch.uint8, device=device)
    flops = 2 * 2 * 282.43e9
    for depth in (1, 2, 3):
        measure(depth, precision="fp32")
else:
    for depth in (1, 2, 3, 4):
        measure(depth)
    measure(2, chunk_long_layers=False)
    measure(3, chunk_long_layers=False)
