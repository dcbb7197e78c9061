"""Kernel timeline of one bench step via torch.profiler (CUPTI activity records; works for graph-launched kernels).
Usage (GPU box): python profiles/timeline.py > gpurun_out/timeline.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from stemseg_b200.pipeline import build_davis_pipeline

device = torch.device("cuda:0")
pipe = build_davis_pipeline(device, num_frames=bench.T)
feats = {s: f.to(device) for s, f in bench.make_features_cpu().items()}
mask = torch.ones((bench.T, bench.H4, bench.W4), dtype=torch.uint8, device=device)
for _ in range(4):
    pipe(feats, fg_mask=mask)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pend = None
    for _ in range(3):
        nxt = pipe.submit(feats, fg_mask=mask)
        if pend is not None:
            pend.result()
        pend = nxt
    pend.result()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    name = e.name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").split("(")[0][:60]
    print("%9.1f %8.1f  s%-3s %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start,
                                     getattr(e, "stream", "?") if hasattr(e, "stream") else "?", name))
