"""Per-kernel device time of one training step (bench.py --workload train geometry) via torch.profiler (CUPTI).
Usage (GPU box): python profiles/train_timeline.py > gpurun_out/train_timeline.txt"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from torch.profiler import ProfilerActivity, profile

import bench
from stemseg_b200 import heads
from stemseg_b200.losses import EmbeddingLoss
from stemseg_b200.training import DecoderTrainer

device = torch.device("cuda:0")
torch.manual_seed(42)
norm = lambda c: nn.GroupNorm(32, c)       # noqa: E731
emb = heads.EmbeddingHead(bench.IN_CH, list(bench.INTER), 4, True, False, "xyff", NormType=norm,
                          num_frames=bench.TRAIN_T).to(device)
seedh = heads.SeedinessHead(bench.IN_CH, list(bench.INTER), NormType=norm, num_frames=bench.TRAIN_T).to(device)
crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3], weight_variance_smoothness=10.0,
                     weight_lovasz=1.0, weight_regularization=0.001, weight_seediness=1.0, weight=1.0)
trainer = DecoderTrainer({"embedding": emb, "seediness": seedh}, crit)
feats_cpu, masks, ignore = bench.make_train_inputs(0)
feats = [f.to(device).requires_grad_(True) for f in feats_cpu]
targets = [{"masks": masks.to(device), "ignore_masks": ignore.to(device)}]
for _ in range(3):
    trainer.step(feats, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.step(feats, targets)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = defaultdict(lambda: [0, 0.0])
for e in evs:
    name = e.name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").split("(")[0][:70]
    agg[name][0] += 1
    agg[name][1] += e.time_range.end - e.time_range.start
total = sum(v[1] for v in agg.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print("# one training step (8x384x640 clip, 2 heads, fp32-parity mode): %d kernels, %.1f us busy, %.1f us span" % (
    len(evs), total, span))
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %4d launches %10.1f us %5.1f%%" % (name, cnt, us, 100.0 * us / total))
if "--timeline" in sys.argv:
    t0 = evs[0].time_range.start
    for e in evs:
        name = e.name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").split("(")[0][:60]
        print("%9.1f %8.1f %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, name))
