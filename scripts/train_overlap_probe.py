"""Data-parallel training step (configs[4]) under torchrun: step time with the backward pass cut in two graphs (prefix
all-reduce overlapped with the block_8x / block_4x backward) against the single-graph schedule, plus a CUPTI timeline of
one step on rank 0.   torchrun --nproc-per-node N scripts/train_overlap_probe.py [timeline]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

import bench  # noqa: E402
from stemseg_b200 import heads  # noqa: E402
from stemseg_b200.losses import EmbeddingLoss  # noqa: E402
from stemseg_b200.training import DecoderTrainer  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dd = bench.Dist(rank, local, world)
device = dd.device


def build(split):
    torch.manual_seed(42)
    norm = lambda c: nn.GroupNorm(32, c)       # noqa: E731
    emb = heads.EmbeddingHead(256, [256, 256, 128, 128], 4, True, False, "xyff", NormType=norm, num_frames=8).to(device)
    seedh = heads.SeedinessHead(256, [256, 256, 128, 128], NormType=norm, num_frames=8).to(device)
    crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3], weight_variance_smoothness=10.0,
                         weight_lovasz=1.0, weight_regularization=0.001, weight_seediness=1.0, weight=1.0)
    tr = DecoderTrainer({"embedding": emb, "seediness": seedh}, crit)
    tr.split_backward = split
    return tr


feats_cpu, masks, ignore = bench.make_train_inputs(rank)
dev_feats = [f.to(device).requires_grad_(True) for f in feats_cpu]
targets = [{"masks": masks.to(device), "ignore_masks": ignore.to(device)}]


def run(tr, n):
    for _ in range(n):
        tr.step(dev_feats, targets)


results = {}
for split in ((False, True, True) if os.environ.get("PROBE_SHORT") else (False, True, False, True)):
    tr = build(split)
    run(tr, 5)
    ms, _ = dd.timed(lambda: run(tr, 30))
    results.setdefault(split, []).append(ms / 30)
    if rank == 0:
        print("world %d split_backward=%s conv_smem_kb=%s nccl_max_nchannels=%s: %.3f ms/step" % (
            world, split, os.environ.get("STEMSEG_CONV_SMEM_KB", "200"), os.environ.get("NCCL_MAX_NCHANNELS", "-"),
            ms / 30), flush=True)
    if split and len(sys.argv) > 1 and sys.argv[1] == "timeline" and len(results[True]) == (2 if not os.environ.get("PROBE_SHORT") else 2):
        from torch.profiler import ProfilerActivity, profile
        dd.barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run(tr, 2)
            torch.cuda.synchronize()
        if rank == 0:
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            evs.sort(key=lambda e: e.time_range.start)
            t0 = evs[0].time_range.start
            with open("gpurun_out/train_timeline_n%d.txt" % world, "w") as f:   # rank 0
                for e in evs:
                    name = e.name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").split("(")[0][:70]
                    f.write("%9.1f %8.1f  %s\n" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, name))
    tr.exchange.close()
    del tr
    torch.cuda.empty_cache()
dd.close()
