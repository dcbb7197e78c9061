"""Developer tool: full-width pipelines of the three shipped configs on one 8x480x864 clip (random init, synthetic
pyramid): checks they run at production widths and times the pipelined step.  usage: python scripts/time_shipped_configs.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from stemseg_b200.pipeline import build_pipeline      # noqa: E402

dev = torch.device("cuda:0")
T, HP, WP = 8, 480, 864
g = torch.Generator().manual_seed(0)
feats = {s: torch.randn(1, 256, T, HP // s, WP // s, generator=g).to(dev) for s in (32, 16, 8, 4)}
flops = {"davis": 2 * 564.9e9, "youtube_vis": 564.9e9 + 1072.6e9, "kitti_mots": 564.9e9 + 564.71e9}
for config in ("davis", "youtube_vis", "kitti_mots"):
    pipe = build_pipeline(config, dev, num_frames=T, min_seediness_prob=0.0)
    for _ in range(3):
        res = pipe(feats)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    steps, queue = 20, []
    a.record()
    for _ in range(steps):
        queue.append(pipe.submit(feats))
        if len(queue) > pipe.steps_in_flight:
            queue.pop(0).result()
    for q in queue:
        q.result()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(json.dumps({"config": config, "ms_per_clip": ms, "clips_per_sec": 1e3 / ms,
                      "algorithmic_tflops": flops[config] / ms / 1e9, "fg_points": res.fg_index.num_points,
                      "clusters": len(res.meta["instance_labels"]),
                      "semseg_logits": None if res.semseg_logits is None else list(res.semseg_logits.shape)}), flush=True)
    del pipe
