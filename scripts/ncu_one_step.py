"""Three eager (no CUDA graph) passes of the heads plan for `ncu -k regex:conv_tc_kernel -s 44 -c 22`: 22 conv launches
per pass, the third pass is the one captured.   python scripts/ncu_one_step.py fp32|bf16"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stemseg_b200.pipeline import build_davis_pipeline  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
t = 8 if precision == "fp32" else 16
device = torch.device("cuda:0")
pipe = build_davis_pipeline(device, num_frames=t, precision=precision)
feats = {s: f.to(device) for s, f in bench.make_features_cpu(seed=0, t=t).items()}
group = pipe._head_group()
group.use_graph, pipe.use_step_graph = False, False
for _ in range(3):
    pipe.run_heads(feats)
torch.cuda.synchronize()
print("done")
