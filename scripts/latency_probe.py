"""Per-call host wall time of the synchronous single-clip call pipe(feats) and of submit / result separately."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stemseg_b200.pipeline import build_davis_pipeline  # noqa: E402

device = torch.device("cuda:0")
pipe = build_davis_pipeline(device, num_frames=bench.T)
feats = {s: f.to(device) for s, f in bench.make_features_cpu().items()}
mask = torch.ones((bench.T, bench.H4, bench.W4), dtype=torch.uint8, device=device)
for _ in range(5):
    pipe(feats, fg_mask=mask)
torch.cuda.synchronize()
for i in range(8):
    t0 = time.perf_counter()
    pend = pipe.submit(feats, fg_mask=mask)
    t1 = time.perf_counter()
    pend._done.synchronize()
    t2 = time.perf_counter()
    res = pend.result()
    t3 = time.perf_counter()
    print("call %d: submit %.3f ms, wait %.3f ms, result() host part %.3f ms" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1),
                                                                                 1e3 * (t3 - t2)), flush=True)
emb, var, seedi, _ = pipe.run_heads(feats)
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe.cluster(emb, var, seedi, mask)
    torch.cuda.synchronize()
    print("eager cluster %d: %.3f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe(feats, fg_mask=mask)
    torch.cuda.synchronize()
    print("pipe() after eager calls %d: %.3f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
