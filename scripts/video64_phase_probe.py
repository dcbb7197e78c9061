"""Where a clip-parallel video spends its time after the sub-clips are done: host wall clock with a synchronisation at
every phase boundary (1 GPU; the exchange + stitch phase is what every rank of an N-GPU run executes)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stemseg_b200 import _lib  # noqa: E402
from stemseg_b200.chaining import DeviceStitcher, get_subsequence_frames  # noqa: E402
from stemseg_b200.parallel import _meta_dict  # noqa: E402
from stemseg_b200.pipeline import build_davis_pipeline  # noqa: E402

dev = torch.device("cuda:0")
t16 = 16
windows, _ = get_subsequence_frames(64, t16, "davis", 9)
pipe = build_davis_pipeline(dev, num_frames=t16, min_seediness_prob=0.0)
g = torch.Generator().manual_seed(1000)
feats = {s: torch.randn(1, 256, t16, bench.HP // s, bench.WP // s, generator=g).to(dev) for s in (32, 16, 8, 4)}
masks = torch.ones((64, bench.H4, bench.W4), dtype=torch.uint8, device=dev)
cap = bench.H4 * bench.W4
mi = pipe.clusterer.max_instances
meta_words = int(_lib.load().stemseg_seq_cluster_meta_words(4, mi))


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    t0 = sync()
    pend = pipe.submit(feats, fg_mask=masks[windows[0]], cluster_label_start=1)
    view = pend.device_view()
    t1 = sync()                                             # one sub-clip, unpipelined (what a rank of an 8-GPU run does)
    labels = [view["labels"].clone() for _ in windows]      # stand-in for the all-gathered buffers
    counts = view["counts"][:t16].contiguous()
    k = view["meta"][0:1].contiguous()
    stitcher = DeviceStitcher(64, cap, dev, max_instances=mi, max_subclips=len(windows))
    t2 = sync()
    for i, frames in enumerate(windows):
        stitcher.add_subclip(frames, labels[i], counts, k)
    t3 = time.perf_counter()
    t4 = sync()
    container, out_labels, out_meta = stitcher.finish()
    t5 = sync()
    print("rep %d: sub-clip latency %.3f ms | setup %.3f | stitch enqueue (host) %.3f, stitch device drain %.3f | "
          "finish %.3f" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4)),
          flush=True)
