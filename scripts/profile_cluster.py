"""Runs SequentialClustering alone (for `ncu -k regex:seq_cluster`, and to compare STEMSEG_CLUSTER_STREAM variants):
default = the HBM-resident cfg3 shape (E=8, 8 learned variances, N = 16x480x864 = 6 635 520).
    python scripts/profile_cluster.py [n] [e] [n_free]
Prints the roofline record and a checksum of the labels (identical across kernel variants)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from stemseg_b200.clusterers import SequentialClustering  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 480 * 864
e = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_free = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
rec = bench.cluster_roofline(dev, bench.load_peaks(), n, e, n_free, [0.3] * n_free, "profile run")
emb, seed, rng = bench.synthetic_points(n, e, dev)
v = e - n_free
bw = torch.full((n, v), 100.0, device=dev) if n_free else torch.from_numpy(
    np.exp(rng.uniform(-1, 1, size=(n, v))).astype(np.float32) * 10).to(dev)
labels, meta = SequentialClustering(0.5, 0.3, 0.0, n_free, [0.3] * n_free, dev)(emb, bandwidths=bw, seediness=seed)
w = torch.arange(n, device=dev, dtype=torch.int64) % 1000003 + 1
rec["labels_checksum"] = int((labels * w).sum())
rec["assigned"] = int((labels >= 0).sum())
rec["variant"] = os.environ.get("STEMSEG_CLUSTER_STREAM", "default")
print(json.dumps(rec))
