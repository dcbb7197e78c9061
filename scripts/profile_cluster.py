"""Runs SequentialClustering alone (for `ncu -k regex:seq_cluster`): default = the HBM-resident cfg3 shape
(E=8, 8 learned variances, N = 16x480x864 = 6 635 520).  python scripts/profile_cluster.py [n] [e] [n_free]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 480 * 864
e = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_free = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
rec = bench.cluster_roofline(dev, bench.load_peaks(), n, e, n_free, [0.3] * n_free, "profile run")
print(rec)
