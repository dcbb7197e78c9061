"""Generate tests/golden/cluster_golden.npz by running the UNMODIFIED reference clusterer on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/gen_cluster_golden.py
For every case of tests/cluster_cases.case_table() it runs
``stemseg.inference.clusterers.SequentialClustering(device="cpu")`` (clusterers.py:34-175), checks that
oracle/cluster_oracle.py reproduces labels / instance_labels / centers / stds bit-for-bit, checks that no decision
is within 8 fp32 ulps of its threshold (the reference's own CPU and CUDA paths disagree in the last ulp -- see the
oracle's docstring), and stores the reference's outputs.  Also pins ``aten_inner_sum_f32`` against torch for E=1..24.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from stemseg.inference.clusterers import SequentialClustering  # noqa: E402

from cluster_cases import case_table, make_points  # noqa: E402
from oracle import cluster_oracle as co  # noqa: E402


def main():
    # pin the summation order against torch itself
    rng = np.random.default_rng(0)
    for e in range(1, 25):
        t = (rng.standard_normal((257, e)) * 10.0 ** rng.integers(-2, 3, size=(1, e))).astype(np.float32)
        ref = torch.from_numpy(t).sum(dim=-1).numpy()
        got = co.aten_inner_sum_f32(t)
        assert (ref == got).all(), "ATen sum order mismatch for E=%d" % e
    print("aten_inner_sum_f32 pinned for E=1..24")

    out = {}
    for name, (pts, clu) in case_table().items():
        emb, bw, seed = make_points(**pts)
        ref = SequentialClustering(clu['primary_prob_thresh'], clu['secondary_prob_thresh'],
                                   clu['min_seediness_prob'], clu['n_free_dims'], clu['free_dim_stds'], "cpu",
                                   max_instances=clu['max_instances'])
        labels, meta = ref(torch.from_numpy(emb), bandwidths=torch.from_numpy(bw), seediness=torch.from_numpy(seed),
                           cluster_label_start=clu['cluster_label_start'], return_label_masks=True)
        o_labels, o_meta = co.sequential_cluster(emb, bw, seed, return_label_masks=True, **clu)
        labels = labels.numpy()
        assert labels.dtype == np.int64
        assert (labels == o_labels).all(), "%s: oracle labels differ from the reference at %d points" % (
            name, int((labels != o_labels).sum()))
        assert meta['instance_labels'] == o_meta['instance_labels'], name
        assert np.array_equal(np.array(meta['instance_centers'], np.float32).reshape(-1),
                              np.array(o_meta['instance_centers'], np.float32).reshape(-1)), name
        # torch's CPU sqrt is not correctly rounded (1 ulp low in ~0.6 % of cases): stds are compared to 2 ulp
        assert np.allclose(np.array(meta['instance_stds'], np.float32).reshape(-1),
                           np.array(o_meta['instance_stds'], np.float32).reshape(-1), rtol=3e-7, atol=0), name
        for a, b in zip(meta['instance_masks'], o_meta['instance_masks']):
            assert np.array_equal(a.numpy(), b), name
        assert o_meta['margin_ulps'] >= 8, "%s: ambiguous case (margin %s ulps) -- change its seed" % (
            name, o_meta['margin_ulps'])
        k = len(meta['instance_labels'])
        e = emb.shape[1]
        out[name + "/labels"] = labels.astype(np.int32)
        out[name + "/instance_labels"] = np.array(meta['instance_labels'], np.int64)
        out[name + "/instance_centers"] = np.array(meta['instance_centers'], np.float32).reshape(k, e)
        out[name + "/instance_stds"] = np.array(meta['instance_stds'], np.float32).reshape(k, e)
        out[name + "/mask_counts"] = np.array([int(m.sum()) for m in meta['instance_masks']], np.int64)
        print("%-22s N=%6d E=%d K=%2d unassigned=%6d margin=%s" % (
            name, emb.shape[0], e, k, int((labels == -1).sum()), o_meta['margin_ulps']))
    path = os.path.join(HERE, "cluster_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
