"""GPU: OnlineChainer.process (CUDA gather + clustering, host stitch) against the reference OnlineChainer goldens."""
import os

import numpy as np
import pytest
import torch

from chain_cases import CASES, make_video

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_online_chainer_matches_reference(name, golden_dir, cuda_device):
    from stemseg_b200.chaining import OnlineChainer
    from stemseg_b200.clusterers import SequentialClustering
    golden = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    masks, subseqs = make_video(**CASES[name])
    chainer = OnlineChainer(SequentialClustering(0.5, 0.3, 0.5, 2, [0.3, 0.3], cuda_device), 1.0)
    t_subseqs = [{"frames": list(s["frames"]), "embeddings": torch.from_numpy(s["embeddings"]),
                  "bandwidths": torch.from_numpy(s["bandwidths"]), "seediness": torch.from_numpy(s["seediness"])}
                 for s in subseqs]
    (track_labels, pt_counts, lifetimes), mask_idxes, subseq_labels, _, metas = chainer.process(
        torch.from_numpy(masks), t_subseqs)
    for t, lab in enumerate(track_labels):
        np.testing.assert_array_equal(lab.numpy().astype(np.int32), golden["%s/track/%d" % (name, t)])
    ids = golden[name + "/ids"].tolist()
    assert sorted(pt_counts.keys()) == ids
    assert [pt_counts[i] for i in ids] == golden[name + "/pt_counts"].tolist()
    assert [lifetimes[i] for i in ids] == golden[name + "/lifetimes"].tolist()
    flat = sum([m["instance_labels"] + [-999] for m in metas], [])
    assert flat == golden[name + "/instance_labels"].tolist()
    # mask_idxes has the reference's list(T) of (y, x) layout
    assert len(mask_idxes) == masks.shape[0]
    y, x = mask_idxes[0]
    ry, rx = np.nonzero(masks[0])
    np.testing.assert_array_equal(y.cpu().numpy(), ry)
    np.testing.assert_array_equal(x.cpu().numpy(), rx)


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_device_stitch_matches_reference(name, golden_dir, cuda_device):
    """Labels stay on the GPU: pair-histogram + LUT kernels, Hungarian on the host -> identical track ids."""
    from test_chaining_cpu import oracle_local_labels
    from stemseg_b200.chaining import stitch_subsequences_device
    golden = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    masks, subseqs = make_video(**CASES[name])
    frames_list, labels_list, metas = oracle_local_labels(masks, subseqs)
    dev_labels = [torch.cat(l).to(cuda_device) for l in labels_list]
    counts = [[x.numel() for x in l] for l in labels_list]
    ks = [len(m["instance_labels"]) for m in metas]
    container, subseq_labels, meta_out = stitch_subsequences_device(masks.shape[0], frames_list, dev_labels, counts,
                                                                    ks, metas)
    track_labels, pt_counts, lifetimes = container.get_track_mask_idxes()
    for t, lab in enumerate(track_labels):
        assert lab.is_cuda
        np.testing.assert_array_equal(lab.cpu().numpy().astype(np.int32), golden["%s/track/%d" % (name, t)])
    ids = golden[name + "/ids"].tolist()
    assert sorted(pt_counts.keys()) == ids
    assert [pt_counts[i] for i in ids] == golden[name + "/pt_counts"].tolist()
    assert [lifetimes[i] for i in ids] == golden[name + "/lifetimes"].tolist()
    for i, labs in enumerate(subseq_labels):
        np.testing.assert_array_equal(labs.cpu().numpy().astype(np.int32), golden["%s/subseq/%d" % (name, i)])
    flat = sum([m["instance_labels"] + [-999] for m in meta_out], [])
    assert flat == golden[name + "/instance_labels"].tolist()


def test_full_resolution_clustering(cuda_device):
    """embedding_resize_factor = 2: maps stay at low resolution, the mask is at 2x; the fused gather evaluates the
    trilinear resize at the foreground voxels; clustering is bit-exact against the oracle on the gathered points."""
    from oracle import cluster_oracle as co
    from stemseg_b200.chaining import OnlineChainer
    from stemseg_b200.clusterers import SequentialClustering
    from stemseg_b200.foreground import compact_foreground, gather_points
    masks, subseqs = make_video(**CASES["three_blobs"])
    s = subseqs[0]
    up = 2
    big = np.repeat(np.repeat(masks[s["frames"]], up, axis=1), up, axis=2)
    chainer = OnlineChainer(SequentialClustering(0.5, 0.3, 0.5, 2, [0.3, 0.3], cuda_device), float(up))
    emb, bw, sd = [torch.from_numpy(s[k]).to(cuda_device) for k in ("embeddings", "bandwidths", "seediness")]
    fg = compact_foreground(torch.from_numpy(big).to(cuda_device))
    labels, emb_flat, meta = chainer.cluster_subsequence(fg, emb, bw, sd, 5, False)
    assert [l.numel() for l in labels] == fg.frame_counts and emb_flat.shape == (fg.num_points, 4)
    e = emb_flat.cpu().numpy()
    b = gather_points(bw, fg, upsample=up).cpu().numpy()
    d = gather_points(sd, fg, upsample=up).cpu().numpy()
    o_labels, o_meta = co.sequential_cluster(e, b, d, 0.5, 0.3, 0.5, 2, [0.3, 0.3], cluster_label_start=5)
    np.testing.assert_array_equal(torch.cat(labels).cpu().numpy(), o_labels)
    assert meta["instance_labels"] == o_meta["instance_labels"]
    with pytest.raises(NotImplementedError):
        OnlineChainer(chainer.clusterer, 1.5)
