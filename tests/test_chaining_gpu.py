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
