"""CPU tests of the stitching host logic against the reference OnlineChainer goldens (tests/golden/chain_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from chain_cases import CASES, make_video
from oracle import cluster_oracle as co
from oracle import gather_oracle as go


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "chain_golden.npz"))


def oracle_local_labels(masks, subseqs):
    """Per-sub-clip cluster labels with cluster_label_start=1 from the CPU oracle (gather + clustering)."""
    coords, _ = go.masks_to_coord_list(masks.astype(bool))
    frames_list, labels_list, metas = [], [], []
    for s in subseqs:
        cs = [coords[t] for t in s["frames"]]
        e, b, sd = go.gather_foreground(cs, s["embeddings"], s["bandwidths"], s["seediness"])
        labels, meta = co.sequential_cluster(e, b, sd, 0.5, 0.3, 0.5, 2, [0.3, 0.3], cluster_label_start=1)
        counts = [len(c[0]) for c in cs]
        split = np.split(labels, np.cumsum(counts)[:-1])
        frames_list.append(list(s["frames"]))
        labels_list.append([torch.from_numpy(x.astype(np.int64)) for x in split])
        metas.append({k: v for k, v in meta.items() if k != "margin_ulps"})
    return frames_list, labels_list, metas


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_stitch_matches_reference(name, golden):
    from stemseg_b200.chaining import stitch_subsequences
    masks, subseqs = make_video(**CASES[name])
    frames_list, labels_list, metas = oracle_local_labels(masks, subseqs)
    container, subseq_labels, meta_out = stitch_subsequences(masks.shape[0], frames_list, labels_list, metas)
    track_labels, pt_counts, lifetimes = container.get_track_mask_idxes()
    for t, lab in enumerate(track_labels):
        np.testing.assert_array_equal(lab.numpy().astype(np.int32), golden["%s/track/%d" % (name, t)])
    ids = golden[name + "/ids"].tolist()
    assert sorted(pt_counts.keys()) == ids
    assert [pt_counts[i] for i in ids] == golden[name + "/pt_counts"].tolist()
    assert [lifetimes[i] for i in ids] == golden[name + "/lifetimes"].tolist()
    for i, labs in enumerate(subseq_labels):
        np.testing.assert_array_equal(torch.cat(labs).numpy().astype(np.int32), golden["%s/subseq/%d" % (name, i)])
    flat = sum([m["instance_labels"] + [-999] for m in meta_out], [])
    assert flat == golden[name + "/instance_labels"].tolist()


def test_subsequence_windows():
    from stemseg_b200.chaining import get_subsequence_frames
    w, pad = get_subsequence_frames(64, 16, "davis", 9)          # exactly 8 windows
    assert [(x[0], x[-1]) for x in w] == [(0, 15), (7, 22), (14, 29), (21, 36), (28, 43), (35, 50), (42, 57), (48, 63)]
    assert pad is None
    w, pad = get_subsequence_frames(5, 8, "davis", 4)            # short video: frame 0 repeated
    assert w == [[0, 0, 0, 0, 1, 2, 3, 4]] and pad == [True] * 3 + [False] * 5
    w, _ = get_subsequence_frames(20, 8, "ytvis")                # default overlap 4
    assert [x[0] for x in w] == [0, 4, 8, 12]
    with pytest.raises(NotImplementedError):
        get_subsequence_frames(20, 8, "unknown")


def test_association_edge_cases():
    from stemseg_b200.chaining import associate_label_sets
    a = np.array([1, 1, 2, 2, -1, 3], np.int64)
    b = np.array([7, 7, 7, 8, 8, -1], np.int64)
    assoc, un1, un2, costs, _ = associate_label_sets(a, b)
    assert (1, 7) in assoc and (2, 8) in assoc and un1 == {3} and un2 == set()
    assoc, un1, un2, _, _ = associate_label_sets(np.array([-1, -1]), np.array([4, 4]))
    assert assoc == [] and un2 == {4}
