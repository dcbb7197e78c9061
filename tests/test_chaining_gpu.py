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


def _run_device_stitcher(cuda_device, num_frames, frames_list, labels_list, ks, cap):
    from stemseg_b200.chaining import DeviceStitcher
    st = DeviceStitcher(num_frames, cap, cuda_device, max_instances=20, max_subclips=len(frames_list))
    for frames, labs, k in zip(frames_list, labels_list, ks):
        flat = torch.cat(labs).to(cuda_device)
        pad = torch.full((37,), 12345, dtype=torch.int64, device=cuda_device)       # capacity > valid points
        counts = torch.tensor([x.numel() for x in labs], dtype=torch.int32, device=cuda_device)
        st.add_subclip(frames, torch.cat([flat, pad]), counts, torch.tensor([k], dtype=torch.int32, device=cuda_device))
    return st.finish()


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_device_stitcher_matches_reference(name, golden_dir, cuda_device):
    """Fully device-resident stitch (assignment + set ordering on the GPU, one read-back per video) against the
    goldens of the reference OnlineChainer."""
    from test_chaining_cpu import oracle_local_labels
    golden = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    masks, subseqs = make_video(**CASES[name])
    frames_list, labels_list, metas = oracle_local_labels(masks, subseqs)
    ks = [len(m["instance_labels"]) for m in metas]
    container, subseq_labels, meta_out = _run_device_stitcher(cuda_device, masks.shape[0], frames_list, labels_list, ks,
                                                              masks.shape[1] * masks.shape[2])
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


@pytest.mark.parametrize("seed", range(6))
def test_device_stitcher_matches_host_stitch_on_tie_heavy_labels(seed, cuda_device):
    """Random label vectors with many zero-IoU pairs (cost ties) and label ids large enough to leave CPython's
    small-set ascending order: the device stitch must reproduce the host stitch (python sets + scipy) exactly."""
    from stemseg_b200.chaining import stitch_subsequences
    rng = np.random.default_rng(100 + seed)
    t_sub, overlap, n_sub = 6, 3, 7
    frames_list = [list(range(i * (t_sub - overlap), i * (t_sub - overlap) + t_sub)) for i in range(n_sub)]
    num_frames = frames_list[-1][-1] + 1
    counts = rng.integers(0, 400, size=num_frames)
    counts[rng.integers(0, num_frames)] = 0                                   # an empty frame
    labels_list, ks = [], []
    for frames in frames_list:
        k = int(rng.integers(1, 21))
        used = rng.choice(np.arange(1, k + 1), size=int(rng.integers(1, k + 1)), replace=False)
        labs = []
        for t in frames:
            # blocky labels so that IoUs are mostly 0 or large; some outliers
            seg = rng.choice(np.concatenate([used, [-1]]), size=max(1, counts[t] // 40 + 1))
            lab = np.repeat(seg, 40)[:counts[t]]
            labs.append(torch.from_numpy(lab.astype(np.int64)))
        labels_list.append(labs)
        ks.append(k)
    metas = [{"instance_labels": list(range(1, k + 1))} for k in ks]
    ref_container, ref_labels, ref_meta = stitch_subsequences(num_frames, frames_list, labels_list, metas)
    container, subseq_labels, meta_out = _run_device_stitcher(cuda_device, num_frames, frames_list, labels_list, ks,
                                                              int(counts.max()) + 1)
    r_tracks, r_counts, r_life = ref_container.get_track_mask_idxes()
    tracks, pt_counts, lifetimes = container.get_track_mask_idxes()
    for t in range(num_frames):
        assert torch.equal(tracks[t].cpu(), r_tracks[t]), "frame %d" % t
    assert dict(pt_counts) == dict(r_counts) and dict(lifetimes) == dict(r_life)
    for got, ref in zip(subseq_labels, ref_labels):
        assert torch.equal(got.cpu(), torch.cat(ref))
    assert [m["instance_labels"] for m in meta_out] == [m["instance_labels"] for m in ref_meta]


def test_device_stitcher_reports_mismatched_overlap(cuda_device):
    """Overlap frames with different point sets in the two sub-clips: loud failure, no OOB read."""
    from stemseg_b200.chaining import DeviceStitcher
    st = DeviceStitcher(6, 64, cuda_device, max_instances=20, max_subclips=2)
    one = torch.ones(4 * 10, dtype=torch.int64, device=cuda_device)
    k = torch.tensor([1], dtype=torch.int32, device=cuda_device)
    st.add_subclip([0, 1, 2, 3], one.clone(), torch.full((4,), 10, dtype=torch.int32, device=cuda_device), k)
    st.add_subclip([2, 3, 4, 5], one.clone(), torch.tensor([10, 9, 10, 11], dtype=torch.int32, device=cuda_device), k)
    with pytest.raises(AssertionError, match="Shape mismatch"):
        st.finish()
