"""GPU parity: stemseg_b200.SequentialClustering (CUDA, through the C ABI) vs the oracle / the reference goldens."""
import os

import numpy as np
import pytest
import torch

from cluster_cases import case_table, make_points
from oracle import cluster_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "cluster_golden.npz"))


def run_cuda(emb, bw, seed, clu, device, return_label_masks=True):
    from stemseg_b200.clusterers import SequentialClustering
    c = SequentialClustering(clu['primary_prob_thresh'], clu['secondary_prob_thresh'], clu['min_seediness_prob'],
                             clu['n_free_dims'], clu['free_dim_stds'], device, max_instances=clu['max_instances'])
    labels, meta = c(torch.from_numpy(emb).to(device), bandwidths=torch.from_numpy(bw).to(device),
                     seediness=torch.from_numpy(seed).to(device), cluster_label_start=clu['cluster_label_start'],
                     return_label_masks=return_label_masks)
    return labels, meta


@pytest.mark.parametrize("name", sorted(case_table().keys()))
def test_cuda_matches_reference_golden(name, golden, cuda_device):
    pts, clu = case_table()[name]
    emb, bw, seed = make_points(**pts)
    labels, meta = run_cuda(emb, bw, seed, clu, cuda_device)
    assert labels.dtype == torch.int64 and labels.device.type == "cuda"
    np.testing.assert_array_equal(labels.cpu().numpy(), golden[name + "/labels"].astype(np.int64))   # bit-exact
    assert meta['instance_labels'] == golden[name + "/instance_labels"].tolist()
    k, e = len(meta['instance_labels']), emb.shape[1]
    np.testing.assert_array_equal(np.array(meta['instance_centers'], np.float32).reshape(k, e),
                                  golden[name + "/instance_centers"])
    np.testing.assert_allclose(np.array(meta['instance_stds'], np.float32).reshape(k, e),
                               golden[name + "/instance_stds"], rtol=3e-7, atol=0)
    assert [int(m.sum()) for m in meta['instance_masks']] == golden[name + "/mask_counts"].tolist()


@pytest.mark.parametrize("e,nf,n", [(4, 2, 207360), (3, 0, 207360), (8, 0, 414720), (5, 2, 100003), (4, 2, 3317760),
                                    (8, 0, 6635520)])
def test_cuda_matches_oracle_large(e, nf, n, cuda_device):
    """BASELINE config sizes: quarter-res 8x480x864 (207 360 points), cfg3 E=8 (414 720), full-res 3.3 M, and the
    HBM-resident cfg3 full-resolution shape (E=8, 8 learned variances, N = 16x480x864 = 6 635 520: bit-mask streaming
    kernel)."""
    emb, bw, seed = make_points(seed=77 + e, n=n, e=e, n_free=nf, n_blobs=14, noise_frac=0.2)
    clu = dict(primary_prob_thresh=0.5, secondary_prob_thresh=0.3, min_seediness_prob=0.0, n_free_dims=nf,
               free_dim_stds=[0.3, 0.3][:nf], max_instances=20, cluster_label_start=3)
    o_labels, o_meta = co.sequential_cluster(emb, bw, seed, **clu)
    labels, meta = run_cuda(emb, bw, seed, clu, cuda_device, return_label_masks=False)
    got = labels.cpu().numpy()
    # identical even when a point sits exactly on a threshold (margin_ulps == 0): both sides test d <= d*
    np.testing.assert_array_equal(got, o_labels)
    assert meta['instance_labels'] == o_meta['instance_labels']
    np.testing.assert_array_equal(np.array(meta['instance_centers'], np.float32),
                                  np.array(o_meta['instance_centers'], np.float32))


def test_properties_full_size(cuda_device):
    """Size-independent properties at full resolution: label-offset invariance, idempotent re-run, label range."""
    emb, bw, seed = make_points(seed=3, n=3317760, e=4, n_free=2, n_blobs=18, noise_frac=0.25)
    clu = dict(primary_prob_thresh=0.5, secondary_prob_thresh=0.3, min_seediness_prob=0.3, n_free_dims=2,
               free_dim_stds=[0.3, 0.3], max_instances=20, cluster_label_start=1)
    a, ma = run_cuda(emb, bw, seed, clu, cuda_device, return_label_masks=False)
    b, mb = run_cuda(emb, bw, seed, dict(clu, cluster_label_start=101), cuda_device, return_label_masks=False)
    a2, _ = run_cuda(emb, bw, seed, clu, cuda_device, return_label_masks=False)
    assert torch.equal(a, a2)                                           # deterministic
    assert torch.equal(torch.where(a < 0, a, a + 100), b)               # offset invariant (SURVEY §8a quirk v)
    k = len(ma['instance_labels'])
    assert int(a.max()) <= k and int(a.min()) >= -1
    # every seed point belongs to its own cluster unless a later secondary pass could not touch it
    assert ma['instance_centers'] == mb['instance_centers']


def test_api_behaviour(cuda_device):
    from stemseg_b200.clusterers import SequentialClustering
    c = SequentialClustering(0.5, 0.3, 0.8, 2, [0.3, 0.3], cuda_device)
    labels, meta = c(torch.zeros(0, 4, device=cuda_device), bandwidths=torch.zeros(0, 2, device=cuda_device),
                     seediness=torch.zeros(0, 1, device=cuda_device))
    assert labels.numel() == 0 and labels.dtype == torch.int64
    assert meta == {'instance_labels': [], 'instance_centers': [], 'instance_stds': [], 'instance_masks': []}
    with pytest.raises(AssertionError):
        c(torch.zeros(4, 4, dtype=torch.float64, device=cuda_device), bandwidths=None, seediness=None)
    # input on the CPU: result comes back on the CPU (clusterers.py:161)
    emb, bw, seed = make_points(seed=21, n=2000, e=4, n_free=2)
    labels, meta = c(torch.from_numpy(emb), bandwidths=torch.from_numpy(bw), seediness=torch.from_numpy(seed))
    assert labels.device.type == "cpu"
    o_labels, _ = co.sequential_cluster(emb, bw, seed, 0.5, 0.3, 0.8, 2, [0.3, 0.3])
    np.testing.assert_array_equal(labels.numpy(), o_labels)
    assert c.average_time > 0
    # broadcast bandwidths ([1,E] -> expand_as, clusterers.py:75-76)
    c0 = SequentialClustering(0.5, 0.3, 0.5, 0, [], cuda_device)
    bw1 = np.full((1, 4), 60.0, np.float32)
    labels, _ = c0(torch.from_numpy(emb).to(cuda_device), bandwidths=torch.from_numpy(bw1).to(cuda_device),
                   seediness=torch.from_numpy(seed).to(cuda_device))
    o_labels, _ = co.sequential_cluster(emb, bw1, seed, 0.5, 0.3, 0.5, 0, [])
    np.testing.assert_array_equal(labels.cpu().numpy(), o_labels)


def test_unaligned_embeddings(cuda_device):
    """A view whose data pointer is not 16-byte aligned takes the scalar-load kernel variant."""
    from stemseg_b200.clusterers import SequentialClustering
    emb, bw, seed = make_points(seed=22, n=5001, e=4, n_free=2)
    big = torch.zeros(5001 * 4 + 1, device=cuda_device)
    view = big[1:].view(5001, 4)
    view.copy_(torch.from_numpy(emb))
    c = SequentialClustering(0.5, 0.3, 0.5, 2, [0.3, 0.3], cuda_device)
    # .contiguous() keeps the (already contiguous) misaligned view
    labels, _ = c(view, bandwidths=torch.from_numpy(bw).to(cuda_device), seediness=torch.from_numpy(seed).to(cuda_device))
    o_labels, _ = co.sequential_cluster(emb, bw, seed, 0.5, 0.3, 0.5, 2, [0.3, 0.3])
    np.testing.assert_array_equal(labels.cpu().numpy(), o_labels)
