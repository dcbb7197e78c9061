"""Property tests (hypothesis) of the clustering kernel against the oracle on random small point sets."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from cluster_cases import make_points
from oracle import cluster_oracle as co

pytestmark = pytest.mark.gpu


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10**6), n=st.integers(1, 3000), e=st.sampled_from([2, 3, 4, 5, 6, 7, 8, 12]),
       nf=st.integers(0, 2), primary=st.sampled_from([0.3, 0.5, 0.7, 0.9]), secondary=st.sampled_from([0.05, 0.3, 0.5]),
       min_seed=st.sampled_from([0.0, 0.4, 0.8]), max_inst=st.sampled_from([1, 3, 20]), start=st.integers(1, 50),
       quant=st.sampled_from([0, 4, 32]))
def test_random_cases_bit_exact(cuda_device, seed, n, e, nf, primary, secondary, min_seed, max_inst, start, quant):
    from stemseg_b200.clusterers import SequentialClustering
    nf = min(nf, e - 1)
    emb, bw, seedi = make_points(seed=seed, n=n, e=e, n_free=nf, n_blobs=5, quantize_seediness=quant)
    clu = dict(primary_prob_thresh=primary, secondary_prob_thresh=secondary, min_seediness_prob=min_seed,
               n_free_dims=nf, free_dim_stds=[0.3, 0.2][:nf], max_instances=max_inst, cluster_label_start=start)
    o_labels, o_meta = co.sequential_cluster(emb, bw, seedi, return_label_masks=True, **clu)
    c = SequentialClustering(primary, secondary, min_seed, nf, clu["free_dim_stds"], cuda_device,
                             max_instances=max_inst)
    labels, meta = c(torch.from_numpy(emb).to(cuda_device), bandwidths=torch.from_numpy(bw).to(cuda_device),
                     seediness=torch.from_numpy(seedi).to(cuda_device), cluster_label_start=start,
                     return_label_masks=True)
    np.testing.assert_array_equal(labels.cpu().numpy(), o_labels)
    assert meta["instance_labels"] == o_meta["instance_labels"]
    np.testing.assert_array_equal(np.array(meta["instance_centers"], np.float32).reshape(-1),
                                  np.array(o_meta["instance_centers"], np.float32).reshape(-1))
    for a, b in zip(meta["instance_masks"], o_meta["instance_masks"]):
        np.testing.assert_array_equal(a.numpy(), b)
