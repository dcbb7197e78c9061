"""CPU tests: the clustering oracle against the golden vectors produced by the reference (tests/golden)."""
import math
import os

import numpy as np
import pytest

from cluster_cases import case_table, make_points
from oracle import cluster_oracle as co


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "cluster_golden.npz"))


@pytest.mark.parametrize("name", sorted(case_table().keys()))
def test_oracle_matches_reference_golden(name, golden):
    pts, clu = case_table()[name]
    emb, bw, seed = make_points(**pts)
    labels, meta = co.sequential_cluster(emb, bw, seed, return_label_masks=True, **clu)
    assert labels.dtype == np.int64
    np.testing.assert_array_equal(labels, golden[name + "/labels"].astype(np.int64))
    np.testing.assert_array_equal(np.array(meta['instance_labels'], np.int64), golden[name + "/instance_labels"])
    k, e = len(meta['instance_labels']), emb.shape[1]
    np.testing.assert_array_equal(np.array(meta['instance_centers'], np.float32).reshape(k, e),
                                  golden[name + "/instance_centers"])
    np.testing.assert_allclose(np.array(meta['instance_stds'], np.float32).reshape(k, e),
                               golden[name + "/instance_stds"], rtol=3e-7, atol=0)
    np.testing.assert_array_equal(np.array([int(m.sum()) for m in meta['instance_masks']], np.int64),
                                  golden[name + "/mask_counts"])


def test_empty_input():
    labels, meta = co.sequential_cluster(np.zeros((0, 4), np.float32), np.zeros((0, 2), np.float32),
                                         np.zeros((0, 1), np.float32), 0.5, 0.3, 0.8, 2, [0.3, 0.3])
    assert labels.shape == (0,) and meta['instance_labels'] == []


def test_label_offset_invariance():
    emb, bw, seed = make_points(seed=5, n=3000, e=4, n_free=2)
    a, _ = co.sequential_cluster(emb, bw, seed, 0.5, 0.3, 0.5, 2, [0.3, 0.3], cluster_label_start=1)
    b, _ = co.sequential_cluster(emb, bw, seed, 0.5, 0.3, 0.5, 2, [0.3, 0.3], cluster_label_start=41)
    np.testing.assert_array_equal(np.where(a < 0, a, a + 40), b)


@pytest.mark.parametrize("p", [0.5, 0.3, 0.8, 0.95, 0.05, 1e-30, 0.999])
def test_threshold_distance_is_the_boundary(p):
    d = co.prob_threshold_to_distance(p)
    nxt = np.nextafter(d, np.float32(np.inf))
    p32 = np.float32(p)
    assert np.float32(math.exp(float(np.float32(-0.5) * d))) > p32
    assert not (np.float32(math.exp(float(np.float32(-0.5) * nxt))) > p32)


def test_threshold_distance_edges():
    assert co.prob_threshold_to_distance(1.0) == -1.0
    assert co.prob_threshold_to_distance(2.0) == -1.0
    assert np.isinf(co.prob_threshold_to_distance(-0.5))


def test_aten_sum_order_small_cases():
    # E < 8: 4 partial sums, tail into partial 0 (pinned against torch in gen_cluster_golden.py)
    t = np.array([[1e8, 1.0, -1e8, 1.0, 1.0]], np.float32)
    # ((((t0 + t4) + t1) + t2) + t3)
    expect = np.float32(np.float32(np.float32(np.float32(t[0, 0] + t[0, 4]) + t[0, 1]) + t[0, 2]) + t[0, 3])
    assert co.aten_inner_sum_f32(t)[0] == expect
