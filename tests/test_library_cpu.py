"""CPU tests of the boundary: the C-ABI library loads here (no GPU) and exports exactly what include/*.h declares."""
import os
import re

import numpy as np
import pytest

from oracle import cluster_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from stemseg_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "stemseg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stemseg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from stemseg_b200 import _lib
    names = declared_symbols()
    assert names, "no declarations parsed"
    assert sorted(_lib.PROTOTYPES.keys()) == names
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n


def test_abi_version(lib):
    from stemseg_b200 import _lib
    assert lib.stemseg_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("p", [0.5, 0.3, 0.8, 0.95, 0.0, 1.0, -1.0, 1e-30, 0.999999, 0.25, 2.0])
def test_threshold_helper_matches_oracle(lib, p):
    """Host helper (C, libm exp) vs the oracle's independent implementation (python math.exp)."""
    got = np.float32(lib.stemseg_prob_threshold_to_distance(p))
    exp = co.prob_threshold_to_distance(p)
    assert got == exp or (np.isinf(got) and np.isinf(exp))


def test_invalid_arguments_are_reported(lib):
    from stemseg_b200 import _lib
    params = _lib.StemsegClusterParams()
    params.n_points = 0
    params.embedding_dims = 4
    rc = lib.stemseg_seq_cluster(None, None, None, params, None, None, None, None, None, 0, None)
    assert rc == -1
    assert b"n_points" in lib.stemseg_last_error()
    params.max_instances = 1000
    ws = _lib.c_size_t(0)
    assert lib.stemseg_seq_cluster_workspace_bytes(params, ws) == -1


def test_cpu_device_is_rejected():
    from stemseg_b200.clusterers import SequentialClustering
    with pytest.raises(ValueError):
        SequentialClustering(0.5, 0.3, 0.8, 2, [0.3, 0.3], "cpu")


def test_first_stage_grouping_rule():
    """Host logic of the launch plan: which heads share one first-stage GEMM (stemseg_b200/decoder.py)."""
    from stemseg_b200.decoder import first_stage_groups
    assert first_stage_groups([128, 128]) == [[0, 1]]                 # DAVIS 4x / 8x: N = 256
    assert first_stage_groups([256, 256]) == [[0, 1]]                 # every config at 32x / 16x: N = 512 -> 2 x 256
    assert first_stage_groups([128, 256]) == [[1], [0]]               # YouTube-VIS 4x / 8x: 384 would need N = 128 tiles
    assert first_stage_groups([128, 128, 256]) == [[0, 1, 2]]         # three heads, 512
    assert first_stage_groups([32, 64]) == [[0, 1]]                   # small test widths stay fused
    assert first_stage_groups([96, 96, 96]) == [[0], [1], [2]]
    for couts in ([128, 256, 256], [64], [256, 128, 128, 256]):
        flat = sorted(i for g in first_stage_groups(couts) for i in g)
        assert flat == list(range(len(couts)))
