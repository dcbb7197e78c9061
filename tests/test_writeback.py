"""Instance-mask writeback: oracle vs the reference goldens (CPU) and the CUDA kernel vs both (GPU)."""
import os

import numpy as np
import pytest
import torch

from chain_cases import CASES, make_video
from oracle import gather_oracle as go
from oracle import writeback_oracle as wo
from writeback_cases import WRITEBACK_CASES


def load_case(name, golden_dir):
    wb = WRITEBACK_CASES[name]
    chain = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    golden = np.load(os.path.join(golden_dir, "writeback_golden.npz"))
    masks, _ = make_video(**CASES[wb["video"]])
    coords, _ = go.masks_to_coord_list(masks.astype(bool))
    labels = [chain["%s/track/%d" % (wb["video"], t)].astype(np.int64) for t in range(masks.shape[0])]
    ids = chain[wb["video"] + "/ids"].tolist()
    life = dict(zip(ids, chain[wb["video"] + "/lifetimes"].tolist()))
    order = []
    for lab in labels:
        for i in np.unique(lab).tolist():
            if i not in order:
                order.append(i)
    return wb, masks, coords, labels, {i: life[i] for i in order}, golden


@pytest.mark.parametrize("name", sorted(WRITEBACK_CASES.keys()))
def test_oracle_matches_reference_golden(name, golden_dir):
    wb, masks, coords, labels, lifetimes, golden = load_case(name, golden_dir)
    keep = wo.instances_to_keep(lifetimes, -1, wb["max_tracks"])
    assert keep == golden[name + "/keep"].tolist()
    maps = wo.id_maps(coords, labels, keep, masks.shape[1:], 4.0, wb["image_dims"], wb["min_dim"], wb["max_dim"])
    np.testing.assert_array_equal(maps, golden[name + "/maps"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(WRITEBACK_CASES.keys()))
def test_cuda_matches_reference_golden(name, golden_dir, cuda_device):
    from stemseg_b200.foreground import compact_foreground
    from stemseg_b200.writeback import instance_id_maps
    wb, masks, coords, labels, lifetimes, golden = load_case(name, golden_dir)
    t_labels = [torch.from_numpy(l).to(cuda_device) for l in labels]
    # (a) the reference's list of (y, x) index tensors
    t_coords = [(torch.from_numpy(y), torch.from_numpy(x)) for y, x in coords]
    maps, keep = instance_id_maps(t_coords, t_labels, lifetimes, masks.shape[1:], 4.0, wb["image_dims"], wb["min_dim"],
                                  wb["max_dim"], wb["max_tracks"], device=cuda_device)
    assert keep == golden[name + "/keep"].tolist()
    assert maps.dtype == torch.uint8 and maps.is_cuda
    np.testing.assert_array_equal(maps.cpu().numpy(), golden[name + "/maps"])          # integer masks: bit-exact
    # (b) a ForegroundIndex from the compaction kernel
    fg = compact_foreground(torch.from_numpy(masks).to(cuda_device))
    maps2, _ = instance_id_maps(fg, t_labels, lifetimes, masks.shape[1:], 4.0, wb["image_dims"], wb["min_dim"],
                                wb["max_dim"], wb["max_tracks"], device=cuda_device)
    assert torch.equal(maps, maps2)


@pytest.mark.gpu
def test_cuda_matches_oracle_480p(cuda_device):
    """Full-size: 8 frames of 120x216 labels -> 480x854 id maps (x4, crop 864 -> 854, no resize)."""
    from stemseg_b200.writeback import instance_id_maps
    rng = np.random.default_rng(3)
    t, h, w = 8, 120, 216
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    inst = np.zeros((t, h, w), np.int64)
    for k in range(12):
        cy, cx, r = rng.uniform(10, h - 10), rng.uniform(10, w - 10), rng.uniform(6, 25)
        for f in range(t):
            inst[f][(ys - cy - f) ** 2 + (xs - cx + 0.5 * f) ** 2 < r * r] = k + 1
    fgm = inst > 0
    coords, _ = go.masks_to_coord_list(fgm)
    labels = [inst[f][fgm[f]] for f in range(t)]
    lifetimes = {k + 1: int(rng.integers(1, 8)) for k in range(12)}
    keep = wo.instances_to_keep(lifetimes, -1, 10)
    ref = wo.id_maps(coords, labels, keep, (h, w), 4.0, (480, 854), 480, 854)
    maps, keep2 = instance_id_maps([(torch.from_numpy(y), torch.from_numpy(x)) for y, x in coords],
                                   [torch.from_numpy(l).to(cuda_device) for l in labels], lifetimes, (h, w), 4.0,
                                   (480, 854), 480, 854, 10, device=cuda_device)
    assert keep2 == keep
    np.testing.assert_array_equal(maps.cpu().numpy(), ref)
