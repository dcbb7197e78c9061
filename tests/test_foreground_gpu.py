"""GPU parity: foreground compaction + gather vs the numpy restatement of online_chainer.py:11-22,258-281."""
import numpy as np
import pytest
import torch

from oracle import gather_oracle as go

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t,h,w,p", [(8, 120, 216, 0.3), (8, 32, 64, 0.5), (3, 7, 5, 0.9), (16, 120, 216, 0.02),
                                     (8, 480, 864, 0.35), (2, 1, 1, 1.0), (4, 33, 4097, 0.5), (5, 64, 64, 0.0)])
def test_compact_and_gather(t, h, w, p, cuda_device):
    from stemseg_b200.foreground import compact_foreground, gather_points
    rng = np.random.default_rng(t * 1000 + h)
    mask = (rng.random((t, h, w)) < p)
    if p > 0:
        mask[t - 1] = False                      # an empty frame
    e, v = 4, 2
    emb = rng.standard_normal((e, t, h, w)).astype(np.float32)
    bw = rng.random((v, t, h, w)).astype(np.float32)
    seed = rng.random((1, t, h, w)).astype(np.float32)
    o_coords, o_counts = go.masks_to_coord_list(mask)
    o_emb, o_bw, o_seed = go.gather_foreground(o_coords, emb, bw, seed)

    for mdtype in (torch.bool, torch.uint8, torch.int64):
        fg = compact_foreground(torch.from_numpy(mask).to(cuda_device).to(mdtype))
        assert fg.frame_counts == o_counts
        coords = fg.coord_list()
        for (y, x), (oy, ox) in zip(coords, o_coords):
            np.testing.assert_array_equal(y.cpu().numpy(), oy)
            np.testing.assert_array_equal(x.cpu().numpy(), ox)
    g_emb = gather_points(torch.from_numpy(emb).to(cuda_device), fg)
    g_bw = gather_points(torch.from_numpy(bw).to(cuda_device), fg)
    g_seed = gather_points(torch.from_numpy(seed).to(cuda_device), fg)
    np.testing.assert_array_equal(g_emb.cpu().numpy(), o_emb)          # pure data movement: bit-exact
    np.testing.assert_array_equal(g_bw.cpu().numpy(), o_bw)
    np.testing.assert_array_equal(g_seed.cpu().numpy(), o_seed)
    # sortedness + round trip: scatter the indices back reproduces the mask
    lin = fg.indices.long()
    assert bool((lin[1:] > lin[:-1]).all()) if lin.numel() > 1 else True
    back = torch.zeros(t * h * w, dtype=torch.bool, device=cuda_device)
    back[lin] = True
    assert torch.equal(back.view(t, h, w).cpu(), torch.from_numpy(mask))


def test_gather_strided_view(cuda_device):
    """inference_model.py:140-146 hands over channel slices of one head-output tensor (views, not copies)."""
    from stemseg_b200.foreground import compact_foreground, gather_points
    rng = np.random.default_rng(1)
    full = rng.standard_normal((7, 4, 16, 24)).astype(np.float32)
    mask = rng.random((4, 16, 24)) < 0.4
    dev = torch.from_numpy(full).to(cuda_device)
    fg = compact_foreground(torch.from_numpy(mask).to(cuda_device))
    coords, _ = go.masks_to_coord_list(mask)
    for sl in (slice(0, 4), slice(4, 6), slice(6, 7)):
        got = gather_points(dev[sl], fg)
        exp = go.gather_map(coords, full[sl])
        np.testing.assert_array_equal(got.cpu().numpy(), exp)


@pytest.mark.parametrize("factor", [1, 2, 4])
def test_upsampled_gather_matches_resized_maps(factor, cuda_device):
    """gather with the fused (1,s,s) trilinear resize == resize the maps (torch CPU), then gather."""
    from stemseg_b200.foreground import compact_foreground, gather_points
    rng = np.random.default_rng(factor)
    t, h, w = 4, 12, 20
    emb = rng.standard_normal((4, t, h, w)).astype(np.float32)
    var = rng.uniform(-1, 1, size=(2, t, h, w)).astype(np.float32)
    mask = rng.random((t, h * factor, w * factor)) < 0.3
    coords, _ = go.masks_to_coord_list(mask)
    fg = compact_foreground(torch.from_numpy(mask).to(cuda_device))
    got = gather_points(torch.from_numpy(emb).to(cuda_device), fg, upsample=factor).cpu().numpy()
    exp = go.gather_map(coords, go.resize_map(emb, factor))
    np.testing.assert_allclose(got, exp, rtol=2e-6, atol=2e-6)          # fp32 interpolation (FMA-order differences)
    got_bw = gather_points(torch.from_numpy(var).to(cuda_device), fg, transform="exp10", upsample=factor).cpu().numpy()
    exp_bw = go.gather_map(coords, go.resize_map((np.exp(var) * 10.0).astype(np.float32), factor))
    np.testing.assert_allclose(got_bw, exp_bw, rtol=5e-6, atol=1e-5)


@pytest.mark.parametrize("factor", [1, 4])
def test_frame_averager_foreground(factor, cuda_device):
    """Seediness averaged over overlapping sub-clips, (up-sampled,) thresholded inside the compaction kernel."""
    from stemseg_b200.foreground import FrameAverager
    rng = np.random.default_rng(10 + factor)
    num_frames, h, w = 12, 24, 32
    frame_lists = [list(range(0, 8)), list(range(4, 12)), [0, 0, 3, 5, 7, 9, 10, 11][:8]]
    frame_lists[2] = sorted(set(frame_lists[2]))
    planes = [rng.random((len(f), h, w)).astype(np.float32) for f in frame_lists]
    avg = FrameAverager(num_frames, h, w, cuda_device)
    for f, p in zip(frame_lists, planes):
        avg.add(f, torch.from_numpy(p).to(cuda_device))
    fg = avg.foreground_index(0.45, upsample=factor)
    mask, mean = go.averaged_foreground(frame_lists, planes, num_frames, 0.45, factor)
    got = torch.zeros(mask.size, dtype=torch.bool)
    got[fg.indices.long().cpu()] = True
    got = got.view(*mask.shape).numpy()
    ambiguous = np.abs(mean - 0.45) < 1e-6            # fp32 rounding can flip values sitting on the threshold
    assert ((got == mask) | ambiguous).all()
    assert ambiguous.sum() < 10
    assert fg.shape == mask.shape and sum(fg.frame_counts) == fg.num_points
