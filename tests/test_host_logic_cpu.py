"""Host-side logic that needs no GPU: graph-cache bound, operand-format selection, first-stage grouping, the flat
gradient prefix used by the overlapped all-reduce, the reference-mirroring window generator."""
import pytest
import torch
import torch.nn as nn


def test_lru_cache_evicts_oldest_and_refreshes_on_get():
    from stemseg_b200._lib import LRUCache
    c = LRUCache(2)
    c.put("a", 1)
    c.put("b", 2)
    assert c.get("a") == 1            # refreshes "a"
    c.put("c", 3)                     # evicts "b"
    assert "b" not in c and c.get("b") is None and len(c) == 2
    assert c.get("a") == 1 and c.get("c") == 3
    c.clear()
    assert len(c) == 0


def test_operand_format_selection():
    from stemseg_b200 import decoder as D
    assert D.FP32_FAST_BLOCKS == ()                                   # parity mode: three products everywhere
    assert D.block_planes(2, "block_8x") == 2 and D.block_planes(1, "block_8x") == 1
    D.set_fast_blocks(("block_8x", "block_16x"))
    try:
        assert D.block_planes(2, "block_8x") == D.PLANES_FP16 and D.block_planes(2, "block_16x") == D.PLANES_FP16
        assert D.block_planes(2, "block_4x") == 2 and D.block_planes(2, "block_32x") == 2
        assert D.block_planes(2, "block_8x", exact=True) == 2         # training / max-pool heads stay exact
        assert D.block_planes(1, "block_8x") == 1                     # bf16 mode is unaffected
    finally:
        D.set_fast_blocks(())
    assert D.plane_count(D.PLANES_FP16) == 1 and D.plane_count(2) == 2


def test_first_stage_grouping_keeps_wide_tiles():
    from stemseg_b200.decoder import first_stage_groups
    assert first_stage_groups([128, 128]) == [[0, 1]]                 # DAVIS: one N=256 GEMM
    assert first_stage_groups([128, 256]) == [[1], [0]]               # YouTube-VIS: N=256 semseg + N=128 embedding
    assert first_stage_groups([256, 256]) == [[0, 1]]


def test_flat_parameters_prefix_is_the_low_resolution_blocks():
    from stemseg_b200 import heads
    from stemseg_b200.training import FlatParameters
    head = heads.SeedinessHead(32, [32, 32, 32, 32], NormType=lambda c: nn.GroupNorm(32, c), num_frames=8)
    flat = FlatParameters(head)
    names = [n for n, _ in head.named_parameters()]
    first_late = names.index("block_8x.0.weight")
    assert flat.prefix_end == flat.offsets[first_late] and 0 < flat.prefix_end < flat.numel
    assert all(n.startswith(("block_32x.", "block_16x.")) for n in names[:first_late])
    assert not any(n.startswith(("block_32x.", "block_16x.")) for n in names[first_late:])
    # parameters are views into the flat buffer
    p0 = next(head.parameters())
    assert p0.data_ptr() == flat.data.data_ptr()


def test_heads_accept_max_pool_and_reject_other_poolers():
    from stemseg_b200 import heads
    h = heads.SeedinessHead(32, [32] * 4, PoolType=nn.MaxPool3d, num_frames=8)
    assert isinstance(h.block_32x[3], nn.MaxPool3d)                  # same module layout as the reference
    with pytest.raises(NotImplementedError):
        heads.SeedinessHead(32, [32] * 4, PoolType=nn.AdaptiveAvgPool3d, num_frames=8)


def test_clip_parallel_needs_video_masks_for_several_subclips():
    from stemseg_b200.parallel import clip_parallel_process
    with pytest.raises(ValueError, match="foreground masks"):
        clip_parallel_process(object(), None, [[0, 1, 2, 3], [2, 3, 4, 5]], lambda i: None)
