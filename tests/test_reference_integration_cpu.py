"""Build-container-only check (skipped where /root/reference is absent, e.g. on the GPU box): the B200 heads and
clusterer drop into the UNMODIFIED reference through its own registries, and a reference state_dict loads strictly."""
import os
import sys

import pytest

REF = os.environ.get("STEMSEG_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "stemseg")), reason="reference tree not present")


def test_install_into_reference_and_strict_state_dict():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import _refshim
    _refshim.install()
    from functools import partial
    import torch.nn as nn
    from stemseg.config import cfg
    from stemseg.modeling.embedding_decoder import EMBEDDING_HEAD_REGISTRY, SqueezingExpandDecoder
    import stemseg_b200.registry as b200
    from stemseg_b200 import heads
    cfg.INPUT.update_param("NUM_FRAMES", 8)
    ref_head = SqueezingExpandDecoder(64, [64, 64, 32, 32], 4, tanh_activation=True, seediness_output=False,
                                      experimental_dims="xyff", PoolType=nn.AvgPool3d,
                                      NormType=partial(nn.GroupNorm, 32))
    b200.install_into_reference()
    cls = EMBEDDING_HEAD_REGISTRY["squeeze_expand_decoder"]
    assert cls is heads.EmbeddingHead and EMBEDDING_HEAD_REGISTRY[b200.B200_KEY] is heads.EmbeddingHead
    mine = cls(64, [64, 64, 32, 32], 4, tanh_activation=True, seediness_output=False, experimental_dims="xyff",
               PoolType=nn.AvgPool3d, NormType=partial(nn.GroupNorm, 32))          # NUM_FRAMES comes from cfg
    assert mine.num_frames == 8
    result = mine.load_state_dict(ref_head.state_dict(), strict=True)
    assert not result.missing_keys and not result.unexpected_keys
    assert (mine.embedding_size, mine.variance_channels, mine.seediness_channels) == \
           (ref_head.embedding_size, ref_head.variance_channels, ref_head.seediness_channels)
    import stemseg.inference.clusterers as ref_clusterers
    from stemseg_b200.clusterers import SequentialClustering
    assert ref_clusterers.SequentialClustering is SequentialClustering
    import stemseg.modeling.model_builder as ref_builder
    from stemseg_b200.losses import EmbeddingLoss
    assert ref_builder.EmbeddingLoss is EmbeddingLoss
    # constructed exactly as model_builder.py:294-298 does (upper-case cfg keys)
    crit = ref_builder.EmbeddingLoss(min(cfg.MODEL.EMBEDDINGS.SCALE), embedding_size=cfg.MODEL.EMBEDDINGS.EMBEDDING_SIZE,
                                     nbr_free_dims=len(cfg.TRAINING.LOSSES.EMBEDDING.FREE_DIM_STDS),
                                     **cfg.TRAINING.LOSSES.EMBEDDING.d())
    assert crit.num_input_channels == 2 * cfg.MODEL.EMBEDDINGS.EMBEDDING_SIZE - crit.n_free_dims + 1
    from stemseg_b200.losses import CrossEntropyLoss
    assert ref_builder.SEMSEG_LOSS_REGISTRY[cfg.TRAINING.LOSSES.SEMSEG] is CrossEntropyLoss     # model_builder.py:334
    assert ref_builder.TrainingModel.compute_fg_loss.__name__ == "<lambda>"


def test_same_seed_gives_identical_init():
    """Sub-modules are created in the reference's order, so build_model's manual_seed(42) initialises identically."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import _refshim
    _refshim.install()
    from functools import partial
    import torch
    import torch.nn as nn
    from stemseg.config import cfg
    import stemseg.modeling.seediness_decoder as ref_sd
    from stemseg_b200 import heads
    cfg.INPUT.update_param("NUM_FRAMES", 8)
    ref_cls = [v for k, v in vars(ref_sd).items() if k == "SqueezingExpandDecoder"][0]
    torch.manual_seed(42)
    a = ref_cls(32, [32, 32, 32, 32], PoolType=nn.AvgPool3d, NormType=partial(nn.GroupNorm, 32))
    torch.manual_seed(42)
    b = heads.SeedinessHead(32, [32, 32, 32, 32], PoolType=nn.AvgPool3d, NormType=partial(nn.GroupNorm, 32),
                            num_frames=8)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
