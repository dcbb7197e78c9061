"""GPU: the sub-clip pipeline wired like the three shipped configs (DAVIS / YouTube-VIS / KITTI-MOTS): heads checked
against the decoder oracle (1e-4 norm-wise), foreground + clustering bit-exact against the gather / cluster oracles on
the device-produced maps, graph path == eager path."""
import numpy as np
import pytest
import torch

from oracle import cluster_oracle as co
from oracle import decoder_oracle as do
from oracle import gather_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_heads(pipe, feats_list, t):
    e = pipe.embedding_head
    sd = {k: v.detach().cpu() for k, v in e.state_dict().items()}
    out = do.embedding_head(sd, feats_list, t, e.embedding_size, e.embedding_dim_mode, e.tanh_activation,
                            e.seediness_channels == 1)[0]
    ev = e.embedding_size + e.variance_channels
    seed = out[ev:ev + 1]
    if pipe.seediness_head is not None:
        ssd = {k: v.detach().cpu() for k, v in pipe.seediness_head.state_dict().items()}
        seed = do.seediness_head(ssd, feats_list, t)[0]
    semseg = None
    if pipe.semseg_head is not None:
        msd = {k: v.detach().cpu() for k, v in pipe.semseg_head.state_dict().items()}
        semseg = do.semseg_head(msd, feats_list[::-1], t)[0]
    return out[:e.embedding_size], out[e.embedding_size:ev], seed, semseg


# (config, semseg widths): embedding 32 + semseg 256 = 288 output channels in the fused first-stage GEMM exercises the
# grouped launches (N=256 for the semseg head, N=32 for the embedding head) that production YouTube-VIS widths take
@pytest.mark.parametrize("config,semseg_widths", [("davis", 64), ("youtube_vis", 64), ("kitti_mots", 64),
                                                  ("youtube_vis", 256)])
def test_shipped_config_pipeline(config, semseg_widths, cuda_device):
    from stemseg_b200.foreground import gather_points
    from stemseg_b200.pipeline import build_pipeline
    t, h4, w4 = 8, 24, 32
    pipe = build_pipeline(config, cuda_device, num_frames=t, in_channels=32, inter_channels=(32, 32, 32, 32),
                          semseg_inter_channels=(semseg_widths,) * 4, num_classes=5, min_seediness_prob=0.0)
    if pipe.semseg_head is not None:
        n_groups = len(pipe._head_group().first_stage["block_4x"])
        assert n_groups == (2 if semseg_widths == 256 else 1)
    feats_list = do.seeded_features(900 + len(config), 1, 32, t, h4, w4)
    feats = {s: f.to(cuda_device) for s, f in zip((32, 16, 8, 4), feats_list)}
    if pipe.semseg_head is None:
        # random-init seediness sits in a narrow band around 0.5: put the foreground threshold at its median so that
        # the compaction sees a non-trivial mask (the threshold is baked into the step graph: set it before capture)
        pipe.seediness_fg_threshold = float(pipe.run_heads(feats)[2].median())
    res = pipe(feats)                                  # whole-step CUDA graph
    torch.cuda.synchronize()
    emb, var, seed, semseg = _oracle_heads(pipe, feats_list, t)
    pairs = [("embedding", res.embeddings, emb), ("variance", res.variances, var), ("seediness", res.seediness, seed)]
    if semseg is not None:
        pairs.append(("semseg", res.semseg_logits, semseg))
    else:
        assert res.semseg_logits is None
    for name, a, b in pairs:
        assert tuple(a.shape) == tuple(b.shape), name
        err = (a.cpu().double() - b.double()).abs().max().item() / b.abs().max().item()
        assert err <= TOL, "%s: norm-wise error %.3e" % (name, err)

    # foreground: semseg foreground logit > 0 when the config has a semseg head with a foreground channel
    # (inference_model.py:212-225), else seediness > 0.25 (inference/main.py:93-103) -- on the DEVICE maps
    if semseg is not None:
        mask = (res.semseg_logits[-1] > 0).cpu().numpy()
    else:
        mask = (res.seediness[0] > pipe.seediness_fg_threshold).cpu().numpy()
    coords, counts = go.masks_to_coord_list(mask)
    assert list(res.fg_index.frame_counts) == list(counts)
    assert 0 < res.fg_index.num_points < mask.size, "degenerate foreground: the case does not test the compaction"
    e_np = gather_points(res.embeddings, res.fg_index).cpu().numpy()
    b_np = gather_points(res.variances, res.fg_index, transform="exp10").cpu().numpy()
    s_np = gather_points(res.seediness, res.fg_index).cpu().numpy()
    ge, gb, gs = go.gather_foreground(coords, res.embeddings.cpu().numpy(), np.exp(res.variances.cpu().numpy()) * 10.0,
                                      res.seediness.cpu().numpy())
    np.testing.assert_array_equal(e_np, ge)
    np.testing.assert_array_equal(s_np.reshape(-1), gs.reshape(-1))
    n_free = pipe.clusterer.n_free_dims
    o_labels, o_meta = co.sequential_cluster(e_np, b_np, s_np, 0.5, 0.3, 0.0, n_free, list(pipe.clusterer.free_dim_stds))
    np.testing.assert_array_equal(res.labels.cpu().numpy(), o_labels)
    assert res.meta["instance_labels"] == o_meta["instance_labels"]

    # eager path gives the same labels
    pipe.use_step_graph = False
    res2 = pipe(feats)
    np.testing.assert_array_equal(res2.labels.cpu().numpy(), res.labels.cpu().numpy())
