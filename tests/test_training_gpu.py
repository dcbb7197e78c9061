"""GPU: fused SGD kernel vs torch.optim.SGD, and one full DecoderTrainer step (heads forward -> embedding loss ->
hand-written backward -> fused SGD) against the float64 oracles + the textbook update rule."""
import pytest
import torch
import torch.nn as nn

import loss_cases as lc
from oracle import decoder_oracle as do
from oracle import loss_oracle as lo
from test_backward_gpu import _cuda_relu_masks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nesterov", [True, False])
def test_fused_sgd_matches_torch_optim(nesterov, cuda_device):
    from stemseg_b200.training import FlatParameters, sgd_step
    torch.manual_seed(3)
    mod = nn.Sequential(nn.Linear(37, 19), nn.Linear(19, 5, bias=False)).to(cuda_device)      # odd sizes: padded slices
    ref = nn.Sequential(nn.Linear(37, 19), nn.Linear(19, 5, bias=False))
    ref.load_state_dict({k: v.cpu() for k, v in mod.state_dict().items()})
    opt = torch.optim.SGD(ref.parameters(), 0.05, 0.9, weight_decay=1e-2, nesterov=nesterov)
    flat = FlatParameters(mod)
    for step in range(4):
        grads = [torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * step + i))
                 for i, p in enumerate(ref.parameters())]
        flat.zero_grad()
        for p, pr, g in zip(mod.parameters(), ref.parameters(), grads):
            p.grad.copy_(2.0 * g)                 # the kernel's grad_scale = 0.5 undoes the factor (rank average)
            pr.grad = g.clone()
        sgd_step(flat, 0.05, 0.9, 1e-2, nesterov, grad_scale=0.5)
        opt.step()
        for p, pr in zip(mod.parameters(), ref.parameters()):
            torch.testing.assert_close(p.detach().cpu(), pr.detach(), rtol=2e-6, atol=1e-7)


def _small_heads(device):
    from stemseg_b200 import heads
    norm = lambda c: nn.GroupNorm(32, c)       # noqa: E731
    emb = heads.EmbeddingHead(32, [32, 32, 32, 32], 4, True, False, "xyff", NormType=norm, num_frames=4).to(device)
    seed = heads.SeedinessHead(32, [32, 32, 32, 32], NormType=norm, num_frames=4).to(device)
    emb_sd = do.seeded_state_dict(do.head_parameter_shapes("embedding", 32, [32] * 4, embedding_size=4, dim_mode="xyff",
                                                           seediness_output=False), 501)
    seed_sd = do.seeded_state_dict(do.head_parameter_shapes("seediness", 32, [32] * 4), 502)
    emb.load_state_dict(emb_sd, strict=True)
    seed.load_state_dict(seed_sd, strict=True)
    return emb, seed, emb_sd, seed_sd


@pytest.mark.parametrize("use_graph,overlap", [(False, False), (True, False), (True, True)])
def test_trainer_step_matches_oracles(use_graph, overlap, cuda_device):
    from stemseg_b200 import autograd as A
    from stemseg_b200.losses import EmbeddingLoss
    from stemseg_b200.training import DecoderTrainer
    t, h4, w4 = 4, 24, 32
    emb, seed, emb_sd, seed_sd = _small_heads(cuda_device)
    feats = do.seeded_features(503, 1, 32, t, h4, w4)
    case = lo.seeded_case(seed=504, t=t, h=h4, w=w4, embedding_size=4, n_free=2, instances=3)
    crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3], weight_variance_smoothness=10.0,
                         weight_lovasz=1.0, weight_regularization=0.001, weight_seediness=1.0, weight=1.0)
    lr, mom, wd = 0.1, 0.9, 1e-4
    trainer = DecoderTrainer({"embedding": emb, "seediness": seed}, crit, lr=lr, momentum=mom, weight_decay=wd,
                             use_graph=use_graph, overlap_heads=overlap)
    before = {("e", k): v.detach().cpu().double().clone() for k, v in emb.named_parameters()}
    before.update({("s", k): v.detach().cpu().double().clone() for k, v in seed.named_parameters()})
    fdev = [f.to(cuda_device).requires_grad_(True) for f in feats]
    targets = [{"masks": case["masks"].to(cuda_device), "ignore_masks": case["ignore"].to(cuda_device)}]
    A.DEBUG_SAVED = []
    try:
        output = trainer.step(fdev, targets)
        # graph mode runs the forward three times (warm-up, capture, replay); the captured buffers (the last two
        # entries) hold the values of the replay.  With overlapping heads the seediness head is issued first.
        saved = list(A.DEBUG_SAVED)[-2:]
        if use_graph and overlap:
            saved = saved[::-1]
    finally:
        A.DEBUG_SAVED = None
    torch.cuda.synchronize()
    assert len(saved) == 2
    feat_grads = output["feature_grads"] if use_graph else [f.grad for f in fdev]

    # float64 oracle of the same step, differentiated along the CUDA forward's ReLU pattern
    e64 = {k: (v.double().clone().requires_grad_(True) if v.dim() > 0 else v) for k, v in emb_sd.items()}
    s64 = {k: v.double().clone().requires_grad_(True) for k, v in seed_sd.items()}
    f64 = [f.double().clone().requires_grad_(True) for f in feats]
    out = torch.cat((do.embedding_head(e64, f64, t, 4, "xyff", True, False, relu_masks=_cuda_relu_masks(saved[0])),
                     do.seediness_head(s64, f64, t, relu_masks=_cuda_relu_masks(saved[1]))), dim=1)
    ref = lo.loss_from_head_output(out, case["masks"], case["ignore"], 4, 2, [0.3, 0.3], **lc.WEIGHTS)
    ref["total"].backward()
    got_total = float(output["optimization_losses"]["embedding_loss"].detach())
    assert abs(got_total - float(ref["total"])) <= 1e-4 * abs(float(ref["total"]))
    for fg, fr in zip(feat_grads, f64):
        assert tuple(fg.shape) == tuple(fr.shape)
        assert float((fg.double().cpu() - fr.grad).norm() / fr.grad.norm()) <= 2e-3

    worst = 0.0
    for tag, mod, sd64 in (("e", emb, e64), ("s", seed, s64)):
        for name, p in mod.named_parameters():
            p0 = before[(tag, name)]
            g = sd64[name].grad + wd * p0                       # torch.optim.SGD, first step: buf = g
            want = -lr * (g + mom * g)                          # nesterov
            delta = p.detach().cpu().double() - p0
            # a conv bias in front of a one-channel-per-group GroupNorm has a ~zero gradient: scale by the weight's update
            floor = 0.0
            if name.endswith(".bias") and name[:-5] + ".weight" in sd64 and sd64[name[:-5] + ".weight"].dim() == 5:
                gw = sd64[name[:-5] + ".weight"].grad
                floor = 1e-2 * lr * float(gw.norm())
            err = float((delta - want).norm()) / max(float(want.norm()), floor, 1e-30)
            worst = max(worst, err)
            assert err <= 3e-3, (tag, name, err)
    print("worst parameter-update error %.3g" % worst)

    # the optimiser wrote the flat buffers behind autograd's back: the next forward must see the new weights
    with torch.no_grad():
        got = emb([f.detach() for f in fdev]).cpu()
    new_sd = {k: v.detach().cpu() for k, v in emb.state_dict().items()}
    want = do.embedding_head(new_sd, feats, t, 4, "xyff", True, False)
    stale = do.embedding_head(emb_sd, feats, t, 4, "xyff", True, False)
    assert float((got - want).abs().max() / want.abs().max()) <= 1e-4
    assert float((stale - want).abs().max()) > 10 * float((got - want).abs().max())


def test_graph_steps_track_the_autograd_steps(cuda_device):
    """Three consecutive steps: the CUDA-graph trainer and the autograd trainer follow the same parameter trajectory
    (same kernels, different launch mechanism), including the momentum state and the per-step weight repack."""
    from stemseg_b200.losses import EmbeddingLoss
    from stemseg_b200.training import DecoderTrainer
    t, h4, w4 = 4, 24, 32
    feats = do.seeded_features(603, 1, 32, t, h4, w4)
    case = lo.seeded_case(seed=604, t=t, h=h4, w=w4, embedding_size=4, n_free=2, instances=2)
    results = []
    for use_graph, overlap in ((False, False), (True, False), (True, True)):
        emb, seed, _, _ = _small_heads(cuda_device)
        crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3],
                             weight_variance_smoothness=10.0, weight_lovasz=1.0, weight_regularization=0.001,
                             weight_seediness=1.0, weight=1.0)
        trainer = DecoderTrainer({"embedding": emb, "seediness": seed}, crit, lr=0.05, use_graph=use_graph,
                                 overlap_heads=overlap)
        targets = [{"masks": case["masks"].to(cuda_device), "ignore_masks": case["ignore"].to(cuda_device)}]
        losses = []
        for step in range(3):
            fdev = [(f * (1.0 + 0.1 * step)).to(cuda_device).requires_grad_(True) for f in feats]
            out = trainer.step(fdev, targets)
            losses.append(float(out["optimization_losses"]["embedding_loss"].detach()))
        torch.cuda.synchronize()
        results.append((losses, torch.cat([f.data.clone() for f in trainer.flats]).cpu()))
    (l_a, p_a) = results[0]
    assert l_a[0] != l_a[2]
    for l_g, p_g in results[1:]:
        for a, g in zip(l_a, l_g):
            assert abs(a - g) <= 1e-5 * abs(a)
        assert float((p_a - p_g).norm() / p_a.norm()) <= 1e-6


def test_split_backward_and_lr_schedule_under_graphs(cuda_device):
    """(1) The data-parallel schedule (backward cut in two graphs so that the all-reduce of the block_32x / block_16x
    gradients overlaps the long block_8x / block_4x backward) follows the same trajectory as the single-graph step.
    (2) set_lr() between steps takes effect in graph mode (the captured SGD launches read the hyper-parameters from
    device memory): a schedule lr_k = lr0 * 0.5^k matches the autograd trainer driven with the same schedule,
    and differs from a constant-lr run."""
    from stemseg_b200.losses import EmbeddingLoss
    from stemseg_b200.training import DecoderTrainer
    t, h4, w4 = 4, 24, 32
    feats = do.seeded_features(703, 1, 32, t, h4, w4)
    case = lo.seeded_case(seed=704, t=t, h=h4, w=w4, embedding_size=4, n_free=2, instances=2)

    def run(use_graph, split, schedule):
        emb, seed, _, _ = _small_heads(cuda_device)
        crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3],
                             weight_variance_smoothness=10.0, weight_lovasz=1.0, weight_regularization=0.001,
                             weight_seediness=1.0, weight=1.0)
        trainer = DecoderTrainer({"embedding": emb, "seediness": seed}, crit, lr=0.05, use_graph=use_graph)
        trainer.split_backward = split
        for flat in trainer.flats:
            assert 0 < flat.prefix_end < flat.numel and flat.prefix_end % 4 == 0
        targets = [{"masks": case["masks"].to(cuda_device), "ignore_masks": case["ignore"].to(cuda_device)}]
        for step in range(3):
            if schedule:
                trainer.set_lr(0.05 * 0.5 ** step)
            fdev = [(f * (1.0 + 0.1 * step)).to(cuda_device).requires_grad_(True) for f in feats]
            trainer.step(fdev, targets)
        torch.cuda.synchronize()
        return torch.cat([f.data.clone() for f in trainer.flats]).cpu()

    ref_sched = run(False, False, True)
    for split in (False, True):
        got = run(True, split, True)
        assert float((ref_sched - got).norm() / ref_sched.norm()) <= 1e-6, split
    const = run(True, True, False)
    assert float((ref_sched - const).norm() / ref_sched.norm()) > 1e-4


def test_semseg_config_trainer_step(cuda_device):
    """YouTube-VIS / KITTI-MOTS wiring: embedding head with its own seediness channel + semseg head with a foreground
    channel.  The CUDA-graph step (three... two heads on two streams) and the autograd step must produce the same
    losses and the same parameter update; the losses must match the float64 oracles."""
    import torch.nn as nn
    from stemseg_b200 import heads
    from stemseg_b200.losses import EmbeddingLoss
    from stemseg_b200.training import DecoderTrainer
    t, h4, w4 = 4, 24, 32
    feats = do.seeded_features(703, 1, 32, t, h4, w4)
    case = lo.seeded_case(seed=704, t=t, h=h4, w=w4, embedding_size=4, n_free=2, instances=2)
    sem_case = lo.seeded_semseg_case(seed=705, t=t, h=h4, w=w4, num_classes=5)
    emb_sd = do.seeded_state_dict(do.head_parameter_shapes("embedding", 32, [32] * 4, embedding_size=4, dim_mode="xyff",
                                                           seediness_output=True), 706)
    sem_sd = do.seeded_state_dict(do.head_parameter_shapes("semseg", 32, [64] * 4, num_out=6), 707)
    norm = lambda c: nn.GroupNorm(32, c)       # noqa: E731
    results = []
    for use_graph in (False, True):
        emb = heads.EmbeddingHead(32, [32] * 4, 4, True, True, "xyff", NormType=norm, num_frames=t).to(cuda_device)
        sem = heads.SemsegHead(32, 5, [64] * 4, (4, 8, 16, 32), foreground_channel=True, NormType=norm,
                               num_frames=t).to(cuda_device)
        emb.load_state_dict(emb_sd, strict=True)
        sem.load_state_dict(sem_sd, strict=True)
        crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3],
                             weight_variance_smoothness=10.0, weight_lovasz=1.0, weight_regularization=0.001,
                             weight_seediness=1.0, weight=1.0)
        trainer = DecoderTrainer({"embedding": emb, "semseg": sem}, crit, lr=0.05, use_graph=use_graph,
                                 weight_semseg=0.7)
        targets = [{"masks": case["masks"].to(cuda_device), "ignore_masks": case["ignore"].to(cuda_device),
                    "semseg_masks": sem_case["semseg_masks"].to(cuda_device)}]
        fdev = [f.to(cuda_device).requires_grad_(True) for f in feats]
        out = trainer.step(fdev, targets)
        torch.cuda.synchronize()
        ol = out["optimization_losses"]
        results.append(({k: float(v.detach()) for k, v in ol.items()},
                        torch.cat([f.data.clone() for f in trainer.flats]).cpu()))
    (l_a, p_a), (l_g, p_g) = results
    assert set(l_a) == set(l_g) == {"embedding_loss", "semantic_segmentation_loss", "foreground"}
    for k in l_a:
        assert abs(l_a[k] - l_g[k]) <= 1e-5 * abs(l_a[k]), k
    assert float((p_a - p_g).norm() / p_a.norm()) <= 1e-6
    # losses against the float64 oracles on the oracle's own forward (tolerance of the fp32-parity forward)
    e64 = {k: v.double() for k, v in emb_sd.items()}
    s64 = {k: v.double() for k, v in sem_sd.items()}
    f64 = [f.double() for f in feats]
    emb_out = do.embedding_head(e64, f64, t, 4, "xyff", True, True)
    sem_out = do.semseg_head(s64, f64[::-1], t)
    ref_e = lo.loss_from_head_output(emb_out, case["masks"], case["ignore"], 4, 2, [0.3, 0.3], **lc.WEIGHTS)
    ref_s = lo.semseg_losses_sequence(sem_out[0], sem_case["semseg_masks"], case["ignore"])
    assert abs(l_g["embedding_loss"] - float(ref_e["total"])) <= 2e-4 * abs(float(ref_e["total"]))
    assert abs(l_g["semantic_segmentation_loss"] - 0.7 * float(ref_s["semseg"])) <= 2e-4 * float(ref_s["semseg"])
    assert abs(l_g["foreground"] - float(ref_s["foreground"])) <= 2e-4 * float(ref_s["foreground"])
