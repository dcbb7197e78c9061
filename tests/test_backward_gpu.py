"""Training backward of the B200 heads vs torch autograd over the CPU oracle.

The oracle (oracle/decoder_oracle.py) is evaluated in float64 on the CPU with autograd; the CUDA path must reproduce
d loss / d feature and d loss / d parameter for loss = <out, R> (R seeded) norm-wise within GRAD_TOL.
"""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle as do
import decoder_cases as dc

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3          # norm-wise relative error of every gradient tensor (fp32 mode: bf16x2 split operands)
FWD_TOL = 1e-4

BACKWARD_CASES = ["semseg_4_t4", "emb_xyff_t8", "emb_xytff_t16", "emb_ff_notanh_t4", "emb_xyff_t2", "seediness_t8",
                  "emb_fullwidth_t8"]


def _build_head(case, device, norm=True):
    import torch.nn as nn
    from stemseg_b200 import heads
    norm_type = (lambda c: nn.GroupNorm(32, c)) if norm else nn.Identity
    if case["kind"] == "embedding":
        head = heads.EmbeddingHead(case["in_channels"], case["inter"], case["embedding_size"], case["tanh"],
                                   case["seediness_output"], case["dim_mode"], NormType=norm_type,
                                   num_frames=case["num_frames"])
    elif case["kind"] == "seediness":
        head = heads.SeedinessHead(case["in_channels"], case["inter"], NormType=norm_type,
                                   num_frames=case["num_frames"])
    else:
        head = heads.SemsegHead(case["in_channels"], case["num_out"], case["inter"], (4, 8, 16, 32),
                                NormType=norm_type, num_frames=case["num_frames"])
    return head.to(device)


def _oracle_forward(case, sd, feats, gn_groups=32, relu_masks=None, trace=None):
    if case["kind"] == "embedding":
        return do.embedding_head(sd, feats, case["num_frames"], case["embedding_size"], case["dim_mode"], case["tanh"],
                                 case["seediness_output"], gn_groups=gn_groups, relu_masks=relu_masks, trace=trace)
    if case["kind"] == "seediness":
        return do.seediness_head(sd, feats, case["num_frames"], gn_groups=gn_groups, relu_masks=relu_masks,
                                 trace=trace)
    return do.semseg_head(sd, feats, case["num_frames"], gn_groups=gn_groups, relu_masks=relu_masks, trace=trace)


def _cuda_relu_masks(saved):
    """Sign decisions of the CUDA forward's ReLUs: fmaf(y, scale, shift) > 0 (csrc/decoder_ops.cu) per conv stage,
    as channels-first 0/1 tensors.  Evaluated in float64: the product of two fp32 is exact there, so the sign equals
    the sign of the correctly rounded fma."""
    masks = {}
    for name, stages in saved["blocks"].items():
        for j, st in enumerate(stages):
            v = st["y"].double()                                       # [n,t,h,w,c]
            if st["scale_shift"] is not None:
                ss = st["scale_shift"].double()                        # [n,c,2]
                v = v * ss[:, None, None, None, :, 0] + ss[:, None, None, None, :, 1]
            masks["%s.%d" % (name, 4 * j)] = (v > 0).permute(0, 4, 1, 2, 3).cpu()
    return masks


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(b.norm().item(), 1e-30))


def _check_case(case, sd, feats, device, norm=True):
    from stemseg_b200 import autograd as A
    head = _build_head(case, device, norm)
    head.load_state_dict(sd, strict=True)
    head.train()
    fdev = [f.to(device).requires_grad_(True) for f in feats]
    A.DEBUG_SAVED = []
    try:
        out = head(fdev)
        saved = A.DEBUG_SAVED[0]
    finally:
        A.DEBUG_SAVED = None
    assert out.requires_grad
    gen = torch.Generator().manual_seed(4242)
    r = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    (out * r.to(device=device, dtype=torch.float32)).sum().backward()
    torch.cuda.synchronize()

    # reference gradients: float64 autograd through the oracle.  A ReLU makes the gradient discontinuous in every
    # element whose pre-activation rounds across zero (one flipped element of a 73 728-element layer already moves the
    # gradient norm by ~2e-3), so the oracle is differentiated along the linear piece the CUDA forward took -- after
    # checking that the two sign patterns only disagree where the oracle's pre-activation is itself ~0.
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and v.dim() > 0}
    sd_ref = dict(sd)
    sd_ref.update(sd64)
    f64 = [f.double().clone().requires_grad_(True) for f in feats]
    masks = _cuda_relu_masks(saved)
    with torch.no_grad():
        trace = {}
        _oracle_forward(case, {k: v.detach() for k, v in sd_ref.items()}, [f.detach() for f in f64],
                        gn_groups=32 if norm else 0, trace=trace)
        flips = 0
        for key, mask in masks.items():
            y = trace[key + ".conv"]
            if norm:
                block, idx = key.split(".")
                gn = "%s.%d" % (block, int(idx) + 1)                  # the GroupNorm follows its conv in the Sequential
                y = torch.nn.functional.group_norm(y, 32, sd_ref[gn + ".weight"].detach(),
                                                   sd_ref[gn + ".bias"].detach(), eps=1e-5)
            differ = (y > 0) != mask
            flips += int(differ.sum())
            assert not differ.any() or float(y[differ].abs().max()) < 1e-4, \
                "%s: the CUDA forward's ReLU disagrees with the oracle on a clearly non-zero pre-activation" % key
        print("ReLU sign decisions that differ from the float64 oracle: %d" % flips)
    ref = _oracle_forward(case, sd_ref, f64, gn_groups=32 if norm else 0, relu_masks=masks)
    assert _rel(out.detach(), ref.detach()) <= FWD_TOL
    (ref * r).sum().backward()

    errors = {}
    for i, (fd, fr) in enumerate(zip(fdev, f64)):
        errors["feature[%d]" % i] = _rel(fd.grad, fr.grad)
    for name, p in head.named_parameters():
        assert p.grad is not None, name
        assert p.grad.shape == p.shape, name
        floor = 0.0
        if name.endswith(".bias") and name[:-5] + ".weight" in sd64:
            # a conv bias in front of a GroupNorm with one channel per group has an exactly-zero gradient: measure
            # the error of bias gradients against the scale of the same layer's weight gradient as well
            floor = 1e-2 * sd64[name[:-5] + ".weight"].grad.norm().item()
        diff = (p.grad.double().cpu() - sd64[name].grad).norm().item()
        errors[name] = diff / max(sd64[name].grad.norm().item(), floor, 1e-30)
    bad = {k: v for k, v in errors.items() if not (v <= GRAD_TOL)}
    print("max gradient error %.3g" % max(errors.values()))
    assert not bad, "gradient mismatch: %s (all: %s)" % (bad, errors)


@pytest.mark.parametrize("name", BACKWARD_CASES)
def test_head_gradients_match_autograd(name, cuda_device):
    sd, feats, case = dc.build_case(name)
    if case["n"] != 1:
        pytest.skip("backward handles one sub-clip per call")
    _check_case(case, sd, feats, cuda_device)


def test_head_gradients_without_normalisation(cuda_device):
    case = dict(kind="seediness", in_channels=32, inter=[32, 32, 32, 32], num_frames=8, n=1, h4=24, w4=24)
    shapes = do.head_parameter_shapes("seediness", 32, [32, 32, 32, 32], gn=False)
    sd = do.seeded_state_dict(shapes, 77)
    feats = do.seeded_features(78, 1, 32, 8, 24, 24)
    _check_case(case, sd, feats, cuda_device, norm=False)


def test_only_parameters_need_grad(cuda_device):
    """Frozen backbone (features without grad) still trains the head; frozen head still back-propagates to features."""
    sd, feats, case = dc.build_case("semseg_4_t4")
    head = _build_head(case, cuda_device)
    head.load_state_dict(sd, strict=True)
    out = head([f.to(cuda_device) for f in feats])
    out.sum().backward()
    assert all(p.grad is not None for p in head.parameters())
    g_ref = {n: p.grad.clone() for n, p in head.named_parameters()}
    head.zero_grad()
    for p in head.parameters():
        p.requires_grad_(False)
    fdev = [f.to(cuda_device).requires_grad_(True) for f in feats]
    head(fdev).sum().backward()
    assert all(f.grad is not None and f.grad.shape == f.shape for f in fdev)
    assert all(p.grad is None for p in head.parameters())
    # an optimiser step invalidates the packed weights (forward after the step sees the new parameters)
    for p in head.parameters():
        p.requires_grad_(True)
    opt = torch.optim.SGD(head.parameters(), lr=1e-2)
    before = head([f.to(cuda_device) for f in feats]).detach().clone()
    for n, p in head.named_parameters():
        p.grad = g_ref[n]
    opt.step()
    after = head([f.to(cuda_device) for f in feats]).detach()
    assert (before - after).abs().max().item() > 0


def test_batched_training_equals_per_sample_runs(cuda_device):
    """Batch 2 under autograd (embedding_decoder.py:101-145 accepts any N): outputs equal the per-sample forwards,
    feature gradients equal the per-sample ones, parameter gradients are their sum."""
    sd, feats, case = dc.build_case("emb_xyt_t8_batch2")
    head = _build_head(case, cuda_device)
    head.load_state_dict(sd, strict=True)
    torch.manual_seed(3)

    def run(inputs):
        for p in head.parameters():
            p.grad = None
        xs = [f.clone().to(cuda_device).requires_grad_(True) for f in inputs]
        out = head(xs)
        weight = torch.linspace(0.5, 1.5, out[0].numel(), device=cuda_device).reshape(out.shape[1:])
        (out * weight).sum().backward()
        return out.detach(), [x.grad.clone() for x in xs], {n: p.grad.clone() for n, p in head.named_parameters()}

    out_b, fg_b, pg_b = run(feats)
    assert out_b.shape[0] == 2
    total = None
    for i in range(2):
        out_i, fg_i, pg_i = run([f[i:i + 1] for f in feats])
        assert torch.equal(out_b[i:i + 1], out_i)
        for a, b in zip(fg_b, fg_i):
            assert torch.equal(a[i:i + 1], b)
        total = pg_i if total is None else {n: total[n] + pg_i[n] for n in total}
    for n in pg_b:
        scale = max(float(total[n].abs().max()), 1e-12)
        assert float((pg_b[n] - total[n]).abs().max()) <= 1e-6 * scale, n


def test_max_pool_training_fails_loudly(cuda_device):
    import torch.nn as nn
    from stemseg_b200 import heads
    head = heads.SeedinessHead(32, [32] * 4, PoolType=nn.MaxPool3d, NormType=lambda c: nn.GroupNorm(32, c),
                               num_frames=8).to(cuda_device)
    x = [torch.randn(1, 32, 8, 96 // s, 96 // s, device=cuda_device, requires_grad=True) for s in (32, 16, 8, 4)]
    with pytest.raises(NotImplementedError):
        head(x)


def test_direct_and_transposed_weight_gradients_agree(cuda_device):
    """The MN-major direct wgrad kernel against the first implementation (zero-padded transposed copies, K-major
    operands): two independent data paths to the same tensor-core GEMM, 3x3x3 and 1x1x1, ragged volume."""
    from stemseg_b200 import autograd as A, decoder as D
    torch.manual_seed(9)
    for ksize, cin, cout, t, h, w in ((3, 64, 96, 3, 10, 20), (1, 96, 32, 4, 12, 16), (3, 128, 128, 2, 24, 40)):
        x = torch.randn(1, cin, t, h, w, device=cuda_device)
        dy = torch.randn(1, t, h, w, cout, device=cuda_device)
        xp = D.pack_activation(x, 2)
        dyp = A._to_planes(dy, 2)
        outs = {}
        for mode in ("direct", "transposed"):
            A.WGRAD_MODE = mode
            try:
                dst = torch.empty(cout, cin, ksize ** 3, device=cuda_device)
                A._wgrad(dy, dyp, xp, ksize, 2, dst, 0)
                torch.cuda.synchronize()
                outs[mode] = dst.double().cpu()
            finally:
                A.WGRAD_MODE = "direct"
        ref = torch.nn.grad.conv3d_weight(x.double().cpu(), (cout, cin, ksize, ksize, ksize),
                                          dy.permute(0, 4, 1, 2, 3).double().cpu(), padding=ksize // 2)
        ref = ref.reshape(cout, cin, ksize ** 3)
        for mode, got in outs.items():
            assert float((got - ref).norm() / ref.norm()) <= 2e-5, (mode, ksize, cin, cout)
        assert float((outs["direct"] - outs["transposed"]).norm() / ref.norm()) <= 2e-5


def test_bf16_precision_gradients(cuda_device):
    """precision='bf16' (one operand plane, one tensor-core product per MAC): same kernels, bf16-level tolerance."""
    import torch.nn as nn
    from stemseg_b200 import heads
    sd, feats, case = dc.build_case("seediness_t8")
    head = heads.SeedinessHead(case["in_channels"], case["inter"], NormType=lambda c: nn.GroupNorm(32, c),
                               num_frames=case["num_frames"], precision="bf16").to(cuda_device)
    head.load_state_dict(sd, strict=True)
    fdev = [f.to(cuda_device).requires_grad_(True) for f in feats]
    out = head(fdev)
    gen = torch.Generator().manual_seed(4242)
    r = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    (out * r.to(device=cuda_device, dtype=torch.float32)).sum().backward()
    torch.cuda.synchronize()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    f64 = [f.double().clone().requires_grad_(True) for f in feats]
    ref = do.seediness_head(sd64, f64, case["num_frames"])
    (ref * r).sum().backward()
    # bf16 operands (2^-9 relative) through seven conv layers, and ReLUs that flip for near-zero pre-activations: the
    # gradients agree to a few percent -- this checks that the one-plane path is wired correctly, not its accuracy
    assert _rel(out.detach(), ref.detach()) <= 2e-2
    for fd, fr in zip(fdev, f64):
        assert _rel(fd.grad, fr.grad) <= 0.15
    for name, p in head.named_parameters():
        if name.endswith(".weight") and p.dim() == 5:
            assert _rel(p.grad, sd64[name].grad) <= 0.15, name
