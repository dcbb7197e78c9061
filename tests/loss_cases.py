"""Seeded embedding-loss parity cases shared by the golden generator (reference, build container) and the tests."""
FREE_DIM_STDS = {0: [], 1: [0.5], 2: [0.3, 0.3], 3: [0.3, 0.3, 0.5]}
WEIGHTS = dict(w_lovasz=1.0, w_variance_smoothness=10.0, w_seediness=1.0, w=1.0)      # defaults.yaml:36-40


def case_table():
    c = {}
    c["xyff_3inst"] = dict(seed=11, t=4, h=24, w=32, embedding_size=4, n_free=2, instances=3)
    c["xyt_2inst"] = dict(seed=12, t=4, h=16, w=24, embedding_size=3, n_free=0, instances=2)
    c["xytff_1inst"] = dict(seed=13, t=2, h=24, w=24, embedding_size=5, n_free=2, instances=1)
    c["empty_first"] = dict(seed=14, t=4, h=24, w=32, embedding_size=4, n_free=2, instances=3, empty_instances=(0,))
    c["empty_last"] = dict(seed=15, t=4, h=24, w=32, embedding_size=4, n_free=2, instances=3, empty_instances=(2,))
    c["overlapping"] = dict(seed=16, t=4, h=24, w=32, embedding_size=4, n_free=2, instances=3, overlap=True)
    c["ragged_8x48x80"] = dict(seed=17, t=8, h=48, w=80, embedding_size=4, n_free=2, instances=4)
    c["half_ignored"] = dict(seed=18, t=4, h=24, w=32, embedding_size=4, n_free=2, instances=2, ignore_frac=0.5)
    c["no_points"] = dict(seed=19, t=2, h=16, w=16, embedding_size=4, n_free=2, instances=2, empty_instances=(0, 1))
    c["six_instances"] = dict(seed=20, t=8, h=32, w=40, embedding_size=4, n_free=2, instances=6)
    return c


def build_case(name):
    from oracle import loss_oracle as lo
    return lo.seeded_case(**case_table()[name])


def run_oracle(name, dtype=None):
    """-> (loss dict, gradient of the total loss wrt the head output [1,C,T,H,W])."""
    import torch
    from oracle import loss_oracle as lo
    case = build_case(name)
    out = case["out"].clone()
    if dtype is not None:
        out = out.to(dtype)
    out.requires_grad_(True)
    losses = lo.loss_from_head_output(out, case["masks"], case["ignore"], case["embedding_size"], case["n_free"],
                                      FREE_DIM_STDS[case["n_free"]], **WEIGHTS)
    if losses["total"].requires_grad:
        losses["total"].backward()
    grad = out.grad if out.grad is not None else torch.zeros_like(out)
    return {k: v.detach() for k, v in losses.items()}, grad


def semseg_case_table():
    c = {}
    c["kitti_3cls_fg"] = dict(seed=31, t=4, h=24, w=32, num_classes=3)
    c["ytvis_41cls_fg"] = dict(seed=32, t=2, h=16, w=24, num_classes=41)
    c["five_cls_no_fg"] = dict(seed=33, t=4, h=12, w=20, num_classes=5, foreground_channel=False)
    c["mostly_ignored"] = dict(seed=34, t=2, h=16, w=16, num_classes=4, ignore_frac=0.9)
    c["ragged_8x30x50"] = dict(seed=35, t=8, h=30, w=50, num_classes=7, ignore_frac=0.0)
    return c


def build_semseg_case(name):
    from oracle import loss_oracle as lo
    return lo.seeded_semseg_case(**semseg_case_table()[name])


def run_semseg_oracle(name, dtype=None):
    """-> (loss dict, gradient of (semseg + foreground) wrt the head output [1,C,T,H,W])."""
    from oracle import loss_oracle as lo
    case = build_semseg_case(name)
    out = case["out"].clone()
    if dtype is not None:
        out = out.to(dtype)
    out.requires_grad_(True)
    losses = lo.semseg_losses_sequence(out[0], case["semseg_masks"], case["ignore"], case["foreground_channel"])
    total = losses["semseg"] if losses["foreground"] is None else losses["semseg"] + losses["foreground"]
    total.backward()
    return {k: (None if v is None else v.detach()) for k, v in losses.items()}, out.grad
