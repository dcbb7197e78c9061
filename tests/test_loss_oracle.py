"""CPU: the embedding-loss oracle against the fixtures written by the reference (tests/golden/gen_loss_golden.py)."""
import os

import numpy as np
import pytest
import torch

import loss_cases as lc


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "loss_golden.npz"))


@pytest.mark.parametrize("name", sorted(lc.case_table()))
def test_oracle_reproduces_reference(name, golden):
    losses, grad = lc.run_oracle(name)
    for key in ("total", "lovasz", "variance_smoothness", "seediness"):
        ref = float(golden["%s/%s" % (name, key)])
        got = float(losses[key])
        # 1e-6 relative, not bit-equal: the reference orders an instance's points by an unstable argsort, so its own
        # fp32 means depend on that permutation (see the generator)
        assert abs(got - ref) <= 1e-6 * max(abs(ref), 1e-3), (key, got, ref)
    ref_grad = torch.from_numpy(golden["%s/grad" % name])
    assert grad.shape == ref_grad.shape
    denom = max(float(ref_grad.norm()), 1e-30)
    assert float((grad - ref_grad).norm()) / denom <= 1e-5 or float(ref_grad.norm()) == 0.0


@pytest.mark.parametrize("name", ["xyff_3inst", "empty_first", "overlapping"])
def test_float64_oracle_agrees_with_float32_reference(name, golden):
    """The GPU tests compare against the oracle evaluated in float64; it must agree with the fp32 reference fixtures."""
    losses, grad = lc.run_oracle(name, dtype=torch.float64)
    for key in ("total", "lovasz", "variance_smoothness", "seediness"):
        ref = float(golden["%s/%s" % (name, key)])
        assert abs(float(losses[key]) - ref) <= 1e-5 * max(abs(ref), 1e-3)
    ref_grad = torch.from_numpy(golden["%s/grad" % name]).double()
    assert float((grad - ref_grad).norm() / ref_grad.norm()) <= 1e-4


def test_empty_first_instance_shifts_targets(golden):
    """Quirk (i): with an empty instance in front the kept instances are scored against the wrong masks."""
    case = lc.build_case("empty_first")
    assert int(case["masks"][0].sum()) == 0 and int(case["masks"][1].sum()) > 0
    # the loss differs from the one obtained after removing the empty instance (which re-aligns slots and masks)
    from oracle import loss_oracle as lo
    aligned = lo.loss_from_head_output(case["out"], case["masks"][1:], case["ignore"], case["embedding_size"],
                                       case["n_free"], lc.FREE_DIM_STDS[case["n_free"]], **lc.WEIGHTS)
    assert abs(float(aligned["lovasz"]) - float(golden["empty_first/lovasz"])) > 1e-3


@pytest.fixture(scope="module")
def semseg_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "semseg_loss_golden.npz"))


@pytest.mark.parametrize("name", sorted(lc.semseg_case_table()))
def test_semseg_oracle_reproduces_reference(name, semseg_golden):
    losses, grad = lc.run_semseg_oracle(name)
    for key in ("semseg", "foreground"):
        if losses[key] is None:
            assert "%s/%s" % (name, key) not in semseg_golden
            continue
        ref = float(semseg_golden["%s/%s" % (name, key)])
        assert abs(float(losses[key]) - ref) <= 1e-6 * max(abs(ref), 1e-3)
    ref_grad = torch.from_numpy(semseg_golden["%s/grad" % name])
    assert float((grad - ref_grad).norm() / ref_grad.norm()) <= 1e-5


def test_ignore_mask_does_not_touch_the_class_loss():
    """Quirk: F.cross_entropy's default mean reduction makes the ignore mask cancel out of the class loss."""
    from oracle import loss_oracle as lo
    case = lc.build_semseg_case("kitti_3cls_fg")
    a = lo.semseg_losses_sequence(case["out"][0], case["semseg_masks"], case["ignore"])
    b = lo.semseg_losses_sequence(case["out"][0], case["semseg_masks"], torch.zeros_like(case["ignore"]))
    assert abs(float(a["semseg"]) - float(b["semseg"])) <= 1e-6 * float(a["semseg"])
    assert abs(float(a["foreground"]) - float(b["foreground"])) > 1e-4
