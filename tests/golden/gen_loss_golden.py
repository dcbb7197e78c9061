"""Generate tests/golden/loss_golden.npz by running the UNMODIFIED reference EmbeddingLoss on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/gen_loss_golden.py
For every case of tests/loss_cases.case_table() the reference criterion (stemseg/modeling/losses/embedding_loss.py)
is built as model_builder.py:294-298 does, called on the seeded head output / targets, and back-propagated; the loss
terms and the gradient wrt the head output are stored.  It also asserts that oracle/loss_oracle.py reproduces the
reference's loss terms to 1e-6 relative and its gradient to 1e-5 norm-wise (fp32 CPU), which pins the oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from stemseg.modeling.losses import EmbeddingLoss  # noqa: E402
from stemseg.utils import ModelOutputConsts, LossConsts  # noqa: E402

import loss_cases as lc  # noqa: E402


def main():
    torch.set_num_threads(8)
    store = {}
    for name in lc.case_table():
        case = lc.build_case(name)
        crit = EmbeddingLoss(4, embedding_size=case["embedding_size"], nbr_free_dims=case["n_free"],
                             free_dim_stds=lc.FREE_DIM_STDS[case["n_free"]], weight_variance_smoothness=10.0,
                             weight_lovasz=1.0, weight_regularization=0.001, weight_seediness=1.0, weight=1.0)
        out = case["out"].clone().requires_grad_(True)
        targets = [{"masks": case["masks"], "ignore_masks": case["ignore"]}]
        od = {}
        crit(out, targets, od)
        total = od[ModelOutputConsts.OPTIMIZATION_LOSSES][LossConsts.EMBEDDING]
        others = od[ModelOutputConsts.OTHERS]
        if total.requires_grad:
            total.backward()
        grad = out.grad if out.grad is not None else torch.zeros_like(out)
        ref = {"total": total.detach(), "lovasz": torch.as_tensor(others[LossConsts.LOVASZ_LOSS]).detach(),
               "variance_smoothness": torch.as_tensor(others[LossConsts.VARIANCE_SMOOTHNESS]).detach(),
               "seediness": torch.as_tensor(others[LossConsts.SEEDINESS_LOSS]).detach()}
        ora, ora_grad = lc.run_oracle(name)
        for k in ref:
            a, b = float(ref[k]), float(ora[k])
            # not bit-equal in general: the reference orders an instance's points by an UNSTABLE argsort of the
            # instance ids (embedding_loss.py:84-85), so its own fp32 means depend on the sort's permutation
            assert abs(a - b) <= 1e-6 * max(abs(a), 1e-3) or (np.isnan(a) and np.isnan(b)), \
                "%s: oracle %s = %r, reference %r" % (name, k, b, a)
        # the gradients agree to rounding (autograd accumulates the index_put /
        # slice gradients in a different order than the reference's permute+split graph)
        gd = (grad - ora_grad).norm().item() / max(grad.norm().item(), 1e-30)
        assert gd <= 1e-5, "%s: oracle gradient differs from the reference (relative %g)" % (name, gd)
        for k in ref:
            store["%s/%s" % (name, k)] = np.float32(float(ref[k]))
        store["%s/grad" % name] = grad.numpy().astype(np.float32)
        print("%-16s total %.6f lovasz %.6f smooth %.6f seed %.6f  |grad| %.4e" % (
            name, float(ref["total"]), float(ref["lovasz"]), float(ref["variance_smoothness"]),
            float(ref["seediness"]), float(grad.norm())))
    path = os.path.join(HERE, "loss_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
