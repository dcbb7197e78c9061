"""Generate tests/golden/semseg_loss_golden.npz by running the UNMODIFIED reference losses of the semantic-segmentation
head on CPU: CrossEntropyLoss (stemseg/modeling/losses/cross_entropy.py) and TrainingModel.compute_fg_loss
(stemseg/modeling/model_builder.py:210-244), fed exactly as TrainingModel.forward feeds them (permute to [N,T,C,H,W],
split off the foreground channel, model_builder.py:119-124,180).  Asserts that oracle/loss_oracle.py reproduces the loss
values (1e-6 relative) and the gradient (1e-5 norm-wise), which pins the oracle.
Run in the build container only:  python tests/golden/gen_semseg_loss_golden.py
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
from stemseg.modeling.losses import CrossEntropyLoss  # noqa: E402
from stemseg.modeling.model_builder import TrainingModel  # noqa: E402
from stemseg.utils import ModelOutputConsts, LossConsts  # noqa: E402

import loss_cases as lc  # noqa: E402


def main():
    torch.set_num_threads(8)
    store = {}
    for name in lc.semseg_case_table():
        case = lc.build_semseg_case(name)
        out = case["out"].clone().requires_grad_(True)
        targets = [{"semseg_masks": case["semseg_masks"], "ignore_masks": case["ignore"]}]
        logits = out.permute(0, 2, 1, 3, 4)                                        # model_builder.py:180
        od = {ModelOutputConsts.OPTIMIZATION_LOSSES: {}, ModelOutputConsts.OTHERS: {}}
        if case["foreground_channel"]:
            logits, fg_logits = logits.split((logits.shape[2] - 1, 1), dim=2)      # model_builder.py:121
            TrainingModel.compute_fg_loss(None, fg_logits.squeeze(2), targets, od)
        CrossEntropyLoss()(logits, targets, od)
        ref = {"semseg": od[ModelOutputConsts.OTHERS][LossConsts.SEMSEG],
               "foreground": od[ModelOutputConsts.OPTIMIZATION_LOSSES].get(LossConsts.FOREGROUND)}
        total = sum(v for v in od[ModelOutputConsts.OPTIMIZATION_LOSSES].values())   # WEIGHT_SEMSEG = 1 (defaults.yaml:34)
        total.backward()
        ora, ora_grad = lc.run_semseg_oracle(name)
        for k in ("semseg", "foreground"):
            if ref[k] is None:
                assert ora[k] is None
                continue
            a, b = float(ref[k]), float(ora[k])
            assert abs(a - b) <= 1e-6 * max(abs(a), 1e-3), (name, k, a, b)
            store["%s/%s" % (name, k)] = np.float32(a)
        gd = (out.grad - ora_grad).norm().item() / out.grad.norm().item()
        assert gd <= 1e-5, (name, gd)
        store["%s/grad" % name] = out.grad.numpy().astype(np.float32)
        print("%-16s semseg %.6f foreground %s |grad| %.4e" % (
            name, float(ref["semseg"]), "-" if ref["foreground"] is None else "%.6f" % float(ref["foreground"]),
            float(out.grad.norm())))
    path = os.path.join(HERE, "semseg_loss_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
