"""Worst per-channel norm-wise error of the fp32-parity plan on every decoder golden, for the current
STEMSEG_FP32_FAST_BLOCKS setting (one process per setting: the variable is read at import).
    STEMSEG_FP32_FAST_BLOCKS=block_8x python scripts/golden_errors.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import decoder_cases as dc  # noqa: E402
from test_decoder_gpu import build_head  # noqa: E402
from stemseg_b200 import decoder as D  # noqa: E402

dev = torch.device("cuda:0")
golden = np.load(os.path.join(ROOT, "tests", "golden", "decoder_golden.npz"))
out = {"fast_blocks": list(D.FP32_FAST_BLOCKS), "cases": {}}
worst = 0.0
for name in sorted(dc.case_table()):
    sd, feats, case = dc.build_case(name)
    head = build_head(case, sd, dev)
    with torch.no_grad():
        got = head([f.to(dev) for f in feats]).cpu().double()
    ref = torch.from_numpy(golden[name]).double()
    errs = [float((got[:, c] - ref[:, c]).abs().max() / ref[:, c].abs().max()) for c in range(ref.shape[1])]
    out["cases"][name] = {"worst": max(errs), "channel": int(np.argmax(errs))}
    worst = max(worst, max(errs))
    print("%-26s worst channel error %.2e (channel %d of %d)" % (name, max(errs), int(np.argmax(errs)), len(errs)))
out["worst"] = worst
print(json.dumps(out))
