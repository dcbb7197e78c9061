"""Generate tests/golden/decoder_golden.npz by running the UNMODIFIED reference decoder heads on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/gen_decoder_golden.py
For every case of tests/decoder_cases.case_table() it instantiates the reference head class from its registry
(embedding_decoder.py:11, seediness_decoder.py:11, semseg_decoder.py:12) with GroupNorm(32) / AvgPool3d exactly as
build_model does (model_builder.py:282-331), loads the seeded state_dict (strict), runs the forward on the seeded
feature pyramid and stores the output.  It also asserts that oracle/decoder_oracle.py reproduces the reference
bit-for-bit (same ATen CPU kernels, same order), which pins the oracle.
"""
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from stemseg.config import cfg  # noqa: E402
from stemseg.modeling.embedding_decoder import EMBEDDING_HEAD_REGISTRY  # noqa: E402
from stemseg.modeling.seediness_decoder import SEEDINESS_HEAD_REGISTRY  # noqa: E402
from stemseg.modeling.semseg_decoder import SEMSEG_HEAD_REGISTRY  # noqa: E402

import decoder_cases as dc  # noqa: E402


def reference_head(case):
    cfg.INPUT.update_param("NUM_FRAMES", case["num_frames"])      # heads read it at construction (common.py:15,28)
    norm = partial(nn.GroupNorm, 32)
    pool = {"avg": nn.AvgPool3d, "max": nn.MaxPool3d}[case.get("pool", "avg")]       # POOLER_REGISTRY, model_builder.py:28-30
    if case["kind"] == "embedding":
        return EMBEDDING_HEAD_REGISTRY["squeeze_expand_decoder"](
            case["in_channels"], case["inter"], case["embedding_size"], tanh_activation=case["tanh"],
            seediness_output=case["seediness_output"], experimental_dims=case["dim_mode"], PoolType=pool,
            NormType=norm)
    if case["kind"] == "seediness":
        return SEEDINESS_HEAD_REGISTRY["squeeze_expand_decoder"](case["in_channels"], case["inter"],
                                                                 PoolType=pool, NormType=norm)
    return SEMSEG_HEAD_REGISTRY["squeeze_expand_decoder"](
        case["in_channels"], case["num_out"] - 1, inter_channels=case["inter"], feature_scales=[4, 8, 16, 32],
        foreground_channel=True, PoolType=pool, NormType=norm)


def main():
    torch.set_num_threads(8)
    out = {}
    for name in dc.case_table():
        sd, feats, case = dc.build_case(name)
        head = reference_head(case).eval()
        missing = head.load_state_dict(sd, strict=True)
        with torch.no_grad():
            ref = head([f.clone() for f in feats])
            ora = dc.run_oracle(name)
        assert ref.shape == ora.shape, (name, ref.shape, ora.shape)
        assert torch.equal(ref, ora), "%s: oracle differs from the reference (max |d| = %g)" % (
            name, (ref - ora).abs().max().item())
        out[name] = ref.numpy()
        print("%-24s out %s  |max| %.4f  oracle == reference: bit-identical" % (
            name, tuple(ref.shape), ref.abs().max().item()))
    path = os.path.join(HERE, "decoder_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
