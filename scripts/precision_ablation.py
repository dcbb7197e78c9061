"""Per-layer operand-precision ablation for the fp32-parity decoder.

Question: the CUDA decoder issues 3 bf16 tensor-core products per MAC (hi*hi + hi*lo + lo*hi) in every layer; can some
layers run with fewer products and stay inside the 1e-4 norm-wise budget?  Emulation on the CPU, exact everywhere
except the operand rounding under test: every convolution is evaluated in fp64 on operands rounded the way the kernel
would round them, everything else (GroupNorm, ReLU, pooling, trilinear, output convs) stays in fp64, and the result is
compared with the all-fp64 forward of the same head (embedding head, DAVIS widths [256,256,128,128], T=8, 96x128 clip,
seeded weights / features as in the golden cases).

Schemes per convolution (A = activations, W = weights; cost = tensor-core products per MAC at the bf16/fp16 rate):
  b3   bf16 split both, A_hi W_hi + A_hi W_lo + A_lo W_hi   (3)  <- shipped
  b2w  A split, W single bf16: (A_hi + A_lo) W_hi            (2)
  b2a  A single bf16, W split: A_hi (W_hi + W_lo)            (2)
  b1   bf16 both single                                      (1)
  h2w  fp16: A split, W single fp16                          (2)
  h1   fp16 both single                                      (1)
Writes profiles/r02_precision_ablation.json.   python scripts/precision_ablation.py [h4 w4]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import decoder_oracle as do  # noqa: E402

IN_CH, INTER = 256, [256, 256, 128, 128]
GMAC_SHARE = {"4x": 0.650, "8x": 0.162, "16x": 0.122, "32x": 0.036, "merge": 0.030}     # of a head at 8x480x864
COST = {"b3": 3, "b2w": 2, "b2a": 2, "b1": 1, "h2w": 2, "h1": 1}


def split(x, dtype):
    hi = x.to(dtype).to(torch.float64)
    lo = (x - hi).to(dtype).to(torch.float64)
    return hi, lo


def make_conv(scheme_of, groups_by_weight):
    real = F.conv3d

    def conv3d(x, w, b=None, stride=1, padding=0):
        group = groups_by_weight.get(id(w))
        scheme = scheme_of.get(group, "exact") if group else "exact"
        if scheme == "exact":
            return real(x, w, b, stride=stride, padding=padding)
        dt = torch.float16 if scheme.startswith("h") else torch.bfloat16
        # operands arrive as fp32 values in the kernel: activations are fp32 results of the previous fp32 stage
        x32, w32 = x.to(torch.float32).to(torch.float64), w.to(torch.float32).to(torch.float64)
        ah, al = split(x32, dt)
        wh, wl = split(w32, dt)
        if scheme in ("b3",):
            y = real(ah, wh + wl, None, stride=stride, padding=padding) + real(al, wh, None, stride=stride, padding=padding)
        elif scheme in ("b2w", "h2w"):
            y = real(ah + al, wh, None, stride=stride, padding=padding)
        elif scheme == "b2a":
            y = real(ah, wh + wl, None, stride=stride, padding=padding)
        else:
            y = real(ah, wh, None, stride=stride, padding=padding)
        if b is not None:
            y = y + b.view(1, -1, 1, 1, 1)
        return y
    return conv3d


def main():
    h4, w4 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 32)
    shapes = do.head_parameter_shapes("embedding", IN_CH, INTER, embedding_size=4, dim_mode="xyff", seediness_output=True)
    sd32 = do.seeded_state_dict(shapes, 42 + int(os.environ.get("ABLATION_SEED", "7")) - 7)
    sd = {k: v.to(torch.float64) for k, v in sd32.items()}
    seed = int(os.environ.get("ABLATION_SEED", "7"))
    feats = [f.to(torch.float64) for f in do.seeded_features(seed, 1, IN_CH, 8, h4, w4)]
    groups = {}
    for key, v in sd.items():
        if v.dim() != 5:
            continue
        if key.startswith("block_"):
            groups[id(v)] = key.split(".")[0].replace("block_", "")
        elif key in ("conv_16.weight", "conv_8.weight", "conv_4.weight"):
            groups[id(v)] = "merge"
    real = F.conv3d

    def forward(scheme_of):
        do.F.conv3d = make_conv(scheme_of, groups)
        try:
            with torch.no_grad():
                return do.embedding_head(sd, feats, 8, 4, "xyff", True, True)[0]
        finally:
            do.F.conv3d = real

    t0 = time.time()
    ref = forward({})
    names = ["emb0(y)", "emb1(x)", "emb2(free)", "emb3(free)", "var0", "var1", "seediness"]

    def errors(out):
        return [float((out[c] - ref[c]).abs().max() / ref[c].abs().max()) for c in range(ref.shape[0])]

    all_groups = ["32x", "16x", "8x", "4x", "merge"]
    rows = []

    def run(label, scheme_of):
        errs = errors(forward(scheme_of))
        products = sum(GMAC_SHARE[g] * COST[scheme_of[g]] for g in all_groups)
        row = {"config": label, "schemes": dict(scheme_of), "worst_channel_error": max(errs),
               "per_channel_error": dict(zip(names, errs)), "products_per_mac": round(products, 3),
               "admissible_1e-4": max(errs) <= 1e-4, "admissible_with_2x_margin": max(errs) <= 5e-5}
        rows.append(row)
        print("%-34s worst %.2e  products/MAC %.2f  %s  (%.0fs)" % (
            label, max(errs), products, "OK" if row["admissible_1e-4"] else "FAILS", time.time() - t0), flush=True)
        return row

    base = {g: "b3" for g in all_groups}
    run("all b3 (shipped)", base)
    for g in all_groups:
        for scheme in ("b2w", "b2a", "b1", "h2w", "h1"):
            cfg = dict(base)
            cfg[g] = scheme
            run("%s -> %s" % (g, scheme), cfg)
    for scheme in ("b2w", "b2a", "h2w"):
        cfg = {g: (scheme if g != "4x" else "b3") for g in all_groups}
        run("all but 4x -> %s" % scheme, cfg)
        cfg = {g: (scheme if g in ("32x", "16x") else "b3") for g in all_groups}
        run("32x+16x -> %s" % scheme, cfg)
    for scheme in ("h1", "h2w"):
        for sel in (("8x",), ("8x", "16x"), ("8x", "16x", "32x")):
            cfg = {g: (scheme if g in sel else "b3") for g in all_groups}
            run("%s -> %s" % ("+".join(sel), scheme), cfg)
    ok = [r for r in rows if r["admissible_with_2x_margin"]]
    best = min(ok, key=lambda r: r["products_per_mac"])
    out = {"what": __doc__.split("\n\n")[0], "clip": [8, h4 * 4, w4 * 4], "tolerance": 1e-4,
           "gmac_share_at_8x480x864": GMAC_SHARE, "rows": rows,
           "cheapest_admissible_with_2x_margin": {"config": best["config"], "products_per_mac": best["products_per_mac"],
                                                  "worst_channel_error": best["worst_channel_error"]},
           "note": "the dominant kernel (4x layer, 65 % of the MACs) needs 3 products in every admissible mix, so "
                   "roofline.frac of that launch stays capped at 1/3 of the bf16 peak",
           "gpu_follow_up": "block_8x + block_16x as single fp16 products were then built (STEMSEG_PLANES_FP16) and "
                            "measured on the B200 over every golden and under the reference's callers with the "
                            "PER-CHANNEL bound: 1.06e-4 on the 2-frame golden, 9.7e-5 on the free-dimension channels of "
                            "a random-init model (even block_8x alone), versus <= 1.1e-5 / 2.9e-5 with three products "
                            "everywhere (profiles/r02_fp16_blocks_golden_errors.txt, r02_fp16_blocks_reference_errors."
                            "txt).  No mix below 3 products per MAC keeps a margin on every case, so the shipped "
                            "fp32-parity plan stays at 3.00; the fp16 blocks remain an opt-in (+16 % clips/s)."}
    with open(os.path.join(ROOT, "profiles", "r02_precision_ablation.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("cheapest admissible (2x margin):", best["config"], best["products_per_mac"])


if __name__ == "__main__":
    main()
