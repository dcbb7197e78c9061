"""CPU restatement (plain PyTorch fp32 functional ops) of the reference's 3-D squeeze-expand decoder heads.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  This is the "plain PyTorch fp32 reference" the CUDA decoder kernels
are compared with (floating-point path: tolerance 1e-4 norm-wise per output group, see tests/test_decoder_gpu.py).
It is table-driven (one ``trunk`` shared by the three head kinds) and follows:
  * embedding head   stemseg/modeling/embedding_decoder.py:12-145
  * seediness head   stemseg/modeling/seediness_decoder.py:12-112
  * semseg head      stemseg/modeling/semseg_decoder.py:13-116      (input list order reversed, :94)
  * pool / temporal-scale tables   stemseg/modeling/common.py:8-35
  * trilinear upsampling           stemseg/modeling/common.py:69-78
  * coordinate offsets             stemseg/modeling/embedding_utils.py:29-120
Pinned against the reference itself by tests/golden/gen_decoder_golden.py (same state_dict, same inputs ->
bit-identical outputs on CPU, stored as fixtures).
"""
import torch
import torch.nn.functional as F

# common.py:15-24: which of the three pooling slots are real AvgPool3d(3, stride=(2,1,1), padding=1) layers
POOL_SLOTS = {2: (False, False, False), 4: (True, False, False), 8: (True, True, False),
              16: (True, True, True), 24: (True, True, True), 32: (True, True, True)}
# common.py:27-35: temporal factor of the three upsampling steps (32->16, 16->8, 8->4)
TEMPORAL_SCALES = {2: (1, 1, 1), 4: (1, 1, 2), 8: (1, 2, 2), 16: (2, 2, 2), 24: (2, 2, 2), 32: (2, 2, 2)}

# (block name, number of conv stages); stage j of a block uses pooling slot j, except block_4x which has no
# pooling layer at all (embedding_decoder.py:20-60)
BLOCKS = (("block_32x", 3), ("block_16x", 2), ("block_8x", 1), ("block_4x", 1))
UNPOOLED_BLOCKS = ("block_4x",)
MERGES = ("conv_16", "conv_8", "conv_4")

EMBEDDING_DIMS = {"xy": 2, "ff": 2, "xyt": 3, "xyf": 3, "xytf": 4, "xyff": 4, "xytff": 5, "xyfff": 5}   # embedding_utils.py:4-14
FREE_DIMS = {"xyf": 1, "xytf": 1, "xyff": 2, "xytff": 2, "xyfff": 3}                                        # embedding_utils.py:17-26
# which coordinate is added to which embedding channel (embedding_utils.py:44-120)
OFFSET_CHANNELS = {"xy": "yx", "ff": "", "xyt": "tyx", "xyf": "yx", "xytf": "tyx", "xyff": "yx", "xytff": "tyx",
                   "xyfff": "yx"}


def conv_stage(x, sd, prefix, conv_idx, gn_groups, pooled, trace=None, relu_masks=None, pool_type="avg"):
    """conv3x3x3(pad 1) -> GroupNorm -> ReLU -> [AvgPool3d(3, (2,1,1), 1)]   (embedding_decoder.py:21-24).

    relu_masks (gradient tests only): {"<block>.<conv_idx>": 0/1 tensor} replaces the ReLU's own sign decision by a
    given one, so that two implementations whose pre-activations differ by rounding are differentiated along the
    SAME linear piece (the ReLU kink makes the gradient discontinuous in elements that round across zero)."""
    w, b = sd["%s.%d.weight" % (prefix, conv_idx)], sd["%s.%d.bias" % (prefix, conv_idx)]
    y = F.conv3d(x, w, b, stride=1, padding=1)
    if trace is not None:
        trace["%s.%d.conv" % (prefix, conv_idx)] = y
    if gn_groups:
        y = F.group_norm(y, gn_groups, sd["%s.%d.weight" % (prefix, conv_idx + 1)],
                         sd["%s.%d.bias" % (prefix, conv_idx + 1)], eps=1e-5)
    key = "%s.%d" % (prefix, conv_idx)
    if relu_masks is not None and key in relu_masks:
        y = y * relu_masks[key].to(y.dtype)
    else:
        y = F.relu(y)
    if pooled and pool_type == "max":                             # POOLER_REGISTRY "max" (model_builder.py:28-30)
        y = F.max_pool3d(y, 3, stride=(2, 1, 1), padding=1)
    elif pooled:
        y = F.avg_pool3d(y, 3, stride=(2, 1, 1), padding=1)      # count_include_pad=True -> always /27
    if trace is not None:
        trace["%s.%d.out" % (prefix, conv_idx)] = y
    return y


def trunk(sd, feats_32_16_8_4, num_frames, gn_groups=32, trace=None, relu_masks=None, pool_type="avg"):
    """Shared trunk of all three heads -> [N, c3, T, H/4, W/4] (embedding_decoder.py:109-129)."""
    pools = POOL_SLOTS[num_frames]
    tscale = TEMPORAL_SCALES[num_frames]
    branch = []
    for (name, stages), f in zip(BLOCKS, feats_32_16_8_4):
        y = f
        for j in range(stages):
            y = conv_stage(y, sd, name, 4 * j, gn_groups, pools[j] and name not in UNPOOLED_BLOCKS, trace,
                           relu_masks, pool_type)    # Sequential indices 0,4,8
        branch.append(y)
    x = branch[0]
    for k, merge in enumerate(MERGES):
        x = F.interpolate(x, scale_factor=(tscale[k], 2, 2), mode="trilinear", align_corners=False)
        x = torch.cat((x, branch[k + 1]), dim=1)
        x = F.conv3d(x, sd[merge + ".weight"], None)
        if trace is not None:
            trace[merge] = x
    return x


def coordinate_grid(t, h, w, time_scale=1.0):
    """embedding_utils.py:29-41 (linspace endpoints max(1, W/H), max(1, H/W), time_scale)."""
    x_abs, y_abs = max(1.0, w / float(h)), max(1.0, h / float(w))
    xs = torch.linspace(-x_abs, x_abs, w, dtype=torch.float32)
    ys = torch.linspace(-y_abs, y_abs, h, dtype=torch.float32)
    ts = torch.linspace(-float(time_scale), float(time_scale), t, dtype=torch.float32)
    return {"t": ts.view(t, 1, 1).expand(t, h, w), "y": ys.view(1, h, 1).expand(t, h, w),
            "x": xs.view(1, 1, w).expand(t, h, w)}


def embedding_head(sd, feats, num_frames, embedding_size, dim_mode, tanh_activation=True, seediness_output=True,
                   gn_groups=32, trace=None, relu_masks=None, pool_type="avg"):
    """-> cat(embeddings, variances[, seediness]) [N, E + (E - free) + {0,1}, T, H/4, W/4]."""
    x = trunk(sd, feats, num_frames, gn_groups, trace, relu_masks, pool_type)
    emb = F.conv3d(x, sd["conv_embedding.weight"], None)
    assert emb.shape[1] == EMBEDDING_DIMS[dim_mode] == embedding_size
    if tanh_activation:
        emb = (emb * 0.25).tanh()
    n, _, t, h, w = emb.shape
    grid = coordinate_grid(t, h, w, float(sd["time_scale"]) if "time_scale" in sd else 1.0)
    offs = torch.zeros_like(emb)
    for ch, axis in enumerate(OFFSET_CHANNELS[dim_mode]):
        offs[:, ch] = grid[axis]
    emb = emb + offs
    var = F.conv3d(x, sd["conv_variance.weight"], sd["conv_variance.bias"])
    assert var.shape[1] == embedding_size - FREE_DIMS.get(dim_mode, 0)
    outs = [emb, var]
    if seediness_output:
        outs.append(F.conv3d(x, sd["conv_seediness.weight"], None).sigmoid())
    return torch.cat(outs, dim=1)


def seediness_head(sd, feats, num_frames, gn_groups=32, trace=None, relu_masks=None, pool_type="avg"):
    return F.conv3d(trunk(sd, feats, num_frames, gn_groups, trace, relu_masks, pool_type), sd["conv_out.weight"],
                    None).sigmoid()


def semseg_head(sd, feats_4_8_16_32, num_frames, gn_groups=32, trace=None, relu_masks=None, pool_type="avg"):
    """Input list arrives highest resolution first and is reversed (semseg_decoder.py:94)."""
    return F.conv3d(trunk(sd, feats_4_8_16_32[::-1], num_frames, gn_groups, trace, relu_masks, pool_type),
                    sd["conv_out.weight"], None)


# --------------------------------------------------------------------------------------------------------------
# deterministic, platform-independent parameters / inputs (numpy PCG64) shared by the golden generator and tests
# --------------------------------------------------------------------------------------------------------------
def head_parameter_shapes(kind, in_channels, inter, embedding_size=None, dim_mode=None, seediness_output=True,
                          num_out=None, gn=True):
    """state_dict key -> shape for one head, in the reference's key order."""
    shapes = {}
    if kind == "embedding":
        shapes["time_scale"] = ()
    for (name, stages), c in zip(BLOCKS, inter):
        cin = in_channels
        for j in range(stages):
            shapes["%s.%d.weight" % (name, 4 * j)] = (c, cin, 3, 3, 3)
            shapes["%s.%d.bias" % (name, 4 * j)] = (c,)
            if gn:
                shapes["%s.%d.weight" % (name, 4 * j + 1)] = (c,)
                shapes["%s.%d.bias" % (name, 4 * j + 1)] = (c,)
            cin = c
    for k, merge in enumerate(MERGES):
        shapes[merge + ".weight"] = (inter[k + 1], inter[k] + inter[k + 1], 1, 1, 1)
    if kind == "embedding":
        shapes["conv_embedding.weight"] = (EMBEDDING_DIMS[dim_mode], inter[3], 1, 1, 1)
        shapes["conv_variance.weight"] = (embedding_size - FREE_DIMS.get(dim_mode, 0), inter[3], 1, 1, 1)
        shapes["conv_variance.bias"] = (embedding_size - FREE_DIMS.get(dim_mode, 0),)
        if seediness_output:
            shapes["conv_seediness.weight"] = (1, inter[3], 1, 1, 1)
    elif kind == "seediness":
        shapes["conv_out.weight"] = (1, inter[3], 1, 1, 1)
    else:
        shapes["conv_out.weight"] = (num_out, inter[3], 1, 1, 1)
    return shapes


def seeded_state_dict(shapes, seed):
    """Weights ~ kaiming-uniform-like scale, GN affine around (1, 0) -- values do not matter, determinism does."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in shapes.items():
        if key == "time_scale":
            sd[key] = torch.tensor(1.0)
            continue
        if len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            bound = (1.0 / fan_in) ** 0.5
            val = rng.uniform(-bound, bound, size=shape)
        elif key.endswith(".weight"):          # GroupNorm gamma
            val = 1.0 + 0.2 * rng.standard_normal(shape)
        else:                                   # conv / GroupNorm bias
            val = 0.1 * rng.standard_normal(shape)
        sd[key] = torch.from_numpy(val.astype("float32"))
    return sd


def seeded_features(seed, n, channels, t, h4, w4, order=(32, 16, 8, 4)):
    """FPN-like feature list [N,C,T,h,w] for the given scale order; (h4, w4) is the stride-4 size (multiples of 8)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    feats = {}
    for s in (4, 8, 16, 32):
        h, w = h4 * 4 // s, w4 * 4 // s
        feats[s] = torch.from_numpy(rng.standard_normal((n, channels, t, h, w)).astype("float32"))
    return [feats[s] for s in order]
