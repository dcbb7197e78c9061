"""Training support: autograd through the B200 decoder heads.

The reference trains its heads through torch autograd over nn.Conv3d / nn.GroupNorm / nn.AvgPool3d / F.interpolate
(stemseg/training/main.py:188-201 -> model_builder.py:171-208).  Here the forward is the CUDA plan of
stemseg_b200/decoder.py run eagerly with its intermediates kept, and the backward is hand-written:

  * dgrad of every convolution = the tcgen05 convolution kernel itself on the gradient planes with flipped /
    transposed weights (``stemseg_pack_conv_weight_dgrad``);
  * wgrad = a tcgen05 GEMM over the voxel index with MN-major operands read straight from the NDHWC planes
    (``stemseg_conv3d_wgrad_direct``; the first implementation on zero-padded transposed planes,
    ``stemseg_conv3d_wgrad``, stays selectable with STEMSEG_WGRAD=transposed as a cross-check);
  * output heads over the fp32 merged feature: csrc/head_train.cu; trilinear adjoint, AvgPool/ReLU and GroupNorm
    backward (whose tail writes the conv-output gradient only as bf16 planes + its channel sums): csrc/backward_ops.cu.
The same two functions (``training_forward`` / ``training_backward``) serve torch autograd (``HeadFunction``) and the
CUDA-graph trainer (stemseg_b200/training.py), which calls them without the autograd engine.

Gradients flow to the four feature maps (so the torch backbone trains as usual) and to every head parameter, which is
all DistributedDataParallel needs: its gradient all-reduce hooks fire on the parameters' ``.grad`` as with the
reference heads.  One sub-clip per call (batch 1), like the reference's MAX_SAMPLES_PER_GPU = 1 (defaults.yaml:20).
"""
import os

import torch

from stemseg_b200 import _lib
from stemseg_b200 import decoder as D


# "direct": weight gradients from the NDHWC planes with MN-major tensor-core operands (stemseg_conv3d_wgrad_direct);
# "transposed": the first implementation (zero-padded transposed copies, K-major operands), kept for cross-checking
WGRAD_MODE = os.environ.get("STEMSEG_WGRAD", "direct")

# tests set this to a list to receive the saved forward state of every training_forward call (to read the ReLU sign
# decisions of the CUDA forward)
DEBUG_SAVED = None


def _check(rc):
    _lib.check(rc)


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def training_forward(head, feats, in_planes=None):
    """Eager forward keeping what the backward needs.  Returns (out, saved).

    in_planes: the four feature maps already packed (D.pack_activation), e.g. static buffers shared by several heads."""
    spec = head.head_spec(exact=True)
    weights, out_spec = spec.weights, spec.out_spec
    planes = D.PRECISION_PLANES[head.precision]
    pools, tscale = D.pool_schedule(head.num_frames)
    has_norm = head._has_norm
    saved = {"blocks": {}, "merges": [], "planes": planes, "tscale": tscale}
    D.KEEP.reset()
    branch = []
    for b, (name, n_stages) in enumerate(D.BLOCKS):
        if in_planes is not None:
            a = in_planes[b]
            if a.n != 1:
                raise NotImplementedError("training through the B200 heads handles one sub-clip per call (batch 1)")
        else:
            if feats[b].shape[0] != 1:
                raise NotImplementedError("training through the B200 heads handles one sub-clip per call (batch 1)")
            a = D.pack_activation(feats[b], planes)
        stages = []
        for j in range(n_stages):
            conv, gamma, beta = weights.stages[name][j]
            y = D.conv3d(a, conv, allow_split=has_norm)
            st = {}
            pool = bool(pools[j] and name != "block_4x")
            a_next = D.group_norm_relu_pool(y, gamma, beta, spec.num_groups, spec.eps, pool, planes, saved=st)
            y0 = y[0] if y.dim() == 6 else y            # the statistics pass summed split-K slices into slice 0
            stages.append({"a_in": a, "y": y0, "scale_shift": st.get("scale_shift"), "mean_rstd": st.get("mean_rstd"),
                           "pool": pool, "gamma": gamma, "conv": conv})
            a = a_next
        saved["blocks"][name] = stages
        branch.append(a)
    x = branch[0]
    out = None
    for k in range(3):
        w_up, w_skip = weights.merges[k]
        y_low = D.conv3d(x, w_up)
        z = D.conv3d(branch[k + 1], w_skip)
        saved["merges"].append({"x_in": x, "f_in": branch[k + 1]})
        if k < 2:
            x = D.upsample_add(z, y_low, tscale[k], planes)
        else:
            # the merged feature x4 = z + up(y_low) is formed once, in fp32, in place of z: the forward and the
            # backward of the output heads both stream it (csrc/head_train.cu)
            n_, t_, h_, w_, c_ = z.shape
            _check(_lib.load().stemseg_upsample_add_f32(_lib.ptr(z), _lib.ptr(y_low), n_, t_, h_, w_, c_, tscale[k],
                                                        _lib.stream_ptr()))
            saved["x4"] = z
            out = D.head_output_x(z, out_spec)
    saved["out_spec"] = out_spec
    D.KEEP.reset()
    if DEBUG_SAVED is not None:
        DEBUG_SAVED.append(saved)
    return out, saved


def _dgrad_weights(head):
    """Flipped / transposed packed weights of every convolution (cached per parameter version)."""
    key = head._cache_key()
    cache = getattr(head, "_dgrad_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    lib = _lib.load()
    planes = D.PRECISION_PLANES[head.precision]
    params = dict(head.named_parameters())
    packed = {}

    def pack(wname, cin_begin, cin_count):
        w = params[wname].detach().contiguous()
        cout, cin_total = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] * w.shape[4]
        dst = _empty((planes, cin_count, taps * cout), torch.bfloat16, w.device)
        _check(lib.stemseg_pack_conv_weight_dgrad(_lib.ptr(w), cout, cin_total, cin_begin, cin_count, taps,
                                                  _lib.ptr(dst), planes, _lib.stream_ptr()))
        return D.PackedConv(dst, None, cout, cin_count, 3 if taps == 27 else 1)

    with torch.cuda.device(next(head.parameters()).device):
        for name, n_stages in D.BLOCKS:
            for j in range(n_stages):
                wname = "%s.%d.weight" % (name, 4 * j)
                packed[wname] = pack(wname, 0, params[wname].shape[1])
        c = head.inter_channels
        for k, merge in enumerate(D.MERGES):
            packed[(merge, "up")] = pack(merge + ".weight", 0, c[k])
            packed[(merge, "skip")] = pack(merge + ".weight", c[k], c[k + 1])
    head._dgrad_cache = (key, packed)
    return packed


def _to_planes(x, planes):
    """fp32 NDHWC [1,t,h,w,c] -> Planes (no activation)."""
    lib = _lib.load()
    n, t, h, w, c = x.shape
    dst = _empty((planes, n, t, h, w, c), torch.bfloat16, x.device)
    _check(lib.stemseg_to_planes(_lib.ptr(x), x.numel(), _lib.ptr(dst), planes, _lib.stream_ptr()))
    return D.Planes(dst, n, t, h, w, c)


def _wgrad(dy, dy_planes, x_planes, kernel_size, planes, dst, cin_begin):
    """dst[cout][cin_total][taps] (a parameter gradient in state_dict layout) <- wgrad(dy [1,t,h,w,co], x Planes).

    dy_planes: the same gradient as bf16 planes (shared with the dgrad convolution)."""
    lib = _lib.load()
    t, h, w, co = dy_planes.t, dy_planes.h, dy_planes.w, dy_planes.c
    ci = x_planes.c
    assert (x_planes.t, x_planes.h, x_planes.w) == (t, h, w)
    taps = 27 if kernel_size == 3 else 1
    dev = dy_planes.tensor.device
    if WGRAD_MODE == "direct":
        ks = lib.stemseg_wgrad_direct_k_splits(co, ci, t, h, w, kernel_size)
        slices = _empty((ks, taps, co, ci), torch.float32, dev)
        _check(lib.stemseg_conv3d_wgrad_direct(_lib.ptr(dy_planes.tensor), _lib.ptr(x_planes.tensor), co, ci, t, h, w,
                                               kernel_size, planes, ks, _lib.ptr(slices), _lib.stream_ptr()))
    else:
        pad = 1 if kernel_size == 3 else 0
        kp = lib.stemseg_transposed_row_length(t, h, w, pad)
        shifts = 3 if kernel_size == 3 else 1      # pre-shifted copies of x: TMA start coordinates must be 16-byte aligned
        dy_t = _empty((planes, co, kp), torch.bfloat16, dev)
        x_t = _empty((planes, shifts, ci, kp), torch.bfloat16, dev)
        _check(lib.stemseg_transpose_pad(_lib.ptr(dy), 0, t, h, w, co, pad, 1, _lib.ptr(dy_t), planes,
                                         _lib.stream_ptr()))
        _check(lib.stemseg_transpose_pad(_lib.ptr(x_planes.tensor), 1, t, h, w, ci, pad, shifts, _lib.ptr(x_t), planes,
                                         _lib.stream_ptr()))
        ks = lib.stemseg_wgrad_k_splits(co, ci, t, h, w, kernel_size, planes)
        slices = _empty((ks, taps, co, ci), torch.float32, dev)
        _check(lib.stemseg_conv3d_wgrad(_lib.ptr(dy_t), _lib.ptr(x_t), co, ci, t, h, w, kernel_size, planes, ks,
                                        _lib.ptr(slices), _lib.stream_ptr()))
    _check(lib.stemseg_wgrad_reduce(_lib.ptr(slices), ks, co, taps, ci, _lib.ptr(dst), dst.shape[1], cin_begin, 0,
                                    _lib.stream_ptr()))


def _dst(grad_dst, name, params):
    if grad_dst is not None:
        out = grad_dst[name]
        if not out.is_contiguous() or out.shape != params[name].shape:
            raise ValueError("gradient slot of %s must be a contiguous tensor of the parameter's shape" % name)
        return out
    return torch.empty_like(params[name])


EARLY_BLOCKS = (0, 1)        # block_32x, block_16x: 81 % of a head's parameters, a few % of its backward time
LATE_BLOCKS = (2, 3)         # block_8x, block_4x: the long dgrad / wgrad launches


def training_backward(head, saved, grad_out, grad_dst=None, need_feature_grads=True, phase="all", carry=None):
    """-> (list of 4 feature gradients [1,C,T,h,w] (None when not needed), {parameter name: gradient}).

    grad_dst: optional {parameter name: tensor of the parameter's shape}; gradients are WRITTEN there (e.g. views
    into a flat all-reduce buffer) instead of into fresh tensors.

    phase: "all", or the two halves used by the data-parallel trainer to overlap the gradient exchange with the long
    part of the backward pass: "early" = output heads, merges and the low-resolution blocks (block_32x / block_16x,
    whose gradients are most of the bytes) -> returns (carry, grads so far); "late" (with that carry) = block_8x /
    block_4x -> returns (feature gradients, remaining grads)."""
    if phase == "late":
        return _backward_blocks(head, saved, carry, grad_dst, need_feature_grads, LATE_BLOCKS, final=True)
    lib = _lib.load()
    planes, tscale = saved["planes"], saved["tscale"]
    spec = saved["out_spec"]
    params = dict(head.named_parameters())
    grads = {}
    dgrad_w = _dgrad_weights(head)
    dev = grad_out.device
    g = grad_out.contiguous().to(torch.float32)

    # ---- output heads ----------------------------------------------------------------------------------------------
    x4 = saved["x4"]
    n, t, h, w, c3 = x4.shape
    dx = _empty((n, t, h, w, c3), torch.float32, dev)
    d_wout = _empty((spec.n_out, c3), torch.float32, dev)
    d_bout = _empty((spec.n_out,), torch.float32, dev)
    ws_bytes = lib.stemseg_head_backward_x_workspace_bytes(c3)
    ws = _empty((ws_bytes,), torch.uint8, dev)
    _check(lib.stemseg_head_backward_x(_lib.ptr(x4), n, t, h, w, c3, _lib.ptr(spec.weight), _lib.ptr(spec.bias),
                                       _lib.ptr(spec.activation), spec.n_out, _lib.ptr(g), _lib.ptr(dx),
                                       _lib.ptr(d_wout), _lib.ptr(d_bout), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    _lib.KERNEL_LAUNCHES[0] += 2 * ((spec.n_out - 1) // 8)
    head._scatter_output_grads(d_wout, d_bout, grads)

    # ---- merges (conv1x1(cat(up(x), f)) = up(W_a x) + W_b f), highest resolution first -----------------------------
    branch_grad = [None] * 4
    d_high = dx
    for k in (2, 1, 0):
        m = saved["merges"][k]
        merge = D.MERGES[k]
        wgrad_dst = _dst(grad_dst, merge + ".weight", params)
        wflat = wgrad_dst.view(wgrad_dst.shape[0], wgrad_dst.shape[1], 1)
        nn_, th, hh, wh, ch = d_high.shape
        d_low = _empty((nn_, th // tscale[k], hh // 2, wh // 2, ch), torch.float32, dev)
        _check(lib.stemseg_upsample_transpose(_lib.ptr(d_high), nn_, th, hh, wh, ch, tscale[k], _lib.ptr(d_low),
                                              _lib.stream_ptr()))
        d_high_p, d_low_p = _to_planes(d_high, planes), _to_planes(d_low, planes)
        branch_grad[k + 1] = D.conv3d(d_high_p, dgrad_w[(merge, "skip")])                           # W_b^T dz
        _wgrad(d_high, d_high_p, m["f_in"], 1, planes, wflat, m["x_in"].c)
        d_xin = D.conv3d(d_low_p, dgrad_w[(merge, "up")])                                           # W_a^T d y_low
        _wgrad(d_low, d_low_p, m["x_in"], 1, planes, wflat, 0)
        grads[merge + ".weight"] = wgrad_dst
        d_high = d_xin
    branch_grad[0] = d_high
    carry = {"branch_grad": branch_grad, "grads": grads, "feat_grads": [None] * 4, "params": params,
             "dgrad_w": dgrad_w}
    if phase == "early":
        _backward_blocks(head, saved, carry, grad_dst, need_feature_grads, EARLY_BLOCKS, final=False)
        return carry, carry["grads"]
    return _backward_blocks(head, saved, carry, grad_dst, need_feature_grads, EARLY_BLOCKS + LATE_BLOCKS, final=True)


def _backward_blocks(head, saved, carry, grad_dst, need_feature_grads, which, final):
    """Backward of the conv stages of the scale blocks `which` (indices into D.BLOCKS), last stage first."""
    lib = _lib.load()
    planes = saved["planes"]
    branch_grad, grads, feat_grads = carry["branch_grad"], carry["grads"], carry["feat_grads"]
    params, dgrad_w = carry["params"], carry["dgrad_w"]
    dev = branch_grad[0].device
    for b, (name, n_stages) in enumerate(D.BLOCKS):
        if b not in which:
            continue
        d = branch_grad[b]
        for j in reversed(range(n_stages)):
            st = saved["blocks"][name][j]
            y = st["y"]
            nn_, t_, h_, w_, c_ = y.shape
            dn = _empty((nn_, t_, h_, w_, c_), torch.float32, dev)
            _check(lib.stemseg_pool_relu_backward(_lib.ptr(d), _lib.ptr(y), _lib.ptr(st["scale_shift"]), nn_, t_, h_, w_,
                                                  c_, 1 if st["pool"] else 0, _lib.ptr(dn), _lib.stream_ptr()))
            wname, bname = "%s.%d.weight" % (name, 4 * j), "%s.%d.bias" % (name, 4 * j)
            d_bias = _empty((c_,), torch.float32, dev)
            fused_tail = st["mean_rstd"] is not None and WGRAD_MODE == "direct"
            if st["mean_rstd"] is not None:
                groups = st["mean_rstd"].shape[1]
                dgb = _empty((nn_, c_, 2), torch.float32, dev)
                gterms = _empty((nn_, groups, 2), torch.float32, dev)
                if fused_tail:
                    # dy only as bf16 planes (+ its channel sums = the bias gradient), never in fp32
                    dy = None
                    dy_p = D.Planes(_empty((planes, nn_, t_, h_, w_, c_), torch.bfloat16, dev), nn_, t_, h_, w_, c_)
                    wsb = lib.stemseg_group_norm_backward_planes_workspace_bytes(nn_, t_ * h_ * w_, c_)
                    wsg = _empty((wsb,), torch.uint8, dev)
                    _check(lib.stemseg_group_norm_backward_planes(
                        _lib.ptr(dn), _lib.ptr(y), _lib.ptr(st["mean_rstd"]), _lib.ptr(st["gamma"]), nn_, t_ * h_ * w_, c_,
                        c_ // groups, _lib.ptr(dgb), _lib.ptr(gterms), _lib.ptr(dy_p.tensor), planes, _lib.ptr(d_bias),
                        _lib.ptr(wsg), wsb, _lib.stream_ptr()))
                else:
                    wsb = lib.stemseg_group_norm_backward_workspace_bytes(nn_, t_ * h_ * w_, c_)
                    wsg = _empty((wsb,), torch.uint8, dev)
                    _check(lib.stemseg_group_norm_backward(_lib.ptr(dn), _lib.ptr(y), _lib.ptr(st["mean_rstd"]),
                                                           _lib.ptr(st["gamma"]), nn_, t_ * h_ * w_, c_, c_ // groups,
                                                           _lib.ptr(dgb), _lib.ptr(gterms), _lib.ptr(wsg), wsb,
                                                           _lib.stream_ptr()))
                grads["%s.%d.weight" % (name, 4 * j + 1)] = dgb[0, :, 0].contiguous()
                grads["%s.%d.bias" % (name, 4 * j + 1)] = dgb[0, :, 1].contiguous()
            if not fused_tail:
                dy = dn                                           # GroupNorm backward ran in place (or no norm)
                wsb = lib.stemseg_channel_sum_workspace_bytes(t_ * h_ * w_, c_)
                wsc = _empty((wsb,), torch.uint8, dev)
                _check(lib.stemseg_channel_sum(_lib.ptr(dy), t_ * h_ * w_, c_, _lib.ptr(d_bias), _lib.ptr(wsc), wsb,
                                               _lib.stream_ptr()))
                dy_p = _to_planes(dy, planes)
            grads[bname] = d_bias
            wgrad_dst = _dst(grad_dst, wname, params)
            _wgrad(dy, dy_p, st["a_in"], 3, planes, wgrad_dst.view(wgrad_dst.shape[0], wgrad_dst.shape[1], 27), 0)
            grads[wname] = wgrad_dst
            if j == 0 and not need_feature_grads:                    # frozen backbone: skip the largest dgrad
                d = None
            else:
                d = D.conv3d(dy_p, dgrad_w[wname])                   # dgrad: conv with flipped / transposed weights
        feat_grads[b] = None if d is None else d.permute(0, 4, 1, 2, 3)          # NDHWC -> NCTHW view
    if grad_dst is not None:                                         # small gradients: copy into their slots
        for name, gr in grads.items():
            if gr.data_ptr() != grad_dst[name].data_ptr():
                grad_dst[name].copy_(gr.reshape(grad_dst[name].shape))
            grads[name] = grad_dst[name]
    return feat_grads, grads


class HeadFunction(torch.autograd.Function):
    """autograd node of one head: inputs = 4 feature maps + all parameters (so DDP / optimisers see the gradients)."""

    @staticmethod
    def forward(ctx, head, n_feats, *tensors):
        feats = list(tensors[:n_feats])
        with torch.no_grad():
            out, saved = training_forward(head, [f.detach() for f in feats])
        ctx.head, ctx.saved_state = head, saved
        ctx.param_names = [n for n, _ in head.named_parameters()]
        ctx.feat_needs = [f.requires_grad for f in feats]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        with torch.no_grad():
            feat_grads, pgrads = training_backward(ctx.head, ctx.saved_state, grad_out,
                                                   need_feature_grads=any(ctx.feat_needs))
        ctx.saved_state = None
        outs = [None, None]
        outs += [fg if need else None for fg, need in zip(feat_grads, ctx.feat_needs)]
        for name in ctx.param_names:
            gr = pgrads.get(name)
            outs.append(None if gr is None else gr)
        return tuple(outs)


def run_head_with_grad(head, feats_32_16_8_4):
    params = [p for _, p in head.named_parameters()]
    n = feats_32_16_8_4[0].shape[0]
    if n == 1:
        return HeadFunction.apply(head, len(feats_32_16_8_4), *feats_32_16_8_4, *params)
    # batch > 1 (embedding_decoder.py:101-145 accepts any N): every op of the head is per sample (GroupNorm statistics
    # included), so the batch is a loop of single-sample nodes whose parameter gradients autograd accumulates
    outs = [HeadFunction.apply(head, len(feats_32_16_8_4), *[f[i:i + 1] for f in feats_32_16_8_4], *params)
            for i in range(n)]
    return torch.cat(outs, dim=0)
