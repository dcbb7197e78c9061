"""Host-side launch plan of the B200 decoder heads (one trunk shared by the embedding / seediness / semseg heads).

This is plumbing only: it owns no arithmetic.  Every tensor op on the path is a kernel of ``libstemseg_b200.so``
(csrc/conv_tc.cu, csrc/decoder_ops.cu) reached through the C ABI; torch is used for device memory and streams.

Reference structure being reproduced (stemseg/modeling/embedding_decoder.py:101-145, same trunk in
seediness_decoder.py:82-112 and semseg_decoder.py:91-116):

    f32 -> [conv3 GN ReLU P0][conv3 GN ReLU P1][conv3 GN ReLU P2]
    f16 -> [conv3 GN ReLU P0][conv3 GN ReLU P1]         x = conv_16(cat(up(x32), f16'))
    f8  -> [conv3 GN ReLU P0]                           x = conv_8 (cat(up(x),   f8'))
    f4  -> [conv3 GN ReLU]                              x = conv_4 (cat(up(x),   f4'))   -> 1x1x1 output convs

The merges use conv1x1(cat(up(x), f)) == up(W_a x) + W_b f (both linear, the conv is pointwise), so the upsampled
half of each merge is a GEMM at the LOW resolution and the concatenated tensor is never materialised.
"""
import torch

from stemseg_b200 import _lib

# stemseg/modeling/common.py:15-24 and :27-35
POOL_SLOTS = {2: (False, False, False), 4: (True, False, False), 8: (True, True, False),
              16: (True, True, True), 24: (True, True, True), 32: (True, True, True)}
TEMPORAL_SCALES = {2: (1, 1, 1), 4: (1, 1, 2), 8: (1, 2, 2), 16: (2, 2, 2), 24: (2, 2, 2), 32: (2, 2, 2)}
BLOCKS = (("block_32x", 3), ("block_16x", 2), ("block_8x", 1), ("block_4x", 1))
MERGES = ("conv_16", "conv_8", "conv_4")

ACT_IDENTITY, ACT_TANH_QUARTER, ACT_SIGMOID = 0, 1, 2
COORD_NONE, COORD_T, COORD_Y, COORD_X = 0, 1, 2, 3

PRECISION_PLANES = {"fp32": 2, "bf16": 1}

# bench.py sets this to a list to collect (shape, start_event, end_event) around every conv launch
PROFILE_EVENTS = None


def pool_schedule(num_frames):
    if num_frames not in POOL_SLOTS:
        raise NotImplementedError("NUM_FRAMES=%r is not supported by the reference decoder tables "
                                  "(stemseg/modeling/common.py:15-35)" % (num_frames,))
    return POOL_SLOTS[num_frames], TEMPORAL_SCALES[num_frames]


class PackedConv(object):
    """Weights of one convolution in kernel layout: bf16 planes [P][rows][taps*cin] (+ fp32 bias)."""

    def __init__(self, planes_tensor, bias, cin, cout, kernel_size):
        self.planes_tensor, self.bias, self.cin, self.cout, self.kernel_size = planes_tensor, bias, cin, cout, kernel_size


def _check(rc):
    _lib.check(rc)


def pack_conv_weight(weight, planes, cin_begin=0, cin_count=None, bias=None):
    """weight: [cout, cin_total, k, k, k] fp32 CUDA parameter -> PackedConv for input channels [cin_begin, +cin_count)."""
    lib = _lib.load()
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_cuda:
        raise ValueError("decoder weights must be fp32 CUDA tensors (got %s on %s)" % (w.dtype, w.device))
    w = w.contiguous()
    cout, cin_total = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    if taps not in (1, 27):
        raise NotImplementedError("only 1x1x1 and 3x3x3 convolutions are on the path (got %s)" % (tuple(w.shape),))
    cin_count = cin_total - cin_begin if cin_count is None else cin_count
    with torch.cuda.device(w.device):
        dst = torch.empty((planes, cout, taps * cin_count), dtype=torch.bfloat16, device=w.device)
        _check(lib.stemseg_pack_conv_weight(_lib.ptr(w), cout, cin_total, cin_begin, cin_count, taps, _lib.ptr(dst), 0,
                                            cout, planes, _lib.stream_ptr()))
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(dst, b, cin_count, cout, 3 if taps == 27 else 1)


class Planes(object):
    """An NDHWC activation stored as bf16 planes [P][n][t][h][w][c]."""

    def __init__(self, tensor, n, t, h, w, c):
        self.tensor, self.n, self.t, self.h, self.w, self.c = tensor, n, t, h, w, c

    @property
    def planes(self):
        return self.tensor.shape[0]


def pack_activation(x, planes):
    """[N,C,T,H,W] fp32 CUDA tensor (H,W contiguous; other strides free) -> Planes."""
    lib = _lib.load()
    if x.dim() != 5:
        raise ValueError("expected a [N,C,T,H,W] feature map, got shape %s" % (tuple(x.shape),))
    if x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("decoder inputs must be fp32 CUDA tensors (got %s on %s); there is no CPU path" % (
            x.dtype, x.device))
    n, c, t, h, w = x.shape
    if not (x.stride(4) == 1 and x.stride(3) == w):
        x = x.contiguous()
    with torch.cuda.device(x.device):
        dst = torch.empty((planes, n, t, h, w, c), dtype=torch.bfloat16, device=x.device)
        _check(lib.stemseg_pack_activation(_lib.ptr(x), x.stride(0), x.stride(1), x.stride(2), n, c, t, h * w,
                                           _lib.ptr(dst), planes, _lib.stream_ptr()))
    return Planes(dst, n, t, h, w, c)


def conv3d(act, packed, max_ctas=0, allow_split=False):
    """Planes x PackedConv -> fp32 NDHWC tensor [n,t,h,w,cout] (tcgen05 implicit GEMM).

    With allow_split the library may split the taps over several CTAs for layers with fewer tiles than SMs; the
    result is then [split_k, n, t, h, w, cout] partial sums (added by the GroupNorm kernels)."""
    lib = _lib.load()
    if act.c != packed.cin:
        raise ValueError("conv input has %d channels, weights expect %d" % (act.c, packed.cin))
    if act.planes != packed.planes_tensor.shape[0]:
        raise ValueError("activation / weight precision mismatch")
    shape = _lib.StemsegConvShape(act.n, act.t, act.h, act.w, packed.cin, packed.cout, packed.kernel_size, act.planes,
                                  1)
    with torch.cuda.device(act.tensor.device):
        if allow_split:
            shape.split_k = lib.stemseg_conv3d_auto_split(shape)
            out = torch.empty((shape.split_k, act.n, act.t, act.h, act.w, packed.cout), dtype=torch.float32,
                              device=act.tensor.device)
        else:
            out = torch.empty((act.n, act.t, act.h, act.w, packed.cout), dtype=torch.float32,
                              device=act.tensor.device)
        ev = None
        if PROFILE_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _check(lib.stemseg_conv3d_bf16_planes(_lib.ptr(act.tensor), _lib.ptr(packed.planes_tensor),
                                              _lib.ptr(packed.bias), _lib.ptr(out), shape, max_ctas,
                                              _lib.stream_ptr()))
        if ev is not None:
            ev[1].record()
            PROFILE_EVENTS.append(((act.n, act.t, act.h, act.w, packed.cin, packed.cout, packed.kernel_size,
                                    act.planes), ev[0], ev[1]))
    return out


def group_norm_relu_pool(y, gamma, beta, num_groups, eps, pool, planes):
    """fp32 NDHWC conv output ([split_k,] n,t,h,w,c) -> relu(GN(y)) [-> avgpool] as Planes.  gamma None = no norm."""
    lib = _lib.load()
    slices = 1
    if y.dim() == 6:
        slices = y.shape[0]
    n, t, h, w, c = y.shape[-5:]
    dev = y.device
    with torch.cuda.device(dev):
        mean_rstd, cpg = None, 1
        if gamma is not None:
            if c % num_groups != 0:
                raise ValueError("channels %d not divisible by %d groups" % (c, num_groups))
            cpg = c // num_groups
            ws_bytes = lib.stemseg_group_norm_workspace_bytes(n, t * h * w, c)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            mean_rstd = torch.empty((n, num_groups, 2), dtype=torch.float32, device=dev)
            _check(lib.stemseg_group_norm_stats(_lib.ptr(y), slices, n, t * h * w, c, cpg, float(eps),
                                                _lib.ptr(mean_rstd), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        t_out = (t - 1) // 2 + 1 if pool else t
        dst = torch.empty((planes, n, t_out, h, w, c), dtype=torch.bfloat16, device=dev)
        _check(lib.stemseg_norm_relu_pool(_lib.ptr(y), slices, _lib.ptr(mean_rstd), _lib.ptr(gamma), _lib.ptr(beta), n, t,
                                          h, w, c, cpg, 1 if pool else 0, _lib.ptr(dst), planes, _lib.stream_ptr()))
    return Planes(dst, n, t_out, h, w, c)


def upsample_add(z, y_low, t_scale, planes):
    lib = _lib.load()
    n, t, h, w, c = z.shape
    if tuple(y_low.shape) != (n, t // t_scale, h // 2, w // 2, c) or t % t_scale or h % 2 or w % 2:
        raise ValueError("upsample_add: low-res %s does not upsample by (%d,2,2) to %s" % (
            tuple(y_low.shape), t_scale, tuple(z.shape)))
    with torch.cuda.device(z.device):
        dst = torch.empty((planes, n, t, h, w, c), dtype=torch.bfloat16, device=z.device)
        _check(lib.stemseg_upsample_add(_lib.ptr(z), _lib.ptr(y_low), n, t, h, w, c, t_scale, _lib.ptr(dst), planes,
                                        _lib.stream_ptr()))
    return Planes(dst, n, t, h, w, c)


class OutputSpec(object):
    """The fused 1x1x1 output convs of one head: weight [J,c3], bias [J], activation / coordinate codes [J]."""

    def __init__(self, weight, bias, activation, coordinate, time_scale=1.0):
        dev = weight.device
        self.weight = weight.detach().to(torch.float32).contiguous()
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()
        self.activation = torch.tensor(list(activation), dtype=torch.int32, device=dev)
        self.coordinate = torch.tensor(list(coordinate), dtype=torch.int32, device=dev)
        self.n_out = int(self.weight.shape[0])
        self.time_scale = float(time_scale)


def head_output(z, y_low, t_scale, spec):
    lib = _lib.load()
    n, t, h, w, c = z.shape
    if tuple(y_low.shape) != (n, t // t_scale, h // 2, w // 2, c) or t % t_scale or h % 2 or w % 2:
        raise ValueError("head_output: low-res %s does not upsample by (%d,2,2) to %s" % (
            tuple(y_low.shape), t_scale, tuple(z.shape)))
    if spec.weight.shape[1] != c:
        raise ValueError("output conv expects %d channels, got %d" % (spec.weight.shape[1], c))
    with torch.cuda.device(z.device):
        out = torch.empty((n, spec.n_out, t, h, w), dtype=torch.float32, device=z.device)
        _check(lib.stemseg_head_output(_lib.ptr(z), _lib.ptr(y_low), n, t, h, w, c, t_scale, _lib.ptr(spec.weight),
                                       _lib.ptr(spec.bias), _lib.ptr(spec.activation), _lib.ptr(spec.coordinate),
                                       spec.n_out, spec.time_scale, _lib.ptr(out), _lib.stream_ptr()))
    return out


class TrunkWeights(object):
    """Kernel-layout weights of one head: per block a list of (PackedConv, gamma, beta); per merge (W_a, W_b)."""

    def __init__(self, state, inter_channels, planes, has_norm):
        self.stages = {}
        for name, n_stages in BLOCKS:
            lst = []
            for j in range(n_stages):
                conv = pack_conv_weight(state["%s.%d.weight" % (name, 4 * j)], planes,
                                        bias=state["%s.%d.bias" % (name, 4 * j)])
                gamma = beta = None
                if has_norm:
                    gamma = state["%s.%d.weight" % (name, 4 * j + 1)].detach().to(torch.float32).contiguous()
                    beta = state["%s.%d.bias" % (name, 4 * j + 1)].detach().to(torch.float32).contiguous()
                lst.append((conv, gamma, beta))
            self.stages[name] = lst
        self.merges = []
        for k, merge in enumerate(MERGES):
            w = state[merge + ".weight"]
            c_up = inter_channels[k]                     # channels of the upsampled (low-res) half come first
            self.merges.append((pack_conv_weight(w, planes, 0, c_up),
                                pack_conv_weight(w, planes, c_up, w.shape[1] - c_up)))


def run_trunk_and_outputs(weights, feats_32_16_8_4, num_frames, num_groups, eps, planes, out_spec, trace=None):
    """Forward of one head; returns the channels-first output tensor [N, J, T, H/4, W/4]."""
    pools, tscale = pool_schedule(num_frames)
    if len(feats_32_16_8_4) != 4:
        raise AssertionError("Expected 4 feature maps, got {}".format(len(feats_32_16_8_4)))
    branch = []
    for (name, n_stages), feat in zip(BLOCKS, feats_32_16_8_4):
        if feat.shape[2] != num_frames:
            raise ValueError("feature map has T=%d but the head was built for NUM_FRAMES=%d" % (
                feat.shape[2], num_frames))
        a = pack_activation(feat, planes)
        for j in range(n_stages):
            conv, gamma, beta = weights.stages[name][j]
            y = conv3d(a, conv, allow_split=True)
            if trace is not None:
                trace["%s.%d.conv" % (name, 4 * j)] = y.sum(0) if y.dim() == 6 else y
            a = group_norm_relu_pool(y, gamma, beta, num_groups, eps, pools[j] and name != "block_4x", planes)
        branch.append(a)
    x = branch[0]
    out = None
    for k in range(3):
        w_up, w_skip = weights.merges[k]
        y_low = conv3d(x, w_up)                    # W_a . x at the low resolution
        z = conv3d(branch[k + 1], w_skip)          # W_b . f'
        if k < 2:
            x = upsample_add(z, y_low, tscale[k], planes)
        else:
            out = head_output(z, y_low, tscale[k], out_spec)
    return out
