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
PLANES_FP16 = 17        # include/stemseg_b200.h STEMSEG_PLANES_FP16: one fp16 plane, one tensor-core product per MAC

# fp32-parity mode: scale blocks whose 3x3x3 convolutions run as ONE fp16 product per MAC instead of three bf16 products.
# OFF by default.  The CPU emulation (profiles/r02_precision_ablation.json, scripts/precision_ablation.py) puts
# block_8x + block_16x in fp16 at 2.6-4.5e-5 of the 1e-4 budget on the full-width 8-frame head (the 4x layer and the 1x1
# merges need the three products: 2-4e-4 otherwise), and the step gets 16 % faster (2.79 -> 2.40 ms per 8x480x864 clip).
# Measured on the B200 over every golden and under the reference's own callers, however, the PER-CHANNEL bound leaves no
# margin: 1.0e-4 on the 2-frame golden (no pooling, nothing averages the operand rounding) and 9.7e-5 on the free-dimension
# channels of a random-init model (profiles/r02_fp16_blocks_golden_errors.txt, r02_fp16_blocks_reference_errors.txt; the
# all-bf16x3 plan: <= 1.1e-5 and 2.9e-5).  So parity mode keeps three products everywhere and the fp16 blocks are an
# opt-in: STEMSEG_FP32_FAST_BLOCKS=block_8x,block_16x or decoder.set_fast_blocks(...) before the heads are packed.
import os as _os
FP32_FAST_BLOCKS = tuple(b for b in _os.environ.get("STEMSEG_FP32_FAST_BLOCKS", "").split(",") if b)


def set_fast_blocks(blocks):
    """Choose the scale blocks that run single fp16 products in fp32-parity mode (takes effect for heads packed / plans
    built afterwards; call head.invalidate_packed_weights() on existing heads)."""
    global FP32_FAST_BLOCKS
    FP32_FAST_BLOCKS = tuple(blocks)


def plane_count(planes):
    return 2 if planes == 2 else 1


def block_planes(planes, block_name, exact=False):
    """Operand format of the 3x3x3 convolutions of one scale block under head precision `planes`."""
    if planes == 2 and not exact and block_name in FP32_FAST_BLOCKS:
        return PLANES_FP16
    return planes

# bench.py sets this to a list to collect (shape, start_event, end_event) around every conv launch
PROFILE_EVENTS = None


def pool_schedule(num_frames):
    if num_frames not in POOL_SLOTS:
        raise NotImplementedError("NUM_FRAMES=%r is not supported by the reference decoder tables "
                                  "(stemseg/modeling/common.py:15-35)" % (num_frames,))
    return POOL_SLOTS[num_frames], TEMPORAL_SCALES[num_frames]


class PackedConv(object):
    """Weights of one convolution in kernel layout: bf16 planes [P][rows][taps*cin] (+ fp32 bias)."""

    def __init__(self, planes_tensor, bias, cin, cout, kernel_size, planes=None):
        self.planes_tensor, self.bias, self.cin, self.cout, self.kernel_size = planes_tensor, bias, cin, cout, kernel_size
        self.planes = planes_tensor.shape[0] if planes is None else planes        # format code (1, 2 or PLANES_FP16)


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
        dst = torch.empty((plane_count(planes), cout, taps * cin_count), dtype=torch.bfloat16, device=w.device)
        _check(lib.stemseg_pack_conv_weight(_lib.ptr(w), cout, cin_total, cin_begin, cin_count, taps, _lib.ptr(dst), 0,
                                            cout, planes, _lib.stream_ptr()))
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(dst, b, cin_count, cout, 3 if taps == 27 else 1, planes)


class Planes(object):
    """An NDHWC activation stored as bf16 planes [P][n][t][h][w][c]."""

    def __init__(self, tensor, n, t, h, w, c, planes=None):
        self.tensor, self.n, self.t, self.h, self.w, self.c = tensor, n, t, h, w, c
        self._planes = tensor.shape[0] if planes is None else planes

    @property
    def planes(self):
        """Format code: 1 (bf16), 2 (bf16 hi + lo) or PLANES_FP16 (one fp16 plane in the same 2-byte storage)."""
        return self._planes


def pack_activation(x, planes, out=None):
    """[N,C,T,H,W] fp32 CUDA tensor (H,W contiguous; other strides free) -> Planes (into `out` if given)."""
    lib = _lib.load()
    if x.dim() != 5:
        raise ValueError("expected a [N,C,T,H,W] feature map, got shape %s" % (tuple(x.shape),))
    if x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("decoder inputs must be fp32 CUDA tensors (got %s on %s); there is no CPU path" % (
            x.dtype, x.device))
    n, c, t, h, w = x.shape
    # channels-last producer (a torch backbone run in channels_last emits [T,C,H,W] tensors whose memory is already
    # T,H,W,C): the NCTHW view of it IS the NDHWC layout of the planes, so D0 degenerates to the elementwise split
    ndhwc = (c > 1 and x.stride(1) == 1 and x.stride(4) == c and x.stride(3) == w * c and x.stride(2) == h * w * c
             and (n == 1 or x.stride(0) == t * h * w * c) and (n * c * t * h * w) % 4 == 0 and x.data_ptr() % 16 == 0)
    if not ndhwc and not (x.stride(4) == 1 and x.stride(3) == w):
        x = x.contiguous()
    with torch.cuda.device(x.device):
        if out is not None:
            if tuple(out.tensor.shape) != (plane_count(planes), n, t, h, w, c) or out.tensor.device != x.device or \
                    out.planes != planes:
                raise ValueError("pack_activation: static buffer %s does not match input %s" % (
                    tuple(out.tensor.shape), (plane_count(planes), n, t, h, w, c)))
            dst = out.tensor
        else:
            dst = torch.empty((plane_count(planes), n, t, h, w, c), dtype=torch.bfloat16, device=x.device)
        if ndhwc:
            _check(lib.stemseg_to_planes(_lib.ptr(x), n * c * t * h * w, _lib.ptr(dst), planes, _lib.stream_ptr()))
        else:
            _check(lib.stemseg_pack_activation(_lib.ptr(x), x.stride(0), x.stride(1), x.stride(2), n, c, t, h * w,
                                               _lib.ptr(dst), planes, _lib.stream_ptr()))
    return Planes(dst, n, t, h, w, c, planes)


# conv launches with at least this many tiles per SM run as short-lived CTAs (CHUNK_TILES consecutive tiles each)
# instead of one persistent CTA per SM, so that higher-priority branches of the CUDA graph can interleave
CHUNK_MIN_TILES_PER_SM = 4
CHUNK_TILES = 2


def conv3d(act, packed, max_ctas=0, allow_split=False, want_stats=False, chunked=False, out_bf16=False):
    """Planes x PackedConv -> fp32 NDHWC tensor [n,t,h,w,cout] (tcgen05 implicit GEMM).

    With allow_split the library may split the taps over several CTAs for layers with fewer tiles than SMs; the
    result is then [split_k, n, t, h, w, cout] partial sums (added by the GroupNorm kernels)."""
    lib = _lib.load()
    if act.c != packed.cin:
        raise ValueError("conv input has %d channels, weights expect %d" % (act.c, packed.cin))
    if act.planes != packed.planes:
        raise ValueError("activation / weight precision mismatch (%s vs %s)" % (act.planes, packed.planes))
    shape = _lib.StemsegConvShape(act.n, act.t, act.h, act.w, packed.cin, packed.cout, packed.kernel_size, act.planes,
                                  1, 0, 0)
    if chunked:
        tiles = lib.stemseg_conv3d_tiles_per_sample(shape) * act.n * max(1, packed.cout // 256)
        if tiles >= CHUNK_MIN_TILES_PER_SM * torch.cuda.get_device_properties(act.tensor.device).multi_processor_count:
            shape.tiles_per_cta = CHUNK_TILES
    stat = None
    with torch.cuda.device(act.tensor.device):
        if allow_split:
            shape.split_k = lib.stemseg_conv3d_auto_split(shape)
        if allow_split and shape.split_k > 1:
            out = torch.empty((shape.split_k, act.n, act.t, act.h, act.w, packed.cout), dtype=torch.float32,
                              device=act.tensor.device)
        else:
            # bf16 mode: the unsplit conv output can be stored as bf16 (the GroupNorm statistics still come from the
            # fp32 accumulators in the epilogue); split layers keep fp32 partial sums
            store_bf16 = bool(out_bf16 and want_stats)
            shape.out_bf16 = 1 if store_bf16 else 0
            out = torch.empty((act.n, act.t, act.h, act.w, packed.cout),
                              dtype=torch.bfloat16 if store_bf16 else torch.float32, device=act.tensor.device)
            if want_stats:     # per-tile channel sums straight from the accumulators (GroupNorm statistics)
                tiles = lib.stemseg_conv3d_tiles_per_sample(shape)
                stat = torch.empty((act.n, packed.cout, tiles, 2), dtype=torch.float32, device=act.tensor.device)
        ev = None
        if PROFILE_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _check(lib.stemseg_conv3d_bf16_planes(_lib.ptr(act.tensor), _lib.ptr(packed.planes_tensor),
                                              _lib.ptr(packed.bias), _lib.ptr(out), _lib.ptr(stat), shape, max_ctas,
                                              _lib.stream_ptr()))
        if ev is not None:
            ev[1].record()
            PROFILE_EVENTS.append(((act.n, act.t, act.h, act.w, packed.cin, packed.cout, packed.kernel_size,
                                    act.planes), ev[0], ev[1]))
    if want_stats:
        return out, stat
    return out


POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2


def group_norm_relu_pool(y, gamma, beta, num_groups, eps, pool, planes, channel_slice=None, stat=None,
                         saved=None):
    """fp32 NDHWC conv output ([split_k,] n,t,h,w,c_total) -> relu(GN(y)) [-> avgpool] as Planes.

    gamma None = no normalisation.  channel_slice=(c0, c) normalises channels [c0, c0+c) of a wider (multi-head)
    conv output."""
    lib = _lib.load()
    slices = y.shape[0] if y.dim() == 6 else 1
    n, t, h, w, c_total = y.shape[-5:]
    c0, c = (0, c_total) if channel_slice is None else channel_slice
    dev = y.device
    y_bf16 = y.dtype == torch.bfloat16
    x_ptr = _lib.c_void_p(y.data_ptr() + (2 if y_bf16 else 4) * c0)
    with torch.cuda.device(dev):
        scale_shift = None
        mean_rstd_out = None
        if gamma is not None and saved is not None:          # training: keep the group statistics for the backward
            mean_rstd_out = torch.empty((n, num_groups, 2), dtype=torch.float32, device=dev)
        if gamma is not None:
            if c % num_groups != 0:
                raise ValueError("channels %d not divisible by %d groups" % (c, num_groups))
            scale_shift = torch.empty((n, c, 2), dtype=torch.float32, device=dev)
            if stat is None and y_bf16:
                raise ValueError("a bf16 conv output needs the epilogue statistics")
            if stat is not None:          # statistics came out of the conv epilogue: finalize only
                chunks = stat.shape[2]
                part_ptr = _lib.c_void_p(stat.data_ptr() + 4 * c0 * chunks * 2)
                _check(lib.stemseg_group_norm_finalize(part_ptr, c_total * chunks * 2, chunks, n, t * h * w, c,
                                                       c // num_groups, float(eps), _lib.ptr(gamma), _lib.ptr(beta),
                                                       _lib.ptr(scale_shift), _lib.ptr(mean_rstd_out),
                                                       _lib.stream_ptr()))
                KEEP.append(scale_shift)
            else:
                ws_bytes = lib.stemseg_group_norm_workspace_bytes(n, t * h * w, c)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _check(lib.stemseg_group_norm_stats(x_ptr, c_total, slices, n, t * h * w, c, c // num_groups,
                                                    float(eps), _lib.ptr(gamma), _lib.ptr(beta),
                                                    _lib.ptr(scale_shift), _lib.ptr(mean_rstd_out), _lib.ptr(ws),
                                                    ws_bytes, _lib.stream_ptr()))
                KEEP.extend((ws, scale_shift))
            slices = 1                               # the statistics pass summed the split-K slices into slice 0
        t_out = (t - 1) // 2 + 1 if pool else t
        dst = torch.empty((plane_count(planes), n, t_out, h, w, c), dtype=torch.bfloat16, device=dev)
        apply = lib.stemseg_norm_relu_pool_bf16in if y_bf16 else lib.stemseg_norm_relu_pool
        _check(apply(x_ptr, c_total, slices, _lib.ptr(scale_shift), n, t, h, w, c, int(pool), _lib.ptr(dst), planes,
                     _lib.stream_ptr()))      # pool: True == POOL_AVG
    if saved is not None:
        saved["scale_shift"], saved["mean_rstd"] = scale_shift, mean_rstd_out
    return Planes(dst, n, t_out, h, w, c, planes)


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
        # code tables may arrive as cached device tensors (no host->device copy: safe under CUDA-graph capture)
        self.activation = activation if torch.is_tensor(activation) else \
            torch.tensor(list(activation), dtype=torch.int32, device=dev)
        self.coordinate = coordinate if torch.is_tensor(coordinate) else \
            torch.tensor(list(coordinate), dtype=torch.int32, device=dev)
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


def head_output_x(x, spec):
    """Training path: output heads over the merged fp32 feature x [n,t,h,w,c] (see csrc/head_train.cu)."""
    lib = _lib.load()
    n, t, h, w, c = x.shape
    if spec.weight.shape[1] != c:
        raise ValueError("output conv expects %d channels, got %d" % (spec.weight.shape[1], c))
    with torch.cuda.device(x.device):
        out = torch.empty((n, spec.n_out, t, h, w), dtype=torch.float32, device=x.device)
        _check(lib.stemseg_head_output_x(_lib.ptr(x), n, t, h, w, c, _lib.ptr(spec.weight), _lib.ptr(spec.bias),
                                         _lib.ptr(spec.activation), _lib.ptr(spec.coordinate), spec.n_out,
                                         spec.time_scale, _lib.ptr(out), _lib.stream_ptr()))
        _lib.KERNEL_LAUNCHES[0] += (spec.n_out - 1) // 8           # one launch per 8 outputs
    return out


def fused_merge_head_output(act, packed, y_low, t_scale, spec, max_ctas=0):
    """Last merge + output heads in one launch: conv1x1(act) stays in TMEM, its epilogue applies the output convs.

    act: Planes of f4' (the skip half of conv_4), packed: W_b, y_low: fp32 [n,tl,hl,wl,c3] = W_a . x_8."""
    lib = _lib.load()
    n, tl, hl, wl, c = y_low.shape
    if (act.n, act.t, act.h, act.w) != (n, tl * t_scale, 2 * hl, 2 * wl) or packed.cout != c:
        raise ValueError("fused_merge_head_output: low-res %s does not upsample by (%d,2,2) to %s" % (
            tuple(y_low.shape), t_scale, (act.n, act.t, act.h, act.w, packed.cout)))
    shape = _lib.StemsegConvShape(act.n, act.t, act.h, act.w, packed.cin, packed.cout, 1, act.planes, 1, 0, 0)
    with torch.cuda.device(y_low.device):
        p_low = torch.empty((n, tl, hl, wl, spec.n_out), dtype=torch.float32, device=y_low.device)
        _check(lib.stemseg_head_lowres(_lib.ptr(y_low), n * tl * hl * wl, c, _lib.ptr(spec.weight), spec.n_out,
                                       _lib.ptr(p_low), _lib.stream_ptr()))
        out = torch.empty((n, spec.n_out, act.t, act.h, act.w), dtype=torch.float32, device=y_low.device)
        _check(lib.stemseg_conv1x1_head_output(_lib.ptr(act.tensor), _lib.ptr(packed.planes_tensor), shape,
                                               _lib.ptr(p_low), t_scale, _lib.ptr(spec.weight), _lib.ptr(spec.bias),
                                               _lib.ptr(spec.activation), _lib.ptr(spec.coordinate), spec.n_out,
                                               spec.time_scale, _lib.ptr(out), max_ctas, _lib.stream_ptr()))
    KEEP.append(p_low)
    return out


class TrunkWeights(object):
    """Kernel-layout weights of one head: per block a list of (PackedConv, gamma, beta); per merge (W_a, W_b)."""

    def __init__(self, state, inter_channels, planes, has_norm, exact=False):
        self.inter_channels = list(inter_channels)
        self.planes, self.exact = planes, exact
        self.stages = {}
        for name, n_stages in BLOCKS:
            lst = []
            for j in range(n_stages):
                conv = pack_conv_weight(state["%s.%d.weight" % (name, 4 * j)], block_planes(planes, name, exact),
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


class HeadSpec(object):
    """Everything the plan needs to know about one head."""

    def __init__(self, weights, out_spec, num_groups, eps, pool_mode=POOL_AVG):
        self.weights, self.out_spec, self.num_groups, self.eps = weights, out_spec, num_groups, eps
        self.pool_mode = pool_mode             # POOL_AVG / POOL_MAX for the slots that pool (cfg POOL_TYPE)


def _fuse_rows(convs):
    """Concatenate the packed weights of several heads along the GEMM N dimension (same input, same K)."""
    if len(convs) == 1:
        return convs[0]
    first = convs[0]
    planes_tensor = torch.cat([c.planes_tensor for c in convs], dim=1).contiguous()
    bias = None
    if first.bias is not None:
        bias = torch.cat([c.bias for c in convs], dim=0).contiguous()
    return PackedConv(planes_tensor, bias, first.cin, sum(c.cout for c in convs), first.kernel_size, first.planes)


def first_stage_groups(couts):
    """Which heads share one first-stage GEMM: lists of head indices.  One group when the concatenated width keeps
    the widest N tile (<= 256 or a multiple of 256); otherwise the heads whose own width is a multiple of 256 form
    one group and the rest another (or one launch each if their sum still does not tile)."""
    total = sum(couts)
    if total <= 256 or total % 256 == 0:
        return [list(range(len(couts)))]
    wide = [i for i, c in enumerate(couts) if c % 256 == 0]
    rest = [i for i in range(len(couts)) if i not in wide]
    rest_total = sum(couts[i] for i in rest)
    rest_groups = [rest] if (rest_total <= 256 or rest_total % 256 == 0) else [[i] for i in rest]
    return ([wide] if wide else []) + [g for g in rest_groups if g]


# intermediates of the plan currently being built / captured: kept alive so that no buffer is recycled while a
# parallel branch of the CUDA graph may still read it.  Thread-local: two pipelines may build / capture plans from
# different host threads.
import threading as _threading


class _KeepAlive(_threading.local):
    def __init__(self):
        self.items = []

    def append(self, x):
        self.items.append(x)

    def extend(self, xs):
        self.items.extend(xs)

    def reset(self):
        self.items = []

    def take(self):
        items, self.items = self.items, []
        return items


KEEP = _KeepAlive()


class HeadSet(object):
    """Launch plan of one or several heads that read the SAME feature pyramid (embedding + seediness (+ semseg)).

    * the first 3x3x3 conv of every block reads the shared FPN tensor, so the heads' weights are concatenated along
      the GEMM N dimension and the (dominant) im2col operand is streamed once for all heads;
    * after packing the inputs into static bf16 planes, the whole plan is captured into ONE CUDA graph whose four
      scale branches are independent sub-graphs (forked onto side streams during capture), so latency-bound small
      layers overlap with the large ones and per-launch host overhead disappears;
    * replaying a graph requires fixed addresses: inputs are packed into static plane buffers outside the graph and
      the outputs are cloned out of the graph's private pool.
    """

    def __init__(self, specs, num_frames, planes, use_graph=True):
        self.specs = list(specs)
        self.num_frames = num_frames
        self.planes = planes
        self.use_graph = use_graph
        self.fuse_output_heads = True       # conv_4 merge GEMM + output heads in one kernel (epilogue fusion)
        self.fuse_stats = True              # GroupNorm statistics from the conv epilogue (unsplit layers)
        self.chunk_long_layers = True       # long layers as short-lived CTAs + high-priority side branches
        self.bf16_conv_outputs = planes == 1    # bf16 mode: unsplit conv outputs stored as bf16 (half the HBM bytes)
        self.pools, self.tscale = pool_schedule(num_frames)
        self.exact = any(s.weights.exact for s in self.specs)      # e.g. max-pool heads: no single-product fp16 blocks
        if any(s.weights.exact != self.exact for s in self.specs):
            raise ValueError("linked heads must agree on exact / fast operand formats")
        # first conv of every block: heads are fused along the GEMM N dimension in GROUPS whose total width keeps the
        # widest N tile (a multiple of 256, or anything up to 256); e.g. embedding (128) + semseg (256) = 384 would fall
        # back to three N=128 tiles, so those two run as separate launches (N=256 and N=128) instead
        self.first_stage = {}
        for name, _ in BLOCKS:
            convs = [s.weights.stages[name][0][0] for s in self.specs]
            groups = first_stage_groups([c.cout for c in convs])
            self.first_stage[name] = [(_fuse_rows([convs[i] for i in g]), g) for g in groups]
        self._entries = _lib.LRUCache(4)      # captured plans per input shape (each owns a private graph pool)

    # ---- the plan itself (eager or under capture) -----------------------------------------------------------
    def _branch(self, name, n_stages, a_in, trace, post_stream=None):
        """All conv stages of one scale block for every head; returns per-head Planes.

        post_stream: stream (high priority) that takes over after the first-stage convolution -- used for block_4x,
        whose long GEMM stays on the normal-priority main stream while everything after it is latency-bound."""
        first = {}                                   # head index -> (conv output, statistics, channel offset)
        for fused, members in self.first_stage[name]:
            res = conv3d(a_in, fused, allow_split=True, want_stats=self.fuse_stats, chunked=self.chunk_long_layers,
                         out_bf16=self.bf16_conv_outputs)
            y, stat = res if self.fuse_stats else (res, None)
            KEEP.extend((y, stat))
            c0 = 0
            for hi in members:
                first[hi] = (y, stat, c0)
                c0 += self.specs[hi].weights.stages[name][0][0].cout
        if post_stream is not None:
            post_stream.wait_event(torch.cuda.current_stream().record_event())
            with torch.cuda.stream(post_stream):
                return self._branch_tail(name, n_stages, first, trace)
        return self._branch_tail(name, n_stages, first, trace)

    def _branch_tail(self, name, n_stages, first, trace):
        outs = []
        for hi, spec in enumerate(self.specs):
            conv, gamma, beta = spec.weights.stages[name][0]
            y, stat, c0 = first[hi]
            if trace is not None and hi == trace[0]:
                full = y.sum(0) if y.dim() == 6 else y.clone()
                trace[1]["%s.0.conv" % name] = full[..., c0:c0 + conv.cout]
            # operands of the block's later stages use the block's format; its last output feeds a merge (head format)
            fmt = block_planes(self.planes, name, self.exact)
            a = group_norm_relu_pool(y, gamma, beta, spec.num_groups, spec.eps,
                                     spec.pool_mode if (self.pools[0] and name != "block_4x") else POOL_NONE,
                                     fmt if n_stages > 1 else self.planes, channel_slice=(c0, conv.cout), stat=stat)
            KEEP.append(a.tensor)
            for j in range(1, n_stages):
                conv, gamma, beta = spec.weights.stages[name][j]
                res = conv3d(a, conv, allow_split=True, want_stats=self.fuse_stats, out_bf16=self.bf16_conv_outputs)
                yj, statj = res if self.fuse_stats else (res, None)
                KEEP.extend((yj, statj))
                if trace is not None and hi == trace[0]:
                    trace[1]["%s.%d.conv" % (name, 4 * j)] = yj.sum(0) if yj.dim() == 6 else yj.clone()
                a = group_norm_relu_pool(yj, gamma, beta, spec.num_groups, spec.eps,
                                         spec.pool_mode if self.pools[j] else POOL_NONE,
                                         fmt if j + 1 < n_stages else self.planes, stat=statj)
                KEEP.append(a.tensor)
            outs.append(a)
        return outs

    def _plan(self, in_planes, trace=None, streams=None):
        main = torch.cuda.current_stream()
        self._tail_stream = None
        branches = [None] * 4
        if streams is None:
            for b, (name, n_stages) in enumerate(BLOCKS):
                branches[b] = self._branch(name, n_stages, in_planes[b], trace)
        else:
            # Priorities (two steps are in flight on two graph instances, pipeline.SubclipPipeline.submit): only the
            # long first-stage GEMM of block_4x runs at normal priority on `main`; the three small scale branches and
            # EVERYTHING after the 4x GEMM (GroupNorm apply, merges, output heads and -- in the step graph -- compaction,
            # gather, clustering) run on high-priority streams.  Otherwise the block scheduler serves the other step's
            # 4x GEMM (thousands of queued CTAs, launched earlier) before this step's small tail kernels, and the two
            # steps run in lock-step: conv phase, conv phase, both tails (profiles/r02_timeline_cfg3_before.txt).
            tail = streams[3] if len(streams) > 3 else None
            fork = main.record_event()
            for b, (name, n_stages) in enumerate(BLOCKS):
                st = main if b == 3 else streams[b]
                if st is not main:
                    st.wait_event(fork)
                with torch.cuda.stream(st):
                    branches[b] = self._branch(name, n_stages, in_planes[b], trace,
                                               post_stream=tail if (b == 3 and tail is not None) else None)
            if tail is not None:
                for b in range(3):
                    tail.wait_event(streams[b].record_event())
                with torch.cuda.stream(tail):
                    outs = self._merge_all(branches, trace, tail, streams[:3])
                self._tail_stream = tail               # the caller continues the step on it and joins `main` at the end
                return outs
            for b in range(3):
                main.wait_event(streams[b].record_event())
        return self._merge_all(branches, trace, main, streams)

    def _merge_all(self, branches, trace, main, streams):
        def merge_chain(hi, spec):
            x = branches[0][hi]
            out = None
            for k in range(3):
                w_up, w_skip = spec.weights.merges[k]
                y_low = conv3d(x, w_up)                      # W_a . x at the low resolution
                KEEP.append(y_low)
                if k == 2 and self.fuse_output_heads and w_skip.cout <= 256:
                    out = fused_merge_head_output(branches[3][hi], w_skip, y_low, self.tscale[k], spec.out_spec)
                    continue
                z = conv3d(branches[k + 1][hi], w_skip)      # W_b . f'
                KEEP.append(z)
                if k < 2:
                    x = upsample_add(z, y_low, self.tscale[k], self.planes)
                    KEEP.append(x.tensor)
                else:
                    out = head_output(z, y_low, self.tscale[k], spec.out_spec)
            return out

        outputs = [None] * len(self.specs)
        if streams is None or len(self.specs) == 1:
            for hi, spec in enumerate(self.specs):
                outputs[hi] = merge_chain(hi, spec)
        else:
            # the heads' merge chains are independent: head 0 on the main stream, the others on side streams
            fork = main.record_event()
            used = []
            for hi, spec in enumerate(self.specs):
                st = main if hi == 0 else streams[(hi - 1) % len(streams)]
                if st is not main:
                    st.wait_event(fork)
                    used.append(st)
                with torch.cuda.stream(st):
                    outputs[hi] = merge_chain(hi, spec)
            for st in used:
                main.wait_event(st.record_event())
        return outputs

    # ---- entry point -------------------------------------------------------------------------------------------
    def run(self, feats_32_16_8_4, trace=None):
        if len(feats_32_16_8_4) != 4:
            raise AssertionError("Expected 4 feature maps, got {}".format(len(feats_32_16_8_4)))
        for f in feats_32_16_8_4:
            if f.dim() != 5 or f.shape[2] != self.num_frames:
                raise ValueError("feature map %s does not have T=NUM_FRAMES=%d" % (tuple(f.shape), self.num_frames))
        dev = feats_32_16_8_4[0].device
        key = tuple((tuple(f.shape), str(f.device)) for f in feats_32_16_8_4)
        if trace is not None or not self.use_graph:
            KEEP.reset()
            with torch.cuda.device(dev):
                in_planes = self.pack_inputs(feats_32_16_8_4)
                outs = self._plan(in_planes, trace=trace)
            KEEP.reset()
            return outs
        entry = self._entries.get(key)
        with torch.cuda.device(dev):
            if entry is None:
                entry = self._capture(feats_32_16_8_4, dev)
                self._entries.put(key, entry)
            self.pack_inputs(feats_32_16_8_4, out=entry["in_planes"])
            entry["graph"].replay()
            _lib.KERNEL_LAUNCHES[0] += entry["kernels"]
            return [o.clone() for o in entry["outputs"]]

    def pack_inputs(self, feats_32_16_8_4, out=None):
        """D0: the four feature maps -> operand planes in the format of the scale block that reads them."""
        return [pack_activation(f, block_planes(self.planes, name, self.exact), out=None if out is None else out[b])
                for b, ((name, _), f) in enumerate(zip(BLOCKS, feats_32_16_8_4))]

    def _capture(self, feats, dev):
        in_planes = self.pack_inputs(feats)
        # warm-up on a side stream (lazy module loading, cudaFuncSetAttribute, ...) before capturing
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            KEEP.reset()
            self._plan(in_planes)
            KEEP.reset()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        # high priority: the three small scale branches + the tail after the 4x GEMM (see _plan)
        streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(4)]
        KEEP.reset()
        before = _lib.KERNEL_LAUNCHES[0]
        with _lib.capture_guard(), torch.cuda.graph(graph):
            outputs = self._plan(in_planes, streams=streams)
            if self._tail_stream is not None:        # join the capture's origin stream
                torch.cuda.current_stream().wait_event(self._tail_stream.record_event())
        kernels = _lib.KERNEL_LAUNCHES[0] - before
        keep = KEEP.take()
        return {"graph": graph, "in_planes": in_planes, "outputs": outputs, "keep": keep, "kernels": kernels,
                "streams": streams}


def run_trunk_and_outputs(weights, feats_32_16_8_4, num_frames, num_groups, eps, planes, out_spec, trace=None):
    """Eager forward of one head (no graph); returns the channels-first output tensor [N, J, T, H/4, W/4]."""
    hs = HeadSet([HeadSpec(weights, out_spec, num_groups, eps)], num_frames, planes, use_graph=False)
    return hs.run(feats_32_16_8_4, trace=None if trace is None else (0, trace))[0]
