"""B200 drop-in for the reference's embedding loss (training path).

Mirrors ``EmbeddingLoss`` (stemseg/modeling/losses/embedding_loss.py:10-185): same constructor keywords
(``embedding_size``, ``weight_variance_smoothness``, ``weight_lovasz``, ``weight_regularization``,
``weight_seediness``, ``weight``, ``nbr_free_dims``, ``free_dim_stds``; case-insensitive like the reference's
``cfg.TRAINING.LOSSES.EMBEDDING.d()`` expansion, model_builder.py:294-298), same ``forward(embedding_map, targets,
output_dict)`` contract (populates ``optimization_losses`` / ``others`` with the keys of stemseg/utils/constants.py)
and the same assertions.  The arithmetic -- masked means, Gaussian probability maps over every voxel, the Lovasz
hinge with its full sort, seediness and smoothness terms AND their gradient -- runs in csrc/embedding_loss.cu in one
C call; torch autograd only carries the pre-computed gradient (scaled on the device by the incoming grad_output).
One sequence per call, like the heads' training backward (the reference trains with MAX_SAMPLES_PER_GPU = 1,
defaults.yaml:20).  No CPU / PyTorch fallback: non-CUDA inputs raise.
"""
import ctypes

import torch
import torch.nn as nn

from stemseg_b200 import _lib

# stemseg/utils/constants.py:15-47
LOSS_EMBEDDING = "embedding_loss"
LOSS_LOVASZ = "lovasz_loss"
LOSS_SEEDINESS = "seediness_loss"
LOSS_VARIANCE_SMOOTHNESS = "variance_smoothness_loss"
LOSS_SEMSEG = "semantic_segmentation_loss"
LOSS_FOREGROUND = "foreground"
OUTPUT_OPTIMIZATION_LOSSES = "optimization_losses"
OUTPUT_OTHERS = "others"

_BITONIC_CHUNK = 4096


def _kernel_count(voxels, n_instances):
    """Kernels one stemseg_embedding_loss call launches (csrc/embedding_loss.cu host code)."""
    n_pad = _BITONIC_CHUNK
    while n_pad < voxels:
        n_pad *= 2
    count = 5                                    # stats, prepare, accumulate, distribute, finalize
    if n_instances > 0:
        count += 4                               # prob, local sort, count, apply
        k = 2 * _BITONIC_CHUNK
        while k <= n_pad:
            j = k // 2
            while j >= _BITONIC_CHUNK:
                count += 1
                j //= 2
            count += 1
            k *= 2
    return count


def embedding_loss_and_gradient(embedding_map, masks, ignore, crit):
    """One C call: (losses [4] = total, lovasz, variance_smoothness, seediness; d total / d embedding_map).

    embedding_map [1, E+V+1, T, H, W] fp32 CUDA, contiguous; masks [I,T,H,W] / ignore [T,H,W] uint8 CUDA (or None)."""
    lib = _lib.load()
    e, v = crit.embedding_size, crit.embedding_size - crit.n_free_dims
    x = embedding_map
    if x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("EmbeddingLoss needs an fp32 CUDA embedding map (got %s on %s); there is no CPU path" % (
            x.dtype, x.device))
    x = x.contiguous()
    voxels = x.shape[2] * x.shape[3] * x.shape[4]
    n_inst = int(masks.shape[0])
    if n_inst > _lib.STEMSEG_MAX_LOSS_INSTANCES:
        raise ValueError("at most %d instances per sequence (got %d)" % (_lib.STEMSEG_MAX_LOSS_INSTANCES, n_inst))
    dev = x.device
    with torch.cuda.device(dev):
        m = masks.to(device=dev, dtype=torch.uint8).contiguous()
        ig = None if ignore is None else ignore.to(device=dev, dtype=torch.uint8).contiguous()
        grad = torch.empty_like(x)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        ws_bytes = lib.stemseg_embedding_loss_workspace_bytes(voxels, n_inst)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        stds = (ctypes.c_float * max(1, crit.n_free_dims))(*[float(s) for s in crit.free_dim_stds])
        base = x.data_ptr()
        seed_off = 4 * (e + v) * voxels
        _lib.check(lib.stemseg_embedding_loss(
            _lib.c_void_p(base), _lib.c_void_p(base + seed_off), _lib.ptr(m), _lib.ptr(ig), voxels, n_inst, e,
            crit.n_free_dims, stds, crit.w_lovasz, crit.w_variance_smoothness, crit.w_seediness, crit.w,
            _lib.ptr(losses), _lib.c_void_p(grad.data_ptr()), _lib.c_void_p(grad.data_ptr() + seed_off),
            _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        _lib.KERNEL_LAUNCHES[0] += _kernel_count(voxels, n_inst)
    return losses, grad


class _EmbeddingLossFunction(torch.autograd.Function):
    """(embedding_map [1,C,T,H,W]) -> losses [4]; the gradient was computed together with the loss."""

    @staticmethod
    def forward(ctx, embedding_map, masks, ignore, crit):
        losses, ctx.grad = embedding_loss_and_gradient(embedding_map.detach(), masks, ignore, crit)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        # only losses[0] (the weighted total) is meant to be optimised (model_output_manager sums
        # optimization_losses); the chain-rule factor is applied on the device without a host round trip
        lib = _lib.load()
        grad, ctx.grad = ctx.grad, None
        g = grad_losses.detach().to(torch.float32).contiguous()
        with torch.cuda.device(grad.device):
            _lib.check(lib.stemseg_scale_by_device_scalar(_lib.ptr(grad), grad.numel(), _lib.ptr(g),
                                                          _lib.stream_ptr()))
        return grad, None, None, None


class EmbeddingLoss(nn.Module):
    def __init__(self, embedding_map_scale, **kwargs):
        super().__init__()
        kwargs = {k.lower(): v for k, v in kwargs.items()}
        self.embedding_map_scale = embedding_map_scale
        self.embedding_size = kwargs["embedding_size"]
        self.w_variance_smoothness = float(kwargs["weight_variance_smoothness"])
        self.w_lovasz = float(kwargs["weight_lovasz"])
        self.w_regularization = kwargs.get("weight_regularization", 0.0)      # read but unused by the reference too
        self.w_seediness = float(kwargs["weight_seediness"])
        self.w = float(kwargs["weight"])
        self.n_free_dims = kwargs["nbr_free_dims"]
        self.free_dim_stds = list(kwargs["free_dim_stds"])
        assert len(self.free_dim_stds) == self.n_free_dims, \
            "List of std values {} does not match number of free dims {}".format(len(self.free_dim_stds),
                                                                                 self.n_free_dims)
        if self.n_free_dims > 0:      # embedding_loss.py:28-29 (kept for state_dict compatibility)
            self.register_buffer("free_dim_bandwidths",
                                 1. / torch.tensor(self.free_dim_stds).float().unsqueeze(0) ** 2)
        self.split_sizes = (self.embedding_size, self.embedding_size - self.n_free_dims, 1)
        self.num_input_channels = sum(self.split_sizes)

    def forward(self, embedding_map, targets, output_dict, *args, **kwargs):
        """embedding_map [1, E+V+1, T, H, W]; targets: list (length 1) of dicts with 'masks' [I,T,H,W] and
        'ignore_masks' [T,H,W] at the embedding resolution.  Populates output_dict like the reference."""
        assert embedding_map.shape[1] == self.num_input_channels, "Expected {} channels in input tensor, got {}".format(
            self.num_input_channels, embedding_map.shape[1])
        if embedding_map.shape[0] != 1 or len(targets) != 1:
            raise NotImplementedError("the B200 embedding loss handles one sequence per call (batch 1), like the "
                                      "reference's MAX_SAMPLES_PER_GPU = 1")
        masks = targets[0]["masks"]
        ignore = targets[0].get("ignore_masks")
        if masks.numel() > 0:
            assert masks.shape[-2:] == embedding_map.shape[-2:], \
                "Masks tensor has shape {} while embedding map has shape {}".format(masks.shape, embedding_map.shape)
            if ignore is not None:
                assert masks.shape[-2:] == ignore.shape[-2:], \
                    "Masks tensor has shape {} while ignore mask has shape {}".format(masks.shape, ignore.shape)
        else:
            masks = masks.reshape((0,) + tuple(embedding_map.shape[2:]))
        losses = _EmbeddingLossFunction.apply(embedding_map, masks, ignore, self)
        output_dict[OUTPUT_OPTIMIZATION_LOSSES] = {LOSS_EMBEDDING: losses[0]}
        output_dict[OUTPUT_OTHERS] = {LOSS_LOVASZ: losses[1].detach(), LOSS_VARIANCE_SMOOTHNESS: losses[2].detach(),
                                      LOSS_SEEDINESS: losses[3].detach()}
        return losses[0]


# ----------------------------------------------------------------------------------------------------------------------
# semantic-segmentation head losses (YouTube-VIS / KITTI-MOTS configs)
# ----------------------------------------------------------------------------------------------------------------------
def semseg_loss_and_gradient(class_logits, fg_logits, class_ids, ignore, w_semseg=1.0, w_foreground=1.0,
                             grad_out=None):
    """One C call (csrc/semseg_loss.cu).  class_logits [T,cls,H,W] or None, fg_logits [T,H,W] or None, class_ids
    [T,H,W] int64, ignore [T,H,W] bool/uint8 or None -- fp32 CUDA.  Returns (losses [2] = class loss, foreground loss;
    d(w_semseg*class + w_foreground*fg)/d class_logits (same shape / strides as a [T,cls,H,W] view of a channels-first
    buffer) or None; d/d fg_logits or None).  grad_out: optional contiguous [cls(+1),T,H,W] tensor that receives both
    gradients (class channels first, then the foreground channel) -- the layout the head's backward consumes."""
    lib = _lib.load()
    ref = class_logits if class_logits is not None else fg_logits
    if ref is None:
        raise ValueError("neither class nor foreground logits given")
    if ref.dtype != torch.float32 or not ref.is_cuda:
        raise ValueError("the semseg losses need fp32 CUDA logits (got %s on %s); there is no CPU path" % (
            ref.dtype, ref.device))
    dev = ref.device
    with torch.cuda.device(dev):
        ids = class_ids.to(device=dev, dtype=torch.int64).contiguous()
        voxels = ids.numel()
        ig = None if ignore is None else ignore.to(device=dev, dtype=torch.uint8).contiguous()
        cls_ptr, cls_stride, n_cls, d_cls, d_cls_view = None, 0, 0, None, None
        if class_logits is not None:
            t, n_cls, h, w = class_logits.shape
            x = class_logits
            if not (x.stride(3) == 1 and x.stride(2) == w and x.stride(0) == h * w and x.stride(1) >= voxels):
                x = x.permute(1, 0, 2, 3).contiguous().permute(1, 0, 2, 3)       # channels-first storage
            cls_ptr, cls_stride = x, x.stride(1)
            d_cls = grad_out[:n_cls] if grad_out is not None else \
                torch.empty((n_cls, t, h, w), dtype=torch.float32, device=dev)
            d_cls_view = d_cls.permute(1, 0, 2, 3)
        fg, d_fg = None, None
        if fg_logits is not None:
            fg = fg_logits.contiguous()
            d_fg = grad_out[n_cls] if grad_out is not None else torch.empty_like(fg)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        ws_bytes = lib.stemseg_semseg_loss_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.stemseg_semseg_loss(_lib.ptr(cls_ptr), cls_stride, n_cls, _lib.ptr(fg), _lib.ptr(ids), _lib.ptr(ig),
                                           voxels, float(w_semseg), float(w_foreground), _lib.ptr(losses), _lib.ptr(d_cls),
                                           voxels, _lib.ptr(d_fg), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    return losses, d_cls_view, d_fg


class _SemsegLossFunction(torch.autograd.Function):
    """(logits) -> losses [2]; term = 0: class logits [T,cls,H,W], term = 1: foreground logits [T,H,W]."""

    @staticmethod
    def forward(ctx, logits, class_ids, ignore, term):
        if term == 0:
            losses, ctx.grad, _ = semseg_loss_and_gradient(logits.detach(), None, class_ids, ignore)
        else:
            losses, _, ctx.grad = semseg_loss_and_gradient(None, logits.detach(), class_ids, ignore)
        ctx.term = term
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        lib = _lib.load()
        grad, ctx.grad = ctx.grad, None
        g = grad_losses.detach().to(torch.float32)[ctx.term:ctx.term + 1].contiguous()
        base = grad if grad.is_contiguous() else grad.permute(1, 0, 2, 3)      # the channels-first storage
        with torch.cuda.device(grad.device):
            _lib.check(lib.stemseg_scale_by_device_scalar(_lib.ptr(base), base.numel(), _lib.ptr(g), _lib.stream_ptr()))
        return grad, None, None, None


def _one_sequence(logits, targets):
    if logits.shape[0] != 1 or len(targets) != 1:
        raise NotImplementedError("the B200 semseg losses handle one sequence per call (batch 1), like the reference's "
                                  "MAX_SAMPLES_PER_GPU = 1")
    gt, ignore = targets[0]["semseg_masks"], targets[0]["ignore_masks"]
    assert gt.shape[-2:] == logits.shape[-2:], \
        "Shape mismatch between ground truth semseg masks {} and predicted semseg masks {}".format(gt.shape, logits.shape)
    assert gt.shape[-2:] == ignore.shape[-2:], \
        "Shape mismatch between ground truth semseg masks {} and ignore masks {} ".format(gt.shape, ignore.shape)
    return gt, ignore


class CrossEntropyLoss(nn.Module):
    """Mirror of stemseg/modeling/losses/cross_entropy.py:9-49 (registered as SEMSEG_LOSS_REGISTRY["CrossEntropy"],
    model_builder.py:26): forward(semseg_logits [N,T,cls,H,W], targets, output_dict)."""

    def __init__(self, weight_semseg=None):
        super().__init__()
        self._weight = weight_semseg

    def _weight_semseg(self):
        if self._weight is not None:
            return float(self._weight)
        try:
            from stemseg.config import cfg                     # cross_entropy.py:49 reads it at call time
            return float(cfg.TRAINING.LOSSES.WEIGHT_SEMSEG)
        except ImportError:
            return 1.0

    def forward(self, semseg_logits, targets, output_dict):
        gt, ignore = _one_sequence(semseg_logits, targets)
        losses = _SemsegLossFunction.apply(semseg_logits[0], gt, ignore, 0)
        output_dict.setdefault(OUTPUT_OTHERS, {})[LOSS_SEMSEG] = losses[0]
        output_dict.setdefault(OUTPUT_OPTIMIZATION_LOSSES, {})[LOSS_SEMSEG] = losses[0] * self._weight_semseg()
        return losses[0]


def compute_fg_loss(fg_logits, targets, output_dict):
    """Mirror of TrainingModel.compute_fg_loss (model_builder.py:210-244): fg_logits [N,T,H,W]."""
    gt, ignore = _one_sequence(fg_logits, targets)
    losses = _SemsegLossFunction.apply(fg_logits[0], gt, ignore, 1)
    output_dict.setdefault(OUTPUT_OPTIMIZATION_LOSSES, {})[LOSS_FOREGROUND] = losses[1]
    return losses[1]
