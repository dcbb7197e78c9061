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
