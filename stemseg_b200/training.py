"""Data-parallel training step of the decoder heads (BASELINE configs[4]; SURVEY.md §3.4, §8e training, §8f rank 3).

What the reference does per iteration (stemseg/training/main.py:187-216): DDP forward -> EmbeddingLoss -> backward with
bucketed NCCL all-reduce -> torch.optim.SGD(momentum 0.9, nesterov, weight decay 1e-4).step() -> zero_grad.  Here the
same step for the heads' parameters is laid out for one process per B200:

  * every head parameter (and its gradient and momentum) is a VIEW into one flat fp32 buffer per head
    (``FlatParameters``), so a head's gradient is one contiguous NCCL message (43 MB at the shipped widths) and the
    optimiser is one fused kernel pass over the buffer (csrc/optim.cu) instead of ~60 per-tensor launches;
  * gradients are exchanged per head: the moment the last parameter of a head has received its gradient (post-
    accumulate hooks) its flat buffer is all-reduced asynchronously on NCCL's stream, overlapping the backward pass
    of the head that autograd runs next (``GradientExchange``; gloo in the CPU tests);
  * the mean over ranks is folded into the SGD pass (grad_scale = 1 / world_size).

Forward/backward of the heads: stemseg_b200.autograd (tcgen05 conv / dgrad / wgrad kernels); loss + its gradient:
stemseg_b200.losses.EmbeddingLoss (csrc/embedding_loss.cu).  torch supplies autograd bookkeeping, streams, NCCL.
The torch ResNet-101 backbone is out of scope: its parameters, if trained, stay with torch's optimiser;
the feature gradients this step produces are what it needs.
"""
import torch
import torch.distributed as dist

from stemseg_b200 import _lib


class FlatParameters(object):
    """Parameters of one module re-seated as views into a flat fp32 buffer, with flat gradient and momentum twins."""

    def __init__(self, module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev, dtype = params[0].device, params[0].dtype
        if dtype != torch.float32 or any(p.device != dev or p.dtype != dtype for p in params):
            raise ValueError("FlatParameters needs fp32 parameters on one device")
        self.module = module
        self.params = params
        # 16-byte aligned slices (float4 access in the fused optimiser; NCCL likes it too)
        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.data = torch.zeros(total, dtype=dtype, device=dev)
        self.grad = torch.zeros(total, dtype=dtype, device=dev)
        self.momentum = torch.zeros(total, dtype=dtype, device=dev)
        with torch.no_grad():
            for p, off in zip(params, offsets):
                view = self.data[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.grad[off:off + p.numel()].view_as(p)
        self.offsets = offsets

    def zero_grad(self):
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):          # re-seat in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view_as(p)


class GradientExchange(object):
    """All-reduce (sum) of each FlatParameters' gradient as soon as it is complete, asynchronously."""

    def __init__(self, flats, group=None):
        self.flats = list(flats)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending = []
        self._remaining = [0] * len(self.flats)
        self._handles = []
        for fi, flat in enumerate(self.flats):
            for p in flat.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(fi)))
        self.reset()

    def _make_hook(self, fi):
        def hook(_param):
            self._remaining[fi] -= 1
            if self._remaining[fi] == 0:
                self._launch(fi)
        return hook

    def reset(self):
        self._remaining = [len(f.params) for f in self.flats]
        self._pending = []

    def _launch(self, fi):
        if self.world > 1:
            self._pending.append(dist.all_reduce(self.flats[fi].grad, op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def finish(self):
        """Flush buffers whose hooks did not all fire (unused parameters), then wait for every reduction."""
        for fi, rem in enumerate(self._remaining):
            if rem > 0:
                self._launch(fi)
        for work in self._pending:
            work.wait()
        self.reset()

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def sgd_step(flat, lr, momentum, weight_decay, nesterov, grad_scale=1.0):
    """One fused pass over the flat buffers (csrc/optim.cu).  CUDA only."""
    if not flat.data.is_cuda:
        raise ValueError("the fused SGD step is a CUDA kernel; there is no CPU path")
    lib = _lib.load()
    with torch.cuda.device(flat.data.device):
        _lib.check(lib.stemseg_sgd_step(_lib.ptr(flat.data), _lib.ptr(flat.grad), _lib.ptr(flat.momentum), flat.numel,
                                        float(lr), float(momentum), float(weight_decay), float(grad_scale),
                                        1 if nesterov else 0, _lib.stream_ptr()))


class DecoderTrainer(object):
    """forward -> loss -> backward -> gradient exchange -> SGD for the decoder heads, one sub-clip per rank per step.

    heads: dict with 'embedding' (EmbeddingHead) and optionally 'seediness' (SeedinessHead), already on the device.
    criterion: stemseg_b200.losses.EmbeddingLoss.  Hyper-parameters default to defaults.yaml:17-31."""

    def __init__(self, heads, criterion, lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True, group=None):
        self.embedding_head = heads["embedding"]
        self.seediness_head = heads.get("seediness")
        self.criterion = criterion
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        mods = [self.embedding_head] + ([self.seediness_head] if self.seediness_head is not None else [])
        for m in mods:
            m.train()
        self.flats = [FlatParameters(m) for m in mods]
        self.exchange = GradientExchange(self.flats, group=group)
        self.world = self.exchange.world
        if self.world > 1:      # identical starting point on every rank, like DistributedDataParallel's constructor
            for flat in self.flats:
                dist.broadcast(flat.data, src=0, group=group)

    def forward_loss(self, feats_32_16_8_4, targets):
        out = self.embedding_head(feats_32_16_8_4)
        if self.seediness_head is not None:         # model_builder.py:198-201: cat(embedding head, seediness head)
            out = torch.cat((out, self.seediness_head(feats_32_16_8_4)), dim=1)
        output = {}
        loss = self.criterion(out, targets, output)
        return loss, output

    def step(self, feats_32_16_8_4, targets):
        """One optimisation step; returns the loss dict (device scalars, no host synchronisation)."""
        for flat in self.flats:
            flat.zero_grad()
        loss, output = self.forward_loss(feats_32_16_8_4, targets)
        loss.backward()
        self.exchange.finish()
        for flat, mod in zip(self.flats, [self.embedding_head, self.seediness_head]):
            sgd_step(flat, self.lr, self.momentum, self.weight_decay, self.nesterov, 1.0 / self.world)
            mod.invalidate_packed_weights()          # the kernel wrote the parameters behind autograd's back
        return output
