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

Two execution modes of ``DecoderTrainer.step``:
  * ``use_graph=False``: through torch autograd (``loss.backward()``), i.e. exactly what the reference's training loop
    drives when the B200 heads / loss are installed into it -- ~500 kernel launches per step issued from Python;
  * ``use_graph=True`` (default on CUDA): the same kernels issued WITHOUT the autograd engine (forward, loss+gradient,
    backward of each head, optimiser) and captured once per input shape into three CUDA graphs -- [weights repack +
    forward of all heads + loss + backward of the last head] / [backward of the first head + feature-gradient sum] /
    [fused SGD] -- so a step is 4 pack launches + 3 graph replays, with the per-head NCCL all-reduce issued between
    the replays (the first reduction overlaps the second graph).

Forward/backward of the heads: stemseg_b200.autograd (tcgen05 conv / dgrad / wgrad kernels); loss + its gradient:
stemseg_b200.losses.EmbeddingLoss (csrc/embedding_loss.cu).  torch supplies autograd bookkeeping, streams, NCCL.
The torch ResNet-101 backbone is out of scope: its parameters, if trained, stay with torch's optimiser;
the feature gradients this step produces are what it needs.
"""
import torch
import torch.distributed as dist

from stemseg_b200 import _lib
from stemseg_b200._lib import LRUCache


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
        # gradients of block_32x / block_16x (complete early in the backward pass) form a contiguous prefix
        names = [n for n, p in module.named_parameters() if p.requires_grad]
        self.prefix_end = total
        for n, off in zip(names, offsets):
            if n.startswith("block_8x."):
                self.prefix_end = off
                break

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


def sgd_step_dev(flat, hyper, nesterov):
    """Same pass with {lr, momentum, weight_decay, grad_scale} read from the device array `hyper` (float32 [4]) when the
    kernel runs -- the form captured into CUDA graphs, so that schedules keep working after capture."""
    lib = _lib.load()
    with torch.cuda.device(flat.data.device):
        _lib.check(lib.stemseg_sgd_step_dev(_lib.ptr(flat.data), _lib.ptr(flat.grad), _lib.ptr(flat.momentum), flat.numel,
                                            _lib.ptr(hyper), 1 if nesterov else 0, _lib.stream_ptr()))


class DecoderTrainer(object):
    """forward -> loss -> backward -> gradient exchange -> SGD for the decoder heads, one sub-clip per rank per step.

    heads: dict with 'embedding' (EmbeddingHead) and optionally 'seediness' (SeedinessHead) and 'semseg' (SemsegHead,
    YouTube-VIS / KITTI-MOTS configs: class cross-entropy + foreground BCE, csrc/semseg_loss.cu; targets then carry
    'semseg_masks'), already on the device.  criterion: stemseg_b200.losses.EmbeddingLoss.  Hyper-parameters default to
    defaults.yaml:17-34."""

    def __init__(self, heads, criterion, lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True, group=None,
                 use_graph=True, need_feature_grads=True, overlap_heads=True, weight_semseg=1.0, max_cached_shapes=4):
        self.use_graph = use_graph
        # graph mode with two heads: run the heads' forward (and, after the loss, their backward) concurrently on two
        # streams inside ONE graph -- the latency-bound small kernels of one head hide under the tensor-core
        # convolutions of the other.  False: three graphs, per-head all-reduce overlapped with the other head's backward.
        self.overlap_heads = overlap_heads
        # data-parallel runs: cut the backward pass in two graphs so that the all-reduce of the low-resolution blocks'
        # gradients (81 % of the bytes, complete after a small part of the backward time) overlaps block_8x / block_4x
        self.split_backward = None            # None = automatically when world > 1
        self.need_feature_grads = need_feature_grads
        self.group = group
        self.embedding_head = heads["embedding"]
        self.seediness_head = heads.get("seediness")
        self.semseg_head = heads.get("semseg")
        self.weight_semseg = float(weight_semseg)            # cfg.TRAINING.LOSSES.WEIGHT_SEMSEG (defaults.yaml:34)
        self.criterion = criterion
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        self._graphs = LRUCache(max_cached_shapes)
        self._hyper = None                    # device float32 [4] read by the captured SGD launches
        self._hyper_values = None
        mods = self._modules()
        for m in mods:
            m.train()
        self.flats = [FlatParameters(m) for m in mods]
        self.exchange = GradientExchange(self.flats, group=group)
        self.world = self.exchange.world
        if self.world > 1:      # identical starting point on every rank, like DistributedDataParallel's constructor
            for flat in self.flats:
                dist.broadcast(flat.data, src=0, group=group)

    def set_lr(self, lr):
        """Learning rate of the next step (the reference calls lr_scheduler.step() every iteration,
        training/main.py:209-210).  Takes effect in graph mode too: the captured SGD launches read it from device memory."""
        self.lr = float(lr)

    def _sync_hyper(self, device):
        """Upload {lr, momentum, weight_decay, 1/world} if they changed since the last step (tiny async H2D copy)."""
        values = (float(self.lr), float(self.momentum), float(self.weight_decay), 1.0 / self.world)
        if self._hyper is None:
            self._hyper = torch.empty(4, dtype=torch.float32, device=device)
        if values != self._hyper_values:
            # pageable source: the 16 bytes are staged by the driver before the call returns, and stream order makes
            # the device-side write wait for the previous step's SGD launch that may still be reading the array
            self._hyper.copy_(torch.tensor(values, dtype=torch.float32), non_blocking=True)
            self._hyper_values = values

    def forward_loss(self, feats_32_16_8_4, targets):
        """TrainingModel.forward after the backbone (model_builder.py:107-126) through torch autograd."""
        from stemseg_b200.losses import CrossEntropyLoss, compute_fg_loss
        out = self.embedding_head(feats_32_16_8_4)
        if self.seediness_head is not None:         # model_builder.py:198-201: cat(embedding head, seediness head)
            out = torch.cat((out, self.seediness_head(feats_32_16_8_4)), dim=1)
        output = {}
        loss = self.criterion(out, targets, output)
        if self.semseg_head is not None:
            logits = self.semseg_head(list(feats_32_16_8_4)[::-1]).permute(0, 2, 1, 3, 4)     # model_builder.py:179-180
            if self.semseg_head.has_foreground_channel:
                logits, fg_logits = logits.split((logits.shape[2] - 1, 1), dim=2)             # model_builder.py:121
                loss = loss + compute_fg_loss(fg_logits.squeeze(2), targets, output)
            loss = loss + CrossEntropyLoss(self.weight_semseg)(logits, targets, output) * self.weight_semseg
        return loss, output

    def step(self, feats_32_16_8_4, targets):
        """One optimisation step; returns the loss dict (device scalars, no host synchronisation).  In graph mode the
        dict also carries 'feature_grads' (4 tensors [1,C,T,h,w], d loss / d feature map, valid until the next step)."""
        if self.use_graph:
            return self._step_graph(feats_32_16_8_4, targets)
        for flat in self.flats:
            flat.zero_grad()
        loss, output = self.forward_loss(feats_32_16_8_4, targets)
        loss.backward()
        self.exchange.finish()
        for flat, mod in zip(self.flats, self._modules()):
            sgd_step(flat, self.lr, self.momentum, self.weight_decay, self.nesterov, 1.0 / self.world)
            mod.invalidate_packed_weights()          # the kernel wrote the parameters behind autograd's back
        return output

    # ---- graph mode ---------------------------------------------------------------------------------------------
    def _modules(self):
        return [m for m in (self.embedding_head, self.seediness_head, self.semseg_head) if m is not None]

    def _grad_slots(self, flat):
        names = {id(p): n for n, p in flat.module.named_parameters()}
        return {names[id(p)]: flat.grad[off:off + p.numel()].view_as(p) for p, off in zip(flat.params, flat.offsets)}

    def _segments(self, entry):
        """The captured segments (closures over the static buffers of `entry`): [both heads concurrently, optimiser]
        or, with overlap_heads=False, [forward + loss + backward of the last head, backward of the first, optimiser]."""
        from stemseg_b200 import autograd as A
        from stemseg_b200.losses import embedding_loss_and_gradient
        mods, flats = self._modules(), self.flats
        state = {}
        split = (self.world > 1) if self.split_backward is None else bool(self.split_backward)

        def seg_forward_loss_backward_last():
            saved, outs = [], []
            for m in mods:
                m.invalidate_packed_weights()                 # the repack kernels become part of the graph
                out, sv = A.training_forward(m, None, in_planes=entry["in_planes"])
                outs.append(out)
                saved.append(sv)
            emb_map = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)      # model_builder.py:198-201
            losses, grad = embedding_loss_and_gradient(emb_map, entry["masks"], entry["ignore"], self.criterion)
            entry["losses"] = losses
            c0 = outs[0].shape[1]
            state["grads_out"] = [grad[:, :c0]] + ([grad[:, c0:]] if len(outs) > 1 else [])
            state["saved"] = saved
            last = len(mods) - 1
            fg, _ = A.training_backward(mods[last], saved[last], state["grads_out"][last],
                                        grad_dst=self._grad_slots(flats[last]),
                                        need_feature_grads=self.need_feature_grads)
            state["fg_last"] = fg

        def seg_backward_first():
            fg = state["fg_last"]
            if len(mods) > 1:
                fg0, _ = A.training_backward(mods[0], state["saved"][0], state["grads_out"][0],
                                             grad_dst=self._grad_slots(flats[0]),
                                             need_feature_grads=self.need_feature_grads)
                if self.need_feature_grads:                   # both heads read the same pyramid
                    fg = [a + b for a, b in zip(fg0, fg)]
            entry["feature_grads"] = fg if self.need_feature_grads else None

        def seg_optimizer():
            for flat in flats:
                sgd_step_dev(flat, self._hyper, self.nesterov)       # hyper-parameters from device memory (set_lr)

        def seg_heads_concurrently():
            """Every head on its own stream: forward, join, losses + gradients on the main stream, fork, backward."""
            from stemseg_b200.losses import semseg_loss_and_gradient
            main = torch.cuda.current_stream()
            sides = entry["side_streams"][:len(mods) - 1]
            streams = [main] + sides

            def fork():
                ev = torch.cuda.Event()
                ev.record(main)
                for st in sides:
                    st.wait_event(ev)

            def join():
                for st in sides:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)

            outs, saved = [None] * len(mods), [None] * len(mods)
            fork()
            for k in reversed(range(len(mods))):              # side streams first, the main-stream head last
                with torch.cuda.stream(streams[k]):
                    mods[k].invalidate_packed_weights()       # the repack kernels become part of the graph
                    outs[k], saved[k] = A.training_forward(mods[k], None, in_planes=entry["in_planes"])
            join()
            n_emb = 2 if self.seediness_head is not None else 1
            emb_map = outs[0] if n_emb == 1 else torch.cat((outs[0], outs[1]), dim=1)      # model_builder.py:198-201
            losses, grad = embedding_loss_and_gradient(emb_map, entry["masks"], entry["ignore"], self.criterion)
            entry["losses"] = losses
            c0 = outs[0].shape[1]
            grads_out = [grad[:, :c0]] + ([grad[:, c0:]] if n_emb == 2 else [])
            if self.semseg_head is not None:
                sem = outs[-1]                                                           # [1, C, T, H, W]
                g_sem = torch.empty_like(sem)
                n_cls = sem.shape[1] - (1 if self.semseg_head.has_foreground_channel else 0)
                cls_view = sem[0, :n_cls].permute(1, 0, 2, 3)                            # [T, cls, H, W] (model_builder.py:180)
                fg = sem[0, n_cls] if self.semseg_head.has_foreground_channel else None
                entry["semseg_losses"], _, _ = semseg_loss_and_gradient(
                    cls_view, fg, entry["semseg_ids"], entry["ignore"], self.weight_semseg, 1.0, grad_out=g_sem[0])
                grads_out.append(g_sem)
            state["grads_out"], state["saved"], state["outs"] = grads_out, saved, outs
            if split:
                carries = [None] * len(mods)
                fork()
                for k in reversed(range(len(mods))):
                    with torch.cuda.stream(streams[k]):
                        carries[k], _ = A.training_backward(mods[k], saved[k], grads_out[k],
                                                            grad_dst=self._grad_slots(flats[k]),
                                                            need_feature_grads=self.need_feature_grads, phase="early")
                join()
                state["carries"] = carries
                return
            backward_rest(None)

        def backward_rest(carries):
            main = torch.cuda.current_stream()
            sides = entry["side_streams"][:len(mods) - 1]
            streams = [main] + sides
            fgs = [None] * len(mods)
            ev = torch.cuda.Event()
            ev.record(main)
            for st in sides:
                st.wait_event(ev)
            for k in reversed(range(len(mods))):
                with torch.cuda.stream(streams[k]):
                    if carries is None:
                        fgs[k], _ = A.training_backward(mods[k], state["saved"][k], state["grads_out"][k],
                                                        grad_dst=self._grad_slots(flats[k]),
                                                        need_feature_grads=self.need_feature_grads)
                    else:
                        fgs[k], _ = A.training_backward(mods[k], state["saved"][k], state["grads_out"][k],
                                                        grad_dst=self._grad_slots(flats[k]),
                                                        need_feature_grads=self.need_feature_grads, phase="late",
                                                        carry=carries[k])
            for st in sides:
                ev2 = torch.cuda.Event()
                ev2.record(st)
                main.wait_event(ev2)
            state["fg"] = fgs
            total = None
            if self.need_feature_grads:                       # every head reads the same pyramid
                total = list(fgs[0])
                for other in fgs[1:]:
                    total = [a + b for a, b in zip(total, other)]
            entry["feature_grads"] = total

        def seg_late_backward():
            backward_rest(state["carries"])

        if len(mods) >= 2 and (self.overlap_heads or self.semseg_head is not None):
            if split:
                entry["mode"] = "concurrent_split"
                return [seg_heads_concurrently, seg_late_backward, seg_optimizer]
            entry["mode"] = "concurrent"
            return [seg_heads_concurrently, seg_optimizer]
        entry["mode"] = "serial"
        return [seg_forward_loss_backward_last, seg_backward_first, seg_optimizer]

    def _capture(self, feats, targets):
        from stemseg_b200 import decoder as D
        dev = feats[0].device
        planes = D.PRECISION_PLANES[self.embedding_head.precision]
        masks = targets[0]["masks"]
        entry = {"in_planes": [D.pack_activation(f.detach(), planes) for f in feats],
                 "masks": masks.to(device=dev, dtype=torch.uint8).contiguous().clone(),
                 "ignore": targets[0]["ignore_masks"].to(device=dev, dtype=torch.uint8).contiguous().clone()}
        entry["side_streams"] = [torch.cuda.Stream(device=dev) for _ in range(2)]
        if self.semseg_head is not None:
            entry["semseg_ids"] = targets[0]["semseg_masks"].to(device=dev, dtype=torch.int64).contiguous().clone()
        segments = self._segments(entry)
        # warm-up on a side stream (lazy module loading, cudaFuncSetAttribute, allocator) -- without the optimiser
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for seg in segments[:-1]:
                seg()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        graphs, pool = [], None
        with _lib.capture_guard():
            for seg in segments:
                g = torch.cuda.CUDAGraph()
                before = _lib.KERNEL_LAUNCHES[0]
                with torch.cuda.graph(g, pool=pool):
                    seg()
                pool = g.pool()
                graphs.append((g, _lib.KERNEL_LAUNCHES[0] - before))
        entry["graphs"] = graphs
        entry["segments"] = segments           # keeps every captured buffer alive
        for flat in self.flats:                # the captures ran nothing: gradients / momentum are still untouched
            flat.grad.zero_()
        return entry

    def _step_graph(self, feats, targets):
        from stemseg_b200 import decoder as D
        if len(targets) != 1 or feats[0].shape[0] != 1:
            raise NotImplementedError("one sub-clip per rank per step (the reference's MAX_SAMPLES_PER_GPU = 1)")
        masks, ignore = targets[0]["masks"], targets[0]["ignore_masks"]
        key = (tuple(tuple(f.shape) for f in feats), tuple(masks.shape), str(feats[0].device))
        dev = feats[0].device
        with torch.no_grad(), torch.cuda.device(dev):
            self._sync_hyper(dev)
            entry = self._graphs.get(key)
            if entry is None:
                entry = self._capture(feats, targets)
                self._graphs.put(key, entry)
            planes = D.PRECISION_PLANES[self.embedding_head.precision]
            for f, pl in zip(feats, entry["in_planes"]):
                D.pack_activation(f.detach(), planes, out=pl)
            entry["masks"].copy_(masks, non_blocking=True)
            entry["ignore"].copy_(ignore, non_blocking=True)
            if self.semseg_head is not None:
                entry["semseg_ids"].copy_(targets[0]["semseg_masks"], non_blocking=True)
            pending = []
            graphs = entry["graphs"]
            if entry["mode"] == "concurrent":  # both heads in one graph, then both reductions
                graphs[0][0].replay()
                if self.world > 1:
                    pending = [dist.all_reduce(f.grad, group=self.group, async_op=True) for f in self.flats]
            elif entry["mode"] == "concurrent_split":
                # forward + loss + early backward | reduce the low-resolution blocks' gradients while the long
                # block_8x / block_4x backward runs | reduce the rest
                graphs[0][0].replay()
                if self.world > 1:
                    pending = [dist.all_reduce(f.grad[:f.prefix_end], group=self.group, async_op=True)
                               for f in self.flats]
                graphs[1][0].replay()
                if self.world > 1:
                    pending += [dist.all_reduce(f.grad[f.prefix_end:], group=self.group, async_op=True)
                                for f in self.flats if f.prefix_end < f.numel]
            else:
                graphs[0][0].replay()
                if self.world > 1:             # gradients of the last head are complete: reduce while g2 runs
                    pending.append(dist.all_reduce(self.flats[-1].grad, group=self.group, async_op=True))
                graphs[1][0].replay()
                if self.world > 1 and len(self.flats) > 1:
                    pending.append(dist.all_reduce(self.flats[0].grad, group=self.group, async_op=True))
            for work in pending:
                work.wait()
            graphs[-1][0].replay()
            _lib.KERNEL_LAUNCHES[0] += sum(k for _, k in graphs)
            for m in self._modules():
                m.invalidate_packed_weights()  # the packed copies inside the graph pool predate this step's update
        losses = entry["losses"]
        out = {"optimization_losses": {"embedding_loss": losses[0]},
               "others": {"lovasz_loss": losses[1], "variance_smoothness_loss": losses[2], "seediness_loss": losses[3]},
               "feature_grads": entry["feature_grads"]}
        if self.semseg_head is not None:
            sl = entry["semseg_losses"]
            out["others"]["semantic_segmentation_loss"] = sl[0]
            out["optimization_losses"]["semantic_segmentation_loss"] = sl[0] * self.weight_semseg
            if self.semseg_head.has_foreground_channel:
                out["optimization_losses"]["foreground"] = sl[1]
        return out
