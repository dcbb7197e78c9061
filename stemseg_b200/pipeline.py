"""The hot path for one sub-clip, end to end on the device:

    FPN feature pyramid -> embedding / seediness (/ semseg) heads -> split -> foreground compaction + gather
    (bandwidths = exp(var) * 10 fused into the gather) -> sequential clustering -> per-frame label vectors

i.e. what ``InferenceModel.forward`` does after the backbone for one sub-clip (stemseg/modeling/inference_model.py:
121-162) followed by ``OnlineChainer.cluster_subsequence`` (stemseg/inference/online_chainer.py:244-289), without
the reference's device -> host -> device round trip of the head outputs (inference_model.py:162 ->
online_chainer.py:174-176).  ``SubclipPipeline.__call__`` is the public call benchmarked by bench.py.
"""
import torch

from stemseg_b200.foreground import compact_foreground, gather_points


class SubclipResult(object):
    __slots__ = ("labels", "frame_labels", "meta", "fg_index", "embeddings", "variances", "seediness", "semseg_logits",
                 "labels_host")


class PendingStep(object):
    """A sub-clip whose GPU work has been enqueued (SubclipPipeline.submit)."""

    def __init__(self, pipe, state, snap, counts_host, meta_host, labels_host, done, label_start):
        self._pipe, self._state, self._snap = pipe, state, snap
        self._counts_host, self._meta_host, self._labels_host = counts_host, meta_host, labels_host
        self._done, self._label_start = done, label_start
        self._result = None
        self.inputs_consumed = None      # CUDA event: the step no longer reads the caller's feature tensors

    def device_view(self):
        """What a device-resident consumer needs, WITHOUT synchronising: local labels (capacity-sized int64, valid
        entries = the first counts[-1]), per-frame counts (int32 [T+1], last = total), clustering meta words (int32,
        word 0 = number of clusters; see stemseg_seq_cluster_meta_words) and the CUDA event that marks them complete."""
        return {"labels": self._snap["labels"], "counts": self._snap["counts"], "meta": self._snap["meta"],
                "done": self._done, "max_instances": self._pipe.clusterer.max_instances,
                "embedding_dims": self._state["pending"]["e"]}

    def result(self):
        if self._result is not None:
            return self._result
        from stemseg_b200.foreground import ForegroundIndex
        self._done.synchronize()                       # this step only
        counts = self._counts_host.tolist()
        pending = dict(self._state["pending"])
        pending["labels"], pending["primary"] = self._snap["labels"], self._snap["primary"]
        pending["label_start"] = self._label_start
        labels, meta = self._pipe.clusterer.finish(pending, meta_host=self._meta_host)
        if self._label_start != 1:                     # labels are offset-invariant (SURVEY §8a quirk v)
            labels = torch.where(labels >= 0, labels + (self._label_start - 1), labels)
        res = SubclipResult()
        res.embeddings, res.variances, res.seediness = self._snap["emb"], self._snap["var"], self._snap["seed"]
        res.semseg_logits = self._snap["semseg"]
        res.labels, res.meta = labels, meta
        res.fg_index = ForegroundIndex(self._snap["indices"][:counts[-1]], counts[:-1], self._state["fg"].shape)
        res.frame_labels = list(labels.split(counts[:-1], 0))
        res.labels_host = None
        if self._labels_host is not None:
            host = self._labels_host[:counts[-1]]
            if self._label_start != 1:
                host = torch.where(host >= 0, host + (self._label_start - 1), host)
            res.labels_host = host
        self._result = res
        return res


class SubclipPipeline(object):
    def __init__(self, embedding_head, seediness_head, clusterer, semseg_head=None, seediness_fg_threshold=0.25,
                 embedding_scales=(32, 16, 8, 4), semseg_scales=(4, 8, 16, 32)):
        self.embedding_head = embedding_head
        self.seediness_head = seediness_head
        self.semseg_head = semseg_head
        self.clusterer = clusterer
        self.seediness_fg_threshold = seediness_fg_threshold
        self.embedding_scales = tuple(embedding_scales)
        self.semseg_scales = tuple(semseg_scales)
        if seediness_head is None and embedding_head.seediness_channels == 0:
            raise ValueError("no seediness source: give a seediness head or an embedding head with seediness_output")
        self.fuse_heads = True          # run all heads as one HeadSet (shared im2col operand, one CUDA graph)
        self.use_step_graph = True      # heads + compaction + gather + clustering replayed as one CUDA graph
        self.steps_in_flight = 2        # independent graph instances used round-robin by submit()
        self.max_cached_shapes = 4      # captured step graphs kept per instance (LRU; older shapes are re-captured)
        self._group = None
        self._group_key = None

    def _head_group(self):
        """One launch plan for all heads (they read the same pyramid): stemseg_b200.decoder.HeadSet."""
        from stemseg_b200 import decoder as D
        heads = [self.embedding_head]
        if not self.embedding_head.seediness_channels:
            heads.append(self.seediness_head)
        if self.semseg_head is not None:
            heads.append(self.semseg_head)
        specs = [h.head_spec() for h in heads]
        key = tuple(id(sp) for sp in specs)
        if self._group is None or self._group_key != key:
            e = self.embedding_head
            for h in heads:
                if h.num_frames != e.num_frames or h.precision != e.precision:
                    raise ValueError("linked heads must share NUM_FRAMES and precision")
            self._group = D.HeadSet(specs, e.num_frames, D.PRECISION_PLANES[e.precision],
                                    use_graph=e.use_cuda_graph)
            self._group_key = key
        return self._group

    @torch.no_grad()
    def run_heads(self, features):
        """features: dict {scale: [1,C,T,h,w]} -> (embeddings [E,T,h,w], variances [V,T,h,w], seediness [1,T,h,w],
        semseg logits or None).  Slices of one output tensor, as in inference_model.py:140-146."""
        emb_in = [features[s] for s in self.embedding_scales]
        if emb_in[0].shape[0] != 1:
            raise ValueError("SubclipPipeline processes one sub-clip at a time (batch dimension must be 1)")
        e, v = self.embedding_head.embedding_size, self.embedding_head.variance_channels
        if self.fuse_heads and tuple(self.embedding_scales) == (32, 16, 8, 4):
            outs = [o.squeeze(0) for o in self._head_group().run(emb_in)]
            out = outs.pop(0)
            seediness = out[e + v:e + v + 1] if self.embedding_head.seediness_channels else outs.pop(0)
            semseg = outs.pop(0) if self.semseg_head is not None else None
            return out[:e], out[e:e + v], seediness, semseg
        out = self.embedding_head(emb_in).squeeze(0)
        embeddings, variances = out[:e], out[e:e + v]
        if self.embedding_head.seediness_channels:
            seediness = out[e + v:e + v + 1]
        else:
            seediness = self.seediness_head(emb_in).squeeze(0)
        semseg = None
        if self.semseg_head is not None:
            semseg = self.semseg_head([features[s] for s in self.semseg_scales]).squeeze(0)
        return embeddings, variances, seediness, semseg

    @torch.no_grad()
    def cluster(self, embeddings, variances, seediness, fg_mask=None, cluster_label_start=1,
                return_label_masks=False):
        """fg_mask: [T,h,w] (non-zero = foreground) or None -> seediness > seediness_fg_threshold
        (stemseg/inference/main.py:93-103 for a single sub-clip)."""
        if fg_mask is None:
            fg = compact_foreground(seediness[0], threshold=self.seediness_fg_threshold)
        else:
            fg = compact_foreground(fg_mask)
        emb_flat = gather_points(embeddings, fg)
        bw_flat = gather_points(variances, fg, transform="exp10")      # inference_model.py:148
        seed_flat = gather_points(seediness, fg)
        labels, meta = self.clusterer(emb_flat, bandwidths=bw_flat, seediness=seed_flat,
                                      cluster_label_start=cluster_label_start,
                                      return_label_masks=return_label_masks)
        assert labels.numel() == emb_flat.shape[0]                     # online_chainer.py:286
        return labels, meta, fg

    # ---- whole step as ONE CUDA graph ---------------------------------------------------------------------------
    def _capture_step(self, emb_in, fg_mask):
        """heads plan -> compaction -> gathers -> clustering captured into one CUDA graph.  Point counts stay on the
        device (compaction writes them, gather / clustering read them), so nothing inside needs the host."""
        from stemseg_b200 import decoder as D
        group = self._head_group()
        dev = emb_in[0].device
        e, v = self.embedding_head.embedding_size, self.embedding_head.variance_channels
        in_planes = group.pack_inputs(emb_in)
        mask_static = None if fg_mask is None else torch.empty_like(fg_mask)
        if mask_static is not None:
            mask_static.copy_(fg_mask)

        def body(streams):
            outs = [o.squeeze(0) for o in group._plan(in_planes, streams=streams)]
            tail = group._tail_stream if streams is not None else None
            if tail is None:
                return rest(outs)
            # everything after the long 4x GEMM stays on the plan's high-priority tail stream (decoder.HeadSet._plan)
            with torch.cuda.stream(tail):
                state = rest(outs)
            torch.cuda.current_stream().wait_event(tail.record_event())
            return state

        def rest(outs):
            out = outs.pop(0)
            seediness = out[e + v:e + v + 1] if self.embedding_head.seediness_channels else outs.pop(0)
            semseg = outs.pop(0) if self.semseg_head is not None else None
            emb, var = out[:e], out[e:e + v]
            if mask_static is not None:
                fg = compact_foreground(mask_static, sync=False)
            elif semseg is not None and self.semseg_head.has_foreground_channel:
                fg = compact_foreground(semseg[-1], threshold=0.0, sync=False)
            else:
                fg = compact_foreground(seediness[0], threshold=self.seediness_fg_threshold, sync=False)
            ef = gather_points(emb, fg)
            bf = gather_points(var, fg, transform="exp10")
            sf = gather_points(seediness, fg)
            pending = self.clusterer.launch(ef, bf, sf.reshape(-1), 1, n_points_dev=fg.total_dev)
            return {"emb": emb, "var": var, "seed": seediness, "semseg": semseg, "fg": fg, "pending": pending,
                    "flat": (ef, bf, sf)}

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            D.KEEP.reset()
            body(None)
            D.KEEP.reset()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(4)]     # small branches + tail: high priority
        from stemseg_b200 import _lib
        D.KEEP.reset()
        before = _lib.KERNEL_LAUNCHES[0]
        with _lib.capture_guard(), torch.cuda.graph(graph):
            state = body(streams)
        kernels = _lib.KERNEL_LAUNCHES[0] - before
        keep = D.KEEP.take()
        return {"graph": graph, "in_planes": in_planes, "mask": mask_static, "state": state, "keep": keep,
                "kernels": kernels, "streams": streams, "planes": group.planes}

    @torch.no_grad()
    def submit(self, features, fg_mask=None, cluster_label_start=1, labels_to_host=False):
        """Enqueue one sub-clip (pack -> graph replay -> stream-ordered copies of the results) and return at once.

        The returned ``PendingStep.result()`` synchronises on that step only, so the host work of step i (and the
        pack / replay of step i+1) overlaps the kernels of step i: ``submit`` the next clip before asking for the
        previous result."""
        from stemseg_b200 import _lib, decoder as D
        emb_in = [features[s] for s in self.embedding_scales]
        if emb_in[0].shape[0] != 1:
            raise ValueError("SubclipPipeline processes one sub-clip at a time (batch dimension must be 1)")
        dev = emb_in[0].device
        if not hasattr(self, "_step_graphs"):
            # bounded: each entry owns a CUDA graph with a private pool of ~1-3 GB (keyed on shapes / weights version)
            self._step_graphs = _lib.LRUCache(self.max_cached_shapes * max(1, self.steps_in_flight))
            self._instance = 0
        # `steps_in_flight` independent instances (own static buffers, own stream) are used round-robin, so the
        # latency-bound tail of step i (merges, clustering) overlaps the convolutions of step i+1
        self._instance = (self._instance + 1) % max(1, self.steps_in_flight)
        key = (tuple(tuple(f.shape) for f in emb_in), str(dev),
               None if fg_mask is None else (tuple(fg_mask.shape), fg_mask.dtype),
               tuple(id(sp) for sp in self._head_group().specs), self._instance)
        with torch.cuda.device(dev):
            entry = self._step_graphs.get(key)
            if entry is None:
                entry = self._capture_step(emb_in, fg_mask)
                # high priority: the operand packing of step i+1 (eager launches on this stream) must not queue behind
                # the thousands of CTAs of step i's 4x GEMM (normal priority, captured in the graph) -- it is HBM work
                # that hides under that GEMM, and the step's own GEMM can then start the moment the previous one ends
                entry["stream"] = torch.cuda.Stream(device=dev, priority=-1)
                self._step_graphs.put(key, entry)
            caller = torch.cuda.current_stream(dev)
            run_stream = entry["stream"] if self.steps_in_flight > 1 else caller
            if run_stream is not caller:
                run_stream.wait_stream(caller)                  # the caller's features / mask are ready
            stream_ctx = torch.cuda.stream(run_stream)
            stream_ctx.__enter__()
            self._head_group().pack_inputs(emb_in, out=entry["in_planes"])
            if fg_mask is not None:
                entry["mask"].copy_(fg_mask)
            consumed = torch.cuda.Event()
            consumed.record(run_stream)                         # inputs may be overwritten after this point
            entry["graph"].replay()
            _lib.KERNEL_LAUNCHES[0] += entry["kernels"]
            st = entry["state"]
            # stream-ordered snapshots: the next replay may overwrite the graph's static buffers right after these
            snap = {"labels": st["pending"]["labels"].clone(), "primary": st["pending"]["primary"].clone(),
                    "indices": st["fg"].capacity_indices.clone(), "emb": st["emb"].clone(), "var": st["var"].clone(),
                    "seed": st["seed"].clone(), "semseg": None if st["semseg"] is None else st["semseg"].clone(),
                    "counts": st["fg"].counts_dev.clone(), "meta": st["pending"]["meta"].clone()}
            counts_host = torch.empty(st["fg"].counts_dev.shape, dtype=torch.int32, pin_memory=True)
            meta_host = torch.empty(st["pending"]["meta"].shape, dtype=torch.int32, pin_memory=True)
            counts_host.copy_(st["fg"].counts_dev, non_blocking=True)
            meta_host.copy_(st["pending"]["meta"], non_blocking=True)
            labels_host = None
            if labels_to_host:
                labels_host = torch.empty(snap["labels"].shape, dtype=torch.int64, pin_memory=True)
                labels_host.copy_(snap["labels"], non_blocking=True)
            done = torch.cuda.Event()
            done.record(run_stream)
            stream_ctx.__exit__(None, None, None)
            for f in emb_in:                      # keep the caller's tensors alive for the side stream (allocator)
                f.record_stream(run_stream)
            if fg_mask is not None:
                fg_mask.record_stream(run_stream)
        pend = PendingStep(self, st, snap, counts_host, meta_host, labels_host, done, int(cluster_label_start))
        pend.inputs_consumed = consumed
        return pend

    @torch.no_grad()
    def run_graphed(self, features, fg_mask=None, cluster_label_start=1):
        """Same result as the eager path, replaying one captured CUDA graph per step."""
        return self.submit(features, fg_mask, cluster_label_start).result()

    @torch.no_grad()
    def __call__(self, features, fg_mask=None, cluster_label_start=1):
        if self.use_step_graph and self.fuse_heads and tuple(self.embedding_scales) == (32, 16, 8, 4) and \
                self.embedding_head.use_cuda_graph:
            return self.run_graphed(features, fg_mask, cluster_label_start)
        res = SubclipResult()
        res.embeddings, res.variances, res.seediness, res.semseg_logits = self.run_heads(features)
        if fg_mask is None and res.semseg_logits is not None and self.semseg_head.has_foreground_channel:
            # foreground channel is the last one (inference_model.py:212-225); prob > 0.5 <=> logit > 0
            fg = compact_foreground(res.semseg_logits[-1], threshold=0.0)
            fg_mask_index = fg
            emb_flat = gather_points(res.embeddings, fg)
            bw_flat = gather_points(res.variances, fg, transform="exp10")
            seed_flat = gather_points(res.seediness, fg)
            res.labels, res.meta = self.clusterer(emb_flat, bandwidths=bw_flat, seediness=seed_flat,
                                                  cluster_label_start=cluster_label_start)
            res.fg_index = fg_mask_index
        else:
            res.labels, res.meta, res.fg_index = self.cluster(res.embeddings, res.variances, res.seediness, fg_mask,
                                                              cluster_label_start)
        res.frame_labels = list(res.labels.split(res.fg_index.frame_counts, 0))   # online_chainer.py:289
        return res


class HostFeatureStream(object):
    """Double-buffered host -> device staging of feature pyramids held in pinned host memory.

    ``submit`` enqueues the copies of one pyramid on a dedicated copy stream and returns a ticket; ``get`` makes the
    current (compute) stream wait for them; ``release`` tells the stager the compute stream is done reading the
    buffer.  The H2D transfer of clip i+1 (282 MB at 480p) thereby overlaps the kernels of clip i."""

    def __init__(self, device, depth=3):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._buffers = [None] * depth
        self._ready = [None] * depth
        self._free = [None] * depth
        self._next = 0

    def submit(self, host_features):
        slot = self._next
        self._next = (self._next + 1) % self.depth
        if self._buffers[slot] is None or any(self._buffers[slot][k].shape != v.shape for k, v in host_features.items()):
            self._buffers[slot] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device)
                                   for k, v in host_features.items()}
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])      # previous consumer of this buffer has finished
            for k, v in host_features.items():
                self._buffers[slot][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready[slot] = ev
        return slot

    def get(self, ticket):
        torch.cuda.current_stream(self.device).wait_event(self._ready[ticket])
        return self._buffers[ticket]

    def release(self, ticket, event=None):
        """The buffer may be refilled once `event` (default: everything enqueued so far on the current stream) is done."""
        if event is None:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(self.device))
        self._free[ticket] = event


def build_davis_pipeline(device, num_frames=8, precision="fp32", in_channels=256,
                         inter_channels=(256, 256, 128, 128), min_seediness_prob=0.0, seed=42):
    """Random-init model of the shipped DAVIS config (davis_1.yaml: E=4 'xyff', separate seediness head, GN32,
    free_dim_stds [0.3, 0.3]; defaults.yaml:114-117 clustering thresholds) on `device`."""
    from functools import partial
    import torch.nn as nn
    from stemseg_b200.clusterers import SequentialClustering
    from stemseg_b200.heads import EmbeddingHead, SeedinessHead
    torch.manual_seed(seed)                                  # model_builder.py:252
    norm = partial(nn.GroupNorm, 32)
    emb = EmbeddingHead(in_channels, list(inter_channels), 4, tanh_activation=True, seediness_output=False,
                        experimental_dims="xyff", PoolType=nn.AvgPool3d, NormType=norm, num_frames=num_frames,
                        precision=precision)
    seedi = SeedinessHead(in_channels, list(inter_channels), PoolType=nn.AvgPool3d, NormType=norm,
                          num_frames=num_frames, precision=precision)
    emb, seedi = emb.to(device).eval(), seedi.to(device).eval()
    clusterer = SequentialClustering(0.5, 0.3, min_seediness_prob, 2, [0.3, 0.3], device)
    return SubclipPipeline(emb, seedi, clusterer)


# head settings of the shipped configs (stemseg/config/davis_1.yaml, youtube_vis.yaml, kitti_mots_2.yaml over
# defaults.yaml:56-86,114-117): embedding mode / size, where seediness comes from, semseg head, clustering threshold
SHIPPED_CONFIGS = {
    "davis": dict(dim_mode="xyff", embedding_size=4, seediness_head=True, semseg=None, free_dim_stds=[0.3, 0.3],
                  min_seediness_prob=0.8),
    "youtube_vis": dict(dim_mode="xyff", embedding_size=4, seediness_head=False,
                        semseg=dict(num_classes=41, inter_channels=(256, 256, 256, 256), foreground_channel=True),
                        free_dim_stds=[0.3, 0.3], min_seediness_prob=0.8),
    "kitti_mots": dict(dim_mode="xyt", embedding_size=3, seediness_head=False,
                       semseg=dict(num_classes=3, inter_channels=(256, 256, 128, 128), foreground_channel=True),
                       free_dim_stds=[], min_seediness_prob=0.95),
}


def build_pipeline(config, device, num_frames=8, precision="fp32", in_channels=256,
                   inter_channels=(256, 256, 128, 128), semseg_inter_channels=None, num_classes=None,
                   min_seediness_prob=None, seed=42):
    """Random-init heads + clusterer wired like build_model() / TrackGenerator do for one of the shipped configs
    (model_builder.py:268-331, inference/main.py:84-91).  Channel widths can be overridden (tests use small ones)."""
    from functools import partial
    import torch.nn as nn
    from stemseg_b200.clusterers import SequentialClustering
    from stemseg_b200.heads import EmbeddingHead, SeedinessHead, SemsegHead, get_nb_free_dims
    cfg = SHIPPED_CONFIGS[config]
    torch.manual_seed(seed)                                  # model_builder.py:252
    norm = partial(nn.GroupNorm, 32)
    emb = EmbeddingHead(in_channels, list(inter_channels), cfg["embedding_size"], tanh_activation=True,
                        seediness_output=not cfg["seediness_head"], experimental_dims=cfg["dim_mode"],
                        PoolType=nn.AvgPool3d, NormType=norm, num_frames=num_frames, precision=precision)
    seedi = None
    if cfg["seediness_head"]:
        seedi = SeedinessHead(in_channels, list(inter_channels), PoolType=nn.AvgPool3d, NormType=norm,
                              num_frames=num_frames, precision=precision).to(device).eval()
    semseg = None
    if cfg["semseg"] is not None:
        sc = cfg["semseg"]
        semseg = SemsegHead(in_channels, num_classes if num_classes is not None else sc["num_classes"],
                            inter_channels=list(semseg_inter_channels or sc["inter_channels"]),
                            feature_scales=[4, 8, 16, 32], foreground_channel=sc["foreground_channel"],
                            PoolType=nn.AvgPool3d, NormType=norm, num_frames=num_frames,
                            precision=precision).to(device).eval()
    emb = emb.to(device).eval()
    n_free = get_nb_free_dims(cfg["dim_mode"])
    assert n_free == len(cfg["free_dim_stds"])
    msp = cfg["min_seediness_prob"] if min_seediness_prob is None else min_seediness_prob
    clusterer = SequentialClustering(0.5, 0.3, msp, n_free, list(cfg["free_dim_stds"]), device)
    return SubclipPipeline(emb, seedi, clusterer, semseg_head=semseg)
