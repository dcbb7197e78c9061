"""Clip-parallel inference: the path shards at sub-clip granularity.

Sub-clips are independent through heads -> gather -> clustering (the only cross-clip state of the reference is
``cluster_label_start``, online_chainer.py:162,183-196, and clustering is label-offset invariant), so sub-clip i is
processed by rank ``i % world_size`` with local labels 1..K_i.  ONE exchange step follows: an all-gather of the
per-frame label vectors (+ counts, + the small clustering metadata); every rank then replays the same sequential
stitch (stemseg_b200.chaining.stitch_subsequences) and obtains bit-identical global track ids.  Backend: NCCL over
NVLink on the GPU box (device tensors), gloo in the CPU tests; payload < 1 MB per sub-clip, so stock all_gather is the
right tool (no custom transport).
"""
import torch
import torch.distributed as dist

from stemseg_b200.chaining import stitch_subsequences, stitch_subsequences_device


def shard_subclips(num_subclips, rank, world_size):
    """Round-robin ownership: sub-clip i -> rank i % world_size."""
    return [i for i in range(num_subclips) if i % world_size == rank]


def exchange_and_stitch(num_frames, subseq_frames, local_results, group=None, device=None):
    """All-gather the per-sub-clip labels of every rank, then stitch.

    subseq_frames: frame lists of ALL sub-clips (every rank knows the windowing).
    local_results: {sub-clip index: (list of per-frame int64 label tensors (local labels, start 1), meta dict)} for
    the sub-clips this rank owns.  Returns (TrackContainer, relabelled per-sub-clip labels, metas) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_sub = len(subseq_frames)
    if world == 1:
        labels = [[l.cpu() for l in local_results[i][0]] for i in range(n_sub)]
        return stitch_subsequences(num_frames, subseq_frames, labels, [local_results[i][1] for i in range(n_sub)])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
    owned = shard_subclips(n_sub, rank, world)
    assert sorted(local_results.keys()) == owned, "rank %d owns %s but got %s" % (rank, owned, sorted(local_results))
    max_local = (n_sub + world - 1) // world
    max_t = max(len(f) for f in subseq_frames)
    # 1) per-frame counts of every owned sub-clip (fixed shape -> plain all_gather)
    counts = torch.full((max_local, max_t), -1, dtype=torch.int64)
    for slot, i in enumerate(owned):
        for j, lab in enumerate(local_results[i][0]):
            counts[slot, j] = lab.numel()
    counts = counts.to(device)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = [c.cpu() for c in all_counts]
    totals = [int(c.clamp(min=0).sum()) for c in all_counts]
    # 2) labels, padded to the largest per-rank total
    pad_to = max(max(totals), 1)
    flat = [lab.reshape(-1).to(torch.int32) for i in owned for lab in local_results[i][0]]
    mine = torch.cat(flat) if flat else torch.zeros(0, dtype=torch.int32)
    buf = torch.full((pad_to,), -2, dtype=torch.int32, device=device)
    buf[:mine.numel()] = mine.to(device)
    all_labels = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(all_labels, buf, group=group)
    # 3) clustering metadata (tiny python dicts)
    metas_mine = {i: local_results[i][1] for i in owned}
    all_metas = [None] * world
    dist.all_gather_object(all_metas, metas_mine, group=group)
    # reassemble in sub-clip order
    labels_by_clip, meta_by_clip = {}, {}
    for r in range(world):
        cursor = 0
        lab_r = all_labels[r].cpu().to(torch.int64)
        for slot, i in enumerate(shard_subclips(n_sub, r, world)):
            per_frame = []
            for j in range(len(subseq_frames[i])):
                c = int(all_counts[r][slot, j])
                per_frame.append(lab_r[cursor:cursor + c].clone())
                cursor += c
            labels_by_clip[i] = per_frame
            meta_by_clip[i] = all_metas[r][i]
    return stitch_subsequences(num_frames, subseq_frames, [labels_by_clip[i] for i in range(n_sub)],
                               [meta_by_clip[i] for i in range(n_sub)])


def exchange_and_stitch_device(num_frames, subseq_frames, local_results, group=None):
    """Device-resident variant: local_results = {i: (int64 CUDA tensor of all local labels of sub-clip i, per-frame
    counts, meta)}.  One NCCL all_gather of the (padded) label vectors + one of the counts; the stitch then runs on the
    device of every rank (stemseg_b200.chaining.stitch_subsequences_device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_sub = len(subseq_frames)
    if world > 1:
        owned = shard_subclips(n_sub, rank, world)
        assert sorted(local_results.keys()) == owned
        device = next(iter(local_results.values()))[0].device if local_results else torch.device(
            "cuda", torch.cuda.current_device())
        max_local = (n_sub + world - 1) // world
        max_t = max(len(f) for f in subseq_frames)
        counts = torch.full((max_local, max_t), -1, dtype=torch.int64)
        for slot, i in enumerate(owned):
            counts[slot, :len(local_results[i][1])] = torch.tensor(local_results[i][1], dtype=torch.int64)
        counts = counts.to(device)
        all_counts = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        all_counts = torch.stack(all_counts).cpu()
        totals = all_counts.clamp(min=0).sum(dim=(1, 2)).tolist()
        pad_to = max(max(totals), 1)
        buf = torch.full((pad_to,), -2, dtype=torch.int64, device=device)
        mine = [local_results[i][0] for i in owned]
        if mine:
            flat = torch.cat(mine)
            buf[:flat.numel()] = flat
        all_labels = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(all_labels, buf, group=group)
        all_metas = [None] * world
        dist.all_gather_object(all_metas, {i: local_results[i][2] for i in owned}, group=group)
        gathered = {}
        for r in range(world):
            cursor = 0
            for slot, i in enumerate(shard_subclips(n_sub, r, world)):
                cnt = [int(c) for c in all_counts[r, slot, :len(subseq_frames[i])]]
                n_i = sum(cnt)
                gathered[i] = (all_labels[r][cursor:cursor + n_i], cnt, all_metas[r][i])
                cursor += n_i
        local_results = gathered
    labels = [local_results[i][0] for i in range(n_sub)]
    counts = [local_results[i][1] for i in range(n_sub)]
    metas = [local_results[i][2] for i in range(n_sub)]
    ks = [len(m["instance_labels"]) for m in metas]
    return stitch_subsequences_device(num_frames, subseq_frames, labels, counts, ks, metas)


class ExchangeLayout(object):
    """Fixed layout of the ONE buffer every rank contributes to the all_gather: per owned sub-clip a row of int64 slots
    ``[labels: max_t * cap | header]`` whose header holds int32 words ``[per-frame counts: max_t | K | clustering meta]``.
    Nothing is pickled and nothing depends on data-dependent sizes, so the exchange needs no host round trip."""

    def __init__(self, n_sub, world, max_t, cap, meta_words):
        self.n_sub, self.world, self.max_t, self.cap, self.meta_words = n_sub, world, max_t, cap, meta_words
        self.max_local = (n_sub + world - 1) // world
        self.label_slots = max_t * cap
        self.hdr = max_t + 1 + meta_words
        self.row = self.label_slots + (self.hdr + 1) // 2

    def alloc(self, device):
        return torch.zeros((self.max_local, self.row), dtype=torch.int64, device=device)

    def views(self, buf):
        """buf [..., row] int64 -> (labels [..., max_t*cap] int64, header [..., hdr] int32), both views of `buf`."""
        labels = buf[..., :self.label_slots]
        head = buf[..., self.label_slots:].view(torch.int32)[..., :self.hdr]
        return labels, head

    def slot_of(self, i):
        """Sub-clip i lives in row `slot` of rank `r` (round-robin ownership, shard_subclips)."""
        return i % self.world, i // self.world

    def write(self, buf, slot, labels, counts, meta):
        """Fill one row: local labels (any length <= max_t*cap), per-frame counts (int32 [t]), clustering meta words."""
        lab, head = self.views(buf)
        lab[slot, :labels.numel()].copy_(labels, non_blocking=True)
        t = counts.numel()
        head[slot, :t].copy_(counts, non_blocking=True)
        head[slot, self.max_t:self.max_t + 1 + self.meta_words].copy_(torch.cat([meta[0:1], meta]), non_blocking=True)

    def gather(self, buf, group=None):
        """All ranks' buffers [world, max_local, row] (the input itself, unsqueezed, for a single process)."""
        if self.world == 1:
            return buf.unsqueeze(0)
        # concatenation along dim 0 (the form every backend accepts), viewed per rank
        out = torch.empty((self.world * buf.shape[0],) + tuple(buf.shape[1:]), dtype=torch.int64, device=buf.device)
        dist.all_gather_into_tensor(out, buf, group=group)
        return out.view((self.world,) + tuple(buf.shape))


def _meta_dict(words, e, max_instances, offset_labels):
    """Clustering meta words (host int32 tensor) -> the reference's meta dict (clusterers.py:161-166) with the given
    (already stitched) instance labels."""
    k = int(words[0])
    floats = words[4 + max_instances:].view(torch.float32)
    centers = floats[:max_instances * e].reshape(max_instances, e)[:k]
    bws = floats[max_instances * e:2 * max_instances * e].reshape(max_instances, e)[:k]
    return {"instance_labels": list(offset_labels), "instance_centers": centers.tolist(),
            "instance_stds": (1. / bws).clamp(min=1e-8).sqrt().tolist(), "instance_masks": []}


@torch.no_grad()
def clip_parallel_process(pipeline, masks, subseq_frames, features_for_clip, group=None):
    """Run the owned sub-clips through ``pipeline`` (stemseg_b200.pipeline.SubclipPipeline) and stitch globally.

    masks: [T,h,w] foreground masks of the whole video (None is only accepted for a single sub-clip: per-sub-clip
    seediness thresholds would give overlap frames different point sets in different sub-clips);
    features_for_clip(i) -> {scale: tensor} pyramid of sub-clip i (only called for owned sub-clips).

    Nothing is synchronised with the host until the whole video is stitched: every owned sub-clip is enqueued
    (``submit``), its local labels / per-frame counts / clustering meta words are copied device-to-device into
    ONE fixed-layout exchange buffer, a single all_gather (NCCL) makes every sub-clip visible on every rank, and the
    sequential stitch runs on the device (``DeviceStitcher``: histogram -> assignment -> relabel kernels per sub-clip).
    The only device->host copy is the final one in ``DeviceStitcher.finish``."""
    from stemseg_b200.chaining import DeviceStitcher
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_sub = len(subseq_frames)
    if masks is None and n_sub > 1:
        # a per-sub-clip seediness threshold gives overlap frames a different point set in each sub-clip; the stitch
        # needs identical points there (the reference averages seediness / semseg per frame over all sub-clips first,
        # inference/main.py:93-103) -> build the video masks with FrameAverager.foreground_index and pass them in
        raise ValueError("clip_parallel_process needs the video's foreground masks when there is more than one "
                         "sub-clip (build them with stemseg_b200.foreground.FrameAverager)")
    if not (hasattr(pipeline, "submit") and getattr(pipeline, "use_step_graph", False)) or masks is None:
        return _clip_parallel_process_eager(pipeline, masks, subseq_frames, features_for_clip, group)
    dev = pipeline.clusterer.device
    num_frames = max(max(f) for f in subseq_frames) + 1
    max_t = max(len(f) for f in subseq_frames)
    cap = int(masks.shape[-2]) * int(masks.shape[-1])
    mi = pipeline.clusterer.max_instances
    e_dims = pipeline.embedding_head.embedding_size
    from stemseg_b200 import _lib
    meta_words = int(_lib.load().stemseg_seq_cluster_meta_words(e_dims, mi))
    layout = ExchangeLayout(n_sub, world, max_t, cap, meta_words)
    with torch.cuda.device(dev):
        # the stitcher (track container, statistics, workspace) is reused from video to video (same geometry)
        cache_key = (num_frames, cap, layout.max_local, max_t, layout.hdr, world, n_sub, str(dev))
        cached = getattr(pipeline, "_clip_parallel_cache", None)
        if cached is None or cached[0] != cache_key:
            cached = (cache_key, DeviceStitcher(num_frames, cap, dev, max_instances=mi, max_subclips=n_sub))
            pipeline._clip_parallel_cache = cached
        stitcher = cached[1]
        stitcher.reset()
        # a fresh exchange buffer per video (caching allocator: no cudaMalloc): the returned per-sub-clip labels are views
        xbuf = layout.alloc(dev)
        main = torch.cuda.current_stream(dev)
        # The copies into the exchange buffer run on their own stream: submit() makes the pipeline's stream wait for
        # the CALLER's stream, so collecting sub-clip i on `main` would serialise sub-clip i+1 behind it and lose the
        # two-steps-in-flight overlap.
        collect = getattr(pipeline, "_collect_stream", None)
        if collect is None or collect.device != dev:
            collect = pipeline._collect_stream = torch.cuda.Stream(device=dev, priority=-1)
        collect.wait_stream(main)                         # the buffer above was allocated / zeroed on `main`
        xbuf.record_stream(collect)
        for slot, i in enumerate(shard_subclips(n_sub, rank, world)):
            frames = subseq_frames[i]
            pend = pipeline.submit(features_for_clip(i), fg_mask=masks[frames], cluster_label_start=1)
            view = pend.device_view()
            collect.wait_event(view["done"])
            with torch.cuda.stream(collect):
                for key in ("labels", "counts", "meta"):    # allocated on the pipeline's stream, consumed on this one
                    view[key].record_stream(collect)
                layout.write(xbuf, slot, view["labels"], view["counts"][:len(frames)], view["meta"])
        main.wait_stream(collect)
        all_labels, all_head = layout.views(layout.gather(xbuf, group=group))
        for i, frames in enumerate(subseq_frames):
            r, slot = layout.slot_of(i)
            stitcher.add_subclip(frames, all_labels[r, slot], all_head[r, slot, :len(frames)],
                                 all_head[r, slot, max_t:max_t + 1])
        head_host = all_head.cpu()                      # stream-ordered after the stitch kernels: the video's one sync
        container, out_labels, out_meta = stitcher.finish()
    metas = []
    for i in range(n_sub):
        r, slot = layout.slot_of(i)
        metas.append(_meta_dict(head_host[r, slot, max_t + 1:], e_dims, mi, out_meta[i]["instance_labels"]))
    return container, out_labels, metas


@torch.no_grad()
def _clip_parallel_process_eager(pipeline, masks, subseq_frames, features_for_clip, group=None):
    """Fallback for pipelines without the graphed submit() path: per-sub-clip results on the host, then
    ``exchange_and_stitch_device``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local = {}
    for i in shard_subclips(len(subseq_frames), rank, world):
        fg_mask = None if masks is None else masks[subseq_frames[i]]
        res = pipeline(features_for_clip(i), fg_mask=fg_mask, cluster_label_start=1)
        local[i] = (res.labels, list(res.fg_index.frame_counts), res.meta)
    num_frames = max(max(f) for f in subseq_frames) + 1
    return exchange_and_stitch_device(num_frames, subseq_frames, local, group=group)
