"""Sub-clip chaining: windows over a video, per-sub-clip clustering on the device, and the sequential stitch.

Host-side mirror of the reference's inference glue around the hot path (same names / arguments / return structure):
  * ``get_subsequence_frames``           stemseg/inference/main.py:23-49
  * ``TrackContainer``                   stemseg/inference/online_chainer.py:25-117
  * ``OnlineChainer.process`` / ``cluster_subsequence`` / ``associate_clusters``   online_chainer.py:120-343
The per-point work of ``cluster_subsequence`` (gather + clustering) runs in the CUDA kernels; the stitch
(``stitch_subsequences``) is integer host logic on label vectors: a K1 x K2 (<= 20 x 20) IoU table from one joint
histogram per overlap and scipy's Hungarian solver, exactly like online_chainer.py:291-343.  It only consumes
labels, which is what makes sub-clips independent and the path clip-parallel (stemseg_b200/parallel.py).

``TrackContainer`` and ``get_subsequence_frames`` are host control-plane mirrors of the reference classes (same method
names, assertions and return values, because ``OnlineChainer.process`` callers and the output generators index into
them); they hold no per-point arithmetic.  The device-resident equivalent used on the fast path is ``DeviceStitcher``
(csrc/stitch.cu), which keeps the same state in GPU memory.
"""
from collections import defaultdict

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from stemseg_b200.foreground import ForegroundIndex, compact_foreground, gather_points

DEFAULT_FRAME_OVERLAP = {"davis": 6, "ytvis": 4, "kittimots": 4}      # defaults.yaml:91,101,110


def get_subsequence_frames(seq_len, subseq_len, dataset_name=None, frame_overlap=-1):
    """Overlapping windows of a video (main.py:23-49).  Returns (list of frame-index lists, padded-frame flags|None)."""
    if dataset_name not in DEFAULT_FRAME_OVERLAP:          # main.py:26-33: unknown datasets raise before anything else
        raise NotImplementedError()
    if frame_overlap <= 0:
        frame_overlap = DEFAULT_FRAME_OVERLAP[dataset_name]
    assert frame_overlap < subseq_len
    if seq_len < subseq_len:                           # short video: repeat frame 0 (main.py:37-39)
        pad = subseq_len - seq_len
        return [[0] * pad + list(range(seq_len))], [True] * pad + [False] * seq_len
    windows, last = [], -1
    for t in range(0, seq_len - subseq_len + 1, subseq_len - frame_overlap):
        windows.append(list(range(t, t + subseq_len)))
        last = windows[-1][-1]
    if last != seq_len - 1:                            # tail window (main.py:46-47)
        windows.append(list(range(seq_len - subseq_len, seq_len)))
    return windows, None


class TrackContainer(object):
    """Final stitched labels of every frame (online_chainer.py:25-117)."""

    def __init__(self, num_frames):
        self._frame_labels = [None for _ in range(num_frames)]
        self._is_frozen = [False for _ in range(num_frames)]
        self._highest_instance_id = 0

    def add_labels(self, frame_nums, labels):
        assert all([self._frame_labels[t] is None for t in frame_nums])
        for t, labels_t in zip(frame_nums, labels):
            self._frame_labels[t] = labels_t
            if labels_t.numel() > 0:
                self._highest_instance_id = max(self._highest_instance_id, labels_t.max().item())
        return self._highest_instance_id + 1

    def labels_exist(self, frame_num):
        return self._frame_labels[frame_num] is not None

    def has_fg_pixels(self, frame_num):
        assert self.labels_exist(frame_num)
        return self._frame_labels[frame_num].numel() > 0

    def get_labels(self, frame_nums):
        assert all(self.labels_exist(t) for t in frame_nums)
        return [self._frame_labels[t] for t in frame_nums]

    def update_labels(self, frame_num, labels):
        assert self.labels_exist(frame_num)
        assert not self._is_frozen[frame_num]
        self._frame_labels[frame_num] = labels
        if labels.numel() > 0:
            self._highest_instance_id = max(self._highest_instance_id, labels.max().item())
        return self._highest_instance_id

    def freeze_frame(self, frame_num):
        assert self.labels_exist(frame_num)
        self._is_frozen[frame_num] = True

    def get_track_mask_idxes(self):
        """-> (per-frame label tensors, {id: point count}, {id: lifetime})  (online_chainer.py:94-117)."""
        counts = defaultdict(lambda: 0)
        span = defaultdict(lambda: [10000, -1])
        for frame_num, labels in enumerate(self._frame_labels):
            ids, cnt = np.unique(labels.numpy(), return_counts=True)
            for i, c in zip(ids.tolist(), cnt.tolist()):
                counts[i] += c
                span[i][0] = min(frame_num, span[i][0])
                span[i][1] = max(frame_num, span[i][1])
        lifetimes = {k: v[1] - v[0] for k, v in span.items()}
        return self._frame_labels, counts, lifetimes


OUTLIER_LABEL = -1


def associate_label_sets(labels_1, labels_2):
    """IoU association of the clusters of two labelings of the same points (online_chainer.py:291-343).

    labels_*: 1-D integer arrays (np or torch).  Returns (associations [(l1, l2)], unassigned_1, unassigned_2,
    costs of the chosen pairs, (recall matrix, unique_1, unique_2)) like the reference."""
    a = labels_1.cpu().numpy() if torch.is_tensor(labels_1) else np.asarray(labels_1)
    b = labels_2.cpu().numpy() if torch.is_tensor(labels_2) else np.asarray(labels_2)
    assert a.shape == b.shape, "Shape mismatch: {}, {}".format(a.shape, b.shape)
    # one joint histogram over the raw label values gives every count that is needed (labels are small ints >= -1)
    a1 = a.astype(np.int64) + 1
    b1 = b.astype(np.int64) + 1
    assert a1.size == 0 or (a1.min() >= 0 and b1.min() >= 0), "labels below the outlier label -1"
    na = int(a1.max()) + 1 if a1.size else 1
    nb = int(b1.max()) + 1 if b1.size else 1
    joint = np.bincount(a1 * nb + b1, minlength=na * nb).reshape(na, nb)
    size_a, size_b = joint.sum(1), joint.sum(0)
    # the reference builds the label lists through python sets (online_chainer.py:307-308); keep its ordering
    unique_1 = list(set((np.nonzero(size_a)[0] - 1).tolist()) - {OUTLIER_LABEL})
    unique_2 = list(set((np.nonzero(size_b)[0] - 1).tolist()) - {OUTLIER_LABEL})
    assert not set(unique_1).intersection(set(unique_2)), "Labels overlap: {}, {}".format(unique_1, unique_2)
    costs = np.zeros((len(unique_1), len(unique_2)), np.float32)
    recall = np.zeros((len(unique_1), len(unique_2)), np.float32)
    if unique_1 and unique_2:
        r = np.array(unique_1, np.int64) + 1
        c = np.array(unique_2, np.int64) + 1
        inter = joint[np.ix_(r, c)]
        size1, size2 = size_a[r], size_b[c]
        union = size1[:, None] + size2[None, :] - inter
        iou = inter.astype(np.float32) / union.astype(np.float32)          # fp32 division like the torch ops
        costs = (1.0 - iou.astype(np.float64)).astype(np.float32)          # 1. - iou.item() stored as fp32
        recall = inter.astype(np.float32) / size1.astype(np.float32)[:, None]
    rows, cols = linear_sum_assignment(costs)                              # online_chainer.py:330
    associations = []
    un1, un2 = set(unique_1), set(unique_2)
    for r, c in zip(rows, cols):
        associations.append((unique_1[r], unique_2[c]))
        un1.remove(unique_1[r])
        un2.remove(unique_2[c])
    return associations, un1, un2, costs[rows, cols], (recall, unique_1, unique_2)


def stitch_subsequences(num_frames, subseq_frames, subseq_local_labels, subseq_meta=None):
    """The sequential stitch of OnlineChainer.process (online_chainer.py:162-236) on per-sub-clip LOCAL labels.

    subseq_local_labels[i]: list (one int64 CPU tensor per frame of sub-clip i) of cluster labels obtained with
    cluster_label_start=1.  Clustering is label-offset invariant, so adding ``next_track_label - 1`` reproduces what the
    reference gets by clustering sub-clip i with cluster_label_start=next_track_label (online_chainer.py:183-185).
    Returns (TrackContainer, list of per-sub-clip relabelled label lists, list of meta dicts)."""
    container = TrackContainer(num_frames)
    next_track_label = 1
    out_labels, out_meta = [], []
    for i, frames in enumerate(subseq_frames):
        offset = next_track_label - 1
        labels = [torch.where(l >= 0, l + offset, l) for l in subseq_local_labels[i]]
        meta = None
        if subseq_meta is not None:
            meta = dict(subseq_meta[i])
            meta['instance_labels'] = [l + offset for l in meta['instance_labels']]
        if i == 0:
            next_track_label = container.add_labels(frames, labels)
            out_labels.append(labels)
            out_meta.append(meta)
            continue
        prev = subseq_frames[i - 1]
        overlapping = sorted(list(set(frames).intersection(set(prev))))
        existing = container.get_labels(overlapping)
        current = [labels[j] for j, t in enumerate(frames) if t in overlapping]
        associations, _, _, _, _ = associate_label_sets(torch.cat(existing), torch.cat(current))
        # relabel the non-overlap frames: the reference applies the (associated <- current) substitutions one after
        # the other (online_chainer.py:219-224); current labels are >= next_track_label and associated ones are
        # smaller, so the substitutions never chain and one lookup table is equivalent
        top = max([int(l.max()) for l in labels if l.numel() > 0] + [0])
        lut = torch.arange(-1, top + 1, dtype=torch.int64)            # index = label + 1
        for associated_label, current_label in associations:
            if current_label <= top:
                lut[current_label + 1] = associated_label
        overlap_set = set(overlapping)
        for j, t in enumerate(frames):
            if t in overlap_set:
                continue
            labels[j] = lut[labels[j] + 1]
            next_track_label = container.add_labels([t], [labels[j]])
        if meta is not None:
            for associated_label, current_label in associations:
                idx = meta['instance_labels'].index(current_label)
                meta['instance_labels'][idx] = associated_label
        out_labels.append(labels)
        out_meta.append(meta)
    return container, out_labels, out_meta


class OnlineChainer(object):
    """Drop-in for stemseg.inference.online_chainer.OnlineChainer with the per-point work on the device."""
    OUTLIER_LABEL = OUTLIER_LABEL

    def __init__(self, clusterer, embedding_resize_factor):
        self.clusterer = clusterer
        self.resize_scale = embedding_resize_factor
        if float(embedding_resize_factor) != int(embedding_resize_factor) or not 1 <= int(embedding_resize_factor) <= 8:
            raise NotImplementedError("embedding_resize_factor must be an integer in [1, 8] (the reference uses 1 or 4)")
        # resize_tensors (online_chainer.py:128-140) is never materialised: the (1, s, s) trilinear interpolation is
        # evaluated by the gather kernel at the foreground voxels of the full-resolution mask only
        self._upsample = int(embedding_resize_factor)

    @torch.no_grad()
    def cluster_subsequence(self, mask_idxes, embeddings, bandwidths, seediness, label_start, return_fg_embeddings):
        """mask_idxes: ForegroundIndex of the sub-clip's frames (see ``process``); maps are [C,T,H,W] CUDA tensors.
        Returns (per-frame label list, [N,E] foreground embeddings, clustering meta) (online_chainer.py:244-289)."""
        assert len(mask_idxes.frame_counts) == embeddings.shape[1]
        emb_flat = gather_points(embeddings, mask_idxes, upsample=self._upsample)
        bw_flat = gather_points(bandwidths, mask_idxes, upsample=self._upsample)
        seed_flat = gather_points(seediness, mask_idxes, upsample=self._upsample)
        labels, meta = self.clusterer(emb_flat, bandwidths=bw_flat, seediness=seed_flat,
                                      cluster_label_start=label_start, return_label_masks=return_fg_embeddings)
        assert labels.numel() == emb_flat.shape[0]
        return list(labels.split(mask_idxes.frame_counts, 0)), emb_flat, meta

    @torch.no_grad()
    def process(self, masks, subsequences, return_fg_embeddings=False):
        """masks [T,H,W]; subsequences: list of dicts with 'frames', 'embeddings' [E,T',H,W], 'bandwidths' (already
        activated), 'seediness'.  Same return tuple as the reference (online_chainer.py:241-242)."""
        device = self.clusterer.device
        num_frames = masks.shape[0]
        fg_all = compact_foreground(masks.to(device))
        mask_idxes = fg_all.coord_list()
        frames_list, local_labels, metas, fg_embeddings = [], [], [], []
        for subseq in subsequences:
            if isinstance(subseq['frames'], dict):
                subseq['frames'] = sorted(subseq['frames'].keys())
            frames = list(subseq['frames'])
            emb = subseq['embeddings'].to(device)
            assert (emb.shape[-2] * self._upsample, emb.shape[-1] * self._upsample) == tuple(masks.shape[-2:]), \
                "Size mismatch between embeddings {} (x{}) and masks {}".format(emb.shape, self._upsample, masks.shape)
            labels, fg_emb, meta = self.cluster_subsequence(
                fg_all.frame_slice(frames), emb, subseq['bandwidths'].to(device), subseq['seediness'].to(device), 1,
                return_fg_embeddings)
            frames_list.append(frames)
            local_labels.append([l.cpu() for l in labels])
            metas.append(meta)
            if return_fg_embeddings:
                fg_embeddings.append(fg_emb.cpu())
        container, subseq_labels_list, meta_out = stitch_subsequences(num_frames, frames_list, local_labels, metas)
        return container.get_track_mask_idxes(), mask_idxes, subseq_labels_list, fg_embeddings, meta_out


# ----------------------------------------------------------------------------------------------------------------
# device-side stitch: labels never leave the GPU; per sub-clip two histogram launches, one
# <= 21 x 21 table on the host for the Hungarian solve, one LUT relabel launch
# ----------------------------------------------------------------------------------------------------------------
class DeviceTrackContainer(object):
    """Result of ``stitch_subsequences_device``: per-frame label tensors on the device + host-side statistics."""

    def __init__(self, frame_labels, counts, spans):
        self._frame_labels, self._counts, self._spans = frame_labels, counts, spans

    def get_labels(self, frame_nums):
        return [self._frame_labels[t] for t in frame_nums]

    def get_track_mask_idxes(self):
        lifetimes = {k: v[1] - v[0] for k, v in self._spans.items()}
        return self._frame_labels, self._counts, lifetimes


def _pair_histogram(a, b, a_base, b_base, na, nb):
    from stemseg_b200 import _lib
    lib = _lib.load()
    dev = a.device
    with torch.cuda.device(dev):
        table = torch.empty((na, nb), dtype=torch.int32, device=dev)
        bad = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(lib.stemseg_label_pair_histogram(_lib.ptr(a), _lib.ptr(b), a.numel(), a_base, b_base, na, nb,
                                                    _lib.ptr(table), _lib.ptr(bad), _lib.stream_ptr()))
    return table, bad


def _relabel(labels, base, lut_host):
    from stemseg_b200 import _lib
    lib = _lib.load()
    if labels.numel() == 0:
        return
    with torch.cuda.device(labels.device):
        lut = torch.tensor(lut_host, dtype=torch.int64, device=labels.device)
        _lib.check(lib.stemseg_relabel_lut(_lib.ptr(labels), labels.numel(), base, _lib.ptr(lut), len(lut_host),
                                           _lib.stream_ptr()))


@torch.no_grad()
def stitch_subsequences_device(num_frames, subseq_frames, subseq_labels, subseq_counts, subseq_num_clusters,
                               subseq_meta=None):
    """Same result as ``stitch_subsequences`` with the label vectors resident on the device.

    subseq_labels[i]: contiguous int64 CUDA tensor with the LOCAL labels (cluster_label_start=1) of all frames of
    sub-clip i; subseq_counts[i]: per-frame point counts (python ints); subseq_num_clusters[i]: K_i.
    Returns (DeviceTrackContainer, per-sub-clip relabelled tensors, metas)."""
    frame_labels = [None] * num_frames
    counts = defaultdict(lambda: 0)
    spans = defaultdict(lambda: [10000, -1])
    highest = 0
    next_track_label = 1
    out_labels, out_meta = [], []

    def account(frames_added, table, lut_local):
        """table[local bin][frame slot + 1] -> global counts / lifetimes / highest id for the frames just added."""
        nonlocal highest
        for slot, t in frames_added:
            col = table[:, slot + 1]
            for bin_, c in enumerate(col.tolist()):
                if c == 0:
                    continue
                label = -1 if bin_ == 0 else lut_local[bin_ - 1]
                counts[label] += c
                spans[label][0] = min(t, spans[label][0])
                spans[label][1] = max(t, spans[label][1])
                if label > highest:
                    highest = label

    for i, frames in enumerate(subseq_frames):
        labels = subseq_labels[i].clone()
        cnts = list(subseq_counts[i])
        k = int(subseq_num_clusters[i])
        dev = labels.device
        offset = next_track_label - 1
        starts = [0]
        for c in cnts:
            starts.append(starts[-1] + c)
        frame_ids = torch.repeat_interleave(torch.arange(len(frames), device=dev),
                                            torch.tensor(cnts, device=dev)).to(torch.int64)
        per_frame, bad0 = _pair_histogram(labels, frame_ids, 1, 0, k + 2, len(frames) + 1)      # [local bin][slot+1]
        meta = None
        if subseq_meta is not None:
            meta = dict(subseq_meta[i])
            meta['instance_labels'] = [l + offset for l in meta['instance_labels']]
        lut_local = [l + offset for l in range(1, k + 2)]           # local label l -> global label (index l - 1)
        if i == 0:
            per_frame_h = per_frame.cpu().numpy()
            assert int(bad0.item()) == 0, "label outside the histogram range"
            _relabel(labels, 1, lut_local)
            for j, t in enumerate(frames):
                frame_labels[t] = labels[starts[j]:starts[j + 1]]
            account(list(enumerate(frames)), per_frame_h, lut_local)
        else:
            prev = subseq_frames[i - 1]
            overlap_set = set(frames).intersection(set(prev))
            overlapping = sorted(list(overlap_set))
            existing = torch.cat([frame_labels[t] for t in overlapping])
            current = torch.cat([labels[starts[j]:starts[j + 1]] for j, t in enumerate(frames) if t in overlap_set])
            assert existing.numel() == current.numel(), \
                "Shape mismatch: {}, {} (overlap frames must hold the same foreground points in both sub-clips)".format(
                    tuple(existing.shape), tuple(current.shape))
            na = next_track_label + 1
            joint, bad1 = _pair_histogram(existing, current, 1, 1, na, k + 2)
            joint_h = joint.cpu().numpy().astype(np.int64)           # sync: the small tables
            per_frame_h = per_frame.cpu().numpy()
            assert int(bad0.item()) == 0 and int(bad1.item()) == 0, "label outside the histogram range"
            size_a, size_b = joint_h.sum(1), joint_h.sum(0)
            # the reference builds these lists through python sets (online_chainer.py:307-308); keep its ordering
            unique_1 = list(set(np.nonzero(size_a[1:])[0] + 1))
            unique_1 = [int(v) for v in unique_1]
            unique_2 = list(set(int(v) + 1 + offset for v in np.nonzero(size_b[1:])[0]))
            assert not set(unique_1).intersection(set(unique_2)), "Labels overlap: {}, {}".format(unique_1, unique_2)
            costs = np.zeros((len(unique_1), len(unique_2)), np.float32)
            if unique_1 and unique_2:
                r = np.array(unique_1, np.int64)
                c = np.array(unique_2, np.int64) - offset
                inter = joint_h[np.ix_(r, c)]
                union = size_a[r][:, None] + size_b[c][None, :] - inter
                iou = inter.astype(np.float32) / union.astype(np.float32)
                costs = (1.0 - iou.astype(np.float64)).astype(np.float32)
            rows, cols = linear_sum_assignment(costs)
            associations = [(unique_1[r_], unique_2[c_]) for r_, c_ in zip(rows, cols)]
            lut_assoc = list(lut_local)
            for associated_label, current_label in associations:
                lut_assoc[current_label - offset - 1] = associated_label
            # overlap frames keep the raw (offset) labels in the returned sub-clip labels and are NOT added to the
            # container; non-overlap frames are relabelled through the association table
            added = []
            for j, t in enumerate(frames):
                seg = labels[starts[j]:starts[j + 1]]
                if t in overlap_set:
                    _relabel(seg, 1, lut_local)
                else:
                    _relabel(seg, 1, lut_assoc)
                    frame_labels[t] = seg
                    added.append((j, t))
            account(added, per_frame_h, lut_assoc)
            if meta is not None:
                for associated_label, current_label in associations:
                    idx = meta['instance_labels'].index(current_label)
                    meta['instance_labels'][idx] = associated_label
        next_track_label = highest + 1
        out_labels.append(labels)
        out_meta.append(meta)
    return DeviceTrackContainer(frame_labels, counts, spans), out_labels, out_meta


# ----------------------------------------------------------------------------------------------------------------
# fully device-resident stitch: nothing is read back per sub-clip (csrc/stitch.cu: stemseg_stitch_subclip)
# ----------------------------------------------------------------------------------------------------------------
STITCH_ERRORS = {1: "label outside the expected range", 2: "Shape mismatch: overlap frames hold different point sets",
                 4: "Labels overlap", 8: "too many labels for the stitch tables", 16: "frame labels already exist",
                 32: "assignment failed"}


class DeviceStitcher(object):
    """``OnlineChainer.process``'s sequential stitch (online_chainer.py:162-236) as three kernel launches per sub-clip
    and ONE device->host copy per video.

    The track container (per-frame label vectors, highest id, per-id point counts / lifetimes), the label association
    (IoU table -> scipy-identical assignment, csrc/assoc.cuh) and ``next_track_label`` all live on the device, so
    sub-clip i+1 can be enqueued before anything about sub-clip i is known on the host.  Inputs per sub-clip are the
    LOCAL labels (cluster_label_start=1; clustering is label-offset invariant) with their per-frame counts and cluster
    count still on the device -- exactly what ``SubclipPipeline.submit`` leaves there."""

    def __init__(self, num_frames, frame_capacity, device, max_instances=20, max_labels=None, max_subclips=None):
        from stemseg_b200 import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.num_frames, self.frame_capacity, self.max_instances = int(num_frames), int(frame_capacity), int(max_instances)
        if max_labels is None:
            max_labels = self.max_instances * (max_subclips if max_subclips is not None else max(1, self.num_frames))
        self.max_labels = min(int(max_labels), 4095)        # csrc/stitch.cu kAssignMaxLabels (shared-memory table)
        with torch.cuda.device(self.device):
            self.frame_labels = torch.empty((self.num_frames, self.frame_capacity), dtype=torch.int64, device=self.device)
            self.frame_count = torch.full((self.num_frames,), -1, dtype=torch.int32, device=self.device)
            # one int32 block so that the end-of-video read-back is a single copy:
            # [state 4][span_lo L+1][span_hi L+1][track_counts (L+1) x int64 as 2 x int32]
            n = self.max_labels + 1
            self._block = torch.empty(4 + 4 * n + 4, dtype=torch.int32, device=self.device)
            self.state = self._block[0:4]
            self.span_lo = self._block[4:4 + n]
            self.span_hi = self._block[4 + n:4 + 2 * n]
            off = 4 + 2 * n
            off += off % 2                                   # 8-byte alignment of the int64 view
            self.track_counts = self._block[off:off + 2 * n].view(torch.int64)
            self._ws = {}
        self.reset()

    def reset(self):
        with torch.cuda.device(self.device):
            self.frame_count.fill_(-1)
            self.state.zero_()
            self.state[0:1].fill_(1)                         # next_track_label starts at 1 (online_chainer.py:162)
            self.span_lo.fill_(10000)
            self.span_hi.fill_(-1)
            self.track_counts.zero_()
        self._subclips = []
        self._prev_frames = None

    @torch.no_grad()
    def add_subclip(self, frames, labels, frame_counts_dev, k_dev):
        """frames: list of global frame numbers; labels: int64 CUDA tensor (capacity >= sum of counts) with the local
        labels, rewritten in place; frame_counts_dev: int32 CUDA tensor [len(frames)]; k_dev: int32 CUDA tensor [1]
        (number of clusters).  Returns the device tensor of this sub-clip's instance_labels (int64 [max_instances],
        -1 padding), valid once the stream has run."""
        lib, _lib = self.lib, self._lib
        frames = [int(t) for t in frames]
        nf = len(frames)
        first = self._prev_frames is None
        prev = set() if first else set(self._prev_frames)
        overlap = [1 if t in prev else 0 for t in frames]
        if not first and not any(overlap):
            raise ValueError("sub-clip shares no frame with the previous one: nothing to associate it through")
        if labels.dtype != torch.int64 or not labels.is_cuda or not labels.is_contiguous():
            raise ValueError("labels must be a contiguous int64 CUDA tensor")
        if frame_counts_dev.dtype != torch.int32 or frame_counts_dev.numel() < nf or k_dev.dtype != torch.int32:
            raise ValueError("frame counts / cluster count must be int32 CUDA tensors")
        arr = (_lib.c_int32 * nf)
        with torch.cuda.device(self.device):
            ws = self._ws.get(nf)
            if ws is None:
                nbytes = lib.stemseg_stitch_workspace_bytes(self.max_instances, nf, self.max_labels)
                ws = self._ws[nf] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            meta = torch.empty(self.max_instances, dtype=torch.int64, device=self.device)
            _lib.check(lib.stemseg_stitch_subclip(
                _lib.ptr(labels), labels.numel(), _lib.ptr(frame_counts_dev), _lib.ptr(k_dev), arr(*frames), arr(*overlap),
                nf, 1 if first else 0, self.max_instances, _lib.ptr(self.frame_labels), _lib.ptr(self.frame_count),
                self.frame_capacity, self.num_frames, _lib.ptr(self.state), _lib.ptr(self.track_counts),
                _lib.ptr(self.span_lo), _lib.ptr(self.span_hi), self.max_labels, _lib.ptr(meta), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr()))
        self._prev_frames = frames
        self._subclips.append((frames, labels, frame_counts_dev, meta))
        return meta

    @torch.no_grad()
    def finish(self, subseq_meta=None):
        """The one synchronisation of the video: fetch state + statistics + per-frame counts, raise on device-side
        errors, and return (DeviceTrackContainer, per-sub-clip relabelled label tensors, metas) like
        ``stitch_subsequences_device``."""
        with torch.cuda.device(self.device):
            host = torch.cat([self._block, self.frame_count]).cpu()
        n = self.max_labels + 1
        state = host[0:4].tolist()
        if state[2]:
            msgs = [m for bit, m in STITCH_ERRORS.items() if state[2] & bit]
            raise AssertionError("device stitch failed: " + "; ".join(msgs))
        lo = host[4:4 + n].tolist()
        hi = host[4 + n:4 + 2 * n].tolist()
        off = 4 + 2 * n
        off += off % 2
        cnt = host[off:off + 2 * n].view(torch.int64).tolist()
        fcount = host[self._block.numel():].tolist()
        counts = defaultdict(lambda: 0)
        spans = defaultdict(lambda: [10000, -1])
        for idx in range(n):
            if cnt[idx] > 0:
                counts[idx - 1] = cnt[idx]
                spans[idx - 1] = [lo[idx], hi[idx]]
        container = self.frame_labels.clone()          # the stitcher may be reset and reused for the next video
        frame_labels = [container[t, :fcount[t]] if fcount[t] >= 0 else None for t in range(self.num_frames)]
        out_labels, out_meta = [], []
        metas_host = torch.stack([m for _, _, _, m in self._subclips]).cpu().tolist() if self._subclips else []
        for i, (frames, labels, _, _) in enumerate(self._subclips):
            total = sum(fcount[t] for t in frames)
            out_labels.append(labels[:total])
            meta = None
            if subseq_meta is not None:
                meta = dict(subseq_meta[i])
                k = len(meta['instance_labels'])
                meta['instance_labels'] = [int(v) for v in metas_host[i][:k]]
            else:
                meta = {'instance_labels': [int(v) for v in metas_host[i] if v >= 0]}
            out_meta.append(meta)
        self.next_track_label = state[0]
        return DeviceTrackContainer(frame_labels, counts, spans), out_labels, out_meta
