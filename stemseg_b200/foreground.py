"""Foreground compaction + gather on the device.

Replaces ``masks_to_coord_list`` (stemseg/inference/online_chainer.py:11-22) and the gather half of
``OnlineChainer.cluster_subsequence`` (online_chainer.py:258-281): T ``nonzero`` syncs + 3T advanced-index gathers +
3 ``cat`` become one ordered stream compaction and one gather per map.
"""
import torch

from stemseg_b200 import _lib


class ForegroundIndex(object):
    """Ordered foreground voxel list of a [T,H,W] mask.

    ``indices``: int32 [N] device tensor of linear voxel ids t*H*W + y*W + x in frame-major / row-major order (the
    order torch.nonzero produces per frame); ``frame_counts``: python list of per-frame counts (one D2H copy).
    ``coord_list()`` reproduces the reference's list(T) of (y, x) index tuples (online_chainer.py:11-22).
    """

    def __init__(self, indices, frame_counts, shape, counts_dev=None):
        self._indices = indices              # device int32; capacity T*H*W while the counts are still on the device
        self._frame_counts = frame_counts    # python list, or None until `resolve()` fetched counts_dev
        self.shape = tuple(shape)
        self.counts_dev = counts_dev         # device int32 [T+1] (per-frame counts, then the total) or None

    def resolve(self):
        """Fetch the counts from the device (one sync) if that has not happened yet."""
        if self._frame_counts is None:
            counts = self.counts_dev.cpu().tolist()
            self._frame_counts = counts[:-1]
            self._indices = self._indices[:counts[-1]]
        return self

    @property
    def indices(self):
        return self.resolve()._indices

    @property
    def frame_counts(self):
        return self.resolve()._frame_counts

    @property
    def capacity_indices(self):
        """Index buffer without forcing a sync (valid entries: the first counts_dev[-1])."""
        return self._indices

    @property
    def total_dev(self):
        """Device pointer holder of the total count, or None when the count is already known on the host."""
        return None if self._frame_counts is not None else self.counts_dev[-1:]

    @property
    def num_points(self):
        return int(self.indices.shape[0])

    def frame_slice(self, frames):
        """ForegroundIndex restricted to (and re-based on) the given ascending frame numbers."""
        t, h, w = self.shape
        offsets = [0]
        for c in self.frame_counts:
            offsets.append(offsets[-1] + c)
        parts, counts = [], []
        for j, f in enumerate(frames):
            seg = self.indices[offsets[f]:offsets[f + 1]]
            parts.append(seg + (j - f) * h * w)
            counts.append(self.frame_counts[f])
        idx = torch.cat(parts) if parts else self.indices[:0]
        return ForegroundIndex(idx.to(torch.int32), counts, (len(frames), h, w))

    def coord_list(self):
        t, h, w = self.shape
        out, start = [], 0
        for f, c in enumerate(self.frame_counts):
            lin = self.indices[start:start + c].long() - f * h * w
            out.append((torch.div(lin, w, rounding_mode="floor"), lin % w))
            start += c
        return out


@torch.no_grad()
def compact_foreground(masks, threshold=None, sync=True):
    """masks: [T,H,W] tensor on a CUDA device, any integer/bool dtype (non-zero = foreground).

    With ``threshold`` given, ``masks`` is an fp32 map and foreground = ``masks > threshold`` evaluated inside the
    compaction kernel (seediness > 0.25, stemseg/inference/main.py:93-103).
    """
    if masks.dim() != 3:
        raise ValueError("expected a [T,H,W] mask, got shape %s" % (tuple(masks.shape),))
    if not masks.is_cuda:
        raise ValueError("compact_foreground needs a CUDA tensor; there is no CPU path")
    t, h, w = masks.shape
    if threshold is not None:
        if masks.dtype != torch.float32:
            raise ValueError("thresholded compaction needs an fp32 map")
        m = masks.contiguous()
    elif masks.dtype == torch.bool:
        m = masks.contiguous().view(torch.uint8)
    elif masks.dtype == torch.uint8:
        m = masks.contiguous()
    else:
        m = (masks != 0).contiguous().view(torch.uint8)
    if t * h * w == 0:
        return ForegroundIndex(torch.zeros(0, dtype=torch.int32, device=masks.device), [0] * t, (t, h, w))
    lib = _lib.load()
    with torch.cuda.device(masks.device):
        ws_bytes = lib.stemseg_fg_compact_workspace_bytes(t, h * w)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=masks.device)
        indices = torch.empty(t * h * w, dtype=torch.int32, device=masks.device)
        counts = torch.empty(t + 1, dtype=torch.int32, device=masks.device)
        if threshold is not None:
            _lib.check(lib.stemseg_fg_compact_threshold(_lib.ptr(m), float(threshold), t, h * w, _lib.ptr(indices),
                                                        _lib.ptr(counts), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        else:
            _lib.check(lib.stemseg_fg_compact(_lib.ptr(m), t, h * w, _lib.ptr(indices), _lib.ptr(counts),
                                              _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        if not sync:     # counts stay on the device: gather / clustering read them there (no host round trip)
            return ForegroundIndex(indices, None, (t, h, w), counts_dev=counts)
        counts_host = counts.cpu().tolist()          # one sync (the reference syncs once per frame)
    return ForegroundIndex(indices[:counts_host[-1]], counts_host[:-1], (t, h, w))


@torch.no_grad()
def gather_points(channel_first_map, fg_index, transform="none", upsample=1):
    """channel_first_map: [C,T,H,W] fp32 CUDA tensor -> [N,C] rows for the foreground voxels (online_chainer.py:265-281)."""
    x = channel_first_map
    code = {"none": 0, "exp10": 1}[transform]        # exp10: bandwidths = exp(v) * 10 (inference_model.py:148)
    expect = fg_index.shape if upsample == 1 else (fg_index.shape[0], fg_index.shape[1] // upsample,
                                                   fg_index.shape[2] // upsample)
    if x.dim() != 4 or tuple(x.shape[1:]) != tuple(expect):
        raise ValueError("map shape %s does not match mask shape %s (upsample %d)" % (
            tuple(x.shape), fg_index.shape, upsample))
    if x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("gather_points needs an fp32 CUDA tensor")
    c = x.shape[0]
    inner = x.shape[1] * x.shape[2] * x.shape[3]
    if not x[0].is_contiguous() or (c > 1 and x.stride(0) < inner):
        x = x.contiguous()
    total_dev = fg_index.total_dev
    idx = fg_index.capacity_indices
    n = int(idx.shape[0])            # capacity when the count is still on the device
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    if n == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(x.device):
        if upsample == 1:
            _lib.check(lib.stemseg_fg_gather(_lib.ptr(x), x.stride(0) if c > 1 else inner, c, _lib.ptr(idx), n,
                                             _lib.ptr(total_dev), code, _lib.ptr(out), _lib.stream_ptr()))
        else:   # the (1, s, s) trilinear resize of online_chainer.py:128-140 evaluated only at the foreground voxels
            _lib.check(lib.stemseg_fg_gather_upsampled(_lib.ptr(x), x.stride(0) if c > 1 else inner, c, x.shape[2],
                                                       x.shape[3], int(upsample), _lib.ptr(idx), n,
                                                       _lib.ptr(total_dev), code, _lib.ptr(out), _lib.stream_ptr()))
    return out


class FrameAverager(object):
    """Per-frame running mean of a map over the sub-clips that cover the frame, kept on the device.

    The reference averages the seediness maps (stemseg/inference/main.py:93-103) or the foreground logits
    (inference_model.py:126-128,207) of overlapping sub-clips on the host, frame by frame; here each sub-clip adds its
    [T',h,w] planes into [T,h,w] sums and ``foreground_index`` thresholds mean (optionally up-sampled x4) inside the
    compaction kernel."""

    def __init__(self, num_frames, height, width, device):
        self.device = torch.device(device)
        self.sum = torch.zeros((num_frames, height, width), dtype=torch.float32, device=self.device)
        self.count = torch.zeros(num_frames, dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def add(self, frames, planes):
        """planes: [len(frames), h, w] fp32 CUDA tensor (e.g. seediness[0] of one sub-clip)."""
        if tuple(planes.shape) != (len(frames),) + tuple(self.sum.shape[1:]):
            raise ValueError("planes %s do not match %d frames of %s" % (tuple(planes.shape), len(frames),
                                                                          tuple(self.sum.shape[1:])))
        lib = _lib.load()
        planes = planes.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            ids = torch.tensor(list(frames), dtype=torch.int32, device=self.device)
            _lib.check(lib.stemseg_frame_accumulate(_lib.ptr(self.sum), _lib.ptr(self.count), _lib.ptr(planes),
                                                    _lib.ptr(ids), len(frames), self.sum.shape[1] * self.sum.shape[2],
                                                    _lib.stream_ptr()))

    @torch.no_grad()
    def foreground_index(self, threshold, upsample=1, sync=True):
        """ForegroundIndex of ``upsample(mean) > threshold`` on the [T, s*h, s*w] grid."""
        lib = _lib.load()
        t, h, w = self.sum.shape
        s = int(upsample)
        with torch.cuda.device(self.device):
            ws_bytes = lib.stemseg_fg_compact_workspace_bytes(t, h * s * w * s)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            indices = torch.empty(t * h * s * w * s, dtype=torch.int32, device=self.device)
            counts = torch.empty(t + 1, dtype=torch.int32, device=self.device)
            _lib.check(lib.stemseg_fg_compact_mean_threshold(_lib.ptr(self.sum), _lib.ptr(self.count), float(threshold),
                                                             t, h, w, s, _lib.ptr(indices), _lib.ptr(counts),
                                                             _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
            if not sync:
                return ForegroundIndex(indices, None, (t, h * s, w * s), counts_dev=counts)
            counts_host = counts.cpu().tolist()
        return ForegroundIndex(indices[:counts_host[-1]], counts_host[:-1], (t, h * s, w * s))
