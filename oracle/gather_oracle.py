"""CPU restatement (numpy) of the foreground gather.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows ``masks_to_coord_list`` (stemseg/inference/online_chainer.py:11-22) and the gather in
``OnlineChainer.cluster_subsequence`` (online_chainer.py:258-281).  Pure index/byte movement, so parity is
bit-exact.  Pinned against the reference by tests/golden/gen_chain_golden.py.
"""
import numpy as np


def masks_to_coord_list(masks):
    """masks [T,H,W] -> (list(T) of (y, x) int64 index arrays in torch.nonzero (row-major) order, per-frame counts)."""
    coords, counts = [], []
    for t in range(masks.shape[0]):
        y, x = np.nonzero(masks[t])                 # row-major, like torch.nonzero (online_chainer.py:18)
        coords.append((y.astype(np.int64), x.astype(np.int64)))
        counts.append(int(y.shape[0]))
    return coords, counts


def gather_map(coords, channel_first):
    """channel_first [C,T,H,W] -> [N,C]: permute(1,2,3,0), per-frame advanced index, cat (online_chainer.py:258-281)."""
    c = channel_first.shape[0]
    rows = [np.transpose(channel_first[:, t], (1, 2, 0))[y, x] for t, (y, x) in enumerate(coords)]
    return np.concatenate(rows, axis=0).reshape(-1, c) if rows else np.zeros((0, c), channel_first.dtype)


def gather_foreground(coords, embeddings, bandwidths, seediness):
    return gather_map(coords, embeddings), gather_map(coords, bandwidths), gather_map(coords, seediness)


def resize_map(channel_first, factor):
    """online_chainer.py:128-140 / inference_model.py:55-61: trilinear (1, s, s), align_corners=False (torch CPU)."""
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(np.ascontiguousarray(channel_first)).unsqueeze(0)
    if factor != 1:
        x = F.interpolate(x, scale_factor=(1.0, float(factor), float(factor)), mode="trilinear", align_corners=False)
    return x.squeeze(0).numpy()


def averaged_foreground(frame_lists, plane_lists, num_frames, threshold, factor=1):
    """Mean over the sub-clips covering each frame, up-sampled, thresholded (main.py:93-103 with resize_output).
    Returns (mask bool [T,H,W], mean float64 at output resolution for ambiguity checks)."""
    import torch
    import torch.nn.functional as F
    h, w = plane_lists[0].shape[-2:]
    acc = np.zeros((num_frames, h, w), np.float32)
    cnt = np.zeros(num_frames, np.float32)
    for frames, planes in zip(frame_lists, plane_lists):
        for j, t in enumerate(frames):
            acc[t] = acc[t] + planes[j]
            cnt[t] += 1.0
    mean = acc / cnt[:, None, None]
    up = torch.from_numpy(mean).double()[None, None]
    if factor != 1:
        up = F.interpolate(up, scale_factor=(1.0, float(factor), float(factor)), mode="trilinear", align_corners=False)
    up = up[0, 0].numpy()
    return up > threshold, up
