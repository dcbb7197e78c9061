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
