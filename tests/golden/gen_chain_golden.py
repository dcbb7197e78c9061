"""Generate tests/golden/chain_golden.npz with the UNMODIFIED reference OnlineChainer + SequentialClustering on CPU.

Run in the build container only:  python tests/golden/gen_chain_golden.py
Pins (a) oracle/gather_oracle.py against masks_to_coord_list / cluster_subsequence's gather
(online_chainer.py:11-22,258-281), (b) the per-sub-clip cluster labels, and (c) the stitched per-frame track labels,
point counts and lifetimes of OnlineChainer.process (online_chainer.py:143-242, 291-343, 94-117) that
stemseg_b200/chaining.py must reproduce.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from stemseg.inference.clusterers import SequentialClustering  # noqa: E402
from stemseg.inference.online_chainer import OnlineChainer, masks_to_coord_list  # noqa: E402

from chain_cases import CASES, make_video  # noqa: E402
from oracle import gather_oracle as go  # noqa: E402


def main():
    out = {}
    for name, kw in CASES.items():
        masks, subseqs = make_video(**kw)
        # (a) gather oracle vs the reference's index lists + gather
        coords = masks_to_coord_list(torch.from_numpy(masks))
        o_coords, o_counts = go.masks_to_coord_list(masks)
        for (y, x), (oy, ox) in zip(coords, o_coords):
            assert np.array_equal(y.numpy(), oy) and np.array_equal(x.numpy(), ox)
        clusterer = SequentialClustering(0.5, 0.3, 0.5, 2, [0.3, 0.3], "cpu")
        chainer = OnlineChainer(clusterer, embedding_resize_factor=1.0)
        t_subseqs = [{"frames": list(s["frames"]), "embeddings": torch.from_numpy(s["embeddings"]),
                      "bandwidths": torch.from_numpy(s["bandwidths"]), "seediness": torch.from_numpy(s["seediness"])}
                     for s in subseqs]
        # the reference's gather on the first sub-clip vs the oracle
        s0 = t_subseqs[0]
        labels0, emb_flat0, _ = chainer.cluster_subsequence(
            [coords[t] for t in s0["frames"]], s0["embeddings"], s0["bandwidths"], s0["seediness"], 1, True)
        o_emb = go.gather_map([o_coords[t] for t in subseqs[0]["frames"]], subseqs[0]["embeddings"])
        assert np.array_equal(emb_flat0.numpy(), o_emb)
        (track_labels, pt_counts, lifetimes), _, subseq_labels, _, meta = chainer.process(
            torch.from_numpy(masks), t_subseqs)
        for t, lab in enumerate(track_labels):
            out["%s/track/%d" % (name, t)] = lab.numpy().astype(np.int32)
        for i, labs in enumerate(subseq_labels):
            # NOTE: process() relabels non-overlap frames of subseq_labels in place; store them as returned
            out["%s/subseq/%d" % (name, i)] = torch.cat(labs).numpy().astype(np.int32)
        ids = sorted(pt_counts.keys())
        out[name + "/ids"] = np.array(ids, np.int64)
        out[name + "/pt_counts"] = np.array([pt_counts[i] for i in ids], np.int64)
        out[name + "/lifetimes"] = np.array([lifetimes[i] for i in ids], np.int64)
        out[name + "/instance_labels"] = np.array(sum([m["instance_labels"] + [-999] for m in meta], []), np.int64)
        print(name, "frames", masks.shape[0], "subclips", len(subseqs), "tracks", ids, "counts",
              [pt_counts[i] for i in ids])
    path = os.path.join(HERE, "chain_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
