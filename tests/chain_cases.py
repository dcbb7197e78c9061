"""Seeded synthetic video (moving blobs) for the stitching / clip-parallel tests: per-sub-clip embedding maps in the
format InferenceModel produces (inference_model.py:161-162) plus the foreground masks."""
import numpy as np

F32 = np.float32


def subsequence_frames(seq_len, subseq_len, overlap):
    """Same windows as get_subsequence_frames for seq_len >= subseq_len (stemseg/inference/main.py:23-49)."""
    idx = []
    last = -1
    for t in range(0, seq_len - subseq_len + 1, subseq_len - overlap):
        idx.append(list(range(t, t + subseq_len)))
        last = idx[-1][-1]
    if last != seq_len - 1:
        idx.append(list(range(seq_len - subseq_len, seq_len)))
    return idx


def make_video(seed, seq_len=20, subseq_len=8, overlap=4, h=24, w=32, n_blobs=3):
    """-> (masks uint8 [L,h,w], list of dicts {frames, embeddings [4,T,h,w], bandwidths [2,T,h,w], seediness [1,T,h,w]})."""
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    start = rng.uniform(-0.6, 0.6, size=(n_blobs, 2))
    vel = rng.uniform(-0.03, 0.03, size=(n_blobs, 2))
    free = rng.uniform(-1, 1, size=(n_blobs, 2))
    radius = rng.uniform(0.18, 0.3, size=n_blobs)
    inst = np.zeros((seq_len, h, w), np.int32)
    centre = np.zeros((seq_len, n_blobs, 2))
    for t in range(seq_len):
        for k in range(n_blobs):
            cy, cx = start[k] + vel[k] * t
            centre[t, k] = (cy, cx)
            inst[t][(ys - cy) ** 2 + (xs - cx) ** 2 < radius[k] ** 2] = k + 1
    masks = (inst > 0).astype(np.uint8)
    subseqs = []
    for frames in subsequence_frames(seq_len, subseq_len, overlap):
        tlen = len(frames)
        emb = np.zeros((4, tlen, h, w), F32)
        bw = np.zeros((2, tlen, h, w), F32)
        seed_map = np.zeros((1, tlen, h, w), F32)
        # the embedding of an instance is the centre it has in the FIRST frame of the sub-clip (+ its free dims)
        ref_centre = centre[frames[0]]
        for i, t in enumerate(frames):
            noise = rng.standard_normal((4, h, w)) * 0.02
            for k in range(n_blobs):
                m = inst[t] == k + 1
                emb[0, i][m] = ref_centre[k, 0]
                emb[1, i][m] = ref_centre[k, 1]
                emb[2, i][m] = free[k, 0]
                emb[3, i][m] = free[k, 1]
                d2 = (ys - centre[t, k, 0]) ** 2 + (xs - centre[t, k, 1]) ** 2
                seed_map[0, i][m] = np.exp(-4.0 * d2[m] / radius[k] ** 2)
            emb[:, i] += noise.astype(F32)
            bw[:, i] = 1.0 / (2 * 0.05) ** 2
        seed_map = np.clip(seed_map * rng.uniform(0.9, 1.0, size=seed_map.shape), 0, 1).astype(F32)
        subseqs.append({"frames": list(frames), "embeddings": emb, "bandwidths": bw, "seediness": seed_map})
    return masks, subseqs


CASES = {
    "three_blobs": dict(seed=1, seq_len=20, subseq_len=8, overlap=4, n_blobs=3),
    "five_blobs_tail": dict(seed=2, seq_len=23, subseq_len=8, overlap=4, n_blobs=5, h=32, w=40),
    "two_blobs_overlap6": dict(seed=3, seq_len=16, subseq_len=8, overlap=6, n_blobs=2),
}
