"""CPU restatement (numpy, fp32, explicit evaluation order) of the reference's sequential clustering.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows, statement by statement:
  * ``SequentialClustering._process``            stemseg/inference/clusterers.py:60-166
  * ``SequentialClustering._get_next_instance_center``   clusterers.py:168-175
  * ``compute_distance`` / ``distances_to_prob``  clusterers.py:53-58

Arithmetic contract (what "bit-exact" means for this path; DESIGN.md "Clustering semantics"):
  * every elementwise op is an IEEE-754 binary32 op with round-to-nearest-even, NO fused multiply-add;
  * the sum over the embedding dimension uses ATen's CPU order for a contiguous inner reduction
    (aten/src/ATen/native/cpu/SumKernel.cpp, torch 2.11, as dispatched in the build container: 4 scalar
    accumulators for E < 8, one 8-lane vector + scalar tail for 8 <= E < 16, ...) -- see ``aten_inner_sum_f32``;
    pinned by tests/golden/gen_cluster_golden.py against torch itself for E = 1..24;
  * ``sqrt`` is the correctly rounded IEEE sqrt and the probability tests ``exp(-0.5 d) > p`` are evaluated with a
    correctly rounded ``exp``.  The reference's own CPU path (MKL-VML sqrt: 0.6 % of results 1 ulp low; Sleef
    expf: 1.1 % of results 1 ulp off) and its CUDA path disagree with each other in the last ulp, so points
    within 1 ulp of a decision threshold are *ambiguous in the reference itself*; golden vectors are generated
    with a >= 8 ulp margin check (tests/golden/gen_cluster_golden.py) so that they are unambiguous.
  * comparisons against python-float thresholds happen in fp32 (the scalar is rounded to fp32), as torch does.
"""
import math

import numpy as np

F32 = np.float32
UNASSIGNED_DISTANCE = F32(1e8)   # clusterers.py:128


def aten_inner_sum_f32(t):
    """Sum a [N, E] fp32 array over E in the order ATen's CPU sum kernel uses (see module docstring)."""
    t = np.ascontiguousarray(t, dtype=F32)
    n, e = t.shape
    if e == 0:
        return np.zeros(n, F32)

    def add(a, b):
        return (a + b).astype(F32)

    def row_sum(cols):
        # SumKernel.cpp row_sum(): ilp_factor = 4 partial sums, tail into partial 0, then partial0 += partial k.
        # `cols` is a list of arrays (scalars per row, or 8-lane vectors per row).
        size = len(cols)
        size_ilp = size // 4
        if size_ilp > 16:
            raise NotImplementedError("cascade levels of multi_row_sum not restated (E too large)")
        partial = [np.zeros_like(cols[0]) for _ in range(4)]
        for i in range(size_ilp):
            for k in range(4):
                partial[k] = add(partial[k], cols[4 * i + k])
        for i in range(size_ilp * 4, size):
            partial[0] = add(partial[0], cols[i])
        for k in range(1, 4):
            partial[0] = add(partial[0], partial[k])
        return partial[0]

    vec = 8   # Vectorized<float>::size() of the SumKernel build that is dispatched (AVX2 variant)
    if e < vec:
        return row_sum([t[:, k] for k in range(e)])
    nvec = e // vec
    vec_acc = row_sum([t[:, v * vec:(v + 1) * vec] for v in range(nvec)])     # [N, 8]
    acc = np.zeros(n, F32)
    for k in range(nvec * vec, e):        # scalar tail first
        acc = add(acc, t[:, k])
    for k in range(vec):                  # then the 8 lanes, in lane order
        acc = add(acc, vec_acc[:, k])
    return acc


def compute_distance(emb, center, bandwidth):
    """clusterers.py:57-58: sqrt(sum_e (emb_e - c_e)^2 * bw_e); fp32, no FMA, ATen sum order, IEEE sqrt."""
    diff = (emb - center[None, :]).astype(F32)
    sq = (diff * diff).astype(F32)
    term = (sq * bandwidth[None, :]).astype(F32)
    return np.sqrt(aten_inner_sum_f32(term)).astype(F32)


def _prob_f32(d):
    """fl32(exp(-0.5 * d)) with a correctly rounded exp (via double)."""
    x = float(F32(-0.5) * F32(d))
    try:
        return F32(math.exp(x))
    except OverflowError:
        return F32(np.inf)


def prob_threshold_to_distance(p):
    """Largest fp32 d >= 0 with fl32(exp(-0.5 d)) > fl32(p); -1 if none, +inf if all (incl. d = +inf) pass.

    The CUDA path and this oracle both test ``d <= d*`` instead of ``exp(-0.5 d) > p`` (clusterers.py:53-54,
    136-138, 156-157): the map is monotone so the two are identical for a correctly rounded exp.
    """
    p32 = F32(p)
    if np.isnan(p32):
        return F32(-1.0)
    if not (_prob_f32(F32(0.0)) > p32):
        return F32(-1.0)
    if F32(0.0) > p32:          # exp(-inf) = 0 > p
        return F32(np.inf)
    lo = 0                                   # bit pattern of +0.0: passes
    hi = int(np.array(np.inf, F32).view(np.uint32))   # +inf: fails (0 > p false since p >= 0)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        d = np.array(mid, np.uint32).view(F32)
        if _prob_f32(d) > p32:
            lo = mid
        else:
            hi = mid
    return np.array(lo, np.uint32).view(F32)[()]


def free_dim_bandwidths(free_dim_stds):
    """clusterers.py:100-102: 1 / std^2 in fp32."""
    s = np.asarray(list(free_dim_stds), dtype=F32)
    return (F32(1.0) / (s * s).astype(F32)).astype(F32)


def bandwidth_to_std(bandwidth):
    """clusterers.py:125: sqrt(clamp(1 / bw, min=1e-8))."""
    inv = (F32(1.0) / bandwidth.astype(F32)).astype(F32)
    return np.sqrt(np.maximum(inv, F32(1e-8))).astype(F32)


def first_argmax(values, mask):
    """argmax over ``values[mask]`` with torch's semantics (NaN is the maximum, first index wins), returned as an
    index into the *uncompacted* array.  clusterers.py:112-114,174 (argmax on the compacted array picks the same
    element because boolean-mask compaction preserves order)."""
    idx = np.nonzero(mask)[0]
    v = values[idx]
    nan = np.isnan(v)
    if nan.any():
        return int(idx[np.argmax(nan)])
    return int(idx[np.argmax(v)])


def sequential_cluster(embeddings, bandwidths, seediness, primary_prob_thresh, secondary_prob_thresh,
                       min_seediness_prob, n_free_dims, free_dim_stds, max_instances=20, cluster_label_start=1,
                       return_label_masks=False):
    """Restatement of SequentialClustering._process (clusterers.py:60-166).

    embeddings [N,E] f32, bandwidths [N,E-n_free] f32 (already ``exp()*10``-activated, inference_model.py:148),
    seediness [N,1] or [N] f32.  Returns (labels int64 [N], meta dict) like the reference; additionally
    meta['margin_ulps'] = smallest distance (in fp32 ulps) of any decision from its threshold (for the golden
    generator's ambiguity check).
    """
    emb = np.ascontiguousarray(embeddings, dtype=F32)
    n, e = emb.shape
    meta = {'instance_labels': [], 'instance_centers': [], 'instance_stds': [], 'instance_masks': [],
            'margin_ulps': float('inf')}
    if emb.size == 0:                                                     # clusterers.py:62-69
        return np.zeros(0, np.int64), meta
    bw = np.ascontiguousarray(bandwidths, dtype=F32)
    if bw.shape[0] != n:                                                  # clusterers.py:75-76
        bw = np.broadcast_to(bw, emb.shape).copy()
    seed = np.ascontiguousarray(seediness, dtype=F32).reshape(n)
    d_primary = prob_threshold_to_distance(primary_prob_thresh)
    d_secondary = prob_threshold_to_distance(secondary_prob_thresh)
    min_seed32 = F32(min_seediness_prob)
    free_bw = free_dim_bandwidths(free_dim_stds) if n_free_dims > 0 else np.zeros(0, F32)

    labels = np.full(n, -1, np.int64)                                     # clusterers.py:96
    label_distances = []
    avail = None
    num_unassigned = n
    margin = float('inf')

    def ulps(d, thr):
        if not np.isfinite(thr) or thr < 0:
            return float('inf')
        finite = np.isfinite(d)
        if not finite.any():
            return float('inf')
        a = d[finite].astype(F32).view(np.int32).astype(np.int64)
        b = int(np.array(thr, F32).view(np.int32))
        # d <= thr passes, d >= nextafter(thr) fails: distance to the boundary between thr and its successor
        return float(np.min(np.where(a <= b, b - a, a - b - 1)))

    for i in range(max_instances):                                        # clusterers.py:106
        avail = labels == -1
        num_unassigned = int(avail.sum())
        if num_unassigned == 0:
            break
        j = first_argmax(seed, avail)                                     # clusterers.py:112-114
        center, prob = emb[j], seed[j]
        if prob < min_seed32:                                             # clusterers.py:116-117
            break
        bandwidth = np.concatenate([bw[j], free_bw]).astype(F32)          # clusterers.py:119
        label = i + cluster_label_start
        meta['instance_labels'].append(label)
        meta['instance_centers'].append(center.tolist())
        meta['instance_stds'].append(bandwidth_to_std(bandwidth).tolist())
        d = np.full(n, UNASSIGNED_DISTANCE, F32)                          # clusterers.py:128-130
        d[avail] = compute_distance(emb[avail], center, bandwidth)
        label_distances.append(d)
        match = avail & (d <= d_primary)                                  # clusterers.py:136-138 (see docstring)
        margin = min(margin, ulps(d[avail], d_primary))
        labels[match] = label                                             # clusterers.py:143
        if return_label_masks:
            meta['instance_masks'].append(match.copy())

    if num_unassigned > 0 and label_distances:                            # clusterers.py:149
        dist = np.stack(label_distances, axis=1)                          # [N, K]
        has_nan = np.isnan(dist).any(axis=1)
        kmax = np.argmax(np.where(np.isnan(dist), -np.inf, dist), axis=1)  # first max (clusterers.py:153)
        dmax = dist[np.arange(n), kmax]
        update = avail & (dmax <= d_secondary) & ~has_nan                 # `avail` is the STALE mask (quirk iii)
        margin = min(margin, ulps(dmax[avail], d_secondary))
        if dist.shape[1] > 1 and avail.any():
            srt = np.sort(dist[avail], axis=1)
            gap = (srt[:, -1].view(np.int32).astype(np.int64) - srt[:, -2].view(np.int32).astype(np.int64))
            near = dmax[avail] <= d_secondary * F32(1.001)
            near &= gap > 0          # exact ties resolve to the first index in every implementation
            if near.any():
                margin = min(margin, float(np.min(gap[near])))
        labels = np.where(update, kmax + cluster_label_start, labels)     # clusterers.py:154,159
    meta['margin_ulps'] = margin
    return labels, meta
