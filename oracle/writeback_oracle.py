"""CPU restatement (torch fp32) of the instance-mask writeback.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows DavisOutputGenerator.process_sequence (stemseg/inference/output_utils/davis.py:38-112; the same chain is in
youtube_vis.py:117-161 and kitti_mots.py:101-166) and compute_resize_params_2 (stemseg/data/common.py:142-159).
Pinned by tests/golden/gen_writeback_golden.py against the PNGs the reference generator writes.
"""
import numpy as np
import torch
import torch.nn.functional as F


def compute_resize_params_2(image_dims_wh, min_resize_dim, max_resize_dim):
    lower, higher = float(min(image_dims_wh)), float(max(image_dims_wh))
    scale = min_resize_dim / lower
    if higher * scale > max_resize_dim:
        scale = max_resize_dim / higher
    width, height = image_dims_wh
    return round(scale * width), round(scale * height), scale


def instances_to_keep(instance_lifetimes, outlier_label, max_tracks):
    """davis.py:58-65: ids by lifetime, descending, stable in dict order; outlier dropped; first max_tracks."""
    ordered = [k for k, _ in sorted([(k, v) for k, v in instance_lifetimes.items()], key=lambda x: x[1], reverse=True)
               if k != outlier_label]
    return ordered[:max_tracks]


def id_maps(track_mask_idxes, track_mask_labels, keep, mask_dims, mask_scale, image_dims, min_dim, max_dim):
    """-> uint8 [T, image_h, image_w] condensed instance-id maps (davis.py:76-112)."""
    mask_h, mask_w = mask_dims
    image_h, image_w = image_dims
    out = []
    for (ys, xs), labels in zip(track_mask_idxes, track_mask_labels):
        m = torch.zeros(mask_h, mask_w, dtype=torch.long)
        m[torch.as_tensor(ys), torch.as_tensor(xs)] = torch.as_tensor(labels)
        if keep:
            one_hot = torch.stack([m == ii for ii in keep], 0).unsqueeze(0).float()
            one_hot = F.interpolate(one_hot, scale_factor=mask_scale, mode="bilinear", align_corners=False)
            rw, rh, _ = compute_resize_params_2((image_w, image_h), min_dim, max_dim)
            assert one_hot.shape[3] >= rw and one_hot.shape[2] >= rh
            one_hot = one_hot[:, :, :rh, :rw]
            on = (F.interpolate(one_hot, (image_h, image_w), mode="bilinear", align_corners=False) > 0.5)[0]
        else:
            on = torch.zeros((0, image_h, image_w), dtype=torch.bool)
        condensed = torch.zeros(image_h, image_w, dtype=torch.uint8)
        for n in range(len(keep)):
            condensed = torch.where(on[n], torch.tensor(n + 1, dtype=torch.uint8), condensed)
        out.append(condensed)
    return torch.stack(out, 0).numpy()


def threshold_margin(track_mask_idxes, track_mask_labels, keep, mask_dims, mask_scale, image_dims, min_dim, max_dim):
    """Smallest |value - 0.5| over pixels whose value is not an exact dyadic tie (for the golden generator)."""
    mask_h, mask_w = mask_dims
    image_h, image_w = image_dims
    best = 1.0
    for (ys, xs), labels in zip(track_mask_idxes, track_mask_labels):
        m = torch.zeros(mask_h, mask_w, dtype=torch.long)
        m[torch.as_tensor(ys), torch.as_tensor(xs)] = torch.as_tensor(labels)
        if not keep:
            continue
        one_hot = torch.stack([m == ii for ii in keep], 0).unsqueeze(0).double()
        one_hot = F.interpolate(one_hot, scale_factor=mask_scale, mode="bilinear", align_corners=False)
        rw, rh, _ = compute_resize_params_2((image_w, image_h), min_dim, max_dim)
        val = F.interpolate(one_hot[:, :, :rh, :rw], (image_h, image_w), mode="bilinear", align_corners=False)
        d = (val - 0.5).abs()
        d = d[d > 0]
        if d.numel():
            best = min(best, float(d.min()))
    return best
