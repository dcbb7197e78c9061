"""Instance-mask writeback on the device: per-point track labels -> full-resolution uint8 instance-id maps.

Mirror of the mask-generation half of the reference's output generators (stemseg/inference/output_utils/davis.py:38-112;
youtube_vis.py:117-161 and kitti_mots.py:101-166 use the same scatter -> one-hot -> x4 bilinear -> crop -> resize ->
"> 0.5" chain; their per-instance binary masks are ``id_map == rank + 1``).  File formats (PNG palettes, RLE JSON,
MOTS txt) stay out of scope.
"""
import torch

from stemseg_b200 import _lib
from stemseg_b200.foreground import ForegroundIndex


def compute_resize_params_2(image_dims_wh, min_resize_dim, max_resize_dim):
    """stemseg/data/common.py:142-159: network-input size (without zero padding) of an image of (width, height)."""
    lower_size = float(min(image_dims_wh))
    higher_size = float(max(image_dims_wh))
    scale_factor = min_resize_dim / lower_size
    if (higher_size * scale_factor) > max_resize_dim:
        scale_factor = max_resize_dim / higher_size
    width, height = image_dims_wh
    return round(scale_factor * width), round(scale_factor * height), scale_factor


def select_instances(instance_lifetimes, outlier_label, max_tracks):
    """davis.py:58-65: track ids by lifetime (descending, stable in dict order), outlier dropped, first max_tracks."""
    assert max_tracks < 256
    ordered = [k for k, _ in sorted([(k, v) for k, v in instance_lifetimes.items()], key=lambda x: x[1], reverse=True)
               if k != outlier_label]
    return ordered[:max_tracks]


@torch.no_grad()
def instance_id_maps(track_mask_idxes, track_mask_labels, instance_lifetimes, mask_dims, mask_scale, image_dims,
                     min_dim, max_dim, max_tracks, outlier_label=-1, device=None, upscaled_inputs=False):
    """-> (uint8 CUDA tensor [T, image_h, image_w] with 0 = background and n+1 = n-th kept instance, kept ids).

    track_mask_idxes: ``ForegroundIndex`` of the whole video or the reference's list(T) of (y, x) index tensors;
    track_mask_labels: list(T) of per-frame label tensors (the stitched track ids)."""
    mask_h, mask_w = mask_dims
    image_h, image_w = image_dims
    keep = select_instances(instance_lifetimes, outlier_label, max_tracks)
    frames = len(track_mask_labels)
    if device is None:
        device = track_mask_labels[0].device if track_mask_labels[0].is_cuda else torch.device("cuda")
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("instance_id_maps runs on a B200 only; there is no CPU path")
    if isinstance(track_mask_idxes, ForegroundIndex):
        assert track_mask_idxes.shape == (frames, mask_h, mask_w)
        indices = track_mask_idxes.indices.to(device)
    else:
        assert len(track_mask_idxes) == frames
        parts = [(t * mask_h + y.to(device).long()) * mask_w + x.to(device).long()
                 for t, (y, x) in enumerate(track_mask_idxes)]
        indices = torch.cat(parts).to(torch.int32)
    labels = torch.cat([l.to(device).long() for l in track_mask_labels]).contiguous()
    assert labels.numel() == indices.numel()
    up = 1 if upscaled_inputs else int(mask_scale)
    assert float(up) == float(mask_scale) or upscaled_inputs, "integer mask scale expected"
    resized_w, resized_h, _ = compute_resize_params_2((image_w, image_h), min_dim, max_dim)
    if mask_w * up < resized_w or mask_h * up < resized_h:                      # davis.py:91-96
        raise RuntimeError("Network input dims without padding {} should be <= padded dims".format(
            (resized_w, resized_h), (mask_h * up, mask_w * up)))
    nlut = (max(keep) + 1) if keep else 1
    lut_host = [0] * nlut
    for rank, inst in enumerate(keep):
        if inst >= 0:
            lut_host[inst] = rank + 1
    lib = _lib.load()
    with torch.cuda.device(device):
        lut = torch.tensor(lut_host, dtype=torch.uint8, device=device)
        rank_map = torch.empty((frames, mask_h, mask_w), dtype=torch.uint8, device=device)
        _lib.check(lib.stemseg_rank_map_scatter(_lib.ptr(indices.contiguous()), _lib.ptr(labels), labels.numel(),
                                                _lib.ptr(lut), nlut, _lib.ptr(rank_map), rank_map.numel(),
                                                _lib.stream_ptr()))
        out = torch.empty((frames, image_h, image_w), dtype=torch.uint8, device=device)
        _lib.check(lib.stemseg_mask_writeback(_lib.ptr(rank_map), frames, mask_h, mask_w, up, resized_h, resized_w,
                                              image_h, image_w, _lib.ptr(out), _lib.stream_ptr()))
    return out, keep
