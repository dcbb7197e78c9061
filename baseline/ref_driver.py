"""Drives the UNMODIFIED reference callers (stemseg/inference/main.py:52-170 ``TrackGenerator``, which builds
``InferenceModel`` through ``build_model()`` and owns an ``OnlineChainer`` + ``SequentialClustering``) on synthetic
videos, so that the same calls can be made (a) through the untouched reference on the host cores and (b) with the B200
plugin installed by ``stemseg_b200.registry.install_into_reference()``.

Test / bench infrastructure only (tests/test_reference_gpu.py, tests/test_reference_cpu.py, bench.py's reference arm).
Nothing here re-implements the reference: every call below is one of its own public entry points.
"""
import os

import numpy as np


class SyntheticSequence(object):
    """What TrackGenerator needs from a sequence object (main.py:132-137): base_dir, image_paths, seq_id, len()."""

    def __init__(self, base_dir, image_paths, seq_id="synthetic"):
        self.base_dir, self.image_paths, self.seq_id = base_dir, list(image_paths), seq_id

    def __len__(self):
        return len(self.image_paths)


def synthetic_frames(num_frames, height, width, seed=0):
    """BGR uint8 frames: a smooth background with three moving discs (deterministic for a seed)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    base = rng.integers(0, 255, size=(height // 16 + 2, width // 16 + 2, 3)).astype(np.float32)
    base = np.kron(base, np.ones((16, 16, 1), np.float32))[:height, :width]
    discs = [(rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.08, 0.2), rng.uniform(-0.02, 0.02),
              rng.uniform(-0.02, 0.02), rng.integers(0, 255, size=3)) for _ in range(3)]
    frames = []
    for t in range(num_frames):
        img = base.copy()
        for cy, cx, r, vy, vx, col in discs:
            m = ((yy - (cy + vy * t) * height) ** 2 + (xx - (cx + vx * t) * width) ** 2) <= (r * min(height, width)) ** 2
            img[m] = col
        frames.append(np.clip(img, 0, 255).astype(np.uint8))
    return frames


def write_synthetic_video(directory, num_frames, height, width, seed=0):
    """PNG frames on disk (InferenceImageLoader reads paths with cv2.imread, inference_image_loader.py:27-28)."""
    import cv2
    os.makedirs(directory, exist_ok=True)
    names = []
    for t, frame in enumerate(synthetic_frames(num_frames, height, width, seed)):
        name = "%05d.png" % t
        cv2.imwrite(os.path.join(directory, name), frame)
        names.append(name)
    return SyntheticSequence(directory, names)


class RecordingOutputGenerator(object):
    """Stands where Davis/YoutubeVIS/KittiMOTSOutputGenerator stand (main.py:166-169): keeps what the chainer produced."""

    def __init__(self):
        self.calls = []

    def process_sequence(self, sequence, framewise_mask_idxes, track_labels, instance_pt_counts, instance_lifetimes,
                         multiclass_masks, fg_mask_dims, mask_scale, max_tracks, device=None):
        self.calls.append({
            "mask_idxes": framewise_mask_idxes, "track_labels": [l.cpu() for l in track_labels],
            "instance_pt_counts": dict(instance_pt_counts), "instance_lifetimes": dict(instance_lifetimes),
            "fg_mask_dims": tuple(fg_mask_dims), "max_tracks": max_tracks})


def make_track_generator(sequence, dataset_name, clustering_device, frame_overlap=-1, seediness_thresh=0.25,
                         resize_scale=1.0, cpu_workers=0):
    """TrackGenerator exactly as main.py:266-271 constructs it (random init: model_ckpt_path=None)."""
    from stemseg.inference.main import TrackGenerator
    recorder = RecordingOutputGenerator()
    tg = TrackGenerator([sequence], dataset_name, recorder, os.path.join(sequence.base_dir, "out"), None, 20,
                        False, resize_scale, str(clustering_device) != "cpu", save_vis=False, seediness_thresh=seediness_thresh,
                        frame_overlap=frame_overlap, clustering_device=clustering_device)
    tg.model.cpu_workers = cpu_workers            # DataLoader worker processes are pointless for 12 frames
    return tg, recorder


def configure(config_name, num_frames=8, min_dim=None, max_dim=None, min_seediness_prob=None, overrides=()):
    """cfg as inference/main.py:185-226 sets it up (merge the dataset YAML, then the CLI overrides)."""
    from baseline import refshim
    cfg = refshim.load_config(config_name)
    cfg.INPUT.update_param("NUM_FRAMES", num_frames)
    if min_dim is not None:
        cfg.INPUT.update_param("MIN_DIM", min_dim)
        cfg.INPUT.update_param("MAX_DIM", max_dim)
    if min_seediness_prob is not None:
        cfg.CLUSTERING.update_param("MIN_SEEDINESS_PROB", min_seediness_prob)
    for section, name, value in overrides:
        node = cfg
        for part in section.split("."):
            node = getattr(node, part)
        node.update_param(name, value)
    return cfg
