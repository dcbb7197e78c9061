"""Writeback parity cases: (video of tests/chain_cases.py, original image size, MIN/MAX_DIM of the resize, max_tracks)."""
WRITEBACK_CASES = {
    # network input 96x128 (4 x 24x32); resized (unpadded) 90x120; original image 60x80 (down-sizing)
    "down_90x120_to_60x80": dict(video="three_blobs", image_dims=(60, 80), min_dim=90, max_dim=1000, max_tracks=20),
    # no padding, no resize: identity second stage -> exact dyadic ties at 0.5 must stay "not > 0.5"
    "identity_96x128": dict(video="three_blobs", image_dims=(96, 128), min_dim=96, max_dim=1000, max_tracks=20),
    # up-sizing to an image larger than the network input, max_dim active, max_tracks truncates the instance list
    "up_to_150x200_maxtracks": dict(video="five_blobs_tail", image_dims=(150, 200), min_dim=500, max_dim=157,
                                    max_tracks=4),
    "two_blobs_odd": dict(video="two_blobs_overlap6", image_dims=(77, 101), min_dim=93, max_dim=1000, max_tracks=20),
}
