"""Seeded synthetic point sets for the clustering parity tests (shared by the golden generator and the tests).

numpy's PCG64 streams are platform independent, so the GPU box regenerates exactly the inputs the golden labels
were computed on in the build container.
"""
import numpy as np

F32 = np.float32


def make_points(seed, n, e, n_free, n_blobs=6, sigma=0.06, noise_frac=0.15, quantize_seediness=0, bw_jitter=0.3,
                seed_floor=0.0):
    """Returns (emb [n,e], bandwidths [n,e-n_free] (already activated, > 0), seediness [n,1]) as fp32."""
    rng = np.random.default_rng(seed)
    v = e - n_free
    centres = rng.uniform(-1.0, 1.0, size=(max(n_blobs, 1), e))
    sig = sigma * rng.uniform(0.6, 1.6, size=(max(n_blobs, 1), e))
    which = rng.integers(0, max(n_blobs, 1), size=n)
    emb = centres[which] + rng.standard_normal((n, e)) * sig[which]
    is_noise = rng.random(n) < noise_frac
    emb[is_noise] = rng.uniform(-1.2, 1.2, size=(int(is_noise.sum()), e))
    # learned bandwidth ~ 1 / (2 sigma)^2 with multiplicative jitter (the model predicts it per point)
    bw = 1.0 / (2.0 * sig[which][:, :v]) ** 2 * np.exp(bw_jitter * rng.standard_normal((n, v)))
    z = (emb - centres[which]) / sig[which]
    seediness = np.exp(-0.5 * (z ** 2).sum(1) / e) * rng.uniform(0.85, 1.0, size=n)
    seediness[is_noise] *= 0.3
    seediness = np.clip(seediness + seed_floor, 0.0, 1.0)
    if quantize_seediness:
        seediness = np.round(seediness * quantize_seediness) / quantize_seediness   # forces ties
    return emb.astype(F32), bw.astype(F32), seediness.astype(F32).reshape(n, 1)


# name -> (make_points kwargs, clusterer kwargs).  Keep N small enough that the whole table runs in seconds.
def case_table():
    cases = {}

    def add(name, pts, **clu):
        base = dict(primary_prob_thresh=0.5, secondary_prob_thresh=0.3, min_seediness_prob=0.8, n_free_dims=0,
                    free_dim_stds=[], max_instances=20, cluster_label_start=1)
        base.update(clu)
        cases[name] = (pts, base)

    s = 1000
    for e, nf in ((2, 0), (3, 0), (3, 1), (4, 0), (4, 2), (5, 2), (5, 3), (8, 0)):
        for n in (1, 57, 1500, 20000):
            s += 1
            fds = [0.3, 0.25, 0.4][:nf]
            add(f"e{e}f{nf}_n{n}", dict(seed=s, n=n, e=e, n_free=nf), n_free_dims=nf, free_dim_stds=fds,
                min_seediness_prob=[0.0, 0.5, 0.8, 0.95][s % 4], cluster_label_start=[1, 7, 100][s % 3])
    # ties in seediness (first-index tie-break), DAVIS / KITTI shaped
    add("ties_xyff", dict(seed=7, n=6000, e=4, n_free=2, quantize_seediness=16), n_free_dims=2,
        free_dim_stds=[0.3, 0.3], min_seediness_prob=0.5)
    add("ties_xyt", dict(seed=8, n=6000, e=3, n_free=0, quantize_seediness=8), min_seediness_prob=0.0)
    # loop exhaustion: more blobs than max_instances -> stale-mask secondary overwrite (quirk iii)
    add("exhaust_small_max", dict(seed=9, n=8000, e=4, n_free=2, n_blobs=12, noise_frac=0.3), n_free_dims=2,
        free_dim_stds=[0.3, 0.3], min_seediness_prob=0.0, max_instances=5)
    add("exhaust_dense", dict(seed=10, n=5000, e=3, n_free=0, n_blobs=1, sigma=0.5, noise_frac=0.0),
        min_seediness_prob=0.0, max_instances=4, primary_prob_thresh=0.9, secondary_prob_thresh=0.05)
    add("exhaust_20", dict(seed=11, n=20000, e=4, n_free=2, n_blobs=40, sigma=0.03, noise_frac=0.4),
        n_free_dims=2, free_dim_stds=[0.3, 0.3], min_seediness_prob=0.0)
    # all points assigned before max_instances (break on n_un == 0 -> no secondary)
    add("all_assigned", dict(seed=12, n=3000, e=4, n_free=2, n_blobs=2, sigma=0.01, noise_frac=0.0),
        n_free_dims=2, free_dim_stds=[0.3, 0.3], min_seediness_prob=0.0, primary_prob_thresh=1e-30)
    # min_seediness stops immediately (no cluster) / after a few
    add("no_cluster", dict(seed=13, n=2000, e=4, n_free=2), n_free_dims=2, free_dim_stds=[0.3, 0.3],
        min_seediness_prob=1.5)
    add("kitti_095", dict(seed=14, n=9000, e=3, n_free=0, seed_floor=0.05), min_seediness_prob=0.95)
    # thresholds at the ends of the range
    add("primary_one", dict(seed=15, n=500, e=4, n_free=0), primary_prob_thresh=1.0, min_seediness_prob=0.0,
        max_instances=3)
    add("secondary_zero", dict(seed=16, n=4000, e=4, n_free=2), n_free_dims=2, free_dim_stds=[0.3, 0.3],
        secondary_prob_thresh=0.0, min_seediness_prob=0.6)
    add("max_instances_0", dict(seed=17, n=100, e=3, n_free=0), max_instances=0)
    add("max_instances_1", dict(seed=18, n=4000, e=4, n_free=2), n_free_dims=2, free_dim_stds=[0.3, 0.3],
        max_instances=1, min_seediness_prob=0.0)
    return cases
