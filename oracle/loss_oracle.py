"""CPU restatement (plain PyTorch, any float dtype, autograd-capable) of the reference's embedding loss.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows, for ONE sequence (the reference trains with
MAX_SAMPLES_PER_GPU = 1, defaults.yaml:20):
  * EmbeddingLoss.forward                      stemseg/modeling/losses/embedding_loss.py:35-157
  * EmbeddingLoss.compute_prob_map             embedding_loss.py:159-178
  * EmbeddingLoss.compute_bandwidth_smoothness_loss   embedding_loss.py:180-185
  * lovasz_hinge_flat / lovasz_grad            stemseg/modeling/losses/_lovasz.py:139-157, :18-31
Pinned against the reference itself by tests/golden/gen_loss_golden.py (same inputs -> identical loss terms and
gradients on CPU fp32, stored as fixtures in tests/golden/loss_golden.npz).

Quirks preserved (the CUDA kernels reproduce them too):
  (i)   instances without a single mask point are dropped from every term (`unique` over the nonzero points, :83-87)
        but the Lovasz target of the n-th KEPT instance is `masks[n]`, indexed by its position in the kept list, not
        by its original id (:128) -- an empty instance in front of a non-empty one shifts the targets;
  (ii)  an instance whose (possibly shifted) target is empty is skipped for the Lovasz and the instance-seediness
        terms but still counted in total_instances (:129-130);
  (iii) the background seediness term divides by ALL background points, ignored ones included (:115-116); no
        background point at all -> mean of an empty tensor = NaN, as in the reference;
  (iv)  seediness regresses to the DETACHED probability (:133).
"""
import torch


def lovasz_grad(gt_sorted):
    """_lovasz.py:18-31."""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1. - intersection / union
    if p > 1:
        jaccard[1:p] = jaccard[1:p] - jaccard[0:-1]
    return jaccard


def lovasz_hinge_flat(logits, labels):
    """_lovasz.py:139-157 (labels: 0/1 tensor [P])."""
    if len(labels) == 0:
        return logits.sum() * 0.
    signs = 2. * labels.to(logits.dtype) - 1.
    errors = 1. - logits * signs
    errors_sorted, perm = torch.sort(errors, dim=0, descending=True)
    gt_sorted = labels[perm]
    grad = lovasz_grad(gt_sorted).to(logits.dtype)
    return torch.dot(torch.relu(errors_sorted), grad)


def prob_map(embeddings, centre_points, bandwidth_points, free_bandwidths):
    """embedding_loss.py:159-178: embeddings [M,E], centre_points [c,E], bandwidth_points [c,V] (activated)."""
    centre = centre_points.mean(dim=0, keepdim=True)
    bw = bandwidth_points.mean(dim=0, keepdim=True)
    if free_bandwidths is not None and free_bandwidths.numel() > 0:
        bw = torch.cat((bw, free_bandwidths.to(bw).reshape(1, -1)), 1)
    return torch.exp(-0.5 * torch.sum(torch.pow(embeddings - centre, 2) * bw, dim=-1))


def embedding_loss_sequence(embeddings, variances, seediness, masks, ignore, free_dim_stds, w_lovasz=1.0,
                            w_variance_smoothness=1.0, w_seediness=1.0, w=1.0):
    """One sequence.  embeddings [M,E], variances [M,V], seediness [M] (float, may require grad); masks [I,M] 0/1
    integer tensor; ignore [M] bool.  -> dict(total, lovasz, variance_smoothness, seediness) of 0-dim tensors."""
    dtype = embeddings.dtype
    free_bw = None
    if len(free_dim_stds) > 0:
        free_bw = 1. / torch.tensor(list(free_dim_stds), dtype=torch.float32) ** 2        # embedding_loss.py:29
    zero = embeddings.sum() * 0
    if masks.numel() == 0 or int(masks.sum()) == 0:                                         # :74-75, :87-89, :136-140
        return {"total": zero, "lovasz": zero, "variance_smoothness": zero, "seediness": zero}
    kept = [i for i in range(masks.shape[0]) if int(masks[i].sum()) > 0]                    # :83-87
    points = [masks[i].nonzero(as_tuple=False)[:, 0] for i in kept]
    inst_emb = [embeddings[p] for p in points]
    inst_var = [variances[p] for p in points]
    inst_seed = [seediness[p] for p in points]
    total_instances = len(kept)

    bg = (masks == 0).all(0).nonzero(as_tuple=False)[:, 0]                                  # :112
    bg_loss = seediness[bg] ** 2
    seed_loss = torch.where(ignore[bg], torch.zeros((), dtype=dtype), bg_loss).mean()       # :113-116

    smooth = 0.
    for v in inst_var:                                                                      # :180-185
        smooth = smooth + torch.pow(v.mean(dim=0, keepdim=True) - v, 2).mean()
    smooth = smooth / float(len(inst_var))

    lovasz = 0.
    for n in range(total_instances):
        probs = prob_map(embeddings, inst_emb[n], inst_var[n].exp() * 10., free_bw)          # :121-124, :127
        logits = probs * 2. - 1.
        target = masks[n]                                                                   # quirk (i)
        if int(target.sum()) == 0:
            continue
        lovasz = lovasz + lovasz_hinge_flat(logits, target)
        inst_probs = probs[points[n]].detach()
        seed_loss = seed_loss + torch.mean((inst_seed[n] - inst_probs) ** 2)                # :133-134
    lovasz = lovasz / total_instances                                                       # :143-145 (batch of 1)
    seed_loss = seed_loss / float(total_instances + 1)
    total = (lovasz * w_lovasz + smooth * w_variance_smoothness + seed_loss * w_seediness) * w
    return {"total": total, "lovasz": lovasz + zero, "variance_smoothness": smooth + zero, "seediness": seed_loss}


# --------------------------------------------------------------------------------------------------------------
# seeded synthetic targets shared by the golden generator and the tests
# --------------------------------------------------------------------------------------------------------------
def seeded_case(seed, t, h, w, embedding_size=4, n_free=2, instances=3, empty_instances=(), ignore_frac=0.05,
                overlap=False):
    """-> dict(out [1, E+V+1, T, H, W] fp32 'head output', masks [I,T,H,W] uint8, ignore [T,H,W] bool).

    Instances are moving ellipses; `empty_instances` lists instance ids left without any point (quirk (i))."""
    import numpy as np
    rng = np.random.default_rng(seed)
    v = embedding_size - n_free
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    masks = np.zeros((instances, t, h, w), dtype=np.uint8)
    taken = np.zeros((t, h, w), dtype=bool)
    for i in range(instances):
        if i in empty_instances:
            continue
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        ry, rx = rng.uniform(0.08, 0.25) * h, rng.uniform(0.08, 0.25) * w
        vy, vx = rng.uniform(-0.02, 0.02) * h, rng.uniform(-0.02, 0.02) * w
        for f in range(t):
            m = ((yy - cy - vy * f) / ry) ** 2 + ((xx - cx - vx * f) / rx) ** 2 <= 1.0
            if not overlap:
                m &= ~taken[f]
            masks[i, f] = m
            taken[f] |= m
    ignore = rng.random((t, h, w)) < ignore_frac
    # a plausible head output: coordinates + noise for the embedding dims, small variances, seediness in (0,1)
    out = np.zeros((1, embedding_size + v + 1, t, h, w), dtype=np.float32)
    out[0, 0] = (yy / max(h - 1, 1) * 2 - 1)[None] + 0.1 * rng.standard_normal((t, h, w))
    out[0, 1] = (xx / max(w - 1, 1) * 2 - 1)[None] + 0.1 * rng.standard_normal((t, h, w))
    for e in range(2, embedding_size):
        out[0, e] = 0.3 * rng.standard_normal((t, h, w))
    out[0, embedding_size:embedding_size + v] = 0.5 * rng.standard_normal((v, t, h, w))
    out[0, embedding_size + v] = rng.random((t, h, w))
    return {"out": torch.from_numpy(out), "masks": torch.from_numpy(masks), "ignore": torch.from_numpy(ignore),
            "embedding_size": embedding_size, "n_free": n_free}


def loss_from_head_output(out, masks, ignore, embedding_size, n_free, free_dim_stds, **weights):
    """out [1, E+V+1, T, H, W] (any float dtype, may require grad) -> the loss dict of embedding_loss_sequence."""
    v = embedding_size - n_free
    flat = out[0].reshape(out.shape[1], -1).t()                                   # [M, C] like the permute at :48
    return embedding_loss_sequence(flat[:, :embedding_size], flat[:, embedding_size:embedding_size + v],
                                   flat[:, embedding_size + v], masks.reshape(masks.shape[0], -1).long(),
                                   ignore.reshape(-1), free_dim_stds, **weights)


# --------------------------------------------------------------------------------------------------------------
# semantic-segmentation head losses (YouTube-VIS / KITTI-MOTS configs)
# --------------------------------------------------------------------------------------------------------------
def semseg_losses_sequence(head_out, semseg_masks, ignore, foreground_channel=True):
    """One sequence.  head_out [C, T, H, W] = the semseg head's output (class logits, then the foreground logit when
    foreground_channel); semseg_masks [T, H, W] int64 class ids; ignore [T, H, W] bool.
    -> dict(semseg, foreground) following
      * CrossEntropyLoss.forward     stemseg/modeling/losses/cross_entropy.py:13-49
      * TrainingModel.compute_fg_loss  stemseg/modeling/model_builder.py:210-244
      * the permute / split in front of them  model_builder.py:180, :121-122
    Quirk preserved: F.cross_entropy is called with its default 'mean' reduction, so the ignore mask multiplies a SCALAR
    and cancels out -- ignored voxels count in the class loss (cross_entropy.py:36-42); only the foreground loss really
    masks them."""
    import torch.nn.functional as F
    logits = head_out.permute(1, 0, 2, 3)                                       # [T, C, H, W]  (model_builder.py:180)
    fg_logits = None
    if foreground_channel:
        logits, fg_logits = logits.split((logits.shape[1] - 1, 1), dim=1)       # model_builder.py:121
        fg_logits = fg_logits.squeeze(1)
    nonignore = 1. - ignore.to(head_out.dtype)
    seq = F.cross_entropy(logits, semseg_masks)
    seq = seq * nonignore
    out = {"semseg": seq.sum() / nonignore.sum().detach(), "foreground": None}
    if fg_logits is not None:
        fg_target = (semseg_masks > 0).to(head_out.dtype)
        bce = F.binary_cross_entropy_with_logits(fg_logits, fg_target, reduction="none")
        out["foreground"] = (bce * nonignore).sum() / nonignore.sum().detach()
    return out


def seeded_semseg_case(seed, t, h, w, num_classes, foreground_channel=True, ignore_frac=0.1):
    """-> dict(out [1, C, T, H, W] fp32 'semseg head output', semseg_masks [T,H,W] int64, ignore [T,H,W] bool)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    c = num_classes + (1 if foreground_channel else 0)
    out = (1.5 * rng.standard_normal((1, c, t, h, w))).astype(np.float32)
    # blocky class map with a large background (class 0) share
    coarse = rng.integers(0, num_classes, size=(t, (h + 3) // 4, (w + 3) // 4))
    coarse[rng.random(coarse.shape) < 0.5] = 0
    masks = np.repeat(np.repeat(coarse, 4, axis=1), 4, axis=2)[:, :h, :w].astype(np.int64)
    ignore = rng.random((t, h, w)) < ignore_frac
    return {"out": torch.from_numpy(out), "semseg_masks": torch.from_numpy(masks), "ignore": torch.from_numpy(ignore),
            "num_classes": num_classes, "foreground_channel": foreground_channel}
