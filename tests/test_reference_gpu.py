"""The UNMODIFIED reference callers running on a B200 with the plugin installed, gated against the untouched reference.

What runs (all of it the reference's own code from baseline/_ref, shipped by baseline/install_reference.py):
``TrackGenerator.__init__`` (stemseg/inference/main.py:52-91: ``InferenceModel`` -> ``build_model()`` through the head
registries, ``create_clusterer``), ``TrackGenerator.do_inference`` (``InferenceModel.forward``,
modeling/inference_model.py:63-194) and ``TrackGenerator.do_clustering`` (``OnlineChainer.process``,
inference/online_chainer.py:143-242).  Three arms on the same synthetic PNG video:

  1. untouched reference on the host cores (``refshim.cpu_only``),
  2. untouched reference on the GPU (torch/cuDNN heads, TF32 off),
  3. ``stemseg_b200.registry.install_into_reference()`` -> same calls, B200 heads + clusterer.

Head tensors: arm 3 vs arm 2 share the torch backbone bit for bit, so only the heads differ: <= 1e-4 norm-wise per
channel (BASELINE north_star).  Arm 3 vs arm 1 additionally carries the CPU-vs-GPU difference of the torch ResNet-101
(out of scope, stays torch): bounded at 1e-3 and printed.  Labels / track ids / point counts: the chainer of arm 1
(reference ``SequentialClustering(device="cpu")``) and of arm 3 are fed the SAME head outputs and must agree exactly.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, NUM_VIDEO_FRAMES, OVERLAP = 128, 240, 12, 4


def _channel_errors(got, ref):
    """max|a-b| / max|ref| per leading channel."""
    out = []
    for c in range(ref.shape[0]):
        scale = max(float(ref[c].abs().max()), 1e-6)
        out.append(float((got[c].double() - ref[c].double()).abs().max()) / scale)
    return out


def _compare_entries(got, ref, tol, what):
    worst = 0.0
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert list(g.subseq_frames) == list(r.subseq_frames)
        pairs = (("embeddings", g.embeddings, r.embeddings),
                 ("variances", (g.bandwidths / 10.).log(), (r.bandwidths / 10.).log()),      # inference_model.py:148
                 ("seediness", g.seediness, r.seediness))
        for name, a, b in pairs:
            assert a.shape == b.shape, (name, a.shape, b.shape)
            errs = _channel_errors(a.cpu(), b.cpu())
            worst = max(worst, max(errs))
            assert max(errs) <= tol, "%s: %s channel errors %s > %g" % (what, name, errs, tol)
    return worst


def _tracks_equal(a, b):
    assert len(a["track_labels"]) == len(b["track_labels"])
    for t, (x, y) in enumerate(zip(a["track_labels"], b["track_labels"])):
        assert torch.equal(x.cpu().long(), y.cpu().long()), "track labels differ in frame %d" % t
    assert a["instance_pt_counts"] == b["instance_pt_counts"]
    assert a["instance_lifetimes"] == b["instance_lifetimes"]


@pytest.fixture()
def reference(cuda_device):
    from baseline import refshim
    if not refshim.available():
        # the tree is git-ignored and travels with the gpurun snapshot (like the built .so); a checkout without it cannot
        # run the reference's callers at all
        pytest.skip("baseline/_ref is missing: run `python baseline/install_reference.py` (or __graft_entry__.build()) "
                    "in the build container before gpurun")
    from baseline import install_reference
    root = refshim.install()
    if os.path.realpath(root) == os.path.realpath(install_reference.DEST):
        assert install_reference.verify(), "baseline/_ref differs from the manifest written at install time"
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import stemseg_b200.registry as b200
    b200.uninstall_from_reference()
    yield refshim
    b200.uninstall_from_reference()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _three_arms(refshim, tmp_path, dataset, config_name):
    from baseline import ref_driver
    import stemseg_b200.registry as b200
    from stemseg_b200 import heads
    from stemseg_b200.clusterers import SequentialClustering as B200Clustering
    ref_driver.configure(config_name, 8, H, W, min_seediness_prob=0.0)   # random init: seediness ~0.5 everywhere
    seq = ref_driver.write_synthetic_video(str(tmp_path / "frames"), NUM_VIDEO_FRAMES, H, W, seed=3)

    with refshim.cpu_only():                                             # arm 1
        tg_cpu, rec_cpu = ref_driver.make_track_generator(seq, dataset, "cpu", frame_overlap=OVERLAP)
        emb_cpu, fg_cpu, mc_cpu = tg_cpu.do_inference(seq)
    assert not next(tg_cpu.model.parameters()).is_cuda

    tg_gpu, _ = ref_driver.make_track_generator(seq, dataset, "cuda:0", frame_overlap=OVERLAP)       # arm 2
    assert next(tg_gpu.model.parameters()).is_cuda
    assert type(tg_gpu.model._model.embedding_head).__module__.startswith("stemseg.")
    emb_gpu, fg_gpu, mc_gpu = tg_gpu.do_inference(seq)

    b200.install_into_reference()                                        # arm 3
    tg_b, rec_b = ref_driver.make_track_generator(seq, dataset, "cuda:0", frame_overlap=OVERLAP)
    model = tg_b.model._model
    assert type(model.embedding_head) is heads.EmbeddingHead
    assert model.seediness_head is None or type(model.seediness_head) is heads.SeedinessHead
    assert model.semseg_head is None or type(model.semseg_head) is heads.SemsegHead
    assert type(tg_b.chainer.clusterer) is B200Clustering                # main.py:84-91 picked the patched name
    # reference checkpoint -> plugin model, strictly, on the device (inference_model.py:27-28)
    res = model.load_state_dict(tg_cpu.model._model.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    emb_b, fg_b, mc_b = tg_b.do_inference(seq)
    return seq, (tg_cpu, rec_cpu, emb_cpu, fg_cpu, mc_cpu), (tg_gpu, emb_gpu, fg_gpu, mc_gpu), \
        (tg_b, rec_b, emb_b, fg_b, mc_b)


def _check_chain(refshim, seq, tg_cpu, rec_cpu, tg_b, rec_b, emb_b, fg_b, mc_b):
    """Same head outputs through (i) unmodified OnlineChainer + reference CPU clusterer, (ii) unmodified OnlineChainer +
    B200 clusterer (the seam at online_chainer.py:283-285), (iii) stemseg_b200.chaining.OnlineChainer."""
    from stemseg_b200.chaining import OnlineChainer as B200Chainer
    tg_b.do_clustering(seq, emb_b, fg_b, mc_b, 20)
    with refshim.cpu_only():
        tg_cpu.do_clustering(seq, emb_b, fg_b, mc_b, 20)
    _tracks_equal(rec_b.calls[-1], rec_cpu.calls[-1])
    assert len(rec_b.calls[-1]["instance_pt_counts"]) > 2, "degenerate clip: nothing was clustered"
    subseqs = [{"frames": list(e.subseq_frames), "embeddings": e.embeddings, "bandwidths": e.bandwidths,
                "seediness": e.seediness} for e in emb_b]
    (labels, counts, lifetimes), _, _, _, _ = B200Chainer(tg_b.chainer.clusterer, 1.0).process(fg_b, subseqs)
    _tracks_equal({"track_labels": labels, "instance_pt_counts": dict(counts), "instance_lifetimes": dict(lifetimes)},
                  rec_cpu.calls[-1])


def test_davis_track_generator(reference, tmp_path):
    seq, (tg_cpu, rec_cpu, emb_cpu, fg_cpu, _), (tg_gpu, emb_gpu, fg_gpu, _), (tg_b, rec_b, emb_b, fg_b, mc_b) = \
        _three_arms(reference, tmp_path, "davis", "davis_1.yaml")
    worst_gpu = _compare_entries(emb_b, emb_gpu, 1e-4, "plugin vs reference-on-GPU")
    worst_cpu = _compare_entries(emb_b, emb_cpu, 1e-3, "plugin vs reference-on-CPU")
    print("davis: worst norm-wise channel error vs reference heads on the same GPU features %.2e, vs the full CPU "
          "reference %.2e" % (worst_gpu, worst_cpu))
    assert fg_b.shape == fg_cpu.shape == (NUM_VIDEO_FRAMES, H // 4, (W + 16) // 4)
    assert float((fg_b != fg_gpu).float().mean()) < 1e-3
    _check_chain(reference, seq, tg_cpu, rec_cpu, tg_b, rec_b, emb_b, fg_b, mc_b)


def test_youtube_vis_track_generator(reference, tmp_path):
    """Semseg-foreground path: embedding head with seediness output + 41(+1)-class semseg head, fg mask from the
    averaged foreground logit (inference_model.py:196-231, main.py:142-147)."""
    seq, (tg_cpu, rec_cpu, emb_cpu, fg_cpu, mc_cpu), (tg_gpu, emb_gpu, fg_gpu, mc_gpu), \
        (tg_b, rec_b, emb_b, fg_b, mc_b) = _three_arms(reference, tmp_path, "ytvis", "youtube_vis.yaml")
    worst_gpu = _compare_entries(emb_b, emb_gpu, 1e-4, "plugin vs reference-on-GPU")
    worst_cpu = _compare_entries(emb_b, emb_cpu, 1e-3, "plugin vs reference-on-CPU")
    assert mc_b.shape == mc_gpu.shape and mc_b.shape[1] == 41            # averaged multi-class logits [T,41,h,w]
    scale = float(mc_gpu.abs().max())
    err = float((mc_b.double() - mc_gpu.double()).abs().max()) / scale
    print("ytvis: heads %.2e (GPU ref) / %.2e (CPU ref), semseg logits %.2e" % (worst_gpu, worst_cpu, err))
    assert err <= 1e-4
    # the foreground mask thresholds a noisy logit at 0: only voxels within the 1e-4 band may flip
    assert float((fg_b != fg_gpu).float().mean()) < 5e-3
    assert 0.02 < float(fg_b.float().mean()) < 0.98
    _check_chain(reference, seq, tg_cpu, rec_cpu, tg_b, rec_b, emb_b, fg_b, mc_b)
