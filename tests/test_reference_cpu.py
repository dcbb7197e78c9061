"""Host-side checks of the reference arm: the install recipe (baseline/install_reference.py), the import shim and the
driver that runs the UNMODIFIED ``TrackGenerator`` (stemseg/inference/main.py:52-170) on a synthetic PNG video.  The
GPU counterpart (tests/test_reference_gpu.py) makes the same calls with the B200 plugin installed."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import install_reference, refshim  # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


def test_install_recipe_copies_unmodified_tree(tmp_path):
    src = refshim.find_reference_root()
    dest = install_reference.install(source=src, dest=str(tmp_path / "_ref"))
    assert os.path.isfile(os.path.join(dest, "stemseg", "inference", "clusterers.py"))
    assert install_reference.verify(dest)
    with open(os.path.join(dest, "stemseg", "inference", "clusterers.py"), "a") as f:
        f.write("\n# tampered\n")
    assert not install_reference.verify(dest)
    # nothing under baseline/_ref may be tracked by git (no reference source in this repository's history)
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "baseline/_ref/" in f.read().split()


@pytest.mark.parametrize("dataset,config", [("davis", "davis_1.yaml"), ("ytvis", "youtube_vis.yaml")])
def test_unmodified_track_generator_runs_on_cpu(tmp_path, dataset, config):
    import torch
    from baseline import ref_driver
    import stemseg_b200.registry as b200
    refshim.install()
    b200.uninstall_from_reference()
    ref_driver.configure(config, 8, 96, 128, min_seediness_prob=0.0)
    seq = ref_driver.write_synthetic_video(str(tmp_path / "frames"), 10, 96, 128, seed=1)
    with refshim.cpu_only():
        tg, rec = ref_driver.make_track_generator(seq, dataset, "cpu", frame_overlap=4)
        assert type(tg.model._model.embedding_head).__module__ == "stemseg.modeling.embedding_decoder"
        assert type(tg.chainer.clusterer).__module__ == "stemseg.inference.clusterers"
        emb, fg, mc = tg.do_inference(seq)
        tg.do_clustering(seq, emb, fg, mc, 20)
    assert len(emb) == 2 and emb[0].embeddings.shape == (4, 8, 24, 32)
    call = rec.calls[-1]
    assert len(call["track_labels"]) == 10 and call["fg_mask_dims"] == (24, 32)
    assert sum(call["instance_pt_counts"].values()) == int(fg.sum())
    if dataset == "ytvis":
        assert torch.is_tensor(mc) and mc.shape == (10, 41, 24, 32)


def test_install_and_uninstall_restore_the_reference_classes():
    refshim.install()
    import stemseg.inference.clusterers as ref_clusterers
    import stemseg.modeling.model_builder as ref_builder
    from stemseg.modeling.embedding_decoder import EMBEDDING_HEAD_REGISTRY
    import stemseg_b200.registry as b200
    b200.uninstall_from_reference()
    before = (ref_clusterers.SequentialClustering, EMBEDDING_HEAD_REGISTRY["squeeze_expand_decoder"],
              ref_builder.EmbeddingLoss, ref_builder.TrainingModel.compute_fg_loss)
    b200.install_into_reference()
    assert ref_clusterers.SequentialClustering.__module__ == "stemseg_b200.clusterers"
    assert b200.B200_KEY in EMBEDDING_HEAD_REGISTRY._obj_map
    b200.uninstall_from_reference()
    after = (ref_clusterers.SequentialClustering, EMBEDDING_HEAD_REGISTRY["squeeze_expand_decoder"],
             ref_builder.EmbeddingLoss, ref_builder.TrainingModel.compute_fg_loss)
    assert before == after and b200.B200_KEY not in EMBEDDING_HEAD_REGISTRY._obj_map
