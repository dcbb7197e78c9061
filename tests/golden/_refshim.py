"""Import shim for running the UNMODIFIED reference (/root/reference) on CPU in the build container.

Only used by the golden-vector generator scripts in this directory (never on the GPU box, where
/root/reference does not exist).  Three non-invasive shims, reference tree untouched:
  1. yaml.load default Loader (PyYAML >= 6 breaks stemseg/config/config.py:183-194)
  2. stub modules for pycocotools / imgaug / tensorboardX (pulled in by stemseg/data/__init__.py:1-6)
  3. torch.Tensor.cuda -> identity on CPU-only hosts (online_chainer.py:174-176, inference_model.py:102)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STEMSEG_REFERENCE_ROOT", "/root/reference")


def install():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "stemseg")):
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import yaml
    if not getattr(yaml.load, "_shimmed", False):
        _orig = yaml.load

        def _load(stream, Loader=None, **kw):
            return _orig(stream, Loader=Loader or yaml.FullLoader, **kw)
        _load._shimmed = True
        yaml.load = _load

    def _stub(name):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(_stub(parent), child, m)
        return m

    for n in ("pycocotools", "pycocotools.mask", "imgaug", "imgaug.augmenters", "imgaug.augmentables",
              "imgaug.augmentables.segmaps", "tensorboardX"):
        try:
            __import__(n)
        except Exception:
            _stub(n)
    sm = sys.modules["imgaug.augmentables.segmaps"]
    if not hasattr(sm, "SegmentationMapsOnImage"):
        sm.SegmentationMapsOnImage = object
    tb = sys.modules["tensorboardX"]
    if not hasattr(tb, "SummaryWriter"):
        tb.SummaryWriter = object
    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
