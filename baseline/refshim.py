"""Import shim for running the UNMODIFIED reference on this image (torch 2.11, PyYAML 6, no pycocotools / imgaug).

Used by the golden-vector generators (tests/golden/gen_*.py, build container), the reference-integration tests and
`bench.py --impl reference` / `cpu_baseline`.  The reference tree is looked up at $STEMSEG_REFERENCE_ROOT, then
/root/reference (build container), then baseline/_ref (the copy `baseline/install_reference.py` ships to the GPU box).
Three non-invasive shims, reference files untouched:
  1. yaml.load default Loader (PyYAML >= 6 breaks stemseg/config/config.py:183-194)
  2. stub modules for pycocotools / imgaug / tensorboardX (pulled in by stemseg/data/__init__.py:1-6)
  3. CPU execution of callers that hard-code .cuda() (online_chainer.py:174-176, inference_model.py:102):
     `cpu_only()` makes Tensor.cuda / Module.cuda the identity for the duration of a `with` block (permanently on
     hosts without a GPU).
"""
import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def find_reference_root():
    cands = [os.environ.get("STEMSEG_REFERENCE_ROOT"), "/root/reference", os.path.join(HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "stemseg")):
            return c
    return None


REFERENCE_ROOT = find_reference_root()


def available():
    return find_reference_root() is not None


def install():
    root = find_reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (looked at $STEMSEG_REFERENCE_ROOT, /root/reference, baseline/_ref; "
                           "run `python baseline/install_reference.py` in the build container)")
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
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


@contextlib.contextmanager
def cpu_only():
    """Run reference callers that hard-code `.cuda()` on the host cores (the reference's CPU path on a GPU box)."""
    import torch
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def load_config(name):
    """cfg.merge_from_file(<reference>/stemseg/config/<name>) like inference/main.py:190-199 does for a checkpoint's
    config.yaml.  Returns the reference's global cfg."""
    from stemseg.config import cfg
    cfg.merge_from_file(os.path.join(find_reference_root(), "stemseg", "config", name))
    return cfg
