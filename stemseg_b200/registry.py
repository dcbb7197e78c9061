"""Plugin tables with the reference's ``GlobalRegistry`` interface (stemseg/utils/global_registry.py:1-74).

The reference selects its heads by name through three registries ("EmbeddingHead", "SeedinessHead", "SemsegHead",
model_builder.py:282,309,325).  ``install_into_reference()`` puts the B200 heads into those tables -- under the
reference's own key ``squeeze_expand_decoder`` (replacing the torch implementation, so unmodified YAML configs and
``build_model()`` pick the CUDA heads) and under ``squeeze_expand_decoder_b200``.
"""


class GlobalRegistry(object):
    _REGISTRIES = dict()

    def __init__(self, name):
        self._name = name
        self._obj_map = dict()

    def __getitem__(self, name):
        if name not in self._obj_map:
            raise KeyError("No object with name '{}' is registered under '{}'".format(name, self._name))
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    @staticmethod
    def exists(name):
        return name in GlobalRegistry._REGISTRIES

    @staticmethod
    def get(name):
        if name not in GlobalRegistry._REGISTRIES:
            GlobalRegistry._REGISTRIES[name] = GlobalRegistry(name)
        return GlobalRegistry._REGISTRIES[name]

    @staticmethod
    def register(registry_name, obj_name=None, obj=None):
        return GlobalRegistry.get(registry_name).add(obj_name, obj)

    def _do_register(self, name, obj):
        assert (name not in self._obj_map), \
            "An object named '{}' was already registered in '{}' registry!".format(name, self._name)
        self._obj_map[name] = obj

    def add(self, name=None, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(name if name is not None else func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(name if name else obj.__name__, obj)
        return None


EMBEDDING_HEAD_REGISTRY = GlobalRegistry.get("EmbeddingHead")
SEEDINESS_HEAD_REGISTRY = GlobalRegistry.get("SeedinessHead")
SEMSEG_HEAD_REGISTRY = GlobalRegistry.get("SemsegHead")

B200_KEY = "squeeze_expand_decoder_b200"


_SAVED = []           # (object, attribute or key, previous value or _MISSING, is_mapping) for uninstall_from_reference
_MISSING = object()


def _swap(obj, name, value, mapping=False):
    if mapping:
        _SAVED.append((obj, name, obj.get(name, _MISSING), True))
        obj[name] = value
    else:
        _SAVED.append((obj, name, getattr(obj, name, _MISSING), False))
        setattr(obj, name, value)


def uninstall_from_reference():
    """Undo install_into_reference(): put the reference's own classes back (tests use this to keep arms separate)."""
    while _SAVED:
        obj, name, prev, mapping = _SAVED.pop()
        if mapping:
            if prev is _MISSING:
                obj.pop(name, None)
            else:
                obj[name] = prev
        elif prev is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, prev)


def install_into_reference(replace_default=True):
    """Register the B200 heads and clusterer inside an importable reference tree (``import stemseg``)."""
    if _SAVED:
        return
    from stemseg.utils.global_registry import GlobalRegistry as RefRegistry
    import stemseg.modeling.embedding_decoder  # noqa: F401  (fills the reference tables first)
    import stemseg.modeling.seediness_decoder  # noqa: F401
    import stemseg.modeling.semseg_decoder  # noqa: F401
    from stemseg_b200 import heads
    for table, cls in (("EmbeddingHead", heads.EmbeddingHead), ("SeedinessHead", heads.SeedinessHead),
                       ("SemsegHead", heads.SemsegHead)):
        reg = RefRegistry.get(table)
        _swap(reg._obj_map, B200_KEY, cls, mapping=True)
        if replace_default:
            _swap(reg._obj_map, "squeeze_expand_decoder", cls, mapping=True)
    import stemseg.inference.clusterers as ref_clusterers
    from stemseg_b200.clusterers import SequentialClustering
    _swap(ref_clusterers, "SequentialClustering", SequentialClustering)
    try:
        import stemseg.inference.main as ref_main
        _swap(ref_main, "SequentialClustering", SequentialClustering)
    except Exception:        # the CLI module needs dataset dependencies that may be absent
        pass
    # training: build_model() instantiates the name `EmbeddingLoss` imported into model_builder (model_builder.py:5,294)
    from stemseg_b200.losses import EmbeddingLoss
    import stemseg.modeling.losses as ref_losses
    import stemseg.modeling.model_builder as ref_builder
    _swap(ref_losses, "EmbeddingLoss", EmbeddingLoss)
    _swap(ref_builder, "EmbeddingLoss", EmbeddingLoss)
    # semseg head losses: the registry entry build_model() looks up (model_builder.py:26,333-335) and the method the
    # model calls for the foreground channel (model_builder.py:122,210-244)
    from stemseg_b200.losses import CrossEntropyLoss, compute_fg_loss
    _swap(ref_losses, "CrossEntropyLoss", CrossEntropyLoss)
    _swap(ref_builder, "CrossEntropyLoss", CrossEntropyLoss)
    _swap(ref_builder.SEMSEG_LOSS_REGISTRY._obj_map, "CrossEntropy", CrossEntropyLoss, mapping=True)
    _swap(ref_builder.TrainingModel, "compute_fg_loss",
          lambda self, fg_logits, targets, output_dict: compute_fg_loss(fg_logits, targets, output_dict))
