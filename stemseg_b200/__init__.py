"""stemseg_b200 -- B200-native (sm_100a) implementation of STEm-Seg's decoder + clustering hot path.

Host side mirrors the reference's plugin surface (head registries, ``SequentialClustering``); all arithmetic on the
path runs in hand-written CUDA kernels behind the C ABI of ``include/stemseg_b200.h`` (``libstemseg_b200.so``).
There is no CPU / PyTorch fallback: a missing library or a non-sm_100 device raises.
"""
from stemseg_b200._lib import StemsegError, load as load_library  # noqa: F401

__version__ = "0.1.0"
