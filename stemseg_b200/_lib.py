"""ctypes binding of include/stemseg_b200.h.  Fails loudly: no library -> ImportError, no fallback of any kind."""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libstemseg_b200.so")

STEMSEG_MAX_EMBEDDING_DIMS = 16
STEMSEG_MAX_INSTANCES = 64
ABI_VERSION = 1

c_void_p, c_size_t, c_int32, c_int64, c_float, c_double = (
    ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double)


class StemsegClusterParams(ctypes.Structure):
    _fields_ = [
        ("n_points", c_int64),
        ("embedding_dims", c_int32),
        ("n_free_dims", c_int32),
        ("free_dim_bandwidths", c_float * STEMSEG_MAX_EMBEDDING_DIMS),
        ("d_primary", c_float),
        ("d_secondary", c_float),
        ("min_seediness_prob", c_float),
        ("max_instances", c_int32),
        ("cluster_label_start", c_int64),
    ]


class StemsegConvShape(ctypes.Structure):
    _fields_ = [("n", c_int32), ("t", c_int32), ("h", c_int32), ("w", c_int32), ("cin", c_int32), ("cout", c_int32),
                ("kernel_size", c_int32), ("planes", c_int32)]


# name -> (restype, argtypes); every symbol include/stemseg_b200.h declares (tests check the two agree)
PROTOTYPES = {
    "stemseg_last_error": (ctypes.c_char_p, []),
    "stemseg_abi_version": (c_int32, []),
    "stemseg_check_device": (c_int32, []),
    "stemseg_seq_cluster_meta_words": (c_size_t, [c_int32, c_int32]),
    "stemseg_seq_cluster_workspace_bytes": (c_int32, [ctypes.POINTER(StemsegClusterParams),
                                                      ctypes.POINTER(c_size_t)]),
    "stemseg_seq_cluster": (c_int32, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(StemsegClusterParams), c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_prob_threshold_to_distance": (c_float, [c_double]),
    "stemseg_fg_compact_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "stemseg_fg_compact": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_fg_gather": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "stemseg_pack_activation": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                          c_void_p, c_int32, c_void_p]),
    "stemseg_pack_conv_weight": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                           c_int32, c_int32, c_void_p]),
    "stemseg_conv3d_bf16_planes": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p,
                                             ctypes.POINTER(StemsegConvShape), c_int32, c_void_p]),
    "stemseg_group_norm_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "stemseg_group_norm_stats": (c_int32, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_float, c_void_p, c_void_p,
                                           c_size_t, c_void_p]),
    "stemseg_norm_relu_pool": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "stemseg_upsample_add": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p, c_int32, c_void_p]),
    "stemseg_head_output": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p]),
}


class StemsegError(RuntimeError):
    pass


_lib = None


def load():
    """Load libstemseg_b200.so (built in-tree by ``python -m stemseg_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "stemseg_b200: %s is missing -- build it with `python -m stemseg_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.stemseg_abi_version() != ABI_VERSION:
        raise ImportError("stemseg_b200: ABI version mismatch (library %d, binding %d) -- rebuild" % (
            lib.stemseg_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().stemseg_last_error()
        raise StemsegError("stemseg_b200 call failed (%d): %s" % (rc, msg.decode() if msg else ""))


def ptr(t):
    """Device/host pointer of a torch tensor as c_void_p (None -> NULL)."""
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
