"""ctypes binding of include/stemseg_b200.h.  Fails loudly: no library -> ImportError, no fallback of any kind."""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libstemseg_b200.so")

STEMSEG_MAX_EMBEDDING_DIMS = 16
STEMSEG_MAX_INSTANCES = 64
STEMSEG_MAX_LOSS_INSTANCES = 32
ABI_VERSION = 22

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
                ("kernel_size", c_int32), ("planes", c_int32), ("split_k", c_int32), ("tiles_per_cta", c_int32),
                ("out_bf16", c_int32)]


# name -> (restype, argtypes); every symbol include/stemseg_b200.h declares (tests check the two agree)
PROTOTYPES = {
    "stemseg_last_error": (ctypes.c_char_p, []),
    "stemseg_abi_version": (c_int32, []),
    "stemseg_check_device": (c_int32, []),
    "stemseg_seq_cluster_meta_words": (c_size_t, [c_int32, c_int32]),
    "stemseg_seq_cluster_workspace_bytes": (c_int32, [ctypes.POINTER(StemsegClusterParams),
                                                      ctypes.POINTER(c_size_t)]),
    "stemseg_seq_cluster": (c_int32, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(StemsegClusterParams), c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_prob_threshold_to_distance": (c_float, [c_double]),
    "stemseg_fg_compact_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "stemseg_fg_compact": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_fg_compact_threshold": (c_int32, [c_void_p, c_float, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                               c_size_t, c_void_p]),
    "stemseg_fg_compact_mean_threshold": (c_int32, [c_void_p, c_void_p, c_float, c_int64, c_int32, c_int32, c_int32,
                                                    c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_frame_accumulate": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p]),
    "stemseg_fg_gather_upsampled": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64,
                                              c_void_p, c_int32, c_void_p, c_void_p]),
    "stemseg_fg_gather": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_int32, c_void_p,
                                    c_void_p]),
    "stemseg_label_pair_histogram": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32,
                                               c_void_p, c_void_p, c_void_p]),
    "stemseg_relabel_lut": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p]),
    "stemseg_stitch_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "stemseg_stitch_subclip": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                         c_int32, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_rank_map_scatter": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "stemseg_mask_writeback": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                         c_int32, c_void_p, c_void_p]),
    "stemseg_pack_activation": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                          c_void_p, c_int32, c_void_p]),
    "stemseg_pack_conv_weight": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                           c_int32, c_int32, c_void_p]),
    "stemseg_conv3d_bf16_planes": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             ctypes.POINTER(StemsegConvShape), c_int32, c_void_p]),
    "stemseg_conv3d_tiles_per_sample": (c_int32, [ctypes.POINTER(StemsegConvShape)]),
    "stemseg_group_norm_finalize": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int64, c_int32, c_int32, c_float,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "stemseg_pack_conv_weight_dgrad": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                                 c_int32, c_void_p]),
    "stemseg_head_backward_workspace_bytes": (c_size_t, [c_int32]),
    "stemseg_head_backward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "stemseg_upsample_add_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                           c_void_p]),
    "stemseg_head_output_x": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p]),
    "stemseg_head_backward_x_workspace_bytes": (c_size_t, [c_int32]),
    "stemseg_head_backward_x": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                          c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                          c_void_p]),
    "stemseg_upsample_transpose": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                             c_void_p]),
    "stemseg_pool_relu_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                             c_int32, c_void_p, c_void_p]),
    "stemseg_group_norm_backward_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "stemseg_group_norm_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32,
                                              c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_group_norm_backward_planes_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "stemseg_group_norm_backward_planes": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32,
                                                     c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                                     c_size_t, c_void_p]),
    "stemseg_channel_sum_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "stemseg_channel_sum": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_to_planes": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p]),
    "stemseg_transposed_row_length": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "stemseg_transpose_pad": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_int32, c_void_p]),
    "stemseg_wgrad_k_splits": (c_int32, [c_int32] * 7),
    "stemseg_conv3d_wgrad": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, c_int32, c_void_p, c_void_p]),
    "stemseg_wgrad_direct_k_splits": (c_int32, [c_int32] * 6),
    "stemseg_conv3d_wgrad_direct": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                              c_int32, c_int32, c_void_p, c_void_p]),
    "stemseg_wgrad_reduce": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32,
                                       c_int32, c_void_p]),
    "stemseg_embedding_loss_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "stemseg_embedding_loss": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                         ctypes.POINTER(c_float), c_float, c_float, c_float, c_float, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_semseg_loss_workspace_bytes": (c_size_t, []),
    "stemseg_semseg_loss": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                      c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_scale_by_device_scalar": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "stemseg_sgd_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int32,
                                   c_void_p]),
    "stemseg_sgd_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p]),
    "stemseg_conv3d_auto_split": (c_int32, [ctypes.POINTER(StemsegConvShape)]),
    "stemseg_group_norm_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "stemseg_group_norm_stats": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int64, c_int32, c_int32, c_float,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemseg_norm_relu_pool": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "stemseg_norm_relu_pool_bf16in": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "stemseg_upsample_add": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p, c_int32, c_void_p]),
    "stemseg_head_lowres": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "stemseg_conv1x1_head_output": (c_int32, [c_void_p, c_void_p, ctypes.POINTER(StemsegConvShape), c_void_p, c_int32,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_void_p,
                                              c_int32, c_void_p]),
    "stemseg_head_output": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p]),
}


class StemsegError(RuntimeError):
    pass


_lib = None
# number of CUDA kernels launched through this binding (each wrapper adds the kernels its C call enqueues);
# bench.py reports it as `gpu_launches`
KERNEL_LAUNCHES = [0]
KERNELS_PER_CALL = {
    "stemseg_seq_cluster": 1, "stemseg_fg_compact": 3, "stemseg_fg_compact_threshold": 3, "stemseg_fg_gather": 1, "stemseg_fg_compact_mean_threshold": 3, "stemseg_frame_accumulate": 1,
    "stemseg_fg_gather_upsampled": 1,
    "stemseg_pack_activation": 1, "stemseg_pack_conv_weight": 1, "stemseg_conv3d_bf16_planes": 1,
    "stemseg_group_norm_stats": 2, "stemseg_group_norm_finalize": 1, "stemseg_norm_relu_pool": 1, "stemseg_norm_relu_pool_bf16in": 1, "stemseg_upsample_add": 1, "stemseg_head_output": 1,
    "stemseg_head_lowres": 1, "stemseg_conv1x1_head_output": 1,
    "stemseg_label_pair_histogram": 1, "stemseg_relabel_lut": 1, "stemseg_stitch_subclip": 3,
    "stemseg_rank_map_scatter": 1, "stemseg_mask_writeback": 1,
    "stemseg_pack_conv_weight_dgrad": 1, "stemseg_head_backward": 3, "stemseg_upsample_transpose": 1,
    "stemseg_pool_relu_backward": 1, "stemseg_group_norm_backward": 3, "stemseg_channel_sum": 2,
    "stemseg_to_planes": 1, "stemseg_transpose_pad": 1, "stemseg_conv3d_wgrad": 1, "stemseg_wgrad_reduce": 1,
    "stemseg_conv3d_wgrad_direct": 1,
    "stemseg_scale_by_device_scalar": 1, "stemseg_sgd_step": 1, "stemseg_sgd_step_dev": 1,
    "stemseg_group_norm_backward_planes": 4, "stemseg_semseg_loss": 2, "stemseg_upsample_add_f32": 1, "stemseg_head_output_x": 1, "stemseg_head_backward_x": 2,
    # stemseg_embedding_loss launches a shape-dependent number of kernels: counted by losses.py
}


class _Counted(object):
    """Callable wrapper around one C entry point that keeps KERNEL_LAUNCHES up to date."""

    def __init__(self, fn, kernels):
        self._fn, self._kernels = fn, kernels

    def __call__(self, *args):
        KERNEL_LAUNCHES[0] += self._kernels
        return self._fn(*args)


class _Library(object):
    def __init__(self, cdll):
        self._cdll = cdll
        for name in PROTOTYPES:
            fn = getattr(cdll, name)
            k = KERNELS_PER_CALL.get(name, 0)
            setattr(self, name, _Counted(fn, k) if k else fn)


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
    _lib = _Library(lib)
    return _lib


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


class capture_guard(object):
    """Keeps Python's cyclic garbage collector out of a CUDA-graph capture.

    Objects that own CUDA graphs (an older pipeline / trainer that became garbage) are destroyed whenever the cyclic
    collector happens to run; destroying a graph while ANOTHER stream capture is in progress invalidates that capture
    ("operation not permitted when stream is capturing").  torch.cuda.graph no longer forces a collection on entry, so
    collect up front and hold the collector off until the capture is over."""

    def __enter__(self):
        import gc
        gc.collect()
        self._was_enabled = gc.isenabled()
        gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self._was_enabled:
            gc.enable()
        return False


class LRUCache(object):
    """Bounded cache of captured CUDA graphs: every entry owns a private memory pool (saved activations, gradient
    buffers), so an unbounded dict keyed on input shapes grows until the GPU is full."""

    def __init__(self, capacity):
        from collections import OrderedDict
        self.capacity = max(1, int(capacity))
        self._d = OrderedDict()

    def get(self, key):
        if key in self._d:
            self._d.move_to_end(key)
            return self._d[key]
        return None

    def put(self, key, value):
        self._d[key] = value
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)              # dropping the entry releases its graphs and their pool

    def clear(self):
        self._d.clear()

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        return key in self._d
