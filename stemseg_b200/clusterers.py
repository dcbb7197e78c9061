"""Drop-in ``SequentialClustering`` running as one persistent CUDA kernel on a B200.

Mirrors stemseg/inference/clusterers.py (class names, constructor and call signatures, returned dict, timing-log
attributes, assertion behaviour) so that ``stemseg/inference/main.py:84-91`` / ``online_chainer.py:283-285`` work
unchanged when this class is substituted.  All per-point arithmetic happens in
``stemseg_seq_cluster`` (csrc/cluster.cu); the host only converts thresholds and formats the metadata lists.
"""
from collections import defaultdict
from time import time as current_time

import torch

from stemseg_b200 import _lib


class ClustererBase(object):
    """Same interface as the reference's ClustererBase (clusterers.py:7-31)."""

    def __init__(self):
        self._time_log = defaultdict(list)

    def __call__(self, embeddings, *args, **kwargs):
        assert embeddings.dtype == torch.float32                                    # clusterers.py:12

        start_time = current_time()
        output = self._process(embeddings, *args, **kwargs)
        duration = current_time() - start_time
        self._time_log[embeddings.shape[0]].append(duration)
        return output

    def _process(self, embeddings, *args, **kwargs):
        raise NotImplementedError("Must be implemented by derived class")

    def reset_time_log(self):
        self._time_log = defaultdict(list)

    @property
    def average_time(self):
        all_times = sum(list(self._time_log.values()), [])
        return sum(all_times) / float(len(all_times))

    name = property(fget=lambda self: self._name)


class SequentialClustering(ClustererBase):
    """clusterers.py:34-175 with the loop body on the device."""

    def __init__(self, primary_prob_thresh, secondary_prob_thresh, min_seediness_prob,
                 n_free_dims, free_dim_stds, device, max_instances=20):
        super().__init__()
        self.thresholding_mode = "probability"
        self.primary_prob_thresh = primary_prob_thresh
        self.secondary_prob_thresh = secondary_prob_thresh
        self.min_seediness_prob = min_seediness_prob
        self.max_instances = max_instances
        self.n_free_dims = n_free_dims
        self.free_dim_stds = free_dim_stds
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("stemseg_b200.SequentialClustering runs on a B200 only (device=%r); there is no CPU "
                             "path -- use the reference implementation for CPU clustering" % (device,))
        if not 0 <= max_instances <= _lib.STEMSEG_MAX_INSTANCES:
            raise ValueError("max_instances must be in [0, %d]" % _lib.STEMSEG_MAX_INSTANCES)
        self._lib = _lib.load()
        self._last_primary = None     # int32 [N]: ordinal of the claiming cluster in the primary pass, -1 if none

    @staticmethod
    def distances_to_prob(distances):
        return (-0.5 * distances).exp()

    def prob_to_distance(self, prob):
        """Largest fp32 distance d with exp(-0.5 d) > prob (clusterers.py:53-54 inverted; see the C header)."""
        return float(self._lib.stemseg_prob_threshold_to_distance(float(prob)))

    @torch.no_grad()
    def _process(self, embeddings, bandwidths, seediness, cluster_label_start=1, *args, **kwargs):
        if embeddings.numel() == 0:                                                  # clusterers.py:62-69
            return torch.zeros(0, dtype=torch.long, device=embeddings.device), {
                'instance_labels': [], 'instance_centers': [], 'instance_stds': [], 'instance_masks': []}

        input_device = embeddings.device
        embeddings = embeddings.to(device=self.device).contiguous()

        assert torch.is_tensor(bandwidths)
        if bandwidths.shape[0] != embeddings.shape[0]:                               # clusterers.py:75-76
            bandwidths = bandwidths.expand_as(embeddings)
        bandwidths = bandwidths.to(device=self.device, dtype=torch.float32).contiguous()

        n, e = embeddings.shape
        if self.n_free_dims == 0:
            assert embeddings.shape == bandwidths.shape                              # clusterers.py:80-81

        assert torch.is_tensor(seediness)
        seediness = seediness.reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()  # [N,1] -> [N]
        assert seediness.shape[0] == n

        return_label_masks = kwargs.get("return_label_masks", False)

        pending = self.launch(embeddings, bandwidths, seediness, cluster_label_start)
        labels, meta = self.finish(pending, return_label_masks)
        return labels.to(input_device), meta

    @torch.no_grad()
    def launch(self, embeddings, bandwidths, seediness, cluster_label_start=1, n_points_dev=None):
        """Enqueue the clustering kernel without synchronising (capture-safe).  All inputs are contiguous fp32 CUDA
        tensors of a fixed *capacity* N; ``n_points_dev`` (device int32 [1]) optionally holds the actual count."""
        n, e = embeddings.shape
        if e > _lib.STEMSEG_MAX_EMBEDDING_DIMS:
            raise ValueError("embedding size %d > %d unsupported" % (e, _lib.STEMSEG_MAX_EMBEDDING_DIMS))
        if bandwidths.shape[1] + self.n_free_dims != e:
            raise ValueError("bandwidths has %d columns, expected %d (E=%d, n_free_dims=%d)" % (
                bandwidths.shape[1], e - self.n_free_dims, e, self.n_free_dims))
        params = _lib.StemsegClusterParams()
        params.n_points = n
        params.embedding_dims = e
        params.n_free_dims = self.n_free_dims
        if self.n_free_dims > 0:                                                     # clusterers.py:100-102
            free_dim_stds = torch.tensor(self.free_dim_stds).to(dtype=torch.float32)
            free_bw = (1. / (free_dim_stds ** 2)).tolist()
            assert len(free_bw) == self.n_free_dims
            for k, val in enumerate(free_bw):
                params.free_dim_bandwidths[k] = val
        params.d_primary = self.prob_to_distance(self.primary_prob_thresh)
        params.d_secondary = self.prob_to_distance(self.secondary_prob_thresh)
        params.min_seediness_prob = float(self.min_seediness_prob)
        params.max_instances = self.max_instances
        params.cluster_label_start = int(cluster_label_start)

        lib = self._lib
        ws_bytes = _lib.c_size_t(0)
        _lib.check(lib.stemseg_seq_cluster_workspace_bytes(params, ws_bytes))
        meta_words = lib.stemseg_seq_cluster_meta_words(e, self.max_instances)
        with torch.cuda.device(self.device):
            labels = torch.empty(n, dtype=torch.int64, device=self.device)
            primary = torch.empty(n, dtype=torch.int32, device=self.device)
            meta = torch.empty(meta_words, dtype=torch.int32, device=self.device)
            workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=self.device)
            _lib.check(lib.stemseg_seq_cluster(
                _lib.ptr(embeddings), _lib.ptr(bandwidths), _lib.ptr(seediness), params, _lib.ptr(n_points_dev),
                _lib.ptr(labels), _lib.ptr(primary), _lib.ptr(meta), _lib.ptr(workspace), ws_bytes.value,
                _lib.stream_ptr()))
        return {"labels": labels, "primary": primary, "meta": meta, "e": e, "label_start": int(cluster_label_start),
                "workspace": workspace}

    @torch.no_grad()
    def finish(self, pending, return_label_masks=False, meta_host=None):
        """Fetch the metadata of a launched clustering (the one device->host sync) and format the reference's dict.
        ``meta_host``: the metadata words already copied to the host by the caller (asynchronous pipelines)."""
        labels, primary, e = pending["labels"], pending["primary"], pending["e"]
        cluster_label_start = pending["label_start"]
        if meta_host is None:
            meta_host = pending["meta"].cpu()
        n_done = int(meta_host[2])
        labels, primary = labels[:n_done], primary[:n_done]
        self._last_primary = primary
        k = int(meta_host[0])
        mi = self.max_instances
        floats = meta_host[4 + mi:].view(torch.float32)
        centers = floats[:mi * e].reshape(mi, e)[:k]
        bws = floats[mi * e:2 * mi * e].reshape(mi, e)[:k]
        unique_labels = [i + cluster_label_start for i in range(k)]                 # clusterers.py:121-123
        label_centers = centers.tolist()                                            # clusterers.py:124
        label_stds = (1. / bws).clamp(min=1e-8).sqrt().tolist()                     # clusterers.py:125 (one op for all K)
        label_masks = []
        if return_label_masks:                                                      # clusterers.py:145-146
            label_masks = [(primary == i).cpu() for i in range(k)]

        return labels, {
            'instance_labels': unique_labels,
            'instance_centers': label_centers,
            'instance_stds': label_stds,
            'instance_masks': label_masks
        }
