"""world_size-2 gloo test (CPU) of the clip-parallel exchange + stitch against the reference goldens."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chain_cases import CASES, make_video


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_chaining_cpu import oracle_local_labels
    from stemseg_b200.parallel import exchange_and_stitch, shard_subclips
    masks, subseqs = make_video(**CASES[name])
    owned = shard_subclips(len(subseqs), rank, world)
    frames_all = [list(s["frames"]) for s in subseqs]
    # each rank clusters ONLY its own sub-clips (CPU oracle stands in for the CUDA pipeline in this host-logic test)
    f, l, m = oracle_local_labels(masks, [subseqs[i] for i in owned])
    local = {i: (l[k], m[k]) for k, i in enumerate(owned)}
    container, subseq_labels, metas = exchange_and_stitch(masks.shape[0], frames_all, local)
    track_labels, pt_counts, lifetimes = container.get_track_mask_idxes()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank),
             **{"track/%d" % t: lab.numpy() for t, lab in enumerate(track_labels)},
             ids=np.array(sorted(pt_counts.keys())), counts=np.array([pt_counts[i] for i in sorted(pt_counts.keys())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["three_blobs", "five_blobs_tail"])
def test_clip_parallel_two_ranks(name, golden_dir, tmp_path):
    golden = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        t = 0
        while "%s/track/%d" % (name, t) in golden:
            np.testing.assert_array_equal(got["track/%d" % t].astype(np.int32), golden["%s/track/%d" % (name, t)])
            t += 1
        assert got["ids"].tolist() == golden[name + "/ids"].tolist()
        assert got["counts"].tolist() == golden[name + "/pt_counts"].tolist()


def _layout_worker(rank, world, port, name, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_chaining_cpu import oracle_local_labels
    from stemseg_b200.chaining import stitch_subsequences
    from stemseg_b200.parallel import ExchangeLayout, shard_subclips
    masks, subseqs = make_video(**CASES[name])
    n_sub = len(subseqs)
    owned = shard_subclips(n_sub, rank, world)
    frames_all = [list(s["frames"]) for s in subseqs]
    max_t = max(len(f) for f in frames_all)
    cap = masks.shape[1] * masks.shape[2]
    mi, e = 20, 4
    meta_words = 4 + mi * (1 + 2 * e)                       # stemseg_seq_cluster_meta_words(e, max_instances)
    layout = ExchangeLayout(n_sub, world, max_t, cap, meta_words)
    buf = layout.alloc("cpu")
    f, l, m = oracle_local_labels(masks, [subseqs[i] for i in owned])
    for slot, i in enumerate(owned):
        labels = torch.cat(l[slot])
        counts = torch.tensor([x.numel() for x in l[slot]], dtype=torch.int32)
        meta = torch.zeros(meta_words, dtype=torch.int32)
        meta[0] = len(m[slot]["instance_labels"])
        layout.write(buf, slot, labels, counts, meta)
    all_labels, all_head = layout.views(layout.gather(buf))
    assert tuple(all_labels.shape) == (world, layout.max_local, max_t * cap)
    local_labels, ks = [], []
    for i in range(n_sub):
        r, slot = layout.slot_of(i)
        cnt = all_head[r, slot, :len(frames_all[i])].tolist()
        ks.append(int(all_head[r, slot, max_t]))
        assert int(all_head[r, slot, max_t + 1]) == ks[-1]            # meta word 0 repeats K
        flat = all_labels[r, slot, :sum(cnt)]
        local_labels.append(list(flat.split(cnt)))
    metas = [{"instance_labels": list(range(1, k + 1))} for k in ks]
    container, _, _ = stitch_subsequences(masks.shape[0], frames_all, local_labels, metas)
    track_labels, pt_counts, _ = container.get_track_mask_idxes()
    np.savez(os.path.join(out_dir, "layout_rank%d.npz" % rank),
             **{"track/%d" % t: lab.numpy() for t, lab in enumerate(track_labels)},
             ids=np.array(sorted(pt_counts.keys())))
    dist.barrier()
    dist.destroy_process_group()


def test_fixed_layout_exchange_two_ranks(golden_dir, tmp_path):
    """The GPU path's exchange (ONE all_gather of a fixed-layout int64 buffer: labels + int32 header words) on gloo:
    both ranks rebuild every sub-clip from the gathered buffer and stitch to the reference's track ids."""
    name = "five_blobs_tail"
    golden = np.load(os.path.join(golden_dir, "chain_golden.npz"))
    port = _free_port()
    mp.spawn(_layout_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "layout_rank%d.npz" % rank))
        t = 0
        while "%s/track/%d" % (name, t) in golden:
            np.testing.assert_array_equal(got["track/%d" % t].astype(np.int32), golden["%s/track/%d" % (name, t)])
            t += 1
        assert t > 0 and got["ids"].tolist() == golden[name + "/ids"].tolist()


def test_sharding():
    from stemseg_b200.parallel import shard_subclips
    assert shard_subclips(8, 3, 8) == [3]
    assert shard_subclips(5, 1, 2) == [1, 3]
    assert sorted(sum([shard_subclips(11, r, 4) for r in range(4)], [])) == list(range(11))
