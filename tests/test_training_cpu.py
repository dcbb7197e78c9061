"""world_size-2 gloo test (CPU) of the training-side host logic: flat parameter buffers + per-module asynchronous
gradient all-reduce (stemseg_b200.training).  The CUDA kernels are not involved (the fused SGD kernel is tested on
the GPU against torch.optim.SGD)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _modules():
    torch.manual_seed(5)
    a = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    b = torch.nn.Linear(6, 3, bias=False)
    return a, b


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stemseg_b200.training import FlatParameters, GradientExchange
    a, b = _modules()
    flats = [FlatParameters(a), FlatParameters(b)]
    ex = GradientExchange(flats)
    for step in range(2):
        for f in flats:
            f.zero_grad()
        x = torch.full((4, 6), float(rank + 1 + step))
        (a(x).sum() + 2.0 * b(x).sum()).backward()
        ex.finish()
    torch.save({"ga": flats[0].grad.clone(), "gb": flats[1].grad.clone(),
                "views": all(p.grad.data_ptr() == f.grad.data_ptr() + 4 * o for f in flats
                             for p, o in zip(f.params, f.offsets))}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(2)]
    # expected: the SUM over ranks of the last step's local gradients, computed without any flat buffers
    want_a, want_b = None, None
    for rank in range(2):
        a, b = _modules()
        x = torch.full((4, 6), float(rank + 1 + 1))
        (a(x).sum() + 2.0 * b(x).sum()).backward()
        ga = torch.cat([p.grad.reshape(-1) for p in a.parameters()])
        gb = torch.cat([p.grad.reshape(-1) for p in b.parameters()])
        want_a = ga if want_a is None else want_a + ga
        want_b = gb if want_b is None else want_b + gb
    for r in range(2):
        assert got[r]["views"]
        # flat buffers pad every parameter to a multiple of 4 elements: compare the packed values
        from stemseg_b200.training import FlatParameters
        a, b = _modules()
        fa, fb = FlatParameters(a), FlatParameters(b)
        pa = torch.cat([got[r]["ga"][o:o + p.numel()] for p, o in zip(fa.params, fa.offsets)])
        pb = torch.cat([got[r]["gb"][o:o + p.numel()] for p, o in zip(fb.params, fb.offsets)])
        torch.testing.assert_close(pa, want_a)
        torch.testing.assert_close(pb, want_b)
    assert torch.equal(got[0]["ga"], got[1]["ga"]) and torch.equal(got[0]["gb"], got[1]["gb"])


def test_flat_parameters_preserve_values_and_alignment():
    from stemseg_b200.training import FlatParameters
    a, _ = _modules()
    before = [p.detach().clone() for p in a.parameters()]
    flat = FlatParameters(a)
    for p, q, off in zip(a.parameters(), before, flat.offsets):
        assert torch.equal(p.detach(), q)
        assert off % 4 == 0 and p.data_ptr() == flat.data.data_ptr() + 4 * off
    out = a(torch.ones(2, 6))
    out.sum().backward()
    assert float(flat.grad.abs().sum()) > 0


def test_fused_sgd_has_no_cpu_path():
    from stemseg_b200.training import FlatParameters, sgd_step
    a, _ = _modules()
    with pytest.raises(ValueError):
        sgd_step(FlatParameters(a), 1e-3, 0.9, 1e-4, True)
