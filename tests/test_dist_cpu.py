"""CPU, gloo, world_size 2: the flat-gradient data-parallel exchange (dist.FlatGradients) — reduced gradients equal the
mean of the per-rank gradients, unused parameters keep a zero gradient, parameters are broadcast from rank 0."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shapeclipper_b200 import dist as scdist
    scdist.setup(backend="gloo")
    torch.manual_seed(rank)                                   # different initial parameters per rank on purpose
    a = torch.nn.Parameter(torch.randn(5, 3)); b = torch.nn.Parameter(torch.randn(7)); unused = torch.nn.Parameter(torch.randn(4))
    flat = scdist.FlatGradients([a, b, unused], extra=11, device="cpu")
    flat.broadcast_parameters()
    x = torch.full((3,), float(rank + 1))
    flat.zero()
    ((a @ x).sum() * (rank + 1) + (b * b).sum()).backward()
    local = [a.grad.clone(), b.grad.clone()]
    flat.all_reduce()
    out[rank] = dict(a=a.detach().clone(), ga=a.grad.clone(), gb=b.grad.clone(), gu=unused.grad.clone(), local=local,
                     n=flat.flat.numel())
    scdist.cleanup()


def test_flat_gradient_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert torch.equal(r0["a"], r1["a"])                                        # broadcast from rank 0
    assert torch.allclose(r0["ga"], (r0["local"][0] + r1["local"][0]) / 2)      # mean of per-rank gradients
    assert torch.allclose(r0["gb"], (r0["local"][1] + r1["local"][1]) / 2)
    assert torch.equal(r0["ga"], r1["ga"]) and torch.equal(r0["gb"], r1["gb"])
    assert (r0["gu"] == 0).all() and r0["n"] == 15 + 7 + 4 + 11
