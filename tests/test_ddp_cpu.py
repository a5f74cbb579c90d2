"""CPU, world_size 2, gloo: the N>1 host logic — gradient SUM-reduction in flat buckets launched from
autograd hooks, one-time parameter broadcast, and the rp_iter exchange that makes rp_mask a global-batch
statistic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from passion_b200.ddp import GradReducer
    torch.manual_seed(100 + rank)                       # different init per rank: broadcast must equalise
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 8), torch.nn.Tanh(),
                                torch.nn.Linear(8, 3))
    red = GradReducer(model, bucket_mb=1e-4)            # ~26 floats per bucket -> several buckets
    red.broadcast_parameters()
    p0 = torch.cat([p.detach().flatten() for p in model.parameters()])
    torch.manual_seed(7 + rank)
    x = torch.randn(5, 6)
    for it in range(2):                                 # twice: buckets must be reusable
        model.zero_grad(set_to_none=True)
        red.prepare()
        loss = model(x).pow(2).sum()                    # SUM over the local samples
        loss.backward()
        red.finish()
    g = torch.cat([p.grad.flatten() for p in model.parameters()])
    # local (unreduced) gradient for the check
    red.enabled = False
    model.zero_grad(set_to_none=True)
    model(x).pow(2).sum().backward()
    g_local = torch.cat([p.grad.flatten() for p in model.parameters()])
    rp = red.allreduce_small(torch.tensor([1.0, -2.0, 0.5, 0.0]) * (rank + 1))
    q.put((rank, p0.tolist(), g.tolist(), g_local.tolist(), rp.tolist(), len(red.buckets)))   # by value
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0a, ga, la, rpa, nb), (_, p0b, gb, lb, rpb, _) = [tuple(torch.tensor(v) if isinstance(v, list) else v for v in r) for r in res]
    assert nb > 1
    assert torch.equal(p0a, p0b)                                   # broadcast once
    assert torch.allclose(ga, gb) and torch.allclose(ga, la + lb, atol=1e-6)   # SUM, not mean
    assert torch.allclose(rpa, torch.tensor([3.0, -6.0, 1.5, 0.0])) and torch.equal(rpa, rpb)
