"""CPU, gloo, world_size 2: the flat-gradient allreduce used for the N > 1 path (one process per GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gapartnet_b200.ddp import FlatGradArena, shard_scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 3))
    pre = torch.zeros(sum(p.numel() for p in net[0].parameters()))
    off = 0
    for p in net[0].parameters():          # pretend the first layer's grads already live in an engine arena
        p.grad = pre[off:off + p.numel()].view_as(p)
        off += p.numel()
    arena = FlatGradArena(net.parameters(), existing=[pre])
    assert len(arena.arenas) == 2 and arena.nbytes() == 4 * sum(p.numel() for p in net.parameters())
    arena.zero_()
    x = torch.randn(32, 8, generator=torch.Generator().manual_seed(100 + rank))
    net(x).square().mean().backward()       # autograd accumulates into the flat views
    local = [p.grad.clone() for p in net.parameters()]
    arena.allreduce_mean()
    out[rank] = (local, [p.grad.clone() for p in net.parameters()])
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (l0, r0), (l1, r1) = out[0], out[1]
    for a, b, ra, rb in zip(l0, l1, r0, r1):
        torch.testing.assert_close(ra, (a + b) / 2)
        torch.testing.assert_close(rb, (a + b) / 2)


def test_shard_scenes_partitions():
    for n in (16, 17, 3):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in shard_scenes(n, r, w)]
            assert got == list(range(n))
