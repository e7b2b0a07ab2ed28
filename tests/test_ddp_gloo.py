"""world_size-2 `gloo` test (CPU) of the multi-GPU row: graph sharding + the single flat gradient all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from signnet_basisnet_b200.ddp import FlatGradAllReduce, shard_batch, shard_graphs
from signnet_basisnet_b200.synth import synth_batch


def test_shard_graphs_partition():
    for B in (1, 7, 128, 1024):
        for W in (1, 2, 3, 8):
            parts = [shard_graphs(B, W, r) for r in range(W)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_rebases_and_covers():
    d = synth_batch(11, "zinc", seed=3)
    d.y = torch.arange(11.0).unsqueeze(1)
    shards = [shard_batch(d, 3, r) for r in range(3)]
    assert sum(s.num_graphs for s in shards) == 11
    assert sum(s.batch.numel() for s in shards) == d.batch.numel()
    assert sum(s.edge_index.shape[1] for s in shards) == d.edge_index.shape[1]
    assert sum(s.eigen_vectors.numel() for s in shards) == d.eigen_vectors.numel()
    for s in shards:
        assert int(s.batch.min()) == 0 and int(s.batch.max()) == s.num_graphs - 1
        assert int(s.edge_index.min()) >= 0 and int(s.edge_index.max()) < s.batch.numel()
        assert (s.batch[s.edge_index[0]] == s.batch[s.edge_index[1]]).all()
    assert torch.equal(torch.cat([s.y for s in shards]), d.y)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # different initial weights per rank: broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    sync = FlatGradAllReduce(net, world)
    sync.broadcast_parameters()
    w0 = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    torch.manual_seed(100 + rank)
    x = torch.randn(4, 5)
    net[2](net[1](net[0](x))).sum().backward()  # the last Linear never runs: its grads stay None -> zero slice
    local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    flat = sync.allreduce().clone()
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    pieces = [torch.zeros(p.numel()) if g is None else g.reshape(-1) for p, g in zip(net.parameters(), local)]
    mine = torch.cat(pieces)
    allg = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allg, mine)
    if rank == 0:
        out["same_init"] = all(torch.equal(g, gathered[0]) for g in gathered)
        out["avg_ok"] = torch.allclose(flat, sum(allg) / world, atol=1e-7)
        out["grads_are_views"] = all(p.grad is not None for p in net.parameters())
        out["untouched_zero"] = bool((net[3].weight.grad == 0).all())
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["same_init"] and out["avg_ok"] and out["grads_are_views"] and out["untouched_zero"]
