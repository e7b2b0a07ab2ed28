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
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.ReLU(), torch.nn.Linear(7, 3),
                              torch.nn.Linear(3, 2))
    with torch.no_grad():
        net[1].running_mean.fill_(float(rank + 1))      # buffers differ per rank too
    sync = FlatGradAllReduce(net, world, buckets=2)
    sync.broadcast_parameters()
    w0 = torch.cat([p.detach().reshape(-1) for p in net.parameters()] + [net[1].running_mean.reshape(-1)])
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    live = [p for p in net.parameters()][:6]            # the last Linear never runs: its grads stay None

    def local_step(seed):
        """-> this rank's own gradients of the step (taken with autograd.grad, which fires no accumulate hooks: once the
        hooks are armed, a bucket's all-reduce may already be rewriting .grad in place when backward() returns)."""
        torch.manual_seed(seed + rank)
        x = torch.randn(4, 5)
        own = torch.autograd.grad(net[3](net[2](net[1](net[0](x)))).sum(), live)
        net[1].num_batches_tracked -= 1
        net[3](net[2](net[1](net[0](x)))).sum().backward()
        return [g.clone() for g in own]

    def expected(local):
        mine = torch.cat([g.reshape(-1) for g in local])
        allg = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allg, mine)
        return sum(allg) / world

    res = {}
    # step 1 (discovery, un-overlapped), grads start as None
    want = expected(local_step(100))
    sync.allreduce()
    got = torch.cat([p.grad.reshape(-1) for p in live])
    res["avg_ok_1"] = torch.allclose(got, want, atol=1e-7)
    res["untouched_none"] = net[4].weight.grad is None and net[4].bias.grad is None
    res["payload"] = sync.last_payload_bytes == 4 * sum(p.numel() for p in live)
    res["grads_are_views"] = all(p.grad.data_ptr() == sync.views[i].data_ptr() for i, p in enumerate(live))
    # step 2 (hooks launch the buckets during backward), grads reset to None
    for p in net.parameters():
        p.grad = None
    want = expected(local_step(200))
    sync.allreduce()
    res["avg_ok_2"] = torch.allclose(torch.cat([p.grad.reshape(-1) for p in live]), want, atol=1e-7)
    # step 3: zero_grad(set_to_none=False) - autograd accumulates IN PLACE into the flat views (this used to
    # produce all-zero gradients)
    torch.optim.SGD(net.parameters(), lr=0.1).zero_grad(set_to_none=False)
    res["views_zeroed"] = bool((sync.flat == 0).all())
    want = expected(local_step(300))
    sync.allreduce()
    got3 = torch.cat([p.grad.reshape(-1) for p in live])
    res["avg_ok_3"] = torch.allclose(got3, want, atol=1e-7) and bool(got3.abs().sum() > 0)
    # step 4: no zeroing at all (gradient accumulation): result = previous average + average of the new gradients
    prev = got3.clone()
    want_new = expected(local_step(400))
    sync.allreduce()
    res["accumulate_ok"] = torch.allclose(torch.cat([p.grad.reshape(-1) for p in live]), prev + want_new, atol=1e-6)
    if rank == 0:
        res["same_init"] = all(torch.equal(g, gathered[0]) for g in gathered)
        out.update(res)
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    bad = [k for k, v in out.items() if not v]
    assert not bad and len(out) == 9, dict(out)
