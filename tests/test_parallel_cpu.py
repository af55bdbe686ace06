"""CPU (gloo, world_size 2) tests of the data-parallel host logic: bucketed gradient all-reduce with post-accumulate
hooks, parameter broadcast, row sharding.  The GPU path uses the same code with the nccl backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hint_b200.parallel import BucketedGradAllReduce, broadcast_parameters, shard_rows
    torch.manual_seed(100 + rank)                       # different init per rank -> broadcast must equalise
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    broadcast_parameters(net)
    ref = [p.detach().clone() for p in net.parameters()]
    red = BucketedGradAllReduce(net)
    torch.manual_seed(7)
    X = torch.randn(12, 5)                               # the GLOBAL batch; each rank takes its row shard
    lo, hi = shard_rows(12, rank, world)
    loss = net(X[lo:hi]).pow(2).sum() / (hi - lo)        # local mean
    loss.backward()
    red.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    # single-process reference on the global batch: mean over all rows == average of the per-rank means (equal shards)
    net2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    with torch.no_grad():
        for p2, r in zip(net2.parameters(), ref):
            p2.copy_(r)
    (net2(X).pow(2).sum() / 12).backward()
    err = max(float((g - p2.grad).abs().max()) for g, p2 in zip(grads, net2.parameters()))
    q.put((rank, err, [float(r.sum()) for r in ref], (lo, hi)))
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][2] == out[1][2]                        # parameters identical after broadcast
    assert out[0][3] == (0, 6) and out[1][3] == (6, 12)
    assert out[0][1] < 1e-6 and out[1][1] < 1e-6         # averaged gradients == global-batch gradients


def test_shard_rows_covers_everything():
    from hint_b200.parallel import shard_rows
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
