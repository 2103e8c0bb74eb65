"""CPU, world_size 2 over gloo: the N>1 plumbing of the hot path — batch sharding (data/data_loader.py:40-43) and the
gradient all-reduce(mean) that replaces apex DDP (structure/model.py:27-34) — reproduces the single-process gradient
of the global batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megreader_b200 import dp
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x, y = torch.randn(8, 6), torch.randn(8, 1)
    lo, hi = dp.shard_range(8, rank, world)
    loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean()     # per-rank mean, like trainer.py:127 `l.mean()`
    loss.backward()
    dp.allreduce_mean_grads_(list(model.parameters()))
    if rank == 0:
        torch.save([p.grad for p in model.parameters()], out)
    dist.destroy_process_group()


def test_sharded_allreduce_mean_equals_global_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x, y = torch.randn(8, 6), torch.randn(8, 1)
    ((model(x) - y) ** 2).mean().backward()
    for g, p in zip(got, model.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-5, atol=1e-6)


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker_flat(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megreader_b200 import dp
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x, y = torch.randn(8, 6), torch.randn(8, 1)
    params = list(model.parameters())[::-1]                 # backward order
    fg = dp.FlatGrads(params, buckets=2)
    assert fg.n_buckets() == 2 and fg.attached()
    for _ in range(2):                                       # second round: views survive zero() and accumulate afresh
        fg.zero()
        lo, hi = dp.shard_range(8, rank, world)
        ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
        assert fg.attached()
        fg.allreduce_()
    if rank == 0:
        torch.save([p.grad.clone() for p in model.parameters()], out)
    dist.destroy_process_group()


def test_flat_grad_views_bucketed_allreduce(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker_flat, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x, y = torch.randn(8, 6), torch.randn(8, 1)
    ((model(x) - y) ** 2).mean().backward()
    for g, p in zip(got, model.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-5, atol=1e-6)


def test_shard_range_matches_reference_split():
    from megreader_b200 import dp
    assert [dp.shard_range(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]
