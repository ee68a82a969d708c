"""The N>1 path on CPU: two gloo ranks, GradBucket all-reduce == average of the per-rank gradients, and a
data-parallel step over a split batch reproduces the single-process gradient (SURVEY §4 item v)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, overlap=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from segtran_b200.parallel import GradBucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    net[2].weight = net[2].weight                      # plain module; tying handled by id() de-duplication
    milestones = overlap == "milestones"
    bucket = GradBucket(net.parameters(), overlap_chunks=0 if milestones else overlap, milestones=milestones)
    x = torch.arange(4 * 6, dtype=torch.float32).view(4, 6) / 10.0
    xs = x[rank * 2:(rank + 1) * 2]                    # batch-split data parallelism (train3d.py:495)
    sent_early = []
    for it in range(2):                                # second iteration checks that .grad views stay attached
        bucket.zero()
        if milestones:                                 # layer-boundary milestone: backward past h => net[2]'s grads final
            from segtran_b200 import ops
            h = net[1](net[0](xs))
            ops.grad_ready(h, net[2].parameters())
            (net[2](h).pow(2).sum() / 2).backward()
            sent_early.append(sum(bucket._sent) == 2 and len(bucket._works) == 1)   # weight+bias merged into one call
        else:
            (net(xs).pow(2).sum() / 2).backward()
        bucket.allreduce_async()
        bucket.wait()
    # the bucket pads every parameter to a 256-byte boundary: compare the parameter views, and the padding must stay zero
    flat = torch.cat([p.grad.reshape(-1) for p in bucket.params])
    pad_ok = abs(float(bucket.flat.sum()) - float(flat.sum())) < 1e-4
    # single-process reference on the full batch: mean over ranks of sum-over-local == (sum over batch) / world
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    (ref(x).pow(2).sum() / 2 / world).backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    ok = pad_ok and torch.allclose(flat, want, rtol=1e-5, atol=1e-6) and all(
        p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters()) and all(sent_early)
    q.put((rank, bool(ok), float((flat - want).abs().max())))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("overlap", [0, 3, "milestones"])
def test_grad_bucket_two_ranks_gloo(overlap):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, overlap)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_grad_bucket_single_process_dedups_tied_parameters():
    from segtran_b200.parallel import GradBucket
    a = torch.nn.Linear(4, 4)
    b = torch.nn.Linear(4, 4)
    b.weight = a.weight
    bucket = GradBucket(list(a.parameters()) + list(b.parameters()))
    assert len(bucket.params) == 3 and bucket.numel == 3 * 64          # tied weight once; 256-byte slots
    (b(a(torch.ones(2, 4))).sum()).backward()
    assert a.weight.grad.data_ptr() == bucket.flat.data_ptr()
    assert float(bucket.flat.abs().sum()) > 0
