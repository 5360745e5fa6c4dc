"""The N>1 path on the CPU: world_size-2 `gloo` processes run the bucket planner and the bucketed gradient all-reduce
(denet_b200/multi/ddp.py) that replaces the reference's host-side parameter averaging (multi/shared.py:105-119)."""
import os
import socket

import numpy
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from denet_b200.multi import ddp


def test_plan_buckets_covers_every_gradient_once_in_backward_order():
    ranges = [(1, 0, 100), (2, 100, 100), (3, 100, 1100), (5, 1100, 1200), (9, 1200, 5000)]
    buckets = ddp.plan_buckets(ranges, 1000)
    covered = sorted((s, e) for _, s, e in buckets)
    assert covered[0][0] == 0 and covered[-1][1] == 5000
    for (s0, e0), (s1, e1) in zip(covered[:-1], covered[1:]):
        assert e0 == s1                                   # contiguous, no overlap
    ready = [r for r, _, _ in buckets]
    assert ready == sorted(ready, reverse=True)           # launched as backward descends through the layers
    for r, s, e in buckets:
        owners = [l for l, ls, le in ranges if ls < e and le > s and le > ls]
        assert min(owners) >= r                           # every owner of the bucket has finished by layer r


def test_shard_batch():
    assert ddp.shard_batch(64, 0, 2) == (0, 32) and ddp.shard_batch(64, 1, 2) == (32, 64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w = ddp.init_process_group("gloo")
    assert (r, w) == (rank, world)
    rs = numpy.random.RandomState(100 + rank)
    flat = torch.from_numpy(rs.randn(5000).astype(numpy.float32))
    mine = flat.clone()
    bn_stat = torch.full((7,), float(rank + 1))
    ranges = [(1, 0, 100), (3, 100, 1100), (5, 1100, 1200), (9, 1200, 5000)]
    red = ddp.GradientAllReduce(flat, ranges, bucket_bytes=4000, extra_mean_tensors=[bn_stat])
    red.begin_step()
    for layer in range(9, 0, -1):                         # the backward pass retires layers in descending order
        red.layer_done(layer)
    scale = red.finish_step()
    # several extra tensors (gathered, reduced, scattered back) and the will_launch() contract the deferred split-K
    # reduction relies on: it announces exactly the layer_done() calls that start an all-reduce
    flat2 = torch.from_numpy(rs.randn(5000).astype(numpy.float32))
    ea, eb = torch.full((3,), float(rank)), torch.full((2, 2), 10.0 * (rank + 1))
    red2 = ddp.GradientAllReduce(flat2, ranges, bucket_bytes=4000, extra_mean_tensors=[ea, eb])
    red2.begin_step()
    announced = []
    for layer in range(9, 0, -1):
        before = red2._next
        will = red2.will_launch(layer)
        red2.layer_done(layer)
        assert will == (red2._next > before), layer
        announced.append(will)
    red2.finish_step()
    assert any(announced)
    assert torch.allclose(ea, torch.full((3,), 0.5)) and torch.allclose(eb, torch.full((2, 2), 15.0))
    out.put((rank, mine.numpy(), flat.numpy().copy(), scale, bn_stat.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for rank, _, reduced, scale, bn in res:
        assert numpy.allclose(reduced, total, rtol=1e-6, atol=1e-6)      # sum over ranks in every bucket
        assert scale == 0.5                                              # the solver applies 1/world
        assert numpy.allclose(bn, 1.5)                                   # running statistics are averaged


def _worker_sync(rank, world, port, out):
    """replicas built from different numpy seeds must be made identical by the data-parallel setup"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    ddp.init_process_group("gloo")
    from denet_b200.model import model_cnn
    numpy.random.seed(1000 + rank)                        # what an unseeded torchrun launch amounts to
    model = model_cnn.ModelCNN()
    model.batch_size, model.class_num = 2, 10
    model.build("C[8,3] BN A P[2] C.B[16,3] BN A P.A R".split(), (3, 8, 8), "relu", "half", ["he-backward"])
    tensors = model.state_tensors()
    for t in tensors[-2:]:
        t.data.add_(float(rank))                          # running statistics differ too
    before = ddp.state_checksum(tensors)
    n = model.sync_state_from_rank0()
    after = ddp.state_checksum(tensors)
    gathered = [None] * world
    dist.all_gather_object(gathered, (before, after, n, len(tensors)))
    out.put((rank, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_built_with_different_seeds_are_synchronised_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sync, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (b0, a0, n0, k0), (b1, a1, n1, k1) = res[0][1]
    assert b0 != b1                     # the replicas really started from different weights
    assert a0 == a1 == b0               # ... and both hold rank 0's state afterwards
    assert n0 == n1 == k0 == k1 and k0 >= 8
