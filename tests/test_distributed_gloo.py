"""N>1 host logic on CPU: world_size 2, gloo.  Scene sharding, the single bucketed gradient
all-reduce and the max-over-ranks timing rule that bench.py relies on."""
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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepviewagg_b200 import distributed as D
    from deepviewagg_b200.modules.multimodal.pooling import GroupBimodalCSRPool
    # identical replicas (same seed), different "scenes" per rank -> different gradients
    torch.manual_seed(0)
    m = GroupBimodalCSRPool(in_map=8, in_mod=16, num_groups=4, use_num=True)
    shard = D.shard_indices(7, rank, world)
    gen = torch.Generator().manual_seed(100 + rank)
    for p in m.parameters():
        p.grad = torch.randn(p.shape, generator=gen)
    local = [p.grad.clone() for p in m.parameters()]
    n = D.allreduce_gradients(m.parameters(), average=True)
    t = D.max_over_ranks(1.0 + rank)
    s = D.sum_over_ranks(10.0)
    # plain python lists: tensors in a Queue are shared through fds that die with the sender
    q.put((rank, shard, n, t, s, [g.reshape(-1).tolist() for g in local],
           [p.grad.reshape(-1).tolist() for p in m.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_grad_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, n0, t0, sum0, l0, g0), (r1, s1, n1, t1, sum1, l1, g1) = res
    assert sorted(s0 + s1) == list(range(7)) and not set(s0) & set(s1)   # disjoint cover
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5]
    assert n0 == n1 == sum(len(g) for g in g0) > 0                          # one bucket, all grads
    assert t0 == t1 == 2.0 and sum0 == sum1 == 20.0                         # max / sum over ranks
    for a, b, x, y in zip(l0, l1, g0, g1):
        a, b, x, y = (torch.tensor(v) for v in (a, b, x, y))
        assert torch.allclose(x, (a + b) / 2, atol=1e-6) and torch.equal(x, y)


def _worker_missing(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepviewagg_b200 import distributed as D
    ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2)),
          torch.nn.Parameter(torch.zeros(5), requires_grad=False), torch.nn.Parameter(torch.zeros(1))]
    # rank 0: gradients for p0 and p1; rank 1 (batch without modality data): only p0.  p3 is unused
    # on both ranks, p2 is frozen.
    ps[0].grad = torch.full((3,), 1.0 + rank)
    if rank == 0:
        ps[1].grad = torch.full((2, 2), 4.0)
    n = D.allreduce_gradients(ps, average=True)
    q.put((rank, n, [None if p.grad is None else p.grad.reshape(-1).tolist() for p in ps]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_layout_does_not_depend_on_local_grads():
    """ranks with different sets of existing gradients must still reduce one identical
    bucket (zeros for missing ones) instead of hanging or mis-aligning."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_missing, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, grads in res:
        assert n == 3 + 4 + 1                               # every requires_grad parameter, frozen one excluded
        assert grads[0] == [1.5] * 3                        # (1 + 2) / 2
        assert grads[1] == [2.0] * 4                        # (4 + 0) / 2 on BOTH ranks
        assert grads[2] is None and grads[3] == [0.0]


def test_numa_binding_is_best_effort():
    from deepviewagg_b200 import distributed as D
    assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    info = D.bind_to_gpu_numa_node(0)                       # no GPU here: reports the error, never raises
    assert isinstance(info, dict) and "node" in info


def test_single_process_is_a_noop():
    from deepviewagg_b200 import distributed as D
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert D.allreduce_gradients([p]) == 0 and D.max_over_ranks(3.5) == 3.5
    assert D.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
