"""N>1 host logic on CPU: two gloo ranks shard requests, receive the broadcast weights,
time with max-over-ranks, and rank 0 collects results in request order."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from lanpaint_b200.replicas import shard_bounds, shard_sizes


def test_shard_bounds_cover_every_request_once():
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    assert shard_sizes(32, 8) == [4] * 8                      # BASELINE config 3: 4 requests per GPU
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_requests, out_q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from lanpaint_b200.replicas import ReplicaGroup
    grp = ReplicaGroup(backend="gloo")
    w = torch.tensor([0.7, 0.1, 0.0, 0.6, -0.05]) if rank == 0 else torch.zeros(5)
    moved = grp.broadcast_weights([w])
    sl = grp.my_slice(n_requests)
    requests = torch.arange(n_requests, dtype=torch.float32).view(-1, 1, 1).expand(n_requests, 2, 3)
    local = requests[sl] * w[0] + rank * 0.0      # stand-in for "run my shard through my replica"
    grp.barrier()
    slowest = grp.max_over_ranks(10.0 + rank)
    total = grp.sum_over_ranks(float(local.shape[0]))
    full = grp.gather_results(local.contiguous(), n_requests)
    out_q.put((rank, w.tolist(), moved, (sl.start, sl.stop), slowest, total,
               None if full is None else full[:, 0, 0].tolist()))
    grp.close()


@pytest.mark.parametrize("n_requests", [5, 8])
def test_two_rank_replicas_over_gloo(n_requests):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_requests, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, moved0, s0, slow0, tot0, full0), (r1, w1, moved1, s1, slow1, tot1, full1) = got
    assert w0 == w1 == pytest.approx([0.7, 0.1, 0.0, 0.6, -0.05]) and moved0 == moved1 == 20
    assert s0[0] == 0 and s0[1] == s1[0] and s1[1] == n_requests
    assert slow0 == slow1 == 11.0 and tot0 == tot1 == n_requests
    assert full1 is None
    assert full0 == pytest.approx([0.7 * i for i in range(n_requests)])
