"""world_size-2 gloo tests (CPU) of the data-parallel host logic: graph sharding and the single flat-gradient
all-reduce.  The kernels are not involved (no GPU here); the GPU path is exercised by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relpose_gnn_b200 import parallel


def test_shard_range_partitions_contiguously():
    for total, world in [(4096, 8), (65536, 8), (10, 3), (7, 8), (2048, 2)]:
        spans = [parallel.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                   # identical replicas
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    bucket = parallel.FlatGradBucket(lin.parameters())
    assert bucket.numel == sum(p.numel() for p in lin.parameters())
    # each rank owns a contiguous block of "graphs"; loss = mean over the rank's block
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lo, hi = parallel.shard_range(8, rank, world)
    for _ in range(2):                                      # second step checks zero() + in-place accumulation
        bucket.zero()
        lin(data[lo:hi]).pow(2).mean().backward()
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())   # still views
        if _ == 0:
            bucket.allreduce()
        else:       # overlapped form: everything but the first layer's weight goes first, the rest at the end
            bucket.begin_allreduce(bucket.ranges_excluding([lin[0].weight]))
            bucket.finish_allreduce()
    if rank == 0:
        torch.save(bucket.flat.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "flat.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lin(data).pow(2).mean().backward()                     # equal shards + mean loss => averaged grads == full-batch grads
    want = torch.cat([p.grad.flatten() for p in lin.parameters()])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)


def test_bucket_reattaches_after_zero_grad_set_to_none():
    """optimizer.zero_grad() (set_to_none=True, train.py:252) detaches p.grad from the flat buffer; the
    bucket must copy the fresh gradients back before the collective instead of reducing a stale buffer."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    params = list(lin.parameters())
    bucket = parallel.FlatGradBucket(params)
    x = torch.randn(5, 4)
    lin(x).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    assert torch.equal(bucket.allreduce(), want)
    torch.optim.SGD(params, lr=0.1).zero_grad()                    # every p.grad is None now
    assert all(p.grad is None for p in params)
    lin(x).sum().backward()                                        # autograd allocates fresh gradient tensors
    assert all(p.grad.data_ptr() != v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    assert torch.equal(bucket.allreduce(), want)                   # copied back and re-pointed
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    bucket.zero()
    assert float(bucket.flat.abs().sum()) == 0.0 and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))


def test_ranges_excluding_late_parameters():
    a, b, c = (torch.nn.Parameter(torch.zeros(n)) for n in (3, 5, 2))
    bucket = parallel.FlatGradBucket([a, b, c])
    assert bucket.ranges_excluding([a]) == [(3, 10)]
    assert bucket.ranges_excluding([b]) == [(0, 3), (8, 10)]
    assert bucket.ranges_excluding([c]) == [(0, 8)]
    assert bucket.ranges_excluding([]) == [(0, 10)]
    bucket.begin_allreduce([(3, 10)])                 # no process group: nothing to do
    assert bucket.finish_allreduce() is bucket.flat
