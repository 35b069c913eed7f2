"""Host-side logic of the data-parallel path on CPU with gloo, world_size 2: bucket planning follows the backward
completion order, gradients written into the flat buckets are averaged over ranks exactly like the mean of the
per-shard reference gradients, and parameters are broadcast from rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_plan_buckets_and_grad_order():
    import unetca_b200
    from unetca_b200 import parallel
    from unetca_b200.model import grad_order
    m = unetca_b200.UNet(3, 2, True)
    names = grad_order(m)
    params = dict(m.named_parameters())
    assert sorted(names) == sorted(params) and len(names) == 100
    assert names[0] == "outc.weight" and names[-1] == "inc.double_conv.0.weight"
    offsets, bounds, total = parallel.plan_buckets(names, [params[n].numel() for n in names], 25.0)
    assert total == 31_261_698
    assert bounds[0][0] == 0 and bounds[-1][1] == total
    for (s0, e0, _), (s1, _, _) in zip(bounds, bounds[1:]):
        assert e0 == s1
    assert 3 <= len(bounds) <= 8
    # contiguous, non-overlapping slices in completion order
    pos = 0
    for n in names:
        assert offsets[n] == (pos, params[n].numel())
        pos += params[n].numel()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    """Different random init per rank: GradBuckets must broadcast rank 0's parameters and buffers."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unetca_b200
    from unetca_b200 import parallel
    torch.manual_seed(100 + rank)
    m = unetca_b200.UNet(3, 2, True)
    parallel.GradBuckets(m, bucket_mb=25.0)
    ok = True
    for t in (m.outc.weight, m.down4[1].double_conv[3].weight, m.inc.double_conv[1].running_var):
        probe = t.detach().flatten()[:64].clone()
        gathered = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        ok = ok and torch.equal(gathered[0], gathered[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _fill(gb, m, rank, scale):
    """Stand-in for one backward: write deterministic per-name values through a fresh sink, in completion order."""
    sink = m._grad_sink_factory()
    params = dict(m.named_parameters())
    for i, n in enumerate(gb.names):
        sink.alloc(n, params[n]).fill_(float(scale * (rank + 1) * (i + 1)))
        sink.put(n)
    assert sink.finish() is None               # the sink owns .grad: nothing is handed to autograd
    return params


def _worker_values(rank, world, port, out):
    """Deterministic per-name values: rank r writes (r+1)*idx -> average must be 1.5*idx.  `.grad` is a view of the flat
    bucket storage, and neither zero_grad flavour double-counts on the next backward (round-1 advisor finding)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unetca_b200
    from unetca_b200 import parallel
    m = unetca_b200.UNet(3, 2, False)
    gb = parallel.GradBuckets(m, bucket_mb=10.0)
    params = _fill(gb, m, rank, 1.0)
    ok = all(torch.all(params[n].grad == 1.5 * (i + 1)).item() for i, n in enumerate(gb.names))
    ok = ok and all(params[n].grad.shape == params[n].shape for n in gb.names)
    ok = ok and all(params[n].grad.data_ptr() == gb.view(n).data_ptr() for n in gb.names)
    # zero_grad(set_to_none=False): grads zeroed in place, the next backward must give the same numbers (not 2x)
    for p in m.parameters():
        p.grad.zero_()
    params = _fill(gb, m, rank, 1.0)
    ok = ok and all(torch.all(params[n].grad == 1.5 * (i + 1)).item() for i, n in enumerate(gb.names))
    # zero_grad(set_to_none=True)
    for p in m.parameters():
        p.grad = None
    params = _fill(gb, m, rank, 2.0)
    ok = ok and all(torch.all(params[n].grad == 3.0 * (i + 1)).item() for i, n in enumerate(gb.names))
    # no zero_grad at all: torch semantics, .grad accumulates (3.0 + 1.5)
    params = _fill(gb, m, rank, 1.0)
    ok = ok and all(torch.all(params[n].grad == 4.5 * (i + 1)).item() for i, n in enumerate(gb.names))
    x = torch.arange(8).view(8, 1)
    ok = ok and parallel.shard_batch(x, rank, world).flatten().tolist() == list(range(rank * 4, rank * 4 + 4))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _worker_accumulate(rank, world, port, out):
    """BASELINE configs[2] mechanics: k micro-steps per optimizer step, only the k-th backward all-reduces.  Micro-step j
    on rank r writes (j+1)*(r+1)*idx; after k = 3 the gradient must be mean_r sum_j = 6 * 1.5 * idx = 9 * idx, and
    before the synchronising backward the ranks must still hold their own local sums."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unetca_b200
    from unetca_b200 import parallel
    m = unetca_b200.UNet(3, 2, False)
    gb = parallel.GradBuckets(m, bucket_mb=10.0).set_accumulation(3)
    ok = True
    for rep in range(2):                                     # two optimizer steps: the micro-step counter wraps
        for p in m.parameters():
            p.grad = None
        params = _fill(gb, m, rank, 1.0)
        params = _fill(gb, m, rank, 2.0)
        local = 3.0 * (rank + 1)
        ok = ok and all(torch.all(params[n].grad == local * (i + 1)).item() for i, n in enumerate(gb.names))
        params = _fill(gb, m, rank, 3.0)
        ok = ok and all(torch.all(params[n].grad == 9.0 * (i + 1)).item() for i, n in enumerate(gb.names))
    # no_sync() context (DDP's name) overrides the counter
    for p in m.parameters():
        p.grad = None
    gb.set_accumulation(1)
    with gb.no_sync():
        params = _fill(gb, m, rank, 1.0)
    ok = ok and all(torch.all(params[n].grad == float(rank + 1) * (i + 1)).item() for i, n in enumerate(gb.names))
    params = _fill(gb, m, rank, 1.0)
    ok = ok and all(torch.all(params[n].grad == 3.0 * (i + 1)).item() for i, n in enumerate(gb.names))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker, _worker_values, _worker_accumulate])
def test_gloo_world2(worker):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] and out[1]
