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


def _worker_values(rank, world, port, out):
    """Deterministic per-name values: rank r writes (r+1)*idx -> average must be 1.5*idx."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unetca_b200
    from unetca_b200 import parallel
    m = unetca_b200.UNet(3, 2, False)
    gb = parallel.GradBuckets(m, bucket_mb=10.0)
    sink = m._grad_sink_factory()
    params = dict(m.named_parameters())
    for i, n in enumerate(gb.names):
        sink.alloc(n, params[n]).fill_(float((rank + 1) * (i + 1)))
        sink.put(n)
    grads = sink.finish()
    ok = all(torch.all(grads[n] == 1.5 * (i + 1)).item() for i, n in enumerate(gb.names))
    ok = ok and all(grads[n].shape == params[n].shape for n in gb.names)
    x = torch.arange(8).view(8, 1)
    ok = ok and parallel.shard_batch(x, rank, world).flatten().tolist() == list(range(rank * 4, rank * 4 + 4))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker, _worker_values])
def test_gloo_world2(worker):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] and out[1]
