"""CPU, world_size 2 over gloo: the data-parallel gradient all-reduce of the training configuration (one flat bucket whose
views are the parameters' .grad) gives every rank the mean gradient, and two ranks stepping on their own halves of a batch
stay identical to one process stepping on the whole batch."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(32, 8, generator=g), torch.randn(32, 1, generator=g)


def _train(rank, world, steps=3):
    from dyglib_b200.utils.dist import GradBucket
    m = _model()
    bucket = GradBucket(m.parameters())
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    x, y = _data()
    n = x.shape[0] // world
    xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
    for _ in range(steps):
        bucket.zero()
        torch.nn.functional.mse_loss(m(xs), ys).backward()
        bucket.check_views()
        bucket.allreduce()
        opt.step()
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ret[rank] = _train(rank, world).numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_allreduce_matches_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    single = _train(0, 1).numpy()
    import numpy as np
    np.testing.assert_allclose(ret[0], ret[1], rtol=0, atol=0)            # ranks stay in lock step
    np.testing.assert_allclose(ret[0], single, rtol=1e-5, atol=1e-6)      # mean of half-batch gradients = whole-batch gradient
