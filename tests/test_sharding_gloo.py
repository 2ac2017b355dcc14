"""CPU, world_size 2 over gloo: the bench's N>1 path (whole reference batches round-robin over ranks, no
data-path collective, final score gather) reproduces the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scores_for(batch_ids):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import small_graph, oracle_factories
    import bench
    g = small_graph(seed=12)
    _, _, odyg, _ = oracle_factories()
    m = odyg(g, 2, 16, 2)
    stream = bench.Stream(g, batch=25, region=0.5)
    out = []
    with torch.no_grad():
        for b in batch_ids:
            src, dst, neg, t, _ = stream.rows([b])
            a, c = m.compute_src_dst_node_temporal_embeddings(src, dst, t)
            out.append((a * c).sum(dim=1))
    return torch.stack(out)


def _worker(rank, world, port, G, steps, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    mine = [b for s in range(steps) for b in bench.shard_batches(s, G, world, rank, 1000)]
    sc = _scores_for(mine)
    gathered = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(gathered, sc)                      # the final score gather
    ids = [torch.empty(len(mine), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ids, torch.tensor(mine))
    if rank == 0:
        ret['ids'] = torch.cat(ids).numpy()
        ret['scores'] = torch.cat(gathered).numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_shards_cover_and_match_single_process():
    world, G, steps = 2, 2, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, G, steps, ret), nprocs=world, join=True)
    ids, scores = ret['ids'], ret['scores']
    assert sorted(ids.tolist()) == list(range(world * G * steps))      # disjoint, complete, whole batches
    want = _scores_for(list(range(world * G * steps))).numpy()
    order = np.argsort(ids)
    np.testing.assert_array_equal(scores[order], want)                 # batch-level sharding is exact


def test_shard_batches_weak_scaling():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += bench.shard_batches(3, 5, world, r, 10 ** 6)
        assert len(set(seen)) == 5 * world
