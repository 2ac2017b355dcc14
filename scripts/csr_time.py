import sys, time, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from dyglib_b200.synthetic import make_config_graph
from dyglib_b200.utils.utils import get_neighbor_sampler
g = make_config_graph('dygformer_wiki')
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = get_neighbor_sampler(g, 'recent', seed=0)
    torch.cuda.synchronize(); print('build %d: %.1f ms' % (i, (time.perf_counter() - t0) * 1e3), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); s = get_neighbor_sampler(g, 'recent', seed=0); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
