import sys, time, torch
sys.path.insert(0, '.')
from dyglib_b200 import ops
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
for n in (315_000, 20_000_000, 200_000_000):
    g=torch.Generator(device='cuda').manual_seed(1)
    owner=torch.randint(0, 10_000_000, (n,), device='cuda', generator=g, dtype=torch.int64)
    o32=owner.to(torch.int32)
    tk=ops.float64_sort_key(torch.arange(n, device='cuda', dtype=torch.float64).repeat_interleave(1))
    print(n, 'own u32 %.2f ms'%t(lambda: ops.stable_argsort(o32)), 'torch %.2f ms'%t(lambda: torch.sort(owner, stable=True)),
          'own f64 %.2f ms'%t(lambda: ops.stable_argsort(tk)), flush=True)
