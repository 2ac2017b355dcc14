#!/bin/bash
# same-box A/B of scratch_libs/lib_old.so vs lib_new.so on the patch projection (micro-bench + headline + lastfm)
for round in 1 2; do
  for v in old new; do
    cp scratch_libs/lib_$v.so dyglib_b200/libdygb200.so
    echo "== $v (round $round)"
    timeout 200 python scripts/pp_bench.py 2>&1 | tail -1
    for wl in dygformer_wiki dygformer_lastfm; do
    timeout 400 python bench.py --workload $wl --only-headline --no-eager --cpu-batches 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$wl', round(d['value']), round(d['ms_per_step'],3), d.get('parity_max_abs_err'), {k:round(v['ms']/10,3) for k,v in d['kernels'].items() if 'patch' in k})"
    done
  done
done
