#!/bin/bash
# A/B of scratch_libs/lib_old.so vs lib_new.so on the same box over bench workloads: ab_libs_wl.sh wl1 wl2 ...
for round in 1 2; do
  for v in old new; do
    cp scratch_libs/lib_$v.so dyglib_b200/libdygb200.so
    for wl in "$@"; do
      timeout 400 python bench.py --workload $wl --only-headline --no-eager --cpu-batches 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', '$wl', round(d['value']), round(d['ms_per_step'],4), d.get('parity_max_abs_err'), {k:round(v['ms']/10,3) for k,v in d.get('kernels',{}).items()})"
    done
  done
done
