#!/bin/bash
# A/B of two prebuilt libraries on the same box: scratch_libs/lib_old.so vs lib_new.so (micro-bench of ln_ffn + the headline step, twice each)
for round in 1 2; do
  for v in old new; do
    cp scratch_libs/lib_$v.so dyglib_b200/libdygb200.so
    echo "== $v (round $round)"
    FFN_M=819200 timeout 200 python scripts/ffn_bench.py 2>&1 | tail -1
    timeout 400 python bench.py --only-headline --no-eager --cpu-batches 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']/10,3) for k,v in d['kernels'].items()})"
  done
done
