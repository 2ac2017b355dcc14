#!/bin/bash
# usage: gpu_scale.sh N   -- N-GPU box: default bench (DyGFormer), TGAT eval, sampler sweep and the training workload under torchrun
N=$1
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/scale_${name}_${N}gpu.json 2> gpurun_out/scale_${name}_${N}gpu.err
  echo "$name N=$N rc=$?"; cut -c1-200 gpurun_out/scale_${name}_${N}gpu.json; tail -2 gpurun_out/scale_${name}_${N}gpu.err | cut -c1-300
}
run dygformer_wiki --steps 5 --warmup 3
run tgat_myket --workload tgat_myket --steps 5 --warmup 3
run sampler --workload sampler_sweep --steps 5
run tgat_train --workload tgat_train --steps 20 --warmup 3
