#!/bin/bash
# 2-GPU box: TGAT data-parallel training bench at N=1 and N=2 (NCCL gradient all-reduce inside the captured step)
mkdir -p gpurun_out
timeout 300 python bench.py --workload tgat_train --steps 20 --warmup 3 > gpurun_out/bench_tgat_train.json 2> gpurun_out/bench_tgat_train.err; echo "n1 rc=$?"; cut -c1-300 gpurun_out/bench_tgat_train.json; tail -3 gpurun_out/bench_tgat_train.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload tgat_train --steps 20 --warmup 3 > gpurun_out/bench_tgat_train_2gpu.json 2> gpurun_out/bench_tgat_train_2gpu.err; echo "n2 rc=$?"; cat gpurun_out/bench_tgat_train_2gpu.json; tail -5 gpurun_out/bench_tgat_train_2gpu.err
if ! grep -q value gpurun_out/bench_tgat_train_2gpu.json; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload tgat_train --no-graph --steps 20 --warmup 3 > gpurun_out/bench_tgat_train_2gpu_nograph.json 2> gpurun_out/bench_tgat_train_2gpu_nograph.err; echo "n2 nograph rc=$?"; cat gpurun_out/bench_tgat_train_2gpu_nograph.json; tail -5 gpurun_out/bench_tgat_train_2gpu_nograph.err
fi
