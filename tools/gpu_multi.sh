#!/bin/bash
# 2-GPU check of the sharded bench (the driver launches it the same way) + reference arm under torchrun
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 100 --warmup 8 > gpurun_out/bench_2gpu.log 2>&1
echo "bench_2gpu rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_ref.log 2>&1
echo "bench_2gpu_ref rc=$?"
grep '^{' gpurun_out/bench_2gpu.log | cut -c1-700
grep '^{' gpurun_out/bench_2gpu_ref.log | cut -c1-300
tail -5 gpurun_out/bench_2gpu.log | cut -c1-300
