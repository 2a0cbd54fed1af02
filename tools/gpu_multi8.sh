#!/bin/bash
# N-GPU check of the sharded bench, launched the way the driver does; usage: tools/gpu_multi8.sh <N>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
   bench.py --gpus $N --steps 100 --warmup 8 > gpurun_out/bench_${N}gpu.log 2>&1
echo "bench_${N}gpu rc=$?"
grep '^{' gpurun_out/bench_${N}gpu.log | cut -c1-900
tail -3 gpurun_out/bench_${N}gpu.log | cut -c1-300
