#!/bin/bash
# BASELINE configs[4] end point: 64k streams sharded over 8 GPUs (8192 per GPU), plus 1k per GPU for the small end
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/sweep8.jsonl
for n in 1024 8192; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 \
     bench.py --gpus 8 --streams $n --steps 100 --warmup 8 --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/sweep8.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep8.jsonl'):
    d = json.loads(l)
    print(d['config']['streams_total'], 'streams on', d['n_gpus'], 'GPUs:', round(d['value']), 'stream-frames/s', round(d['ms_per_step'], 3), 'ms/step, e2e', round(d['e2e']['value']))
PY
