#!/bin/bash
# stream-count sweep on one GPU (BASELINE configs[4] per-GPU points: 1k..8k streams), CoST-GCN
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for n in 128 512 1024 2048 4096 8192; do
  timeout 600 python bench.py --streams $n --steps ${SWEEP_STEPS:-100} --warmup 8 --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d = json.loads(l)
    print(d['config']['streams_per_gpu'], round(d['value']), round(d['ms_per_step'], 3), 'p50', round(d['p50_ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
          'step hbm frac', round(d['step_roofline']['hbm_frac'], 3))
PY
