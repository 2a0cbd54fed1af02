#!/bin/bash
# A/B helper: bench with an environment knob at several values (usage: gpu_ab_vals.sh KNOB v1 v2 ...), two rounds, interleaved
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KNOB=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    env $KNOB=$v timeout 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline ${AB_ARGS:-} > gpurun_out/ab_${v}_$rep.log 2>&1
    python - "$KNOB=$v rep$rep" gpurun_out/ab_${v}_$rep.log <<'PY'
import json, sys
l = [x for x in open(sys.argv[2]) if x.startswith('{')]
if not l:
    print(sys.argv[1], 'FAILED'); sys.exit(0)
d = json.loads(l[-1])
pb = d['kernel_time_per_block_ms']
print(sys.argv[1], round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'gcn', [round(b['gcn_ms'] / max(1, b['gcn_n']), 4) for b in pb], 'clk', d['clocks']['sm_mhz'])
PY
  done
done
