#!/bin/bash
# BASELINE configs[4]: CoST-GCN NTU60 stream-scaling sweep, total streams sharded evenly over G GPUs of one box.
#   gpurun --gpus G -- 'bash tools/gpu_sweep_r2.sh G "1024 2048 ..."'   (totals; per-GPU count = total / G)
# Appends one bench.py JSON line per point to gpurun_out/sweep_r2_g<G>.jsonl
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-1}
TOTALS=${2:-"1024 2048 4096 8192 16384 32768"}
out=gpurun_out/sweep_r2_g$G.jsonl
: > $out
for total in $TOTALS; do
  per=$((total / G))
  if [ $G -eq 1 ]; then
    timeout 900 python bench.py --streams $per --steps 40 --warmup 4 --prewarm-s 0.5 --no-cpu-baseline > gpurun_out/sweep_tmp.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29650 \
      bench.py --gpus $G --streams $per --steps 40 --warmup 4 --prewarm-s 0.5 --no-cpu-baseline > gpurun_out/sweep_tmp.log 2>&1
  fi
  rc=$?
  line=$(grep '^{' gpurun_out/sweep_tmp.log | tail -1)
  if [ -z "$line" ]; then  # one retry (a failed rendezvous on a fresh box is not a result)
    cp gpurun_out/sweep_tmp.log gpurun_out/sweep_fail_${G}_${total}.log
    sleep 3
    if [ $G -eq 1 ]; then
      timeout 900 python bench.py --streams $per --steps 40 --warmup 4 --prewarm-s 0.5 --no-cpu-baseline > gpurun_out/sweep_tmp.log 2>&1
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29651 \
        bench.py --gpus $G --streams $per --steps 40 --warmup 4 --prewarm-s 0.5 --no-cpu-baseline > gpurun_out/sweep_tmp.log 2>&1
    fi
    rc=$?
    line=$(grep '^{' gpurun_out/sweep_tmp.log | tail -1)
  fi
  if [ -n "$line" ]; then echo "$line" >> $out; else echo "{\"streams_total\": $total, \"n_gpus\": $G, \"failed\": $rc, \"tail\": \"$(tail -3 gpurun_out/sweep_tmp.log | tr '\n"' ' .' | cut -c1-300)\"}" >> $out; fi
done
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    if "value" not in d:
        print(d); continue
    print(f"G={d['n_gpus']} total={d['config']['streams_total']:6d} per_gpu={d['config']['streams_per_gpu']:6d} value={d['value']/1e6:7.3f} M sf/s  e2e={d['e2e']['value']/1e6:7.3f} M  ms/step={d['ms_per_step']:.3f} p50={d['p50_ms_per_step']:.3f}  state={d['state_bytes']/1e9:.1f} GB/GPU  step hbm frac={d['step_roofline']['hbm_frac']:.3f}")
PY
