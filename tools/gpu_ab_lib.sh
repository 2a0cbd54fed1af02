#!/bin/bash
# A/B of two builds of libcosk.so: the in-tree one and tools/<variant>.so (swapped in for its run, then restored).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=${1:-libcosk_h5b2.so}
L=continual-skeletons_b200/csrc/libcosk.so
cp $L /tmp/libcosk_base.so
for rep in 1 2; do
  for which in base variant; do
    if [ $which = variant ]; then cp tools/$V $L; else cp /tmp/libcosk_base.so $L; fi
    touch $L
    timeout 300 python bench.py --steps 200 --warmup 8 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_${which}_$rep.log 2>&1
  done
done
cp /tmp/libcosk_base.so $L; touch $L
if [ -n "$QUICK_K" ]; then cp tools/$V $L; touch $L; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 60 -k "$QUICK_K" 2>&1 | tail -3; cp /tmp/libcosk_base.so $L; touch $L; fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.log')):
    txt=open(f).read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'p50', round(d.get('p50_ms_per_step'),3), d.get('clocks',{}).get('sm_mhz'),
                  'blk', [round(b['block_ms']/max(b['block_n'],1),4) for b in pb][1:4], 'gcn', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb][4:8])
    if 'Traceback' in txt: print(txt[-800:])
PY
