#!/bin/bash
# Round 2: validation + profiling artifacts (full tests, parity report, benches, scripts, ncu full + launch list).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_all 1800 python -m pytest tests -m gpu -q
run parity_report 900 python tools/gpu_parity_report.py
run bench_auto 900 python bench.py --steps 200 --warmup 8
run bench_auto_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
run bench_coa 900 python bench.py --workload coa_gcn --steps 100 --warmup 8 --no-cpu-baseline
run bench_cos 900 python bench.py --workload cos_tr --streams 2048 --steps 100 --warmup 8 --no-cpu-baseline
run bench_reference 900 python bench.py --impl reference --steps 3 --warmup 1
run bench_script_ntu 900 python scripts/benchmark_all_ntu60.py
run bench_script_kin 900 python scripts/benchmark_all_kinetics.py
cat > /tmp/batched.py <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for tc in (1, 4, 8):
    m = cs.CoStGcn({"dataset_name": "dummy_ntu", "forward_mode": "frame", "time_chunk": tc})
    N, T = 4096, 64
    x = torch.rand(N, 3, T, 25, 2, device="cuda")
    w = torch.rand(N, 3, 296, 25, 2, device="cuda")
    m.forward_steps(w)
    del w
    m.forward_steps(x)
    torch.cuda.synchronize()
    l0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.forward_steps(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"forward_steps_time_chunk": tc, "streams": N, "frames_per_call": T, "stream_frames_per_s": N * T / (ms * 1e-3), "ms_per_frame": ms / T,
                      "launches_per_call": (m.launch_count() - l0) / 3, "state_GB": m.state_bytes() / 1e9}), flush=True)
    del m, x
    torch.cuda.empty_cache()
PY
run batched_throughput 900 python /tmp/batched.py
COSK_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/summary.txt
COSK_NCU=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none \
   -k regex:"k_tc_|k_gcn_small|k_head|k_input" -c 46 -o gpurun_out/main_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/main_full.log 2>&1
echo "ncu_full rc=$?" >> gpurun_out/summary.txt
ls -la gpurun_out/main_full.ncu-rep >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -v "^E   " gpurun_out/pytest_all.log | tail -8 | cut -c1-300
cat gpurun_out/parity_report.log | cut -c1-220
cat gpurun_out/batched_throughput.log | cut -c1-300
cat gpurun_out/bench_script_ntu.log gpurun_out/bench_script_kin.log | cut -c1-300
