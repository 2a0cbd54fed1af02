#!/usr/bin/env python
"""Benchmark of the continual ST-GCN per-step forward (BASELINE.json metric: CoST-GCN NTU60
stream-frames/s at 1/2/4/8 B200; p50 per-step latency).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference algorithm on host cores

One "step" = one new frame for every concurrent stream pushed through ``forward_step``.  Streams are
sharded over ranks (one process per GPU, weak scaling: ``--streams`` per GPU); the only collective is
the all-gather of logits on emitting steps.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "CoST-GCN NTU60 stream-frames/s"
UNIT = "stream-frames/s"
V, S, C_IN, CLASSES = 25, 2, 3, 60
# algorithmic bytes per stream-frame (SURVEY.md section 8d / BASELINE.md section 3): state bytes moved
# + input frame + logits per emission + pool running-sum minimum
ALGO = {
    "cost_gcn": {"state": 1.408e6, "io": 600 + 240 / 4 + 2048 / 4, "flops": 115.0e6, "warm": 297, "period": 4},
    "cost_gcn_mod": {"state": 3.072e6, "io": 600 + 240 + 2048, "flops": 318.0e6, "warm": 300, "period": 1},
    # CoA-GCN: same rings and schedule as CoST-GCN; FLOPs scaled by the paper's
    # per-prediction costs, 0.30 G vs 0.27 G (figures/table-2.png)
    "coa_gcn": {"state": 1.408e6, "io": 600 + 240 / 4 + 2048 / 4, "flops": 115.0e6 * 0.30 / 0.27, "warm": 297, "period": 4},
    # CoS-TR on the 18-joint Kinetics skeleton (BASELINE configs[3]): CoST-GCN's rings
    # scaled by 18/25 vertices; FLOPs from the paper's 0.16 G per prediction on Kinetics (figures/table-5.png)
    "cos_tr": {"state": 1.408e6 * 18 / 25, "io": 432 + 1600 / 4 + 2048 / 4, "flops": 0.16e9 / 4 * 2, "warm": 297, "period": 4},
}
NAMES = {"cost_gcn": "CoST-GCN", "cost_gcn_mod": "CoST-GCN*", "coa_gcn": "CoA-GCN", "cos_tr": "CoS-TR"}
# workload -> V, classes, dataset, label (BASELINE configs[2] quotes CoA-GCN on NTU RGB+D 120, configs[3] CoS-TR on Kinetics)
GEOMETRY = {"cos_tr": (18, 400, "dummy_kin", "Kinetics-400 skeleton"), "coa_gcn": (25, 120, "ntu120", "NTU RGB+D 120 joint stream")}
DATASET, DATA_LABEL = "dummy_ntu", "NTU RGB+D 60 joint stream"


def set_geometry(workload):
    """Vertex count / class count of the workload's dataset (module-level, read by every arm)."""
    global V, CLASSES, DATASET, DATA_LABEL
    V, CLASSES, DATASET, DATA_LABEL = GEOMETRY.get(workload, (25, 60, "dummy_ntu", "NTU RGB+D 60 joint stream"))


# dram__bytes_read.sum + dram__bytes_write.sum per launch at 4096 streams (bytes), averaged over the launches listed in
# profiles/r1n_ncu_full_main_summary.csv (ncu --set full of one full step; earlier capture with --cache-control none:
# profiles/r1h_dram_bytes_per_launch.csv); None where no capture exists.
NCU_TRAFFIC = {"tcn<64>": 552.0e6, "tcn<128>": 1115.0e6, "tcn<256>": 2256.0e6, "gcn<64>": 59.0e6, "gcn<128>": 158.0e6,
               "gcn<256>": 387.0e6}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(workload, n_streams, steps, warm_extra=0, threads=None):
    """The oracle's per-step restatement (oracle/step.py) timed on the host cores: stream-frames/s."""
    from oracle import step, weights

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    arch = {"cost_gcn": weights.cost_gcn_arch, "cost_gcn_mod": weights.cost_gcn_mod_arch, "cos_tr": weights.cos_tr_arch,
            "coa_gcn": lambda: weights.coa_gcn_arch(classes=120)}[workload]()
    sd = weights.make_state_dict(arch, seed=0)
    model = step.StepModel(sd, arch)
    frames = [torch.rand(n_streams, C_IN, V, S) for _ in range(4)]
    per = []
    with torch.no_grad():
        for t in range(ALGO[workload]["warm"] + warm_extra):
            model.forward_step(frames[t % 4])
        model.trace = []
        t0 = time.perf_counter()
        for t in range(steps):
            s0 = time.perf_counter()
            model.forward_step(frames[t % 4])
            per.append(time.perf_counter() - s0)
        dt = time.perf_counter() - t0
    return n_streams * steps / dt, dt / steps * 1e3, statistics.median(per) * 1e3, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = args.ref_streams
    config = make_config(args, world)
    rate, ms, p50, cores = cpu_port_rate(args.workload, n_sample, args.steps, warm_extra=args.warmup)
    sample = (f"{n_sample} concurrent streams per step (bounded sample of the {args.streams}-stream workload), steady state "
              f"after {ALGO[args.workload]['warm']} warm frames, oracle/step.py eager torch fp32")
    line = {
        "impl": "reference", "metric": METRIC.replace("CoST-GCN", NAMES[args.workload]).replace("NTU60", "Kinetics" if V == 18 else "NTU120" if CLASSES == 120 else "NTU60"), "value": rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "p50_ms_per_step": p50, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def kernel_label(path, kname, cout, kblock, workload="cost_gcn"):
    if path != "auto":
        return f"k_{kname}_simt (layer {kblock + 1}, C={cout})"
    if kname == "tcn":
        return (f"k_tc_tcn<{cout}>" if cout == 64 else f"k_tc_tcn2<{cout}> (CTA pairs)") + f" (layer {kblock + 1})"
    if workload == "coa_gcn":
        return f"k_tc_agcn C={cout} (layer {kblock + 1}; dense per-skeleton mix, attention kernel timed separately)"
    if workload == "cos_tr" and kblock >= 3:
        return f"attention-unit output conv, k_tc_tcn with one tap, C={cout} (layer {kblock + 1}; qkv + attention timed separately)"
    return f"k_tc_gcn<4> C={cout} (layer {kblock + 1})"


def make_config(args, world):
    return {
        "workload": f"{NAMES[args.workload]} {DATA_LABEL}, per-step forward_step, {args.streams} concurrent streams per GPU, "
                    f"random-init weights, synthetic U[0,1) frames (N,C=3,V={V},S=2)",
        "model_variant": args.workload, "streams_per_gpu": args.streams, "streams_total": args.streams * world,
        "V": V, "S": S, "classes": CLASSES, "sharding": f"streams sharded over {world} rank(s), logits all-gathered on emitting steps",
        "l2": "per-step state traffic is GBs (>> 126 MB L2) and 8 distinct input frames are cycled, so no L2 flush is needed",
    }


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    import continual_skeletons_b200 as cs

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    algo = ALGO[args.workload]
    cls = {"cost_gcn": cs.CoStGcn, "cost_gcn_mod": cs.CoStGcnMod, "coa_gcn": cs.CoAGcn, "cos_tr": cs.CoSTr}[args.workload]
    torch.manual_seed(0)
    model = cls({"dataset_name": DATASET, "forward_mode": "frame", "kernel_path": args.kernel_path})
    n_local = args.streams
    n_total = n_local * world
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_frames = [torch.rand((n_local, C_IN, V, S), generator=gen).pin_memory() for _ in range(8)]
    dev_frames = [f.to(dev) for f in host_frames]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(t):
        out = model.forward_step(dev_frames[t % 8])
        if out is not None and world > 1:
            out = cs.all_gather_logits(out, n_total)
        return out

    # state warm-up to steady state (every ring full, logits emitting), then W untimed steps
    t = 0
    for _ in range(algo["warm"]):
        step_resident(t)
        t += 1
    for _ in range(max(args.warmup, 3)):
        step_resident(t)
        t += 1
    assert model.device_error() == 0, hex(model.device_error())

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = model.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu = os.environ.get("COSK_NCU") == "1"  # profile only the timed steps: ncu --profile-from-start off
    if ncu:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for _ in range(args.steps):
        step_resident(t)
        t += 1
    ev1.record()
    barrier()
    if ncu:
        torch.cuda.cudart().cudaProfilerStop()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = model.launch_count() - launches0
    # ---- timed region 1b: the same K steps again with the library's per-kernel CUDA events on the
    # launching stream (an event between two kernels serialises them, so this pass is kept out of `value`)
    model.profile(True)
    for _ in range(args.steps):
        step_resident(t)
        t += 1
    barrier()
    prof = {}
    for kind, name in ((0, "input"), (1, "gcn"), (2, "tcn"), (3, "head"), (4, "attn")):
        ms, n = model.profile_read(kind)
        prof[name] = {"ms": ms, "launches": n}
    per_block = []
    for b in range(10):
        g_ms, g_n = model.profile_read(1, b)
        t_ms, t_n = model.profile_read(2, b)
        per_block.append({"gcn_ms": g_ms, "gcn_n": g_n, "tcn_ms": t_ms, "tcn_n": t_n})
        if args.workload in ("coa_gcn", "cos_tr"):
            per_block[-1]["attn_ms"], per_block[-1]["attn_n"] = model.profile_read(4, b)
    model.profile(False)
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    value = n_total * args.steps / (elapsed_ms * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers ------------------
    stage = torch.empty((n_local, C_IN, V, S), device=dev)
    host_out = torch.empty((n_total if world > 1 else n_local, CLASSES)).pin_memory()
    h2d = n_local * C_IN * V * S * 4
    d2h_total = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stage.copy_(host_frames[t % 8], non_blocking=True)
        out = model.forward_step(stage)
        if out is not None:
            if world > 1:
                out = cs.all_gather_logits(out, n_total)
            host_out.copy_(out, non_blocking=True)
            d2h_total += out.numel() * 4
        t += 1
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = n_total * args.steps / (e2e_ms * 1e-3)

    # ---- per-step latency (events per step, separate pass) -----------------------------------
    lat = []
    n_lat = min(args.steps, 200)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_lat + 1)]
    barrier()
    evs[0].record()
    for i in range(n_lat):
        step_resident(t)
        t += 1
        evs[i + 1].record()
    barrier()
    lat = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_lat))
    assert model.device_error() == 0, hex(model.device_error())

    if rank != 0:
        return
    peaks = load_peaks()
    # dominant kernel: the (kind, block) with the largest summed device time
    best = max(((pb["tcn_ms"], "tcn", i) for i, pb in enumerate(per_block)), default=(0, "tcn", 9))
    bg = max(((pb["gcn_ms"], "gcn", i) for i, pb in enumerate(per_block)), default=(0, "gcn", 9))
    if bg[0] > best[0]:
        best = bg
    _, kname, kblock = best
    pb = per_block[kblock]
    k_ms, k_n = (pb["tcn_ms"], pb["tcn_n"]) if kname == "tcn" else (pb["gcn_ms"], pb["gcn_n"])
    cout = [64, 64, 64, 64, 128, 128, 128, 256, 256, 256][kblock]
    tokens = n_local * S * V
    # algorithmic frames of (tokens x C x 4 B) per launch: temporal conv = 8 ring reads +
    # 1 delayed-residual read + 1 output write; graph conv = 1 ring write
    frames_moved = 10 if kname == "tcn" else 1
    k_bytes = frames_moved * tokens * cout * 4.0
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = k_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    # tensor side of the same kernel: credited = 1-product FLOPs of the reference math; issued = the three
    # split-precision products actually executed on the 128-row tiles (125 valid rows per tile)
    cin = [C_IN, 64, 64, 64, 64, 128, 128, 128, 256, 256][kblock]
    res_k = cin if kblock in (4, 7) else 0
    k_macs = tokens * cout * ((9 * cout + res_k) if kname == "tcn" else (3 * cin + (cin if cin != cout else 0)))
    k_tf_credit = 2.0 * k_macs / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0
    issued_k = (9 * cout + res_k) if kname == "tcn" else 4 * cin
    k_tf_issued = 3 * 2.0 * (tokens * 128.0 / 125.0) * cout * issued_k / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0
    traffic = NCU_TRAFFIC.get(f"{kname}<{cout}>") if n_local == 4096 and V == 25 and not (args.workload == "coa_gcn" and kname == "gcn") else None
    step_bytes = algo["state"] + algo["io"]
    line = {
        "metric": METRIC.replace("CoST-GCN", NAMES[args.workload]).replace("NTU60", "Kinetics" if V == 18 else "NTU120" if CLASSES == 120 else "NTU60"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 operands, f32 accumulate)" if args.kernel_path == "auto" else "f32",
        "data": "synthetic", "config": make_config(args, world),
        "p50_ms_per_step": lat[len(lat) // 2] if lat else None, "p95_ms_per_step": lat[int(len(lat) * 0.95)] if lat else None,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_total / args.steps,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "kernel": kernel_label(args.kernel_path, kname, cout, kblock, args.workload),
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "algorithmic_bytes_per_launch": k_bytes, "avg_launch_ms": k_avg_ms, "launches_timed": k_n,
            "peak_source": peaks["source"],
            "tensor": {"credited_tflops_1product": k_tf_credit, "issued_tflops_3product": k_tf_issued, "peak_tflops": peaks["bf16_tflops"],
                       "frac_credited": k_tf_credit / peaks["bf16_tflops"], "frac_issued": k_tf_issued / peaks["bf16_tflops"]},
        },
        "step_roofline": {
            "hbm_frac": step_bytes * (value / world) / 1e9 / peaks["hbm_gbs"],
            "tensor_frac": algo["flops"] * (value / world) / 1e12 / peaks["bf16_tflops"],
            "algorithmic_bytes_per_stream_frame": step_bytes, "algorithmic_flops_per_stream_frame": algo["flops"],
            "note": "1-product FLOPs and minimal state traffic; the 3 split-precision products and unfused inter-kernel traffic are not credited",
        },
        "kernel_time_ms": {k: v for k, v in prof.items()},
        "kernel_time_per_block_ms": per_block,
        "state_bytes": model.state_bytes(),
        "tensor_core_blocks": model.tensor_core_blocks(),
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, p50, cores = cpu_port_rate(args.workload, args.ref_streams, 40)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{args.ref_streams} concurrent streams x 40 steady-state steps of the same model on the host (oracle/step.py, "
                      f"eager torch fp32, {ms:.1f} ms/step)",
        }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cost_gcn", choices=list(ALGO))
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams per GPU")
    ap.add_argument("--ref-streams", type=int, default=64, help="streams per step of the CPU sample")
    ap.add_argument("--kernel-path", default="auto", choices=["auto", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    set_geometry(args.workload)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
