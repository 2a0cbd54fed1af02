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
# warm frames to the first logits and frames per prediction of each workload
SCHED = {"cost_gcn": (297, 4), "cost_gcn_mod": (300, 1), "coa_gcn": (297, 4), "cos_tr": (297, 4)}


def stack_blocks(workload):
    """(cin, cout, stride, has_residual, graph-conv kind) of the ten blocks (models/cost_gcn/cost_gcn.py:30-41,
    models/cost_gcn_mod/cost_gcn_mod.py:29-40, models/coa_gcn/coa_gcn.py:17-46, models/cos_tr/cos_tr.py:24-41)."""
    s = 1 if workload == "cost_gcn_mod" else 2
    chans = [(C_IN, 64, 1), (64, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, s), (128, 128, 1), (128, 128, 1), (128, 256, s),
             (256, 256, 1), (256, 256, 1)]
    kind = {"coa_gcn": "adaptive"}.get(workload, "plain")
    return [(ci, co, st, i > 0, "attention" if (workload == "cos_tr" and i >= 3) else kind) for i, (ci, co, st) in enumerate(chans)]


def algorithmic_cost(workload, v=None):
    """Algorithmic work per stream-frame, derived from the block table (minimal input-ring algorithm,
    fp32 state, S = 2 skeletons): per executed block 1 ring-frame write, per emitting step 8 ring-frame reads and -- for
    blocks with a residual -- a delay-line read + write; MACs of the reference math with stride-phase skipping
    (models/base.py:260-270 graph conv, :307-334 temporal conv; models/a_gcn/a_gcn.py:48-69 embeddings + per-frame V x V
    attention; models/s_tr/s_tr.py:134-231 qkv / 8-head attention / output conv).  Checked against the SURVEY totals in
    tests/test_boundary_cpu.py.  Returns per-block lists as well (bytes, MACs, executions and emissions per frame)."""
    v = V if v is None else v
    rate, state, macs, per_block = 1.0, 0.0, 0.0, []
    for cin, cout, stride, res, kind in stack_blocks(workload):
        frame = v * cout * 4.0
        r_in, r_out = rate, rate / stride
        b_gcn, b_tcn = frame, 8 * frame + (2 * frame if res else 0.0)  # per gcn execution / per tcn emission, one skeleton
        if kind == "attention":
            dk, dv = cout // 4, cout
            g = cin * (2 * dk + dv) * v + v * v * (dk + dv) + dv * dv * v
        else:
            g = 3 * cin * v * v + 3 * cin * cout * v + (cin * cout * v if cin != cout else 0)
            if kind == "adaptive":
                g += 6 * cin * (cout // 4) * v + 3 * v * v * (cout // 4)
        t = 9 * cout * cout * v + (cin * cout * v if (res and (cin != cout or stride != 1)) else 0)
        state += S * (r_in * b_gcn + r_out * b_tcn)
        macs += S * (r_in * g + r_out * t)
        per_block.append({"cin": cin, "cout": cout, "gcn_bytes": S * b_gcn, "tcn_bytes": S * b_tcn, "gcn_macs": S * g, "tcn_macs": S * t,
                          "gcn_per_frame": r_in, "tcn_per_frame": r_out})
        rate = r_out
    warm, period = SCHED[workload]
    io = C_IN * v * S * 4 + (CLASSES * 4 + 2048) / period  # input frame + logits and pool running-sum minimum per emission
    return {"state": state, "io": io, "flops": 2.0 * macs, "warm": warm, "period": period, "blocks": per_block}


NAMES = {"cost_gcn": "CoST-GCN", "cost_gcn_mod": "CoST-GCN*", "coa_gcn": "CoA-GCN", "cos_tr": "CoS-TR"}
# workload -> V, classes, dataset, label (BASELINE configs[2] quotes CoA-GCN on NTU RGB+D 120, configs[3] CoS-TR on Kinetics)
GEOMETRY = {"cos_tr": (18, 400, "dummy_kin", "Kinetics-400 skeleton"), "coa_gcn": (25, 120, "ntu120", "NTU RGB+D 120 joint stream")}
DATASET, DATA_LABEL = "dummy_ntu", "NTU RGB+D 60 joint stream"


def set_geometry(workload):
    """Vertex count / class count of the workload's dataset (module-level, read by every arm)."""
    global V, CLASSES, DATASET, DATA_LABEL
    V, CLASSES, DATASET, DATA_LABEL = GEOMETRY.get(workload, (25, 60, "dummy_ntu", "NTU RGB+D 60 joint stream"))


NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_full_main_summary.csv")  # newest `ncu --set full` export of one full step


def ncu_traffic(kernel_regex, streams):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels matching `kernel_regex`, averaged over the
    launches in the committed ncu export (profiles/ncu_full_main_summary.csv, produced from an `ncu --set full` capture of
    `bench.py --streams 4096` by profiles/ncu_export.py).  None when there is no capture for this kernel / stream count."""
    import csv
    import re

    if streams != 4096 or not os.path.exists(NCU_SUMMARY):
        return None
    rows = list(csv.reader(open(NCU_SUMMARY)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    try:
        k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(x[r]) * mult.get(units[r], 1.0) + float(x[w]) * mult.get(units[w], 1.0) for x in rows[2:]
            if len(x) > max(k, r, w) and re.search(kernel_regex, x[k])]
    return sum(vals) / len(vals) if vals else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  The sampler is started well before the timed
    regions (spawning nvidia-smi next to a 40 ms region perturbs it) and every sample carries a host timestamp, so the
    summary can be restricted to the samples that fell inside a timed window."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        """`windows`: list of (t0, t1) host times of the timed regions; falls back to all samples under load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        good = [(ts, r) for ts, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        # nvidia-smi prints a sample at the END of its interval: accept samples up to one period after a window
        inside = [r for ts, r in good if any(t0 <= ts <= t1 + 0.06 for t0, t1 in windows)]
        where = "timed regions"
        if len(inside) < 3:
            inside, where = [r for ts, r in good if windows and windows[0][0] - 1.0 <= ts <= windows[-1][1] + 0.06], "warm-up + timed regions"
        if not inside:
            inside, where = [r for _, r in good], "whole run"
        sm = [float(r[1]) for r in inside]
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "sampled_over": where}


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
        for t in range(SCHED[workload][0] + warm_extra):
            model.forward_step(frames[t % 4])
        model.trace = []
        t0 = time.perf_counter()
        for t in range(steps):
            s0 = time.perf_counter()
            model.forward_step(frames[t % 4])
            per.append(time.perf_counter() - s0)
        dt = time.perf_counter() - t0
    return n_streams * steps / dt, dt / steps * 1e3, statistics.median(per) * 1e3, torch.get_num_threads()


def metric_name(workload):
    return METRIC.replace("CoST-GCN", NAMES[workload]).replace("NTU60", "Kinetics" if V == 18 else "NTU120" if CLASSES == 120 else "NTU60")


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = args.ref_streams
    config = make_config(args, world)
    rate, ms, p50, cores = cpu_port_rate(args.workload, n_sample, args.steps, warm_extra=args.warmup)
    sample = (f"{n_sample} concurrent streams per step (bounded sample of the {args.streams}-stream workload), steady state "
              f"after {SCHED[args.workload][0]} warm frames, oracle/step.py eager torch fp32")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "p50_ms_per_step": p50, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def kernel_label(path, kname, cin, cout, kblock, workload, knobs):
    if path != "auto":
        return f"k_{kname}_simt (layer {kblock + 1}, C={cout})", None
    if kname == "block":
        return f"k_tc_block64 (layer {kblock + 1}; graph conv + temporal conv of the block step in one kernel)", r"k_tc_block64"
    if kname == "tcn":
        pair = cout >= 128
        name = f"k_tc_tcn2<{cout}>" if pair else f"k_tc_tcn<{cout}>"
        return name + (" (CTA pairs)" if pair else "") + f" (layer {kblock + 1})", rf"k_tc_tcn2?<{cout}>"
    if workload == "coa_gcn":
        return f"k_tc_agcn C={cout} (layer {kblock + 1}; dense per-skeleton mix, attention kernel timed separately)", None
    if workload == "cos_tr" and kblock >= 3:
        return f"attention-unit output conv, k_tc_tcn with one tap, C={cout} (layer {kblock + 1}; qkv + attention timed separately)", None
    if cin < 64:
        return f"k_gcn_small (layer {kblock + 1})", r"k_gcn_small"
    if "gcnp" in knobs.get("graph_conv", {}).get(str(cout), ""):
        return f"k_tc_gcnp<{cout}> (layer {kblock + 1}; pre-mix, A operand in TMEM)", rf"k_tc_gcnp<{cout},"
    if "gcnt" in knobs.get("graph_conv", {}).get(str(cout), "") and cin <= 128:
        return f"k_tc_gcnt C={cout} (layer {kblock + 1}; channel-major GEMM, adjacency mix in registers)", r"k_tc_gcnt<"
    return f"k_tc_gcn C={cout} (layer {kblock + 1}; GEMM-then-mix)", r"k_tc_gcn<"


def make_config(args, world):
    return {
        "workload": f"{NAMES[args.workload]} {DATA_LABEL}, per-step forward_step, {args.streams} concurrent streams per GPU, "
                    f"random-init weights, synthetic U[0,1) frames (N,C=3,V={V},S=2)",
        "model_variant": args.workload, "streams_per_gpu": args.streams, "streams_total": args.streams * world,
        "V": V, "S": S, "classes": CLASSES,
        "sharding": f"streams sharded over {world} rank(s); on emitting steps the logits are all-gathered (NCCL) on a side stream into "
                    f"pre-allocated device buffers; every rank reads back its own shard's logits in the e2e pass",
        "l2": "per-step state traffic is GBs (>> 126 MB L2) and 8 distinct input frames are cycled, so no L2 flush is needed",
    }


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    import continual_skeletons_b200 as cs

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    algo = algorithmic_cost(args.workload)
    cls = {"cost_gcn": cs.CoStGcn, "cost_gcn_mod": cs.CoStGcnMod, "coa_gcn": cs.CoAGcn, "cos_tr": cs.CoSTr}[args.workload]
    torch.manual_seed(0)
    model = cls({"dataset_name": DATASET, "forward_mode": "frame", "kernel_path": args.kernel_path})
    n_local = args.streams
    n_total = n_local * world
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_frames = [torch.rand((n_local, C_IN, V, S), generator=gen).pin_memory() for _ in range(8)]
    dev_frames = [f.to(dev) for f in host_frames]
    gather = cs.LogitGather(n_total, CLASSES, dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # long before the first timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(t):
        out = model.forward_step(dev_frames[t % 8])
        if out is not None:
            gather.launch(out)  # side stream: the next step does not wait for the collective
        return out

    def join_side_stream():
        r = gather.result()  # current stream waits for the newest gather: it is inside whatever is being timed
        return r

    # state warm-up to steady state (every ring full, logits emitting), W untimed steps, then steady running for at least
    # --prewarm-s seconds so that clocks and power state have settled before anything is timed
    t = 0
    for _ in range(algo["warm"]):
        step_resident(t)
        t += 1
    for _ in range(max(args.warmup, 3)):
        step_resident(t)
        t += 1
    torch.cuda.synchronize()
    t_pre = time.time()
    # The steps carry a collective (the logit all-gather), so every rank must run the SAME number of them: the decision
    # to go on is taken by rank-local wall clock and then agreed on (cs.any_rank: MAX over ranks).  A purely local `while time < limit`
    # lets one rank leave the loop a batch earlier than its peers and meet their all-gather with the barrier's all-reduce.
    while cs.any_rank(time.time() - t_pre < args.prewarm_s, dev):
        for _ in range(16):
            step_resident(t)
            t += 1
        torch.cuda.synchronize()
    assert model.device_error() == 0, hex(model.device_error())
    windows = []

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------
    barrier()
    launches0 = model.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu = os.environ.get("COSK_NCU") == "1"  # profile only the timed steps: ncu --profile-from-start off
    if ncu:
        torch.cuda.cudart().cudaProfilerStart()
    w0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step_resident(t)
        t += 1
    join_side_stream()
    ev1.record()
    barrier()
    windows.append((w0, time.time()))
    if ncu:
        torch.cuda.cudart().cudaProfilerStop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = model.launch_count() - launches0
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    value = n_total * args.steps / (elapsed_ms * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers ------------------
    # every step: pinned host frame -> device, forward_step, and on emitting steps the rank's logits -> pinned host
    # (the all-gather runs beside it on the side stream, as in region 1)
    stage = [torch.empty((n_local, C_IN, V, S), device=dev) for _ in range(2)]
    host_out = torch.empty((n_local, CLASSES)).pin_memory()
    h2d = n_local * C_IN * V * S * 4
    d2h_total = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(args.steps):
        buf = stage[i & 1]
        buf.copy_(host_frames[t % 8], non_blocking=True)
        out = model.forward_step(buf)
        if out is not None:
            gather.launch(out)
            host_out.copy_(out, non_blocking=True)
            d2h_total += out.numel() * 4
        t += 1
    join_side_stream()
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = n_total * args.steps / (e2e_ms * 1e-3)
    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- the same K steps again with the library's per-kernel CUDA events on the launching stream (an event between
    # two kernels serialises them, so this pass is kept out of `value`) ----------------------------------------------
    model.profile(True)
    for _ in range(args.steps):
        step_resident(t)
        t += 1
    barrier()
    prof = {}
    for kind, name in ((0, "input"), (1, "gcn"), (2, "tcn"), (3, "head"), (4, "attn"), (5, "block")):
        ms, n = model.profile_read(kind)
        prof[name] = {"ms": ms, "launches": n}
    per_block = []
    for b in range(10):
        g_ms, g_n = model.profile_read(1, b)
        t_ms, t_n = model.profile_read(2, b)
        f_ms, f_n = model.profile_read(5, b)  # fused block step: one launch = one graph conv + one temporal conv
        per_block.append({"gcn_ms": g_ms, "gcn_n": g_n, "tcn_ms": t_ms, "tcn_n": t_n, "block_ms": f_ms, "block_n": f_n})
        if args.workload in ("coa_gcn", "cos_tr"):
            per_block[-1]["attn_ms"], per_block[-1]["attn_n"] = model.profile_read(4, b)
    model.profile(False)

    # ---- per-step latency (events per step, separate pass) -----------------------------------
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
    knobs = model.knobs()
    tokens = n_local * S * V
    blocks = algo["blocks"]
    # per block: the north star's own unit.  Algorithmic bytes / 1-product FLOPs of the launches timed in the profiled
    # pass over their summed device time (graph conv + attention half + temporal conv of that block).
    per_block_roofline = []
    for i, (pb, ab) in enumerate(zip(per_block, blocks)):
        ms = pb["gcn_ms"] + pb["tcn_ms"] + pb.get("attn_ms", 0.0) + pb["block_ms"]
        n_g, n_t = pb["gcn_n"] + pb["block_n"], pb["tcn_n"] + pb["block_n"]  # timed pass is steady state: a fused launch does both
        by = n_local * (n_g * ab["gcn_bytes"] + n_t * ab["tcn_bytes"])
        fl = 2.0 * n_local * (n_g * ab["gcn_macs"] + n_t * ab["tcn_macs"])
        per_block_roofline.append({
            "layer": i + 1, "cin": ab["cin"], "cout": ab["cout"], "ms_per_block_step": ms / max(n_g, 1),
            "kernels": "fused" if pb["block_n"] else "gcn + tcn",
            "hbm_frac": by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if ms > 0 else None,
            "tensor_frac_1product": fl / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"] if ms > 0 else None,
            "tensor_frac_issued_3product": 3 * fl * (128.0 / 125.0 if V == 25 else 128.0 / 126.0) / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"] if ms > 0 else None,
        })
    # dominant kernel: the (kind, block) with the largest summed device time
    best = max(((pb[k + "_ms"], k, i) for i, pb in enumerate(per_block) for k in ("tcn", "gcn", "block")), default=(0, "tcn", 9))
    _, kname, kblock = best
    pb, ab = per_block[kblock], blocks[kblock]
    k_ms, k_n = pb[kname + "_ms"], pb[kname + "_n"]
    cin, cout = ab["cin"], ab["cout"]
    # algorithmic bytes per launch: temporal conv = 8 ring-frame reads + delayed-residual read +
    # output write (the delay-line read + write of SURVEY 8d); graph conv = 1 ring-frame write; fused block step = both
    k_bytes = n_local * {"tcn": ab["tcn_bytes"], "gcn": ab["gcn_bytes"], "block": ab["tcn_bytes"] + ab["gcn_bytes"]}[kname]
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = k_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    k_flops = 2.0 * n_local * {"tcn": ab["tcn_macs"], "gcn": ab["gcn_macs"], "block": ab["tcn_macs"] + ab["gcn_macs"]}[kname]
    k_tf_credit = k_flops / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0
    k_tf_issued = 3 * k_flops * (128.0 / 125.0) / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0
    label, regex = kernel_label(args.kernel_path, kname, cin, cout, kblock, args.workload, knobs)
    traffic = ncu_traffic(regex, n_local) if (regex and V == 25) else None
    step_bytes = algo["state"] + algo["io"]
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 operands, f32 accumulate)" if args.kernel_path == "auto" else "f32",
        "data": "synthetic", "config": make_config(args, world),
        "p50_ms_per_step": lat[len(lat) // 2] if lat else None, "p95_ms_per_step": lat[int(len(lat) * 0.95)] if lat else None,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_total / args.steps,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "kernel": label,
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "traffic_source": os.path.relpath(NCU_SUMMARY, ROOT) if traffic else None,
            "algorithmic_bytes_per_launch": k_bytes, "avg_launch_ms": k_avg_ms, "launches_timed": k_n,
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
        "per_block_roofline": per_block_roofline,
        "kernel_time_ms": {k: v for k, v in prof.items()},
        "kernel_time_per_block_ms": per_block,
        "state_bytes": model.state_bytes(),
        "tensor_core_blocks": model.tensor_core_blocks(),
        "kernel_knobs": knobs,
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
    ap.add_argument("--workload", default="cost_gcn", choices=list(SCHED))
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams per GPU")
    ap.add_argument("--ref-streams", type=int, default=64, help="streams per step of the CPU sample")
    ap.add_argument("--kernel-path", default="auto", choices=["auto", "simt"])
    ap.add_argument("--prewarm-s", type=float, default=1.5, help="seconds of steady stepping before the first timed region")
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
