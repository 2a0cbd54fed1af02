"""Pretty-print the per-kernel timing block of a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("{")][-1]
    d = json.loads(line)
    print("==", path, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "p50", round(d.get("p50_ms_per_step") or 0, 3),
          "e2e %.0f" % d["e2e"]["value"], "launches", d["gpu_launches"])
    print("  clocks", d.get("clocks"))
    r = d["roofline"]
    print("  roofline %s achieved %.0f GB/s frac %.3f avg %.3f ms" % (r["kernel"], r["achieved"], r["frac"], r["avg_launch_ms"]))
    print("  step_roofline hbm %.3f tensor %.3f" % (d["step_roofline"]["hbm_frac"], d["step_roofline"]["tensor_frac"]))
    print("  kernel_time", {k: round(v["ms"], 2) for k, v in d["kernel_time_ms"].items()})
    for i, pb in enumerate(d["kernel_time_per_block_ms"]):
        g = pb["gcn_ms"] / max(pb["gcn_n"], 1)
        t = pb["tcn_ms"] / max(pb["tcn_n"], 1)
        print("   L%-2d gcn %.3f ms x%-4d tcn %.3f ms x%d" % (i + 1, g, pb["gcn_n"], t, pb["tcn_n"]))
    if "cpu_baseline" in d:
        print("  cpu", d["cpu_baseline"])
