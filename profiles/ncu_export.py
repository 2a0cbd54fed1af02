"""Reduce an `ncu --set full` report to the per-launch summary CSV that bench.py reads for `roofline.traffic`
(profiles/ncu_full_main_summary.csv): one row per launch, the metrics the roofline discussion needs, units in row 2.

    ncu -i gpurun_out/main_full.ncu-rep --page raw --csv > /tmp/raw.csv     # here, no GPU needed
    python profiles/ncu_export.py /tmp/raw.csv profiles/ncu_full_main_summary.csv
"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "launch__registers_per_thread",
]


def main(src, dst):
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            if len(r) == len(hdr):
                w.writerow([r[i].replace("void cosk::", "void ").replace("cosk::", "").replace("(int)", "").replace("(bool)", "") for i in idx])
    print(f"{dst}: {len(data)} launches, {len(idx)} columns")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
