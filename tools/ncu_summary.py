#!/usr/bin/env python
"""Selected metrics of an `ncu --set full` report, one column per captured launch (what profiles/*_ncu_full_*.csv hold).

    python tools/ncu_summary.py gpurun_out/prof_gemm_r01j.ncu-rep > profiles/r01j_ncu_full_linear_umma2.csv
"""
import csv
import io
import subprocess
import sys

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in launches])


if __name__ == "__main__":
    main()
