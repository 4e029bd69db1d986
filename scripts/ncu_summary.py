"""Key metrics of one or more .ncu-rep captures as text (run where ncu is installed; no GPU needed).
    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("---", rep)
        print("Kernel Name =", d.get("Kernel Name"))
        for k in hdr:
            if k in KEEP or ("warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                print(f"{k} [{u.get(k, '')}] = {d[k]}")
